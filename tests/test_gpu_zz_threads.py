"""Runs last in the GPU suite (file order): multi-threaded use of PrototypeMemory."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def acb(cabi):
    import adaptive_classifier_b200 as m
    return m


def _mem(acb, dim=64, **cfg):
    return acb.PrototypeMemory(dim, config=acb.ModelConfig(cfg))


def test_memory_concurrent_adds(acb):
    """mirror of /root/reference/tests/test_memory.py:226-256: 3 threads x 100 adds (GIL-level safety, one CUDA stream)"""
    import threading
    mem = _mem(acb, 32, max_examples_per_class=1000, prototype_update_frequency=50)
    errs = []

    def worker(t):
        try:
            g = torch.Generator().manual_seed(t)
            for i in range(100):
                mem.add_example(acb.Example(f"t{t}_{i}", f"class_{t}", torch.randn(32, generator=g)), f"class_{t}")
        except Exception as e:       # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(3)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    st = mem.get_stats()
    assert st["total_examples"] == 300 and st["num_classes"] == 3
    mem._rebuild_index()
    for t in range(3):
        ex = torch.stack([e.embedding for e in mem.examples[f"class_{t}"]]).mean(0)
        assert torch.allclose(mem.prototypes[f"class_{t}"], ex, atol=1e-5)
    assert len(mem.get_nearest_prototypes(torch.randn(32), k=3)) == 3
