"""CPU tests of the host-side mirror (no kernels): config, examples, seeded head init, label rules, sharding."""
import os

import numpy as np
import torch

import adaptive_classifier_b200 as acb
from adaptive_classifier_b200.parallel import shard_bounds
from adaptive_classifier_b200.persistence import select_representative_examples

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_public_names_match_reference_init():
    for n in ["AdaptiveClassifier", "MultiLabelAdaptiveClassifier", "MultiLabelAdaptiveHead", "Example", "AdaptiveHead",
              "ModelConfig", "PrototypeMemory"]:
        assert hasattr(acb, n)


def test_model_config_defaults_and_roundtrip():
    c = acb.ModelConfig()
    assert (c.max_length, c.max_examples_per_class, c.prototype_update_frequency, c.ewc_lambda) == (512, 1000, 100, 100.0)
    assert (c.prototype_weight, c.neural_weight, c.num_representative_examples) == (0.7, 0.3, 5)
    d = c.to_dict()
    assert len(d) == 29 and d["cost_coefficients"] == {}
    c2 = acb.ModelConfig({"max_length": 64, "b200_max_tokens": 1024})
    assert c2.max_length == 64 and c2.config["b200_max_tokens"] == 1024
    c2.update(max_length=32, nonsense=1)
    assert c2.max_length == 32 and not hasattr(c2, "nonsense")


def test_example_roundtrip():
    e = acb.Example("hello", "greet", torch.tensor([0.5, -1.0]))
    e2 = acb.Example.from_dict(e.to_dict())
    assert e2.text == "hello" and e2.label == "greet" and torch.equal(e2.embedding, e.embedding)
    assert acb.Example.from_dict(acb.Example("x", "y").to_dict()).embedding is None


def test_adaptive_head_init_is_the_reference_seeded_init():
    g = np.load(os.path.join(GOLD, "golden_head.npz"))
    head = acb.AdaptiveHead(64, 5, hidden_dims=[64, 32])
    sd = head.state_dict()
    assert sorted(sd) == ["model.0.bias", "model.0.weight", "model.3.bias", "model.3.weight", "model.6.bias", "model.6.weight"]
    for k, v in sd.items():
        assert torch.equal(v, torch.from_numpy(g[k])), k
    # growth keeps the old rows and re-seeds the new ones (models.py:82-98)
    w_old = head.model[-1].weight.detach().clone()
    head.update_num_classes(7)
    assert head.model[-1].weight.shape == (7, 32) and torch.equal(head.model[-1].weight[:5], w_old)
    head.update_num_classes(6)
    assert head.model[-1].weight.shape == (7, 32)


def test_multilabel_head_growth_preserves_rows():
    h = acb.MultiLabelAdaptiveHead(16, 3, hidden_dims=[16, 8])
    w = h.model[-1].weight.detach().clone()
    h.update_num_classes(5)
    assert h.num_classes == 5 and torch.equal(h.model[-1].weight[:3], w)


def test_representative_examples_selection():
    g = torch.Generator().manual_seed(0)
    exs = [acb.Example(f"t{i}", "a", torch.randn(8, generator=g)) for i in range(20)]
    sel = select_representative_examples(exs, k=5)
    assert len(sel) == 5 and all(s in exs for s in sel)
    assert select_representative_examples(exs[:3], k=5) == exs[:3]


def test_shard_bounds_cover_rows_exactly():
    for N, G in ((1_000_000, 8), (10, 3), (7, 8), (500_000, 4)):
        spans = [shard_bounds(N, r, G) for r in range(G)]
        assert spans[0][0] == 0 and spans[-1][1] == N
        assert all(spans[i][1] == spans[i + 1][0] for i in range(G - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_distilbert_state_dict_mapping_matches_hf():
    """DistilBERT -> BERT-named parameters (+ zero type table): the oracle restatement on the mapped weights equals HF"""
    from transformers import DistilBertConfig, DistilBertModel
    from adaptive_classifier_b200._cabi import distilbert_to_bert_state_dict
    from oracle import encoder_oracle as eo
    torch.manual_seed(3)
    cfg = DistilBertConfig(vocab_size=400, dim=128, n_heads=2, n_layers=2, hidden_dim=256, max_position_embeddings=64)
    m = DistilBertModel(cfg).eval()
    ids = eo.synthetic_ids(3, 20, vocab=400)
    mask = torch.ones_like(ids)
    mask[1, 12:] = 0
    with torch.no_grad():
        ref = torch.nn.functional.normalize(m(input_ids=ids, attention_mask=mask).last_hidden_state[:, 0, :], dim=1)
    sd, dims = distilbert_to_bert_state_dict(dict(m.state_dict()), cfg)
    out = eo.encoder_forward_cls({k: v.float() for k, v in sd.items()}, ids, mask, num_heads=2, ln_eps=1e-12)
    assert (out - ref).abs().max() < 1e-6 and dims["type_vocab"] == 1 and dims["layers"] == 2


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs first) on a reduced index: one JSON line with the
    contract's keys; rank != 0 exits silently."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--rows", "3000"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-800:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "queries/s" and j["value"] > 0 and j["higher_is_better"] is True
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    env["RANK"] = "1"
    env["WORLD_SIZE"] = "2"
    out1 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                           "--rows", "3000"], capture_output=True, text=True, env=env, timeout=120)
    assert out1.returncode == 0 and out1.stdout.strip() == ""


def _reference_memory_semantics(D, cap, freq):
    """plain restatement of memory.py:41-83 / :138-153 / :196-217 (per example: append, prune to the `cap` nearest the mean of the
    cap + 1 in distance order, prototype = mean of the retained) used as the expectation for the batched device-store path"""
    import numpy as np
    import torch

    class Ref:
        def __init__(self):
            self.examples, self.prototypes = {}, {}

        def add(self, ex, label):
            lst = self.examples.setdefault(label, [])
            lst.append(ex)
            if len(lst) > cap:
                E = torch.stack([e.embedding for e in lst])
                mean = E.mean(0)
                dist = [torch.norm(e.embedding - mean).item() for e in lst]
                keep = np.argsort(dist, kind="stable")[:cap]
                self.examples[label] = [lst[i] for i in keep]
            self.prototypes[label] = torch.stack([e.embedding for e in self.examples[label]]).mean(0)
    return Ref()


def _torch_memory_append_prune(rows, order, count, new_rows, new_index, cls_start, touched):
    """torch-CPU stand-in with the contract of ac_memory_append_prune (host-logic tests only)"""
    import torch
    cap = rows.shape[1] - 1
    nt = touched.numel()
    src = torch.full((nt, cap), -1, dtype=torch.int32)
    proto = torch.zeros((nt, rows.shape[2]))
    for t in range(nt):
        c = int(touched[t])
        n = int(count[c])
        n_old = n
        ords = order[c].tolist()
        srcs = list(range(n)) + [-1] * (cap + 1 - n)
        for jj, j in enumerate(range(int(cls_start[t]), int(cls_start[t + 1]))):
            rows[c, ords[n]] = new_rows[int(new_index[j])]
            srcs[n] = n_old + jj
            n += 1
            if n > cap:
                E = rows[c, ords[:n]]
                dist = (E - E.double().mean(0).float()).norm(dim=1)
                perm = sorted(range(n), key=lambda i: (float(dist[i]), i))
                ords = [ords[i] for i in perm] + ords[n:]
                srcs = [srcs[i] for i in perm][:cap] + [-1]
                n = cap
        order[c] = torch.tensor(ords, dtype=torch.int32)
        count[c] = n
        src[t, :n] = torch.tensor(srcs[:n], dtype=torch.int32)
        proto[t] = rows[c, ords[:n]].double().mean(0).float()
    return src, proto


def test_device_store_host_logic_reproduces_the_per_example_semantics(monkeypatch):
    """add_examples_batch through the device-resident class stores (grouping by class, provenance ->
    Example lists, store validity after clear / direct edits) == the reference's per-example append / prune / mean sequence.
    The device entry points are replaced by torch-CPU stand-ins here (host logic only; the kernel itself is GPU-tested)."""
    import torch
    import adaptive_classifier_b200 as acb
    from adaptive_classifier_b200 import memory as mem_mod

    def segment_mean(X, cls, C):
        mean = torch.zeros((C, X.shape[1]), dtype=torch.float32)
        cnt = torch.zeros((C,), dtype=torch.int32)
        for c in range(C):
            r = X[cls == c]
            cnt[c] = r.shape[0]
            if r.shape[0]:
                mean[c] = r.sum(0) / r.shape[0]
        return mean, cnt

    monkeypatch.setattr(mem_mod, "_device", lambda: torch.device("cpu"))
    monkeypatch.setattr(mem_mod._cabi, "segment_mean", segment_mean)
    monkeypatch.setattr(mem_mod._cabi, "memory_append_prune", _torch_memory_append_prune)
    D, cap = 16, 12
    g = torch.Generator().manual_seed(0)
    mem = acb.PrototypeMemory(D, config=acb.ModelConfig({"max_examples_per_class": cap, "prototype_update_frequency": 7}))
    ref = _reference_memory_semantics(D, cap, 7)
    pool = ["a", "b", "c"]
    for call in range(16):
        n = int(torch.randint(1, 9, (1,), generator=g))
        labs = [pool[int(torch.randint(0, 3, (1,), generator=g))] for _ in range(n)]
        embs = [torch.nn.functional.normalize(torch.randn(D, generator=g), dim=0) for _ in range(n)]
        exs = [acb.Example(f"t{call}_{i}", l, e) for i, (l, e) in enumerate(zip(labs, embs))]
        if call == 9:
            mem.add_example(exs[0], labs[0])                       # the single-example entry is a batch of one
            mem.add_examples_batch(exs[1:], labs[1:])
        else:
            mem.add_examples_batch(exs, labs)
        for e, l in zip(exs, labs):
            ref.add(e, l)
        for l in ref.examples:
            assert [e.text for e in mem.examples[l]] == [e.text for e in ref.examples[l]], (call, l)   # same retained set, same ORDER
            assert (mem.prototypes[l] - ref.prototypes[l]).abs().max() < 1e-6
        if call == 11:                                              # an edit behind the store's back must be noticed
            mem.examples["a"] = list(reversed(mem.examples["a"]))
            ref.examples["a"] = list(reversed(ref.examples["a"]))
    assert max(len(v) for v in mem.examples.values()) == cap        # classes went through pruning
    assert mem.updates_since_rebuild < 7 and mem.index.ntotal in (0, 3)
    mem.clear()
    assert "_dev_store" not in mem.__dict__ and len(mem.examples) == 0
