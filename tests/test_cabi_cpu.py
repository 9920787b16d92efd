"""CPU tests of the boundary: the C-ABI library builds, loads and exports every symbol include/*.h declares;
without a GPU every compute entry fails loudly (no CPU fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "adaptive_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ac_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(cabi):
    L = cabi.load_library()
    declared = _header_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/adaptive_b200.h but not exported"
    for name in cabi.EXPORTS:
        assert name in declared, f"{name} bound in _cabi.py but not declared in the header"
    assert L.ac_version() == 1


def test_header_cites_reference_call_sites():
    src = open(os.path.join(ROOT, "include", "adaptive_b200.h")).read()
    for cite in ("memory.py:110-114", "models.py:71-80", "classifier.py:1271-1275", "ewc.py", "classifier.py:1358-1384"):
        assert cite in src


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(cabi):
    L = cabi.load_library()
    assert L.ac_device_check() != 0
    assert b"no CPU fallback" in L.ac_last_error() or b"CUDA" in L.ac_last_error()
    import adaptive_classifier_b200 as acb
    with pytest.raises(acb.AdaptiveB200Error):
        acb.AdaptiveClassifier("bert-base-uncased")
    with pytest.raises(acb.AdaptiveB200Error):
        acb.AdaptiveClassifier("bert-base-uncased", device="cpu")
    head = acb.AdaptiveHead(16, 3, hidden_dims=[16, 8]).eval()
    with pytest.raises(acb.AdaptiveB200Error):
        head(torch.zeros(2, 16))
    mem = acb.PrototypeMemory(16)
    with pytest.raises(acb.AdaptiveB200Error):
        mem.add_example(acb.Example("t", "a", torch.zeros(16)), "a")


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "adaptive_classifier_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace("oracle/knn_oracle.c", "").replace("oracle/precision_study.py", "").replace("oracle/deferred_ln_study.py", ""), f   # comments citing the studies


def test_options_roundtrip_without_a_gpu(cabi):
    """ac_set_option / ac_get_option are host-only state: they work without a device; unknown names are rejected"""
    for name in ("gemm_pair", "knn_pair", "ln_defer", "head_fused", "epi16", "attn_pipe", "pdl", "knn_epi", "cls_attn"):
        prev = cabi.get_option(name)
        with cabi.option(name, prev + 2):
            assert cabi.get_option(name) == prev + 2
        assert cabi.get_option(name) == prev
    with pytest.raises(cabi.AdaptiveB200Error):
        cabi.set_option("not_an_option", 1)
    with pytest.raises(cabi.AdaptiveB200Error):
        cabi.get_option("not_an_option")


def test_ac_options_environment_is_applied_on_load():
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); from adaptive_classifier_b200 import _cabi; _cabi.load_library(); "
            "print(_cabi.get_option('ln_defer'), _cabi.get_option('epi16'), _cabi.get_option('gemm_pair'))" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], env={**os.environ, "AC_OPTIONS": "ln_defer=1,epi16=3"},
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    assert out.stdout.split() == ["1", "3", "0"]
    bad = subprocess.run([sys.executable, "-c", code], env={**os.environ, "AC_OPTIONS": "bogus=1"}, capture_output=True, text=True, timeout=300)
    assert bad.returncode != 0 and "bogus" in bad.stderr
