"""SIMT kernels that were written without GPU access, executed on the CPU (tests/cpu_shim: every CUDA thread is a fiber,
barriers / shuffles / cooperative grid sync are fiber barriers, the scheduler shuffles the thread order).

deferred LayerNorm / 16 epilogue warps / peer scatter: the epilogue functors of csrc/encoder.cu are driven tile by tile
with CPU-computed accumulators in the thread numbering of the GEMM kernels and compared element by element (ragged
shapes, sentinel-filled buffers), together with ln_stats / pack_defer / gather_cls_ln / cls_normalize_scatter.

head_fused: the device code of csrc/head.cu is cut out of the .cu file and compiled as C++; one training epoch is run through
the launch-per-kernel sequence and through fused::head_epoch_kernel (cooperative, several blocks) from identical states and
must give bit-identical parameters, AdamW moments and loss.  A mutant without one grid barrier must fail, otherwise the
emulation would prove nothing."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "cpu_shim")
sys.path.insert(0, SHIM)


CUDA_INC = "/usr/local/cuda/include"       # host-usable vector / fp16 headers


def _gxx(gen, driver, exe):
    cmd = ["g++", "-std=c++17", "-O1", f"-I{gen}", f"-I{SHIM}", f"-I{CUDA_INC}", os.path.join(SHIM, driver),
           os.path.join(SHIM, "cuda_shim.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


def _build(tmp_path, mutate=None):
    import extract_device_code as ex
    src = ex.extract(os.path.join(ROOT, "adaptive_classifier_b200", "csrc", "head.cu"))
    if mutate:
        src = mutate(src)
    gen = tmp_path / "gen"
    gen.mkdir(exist_ok=True)
    (gen / "_gen_head_device.inc").write_text(src)
    return _gxx(gen, "head_epoch_emul.cpp", str(tmp_path / ("emul_mut" if mutate else "emul")))


def _build_epilogues(tmp_path, mutate=None):
    import extract_device_code as ex
    gen = tmp_path / "gen_enc"
    gen.mkdir(exist_ok=True)
    for f in ("common.cuh", "gemm_tc.cuh", "peer.cuh", "encoder.cu"):
        src = ex.extract(os.path.join(ROOT, "adaptive_classifier_b200", "csrc", f))
        if mutate and f == "encoder.cu":
            src = mutate(src)
        (gen / f"_gen_{f.split('.')[0]}.inc").write_text(src)
    return _gxx(gen, "epilogue_emul.cpp", str(tmp_path / ("epi_mut" if mutate else "epi")))


def test_deferred_layernorm_epilogues_on_the_cpu_emulation(tmp_path):
    exe = _build_epilogues(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2500:] + r.stderr[-500:]


def test_the_epilogue_emulation_detects_a_wrong_column_mask(tmp_path):
    """mutant: the 16-warp consumer epilogue detects its first chunk with the 8-warp mask -> half of the warps never load
    their row statistics"""
    def wrong_mask(src):
        assert "(COLS - 1)) == 0" in src
        return src.replace("(COLS - 1)) == 0", "(GEMM_BLOCK_N / 2 - 1)) == 0")
    exe = _build_epilogues(tmp_path, mutate=wrong_mask)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode != 0 and "16 epilogue warps) M=300 N=392: FAIL" in r.stdout, r.stdout[-1500:]


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    return _build(tmp_path_factory.mktemp("head_emul"))


#          D   H0  H1  C   n   batch G  loss dropout ewc seed
CONFIGS = [(64, 64, 32, 5, 100, 32, 3, 0, 0.1, 0, 2),       # CE, dropout, 3 blocks, last batch partial
           (96, 80, 40, 7, 90, 40, 5, 1, 0.1, 0, 4),        # BCE, batch 40 (two 32-row blocks), ragged dims
           (64, 64, 32, 70, 70, 32, 7, 0, 0.1, 1, 5),       # C > 64 (two sgemm row blocks), EWC with a grown head
           (72, 72, 36, 3, 33, 64, 1, 1, 0.2, 1, 7)]        # a single block walks every virtual block


@pytest.mark.parametrize("cfg", CONFIGS)
def test_fused_head_epoch_equals_launch_per_kernel_on_the_cpu_emulation(emul, cfg):
    r = subprocess.run([emul] + [str(x) for x in cfg], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "FUSED == LAUNCH-PER-KERNEL" in r.stdout, r.stdout[-1500:] + r.stderr[-500:]


def test_the_emulation_detects_a_missing_grid_barrier(tmp_path):
    def drop_second_grid_sync(src):
        sites = [m.start() for m in re.finditer(r"grid\.sync\(\);", src)]
        assert len(sites) >= 9
        i = sites[1]                                        # the barrier between the h0 and h1 phases
        return src[:i] + "/* mutant: barrier removed */" + src[i + len("grid.sync();"):]
    exe = _build(tmp_path, mutate=drop_second_grid_sync)
    r = subprocess.run([exe] + [str(x) for x in CONFIGS[0]], capture_output=True, text=True, timeout=900)
    assert r.returncode != 0 and "MISMATCH" in r.stdout, r.stdout[-800:]
