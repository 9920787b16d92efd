"""SIMT kernels that were written without GPU access, executed on the CPU (tests/cpu_shim: every CUDA thread is a fiber,
barriers / shuffles / cooperative grid sync are fiber barriers, the scheduler shuffles the thread order).

deferred LayerNorm / 16 epilogue warps / peer scatter: the epilogue functors of csrc/encoder.cu are driven tile by tile
with CPU-computed accumulators in the thread numbering of the GEMM kernels and compared element by element (ragged
shapes, sentinel-filled buffers), together with ln_stats / pack_defer / gather_cls_ln / cls_normalize_scatter.

encoder host logic: ac_encoder_create / ac_encoder_forward_cls (default flow, deferred-LayerNorm flow, CLS-only tail, 16-warp
and pair dispatch) run against a fake CUDA runtime with the real epilogue functors and small kernels; the tensor-core mainloop
and the attention kernels are replaced by plain CPU stand-ins.  Unit CLS rows and hidden states are compared with the fp32
oracle.

tensor path: tests/cpu_shim/tc_emul.h models mbarrier / TMA (128-byte swizzle) / tcgen05.mma / TMEM functionally.  The
B200-verified attention_kernel must reproduce a plain reference through the model (this validates the MODEL), then
attention_pipe_kernel must give the same bits for several grid sizes, and the real gemm_tc_kernel runs all its warp roles
with the deferred-LayerNorm GELU epilogue.

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
    (gen / "_gen_peer_cu.inc").write_text(ex.extract(os.path.join(ROOT, "adaptive_classifier_b200", "csrc", "peer.cu")))
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


def _build_encoder_host(tmp_path):
    import extract_device_code as ex
    gen = tmp_path / "gen_host"
    gen.mkdir(exist_ok=True)
    csrc = os.path.join(ROOT, "adaptive_classifier_b200", "csrc")
    (gen / "_gen_common_host.inc").write_text(ex.extract(os.path.join(csrc, "common.cuh"), host=True))
    (gen / "_gen_gemm_tc.inc").write_text(ex.extract(os.path.join(csrc, "gemm_tc.cuh")))
    (gen / "_gen_peer.inc").write_text(ex.extract(os.path.join(csrc, "peer.cuh")))
    (gen / "_gen_encoder_host.inc").write_text(ex.extract(os.path.join(csrc, "encoder.cu"), host=True))
    return _gxx(gen, "encoder_emul.cpp", str(tmp_path / "enc_host"))


def test_encoder_host_logic_and_deferred_layernorm_flow_on_the_cpu_emulation(tmp_path):
    import struct
    import numpy as np
    import torch
    from oracle import encoder_oracle as eo
    exe = _build_encoder_host(tmp_path)
    L, H, heads, I, V, maxpos, B, S = 3, 128, 2, 256, 200, 64, 3, 20
    sd, cfg, _ = eo.make_bert_state_dict(1234, vocab_size=V, hidden_size=H, num_hidden_layers=L, num_attention_heads=heads,
                                         intermediate_size=I, max_position_embeddings=maxpos, type_vocab_size=2)
    g = torch.Generator().manual_seed(5)
    for k in list(sd.keys()):                      # LayerNorm parameters away from (1, 0), row means away from 0
        if k.endswith("LayerNorm.weight"):
            sd[k] = 1.0 + 0.3 * torch.randn(sd[k].shape, generator=g)
        elif k.endswith("LayerNorm.bias"):
            sd[k] = 0.2 * torch.randn(sd[k].shape, generator=g)
        elif k.endswith("output.dense.bias"):
            sd[k] = sd[k] + 0.5
    ids = torch.randint(5, V, (B, S), generator=g)
    mask = torch.ones_like(ids)
    for b in range(B):
        n = S - 1 - 3 * b
        mask[b, n:] = 0
        ids[b, n:] = 0
    ref, ref_hidden = eo.encoder_forward_cls(sd, ids, mask, num_heads=heads, ln_eps=cfg.layer_norm_eps, return_hidden=True)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("11i", L, H, heads, I, V, maxpos, 2, B, S, 1, 0))
        f.write(struct.pack("f", cfg.layer_norm_eps))
        w = lambda t: f.write(t.detach().float().contiguous().numpy().tobytes())
        for k in ("word_embeddings.weight", "position_embeddings.weight", "token_type_embeddings.weight", "LayerNorm.weight", "LayerNorm.bias"):
            w(sd["embeddings." + k])
        for l in range(L):
            p = f"encoder.layer.{l}."
            for name in ("attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense"):
                w(sd[p + name + ".weight"]); w(sd[p + name + ".bias"])
            for name in ("attention.output.LayerNorm.weight", "attention.output.LayerNorm.bias", "intermediate.dense.weight",
                         "intermediate.dense.bias", "output.dense.weight", "output.dense.bias", "output.LayerNorm.weight", "output.LayerNorm.bias"):
                w(sd[p + name])
        f.write(ids.to(torch.int32).numpy().tobytes())
        f.write(mask.to(torch.int32).numpy().tobytes())
    r = subprocess.run([exe, str(fin), str(fout)], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "forward_cls_scatter: ok" in r.stdout, r.stdout[-1500:] + r.stderr[-500:]
    raw = open(fout, "rb").read()
    off, seen, base = 0, set(), None
    keep = mask.bool()
    while off < len(raw):
        tag = struct.unpack_from("6i", raw, off); off += 24
        cls = torch.from_numpy(np.frombuffer(raw, np.float32, B * H, off).reshape(B, H).copy()); off += 4 * B * H
        hid = torch.from_numpy(np.frombuffer(raw, np.float32, B * S * H, off).reshape(B, S, H).copy()); off += 4 * B * S * H
        seen.add(tag[:5])
        assert (cls - ref).norm(dim=1).max() < 2e-4, (tag, (cls - ref).norm(dim=1).max())      # encoder tolerance is 1e-3
        assert (cls.norm(dim=1) - 1).abs().max() < 1e-5
        if tag[5]:
            assert (hid[keep] - ref_hidden[keep]).abs().max() < 2e-3, tag
        if tag[:5] == (0, 0, 0, 0, 0):
            base = cls
        elif tag[4] and not tag[1]:                           # CLS-only attention: the same arithmetic on one query row
            assert torch.equal(cls, base), tag
        elif tag[1]:                                          # deferred flow vs the LayerNorm-kernel flow of the same build
            assert (cls - base).norm(dim=1).max() < 1e-4, tag
    assert {(0, 0, 0, 0, 0), (1, 1, 0, 0, 0), (0, 1, 0, 0, 0), (1, 1, 3, 0, 0), (0, 1, 3, 1, 0), (1, 0, 0, 0, 1), (1, 1, 0, 0, 1)} <= seen
    assert "attention_cls_kernel == attention stand-in on the CLS rows: ok" in r.stdout


def _build_tensor_path(tmp_path, mutate=None, driver="attention_emul.cpp", mutate_gemm2=None, mutate_knn=None):
    import extract_device_code as ex
    gen = tmp_path / "gen_tc"
    gen.mkdir(exist_ok=True)
    csrc = os.path.join(ROOT, "adaptive_classifier_b200", "csrc")
    (gen / "_gen_common_tc.inc").write_text(ex.extract(os.path.join(csrc, "common.cuh"), tc=True))
    (gen / "_gen_gemm_tc_tc.inc").write_text(ex.extract(os.path.join(csrc, "gemm_tc.cuh"), tc=True))
    g2 = ex.extract(os.path.join(csrc, "gemm_tc2.cuh"), tc=True)
    (gen / "_gen_gemm_tc2_tc.inc").write_text(mutate_gemm2(g2) if mutate_gemm2 else g2)
    (gen / "_gen_peer.inc").write_text(ex.extract(os.path.join(csrc, "peer.cuh")))
    src = ex.extract(os.path.join(csrc, "encoder.cu"), tc=True)
    (gen / "_gen_encoder_tc.inc").write_text(mutate(src) if mutate else src)
    knn = ex.extract(os.path.join(csrc, "knn_tc.cu"), tc=True)
    (gen / "_gen_knn_tc_tc.inc").write_text(mutate_knn(knn) if mutate_knn else knn)
    return _gxx(gen, driver, str(tmp_path / (driver.split(".")[0] + ("_mut" if (mutate or mutate_gemm2 or mutate_knn) else ""))))


def test_pipelined_attention_equals_the_verified_kernel_on_the_blackwell_model(tmp_path):
    exe = _build_tensor_path(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "attention_emul: ALL OK" in r.stdout, r.stdout[-2500:] + r.stderr[-500:]
    assert r.stdout.count("attention_pipe == attention_kernel") == 3 and "gemm_tc_kernel<EpiLinear<GELU, DEFER>>" in r.stdout


def test_the_blackwell_model_detects_a_wrong_buffer_offset(tmp_path):
    """mutant: the PV MMA of attention_pipe_kernel reads V^T from the other buffer's slab"""
    def wrong_slab(src):
        old = "umma_desc_sw128(smem_u32(buf + 32 * 1024 + slab * 8192))"
        assert old in src
        return src.replace(old, "umma_desc_sw128(smem_u32(smem + (b ^ 1) * ATTP_BUF + 32 * 1024 + slab * 8192))")
    exe = _build_tensor_path(tmp_path, mutate=wrong_slab)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=1500)
    assert r.returncode != 0 and "outputs differ from attention_kernel" in r.stdout, r.stdout[-1500:]


def test_pair_gemm_variants_equal_the_one_cta_kernel_on_the_blackwell_model(tmp_path):
    """cluster of two concurrently running CTAs: the pair kernel with 8 epilogue warps (ran bit-identically on a B200), the
    relay variant, 16 epilogue warps / 5 stages with the 64-column functors, and the in-place deferred residual epilogue all
    give the bits of the 1-CTA kernel"""
    exe = _build_tensor_path(tmp_path, driver="pair_emul.cpp")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "pair_emul: ALL OK" in r.stdout, r.stdout[-2500:] + r.stderr[-800:]


def test_the_pair_model_detects_a_wrong_barrier_count(tmp_path):
    """mutant: the leader's accumulator-drained barrier still expects 2 x 8 arrivals when 16 epilogue warps arrive per CTA"""
    def wrong_count(src):
        assert "mbar_init(&tmem_empty[0], 2 * kEpiWarps);" in src
        return src.replace("mbar_init(&tmem_empty[0], 2 * kEpiWarps);", "mbar_init(&tmem_empty[0], 2 * GEMM_EPI_WARPS);")
    exe = _build_tensor_path(tmp_path, driver="pair_emul.cpp", mutate_gemm2=wrong_count)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=1500)
    assert r.returncode != 0, r.stdout[-800:]


def test_knn_coarse_pass_epilogues_on_the_blackwell_model(tmp_path):
    """gemm_tc_kernel<EpiKnn> (verified on a B200) and <EpiKnnLane> (per-lane slow path, option "knn_epi"): per query, the best k
    of the union of its candidate lists equals a brute-force scan with the same arithmetic; lists sorted, no row twice"""
    exe = _build_tensor_path(tmp_path, driver="knn_emul.cpp")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "knn_emul: ALL OK" in r.stdout, r.stdout[-2500:] + r.stderr[-800:]


def test_the_knn_model_detects_a_dropped_column_half(tmp_path):
    """mutant: the per-lane slow path forgets the hits in columns 16..31 of every chunk"""
    def drop_half(src):
        old = "for (int half = 0; half < 2; ++half) {\n            uint32_t hh"
        assert old in src
        return src.replace(old, "for (int half = 0; half < 1; ++half) {\n            uint32_t hh")
    exe = _build_tensor_path(tmp_path, driver="knn_emul.cpp", mutate_knn=drop_half)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=1500)
    assert r.returncode != 0 and "EpiKnnLane" in r.stdout and "FAIL" in r.stdout, r.stdout[-1500:]
