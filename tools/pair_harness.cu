// pair_harness.cu -- torch-free check of the opt-in CTA-pair (tcgen05 cta_group::2) kernels against the measured 1-CTA
// kernels, through the C ABI only (include/adaptive_b200.h).  Starts in seconds (no Python import), so it fits in a very
// small GPU budget:
//     ./pair_harness linear    ac_linear_tc on the four encoder projection shapes: outputs compared bit for bit, both timed
//     ./pair_harness knn       ac_knn_l2_topk (tensor path, fp16 shadow): (d, id) compared bit for bit, both timed
//     ./pair_harness encoder   full bert-base-shaped forward (random weights), CLS rows compared, both timed
//     ./pair_harness epoch     option "head_fused": one training epoch of the head, fused cooperative kernel vs launch-per-kernel
//     ./pair_harness defer     the same forward with option "ln_defer" (deferred LayerNorm) against the LayerNorm-kernel flow:
//                              different association order, so a tolerance check; defer_full = without the CLS-only tail
// Every line is flushed as it is produced: if the experimental kernel traps (mbarrier watchdog), the baseline numbers
// printed before it are still in the log.  Build: tools/build_harness.sh (nvcc, links ../adaptive_classifier_b200/libadaptive_b200.so).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <cmath>
#include "../include/adaptive_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); fflush(stdout); exit(2); } } while (0)
#define AC(x) do { int rc_ = (x); if (rc_ != 0) { printf("ac error %d (%s) at %s:%d\n", rc_, ac_last_error(), __FILE__, __LINE__); fflush(stdout); exit(3); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return static_cast<uint32_t>(x);
}
__global__ void fill_f32(float *p, int64_t n, uint64_t seed, float scale) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        p[i] = (static_cast<float>(hash32(seed * 0x9E3779B97F4A7C15ull + i) >> 8) * (1.f / 8388608.f) - 1.f) * scale;
}
__global__ void fill_f16(__half *p, int64_t n, uint64_t seed, float scale) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        p[i] = __float2half_rn((static_cast<float>(hash32(seed * 0x9E3779B97F4A7C15ull + i) >> 8) * (1.f / 8388608.f) - 1.f) * scale);
}
__global__ void fill_const(float *p, int64_t n, float v) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}
__global__ void fill_ids(int32_t *p, int64_t n, int S, int vocab, uint64_t seed) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int s = static_cast<int>(i % S);
        p[i] = s == 0 ? 101 : (s == S - 1 ? 102 : 1000 + static_cast<int>(hash32(seed + i) % (vocab - 1000)));
    }
}
// unit-norm rows around class centres is not needed here: any well-spread rows exercise the same code
__global__ void normalize_rows(float *p, int64_t rows, int D) {
    const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) s += p[r * D + i] * p[r * D + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = rsqrtf(s);
    for (int i = lane; i < D; i += 32) p[r * D + i] *= inv;
}

template <class T> static T *dmalloc(size_t n) { void *p; CK(cudaMalloc(&p, n * sizeof(T))); return static_cast<T *>(p); }

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms = 0; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

// compares two device buffers on the device: number of differing 16-bit units and the max abs difference (NaN counts as inf)
__global__ void cmp_kernel(const uint16_t *a, const uint16_t *b, int64_t n16, int as_f32, unsigned long long *ndiff, unsigned int *maxbits) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    unsigned long long d = 0;
    float m = 0.f;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
        d += a[i] != b[i];
        float x, y;
        if (as_f32) {
            if (i & 1) continue;
            x = reinterpret_cast<const float *>(a)[i >> 1]; y = reinterpret_cast<const float *>(b)[i >> 1];
        } else {
            x = __half2float(reinterpret_cast<const __half *>(a)[i]); y = __half2float(reinterpret_cast<const __half *>(b)[i]);
        }
        float e = fabsf(x - y);
        if (!(e == e)) e = (a[i] == b[i] && (!as_f32 || a[i + 1] == b[i + 1])) ? 0.f : __int_as_float(0x7f800000);
        m = fmaxf(m, e);
    }
    if (d) atomicAdd(ndiff, d);
    atomicMax(maxbits, __float_as_uint(m));   // non-negative floats order like their bit patterns
}
static long long compare(const char *what, const void *d0, const void *d1, size_t bytes, bool as_f32) {
    static unsigned long long *nd = nullptr; static unsigned int *mb = nullptr;
    if (!nd) { nd = dmalloc<unsigned long long>(1); mb = dmalloc<unsigned int>(1); }
    CK(cudaMemset(nd, 0, 8)); CK(cudaMemset(mb, 0, 4));
    cmp_kernel<<<1184, 256>>>(static_cast<const uint16_t *>(d0), static_cast<const uint16_t *>(d1), static_cast<int64_t>(bytes / 2), as_f32 ? 1 : 0, nd, mb);
    unsigned long long hd = 0; unsigned int hm = 0;
    CK(cudaMemcpy(&hd, nd, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&hm, mb, 4, cudaMemcpyDeviceToHost));
    float maxabs; memcpy(&maxabs, &hm, 4);
    printf("  compare %-10s bytes=%zu differing_16bit_units=%llu max_abs_diff=%.3e %s\n", what, bytes, hd, maxabs, hd == 0 ? "BIT-IDENTICAL" : "DIFFERENT");
    fflush(stdout);
    return static_cast<long long>(hd);
}

static int g_pair = 1;   // option value under test: 1 = TMA-signalled pair kernel, 2 = relay variant

static int run_linear() {
    struct Shape { const char *name; int N, K, epi, out_half; };
    const Shape shapes[4] = {{"qkv", 2304, 768, 0, 1}, {"attn_out", 768, 768, 2, 0}, {"ffn1_gelu", 3072, 768, 1, 1}, {"ffn2", 768, 3072, 2, 0}};
    const int M = 65536;   // B = 512 sequences x 128 tokens
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    __half *X = dmalloc<__half>(static_cast<size_t>(M) * 3072);
    __half *W = dmalloc<__half>(static_cast<size_t>(3072) * 3072);
    float *bias = dmalloc<float>(3072), *res = dmalloc<float>(static_cast<size_t>(M) * 768);
    void *Y0 = dmalloc<uint8_t>(static_cast<size_t>(M) * 3072 * 2), *Y1 = dmalloc<uint8_t>(static_cast<size_t>(M) * 3072 * 2);
    fill_f16<<<1184, 256>>>(X, static_cast<int64_t>(M) * 3072, 1, 1.0f);
    fill_f16<<<1184, 256>>>(W, 3072ll * 3072, 2, 0.05f);
    fill_f32<<<64, 256>>>(bias, 3072, 3, 0.1f);
    fill_f32<<<1184, 256>>>(res, static_cast<int64_t>(M) * 768, 4, 1.0f);
    CK(cudaDeviceSynchronize());
    auto time_variant = [&](int variant) {
        AC(ac_set_option("gemm_pair", variant ? g_pair : 0));
        for (const Shape &sh : shapes) {
            void *Y = variant ? Y1 : Y0;
            for (int it = 0; it < 2; ++it) AC(ac_linear_tc(X, W, bias, res, Y, M, sh.N, sh.K, sh.epi, 0, AC_PREC_F16, sh.out_half, nullptr));
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            const int reps = 5;
            for (int it = 0; it < reps; ++it) AC(ac_linear_tc(X, W, bias, res, Y, M, sh.N, sh.K, sh.epi, 0, AC_PREC_F16, sh.out_half, nullptr));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            const double ms = time_ms(e0, e1) / reps;
            printf("linear %-10s variant=%s M=%d N=%d K=%d: %.1f us  %.1f TFLOP/s\n", sh.name, variant ? "pair" : "1cta", M, sh.N, sh.K,
                   ms * 1e3, 2.0 * M * sh.N * sh.K / (ms * 1e-3) / 1e12);
            fflush(stdout);
        }
    };
    time_variant(0);
    // correctness: per shape, baseline into Y0, pair into Y1, compare
    long long bad = 0;
    for (const Shape &sh : shapes) {
        const size_t ybytes = static_cast<size_t>(M) * sh.N * (sh.out_half ? 2 : 4);
        CK(cudaMemset(Y0, 0, ybytes)); CK(cudaMemset(Y1, 0xFF, ybytes));
        AC(ac_set_option("gemm_pair", 0));
        AC(ac_linear_tc(X, W, bias, res, Y0, M, sh.N, sh.K, sh.epi, 0, AC_PREC_F16, sh.out_half, nullptr));
        AC(ac_set_option("gemm_pair", g_pair));
        AC(ac_linear_tc(X, W, bias, res, Y1, M, sh.N, sh.K, sh.epi, 0, AC_PREC_F16, sh.out_half, nullptr));
        CK(cudaDeviceSynchronize());
        bad += compare(sh.name, Y0, Y1, ybytes, !sh.out_half);
    }
    // ragged shape: M not a multiple of 256, N not a multiple of 256 (edge tiles, peer CTA partly / fully out of range)
    {
        const int Mr = 300, Nr = 392, Kr = 768;
        const size_t ybytes = static_cast<size_t>(Mr) * Nr * 4;
        CK(cudaMemset(Y0, 0, ybytes)); CK(cudaMemset(Y1, 0, ybytes));
        AC(ac_set_option("gemm_pair", 0));
        AC(ac_linear_tc(X, W, bias, res, Y0, Mr, Nr, Kr, 1, 0, AC_PREC_F16, 0, nullptr));
        AC(ac_set_option("gemm_pair", g_pair));
        AC(ac_linear_tc(X, W, bias, res, Y1, Mr, Nr, Kr, 1, 0, AC_PREC_F16, 0, nullptr));
        CK(cudaDeviceSynchronize());
        bad += compare("ragged", Y0, Y1, ybytes, true);
    }
    time_variant(1);
    printf("linear: %s\n", bad == 0 ? "PAIR == 1CTA" : "MISMATCH");
    fflush(stdout);
    return bad == 0 ? 0 : 1;
}

static const char *g_knn_option = "knn_pair";   // or "knn_epi" (per-lane slow path of the scan epilogue)

static int run_knn() {
    const int B = 512, D = 768, k = 5;
    const int64_t N = 1000000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float *P = dmalloc<float>(N * D), *Q = dmalloc<float>(static_cast<size_t>(B) * D), *pn = dmalloc<float>(N);
    __half *Ph = dmalloc<__half>(N * D);
    fill_f32<<<1184, 256>>>(P, N * D, 11, 1.0f);
    normalize_rows<<<static_cast<unsigned>((N + 7) / 8), 256>>>(P, N, D);
    // queries = perturbed copies of some rows so that neighbour gaps are not degenerate
    fill_f32<<<256, 256>>>(Q, static_cast<int64_t>(B) * D, 12, 0.02f);
    CK(cudaDeviceSynchronize());
    {
        std::vector<float> hq(static_cast<size_t>(B) * D), hp(D);
        CK(cudaMemcpy(hq.data(), Q, hq.size() * 4, cudaMemcpyDeviceToHost));
        for (int b = 0; b < B; ++b) {
            CK(cudaMemcpy(hp.data(), P + (static_cast<int64_t>(b) * 1777 % N) * D, D * 4, cudaMemcpyDeviceToHost));
            for (int i = 0; i < D; ++i) hq[static_cast<size_t>(b) * D + i] += hp[i];
        }
        CK(cudaMemcpy(Q, hq.data(), hq.size() * 4, cudaMemcpyHostToDevice));
    }
    normalize_rows<<<(B + 7) / 8, 256>>>(Q, B, D);
    AC(ac_row_sqnorm(P, N, D, pn, nullptr));
    AC(ac_knn_make_shadow(P, N, D, Ph, nullptr));
    size_t wsb = 0;
    AC(ac_knn_workspace_bytes(B, N, D, k, AC_KNN_TENSOR, &wsb));
    void *ws = dmalloc<uint8_t>(wsb);
    float *d0 = dmalloc<float>(B * k), *d1 = dmalloc<float>(B * k);
    int64_t *i0 = dmalloc<int64_t>(B * k), *i1 = dmalloc<int64_t>(B * k);
    CK(cudaDeviceSynchronize());
    for (int variant = 0; variant < 2; ++variant) {
        AC(ac_set_option(g_knn_option, variant ? g_pair : 0));
        float *dd = variant ? d1 : d0; int64_t *ii = variant ? i1 : i0;
        for (int it = 0; it < 2; ++it) AC(ac_knn_l2_topk(Q, P, pn, Ph, B, N, D, k, dd, ii, 0, ws, wsb, AC_KNN_TENSOR, nullptr));
        CK(cudaDeviceSynchronize());
        AC(ac_profile_enable(1));
        CK(cudaEventRecord(e0));
        const int reps = 5;
        for (int it = 0; it < reps; ++it) AC(ac_knn_l2_topk(Q, P, pn, Ph, B, N, D, k, dd, ii, 0, ws, wsb, AC_KNN_TENSOR, nullptr));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        AC(ac_profile_enable(0));
        double ms = 0, fl = 0, by = 0; long long n = 0;
        AC(ac_profile_read(2, &ms, &fl, &by, &n));
        printf("knn variant=%s B=%d N=%lld D=%d k=%d: whole search %.3f ms; coarse scan %.3f ms/launch = %.0f GB/s algorithmic (4ND), %.0f TFLOP/s\n",
               variant ? g_knn_option : "default", B, static_cast<long long>(N), D, k, time_ms(e0, e1) / reps, ms / n, by / n / (ms / n * 1e-3) / 1e9,
               fl / n / (ms / n * 1e-3) / 1e12);
        fflush(stdout);
    }
    long long bad = compare("knn_d", d0, d1, static_cast<size_t>(B) * k * 4, true);
    bad += compare("knn_i", i0, i1, static_cast<size_t>(B) * k * 8, true);
    printf("knn: %s\n", bad == 0 ? "PAIR == 1CTA" : "MISMATCH");
    fflush(stdout);
    return bad == 0 ? 0 : 1;
}

__global__ void fill_affine(float *p, int64_t n, uint64_t seed, float base, float scale) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        p[i] = base + (static_cast<float>(hash32(seed * 0x9E3779B97F4A7C15ull + i) >> 8) * (1.f / 8388608.f) - 1.f) * scale;
}

// opt = "gemm_pair" (value g_pair; results must be bit-identical) or "ln_defer" (value 1; same math in a different
// association order, so the check is a tolerance on the unit CLS rows: the north_star bound is 1e-3 on distances)
struct EncoderWeightsOwner {
    ac_encoder_weights w{};
    std::vector<const float *> qw, qb, kw, kb, vw, vb, aow, aob, alw, alb, f1w, f1b, f2w, f2b, olw, olb;
};
// bert-base-shaped random weights; LayerNorm parameters away from (1, 0) so that the rank-1 corrections of the deferred
// flow are exercised
static void make_weights(EncoderWeightsOwner &o, int L, int H, int I, int V) {
    auto mk = [&](size_t n, uint64_t seed, float scale) { float *p = dmalloc<float>(n); fill_f32<<<592, 256>>>(p, static_cast<int64_t>(n), seed, scale); return p; };
    uint64_t lnseed = 5000;
    auto mkc = [&](size_t n, float v) { float *p = dmalloc<float>(n); fill_affine<<<64, 256>>>(p, static_cast<int64_t>(n), ++lnseed, v, v == 0.f ? 0.2f : 0.3f); return p; };
    ac_encoder_weights &w = o.w;
    w.word_emb = mk(static_cast<size_t>(V) * H, 100, 0.035f); w.pos_emb = mk(512ull * H, 101, 0.035f); w.type_emb = mk(2ull * H, 102, 0.035f);
    w.emb_ln_w = mkc(H, 1.f); w.emb_ln_b = mkc(H, 0.f);
    for (auto *v : {&o.qw, &o.qb, &o.kw, &o.kb, &o.vw, &o.vb, &o.aow, &o.aob, &o.alw, &o.alb, &o.f1w, &o.f1b, &o.f2w, &o.f2b, &o.olw, &o.olb}) v->resize(L);
    for (int l = 0; l < L; ++l) {
        const uint64_t s = 1000 + 20 * l;
        o.qw[l] = mk(static_cast<size_t>(H) * H, s + 0, 0.035f); o.qb[l] = mk(H, s + 1, 0.02f);
        o.kw[l] = mk(static_cast<size_t>(H) * H, s + 2, 0.035f); o.kb[l] = mk(H, s + 3, 0.02f);
        o.vw[l] = mk(static_cast<size_t>(H) * H, s + 4, 0.035f); o.vb[l] = mk(H, s + 5, 0.02f);
        o.aow[l] = mk(static_cast<size_t>(H) * H, s + 6, 0.035f); o.aob[l] = mk(H, s + 7, 0.02f);
        o.alw[l] = mkc(H, 1.f); o.alb[l] = mkc(H, 0.f);
        o.f1w[l] = mk(static_cast<size_t>(I) * H, s + 8, 0.035f); o.f1b[l] = mk(I, s + 9, 0.02f);
        o.f2w[l] = mk(static_cast<size_t>(H) * I, s + 10, 0.035f); o.f2b[l] = mk(H, s + 11, 0.02f);
        o.olw[l] = mkc(H, 1.f); o.olb[l] = mkc(H, 0.f);
    }
    w.q_w = o.qw.data(); w.q_b = o.qb.data(); w.k_w = o.kw.data(); w.k_b = o.kb.data(); w.v_w = o.vw.data(); w.v_b = o.vb.data();
    w.ao_w = o.aow.data(); w.ao_b = o.aob.data(); w.ao_ln_w = o.alw.data(); w.ao_ln_b = o.alb.data();
    w.ff1_w = o.f1w.data(); w.ff1_b = o.f1b.data(); w.ff2_w = o.f2w.data(); w.ff2_b = o.f2b.data(); w.out_ln_w = o.olw.data(); w.out_ln_b = o.olb.data();
    CK(cudaDeviceSynchronize());
}

// debugging aid for option "ln_defer": encoders of 1, 2, 3, 4 and 12 layers over the same weights, full hidden state of both
// flows compared -> the first depth at which they part says which epilogue (or the statistics) is wrong
static int run_defer_layers() {
    const int H = 768, I = 3072, V = 30522, B = 16, S = 128, Lmax = 12;
    EncoderWeightsOwner own;
    make_weights(own, Lmax, H, I, V);
    int32_t *ids = dmalloc<int32_t>(static_cast<size_t>(B) * S);
    fill_ids<<<64, 256>>>(ids, static_cast<int64_t>(B) * S, S, V, 7);
    const size_t nh = static_cast<size_t>(B) * S * H;
    float *h0 = dmalloc<float>(nh), *h1 = dmalloc<float>(nh), *o = dmalloc<float>(static_cast<size_t>(B) * H);
    int rc_all = 0;
    for (int L : {1, 2, 3, 4, 12}) {
        ac_encoder_config cfg{};
        cfg.arch = AC_ARCH_BERT; cfg.layers = L; cfg.hidden = H; cfg.heads = 12; cfg.intermediate = I; cfg.vocab = V; cfg.max_pos = 512;
        cfg.type_vocab = 2; cfg.pad_idx = 0; cfg.ln_eps = 1e-12f; cfg.precision = AC_PREC_F16; cfg.max_tokens = B * S; cfg.cls_only = 0;
        ac_encoder *enc = nullptr;
        AC(ac_encoder_create(&cfg, &own.w, &enc));
        AC(ac_set_option("ln_defer", 0));
        AC(ac_encoder_forward_cls(enc, ids, nullptr, nullptr, B, S, o, nullptr));
        AC(ac_encoder_last_hidden(enc, h0, static_cast<int64_t>(nh), nullptr));
        AC(ac_set_option("ln_defer", 1));
        AC(ac_encoder_forward_cls(enc, ids, nullptr, nullptr, B, S, o, nullptr));
        AC(ac_encoder_last_hidden(enc, h1, static_cast<int64_t>(nh), nullptr));
        AC(ac_set_option("ln_defer", 0));
        CK(cudaDeviceSynchronize());
        printf("layers=%d ", L);
        compare("hidden", h0, h1, nh * 4, true);       // LayerNorm outputs are O(1): expect max_abs_diff ~1e-3 or below
        AC(ac_encoder_destroy(enc));
    }
    return rc_all;
}

static int run_encoder(const char *opt, int cls_only) {
    // pair / 16-epilogue-warp kernels: same arithmetic per element; ln_defer and cls_attn reorder fp32 sums
    const bool exact = strcmp(opt, "ln_defer") != 0 && strcmp(opt, "cls_attn") != 0;
    const int optval = !strcmp(opt, "gemm_pair") ? g_pair : (!strcmp(opt, "epi16") ? 3 : 1);
    const int L = 12, H = 768, I = 3072, V = 30522, B = 512, S = 128;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    EncoderWeightsOwner own;
    make_weights(own, L, H, I, V);
    ac_encoder_weights &w = own.w;
    ac_encoder_config cfg{};
    cfg.arch = AC_ARCH_BERT; cfg.layers = L; cfg.hidden = H; cfg.heads = 12; cfg.intermediate = I; cfg.vocab = V; cfg.max_pos = 512;
    cfg.type_vocab = 2; cfg.pad_idx = 0; cfg.ln_eps = 1e-12f; cfg.precision = AC_PREC_F16; cfg.max_tokens = B * S; cfg.cls_only = cls_only;
    ac_encoder *enc = nullptr;
    AC(ac_encoder_create(&cfg, &w, &enc));
    int32_t *ids = dmalloc<int32_t>(static_cast<size_t>(B) * S);
    fill_ids<<<256, 256>>>(ids, static_cast<int64_t>(B) * S, S, V, 7);
    float *o0 = dmalloc<float>(static_cast<size_t>(B) * H), *o1 = dmalloc<float>(static_cast<size_t>(B) * H);
    CK(cudaDeviceSynchronize());
    for (int variant = 0; variant < 2; ++variant) {
        AC(ac_set_option(opt, variant ? optval : 0));
        float *o = variant ? o1 : o0;
        for (int it = 0; it < 2; ++it) AC(ac_encoder_forward_cls(enc, ids, nullptr, nullptr, B, S, o, nullptr));
        CK(cudaDeviceSynchronize());
        AC(ac_profile_enable(1));
        CK(cudaEventRecord(e0));
        const int reps = 5;
        for (int it = 0; it < reps; ++it) AC(ac_encoder_forward_cls(enc, ids, nullptr, nullptr, B, S, o, nullptr));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        AC(ac_profile_enable(0));
        double ms = 0, fl = 0, by = 0; long long n = 0;
        AC(ac_profile_read(0, &ms, &fl, &by, &n));
        printf("encoder %s=%d cls_only=%d B=%d S=%d: forward %.3f ms (%.0f seq/s); projection GEMMs %.3f ms per forward = %.0f TFLOP/s over %lld launches\n",
               opt, variant ? optval : 0, cls_only, B, S, time_ms(e0, e1) / reps, B / (time_ms(e0, e1) / reps * 1e-3), ms / reps, fl / (ms * 1e-3) / 1e12, n);
        fflush(stdout);
    }
    AC(ac_set_option(opt, 0));
    const long long bad = compare("cls_rows", o0, o1, static_cast<size_t>(B) * H * 4, true);
    if (exact) {
        printf("encoder %s: %s\n", opt, bad == 0 ? "VARIANT == DEFAULT (bit-identical)" : "MISMATCH");
        fflush(stdout);
        return bad == 0 ? 0 : 1;
    }
    // tolerance check: largest row-wise L2 distance between the two sets of unit CLS rows
    std::vector<float> h0(static_cast<size_t>(B) * H), h1(static_cast<size_t>(B) * H);
    CK(cudaMemcpy(h0.data(), o0, h0.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h1.data(), o1, h1.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0;
    bool finite = true;
    for (int b = 0; b < B; ++b) {
        double d2 = 0;
        for (int i = 0; i < H; ++i) { const double d = static_cast<double>(h0[static_cast<size_t>(b) * H + i]) - h1[static_cast<size_t>(b) * H + i]; d2 += d * d; finite &= (d == d); }
        worst = fmax(worst, sqrt(d2));
    }
    const bool ok = finite && worst < 1e-3;
    printf("encoder %s: max row ||dq||_2 = %.3e (both flows carry ~4e-4 of fp16 operand rounding against fp32) -> %s\n", opt, worst, ok ? "WITHIN 1e-3" : "OUT OF TOLERANCE");
    fflush(stdout);
    return ok ? 0 : 1;
}

// option "head_fused": one epoch of the head's training loop (batch 32, dropout 0.1, CE) through the launch-per-kernel path
// and through the fused cooperative kernel, from identical initial states: parameters and moments must match bit for bit
static int run_epoch() {
    const int D = 768, H0 = 768, H1 = 384, C = 20, n = 20000, batch = 32;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float *X = dmalloc<float>(static_cast<size_t>(n) * D);
    fill_f32<<<592, 256>>>(X, static_cast<int64_t>(n) * D, 21, 1.0f);
    normalize_rows<<<(n + 7) / 8, 256>>>(X, n, D);
    std::vector<int64_t> hy(n), hperm(n);
    for (int i = 0; i < n; ++i) { hy[i] = (i * 7 + i / 13) % C; hperm[i] = (static_cast<int64_t>(i) * 7919) % n; }
    int64_t *y = dmalloc<int64_t>(n), *perm = dmalloc<int64_t>(n);
    CK(cudaMemcpy(y, hy.data(), n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(perm, hperm.data(), n * 8, cudaMemcpyHostToDevice));
    const size_t sizes[6] = {static_cast<size_t>(H0) * D, static_cast<size_t>(H0), static_cast<size_t>(H1) * H0, static_cast<size_t>(H1),
                             static_cast<size_t>(C) * H1, static_cast<size_t>(C)};
    auto mkhead = [&](bool zero, uint64_t seed) {
        ac_head_params h{}; h.D = D; h.H0 = H0; h.H1 = H1; h.C = C;
        float **slots[6] = {&h.W0, &h.b0, &h.W1, &h.b1, &h.W2, &h.b2};
        for (int t = 0; t < 6; ++t) {
            *slots[t] = dmalloc<float>(sizes[t]);
            if (zero) CK(cudaMemset(*slots[t], 0, sizes[t] * 4));
            else fill_f32<<<128, 256>>>(*slots[t], static_cast<int64_t>(sizes[t]), seed + t, 0.05f);
        }
        return h;
    };
    ac_head_params P[2] = {mkhead(false, 300), mkhead(false, 300)}, M[2] = {mkhead(true, 0), mkhead(true, 0)}, V[2] = {mkhead(true, 0), mkhead(true, 0)};
    size_t wsb = 0;
    AC(ac_head_train_epoch_workspace_bytes(batch, &P[0], &wsb));
    void *ws = dmalloc<uint8_t>(wsb);
    float *acc = dmalloc<float>(2);
    CK(cudaMemset(acc, 0, 8));
    ac_train_cfg cfg{};
    cfg.lr = 1e-3f; cfg.beta1 = 0.9f; cfg.beta2 = 0.999f; cfg.eps = 1e-8f; cfg.weight_decay = 0.01f; cfg.max_norm = 1.0f;
    cfg.step = 1; cfg.loss_kind = AC_LOSS_CE; cfg.dropout_p = 0.1f; cfg.seed = 42;
    CK(cudaDeviceSynchronize());
    for (int variant = 0; variant < 2; ++variant) {
        AC(ac_set_option("head_fused", variant));
        CK(cudaEventRecord(e0));
        AC(ac_head_train_epoch(X, y, perm, n, batch, &P[variant], &M[variant], &V[variant], &cfg, acc + variant, ws, wsb, nullptr));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        const int steps = (n + batch - 1) / batch;
        float loss = 0; CK(cudaMemcpy(&loss, acc + variant, 4, cudaMemcpyDeviceToHost));
        printf("epoch head_fused=%d: %d steps of batch %d in %.2f ms = %.0f steps/s (%.1f us/step), mean loss %.5f\n", variant, steps, batch,
               time_ms(e0, e1), steps / (time_ms(e0, e1) * 1e-3), time_ms(e0, e1) * 1e3 / steps, loss / steps);
        fflush(stdout);
    }
    AC(ac_set_option("head_fused", 0));
    long long bad = 0;
    float *pa[6] = {P[0].W0, P[0].b0, P[0].W1, P[0].b1, P[0].W2, P[0].b2}, *pb[6] = {P[1].W0, P[1].b0, P[1].W1, P[1].b1, P[1].W2, P[1].b2};
    float *va[6] = {V[0].W0, V[0].b0, V[0].W1, V[0].b1, V[0].W2, V[0].b2}, *vb[6] = {V[1].W0, V[1].b0, V[1].W1, V[1].b1, V[1].W2, V[1].b2};
    const char *names[6] = {"W0", "b0", "W1", "b1", "W2", "b2"};
    for (int t = 0; t < 6; ++t) {
        bad += compare(names[t], pa[t], pb[t], sizes[t] * 4, true);
        bad += compare("  adam v", va[t], vb[t], sizes[t] * 4, true);
    }
    printf("epoch: %s\n", bad == 0 ? "FUSED == LAUNCH-PER-KERNEL" : "MISMATCH");
    fflush(stdout);
    return bad == 0 ? 0 : 1;
}

int main(int argc, char **argv) {
    setvbuf(stdout, nullptr, _IOLBF, 0);
    if (argc < 2) { printf("usage: %s linear|knn|encoder|defer|defer_full|defer_layers|epoch|epi16|attn|pdl|knn_epi|cls_attn [pair option value: 1 (default) | 2 = relay variant]\n", argv[0]); return 64; }
    if (argc > 2) g_pair = atoi(argv[2]);
    AC(ac_device_check());
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    printf("device: %s, %d SMs, ABI v%d, test %s, pair option %d\n", pr.name, pr.multiProcessorCount, ac_version(), argv[1], g_pair);
    if (!strcmp(argv[1], "linear")) return run_linear();
    if (!strcmp(argv[1], "knn")) return run_knn();
    if (!strcmp(argv[1], "knn_epi")) { g_knn_option = "knn_epi"; g_pair = 1; return run_knn(); }
    if (!strcmp(argv[1], "encoder")) return run_encoder("gemm_pair", 1);
    if (!strcmp(argv[1], "epoch")) return run_epoch();
    if (!strcmp(argv[1], "defer_layers")) return run_defer_layers();
    if (!strcmp(argv[1], "epi16")) return run_encoder("epi16", 1);             // FFN1 + QKV with 16 epilogue warps
    if (!strcmp(argv[1], "attn")) return run_encoder("attn_pipe", 1);          // persistent pipelined attention
    if (!strcmp(argv[1], "pdl")) {                                             // programmatic dependent launch on the opted-in chain
        AC(ac_set_option("gemm_pair", 1)); AC(ac_set_option("ln_defer", 1)); AC(ac_set_option("attn_pipe", 1));
        return run_encoder("pdl", 1);                                          // same kernels, only the launch attribute changes
    }
    if (!strcmp(argv[1], "cls_attn")) return run_encoder("cls_attn", 1);       // last layer: attention on the CLS query row only
    if (!strcmp(argv[1], "defer")) return run_encoder("ln_defer", 1);          // production shape: CLS-only tail
    if (!strcmp(argv[1], "defer_full")) return run_encoder("ln_defer", 0);     // every layer through the deferred epilogues
    printf("unknown test %s\n", argv[1]);
    return 64;
}
