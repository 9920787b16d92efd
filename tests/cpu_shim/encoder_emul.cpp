// encoder_emul.cpp -- the HOST logic of csrc/encoder.cu (ac_encoder_create, ac_encoder_forward_cls, the deferred-LayerNorm
// layer loop, the CLS-only tail) executed on the CPU.  TEST INFRASTRUCTURE ONLY.
//
// tests/cpu_shim/extract_device_code.py --host keeps the host functions of encoder.cu and rewrites every <<<...>>> launch
// into a shim launch; this file supplies a fake CUDA runtime (cudaMalloc = malloc ...), a fake CUtensorMap that simply
// remembers the matrix it describes, and stand-ins for what needs the real hardware:
//   * the tensor-core GEMM mainloop: the accumulator is computed here (fp16 operands, fp32 sums) and handed, tile by tile,
//     to the REAL epilogue functors in the thread numbering of gemm_tc_kernel / gemm_tc2_kernel;
//   * the attention kernels: a plain per-(sequence, head) softmax(QK^T / 8 + mask) V with the kernels' rounding points.
// Everything else (weight packing, embeddings, LayerNorm kernels, statistics, which buffer / gamma / beta / statistics
// array feeds which launch) is the product code.  The Python test compares the unit CLS rows with the fp32 oracle for the
// default flow and for the opt-in variants (ln_defer, epi16) that were written without GPU access.
//   usage: encoder_emul <input.bin> <output.bin>
#include "cuda_shim.h"
#include "../../include/adaptive_b200.h"
#include <cstdarg>
#include <string>

// ---------------------------------------------------------------- fake CUDA runtime
// (cuda_fp16.h already pulled in the declarations of cuda_runtime_api.h: these are the definitions the linker will use)
extern "C" {
cudaError_t cudaMalloc(void **p, size_t n) {
    const size_t bytes = (n + 255) / 256 * 256 + 256;
    *p = aligned_alloc(256, bytes);
    memset(*p, 0xCD, bytes);                        // uninitialised device memory is garbage, not zeros
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, enum cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return "fake"; }
}
template <class F> static cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// a tensor map that remembers what it describes
typedef struct CUtensorMap_st {
    const void *ptr;
    int elem_bytes;
    uint64_t rows, cols, stride_bytes;
    uint32_t box_rows, box_cols;
} CUtensorMap;
#define __grid_constant__

static void __threadfence_system() {}
static unsigned int atomicAdd(unsigned int *p, unsigned int v) { unsigned int o = *p; *p += v; return o; }

static long long g_options[16] = {0};
static char g_err[512];
namespace ac {
static inline float ex2_approx(float x) { return exp2f(x); }
static inline float rcp_approx(float x) { return 1.f / x; }
static inline void griddep_wait() {}
static inline void griddep_launch_dependents() {}
static inline void st_release_sys(uint32_t *p, uint32_t v) { *p = v; }
static inline uint32_t ld_acquire_sys(const uint32_t *p) { return *p; }
}  // namespace ac

#define AC_CUDA(call) do { if ((call) != cudaSuccess) return AC_E_CUDA; } while (0)
#define AC_REQUIRE(cond, ...) do { if (!(cond)) { ::ac::set_error(__VA_ARGS__); return AC_E_INVALID; } } while (0)
#define AC_LAUNCH_CHECK() do { ::ac::count_launch(); } while (0)

#include "_gen_common_host.inc"
namespace ac {
void set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap); }
int check_cuda(cudaError_t e, const char *) { return e == cudaSuccess ? AC_OK : AC_E_CUDA; }
static long long g_launches = 0;
void count_launch() { ++g_launches; }
int prof_begin(int, double, double, cudaStream_t) { return -1; }
void prof_end(int, cudaStream_t) {}
long long option(int id) { return g_options[id]; }
int sm_count() { return 4; }
int make_tmap_2d(CUtensorMap *out, const void *gptr, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                 uint32_t box_rows, uint32_t box_cols) {
    if (box_cols * static_cast<uint32_t>(elem_bytes) != 128 || box_rows > 256 || row_stride_bytes % 16 != 0) { set_error("bad tensor map"); return AC_E_INVALID; }
    *out = CUtensorMap{gptr, elem_bytes, rows, cols, row_stride_bytes, box_rows, box_cols};
    return AC_OK;
}
template <class... KA, class... A>
static inline cudaError_t launch_maybe_pdl(void (*kern)(KA...), dim3 grid, dim3 block, size_t, cudaStream_t, bool, A &&...a) {
    shim::launch(grid, block, [&] { kern(static_cast<KA>(a)...); });
    return cudaSuccess;
}
}  // namespace ac
extern "C" int ac_device_check(void) { return 0; }

struct LaunchCfg { dim3 g, b; };
static inline LaunchCfg shim_cfg(dim3 g, dim3 b, size_t = 0, cudaStream_t = nullptr) { return {g, b}; }
#define SHIM_CFG(...) shim_cfg(__VA_ARGS__)
template <class... KA, class... A>
static inline void shim_launch(LaunchCfg c, void (*k)(KA...), A &&...a) {
    shim::launch(c.g, c.b, [&] { k(static_cast<KA>(a)...); });
}

#include "_gen_gemm_tc.inc"
namespace ac {
constexpr int GEMM2_B_ROWS = GEMM_BLOCK_N / 2;

// ---------------------------------------------------------------- stand-in for the tensor-core mainloop
static float h2f(const void *base, uint64_t stride_bytes, uint64_t rows, uint64_t cols, int64_t r, int64_t c) {
    if (r < 0 || static_cast<uint64_t>(r) >= rows || static_cast<uint64_t>(c) >= cols) return 0.f;       // TMA zero fill
    return __half2float(*reinterpret_cast<const __half *>(static_cast<const uint8_t *>(base) + r * stride_bytes + c * 2));
}
template <class Epi, int kEpiWarps>
static int emul_gemm(const CUtensorMap &ta, const CUtensorMap &tb, int M, int N, int K, const Epi &epi, uint32_t want_b_box) {
    if (ta.elem_bytes != 2 || tb.elem_bytes != 2) { set_error("emulation handles the fp16 GEMMs"); return AC_E_UNSUPPORTED; }
    if (tb.box_rows != want_b_box || ta.box_rows != GEMM_BLOCK_M) { set_error("tensor map box %u does not fit this kernel (wants %u)", tb.box_rows, want_b_box); return AC_E_INVALID; }
    const int Mp = (M + 127) / 128 * 128, Np = (N + 255) / 256 * 256;
    std::vector<float> acc(static_cast<size_t>(Mp) * Np), Af(static_cast<size_t>(Mp) * K), Bf(static_cast<size_t>(Np) * K);
    for (int m = 0; m < Mp; ++m)
        for (int k = 0; k < K; ++k) Af[static_cast<size_t>(m) * K + k] = h2f(ta.ptr, ta.stride_bytes, ta.rows, ta.cols, m, k);
    for (int n = 0; n < Np; ++n)
        for (int k = 0; k < K; ++k) Bf[static_cast<size_t>(n) * K + k] = h2f(tb.ptr, tb.stride_bytes, tb.rows, tb.cols, n, k);
    for (int m = 0; m < Mp; ++m)
        for (int n = 0; n < Np; ++n) {
            float s = 0.f;
            const float *ar = &Af[static_cast<size_t>(m) * K], *br = &Bf[static_cast<size_t>(n) * K];
            for (int k = 0; k < K; ++k) s = fmaf(ar[k], br[k], s);
            acc[static_cast<size_t>(m) * Np + n] = s;
        }
    static uint8_t epi_stage[16 * GEMM_EPI_STAGE_BYTES];
    constexpr int kCols = GEMM_BLOCK_N / (kEpiWarps / 4);
    int it = 0;
    for (int m0 = 0; m0 < M; m0 += GEMM_BLOCK_M)
        for (int n0 = 0; n0 < N; n0 += GEMM_BLOCK_N, ++it)
            shim::launch(dim3(1), dim3(64 + 32 * kEpiWarps), [&] {
                const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
                if (warp < 2) return;
                const int q = warp & 3, cpart = (warp - 2) >> 2;
                typename Epi::State est;
                epi.begin_cta(est, q, lane);
                GemmTileInfo ti;
                ti.m0 = m0; ti.n0 = n0; ti.tile_iter = it;
                const int row = m0 + q * 32 + lane, c_lo = cpart * kCols;
                epi.prefetch(est, ti, row, n0 + c_lo, lane, 0);
                for (int ci = 0; ci < kCols / 32; ++ci) {
                    const int c = c_lo + 32 * ci;
                    if (ci + 1 < kCols / 32) epi.prefetch(est, ti, row, n0 + c + 32, lane, (ci + 1) & 1);
                    float v[32];
                    for (int j = 0; j < 32; ++j) v[j] = acc[static_cast<size_t>(row) * Np + n0 + c + j];
                    epi.tile(est, ti, row, n0 + c, v, epi_stage + (warp - 2) * GEMM_EPI_STAGE_BYTES, lane, ci & 1, 0u);
                }
                epi.end_cta(est, q, lane);
            }, 11 + it);
    count_launch();
    return AC_OK;
}
template <class Epi, bool kMFastest = false, int kKind = GEMM_KIND_TF32>
int launch_gemm_tc(const CUtensorMap &ta, const CUtensorMap &tb, int M, int N, int K, const Epi &epi, cudaStream_t, int = 0, int = 0, double = 0.0) {
    return emul_gemm<Epi, 8>(ta, tb, M, N, K, epi, GEMM_BLOCK_N);
}
template <class Epi, bool kMFastest = false, int kKind = GEMM_KIND_TF32, int kEpiWarps = GEMM_EPI_WARPS>
int launch_gemm_tc2(const CUtensorMap &ta, const CUtensorMap &tb, int M, int N, int K, const Epi &epi, cudaStream_t, int = 0, int = 0, double = 0.0) {
    return emul_gemm<Epi, kEpiWarps>(ta, tb, M, N, K, epi, GEMM2_B_ROWS);
}

// ---------------------------------------------------------------- stand-ins for the attention kernels (one thread per item)
static void attention_item(const CUtensorMap &tqk, const CUtensorMap &tvt, const int32_t *mask, int b, int h, int S, int H, __half *ctx) {
    const int S_pad = static_cast<int>(tvt.cols);
    const int64_t row0 = static_cast<int64_t>(b) * S;
    const float scale_log2 = (1.f / sqrtf(64.f)) * 1.44269504088896340736f;
    std::vector<float> sc(S), p(S);
    for (int qi = 0; qi < S; ++qi) {
        float mx = -INFINITY;
        for (int k = 0; k < S; ++k) {
            float s = 0.f;
            for (int d = 0; d < 64; ++d)
                s = fmaf(h2f(tqk.ptr, tqk.stride_bytes, tqk.rows, tqk.cols, row0 + qi, h * 64 + d), h2f(tqk.ptr, tqk.stride_bytes, tqk.rows, tqk.cols, row0 + k, H + h * 64 + d), s);
            sc[k] = s;
            if (!mask || mask[row0 + k] != 0) mx = fmaxf(mx, s);
        }
        float sum = 0.f;
        for (int k = 0; k < S; ++k) {
            const bool ok = !mask || mask[row0 + k] != 0;
            const float e = ok ? exp2f(fmaf(sc[k], scale_log2, -mx * scale_log2)) : 0.f;
            sum += e;
            p[k] = __half2float(__float2half_rn(e));
        }
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
        for (int d = 0; d < 64; ++d) {
            float o = 0.f;
            for (int k = 0; k < S; ++k)
                o = fmaf(p[k], h2f(tvt.ptr, tvt.stride_bytes, tvt.rows, S_pad, static_cast<int64_t>(b) * H + h * 64 + d, k), o);
            ctx[(row0 + qi) * H + h * 64 + d] = __float2half_rn(o * inv);
        }
    }
}
void attention_kernel(const CUtensorMap tqk, const CUtensorMap tvt, const int32_t *mask, int B, int S, int heads, int H, __half *ctx) {
    if (threadIdx.x == 0) attention_item(tqk, tvt, mask, blockIdx.x / heads, blockIdx.x % heads, S, H, ctx);
}
void attention_long_kernel(const CUtensorMap tqk, const CUtensorMap tvt, const int32_t *mask, int B, int S, int heads, int H, __half *ctx) {
    if (threadIdx.x == 0 && blockIdx.y == 0) attention_item(tqk, tvt, mask, blockIdx.x / heads, blockIdx.x % heads, S, H, ctx);
}
void attention_pipe_kernel(const CUtensorMap tqk, const CUtensorMap tvt, const int32_t *mask, int B, int S, int heads, int H, __half *ctx) {
    if (threadIdx.x == 0)
        for (int item = blockIdx.x; item < B * heads; item += gridDim.x) attention_item(tqk, tvt, mask, item / heads, item % heads, S, H, ctx);
}
}  // namespace ac

#include "_gen_peer.inc"
#include "_gen_encoder_host.inc"

// ---------------------------------------------------------------- driver
template <class T> static std::vector<T> rd(FILE *f, size_t n) { std::vector<T> v(n); if (fread(v.data(), sizeof(T), n, f) != n) { printf("short read\n"); exit(2); } return v; }

int main(int argc, char **argv) {
    if (argc < 3) return 64;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 65;
    const std::vector<int32_t> hd = rd<int32_t>(f, 12);
    const int L = hd[0], H = hd[1], heads = hd[2], I = hd[3], V = hd[4], maxpos = hd[5], typev = hd[6], B = hd[7], S = hd[8], use_mask = hd[9], arch = hd[10];
    const float eps = *reinterpret_cast<const float *>(&hd[11]);
    std::vector<std::vector<float>> keep;
    auto take = [&](size_t n) { keep.push_back(rd<float>(f, n)); return const_cast<const float *>(keep.back().data()); };
    ac_encoder_weights w{};
    w.word_emb = take(size_t(V) * H); w.pos_emb = take(size_t(maxpos) * H); w.type_emb = take(size_t(typev) * H);
    w.emb_ln_w = take(H); w.emb_ln_b = take(H);
    std::vector<const float *> a[16];
    for (int l = 0; l < L; ++l) {
        const size_t sz[16] = {size_t(H) * H, size_t(H), size_t(H) * H, size_t(H), size_t(H) * H, size_t(H), size_t(H) * H, size_t(H), size_t(H), size_t(H),
                               size_t(I) * H, size_t(I), size_t(H) * I, size_t(H), size_t(H), size_t(H)};
        for (int t = 0; t < 16; ++t) a[t].push_back(take(sz[t]));
    }
    w.q_w = a[0].data(); w.q_b = a[1].data(); w.k_w = a[2].data(); w.k_b = a[3].data(); w.v_w = a[4].data(); w.v_b = a[5].data();
    w.ao_w = a[6].data(); w.ao_b = a[7].data(); w.ao_ln_w = a[8].data(); w.ao_ln_b = a[9].data();
    w.ff1_w = a[10].data(); w.ff1_b = a[11].data(); w.ff2_w = a[12].data(); w.ff2_b = a[13].data(); w.out_ln_w = a[14].data(); w.out_ln_b = a[15].data();
    const std::vector<int32_t> ids = rd<int32_t>(f, size_t(B) * S), mask = rd<int32_t>(f, size_t(B) * S);
    fclose(f);

    FILE *out = fopen(argv[2], "wb");
    // variants: {cls_only, ln_defer, epi16, gemm_pair, cls_attn}
    const int variants[][5] = {{0, 0, 0, 0, 0}, {1, 0, 0, 0, 0}, {0, 1, 0, 0, 0}, {1, 1, 0, 0, 0}, {1, 1, 3, 0, 0}, {1, 0, 1, 1, 0}, {0, 1, 3, 1, 0},
                               {1, 0, 0, 0, 1}, {1, 1, 0, 0, 1}, {0, 0, 0, 0, 1}};
    for (const auto &v : variants) {
        ac_encoder_config cfg{};
        cfg.arch = arch; cfg.layers = L; cfg.hidden = H; cfg.heads = heads; cfg.intermediate = I; cfg.vocab = V; cfg.max_pos = maxpos;
        cfg.type_vocab = typev; cfg.pad_idx = arch == AC_ARCH_ROBERTA ? 1 : 0; cfg.ln_eps = eps; cfg.precision = AC_PREC_F16; cfg.max_tokens = B * S;
        cfg.cls_only = v[0];
        g_options[ac::OPT_LN_DEFER] = v[1]; g_options[ac::OPT_EPI16] = v[2]; g_options[ac::OPT_GEMM_PAIR] = v[3];
        g_options[ac::OPT_CLS_ATTN] = v[4];
        ac_encoder *enc = nullptr;
        int rc = ac_encoder_create(&cfg, &w, &enc);
        if (rc) { printf("create failed: %s\n", g_err); return 3; }
        std::vector<float> cls(size_t(B) * H, -5.f), hidden(size_t(B) * S * H, -5.f);
        rc = ac_encoder_forward_cls(enc, ids.data(), use_mask ? mask.data() : nullptr, nullptr, B, S, cls.data(), nullptr);
        if (rc) { printf("forward failed: %s\n", g_err); return 4; }
        int32_t have_hidden = 0;
        if (!v[0]) { rc = ac_encoder_last_hidden(enc, hidden.data(), int64_t(B) * S * H, nullptr); have_hidden = rc == 0; }
        const int32_t tag[6] = {v[0], v[1], v[2], v[3], v[4], have_hidden};
        fwrite(tag, 4, 6, out);
        fwrite(cls.data(), 4, cls.size(), out);
        fwrite(hidden.data(), 4, hidden.size(), out);
        ac_encoder_destroy(enc);
        printf("variant cls_only=%d ln_defer=%d epi16=%d gemm_pair=%d cls_attn=%d: %lld launches\n", v[0], v[1], v[2], v[3], v[4], ac::g_launches);
        ac::g_launches = 0;
    }
    // ac_encoder_forward_cls_scatter: the final normalise kernel also stores the rows into every peer's buffer + flags
    {
        ac_encoder_config cfg{};
        cfg.arch = arch; cfg.layers = L; cfg.hidden = H; cfg.heads = heads; cfg.intermediate = I; cfg.vocab = V; cfg.max_pos = maxpos;
        cfg.type_vocab = typev; cfg.pad_idx = 0; cfg.ln_eps = eps; cfg.precision = AC_PREC_F16; cfg.max_tokens = B * S; cfg.cls_only = 1;
        for (auto &o : g_options) o = 0;
        ac_encoder *enc = nullptr;
        if (ac_encoder_create(&cfg, &w, &enc)) return 5;
        const int G = 3, rank = 2;
        const size_t off = 512;
        std::vector<std::vector<float>> peer(G, std::vector<float>(off / 4 + size_t(B) * H + 16, -9.f));
        std::vector<std::vector<uint32_t>> flags(G, std::vector<uint32_t>(G, 0));
        ac_peer_table t{};
        t.world = G; t.rank = rank;
        for (int p = 0; p < G; ++p) { t.buf[p] = peer[p].data(); t.flag[p] = flags[p].data(); }
        uint32_t counter = 0;
        std::vector<float> cls(size_t(B) * H, -5.f), ref(size_t(B) * H, -5.f);
        if (ac_encoder_forward_cls(enc, ids.data(), use_mask ? mask.data() : nullptr, nullptr, B, S, ref.data(), nullptr)) return 6;
        if (ac_encoder_forward_cls_scatter(enc, ids.data(), use_mask ? mask.data() : nullptr, nullptr, B, S, cls.data(), &t, off, 17u, &counter, nullptr)) return 7;
        int bad = memcmp(cls.data(), ref.data(), cls.size() * 4) != 0;
        for (int p = 0; p < G; ++p) {
            bad += memcmp(peer[p].data() + off / 4, ref.data(), ref.size() * 4) != 0;
            bad += peer[p][off / 4 - 1] != -9.f || peer[p][off / 4 + ref.size()] != -9.f;
            for (int r = 0; r < G; ++r) bad += flags[p][r] != (r == rank ? 17u : 0u);
        }
        // the sink must not leak into the next ordinary forward
        std::fill(peer[0].begin(), peer[0].end(), -9.f);
        if (ac_encoder_forward_cls(enc, ids.data(), use_mask ? mask.data() : nullptr, nullptr, B, S, cls.data(), nullptr)) return 8;
        for (float x : peer[0]) bad += x != -9.f;
        printf("forward_cls_scatter: %s\n", bad ? "FAIL" : "ok");
        ac_encoder_destroy(enc);
        if (bad) return 9;
    }
    // attention_cls_kernel (one query row per (sequence, head)) against the stand-in of the full kernel, on the Q | K and V^T
    // buffers the last layer of a full forward left behind: CLS rows bit for bit, every other row of ctx untouched
    {
        ac_encoder_config cfg{};
        cfg.arch = arch; cfg.layers = L; cfg.hidden = H; cfg.heads = heads; cfg.intermediate = I; cfg.vocab = V; cfg.max_pos = maxpos;
        cfg.type_vocab = typev; cfg.pad_idx = 0; cfg.ln_eps = eps; cfg.precision = AC_PREC_F16; cfg.max_tokens = B * S; cfg.cls_only = 0;
        for (auto &o : g_options) o = 0;
        ac_encoder *enc = nullptr;
        if (ac_encoder_create(&cfg, &w, &enc)) return 10;
        std::vector<float> cls(size_t(B) * H);
        if (ac_encoder_forward_cls(enc, ids.data(), use_mask ? mask.data() : nullptr, nullptr, B, S, cls.data(), nullptr)) return 11;
        const int S_pad = (S + 7) / 8 * 8;
        int bad = 0;
        for (int round = 0; round < 3; ++round) {
            std::vector<int32_t> m2 = mask;
            if (round == 1) for (int k = 0; k < S; ++k) m2[size_t(B > 1 ? 1 : 0) * S + k] = 0;      // one sequence with no valid key
            const int32_t *mp = round == 2 ? nullptr : m2.data();
            std::vector<__half> full(size_t(B) * S * H), one(size_t(B) * S * H);
            memset(full.data(), 0x7b, full.size() * 2);
            memset(one.data(), 0x7b, one.size() * 2);
            shim::launch(dim3(B * heads), dim3(128), [&] { ac::attention_kernel(enc->m_qk_att, enc->m_vt_att, mp, B, S, heads, H, full.data()); });
            shim::launch(dim3((B * heads + ac::ATTC_WARPS - 1) / ac::ATTC_WARPS), dim3(ac::ATTC_WARPS * 32),
                         [&] { ac::attention_cls_kernel(enc->qk, enc->vT, mp, B, S, S_pad, heads, H, one.data()); });
            for (int b = 0; b < B; ++b)
                for (int t = 0; t < S; ++t) {
                    const __half *x = &one[(size_t(b) * S + t) * H], *y = &full[(size_t(b) * S + t) * H];
                    if (t == 0) bad += memcmp(x, y, size_t(H) * 2) != 0;
                    else
                        for (int j = 0; j < H; ++j) { uint16_t u; memcpy(&u, &x[j], 2); bad += u != 0x7b7b; }
                }
        }
        printf("attention_cls_kernel == attention stand-in on the CLS rows: %s\n", bad ? "FAIL" : "ok");
        ac_encoder_destroy(enc);
        if (bad) return 12;
    }
    fclose(out);
    return 0;
}
