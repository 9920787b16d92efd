// attention_emul.cpp -- attention_kernel (verified on a B200) and attention_pipe_kernel (written without GPU access) executed
// on the functional Blackwell model of tests/cpu_shim/tc_emul.h.  TEST INFRASTRUCTURE ONLY.
//   1. attention_kernel through the model must reproduce a plain softmax(QK^T / 8 + mask) V reference: this validates the
//      MODEL (TMA swizzle, SWIZZLE_128B descriptors, TMEM addressing) against a kernel that is known to be right.
//   2. attention_pipe_kernel, for several grid sizes (1 CTA walking all items ... one item per CTA), must give the SAME BITS
//      as attention_kernel: same arithmetic per (sequence, head), only the pipelining differs.
//   3. (bonus) gemm_tc_kernel with the deferred-LayerNorm GELU epilogue runs through its real TMA / MMA / epilogue warp roles.
#include "tc_emul.h"
#include "../../include/adaptive_b200.h"

static void __threadfence_system() {}
static unsigned int atomicAdd(unsigned int *p, unsigned int v) { unsigned int o = *p; *p += v; return o; }
namespace ac {
static inline void st_release_sys(uint32_t *p, uint32_t v) { *p = v; }
static inline uint32_t ld_acquire_sys(const uint32_t *p) { return *p; }
static inline float ex2_approx(float x) { return exp2f(x); }
static inline float rcp_approx(float x) { return 1.f / x; }
static inline void griddep_wait() {}
static inline void griddep_launch_dependents() {}
}  // namespace ac

#include "_gen_common_tc.inc"
#include "_gen_gemm_tc_tc.inc"
#include "_gen_peer.inc"
#include "_gen_encoder_tc.inc"

using namespace ac;

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (g_fail < 20) { printf("  FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } ++g_fail; } } while (0)
static std::mt19937 g_rng(7);
static float urand(float s) { return std::uniform_real_distribution<float>(-s, s)(g_rng); }
static uint16_t hbits(__half h) { uint16_t u; memcpy(&u, &h, 2); return u; }

static void test_attention(int B, int S, int heads, bool use_mask) {
    const int H = heads * 64, T = B * S, S_pad = (S + 7) / 8 * 8;
    std::vector<__half> qk(static_cast<size_t>(T) * 2 * H), vT(static_cast<size_t>(B) * H * S_pad);
    for (auto &h : qk) h = __float2half_rn(urand(1.5f));
    for (auto &h : vT) h = __float2half_rn(urand(1.0f));
    std::vector<int32_t> mask(static_cast<size_t>(T), 1);
    if (use_mask)
        for (int b = 0; b < B; ++b)
            for (int s = S - 1 - 3 * b; s < S; ++s) if (s > 0) mask[static_cast<size_t>(b) * S + s] = 0;
    CUtensorMap tqk{qk.data(), 2, static_cast<uint64_t>(T), static_cast<uint64_t>(2 * H), static_cast<uint64_t>(2 * H) * 2, 128, 64};
    CUtensorMap tvt{vT.data(), 2, static_cast<uint64_t>(B) * H, static_cast<uint64_t>(S_pad), static_cast<uint64_t>(S_pad) * 2, 64, 64};
    const uint16_t sentinel = 0x7bad;
    auto fresh = [&] { std::vector<__half> c(static_cast<size_t>(T) * H); for (auto &h : c) memcpy(&h, &sentinel, 2); return c; };
    const int32_t *mp = use_mask ? mask.data() : nullptr;

    std::vector<__half> ctx0 = fresh();
    shim::launch(dim3(B * heads), dim3(ATT_THREADS), [&] { attention_kernel(tqk, tvt, mp, B, S, heads, H, ctx0.data()); });
    // (1) the model + the verified kernel against the plain reference
    double worst = 0;
    for (int b = 0; b < B; ++b)
        for (int h = 0; h < heads; ++h)
            for (int qi = 0; qi < S; ++qi) {
                std::vector<double> sc(S);
                double mx = -1e30;
                for (int k = 0; k < S; ++k) {
                    double s = 0;
                    for (int d = 0; d < 64; ++d)
                        s += static_cast<double>(__half2float(qk[(static_cast<size_t>(b) * S + qi) * 2 * H + h * 64 + d])) * __half2float(qk[(static_cast<size_t>(b) * S + k) * 2 * H + H + h * 64 + d]);
                    sc[k] = s / 8.0;
                    if (!mp || mp[b * S + k]) mx = std::max(mx, sc[k]);
                }
                double sum = 0;
                for (int k = 0; k < S; ++k) { sc[k] = (!mp || mp[b * S + k]) ? exp(sc[k] - mx) : 0.0; sum += sc[k]; }
                for (int d = 0; d < 64; ++d) {
                    double o = 0;
                    for (int k = 0; k < S; ++k) o += sc[k] * __half2float(vT[(static_cast<size_t>(b) * H + h * 64 + d) * S_pad + k]);
                    const double got = __half2float(ctx0[(static_cast<size_t>(b) * S + qi) * H + h * 64 + d]);
                    worst = std::max(worst, fabs(got - o / sum));
                }
            }
    CHECK(worst < 3e-3, "attention_kernel on the model deviates from the reference by %g", worst);
    // (2) the pipelined kernel, several grid sizes
    const int items = B * heads;
    for (int G : {1, 2, 3, items}) {
        if (G > items) continue;
        std::vector<__half> ctx1 = fresh();
        shim::launch(dim3(G), dim3(ATTP_THREADS), [&] { attention_pipe_kernel(tqk, tvt, mp, B, S, heads, H, ctx1.data()); });
        long long diff = 0;
        for (size_t i = 0; i < ctx0.size(); ++i) diff += hbits(ctx0[i]) != hbits(ctx1[i]);
        CHECK(diff == 0, "attention_pipe_kernel with %d CTAs: %lld of %zu outputs differ from attention_kernel", G, diff, ctx0.size());
    }
    printf("attention B=%d S=%d heads=%d mask=%d: model vs reference %.2e; attention_pipe == attention_kernel for 1/2/3/%d CTAs: %s\n",
           B, S, heads, use_mask, worst, items, g_fail ? "FAIL" : "ok");
}

// (3) the real gemm_tc_kernel (TMA producer warp, MMA issuer, 8 epilogue warps, double-buffered accumulators) with the
// deferred-LayerNorm GELU epilogue, persistent over several tiles on 2 CTAs
static void test_gemm_kernel() {
    using Epi = EpiLinear<1, true, false, true>;
    const int M = 300, N = 392, K = 128;
    std::vector<__half> A(static_cast<size_t>(M) * K), W(static_cast<size_t>(N) * K);
    for (auto &h : A) h = __float2half_rn(urand(1.f));
    for (auto &h : W) h = __float2half_rn(urand(0.2f));
    std::vector<float> c0(512), c1(512);
    std::vector<float2> stats(512);
    for (auto &x : c0) x = urand(0.5f);
    for (auto &x : c1) x = urand(1.f);
    for (auto &s : stats) s = make_float2(urand(0.3f), 0.5f + fabsf(urand(1.f)));
    const uint16_t sentinel = 0x7bad;
    std::vector<__half> Y(static_cast<size_t>(M + 8) * N);
    for (auto &h : Y) memcpy(&h, &sentinel, 2);
    CUtensorMap ta{A.data(), 2, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(K) * 2, 128, 64};
    CUtensorMap tb{W.data(), 2, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(K) * 2, 256, 64};
    Epi epi{c0.data(), nullptr, Y.data(), M, N, N, 0, nullptr, 0, 0, 0, 0, c1.data(), stats.data()};
    shim::launch(dim3(2), dim3(GEMM_THREADS), [&] { gemm_tc_kernel<Epi, false, GEMM_KIND_F16>(ta, tb, M, N, K, epi); });
    double worst = 0;
    for (int m = 0; m < M + 8; ++m)
        for (int n = 0; n < N; ++n) {
            const __half got = Y[static_cast<size_t>(m) * N + n];
            if (m >= M) { CHECK(hbits(got) == sentinel, "gemm: row %d beyond M written", m); continue; }
            double acc = 0;
            for (int k = 0; k < K; ++k) acc += static_cast<double>(__half2float(A[static_cast<size_t>(m) * K + k])) * __half2float(W[static_cast<size_t>(n) * K + k]);
            const float y = gelu_erf(static_cast<float>(stats[m].y * (acc - stats[m].x * c1[n]) + c0[n]));
            worst = std::max(worst, fabs(static_cast<double>(__half2float(got)) - y));
        }
    CHECK(worst < 5e-3, "gemm_tc_kernel + deferred GELU epilogue deviates by %g", worst);
    printf("gemm_tc_kernel<EpiLinear<GELU, DEFER>> M=%d N=%d K=%d on 2 persistent CTAs: max abs deviation %.2e: %s\n", M, N, K, worst, g_fail ? "FAIL" : "ok");
}

int main() {
    test_attention(3, 128, 2, false);
    test_attention(5, 50, 2, true);          // ragged S (keys beyond S masked, zero-filled tiles), padding masks, odd item count
    test_attention(2, 17, 1, true);
    test_gemm_kernel();
    printf("attention_emul: %s (%d failed checks)\n", g_fail ? "FAIL" : "ALL OK", g_fail);
    return g_fail ? 1 : 0;
}
