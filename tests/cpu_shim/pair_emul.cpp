// pair_emul.cpp -- gemm_tc2_kernel (CTA pair, tcgen05 cta_group::2) on the functional Blackwell model of tc_emul.h with a
// cluster of two concurrently running CTAs.  TEST INFRASTRUCTURE ONLY.
// The 8-epilogue-warp pair kernel ran bit-identically to the 1-CTA kernel on a B200; in the model it must therefore ALSO be
// bit-identical to gemm_tc_kernel (which validates the model's reading of the pair semantics: A rows / accumulator lanes and
// B row halves split across the two CTAs at equal shared-memory offsets, bytes counted on the leader's barrier, multicast
// commits, remote arrives).  The variants that never ran on hardware are then held to the same bits: 16 epilogue warps with 5
// stages, the relay variant, and the deferred-LayerNorm functors with 64 columns per warp.
#include "tc_emul.h"
#include "../../include/adaptive_b200.h"

static void __threadfence_system() {}
static unsigned int atomicAdd(unsigned int *p, unsigned int v) { unsigned int o = *p; *p += v; return o; }
namespace ac {
static inline float ex2_approx(float x) { return exp2f(x); }
static inline float rcp_approx(float x) { return 1.f / x; }
static inline void griddep_wait() {}
static inline void griddep_launch_dependents() {}
static inline void st_release_sys(uint32_t *p, uint32_t v) { *p = v; }
static inline uint32_t ld_acquire_sys(const uint32_t *p) { return *p; }
}  // namespace ac

#include "_gen_common_tc.inc"
#include "_gen_gemm_tc_tc.inc"
#include "_gen_gemm_tc2_tc.inc"
#include "_gen_peer.inc"
#include "_gen_encoder_tc.inc"

using namespace ac;

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (g_fail < 20) { printf("  FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } ++g_fail; } } while (0)
static std::mt19937 g_rng(11);
static float urand(float s) { return std::uniform_real_distribution<float>(-s, s)(g_rng); }

struct Problem {
    int M, N, K;
    std::vector<__half> A, W;
    std::vector<float> c0, c1, bias, resid;
    std::vector<float2> stats;
    CUtensorMap ta, tb256, tb128;
};
static Problem make_problem(int M, int N, int K) {
    Problem p;
    p.M = M; p.N = N; p.K = K;
    p.A.resize(static_cast<size_t>(M) * K); p.W.resize(static_cast<size_t>(N) * K);
    for (auto &h : p.A) h = __float2half_rn(urand(1.f));
    for (auto &h : p.W) h = __float2half_rn(urand(0.2f));
    const int Np = (N + 255) / 256 * 256, Mp = (M + 255) / 256 * 256;
    p.c0.resize(Np); p.c1.resize(Np); p.bias.resize(Np); p.stats.resize(Mp); p.resid.resize(static_cast<size_t>(Mp) * N);
    for (auto &x : p.c0) x = urand(0.5f);
    for (auto &x : p.c1) x = urand(1.f);
    for (auto &x : p.bias) x = urand(0.5f);
    for (auto &x : p.resid) x = urand(1.f);
    for (auto &s : p.stats) s = make_float2(urand(0.3f), 0.5f + fabsf(urand(1.f)));
    p.ta = CUtensorMap{p.A.data(), 2, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(K) * 2, 128, 64};
    p.tb256 = CUtensorMap{p.W.data(), 2, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(K) * 2, 256, 64};
    p.tb128 = CUtensorMap{p.W.data(), 2, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(K) * 2, 128, 64};
    return p;
}
static long long differing(const std::vector<uint8_t> &a, const std::vector<uint8_t> &b) {
    long long d = 0;
    for (size_t i = 0; i < a.size(); ++i) d += a[i] != b[i];
    return d;
}

// fp16-output epilogues (GELU, optionally deferred): 1-CTA kernel vs pair kernel with 8 / 16 epilogue warps and the relay variant
template <bool DEFER>
static void test_half_out(int M, int N, int K, int clusters) {
    Problem p = make_problem(M, N, K);
    using E128 = EpiLinear<1, true, false, DEFER, 128>;
    using E64 = EpiLinear<1, true, false, DEFER, 64>;
    const size_t ybytes = static_cast<size_t>(M + 8) * N * 2;
    auto run = [&](auto kernel_launch) { std::vector<uint8_t> y(ybytes, 0x7b); kernel_launch(reinterpret_cast<__half *>(y.data())); return y; };
    auto e128 = [&](__half *Y) { return E128{p.c0.data(), nullptr, Y, M, N, N, 0, nullptr, 0, 0, 0, 0, p.c1.data(), p.stats.data()}; };
    auto e64 = [&](__half *Y) { return E64{p.c0.data(), nullptr, Y, M, N, N, 0, nullptr, 0, 0, 0, 0, p.c1.data(), p.stats.data()}; };
    const auto y1 = run([&](__half *Y) { auto e = e128(Y); shim::launch(dim3(3), dim3(GEMM_THREADS), [&] { gemm_tc_kernel<E128, false, GEMM_KIND_F16>(p.ta, p.tb256, M, N, K, e); }); });
    const auto y8 = run([&](__half *Y) { auto e = e128(Y); shim::launch_cluster(dim3(2 * clusters), dim3(64 + 32 * 8), 2, [&] { gemm_tc2_kernel<E128, false, GEMM_KIND_F16, 6, false, 8>(p.ta, p.tb128, M, N, K, e); }); });
    const auto y8r = run([&](__half *Y) { auto e = e128(Y); shim::launch_cluster(dim3(2 * clusters), dim3(64 + 32 * 8), 2, [&] { gemm_tc2_kernel<E128, false, GEMM_KIND_F16, 6, true, 8>(p.ta, p.tb128, M, N, K, e); }); });
    const auto y16 = run([&](__half *Y) { auto e = e64(Y); shim::launch_cluster(dim3(2 * clusters), dim3(64 + 32 * 16), 2, [&] { gemm_tc2_kernel<E64, false, GEMM_KIND_F16, 5, false, 16>(p.ta, p.tb128, M, N, K, e); }); });
    CHECK(differing(y1, y8) == 0, "pair kernel (8 epilogue warps) differs from the 1-CTA kernel in %lld bytes", differing(y1, y8));
    CHECK(differing(y1, y8r) == 0, "relay pair kernel differs from the 1-CTA kernel in %lld bytes", differing(y1, y8r));
    CHECK(differing(y1, y16) == 0, "pair kernel with 16 epilogue warps / 5 stages differs from the 1-CTA kernel in %lld bytes", differing(y1, y16));
    // and the 1-CTA result is the intended function
    const __half *Y = reinterpret_cast<const __half *>(y1.data());
    double worst = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double acc = 0;
            for (int k = 0; k < K; ++k) acc += static_cast<double>(__half2float(p.A[static_cast<size_t>(m) * K + k])) * __half2float(p.W[static_cast<size_t>(n) * K + k]);
            const float pre = DEFER ? static_cast<float>(p.stats[m].y * (acc - p.stats[m].x * p.c1[n]) + p.c0[n]) : static_cast<float>(acc + p.c0[n]);
            worst = std::max(worst, fabs(static_cast<double>(__half2float(Y[static_cast<size_t>(m) * N + n])) - gelu_erf(pre)));
        }
    CHECK(worst < 5e-3, "GELU epilogue deviates by %g", worst);
    for (size_t i = static_cast<size_t>(M) * N * 2; i < ybytes; ++i) CHECK(y16[i] == 0x7b, "rows beyond M written");
    printf("GELU%s M=%d N=%d K=%d, %d clusters: 1-CTA == pair(8 warps) == relay == pair(16 warps, 5 stages); deviation from fp64 %.2e: %s\n",
           DEFER ? " deferred" : "", M, N, K, clusters, worst, g_fail ? "FAIL" : "ok");
}

// fp32 residual epilogues through the pair kernel: plain (EpiLinear MODE 2) and deferred (EpiResidDefer, in place)
static void test_resid(int M, int H, int K, int clusters) {
    Problem p = make_problem(M, H, K);
    const int nparts = H / 128;
    using ER = EpiLinear<2, false, false>;
    std::vector<float> Y1(static_cast<size_t>(M) * H, -3.f), Y2(static_cast<size_t>(M) * H, -3.f);
    ER e1{p.bias.data(), p.resid.data(), Y1.data(), M, H, H, 0, nullptr, 0, 0, 0, 0};
    ER e2{p.bias.data(), p.resid.data(), Y2.data(), M, H, H, 0, nullptr, 0, 0, 0, 0};
    shim::launch(dim3(2), dim3(GEMM_THREADS), [&] { gemm_tc_kernel<ER, false, GEMM_KIND_F16>(p.ta, p.tb256, M, H, K, e1); });
    shim::launch_cluster(dim3(2 * clusters), dim3(GEMM_THREADS), 2, [&] { gemm_tc2_kernel<ER, false, GEMM_KIND_F16, 6, false, 8>(p.ta, p.tb128, M, H, K, e2); });
    CHECK(memcmp(Y1.data(), Y2.data(), Y1.size() * 4) == 0, "residual epilogue: pair kernel differs from the 1-CTA kernel");
    // deferred: y in place, yh, parts -- 1-CTA vs pair
    std::vector<float> gamma(H), beta(H);
    for (auto &x : gamma) x = 1.f + urand(0.3f);
    for (auto &x : beta) x = urand(0.2f);
    std::vector<float> ya = p.resid, yb = p.resid;
    ya.resize(static_cast<size_t>(M) * H); yb.resize(static_cast<size_t>(M) * H);
    std::vector<__half> ha(ya.size()), hb(yb.size());
    std::vector<float2> pa(static_cast<size_t>(nparts) * (M + 8), make_float2(-7.f, -7.f)), pb = pa;
    EpiResidDefer da{p.bias.data(), ya.data(), ha.data(), p.stats.data(), gamma.data(), beta.data(), pa.data(), static_cast<int64_t>(M + 8), M, H, H};
    EpiResidDefer db{p.bias.data(), yb.data(), hb.data(), p.stats.data(), gamma.data(), beta.data(), pb.data(), static_cast<int64_t>(M + 8), M, H, H};
    shim::launch(dim3(2), dim3(GEMM_THREADS), [&] { gemm_tc_kernel<EpiResidDefer, false, GEMM_KIND_F16>(p.ta, p.tb256, M, H, K, da); });
    shim::launch_cluster(dim3(2 * clusters), dim3(GEMM_THREADS), 2, [&] { gemm_tc2_kernel<EpiResidDefer, false, GEMM_KIND_F16, 6, false, 8>(p.ta, p.tb128, M, H, K, db); });
    CHECK(memcmp(ya.data(), yb.data(), ya.size() * 4) == 0 && memcmp(ha.data(), hb.data(), ha.size() * 2) == 0 && memcmp(pa.data(), pb.data(), pa.size() * 8) == 0,
          "deferred residual epilogue: pair kernel differs from the 1-CTA kernel");
    double worst = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < H; ++n) {
            double acc = 0;
            for (int k = 0; k < K; ++k) acc += static_cast<double>(__half2float(p.A[static_cast<size_t>(m) * K + k])) * __half2float(p.W[static_cast<size_t>(n) * K + k]);
            const double want = acc + p.bias[n] + ((p.resid[static_cast<size_t>(m) * H + n] - p.stats[m].x) * p.stats[m].y * gamma[n] + beta[n]);
            worst = std::max(worst, fabs(ya[static_cast<size_t>(m) * H + n] - want));
        }
    CHECK(worst < 1e-4, "deferred residual epilogue deviates by %g", worst);
    printf("residual epilogues M=%d H=%d K=%d: pair == 1-CTA (plain and deferred, in place); deferred deviation from fp64 %.2e: %s\n", M, H, K, worst, g_fail ? "FAIL" : "ok");
}

int main() {
    test_half_out<true>(300, 392, 128, 1);      // ragged M and N: the peer CTA's rows / B half partly and fully out of range
    test_half_out<false>(520, 768, 192, 2);     // several tiles per cluster, both accumulator buffers, the stage ring wraps
    test_half_out<true>(129, 256, 64, 3);       // more clusters than tiles
    test_resid(300, 384, 128, 2);
    printf("pair_emul: %s (%d failed checks)\n", g_fail ? "FAIL" : "ALL OK", g_fail);
    return g_fail ? 1 : 0;
}
