// knn_emul.cpp -- the coarse pass of the prototype scan (gemm_tc_kernel<EpiKnn / EpiKnnLane, m-fastest, kind::f16>) on the
// functional Blackwell model.  TEST INFRASTRUCTURE ONLY.
// The per-(query, CTA, column half) candidate lists are only specified below the exclusion bound (what gets rejected above it
// depends on when other lists publish their bounds), so the check is the property the pipeline relies on: per query, the best
// k (key, row) pairs over the union of its lists are exactly the best k of a brute-force scan computed with the same
// arithmetic (fp16 operands, the model's accumulation order, key = fma(-2, q.p, ||p||^2)).  EpiKnn (verified on a B200) must
// pass -- that checks the harness -- and so must EpiKnnLane, the per-lane slow path written without GPU access.
#include "tc_emul.h"
#include "../../include/adaptive_b200.h"

static void __threadfence_system() {}
static unsigned int atomicAdd(unsigned int *p, unsigned int v) { unsigned int o = *p; *p += v; return o; }
static int atomicAdd(int *p, int v) { int o = *p; *p += v; return o; }
static uint32_t atomicMin(uint32_t *p, uint32_t v) { uint32_t o = *p; if (v < o) *p = v; return o; }
static inline int __ffs(uint32_t x) { return x ? __builtin_ctz(x) + 1 : 0; }
static inline uint32_t shim_reduce_or(uint32_t v) { uint32_t o = 0; for (int l = 0; l < 32; ++l) o |= shim_shfl(v, l); return o; }
#define __reduce_or_sync(mask, v) shim_reduce_or(v)
#define __any_sync(mask, pred) (shim_ballot(pred) != 0)
namespace ac {
static inline void griddep_wait() {}
static inline void griddep_launch_dependents() {}
int sm_count() { return 4; }
size_t topk_select_workspace(int, int64_t, int) { return 0; }
}  // namespace ac

#include "_gen_common_tc.inc"
#include "_gen_gemm_tc_tc.inc"
#include "_gen_gemm_tc2_tc.inc"
#include "_gen_knn_tc_tc.inc"

using namespace ac;

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (g_fail < 20) { printf("  FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } ++g_fail; } } while (0)

template <class Epi>
static void run_scan(const char *name, int B, int N, int D, int ctas_per_tile, int k, bool pair = false) {
    std::mt19937 rng(5);
    std::normal_distribution<float> nd(0.f, 1.f);
    const int tiles_m = (B + 127) / 128, Bp = tiles_m * 128;
    std::vector<float> Pf(static_cast<size_t>(N) * D), Qf(static_cast<size_t>(Bp) * D, 0.f), pn(N);
    for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int d = 0; d < D; ++d) { Pf[static_cast<size_t>(n) * D + d] = nd(rng); s += Pf[static_cast<size_t>(n) * D + d] * Pf[static_cast<size_t>(n) * D + d]; }
        for (int d = 0; d < D; ++d) Pf[static_cast<size_t>(n) * D + d] /= static_cast<float>(sqrt(s));
    }
    for (int b = 0; b < B; ++b) {                      // queries near some prototype: distinct neighbour structure per query
        const int src = (b * 37) % N;
        double s = 0;
        for (int d = 0; d < D; ++d) { Qf[static_cast<size_t>(b) * D + d] = Pf[static_cast<size_t>(src) * D + d] + 0.15f * nd(rng); s += Qf[static_cast<size_t>(b) * D + d] * Qf[static_cast<size_t>(b) * D + d]; }
        for (int d = 0; d < D; ++d) Qf[static_cast<size_t>(b) * D + d] /= static_cast<float>(sqrt(s));
    }
    std::vector<__half> Ph(Pf.size()), Qh(Qf.size());
    for (size_t i = 0; i < Pf.size(); ++i) Ph[i] = __float2half_rn(Pf[i]);
    for (size_t i = 0; i < Qf.size(); ++i) Qh[i] = __float2half_rn(Qf[i]);
    for (int n = 0; n < N; ++n) { float s = 0.f; for (int d = 0; d < D; ++d) s = fmaf(Pf[static_cast<size_t>(n) * D + d], Pf[static_cast<size_t>(n) * D + d], s); pn[n] = s; }

    const int slots = 2 * ctas_per_tile, grid = ctas_per_tile * tiles_m;
    int kt = k + 3 > 8 ? k + 3 : 8;
    if (kt > KNN_KC) kt = KNN_KC;
    std::vector<float> ckey(static_cast<size_t>(B) * slots * KNN_KC, -1.f);
    std::vector<int32_t> cidx(static_cast<size_t>(B) * slots * KNN_KC, -7);
    std::vector<uint32_t> gthr(Bp, 0xFFFFFFFFu);
    CUtensorMap ta{Qh.data(), 2, static_cast<uint64_t>(Bp), static_cast<uint64_t>(D), static_cast<uint64_t>(D) * 2, 128, 64};
    CUtensorMap tb{Ph.data(), 2, static_cast<uint64_t>(N), static_cast<uint64_t>(D), static_cast<uint64_t>(D) * 2, 256, 64};
    const EpiKnn base{pn.data(), ckey.data(), cidx.data(), gthr.data(), B, static_cast<int64_t>(N), tiles_m, slots, kt, pair ? 1 : 0};
    Epi epi{base};
    if (pair) {      // a cluster owns two consecutive query tiles; B rows split between the two CTAs (128-row boxes)
        CUtensorMap tb2 = tb;
        tb2.box_rows = 128;
        shim::launch_cluster(dim3(grid), dim3(GEMM_THREADS), 2, [&] { gemm_tc2_kernel<Epi, true, GEMM_KIND_F16, 6, false, 8>(ta, tb2, Bp, N, D, epi); });
    } else {
        shim::launch(dim3(grid), dim3(GEMM_THREADS), [&] { gemm_tc_kernel<Epi, true, GEMM_KIND_F16>(ta, tb, Bp, N, D, epi); });
    }

    long long inserted = 0;
    for (int b = 0; b < B; ++b) {
        // brute force with the model's arithmetic: per MMA instruction 16 sequential fmas, instructions accumulate in order
        std::vector<std::pair<float, int>> all(N);
        for (int n = 0; n < N; ++n) {
            float acc = 0.f;
            for (int d = 0; d < D; ++d) acc = fmaf(__half2float(Qh[static_cast<size_t>(b) * D + d]), __half2float(Ph[static_cast<size_t>(n) * D + d]), acc);
            all[n] = {fmaf(-2.f, acc, pn[n]), n};
        }
        std::sort(all.begin(), all.end());
        std::vector<std::pair<float, int>> got;
        for (int s = 0; s < slots; ++s)
            for (int i = 0; i < KNN_KC; ++i) {
                const size_t at = (static_cast<size_t>(b) * slots + s) * KNN_KC + i;
                if (cidx[at] >= 0) { got.push_back({ckey[at], cidx[at]}); ++inserted; }
                else CHECK(cidx[at] == -1, "%s: list entry (%d,%d,%d) was never written", name, b, s, i);
                if (i > 0) CHECK(!(ckey[at] < ckey[at - 1]), "%s: list (%d,%d) is not sorted", name, b, s);
            }
        std::sort(got.begin(), got.end());
        for (size_t i = 1; i < got.size(); ++i) CHECK(got[i].second != got[i - 1].second, "%s: row %d appears twice for query %d", name, got[i].second, b);
        for (int j = 0; j < k; ++j)
            CHECK(j < static_cast<int>(got.size()) && got[j] == all[j], "%s: query %d neighbour %d: lists give (%g, %d), scan gives (%g, %d)", name, b, j,
                  j < static_cast<int>(got.size()) ? got[j].first : 0.f, j < static_cast<int>(got.size()) ? got[j].second : -1, all[j].first, all[j].second);
    }
    printf("%s B=%d N=%d D=%d, %d CTAs per query tile: top-%d of every query's lists == brute-force scan (%lld list entries): %s\n", name, B, N, D,
           ctas_per_tile, k, inserted, g_fail ? "FAIL" : "ok");
}

int main() {
    run_scan<EpiKnn>("EpiKnn    ", 200, 3000, 128, 2, 5);
    run_scan<EpiKnnLane>("EpiKnnLane", 200, 3000, 128, 2, 5);
    run_scan<EpiKnnLane>("EpiKnnLane", 70, 1000, 64, 1, 10);       // ragged query tile, ragged last prototype tile, kt = 13
    run_scan<EpiKnn>("EpiKnn     pair", 200, 3000, 128, 2, 5, true);     // pair kernel (ran on the B200 with EpiKnn)
    run_scan<EpiKnnLane>("EpiKnnLane pair", 200, 3000, 128, 2, 5, true);
    printf("knn_emul: %s (%d failed checks)\n", g_fail ? "FAIL" : "ALL OK", g_fail);
    return g_fail ? 1 : 0;
}
