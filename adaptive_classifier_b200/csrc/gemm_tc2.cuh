// gemm_tc2.cuh -- CTA-pair (tcgen05 cta_group::2) variant of the persistent GEMM mainloop of gemm_tc.cuh.
//
//   D[M,N] (fp32, TMEM) = A[M,K] * B[N,K]^T, same operand kinds, same epilogue functor concept as gemm_tc.cuh.
//
// Why a pair: one 128 x 256 x 16 tcgen05.mma reads 4 KB of A and 8 KB of B from shared memory; at the fp16 tensor rate
// that is ~90 B/clk of the SM's 128 B/clk shared-memory bandwidth, and the 1-CTA mainloop measured 72 % tensor-pipe
// active on the QKV / FFN2 shapes (profiles/r01_gemm_v3_fp16_ncu.txt).  With cta_group::2 the two SMs of a TPC compute one
// 256 x 256 tile: each CTA stages ITS 128 rows of A and ITS 128 rows (N half) of B, i.e. 8 KB per MMA per SM, and each
// B byte is fetched from L2 once per pair instead of once per CTA.  32 KB per stage per CTA -> 6 stages in flight.
//
//   cluster = 2 CTAs (rank 0 = leader).  Per CTA:
//   warp 0      : TMA producer for this CTA's A rows / B rows; completion bytes are counted on the LEADER's full barrier
//                 (cp.async.bulk.tensor .cta_group::2 with a mapa'd barrier address)
//   warp 1      : TMEM allocator (both CTAs, cta_group::2); in the leader, one thread issues tcgen05.mma.cta_group::2
//                 (M 256) and commits with .multicast::cluster so "stage free" / "accumulator full" arrive in both CTAs
//   warps 2..9  : epilogue of this CTA's 128 accumulator rows (identical to gemm_tc.cuh); "accumulator drained" arrives on
//                 the leader's barrier (remote arrive from the peer)
//
// Status: every encoder projection runs through this kernel.  Measured on a B200 (profiles/r02_variants.md): bit-identical to
// the 1-CTA kernel on all encoder shapes; 12-layer forward at B*S = 65536 with deferred LayerNorm and the 16-warp GELU epilogue
// 15.02 -> 14.31 ms per step against the 1-CTA mainloop.  The kNN coarse pass stays on gemm_tc.cuh (no gain there).
#pragma once
#include "gemm_tc.cuh"

namespace ac {

constexpr int GEMM2_PAIR_M = 2 * GEMM_BLOCK_M;                          // 256 rows per cluster tile
constexpr int GEMM2_B_ROWS = GEMM_BLOCK_N / 2;                          // B rows staged per CTA
constexpr int GEMM2_A_STAGE_BYTES = GEMM_BLOCK_M * 128;                 // 16 KB (128 rows x one 128-byte swizzle row)
constexpr int GEMM2_B_STAGE_BYTES = GEMM2_B_ROWS * 128;                 // 16 KB
constexpr int GEMM2_STAGE_BYTES = GEMM2_A_STAGE_BYTES + GEMM2_B_STAGE_BYTES;
constexpr int GEMM2_STAGES = 6;
__host__ __device__ constexpr int gemm2_smem_bytes(int stages, int epi_warps = GEMM_EPI_WARPS) {
    return stages * GEMM2_STAGE_BYTES + epi_warps * GEMM_EPI_STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
}
// 16 epilogue warps (4 per scheduler) for issue-bound epilogues such as bias + GELU: their staging tiles cost one stage
__host__ __device__ constexpr int gemm2_stages_for(int epi_warps) { return epi_warps > 8 ? GEMM2_STAGES - 1 : GEMM2_STAGES; }
static_assert(gemm2_smem_bytes(GEMM2_STAGES) <= 227 * 1024, "stage ring does not fit");
static_assert(gemm2_smem_bytes(gemm2_stages_for(16), 16) <= 227 * 1024, "stage ring does not fit with 16 epilogue warps");
static_assert(GEMM2_A_STAGE_BYTES == GEMM_A_STAGE_BYTES, "A stage layout is shared with the 1-CTA kernel");

// wait sites, reported by the watchdog so that a broken protocol names the barrier that never completed
enum { PAIR_SITE_PRODUCER_EMPTY = 0, PAIR_SITE_MMA_TMEM_EMPTY = 1, PAIR_SITE_MMA_FULL = 2, PAIR_SITE_EPI_TMEM_FULL = 3 };
__device__ __forceinline__ void mbar_wait_guarded_cluster(uint64_t *bar, uint32_t parity, int site, int index) {
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (++spins > (1u << 24)) {
            printf("ac: mbarrier watchdog fired (pair kernel, block %d rank %u thread %d, site %d, index %d, parity %u)\n", blockIdx.x,
                   cluster_ctarank(), threadIdx.x, site, index, parity);
            __trap();
        }
    }
}

// Both CTAs' TMA loads count their bytes on the leader's full barrier (.cta_group::2 loads with a mapa'd barrier address).
// kEpiWarps = 8 (warp w drains lane quarter w%4, column half (w-2)/4) or 16 (column quarter (w-2)/4, 64 columns = 2 chunks):
//                 the bias + GELU epilogue issues ~17 instructions per element; two warps per scheduler cannot hide its MUFU /
//                 dependency latency behind a K = 768 mainloop (FFN1 measured 56 % tensor-pipe active), four can.
template <class Epi, bool kMFastest = false, int kKind = GEMM_KIND_TF32, int kStages = GEMM2_STAGES, int kEpiWarps = GEMM_EPI_WARPS>
// registers: warps are allocated in groups of 4, so the 10 (18) warps of a CTA cost 12 (20) warps of registers: 168 (96) per thread.
// (__maxnreg__(200) compiled and then failed to launch: 12 x 32 x 200 > 65536.)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * kEpiWarps, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                int M, int N, int K, Epi epi) {
    extern __shared__ uint8_t smem_raw[];
    // identical carve-up in both CTAs: the MMA addresses the peer's operands at the SAME shared-memory offsets
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem + kStages * GEMM2_A_STAGE_BYTES;
    uint8_t *epi_stage = smem + kStages * GEMM2_STAGE_BYTES;
    static_assert(kEpiWarps == 8 || kEpiWarps == 16, "epilogue warps come in multiples of the 4 TMEM lane quarters");
    constexpr int kColsPerWarp = GEMM_BLOCK_N / (kEpiWarps / 4);     // 128 or 64 accumulator columns per epilogue warp
    uint64_t *bars = reinterpret_cast<uint64_t *>(epi_stage + kEpiWarps * GEMM_EPI_STAGE_BYTES);
    uint64_t *full_bar = bars;                      // [kStages]  used in the leader only (both CTAs' bytes land here)
    uint64_t *empty_bar = bars + kStages;           // [kStages]  one per CTA, multicast commit
    uint64_t *tmem_full = bars + 2 * kStages;       // [2]        one per CTA, multicast commit
    uint64_t *tmem_empty = bars + 2 * kStages + 2;  // [2]        leader only: 2 x GEMM_EPI_WARPS arrivals
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();         // 0 = leader
    const int cluster = static_cast<int>(blockIdx.x >> 1);
    const int num_clusters = static_cast<int>(gridDim.x >> 1);
    const int tiles_m = (M + GEMM2_PAIR_M - 1) / GEMM2_PAIR_M;
    const int tiles_n = (N + GEMM_BLOCK_N - 1) / GEMM_BLOCK_N;
    const int num_tiles = tiles_m * tiles_n;
    constexpr int BK = gemm_block_k(kKind);
    const int num_kb = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full[0], 1);
        mbar_init(&tmem_full[1], 1);
        mbar_init(&tmem_empty[0], 2 * kEpiWarps);
        mbar_init(&tmem_empty[1], 2 * kEpiWarps);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc_pair(tmem_slot, GEMM_TMEM_COLS);   // one warp of EACH CTA of the pair
        tmem_relinquish_pair();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();       // barriers of both CTAs are initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer (both CTAs) ----------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster; tile < num_tiles; tile += num_clusters) {
                const int m0 = (kMFastest ? tile % tiles_m : tile / tiles_n) * GEMM2_PAIR_M + static_cast<int>(rank) * GEMM_BLOCK_M;
                const int n0 = (kMFastest ? tile / tiles_m : tile % tiles_n) * GEMM_BLOCK_N + static_cast<int>(rank) * GEMM2_B_ROWS;
                for (int kb = 0; kb < num_kb; ++kb) {
                    // arrival = the leader's multicast commit
                    mbar_wait_guarded_cluster(&empty_bar[stage], phase ^ 1, PAIR_SITE_PRODUCER_EMPTY, stage);
                    const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[stage]), 0);
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * GEMM2_STAGE_BYTES);
                    tma_load_2d_pair(smem_a + stage * GEMM2_A_STAGE_BYTES, &tmap_a, full_leader, kb * BK, m0);
                    tma_load_2d_pair(smem_b + stage * GEMM2_B_STAGE_BYTES, &tmap_b, full_leader, kb * BK, n0);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (one thread of the leader CTA) ----------------
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc = umma_idesc(kKind == GEMM_KIND_F16 ? 0u : 2u, GEMM2_PAIR_M, GEMM_BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = cluster; tile < num_tiles; tile += num_clusters) {
                // both CTAs' epilogues drained this buffer
                mbar_wait_guarded_cluster(&tmem_empty[acc], acc_phase ^ 1, PAIR_SITE_MMA_TMEM_EMPTY, acc);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * GEMM_BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    // both CTAs' operand bytes have landed
                    mbar_wait_guarded_cluster(&full_bar[stage], phase, PAIR_SITE_MMA_FULL, stage);
                    tc_fence_after();
                    const uint64_t a_desc = umma_desc_sw128(smem_u32(smem_a + stage * GEMM2_A_STAGE_BYTES));
                    const uint64_t b_desc = umma_desc_sw128(smem_u32(smem_b + stage * GEMM2_B_STAGE_BYTES));
#pragma unroll
                    for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
                        if (kKind == GEMM_KIND_F16) umma_f16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                        else umma_tf32_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    }
                    tc_commit_pair(&empty_bar[stage], 0x3);    // stage free in both CTAs once these MMAs retire
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                tc_commit_pair(&tmem_full[acc], 0x3);          // accumulator complete -> both epilogues
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ---------------- epilogue warps (both CTAs, this CTA's 128 rows) ----------------
        const int q = warp & 3;
        const int cpart = (warp - 2) >> 2;            // which kColsPerWarp-column slice of the accumulator this warp drains
        typename Epi::State est;
        epi.begin_cta(est, q, lane);
        int acc = 0;
        uint32_t acc_phase = 0;
        int it = 0;
        for (int tile = cluster; tile < num_tiles; tile += num_clusters, ++it) {
            GemmTileInfo ti;
            ti.m0 = (kMFastest ? tile % tiles_m : tile / tiles_n) * GEMM2_PAIR_M + static_cast<int>(rank) * GEMM_BLOCK_M;
            ti.n0 = (kMFastest ? tile / tiles_m : tile % tiles_n) * GEMM_BLOCK_N;
            ti.tile_iter = it;
            const int row = ti.m0 + q * 32 + lane;
            const int c_lo = cpart * kColsPerWarp;
            // operands the epilogue needs from global memory (residual rows) are requested kDist chunks ahead into kDist + 1
            // register buffers; the first requests go out before the accumulator is even complete
            constexpr int kDist = Epi::kPrefetchDist, kBufs = kDist + 1, kChunks = kColsPerWarp / 32;
#pragma unroll
            for (int d = 0; d < kDist; ++d)
                if (d < kChunks) epi.prefetch(est, ti, row, ti.n0 + c_lo + 32 * d, lane, d % kBufs);
            mbar_wait_guarded_cluster(&tmem_full[acc], acc_phase, PAIR_SITE_EPI_TMEM_FULL, acc);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * GEMM_BLOCK_N;
#pragma unroll (Epi::kUnrollChunks)
            for (int ci = 0; ci < kChunks; ++ci) {
                const int c = c_lo + 32 * ci;
                if (ci + kDist < kChunks) epi.prefetch(est, ti, row, ti.n0 + c + 32 * kDist, lane, (ci + kDist) % kBufs);
                uint32_t r[32];
                tmem_ld_32x32(taddr + c, r);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                epi.tile(est, ti, row, ti.n0 + c, v, epi_stage + (warp - 2) * GEMM_EPI_STAGE_BYTES, lane, ci % kBufs, taddr + c);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        epi.end_cta(est, q, lane);
    }

    // teardown: neither CTA may free TMEM or exit while the pair's MMAs / remote arrivals can still touch it
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, GEMM_TMEM_COLS);
    }
}

// host-side launcher.  tb must be a tensor map over B with a 128-row box (GEMM2_B_ROWS), ta the usual 128-row A box.
template <class Epi, bool kMFastest = false, int kKind = GEMM_KIND_TF32, int kEpiWarps = GEMM_EPI_WARPS>
int launch_gemm_tc2(const CUtensorMap &ta, const CUtensorMap &tb, int M, int N, int K, const Epi &epi,
                    cudaStream_t stream, int max_ctas = 0, int prof_cls = PROF_GEMM_LINEAR, double prof_bytes = 0.0) {
    constexpr int kStg = gemm2_stages_for(kEpiWarps);
    constexpr int smem = gemm2_smem_bytes(kStg, kEpiWarps);
    auto kern = gemm_tc2_kernel<Epi, kMFastest, kKind, kStg, kEpiWarps>;
    static bool attr_set[64] = {};   // per instantiation and per device (the attribute belongs to the (function, device) pair)
    int dev = 0;
    AC_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        AC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const int tiles = ((M + GEMM2_PAIR_M - 1) / GEMM2_PAIR_M) * ((N + GEMM_BLOCK_N - 1) / GEMM_BLOCK_N);
    int clusters = sm_count() / 2;
    if (max_ctas > 0 && max_ctas / 2 < clusters) clusters = max_ctas / 2;
    if (tiles < clusters) clusters = tiles;
    if (clusters <= 0) return AC_OK;
    const int slot = prof_begin(prof_cls, 2.0 * M * static_cast<double>(N) * K, prof_bytes, stream);
    kern<<<2 * clusters, 64 + 32 * kEpiWarps, smem, stream>>>(ta, tb, M, N, K, epi);
    prof_end(slot, stream);
    AC_LAUNCH_CHECK();
    return AC_OK;
}

}  // namespace ac
