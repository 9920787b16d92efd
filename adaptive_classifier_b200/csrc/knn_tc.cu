// knn_tc.cu -- stage K, tensor path: brute-force squared-L2 kNN over a large prototype matrix.
//
//   coarse pass   d~(q,p) = ||q||^2 + ||p||^2 - 2 q.p   with q.p on tcgen05 (kind::tf32) through the GEMM
//                 mainloop of gemm_tc.cuh: M = queries (128 per CTA, fixed per CTA), N = prototype rows
//                 streamed ONCE from HBM by TMA (4.N.D algorithmic bytes), fp32 accumulators in TMEM.
//                 Epilogue: thread = query row; running top-16 (coarse key, row id) in registers.
//   merge         per query, the per-CTA lists are sorted by (d~, id); the best KP = 32 become candidates.
//   re-rank       exact fp32 distances of the candidates in the oracle's lane order (knn_exact.cu).
//   certify       T = smallest coarse distance any non-candidate row can have, eps = rigorous bound on
//                 |d~ - d| for tf32 operands (q rounded RNE: 2^-11, p truncated by the MMA: 2^-10);
//                 a query is certified when d_exact[k-1] < T - eps: then no excluded row can enter or tie
//                 the top-k, so indices are identical to the exact scan.  Uncertified queries are recomputed
//                 by the exact scan (knn_exact.cu).  Result: bit-identical (d, id) to the exact path.
//
// Replaces faiss.IndexFlatL2.search (/root/reference/src/adaptive_classifier/memory.py:110-114) for the
// batched, large-N configuration of BASELINE.json (configs[1], configs[2]).
#include "gemm_tc2.cuh"
#include <cuda_fp16.h>
#include <math_constants.h>

namespace ac {

int launch_knn_rerank(const float *Q, const float *P, int B, int64_t N, int D, int kc, const int32_t *cand, float *out_d,
                      int64_t *out_i, int64_t row_offset, cudaStream_t stream);
size_t topk_select_workspace(int B, int64_t L, int k);
int topk_select(const float *d, const int64_t *idx, int B, int64_t L, int64_t in_stride, int64_t id_offset, int k,
                float *out_d, int64_t *out_i, void *ws, size_t ws_bytes, cudaStream_t stream);
int knn_exact_subset(const float *Q, const float *P, int B, int64_t N, int D, int k, float *out_d, int64_t *out_i,
                     int64_t row_offset, void *ws, size_t ws_bytes, cudaStream_t s);
size_t knn_exact_workspace_pub(int B, int64_t N, int k);

constexpr int KNN_KC = 16;   // candidates kept per (query, CTA)
constexpr int KNN_KP = 32;   // candidates re-ranked per query

// ------------------------------------------------------------------------------------------------
// coarse-pass epilogue
// ------------------------------------------------------------------------------------------------
// order-preserving map float -> uint32 (and back) so that atomicMin works on signed keys
__device__ __forceinline__ uint32_t key_to_ord(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_to_key(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

struct EpiKnn {
    const float *p_sqnorm;   // [N]
    float *cand_key;         // [B, slots, KC]  coarse key = ||p||^2 - 2 q.p  (||q||^2 added later)
    int32_t *cand_idx;       // [B, slots, KC]  local row id, -1 = empty
    uint32_t *gthr;          // [Bp] per query: smallest "worst kept key" published by any list so far (ordered uint)
    int B;                   // queries
    int64_t N;               // rows
    int tiles_m, slots;      // grid = slots/2 * tiles_m CTAs; every CTA owns one query tile and two lists per query
    int kt;                  // a list publishes its kt-th best key (k + 3 <= kt <= KC): see prefetch()
    int pair;                // != 0: launched as CTA pairs (gemm_tc2.cuh): a cluster owns two consecutive query tiles

    static constexpr int kUnrollChunks = 1;
    struct State {
        float key[KNN_KC];
        int32_t idx[KNN_KC];
        float pn[2];          // ||p||^2 of the 32 rows of a chunk, one per lane, requested one chunk ahead
        float gt;             // global bound for this thread's query, refreshed once per tile
    };

    __device__ __forceinline__ void begin_cta(State &st, int, int) const {
#pragma unroll
        for (int i = 0; i < KNN_KC; ++i) { st.key[i] = CUDART_INF_F; st.idx[i] = -1; }
        st.pn[0] = st.pn[1] = CUDART_INF_F;
        st.gt = CUDART_INF_F;
    }

    // Every list (74 per query at B = 512) would on its own perform ~KC ln(n/KC) sorted inserts; sharing a bound
    // across lists makes all of them reject what cannot matter any more.  Each list publishes its kt-th best key;
    // gt = min over lists.  Exclusion bound for the certification: T = min over lists of their FINAL kt-th best
    // (<= every published value, gt only decreases).  A row with key < T is never rejected (key < T <= gt(t) and
    // key < T <= list's kt-th best <= list's KC-th best) and never evicted (eviction would put KC better rows in its
    // list, i.e. that list's kt-th best < key, contradicting key < T).  T >= the global kt-th smallest key, so the
    // certification is at least as strong as with one global top-kt list.
    __device__ __forceinline__ void prefetch(State &st, const GemmTileInfo &, int row, int col0, int lane, int buf) const {
        const int64_t n = static_cast<int64_t>(col0) + lane;
        const float x = (n < N) ? __ldg(p_sqnorm + n) : CUDART_INF_F;
        if (buf) st.pn[1] = x; else st.pn[0] = x;
        if (buf == 0) {        // first chunk of a tile: publish this list's bound, pick up the others'
            float pub = CUDART_INF_F;
#pragma unroll
            for (int i = 0; i < KNN_KC; ++i) pub = (i == kt - 1) ? st.key[i] : pub;
            if (pub < CUDART_INF_F) atomicMin(gthr + row, key_to_ord(pub));
            st.gt = ord_to_key(*reinterpret_cast<volatile uint32_t *>(gthr + row));
        }
    }

    __device__ __forceinline__ void insert(State &st, float key, int32_t n) const {
        // sorted insertion (ascending); strict '<' keeps the earlier (lower id) row on ties.  Slots are visited from the
        // tail towards the head, so every read sees the pre-insertion value.
#pragma unroll
        for (int i = KNN_KC - 1; i > 0; --i) {
            const bool shift = key < st.key[i - 1];
            const bool here = !shift && (key < st.key[i]);
            const float nk = shift ? st.key[i - 1] : (here ? key : st.key[i]);
            const int32_t ni = shift ? st.idx[i - 1] : (here ? n : st.idx[i]);
            st.key[i] = nk;
            st.idx[i] = ni;
        }
        if (key < st.key[0]) { st.key[0] = key; st.idx[0] = n; }
    }

    __device__ __forceinline__ void tile(State &st, const GemmTileInfo &, int /*row*/, int col0, const float (&v)[32],
                                         uint8_t * /*stage*/, int /*lane*/, int buf, uint32_t taddr) const {
        // fast path: all lanes walk the same 32 prototype rows and only record which ones beat their query's bound
        const float pn_lane = buf ? st.pn[1] : st.pn[0];
        const float thr = fminf(st.key[KNN_KC - 1], st.gt);
        uint32_t hits = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float pn = __shfl_sync(0xffffffffu, pn_lane, j);
            const float key = fmaf(-2.f, v[j], pn);
            hits |= (key < thr) ? (1u << j) : 0u;
        }
        // slow path (rare once the bounds are tight): one copy of the insert network, the hit column is read again
        // from TMEM because v[] cannot be indexed dynamically
        uint32_t uni = __reduce_or_sync(0xffffffffu, hits);
        while (uni) {
            const int j = __ffs(uni) - 1;
            uni &= uni - 1;
            const uint32_t r = tmem_ld_32x1(taddr + j);
            tmem_ld_wait();
            const float key = fmaf(-2.f, __uint_as_float(r), __shfl_sync(0xffffffffu, pn_lane, j));
            const int64_t n = static_cast<int64_t>(col0) + j;
            if (((hits >> j) & 1u) && key < fminf(st.key[KNN_KC - 1], st.gt) && n < N) insert(st, key, static_cast<int32_t>(n));
        }
    }

    __device__ __forceinline__ void end_cta(State &st, int q, int lane) const {
        // two epilogue warps share a query row (one per 128-column half of every tile): each owns a slot
        const int chalf = ((threadIdx.x >> 5) - 2) >> 2;
        int mt, slot;
        if (pair) {
            const int cluster = blockIdx.x >> 1, tm2 = tiles_m >> 1;     // tiles_m is even in pair mode
            mt = (cluster % tm2) * 2 + static_cast<int>(cluster_ctarank());
            slot = (cluster / tm2) * 2 + chalf;
        } else {
            mt = blockIdx.x % tiles_m;
            slot = (blockIdx.x / tiles_m) * 2 + chalf;
        }
        const int row = mt * GEMM_BLOCK_M + q * 32 + lane;
        if (row >= B) return;
        float *ck = cand_key + (static_cast<int64_t>(row) * slots + slot) * KNN_KC;
        int32_t *ci = cand_idx + (static_cast<int64_t>(row) * slots + slot) * KNN_KC;
#pragma unroll
        for (int i = 0; i < KNN_KC; ++i) { ck[i] = st.key[i]; ci[i] = st.idx[i]; }
    }
};


// Per-lane slow path (opt-in, option "knn_epi"): same lists, cheaper to build.
//
// EpiKnn::tile walks the UNION of the 32 lanes' hit columns; every iteration re-reads one accumulator column from TMEM and
// runs the ~100-instruction insert network for the one or two lanes that hit there, the other lanes idle.  In the middle of
// the scan every lane still has about one insert per chunk, at a different column than its neighbours, so the union has 20-32
// members: the ncu capture of the scan shows 19 thread-instructions per accumulator element at 16 of 32 lanes active and a
// tensor pipe that is busy 48 % of the time (profiles/r01_knn_v2_ncu.txt).  Here every lane walks ITS OWN hits: the chunk's
// accumulator values go to the warp's private staging tile 16 columns at a time (lane-private rows, stride 17 floats), each
// lane pops its lowest hit column, reads its own value back and inserts; the loop runs max-over-lanes(hits) times instead of
// |union| times.  Every lane still inserts its hits in ascending column order, so the lists -- and everything downstream,
// including the certification argument -- are unchanged bit for bit.  Status: NOT yet run on hardware.
struct EpiKnnLane : EpiKnn {
    __device__ __forceinline__ void tile(State &st, const GemmTileInfo &, int /*row*/, int col0, const float (&v)[32],
                                         uint8_t *stage, int lane, int buf, uint32_t /*taddr*/) const {
        const float pn_lane = buf ? st.pn[1] : st.pn[0];
        const float thr = fminf(st.key[KNN_KC - 1], st.gt);
        uint32_t hits = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float pn = __shfl_sync(0xffffffffu, pn_lane, j);
            const float key = fmaf(-2.f, v[j], pn);
            hits |= (key < thr) ? (1u << j) : 0u;
        }
        if (__reduce_or_sync(0xffffffffu, hits) == 0) return;
        float *srow = reinterpret_cast<float *>(stage) + lane * 17;          // 32 x 17 floats = 2176 B of the 2560 B tile
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t hh = (hits >> (16 * half)) & 0xffffu;
            if (!__any_sync(0xffffffffu, hh != 0)) continue;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) srow[jj] = v[16 * half + jj];    // lane-private row: no cross-lane hazard
            while (__any_sync(0xffffffffu, hh != 0)) {
                const bool act = hh != 0;
                const int jj = act ? __ffs(hh) - 1 : 0;
                hh &= hh - 1;
                const float x = srow[jj];
                const float pn = __shfl_sync(0xffffffffu, pn_lane, 16 * half + jj);
                const float key = fmaf(-2.f, x, pn);
                const int64_t n = static_cast<int64_t>(col0) + 16 * half + jj;
                if (act && key < fminf(st.key[KNN_KC - 1], st.gt) && n < N) insert(st, key, static_cast<int32_t>(n));
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// small kernels around the coarse pass
// ------------------------------------------------------------------------------------------------
// Qr = tf32(RNE)(Q), qn[b] = ||Q_b||^2 (fp32), one warp per query
__global__ void knn_prep_queries_kernel(const float *__restrict__ Q, int B, int D, float *__restrict__ Qr,
                                        __half *__restrict__ Qh, float *__restrict__ qn) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B) return;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) {
        const float x = Q[static_cast<int64_t>(row) * D + i];
        if (Qh) Qh[static_cast<int64_t>(row) * D + i] = __float2half_rn(x);
        else Qr[static_cast<int64_t>(row) * D + i] = round_tf32(x);
        s = fmaf(x, x, s);
    }
    s = warp_sum(s);
    if (lane == 0) qn[row] = s;
}

// pn[n] = ||P_n||^2 and block maxima of it (for the error bound)
__global__ void knn_prep_rows_kernel(const float *__restrict__ P, int64_t N, int D, float *__restrict__ pn,
                                     float *__restrict__ block_max, int have_pn) {
    __shared__ float smax[8];
    const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    float s = 0.f;
    if (row < N) {
        if (have_pn) {
            s = pn[row];
        } else {
            const float *p = P + row * D;
            for (int i = lane; i < D; i += 32) s = fmaf(p[i], p[i], s);
            s = warp_sum(s);
            if (lane == 0) pn[row] = s;
        }
    }
    if (lane == 0) smax[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int i = 0; i < 8; ++i) m = fmaxf(m, smax[i]);
        block_max[blockIdx.x] = m;
    }
}
__global__ void knn_reduce_max_kernel(const float *__restrict__ v, int64_t n, float *__restrict__ out) {
    __shared__ float red[256];
    float m = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += 256) m = fmaxf(m, v[i]);
    red[threadIdx.x] = m;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

// widen the int32 candidate ids of the per-CTA lists for the (key, id) sort
__global__ void knn_widen_kernel(const int32_t *__restrict__ in, int64_t n, int64_t *__restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

// after the (key,id) sort of all per-CTA candidates: take the best KP ids for the re-rank and compute
//   T[b] = ||q||^2 + min( key of the first entry NOT re-ranked, min over CTAs of their worst kept key )
// sorted_*: [B, KP+1] ascending.  cand_key: [B, slots, KC] (entry kt-1 = the bound that list published).
__global__ void knn_pick_kernel(const float *__restrict__ sorted_key, const int64_t *__restrict__ sorted_idx,
                                const float *__restrict__ cand_key, const float *__restrict__ qn, int B, int slots, int kt,
                                int32_t *__restrict__ rerank_idx /*[B,KP]*/, float *__restrict__ T) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float *sk = sorted_key + static_cast<int64_t>(b) * (KNN_KP + 1);
    const int64_t *si = sorted_idx + static_cast<int64_t>(b) * (KNN_KP + 1);
    for (int j = 0; j < KNN_KP; ++j) rerank_idx[static_cast<int64_t>(b) * KNN_KP + j] = static_cast<int32_t>(si[j]);
    float t = (si[KNN_KP] >= 0) ? sk[KNN_KP] : CUDART_INF_F;
    for (int s = 0; s < slots; ++s) t = fminf(t, cand_key[(static_cast<int64_t>(b) * slots + s) * KNN_KC + (kt - 1)]);
    T[b] = t + qn[b];
}

// certified[b] = out_d[b,k-1] < T[b] - eps(b);  eps = 2*rel*||q||*max||p|| * 1.02 + 4e-5*(1+||q||^2+max||p||^2) with
//   rel = 2^-10 + 2^-11 + 2^-21 (fp32 rows truncated to tf32 by the MMA, queries rounded RNE), or
//   rel = 2^-11 + 2^-11 + 2^-22 (fp16 shadow rows and fp16 queries, both RNE; the subnormal tail adds < 1e-6)
__global__ void knn_certify_kernel(const float *__restrict__ out_d, const int64_t *__restrict__ out_i, const float *__restrict__ T,
                                   const float *__restrict__ qn, const float *__restrict__ pmax2, int B, int k, float rel,
                                   int32_t *__restrict__ fail_list, int32_t *__restrict__ fail_count) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float qn2 = qn[b], pm2 = pmax2[0];
    const float eps = 2.f * rel * 1.02f * sqrtf(qn2) * sqrtf(pm2) + 4e-5f * (1.f + qn2 + pm2);
    const float dk = out_d[static_cast<int64_t>(b) * k + (k - 1)];
    const bool full = out_i[static_cast<int64_t>(b) * k + (k - 1)] >= 0;
    // T == +inf: every row of the index was a candidate (nothing excluded) -> exact by construction
    const bool ok = (T[b] == CUDART_INF_F) || (full && dk < T[b] - eps);
    if (!ok) {
        const int slot = atomicAdd(fail_count, 1);
        fail_list[slot] = b;
    }
}

__global__ void knn_gather_rows_kernel(const float *__restrict__ src, const int32_t *__restrict__ list, int n, int D,
                                       float *__restrict__ dst) {
    const int r = blockIdx.x;
    if (r >= n) return;
    const float *s = src + static_cast<int64_t>(list[r]) * D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) dst[static_cast<int64_t>(r) * D + i] = s[i];
}
__global__ void knn_scatter_results_kernel(const float *__restrict__ d, const int64_t *__restrict__ idx,
                                           const int32_t *__restrict__ list, int n, int k, float *__restrict__ out_d,
                                           int64_t *__restrict__ out_i) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * k) return;
    const int r = t / k, j = t % k;
    out_d[static_cast<int64_t>(list[r]) * k + j] = d[t];
    out_i[static_cast<int64_t>(list[r]) * k + j] = idx[t];
}

// ------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------
struct KnnTcPlan {
    int tiles_m, slots, grid, grid_ctas;
    size_t off_qr, off_qn, off_gthr, off_pn, off_bmax, off_pmax, off_ckey, off_cidx, off_cidx64, off_skey, off_sidx, off_ridx, off_T,
        off_rd, off_ri, off_fail, off_fq, off_fd, off_fi, off_sel, sel_bytes, off_exact, exact_bytes, total;
};

static KnnTcPlan plan_knn_tc(int B, int64_t N, int D, int k) {
    KnnTcPlan p;
    p.tiles_m = (B + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
    const int sms = sm_count();
    p.slots = sms / p.tiles_m;
    if (p.slots < 1) p.slots = 1;
    const int64_t tiles_n = (N + GEMM_BLOCK_N - 1) / GEMM_BLOCK_N;
    if (p.slots > tiles_n) p.slots = static_cast<int>(tiles_n);
    p.grid = p.slots * p.tiles_m;
    p.grid_ctas = p.grid;
    p.slots *= 2;   // candidate lists per query: one per (CTA, accumulator column half)
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    const size_t Bp = static_cast<size_t>(p.tiles_m) * GEMM_BLOCK_M;
    p.off_qr = take(Bp * D * 4);
    p.off_qn = take(Bp * 4);
    p.off_gthr = take(Bp * 4);
    p.off_pn = take(static_cast<size_t>(N) * 4);
    p.off_bmax = take(static_cast<size_t>((N + 7) / 8) * 4);
    p.off_pmax = take(256);
    const size_t nc = static_cast<size_t>(B) * p.slots * KNN_KC;
    p.off_ckey = take(nc * 4);
    p.off_cidx = take(nc * 4);
    p.off_cidx64 = take(nc * 8);
    p.off_skey = take(static_cast<size_t>(B) * (KNN_KP + 1) * 4);
    p.off_sidx = take(static_cast<size_t>(B) * (KNN_KP + 1) * 8);
    p.off_ridx = take(static_cast<size_t>(B) * KNN_KP * 4);
    p.off_T = take(static_cast<size_t>(B) * 4);
    p.off_rd = take(static_cast<size_t>(B) * KNN_KP * 4);
    p.off_ri = take(static_cast<size_t>(B) * KNN_KP * 8);
    p.off_fail = take(static_cast<size_t>(B + 1) * 4);
    p.sel_bytes = topk_select_workspace(B, static_cast<int64_t>(p.slots) * KNN_KC, KNN_KP + 1) + 256;
    p.off_sel = take(p.sel_bytes);
    p.total = off;
    (void)k;
    return p;
}

size_t knn_tc_workspace(int B, int64_t N, int D, int k) {
    KnnTcPlan p = plan_knn_tc(B, N, D, k);
    // fallback staging for uncertified queries: gathered queries + their results
    const size_t fb = align_up(static_cast<size_t>(B) * D * 4, 256) + align_up(static_cast<size_t>(B) * k * 4, 256) +
                      align_up(static_cast<size_t>(B) * k * 8, 256);
    return p.total + fb + 1024;
}

int knn_tc_search(const float *Q, const float *P, const float *p_sqnorm, const void *p_half, int B, int64_t N, int D, int k,
                  float *out_d, int64_t *out_i, int64_t row_offset, void *ws, size_t ws_bytes, cudaStream_t s) {
    int rc = ac_device_check();
    if (rc) return rc;
    KnnTcPlan pl = plan_knn_tc(B, N, D, k);
    const size_t fb_q = align_up(static_cast<size_t>(B) * D * 4, 256), fb_d = align_up(static_cast<size_t>(B) * k * 4, 256),
                 fb_i = align_up(static_cast<size_t>(B) * k * 8, 256);
    const size_t exact_ws = knn_exact_workspace_pub(B, N, k);
    if (pl.total + fb_q + fb_d + fb_i + exact_ws > ws_bytes) {
        set_error("knn_tc_search: workspace needs %zu bytes, have %zu", pl.total + fb_q + fb_d + fb_i + exact_ws, ws_bytes);
        return AC_E_WORKSPACE;
    }
    uint8_t *w = static_cast<uint8_t *>(ws);
    float *Qr = reinterpret_cast<float *>(w + pl.off_qr);
    float *qn = reinterpret_cast<float *>(w + pl.off_qn);
    uint32_t *gthr = reinterpret_cast<uint32_t *>(w + pl.off_gthr);
    float *pn = reinterpret_cast<float *>(w + pl.off_pn);
    float *bmax = reinterpret_cast<float *>(w + pl.off_bmax);
    float *pmax = reinterpret_cast<float *>(w + pl.off_pmax);
    float *ckey = reinterpret_cast<float *>(w + pl.off_ckey);
    int32_t *cidx = reinterpret_cast<int32_t *>(w + pl.off_cidx);
    int64_t *cidx64 = reinterpret_cast<int64_t *>(w + pl.off_cidx64);
    float *skey = reinterpret_cast<float *>(w + pl.off_skey);
    int64_t *sidx = reinterpret_cast<int64_t *>(w + pl.off_sidx);
    int32_t *ridx = reinterpret_cast<int32_t *>(w + pl.off_ridx);
    float *T = reinterpret_cast<float *>(w + pl.off_T);
    float *rd = reinterpret_cast<float *>(w + pl.off_rd);
    int64_t *ri = reinterpret_cast<int64_t *>(w + pl.off_ri);
    int32_t *fail = reinterpret_cast<int32_t *>(w + pl.off_fail);   // [0] = count, [1..] = list
    uint8_t *selws = w + pl.off_sel;
    uint8_t *fbq = w + pl.total;
    uint8_t *fbd = fbq + fb_q;
    uint8_t *fbi = fbd + fb_d;
    uint8_t *exws = fbi + fb_i;

    const size_t Bp = static_cast<size_t>(pl.tiles_m) * GEMM_BLOCK_M;
    AC_CUDA(cudaMemsetAsync(Qr, 0, Bp * D * 4, s));
    AC_CUDA(cudaMemsetAsync(fail, 0, 4, s));
    AC_CUDA(cudaMemsetAsync(gthr, 0xFF, Bp * 4, s));   // ordered-uint +max: no bound published yet
    __half *Qh = p_half ? reinterpret_cast<__half *>(Qr) : nullptr;   // the fp16 queries reuse the fp32 query slot
    knn_prep_queries_kernel<<<(B + 3) / 4, 128, 0, s>>>(Q, B, D, Qr, Qh, qn);
    AC_LAUNCH_CHECK();
    const float *pn_use = p_sqnorm;
    const unsigned rb = static_cast<unsigned>((N + 7) / 8);
    if (p_sqnorm) {
        knn_prep_rows_kernel<<<rb, 256, 0, s>>>(P, N, D, const_cast<float *>(p_sqnorm), bmax, 1);
    } else {
        knn_prep_rows_kernel<<<rb, 256, 0, s>>>(P, N, D, pn, bmax, 0);
        pn_use = pn;
    }
    AC_LAUNCH_CHECK();
    knn_reduce_max_kernel<<<1, 256, 0, s>>>(bmax, rb, pmax);
    AC_LAUNCH_CHECK();

    // ---- coarse pass on the tensor cores
    CUtensorMap ta, tb;
    int kt = k + 3 > 8 ? k + 3 : 8;
    if (kt > KNN_KC) kt = KNN_KC;
    // CTA-pair variant (option "knn_pair"): needs an even number of query tiles so that both CTAs of a pair own real queries
    const bool pair = option(OPT_KNN_PAIR) != 0 && pl.tiles_m % 2 == 0 && pl.grid_ctas % 2 == 0;
    EpiKnn epi{pn_use, ckey, cidx, gthr, B, N, pl.tiles_m, pl.slots, kt, pair ? 1 : 0};   // slots = 2 per CTA
    const uint32_t b_box = pair ? GEMM2_B_ROWS : GEMM_BLOCK_N;
    const bool lane_epi = option(OPT_KNN_EPI) != 0;                        // per-lane slow path of the epilogue
    EpiKnnLane epi_lane{epi};
    // algorithmic work of the scan: 2.B.N.D flops, one read of the fp32 prototype matrix (4.N.D bytes); with the fp16
    // shadow the kernel actually streams 2.N.D bytes (the exact re-rank below still reads fp32 rows)
    if (p_half) {
        if ((rc = make_tmap_2d(&ta, Qh, 2, Bp, D, static_cast<uint64_t>(D) * 2, GEMM_BLOCK_M, 64))) return rc;
        if ((rc = make_tmap_2d(&tb, p_half, 2, static_cast<uint64_t>(N), D, static_cast<uint64_t>(D) * 2, b_box, 64))) return rc;
        rc = pair && lane_epi ? launch_gemm_tc2<EpiKnnLane, true, GEMM_KIND_F16>(ta, tb, static_cast<int>(Bp), static_cast<int>(N), D, epi_lane, s,
                                                                                   pl.grid_ctas, PROF_KNN_COARSE, 4.0 * static_cast<double>(N) * D)
             : pair ? launch_gemm_tc2<EpiKnn, true, GEMM_KIND_F16>(ta, tb, static_cast<int>(Bp), static_cast<int>(N), D, epi, s,
                                                                   pl.grid_ctas, PROF_KNN_COARSE, 4.0 * static_cast<double>(N) * D)
             : lane_epi ? launch_gemm_tc<EpiKnnLane, true, GEMM_KIND_F16>(ta, tb, static_cast<int>(Bp), static_cast<int>(N), D, epi_lane, s,
                                                                           pl.grid_ctas, PROF_KNN_COARSE, 4.0 * static_cast<double>(N) * D)
                  : launch_gemm_tc<EpiKnn, true, GEMM_KIND_F16>(ta, tb, static_cast<int>(Bp), static_cast<int>(N), D, epi, s,
                                                                  pl.grid_ctas, PROF_KNN_COARSE, 4.0 * static_cast<double>(N) * D);
        if (rc) return rc;
    } else {
        if ((rc = make_tmap_2d(&ta, Qr, 4, Bp, D, static_cast<uint64_t>(D) * 4, GEMM_BLOCK_M, GEMM_BLOCK_K))) return rc;
        if ((rc = make_tmap_2d(&tb, P, 4, static_cast<uint64_t>(N), D, static_cast<uint64_t>(D) * 4, b_box, GEMM_BLOCK_K))) return rc;
        rc = pair && lane_epi ? launch_gemm_tc2<EpiKnnLane, true>(ta, tb, static_cast<int>(Bp), static_cast<int>(N), D, epi_lane, s, pl.grid_ctas,
                                                                  PROF_KNN_COARSE, 4.0 * static_cast<double>(N) * D)
             : pair ? launch_gemm_tc2<EpiKnn, true>(ta, tb, static_cast<int>(Bp), static_cast<int>(N), D, epi, s, pl.grid_ctas,
                                                  PROF_KNN_COARSE, 4.0 * static_cast<double>(N) * D)
             : lane_epi ? launch_gemm_tc<EpiKnnLane, true>(ta, tb, static_cast<int>(Bp), static_cast<int>(N), D, epi_lane, s, pl.grid_ctas,
                                                           PROF_KNN_COARSE, 4.0 * static_cast<double>(N) * D)
                  : launch_gemm_tc<EpiKnn, true>(ta, tb, static_cast<int>(Bp), static_cast<int>(N), D, epi, s, pl.grid_ctas,
                                                 PROF_KNN_COARSE, 4.0 * static_cast<double>(N) * D);
        if (rc) return rc;
    }

    // ---- merge per-CTA lists, pick KP candidates + exclusion threshold
    const int64_t nc = static_cast<int64_t>(B) * pl.slots * KNN_KC;
    knn_widen_kernel<<<static_cast<unsigned>((nc + 255) / 256), 256, 0, s>>>(cidx, nc, cidx64);
    AC_LAUNCH_CHECK();
    const int64_t L = static_cast<int64_t>(pl.slots) * KNN_KC;
    if ((rc = topk_select(ckey, cidx64, B, L, L, 0, KNN_KP + 1, skey, sidx, selws, pl.sel_bytes, s))) return rc;
    knn_pick_kernel<<<(B + 127) / 128, 128, 0, s>>>(skey, sidx, ckey, qn, B, pl.slots, kt, ridx, T);
    AC_LAUNCH_CHECK();

    // ---- exact re-rank of the candidates, final (d, id) order
    if ((rc = launch_knn_rerank(Q, P, B, N, D, KNN_KP, ridx, rd, ri, row_offset, s))) return rc;
    if ((rc = topk_select(rd, ri, B, KNN_KP, KNN_KP, 0, k, out_d, out_i, selws, pl.sel_bytes, s))) return rc;

    // ---- certification + exact recomputation of the (rare) uncertified queries
    const float rel = p_half ? (2.f * 4.8828125e-4f + 2.4e-7f) : (9.765625e-4f + 4.8828125e-4f + 4.8e-7f);
    knn_certify_kernel<<<(B + 127) / 128, 128, 0, s>>>(out_d, out_i, T, qn, pmax, B, k, rel, fail + 1, fail);
    AC_LAUNCH_CHECK();
    int32_t nfail = 0;
    AC_CUDA(cudaMemcpyAsync(&nfail, fail, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    AC_CUDA(cudaStreamSynchronize(s));
    if (nfail > 0) {
        float *gq = reinterpret_cast<float *>(fbq);
        float *gd = reinterpret_cast<float *>(fbd);
        int64_t *gi = reinterpret_cast<int64_t *>(fbi);
        knn_gather_rows_kernel<<<nfail, 128, 0, s>>>(Q, fail + 1, nfail, D, gq);
        AC_LAUNCH_CHECK();
        if ((rc = knn_exact_subset(gq, P, nfail, N, D, k, gd, gi, row_offset, exws, exact_ws, s))) return rc;
        knn_scatter_results_kernel<<<(nfail * k + 127) / 128, 128, 0, s>>>(gd, gi, fail + 1, nfail, k, out_d, out_i);
        AC_LAUNCH_CHECK();
    }
    return AC_OK;
}

__global__ void knn_to_half_kernel(const float *__restrict__ in, __half *__restrict__ out, int64_t n) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = __float2half_rn(in[i]);
}

}  // namespace ac

extern "C" int ac_knn_make_shadow(const float *P, int64_t N, int D, void *out_half, ac_stream_t stream) {
    AC_REQUIRE(P && out_half && N >= 0 && D > 0, "ac_knn_make_shadow: bad arguments");
    if (N == 0) return AC_OK;
    ac::knn_to_half_kernel<<<1184, 256, 0, static_cast<cudaStream_t>(stream)>>>(P, static_cast<__half *>(out_half), N * D);
    AC_LAUNCH_CHECK();
    return AC_OK;
}

