// knn_exact.cu -- stage K, exact path: fp32 SIMT squared-L2 scan in the oracle's lane order,
// chunked bitonic top-k by (d, id), exp(-d) softmax scores, shard merge, segment mean.
//
// Replaces faiss.IndexFlatL2.search at /root/reference/src/adaptive_classifier/memory.py:110-114 and the
// post-processing at memory.py:117,128-134.  Distances reproduce oracle/knn_oracle.c::oracle_l2sqr bit
// for bit: lane j accumulates (x[i]-y[i])^2 for i == j (mod 8) in index order (separate mul and add
// roundings), lanes combined ((l0+l4)+(l1+l5))+((l2+l6)+(l3+l7)).
#include "common.cuh"
#include <math_constants.h>

namespace ac {

// ------------------------------------------------------------------------------------------------
// exact distances: Dout[b, n] for b in [q0, q0+nq), all n.   CTA = 128 rows x QB queries.
// P tile streamed through smem in D-chunks of 32 floats with cp.async double buffering; a thread owns one
// row and keeps 8 lane accumulators per query.
// ------------------------------------------------------------------------------------------------
constexpr int KE_ROWS = 128;
constexpr int KE_DC = 32;              // floats per D chunk (static smem stays under 48 KB)
constexpr int KE_STRIDE = KE_DC + 4;   // padded row stride (floats): conflict-free float4 reads

__device__ __forceinline__ void cp_async_16(void *dst, const void *src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_4(void *dst, const void *src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int QB>
__global__ void __launch_bounds__(KE_ROWS)
knn_dist_exact_kernel(const float *__restrict__ Q, const float *__restrict__ P, int nq_total, int64_t N, int D,
                      float *__restrict__ Dout /* [nq_total, N] */) {
    __shared__ __align__(16) float sP[2][KE_ROWS * KE_STRIDE];
    __shared__ __align__(16) float sQ[2][QB * KE_DC];

    const int tid = threadIdx.x;
    const int64_t row0 = static_cast<int64_t>(blockIdx.x) * KE_ROWS;
    const int qb0 = blockIdx.y * QB;
    const int nchunks = (D + KE_DC - 1) / KE_DC;
    const bool vec_ok = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(P) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(Q) & 15) == 0);

    auto issue_chunk = [&](int c, int buf) {
        const int d0 = c * KE_DC;
        if (vec_ok) {
            // 128 rows x 8 float4 16-byte copies, coalesced along D
            for (int e = tid; e < KE_ROWS * (KE_DC / 4); e += KE_ROWS) {
                const int r = e / (KE_DC / 4), v = e % (KE_DC / 4);
                const int64_t row = row0 + r;
                const int d = d0 + v * 4;
                const bool ok = (row < N) && (d < D);
                const float *src = ok ? (P + row * D + d) : P;
                cp_async_16(&sP[buf][r * KE_STRIDE + v * 4], src, ok ? 16 : 0);
            }
            for (int e = tid; e < QB * (KE_DC / 4); e += KE_ROWS) {
                const int r = e / (KE_DC / 4), v = e % (KE_DC / 4);
                const int qi = qb0 + r;
                const int d = d0 + v * 4;
                const bool ok = (qi < nq_total) && (d < D);
                const float *src = ok ? (Q + static_cast<int64_t>(qi) * D + d) : Q;
                cp_async_16(&sQ[buf][r * KE_DC + v * 4], src, ok ? 16 : 0);
            }
        } else {
            for (int e = tid; e < KE_ROWS * KE_DC; e += KE_ROWS) {
                const int r = e / KE_DC, v = e % KE_DC;
                const int64_t row = row0 + r;
                const int d = d0 + v;
                const bool ok = (row < N) && (d < D);
                const float *src = ok ? (P + row * D + d) : P;
                cp_async_4(&sP[buf][r * KE_STRIDE + v], src, ok ? 4 : 0);
            }
            for (int e = tid; e < QB * KE_DC; e += KE_ROWS) {
                const int r = e / KE_DC, v = e % KE_DC;
                const int qi = qb0 + r;
                const int d = d0 + v;
                const bool ok = (qi < nq_total) && (d < D);
                const float *src = ok ? (Q + static_cast<int64_t>(qi) * D + d) : Q;
                cp_async_4(&sQ[buf][r * KE_DC + v], src, ok ? 4 : 0);
            }
        }
        cp_async_commit();
    };

    float acc[QB][8];
#pragma unroll
    for (int q = 0; q < QB; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;

    issue_chunk(0, 0);
    for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        if (c + 1 < nchunks) {
            issue_chunk(c + 1, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float4 *prow = reinterpret_cast<const float4 *>(&sP[buf][tid * KE_STRIDE]);
#pragma unroll 4
        for (int v = 0; v < KE_DC / 8; ++v) {   // 8 floats (one lane round) per iteration
            const float4 p0 = prow[2 * v], p1 = prow[2 * v + 1];
#pragma unroll
            for (int q = 0; q < QB; ++q) {
                const float4 q0 = *reinterpret_cast<const float4 *>(&sQ[buf][q * KE_DC + 8 * v]);
                const float4 q1 = *reinterpret_cast<const float4 *>(&sQ[buf][q * KE_DC + 8 * v + 4]);
                float t;
                // zero-padded tail elements contribute (0-0)^2 = +0 exactly, as in the oracle
                t = __fsub_rn(q0.x, p0.x); acc[q][0] = __fadd_rn(acc[q][0], __fmul_rn(t, t));
                t = __fsub_rn(q0.y, p0.y); acc[q][1] = __fadd_rn(acc[q][1], __fmul_rn(t, t));
                t = __fsub_rn(q0.z, p0.z); acc[q][2] = __fadd_rn(acc[q][2], __fmul_rn(t, t));
                t = __fsub_rn(q0.w, p0.w); acc[q][3] = __fadd_rn(acc[q][3], __fmul_rn(t, t));
                t = __fsub_rn(q1.x, p1.x); acc[q][4] = __fadd_rn(acc[q][4], __fmul_rn(t, t));
                t = __fsub_rn(q1.y, p1.y); acc[q][5] = __fadd_rn(acc[q][5], __fmul_rn(t, t));
                t = __fsub_rn(q1.z, p1.z); acc[q][6] = __fadd_rn(acc[q][6], __fmul_rn(t, t));
                t = __fsub_rn(q1.w, p1.w); acc[q][7] = __fadd_rn(acc[q][7], __fmul_rn(t, t));
            }
        }
        __syncthreads();
    }

    const int64_t row = row0 + tid;
    if (row < N) {
#pragma unroll
        for (int q = 0; q < QB; ++q) {
            const int qi = qb0 + q;
            if (qi < nq_total) {
                const float a = __fadd_rn(acc[q][0], acc[q][4]);
                const float b = __fadd_rn(acc[q][1], acc[q][5]);
                const float c2 = __fadd_rn(acc[q][2], acc[q][6]);
                const float e = __fadd_rn(acc[q][3], acc[q][7]);
                Dout[static_cast<int64_t>(qi) * N + row] = __fadd_rn(__fadd_rn(a, b), __fadd_rn(c2, e));
            }
        }
    }
}

int launch_knn_dist_exact(const float *Q, const float *P, int nq, int64_t N, int D, float *Dout,
                          cudaStream_t stream) {
    if (nq <= 0 || N <= 0) return AC_OK;
    const unsigned gx = static_cast<unsigned>((N + KE_ROWS - 1) / KE_ROWS);
    if (nq == 1) {
        knn_dist_exact_kernel<1><<<dim3(gx, 1), KE_ROWS, 0, stream>>>(Q, P, nq, N, D, Dout);
    } else if (nq <= 2) {
        knn_dist_exact_kernel<2><<<dim3(gx, 1), KE_ROWS, 0, stream>>>(Q, P, nq, N, D, Dout);
    } else if (nq <= 4) {
        knn_dist_exact_kernel<4><<<dim3(gx, 1), KE_ROWS, 0, stream>>>(Q, P, nq, N, D, Dout);
    } else {
        knn_dist_exact_kernel<8><<<dim3(gx, (nq + 7) / 8), KE_ROWS, 0, stream>>>(Q, P, nq, N, D, Dout);
    }
    AC_LAUNCH_CHECK();
    return AC_OK;
}

// exact distance of explicit (query, row) candidate pairs -- the re-rank of the tensor path.
// cand_idx[b, kc] local row ids (or < 0 = empty); one 8-lane group per candidate, lane j owns residue j.
__global__ void knn_rerank_kernel(const float *__restrict__ Q, const float *__restrict__ P, int B, int64_t N, int D,
                                  int kc, const int32_t *__restrict__ cand_idx, float *__restrict__ out_d,
                                  int64_t *__restrict__ out_i, int64_t row_offset) {
    const int64_t gid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3;  // candidate id
    const int j = threadIdx.x & 7;
    const int64_t total = static_cast<int64_t>(B) * kc;
    const bool valid = gid < total;
    int32_t r = -1;
    int b = 0;
    if (valid) {
        b = static_cast<int>(gid / kc);
        r = cand_idx[gid];
    }
    float acc = 0.f;
    if (valid && r >= 0 && r < N) {
        const float *q = Q + static_cast<int64_t>(b) * D;
        const float *p = P + static_cast<int64_t>(r) * D;
        for (int i = j; i < D; i += 8) {
            const float t = __fsub_rn(q[i], p[i]);
            acc = __fadd_rn(acc, __fmul_rn(t, t));
        }
    }
    // ((l0+l4)+(l1+l5))+((l2+l6)+(l3+l7)) inside each aligned 8-lane group
    const unsigned m = 0xffffffffu;
    const float s4 = __fadd_rn(acc, __shfl_xor_sync(m, acc, 4));   // lanes 0..3 hold l_j + l_{j+4}
    const float s1 = __fadd_rn(s4, __shfl_xor_sync(m, s4, 1));     // lane 0: (l0+l4)+(l1+l5); lane 2: (l2+l6)+(l3+l7)
    const float s2 = __fadd_rn(s1, __shfl_xor_sync(m, s1, 2));
    if (valid && j == 0) {
        const bool ok = (r >= 0 && r < N);
        out_d[gid] = ok ? s2 : CUDART_INF_F;
        out_i[gid] = ok ? (static_cast<int64_t>(r) + row_offset) : -1;
    }
}

int launch_knn_rerank(const float *Q, const float *P, int B, int64_t N, int D, int kc, const int32_t *cand,
                      float *out_d, int64_t *out_i, int64_t row_offset, cudaStream_t stream) {
    const int64_t threads = static_cast<int64_t>(B) * kc * 8;
    if (threads <= 0) return AC_OK;
    const int bs = 256;
    knn_rerank_kernel<<<static_cast<unsigned>((threads + bs - 1) / bs), bs, 0, stream>>>(Q, P, B, N, D, kc, cand,
                                                                                         out_d, out_i, row_offset);
    AC_LAUNCH_CHECK();
    return AC_OK;
}

// ------------------------------------------------------------------------------------------------
// chunked bitonic top-k by (d, id).  Input list per query: d[b, L] and either explicit ids idx[b, L]
// (entries < 0 are padding) or implicit ids (position + id_offset).  Each CTA sorts one chunk of up to
// SEL_CHUNK entries in shared memory and writes its k best to out[b, chunk, k].  Levels are chained by the
// host until one chunk remains.
// ------------------------------------------------------------------------------------------------
constexpr int SEL_CHUNK = 4096;
constexpr int SEL_THREADS = 512;
constexpr long long SEL_PAD_ID = 0x7fffffffffffffffLL;

__device__ __forceinline__ bool cand_less(float da, long long ia, float db, long long ib) {
    return (da < db) || (da == db && ia < ib);
}

__global__ void __launch_bounds__(SEL_THREADS)
topk_chunk_kernel(const float *__restrict__ d, const int64_t *__restrict__ idx, int64_t L, int64_t in_stride,
                  int64_t id_offset, int k, float *__restrict__ out_d, int64_t *__restrict__ out_i,
                  int64_t out_stride /* per query */, int n2 /* power of two >= entries of a chunk, <= SEL_CHUNK */,
                  const float *__restrict__ row_gate /* nullable: rows with gate == -inf are left untouched */) {
    extern __shared__ __align__(16) uint8_t sel_smem[];
    float *sd = reinterpret_cast<float *>(sel_smem);
    long long *si = reinterpret_cast<long long *>(sel_smem + SEL_CHUNK * sizeof(float));

    const int b = blockIdx.y;
    if (row_gate && row_gate[b] == -CUDART_INF_F) return;        // block-uniform
    const int chunk = blockIdx.x;
    const int64_t base = static_cast<int64_t>(chunk) * SEL_CHUNK;
    const float *db = d + static_cast<int64_t>(b) * in_stride;
    const int64_t *ib = idx ? idx + static_cast<int64_t>(b) * in_stride : nullptr;

    for (int e = threadIdx.x; e < n2; e += SEL_THREADS) {
        const int64_t pos = base + e;
        float dv = CUDART_INF_F;
        long long iv = SEL_PAD_ID;
        if (pos < L) {
            const long long id = ib ? static_cast<long long>(ib[pos]) : static_cast<long long>(pos + id_offset);
            if (id >= 0) {
                dv = db[pos];
                iv = id;
                if (dv != dv) dv = CUDART_INF_F;   // NaN distances sort last
            }
        }
        sd[e] = dv;
        si[e] = iv;
    }
    __syncthreads();

    // bitonic sort ascending by (d, id)
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < n2 / 2; t += SEL_THREADS) {
                const int lo = 2 * t - (t & (stride - 1));   // index with bit `stride` clear
                const int hi = lo + stride;
                const bool asc = ((lo & size) == 0);
                const float dl = sd[lo], dh = sd[hi];
                const long long il = si[lo], ih = si[hi];
                const bool swap = asc ? cand_less(dh, ih, dl, il) : cand_less(dl, il, dh, ih);
                if (swap) {
                    sd[lo] = dh; sd[hi] = dl;
                    si[lo] = ih; si[hi] = il;
                }
            }
            __syncthreads();
        }
    }

    float *od = out_d + static_cast<int64_t>(b) * out_stride + static_cast<int64_t>(chunk) * k;
    int64_t *oi = out_i + static_cast<int64_t>(b) * out_stride + static_cast<int64_t>(chunk) * k;
    for (int e = threadIdx.x; e < k; e += SEL_THREADS) {
        const bool pad = (e >= n2) || (si[e] == SEL_PAD_ID);
        od[e] = pad ? CUDART_INF_F : sd[e];
        oi[e] = pad ? -1 : static_cast<int64_t>(si[e]);
    }
}

// workspace needed by topk_select for a list of length L (per query) at batch B
size_t topk_select_workspace(int B, int64_t L, int k) {
    size_t total = 0;
    int64_t len = L;
    while (len > SEL_CHUNK) {
        const int64_t chunks = (len + SEL_CHUNK - 1) / SEL_CHUNK;
        const int64_t out_len = chunks * k;
        total += align_up(static_cast<size_t>(B) * out_len * sizeof(float), 256);
        total += align_up(static_cast<size_t>(B) * out_len * sizeof(int64_t), 256);
        len = out_len;
    }
    return total + 256;
}

// d[B, L] (+ idx or implicit ids) -> out_d[B,k], out_i[B,k] sorted ascending by (d, id)
// row_gate (nullable, [B]): queries whose gate is -inf are skipped (their output rows keep their contents)
int topk_select(const float *d, const int64_t *idx, int B, int64_t L, int64_t in_stride, int64_t id_offset, int k,
                float *out_d, int64_t *out_i, void *ws, size_t ws_bytes, cudaStream_t stream, const float *row_gate = nullptr) {
    AC_REQUIRE(k >= 1 && k <= AC_KNN_MAX_K, "topk_select: k=%d outside [1,%d]", k, AC_KNN_MAX_K);
    const int smem = SEL_CHUNK * (sizeof(float) + sizeof(long long));
    AC_CUDA(cudaFuncSetAttribute(topk_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (B <= 0) return AC_OK;
    uint8_t *wp = static_cast<uint8_t *>(ws);
    size_t used = 0;
    const float *cur_d = d;
    const int64_t *cur_i = idx;
    int64_t cur_L = L, cur_stride = in_stride, cur_off = id_offset;
    while (cur_L > SEL_CHUNK) {
        const int64_t chunks = (cur_L + SEL_CHUNK - 1) / SEL_CHUNK;
        const int64_t out_len = chunks * k;
        const size_t bd = align_up(static_cast<size_t>(B) * out_len * sizeof(float), 256);
        const size_t bi = align_up(static_cast<size_t>(B) * out_len * sizeof(int64_t), 256);
        if (used + bd + bi > ws_bytes) {
            set_error("topk_select: workspace too small (%zu needed > %zu)", used + bd + bi, ws_bytes);
            return AC_E_WORKSPACE;
        }
        float *nd = reinterpret_cast<float *>(wp + used);
        int64_t *ni = reinterpret_cast<int64_t *>(wp + used + bd);
        used += bd + bi;
        topk_chunk_kernel<<<dim3(static_cast<unsigned>(chunks), B), SEL_THREADS, smem, stream>>>(
            cur_d, cur_i, cur_L, cur_stride, cur_off, k, nd, ni, out_len, SEL_CHUNK, row_gate);
        AC_LAUNCH_CHECK();
        cur_d = nd; cur_i = ni; cur_L = out_len; cur_stride = out_len; cur_off = 0;
    }
    int n2 = 32;
    while (n2 < cur_L) n2 <<= 1;            // the final list is short (candidates, head classes): sort only that much
    topk_chunk_kernel<<<dim3(1, B), SEL_THREADS, smem, stream>>>(cur_d, cur_i, cur_L, cur_stride, cur_off, k, out_d,
                                                                 out_i, k, n2, row_gate);
    AC_LAUNCH_CHECK();
    return AC_OK;
}

// ------------------------------------------------------------------------------------------------
// memory.py:117,128-134: scores = softmax_k(exp(-d)); one warp per query row.
// ------------------------------------------------------------------------------------------------
__global__ void proto_scores_kernel(const float *__restrict__ d, const int64_t *__restrict__ idx, int B, int k,
                                    float *__restrict__ scores) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B) return;
    const float *dr = d + static_cast<int64_t>(row) * k;
    const int64_t *ir = idx ? idx + static_cast<int64_t>(row) * k : nullptr;
    float *sr = scores + static_cast<int64_t>(row) * k;
    float mx = -CUDART_INF_F;
    for (int j = lane; j < k; j += 32) {
        const bool ok = !ir || ir[j] >= 0;
        if (ok) mx = fmaxf(mx, expf(-dr[j]));
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < k; j += 32) {
        const bool ok = !ir || ir[j] >= 0;
        const float e = ok ? expf(expf(-dr[j]) - mx) : 0.f;
        sr[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    for (int j = lane; j < k; j += 32) sr[j] = sr[j] / sum;
}

// ||p||^2 per row (coarse-pass operand; summation order is not pinned)
__global__ void row_sqnorm_kernel(const float *__restrict__ P, int64_t N, int D, float *__restrict__ out) {
    const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= N) return;
    const float *p = P + row * D;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) s = fmaf(p[i], p[i], s);
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}

// memory.py:149-150: mean over the rows of each class, rows summed in index order (deterministic)
__global__ void segment_mean_kernel(const float *__restrict__ X, const int32_t *__restrict__ cls, int64_t n, int D,
                                    int C, float *__restrict__ mean, int32_t *__restrict__ count) {
    const int c = blockIdx.x;
    if (c >= C) return;
    for (int col = threadIdx.x; col < D; col += blockDim.x) {
        float s = 0.f;
        int cnt = 0;
        for (int64_t r = 0; r < n; ++r) {
            if (cls[r] == c) {
                s = __fadd_rn(s, X[r * D + col]);
                ++cnt;
            }
        }
        mean[static_cast<int64_t>(c) * D + col] = cnt > 0 ? s / static_cast<float>(cnt) : 0.f;
        if (col == 0 && count) count[c] = cnt;
    }
}

__global__ void __launch_bounds__(SEL_THREADS)
topk_merge_kernel(const float *d, const int64_t *idx, int G, int B, int k, float *out_d, int64_t *out_i) {
    extern __shared__ __align__(16) uint8_t sel_smem[];
    float *sd = reinterpret_cast<float *>(sel_smem);
    long long *si = reinterpret_cast<long long *>(sel_smem + SEL_CHUNK * sizeof(float));
    const int b = blockIdx.x;
    const int total = G * k;
    int n2 = 1;
    while (n2 < total) n2 <<= 1;
    for (int e = threadIdx.x; e < n2; e += SEL_THREADS) {
        float dv = CUDART_INF_F;
        long long iv = SEL_PAD_ID;
        if (e < total) {
            const int g = e / k, j = e % k;
            const int64_t off = (static_cast<int64_t>(g) * B + b) * k + j;
            const long long id = idx[off];
            if (id >= 0) { dv = d[off]; iv = id; if (dv != dv) dv = CUDART_INF_F; }
        }
        sd[e] = dv; si[e] = iv;
    }
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < n2 / 2; t += SEL_THREADS) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool asc = ((lo & size) == 0);
                const float dl = sd[lo], dh = sd[hi];
                const long long il = si[lo], ih = si[hi];
                const bool swap = asc ? cand_less(dh, ih, dl, il) : cand_less(dl, il, dh, ih);
                if (swap) { sd[lo] = dh; sd[hi] = dl; si[lo] = ih; si[hi] = il; }
            }
            __syncthreads();
        }
    }
    for (int e = threadIdx.x; e < k; e += SEL_THREADS) {
        const bool pad = (e >= n2) || (si[e] == SEL_PAD_ID);
        out_d[static_cast<int64_t>(b) * k + e] = pad ? CUDART_INF_F : sd[e];
        out_i[static_cast<int64_t>(b) * k + e] = pad ? -1 : static_cast<int64_t>(si[e]);
    }
}


}  // namespace ac

// ================================================================================================
// C ABI
// ================================================================================================
using namespace ac;

extern "C" int ac_row_sqnorm(const float *P, int64_t N, int D, float *out, ac_stream_t stream) {
    AC_REQUIRE(P && out && N >= 0 && D > 0, "ac_row_sqnorm: bad arguments");
    if (N == 0) return AC_OK;
    const int wpb = 8;
    row_sqnorm_kernel<<<static_cast<unsigned>((N + wpb - 1) / wpb), wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        P, N, D, out);
    AC_LAUNCH_CHECK();
    return AC_OK;
}

extern "C" int ac_topk_merge(const float *d, const int64_t *i, int G, int B, int k, float *out_d, int64_t *out_i,
                             ac_stream_t stream) {
    AC_REQUIRE(d && i && out_d && out_i && G >= 1 && B >= 0 && k >= 1, "ac_topk_merge: bad arguments");
    AC_REQUIRE(static_cast<int64_t>(G) * k <= SEL_CHUNK, "ac_topk_merge: G*k = %lld exceeds %d",
               static_cast<long long>(G) * k, SEL_CHUNK);
    if (B == 0) return AC_OK;
    // one CTA per query gathers its G*k candidates from the [G,B,k] slabs and sorts them by (d, id)
    const int smem = SEL_CHUNK * (sizeof(float) + sizeof(long long));
    AC_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    topk_merge_kernel<<<B, SEL_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(d, i, G, B, k, out_d, out_i);
    AC_LAUNCH_CHECK();
    return AC_OK;
}

extern "C" int ac_proto_scores(const float *d, const int64_t *idx, int B, int k, float *scores, ac_stream_t stream) {
    AC_REQUIRE(d && scores && B >= 0 && k >= 1, "ac_proto_scores: bad arguments");
    if (B == 0) return AC_OK;
    const int wpb = 4;
    proto_scores_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(d, idx, B, k, scores);
    AC_LAUNCH_CHECK();
    return AC_OK;
}

extern "C" int ac_segment_mean(const float *X, const int32_t *cls, int64_t n, int D, int C, float *mean,
                               int32_t *count, ac_stream_t stream) {
    AC_REQUIRE(X && cls && mean && n >= 0 && D > 0 && C > 0, "ac_segment_mean: bad arguments");
    segment_mean_kernel<<<C, 256, 0, static_cast<cudaStream_t>(stream)>>>(X, cls, n, D, C, mean, count);
    AC_LAUNCH_CHECK();
    return AC_OK;
}

// ================================================================================================
// Device-resident prototype memory maintenance.
//
// Replaces, for a whole add_examples() call at once, the per-example sequence of
//   /root/reference/src/adaptive_classifier/memory.py:60-72   append; if over max_examples_per_class -> _prune_examples
//   memory.py:196-217  _prune_examples: mean of the cap + 1 stored embeddings, L2 distance of each to it, argsort ascending,
//                      keep the first `cap` IN SORTED ORDER (the list is reordered by distance, the farthest is dropped)
//   memory.py:138-153  _update_prototype: prototype = mean of the retained embeddings
// The reference restacks all embeddings of the class on every add (O(n^2) over a continual loop, a Python .item() loop per
// prune).  Here every class keeps its rows in HBM ([cap + 1, D] slots + a logical order), one CTA per touched class walks ITS new
// examples sequentially (the prune of example j sees the list example j-1 left, as in the reference), classes run in parallel:
//   sum S (fp64, recomputed from the stored rows once per call, then updated incrementally) -> mean -> one pass of distances over
//   the cap + 1 rows -> bitonic sort of (distance, logical position) in shared memory -> new logical order, farthest row's slot freed.
// Output per class: where every retained position came from (old position or new example), so the host mirrors the same order on
// its Example lists, and the prototype (mean of the retained rows).
// ================================================================================================
namespace ac {
constexpr int MEM_THREADS = 1024;
constexpr int MEM_MAX_CAP = 2047;          // cap + 1 <= 2048 sort slots

__global__ void __launch_bounds__(MEM_THREADS)
memory_append_prune_kernel(float *__restrict__ rows, int32_t *__restrict__ order, int32_t *__restrict__ count, int cap, int D,
                           const float *__restrict__ new_rows, const int32_t *__restrict__ new_index, const int32_t *__restrict__ cls_start,
                           const int32_t *__restrict__ touched, int32_t *__restrict__ src_out, float *__restrict__ proto_out,
                           double *__restrict__ sum_ws /* [n_touched, D] */) {
    extern __shared__ __align__(16) uint8_t mem_smem[];
    float *skey = reinterpret_cast<float *>(mem_smem);                       // [2048] distances
    int32_t *spos = reinterpret_cast<int32_t *>(mem_smem + 2048 * 4);        // [2048] logical positions
    int32_t *ssrc = reinterpret_cast<int32_t *>(mem_smem + 2048 * 8);        // [2048] source id of every logical position
    int32_t *sord = reinterpret_cast<int32_t *>(mem_smem + 2048 * 12);       // [2048] logical position -> physical slot
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = touched[t];
    float *R = rows + static_cast<int64_t>(c) * (cap + 1) * D;
    int32_t *ord = order + static_cast<int64_t>(c) * (cap + 1);
    double *S = sum_ws + static_cast<int64_t>(t) * D;
    int n = count[c];
    const int n_old = n;
    // invariant (kept by the host when it builds a store and by every step below): order[0..cap] is a permutation of the
    // physical slots 0..cap; positions [0, n) are the stored rows in list order, positions [n, cap] the free slots
    for (int i = tid; i <= cap; i += MEM_THREADS) { sord[i] = ord[i]; ssrc[i] = i < n ? i : -1; }
    __syncthreads();
    // fresh fp64 column sums of the stored rows (one pass; coalesced along D)
    for (int d = tid; d < D; d += MEM_THREADS) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += static_cast<double>(R[static_cast<int64_t>(sord[i]) * D + d]);
        S[d] = s;
    }
    __syncthreads();
    const int j0 = cls_start[t], j1 = cls_start[t + 1];
    for (int j = j0; j < j1; ++j) {
        // ---- append the new example into the first free physical slot
        if (tid == 0) ssrc[n] = n_old + (j - j0);
        const int slot = sord[n];
        const float *nr = new_rows + static_cast<int64_t>(new_index[j]) * D;
        for (int d = tid; d < D; d += MEM_THREADS) {
            const float v = nr[d];
            R[static_cast<int64_t>(slot) * D + d] = v;
            S[d] += static_cast<double>(v);
        }
        ++n;
        __syncthreads();
        if (n <= cap) continue;
        // ---- over the cap: distance of every stored row to the mean of all cap + 1, one warp per row
        const double inv = 1.0 / static_cast<double>(n);
        for (int i = warp; i < n; i += MEM_THREADS / 32) {
            const float *r = R + static_cast<int64_t>(sord[i]) * D;
            float acc = 0.f;
            for (int d = lane; d < D; d += 32) {
                const float m = static_cast<float>(S[d] * inv);
                const float df = r[d] - m;
                acc = fmaf(df, df, acc);
            }
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) { skey[i] = sqrtf(acc); spos[i] = i; }
        }
        for (int i = n + tid; i < 2048; i += MEM_THREADS) { skey[i] = CUDART_INF_F; spos[i] = 0x7fffffff; }
        __syncthreads();
        // ---- ascending bitonic sort by (distance, logical position): ties keep list order
        for (int size = 2; size <= 2048; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                const int lo = 2 * tid - (tid & (stride - 1));
                const int hi = lo + stride;
                const bool asc = ((lo & size) == 0);
                const float kl = skey[lo], kh = skey[hi];
                const int pl = spos[lo], ph = spos[hi];
                const bool less_hl = (kh < kl) || (kh == kl && ph < pl);
                const bool less_lh = (kl < kh) || (kl == kh && pl < ph);
                if (asc ? less_hl : less_lh) { skey[lo] = kh; skey[hi] = kl; spos[lo] = ph; spos[hi] = pl; }
                __syncthreads();
            }
        }
        // ---- new logical order = sorted order; the farthest row (sorted position cap) is evicted and its slot freed
        int new_slot = -1, new_src = -1;
        if (tid <= cap) { new_slot = sord[spos[tid]]; new_src = ssrc[spos[tid]]; }
        __syncthreads();
        if (tid <= cap) { sord[tid] = new_slot; ssrc[tid] = tid < cap ? new_src : -1; }
        __syncthreads();
        {
            const float *ev = R + static_cast<int64_t>(sord[cap]) * D;        // sord[cap] now holds the evicted row's slot (free)
            for (int d = tid; d < D; d += MEM_THREADS) S[d] -= static_cast<double>(ev[d]);
        }
        n = cap;
        __syncthreads();
    }
    // ---- results: order, count, provenance of the retained positions, prototype = mean of the retained rows
    for (int i = tid; i <= cap; i += MEM_THREADS) ord[i] = sord[i];
    for (int i = tid; i < cap; i += MEM_THREADS) src_out[static_cast<int64_t>(t) * cap + i] = i < n ? ssrc[i] : -1;
    // the incremental fp64 sums carry ~1e-16 relative error per update; the prototype is taken from a fresh pass all the same
    for (int d = tid; d < D; d += MEM_THREADS) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += static_cast<double>(R[static_cast<int64_t>(sord[i]) * D + d]);
        proto_out[static_cast<int64_t>(t) * D + d] = static_cast<float>(s / static_cast<double>(n > 0 ? n : 1));
    }
    if (tid == 0) count[c] = n;
}
}  // namespace ac

extern "C" int ac_memory_append_prune(float *rows, int32_t *order, int32_t *count, int cap, int D, const float *new_rows,
                                      const int32_t *new_index, const int32_t *cls_start, const int32_t *touched, int n_touched,
                                      int32_t *src_out, float *proto_out, void *workspace, size_t workspace_bytes, ac_stream_t stream) {
    AC_REQUIRE(rows && order && count && new_rows && new_index && cls_start && touched && src_out && proto_out && workspace,
               "ac_memory_append_prune: null argument");
    AC_REQUIRE(cap >= 1 && cap <= MEM_MAX_CAP && D >= 1 && n_touched >= 0, "ac_memory_append_prune: cap=%d outside [1,%d] or bad sizes", cap,
               MEM_MAX_CAP);
    if (n_touched == 0) return AC_OK;
    const size_t need = static_cast<size_t>(n_touched) * D * sizeof(double);
    if (need > workspace_bytes) { set_error("ac_memory_append_prune: workspace needs %zu bytes", need); return AC_E_WORKSPACE; }
    const int smem = 2048 * 16;
    AC_CUDA(cudaFuncSetAttribute(memory_append_prune_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    memory_append_prune_kernel<<<n_touched, MEM_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
        rows, order, count, cap, D, new_rows, new_index, cls_start, touched, src_out, proto_out, static_cast<double *>(workspace));
    AC_LAUNCH_CHECK();
    return AC_OK;
}
