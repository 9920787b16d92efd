// head.cu -- stage H: AdaptiveHead forward, one fused-sequence optimizer step (fwd + CE/BCE + bwd +
// EWC gradient + global-norm clip + AdamW), Fisher accumulation and the EWC penalty.  fp32 SIMT.
//
// Replaces (paths relative to /root/reference/src/adaptive_classifier/):
//   models.py:71-80                 AdaptiveHead.forward (Linear-ReLU-Dropout x2, Linear)
//   classifier.py:333-351,1489-1505 zero_grad / forward / CrossEntropyLoss / backward /
//                                   clip_grad_norm_(1.0) / AdamW.step
//   multilabel.py:41-44,387-397     sigmoid head + BCELoss
//   ewc.py:67-92, :96-115           Fisher accumulation and penalty
// The head is ~0.9 M parameters and M = 32 rows per step: latency-bound, so the kernels are small and
// the step is a fixed launch sequence (graph-capturable: no host sync inside).
#include "common.cuh"
#include <math_constants.h>
#include <vector>

namespace ac {

// ------------------------------------------------------------------------------------------------
// generic strided SIMT GEMM:  C[m,n] = epi( sum_k A(m,k) * B(k,n) )
//   A(m,k) = A[m*sam + k*sak],  B(k,n) = B[k*sbk + n*sbn]
// 64x64 tile, 16x16 threads, 4x4 micro-tile, BK = 16.  k is summed in ascending order per thread.
// ------------------------------------------------------------------------------------------------
enum { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_RELU = 2, EPI_BIAS_RELU_MASK = 3, EPI_RELUGRAD_MASK = 4 };

struct SgemmEpi {
    int kind;
    const float *bias;   // [n]
    const float *mask;   // [m,n] dropout mask (0 or 1/(1-p)), nullable
    const float *act;    // [m,n] saved activation for relu-grad
};

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

__global__ void __launch_bounds__(256)
sgemm_kernel(const float *__restrict__ A, int64_t sam, int64_t sak, const float *__restrict__ B, int64_t sbk,
             int64_t sbn, float *__restrict__ C, int64_t ldc, int M, int N, int K, SgemmEpi epi) {
    __shared__ float sA[SG_BK][SG_BM + 1];
    __shared__ float sB[SG_BK][SG_BN + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += SG_BK) {
        for (int e = threadIdx.x; e < SG_BM * SG_BK; e += 256) {
            // pick the faster-varying index to follow the contiguous stride of the operand
            int mm, kk;
            if (sak == 1) { kk = e % SG_BK; mm = e / SG_BK; } else { mm = e % SG_BM; kk = e / SG_BM; }
            const int m = m0 + mm, k = k0 + kk;
            sA[kk][mm] = (m < M && k < K) ? A[m * sam + k * sak] : 0.f;
        }
        for (int e = threadIdx.x; e < SG_BN * SG_BK; e += 256) {
            int nn, kk;
            if (sbk == 1) { kk = e % SG_BK; nn = e / SG_BK; } else { nn = e % SG_BN; kk = e / SG_BN; }
            const int n = n0 + nn, k = k0 + kk;
            sB[kk][nn] = (n < N && k < K) ? B[k * sbk + n * sbn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SG_BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty + 16 * i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx + 16 * j;
            if (n >= N) continue;
            float v = acc[i][j];
            const int64_t off = static_cast<int64_t>(m) * ldc + n;
            switch (epi.kind) {
                case EPI_BIAS: v += epi.bias[n]; break;
                case EPI_BIAS_RELU: v = fmaxf(v + epi.bias[n], 0.f); break;
                case EPI_BIAS_RELU_MASK:
                    v = fmaxf(v + epi.bias[n], 0.f);
                    if (epi.mask) v *= epi.mask[off];
                    break;
                case EPI_RELUGRAD_MASK:
                    // d(pre-activation) = d(out) * mask * [act > 0]   (act = relu(a)*mask; mask = 0 kills it)
                    if (epi.mask) v *= epi.mask[off];
                    v = (epi.act[off] > 0.f) ? v : 0.f;
                    break;
                default: break;
            }
            C[off] = v;
        }
    }
}

static int sgemm(const float *A, int64_t sam, int64_t sak, const float *B, int64_t sbk, int64_t sbn, float *C,
                 int64_t ldc, int M, int N, int K, SgemmEpi epi, cudaStream_t s) {
    if (M <= 0 || N <= 0) return AC_OK;
    dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM);
    sgemm_kernel<<<grid, 256, 0, s>>>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, epi);
    AC_LAUNCH_CHECK();
    return AC_OK;
}

// ------------------------------------------------------------------------------------------------
// skinny linears (M = batch rows, 32 per pass): the 64x64-tile SGEMM above runs them on a dozen CTAs and is
// latency-bound (~50 us each at M = 32); these two kernels keep every load independent instead.
//
// rowdot   Y[b,n] = epi( sum_k X[b,k] * W[n,k] + bias[n] )      W row-major [N,K] (nn.Linear layout), forward
//          warp = 4 output columns x 32 batch rows, lanes stride K, warp transpose-reduce at the end
// colacc   Z[b,j] = epi( sum_r G[b,r] * W[r,j] )                W row-major [R,J], backward w.r.t. the input
//          thread = output column j (coalesced W rows), 8 warps split the reduction, fixed-order smem combine
// Both sum in a fixed order (deterministic across runs).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float epi_apply(const SgemmEpi &epi, float v, int n, int64_t off) {
    switch (epi.kind) {
        case EPI_BIAS: v += epi.bias[n]; break;
        case EPI_BIAS_RELU: v = fmaxf(v + epi.bias[n], 0.f); break;
        case EPI_BIAS_RELU_MASK:
            v = fmaxf(v + epi.bias[n], 0.f);
            if (epi.mask) v *= epi.mask[off];
            break;
        case EPI_RELUGRAD_MASK:
            if (epi.mask) v *= epi.mask[off];
            v = (epi.act[off] > 0.f) ? v : 0.f;
            break;
        default: break;
    }
    return v;
}

// after the call lane l holds the sum over all lanes of their v[l]
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = upper ? v[i] : v[i + off];
            const float keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

constexpr int RD_WARPS = 8;

// RD_COLS output columns per warp: 4 when there are many row blocks (re-use of the X loads), 1 for a single 32-row
// batch so that a 768-column layer still spreads over 96 CTAs
template <int RD_COLS>
__global__ void __launch_bounds__(RD_WARPS * 32)
rowdot_kernel(const float *__restrict__ X, const float *__restrict__ W, float *__restrict__ Y, int M, int N, int K,
              SgemmEpi epi) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = (blockIdx.x * RD_WARPS + warp) * RD_COLS;
    const int b0 = blockIdx.y * 32;
    if (n0 >= N) return;
    float acc[RD_COLS][32];
#pragma unroll
    for (int c = 0; c < RD_COLS; ++c)
#pragma unroll
        for (int b = 0; b < 32; ++b) acc[c][b] = 0.f;
    const int rows = min(32, M - b0);
#pragma unroll 2
    for (int k = lane; k < K; k += 32) {
        float w[RD_COLS];
#pragma unroll
        for (int c = 0; c < RD_COLS; ++c) w[c] = (n0 + c < N) ? __ldg(W + static_cast<int64_t>(n0 + c) * K + k) : 0.f;
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const float x = (b < rows) ? __ldg(X + static_cast<int64_t>(b0 + b) * K + k) : 0.f;
#pragma unroll
            for (int c = 0; c < RD_COLS; ++c) acc[c][b] = fmaf(x, w[c], acc[c][b]);
        }
    }
#pragma unroll
    for (int c = 0; c < RD_COLS; ++c) {
        const float s = warp_transpose_reduce(acc[c], lane);     // lane = batch row
        const int n = n0 + c, m = b0 + lane;
        if (n < N && lane < rows) {
            const int64_t off = static_cast<int64_t>(m) * N + n;
            Y[off] = epi_apply(epi, s, n, off);
        }
    }
}

constexpr int CA_GROUPS = 8;    // reduction split

__global__ void __launch_bounds__(CA_GROUPS * 32)
colacc_kernel(const float *__restrict__ G, const float *__restrict__ W, float *__restrict__ Z, int M, int R, int J,
              SgemmEpi epi) {
    __shared__ float part[CA_GROUPS][32][33];
    const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 32 + lane;
    const int b0 = blockIdx.y * 32;
    const int rows = min(32, M - b0);
    float acc[32];
#pragma unroll
    for (int b = 0; b < 32; ++b) acc[b] = 0.f;
    for (int r = g; r < R; r += CA_GROUPS) {
        const float w = (j < J) ? __ldg(W + static_cast<int64_t>(r) * J + j) : 0.f;
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const float gv = (b < rows) ? __ldg(G + static_cast<int64_t>(b0 + b) * R + r) : 0.f;   // warp-uniform address
            acc[b] = fmaf(gv, w, acc[b]);
        }
    }
#pragma unroll
    for (int b = 0; b < 32; ++b) part[g][b][lane] = acc[b];
    __syncthreads();
    // warp g finishes batch rows g, g+8, ...: sum over the groups in index order
    for (int b = g; b < rows; b += CA_GROUPS) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < CA_GROUPS; ++q) s += part[q][b][lane];
        if (j < J) {
            const int64_t off = static_cast<int64_t>(b0 + b) * J + j;
            Z[off] = epi_apply(epi, s, j, off);
        }
    }
}

static int rowdot(const float *X, const float *W, float *Y, int M, int N, int K, SgemmEpi epi, cudaStream_t s) {
    if (M <= 0 || N <= 0) return AC_OK;
    if (M <= 64) {
        dim3 grid((N + RD_WARPS - 1) / RD_WARPS, (M + 31) / 32);
        rowdot_kernel<1><<<grid, RD_WARPS * 32, 0, s>>>(X, W, Y, M, N, K, epi);
    } else {
        dim3 grid((N + 4 * RD_WARPS - 1) / (4 * RD_WARPS), (M + 31) / 32);
        rowdot_kernel<4><<<grid, RD_WARPS * 32, 0, s>>>(X, W, Y, M, N, K, epi);
    }
    AC_LAUNCH_CHECK();
    return AC_OK;
}
static int colacc(const float *G, const float *W, float *Z, int M, int R, int J, SgemmEpi epi, cudaStream_t s) {
    if (M <= 0 || J <= 0) return AC_OK;
    dim3 grid((J + 31) / 32, (M + 31) / 32);
    colacc_kernel<<<grid, CA_GROUPS * 32, 0, s>>>(G, W, Z, M, R, J, epi);
    AC_LAUNCH_CHECK();
    return AC_OK;
}

// y[m,n] = act(X W^T + b):  A = X (sam = K, sak = 1), B(k,n) = W[n*K + k] (sbk = 1, sbn = K)
static int linear_fwd(const float *X, const float *W, const float *b, float *Y, int M, int N, int K, int kind,
                      const float *mask, cudaStream_t s) {
    SgemmEpi e{kind, b, mask, nullptr};
    return rowdot(X, W, Y, M, N, K, e, s);
}

// ------------------------------------------------------------------------------------------------
// row-wise output activations / losses
// ------------------------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(const float *__restrict__ z, int B, int C, float *__restrict__ out, int act) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B) return;
    const float *zr = z + static_cast<int64_t>(row) * C;
    float *orow = out + static_cast<int64_t>(row) * C;
    if (act == AC_ACT_SIGMOID) {
        for (int j = lane; j < C; j += 32) orow[j] = 1.f / (1.f + expf(-zr[j]));
        return;
    }
    float mx = -CUDART_INF_F;
    for (int j = lane; j < C; j += 32) mx = fmaxf(mx, zr[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < C; j += 32) sum += expf(zr[j] - mx);
    sum = warp_sum(sum);
    for (int j = lane; j < C; j += 32) orow[j] = expf(zr[j] - mx) / sum;
}

// loss + dz.  CE: loss_b = -(z_y - lse); dz = (softmax - onehot)/B.
// BCE on sigmoid: loss = mean over B*C; dz = (s - y)/(B*C).  One warp per row; per-row losses are
// written to row_loss[B] and reduced in index order by reduce_loss_kernel (deterministic).
__global__ void loss_grad_kernel(const float *__restrict__ z, const void *__restrict__ targets, int B, int C,
                                 int loss_kind, float *__restrict__ dz, float *__restrict__ row_loss) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B) return;
    const float *zr = z + static_cast<int64_t>(row) * C;
    float *dr = dz + static_cast<int64_t>(row) * C;
    if (loss_kind == AC_LOSS_CE) {
        const int64_t y = static_cast<const int64_t *>(targets)[row];
        float mx = -CUDART_INF_F;
        for (int j = lane; j < C; j += 32) mx = fmaxf(mx, zr[j]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < C; j += 32) sum += expf(zr[j] - mx);
        sum = warp_sum(sum);
        const float lse = mx + logf(sum);
        const float invB = 1.f / static_cast<float>(B);
        for (int j = lane; j < C; j += 32) {
            const float p = expf(zr[j] - mx) / sum;
            dr[j] = (p - (j == y ? 1.f : 0.f)) * invB;
        }
        if (lane == 0) row_loss[row] = (y >= 0 && y < C) ? (lse - zr[y]) : 0.f;
    } else {
        const float *yr = static_cast<const float *>(targets) + static_cast<int64_t>(row) * C;
        const float inv = 1.f / (static_cast<float>(B) * static_cast<float>(C));
        float l = 0.f;
        for (int j = lane; j < C; j += 32) {
            const float s = 1.f / (1.f + expf(-zr[j]));
            const float y = yr[j];
            // nn.BCELoss clamps log at -100
            l -= y * fmaxf(logf(s), -100.f) + (1.f - y) * fmaxf(logf(1.f - s), -100.f);
            // gradient through BCELoss(sigmoid(z)) = (s - y) / (B*C)
            dr[j] = (s - y) * inv;
        }
        l = warp_sum(l);
        if (lane == 0) row_loss[row] = l / static_cast<float>(C);
    }
}

__global__ void reduce_loss_kernel(const float *__restrict__ row_loss, int B, float *__restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < B; ++i) s += row_loss[i];
        out[0] = s / static_cast<float>(B);
    }
}

// column sums of dY[B,N] -> gb[N] (bias gradient), rows added in index order
__global__ void colsum_kernel(const float *__restrict__ dY, int B, int N, float *__restrict__ gb) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dY[static_cast<int64_t>(b) * N + n];
    gb[n] = s;
}

// Philox-free counter hash dropout mask (used only when the caller does not inject masks)
__device__ __forceinline__ uint32_t mix32(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return static_cast<uint32_t>(x);
}
__global__ void dropout_mask_kernel(float *__restrict__ mask, int64_t n, float p, uint64_t seed, uint64_t stream_id) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = mix32(seed * 0x9E3779B97F4A7C15ULL + stream_id * 0xD1B54A32D192ED03ULL + static_cast<uint64_t>(i));
    const float u = (r >> 8) * (1.0f / 16777216.0f);
    mask[i] = (u < p) ? 0.f : 1.f / (1.f - p);
}

// ------------------------------------------------------------------------------------------------
// flat parameter views: 6 tensors {W0,b0,W1,b1,W2,b2}
// ------------------------------------------------------------------------------------------------
struct Flat6 {
    float *p[6];
    int64_t n[6];
};
static Flat6 flat_of(const ac_head_params *h) {
    Flat6 f;
    f.p[0] = h->W0; f.n[0] = static_cast<int64_t>(h->H0) * h->D;
    f.p[1] = h->b0; f.n[1] = h->H0;
    f.p[2] = h->W1; f.n[2] = static_cast<int64_t>(h->H1) * h->H0;
    f.p[3] = h->b1; f.n[3] = h->H1;
    f.p[4] = h->W2; f.n[4] = static_cast<int64_t>(h->C) * h->H1;
    f.p[5] = h->b2; f.n[5] = h->C;
    return f;
}

// EWC: g += 2*lam*invB * F * (theta - theta*) over the first `n_lim` elements of each tensor;
// penalty partial sums (F*(theta-theta*)^2) go to partial[blockIdx] for a deterministic second stage.
__global__ void ewc_grad_penalty_kernel(Flat6 theta, Flat6 fisher, Flat6 star, Flat6 grad, Flat6 lim, float scale2,
                                        float *__restrict__ partial, int add_grad) {
    __shared__ float red[256];
    float local = 0.f;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int t = 0; t < 6; ++t) {
        const int64_t n = lim.n[t];
        for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
            const float diff = theta.p[t][i] - star.p[t][i];
            const float f = fisher.p[t][i];
            local += f * diff * diff;
            if (add_grad) grad.p[t][i] += scale2 * f * diff;
        }
    }
    red[threadIdx.x] = local;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// sum of squares of all gradients -> partial[blockIdx]
__global__ void sumsq_kernel(Flat6 g, float *__restrict__ partial) {
    __shared__ float red[256];
    float local = 0.f;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int t = 0; t < 6; ++t)
        for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < g.n[t]; i += stride) {
            const float v = g.p[t][i];
            local = fmaf(v, v, local);
        }
    red[threadIdx.x] = local;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// finalize: out[0] = scale * sum(partial) (ewc penalty) or sqrt(sum) (grad norm)
__global__ void finalize_kernel(const float *__restrict__ partial, int n, float scale, int take_sqrt,
                                float *__restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < n; ++i) s += partial[i];
        out[0] = take_sqrt ? sqrtf(s) : scale * s;
    }
}

// clip_grad_norm_ + AdamW (decoupled weight decay), one pass over all parameters
__global__ void adamw_kernel(Flat6 theta, Flat6 grad, Flat6 m, Flat6 v, const float *__restrict__ gnorm, float lr,
                             float b1, float b2, float eps, float wd, float max_norm, float bc1, float bc2_sqrt) {
    const float total = gnorm[0];
    float coef = max_norm / (total + 1e-6f);
    coef = coef < 1.f ? coef : 1.f;
    if (!(max_norm > 0.f)) coef = 1.f;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int t = 0; t < 6; ++t)
        for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < theta.n[t]; i += stride) {
            const float g = grad.p[t][i] * coef;
            float p = theta.p[t][i];
            p = p * (1.f - lr * wd);
            const float mi = m.p[t][i] * b1 + g * (1.f - b1);
            const float vi = v.p[t][i] * b2 + g * g * (1.f - b2);
            const float denom = sqrtf(vi) / bc2_sqrt + eps;
            p = p - (lr / bc1) * (mi / denom);
            theta.p[t][i] = p;
            m.p[t][i] = mi;
            v.p[t][i] = vi;
        }
}

__global__ void fisher_accum_kernel(Flat6 grad, Flat6 fisher, float inv_n) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int t = 0; t < 6; ++t)
        for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < grad.n[t]; i += stride) {
            const float g = grad.p[t][i];
            fisher.p[t][i] += g * g * inv_n;
        }
}

// ------------------------------------------------------------------------------------------------
// workspace layout for a train / grad step
// ------------------------------------------------------------------------------------------------
struct TrainWs {
    float *h0, *h1, *z, *dz, *dh1, *dh0, *mask0, *mask1, *row_loss, *partial, *gnorm;
    ac_head_params g;   // gradients
    size_t bytes;
};
constexpr int RED_BLOCKS = 64;

static size_t carve(TrainWs &w, void *base, int B, const ac_head_params *p) {
    uint8_t *ptr = static_cast<uint8_t *>(base);
    size_t off = 0;
    auto take = [&](size_t floats) {
        float *r = base ? reinterpret_cast<float *>(ptr + off) : nullptr;
        off += align_up(floats * sizeof(float), 256);
        return r;
    };
    w.h0 = take(static_cast<size_t>(B) * p->H0);
    w.h1 = take(static_cast<size_t>(B) * p->H1);
    w.z = take(static_cast<size_t>(B) * p->C);
    w.dz = take(static_cast<size_t>(B) * p->C);
    w.dh1 = take(static_cast<size_t>(B) * p->H1);
    w.dh0 = take(static_cast<size_t>(B) * p->H0);
    w.mask0 = take(static_cast<size_t>(B) * p->H0);
    w.mask1 = take(static_cast<size_t>(B) * p->H1);
    w.row_loss = take(B);
    w.partial = take(RED_BLOCKS);
    w.gnorm = take(4);
    w.g = *p;
    w.g.W0 = take(static_cast<size_t>(p->H0) * p->D);
    w.g.b0 = take(p->H0);
    w.g.W1 = take(static_cast<size_t>(p->H1) * p->H0);
    w.g.b1 = take(p->H1);
    w.g.W2 = take(static_cast<size_t>(p->C) * p->H1);
    w.g.b2 = take(p->C);
    w.bytes = off;
    return off;
}

static int check_params(const ac_head_params *p, const char *who) {
    AC_REQUIRE(p && p->D > 0 && p->H0 > 0 && p->H1 > 0 && p->C > 0, "%s: bad head dims", who);
    AC_REQUIRE(p->W0 && p->b0 && p->W1 && p->b1 && p->W2 && p->b2, "%s: null head parameter", who);
    return AC_OK;
}

// forward (optionally train mode with masks) + loss + backward into w.g
static int fwd_bwd(const float *X, const void *targets, int B, const ac_head_params *p, int loss_kind,
                   const float *mask0, const float *mask1, TrainWs &w, float *out_loss, cudaStream_t s) {
    int rc;
    const int D = p->D, H0 = p->H0, H1 = p->H1, C = p->C;
    if ((rc = linear_fwd(X, p->W0, p->b0, w.h0, B, H0, D, EPI_BIAS_RELU_MASK, mask0, s))) return rc;
    if ((rc = linear_fwd(w.h0, p->W1, p->b1, w.h1, B, H1, H0, EPI_BIAS_RELU_MASK, mask1, s))) return rc;
    if ((rc = linear_fwd(w.h1, p->W2, p->b2, w.z, B, C, H1, EPI_BIAS, nullptr, s))) return rc;
    const int wpb = 4;
    loss_grad_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, s>>>(w.z, targets, B, C, loss_kind, w.dz, w.row_loss);
    AC_LAUNCH_CHECK();
    reduce_loss_kernel<<<1, 32, 0, s>>>(w.row_loss, B, out_loss);
    AC_LAUNCH_CHECK();
    SgemmEpi none{EPI_NONE, nullptr, nullptr, nullptr};
    // gW2[C,H1] = dz^T h1 : A(m=c,k=b) = dz[b*C+c], B(k=b,n=j) = h1[b*H1+j]
    if ((rc = sgemm(w.dz, 1, C, w.h1, H1, 1, w.g.W2, H1, C, H1, B, none, s))) return rc;
    colsum_kernel<<<(C + 127) / 128, 128, 0, s>>>(w.dz, B, C, w.g.b2);
    AC_LAUNCH_CHECK();
    // dh1[B,H1] = dz W2, then relu-grad + mask -> da1
    SgemmEpi rg1{EPI_RELUGRAD_MASK, nullptr, mask1, w.h1};
    if ((rc = colacc(w.dz, p->W2, w.dh1, B, C, H1, rg1, s))) return rc;
    if ((rc = sgemm(w.dh1, 1, H1, w.h0, H0, 1, w.g.W1, H0, H1, H0, B, none, s))) return rc;
    colsum_kernel<<<(H1 + 127) / 128, 128, 0, s>>>(w.dh1, B, H1, w.g.b1);
    AC_LAUNCH_CHECK();
    SgemmEpi rg0{EPI_RELUGRAD_MASK, nullptr, mask0, w.h0};
    if ((rc = colacc(w.dh1, p->W1, w.dh0, B, H1, H0, rg0, s))) return rc;
    if ((rc = sgemm(w.dh0, 1, H0, X, D, 1, w.g.W0, D, H0, D, B, none, s))) return rc;
    colsum_kernel<<<(H0 + 127) / 128, 128, 0, s>>>(w.dh0, B, H0, w.g.b0);
    AC_LAUNCH_CHECK();
    return AC_OK;
}

}  // namespace ac

using namespace ac;

extern "C" int ac_head_forward(const float *X, int B, const ac_head_params *p, int act, float *out, float *scratch,
                               size_t scratch_floats, ac_stream_t stream) {
    int rc = check_params(p, "ac_head_forward");
    if (rc) return rc;
    AC_REQUIRE(X && out && B >= 0, "ac_head_forward: bad arguments");
    if (B == 0) return AC_OK;
    const size_t need = static_cast<size_t>(B) * (p->H0 + p->H1);
    if (!scratch || scratch_floats < need) {
        set_error("ac_head_forward: scratch needs %zu floats", need);
        return AC_E_WORKSPACE;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float *h0 = scratch, *h1 = scratch + static_cast<size_t>(B) * p->H0;
    if ((rc = linear_fwd(X, p->W0, p->b0, h0, B, p->H0, p->D, EPI_BIAS_RELU, nullptr, s))) return rc;
    if ((rc = linear_fwd(h0, p->W1, p->b1, h1, B, p->H1, p->H0, EPI_BIAS_RELU, nullptr, s))) return rc;
    if ((rc = linear_fwd(h1, p->W2, p->b2, out, B, p->C, p->H1, EPI_BIAS, nullptr, s))) return rc;
    if (act == AC_ACT_SOFTMAX || act == AC_ACT_SIGMOID) {
        const int wpb = 4;
        softmax_rows_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, s>>>(out, B, p->C, out, act);
        AC_LAUNCH_CHECK();
    }
    return AC_OK;
}

extern "C" int ac_head_train_workspace_bytes(int B, const ac_head_params *p, size_t *bytes) {
    AC_REQUIRE(p && bytes && B > 0, "ac_head_train_workspace_bytes: bad arguments");
    TrainWs w;
    *bytes = carve(w, nullptr, B, p) + 256;
    return AC_OK;
}

extern "C" int ac_head_grad(const float *X, const void *targets, int B, const ac_head_params *p, int loss_kind,
                            ac_head_params *grad_out, ac_head_params *fisher_accum, float inv_n_batches,
                            float *out_loss, void *workspace, size_t workspace_bytes, ac_stream_t stream) {
    int rc = check_params(p, "ac_head_grad");
    if (rc) return rc;
    AC_REQUIRE(X && targets && B > 0 && out_loss && workspace, "ac_head_grad: bad arguments");
    TrainWs w;
    const size_t need = carve(w, workspace, B, p);
    if (need > workspace_bytes) { set_error("ac_head_grad: workspace needs %zu bytes", need); return AC_E_WORKSPACE; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if ((rc = fwd_bwd(X, targets, B, p, loss_kind, nullptr, nullptr, w, out_loss, s))) return rc;
    Flat6 g = flat_of(&w.g);
    if (fisher_accum) {
        Flat6 f = flat_of(fisher_accum);
        fisher_accum_kernel<<<RED_BLOCKS, 256, 0, s>>>(g, f, inv_n_batches);
        AC_LAUNCH_CHECK();
    }
    if (grad_out) {
        Flat6 o = flat_of(grad_out);
        for (int t = 0; t < 6; ++t)
            AC_CUDA(cudaMemcpyAsync(o.p[t], g.p[t], g.n[t] * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    return AC_OK;
}

static Flat6 ewc_limits(const ac_head_params *p, int C_old) {
    // the head may have grown since theta* was taken: only the first C_old output rows are penalised
    ac_head_params q = *p;
    Flat6 f = flat_of(&q);
    if (C_old > 0 && C_old < p->C) {
        f.n[4] = static_cast<int64_t>(C_old) * p->H1;
        f.n[5] = C_old;
    }
    return f;
}

extern "C" int ac_ewc_penalty(const ac_head_params *p, const ac_head_params *fisher, const ac_head_params *star,
                              float lambda, float inv_batch, int C_old, float *out, ac_stream_t stream) {
    int rc = check_params(p, "ac_ewc_penalty");
    if (rc) return rc;
    AC_REQUIRE(fisher && star && out, "ac_ewc_penalty: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float *partial = nullptr;
    AC_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&partial), RED_BLOCKS * sizeof(float), s));
    Flat6 lim = ewc_limits(p, C_old);
    ewc_grad_penalty_kernel<<<RED_BLOCKS, 256, 0, s>>>(flat_of(p), flat_of(fisher), flat_of(star), flat_of(p), lim, 0.f,
                                                       partial, 0);
    AC_LAUNCH_CHECK();
    finalize_kernel<<<1, 32, 0, s>>>(partial, RED_BLOCKS, lambda * inv_batch, 0, out);
    AC_LAUNCH_CHECK();
    AC_CUDA(cudaFreeAsync(partial, s));
    return AC_OK;
}

static int train_step_impl(const float *X, const void *targets, int B, ac_head_params *p, ac_head_params *m,
                           ac_head_params *v, const ac_train_cfg *cfg, int step, float *out_stats, void *workspace,
                           size_t workspace_bytes, cudaStream_t s);

extern "C" int ac_head_train_step(const float *X, const void *targets, int B, ac_head_params *p, ac_head_params *m,
                                  ac_head_params *v, const ac_train_cfg *cfg, float *out_stats, void *workspace,
                                  size_t workspace_bytes, ac_stream_t stream) {
    int rc = check_params(p, "ac_head_train_step");
    if (rc) return rc;
    AC_REQUIRE(X && targets && B > 0 && m && v && cfg && out_stats && workspace, "ac_head_train_step: bad arguments");
    AC_REQUIRE(cfg->step >= 1, "ac_head_train_step: step must be >= 1");
    return train_step_impl(X, targets, B, p, m, v, cfg, cfg->step, out_stats, workspace, workspace_bytes,
                           static_cast<cudaStream_t>(stream));
}

static int train_step_impl(const float *X, const void *targets, int B, ac_head_params *p, ac_head_params *m,
                           ac_head_params *v, const ac_train_cfg *cfg, int step, float *out_stats, void *workspace,
                           size_t workspace_bytes, cudaStream_t s) {
    int rc;
    TrainWs w;
    const size_t need = carve(w, workspace, B, p);
    if (need > workspace_bytes) { set_error("ac_head_train_step: workspace needs %zu bytes", need); return AC_E_WORKSPACE; }

    const float *mask0 = cfg->mask0, *mask1 = cfg->mask1;
    if (cfg->dropout_p > 0.f && (!mask0 || !mask1)) {
        const int64_t n0 = static_cast<int64_t>(B) * p->H0, n1 = static_cast<int64_t>(B) * p->H1;
        dropout_mask_kernel<<<static_cast<unsigned>((n0 + 255) / 256), 256, 0, s>>>(w.mask0, n0, cfg->dropout_p, cfg->seed,
                                                                                  2ull * step);
        dropout_mask_kernel<<<static_cast<unsigned>((n1 + 255) / 256), 256, 0, s>>>(w.mask1, n1, cfg->dropout_p, cfg->seed,
                                                                                  2ull * step + 1);
        AC_LAUNCH_CHECK();
        mask0 = w.mask0;
        mask1 = w.mask1;
    } else if (!(cfg->dropout_p > 0.f)) {
        mask0 = mask1 = nullptr;
    }
    if ((rc = fwd_bwd(X, targets, B, p, cfg->loss_kind, mask0, mask1, w, out_stats + 0, s))) return rc;

    Flat6 g = flat_of(&w.g);
    if (cfg->ewc_fisher && cfg->ewc_star) {
        Flat6 lim = ewc_limits(p, cfg->ewc_C_old);
        const float scale = cfg->ewc_lambda / static_cast<float>(B);
        ewc_grad_penalty_kernel<<<RED_BLOCKS, 256, 0, s>>>(flat_of(p), flat_of(cfg->ewc_fisher), flat_of(cfg->ewc_star), g,
                                                           lim, 2.f * scale, w.partial, 1);
        AC_LAUNCH_CHECK();
        finalize_kernel<<<1, 32, 0, s>>>(w.partial, RED_BLOCKS, scale, 0, out_stats + 1);
        AC_LAUNCH_CHECK();
    } else {
        AC_CUDA(cudaMemsetAsync(out_stats + 1, 0, sizeof(float), s));
    }
    sumsq_kernel<<<RED_BLOCKS, 256, 0, s>>>(g, w.partial);
    AC_LAUNCH_CHECK();
    finalize_kernel<<<1, 32, 0, s>>>(w.partial, RED_BLOCKS, 1.f, 1, out_stats + 2);
    AC_LAUNCH_CHECK();
    const float bc1 = 1.f - powf(cfg->beta1, static_cast<float>(step));
    const float bc2 = 1.f - powf(cfg->beta2, static_cast<float>(step));
    adamw_kernel<<<RED_BLOCKS * 2, 256, 0, s>>>(flat_of(p), g, flat_of(m), flat_of(v), out_stats + 2, cfg->lr, cfg->beta1,
                                                cfg->beta2, cfg->eps, cfg->weight_decay, cfg->max_norm, bc1, sqrtf(bc2));
    AC_LAUNCH_CHECK();
    return AC_OK;
}


// ------------------------------------------------------------------------------------------------
// fused epoch (opt-in, option "head_fused"): ONE cooperative persistent kernel runs every optimizer step of an epoch.
//
// The step above is ~21 dependent launches of latency-bound kernels (measured 2.9 k steps/s at batch 32, i.e. ~350 us per
// step for 171 MFLOP: profiles/r01_bench_add_examples_v3.json).  Here the same kernels become PHASES of one kernel,
// separated by grid barriers (cooperative_groups grid.sync), and every CTA walks the phase's original grid as "virtual
// blocks".  The bodies below are copies of the kernels above with (a) virtual block indices and (b) plain coherent loads
// instead of __ldg / __restrict__ (weights, activations and gradients are rewritten by other CTAs inside this kernel, so
// the non-coherent path is not allowed).  Operation order inside every virtual block is unchanged, so an epoch through
// this kernel is expected to give the same bits as the launch-per-kernel path (tests/test_gpu_zzz_variants.py compares them).
//
//   per step:  gather+masks | h0 | h1 | z | loss,dz | gW2,gb2,dh1,loss | gW1,gb1,dh0 | gW0,gb0 | [EWC] | sumsq | AdamW
//              (9 grid barriers, 10 with EWC; AdamW of step t overlaps the gather of step t+1)
// Status: written after the round-1 GPU budget was spent; compiles for sm_100a, NOT yet run on hardware.
// ------------------------------------------------------------------------------------------------
#include <cooperative_groups.h>
namespace ac {
namespace fused {
namespace cg = cooperative_groups;

constexpr int FT = 256;                              // threads per CTA
constexpr int F_SMEM_FLOATS = CA_GROUPS * 32 * 33;   // colacc's partials are the largest user (33.8 KB)

struct EpochArgs {
    const float *X;            // [n, D]
    const void *targets;       // int64[n] or float[n, C]
    const int64_t *perm;       // [n]
    int n, batch, first_step;
    ac_head_params p, m, v;    // parameters and AdamW moments (device pointers)
    TrainWs w;                 // activations, gradients, partials (device pointers)
    float *xb;                 // [batch, D] gathered rows
    void *yb;                  // gathered targets
    float *stats;              // [3] task loss, ewc penalty, grad norm of the current step
    float *loss_accum;         // += task loss + ewc penalty per step
    float *partial_ewc;        // [RED_BLOCKS]
    const float2 *bias_corr;   // [steps] (1 - beta1^t, sqrt(1 - beta2^t)) computed on the host like the per-step path
    float lr, beta1, beta2, eps, weight_decay, max_norm, dropout_p;
    int loss_kind;
    uint64_t seed;
    int use_ewc, ewc_C_old;
    float ewc_lambda;
    ac_head_params fisher, star;
};

__device__ __forceinline__ void sgemm_vb(const float *A, int64_t sam, int64_t sak, const float *B, int64_t sbk, int64_t sbn,
                                         float *C, int64_t ldc, int M, int N, int K, int vbx, int vby, float *smem) {
    float(*sA)[SG_BM + 1] = reinterpret_cast<float(*)[SG_BM + 1]>(smem);
    float(*sB)[SG_BN + 1] = reinterpret_cast<float(*)[SG_BN + 1]>(smem + SG_BK * (SG_BM + 1));
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = vby * SG_BM, n0 = vbx * SG_BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += SG_BK) {
        for (int e = threadIdx.x; e < SG_BM * SG_BK; e += FT) {
            int mm, kk;
            if (sak == 1) { kk = e % SG_BK; mm = e / SG_BK; } else { mm = e % SG_BM; kk = e / SG_BM; }
            const int m = m0 + mm, k = k0 + kk;
            sA[kk][mm] = (m < M && k < K) ? A[m * sam + k * sak] : 0.f;
        }
        for (int e = threadIdx.x; e < SG_BN * SG_BK; e += FT) {
            int nn, kk;
            if (sbk == 1) { kk = e % SG_BK; nn = e / SG_BK; } else { nn = e % SG_BN; kk = e / SG_BN; }
            const int n = n0 + nn, k = k0 + kk;
            sB[kk][nn] = (n < N && k < K) ? B[k * sbk + n * sbn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SG_BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty + 16 * i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx + 16 * j;
            if (n >= N) continue;
            C[static_cast<int64_t>(m) * ldc + n] = acc[i][j];      // the weight-gradient GEMMs use EPI_NONE
        }
    }
}

// rowdot_kernel<1>: warp = one output column x 32 batch rows
__device__ __forceinline__ void rowdot_vb(const float *X, const float *W, float *Y, int M, int N, int K, const SgemmEpi &epi,
                                          int vbx, int vby) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = vbx * RD_WARPS + warp;
    const int b0 = vby * 32;
    if (n0 >= N) return;
    float acc[32];
#pragma unroll
    for (int b = 0; b < 32; ++b) acc[b] = 0.f;
    const int rows = min(32, M - b0);
#pragma unroll 2
    for (int k = lane; k < K; k += 32) {
        const float w = W[static_cast<int64_t>(n0) * K + k];
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const float x = (b < rows) ? X[static_cast<int64_t>(b0 + b) * K + k] : 0.f;
            acc[b] = fmaf(x, w, acc[b]);
        }
    }
    const float sum = warp_transpose_reduce(acc, lane);     // lane = batch row
    const int m = b0 + lane;
    if (lane < rows) {
        const int64_t off = static_cast<int64_t>(m) * N + n0;
        Y[off] = epi_apply(epi, sum, n0, off);
    }
}

__device__ __forceinline__ void colacc_vb(const float *G, const float *W, float *Z, int M, int R, int J, const SgemmEpi &epi,
                                          int vbx, int vby, float *smem) {
    float(*part)[32][33] = reinterpret_cast<float(*)[32][33]>(smem);
    const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = vbx * 32 + lane;
    const int b0 = vby * 32;
    const int rows = min(32, M - b0);
    float acc[32];
#pragma unroll
    for (int b = 0; b < 32; ++b) acc[b] = 0.f;
    for (int r = g; r < R; r += CA_GROUPS) {
        const float w = (j < J) ? W[static_cast<int64_t>(r) * J + j] : 0.f;
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const float gv = (b < rows) ? G[static_cast<int64_t>(b0 + b) * R + r] : 0.f;
            acc[b] = fmaf(gv, w, acc[b]);
        }
    }
#pragma unroll
    for (int b = 0; b < 32; ++b) part[g][b][lane] = acc[b];
    __syncthreads();
    for (int b = g; b < rows; b += CA_GROUPS) {
        float sum = 0.f;
#pragma unroll
        for (int q = 0; q < CA_GROUPS; ++q) sum += part[q][b][lane];
        if (j < J) {
            const int64_t off = static_cast<int64_t>(b0 + b) * J + j;
            Z[off] = epi_apply(epi, sum, j, off);
        }
    }
    __syncthreads();      // the partials are reused by the next virtual block of this CTA
}

// loss_grad_kernel for one row (one warp)
__device__ __forceinline__ void loss_grad_row(const float *z, const void *targets, int B, int C, int loss_kind, float *dz,
                                              float *row_loss, int row, int lane) {
    const float *zr = z + static_cast<int64_t>(row) * C;
    float *dr = dz + static_cast<int64_t>(row) * C;
    if (loss_kind == AC_LOSS_CE) {
        const int64_t y = static_cast<const int64_t *>(targets)[row];
        float mx = -CUDART_INF_F;
        for (int j = lane; j < C; j += 32) mx = fmaxf(mx, zr[j]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < C; j += 32) sum += expf(zr[j] - mx);
        sum = warp_sum(sum);
        const float lse = mx + logf(sum);
        const float invB = 1.f / static_cast<float>(B);
        for (int j = lane; j < C; j += 32) {
            const float pr = expf(zr[j] - mx) / sum;
            dr[j] = (pr - (j == y ? 1.f : 0.f)) * invB;
        }
        if (lane == 0) row_loss[row] = (y >= 0 && y < C) ? (lse - zr[y]) : 0.f;
    } else {
        const float *yr = static_cast<const float *>(targets) + static_cast<int64_t>(row) * C;
        const float inv = 1.f / (static_cast<float>(B) * static_cast<float>(C));
        float l = 0.f;
        for (int j = lane; j < C; j += 32) {
            const float sg = 1.f / (1.f + expf(-zr[j]));
            const float y = yr[j];
            l -= y * fmaxf(logf(sg), -100.f) + (1.f - y) * fmaxf(logf(1.f - sg), -100.f);
            dr[j] = (sg - y) * inv;
        }
        l = warp_sum(l);
        if (lane == 0) row_loss[row] = l / static_cast<float>(C);
    }
}

__device__ __forceinline__ void colsum_vb(const float *dY, int B, int N, float *gb, int vb) {
    const int n = vb * FT + threadIdx.x;
    if (n >= N) return;
    float sum = 0.f;
    for (int b = 0; b < B; ++b) sum += dY[static_cast<int64_t>(b) * N + n];
    gb[n] = sum;
}

// block tree reduction of sumsq_kernel / ewc_grad_penalty_kernel (256 threads)
__device__ __forceinline__ float block_reduce_256(float local, float *red) {
    red[threadIdx.x] = local;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
        __syncthreads();
    }
    const float r = red[0];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(FT, 1) head_epoch_kernel(const EpochArgs a) {
    __shared__ float smem[F_SMEM_FLOATS];
    __shared__ float s_bcast[2];
    cg::grid_group grid = cg::this_grid();
    const int G = gridDim.x, cta = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = a.p.D, H0 = a.p.H0, H1 = a.p.H1, C = a.p.C;
    const bool drop = a.dropout_p > 0.f;
    const float *mask0 = drop ? a.w.mask0 : nullptr, *mask1 = drop ? a.w.mask1 : nullptr;
    const Flat6 theta = {{a.p.W0, a.p.b0, a.p.W1, a.p.b1, a.p.W2, a.p.b2},
                         {static_cast<int64_t>(H0) * D, H0, static_cast<int64_t>(H1) * H0, H1, static_cast<int64_t>(C) * H1, C}};
    const Flat6 mom = {{a.m.W0, a.m.b0, a.m.W1, a.m.b1, a.m.W2, a.m.b2}, {0, 0, 0, 0, 0, 0}};
    const Flat6 var = {{a.v.W0, a.v.b0, a.v.W1, a.v.b1, a.v.W2, a.v.b2}, {0, 0, 0, 0, 0, 0}};
    const Flat6 grad = {{a.w.g.W0, a.w.g.b0, a.w.g.W1, a.w.g.b1, a.w.g.W2, a.w.g.b2}, {0, 0, 0, 0, 0, 0}};

    int step = a.first_step;
    for (int off = 0; off < a.n; off += a.batch, ++step) {
        const int nb = (a.n - off < a.batch) ? a.n - off : a.batch;
        const int rb = (nb + 31) / 32;                                   // 32-row blocks of the batch (1 at batch 32)
        // ---- phase 0: gather the batch rows / targets, dropout masks of this step
        for (int r = cta; r < nb; r += G) {
            const int64_t src = a.perm[off + r];
            for (int i = threadIdx.x; i < D; i += FT) a.xb[static_cast<int64_t>(r) * D + i] = a.X[src * D + i];
            if (a.loss_kind == AC_LOSS_CE) {
                if (threadIdx.x == 0) static_cast<int64_t *>(a.yb)[r] = static_cast<const int64_t *>(a.targets)[src];
            } else {
                for (int i = threadIdx.x; i < C; i += FT)
                    static_cast<float *>(a.yb)[static_cast<int64_t>(r) * C + i] = static_cast<const float *>(a.targets)[src * C + i];
            }
        }
        if (drop) {
            const int64_t n0 = static_cast<int64_t>(nb) * H0, n1 = static_cast<int64_t>(nb) * H1;
            for (int64_t i = static_cast<int64_t>(cta) * FT + threadIdx.x; i < n0 + n1; i += static_cast<int64_t>(G) * FT) {
                const bool first = i < n0;
                const int64_t e = first ? i : i - n0;
                const uint32_t r = mix32(a.seed * 0x9E3779B97F4A7C15ULL + (2ull * step + (first ? 0 : 1)) * 0xD1B54A32D192ED03ULL +
                                         static_cast<uint64_t>(e));
                const float u = (r >> 8) * (1.0f / 16777216.0f);
                (first ? a.w.mask0 : a.w.mask1)[e] = (u < a.dropout_p) ? 0.f : 1.f / (1.f - a.dropout_p);
            }
        }
        grid.sync();
        // ---- phase 1..3: forward
        {
            const SgemmEpi e0{EPI_BIAS_RELU_MASK, a.p.b0, mask0, nullptr};
            const int nx = (H0 + RD_WARPS - 1) / RD_WARPS;
            for (int vb = cta; vb < nx * rb; vb += G) rowdot_vb(a.xb, a.p.W0, a.w.h0, nb, H0, D, e0, vb % nx, vb / nx);
        }
        grid.sync();
        {
            const SgemmEpi e1{EPI_BIAS_RELU_MASK, a.p.b1, mask1, nullptr};
            const int nx = (H1 + RD_WARPS - 1) / RD_WARPS;
            for (int vb = cta; vb < nx * rb; vb += G) rowdot_vb(a.w.h0, a.p.W1, a.w.h1, nb, H1, H0, e1, vb % nx, vb / nx);
        }
        grid.sync();
        {
            const SgemmEpi e2{EPI_BIAS, a.p.b2, nullptr, nullptr};
            const int nx = (C + RD_WARPS - 1) / RD_WARPS;
            for (int vb = cta; vb < nx * rb; vb += G) rowdot_vb(a.w.h1, a.p.W2, a.w.z, nb, C, H1, e2, vb % nx, vb / nx);
        }
        grid.sync();
        // ---- phase 4: loss + dz, one warp per batch row
        for (int row = cta * (FT / 32) + warp; row < nb; row += G * (FT / 32))
            loss_grad_row(a.w.z, a.yb, nb, C, a.loss_kind, a.w.dz, a.w.row_loss, row, lane);
        grid.sync();
        // ---- phase 5: task loss; gW2 = dz^T h1, gb2, dh1 = (dz W2) * relu' * mask1
        {
            const int sx = (H1 + SG_BN - 1) / SG_BN, sy = (C + SG_BM - 1) / SG_BM;     // sgemm grid
            const int cs = (C + FT - 1) / FT;                                          // colsum blocks
            const int ax = (H1 + 31) / 32;                                             // colacc grid (x), rb in y
            const int total = sx * sy + cs + ax * rb + 1;
            const SgemmEpi rg1{EPI_RELUGRAD_MASK, nullptr, mask1, a.w.h1};
            for (int it = cta; it < total; it += G) {
                int t = it;
                if (t < sx * sy) { sgemm_vb(a.w.dz, 1, C, a.w.h1, H1, 1, a.w.g.W2, H1, C, H1, nb, t % sx, t / sx, smem); continue; }
                t -= sx * sy;
                if (t < cs) { colsum_vb(a.w.dz, nb, C, a.w.g.b2, t); continue; }
                t -= cs;
                if (t < ax * rb) { colacc_vb(a.w.dz, a.p.W2, a.w.dh1, nb, C, H1, rg1, t % ax, t / ax, smem); continue; }
                if (threadIdx.x == 0) {                       // reduce_loss_kernel
                    float sum = 0.f;
                    for (int i = 0; i < nb; ++i) sum += a.w.row_loss[i];
                    a.stats[0] = sum / static_cast<float>(nb);
                }
            }
        }
        grid.sync();
        // ---- phase 6: gW1 = dh1^T h0, gb1, dh0 = (dh1 W1) * relu' * mask0
        {
            const int sx = (H0 + SG_BN - 1) / SG_BN, sy = (H1 + SG_BM - 1) / SG_BM;
            const int cs = (H1 + FT - 1) / FT;
            const int ax = (H0 + 31) / 32;
            const int total = sx * sy + cs + ax * rb;
            const SgemmEpi rg0{EPI_RELUGRAD_MASK, nullptr, mask0, a.w.h0};
            for (int it = cta; it < total; it += G) {
                int t = it;
                if (t < sx * sy) { sgemm_vb(a.w.dh1, 1, H1, a.w.h0, H0, 1, a.w.g.W1, H0, H1, H0, nb, t % sx, t / sx, smem); continue; }
                t -= sx * sy;
                if (t < cs) { colsum_vb(a.w.dh1, nb, H1, a.w.g.b1, t); continue; }
                t -= cs;
                colacc_vb(a.w.dh1, a.p.W1, a.w.dh0, nb, H1, H0, rg0, t % ax, t / ax, smem);
            }
        }
        grid.sync();
        // ---- phase 7: gW0 = dh0^T x, gb0
        {
            const int sx = (D + SG_BN - 1) / SG_BN, sy = (H0 + SG_BM - 1) / SG_BM;
            const int cs = (H0 + FT - 1) / FT;
            const int total = sx * sy + cs;
            for (int it = cta; it < total; it += G) {
                if (it < sx * sy) sgemm_vb(a.w.dh0, 1, H0, a.xb, D, 1, a.w.g.W0, D, H0, D, nb, it % sx, it / sx, smem);
                else colsum_vb(a.w.dh0, nb, H0, a.w.g.b0, it - sx * sy);
            }
        }
        grid.sync();
        // ---- phase 8 (EWC): g += 2 lambda / B * F (theta - theta*), penalty partials (ewc_grad_penalty_kernel's virtual grid)
        if (a.use_ewc) {
            const Flat6 fis = {{a.fisher.W0, a.fisher.b0, a.fisher.W1, a.fisher.b1, a.fisher.W2, a.fisher.b2}, {0, 0, 0, 0, 0, 0}};
            const Flat6 sta = {{a.star.W0, a.star.b0, a.star.W1, a.star.b1, a.star.W2, a.star.b2}, {0, 0, 0, 0, 0, 0}};
            const bool grown = a.ewc_C_old > 0 && a.ewc_C_old < C;
            const float scale2 = 2.f * (a.ewc_lambda / static_cast<float>(nb));
            for (int vb = cta; vb < RED_BLOCKS; vb += G) {
                float local = 0.f;
                for (int t = 0; t < 6; ++t) {
                    int64_t lim = theta.n[t];
                    if (grown && t == 4) lim = static_cast<int64_t>(a.ewc_C_old) * H1;
                    if (grown && t == 5) lim = a.ewc_C_old;
                    for (int64_t i = static_cast<int64_t>(vb) * FT + threadIdx.x; i < lim; i += static_cast<int64_t>(RED_BLOCKS) * FT) {
                        const float diff = theta.p[t][i] - sta.p[t][i];
                        const float f = fis.p[t][i];
                        local += f * diff * diff;
                        grad.p[t][i] += scale2 * f * diff;
                    }
                }
                const float tot = block_reduce_256(local, smem);
                if (threadIdx.x == 0) a.partial_ewc[vb] = tot;
            }
            grid.sync();
        }
        // ---- phase 9: partial sums of squares of all gradients (sumsq_kernel's virtual grid); EWC penalty value
        for (int vb = cta; vb < RED_BLOCKS; vb += G) {
            float local = 0.f;
            for (int t = 0; t < 6; ++t)
                for (int64_t i = static_cast<int64_t>(vb) * FT + threadIdx.x; i < theta.n[t]; i += static_cast<int64_t>(RED_BLOCKS) * FT) {
                    const float gv = grad.p[t][i];
                    local = fmaf(gv, gv, local);
                }
            const float tot = block_reduce_256(local, smem);
            if (threadIdx.x == 0) a.w.partial[vb] = tot;
        }
        if (cta == G - 1 && threadIdx.x == 0) {
            float pen = 0.f;
            if (a.use_ewc) {
                float sum = 0.f;
                for (int i = 0; i < RED_BLOCKS; ++i) sum += a.partial_ewc[i];
                pen = (a.ewc_lambda / static_cast<float>(nb)) * sum;
            }
            a.stats[1] = pen;
        }
        grid.sync();
        // ---- phase 10: global-norm clip + AdamW (every CTA recomputes the norm from the partials in finalize_kernel's order)
        if (threadIdx.x == 0) {
            float sum = 0.f;
            for (int i = 0; i < RED_BLOCKS; ++i) sum += a.w.partial[i];
            s_bcast[0] = sqrtf(sum);
        }
        __syncthreads();
        {
            const float total = s_bcast[0];
            float coef = a.max_norm / (total + 1e-6f);
            coef = coef < 1.f ? coef : 1.f;
            if (!(a.max_norm > 0.f)) coef = 1.f;
            const float2 bc = a.bias_corr[step - a.first_step];
            const float bc1 = bc.x, bc2_sqrt = bc.y;
            for (int t = 0; t < 6; ++t)
                for (int64_t i = static_cast<int64_t>(cta) * FT + threadIdx.x; i < theta.n[t]; i += static_cast<int64_t>(G) * FT) {
                    const float g = grad.p[t][i] * coef;
                    float pv = theta.p[t][i];
                    pv = pv * (1.f - a.lr * a.weight_decay);
                    const float mi = mom.p[t][i] * a.beta1 + g * (1.f - a.beta1);
                    const float vi = var.p[t][i] * a.beta2 + g * g * (1.f - a.beta2);
                    const float denom = sqrtf(vi) / bc2_sqrt + a.eps;
                    pv = pv - (a.lr / bc1) * (mi / denom);
                    theta.p[t][i] = pv;
                    mom.p[t][i] = mi;
                    var.p[t][i] = vi;
                }
            if (cta == 0 && threadIdx.x == 0) {
                a.stats[2] = total;
                a.loss_accum[0] += a.stats[0] + a.stats[1];
            }
        }
        __syncthreads();    // s_bcast is rewritten next step
        // no grid barrier here: the next step's gather/mask phase touches nothing AdamW reads or writes, and the barrier
        // that closes it orders these parameter writes before the next forward
    }
}

}  // namespace fused
}  // namespace ac

// ------------------------------------------------------------------------------------------------
// one epoch of the training loops (classifier.py:329-353, :1485-1507; multilabel.py:381-399) in a single call:
// batches are gathered on the device from a shuffled index list, every optimizer step is launched from here.
// ------------------------------------------------------------------------------------------------
namespace ac {
__global__ void gather_batch_kernel(const float *__restrict__ X, const void *__restrict__ targets, const int64_t *__restrict__ perm,
                                    int nb, int D, int C, int loss_kind, float *__restrict__ xb, void *__restrict__ yb) {
    const int r = blockIdx.x;
    if (r >= nb) return;
    const int64_t src = perm[r];
    for (int i = threadIdx.x; i < D; i += blockDim.x) xb[static_cast<int64_t>(r) * D + i] = X[src * D + i];
    if (loss_kind == AC_LOSS_CE) {
        if (threadIdx.x == 0) static_cast<int64_t *>(yb)[r] = static_cast<const int64_t *>(targets)[src];
    } else {
        for (int i = threadIdx.x; i < C; i += blockDim.x)
            static_cast<float *>(yb)[static_cast<int64_t>(r) * C + i] = static_cast<const float *>(targets)[src * C + i];
    }
}
__global__ void accum_loss_kernel(const float *__restrict__ stats, float *__restrict__ accum) {
    if (threadIdx.x == 0 && blockIdx.x == 0) accum[0] += stats[0] + stats[1];
}
}  // namespace ac

extern "C" int ac_head_train_epoch_workspace_bytes(int batch, const ac_head_params *p, size_t *bytes) {
    AC_REQUIRE(p && bytes && batch > 0, "ac_head_train_epoch_workspace_bytes: bad arguments");
    TrainWs w;
    const size_t yb = static_cast<size_t>(batch) * (p->C > 2 ? p->C : 2) * sizeof(float);
    *bytes = carve(w, nullptr, batch, p) + align_up(static_cast<size_t>(batch) * p->D * sizeof(float), 256) + align_up(yb, 256) + 1024;
    return AC_OK;
}

extern "C" int ac_head_train_epoch(const float *X, const void *targets, const int64_t *perm, int n, int batch,
                                   ac_head_params *p, ac_head_params *m, ac_head_params *v, const ac_train_cfg *cfg,
                                   float *loss_accum, void *workspace, size_t workspace_bytes, ac_stream_t stream) {
    int rc = check_params(p, "ac_head_train_epoch");
    if (rc) return rc;
    AC_REQUIRE(X && targets && perm && n > 0 && batch > 0 && m && v && cfg && loss_accum && workspace,
               "ac_head_train_epoch: bad arguments");
    AC_REQUIRE(cfg->step >= 1 && !cfg->mask0 && !cfg->mask1, "ac_head_train_epoch: step >= 1 and no injected masks");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TrainWs w;
    const size_t step_bytes = carve(w, nullptr, batch, p);
    const size_t xb_bytes = align_up(static_cast<size_t>(batch) * p->D * sizeof(float), 256);
    const size_t yb_bytes = align_up(static_cast<size_t>(batch) * (p->C > 2 ? p->C : 2) * sizeof(float), 256);
    if (step_bytes + xb_bytes + yb_bytes + 256 > workspace_bytes) {
        set_error("ac_head_train_epoch: workspace needs %zu bytes", step_bytes + xb_bytes + yb_bytes + 256);
        return AC_E_WORKSPACE;
    }
    uint8_t *base = static_cast<uint8_t *>(workspace);
    float *xb = reinterpret_cast<float *>(base + step_bytes);
    void *yb = base + step_bytes + xb_bytes;
    float *stats = reinterpret_cast<float *>(base + step_bytes + xb_bytes + yb_bytes);
    int step = cfg->step;
    if (option(OPT_HEAD_FUSED) && batch <= 64) {
        // one cooperative persistent kernel for the whole epoch (see fused::head_epoch_kernel)
        if ((rc = ac_device_check())) return rc;
        carve(w, workspace, batch, p);
        const int steps = (n + batch - 1) / batch;
        std::vector<float2> bc(steps);
        for (int i = 0; i < steps; ++i)
            bc[i] = make_float2(1.f - powf(cfg->beta1, static_cast<float>(step + i)),
                                sqrtf(1.f - powf(cfg->beta2, static_cast<float>(step + i))));
        float2 *bc_dev = nullptr;
        float *partial_ewc = nullptr;
        AC_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&bc_dev), steps * sizeof(float2), s));
        AC_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&partial_ewc), RED_BLOCKS * sizeof(float), s));
        // pageable source: the runtime stages the copy before returning, so `bc` may go out of scope afterwards
        AC_CUDA(cudaMemcpyAsync(bc_dev, bc.data(), steps * sizeof(float2), cudaMemcpyHostToDevice, s));
        fused::EpochArgs a{};
        a.X = X; a.targets = targets; a.perm = perm; a.n = n; a.batch = batch; a.first_step = step;
        a.p = *p; a.m = *m; a.v = *v; a.w = w; a.xb = xb; a.yb = yb; a.stats = stats; a.loss_accum = loss_accum;
        a.partial_ewc = partial_ewc; a.bias_corr = bc_dev;
        a.lr = cfg->lr; a.beta1 = cfg->beta1; a.beta2 = cfg->beta2; a.eps = cfg->eps; a.weight_decay = cfg->weight_decay;
        a.max_norm = cfg->max_norm; a.dropout_p = cfg->dropout_p; a.loss_kind = cfg->loss_kind; a.seed = cfg->seed;
        a.use_ewc = (cfg->ewc_fisher && cfg->ewc_star) ? 1 : 0;
        a.ewc_C_old = cfg->ewc_C_old; a.ewc_lambda = cfg->ewc_lambda;
        if (a.use_ewc) { a.fisher = *cfg->ewc_fisher; a.star = *cfg->ewc_star; }
        int per_sm = 0;
        AC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused::head_epoch_kernel, fused::FT, 0));
        AC_REQUIRE(per_sm >= 1, "ac_head_train_epoch: the fused epoch kernel does not fit on an SM");
        void *args[] = {&a};
        AC_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(fused::head_epoch_kernel), dim3(sm_count()),
                                            dim3(fused::FT), args, 0, s));
        count_launch();
        AC_CUDA(cudaFreeAsync(bc_dev, s));
        AC_CUDA(cudaFreeAsync(partial_ewc, s));
        return AC_OK;
    }
    for (int off = 0; off < n; off += batch, ++step) {
        const int nb = (n - off < batch) ? n - off : batch;      // DataLoader keeps the last partial batch
        gather_batch_kernel<<<nb, 128, 0, s>>>(X, targets, perm + off, nb, p->D, p->C, cfg->loss_kind, xb, yb);
        AC_LAUNCH_CHECK();
        if ((rc = train_step_impl(xb, yb, nb, p, m, v, cfg, step, stats, workspace, step_bytes, s))) return rc;
        accum_loss_kernel<<<1, 32, 0, s>>>(stats, loss_accum);
        AC_LAUNCH_CHECK();
    }
    return AC_OK;
}
