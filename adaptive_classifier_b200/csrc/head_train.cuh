// head_train.cuh -- the optimizer loop of the adaptive head as ONE persistent cooperative kernel (fp32 SIMT).
//
// Replaces, per optimizer step (paths relative to /root/reference/src/adaptive_classifier/):
//   classifier.py:333-351, :1489-1505   zero_grad / forward (train mode: Dropout 0.1) / CrossEntropyLoss / backward /
//                                       clip_grad_norm_(1.0) / AdamW(lr 1e-3, wd 0.01).step()
//   multilabel.py:387-397               the same with sigmoid outputs + BCELoss
//   ewc.py:67-92, :96-115               gradient of the sampled-label NLL (Fisher) and the EWC penalty gradient
// and the DataLoader batching around it (one launch runs all steps of an epoch from a shuffled index list).
//
// Why one kernel: the head is 0.9 M parameters and a batch is 32 rows -- 171 MFLOP and 25 MB of optimizer traffic per step,
// i.e. microseconds of work; the round-1 path launched ~21 dependent kernels per step (310 us measured on a B200).  Here the
// grid stays resident for the whole epoch and a step is seven phases separated by six grid barriers:
//
//   ownership   the rows of every weight matrix are dealt to the CTAs in blocks of HT_RB = 8 rows (block q -> CTA q % G).
//               A CTA keeps ITS rows of W0, W1, W2 (+ biases) in shared memory for the whole launch, computes the
//               activations / gradients of exactly those rows and applies AdamW to them: parameters never move between
//               CTAs, gradients never leave shared memory, AdamW of step t needs no barrier before the forward of step t+1.
//   P1  h0 = dropout(relu(X W0^T + b0))        own rows of layer 0; X rows gathered through the shuffled index list
//   P2  h1 = dropout(relu(h0 W1^T + b1))       own rows of layer 1
//   P3a z  = h1 W2^T + b2                      own rows of layer 2
//   P3b loss, dz per batch row                 one warp per row (softmax-CE or sigmoid-BCE)
//   P4  gW2, gb2 (own rows of layer 2);  da1 = (dz W2) * relu' * mask   (own columns = own rows of layer 1)
//   P5  gW1, gb1 (own rows of layer 1);  da0 = (da1 W1) * relu' * mask  (own rows of layer 0)
//   P6  gW0, gb0;  [EWC: g += 2 lambda / B * F (theta - theta*)];  partial sum of squares of the own gradients
//   P7  global grad norm (every CTA adds the G partials in the same order), clip, AdamW on the own rows
//   Activations cross CTAs through small global (L2-resident) buffers; every product streams its [B x K] operand through shared
//   memory in chunks of HT_KC columns (register-prefetched), so shared memory holds only the own parameter / gradient rows.
//
// All sums have a fixed order: results are deterministic run to run and independent of the grid size up to fp32 rounding of
// the (grid-size dependent) partial-sum order of the gradient norm.  Parity: the CPU restatement of the optimizer step (tests/test_gpu_parity.py,
// tests/test_gpu_training_golden.py: the reference's own per-step losses to 1e-5 over 60 steps).
//
// This header is plain SIMT C++ (no inline PTX): tests/cpu_shim/head_train_emul.cpp compiles it for the CPU (every CUDA
// thread a fiber, grid barriers real) and checks it against a straightforward restatement before any GPU time is spent.
#pragma once
#include <stdint.h>
#if !defined(AC_CPU_SHIM)
#include <cuda_pipeline.h>
#endif

namespace ac {
namespace ht {

constexpr int HT_THREADS = 256;
constexpr int HT_RB = 8;              // rows per ownership block
constexpr int HT_KC = 128;            // columns per streamed chunk
constexpr int HT_AS = HT_KC + 4;      // padded row stride of the chunk buffer (floats): conflict-free float4 rows
constexpr int HT_MAXB = 64;           // rows per batch
constexpr int HT_KPARTS = HT_THREADS / 32;   // 8 warps split a chunk's columns: 16 each
static_assert(HT_KC == HT_KPARTS * 16, "a warp owns 16 columns of a chunk");

struct Layer {
    float *W, *b;                 // [rows, K], [rows]   parameters (global; updated in place)
    float *mW, *mb, *vW, *vb;     // AdamW moments (update mode)
    const float *fW, *fb, *sW, *sb;   // EWC Fisher / theta* (nullable)
    float *xW;                    // [rows, K] gradient scratch in global memory (L2-resident: written and read by the owner CTA only)
    float *gW, *gb;               // gradient outputs (gradient-only mode, nullable)
    float *qW, *qb;               // Fisher accumulators: q += g^2 * fisher_scale (gradient-only mode, nullable)
    int rows, K, ewc_rows;        // ewc_rows: only the first ewc_rows rows carry the EWC term (the head may have grown)
};

struct Args {
    const float *X;               // [n, D]
    const void *targets;          // int64 [n] (CE) or float [n, C] (BCE)
    const int64_t *perm;          // [n] shuffled row order, NULL = identity
    int n, batch, n_steps, first_step;
    Layer L[3];
    float lr, beta1, beta2, eps, wd, max_norm, dropout_p;
    int loss_kind;                // 0 CE, 1 BCE
    unsigned long long seed;
    const float *mask0, *mask1;   // injected dropout masks [B,H0], [B,H1] (single step) or NULL
    int use_ewc;
    float ewc_lambda;
    int update;                   // 1: clip + AdamW;  0: gradient only (Fisher)
    float fisher_scale;
    // global scratch
    float *h0d, *h1d, *z, *dz, *da1, *rowloss, *part, *pen;
    float *stats;                 // [n_steps, 3] (task loss, EWC penalty, grad norm before clipping), nullable
    float *loss_accum;            // [1] += loss + penalty per step, nullable
    unsigned *bar;                // [2] grid barrier state (count, generation), zero-initialised
    int slots[3];                 // ownership blocks per CTA of each layer = ceil(ceil(rows / 8) / G)
    int nst;                      // stages of the streamed-operand ring (2..4)
    unsigned long long *timing;   // nullable, [16]: nanoseconds CTA 0 spent up to each phase boundary, summed over the steps
};

// ---------------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------------
#if !defined(AC_CPU_SHIM)
#define HT_LDCG(p) __ldcg(p)      // L2 (coherent across CTAs): everything another CTA wrote inside this launch
#else
#define HT_LDCG(p) (*(p))
#endif

// asynchronous global -> shared copies (cp.async through the pipeline intrinsics); `valid` = false zero-fills the destination
#if !defined(AC_CPU_SHIM)
__device__ __forceinline__ void ht_async16(float *dst, const float *src, bool valid) { __pipeline_memcpy_async(dst, src, 16, valid ? 0 : 16); }
__device__ __forceinline__ void ht_async4(float *dst, const float *src, bool valid) { __pipeline_memcpy_async(dst, src, 4, valid ? 0 : 4); }
__device__ __forceinline__ void ht_async_commit() { __pipeline_commit(); }
template <int N> __device__ __forceinline__ void ht_async_wait() { __pipeline_wait_prior(N); }
#else
static inline void ht_async16(float *dst, const float *src, bool valid) { for (int i = 0; i < 4; ++i) dst[i] = valid ? src[i] : 0.f; }
static inline void ht_async4(float *dst, const float *src, bool valid) { dst[0] = valid ? src[0] : 0.f; }
static inline void ht_async_commit() {}
template <int N> static inline void ht_async_wait() {}
#endif
__device__ __forceinline__ void ht_async_wait_n(int n) {       // n = stages - 2 in [0, 6]
    switch (n) {
        case 0: ht_async_wait<0>(); break;
        case 1: ht_async_wait<1>(); break;
        case 2: ht_async_wait<2>(); break;
        case 3: ht_async_wait<3>(); break;
        case 4: ht_async_wait<4>(); break;
        case 5: ht_async_wait<5>(); break;
        default: ht_async_wait<6>(); break;
    }
}

__device__ __forceinline__ uint32_t ht_mix32(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return static_cast<uint32_t>(x);
}
// counter-hash dropout mask of element i of mask stream `stream_id` (0 or 1/(1-p))
__device__ __forceinline__ float ht_mask(float p, unsigned long long seed, unsigned long long stream_id, unsigned long long i) {
    const uint32_t r = ht_mix32(seed * 0x9E3779B97F4A7C15ULL + stream_id * 0xD1B54A32D192ED03ULL + i);
    const float u = (r >> 8) * (1.0f / 16777216.0f);
    return (u < p) ? 0.f : 1.f / (1.f - p);
}

// phase timing (diagnostic, tools/head_phase_times.py): CTA 0 accumulates the global-timer delta since the previous stamp
#if !defined(AC_CPU_SHIM)
__device__ __forceinline__ unsigned long long ht_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define HT_STAMP(i)                                                                  \
    do {                                                                             \
        if (a.timing && blockIdx.x == 0 && threadIdx.x == 0) {                       \
            const unsigned long long now_ = ht_now();                               \
            a.timing[i] += now_ - t_prev;                                            \
            t_prev = now_;                                                           \
        }                                                                            \
    } while (0)
#else
static inline unsigned long long ht_now() { return 0; }
#define HT_STAMP(i) do { (void)t_prev; } while (0)
#endif

// grid barrier: all threads of all CTAs.  Cooperative launch guarantees co-residency; a watchdog turns a protocol bug into a
// launch error instead of a hung GPU.
__device__ __forceinline__ void ht_grid_sync(unsigned *bar, unsigned &gen) {
#if !defined(AC_CPU_SHIM)
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned target = gen + 1;
        __threadfence();                                        // publish this CTA's global writes
        if (atomicAdd(bar, 1u) == gridDim.x - 1) {
            atomicExch(bar, 0u);
            __threadfence();
            atomicExch(bar + 1, target);
        } else {
            unsigned spins = 0;
            while (*reinterpret_cast<volatile unsigned *>(bar + 1) != target) {
                if (++spins > (1u << 28)) { printf("ac: head_train grid barrier watchdog (block %d)\n", blockIdx.x); __trap(); }
            }
        }
        __threadfence();
    }
    __syncthreads();
    gen += 1;
#else
    (void)bar;
    cooperative_groups::this_grid().sync();
    gen += 1;
#endif
}

// block-wide sum in a fixed tree order; red: HT_THREADS floats of shared memory; every thread gets the result
__device__ __forceinline__ float ht_block_sum(float v, float *red) {
    __syncthreads();
    red[threadIdx.x] = v;
    __syncthreads();
    for (int s = HT_THREADS / 2; s > 0; s >>= 1) {
        if (static_cast<int>(threadIdx.x) < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    const float r = red[0];
    __syncthreads();
    return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// streamed operand: chunk [rows x HT_KC] of a row-major [*, K] matrix (rows direct or through ridx) -> a stage of the shared
// memory ring, by asynchronous copies.  `nst` stages are in flight (2..4, whatever the parameter rows leave room for): at
// B = 32 the FMAs of a chunk take ~0.3 us while an L2 round trip takes ~1 us, so the ring depth, not the arithmetic, sets the
// time of a phase.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ht_issue_chunk(float *stage, const float *src, int64_t ld, const int64_t *ridx, int rows, int k0, int K) {
    const bool vec = (K & 3) == 0 && (ld & 3) == 0;
    for (int e = threadIdx.x; e < rows * (HT_KC / 4); e += HT_THREADS) {
        const int row = e / (HT_KC / 4), c4 = e % (HT_KC / 4);
        const int k = k0 + 4 * c4;
        const float *p = src + (ridx ? ridx[row] : static_cast<int64_t>(row)) * ld + k;
        float *d = stage + row * HT_AS + 4 * c4;
        if (vec) {
            ht_async16(d, k < K ? p : src, k < K);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) ht_async4(d + i, k + i < K ? p + i : src, k + i < K);
        }
    }
}
// weight chunk of an input-gradient product: Wt[jj][kk] = gW[(k0 + kk) * gld + gcol0 + jj] (zero outside the matrix)
__device__ __forceinline__ void ht_issue_wt(float *wt_stage, const float *gW, int64_t gld, int gcol0, int gcols, int k0, int K) {
    for (int e = threadIdx.x; e < HT_KC * HT_RB; e += HT_THREADS) {
        const int kk = e / HT_RB, jj = e % HT_RB;
        const bool ok = k0 + kk < K && jj < gcols;
        ht_async4(wt_stage + jj * HT_KC + kk, ok ? gW + static_cast<int64_t>(k0 + kk) * gld + gcol0 + jj : gW, ok);
    }
}

// Y[b, j] = sum_k A[b, k] * Wt[j][k]  for the 8 rows j of one ownership block, A streamed in chunks.
//   A: [rows x K] global (ld, optional row index list)
//   w_resident != nullptr: the weights are the block's resident parameter rows [8][K] in shared memory (forward products)
//   otherwise the weight chunk is gathered from global W[(k0 + kk) * gld + gcol0 + jj] (input-gradient products)
// Result: out[b * 8 + j] in shared memory (valid for b < rows), summed over the 8 column parts in a fixed order.
__device__ __forceinline__ void ht_rows_dot(float *out, float *As, float *red, float *Wt, int nst, const float *A, int64_t ld,
                                            const int64_t *ridx, int rows, int K, const float *w_resident /* [8][K] or null */,
                                            const float *gW, int64_t gld, int gcol0, int gcols /* valid columns <= 8 */) {
    const int lane = threadIdx.x & 31, kpart = threadIdx.x >> 5;
    const int nb = (rows + 31) >> 5;
    const int stage_floats = rows * HT_AS;
    float acc[2][HT_RB];
#pragma unroll
    for (int bi = 0; bi < 2; ++bi)
#pragma unroll
        for (int j = 0; j < HT_RB; ++j) acc[bi][j] = 0.f;
    const int nchunks = (K + HT_KC - 1) / HT_KC;
    for (int p = 0; p < nst - 1; ++p) {
        if (p < nchunks) {
            ht_issue_chunk(As + p * stage_floats, A, ld, ridx, rows, p * HT_KC, K);
            if (!w_resident) ht_issue_wt(Wt + p * HT_RB * HT_KC, gW, gld, gcol0, gcols, p * HT_KC, K);
        }
        ht_async_commit();
    }
    for (int c = 0; c < nchunks; ++c) {
        const int k0 = c * HT_KC;
        ht_async_wait_n(nst - 2);                       // chunk c has landed (for this thread's copies) ...
        __syncthreads();                                // ... and for everybody's; the stage read in iteration c - 1 is free again
        const int nx = c + nst - 1;
        if (nx < nchunks) {
            ht_issue_chunk(As + (nx % nst) * stage_floats, A, ld, ridx, rows, nx * HT_KC, K);
            if (!w_resident) ht_issue_wt(Wt + (nx % nst) * HT_RB * HT_KC, gW, gld, gcol0, gcols, nx * HT_KC, K);
        }
        ht_async_commit();
        const float *Ac = As + (c % nst) * stage_floats;
        const float *wbase = w_resident ? (w_resident + k0) : (Wt + (c % nst) * HT_RB * HT_KC);
        const int wld = w_resident ? K : HT_KC;
        const int kcols = (K - k0 < HT_KC) ? (K - k0) : HT_KC;                           // resident rows: stay inside the row
#pragma unroll
        for (int bi = 0; bi < 2; ++bi) {
            if (bi < nb) {
                const int b = lane + 32 * bi;
                const float *a = Ac + b * HT_AS + kpart * 16;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int kk = kpart * 16 + 4 * q4;
                    if (kk < kcols) {
                        const float4 av = *reinterpret_cast<const float4 *>(a + 4 * q4);
#pragma unroll
                        for (int j = 0; j < HT_RB; ++j) {
                            // one 16-byte broadcast load per 4 weights (rows are 16-byte aligned: K, HT_KC and kk are multiples
                            // of 4); resident rows: K % 4 == 0 (host-checked); gathered chunks are zero-padded to HT_KC
                            const float4 wv = *reinterpret_cast<const float4 *>(wbase + j * wld + kk);
                            float s = acc[bi][j];
                            s = fmaf(av.x, wv.x, s);
                            s = fmaf(av.y, wv.y, s);
                            s = fmaf(av.z, wv.z, s);
                            s = fmaf(av.w, wv.w, s);
                            acc[bi][j] = s;
                        }
                    }
                }
            }
        }
    }
    // combine the 8 column parts in order
#pragma unroll
    for (int bi = 0; bi < 2; ++bi)
        if (bi < nb) {
            const int b = lane + 32 * bi;
#pragma unroll
            for (int j = 0; j < HT_RB; ++j) red[(kpart * HT_MAXB + b) * HT_RB + j] = acc[bi][j];
        }
    __syncthreads();
    for (int e = threadIdx.x; e < rows * HT_RB; e += HT_THREADS) {
        const int b = e / HT_RB, j = e % HT_RB;
        float s = 0.f;
#pragma unroll
        for (int kp = 0; kp < HT_KPARTS; ++kp) s += red[(kp * HT_MAXB + b) * HT_RB + j];
        out[b * HT_RB + j] = s;
    }
    __syncthreads();
}

// g[j][k] = sum_b dA[b][j] * A[b][k]  for the nrow (<= 8) rows j of one ownership block (weight gradient), A streamed in chunks;
// g: the block's rows of the gradient scratch in GLOBAL memory ([8][K], owner-private; shared memory is kept for the operand
// ring); dA: shared memory [rows][8]; batch rows are added in index order.
__device__ __forceinline__ void ht_outer_acc(float *g, int nrow, float *As, int nst, const float *dA, const float *A, int64_t ld,
                                             const int64_t *ridx, int rows, int K) {
    const int kk = threadIdx.x % HT_KC, jh = threadIdx.x / HT_KC;      // 2 x 128 threads: columns x row halves
    const int nchunks = (K + HT_KC - 1) / HT_KC;
    const int stage_floats = rows * HT_AS;
    for (int p = 0; p < nst - 1; ++p) {
        if (p < nchunks) ht_issue_chunk(As + p * stage_floats, A, ld, ridx, rows, p * HT_KC, K);
        ht_async_commit();
    }
    for (int c = 0; c < nchunks; ++c) {
        const int k0 = c * HT_KC;
        ht_async_wait_n(nst - 2);
        __syncthreads();
        const int nx = c + nst - 1;
        if (nx < nchunks) ht_issue_chunk(As + (nx % nst) * stage_floats, A, ld, ridx, rows, nx * HT_KC, K);
        ht_async_commit();
        const float *Ac = As + (c % nst) * stage_floats;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int b = 0; b < rows; ++b) {
            const float a = Ac[b * HT_AS + kk];
            const float4 d = *reinterpret_cast<const float4 *>(dA + b * HT_RB + 4 * jh);
            a0 = fmaf(d.x, a, a0);
            a1 = fmaf(d.y, a, a1);
            a2 = fmaf(d.z, a, a2);
            a3 = fmaf(d.w, a, a3);
        }
        if (k0 + kk < K) {
            if (4 * jh + 0 < nrow) g[static_cast<int64_t>(4 * jh + 0) * K + k0 + kk] = a0;
            if (4 * jh + 1 < nrow) g[static_cast<int64_t>(4 * jh + 1) * K + k0 + kk] = a1;
            if (4 * jh + 2 < nrow) g[static_cast<int64_t>(4 * jh + 2) * K + k0 + kk] = a2;
            if (4 * jh + 3 < nrow) g[static_cast<int64_t>(4 * jh + 3) * K + k0 + kk] = a3;
        }
    }
    __syncthreads();
}

// shared-memory carve-up (floats), identical on host and device
struct Smem {
    int th[3], bs[3], gb[3];           // parameter rows / biases / bias gradients per layer: [slots * 8][K], [slots * 8], [slots * 8]
    int f0, f1;                        // relu' * mask factors, later da0 / da1 of the own rows: [slots][batch][8]
    int dA, out, As, Wt, red, rsum, ridx /* int64 */, scal, total;
};
__host__ __device__ inline Smem ht_smem_layout(const Args &a) {
    Smem s;
    int off = 0;
    auto take = [&](int n) { const int o = off; off += (n + 3) & ~3; return o; };
    for (int l = 0; l < 3; ++l) { s.th[l] = take(a.slots[l] * HT_RB * a.L[l].K); s.bs[l] = take(a.slots[l] * HT_RB); }
    for (int l = 0; l < 3; ++l) s.gb[l] = take(a.slots[l] * HT_RB);
    s.f0 = take(a.slots[0] * a.batch * HT_RB);
    s.f1 = take(a.slots[1] * a.batch * HT_RB);
    s.dA = take(a.batch * HT_RB);
    s.out = take(a.batch * HT_RB);
    s.As = take(a.nst * a.batch * HT_AS);
    s.Wt = take(a.nst * HT_RB * HT_KC);
    s.red = take(HT_KPARTS * HT_MAXB * HT_RB);
    s.rsum = take(HT_THREADS);
    s.ridx = take(2 * HT_MAXB);
    s.scal = take(16);
    s.total = off;
    return s;
}

// ---------------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HT_THREADS, 1) head_train_kernel(const Args a) {
#if !defined(AC_CPU_SHIM)
    extern __shared__ __align__(16) float ht_smem[];
#else
    float *ht_smem = reinterpret_cast<float *>(shim_dyn_smem());
#endif
    const Smem sm = ht_smem_layout(a);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const int D = a.L[0].K, H0 = a.L[0].rows, H1 = a.L[1].rows, C = a.L[2].rows;
    float *As = ht_smem + sm.As, *Wt = ht_smem + sm.Wt, *red = ht_smem + sm.red, *rsum = ht_smem + sm.rsum;
    float *dA = ht_smem + sm.dA, *out = ht_smem + sm.out, *scal = ht_smem + sm.scal;
    int64_t *ridx = reinterpret_cast<int64_t *>(ht_smem + sm.ridx);
    unsigned gen = 0;

    // ---- resident parameter rows
    for (int l = 0; l < 3; ++l) {
        const Layer &L = a.L[l];
        const int nblk = (L.rows + HT_RB - 1) / HT_RB;
        for (int s = 0; s < a.slots[l]; ++s) {
            const int q = cta + s * G;
            for (int e = tid; e < HT_RB * L.K; e += HT_THREADS) {
                const int j = e / L.K, k = e % L.K;
                const int r = q * HT_RB + j;
                ht_smem[sm.th[l] + (s * HT_RB + j) * L.K + k] = (q < nblk && r < L.rows) ? L.W[static_cast<int64_t>(r) * L.K + k] : 0.f;
            }
            if (tid < HT_RB) {
                const int r = q * HT_RB + tid;
                ht_smem[sm.bs[l] + s * HT_RB + tid] = (q < nblk && r < L.rows) ? L.b[r] : 0.f;
            }
        }
    }
    __syncthreads();

    unsigned long long t_prev = ht_now();
    for (int t = 0; t < a.n_steps; ++t) {
        const int step = a.first_step + t;
        const int off = t * a.batch;
        const int Bt = (a.n - off < a.batch) ? (a.n - off) : a.batch;
        if (tid < Bt) ridx[tid] = a.perm ? a.perm[off + tid] : static_cast<int64_t>(off + tid);
        if (tid == HT_THREADS - 1) {      // AdamW bias corrections of this step (double precision like Python's 1 - beta ** step), off the critical path
            scal[1] = static_cast<float>(1.0 - pow(static_cast<double>(a.beta1), static_cast<double>(step)));
            scal[2] = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(a.beta2), static_cast<double>(step))));
        }
        __syncthreads();
        const bool drop = a.dropout_p > 0.f;

        // ================= P1 / P2 / P3a: forward of the own rows =================
        for (int l = 0; l < 3; ++l) {
            const Layer &L = a.L[l];
            const int nblk = (L.rows + HT_RB - 1) / HT_RB;
            const float *A = l == 0 ? a.X : (l == 1 ? a.h0d : a.h1d);
            float *dst = l == 0 ? a.h0d : (l == 1 ? a.h1d : a.z);
            float *fac = l == 0 ? ht_smem + sm.f0 : ht_smem + sm.f1;
            const float *inj = l == 0 ? a.mask0 : a.mask1;
            for (int s = 0; s < a.slots[l]; ++s) {
                const int q = cta + s * G;
                if (q >= nblk) break;
                ht_rows_dot(out, As, red, Wt, a.nst, A, L.K, l == 0 ? ridx : nullptr, Bt, L.K, ht_smem + sm.th[l] + s * HT_RB * L.K, nullptr, 0, 0, 0);
                for (int e = tid; e < Bt * HT_RB; e += HT_THREADS) {
                    const int b = e / HT_RB, j = e % HT_RB;
                    const int r = q * HT_RB + j;
                    if (r >= L.rows) continue;
                    const float pre = out[e] + ht_smem[sm.bs[l] + s * HT_RB + j];
                    if (l == 2) {
                        dst[static_cast<int64_t>(b) * C + r] = pre;
                    } else {
                        float mk = 1.f;
                        if (drop) mk = inj ? inj[static_cast<int64_t>(b) * L.rows + r]
                                           : ht_mask(a.dropout_p, a.seed, 2ull * step + l, static_cast<unsigned long long>(b) * L.rows + r);
                        const float h = pre > 0.f ? pre : 0.f;
                        dst[static_cast<int64_t>(b) * L.rows + r] = h * mk;
                        fac[(s * a.batch + b) * HT_RB + j] = pre > 0.f ? mk : 0.f;
                    }
                }
                __syncthreads();
            }
            HT_STAMP(2 * l);
            ht_grid_sync(a.bar, gen);
            HT_STAMP(2 * l + 1);
        }

        // ================= P3b: loss and dz, one warp per batch row =================
        for (int b = cta + G * warp; b < Bt; b += G * HT_KPARTS) {
            const float *zr = a.z + static_cast<int64_t>(b) * C;
            float *dr = a.dz + static_cast<int64_t>(b) * C;
            if (a.loss_kind == 0) {
                const int64_t y = static_cast<const int64_t *>(a.targets)[ridx[b]];
                float mx = -3.402823466e38f;
                for (int j = lane; j < C; j += 32) mx = fmaxf(mx, HT_LDCG(zr + j));
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                float sum = 0.f;
                for (int j = lane; j < C; j += 32) sum += expf(HT_LDCG(zr + j) - mx);
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                const float lse = mx + logf(sum);
                const float invB = 1.f / static_cast<float>(Bt);
                for (int j = lane; j < C; j += 32) {
                    const float p = expf(HT_LDCG(zr + j) - mx) / sum;
                    dr[j] = (p - (j == y ? 1.f : 0.f)) * invB;
                }
                if (lane == 0) a.rowloss[b] = (y >= 0 && y < C) ? (lse - HT_LDCG(zr + y)) : 0.f;
            } else {
                const float *yr = static_cast<const float *>(a.targets) + ridx[b] * C;
                const float inv = 1.f / (static_cast<float>(Bt) * static_cast<float>(C));
                float l = 0.f;
                for (int j = lane; j < C; j += 32) {
                    const float s = 1.f / (1.f + expf(-HT_LDCG(zr + j)));
                    const float y = yr[j];
                    l -= y * fmaxf(logf(s), -100.f) + (1.f - y) * fmaxf(logf(1.f - s), -100.f);   // nn.BCELoss clamps log at -100
                    dr[j] = (s - y) * inv;
                }
                for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
                if (lane == 0) a.rowloss[b] = l / static_cast<float>(C);
            }
        }
        HT_STAMP(6);
        ht_grid_sync(a.bar, gen);
        HT_STAMP(7);

        // ================= P4: layer-2 weight gradients; da1 of the own layer-1 rows =================
        {
            const Layer &L2 = a.L[2], &L1 = a.L[1];
            const int nblk2 = (C + HT_RB - 1) / HT_RB, nblk1 = (H1 + HT_RB - 1) / HT_RB;
            for (int s = 0; s < a.slots[2]; ++s) {
                const int q = cta + s * G;
                if (q >= nblk2) break;
                for (int e = tid; e < Bt * HT_RB; e += HT_THREADS) {
                    const int b = e / HT_RB, r = q * HT_RB + e % HT_RB;
                    dA[e] = r < C ? HT_LDCG(a.dz + static_cast<int64_t>(b) * C + r) : 0.f;
                }
                __syncthreads();
                if (tid < HT_RB) {
                    float sb = 0.f;
                    for (int b = 0; b < Bt; ++b) sb += dA[b * HT_RB + tid];
                    ht_smem[sm.gb[2] + s * HT_RB + tid] = sb;
                }
                ht_outer_acc(L2.xW + static_cast<int64_t>(q) * HT_RB * L2.K, (C - q * HT_RB < HT_RB) ? (C - q * HT_RB) : HT_RB, As, a.nst, dA, a.h1d, H1, nullptr, Bt, H1);
            }
            for (int s = 0; s < a.slots[1]; ++s) {
                const int q = cta + s * G;
                if (q >= nblk1) break;
                const int cols = (H1 - q * HT_RB < HT_RB) ? (H1 - q * HT_RB) : HT_RB;
                ht_rows_dot(out, As, red, Wt, a.nst, a.dz, C, nullptr, Bt, C, nullptr, L2.W, H1, q * HT_RB, cols);
                float *f1 = ht_smem + sm.f1 + s * a.batch * HT_RB;
                for (int e = tid; e < Bt * HT_RB; e += HT_THREADS) {
                    const int b = e / HT_RB, j = e % HT_RB, r = q * HT_RB + j;
                    const float d = out[e] * f1[b * HT_RB + j];
                    f1[b * HT_RB + j] = r < H1 ? d : 0.f;                       // da1 of the own rows (for gW1 in P5)
                    if (r < H1) a.da1[static_cast<int64_t>(b) * H1 + r] = d;
                }
                __syncthreads();
            }
            (void)L1;
        }
        HT_STAMP(8);
        ht_grid_sync(a.bar, gen);
        HT_STAMP(9);

        // ================= P5: layer-1 weight gradients; da0 of the own layer-0 rows.  P6: layer-0 weight gradients ==========
        {
            const Layer &L1 = a.L[1], &L0 = a.L[0];
            const int nblk1 = (H1 + HT_RB - 1) / HT_RB, nblk0 = (H0 + HT_RB - 1) / HT_RB;
            for (int s = 0; s < a.slots[1]; ++s) {
                const int q = cta + s * G;
                if (q >= nblk1) break;
                const float *d1 = ht_smem + sm.f1 + s * a.batch * HT_RB;
                if (tid < HT_RB) {
                    float sb = 0.f;
                    for (int b = 0; b < Bt; ++b) sb += d1[b * HT_RB + tid];
                    ht_smem[sm.gb[1] + s * HT_RB + tid] = sb;
                }
                ht_outer_acc(L1.xW + static_cast<int64_t>(q) * HT_RB * L1.K, (H1 - q * HT_RB < HT_RB) ? (H1 - q * HT_RB) : HT_RB, As, a.nst, d1, a.h0d, H0, nullptr, Bt, H0);
            }
            for (int s = 0; s < a.slots[0]; ++s) {
                const int q = cta + s * G;
                if (q >= nblk0) break;
                const int cols = (H0 - q * HT_RB < HT_RB) ? (H0 - q * HT_RB) : HT_RB;
                ht_rows_dot(out, As, red, Wt, a.nst, a.da1, H1, nullptr, Bt, H1, nullptr, L1.W, H0, q * HT_RB, cols);
                float *f0 = ht_smem + sm.f0 + s * a.batch * HT_RB;
                for (int e = tid; e < Bt * HT_RB; e += HT_THREADS) {
                    const int j = e % HT_RB, r = q * HT_RB + j;
                    f0[e] = r < H0 ? out[e] * f0[e] : 0.f;                      // da0 of the own rows
                }
                __syncthreads();
                if (tid < HT_RB) {
                    float sb = 0.f;
                    for (int b = 0; b < Bt; ++b) sb += f0[b * HT_RB + tid];
                    ht_smem[sm.gb[0] + s * HT_RB + tid] = sb;
                }
                ht_outer_acc(L0.xW + static_cast<int64_t>(q) * HT_RB * L0.K, (H0 - q * HT_RB < HT_RB) ? (H0 - q * HT_RB) : HT_RB, As, a.nst, f0, a.X, D, ridx, Bt, D);
            }
        }
        __syncthreads();

        // ---- EWC gradient on the own rows, partial sum of squares of the own gradients (row loops: no integer divisions)
        float ss = 0.f, pen = 0.f;
        const float ewc2 = a.use_ewc ? 2.f * a.ewc_lambda / static_cast<float>(Bt) : 0.f;
        for (int l = 0; l < 3; ++l) {
            const Layer &L = a.L[l];
            const int nblk = (L.rows + HT_RB - 1) / HT_RB;
            const int K = L.K;
            for (int s = 0; s < a.slots[l]; ++s) {
                const int q = cta + s * G;
                if (q >= nblk) break;
                float *g = L.xW + static_cast<int64_t>(q) * HT_RB * K;         // written by this CTA's ht_outer_acc above
                const float *th = ht_smem + sm.th[l] + s * HT_RB * K;
                const int nrow = (L.rows - q * HT_RB < HT_RB) ? (L.rows - q * HT_RB) : HT_RB;
                for (int j = 0; j < nrow; ++j) {
                    const int r = q * HT_RB + j;
                    const bool ew = a.use_ewc && r < L.ewc_rows;
                    for (int k = tid; k < K; k += HT_THREADS) {
                        float gv = HT_LDCG(g + j * K + k);
                        if (ew) {
                            const int64_t gi = static_cast<int64_t>(r) * K + k;
                            const float dlt = th[j * K + k] - L.sW[gi];
                            const float f = L.fW[gi];
                            gv = fmaf(ewc2 * f, dlt, gv);
                            pen = fmaf(f * dlt, dlt, pen);
                            g[j * K + k] = gv;
                        }
                        ss = fmaf(gv, gv, ss);
                    }
                }
                if (tid < nrow) {
                    const int r = q * HT_RB + tid;
                    float gv = ht_smem[sm.gb[l] + s * HT_RB + tid];
                    if (a.use_ewc && r < L.ewc_rows) {
                        const float dlt = ht_smem[sm.bs[l] + s * HT_RB + tid] - L.sb[r];
                        const float f = L.fb[r];
                        gv = fmaf(ewc2 * f, dlt, gv);
                        pen = fmaf(f * dlt, dlt, pen);
                        ht_smem[sm.gb[l] + s * HT_RB + tid] = gv;
                    }
                    ss = fmaf(gv, gv, ss);
                }
            }
        }
        ss = ht_block_sum(ss, rsum);
        pen = ht_block_sum(pen, rsum);
        if (tid == 0) { a.part[cta] = ss; a.pen[cta] = pen; }
        HT_STAMP(10);
        ht_grid_sync(a.bar, gen);
        HT_STAMP(11);

        // ================= P7: global norm, clip, AdamW on the own rows =================
        float tot = 0.f, pt = 0.f, ls = 0.f;
        if (warp == 0) {
            // every CTA adds the G partials in the same fixed order: lane-strided sums, then a shuffle tree
            for (int i = lane; i < G; i += 32) { tot += HT_LDCG(a.part + i); pt += HT_LDCG(a.pen + i); }
            for (int o = 16; o > 0; o >>= 1) { tot += __shfl_xor_sync(0xffffffffu, tot, o); pt += __shfl_xor_sync(0xffffffffu, pt, o); }
            if (cta == 0) {                 // batch loss: the row losses are fetched in parallel, added in row order
                float r0 = lane < Bt ? HT_LDCG(a.rowloss + lane) : 0.f, r1 = lane + 32 < Bt ? HT_LDCG(a.rowloss + lane + 32) : 0.f;
                for (int b = 0; b < 32; ++b) ls += __shfl_sync(0xffffffffu, r0, b);
                for (int b = 0; b < 32; ++b) ls += __shfl_sync(0xffffffffu, r1, b);
            }
        }
        if (tid == 0) {
            const float norm = sqrtf(tot);
            float coef = a.max_norm / (norm + 1e-6f);
            coef = coef < 1.f ? coef : 1.f;
            if (!(a.max_norm > 0.f)) coef = 1.f;
            scal[0] = coef;
            if (cta == 0) {
                const float loss = ls / static_cast<float>(Bt);
                const float penalty = a.use_ewc ? a.ewc_lambda / static_cast<float>(Bt) * pt : 0.f;
                if (a.stats) { a.stats[3 * t + 0] = loss; a.stats[3 * t + 1] = penalty; a.stats[3 * t + 2] = norm; }
                if (a.loss_accum) a.loss_accum[0] += loss + penalty;
            }
        }
        __syncthreads();
        const float coef = scal[0], bc1 = scal[1], bc2s = scal[2];
        const float decay = 1.f - a.lr * a.wd, lr_c = a.lr / bc1, omb1 = 1.f - a.beta1, omb2 = 1.f - a.beta2;
        for (int l = 0; l < 3; ++l) {
            const Layer &L = a.L[l];
            const int nblk = (L.rows + HT_RB - 1) / HT_RB;
            const int K = L.K;
            for (int s = 0; s < a.slots[l]; ++s) {
                const int q = cta + s * G;
                if (q >= nblk) break;
                float *th = ht_smem + sm.th[l] + s * HT_RB * K;
                const int nrow = (L.rows - q * HT_RB < HT_RB) ? (L.rows - q * HT_RB) : HT_RB;
                const int64_t g00 = static_cast<int64_t>(q) * HT_RB * K;
                // thread = column k of the block's 8 rows.  Gradient and moments live in global memory (L2): the 24 loads of a
                // column are issued together (coalesced along k), and the loads of column k + 256 are in flight while column k
                // is computed (two register buffers; the loop is fully unrolled so the buffer index is a constant)
                constexpr int HT_NIT = 8;                             // K <= 8 * 256
                float gr[2][HT_RB], mi[2][HT_RB], vi[2][HT_RB];
#pragma unroll
                for (int it = 0; it <= HT_NIT; ++it) {
                    if (it < HT_NIT) {
                        const int k = tid + it * HT_THREADS;
                        if (k < K) {
#pragma unroll
                            for (int j = 0; j < HT_RB; ++j) {
                                const int64_t gi = g00 + static_cast<int64_t>(j) * K + k;
                                gr[it & 1][j] = j < nrow ? HT_LDCG(L.xW + gi) : 0.f;
                                mi[it & 1][j] = (a.update && j < nrow) ? L.mW[gi] : 0.f;
                                vi[it & 1][j] = (a.update && j < nrow) ? L.vW[gi] : 0.f;
                            }
                        }
                    }
                    if (it >= 1) {
                        const int k = tid + (it - 1) * HT_THREADS;
                        if (k < K) {
#pragma unroll
                            for (int j = 0; j < HT_RB; ++j) {
                                if (j >= nrow) continue;
                                const int64_t gi = g00 + static_cast<int64_t>(j) * K + k;
                                const float graw = gr[(it - 1) & 1][j];
                                if (a.update) {
                                    const float gv = graw * coef;
                                    float p = th[j * K + k] * decay;
                                    const float m1 = mi[(it - 1) & 1][j] * a.beta1 + gv * omb1;
                                    const float v1 = vi[(it - 1) & 1][j] * a.beta2 + gv * gv * omb2;
                                    const float denom = sqrtf(v1) / bc2s + a.eps;
                                    p = p - lr_c * (m1 / denom);
                                    th[j * K + k] = p;
                                    L.W[gi] = p;
                                    L.mW[gi] = m1;
                                    L.vW[gi] = v1;
                                } else {
                                    if (L.gW) L.gW[gi] = graw;
                                    if (L.qW) L.qW[gi] += graw * graw * a.fisher_scale;
                                }
                            }
                        }
                    }
                }
                if (tid < nrow) {
                    const int r = q * HT_RB + tid;
                    const float graw = ht_smem[sm.gb[l] + s * HT_RB + tid];
                    if (a.update) {
                        const float gv = graw * coef;
                        float p = ht_smem[sm.bs[l] + s * HT_RB + tid] * decay;
                        const float m1 = L.mb[r] * a.beta1 + gv * omb1;
                        const float v1 = L.vb[r] * a.beta2 + gv * gv * omb2;
                        const float denom = sqrtf(v1) / bc2s + a.eps;
                        p = p - lr_c * (m1 / denom);
                        ht_smem[sm.bs[l] + s * HT_RB + tid] = p;
                        L.b[r] = p;
                        L.mb[r] = m1;
                        L.vb[r] = v1;
                    } else {
                        if (L.gb) L.gb[r] = graw;
                        if (L.qb) L.qb[r] += graw * graw * a.fisher_scale;
                    }
                }
            }
        }
        __syncthreads();
        HT_STAMP(12);
    }
}

}  // namespace ht
}  // namespace ac
