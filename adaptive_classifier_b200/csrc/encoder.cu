// encoder.cu -- stage E: BERT / RoBERTa post-LN encoder forward -> unit-norm CLS rows.
//
// Replaces `self.model(**inputs).last_hidden_state[:, 0, :]` + F.normalize at
// /root/reference/src/adaptive_classifier/classifier.py:1271-1275 (HF BertModel.forward:
// embeddings modeling_bert.py:53-113, self-attention :143-207, output+LN :287-298, FFN :330-356).
//
// Dense projections run on the tcgen05 GEMM of gemm_tc.cuh (kind::tf32, operands RNE-rounded to tf32 by
// the producing kernel, fp32 accumulation in TMEM) with fused bias / exact-erf GELU / residual epilogues.
// Attention is one CTA per (sequence, head): QK^T and PV as tcgen05 MMAs with the score tile and the
// output tile in TMEM and a thread-per-query-row softmax in between (S <= 128, head_dim 64).
// LayerNorm keeps the fp32 residual stream and also emits the tf32-rounded copy the next GEMM reads.
#include "gemm_tc.cuh"
#include <math_constants.h>
#include <vector>

namespace ac {

// ------------------------------------------------------------------------------------------------
// fused epilogue of the encoder linears
// ------------------------------------------------------------------------------------------------
struct EpiLinear {
    const float *bias;       // [N] nullable
    const float *residual;   // [M,N] nullable (mode 2)
    float *Y;                // [M,N]
    int M, N;
    int mode;                // 0 bias, 1 bias+GELU(erf), 2 bias+residual
    int round_out;           // round result to tf32 (RNE): the result feeds another tcgen05 GEMM

    struct State {};
    __device__ __forceinline__ void begin_cta(State &, int, int) const {}
    __device__ __forceinline__ void end_cta(State &, int, int) const {}

    __device__ __forceinline__ float apply(float a, float b, float r) const {
        float y = a + b;
        if (mode == 1) y = 0.5f * y * (1.f + erff(y * 0.70710678118654752440f));
        if (mode == 2) y += r;
        if (round_out) y = round_tf32(y);
        return y;
    }

    // v[] = this thread's row (TMEM lane), 32 consecutive columns.  Transposed through the warp's staging tile so
    // that every global access is a full 128-byte row segment: lane (r4 = lane/8, c4 = lane%8) handles rows
    // r4 + 4*i and the 16-byte column group c4 -> one warp instruction touches 4 rows x 128 B.
    __device__ __forceinline__ void tile(State &, const GemmTileInfo &ti, int row, int col0, const float (&v)[32],
                                         float *stage, int lane) const {
        (void)row;
        const int row_base = ti.m0 + ((threadIdx.x >> 5) & 3) * 32;       // first row of this warp's TMEM quarter
        if (row_base >= M || col0 >= N) return;                             // warp-uniform
        float4 *srow = reinterpret_cast<float4 *>(stage + lane * GEMM_EPI_STAGE_STRIDE);
#pragma unroll
        for (int j = 0; j < 8; ++j) srow[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const int r4 = lane >> 3, c4 = lane & 7;
        const int col = col0 + 4 * c4;
        if (col + 4 <= N) {
            const float4 b4 = bias ? __ldg(reinterpret_cast<const float4 *>(bias + col)) : make_float4(0, 0, 0, 0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = r4 + 4 * i;
                const int grow = row_base + rr;
                if (grow < M) {
                    const float4 a = *reinterpret_cast<const float4 *>(stage + rr * GEMM_EPI_STAGE_STRIDE + 4 * c4);
                    float4 r = make_float4(0, 0, 0, 0);
                    if (mode == 2) r = *reinterpret_cast<const float4 *>(residual + static_cast<int64_t>(grow) * N + col);
                    float4 o;
                    o.x = apply(a.x, b4.x, r.x);
                    o.y = apply(a.y, b4.y, r.y);
                    o.z = apply(a.z, b4.z, r.z);
                    o.w = apply(a.w, b4.w, r.w);
                    *reinterpret_cast<float4 *>(Y + static_cast<int64_t>(grow) * N + col) = o;
                }
            }
        } else {
            for (int i = 0; i < 8; ++i) {
                const int rr = r4 + 4 * i;
                const int grow = row_base + rr;
                if (grow >= M) continue;
                for (int t = 0; t < 4; ++t) {
                    const int cc = col + t;
                    if (cc < N)
                        Y[static_cast<int64_t>(grow) * N + cc] =
                            apply(stage[rr * GEMM_EPI_STAGE_STRIDE + 4 * c4 + t], bias ? bias[cc] : 0.f,
                                  mode == 2 ? residual[static_cast<int64_t>(grow) * N + cc] : 0.f);
                }
            }
        }
        __syncwarp();   // staging tile is rewritten by the next chunk
    }
};

// ------------------------------------------------------------------------------------------------
// elementwise / normalisation kernels (one warp per row, float4 lanes; H % 128 == 0, H <= 1024)
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 8;

__device__ __forceinline__ void ln_row(float4 (&x)[LN_MAXV], int nv, int H, const float *__restrict__ w,
                                       const float *__restrict__ b, float eps, int lane, float *out_full,
                                       float *out_round) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
        if (i < nv) s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    const float mean = warp_sum(s) / static_cast<float>(H);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
        if (i < nv) {
            const float a = x[i].x - mean, c = x[i].y - mean, d = x[i].z - mean, e = x[i].w - mean;
            q += (a * a + c * c) + (d * d + e * e);
        }
    const float var = warp_sum(q) / static_cast<float>(H);
    const float rstd = 1.f / sqrtf(var + eps);
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
        if (i < nv) {
            const int col = (lane + 32 * i) * 4;
            const float4 w4 = __ldg(reinterpret_cast<const float4 *>(w + col));
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(b + col));
            float4 o;
            o.x = (x[i].x - mean) * rstd * w4.x + b4.x;
            o.y = (x[i].y - mean) * rstd * w4.y + b4.y;
            o.z = (x[i].z - mean) * rstd * w4.z + b4.z;
            o.w = (x[i].w - mean) * rstd * w4.w + b4.w;
            if (out_full) *reinterpret_cast<float4 *>(out_full + col) = o;
            if (out_round) {
                float4 r;
                r.x = round_tf32(o.x); r.y = round_tf32(o.y); r.z = round_tf32(o.z); r.w = round_tf32(o.w);
                *reinterpret_cast<float4 *>(out_round + col) = r;
            }
        }
}

__global__ void layernorm_kernel(const float *__restrict__ in, const float *__restrict__ w, const float *__restrict__ b,
                                 float eps, int rows, int H, float *__restrict__ out_full,
                                 float *__restrict__ out_round) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int nv = H / 128;
    float4 x[LN_MAXV];
    const float *src = in + static_cast<int64_t>(row) * H;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
        if (i < nv) x[i] = *reinterpret_cast<const float4 *>(src + (lane + 32 * i) * 4);
    ln_row(x, nv, H, w, b, eps, lane, out_full ? out_full + static_cast<int64_t>(row) * H : nullptr,
           out_round ? out_round + static_cast<int64_t>(row) * H : nullptr);
}

// modeling_bert.py:53-113 / modeling_roberta.py:146-159: (word + type) + position -> LayerNorm
__global__ void embed_ln_kernel(const int32_t *__restrict__ ids, const int32_t *__restrict__ type_ids,
                                const float *__restrict__ word, const float *__restrict__ pos,
                                const float *__restrict__ type, const float *__restrict__ w,
                                const float *__restrict__ b, float eps, int B, int S, int H, int arch, int pad_idx,
                                int vocab, int max_pos, int type_vocab, float *__restrict__ out_full,
                                float *__restrict__ out_round) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B * S) return;
    const int bq = row / S, s = row % S;
    int id = ids[row];
    id = min(max(id, 0), vocab - 1);
    int tt = type_ids ? type_ids[row] : 0;
    tt = min(max(tt, 0), type_vocab - 1);
    int p = s;
    if (arch == AC_ARCH_ROBERTA) {
        // position = cumsum(ids != pad)[s] * (id != pad) + pad_idx
        int cnt = 0;
        for (int j = lane; j <= s; j += 32) cnt += (ids[bq * S + j] != pad_idx) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        p = (id != pad_idx) ? cnt + pad_idx : pad_idx;
    }
    p = min(p, max_pos - 1);
    const int nv = H / 128;
    float4 x[LN_MAXV];
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
        if (i < nv) {
            const int col = (lane + 32 * i) * 4;
            const float4 a = __ldg(reinterpret_cast<const float4 *>(word + static_cast<int64_t>(id) * H + col));
            const float4 t = __ldg(reinterpret_cast<const float4 *>(type + static_cast<int64_t>(tt) * H + col));
            const float4 q = __ldg(reinterpret_cast<const float4 *>(pos + static_cast<int64_t>(p) * H + col));
            x[i].x = (a.x + t.x) + q.x;
            x[i].y = (a.y + t.y) + q.y;
            x[i].z = (a.z + t.z) + q.z;
            x[i].w = (a.w + t.w) + q.w;
        }
    ln_row(x, nv, H, w, b, eps, lane, out_full + static_cast<int64_t>(row) * H, out_round + static_cast<int64_t>(row) * H);
}

// classifier.py:1272,1275: CLS row -> x / max(||x||_2, 1e-12)
__global__ void cls_normalize_kernel(const float *__restrict__ x, int B, int S, int H, float *__restrict__ out) {
    const int bq = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (bq >= B) return;
    const float *src = x + static_cast<int64_t>(bq) * S * H;
    float s = 0.f;
    for (int i = lane; i < H; i += 32) s = fmaf(src[i], src[i], s);
    const float nrm = fmaxf(sqrtf(warp_sum(s)), 1e-12f);
    for (int i = lane; i < H; i += 32) out[static_cast<int64_t>(bq) * H + i] = src[i] / nrm;
}

__global__ void round_copy_kernel(const float *__restrict__ in, float *__restrict__ out, int64_t n, int do_round) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = do_round ? round_tf32(in[i]) : in[i];
}

// ------------------------------------------------------------------------------------------------
// attention: one CTA (128 threads) per (sequence b, head h); S <= 128, head_dim == 64.
//   scores[128x128] = Q K^T        8 x tcgen05.mma kind::tf32 (M128 N128 K8), accumulator TMEM cols [0,128)
//   P = exp(scale*(s - max)) masked  thread = query row, tcgen05.ld 32x32b; P -> smem (swizzled, tf32-rounded)
//   out[128x64] = P V              16 x tcgen05.mma (M128 N64 K8), accumulator TMEM cols [128,192)
//   ctx[row, h*64 + :] = out / rowsum (rounded to tf32: it is the A operand of the output projection)
// smem: Q|K tiles (2 x 32 KB, TMA, 128B swizzle) reused for P (64 KB); V^T (32 KB) staged by the threads into the
// K-major 128B-swizzled layout (all 16 global loads of a thread are issued before the first shared store).
// (An MN-major descriptor for V straight from TMA was tried in round 1 and produced wrong results; see DESIGN.md.)
// ------------------------------------------------------------------------------------------------
constexpr int ATT_THREADS = 128;
constexpr int ATT_SMEM = 64 * 1024 + 32 * 1024 + 1024 /*align*/ + 64;
constexpr int ATT_TMEM_COLS = 256;

__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const float *__restrict__ qkv,
                 const int32_t *__restrict__ mask, int B, int S, int heads, int H, float *__restrict__ ctx) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem;                    // 2 slabs x [128 rows x 128 B]
    uint8_t *sK = smem + 32 * 1024;        // 2 slabs
    uint8_t *sP = smem;                    // 4 slabs x [128 rows x 128 B]   (after QK^T retired)
    uint8_t *sVt = smem + 64 * 1024;       // 4 slabs x [64 rows (d) x 128 B (32 keys)]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 96 * 1024);
    uint64_t *bar_load = bars, *bar_s = bars + 1, *bar_o = bars + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int64_t row0 = static_cast<int64_t>(b) * S;
    const int ld = 3 * H;

    if (tid == 0) {
        tma_prefetch_desc(&tmap_qkv);
        mbar_init(bar_load, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_o, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, ATT_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (tid == 0) {
        mbar_arrive_expect_tx(bar_load, 64 * 1024);
        const int r = static_cast<int>(row0);
        tma_load_2d(sQ, &tmap_qkv, bar_load, h * 64, r);
        tma_load_2d(sQ + 16 * 1024, &tmap_qkv, bar_load, h * 64 + 32, r);
        tma_load_2d(sK, &tmap_qkv, bar_load, H + h * 64, r);
        tma_load_2d(sK + 16 * 1024, &tmap_qkv, bar_load, H + h * 64 + 32, r);
    }

    // stage V^T (K-major B operand: row = d, contiguous = key) with the 128B swizzle applied by hand
    {
        const float *vbase = qkv + row0 * ld + 2 * H + h * 64;
        float4 vreg[16];
#pragma unroll
        for (int it = 0; it < 16; ++it) {          // all loads in flight first
            const int e = tid + it * ATT_THREADS;
            const int key = e >> 4, d4 = e & 15;
            vreg[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (key < S) vreg[it] = __ldg(reinterpret_cast<const float4 *>(vbase + static_cast<int64_t>(key) * ld + d4 * 4));
        }
        const uint32_t sv_base = smem_u32(sVt);
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            const int e = tid + it * ATT_THREADS;
            const int key = e >> 4, d4 = e & 15;
            const int slab = key >> 5, c = (key & 31) >> 2, wi = key & 3;
            const float vv[4] = {vreg[it].x, vreg[it].y, vreg[it].z, vreg[it].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int d = d4 * 4 + i;
                const uint32_t off = slab * 8192 + (d >> 3) * 1024 + (d & 7) * 128 + ((c ^ (d & 7)) << 4) + wi * 4;
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(sv_base + off), "f"(vv[i]) : "memory");
            }
        }
    }

    // ---- S = Q K^T
    if (tid == 0) {
        mbar_wait_guarded(bar_load, 0);
        tc_fence_after();
        constexpr uint32_t idesc_s = umma_idesc(2, 128, 128);
#pragma unroll
        for (int slab = 0; slab < 2; ++slab) {
            const uint64_t a = umma_desc_sw128(smem_u32(sQ + slab * 16 * 1024));
            const uint64_t bdesc = umma_desc_sw128(smem_u32(sK + slab * 16 * 1024));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, a + 2 * k, bdesc + 2 * k, idesc_s, (slab | k) != 0);
        }
        tc_commit(bar_s);
    }
    __syncwarp();
    mbar_wait_guarded(bar_s, 0);
    tc_fence_after();

    // ---- softmax: thread = query row (TMEM lane), two passes over the 128 score columns
    const int qrow = warp * 32 + lane;
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int32_t *mrow = mask ? mask + row0 : nullptr;
    const float scale_log2 = rsqrtf(64.f) * 1.44269504088896340736f;
    float mx = -CUDART_INF_F;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(t_s + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int key = c + j;
            const bool valid = (key < S) && (!mrow || mrow[key] != 0);
            if (valid) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
    }
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(t_s + c, r);
        tmem_ld_wait();
        float p[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int key = c + j;
            const bool valid = (key < S) && (!mrow || mrow[key] != 0);
            const float e = valid ? exp2f((__uint_as_float(r[j]) - mx) * scale_log2) : 0.f;
            sum += e;
            p[j] = round_tf32(e);
        }
        // slab (c/32), row qrow: 8 x 16-byte chunks at the swizzled positions
        const uint32_t prow = smem_u32(sP) + (c >> 5) * 16384 + (qrow >> 3) * 1024 + (qrow & 7) * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(prow + ((ch ^ (qrow & 7)) << 4)), "f"(p[4 * ch]),
                         "f"(p[4 * ch + 1]), "f"(p[4 * ch + 2]), "f"(p[4 * ch + 3])
                         : "memory");
        }
    }
    // generic-proxy smem writes (P, V^T) -> visible to the tensor-core (async) proxy
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- O = P V
    if (tid == 0) {
        constexpr uint32_t idesc_o = umma_idesc(2, 128, 64);
#pragma unroll
        for (int slab = 0; slab < 4; ++slab) {
            const uint64_t a = umma_desc_sw128(smem_u32(sP + slab * 16384));
            const uint64_t bdesc = umma_desc_sw128(smem_u32(sVt + slab * 8192));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_tf32(tmem_base + 128, a + 2 * k, bdesc + 2 * k, idesc_o, (slab | k) != 0);
        }
        tc_commit(bar_o);
    }
    __syncwarp();
    mbar_wait_guarded(bar_o, 0);
    tc_fence_after();

    const float inv = (sum > 0.f) ? 1.f / sum : 0.f;
#pragma unroll 1
    for (int c = 0; c < 64; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(t_s + 128 + c, r);
        tmem_ld_wait();
        if (qrow < S) {
            float *dst = ctx + (row0 + qrow) * H + h * 64 + c;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 o;
                o.x = round_tf32(__uint_as_float(r[j]) * inv);
                o.y = round_tf32(__uint_as_float(r[j + 1]) * inv);
                o.z = round_tf32(__uint_as_float(r[j + 2]) * inv);
                o.w = round_tf32(__uint_as_float(r[j + 3]) * inv);
                *reinterpret_cast<float4 *>(dst + j) = o;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, ATT_TMEM_COLS);
    }
}

}  // namespace ac

// ================================================================================================
// encoder handle
// ================================================================================================
using namespace ac;

struct ac_encoder {
    ac_encoder_config cfg;
    // packed weights (device): tf32-rounded GEMM operands, fp32 everything else
    float *word = nullptr, *pos = nullptr, *type = nullptr, *emb_ln_w = nullptr, *emb_ln_b = nullptr;
    std::vector<float *> wqkv, bqkv, wo, bo, ln1w, ln1b, w1, b1, w2, b2, ln2w, ln2b;
    // activations
    float *x = nullptr, *xr = nullptr, *qkv = nullptr, *ctx = nullptr, *tmp = nullptr, *ffn = nullptr;
    // cached TMA descriptors
    CUtensorMap m_xr, m_ctx, m_ffn, m_qkv_att;
    std::vector<CUtensorMap> m_wqkv, m_wo, m_w1, m_w2;
    std::vector<void *> allocs;
    int last_B = 0, last_S = 0;
};

static int dev_alloc(ac_encoder *e, float **p, size_t floats) {
    void *q = nullptr;
    AC_CUDA(cudaMalloc(&q, floats * sizeof(float)));
    e->allocs.push_back(q);
    *p = static_cast<float *>(q);
    return AC_OK;
}

static int pack(ac_encoder *e, float **dst, const float *src, size_t n, bool round) {
    int rc = dev_alloc(e, dst, n);
    if (rc) return rc;
    round_copy_kernel<<<256, 256>>>(src, *dst, static_cast<int64_t>(n), round ? 1 : 0);
    AC_LAUNCH_CHECK();
    return AC_OK;
}

extern "C" int ac_encoder_destroy(ac_encoder *enc) {
    if (!enc) return AC_OK;
    for (void *p : enc->allocs) cudaFree(p);
    delete enc;
    return AC_OK;
}

extern "C" int ac_encoder_create(const ac_encoder_config *cfg, const ac_encoder_weights *w, ac_encoder **out) {
    AC_REQUIRE(cfg && w && out, "ac_encoder_create: null argument");
    AC_REQUIRE(cfg->precision == AC_PREC_TF32, "ac_encoder_create: only AC_PREC_TF32 is implemented");
    AC_REQUIRE(cfg->hidden % 128 == 0 && cfg->hidden <= 1024, "ac_encoder_create: hidden=%d must be a multiple of 128, <= 1024", cfg->hidden);
    AC_REQUIRE(cfg->heads > 0 && cfg->hidden / cfg->heads == 64 && cfg->hidden % cfg->heads == 0,
               "ac_encoder_create: head_dim must be 64 (hidden=%d heads=%d)", cfg->hidden, cfg->heads);
    AC_REQUIRE(cfg->intermediate % 32 == 0 && cfg->layers > 0 && cfg->max_tokens > 0, "ac_encoder_create: bad dims");
    int rc = ac_device_check();
    if (rc) return rc;
    ac_encoder *e = new ac_encoder();
    e->cfg = *cfg;
    const int H = cfg->hidden, I = cfg->intermediate, L = cfg->layers;
    const size_t T = static_cast<size_t>((cfg->max_tokens + 127) / 128 * 128);
#define TRY(x) do { rc = (x); if (rc) { ac_encoder_destroy(e); return rc; } } while (0)
    TRY(pack(e, &e->word, w->word_emb, static_cast<size_t>(cfg->vocab) * H, false));
    TRY(pack(e, &e->pos, w->pos_emb, static_cast<size_t>(cfg->max_pos) * H, false));
    TRY(pack(e, &e->type, w->type_emb, static_cast<size_t>(cfg->type_vocab) * H, false));
    TRY(pack(e, &e->emb_ln_w, w->emb_ln_w, H, false));
    TRY(pack(e, &e->emb_ln_b, w->emb_ln_b, H, false));
    auto resize_all = [&](std::vector<float *> &v) { v.assign(L, nullptr); };
    resize_all(e->wqkv); resize_all(e->bqkv); resize_all(e->wo); resize_all(e->bo); resize_all(e->ln1w); resize_all(e->ln1b);
    resize_all(e->w1); resize_all(e->b1); resize_all(e->w2); resize_all(e->b2); resize_all(e->ln2w); resize_all(e->ln2b);
    const size_t HH = static_cast<size_t>(H) * H;
    for (int l = 0; l < L; ++l) {
        // fused QKV operand [3H, H] and bias [3H]
        TRY(dev_alloc(e, &e->wqkv[l], 3 * HH));
        TRY(dev_alloc(e, &e->bqkv[l], 3 * static_cast<size_t>(H)));
        const float *ws[3] = {w->q_w[l], w->k_w[l], w->v_w[l]};
        const float *bs[3] = {w->q_b[l], w->k_b[l], w->v_b[l]};
        for (int j = 0; j < 3; ++j) {
            round_copy_kernel<<<256, 256>>>(ws[j], e->wqkv[l] + j * HH, static_cast<int64_t>(HH), 1);
            round_copy_kernel<<<8, 256>>>(bs[j], e->bqkv[l] + j * H, H, 0);
        }
        TRY(pack(e, &e->wo[l], w->ao_w[l], HH, true));
        TRY(pack(e, &e->bo[l], w->ao_b[l], H, false));
        TRY(pack(e, &e->ln1w[l], w->ao_ln_w[l], H, false));
        TRY(pack(e, &e->ln1b[l], w->ao_ln_b[l], H, false));
        TRY(pack(e, &e->w1[l], w->ff1_w[l], static_cast<size_t>(I) * H, true));
        TRY(pack(e, &e->b1[l], w->ff1_b[l], I, false));
        TRY(pack(e, &e->w2[l], w->ff2_w[l], static_cast<size_t>(H) * I, true));
        TRY(pack(e, &e->b2[l], w->ff2_b[l], H, false));
        TRY(pack(e, &e->ln2w[l], w->out_ln_w[l], H, false));
        TRY(pack(e, &e->ln2b[l], w->out_ln_b[l], H, false));
    }
    TRY(dev_alloc(e, &e->x, T * H));
    TRY(dev_alloc(e, &e->xr, T * H));
    TRY(dev_alloc(e, &e->qkv, T * 3 * H));
    TRY(dev_alloc(e, &e->ctx, T * H));
    TRY(dev_alloc(e, &e->tmp, T * H));
    TRY(dev_alloc(e, &e->ffn, T * I));
    TRY(check_cuda(cudaMemset(e->qkv, 0, T * 3 * H * sizeof(float)), "memset qkv"));
    TRY(check_cuda(cudaMemset(e->xr, 0, T * H * sizeof(float)), "memset xr"));
    TRY(check_cuda(cudaMemset(e->ctx, 0, T * H * sizeof(float)), "memset ctx"));
    TRY(check_cuda(cudaMemset(e->ffn, 0, T * I * sizeof(float)), "memset ffn"));
    // TMA descriptors
    TRY(make_tmap_2d(&e->m_xr, e->xr, 4, T, H, static_cast<uint64_t>(H) * 4, GEMM_BLOCK_M, GEMM_BLOCK_K));
    TRY(make_tmap_2d(&e->m_ctx, e->ctx, 4, T, H, static_cast<uint64_t>(H) * 4, GEMM_BLOCK_M, GEMM_BLOCK_K));
    TRY(make_tmap_2d(&e->m_ffn, e->ffn, 4, T, I, static_cast<uint64_t>(I) * 4, GEMM_BLOCK_M, GEMM_BLOCK_K));
    TRY(make_tmap_2d(&e->m_qkv_att, e->qkv, 4, T, 3 * H, static_cast<uint64_t>(3 * H) * 4, 128, 32));
    e->m_wqkv.resize(L); e->m_wo.resize(L); e->m_w1.resize(L); e->m_w2.resize(L);
    for (int l = 0; l < L; ++l) {
        TRY(make_tmap_2d(&e->m_wqkv[l], e->wqkv[l], 4, 3 * H, H, static_cast<uint64_t>(H) * 4, GEMM_BLOCK_N, GEMM_BLOCK_K));
        TRY(make_tmap_2d(&e->m_wo[l], e->wo[l], 4, H, H, static_cast<uint64_t>(H) * 4, GEMM_BLOCK_N, GEMM_BLOCK_K));
        TRY(make_tmap_2d(&e->m_w1[l], e->w1[l], 4, I, H, static_cast<uint64_t>(H) * 4, GEMM_BLOCK_N, GEMM_BLOCK_K));
        TRY(make_tmap_2d(&e->m_w2[l], e->w2[l], 4, H, I, static_cast<uint64_t>(I) * 4, GEMM_BLOCK_N, GEMM_BLOCK_K));
    }
    TRY(check_cuda(cudaDeviceSynchronize(), "encoder_create sync"));
#undef TRY
    *out = e;
    return AC_OK;
}

extern "C" int ac_encoder_forward_cls(ac_encoder *e, const int32_t *ids, const int32_t *mask, const int32_t *type_ids,
                                      int B, int S, float *out_unit_cls, ac_stream_t stream) {
    AC_REQUIRE(e && ids && out_unit_cls, "ac_encoder_forward_cls: null argument");
    AC_REQUIRE(B > 0 && S > 0, "ac_encoder_forward_cls: B=%d S=%d", B, S);
    if (S > 128) {
        set_error("ac_encoder_forward_cls: S=%d > 128 is not implemented yet (attention tile)", S);
        return AC_E_UNSUPPORTED;
    }
    AC_REQUIRE(static_cast<int64_t>(B) * S <= e->cfg.max_tokens, "ac_encoder_forward_cls: B*S=%lld exceeds max_tokens=%d",
               static_cast<long long>(B) * S, e->cfg.max_tokens);
    AC_REQUIRE(S <= e->cfg.max_pos, "ac_encoder_forward_cls: S exceeds max_position_embeddings");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const ac_encoder_config &c = e->cfg;
    const int H = c.hidden, I = c.intermediate, M = B * S;
    const int wpb = 8;
    const int row_blocks = (M + wpb - 1) / wpb;
    int rc;

    embed_ln_kernel<<<row_blocks, wpb * 32, 0, s>>>(ids, type_ids, e->word, e->pos, e->type, e->emb_ln_w, e->emb_ln_b,
                                                    c.ln_eps, B, S, H, c.arch, c.pad_idx, c.vocab, c.max_pos,
                                                    c.type_vocab, e->x, e->xr);
    AC_LAUNCH_CHECK();
    static bool att_attr = false;
    if (!att_attr) {
        AC_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        att_attr = true;
    }
    for (int l = 0; l < c.layers; ++l) {
        EpiLinear eq{e->bqkv[l], nullptr, e->qkv, M, 3 * H, 0, 1};
        if ((rc = launch_gemm_tf32(e->m_xr, e->m_wqkv[l], M, 3 * H, H, eq, s))) return rc;
        {
            // algorithmic flops of softmax(QK^T)V at the true sequence length (the 128-wide tile does more)
            const int slot = prof_begin(PROF_ATTENTION, 4.0 * B * c.heads * static_cast<double>(S) * S * 64, 0.0, s);
            attention_kernel<<<B * c.heads, ATT_THREADS, ATT_SMEM, s>>>(e->m_qkv_att, e->qkv, mask, B, S, c.heads, H, e->ctx);
            prof_end(slot, s);
        }
        AC_LAUNCH_CHECK();
        EpiLinear eo{e->bo[l], e->x, e->tmp, M, H, 2, 0};
        if ((rc = launch_gemm_tf32(e->m_ctx, e->m_wo[l], M, H, H, eo, s))) return rc;
        layernorm_kernel<<<row_blocks, wpb * 32, 0, s>>>(e->tmp, e->ln1w[l], e->ln1b[l], c.ln_eps, M, H, e->x, e->xr);
        AC_LAUNCH_CHECK();
        EpiLinear e1{e->b1[l], nullptr, e->ffn, M, I, 1, 1};
        if ((rc = launch_gemm_tf32(e->m_xr, e->m_w1[l], M, I, H, e1, s))) return rc;
        EpiLinear e2{e->b2[l], e->x, e->tmp, M, H, 2, 0};
        if ((rc = launch_gemm_tf32(e->m_ffn, e->m_w2[l], M, H, I, e2, s))) return rc;
        layernorm_kernel<<<row_blocks, wpb * 32, 0, s>>>(e->tmp, e->ln2w[l], e->ln2b[l], c.ln_eps, M, H, e->x, e->xr);
        AC_LAUNCH_CHECK();
    }
    cls_normalize_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, s>>>(e->x, B, S, H, out_unit_cls);
    AC_LAUNCH_CHECK();
    e->last_B = B;
    e->last_S = S;
    return AC_OK;
}

extern "C" int ac_encoder_last_hidden(ac_encoder *e, float *out, int64_t n_floats, ac_stream_t stream) {
    AC_REQUIRE(e && out, "ac_encoder_last_hidden: null argument");
    const int64_t have = static_cast<int64_t>(e->last_B) * e->last_S * e->cfg.hidden;
    AC_REQUIRE(n_floats <= have, "ac_encoder_last_hidden: asked %lld floats, have %lld", (long long)n_floats, (long long)have);
    AC_CUDA(cudaMemcpyAsync(out, e->x, n_floats * sizeof(float), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return AC_OK;
}

extern "C" int ac_linear_tc(const float *X, const float *W, const float *bias, const float *residual, float *Y, int M,
                            int N, int K, int epi, int round_out, ac_stream_t stream) {
    AC_REQUIRE(X && W && Y && M > 0 && N > 0 && K > 0, "ac_linear_tc: bad arguments");
    AC_REQUIRE(K % 4 == 0 && N % 4 == 0, "ac_linear_tc: K and N must be multiples of 4 (16-byte rows)");
    AC_REQUIRE(epi >= 0 && epi <= 2 && (epi != 2 || residual), "ac_linear_tc: bad epilogue");
    int rc = ac_device_check();
    if (rc) return rc;
    CUtensorMap ta, tb;
    if ((rc = make_tmap_2d(&ta, X, 4, M, K, static_cast<uint64_t>(K) * 4, GEMM_BLOCK_M, GEMM_BLOCK_K))) return rc;
    if ((rc = make_tmap_2d(&tb, W, 4, N, K, static_cast<uint64_t>(K) * 4, GEMM_BLOCK_N, GEMM_BLOCK_K))) return rc;
    EpiLinear e{bias, residual, Y, M, N, epi, round_out};
    return launch_gemm_tf32(ta, tb, M, N, K, e, static_cast<cudaStream_t>(stream));
}
