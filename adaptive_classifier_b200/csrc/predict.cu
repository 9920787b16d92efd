// predict.cu -- the glue of predict_batch() on the device and the end-to-end pipeline handle.
//
//   ac_proto_class_scores   memory.py:117-134 generalised to many rows per class: the k
//                           nearest ROWS are mapped to classes, a class keeps its nearest row, scores =
//                           softmax(exp(-d)) over the distinct classes returned.  With one row per class
//                           (the reference's own usage) this is exactly memory.py:117-134.
//   ac_topk_desc            classifier.py:1347-1350 (torch.topk of the head probabilities)
//   ac_blend_topk           classifier.py:1358-1384: 0.7 * prototype score + 0.3 * head probability per label,
//                           stable descending sort, normalise by the sum, keep k
//   ac_pipeline_*           E -> K -> H -> blend with device or host (pinned) buffers at the boundary
#include "common.cuh"
#include <math_constants.h>

namespace ac {
size_t topk_select_workspace(int B, int64_t L, int k);
int topk_select(const float *d, const int64_t *idx, int B, int64_t L, int64_t in_stride, int64_t id_offset, int k,
                float *out_d, int64_t *out_i, void *ws, size_t ws_bytes, cudaStream_t stream, const float *row_gate = nullptr);

constexpr int BLEND_MAX_K = 32;

// one thread per query (k <= 32): dedupe by class keeping the first (nearest) row
__global__ void proto_class_scores_kernel(const float *__restrict__ d, const int64_t *__restrict__ idx,
                                          const int32_t *__restrict__ row_class, int B, int k,
                                          int32_t *__restrict__ out_cls, float *__restrict__ out_score) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int32_t cls[BLEND_MAX_K];
    float sc[BLEND_MAX_K];
    int n = 0;
    for (int j = 0; j < k; ++j) {
        const int64_t id = idx[static_cast<int64_t>(b) * k + j];
        if (id < 0) continue;
        const int32_t c = row_class ? row_class[id] : static_cast<int32_t>(id);
        bool seen = false;
        for (int t = 0; t < n; ++t) seen |= (cls[t] == c);
        if (seen) continue;
        cls[n] = c;
        sc[n] = expf(-d[static_cast<int64_t>(b) * k + j]);
        ++n;
    }
    float mx = -CUDART_INF_F, sum = 0.f;
    for (int t = 0; t < n; ++t) mx = fmaxf(mx, sc[t]);
    for (int t = 0; t < n; ++t) { sc[t] = expf(sc[t] - mx); sum += sc[t]; }
    for (int j = 0; j < k; ++j) {
        out_cls[static_cast<int64_t>(b) * k + j] = j < n ? cls[j] : -1;
        out_score[static_cast<int64_t>(b) * k + j] = j < n ? sc[j] / sum : 0.f;
    }
}

// the same for any k <= 1024 (predict(): k = num_classes, classifier.py:424-425): one CTA per query.  A class keeps its nearest
// row = the entry of lowest rank among its rows (atomicMin of the rank per class in shared memory); the kept entries stay in
// rank order (compacted by a block scan), softmax(exp(-d)) over them.  classes: ids < n_classes <= PCS_MAX_CLASSES.
constexpr int PCS_THREADS = 256;
constexpr int PCS_MAX_CLASSES = 4096;
__global__ void __launch_bounds__(PCS_THREADS)
proto_class_scores_block_kernel(const float *__restrict__ d, const int64_t *__restrict__ idx, const int32_t *__restrict__ row_class,
                                int k, int n_classes, int32_t *__restrict__ out_cls, float *__restrict__ out_score) {
    __shared__ int first_rank[PCS_MAX_CLASSES];
    __shared__ float red[PCS_THREADS];
    __shared__ int scan[PCS_THREADS];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float *db = d + static_cast<int64_t>(b) * k;
    const int64_t *ib = idx + static_cast<int64_t>(b) * k;
    int32_t *oc = out_cls + static_cast<int64_t>(b) * k;
    float *os = out_score + static_cast<int64_t>(b) * k;
    for (int c = tid; c < n_classes; c += PCS_THREADS) first_rank[c] = 0x7fffffff;
    __syncthreads();
    constexpr int PER = 4;                                    // k <= 1024 = 256 threads x 4 consecutive ranks
    int cls[PER];
    for (int i = 0; i < PER; ++i) {
        const int j = tid * PER + i;
        cls[i] = -1;
        if (j < k) {
            const int64_t id = ib[j];
            if (id >= 0) {
                const int c = row_class ? row_class[id] : static_cast<int>(id);
                if (c >= 0 && c < n_classes) { cls[i] = c; atomicMin(&first_rank[c], j); }
            }
        }
    }
    __syncthreads();
    float e[PER];
    int keep = 0;
    float mx = -CUDART_INF_F;
    for (int i = 0; i < PER; ++i) {
        const int j = tid * PER + i;
        const bool kept = cls[i] >= 0 && first_rank[cls[i]] == j;
        e[i] = kept ? expf(-db[j]) : -CUDART_INF_F;
        if (!kept) cls[i] = -1;
        keep += kept ? 1 : 0;
        mx = fmaxf(mx, e[i]);
    }
    red[tid] = mx;
    scan[tid] = keep;
    __syncthreads();
    for (int s2 = PCS_THREADS / 2; s2 > 0; s2 >>= 1) {
        if (tid < s2) red[tid] = fmaxf(red[tid], red[tid + s2]);
        __syncthreads();
    }
    mx = red[0];
    __syncthreads();
    // exclusive scan of the kept counts (rank order) -> output positions
    for (int off = 1; off < PCS_THREADS; off <<= 1) {
        const int v = tid >= off ? scan[tid - off] : 0;
        __syncthreads();
        scan[tid] += v;
        __syncthreads();
    }
    const int total = scan[PCS_THREADS - 1];
    int pos = scan[tid] - keep;
    float sum = 0.f;
    for (int i = 0; i < PER; ++i)
        if (cls[i] >= 0) { e[i] = expf(e[i] - mx); sum += e[i]; }
    red[tid] = sum;
    __syncthreads();
    for (int s2 = PCS_THREADS / 2; s2 > 0; s2 >>= 1) {
        if (tid < s2) red[tid] += red[tid + s2];
        __syncthreads();
    }
    sum = red[0];
    for (int i = 0; i < PER; ++i)
        if (cls[i] >= 0) { oc[pos] = cls[i]; os[pos] = e[i] / sum; ++pos; }
    for (int j = total + tid; j < k; j += PCS_THREADS) { oc[j] = -1; os[j] = 0.f; }
}

// predict() blend over ALL classes (classifier.py:446-480): combined[c] = proto[c] * wp[c] + head[c] * wh[c] with per-class
// weights (training_history < 10 -> 0.3 / 0.7, else 0.7 / 0.3), normalised by the sum, top kout by value; ties keep the
// reference's insertion order (prototype entries by rank, then head-only classes -- by class id here).
// One CTA per query; the scores of all classes live in shared memory.
constexpr int BD_THREADS = 256;
__global__ void __launch_bounds__(BD_THREADS)
blend_dense_kernel(const int32_t *__restrict__ p_cls, const float *__restrict__ p_score, int kp, const float *__restrict__ probs,
                   int C, const float *__restrict__ w_proto, const float *__restrict__ w_head, int kout,
                   int32_t *__restrict__ out_cls, float *__restrict__ out_score) {
    __shared__ float comb[PCS_MAX_CLASSES];
    __shared__ int order[PCS_MAX_CLASSES];
    __shared__ float redv[BD_THREADS];
    __shared__ int redi[BD_THREADS];
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int c = tid; c < C; c += BD_THREADS) {
        comb[c] = probs ? probs[static_cast<int64_t>(b) * C + c] * w_head[c] : 0.f;
        order[c] = kp + c;                                    // head-only classes come after every prototype entry
    }
    __syncthreads();
    for (int j = tid; j < kp; j += BD_THREADS) {
        const int c = p_cls[static_cast<int64_t>(b) * kp + j];
        if (c >= 0 && c < C) {                                // a class appears at most once in p_cls
            comb[c] += p_score[static_cast<int64_t>(b) * kp + j] * w_proto[c];
            order[c] = j;
        }
    }
    __syncthreads();
    float tot = 0.f;
    for (int c = tid; c < C; c += BD_THREADS) tot += comb[c];
    redv[tid] = tot;
    __syncthreads();
    for (int s2 = BD_THREADS / 2; s2 > 0; s2 >>= 1) {
        if (tid < s2) redv[tid] += redv[tid + s2];
        __syncthreads();
    }
    tot = redv[0];
    __syncthreads();
    for (int r = 0; r < kout; ++r) {
        float bv = -CUDART_INF_F;
        int bi = -1, bo = 0x7fffffff;
        for (int c = tid; c < C; c += BD_THREADS) {
            const float v = comb[c];
            if (v > bv || (v == bv && order[c] < bo)) { bv = v; bi = c; bo = order[c]; }
        }
        redv[tid] = bv; redi[tid] = bi;
        __syncthreads();
        for (int s2 = BD_THREADS / 2; s2 > 0; s2 >>= 1) {
            if (tid < s2) {
                const float ov = redv[tid + s2];
                const int oi = redi[tid + s2];
                const bool better = oi >= 0 && (redi[tid] < 0 || ov > redv[tid] || (ov == redv[tid] && order[oi] < order[redi[tid]]));
                if (better) { redv[tid] = ov; redi[tid] = oi; }
            }
            __syncthreads();
        }
        if (tid == 0) {
            const int c = redi[0];
            out_cls[static_cast<int64_t>(b) * kout + r] = c;
            out_score[static_cast<int64_t>(b) * kout + r] = c >= 0 ? (tot > 0.f ? redv[0] / tot : redv[0]) : 0.f;
            if (c >= 0) comb[c] = -CUDART_INF_F;
        }
        __syncthreads();
    }
}

__global__ void negate_kernel(const float *__restrict__ in, int64_t n, float *__restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = -in[i];
}

// classifier.py:1358-1384 with integer labels; insertion order = prototype entries then head-only entries,
// stable descending sort (Python's sorted(..., reverse=True) keeps insertion order on ties)
__global__ void blend_topk_kernel(const int32_t *__restrict__ p_cls, const float *__restrict__ p_score,
                                  const int64_t *__restrict__ h_idx, const float *__restrict__ h_val, int B, int k,
                                  int kh, float w_proto, float w_head, int32_t *__restrict__ out_cls,
                                  float *__restrict__ out_score) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int32_t cls[2 * BLEND_MAX_K];
    float sc[2 * BLEND_MAX_K];
    int n = 0;
    for (int j = 0; j < k; ++j) {
        const int32_t c = p_cls[static_cast<int64_t>(b) * k + j];
        if (c < 0) continue;
        cls[n] = c;
        sc[n] = p_score[static_cast<int64_t>(b) * k + j] * w_proto;
        ++n;
    }
    for (int j = 0; j < kh; ++j) {
        const int32_t c = static_cast<int32_t>(h_idx[static_cast<int64_t>(b) * kh + j]);
        if (c < 0) continue;
        const float v = -h_val[static_cast<int64_t>(b) * kh + j] * w_head;   // h_val holds the NEGATED probabilities
        int t = 0;
        for (; t < n; ++t)
            if (cls[t] == c) break;
        if (t < n) sc[t] += v;
        else { cls[n] = c; sc[n] = v; ++n; }
    }
    // stable insertion sort, descending
    for (int i = 1; i < n; ++i) {
        const int32_t c = cls[i];
        const float v = sc[i];
        int j = i - 1;
        while (j >= 0 && sc[j] < v) { cls[j + 1] = cls[j]; sc[j + 1] = sc[j]; --j; }
        cls[j + 1] = c; sc[j + 1] = v;
    }
    float total = 0.f;
    for (int t = 0; t < n; ++t) total += sc[t];
    for (int j = 0; j < k; ++j) {
        const bool ok = j < n;
        out_cls[static_cast<int64_t>(b) * k + j] = ok ? cls[j] : -1;
        out_score[static_cast<int64_t>(b) * k + j] = ok ? (total > 0.f ? sc[j] / total : sc[j]) : 0.f;
    }
}

}  // namespace ac

using namespace ac;

extern "C" int ac_proto_class_scores_n(const float *d, const int64_t *idx, const int32_t *row_class, int B, int k, int n_classes,
                                       int32_t *out_cls, float *out_score, ac_stream_t stream) {
    AC_REQUIRE(d && idx && out_cls && out_score && B >= 0, "ac_proto_class_scores_n: bad arguments");
    AC_REQUIRE(k >= 1 && k <= 1024 && n_classes >= 1 && n_classes <= PCS_MAX_CLASSES,
               "ac_proto_class_scores_n: k=%d outside [1,1024] or n_classes=%d outside [1,%d]", k, n_classes, PCS_MAX_CLASSES);
    if (B == 0) return AC_OK;
    proto_class_scores_block_kernel<<<B, PCS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(d, idx, row_class, k, n_classes, out_cls,
                                                                                          out_score);
    AC_LAUNCH_CHECK();
    return AC_OK;
}

extern "C" int ac_blend_dense(const int32_t *proto_cls, const float *proto_score, int kp, const float *head_probs, int B, int C,
                              const float *w_proto, const float *w_head, int kout, int32_t *out_cls, float *out_score,
                              ac_stream_t stream) {
    AC_REQUIRE(proto_cls && proto_score && w_proto && out_cls && out_score && B >= 0 && kp >= 0, "ac_blend_dense: bad arguments");
    AC_REQUIRE(C >= 1 && C <= PCS_MAX_CLASSES && kout >= 1 && kout <= C, "ac_blend_dense: C=%d outside [1,%d] or kout=%d", C,
               PCS_MAX_CLASSES, kout);
    AC_REQUIRE(!head_probs || w_head, "ac_blend_dense: head weights missing");
    if (B == 0) return AC_OK;
    blend_dense_kernel<<<B, BD_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(proto_cls, proto_score, kp, head_probs, C, w_proto, w_head,
                                                                              kout, out_cls, out_score);
    AC_LAUNCH_CHECK();
    return AC_OK;
}

extern "C" int ac_proto_class_scores(const float *d, const int64_t *idx, const int32_t *row_class, int B, int k,
                                     int32_t *out_cls, float *out_score, ac_stream_t stream) {
    AC_REQUIRE(d && idx && out_cls && out_score && B >= 0, "ac_proto_class_scores: bad arguments");
    AC_REQUIRE(k >= 1 && k <= BLEND_MAX_K, "ac_proto_class_scores: k=%d outside [1,%d] (larger k: ac_proto_class_scores_n)", k, BLEND_MAX_K);
    if (B == 0) return AC_OK;
    proto_class_scores_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(d, idx, row_class, B, k,
                                                                                              out_cls, out_score);
    AC_LAUNCH_CHECK();
    return AC_OK;
}

extern "C" int ac_topk_desc_workspace_bytes(int B, int C, int k, size_t *bytes) {
    AC_REQUIRE(bytes && B >= 0 && C >= 1 && k >= 1, "ac_topk_desc_workspace_bytes: bad arguments");
    *bytes = align_up(static_cast<size_t>(B) * C * sizeof(float), 256) + topk_select_workspace(B, C, k) + 512;
    return AC_OK;
}

// out_neg_vals[B,k] = NEGATED values in ascending order (i.e. the k largest values, descending, negated);
// ties -> lower index.  (Negated so that the (d, id) selection kernel is reused unchanged.)
extern "C" int ac_topk_desc(const float *values, int B, int C, int k, float *out_neg_vals, int64_t *out_idx,
                            void *workspace, size_t workspace_bytes, ac_stream_t stream) {
    AC_REQUIRE(values && out_neg_vals && out_idx && workspace && B >= 0 && C >= 1 && k >= 1 && k <= AC_KNN_MAX_K,
               "ac_topk_desc: bad arguments");
    if (B == 0) return AC_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint8_t *w = reinterpret_cast<uint8_t *>(align_up(reinterpret_cast<uintptr_t>(workspace), 256));
    const size_t slack = w - static_cast<uint8_t *>(workspace);
    const size_t nb = align_up(static_cast<size_t>(B) * C * sizeof(float), 256);
    if (slack + nb > workspace_bytes) { set_error("ac_topk_desc: workspace too small"); return AC_E_WORKSPACE; }
    float *neg = reinterpret_cast<float *>(w);
    const int64_t n = static_cast<int64_t>(B) * C;
    negate_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(values, n, neg);
    AC_LAUNCH_CHECK();
    return topk_select(neg, nullptr, B, C, C, 0, k, out_neg_vals, out_idx, w + nb, workspace_bytes - slack - nb, s);
}

extern "C" int ac_blend_topk(const int32_t *proto_cls, const float *proto_score, const int64_t *head_idx,
                             const float *head_neg_val, int B, int k, int kh, float w_proto, float w_head,
                             int32_t *out_cls, float *out_score, ac_stream_t stream) {
    AC_REQUIRE(proto_cls && proto_score && out_cls && out_score && B >= 0, "ac_blend_topk: bad arguments");
    AC_REQUIRE(k >= 1 && k <= BLEND_MAX_K && kh >= 0 && kh <= BLEND_MAX_K, "ac_blend_topk: k/kh outside [1,%d]", BLEND_MAX_K);
    AC_REQUIRE(kh == 0 || (head_idx && head_neg_val), "ac_blend_topk: null head inputs");
    if (B == 0) return AC_OK;
    blend_topk_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        proto_cls, proto_score, head_idx, head_neg_val, B, k, kh, w_proto, w_head, out_cls, out_score);
    AC_LAUNCH_CHECK();
    return AC_OK;
}

// ================================================================================================
// pipeline: ids -> E -> K -> class scores -> H -> top-k -> blend
// ================================================================================================
struct ac_pipeline {
    ac_encoder *enc;
    const float *P, *p_sqnorm;
    const void *p_half;
    const int32_t *row_class;
    ac_head_params head;
    bool has_head;
    int64_t N, row_offset;
    int D, max_B, S, k, kh, shards;
    int32_t *ids_dev, *p_cls, *out_cls;
    float *emb, *knn_d, *p_score, *probs, *h_val, *out_score, *scratch;
    int64_t *knn_i, *h_idx;
    float *loc_d;            // [shards * max_B, k]  this shard's candidates for every rank's queries (sharded mode)
    int64_t *loc_i;
    void *ws, *ws_topk;
    size_t ws_bytes, ws_topk_bytes, scratch_floats;
    // the head (fp32 SIMT) and the prototype scan (tensor cores + HBM) are independent given the embeddings: the head
    // runs on a side stream forked after the encoder and joined before the blend
    cudaStream_t side;
    cudaEvent_t ev_emb, ev_head;
    // search statistics accumulated on the device (ac_knn_l2_topk stats): no host synchronisation inside a predict call
    int32_t *knn_stats;
    // ac_pipeline_predict_host replays the device part of the step as a CUDA graph, one per batch size: at B = 1 the ~110 launches
    // of a step cost more host time than GPU time.  First call with a batch size: eager; second: captured on `cap`; then replayed.
    struct GraphSlot { int B, calls; long long launches; cudaGraphExec_t exec; };
    GraphSlot gslot[8];
    int n_gslot;
    bool graph_off;
    cudaStream_t cap;
};

extern "C" int ac_pipeline_destroy(ac_pipeline *pl) {
    if (!pl) return AC_OK;
    void *ptrs[] = {pl->ids_dev, pl->p_cls, pl->out_cls, pl->emb, pl->knn_d, pl->p_score, pl->probs, pl->h_val,
                    pl->out_score, pl->scratch, pl->knn_i, pl->h_idx, pl->ws, pl->ws_topk, pl->knn_stats, pl->loc_d, pl->loc_i};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (int i = 0; i < pl->n_gslot; ++i) if (pl->gslot[i].exec) cudaGraphExecDestroy(pl->gslot[i].exec);
    if (pl->cap) cudaStreamDestroy(pl->cap);
    if (pl->side) cudaStreamDestroy(pl->side);
    if (pl->ev_emb) cudaEventDestroy(pl->ev_emb);
    if (pl->ev_head) cudaEventDestroy(pl->ev_head);
    delete pl;
    return AC_OK;
}

extern "C" int ac_pipeline_create(ac_encoder *enc, const float *P, const float *p_sqnorm, const void *p_half,
                                  const int32_t *row_class, int64_t N, int D, const ac_head_params *head, int max_B, int S,
                                  int k, int64_t row_offset, int shards, ac_pipeline **out) {
    AC_REQUIRE(enc && P && out && N > 0 && D > 0 && max_B > 0 && S > 0, "ac_pipeline_create: bad arguments");
    AC_REQUIRE(k >= 1 && k <= 16, "ac_pipeline_create: k=%d outside [1,16]", k);
    AC_REQUIRE(shards >= 1 && shards <= 64 && static_cast<int64_t>(shards) * k <= 4096, "ac_pipeline_create: shards=%d", shards);
    ac_pipeline *pl = new ac_pipeline();
    memset(pl, 0, sizeof(*pl));
    pl->enc = enc; pl->P = P; pl->p_sqnorm = p_sqnorm; pl->p_half = p_half; pl->row_class = row_class; pl->N = N; pl->row_offset = row_offset;
    pl->D = D; pl->max_B = max_B; pl->S = S; pl->k = k; pl->shards = shards;
    pl->has_head = head != nullptr;
    if (head) { pl->head = *head; pl->kh = k < head->C ? k : head->C; }
    int rc = ac_knn_workspace_bytes(max_B * shards, N, D, k, AC_KNN_AUTO, &pl->ws_bytes);
    if (rc) { delete pl; return rc; }
    const size_t C = head ? head->C : 1;
    pl->scratch_floats = head ? static_cast<size_t>(max_B) * (head->H0 + head->H1) : 1;
    if (head) ac_topk_desc_workspace_bytes(max_B, head->C, pl->kh, &pl->ws_topk_bytes); else pl->ws_topk_bytes = 256;
    cudaError_t e = cudaSuccess;
    auto al = [&](void **p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes ? bytes : 256); };
    al(reinterpret_cast<void **>(&pl->ids_dev), sizeof(int32_t) * max_B * S);
    al(reinterpret_cast<void **>(&pl->emb), sizeof(float) * max_B * D);
    al(reinterpret_cast<void **>(&pl->knn_d), sizeof(float) * max_B * k);
    al(reinterpret_cast<void **>(&pl->knn_i), sizeof(int64_t) * max_B * k);
    al(reinterpret_cast<void **>(&pl->p_cls), sizeof(int32_t) * max_B * k);
    al(reinterpret_cast<void **>(&pl->p_score), sizeof(float) * max_B * k);
    al(reinterpret_cast<void **>(&pl->probs), sizeof(float) * max_B * C);
    al(reinterpret_cast<void **>(&pl->h_val), sizeof(float) * max_B * k);
    al(reinterpret_cast<void **>(&pl->h_idx), sizeof(int64_t) * max_B * k);
    al(reinterpret_cast<void **>(&pl->out_cls), sizeof(int32_t) * max_B * k);
    al(reinterpret_cast<void **>(&pl->out_score), sizeof(float) * max_B * k);
    al(reinterpret_cast<void **>(&pl->scratch), sizeof(float) * pl->scratch_floats);
    al(&pl->ws, pl->ws_bytes);
    al(&pl->ws_topk, pl->ws_topk_bytes);
    if (shards > 1) {
        al(reinterpret_cast<void **>(&pl->loc_d), sizeof(float) * max_B * shards * k);
        al(reinterpret_cast<void **>(&pl->loc_i), sizeof(int64_t) * max_B * shards * k);
    }
    al(reinterpret_cast<void **>(&pl->knn_stats), 4 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemset(pl->knn_stats, 0, 4 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&pl->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&pl->cap, cudaStreamNonBlocking);
    const char *genv = getenv("AC_PIPELINE_GRAPH");
    pl->graph_off = genv && genv[0] == '0';
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev_emb, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev_head, cudaEventDisableTiming);
    if (e != cudaSuccess) { ac_pipeline_destroy(pl); return check_cuda(e, "ac_pipeline_create cudaMalloc"); }
    *out = pl;
    return AC_OK;
}

// ---- phases of one predict step (the single-GPU entry below chains them; the row-sharded multi-GPU step interleaves the two
// NCCL exchanges of parallel.py between them -- same kernels, same side-stream overlap of the head)
// E: ids -> unit CLS rows (pl->emb); the head (fp32 SIMT) is forked onto the side stream as soon as the embeddings exist
extern "C" int ac_pipeline_encode(ac_pipeline *pl, const int32_t *ids_dev, const int32_t *mask_dev, int B, ac_stream_t stream) {
    AC_REQUIRE(pl && ids_dev && B > 0 && B <= pl->max_B, "ac_pipeline_encode: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = ac_encoder_forward_cls(pl->enc, ids_dev, mask_dev, nullptr, B, pl->S, pl->emb, stream);
    if (rc) return rc;
    if (pl->has_head) {
        AC_CUDA(cudaEventRecord(pl->ev_emb, s));
        AC_CUDA(cudaStreamWaitEvent(pl->side, pl->ev_emb, 0));
        rc = ac_head_forward(pl->emb, B, &pl->head, AC_ACT_SOFTMAX, pl->probs, pl->scratch, pl->scratch_floats, pl->side);
        if (rc) return rc;
        rc = ac_topk_desc(pl->probs, B, pl->head.C, pl->kh, pl->h_val, pl->h_idx, pl->ws_topk, pl->ws_topk_bytes, pl->side);
        if (rc) return rc;
        AC_CUDA(cudaEventRecord(pl->ev_head, pl->side));
    }
    return AC_OK;
}
extern "C" int ac_pipeline_embeddings(ac_pipeline *pl, const float **emb_dev) {
    AC_REQUIRE(pl && emb_dev, "ac_pipeline_embeddings: bad arguments");
    *emb_dev = pl->emb;
    return AC_OK;
}

// class scores of pl->knn_d / pl->knn_i -> join the head -> blend
static int pipeline_finish(ac_pipeline *pl, int B, int32_t *out_cls_dev, float *out_score_dev, ac_stream_t stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = ac_proto_class_scores(pl->knn_d, pl->knn_i, pl->row_class, B, pl->k, pl->p_cls, pl->p_score, stream);
    if (rc) return rc;
    if (pl->has_head) AC_CUDA(cudaStreamWaitEvent(s, pl->ev_head, 0));
    return ac_blend_topk(pl->p_cls, pl->p_score, pl->h_idx, pl->h_val, B, pl->k, pl->has_head ? pl->kh : 0, 0.7f, 0.3f,
                         out_cls_dev, out_score_dev, stream);
}

namespace ac {
// per-shard candidate lists [G*B, k] (block g = the queries of rank g) -> one byte buffer of G chunks, chunk g =
// d[B,k] fp32 | id[B,k] int64: ONE all-to-all moves distances and ids together
__global__ void pack_candidates_kernel(const float *__restrict__ d, const int64_t *__restrict__ idx, int G, int bk /* B*k */,
                                       uint8_t *__restrict__ packed) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= static_cast<int64_t>(G) * bk) return;
    const int g = static_cast<int>(t / bk), e = static_cast<int>(t % bk);
    uint8_t *chunk = packed + static_cast<int64_t>(g) * bk * 12;
    reinterpret_cast<float *>(chunk)[e] = d[t];
    reinterpret_cast<int64_t *>(chunk + static_cast<int64_t>(bk) * 4)[e] = idx[t];
}
// received buffer (chunk g = shard g's list of MY queries) -> [G, B, k] slabs for ac_topk_merge
__global__ void unpack_candidates_kernel(const uint8_t *__restrict__ packed, int G, int bk, float *__restrict__ d,
                                         int64_t *__restrict__ idx) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= static_cast<int64_t>(G) * bk) return;
    const int g = static_cast<int>(t / bk), e = static_cast<int>(t % bk);
    const uint8_t *chunk = packed + static_cast<int64_t>(g) * bk * 12;
    d[t] = reinterpret_cast<const float *>(chunk)[e];
    idx[t] = reinterpret_cast<const int64_t *>(chunk + static_cast<int64_t>(bk) * 4)[e];
}
}  // namespace ac

// K over THIS shard for the queries of every rank: q_all[G*B, D] (all-gathered unit embeddings, rank-major) -> packed[G * B*k*12]
extern "C" int ac_pipeline_search_shard(ac_pipeline *pl, const float *q_all, int G, int B, void *packed, ac_stream_t stream) {
    AC_REQUIRE(pl && q_all && packed && G >= 1 && G <= pl->shards && B > 0 && B <= pl->max_B, "ac_pipeline_search_shard: bad arguments");
    AC_REQUIRE((static_cast<int64_t>(B) * pl->k) % 2 == 0, "ac_pipeline_search_shard: B*k must be even (8-byte alignment of the id block)");
    int rc = ac_knn_l2_topk(q_all, pl->P, pl->p_sqnorm, pl->p_half, G * B, pl->N, pl->D, pl->k, pl->loc_d, pl->loc_i, pl->row_offset,
                            pl->ws, pl->ws_bytes, AC_KNN_AUTO, pl->knn_stats, stream);
    if (rc) return rc;
    const int64_t n = static_cast<int64_t>(G) * B * pl->k;
    pack_candidates_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        pl->loc_d, pl->loc_i, G, B * pl->k, static_cast<uint8_t *>(packed));
    AC_LAUNCH_CHECK();
    return AC_OK;
}

// received[G * B*k*12] (chunk g = shard g's candidates for MY queries) -> merge by (d, global id) -> class scores -> H -> blend
extern "C" int ac_pipeline_finish_sharded(ac_pipeline *pl, const void *received, int G, int B, int32_t *out_cls_dev,
                                          float *out_score_dev, ac_stream_t stream) {
    AC_REQUIRE(pl && received && out_cls_dev && out_score_dev && G >= 1 && G <= pl->shards && B > 0 && B <= pl->max_B,
               "ac_pipeline_finish_sharded: bad arguments");
    const int64_t n = static_cast<int64_t>(G) * B * pl->k;
    unpack_candidates_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint8_t *>(received), G, B * pl->k, pl->loc_d, pl->loc_i);
    AC_LAUNCH_CHECK();
    int rc = ac_topk_merge(pl->loc_d, pl->loc_i, G, B, pl->k, pl->knn_d, pl->knn_i, stream);
    if (rc) return rc;
    return pipeline_finish(pl, B, out_cls_dev, out_score_dev, stream);
}

// device entry: ids_dev[B,S] int32 (device) -> out_cls_dev[B,k] int32, out_score_dev[B,k] fp32 (device)
extern "C" int ac_pipeline_predict_device(ac_pipeline *pl, const int32_t *ids_dev, const int32_t *mask_dev, int B,
                                          int32_t *out_cls_dev, float *out_score_dev, ac_stream_t stream) {
    AC_REQUIRE(pl && ids_dev && out_cls_dev && out_score_dev && B > 0 && B <= pl->max_B, "ac_pipeline_predict_device: bad arguments");
    int rc = ac_pipeline_encode(pl, ids_dev, mask_dev, B, stream);
    if (rc) return rc;
    rc = ac_knn_l2_topk(pl->emb, pl->P, pl->p_sqnorm, pl->p_half, B, pl->N, pl->D, pl->k, pl->knn_d, pl->knn_i,
                        pl->row_offset, pl->ws, pl->ws_bytes, AC_KNN_AUTO, pl->knn_stats, stream);
    if (rc) return rc;
    return pipeline_finish(pl, B, out_cls_dev, out_score_dev, stream);
}

// host entry: ids_host[B,S] (pinned) -> H2D -> predict -> D2H of [B,k] class ids + scores, stream-synchronised
// the device part of a host-boundary step (pl->ids_dev -> pl->out_cls / pl->out_score) on stream s: a CUDA graph replay once the
// batch size has been seen twice, otherwise the ordinary launches.  The graph is recorded on the pipeline's own stream (the
// caller's may be the legacy default stream, which cannot capture) and launched into s.  Timed profiling (bench.py's per-kernel
// events) and AC_PIPELINE_GRAPH=0 keep the step eager; any capture failure turns graphs off for this pipeline.
static int pipeline_step_host(ac_pipeline *pl, int B, cudaStream_t s) {
    auto eager = [&]() { return ac_pipeline_predict_device(pl, pl->ids_dev, nullptr, B, pl->out_cls, pl->out_score, s); };
    if (pl->graph_off || prof_is_on()) return eager();
    ac_pipeline::GraphSlot *g = nullptr;
    for (int i = 0; i < pl->n_gslot; ++i) if (pl->gslot[i].B == B) g = &pl->gslot[i];
    if (!g) {
        if (pl->n_gslot == 8) return eager();
        g = &pl->gslot[pl->n_gslot++];
        g->B = B; g->calls = 0; g->launches = 0; g->exec = nullptr;
    }
    g->calls += 1;
    if (!g->exec) {
        if (g->calls < 2) return eager();            // one-time function attributes and tensor maps are set by an ordinary call
        const long long before = launch_count_now();
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamBeginCapture(pl->cap, cudaStreamCaptureModeThreadLocal);
        int rc = AC_OK;
        if (e == cudaSuccess) {
            rc = ac_pipeline_predict_device(pl, pl->ids_dev, nullptr, B, pl->out_cls, pl->out_score, pl->cap);
            e = cudaStreamEndCapture(pl->cap, &graph);
        }
        const long long recorded = launch_count_now() - before;
        count_launch_n(-recorded);                   // recorded, not executed
        cudaGraphExec_t exec = nullptr;
        if (rc == AC_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (rc != AC_OK || e != cudaSuccess || !exec) {
            cudaGetLastError();                      // clear the capture error; this pipeline stays eager from now on
            pl->graph_off = true;
            return eager();
        }
        g->exec = exec;
        g->launches = recorded;
    }
    AC_CUDA(cudaGraphLaunch(g->exec, s));
    count_launch_n(g->launches);
    return AC_OK;
}

extern "C" int ac_pipeline_predict_host(ac_pipeline *pl, const int32_t *ids_host, int B, int32_t *out_cls_host,
                                        float *out_score_host, ac_stream_t stream) {
    AC_REQUIRE(pl && ids_host && out_cls_host && out_score_host && B > 0 && B <= pl->max_B, "ac_pipeline_predict_host: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    AC_CUDA(cudaMemcpyAsync(pl->ids_dev, ids_host, sizeof(int32_t) * B * pl->S, cudaMemcpyHostToDevice, s));
    int rc = pipeline_step_host(pl, B, s);
    if (rc) return rc;
    AC_CUDA(cudaMemcpyAsync(out_cls_host, pl->out_cls, sizeof(int32_t) * B * pl->k, cudaMemcpyDeviceToHost, s));
    AC_CUDA(cudaMemcpyAsync(out_score_host, pl->out_score, sizeof(float) * B * pl->k, cudaMemcpyDeviceToHost, s));
    AC_CUDA(cudaStreamSynchronize(s));
    return AC_OK;
}

// search statistics since the previous read (synchronises the stream): out[0] queries that took the second tensor pass,
// out[1] queries whose candidate buffer overflowed (their results are not exact: redo with AC_KNN_EXACT), out[2] max rows
// collected for one query, out[3] searches.  reset != 0 clears the counters.
extern "C" int ac_pipeline_knn_stats(ac_pipeline *pl, int32_t *out4_host, int reset, ac_stream_t stream) {
    AC_REQUIRE(pl && out4_host, "ac_pipeline_knn_stats: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    AC_CUDA(cudaMemcpyAsync(out4_host, pl->knn_stats, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (reset) AC_CUDA(cudaMemsetAsync(pl->knn_stats, 0, 4 * sizeof(int32_t), s));
    AC_CUDA(cudaStreamSynchronize(s));
    return AC_OK;
}

// intermediate results of the last predict call (parity tests): copied into caller-owned device buffers
extern "C" int ac_pipeline_debug_copy(ac_pipeline *pl, int B, float *emb_out, float *knn_d_out, int64_t *knn_i_out,
                                      ac_stream_t stream) {
    AC_REQUIRE(pl && B > 0 && B <= pl->max_B, "ac_pipeline_debug_copy: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (emb_out) AC_CUDA(cudaMemcpyAsync(emb_out, pl->emb, sizeof(float) * B * pl->D, cudaMemcpyDeviceToDevice, s));
    if (knn_d_out) AC_CUDA(cudaMemcpyAsync(knn_d_out, pl->knn_d, sizeof(float) * B * pl->k, cudaMemcpyDeviceToDevice, s));
    if (knn_i_out) AC_CUDA(cudaMemcpyAsync(knn_i_out, pl->knn_i, sizeof(int64_t) * B * pl->k, cudaMemcpyDeviceToDevice, s));
    return AC_OK;
}
