/*
 * ORACLE (test infrastructure only -- never linked or called by the product path).
 *
 * CPU restatement of stage K of the hot path:
 *   PrototypeMemory.get_nearest_prototypes   /root/reference/src/adaptive_classifier/memory.py:85-136
 *     index.search(query[1,D], k)            memory.py:110-114   (faiss.IndexFlatL2, nq = 1 always)
 *     similarities = exp(-distances)         memory.py:117
 *     softmax over the k returned scores     memory.py:128-134
 *
 * faiss-cpu (requirements.txt:4, `faiss-cpu>=1.7.4`, lower bound only, no lock file) is a third-party
 * dependency that is absent from /root/reference and from this image, so its published algorithm is
 * restated: IndexFlatL2::search for nq < distance_compute_blas_threshold (20) scans every stored row
 * with fvec_L2sqr (no ||x||^2+||y||^2-2xy expansion) and keeps the k smallest in a max-heap, returned
 * ascending; labels are int64 row ids, -1 padded when fewer than k rows exist.  fvec_L2sqr's AVX2 form
 * (faiss 1.7.x utils/distances_simd.cpp, restated FROM MEMORY): eight lane accumulators, lane j sums
 * (x[i]-y[i])^2 over i == j (mod 8) in index order with separate multiply and add roundings; the
 * tail is zero-padded into the same lanes; lanes are combined as
 *      ((l0+l4)+(l1+l5)) + ((l2+l6)+(l3+l7)).
 * Ties are ordered by the lower row id.
 *
 * PARITY UNPINNED against real FAISS: no reference test pins a distance value, an ordering on real
 * data or tie behaviour; what is pinned is (a) tests/test_memory.py semantics
 * (3 results, softmax sums to 1) and (b) agreement with float64 ground truth to < 1e-5.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -o liboracle_knn.so knn_oracle.c -lm   (see oracle/build.py)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* fvec_L2sqr, AVX2 8-lane order (see header). */
float oracle_l2sqr(const float *x, const float *y, int d)
{
    float l[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int i;
    for (i = 0; i < d; ++i) {
        float t = x[i] - y[i];
        float sq = t * t;
        l[i & 7] = l[i & 7] + sq;
    }
    {
        float a = l[0] + l[4], b = l[1] + l[5], c = l[2] + l[6], e = l[3] + l[7];
        return (a + b) + (c + e);
    }
}

/* all distances for one query */
void oracle_l2sqr_ny(const float *q, const float *P, int64_t n, int d, float *out)
{
    int64_t j;
    for (j = 0; j < n; ++j) out[j] = oracle_l2sqr(q, P + j * (int64_t)d, d);
}

typedef struct { float d; int64_t i; } cand_t;

static int cand_less(const cand_t *a, const cand_t *b)
{
    return (a->d < b->d) || (a->d == b->d && a->i < b->i);
}

/* max-heap on (d, i): root is the WORST kept candidate */
static void heap_sift_down(cand_t *h, int n, int pos)
{
    for (;;) {
        int l = 2 * pos + 1, r = l + 1, m = pos;
        if (l < n && cand_less(&h[m], &h[l])) m = l;
        if (r < n && cand_less(&h[m], &h[r])) m = r;
        if (m == pos) return;
        { cand_t t = h[m]; h[m] = h[pos]; h[pos] = t; }
        pos = m;
    }
}

static int cand_cmp(const void *a, const void *b)
{
    const cand_t *x = (const cand_t *)a, *y = (const cand_t *)b;
    if (cand_less(x, y)) return -1;
    if (cand_less(y, x)) return 1;
    return 0;
}

/*
 * IndexFlatL2.search restatement.  Q[nq,d], P[n,d] row-major fp32.
 * out_d[nq,k] ascending, out_i[nq,k] int64 (row_offset added), padded with (+inf? no: FAISS pads
 * distances with FLT_MAX-like +inf and labels with -1) when n < k.
 */
void oracle_knn_l2(const float *Q, const float *P, int nq, int64_t n, int d, int k,
                   float *out_d, int64_t *out_i, int64_t row_offset)
{
    int b;
    cand_t *heap = (cand_t *)malloc(sizeof(cand_t) * (size_t)(k > 0 ? k : 1));
    for (b = 0; b < nq; ++b) {
        const float *q = Q + (int64_t)b * d;
        int hs = 0;
        int64_t j;
        for (j = 0; j < n; ++j) {
            cand_t c;
            c.d = oracle_l2sqr(q, P + j * (int64_t)d, d);
            c.i = j;
            if (hs < k) {
                /* push */
                int pos = hs++;
                heap[pos] = c;
                while (pos > 0) {
                    int par = (pos - 1) / 2;
                    if (cand_less(&heap[par], &heap[pos])) {
                        cand_t t = heap[par]; heap[par] = heap[pos]; heap[pos] = t;
                        pos = par;
                    } else break;
                }
            } else if (k > 0 && cand_less(&c, &heap[0])) {
                heap[0] = c;
                heap_sift_down(heap, hs, 0);
            }
        }
        qsort(heap, (size_t)hs, sizeof(cand_t), cand_cmp);
        for (j = 0; j < k; ++j) {
            if (j < hs) {
                out_d[(int64_t)b * k + j] = heap[j].d;
                out_i[(int64_t)b * k + j] = heap[j].i + row_offset;
            } else {
                out_d[(int64_t)b * k + j] = INFINITY;
                out_i[(int64_t)b * k + j] = -1;
            }
        }
    }
    free(heap);
}

/* memory.py:117,128-134: scores = softmax_k(exp(-d)) computed like torch.softmax on fp32
 * (max-subtracted, fp32 exp, fp32 sum in index order).  Entries with index < 0 are skipped by the
 * reference (memory.py:121-125); here they get score 0 and do not enter the softmax. */
void oracle_proto_scores(const float *d, const int64_t *idx, int nq, int k, float *scores)
{
    int b, j;
    for (b = 0; b < nq; ++b) {
        const float *db = d + (int64_t)b * k;
        const int64_t *ib = idx + (int64_t)b * k;
        float *sb = scores + (int64_t)b * k;
        float mx = -INFINITY, sum = 0.f;
        for (j = 0; j < k; ++j) {
            if (ib[j] < 0) { sb[j] = 0.f; continue; }
            sb[j] = expf(-db[j]);
            if (sb[j] > mx) mx = sb[j];
        }
        for (j = 0; j < k; ++j) {
            if (ib[j] < 0) continue;
            sb[j] = expf(sb[j] - mx);
            sum += sb[j];
        }
        for (j = 0; j < k; ++j) {
            if (ib[j] < 0) continue;
            sb[j] = sb[j] / sum;
        }
    }
}

/* deterministic merge of per-shard top-k lists (multi-GPU row sharding):
 * candidates d[G,nq,k], i[G,nq,k] -> k smallest by (d, i); idx < 0 entries ignored. */
void oracle_topk_merge(const float *d, const int64_t *idx, int G, int nq, int k,
                       float *out_d, int64_t *out_i)
{
    int b, g, j;
    cand_t *all = (cand_t *)malloc(sizeof(cand_t) * (size_t)G * (size_t)k);
    for (b = 0; b < nq; ++b) {
        int n = 0;
        for (g = 0; g < G; ++g)
            for (j = 0; j < k; ++j) {
                int64_t off = ((int64_t)g * nq + b) * k + j;
                if (idx[off] < 0) continue;
                all[n].d = d[off]; all[n].i = idx[off]; ++n;
            }
        qsort(all, (size_t)n, sizeof(cand_t), cand_cmp);
        for (j = 0; j < k; ++j) {
            if (j < n) { out_d[(int64_t)b * k + j] = all[j].d; out_i[(int64_t)b * k + j] = all[j].i; }
            else { out_d[(int64_t)b * k + j] = INFINITY; out_i[(int64_t)b * k + j] = -1; }
        }
    }
    free(all);
}
