"""kNN-only workload for ncu: the only gemm_tc_kernel launches are the prototype scan (fp16 shadow and tf32 variants)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptive_classifier_b200 import _cabi, workload as wl
N, D, C, B, k = 1_000_000, 768, 1000, 512, 5
P = wl.synthetic_rows(0, N, D, C, seed=0, device="cuda")
Q = wl.synthetic_queries_embeddings(B, D, C, device="cuda")
pn = _cabi.row_sqnorm(P)
Ph = _cabi.knn_make_shadow(P)
for it in range(3):
    d, i = _cabi.knn_l2_topk(Q, P, k, p_sqnorm=pn, p_half=Ph, algo=_cabi.AC_KNN_TENSOR)
    d2, i2 = _cabi.knn_l2_topk(Q, P, k, p_sqnorm=pn, algo=_cabi.AC_KNN_TENSOR)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
for _ in range(5):
    _cabi.knn_l2_topk(Q, P, k, p_sqnorm=pn, p_half=Ph, algo=_cabi.AC_KNN_TENSOR)
e1.record()
for _ in range(5):
    _cabi.knn_l2_topk(Q, P, k, p_sqnorm=pn, algo=_cabi.AC_KNN_TENSOR)
e2.record()
torch.cuda.synchronize()
print(f"kNN tensor path, B={B}, N={N}: fp16 shadow {e0.elapsed_time(e1)/5:.3f} ms/call, tf32 on fp32 rows {e1.elapsed_time(e2)/5:.3f} ms/call; "
      f"identical results: {bool(torch.equal(i, i2) and torch.equal(d, d2))}; top-1 own class: {bool(torch.equal(i[:,0].cpu() % C, torch.arange(B) % C))}")
