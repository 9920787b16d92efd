#!/usr/bin/env python
"""k = C search at the benchmark shape (512 queries, 1 M x 768 rows, k = 1000): time per search, overflow statistics, equality with
the exact scan on 8 queries (development probe for the queries-per-pass constant of csrc/knn_tc.cu)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptive_classifier_b200 import _cabi, workload as wl
N, D, C, B, k = 1_000_000, 768, 1000, 512, 1000
P = wl.synthetic_rows(0, N, D, C, seed=0, device="cuda")
g = torch.Generator(device="cuda").manual_seed(3)
Q = torch.nn.functional.normalize(P[torch.randint(0, N, (B,), device="cuda", generator=g)] + 0.05 * torch.randn(B, D, device="cuda", generator=g), dim=1)
ph = _cabi.knn_make_shadow(P); sq = _cabi.row_sqnorm(P)
stats = torch.zeros(4, dtype=torch.int32, device="cuda")
for _ in range(2):
    d, i = _cabi.knn_l2_topk(Q, P, k, p_sqnorm=sq, p_half=ph, stats=stats)
torch.cuda.synchronize()
stats.zero_()
t0 = time.time()
for _ in range(5):
    d, i = _cabi.knn_l2_topk(Q, P, k, p_sqnorm=sq, p_half=ph, stats=stats)
torch.cuda.synchronize()
ms = (time.time() - t0) / 5 * 1e3
de, ie = _cabi.knn_l2_topk(Q[:8], P, k, p_sqnorm=sq, algo=_cabi.AC_KNN_EXACT)
print(json.dumps({"ms_per_search": round(ms, 3), "stats_second_pass_overflow_maxcollected_searches": stats.tolist(),
                  "equals_exact_scan": bool(torch.equal(i[:8], ie) and torch.equal(d[:8], de))}))
