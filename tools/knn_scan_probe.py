#!/usr/bin/env python
"""Pass 1 of the prototype scan timed alone at the benchmark shape (512 queries, 1 M x 768 rows, k = 5): ms per launch from the
library's per-kernel events (development probe)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptive_classifier_b200 import _cabi, workload as wl
N, D, C, B, k = 1_000_000, 768, 1000, 512, 5
P = wl.synthetic_rows(0, N, D, C, seed=0, device="cuda")
g = torch.Generator(device="cuda").manual_seed(3)
Q = torch.nn.functional.normalize(torch.randn(B, D, device="cuda", generator=g), dim=1)
ph = _cabi.knn_make_shadow(P); sq = _cabi.row_sqnorm(P)
for _ in range(5):
    _cabi.knn_l2_topk(Q, P, k, p_sqnorm=sq, p_half=ph)
torch.cuda.synchronize()
_cabi.profile_enable(True)
for _ in range(20):
    _cabi.knn_l2_topk(Q, P, k, p_sqnorm=sq, p_half=ph)
_cabi.profile_enable(False)
pa = _cabi.profile_read(2)
print(json.dumps({"scan_ms_per_launch": pa["ms"] / max(1, pa["launches"]), "launches": pa["launches"]}))
