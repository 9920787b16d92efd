#!/usr/bin/env python
"""Where a head optimizer step spends its time: per-phase / per-barrier microseconds of CTA 0 of head_train_kernel
(csrc/head_train.cuh), averaged over one epoch of 625 steps at batch 32 (bert-base head, C classes)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptive_classifier_b200 import _cabi
from adaptive_classifier_b200.models import AdaptiveHead

C = int(sys.argv[1]) if len(sys.argv) > 1 else 20
head = AdaptiveHead(768, C, hidden_dims=[768, 384]).cuda()
p = head._param_dict()
m = {k: torch.zeros_like(v) for k, v in p.items()}
v = {k: torch.zeros_like(t) for k, t in p.items()}
n = 20000
X = torch.nn.functional.normalize(torch.randn(n, 768, device="cuda"), dim=1)
y = torch.randint(0, C, (n,), device="cuda")
perm = torch.randperm(n)
_cabi.head_train_epoch(X, y, perm, p, m, v, first_step=1, batch=32)
torch.cuda.synchronize()
t0 = time.time()
for i in range(4):
    _cabi.head_train_epoch(X, y, perm, p, m, v, first_step=1 + 625 * (i + 1), batch=32)
torch.cuda.synchronize()
untimed = 1e6 * (time.time() - t0) / (4 * 625)
_cabi.head_phase_timing(True)
t0 = time.time()
_, nb = _cabi.head_train_epoch(X, y, perm, p, m, v, first_step=626, batch=32)
torch.cuda.synchronize()
wall = time.time() - t0
ns = _cabi.head_phase_timing(False)
names = ["P1 h0", "B1", "P2 h1", "B2", "P3a z", "B3a", "P3b loss/dz", "B3b", "P4 gW2|da1", "B4", "P5/P6 gW1|da0,gW0,norm", "B6", "P7 adamw"]
detail = ["dot wait", "dot issue", "dot multiply", "dot combine", "outer wait", "outer issue", "outer multiply"]
rows = {}
for cls, r in zip(("layer0 CTA", "layer1 CTA", "layer2 CTA (last)"), ns):
    rows[cls] = {"phase_us": {nm: round(x / nb / 1e3, 2) for nm, x in zip(names, r[:13])}, "sum_us": round(sum(r[:13]) / nb / 1e3, 2),
                 "inside_products_us": {nm: round(x / nb / 1e3, 2) for nm, x in zip(detail, r[14:21])}}
print(json.dumps({"classes": C, "plan": _cabi.head_train_plan(p, 32), "steps": nb, "us_per_step_untimed": untimed, "us_per_step_wall": 1e6 * wall / nb, "observed": rows}))
