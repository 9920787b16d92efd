#!/usr/bin/env python
"""Single-query predict() latency (the figure the reference's README quotes for CPU: 8.3 ms PyTorch / 2.1 ms ONNX per query,
README.md:256-261, 3 prototypes).  bert-base, host ids in, top-5 (class, score) out, through ac_pipeline_predict_host."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptive_classifier_b200 import _cabi, workload as wl
from adaptive_classifier_b200.models import AdaptiveHead

out = {}
model, cfg = wl.bert_base_state_dict(1234)
enc = _cabi.Encoder.from_hf(model, max_tokens=8 * 128, device="cuda")
for (name, N, C) in (("reference-scale index: 20 prototypes", 20, 20), ("1M x 768 prototypes, 1000 classes", 1_000_000, 1000)):
    P = wl.synthetic_rows(0, N, 768, C, seed=0, device="cuda")
    head = AdaptiveHead(768, C, hidden_dims=[768, 384]).cuda().eval()
    rc = (torch.arange(N, device="cuda") % C).to(torch.int32)
    for S in (16, 128):
        pipe = _cabi.Pipeline(enc, P, 8, S, min(5, C), head=head._param_dict(), row_class=rc,
                              p_sqnorm=_cabi.row_sqnorm(P), p_half=_cabi.knn_make_shadow(P))
        ids = wl.synthetic_ids(1, S).pin_memory()
        for _ in range(20):
            pipe.predict_host(ids)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 200
        for _ in range(n):
            pipe.predict_host(ids)
        dt = (time.perf_counter() - t0) / n
        out[f"{name}, S={S}"] = round(dt * 1e3, 4)
        pipe.close()
print(json.dumps({"metric": "ms per single-query predict (B=1, host ids -> top-5 on host)", "values_ms": out}))
