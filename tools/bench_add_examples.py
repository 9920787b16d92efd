#!/usr/bin/env python
"""BASELINE.json configs[3]: add_examples() continual loop -- 50k new examples, EWC-penalised head update, bert-base, 1xB200.

Synthetic pre-tokenised sequences: 20 classes, calls of 256 examples, a 21st class introduced at call
100 so that _train_new_classes (+EWC) is traversed.  Reports examples/s end to end, and isolated head optimizer steps/s at
batch 32 (the reference's loop shape).  `--examples` bounds the run (the full 50k takes minutes because, like the
reference, every call retrains the head on the whole memory: classifier.py:1428-1522).

    python tools/bench_add_examples.py --examples 5120
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--examples", type=int, default=5120)
    ap.add_argument("--call", type=int, default=256)
    ap.add_argument("--seq", type=int, default=128)
    args = ap.parse_args()
    import numpy as np
    import torch
    import adaptive_classifier_b200 as acb
    from adaptive_classifier_b200 import _cabi, workload as wl
    from adaptive_classifier_b200.models import AdaptiveHead

    dev = "cuda"
    model, cfg = wl.bert_base_state_dict(1234)
    enc = _cabi.Encoder.from_hf(model, max_tokens=args.call * args.seq, device=dev)

    # an AdaptiveClassifier without the HF/tokenizer constructor: the loop below drives the same methods
    clf = acb.AdaptiveClassifier.__new__(acb.AdaptiveClassifier)
    clf.config = acb.ModelConfig()
    clf.device, clf.use_onnx, clf.model_name = dev, False, "synthetic-bert-base"
    clf.encoder, clf._max_tokens, clf.embedding_dim = enc, args.call * args.seq, 768
    clf.memory = acb.PrototypeMemory(768, config=clf.config)
    clf.adaptive_head, clf.label_to_id, clf.id_to_label = None, {}, {}
    clf.train_steps, clf.training_history = 0, {}
    np.random.seed(0)

    n_calls = args.examples // args.call
    t_embed = t_mem = t_train = 0.0
    torch.cuda.synchronize()
    t0 = time.time()
    for c in range(n_calls):
        ids = wl.synthetic_ids(args.call, args.seq, seed=1000 + c)
        ncls = 21 if c >= 100 else 20
        labels = [f"class_{(c * args.call + i) % ncls:02d}" for i in range(args.call)]
        ta = time.time()
        emb = clf._embed_ids_device(ids, None, None).cpu()
        tb = time.time()
        has_existing = len(clf.label_to_id) > 0
        new = set(labels) - set(clf.label_to_id)
        for l in sorted(new):
            clf.label_to_id[l] = len(clf.label_to_id); clf.id_to_label[clf.label_to_id[l]] = l
        clf.memory.add_examples_batch([acb.Example(f"t{c}_{i}", l, e) for i, (l, e) in enumerate(zip(labels, emb))], labels)
        for l in labels:
            clf.training_history[l] = clf.training_history.get(l, 0) + 1
        tc = time.time()
        if new and has_existing:
            import copy
            old = copy.deepcopy(clf.adaptive_head)
            clf.adaptive_head.update_num_classes(len(clf.label_to_id))
            clf.adaptive_head = clf.adaptive_head.to(dev)
            clf._train_new_classes(old, new)
        else:
            if clf.adaptive_head is None:
                clf._initialize_adaptive_head()
            clf._train_adaptive_head()
        clf.memory._rebuild_index()
        torch.cuda.synchronize()
        td = time.time()
        t_embed += tb - ta; t_mem += tc - tb; t_train += td - tc
    total = time.time() - t0

    # isolated head optimizer steps/s at batch 32
    p = clf.adaptive_head._param_dict()
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v = {k: torch.zeros_like(t) for k, t in p.items()}
    X = torch.nn.functional.normalize(torch.randn(32, 768, device=dev), dim=1)
    y = torch.randint(0, len(clf.label_to_id), (32,), device=dev)
    for s in range(20):
        _cabi.head_train_step(X, y, p, m, v, step=s + 1)
    torch.cuda.synchronize()
    t1 = time.time()
    n_steps = 500
    for s in range(n_steps):
        _cabi.head_train_step(X, y, p, m, v, step=21 + s)
    torch.cuda.synchronize()
    steps_per_s = n_steps / (time.time() - t1)
    print(json.dumps({
        "metric": "examples/sec add_examples() continual loop (bert-base, S=128, calls of 256, EWC path traversed at call 100)",
        "value": n_calls * args.call / total, "unit": "examples/s", "examples": n_calls * args.call,
        "seconds": {"encoder": round(t_embed, 3), "memory_update": round(t_mem, 3), "head_training": round(t_train, 3), "total": round(total, 3)},
        "head_steps_per_s_batch32": steps_per_s, "classes": len(clf.label_to_id), "stored_examples": clf.memory.get_stats()["total_examples"],
        "note": "every call retrains the head on the whole memory for <= 10 epochs like classifier.py:1428-1522; memory is capped at 1000 examples per class"}))


if __name__ == "__main__":
    main()
