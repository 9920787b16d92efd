"""Golden-vector generator (test infrastructure; runs ONLY in the dev container).

Imports the UNMODIFIED Python reference from /root/reference/src with oracle/shim/faiss.py standing in for
faiss-cpu (absent offline), drives it on seeded inputs and commits the resulting tensors as small fixtures
under tests/golden/.  /root/reference does not exist on the GPU box, so tests only read the fixtures.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

Fixtures
  golden_memory.npz      PrototypeMemory: adds -> prototypes, get_nearest_prototypes labels/scores
  golden_router.npz      the two real prototypes of scripts/adaptive_router/tensors.safetensors (6 KB,
                         inter-prototype d = 0.001965: near-tie stress) + reference search results
  golden_head.npz        AdaptiveHead forward / EWC loss values of the reference on seeded inputs
  golden_classifier.npz  tiny seeded BERT checkpoint + vocab, reference _get_embeddings / add_examples /
                         predict / predict_batch outputs and the reference-trained head
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "shim"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def gen_memory():
    from adaptive_classifier.memory import PrototypeMemory
    from adaptive_classifier.models import Example, ModelConfig
    g = torch.Generator().manual_seed(11)
    D = 64
    labels = ["alpha", "beta", "gamma"]
    centres = unit(torch.randn(3, D, generator=g))
    embs, labs = [], []
    for i in range(60):
        c = i % 3
        embs.append(unit(centres[c] + 0.3 * torch.randn(D, generator=g) / D ** 0.5 * 4))
        labs.append(labels[c])
    mem = PrototypeMemory(D, config=ModelConfig({"prototype_update_frequency": 25, "max_examples_per_class": 15}))
    for e, l in zip(embs, labs):
        mem.add_example(Example(f"t{len(l)}", l, e), l)
    mem._rebuild_index()
    queries = unit(torch.randn(8, D, generator=g) * 0.2 + centres[torch.arange(8) % 3])
    res_labels, res_scores = [], []
    for q in queries:
        r = mem.get_nearest_prototypes(q, k=3)
        res_labels.append([labels.index(l) for l, _ in r])
        res_scores.append([s for _, s in r])
    r2 = mem.get_nearest_prototypes(queries[0], k=2)
    np.savez_compressed(
        os.path.join(OUT, "golden_memory.npz"),
        embeddings=torch.stack(embs).numpy(), label_ids=np.array([labels.index(l) for l in labs]),
        labels=np.array(labels), queries=queries.numpy(),
        prototypes=torch.stack([mem.prototypes[l] for l in sorted(mem.prototypes)]).numpy(),
        kept=np.array([len(mem.examples[l]) for l in labels]),
        res_labels=np.array(res_labels), res_scores=np.array(res_scores, dtype=np.float64),
        res_k2_labels=np.array([labels.index(l) for l, _ in r2]), res_k2_scores=np.array([s for _, s in r2]),
        updates_since_rebuild=np.array(mem.updates_since_rebuild))
    print("golden_memory ok")


def gen_router():
    import faiss
    from safetensors.torch import load_file
    t = load_file("/root/reference/scripts/adaptive_router/tensors.safetensors")
    P = torch.stack([t["prototype_HIGH"], t["prototype_LOW"]]).float()
    g = torch.Generator().manual_seed(5)
    Q = unit(P[torch.arange(12) % 2] + 0.02 * torch.randn(12, P.shape[1], generator=g))
    idx = faiss.IndexFlatL2(P.shape[1])
    idx.add(P.numpy())
    d, i = idx.search(Q.numpy(), 2)
    np.savez_compressed(os.path.join(OUT, "golden_router.npz"), P=P.numpy(), Q=Q.numpy(), d=d, i=i)
    print("golden_router ok; inter-prototype d =", float(((P[0] - P[1]) ** 2).sum()))


def gen_head():
    from adaptive_classifier.models import AdaptiveHead
    from adaptive_classifier.ewc import EWC
    D, C = 64, 5
    head = AdaptiveHead(D, C, hidden_dims=[D, D // 2])
    head.eval()
    g = torch.Generator().manual_seed(21)
    X = unit(torch.randn(16, D, generator=g))
    with torch.no_grad():
        logits = head(X)
    sd = {k: v.numpy().copy() for k, v in head.state_dict().items()}
    # EWC identities of tests/test_ewc.py:128-153 on the reference: loss == 0 at theta*, > 0 after +0.1
    ds = torch.utils.data.TensorDataset(X, torch.randint(0, C, (16,), generator=g))
    torch.manual_seed(0)
    ewc = EWC(head, ds, device="cpu", ewc_lambda=100.0)
    loss0 = float(ewc.ewc_loss())
    with torch.no_grad():
        for p in head.parameters():
            p.add_(0.1)
    loss1 = float(ewc.ewc_loss())
    loss1_b32 = float(ewc.ewc_loss(batch_size=32))
    fisher = {("fisher_" + k): v.numpy().copy() for k, v in ewc.fisher_info.items()}
    np.savez_compressed(os.path.join(OUT, "golden_head.npz"), X=X.numpy(), logits=logits.numpy().copy(),
                        ewc_loss0=loss0, ewc_loss1=loss1, ewc_loss1_b32=loss1_b32, **sd, **fisher)
    print("golden_head ok", loss0, loss1, loss1_b32)


def gen_classifier():
    from transformers import BertConfig, BertModel, BertTokenizerFast
    from adaptive_classifier import AdaptiveClassifier
    words = [f"w{i}" for i in range(195)]
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + words
    cfg = BertConfig(vocab_size=len(vocab), hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                     intermediate_size=256, max_position_embeddings=64, type_vocab_size=2, pad_token_id=0)
    torch.manual_seed(1234)
    model = BertModel(cfg)
    # random-init LayerNorm/bias are trivial (1/0): perturb them so every parameter is exercised
    g = torch.Generator().manual_seed(99)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "LayerNorm" in n or n.endswith(".bias"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif "weight" in n and p.dim() == 2:
                p.mul_(3.0)       # std 0.06: attention / FFN actually mix
    tmp = tempfile.mkdtemp(prefix="golden_ckpt_")
    model.save_pretrained(tmp)
    with open(os.path.join(tmp, "vocab.txt"), "w") as f:
        f.write("\n".join(vocab) + "\n")
    BertTokenizerFast(vocab_file=os.path.join(tmp, "vocab.txt"), do_lower_case=True).save_pretrained(tmp)

    rng = np.random.default_rng(7)
    class_words = {"sports": words[0:40], "finance": words[40:80], "cooking": words[80:120]}

    def sentence(label, n):
        own = rng.choice(class_words[label], size=n, replace=True)
        noise = rng.choice(words[120:], size=max(1, n // 4), replace=True)
        toks = list(own) + list(noise)
        rng.shuffle(toks)
        return " ".join(toks)

    texts, labels = [], []
    for label in ["sports", "finance", "cooking"]:
        for _ in range(12):
            texts.append(sentence(label, int(rng.integers(4, 14))))
            labels.append(label)
    test_texts = [sentence(l, 9) for l in ["sports", "finance", "cooking", "finance", "sports", "cooking"]]

    torch.manual_seed(0)
    np.random.seed(0)
    clf = AdaptiveClassifier(tmp, device="cpu", use_onnx=False)
    clf.add_examples(texts[:24], labels[:24])           # sports + finance -> _train_adaptive_head
    clf.add_examples(texts[24:], labels[24:])           # new class cooking -> _train_new_classes (+EWC)
    emb_train = torch.stack(clf._get_embeddings(texts)).numpy()
    emb_test = torch.stack(clf._get_embeddings(test_texts)).numpy()
    enc = clf.tokenizer(texts + test_texts, max_length=512, truncation=True, padding=True, return_tensors="pt")
    label_names = [clf.id_to_label[i] for i in range(len(clf.id_to_label))]
    pred = [clf.predict(t, k=3) for t in test_texts]
    pred_k1 = [clf.predict(t, k=1) for t in test_texts]
    pred_b = clf.predict_batch(test_texts, k=2)
    train_top1 = [p[0][0] for p in clf.predict_batch(texts, k=1)]     # end metric of the reference's own training

    def pack(preds, k):
        L = np.full((len(preds), k), -1, dtype=np.int64)
        S = np.zeros((len(preds), k), dtype=np.float64)
        for i, p in enumerate(preds):
            for j, (l, s) in enumerate(p):
                L[i, j] = label_names.index(l)
                S[i, j] = s
        return L, S

    pl, ps = pack(pred, 3)
    p1l, p1s = pack(pred_k1, 1)
    pbl, pbs = pack(pred_b, 2)
    head_sd = {("head_" + k): v.detach().numpy() for k, v in clf.adaptive_head.state_dict().items()}
    model_sd = {("bert_" + k): v.detach().numpy() for k, v in model.state_dict().items()}
    protos = np.stack([clf.memory.prototypes[l].numpy() for l in sorted(clf.memory.prototypes)])
    np.savez_compressed(
        os.path.join(OUT, "golden_classifier.npz"),
        vocab=np.array(vocab), texts=np.array(texts), labels=np.array(labels), test_texts=np.array(test_texts),
        label_names=np.array(label_names), input_ids=enc["input_ids"].numpy(), attention_mask=enc["attention_mask"].numpy(),
        emb_train=emb_train, emb_test=emb_test, prototypes=protos, proto_labels=np.array(sorted(clf.memory.prototypes)),
        training_history=json.dumps(clf.training_history), train_steps=clf.train_steps,
        pred_labels=pl, pred_scores=ps, pred_k1_labels=p1l, pred_k1_scores=p1s, predb_labels=pbl, predb_scores=pbs,
        train_top1=np.array([label_names.index(l) for l in train_top1]),
        bert_config=json.dumps(cfg.to_dict()), **head_sd, **model_sd)
    print("golden_classifier ok; labels", label_names, "pred[0]", pred[0])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_memory()
    gen_router()
    gen_head()
    gen_classifier()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
