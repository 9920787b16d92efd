"""Save / load in the reference's on-disk format.

    config.json        model_name, embedding_dim, label_to_id, id_to_label (str keys), train_steps,
                       training_history, config, library_name          classifier.py:546-556
    examples.json      <= num_representative_examples per class (k-means representatives, :1533-1571)
    model.safetensors  prototype_{label} + adaptive_head_{state_dict key}                  :569-578

ONNX export, the model card and Hub transfer are out of scope.
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import List, Optional

import torch
import torch.nn.functional as F

from .models import Example


def select_representative_examples(examples: List[Example], k: int = 5) -> List[Example]:
    """classifier.py:1533-1571: k-means (sklearn, random_state 42, n_init 10) on normalised embeddings,
    example nearest to each centroid."""
    if len(examples) <= k:
        return examples
    from sklearn.cluster import KMeans
    embeddings = F.normalize(torch.stack([ex.embedding for ex in examples]), p=2, dim=1)
    kmeans = KMeans(n_clusters=k, random_state=42, n_init=10)
    kmeans.fit(embeddings.numpy())
    centroids = torch.tensor(kmeans.cluster_centers_, dtype=embeddings.dtype)
    selected = []
    for centroid in centroids:
        selected.append(torch.argmin(torch.norm(embeddings - centroid, dim=1)).item())
    return [examples[i] for i in selected]


def save_classifier(clf, save_directory):
    from safetensors.torch import save_file
    save_directory = Path(save_directory)
    os.makedirs(save_directory, exist_ok=True)
    config_dict = {
        "model_name": getattr(clf.model.config, "_name_or_path", clf.model_name) or clf.model_name,
        "embedding_dim": clf.embedding_dim,
        "label_to_id": clf.label_to_id,
        "id_to_label": {str(k): v for k, v in clf.id_to_label.items()},
        "train_steps": clf.train_steps,
        "training_history": clf.training_history,
        "config": clf.config.to_dict(),
        "library_name": "adaptive-classifier",
    }
    saved_examples = {}
    for label, examples in clf.memory.examples.items():
        saved_examples[label] = [ex.to_dict() for ex in
                                 select_representative_examples(examples, k=clf.config.num_representative_examples)]
    tensor_dict = {}
    for label, proto in clf.memory.prototypes.items():
        tensor_dict[f"prototype_{label}"] = proto.detach().cpu().contiguous()
    if clf.adaptive_head is not None:
        for name, param in clf.adaptive_head.state_dict().items():
            tensor_dict[f"adaptive_head_{name}"] = param.detach().cpu().contiguous()
    with open(save_directory / "config.json", "w", encoding="utf-8") as f:
        json.dump(config_dict, f, indent=2, sort_keys=True)
    with open(save_directory / "examples.json", "w", encoding="utf-8") as f:
        json.dump(saved_examples, f, indent=2, sort_keys=True)
    save_file(tensor_dict, str(save_directory / "model.safetensors"))
    return {"config": "config.json", "examples": "examples.json", "model": "model.safetensors"}, {}


def load_classifier(cls, model_id, device: Optional[str] = None, trust_remote_code: bool = False):
    from safetensors.torch import load_file
    model_path = Path(model_id)
    if not model_path.is_dir():
        raise FileNotFoundError(f"{model_id}: only local directories can be loaded (no network / Hub here)")
    with open(model_path / "config.json", "r", encoding="utf-8") as f:
        config_dict = json.load(f)
    with open(model_path / "examples.json", "r", encoding="utf-8") as f:
        saved_examples = json.load(f)
    kwargs = {}
    if device is not None:
        kwargs["device"] = device
    clf = cls(config_dict["model_name"], config=config_dict.get("config", None),
              trust_remote_code=trust_remote_code, **kwargs)
    clf.label_to_id = config_dict["label_to_id"]
    clf.id_to_label = {int(k): v for k, v in config_dict["id_to_label"].items()}
    clf.train_steps = config_dict["train_steps"]
    clf.training_history = config_dict.get("training_history", {})
    tensors = load_file(str(model_path / "model.safetensors"))
    for label, examples_data in saved_examples.items():
        clf.memory.examples[label] = [Example.from_dict(d) for d in examples_data]
    for label in clf.label_to_id.keys():           # prototypes are restored as saved, not recomputed (:888-895)
        key = f"prototype_{label}"
        if key in tensors:
            clf.memory.prototypes[label] = tensors[key]
    clf.memory._restore_from_save()
    head_params = {k.replace("adaptive_head_", ""): v for k, v in tensors.items() if k.startswith("adaptive_head_")}
    if head_params:
        clf._initialize_adaptive_head()
        clf.adaptive_head.load_state_dict(head_params)
        clf.adaptive_head = clf.adaptive_head.to(clf.device)
    if not clf.training_history:                   # back-compat estimate (:909-913)
        for label, examples in saved_examples.items():
            clf.training_history[label] = len(examples) * 20
    return clf
