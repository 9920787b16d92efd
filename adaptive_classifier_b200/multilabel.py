"""Host-side mirror of /root/reference/src/adaptive_classifier/multilabel.py.

MultiLabelAdaptiveHead (sigmoid inside forward, default-initialised Linear layers) and
MultiLabelAdaptiveClassifier (adaptive / per-label thresholds, min/max predictions, BCE training on
multi-hot targets).  Sigmoid epilogue and the BCE step are the AC_ACT_SIGMOID / AC_LOSS_BCE variants of
csrc/head.cu; thresholds stay on the host.
"""
from __future__ import annotations

import logging
from collections import defaultdict
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi
from .classifier import AdaptiveClassifier
from .models import _CudaHeadMixin

logger = logging.getLogger(__name__)


class MultiLabelAdaptiveHead(_CudaHeadMixin, nn.Module):
    """multilabel.py:15-68."""

    _act = _cabi.AC_ACT_SIGMOID

    def __init__(self, input_dim: int, num_classes: int, hidden_dims: List[int] = None):
        super().__init__()
        if hidden_dims is None:
            hidden_dims = [input_dim // 2]
        layers = []
        prev_dim = input_dim
        for dim in hidden_dims:
            layers.extend([nn.Linear(prev_dim, dim), nn.ReLU(), nn.Dropout(0.1)])
            prev_dim = dim
        layers.append(nn.Linear(prev_dim, num_classes))
        self.model = nn.Sequential(*layers)
        self.num_classes = num_classes

    def forward(self, x):
        return self._forward_cuda(x, _cabi.AC_ACT_SIGMOID)     # probabilities (multilabel.py:41-44)

    def update_num_classes(self, new_num_classes: int):
        if new_num_classes <= self.num_classes:
            return
        final_layer = self.model[-1]
        new_final_layer = nn.Linear(final_layer.in_features, new_num_classes)
        with torch.no_grad():
            new_final_layer.weight[: self.num_classes] = final_layer.weight.detach().cpu()
            new_final_layer.bias[: self.num_classes] = final_layer.bias.detach().cpu()
            nn.init.xavier_uniform_(new_final_layer.weight[self.num_classes:])
            nn.init.zeros_(new_final_layer.bias[self.num_classes:])
        self.model[-1] = new_final_layer.to(final_layer.weight.device)
        self.num_classes = new_num_classes


class MultiLabelAdaptiveClassifier(AdaptiveClassifier):
    """multilabel.py:71-426."""

    def __init__(self, model_name: str, device: Optional[str] = None, config: Optional[Dict[str, Any]] = None,
                 seed: int = 42, default_threshold: float = 0.5, min_predictions: int = 1,
                 max_predictions: Optional[int] = None, use_onnx="auto", trust_remote_code: bool = False):
        # use_onnx / trust_remote_code accepted so the non-ONNX load path works
        super().__init__(model_name, device, config, seed, use_onnx=use_onnx, trust_remote_code=trust_remote_code)
        self.default_threshold = default_threshold
        self.min_predictions = min_predictions
        self.max_predictions = max_predictions
        self.label_thresholds = {}
        self.adaptive_head = None

    def _initialize_adaptive_head(self):
        num_classes = len(self.label_to_id)
        hidden_dims = [self.embedding_dim, self.embedding_dim // 2]
        self.adaptive_head = MultiLabelAdaptiveHead(self.embedding_dim, num_classes, hidden_dims=hidden_dims).to(self.device)

    def _get_adaptive_threshold(self, num_labels: int) -> float:
        """multilabel.py:113-130."""
        if num_labels <= 2:
            return self.default_threshold
        elif num_labels <= 5:
            return self.default_threshold * 0.8
        elif num_labels <= 10:
            return self.default_threshold * 0.6
        elif num_labels <= 20:
            return self.default_threshold * 0.4
        return self.default_threshold * 0.2

    def _sigmoid_probs(self, emb: torch.Tensor) -> torch.Tensor:
        self.adaptive_head.eval()
        return _cabi.head_forward(emb.contiguous(), self.adaptive_head._param_dict(), _cabi.AC_ACT_SIGMOID)

    def predict_multilabel(self, text: str, threshold: Optional[float] = None,
                           max_labels: Optional[int] = None) -> List[Tuple[str, float]]:
        """multilabel.py:132-229."""
        if not text:
            raise ValueError("Empty input text")
        num_labels = len(self.label_to_id)
        if num_labels == 0:
            return []
        if threshold is None:
            threshold = self._get_adaptive_threshold(num_labels)
        max_labels = max_labels or self.max_predictions
        emb = self._embed_device([text])
        probabilities = None
        if self.adaptive_head is not None:
            probabilities = self._sigmoid_probs(emb)[0].cpu()
            predictions = []
            for i, prob in enumerate(probabilities.tolist()):
                if i < len(self.id_to_label):
                    label = self.id_to_label[i]
                    if prob >= self.label_thresholds.get(label, threshold):
                        predictions.append((label, prob))
            predictions.sort(key=lambda x: x[1], reverse=True)
            if max_labels and len(predictions) > max_labels:
                predictions = predictions[:max_labels]
        else:
            proto_predictions = self.memory.get_nearest_prototypes_batch(
                emb, min(num_labels, max_labels) if max_labels else num_labels)[0]
            predictions = [(label, score) for label, score in proto_predictions if score >= threshold]
        if len(predictions) < self.min_predictions and self.adaptive_head is not None:
            values, indices = torch.topk(probabilities, min(self.min_predictions, len(self.id_to_label)))
            additional = []
            for val, idx in zip(values.tolist(), indices.tolist()):
                if idx < len(self.id_to_label):
                    label = self.id_to_label[idx]
                    if not any(pred[0] == label for pred in predictions):
                        additional.append((label, val))
            predictions.extend(additional[: self.min_predictions - len(predictions)])
            predictions.sort(key=lambda x: x[1], reverse=True)
        return predictions

    def predict(self, text: str, k: int = 5) -> List[Tuple[str, float]]:
        """multilabel.py:231-243."""
        preds = self.predict_multilabel(text, max_labels=k)
        if preds:
            return preds[:k]
        return super().predict(text, k)

    def _head_probs(self, emb):
        # base-class blending applies softmax to the module output; for this head the module output is the
        # sigmoid probability vector (multilabel.py:41-44 + classifier.py:435)
        if self.adaptive_head is None:
            return None
        return torch.softmax(self._sigmoid_probs(emb), dim=1)

    def add_examples(self, texts: List[str], labels: List[List[str]]):
        """multilabel.py:245-279: one (text, label) pair per label, then per-label thresholds."""
        if not texts or not labels:
            raise ValueError("Empty input lists")
        if len(texts) != len(labels):
            raise ValueError("Mismatched text and label lists")
        flat_texts, flat_labels = [], []
        for text, text_labels in zip(texts, labels):
            if not text_labels:
                continue
            for label in text_labels:
                flat_texts.append(text)
                flat_labels.append(label)
        if flat_texts:
            super().add_examples(flat_texts, flat_labels)
        self._update_label_thresholds()

    def _update_label_thresholds(self):
        """multilabel.py:280-307."""
        if not self.memory.examples:
            return
        counts = {label: len(ex) for label, ex in self.memory.examples.items()}
        total = sum(counts.values())
        for label, count in counts.items():
            freq = count / total
            if freq < 0.05:
                self.label_thresholds[label] = self.default_threshold * 0.3
            elif freq < 0.1:
                self.label_thresholds[label] = self.default_threshold * 0.5
            elif freq > 0.3:
                self.label_thresholds[label] = self.default_threshold * 1.2
            else:
                self.label_thresholds[label] = self.default_threshold

    def _train_adaptive_head(self, epochs: int = 10):
        """multilabel.py:309-413: multi-hot targets per unique text, BCE on sigmoid outputs, no scheduler."""
        if not self.memory.examples:
            return
        num_classes = len(self.label_to_id)
        text_to_labels = defaultdict(set)
        first_embedding = {}
        for label, examples in self.memory.examples.items():
            for ex in examples:
                text_to_labels[ex.text].add(label)
        for text, labels in text_to_labels.items():
            emb = None
            for label in labels:
                for ex in self.memory.examples[label]:
                    if ex.text == text:
                        emb = ex.embedding
                        break
                if emb is not None:
                    break
            first_embedding[text] = emb
        embs, targets = [], []
        for text, labels in text_to_labels.items():
            if first_embedding[text] is None:
                continue
            embs.append(first_embedding[text])
            vec = torch.zeros(num_classes)
            for label in labels:
                if label in self.label_to_id:
                    vec[self.label_to_id[label]] = 1.0
            targets.append(vec)
        if not embs:
            return
        X = F.normalize(torch.stack(embs).to(self.device, dtype=torch.float32), p=2, dim=1)
        Y = torch.stack(targets).to(self.device)
        self._loss_kind = _cabi.AC_LOSS_BCE
        try:
            self._run_epochs(X, Y, epochs=epochs, batch_size=min(32, X.shape[0]), use_scheduler=False)
        finally:
            self._loss_kind = _cabi.AC_LOSS_CE
        self.train_steps += 1

    def _train_new_classes(self, old_head, new_classes):
        """The reference inherits classifier.py:202-367 here, applying CrossEntropyLoss to the sigmoid
        OUTPUTS of this head.  That quirk needs a CE-on-probabilities step the B200
        head kernels do not provide; the multilabel path retrains with its own BCE loop instead."""
        self._train_adaptive_head()

    def get_label_statistics(self) -> Dict[str, Any]:
        stats = super().get_example_statistics()
        stats["label_thresholds"] = dict(self.label_thresholds)
        stats["adaptive_threshold"] = self._get_adaptive_threshold(len(self.label_to_id))
        stats["default_threshold"] = self.default_threshold
        stats["min_predictions"] = self.min_predictions
        stats["max_predictions"] = self.max_predictions
        return stats
