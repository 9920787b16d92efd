"""adaptive_classifier_b200 -- B200-native predict()/add_examples() hot path of codelion/adaptive-classifier.

Same public names as /root/reference/src/adaptive_classifier/__init__.py:1-16.  Importing the package needs
no GPU; constructing a classifier or calling any kernel does (there is no CPU fallback).
"""
from .models import Example, AdaptiveHead, ModelConfig
from .memory import PrototypeMemory, FlatL2Index
from .ewc import EWC
from .classifier import AdaptiveClassifier
from .multilabel import MultiLabelAdaptiveClassifier, MultiLabelAdaptiveHead
from ._cabi import AdaptiveB200Error

__version__ = "0.1.0"

__all__ = [
    "AdaptiveClassifier", "MultiLabelAdaptiveClassifier", "MultiLabelAdaptiveHead", "Example", "AdaptiveHead",
    "ModelConfig", "PrototypeMemory", "EWC", "FlatL2Index", "AdaptiveB200Error",
]
