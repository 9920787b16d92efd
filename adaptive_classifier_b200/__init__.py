"""adaptive_classifier_b200 -- B200-native predict()/add_examples() hot path of codelion/adaptive-classifier."""
__version__ = "0.1.0"
