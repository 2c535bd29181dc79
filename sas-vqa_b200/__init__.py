"""sasvqa_b200 -- B200-native SAS-VQA frame-sampling hot path."""
from . import synth  # noqa: F401
