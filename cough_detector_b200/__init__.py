"""B200-native feature + CoughDetectorResidual hot path of cough_detector.

Importing the package does not need a GPU or the shared library; the first call that computes
anything loads ``lib/libcough_b200.so`` and fails loudly if it (or a B200) is missing.
"""
__version__ = "0.1.0"

from .preprocessing import AudioPreprocessor, RealtimePreprocessor, create_preprocessor  # noqa: F401
from .model import (CoughDetector, CoughDetectorResidual, CoughDetectorSmall, ConvBlock, ResidualBlock, create_model,  # noqa: F401
                    count_parameters)
from .inference import CoughDetectorInference, smooth_and_debounce  # noqa: F401
from .augmentation import SpecAugment  # noqa: F401
from .dataset import CoughDataset  # noqa: F401

__all__ = ["AudioPreprocessor", "RealtimePreprocessor", "create_preprocessor", "CoughDetectorResidual",
           "ResidualBlock", "CoughDetector", "CoughDetectorSmall", "ConvBlock", "create_model", "count_parameters", "CoughDetectorInference", "smooth_and_debounce", "SpecAugment", "CoughDataset"]
