"""B200-native feature + CoughDetectorResidual hot path of cough_detector."""
__version__ = "0.1.0"
