from cough_detector_b200.model import CoughDetectorResidual, ResidualBlock, create_model, count_parameters  # noqa: F401
