from cough_detector_b200.model import (CoughDetector, CoughDetectorSmall, CoughDetectorResidual, ConvBlock, ResidualBlock,  # noqa: F401
                                        create_model, count_parameters)
