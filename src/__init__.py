"""Import-path shim: ``from src.preprocessing import AudioPreprocessor`` etc. resolve to the B200
implementation, so the reference's entry scripts (run_detection.py, train_with_data.py) import it
unchanged.  Only the hot-path names are provided."""
from cough_detector_b200 import (AudioPreprocessor, RealtimePreprocessor, create_preprocessor,  # noqa: F401
                                 CoughDetector, CoughDetectorSmall, CoughDetectorResidual, create_model, count_parameters)
