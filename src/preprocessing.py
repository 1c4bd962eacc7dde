from cough_detector_b200.preprocessing import *  # noqa: F401,F403
from cough_detector_b200.preprocessing import AudioPreprocessor, RealtimePreprocessor, create_preprocessor  # noqa: F401
