from cough_detector_b200.inference import CoughDetectorInference, smooth_and_debounce  # noqa: F401
