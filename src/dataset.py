from cough_detector_b200.dataset import CoughDataset, decode_audio  # noqa: F401
