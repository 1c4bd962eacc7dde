from cough_detector_b200.augmentation import SpecAugment  # noqa: F401
