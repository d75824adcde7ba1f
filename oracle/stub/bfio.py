"""TEST INFRASTRUCTURE ONLY -- stand-in for the `bfio` package the reference's Python test-suite imports to read its four
sample OME-TIFFs (tests/python/test_tissuenet_data.py); this image has no bfio.  BioReader.read() decodes with OpenCV."""
import numpy as np


class BioReader:
    def __init__(self, path, *a, **k):
        self.path = str(path)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def read(self, *a, **k):
        import cv2
        img = cv2.imread(self.path, cv2.IMREAD_UNCHANGED)
        if img is None:
            from PIL import Image
            img = np.asarray(Image.open(self.path))
        return img
