"""Known-answer fixtures of the reference's OWN test-suite for this path, restated as data (small literal arrays).
Sources (under /root/reference/tests):
  test_data.h:122-171   IBSI digital phantom, 4 slices of 5 x 4 (intensity + mask)
  test_data.h:183-199   IBSI fig 3.19 NGLDM sample image + reference matrix
  test_data.h:212-230   non-IBSI NGLDM sample image + reference matrix
  test_data.h:243-265   IBSI fig 3.17 GLDZM sample image + reference matrix
  test_2d_glcm_ibsi.h:17-49  IBSI consensus GLCM values (2D, averaged over slices and angles; checked at 1 %)
"""
import numpy as np

IBSI_PHANTOM_INT = [
    np.array([[1, 4, 4, 1, 1], [1, 4, 6, 1, 1], [4, 1, 6, 4, 1], [4, 4, 6, 4, 1]]),
    np.array([[1, 4, 4, 1, 1], [1, 1, 6, 1, 1], [1, 1, 3, 1, 1], [4, 4, 6, 1, 1]]),
    np.array([[1, 4, 4, 1, 1], [1, 1, 1, 1, 1], [1, 1, 9, 1, 1], [1, 1, 6, 1, 1]]),
    np.array([[1, 4, 4, 1, 1], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, 6, 1, 1]]),
]
IBSI_PHANTOM_MASK = [
    np.array([[1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1]]),
    np.array([[1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [0, 1, 1, 1, 1], [1, 1, 1, 1, 1]]),
    np.array([[1, 1, 1, 0, 0], [1, 1, 1, 1, 1], [1, 1, 0, 1, 1], [1, 1, 1, 1, 1]]),
    np.array([[1, 1, 1, 0, 0], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1]]),
]

# IBSI consensus values, GLCM, "2D, averaged" (test_2d_glcm_ibsi.h:17-49); the reference asserts them at 1 %
GLCM_IBSI_CONSENSUS = {
    "GLCM_ACOR": 5.09, "GLCM_ASM": 0.368, "GLCM_CLUPROM": 79.1, "GLCM_CLUSHADE": 7, "GLCM_CLUTEND": 5.47, "GLCM_CONTRAST": 5.28,
    "GLCM_CORRELATION": -0.0121, "GLCM_DIFAVE": 1.42, "GLCM_DIFENTRO": 1.4, "GLCM_DIFVAR": 2.9, "GLCM_DIS": 1.42, "GLCM_ID": 0.678,
    "GLCM_IDN": 0.851, "GLCM_IDM": 0.619, "GLCM_IDMN": 0.899, "GLCM_INFOMEAS1": -0.155, "GLCM_INFOMEAS2": 0.487, "GLCM_HOM2": 0.619,
    "GLCM_ENTROPY": 2.05, "GLCM_IV": 0.0567, "GLCM_JAVE": 2.14, "GLCM_JE": 2.05, "GLCM_JMAX": 0.519, "GLCM_JVAR": 2.69,
    "GLCM_SUMAVERAGE": 4.28, "GLCM_SUMENTROPY": 1.6, "GLCM_SUMVARIANCE": 5.47, "GLCM_VARIANCE": 2.69,
}

NGLDM_IBSI_IMAGE = np.array([[1, 2, 2, 3], [1, 2, 3, 3], [4, 2, 4, 1], [4, 1, 2, 3]])
# rows = grey level 1..4, columns = number of matching neighbours 0..3 (test_data.h:192-199)
NGLDM_IBSI_MATRIX = np.array([[2, 2, 0, 0], [0, 1, 3, 1], [1, 0, 3, 0], [1, 2, 0, 0]])

NGLDM_NONIBSI_IMAGE = np.array([[91, 91, 92, 93, 91], [90, 91, 91, 92, 92], [90, 90, 92, 92, 91], [93, 93, 92, 92, 91], [90, 90, 92, 90, 91]])
# rows = sorted unique binned levels, columns = matching neighbours 0..5 (test_data.h:223-230)
NGLDM_NONIBSI_MATRIX = np.array([[1, 2, 3, 0, 0, 0], [1, 2, 3, 2, 0, 0], [0, 1, 2, 0, 4, 1], [1, 2, 0, 0, 0, 0]])

GLDZM_IBSI_IMAGE = np.array([[1, 2, 2, 3], [1, 2, 3, 3], [4, 2, 4, 1], [4, 1, 2, 3]])
# rows = grey level 1..4, columns = zone distance 1..2 (test_data.h:258-265)
GLDZM_IBSI_MATRIX = np.array([[3, 0], [2, 0], [2, 0], [1, 1]])

# tests/python/test_data.py:3-46 -- `intens` / `seg` of the reference's Python test-suite (the IBSI phantom as the Python
# tests hold it: slice 3 carries a 6 where test_data.h has a 9)
PY_INTENS = np.array([
    [[1, 4, 4, 1, 1], [1, 4, 6, 1, 1], [4, 1, 6, 4, 1], [4, 4, 6, 4, 1]],
    [[1, 4, 4, 1, 1], [1, 1, 6, 1, 1], [1, 1, 3, 1, 1], [4, 4, 6, 1, 1]],
    [[1, 4, 4, 1, 1], [1, 1, 1, 1, 1], [1, 1, 6, 1, 1], [1, 1, 6, 1, 1]],
    [[1, 4, 4, 1, 1], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, 6, 1, 1]],
])
PY_SEG = np.array([
    [[1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1]],
    [[1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [0, 1, 1, 1, 1], [1, 1, 1, 1, 1]],
    [[1, 1, 1, 0, 0], [1, 1, 1, 1, 1], [1, 1, 0, 1, 1], [1, 1, 1, 1, 1]],
    [[1, 1, 1, 0, 0], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1]],
])
