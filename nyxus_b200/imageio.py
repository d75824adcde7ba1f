"""2D image decoding for the file-based entry points (`featurize_directory`, `featurize_files`).

File decoding is host IO outside the B200 hot path.  The reference links libtiff / z5 / dcmtk
(src/nyx/image_loader.cpp, raw_tiff.h); here the interpreter's OpenCV decodes (tiled / stripped / compressed TIFF incl.
OME-TIFF pages, PNG, ...), Pillow is the fall-back.  Integer grey images only: the floating-point dynamic-range remap of
the reference's loaders is outside this path."""
import numpy as np


def read_image(path, what="intensity"):
    """-> 2D array of an unsigned integer type (uint16 kept, everything else widened to uint32)"""
    a = None
    try:
        import cv2
        a = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    except ImportError:
        a = None
    if a is None:
        try:
            from PIL import Image
            a = np.asarray(Image.open(str(path)))
        except Exception as e:
            raise IOError(f"Error: cannot decode {what} image {path}: {e}")
    if a.ndim != 2:
        raise ValueError(f"{what} image {path} is not a single-channel 2D image (shape {a.shape})")
    if np.issubdtype(a.dtype, np.floating):
        raise ValueError(f"{what} image {path} holds floating-point pixels: the dynamic-range remap of the reference's loaders "
                         "is outside the path implemented by nyxus_b200")
    if np.issubdtype(a.dtype, np.signedinteger) and a.min() < 0:
        raise ValueError(f"{what} image {path} holds negative pixels (Hounsfield-type scale): outside the path implemented by nyxus_b200")
    if a.dtype == np.uint16:
        return np.ascontiguousarray(a)
    return np.ascontiguousarray(a.astype(np.uint32))
