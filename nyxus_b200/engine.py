"""Thin object wrapper over the C ABI: one Engine == one nyxb200_ctx (one CUDA device)."""
import ctypes as C
import weakref

import numpy as np
from . import _lib

FAMILY_IDS = {"INT": 0, "GLCM": 1, "GLRLM": 2, "GLDZM": 3, "GLSZM": 4, "GLDM": 5, "NGLDM": 6, "NGTDM": 7}


class NyxusB200Error(RuntimeError):
    pass


class Engine:
    def __init__(self, features, device=0, coarse_gray_depth=64, ibsi=False, glcm_offset=1, glcm_symmetric=False,
                 soft_nan=0.0, verbose=0, dim=2):
        self._L = _lib.lib()
        self._ctx = C.c_void_p()
        st = _lib.Settings()
        self._L.nyxb200_default_settings(C.byref(st))
        st.device = int(device)
        st.coarse_gray_depth = int(coarse_gray_depth)
        st.ibsi = int(bool(ibsi))
        st.glcm_offset = int(glcm_offset)
        st.glcm_symmetric = int(bool(glcm_symmetric))
        st.soft_nan = float(soft_nan)
        st.verbose = int(verbose)
        st.dim = int(dim)
        self.dim = int(dim)
        rc = self._L.nyxb200_create(C.byref(st), C.byref(self._ctx))
        if rc != 0:
            msg = self._L.nyxb200_last_error(None).decode()
            self._ctx = C.c_void_p()
            if rc == -1:
                raise ValueError(msg)
            raise NyxusB200Error(msg)
        self.select(features)

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._L.nyxb200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self._L.nyxb200_last_error(self._ctx).decode()
            if rc == -1:
                raise ValueError(msg)      # reference: std::invalid_argument -> ValueError through pybind11
            raise NyxusB200Error(msg)

    def select(self, features):
        arr = (C.c_char_p * len(features))(*[f.encode() for f in features])
        self._check(self._L.nyxb200_select_features(self._ctx, arr, len(features)))

    def set_param(self, key, value):
        self._check(self._L.nyxb200_set_param(self._ctx, key.encode(), float(value)))

    @property
    def columns(self):
        n = self._L.nyxb200_num_columns(self._ctx)
        return [self._L.nyxb200_column_name(self._ctx, i).decode() for i in range(n)]

    @property
    def column_codes(self):
        n = self._L.nyxb200_num_columns(self._ctx)
        return [self._L.nyxb200_column_code(self._ctx, i) for i in range(n)]

    def _collect(self, tab, rimg, copy=True):
        """copy=True: private copies; copy=False: views of the library's pinned buffer (valid until the next call on this
        context); copy="detach": the pinned table changes owner (nyxb200_table_detach) -- the returned array keeps it
        alive and hands it back to the library's pool when it is garbage-collected, like the reference's zero-copy
        capsule (new_bindings_py.cpp:63-73)"""
        n, nc = tab.nrows, tab.ncols
        if n:
            img = np.ctypeslib.as_array(rimg, shape=(n,))
            if copy == "detach":
                img = img.copy()
                self._check(self._L.nyxb200_table_detach(self._ctx, C.byref(tab)))
                addr = C.cast(tab.data, C.c_void_p).value
                buf = (C.c_double * (n * nc)).from_address(addr)
                weakref.finalize(buf, self._L.nyxb200_host_release, addr)
                data = np.frombuffer(buf, dtype=np.float64).reshape(n, nc)
                return data, img
            data = np.ctypeslib.as_array(tab.data, shape=(n, nc))
            if copy:
                data, img = data.copy(), img.copy()
        else:
            data = np.zeros((0, nc))
            img = np.zeros((0,), np.int32)
        self._L.nyxb200_table_free(self._ctx, C.byref(tab))
        return data, img

    def featurize(self, inten, label, copy=True):
        """inten: uint16 or uint32 [N,H,W] (or [H,W]); label: uint32 same shape.  Returns (table[nrows, 2+F], image index per row)."""
        inten = np.asarray(inten)
        label = np.asarray(label)
        if inten.ndim == 2:
            inten, label = inten[None], label[None]
        if inten.shape != label.shape or inten.ndim != 3:
            raise ValueError("Intensity and label image arrays must have the same number of images with matching dimensions")
        if inten.dtype not in (np.uint16, np.uint32):
            inten = inten.astype(np.uint32)
        inten = np.ascontiguousarray(inten)
        label = np.ascontiguousarray(label, dtype=np.uint32)
        N, H, W = inten.shape
        tab = _lib.Table()
        rimg = C.POINTER(C.c_int32)()
        fn = self._L.nyxb200_featurize_2d_u16 if inten.dtype == np.uint16 else self._L.nyxb200_featurize_2d_u32
        self._check(fn(self._ctx, inten.ctypes.data, label.ctypes.data, N, H, W, C.byref(tab), C.byref(rimg)))
        return self._collect(tab, rimg, copy)

    def featurize_volume(self, inten, label):
        """3D (dim=3 engines).  inten: uint16 or uint32 [D,H,W] or [N,D,H,W]; label: uint32 same shape.
        Returns (table[nrows, 2+F], volume index per row)."""
        inten = np.asarray(inten)
        label = np.asarray(label)
        if inten.ndim == 3:
            inten, label = inten[None], label[None]
        if inten.shape != label.shape or inten.ndim != 4:
            raise ValueError("Intensity and label volumes must have matching dimensions")
        if inten.dtype not in (np.uint16, np.uint32):
            inten = inten.astype(np.uint32)
        inten = np.ascontiguousarray(inten)
        label = np.ascontiguousarray(label, dtype=np.uint32)
        N, D, H, W = inten.shape
        tab = _lib.Table()
        rimg = C.POINTER(C.c_int32)()
        fn = self._L.nyxb200_featurize_3d_u16 if inten.dtype == np.uint16 else self._L.nyxb200_featurize_3d_u32
        self._check(fn(self._ctx, inten.ctypes.data, label.ctypes.data, N, D, H, W, C.byref(tab), C.byref(rimg)))
        return self._collect(tab, rimg)

    def featurize_ptr(self, inten_ptr, inten_is_u16, label_ptr, n_img, H, W, on_device, copy=True):
        """Raw-pointer entry (bench / torch interop).  on_device: pointers are device memory of this engine's GPU.
        copy=False returns views of the library-owned pinned result buffer (valid until the next call)."""
        tab = _lib.Table()
        rimg = C.POINTER(C.c_int32)()
        if on_device:
            rc = self._L.nyxb200_featurize_2d_device(self._ctx, inten_ptr, int(inten_is_u16), label_ptr, n_img, H, W,
                                                     C.byref(tab), C.byref(rimg))
        elif inten_is_u16:
            rc = self._L.nyxb200_featurize_2d_u16(self._ctx, inten_ptr, label_ptr, n_img, H, W, C.byref(tab), C.byref(rimg))
        else:
            rc = self._L.nyxb200_featurize_2d_u32(self._ctx, inten_ptr, label_ptr, n_img, H, W, C.byref(tab), C.byref(rimg))
        self._check(rc)
        return self._collect(tab, rimg, copy)

    def featurize_list(self, inten_ptrs, inten_is_u16, label_ptrs, H, W, on_device=False, copy=True, D=None):
        """A list of same-shape slides (2D) or volumes (D given), one raw pointer each (host unless on_device).  Host
        slides are pipelined: upload of slide k+1 overlaps the compute of slide k."""
        n = len(inten_ptrs)
        if n != len(label_ptrs) or n == 0:
            raise ValueError("Intensity and label lists must have the same, non-zero length")
        ip = (C.c_void_p * n)(*[int(p) for p in inten_ptrs])
        lp = (C.c_void_p * n)(*[int(p) for p in label_ptrs])
        tab = _lib.Table()
        rimg = C.POINTER(C.c_int32)()
        if D is None:
            rc = self._L.nyxb200_featurize_2d_list(self._ctx, ip, int(inten_is_u16), lp, int(on_device), n, H, W,
                                                   C.byref(tab), C.byref(rimg))
        else:
            rc = self._L.nyxb200_featurize_3d_list(self._ctx, ip, int(inten_is_u16), lp, int(on_device), n, D, H, W,
                                                   C.byref(tab), C.byref(rimg))
        self._check(rc)
        return self._collect(tab, rimg, copy)

    def get_param(self, key):
        v = C.c_double()
        if self._L.nyxb200_get_param(self._ctx, key.encode(), C.byref(v)) != 0:
            raise ValueError("unknown parameter '%s'" % key)
        return v.value

    def timings(self):
        ms = (C.c_float * _lib.NSTAGES)()
        self._L.nyxb200_last_timings(self._ctx, ms, _lib.NSTAGES)
        return {self._L.nyxb200_stage_name(i).decode(): float(ms[i]) for i in range(_lib.NSTAGES)}

    def counts(self):
        v = (C.c_int64 * 4)()
        self._L.nyxb200_last_counts(self._ctx, v, 4)
        return {"launches": int(v[0]), "rois": int(v[1]), "fg_pixels": int(v[2]), "pixels": int(v[3])}

    # ---- test hooks
    def debug_rois(self):
        n = C.c_int64()
        self._check(self._L.nyxb200_debug_rois(self._ctx, None, 0, C.byref(n)))
        out = np.zeros((n.value, 8), np.int64)
        if n.value:
            self._check(self._L.nyxb200_debug_rois(self._ctx, out.ctypes.data, n.value, C.byref(n)))
        return out

    def debug_cloud(self, label):
        n = C.c_int64()
        self._check(self._L.nyxb200_debug_cloud(self._ctx, int(label), None, 0, C.byref(n)))
        out = np.zeros(n.value, np.uint32)
        self._check(self._L.nyxb200_debug_cloud(self._ctx, int(label), out.ctypes.data, n.value, C.byref(n)))
        return out

    def debug_matrix(self, label, family):
        """Sparse matrix of one ROI: sorted list of (angle, a, b, count); for NGTDM also the S values."""
        fid = FAMILY_IDS[family]
        n = C.c_int64()
        self._check(self._L.nyxb200_debug_matrix(self._ctx, int(label), fid, None, 0, C.byref(n), None))
        trip = np.zeros((max(n.value, 1), 4), np.int64)
        sv = np.zeros(max(n.value, 1), np.float64)
        self._check(self._L.nyxb200_debug_matrix(self._ctx, int(label), fid, trip.ctypes.data, n.value, C.byref(n), sv.ctypes.data))
        t = [tuple(int(x) for x in row) for row in trip[:n.value]]
        return t, sv[:n.value].tolist()


def shard_slides(n_dev, n_img, weight=None):
    """device that nyxb200_featurize_2d_list_multi gives every slide (LPT on `weight`; equal weights: k % n_dev)"""
    L = _lib.lib()
    owner = (C.c_int32 * max(n_img, 1))()
    w = None
    if weight is not None:
        w = (C.c_int64 * n_img)(*[int(x) for x in weight])
    if L.nyxb200_shard_slides(int(n_dev), int(n_img), w, owner) != 0:
        raise ValueError("bad shard arguments")
    return [int(owner[i]) for i in range(n_img)]


def featurize_list_multi(engines, inten_ptrs, inten_is_u16, label_ptrs, H, W, on_device=False, weight=None, copy=True):
    """One job over several devices (one Engine each, same features): the slides are sharded by `weight` (pixel count),
    the rows of all slides land in ONE host table in slide order.  Returns (table, slide index per row)."""
    n = len(inten_ptrs)
    if n != len(label_ptrs) or n == 0 or not engines:
        raise ValueError("Intensity and label lists must have the same, non-zero length")
    L = engines[0]._L
    cs = (C.c_void_p * len(engines))(*[e._ctx.value for e in engines])
    ip = (C.c_void_p * n)(*[int(p) for p in inten_ptrs])
    lp = (C.c_void_p * n)(*[int(p) for p in label_ptrs])
    w = None
    if weight is not None:
        w = (C.c_int64 * n)(*[int(x) for x in weight])
    tab = _lib.Table()
    rimg = C.POINTER(C.c_int32)()
    rc = L.nyxb200_featurize_2d_list_multi(cs, len(engines), ip, int(inten_is_u16), lp, int(on_device), n, H, W, w,
                                           C.byref(tab), C.byref(rimg))
    engines[0]._check(rc)
    return engines[0]._collect(tab, rimg, copy)
