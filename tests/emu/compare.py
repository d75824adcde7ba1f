"""Dev tool: compare the host-emulated feature bodies with the compiled reference on random ROIs."""
import ctypes as C, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref
E = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libnyx_emu.so"))
E.nyxemu_roi_features.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_void_p, C.c_int]
cols = ref.columns()

def emu(tile, mask, g=64, ibsi=False, off=1, fam=255, p16=1):
    T = np.ascontiguousarray(tile, np.uint32); K = np.ascontiguousarray(mask, np.uint8)
    out = np.zeros(337)
    rc = E.nyxemu_roi_features(T.ctypes.data, K.ctypes.data, T.shape[1], T.shape[0], g, int(ibsi), off, fam, out.ctypes.data, p16)
    assert rc == 0
    return out

def rand_roi(rng, kind):
    h = int(rng.integers(1, 30)); w = int(rng.integers(1, 30))
    if kind == 0:   # noisy 16-bit
        t = rng.integers(1, 5000, (h, w))
    elif kind == 1: # few levels
        t = rng.integers(1, 6, (h, w))
    elif kind == 2: # smooth
        yy, xx = np.mgrid[0:h, 0:w]
        t = (1000 + 800*np.sin(xx/3.0) + 500*np.cos(yy/4.0) + rng.normal(0, 50, (h, w))).clip(1, 65535).astype(np.int64)
    else:
        t = rng.integers(0, 256, (h, w))
    m = rng.random((h, w)) < rng.choice([0.5, 0.8, 1.0])
    if not m.any(): m[0, 0] = True
    t = t.copy(); z = rng.random((h, w)) < 0.03; t[z] = 0
    return t.astype(np.uint32), m.astype(np.uint8)

def main(n=300, seed=0):
    rng = np.random.default_rng(seed)
    worst = {}
    nbad = 0
    for it in range(n):
        kind = it % 4
        t, m = rand_roi(rng, kind)
        ibsi = bool(it % 3 == 0 and kind in (1, 3))
        g = int(rng.choice([8, 64, 256])) if not ibsi else 64
        a = ref.roi_features(t, m, g, ibsi)
        b = emu(t, m, g, ibsi, p16=it % 2)
        a = np.where(np.isfinite(a), a, 0.0); b = np.where(np.isfinite(b), b, 0.0)
        err = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300)
        err = np.where((np.abs(a - b) < 1e-12), 0, err)
        for j in np.nonzero(err > 1e-9)[0]:
            nbad += 1
            key = cols[j]
            if key not in worst or worst[key][0] < err[j]:
                worst[key] = (err[j], a[j], b[j], it, kind, ibsi, g, t.shape)
    print("bad entries:", nbad)
    for k, v in sorted(worst.items(), key=lambda kv: -kv[1][0])[:60]:
        print(k, v)

if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 300, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
