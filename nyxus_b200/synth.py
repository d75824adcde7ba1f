"""Synthetic inputs of the shapes BASELINE.json names, numpy PCG64 with fixed seeds."""
import numpy as np


def _smooth_noise(rng, H, W, sigma):
    """cheap separable box-blur approximation of Gaussian-blurred white noise (no scipy dependency)"""
    a = rng.standard_normal((H, W)).astype(np.float32)
    k = max(1, int(sigma))
    for _ in range(3):
        c = np.cumsum(np.pad(a, ((k, k), (0, 0)), mode="wrap"), axis=0)
        a = (c[2 * k:] - c[:-2 * k]) / (2 * k)
        c = np.cumsum(np.pad(a, ((0, 0), (k, k)), mode="wrap"), axis=1)
        a = (c[:, 2 * k:] - c[:, :-2 * k]) / (2 * k)
    a /= (a.std() + 1e-12)
    return a


def blob_slide(H, W, pitch, rmin, rmax, seed, jitter=None, lobes=True, zero_frac=0.01, dtype=np.uint16, sigma=6):
    """Grid of blob ROIs (disks with a 5-lobe sinusoidal boundary, amplitude 3 px), label id = 1 + row-major grid index.
    intensity = clip(2000 + 1500*smooth + 800*N(0,1), 1, 65535), a fraction of in-ROI pixels forced to 0.
    Vectorised per band of grid rows so a 10k x 10k slide takes seconds."""
    rng = np.random.default_rng(seed)
    ny, nx = H // pitch, W // pitch
    if jitter is None:
        jitter = max(0, pitch // 2 - rmax - (3 if lobes else 0) - 1)
    r = rng.integers(rmin, rmax + 1, (ny, nx)).astype(np.float32)
    dx = rng.integers(-jitter, jitter + 1, (ny, nx)).astype(np.float32)
    dy = rng.integers(-jitter, jitter + 1, (ny, nx)).astype(np.float32)
    ph = rng.uniform(0, 2 * np.pi, (ny, nx)).astype(np.float32)
    label = np.zeros((H, W), np.uint32)
    xs = np.arange(nx * pitch)
    ci = xs // pitch
    lx = (xs % pitch - pitch // 2).astype(np.float32)
    for j in range(ny):
        ly = (np.arange(pitch) - pitch // 2).astype(np.float32)[:, None]
        X = lx[None, :] - dx[j, ci][None, :]
        Y = ly - dy[j, ci][None, :]
        rad = np.hypot(X, Y)
        rr = r[j, ci][None, :]
        if lobes:
            rr = rr + 3.0 * np.sin(5.0 * np.arctan2(Y, X) + ph[j, ci][None, :])
        ids = (1 + j * nx + ci).astype(np.uint32)[None, :]
        label[j * pitch:(j + 1) * pitch, :nx * pitch] = np.where(rad <= rr, ids, 0)
    smooth = _smooth_noise(rng, H, W, sigma)
    inten = 2000.0 + 1500.0 * smooth + 800.0 * rng.standard_normal((H, W)).astype(np.float32)
    inten = np.clip(inten, 1, 65535).astype(dtype)
    if zero_frac > 0:
        z = (rng.random((H, W), dtype=np.float32) < zero_frac) & (label != 0)
        inten[z] = 0
    return inten, label


def config1(seed=1):
    """C1: 1024x1024 uint16, 16x16 grid of 256 blob ROIs"""
    return blob_slide(1024, 1024, 64, 12, 22, seed)


def config2(seed=2, size=10000):
    """C2: 10k x 10k, pitch 31 -> 322x322 = 103684 nucleus-sized ROIs"""
    return blob_slide(size, size, 31, 7, 12, seed, jitter=0, lobes=True)


def config3(seed=3, size=16384):
    """C3: size x size uint16 slide, label == 1 everywhere (whole-image single ROI), intensity model of blob_slide with
    sigma = 12; run with coarse_gray_depth = 256, GLCM + GLRLM + GLSZM"""
    rng = np.random.default_rng(seed)
    smooth = _smooth_noise(rng, size, size, 12)
    inten = 2000.0 + 1500.0 * smooth
    del smooth
    inten += 800.0 * rng.standard_normal((size, size), dtype=np.float32)
    inten = np.clip(inten, 1, 65535).astype(np.uint16)
    return inten, np.ones((size, size), np.uint32)


def random_rois(H, W, nroi, seed, maxval=5000, levels=None):
    """random rectangles / noisy masks with assorted intensity statistics, for parity tests"""
    rng = np.random.default_rng(seed)
    label = np.zeros((H, W), np.uint32)
    inten = rng.integers(1, maxval, (H, W)).astype(np.uint32)
    if levels:
        inten = rng.integers(0, levels, (H, W)).astype(np.uint32)
    lab = 1
    for _ in range(nroi):
        h, w = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        y, x = int(rng.integers(0, H - h)), int(rng.integers(0, W - w))
        m = rng.random((h, w)) < rng.choice([0.4, 0.8, 1.0])
        sub = label[y:y + h, x:x + w]
        sub[m & (sub == 0)] = lab * 3 + 1      # sparse label values
        lab += 1
    z = rng.random((H, W)) < 0.02
    inten[z] = 0
    return inten, label


def ellipsoid_volume(D, H, W, pitch, amin, amax, seed, dtype=np.uint16, zero_frac=0.01, sigma=4):
    """Jittered lattice of ellipsoid ROIs (semi-axes U{amin..amax}); label id = 1 + lattice index.  Intensity model as in
    blob_slide.  Returns (inten[D,H,W], label[D,H,W] uint32)."""
    rng = np.random.default_rng(seed)
    nz, ny, nx = D // pitch, H // pitch, W // pitch
    label = np.zeros((D, H, W), np.uint32)
    g = np.arange(pitch) - pitch // 2
    Z, Y, X = np.meshgrid(g, g, g, indexing="ij")
    lab = 1
    jit = max(0, pitch // 2 - amax - 1)
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                a = rng.integers(amin, amax + 1, 3).astype(np.float32)
                dz, dy, dx = rng.integers(-jit, jit + 1, 3) if jit else (0, 0, 0)
                m = ((Z - dz) / a[0]) ** 2 + ((Y - dy) / a[1]) ** 2 + ((X - dx) / a[2]) ** 2 <= 1.0
                sub = label[k * pitch:(k + 1) * pitch, j * pitch:(j + 1) * pitch, i * pitch:(i + 1) * pitch]
                sub[m] = lab
                lab += 1
    smooth = rng.standard_normal((D, H, W)).astype(np.float32)
    kk = max(1, int(sigma))
    for ax in range(3):
        for _ in range(2):
            c = np.cumsum(np.concatenate([smooth.take(range(-kk, 0), axis=ax), smooth, smooth.take(range(0, kk), axis=ax)], axis=ax), axis=ax)
            n = smooth.shape[ax]
            smooth = (c.take(range(2 * kk, 2 * kk + n), axis=ax) - c.take(range(0, n), axis=ax)) / (2 * kk)
    smooth /= (smooth.std() + 1e-12)
    inten = 2000.0 + 1500.0 * smooth + 800.0 * rng.standard_normal((D, H, W)).astype(np.float32)
    inten = np.clip(inten, 1, 65535).astype(dtype)
    if zero_frac > 0:
        z = (rng.random((D, H, W), dtype=np.float32) < zero_frac) & (label != 0)
        inten[z] = 0
    return inten, label


def config4(seed=4, size=512):
    """C4: size^3 uint16 volume, ellipsoid ROIs (semi-axes 6..14) on a pitch-39 lattice: 13^3 = 2197 ROIs at 512^3"""
    return ellipsoid_volume(size, size, size, 39, 6, 14, seed)


def config2_device(seed, device, size=10000, pitch=31, rmin=7, rmax=12):
    """configs[4] needs 64 distinct configs[1]-shaped slides; blob_slide (numpy PCG64) takes ~25 s per slide on the host, so
    the 64-slide job generates them ON THE DEVICE: same geometry model (pitch-31 grid of 5-lobe blobs, radius U{7..12},
    label = 1 + row-major grid index) and intensity model (2000 + 1500 * smoothed noise + 800 * N(0,1), clipped to
    [1, 65535], 1 % of the in-ROI pixels forced to 0), torch's Philox generator seeded with `seed` instead of PCG64.
    Returns (uint16 [size, size], int32 [size, size]) torch tensors on `device` (the int32 bits are the uint32 labels)."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    H = W = int(size)
    ny, nx = H // pitch, W // pitch
    r = torch.randint(rmin, rmax + 1, (ny, nx), generator=g, device=dev).float()
    ph = torch.rand((ny, nx), generator=g, device=dev) * (2 * 3.141592653589793)
    ys = torch.arange(H, device=dev)
    xs = torch.arange(W, device=dev)
    cj = torch.clamp(ys // pitch, max=ny - 1)
    ci = torch.clamp(xs // pitch, max=nx - 1)
    Y = (ys % pitch - pitch // 2).float()[:, None]
    X = (xs % pitch - pitch // 2).float()[None, :]
    inside = ((ys < ny * pitch)[:, None]) & ((xs < nx * pitch)[None, :])
    rr = r[cj][:, ci] + 3.0 * torch.sin(5.0 * torch.atan2(Y, X).expand(H, W) + ph[cj][:, ci])
    ids = (1 + cj[:, None] * nx + ci[None, :]).int()
    label = torch.where((torch.hypot(Y, X).expand(H, W) <= rr) & inside, ids, torch.zeros((), dtype=torch.int32, device=dev))
    del rr, ids, inside
    a = torch.randn((1, 1, H, W), generator=g, device=dev)
    for _ in range(3):
        a = torch.nn.functional.avg_pool2d(a, 13, stride=1, padding=6, count_include_pad=True)
    a = a[0, 0]
    a = a / (a.std() + 1e-12)
    inten = 2000.0 + 1500.0 * a + 800.0 * torch.randn((H, W), generator=g, device=dev)
    del a
    inten = inten.clamp_(1, 65535)
    z = (torch.rand((H, W), generator=g, device=dev) < 0.01) & (label != 0)
    inten[z] = 0
    return inten.to(torch.uint16), label
