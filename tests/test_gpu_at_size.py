"""GPU parity tests (-m gpu) AT THE SIZES BASELINE.json names, against the compiled reference (oracle/_ref): the code that
only breaks at scale -- 32-bit counters, the band-GLCM spill path, multi-pass zone labelling, ROI batches, the size
classes -- is compared entry by entry, not through properties.

  C2  configs[1]: the full 10k x 10k slide, 103 684 ROIs, all 337 columns
  C3  configs[2] generator: single whole-image ROI, 256 grey levels: GLCM + GLRLM at 4096^2 and GLSZM at 2048^2 (the
      reference's GLSZM allocates w*h*Ng ints, glszm.cpp:211-217: 4.3 GB at 2048^2; it cannot run at 16k^2), table
      within 1e-9 and every matrix entry exact
  C4  configs[3]: the full 512^3 volume, 2 197 ROIs, 3D GLCM / GLSZM / NGTDM + first-order
Run time is dominated by the reference on the host cores (about 3 minutes in total)."""
import os

import numpy as np
import pytest

import nyx_testutil as tu
from oracle import ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnyxus_ref.so did not travel")]
NT = os.cpu_count() or 8


def _engine(features, **kw):
    from nyxus_b200 import Engine
    return Engine(list(features), **kw)


def test_config2_full_slide_against_compiled_reference():
    from nyxus_b200 import synth
    I, L = synth.config2(seed=2, size=10000)
    e = _engine(["*ALL_TEXTURE*", "*ALL_INTENSITY*"], coarse_gray_depth=64)
    try:
        tab, _ = e.featurize(I, L)
        cols = e.columns
    finally:
        e.close()
    rl, rt, _ = ref.featurize_2d(I.astype(np.uint32), L, greydepth=64, nthreads=NT, cap=110000)
    assert cols == ref.columns()
    assert tab.shape == (103684, 339) and np.array_equal(tab[:, 0].astype(np.int64), rl.astype(np.int64))
    bad = tu.mismatches(rt, tab[:, 2:], cols)
    assert not bad, (len(bad), bad[:8])


def _giant_case(size, fams, fam_mask):
    from nyxus_b200 import synth
    I, L = synth.config3(size=size)
    e = _engine(["*ALL_%s*" % f for f in fams], coarse_gray_depth=256)
    try:
        tab, _ = e.featurize(I, L)
        cols = e.columns
        rl, rt, _ = ref.featurize_2d(I.astype(np.uint32), L, greydepth=256, fam_mask=fam_mask, nthreads=NT)
        rcols = ref.columns()
        idx = [rcols.index(c) for c in cols]
        assert tab.shape[0] == 1 and int(tab[0, 0]) == 1 and list(rl) == [1]
        bad = tu.mismatches(rt[:, idx], tab[:, 2:], cols)
        assert not bad, bad[:8]
        rm = ref.roi_matrices(I.astype(np.uint32), np.ones(I.shape, bool), 256, False, fam_mask=fam_mask)
        for fam in fams:
            got, _ = e.debug_matrix(1, fam)
            assert len(got) == len(rm[fam]) and sorted(got) == sorted(rm[fam]), fam
    finally:
        e.close()


def test_config3_glcm_glrlm_4096_against_compiled_reference():
    _giant_case(4096, ["GLCM", "GLRLM"], ref.FAM["GLCM"] | ref.FAM["GLRLM"])


def test_config3_glszm_2048_against_compiled_reference():
    _giant_case(2048, ["GLSZM"], ref.FAM["GLSZM"])


def test_config4_full_volume_against_compiled_reference():
    from nyxus_b200 import Engine, synth
    I, L = synth.config4(seed=4, size=512)
    e = Engine(["*3D_ALL_INTENSITY*", "*3D_GLCM*", "*3D_NGTDM*", "*3D_GLSZM*"], coarse_gray_depth=64, dim=3)
    try:
        e.set_param("3glszm/greydepth", 64)
        e.set_param("3ngtdm/greydepth", 64)
        e.set_param("3ngtdm/radius", 1)
        tab, _ = e.featurize_volume(I, L)
        cols = e.columns
    finally:
        e.close()
    rl, rt, _ = ref.featurize_3d(I.astype(np.uint32), L, greydepth=64, glszm_greydepth=64, ngtdm_greydepth=64, ngtdm_radius=1, nthreads=NT)
    assert tab.shape[0] == len(rl) and tab.shape[0] > 2000
    assert np.array_equal(tab[:, 0].astype(np.int64), rl.astype(np.int64))
    bad = tu.mismatches(rt, tab[:, 2:], cols)
    assert not bad, (len(bad), bad[:8])
