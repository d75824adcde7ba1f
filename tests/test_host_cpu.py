"""CPU tests (no GPU): the C-ABI library loads and exports every declared symbol, the host-side mirror validates its
arguments like the reference, the product refuses to run without a device, and the multi-rank sharding works (gloo)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import nyx_testutil as tu

ROOT = tu.ROOT


def _lib():
    from nyxus_b200 import build
    build.build()
    from nyxus_b200 import _lib
    return _lib.lib(), _lib


def test_cabi_exports_every_declared_symbol():
    L, mod = _lib()
    hdr = open(os.path.join(ROOT, "include", "nyxus_b200.h")).read()
    declared = set(re.findall(r"\b(nyxb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(L, name), "libnyxus_b200.so does not export " + name
    assert declared == set(mod.EXPORTS)


def test_library_contains_sm100a_code_only():
    _lib()
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "nyxus_b200", "libnyxus_b200.so")], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out, out


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from nyxus_b200 import Engine, NyxusB200Error
    with pytest.raises(NyxusB200Error) as e:
        Engine(["*ALL_GLCM*"])
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nyxus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", "").replace("the oracle", "") or f in ("nyx_common.h",), f


def test_nyxus_argument_validation_matches_reference():
    """reference nyxus.py:200-224 -- these raise before the backend is touched"""
    from nyxus_b200 import Nyxus
    with pytest.raises(ValueError, match="Neighbor distance"):
        Nyxus(["*ALL_GLCM*"], neighbor_distance=0)
    with pytest.raises(ValueError, match="Pixels per micron"):
        Nyxus(["*ALL_GLCM*"], pixels_per_micron=0)
    with pytest.raises(ValueError, match="grayscale levels"):
        Nyxus(["*ALL_GLCM*"], coarse_gray_depth=0)
    with pytest.raises(ValueError, match="at least one feature calculation thread"):
        Nyxus(["*ALL_GLCM*"], n_feature_calc_threads=0)
    with pytest.raises(ValueError, match="verbosity"):
        Nyxus(["*ALL_GLCM*"], verbose=-1)


def test_feature_table_layout():
    from nyxus_b200 import feature_table as ft
    # + *BASIC_MORPHOLOGY*, *ALL_IH*, the unweighted and the distance-to-contour weighted geometric moments (appended);
    # + the multi-valued HISTOGRAM feature (no fixed columns: one per grey-depth bin, after the fixed blocks)
    assert len(ft.CANONICAL_COLUMNS) == 337 + 15 + 46 + 118 + 62 and len(ft.FEATURES) == 199 + 15 + 46 + 118 + 62 + 1
    assert [f for f in ft.FEATURES if f[4] == 0] == [("HISTOGRAM", 492, "HIST", 578, 0)]
    assert ft.CANONICAL_COLUMNS[0] == "COV" and ft.CANONICAL_COLUMNS[35] == "UNIFORMITY_PIU"
    assert ft.CANONICAL_COLUMNS[36:40] == ["GLCM_ASM_0", "GLCM_ASM_45", "GLCM_ASM_90", "GLCM_ASM_135"]
    assert ft.CANONICAL_COLUMNS[332:337] == ["NGTDM_COARSENESS", "NGTDM_CONTRAST", "NGTDM_BUSYNESS", "NGTDM_COMPLEXITY", "NGTDM_STRENGTH"]
    assert ft.CANONICAL_COLUMNS[337] == "AREA_PIXELS_COUNT" and ft.CANONICAL_COLUMNS[351] == "ASPECT_RATIO"
    assert ft.CANONICAL_COLUMNS[352] == "IH_MEAN_VAL" and ft.CANONICAL_COLUMNS[397] == "IH_BIN_SIZE"
    assert ft.CANONICAL_COLUMNS[398] == "SPAT_MOMENT_00" and ft.CANONICAL_COLUMNS[515] == "IMOM_HU7"
    assert ft.CANONICAL_COLUMNS[516] == "WEIGHTED_SPAT_MOMENT_00" and ft.CANONICAL_COLUMNS[-1] == "IMOM_WHU7"
    codes = [f[1] for f in ft.FEATURES]
    assert codes == sorted(codes) and codes[0] == 0 and codes[36] == 36 and codes[51] == 144   # COV, AREA_PIXELS_COUNT, GLCM_ASM


def test_lpt_partition_balances_and_is_deterministic():
    from nyxus_b200.shard import lpt_partition
    w = [100, 90, 80, 10, 10, 10, 5, 5]
    parts = lpt_partition(w, 2)
    assert sorted(i for p in parts for i in p) == list(range(len(w)))
    loads = [sum(w[i] for i in p) for p in parts]
    assert max(loads) <= (4.0 / 3.0) * (sum(w) / 2.0)      # Graham's LPT bound
    assert parts == lpt_partition(w, 2)
    assert lpt_partition([], 3) == [[], [], []]


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
from nyxus_b200.shard import featurize_sharded
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
rng = np.random.default_rng(3)
slides = []
for k in range(5):
    h = 8 + 4 * k
    lab = np.zeros((h, 16), np.uint32); lab[1:h-1, 2:9] = 3; lab[2:5, 10:15] = 7 + k
    slides.append((rng.integers(1, 100, (h, 16)).astype(np.uint32), lab))
def fake_featurize(inten, label):       # stands in for Engine.featurize: [label, t_index, area, sum]
    rows = []
    for l in np.unique(label[label > 0]):
        m = label == l
        rows.append([float(l), 0.0, float(m.sum()), float(inten[m].sum())])
    return np.array(rows)
res = featurize_sharded(slides, fake_featurize, rank, world)
if rank == 0:
    tab, idx = res
    exp = np.concatenate([fake_featurize(*s) for s in slides])
    exp_idx = np.concatenate([np.full(2, i) for i in range(5)])
    assert np.array_equal(tab, exp) and np.array_equal(idx, exp_idx), (tab, exp)
    print("SHARD_OK")
else:
    assert res is None
dist.destroy_process_group()
'''


def test_sharded_featurize_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SHARD_OK" in r.stdout, r.stdout + r.stderr


def test_cabi_shard_slides_is_lpt_and_ascending():
    """nyxb200_shard_slides (the device-free half of nyxb200_featurize_2d_list_multi): equal weights -> slide k on device k % n;
    weights -> longest-processing-time-first within Graham's bound; deterministic"""
    from nyxus_b200.engine import shard_slides
    assert shard_slides(4, 10) == [k % 4 for k in range(10)]
    assert shard_slides(1, 5) == [0] * 5
    w = [100, 90, 80, 10, 10, 10, 5, 5, 70, 1]
    own = shard_slides(3, len(w), w)
    assert own == shard_slides(3, len(w), w) and set(own) == {0, 1, 2}
    loads = [sum(w[i] for i in range(len(w)) if own[i] == d) for d in range(3)]
    assert max(loads) <= (4.0 / 3.0) * (sum(w) / 3.0) + max(w) / 3.0
    with pytest.raises(ValueError):
        shard_slides(0, 3)


def test_cabi_device_free_validation_and_device_count():
    L, _ = _lib()
    msg = C.create_string_buffer(512)

    def val(dim, names):
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        return L.nyxb200_validate_features(dim, arr, len(names), msg, 512), msg.value.decode()
    assert val(2, ["*ALL_TEXTURE*", "*ALL_INTENSITY*", "GLCM_ASM", "MEAN"])[0] == 0
    rc, m = val(2, ["GLCM_NOPE"])
    assert rc == -1 and "not a valid 2D feature name" in m
    rc, m = val(3, ["*3D_ALL_MORPHOLOGY*"])
    assert rc == -3 and "outside this path" in m
    assert val(3, ["*3D_GLCM*", "3MEAN"])[0] == 0
    assert val(3, ["*3D_ALL_TEXTURE*", "3GLRLM_SRE_AVE", "3GLDZM_ZDE", "3NGLDM_DCP"])[0] == 0
    assert val(4, ["MEAN"])[0] == -1
    import torch
    if not torch.cuda.is_available():
        assert L.nyxb200_device_count() == 0


def test_arrow_file_name_rules_and_writer(tmp_path):
    """the reference's get_arrow_filename cases 1-6 (arrow_helpers.cpp:11-31) and the typed table (output_writers.cpp:21-30)"""
    import pandas as pd
    import pyarrow as pa
    from nyxus_b200 import nyxus as N
    d = str(tmp_path)
    assert N.arrow_filename("", False) == "NyxusFeatures.arrow"                                     # case 6
    assert N.arrow_filename(d, True) == os.path.join(d, "NyxusFeatures.parquet")                    # case 1
    assert N.arrow_filename(d + "/new/", False) == os.path.join(d, "new", "NyxusFeatures.arrow")    # case 1 (trailing separator)
    assert N.arrow_filename(d + "/bar", False) == d + "/bar.arrow"                                  # case 2
    assert N.arrow_filename(d + "/foo/bar", False) == d + "/foo/bar/NyxusFeatures.arrow"            # case 3
    assert N.arrow_filename(d + "/bar.csv", True) == d + "/bar.parquet"                             # case 5

    class Holder:
        pass
    df = pd.DataFrame({"intensity_image": ["a", "b"], "mask_image": ["m", "n"], "ROI_label": np.array([1, 2], np.uint32),
                       "t_index": [0., 0.], "MEAN": [1.5, 2.5]})
    h = Holder()
    p = N._emit(h, df, "arrowipc", d + "/sub/f.arrow")
    t = pa.ipc.open_file(p).read_all()
    assert [str(f.type) for f in t.schema] == ["string", "string", "int32", "double", "double"] and h._arrow_path == p
    assert t.to_pandas().MEAN.tolist() == [1.5, 2.5]
