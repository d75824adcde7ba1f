"""The compiled pybind11 module `nyxus_b200.backend` (csrc/backend_py.cpp) on a machine WITHOUT a GPU: it loads, exports
the 23 functions of the reference module (new_bindings_py.cpp:1093-1115), and the host-side protocol -- environment per
instance id, feature-name validation, parameter and metaparameter round trips -- behaves as the reference's own Python
layer (src/nyx/python/nyxus/nyxus.py, staged unmodified by oracle/Makefile.backend) expects.  No compute calls."""
import numpy as np
import pytest

from oracle import pyref

# new_bindings_py.cpp:1093-1115
REFERENCE_EXPORTS = [
    "initialize_environment", "release_environment_imp", "featurize_directory_imp", "featurize_directory_3D_imp",
    "featurize_montage_imp", "featurize_fname_lists_imp", "featurize_fname_lists_3D_imp", "findrelations_imp",
    "gpu_available_imp", "use_gpu", "get_gpu_props", "blacklist_roi_imp", "clear_roi_blacklist_imp",
    "roi_blacklist_get_summary_imp", "customize_gabor_feature_imp", "set_if_ibsi_imp", "set_environment_params_imp",
    "get_params_imp", "arrow_is_enabled_imp", "get_arrow_file_imp", "get_parquet_file_imp", "set_metaparam_imp",
    "get_metaparam_imp",
]


def _backend():
    import nyxus_b200.backend as be
    return be


def test_backend_module_exports_the_reference_functions():
    be = _backend()
    for name in REFERENCE_EXPORTS:
        assert callable(getattr(be, name, None)), name
    if pyref.reference_available():       # the compiled reference module exports exactly the same names
        ref_be = pyref.reference().backend
        ours = {n for n in dir(be) if not n.startswith("_")}
        theirs = {n for n in dir(ref_be) if not n.startswith("_")}
        assert theirs <= ours, theirs - ours


def test_environment_protocol_without_a_device():
    be = _backend()
    iid = 0xB200_0001
    be.initialize_environment(iid, 2, ["*ALL_GLCM*", "MEAN"], 5, 1.0, 64, 4, -1, True, 10000.0, 0.0, 1.0, False, -1, 0, 1.0, 1.0, 1.0, False, False)
    be.customize_gabor_feature_imp(iid, "16", "0.1", "0.8", "0.1", "0,45,90,135", "0.025", "4,16,32,64")
    p = be.get_params_imp(iid, [])
    assert p["features"] == ["*ALL_GLCM*", "MEAN"] and p["coarse_gray_depth"] == 64 and p["ibsi"] == 1
    assert p["gabor_thetas"] == [0.0, 45.0, 90.0, 135.0] and p["gabor_kersize"] == 16
    assert be.get_params_imp(iid, ["neighbor_distance"]) == {"neighbor_distance": 5}
    be.set_environment_params_imp(iid, [], 7, -1, 128, 0, -2, -1, -1, -1, -1, 0)
    p = be.get_params_imp(iid, ["neighbor_distance", "coarse_gray_depth"])
    assert p == {"neighbor_distance": 7, "coarse_gray_depth": 128}
    be.set_if_ibsi_imp(iid, False)
    assert be.get_params_imp(iid, ["ibsi"])["ibsi"] == 0
    # metaparameters: (ok, msg) / (value, msg) tuples (new_bindings_py.cpp:1067-1086, env_metaparams.cpp:63-247)
    assert be.get_metaparam_imp(iid, "glcm/greydepth") == (64.0, "")
    assert be.set_metaparam_imp(iid, "glcm/greydepth=100") == (True, "success")
    assert be.get_metaparam_imp(iid, "glcm/greydepth") == (100.0, "")
    assert be.set_metaparam_imp(iid, "glcm/offset=2") == (True, "success")
    assert be.get_metaparam_imp(iid, "glcm/offset") == (2.0, "")
    ok, msg = be.set_metaparam_imp(iid, "glcm/greydepth=abc")
    assert ok is False and "expecting an integer" in msg
    ok, msg = be.set_metaparam_imp(iid, "glcm/nope=1")
    assert ok is False and "unrecognized feature parameter of feature glcm" in msg
    ok, msg = be.set_metaparam_imp(iid, "zzz/greydepth=1")
    assert ok is False and 'unrecognized feature "zzz"' in msg
    ok, msg = be.set_metaparam_imp(iid, "glcm/greydepth")
    assert ok is False and "expecting <paramName>=<paramVal>" in msg
    ok, msg = be.set_metaparam_imp(iid, "3ngtdm/radius=0")
    assert ok is False and "positive integer" in msg
    assert be.set_metaparam_imp(iid, "3glcm/sparseintensities=true") == (True, "success")
    assert be.arrow_is_enabled_imp(iid) is True         # Arrow IPC / Parquet through the interpreter's pyarrow
    assert be.get_arrow_file_imp(iid) == "" and be.get_parquet_file_imp(iid) == ""      # nothing written yet
    be.release_environment_imp(iid)
    # an invalid feature name is an invalid_argument -> ValueError (env_features.cpp:496,507)
    with pytest.raises(ValueError, match="not a valid 2D feature name"):
        be.initialize_environment(iid, 2, ["GLCM_NOPE"], 5, 1.0, 64, 4, -1, False, 10000.0, 0.0, 1.0, False, -1, 0, 1.0, 1.0, 1.0, False, False)
    with pytest.raises(ValueError, match="not a valid 3D feature name"):
        be.initialize_environment(iid, 3, ["GLCM_ASM"], 5, 1.0, 64, 4, -1, False, 10000.0, 0.0, 1.0, False, -1, 0, 1.0, 1.0, 1.0, False, False)
    be.release_environment_imp(iid)


@pytest.mark.skipif(not pyref.staged_available(), reason="the reference's Python layer is staged by oracle/Makefile.backend (needs /root/reference)")
def test_reference_python_layer_constructs_over_this_backend():
    nyxus = pyref.reference_on_b200()
    nyx = nyxus.Nyxus(["*ALL_GLCM*", "*ALL_INTENSITY*"], ibsi=True, coarse_gray_depth=32)
    params = nyx.get_params()
    assert params["coarse_gray_depth"] == 32 and params["ibsi"] == 1 and params["features"] == ["*ALL_GLCM*", "*ALL_INTENSITY*"]
    nyx.set_params(coarse_gray_depth=100, neighbor_distance=3)
    assert nyx.get_params("coarse_gray_depth", "neighbor_distance") == {"coarse_gray_depth": 100, "neighbor_distance": 3}
    nyx.set_metaparam("glcm/offset=3")
    assert nyx.get_metaparam("glcm/offset") == 3
    with pytest.raises(Exception):
        nyx.set_metaparam("glcm/offset=x")
    with pytest.raises(ValueError):
        nyxus.Nyxus(["NOT_A_FEATURE"])
    n3 = nyxus.Nyxus3D(["*3D_GLCM*", "3MEAN"])
    n3.set_metaparam("3glszm/greydepth=20")
    assert n3.get_metaparam("3ngtdm/radius") == 0
    # no device here: the compute call must fail loudly (no CPU fallback), never return a table
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            nyx.featurize(np.ones((8, 8), np.uint16), np.ones((8, 8), np.uint32))
