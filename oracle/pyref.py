"""TEST INFRASTRUCTURE ONLY -- loaders for the reference's own Python package (src/nyx/python/nyxus, staged UNMODIFIED
under oracle/_ref by oracle/Makefile.backend; nothing of it is in the repository's history):

  reference()          the complete reference: its nyxus.py over ITS OWN compiled pybind `backend` (SURVEY 8c option B)
  reference_on_b200()  the same unmodified nyxus.py over the PRODUCT's pybind module nyxus_b200/backend*.so -- the
                       drop-in claim: `from .backend import ...` in the reference's nyxus.py resolves to our module.
Only tests/ may import this."""
import importlib
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
PYREF = os.path.join(_HERE, "_ref", "pyref", "nyxus")
PYB200 = os.path.join(_HERE, "_ref", "pyb200", "nyxus")


def _load(pkgname, pkgdir, backend=None):
    if pkgname in sys.modules:
        return sys.modules[pkgname]
    spec = importlib.util.spec_from_file_location(pkgname, os.path.join(pkgdir, "__init__.py"), submodule_search_locations=[pkgdir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[pkgname] = mod
    if backend is not None:
        sys.modules[pkgname + ".backend"] = backend     # what `from .backend import (...)` will find
    try:
        spec.loader.exec_module(mod)
    except Exception:
        sys.modules.pop(pkgname, None)
        sys.modules.pop(pkgname + ".backend", None)
        raise
    return mod


def reference_available():
    return os.path.isdir(PYREF) and any(f.startswith("backend") and f.endswith(".so") for f in os.listdir(PYREF))


def staged_available():
    return os.path.exists(os.path.join(PYB200, "nyxus.py"))


def reference():
    return _load("nyxus_reference", PYREF)


def reference_on_b200():
    backend = importlib.import_module("nyxus_b200.backend")
    return _load("nyxus_on_b200", PYB200, backend)
