"""TEST INFRASTRUCTURE ONLY -- prints the translation units of the reference's `backend` pybind11 target, read from the
reference's own CMakeLists.txt (set(SOURCE ...) + the extra files of pybind11_add_module(backend ...)), relative to src/nyx."""
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
t = open(ref + "/CMakeLists.txt").read()
src = [ln.strip() for ln in re.search(r"set\(SOURCE(.*?)\n\)", t, re.S).group(1).splitlines() if ln.strip().endswith(".cpp")]
src += [ln.strip() for ln in re.search(r"pybind11_add_module\(backend(.*?)\n\t\)", t, re.S).group(1).splitlines() if ln.strip().endswith(".cpp")]
print(" ".join(s.replace("src/nyx/", "") for s in src))
