"""Compares the device code (normalised SASS, tests/test_sass_pin_cpu.py::sass_digest) of every csrc/*.cu of the working
tree with the same file at a given commit:   python tools/check_device_code_unchanged.py e8f944c
Cross-compiles with nvcc (no GPU needed).  Use it after host-side edits made without a GPU to show that the kernels the
last GPU run verified are still byte-identical."""
import importlib.util
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("pin", os.path.join(ROOT, "tests", "test_sass_pin_cpu.py"))
pin = importlib.util.module_from_spec(spec)
spec.loader.exec_module(pin)
commit = sys.argv[1] if len(sys.argv) > 1 else "HEAD"
CSRC = "deepsphere-cosmo-tf2_b200/csrc"
tmp = tempfile.mkdtemp()
old = os.path.join(tmp, CSRC)
os.makedirs(old)
os.makedirs(os.path.join(tmp, "include"))
names = subprocess.run(["git", "ls-tree", "--name-only", commit, CSRC + "/"], cwd=ROOT, check=True, capture_output=True,
                       text=True).stdout.split()
for n in names + ["include/deepsphere_b200.h"]:
    with open(os.path.join(tmp, n), "w") as f:
        f.write(subprocess.run(["git", "show", f"{commit}:{n}"], cwd=ROOT, check=True, capture_output=True, text=True).stdout)


def digests(d):
    out = {}
    for f in sorted(os.listdir(d)):
        if f.endswith(".cu"):
            obj = os.path.join(tmp, f"{abs(hash(d))}_{f}.o")
            r = subprocess.run([pin.NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                                "-c", f, "-o", obj], cwd=d, capture_output=True, text=True)
            out[f] = pin.sass_digest(obj) if r.returncode == 0 else "compile failed"
    return out


a, b = digests(old), digests(os.path.join(ROOT, CSRC))
bad = 0
for f in sorted(set(a) | set(b)):
    state = "new file" if f not in a else ("removed" if f not in b else ("same" if a[f] == b[f] else "DIFFERENT"))
    bad += state == "DIFFERENT"
    print(f"{f:26s} {state}")
sys.exit(1 if bad else 0)
