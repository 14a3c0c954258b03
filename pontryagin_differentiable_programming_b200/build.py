"""nvcc driver: compiles generated translation units and the C-ABI library for sm_100a, IN-TREE.

Built objects live under ``pontryagin_differentiable_programming_b200/_modules`` (system modules,
keyed by a hash of their source) and ``pontryagin_differentiable_programming_b200/csrc``
(``libpdp_b200.so``) so they travel to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
MODULE_DIR = os.path.join(PKG_DIR, "_modules")
CSRC_DIR = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(CSRC_DIR, "libpdp_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]


def nvcc_path():
    for cand in (os.environ.get("PDP_B200_NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: the PDP B200 engine compiles its per-system CUDA modules with nvcc")


def _run(cmd, what):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError("%s failed (%d):\n%s\n%s" % (what, p.returncode, " ".join(cmd), p.stdout[-4000:]))
    return p.stdout


def compile_module(source: str, key: str, verbose: bool = False, extra_flags=()) -> str:
    """Compile a generated module; returns the path of the .so (cached by ``key``)."""
    os.makedirs(MODULE_DIR, exist_ok=True)
    so = os.path.join(MODULE_DIR, "pdpmod_%s.so" % key)
    cu = os.path.join(MODULE_DIR, "pdpmod_%s.cu" % key)
    if os.path.isfile(so) and os.path.isfile(cu) and open(cu).read() == source:
        return so
    # several processes (torchrun ranks, pytest-xdist workers) may generate the same system at the same time: each writes
    # its OWN temporary copy of the source and publishes it under the final name by an atomic rename (identical content, so
    # the last rename wins harmlessly and nvcc never reads a half-written file; compiling the final path keeps -lineinfo
    # pointing at a file that exists); the .so is published the same way
    tmp_cu = os.path.join(MODULE_DIR, "pdpmod_%s.tmp%d.cu" % (key, os.getpid()))
    tmp = so + ".tmp.%d" % os.getpid()
    with open(tmp_cu, "w") as f:
        f.write(source)
    os.replace(tmp_cu, cu)
    try:
        cmd = [nvcc_path()] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, cu]
        out = _run(cmd, "nvcc (system module %s)" % key)
        os.replace(tmp, so)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    if verbose:
        print(out)
    return so


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/pdp_b200.cu -> libpdp_b200.so (the C-ABI of include/pdp_b200.h)."""
    src = os.path.join(CSRC_DIR, "pdp_b200.cu")
    inc = os.path.join(os.path.dirname(PKG_DIR), "include")
    if not force and os.path.isfile(LIB_PATH) and os.path.getmtime(LIB_PATH) >= max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(inc, "pdp_b200.h"))):
        return LIB_PATH
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-I", inc] + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, src, "-ldl"]
    out = _run(cmd, "nvcc (libpdp_b200.so)")
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(out)
    return LIB_PATH


def build_microbench(force: bool = False) -> str:
    """tools/microbench/fp64_peak.cu -> libpdp_microbench.so (measures the FP64 FMA peak that bench.py reports beside the
    HBM roofline; bench / test infrastructure, not part of the product path)."""
    d = os.path.join(os.path.dirname(PKG_DIR), "tools", "microbench")
    src, so = os.path.join(d, "fp64_peak.cu"), os.path.join(d, "libpdp_microbench.so")
    if not force and os.path.isfile(so) and os.path.getmtime(so) >= os.path.getmtime(src):
        return so
    tmp = so + ".tmp.%d" % os.getpid()
    _run([nvcc_path()] + NVCC_FLAGS + ["-o", tmp, src], "nvcc (libpdp_microbench.so)")
    os.replace(tmp, so)
    return so
