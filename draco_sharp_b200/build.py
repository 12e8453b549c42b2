"""Builds the native pieces in-tree (nvcc cross-compiles sm_100a without a GPU).

  libdracob200.so   the product: C ABI (include/dracob200.h) + sm_100a kernels   (csrc/)
  libdrcsynth.so    synthetic .drc generator used by tests and bench.py          (synth/)

The CPU oracle under oracle/ is test infrastructure and is built by oracle/Makefile
(`build_oracle`), never linked into the product library.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdracob200.so")
SYNTH_SRC = os.path.join(HERE, "synth", "drc_synth.cpp")
SYNTH_LIB = os.path.join(HERE, "synth", "libdrcsynth.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
OBJ_DIR = os.path.join(HERE, "build")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout


def build_lib(force=False):
    """One object per .cu (compiled in parallel, rebuilt only when it or a header changed), then one link."""
    from concurrent.futures import ThreadPoolExecutor
    headers = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(ROOT, "include", "dracob200.h"))
    cus = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs = [os.path.join(OBJ_DIR, os.path.basename(c)[:-3] + ".o") for c in cus]
    todo = [(c, o) for c, o in zip(cus, objs) if force or _newer(o, [c] + headers)]
    if todo:
        with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
            list(ex.map(lambda co: _run([nvcc] + NVCC_FLAGS + ["-c", "-o", co[1], co[0]]), todo))
    if todo or _newer(LIB, objs):
        _run([nvcc] + NVCC_FLAGS + ["-shared", "-o", LIB] + objs)
    return LIB


def build_synth(force=False):
    if force or _newer(SYNTH_LIB, [SYNTH_SRC]):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", SYNTH_LIB, SYNTH_SRC])
    return SYNTH_LIB


def build_oracle(force=False):
    odir = os.path.join(ROOT, "oracle")
    if force:
        _run(["make", "-C", odir, "clean"])
    _run(["make", "-C", odir])
    return os.path.join(odir, "liboracle.so")


def build_all(force=False):
    return build_lib(force), build_synth(force)


if __name__ == "__main__":
    print(build_all("--force" in sys.argv))
