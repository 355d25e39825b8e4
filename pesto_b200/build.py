"""Build libpesto_b200.so in-tree with nvcc for sm_100a (no torch headers, plain C ABI).

    python -m pesto_b200.build [--force]

The shared library is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpesto_b200.so")
SOURCES = ["cabi.cu", "knn.cu", "prologue.cu", "state_update.cu", "state_update_tc.cu", "node_umma.cu", "pool.cu", "pdb_io.cu", "rmma_probe.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "pesto_b200.h"))
    return hdrs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src):
    obj = os.path.join(CSRC, src[:-3] + ".o")
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(obj + ".log", "w") as fh:
        fh.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build_library(force=False, verbose=False):
    hdrs = _deps()
    todo = [s for s in SOURCES
            if force or _stale(os.path.join(CSRC, s[:-3] + ".o"), [os.path.join(CSRC, s)] + hdrs)]
    if todo:
        if verbose:
            print("nvcc:", " ".join(todo), flush=True)
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1)) as ex:
            list(ex.map(_compile, todo))
    objs = [os.path.join(CSRC, s[:-3] + ".o") for s in SOURCES]
    if todo or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def build_variant(name, defines, src="state_update_tc.cu"):
    """Experiment aid: the library with `src` (one file name or a list) recompiled under extra -D defines
    -> libpesto_b200.<name>.so."""
    build_library()
    srcs = [src] if isinstance(src, str) else list(src)
    objs = {}
    for sfile in srcs:
        obj = os.path.join(CSRC, f"{sfile[:-3]}.{name}.o")
        cmd = [NVCC] + FLAGS + [f"-D{d}" for d in defines] + ["-c", os.path.join(CSRC, sfile), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as fh:
            fh.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for variant {name}:\n{r.stdout}\n{r.stderr}")
        objs[sfile] = obj
    link = [objs.get(s, os.path.join(CSRC, s[:-3] + ".o")) for s in SOURCES]
    lib = os.path.join(HERE, f"libpesto_b200.{name}.so")
    r = subprocess.run([NVCC, "-shared", "-o", lib] + link + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
