"""nvcc recipe for libshineon_b200.so (sm_100a only, in-tree, -lineinfo for ncu source pages).

Every csrc/*.cu is compiled to csrc/_build/<name>.o in parallel (only the stale ones) and linked into
csrc/libshineon_b200.so.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(CSRC, "libshineon_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "shineon_b200.h"))
    return max(os.path.getmtime(d) for d in deps)


def _compile(nvcc, src, obj, verbose):
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    return src, proc


def build(force=False, verbose=False):
    """Compile csrc/*.cu into csrc/libshineon_b200.so.  Returns the library path."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdr = _headers_mtime()
    jobs, objs = [], []
    for s in sources():
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        src = os.path.join(CSRC, s)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr):
            jobs.append((src, obj))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            results = list(ex.map(lambda j: _compile(nvcc, j[0], j[1], verbose), jobs))
        failed = [(s, p) for s, p in results if p.returncode != 0]
        for s, p in results:
            if p.returncode != 0 or verbose:
                sys.stderr.write(f"---- {os.path.basename(s)}\n{p.stdout}{p.stderr}")
        if failed:
            raise RuntimeError("nvcc failed on: " + ", ".join(os.path.basename(s) for s, _ in failed))
    if jobs or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        proc = subprocess.run([nvcc, "-shared", "-o", LIB] + objs, cwd=CSRC, capture_output=True, text=True)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
            raise RuntimeError("nvcc failed linking libshineon_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
