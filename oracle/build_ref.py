"""Builds the reference's OWN three CUDA extensions (resample2d_cuda, channelnorm_cuda, correlation_cuda) from
the sources where they lie under /root/reference, for sm_100, into oracle/_ref/ (git-ignored, travels to the
GPU box with the snapshot).  Build-container only; nothing is copied out of the reference tree.

The resulting modules are the *real* reference kernels: tests/test_ref_ext_gpu.py uses them on the GPU box to
pin oracle/flow_ops.py (and, transitively, our kernels) against the reference itself.

    python -m oracle.build_ref
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("SHINEON_REFERENCE_ROOT", "/root/reference")
PKG = os.path.join(REF, "models", "flownet2_pytorch", "networks")
EXTS = {
    "resample2d_cuda": ("resample2d_package", ["resample2d_cuda.cc", "resample2d_kernel.cu"]),
    "channelnorm_cuda": ("channelnorm_package", ["channelnorm_cuda.cc", "channelnorm_kernel.cu"]),
    "correlation_cuda": ("correlation_package", ["correlation_cuda.cc", "correlation_cuda_kernel.cu"]),
}


def built(name):
    return os.path.exists(os.path.join(OUT, name + ".so"))


def build_one(name, verbose=False):
    from torch.utils import cpp_extension

    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    sub, files = EXTS[name]
    bdir = os.path.join(OUT, "build_" + name)
    os.makedirs(bdir, exist_ok=True)
    cpp_extension.load(
        name=name, sources=[os.path.join(PKG, sub, f) for f in files], build_directory=bdir,
        extra_cflags=["-O2", "-w"],
        extra_cuda_cflags=["-O2", "-w", "-gencode", "arch=compute_100,code=sm_100", "-include",
                           os.path.join(HERE, "ref_compat.h")],
        is_python_module=False, verbose=verbose)
    so = os.path.join(bdir, name + ".so")
    os.replace(so, os.path.join(OUT, name + ".so"))


def build_if_possible(verbose=False):
    if not os.path.isdir(PKG):
        return [n for n in EXTS if built(n)]
    done = []
    for name in EXTS:
        if not built(name):
            try:
                build_one(name, verbose)
            except Exception as e:  # noqa: BLE001
                print(f"[oracle/_ref] {name}: build failed: {str(e)[-400:]}")
                continue
        done.append(name)
    return done


def load(name):
    """Import a built reference extension (GPU box or build container)."""
    import importlib.util

    import torch  # noqa: F401  (libtorch must be loaded first)

    path = os.path.join(OUT, name + ".so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build_if_possible(verbose="-v" in sys.argv))
