"""The reference's three pybind11 extension modules — `resample2d_cuda`, `channelnorm_cuda`, `correlation_cuda`
(models/flownet2_pytorch/networks/*_package/*_cuda.cc) — as thin modules over the C ABI, so the reference's UNTOUCHED
Python wrappers (resample2d.py, channelnorm.py, correlation.py) run on libshineon_b200.so.

    import shineon_virtual_tryon_b200.legacy_shims as shims
    shims.install()            # registers the three module names in sys.modules (before the reference imports them)

Same calling contract as the originals: torch CUDA tensors, caller-allocated outputs (correlation's outputs and scratch
are resized here like correlation_cuda.cc:36-42 does), work enqueued on the current stream, `forward` / `backward`
return 1.  Stricter on purpose: a failed launch or bad argument raises RuntimeError (the reference swallows them,
correlation_cuda.cc:80-83).
"""
import ctypes as C
import sys
import types

import torch

from . import _lib


def _p(t):
    return C.c_void_p(t.data_ptr())


def _s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(rc, what):
    if rc != 0:
        msg = _lib.load().shineon_last_error()
        raise RuntimeError(f"{what}: {msg.decode() if msg else rc}")


def _cuda_f32(*ts):
    for t in ts:
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError("shineon legacy shim: expected contiguous float32 CUDA tensors")


# ---- resample2d_cuda.cc:6-31
def resample2d_forward(input1, input2, output, kernel_size, bilinear):
    _cuda_f32(input1, input2, output)
    b, d, h, w = output.shape
    _check(_lib.load().shineon_resample2d_fwd(_p(input1), _p(input2), _p(output), b, d, input1.shape[2], input1.shape[3],
                                              h, w, int(kernel_size), int(bool(bilinear)), _s()), "resample2d_cuda.forward")
    return 1


def resample2d_backward(input1, input2, grad_output, grad_input1, grad_input2, kernel_size, bilinear):
    _cuda_f32(input1, input2, grad_output, grad_input1, grad_input2)
    b, d, h, w = grad_output.shape
    _check(_lib.load().shineon_resample2d_bwd(_p(input1), _p(input2), _p(grad_output), _p(grad_input1), _p(grad_input2), b, d,
                                              input1.shape[2], input1.shape[3], h, w, int(kernel_size), int(bool(bilinear)),
                                              _s()), "resample2d_cuda.backward")
    return 1


# ---- channelnorm_cuda.cc:6-30
def channelnorm_forward(input1, output, norm_deg):
    _cuda_f32(input1, output)
    b, c, h, w = input1.shape
    _check(_lib.load().shineon_channelnorm_fwd(_p(input1), _p(output), b, c, h, w, int(norm_deg), _s()),
           "channelnorm_cuda.forward")
    return 1


def channelnorm_backward(input1, output, grad_output, grad_input1, norm_deg):
    _cuda_f32(input1, output, grad_output, grad_input1)
    b, c, h, w = input1.shape
    _check(_lib.load().shineon_channelnorm_bwd(_p(input1), _p(output), _p(grad_output), _p(grad_input1), b, c, h, w,
                                               int(norm_deg), _s()), "channelnorm_cuda.backward")
    return 1


# ---- correlation_cuda.cc:10-167 (rbot1 / rbot2 = the reference's padded NHWC scratch copies: not needed, left empty)
def correlation_forward(input1, input2, rbot1, rbot2, output, pad_size, kernel_size, max_displacement, stride1, stride2,
                        corr_type_multiply):
    _cuda_f32(input1, input2)
    b, c, h, w = input1.shape
    oc, oh, ow = C.c_int(), C.c_int(), C.c_int()
    lib = _lib.load()
    _check(lib.shineon_correlation_out_shape(c, h, w, pad_size, kernel_size, max_displacement, stride1, stride2,
                                             C.byref(oc), C.byref(oh), C.byref(ow)), "correlation_cuda.forward (shape)")
    output.resize_(b, oc.value, oh.value, ow.value)
    _check(lib.shineon_correlation_fwd(_p(input1), _p(input2), _p(output), b, c, h, w, pad_size, kernel_size,
                                       max_displacement, stride1, stride2, _s()), "correlation_cuda.forward")
    return 1


def correlation_backward(input1, input2, rbot1, rbot2, grad_output, grad_input1, grad_input2, pad_size, kernel_size,
                         max_displacement, stride1, stride2, corr_type_multiply):
    _cuda_f32(input1, input2)
    b, c, h, w = input1.shape
    grad_input1.resize_(b, c, h, w)
    grad_input2.resize_(b, c, h, w)
    go = grad_output.contiguous()
    _check(_lib.load().shineon_correlation_bwd(_p(input1), _p(input2), _p(go), _p(grad_input1), _p(grad_input2), b, c, h, w,
                                               pad_size, kernel_size, max_displacement, stride1, stride2, _s()),
           "correlation_cuda.backward")
    return 1


MODULES = {
    "resample2d_cuda": (resample2d_forward, resample2d_backward),
    "channelnorm_cuda": (channelnorm_forward, channelnorm_backward),
    "correlation_cuda": (correlation_forward, correlation_backward),
}


def install(force=True):
    """Register the three extension-module names.  force=False keeps modules that are already importable."""
    out = {}
    for name, (fwd, bwd) in MODULES.items():
        if not force and name in sys.modules:
            out[name] = sys.modules[name]
            continue
        m = types.ModuleType(name)
        m.__doc__ = f"shineon_b200 shim of the reference's {name} extension"
        m.forward, m.backward = fwd, bwd
        sys.modules[name] = m
        out[name] = m
    return out
