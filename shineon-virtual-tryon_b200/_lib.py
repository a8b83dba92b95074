"""ctypes binding of the C ABI declared in include/shineon_b200.h.

There is no CPU or PyTorch fallback: if libshineon_b200.so is missing or does not load, every op raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libshineon_b200.so")

c_p = C.c_void_p
c_i = C.c_int
c_f = C.c_float

ACT = {None: 0, "none": 0, "relu": 1, "leaky": 2, "gelu": 3, "swish": 4, "sine": 5, "tanh": 6, "sigmoid": 7}
PAD = {"zeros": 0, "border": 1}
FMT_BF16, FMT_FP16 = 0, 1


class TpsTables(C.Structure):
    _fields_ = [("Li", c_p), ("P_X", c_p), ("P_Y", c_p), ("grid_X", c_p), ("grid_Y", c_p), ("grid_size", c_i)]


class Conv2dParams(C.Structure):
    _fields_ = [
        ("x_hi", c_p), ("x_lo", c_p), ("N", c_i), ("H", c_i), ("W", c_i), ("cin_pad", c_i), ("x_cstride", c_i),
        ("w_hi", c_p), ("w_lo", c_p), ("w_per_image", c_i), ("Cout", c_i), ("kh", c_i), ("kw", c_i), ("stride", c_i), ("pad_h", c_i), ("pad_w", c_i),
        ("Ho", c_i), ("Wo", c_i),
        ("bias", c_p), ("scale", c_p), ("shift", c_p), ("pre_act", c_i), ("post_act", c_i), ("act_param", c_f),
        ("acc_scale", c_f), ("plane_fmt", c_i),
        ("y_f32", c_p), ("y_hi", c_p), ("y_lo", c_p),
        ("out_H", c_i), ("out_W", c_i), ("out_cstride", c_i), ("out_coffset", c_i),
        ("oh_mul", c_i), ("oh_off", c_i), ("ow_mul", c_i), ("ow_off", c_i),
        ("tile_n", c_i), ("stages", c_i), ("stats_ws", c_p), ("acc_chunk_kb", c_i),
        ("splitk_ws", c_p), ("splitk_ws_bytes", C.c_size_t), ("deconv_phases", c_i),
    ]


class Conv2dWgradParams(C.Structure):
    _fields_ = [
        ("g_hi", c_p), ("g_lo", c_p), ("g_cpad", c_i), ("g_cstride", c_i),
        ("x_hi", c_p), ("x_lo", c_p), ("N", c_i), ("H", c_i), ("W", c_i), ("cin_pad", c_i), ("x_cstride", c_i),
        ("Cout", c_i), ("Cin", c_i), ("kh", c_i), ("kw", c_i), ("stride", c_i), ("pad_h", c_i), ("pad_w", c_i),
        ("Ho", c_i), ("Wo", c_i), ("plane_fmt", c_i), ("chan_map", c_p), ("mode", c_i), ("grad_w", c_p),
        ("alpha", c_f), ("beta", c_f), ("workspace", c_p), ("workspace_bytes", C.c_size_t), ("splits", c_i),
        ("desc_variant", c_i), ("g_coffset", c_i),
    ]


class FramePrepParams(C.Structure):
    _fields_ = [
        ("image", c_p), ("parse", c_p), ("cloth", c_p), ("densepose", c_p), ("pose", c_p),
        ("image_out", c_p), ("cloth_out", c_p), ("cloth_mask_out", c_p), ("densepose_out", c_p), ("agnostic_out", c_p),
        ("cocopose_out", c_p), ("im_cocopose_out", c_p),
        ("tab_bounds", c_p * 4), ("tab_kk", c_p * 4), ("tab_ksize", c_i * 4),
        ("F", c_i), ("H", c_i), ("W", c_i), ("n_joints", c_i), ("radius", c_i), ("cloth_mask_threshold", c_f),
    ]


class FramePrepPlanesParams(C.Structure):
    _fields_ = [
        ("image", c_p), ("parse", c_p), ("cloth", c_p), ("densepose", c_p), ("silhouette_scratch", c_p),
        ("gmm_hi", c_p), ("gmm_lo", c_p), ("unet_hi", c_p), ("unet_lo", c_p), ("cloth_hi", c_p), ("cloth_lo", c_p),
        ("gmm_cpad", c_i), ("unet_cpad", c_i), ("cloth_cpad", c_i), ("plane_fmt", c_i),
        ("tab_bounds", c_p * 4), ("tab_kk", c_p * 4), ("tab_ksize", c_i * 4),
        ("F", c_i), ("H", c_i), ("W", c_i), ("n_joints", c_i),
    ]


# name -> argtypes (every function returns int status unless listed in _RESTYPES)
SIGNATURES = {
    "shineon_version": [],
    "shineon_last_error": [],
    "shineon_launch_count": [],
    "shineon_tps_grid_fwd": [c_p, C.POINTER(TpsTables), c_p, c_i, c_i, c_i, c_p],
    "shineon_grid_sample_fwd": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_tps_grid_sample_fwd": [c_p, C.POINTER(TpsTables), c_i, c_i, c_i,
                                    c_p, c_i, c_i, c_p, c_p, c_i, c_i, c_p, c_p, c_i, c_i, c_p, c_p, c_p],
    "shineon_resample2d_fwd": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_resample2d_bwd": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_channelnorm_fwd": [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_channelnorm_bwd": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_correlation_out_shape": [c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i,
                                      C.POINTER(c_i), C.POINTER(c_i), C.POINTER(c_i)],
    "shineon_correlation_fwd": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_correlation_gather": [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_correlation_gather_planes": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_p],
    "shineon_correlation_bwd": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_pack_conv_weight": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_f, c_p],
    "shineon_pack_deconv4x4s2_weight": [c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_i, c_f, c_p],
    "shineon_conv2d_igemm_fwd": [C.POINTER(Conv2dParams), c_p],
    "shineon_conv2d_splitk_workspace_bytes": [C.POINTER(Conv2dParams)],
    "shineon_conv2d_im2col_fwd": [C.POINTER(Conv2dParams), c_p, c_i, c_p, c_i, c_p],
    "shineon_conv2d_direct_fwd": [C.POINTER(Conv2dParams), c_p],
    "shineon_nchw_to_planes": [c_p, c_i, c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_p],
    "shineon_planes_to_nchw": [c_p, c_p, c_i, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_nchw_im2col_planes": [c_p, c_i, c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i,
                                   c_f, c_i, c_p],
    "shineon_nchw_s2d_planes": [c_p, c_i, c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_col2im3x3": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_upconv3x3_gather": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_pil_bilinear_coeffs": [c_i, c_i, c_p, c_p],
    "shineon_frame_prep": [C.POINTER(FramePrepParams), c_p],
    "shineon_frame_prep_planes": [C.POINTER(FramePrepPlanesParams), c_p],
    "shineon_tps_warp_u8_planes": [c_p, C.POINTER(TpsTables), c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_flo_decode": [c_p, c_p, c_i, c_i, c_p],
    "shineon_instnorm_act": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_i, c_i, c_f, c_i, c_p],
    "shineon_upsample2x_cat": [c_p, c_p, c_i, c_p, c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_i, c_p],
    "shineon_sagan_attention": [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_p],
    "shineon_l2norm_correlation": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_l2norm_planes": [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_p],
    "shineon_feature_l2norm": [c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "shineon_linear_tanh": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_flownet_normalize": [c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_p],
    "shineon_upsample4x_flow": [c_p, c_i, c_p, c_i, c_i, c_i, c_f, c_i, c_p],
    "shineon_flow_deconv4x4s2_planes": [c_p, c_i, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_flownet_warp_concat": [c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_p],
    "shineon_flownet_fusion_concat": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    "shineon_bilinear_resize": [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_p],
    "shineon_flow_confidence": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_p],
    "shineon_adam_step": [c_p, c_p, c_p, c_p, C.c_long, c_f, c_f, c_f, c_f, c_f, c_i, c_f, c_p],
    "shineon_conv2d_wgrad_workspace_bytes": [C.POINTER(Conv2dWgradParams)],
    "shineon_conv2d_wgrad": [C.POINTER(Conv2dWgradParams), c_p],
    "shineon_channel_sum": [c_p, c_p, c_p, C.c_long, c_i, c_i, c_f, c_f, c_p],
    "shineon_instnorm_act_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_i, c_f, c_i, c_p],
    "shineon_act_bwd": [c_p, c_p, c_p, c_p, C.c_long, c_i, c_f, c_p],
    "shineon_upsample2x_cat_bwd": [c_p, c_i, c_p, c_i, c_p, c_i, c_i, c_i, c_i, c_p],
    "shineon_sagan_attention_bwd_workspace_bytes": [c_i, c_i],
    "shineon_sagan_attention_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, C.c_size_t, c_i, c_i, c_i, c_i, c_f, c_p],
    "shineon_tom_compose_bwd": [c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_l1_loss": [c_p, c_p, c_p, c_p, c_p, C.c_long, c_f, c_f, c_i, c_p],
    "shineon_maxpool2x2_fwd": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_nhwc_to_nchw_add": [c_p, c_i, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_maxpool2x2_bwd": [c_p, c_p, c_i, c_p, c_i, c_i, c_i, c_i, c_p],
    "shineon_tom_compose": [c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "shineon_image_to_u8": [c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "shineon_chan_stats": [c_p, c_p, c_i, c_i, c_i, c_p],
    "shineon_spade_modulate": [c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_i, c_f,
                               c_i, c_p],
    "shineon_nearest_resize_nhwc": [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_p],
    "shineon_nearest_resize_planes": [c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_i, c_i, c_i, c_f, c_f, c_i, c_p],
    "shineon_nearest_im2col_planes": [c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_f, c_i, c_p],
    "shineon_add_nhwc": [c_p, c_p, c_p, C.c_long, c_p],
    "shineon_sams_flow_blend": [c_p, c_i, c_p, c_p, C.c_long, c_i, c_i, c_i, c_p],
}
_RESTYPES = {"shineon_last_error": C.c_char_p, "shineon_launch_count": C.c_uint64,
             "shineon_conv2d_splitk_workspace_bytes": C.c_size_t,
             "shineon_conv2d_wgrad_workspace_bytes": C.c_size_t,
             "shineon_sagan_attention_bwd_workspace_bytes": C.c_size_t}

_lib = None


class ShineonError(RuntimeError):
    pass


def load():
    """dlopen libshineon_b200.so (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ShineonError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "There is no CPU/PyTorch fallback for the shineon ops.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == symbol missing from the build
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_i)
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().shineon_last_error()
        raise ShineonError(f"{what} failed with status {status}: {msg.decode() if msg else ''}")


def launch_count():
    return int(load().shineon_launch_count())
