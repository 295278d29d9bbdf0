"""ctypes binding of libcpd_b200.so (the C ABI declared in include/cpd_b200.h).

There is NO fallback: if the CUDA library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CPD_B200_LIB", os.path.join(_HERE, "libcpd_b200.so"))  # the override is for A/B measurements of two builds

CPD_F32, CPD_F16, CPD_BF16 = 0, 1, 2
CPD_EULER, CPD_EULER_ANCESTRAL, CPD_DPMPP_2M, CPD_DENOISE_ONLY, CPD_HEUN2, CPD_LMS = 0, 1, 2, 3, 4, 5
CPD_THRESH_DYNAMIC, CPD_THRESH_STATIC = 0, 1
(CPD_THRESH_DYNANORMIC, CPD_THRESH_SCALED_DYNAMIC_PERC, CPD_THRESH_RENORM, CPD_THRESH_SCALED_NORM, CPD_THRESH_SPATIAL_NORM,
 CPD_THRESH_SCALED_SPATIAL_NORM) = 2, 3, 4, 5, 6, 7
CPD_PRED_EPSILON, CPD_PRED_VELOCITY = 0, 1
CPD_EPI_NONE, CPD_EPI_GEGLU = 0, 1
CPD_MAX_SUBPROMPTS = 16

DTYPE_CODE = {torch.float32: CPD_F32, torch.float16: CPD_F16, torch.bfloat16: CPD_BF16}


class StepParams(C.Structure):
    _fields_ = [
        ("eps", C.c_void_p), ("eps_dtype", C.c_int), ("eps_image_stride", C.c_int64), ("eps_row_stride", C.c_int64),
        ("x", C.c_void_p), ("old_denoised", C.c_void_p), ("noise", C.c_void_p), ("denoised_out", C.c_void_p),
        ("eps_out", C.c_void_p),
        ("n_images", C.c_int), ("n_sub", C.c_int), ("hw", C.c_int),
        ("weights", C.c_float * CPD_MAX_SUBPROMPTS), ("mask_scalar", C.c_float * CPD_MAX_SUBPROMPTS),
        ("masks", C.c_void_p * CPD_MAX_SUBPROMPTS),
        ("guidance", C.c_float), ("sampler", C.c_int), ("pred_type", C.c_int),
        ("sigma_hat", C.c_float), ("v_c_eps", C.c_float), ("v_c_x_div", C.c_float), ("dt", C.c_float),
        ("sigma_up", C.c_float), ("dpm_ratio", C.c_float), ("dpm_expm1", C.c_float), ("dpm_c1", C.c_float),
        ("dpm_c2", C.c_float), ("dpm_first", C.c_int), ("write_old", C.c_int),
        ("x_base", C.c_void_p), ("x_out", C.c_void_p), ("d_out", C.c_void_p), ("d_prev", C.c_void_p * 3),
        ("lms_coeff", C.c_float * 4), ("lms_order", C.c_int), ("noise_mul", C.c_float),
        ("clip_scaled", C.c_void_p), ("scaled_out", C.c_void_p), ("scaled_in", C.c_void_p),
        ("dyn", C.c_void_p),
    ]


class StepScalars(C.Structure):  # cpd_step_scalars: one row of the per-schedule device table (64 bytes)
    _fields_ = [
        ("c_in", C.c_float), ("t", C.c_float), ("guidance", C.c_float), ("sigma_hat", C.c_float), ("v_c_eps", C.c_float),
        ("v_c_x_div", C.c_float), ("dt", C.c_float), ("sigma_up", C.c_float), ("dpm_ratio", C.c_float), ("dpm_expm1", C.c_float),
        ("dpm_c1", C.c_float), ("dpm_c2", C.c_float), ("dpm_first", C.c_int), ("write_old", C.c_int), ("noise_mul", C.c_float),
        ("reserved", C.c_int),
    ]


class GemmParams(C.Structure):
    _fields_ = [
        ("a0", C.c_void_p), ("a1", C.c_void_p), ("c0", C.c_int), ("c1", C.c_int),
        ("n_img", C.c_int), ("h_in", C.c_int), ("w_in", C.c_int), ("ksize", C.c_int), ("stride", C.c_int),
        ("wt", C.c_void_p), ("n_out", C.c_int), ("bias", C.c_void_p),
        ("rowvec", C.c_void_p), ("rowvec_stride", C.c_int),
        ("residual", C.c_void_p), ("ld_res", C.c_int),
        ("d", C.c_void_p), ("ldd", C.c_int), ("epilogue", C.c_int), ("variant", C.c_int), ("m_valid", C.c_int),
        ("a_fp16", C.c_int), ("b_fp16", C.c_int), ("out_fp16", C.c_int), ("geglu_block", C.c_int),
        ("splitk_ws", C.c_void_p), ("splitk_ws_floats", C.c_int64),
        ("tune_scratch", C.c_void_p), ("tune_scratch_bytes", C.c_int64),
        ("ln_sums_out", C.c_void_p), ("ln_parts_out", C.POINTER(C.c_int)), ("ln_sums", C.c_void_p), ("ln_parts", C.c_int),
        ("ln_ld", C.c_int64), ("ln_g", C.c_void_p), ("ln_c", C.c_int), ("ln_eps", C.c_float),
        ("d_t", C.c_void_p), ("dt_col0", C.c_int), ("ldd_t", C.c_int64), ("gn_sums_out", C.c_void_p),
    ]


CPD_UNET_MAX_LEVELS = 8


class UNetConfig(C.Structure):  # cpd_unet_config
    _fields_ = [
        ("in_channels", C.c_int), ("out_channels", C.c_int), ("model_channels", C.c_int), ("num_res_blocks", C.c_int),
        ("n_levels", C.c_int), ("channel_mult", C.c_int * CPD_UNET_MAX_LEVELS),
        ("n_attention_resolutions", C.c_int), ("attention_resolutions", C.c_int * CPD_UNET_MAX_LEVELS),
        ("num_heads", C.c_int), ("num_head_channels", C.c_int), ("transformer_depth", C.c_int * CPD_UNET_MAX_LEVELS),
        ("context_dim", C.c_int), ("use_linear_in_transformer", C.c_int), ("adm_in_channels", C.c_int),
        ("act_fp16", C.c_int), ("eps_dtype", C.c_int), ("use_cuda_graph", C.c_int),
    ]


class UNetIO(C.Structure):  # cpd_unet_io
    _fields_ = [
        ("x", C.c_void_p), ("n_images", C.c_int), ("h", C.c_int), ("w", C.c_int), ("rows_per_image", C.c_int),
        ("c_in", C.c_void_p), ("t", C.c_void_p), ("t_count", C.c_int), ("eps", C.c_void_p),
        ("inject_skips", C.POINTER(C.c_void_p)), ("inject_feats", C.POINTER(C.c_void_p)), ("no_graph", C.c_int),
    ]


class AttnParams(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("ldq", C.c_int), ("k", C.c_void_p), ("ldk", C.c_int), ("vt", C.c_void_p), ("ldvt", C.c_int),
        ("o", C.c_void_p), ("ldo", C.c_int),
        ("batch", C.c_int), ("heads", C.c_int), ("nq", C.c_int), ("nk", C.c_int), ("nk_pad", C.c_int), ("dpad", C.c_int),
        ("scale", C.c_float), ("kv_batch", C.c_int), ("act_fp16", C.c_int), ("d_head", C.c_int),
    ]


_SIGS = {
    "cpd_last_error": (C.c_char_p, []),
    "cpd_abi_version": (C.c_int, []),
    "cpd_sampler_step": (C.c_int, [C.POINTER(StepParams), C.c_void_p]),
    "cpd_step_select": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cpd_add_noise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int64, C.c_void_p]),
    "cpd_threshold": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]),
    "cpd_threshold_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "cpd_gemm_conv": (C.c_int, [C.POINTER(GemmParams), C.c_void_p]),
    "cpd_gemm_set_autotune": (None, [C.c_int]),
    "cpd_gemm_tune_clear": (None, []),
    "cpd_groupnorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cpd_groupnorm_launches": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "cpd_groupnorm_apply": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "cpd_layernorm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p,
                                C.c_void_p]),
    "cpd_timestep_embedding": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cpd_small_linear": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_void_p]),
    "cpd_conv_in": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                              C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cpd_conv_out": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                               C.c_int, C.c_int, C.c_void_p]),
    "cpd_upsample2x": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cpd_attention": (C.c_int, [C.POINTER(AttnParams), C.c_void_p]),
    "cpd_softmax_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_float, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "cpd_pointwise_small": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                                      C.c_void_p]),
    "cpd_gaussian_blur": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                    C.c_void_p]),
    "cpd_channel_mean": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cpd_percentile": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "cpd_attn_guide": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "cpd_images_to_uint8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "cpd_gemm_tune_export": (C.c_int64, [C.c_char_p, C.c_int64]),
    "cpd_gemm_tune_import": (C.c_int, [C.c_char_p]),
    # plan-level UNet
    "cpd_unet_plan_create": (C.c_int, [C.POINTER(UNetConfig), C.POINTER(C.c_void_p)]),
    "cpd_unet_plan_destroy": (None, [C.c_void_p]),
    "cpd_pack_weights": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int64, C.c_int]),
    "cpd_unet_plan_missing_weights": (C.c_int, [C.c_void_p]),
    "cpd_cache_context_kv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cpd_unet_set_vector": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "cpd_unet_forward": (C.c_int, [C.c_void_p, C.POINTER(UNetIO), C.c_void_p]),
    "cpd_unet_plan_buffer": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "cpd_unet_plan_launches": (C.c_int64, [C.c_void_p]),
    "cpd_unet_plan_tap": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.POINTER(C.c_int)]),
    "cpd_unet_plan_blocks": (C.c_int, [C.c_void_p, C.c_int]),
    "cpd_unet_plan_set_profile": (None, [C.c_void_p, C.c_int]),
    "cpd_unet_plan_profile_dump": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_int64]),
    "cpd_debug_attention_cross_timeline": (None, [C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)
_lib = None


def load():
    """Load libcpd_b200.so; raises RuntimeError (no fallback) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the hot path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status, what):
    if status != 0:
        msg = load().cpd_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {status}): {msg}")


def stream_ptr(device=None):
    """The current torch stream of `device` (default: the current device).  Callers that own a device (UNetModel, Denoiser)
    wrap their launches in `torch.cuda.device(...)`, so the current device IS the tensors' device."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class DeviceView:
    """A raw device pointer owned by the C library (a plan buffer) as a zero-copy torch tensor, through
    __cuda_array_interface__.  16-bit buffers are exposed as int16 and re-viewed in the activation dtype."""

    def __init__(self, ptr, numel, itemsize):
        typestr = {2: "<i2", 4: "<f4", 8: "<f8"}[itemsize]
        self.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_tensor(ptr, numel, dtype, device):
    with torch.cuda.device(device):
        t = torch.as_tensor(DeviceView(ptr, numel, torch.empty(0, dtype=dtype).element_size()), device=device)
    return t.view(dtype) if t.dtype != dtype else t


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
