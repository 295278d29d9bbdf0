"""Thin torch-tensor wrappers over the C ABI (include/cpd_b200.h).  Device memory and the current CUDA
stream come from torch; all arithmetic happens in libcpd_b200.so.  No fallbacks.

Activation tensors are 16-bit: torch.float16 or torch.bfloat16 (the `act_fp16` flag of the C ABI is derived
from the tensor dtype); weights are always bf16."""
import ctypes as C
import os

import torch

from ._lib import (AttnParams, GemmParams, StepParams, check, load, ptr, stream_ptr, CPD_BF16, CPD_F32, DTYPE_CODE,  # noqa: F401
                   CPD_EPI_NONE, CPD_EPI_GEGLU, CPD_MAX_SUBPROMPTS)

LAUNCHES = 0    # kernels launched through this module (bench.py reports it as gpu_launches)
PROFILE = None  # when a list: (kind, start_event, end_event, flops) per tensor-core launch (bench.py roofline leg)

_ACT = (torch.float16, torch.bfloat16)


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


class _Prof:
    def __init__(self, kind, flops, label=""):
        self.kind, self.flops, self.label = kind, flops, label

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.append((self.kind, self.e0, self.e1, self.flops, self.label))


def _req(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (the hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def _act(t, name, like=None):
    """Validate a 16-bit activation tensor; returns 1 for fp16, 0 for bf16."""
    if t is None:
        return None
    if not t.is_cuda or t.dtype not in _ACT:
        raise RuntimeError(f"{name} must be a CUDA fp16/bf16 tensor, got {t.dtype} on {t.device}")
    if like is not None and t.dtype != like:
        raise RuntimeError(f"{name} must be {like}, got {t.dtype}")
    return int(t.dtype == torch.float16)


def sampler_step(eps, x, *, n_sub, weights, mask_scalars, masks, guidance, sampler, pred_type, sigma_hat, v_c_eps=0.0,
                 v_c_x_div=1.0, dt=0.0, sigma_up=0.0, dpm_ratio=0.0, dpm_expm1=0.0, dpm_c1=0.0, dpm_c2=0.0, dpm_first=1,
                 write_old=0, old_denoised=None, noise=None, denoised_out=None, eps_out=None, x_base=None, x_out=None, d_out=None,
                 d_prev=(), lms_coeff=(), noise_mul=1.0, clip_scaled=None, scaled_out=None, scaled_in=None, dyn=None):
    """eps: [n_images * (1 + n_sub), 4, h, w] (image-major rows); x: [n_images, 4, h, w] fp32, the UNet input; the updated
    sample goes to x_out (default: x, in place) and starts from x_base (default: x)."""
    _req(x, torch.float32, "x")
    for name, t in (("x_base", x_base), ("x_out", x_out), ("d_out", d_out), ("clip_scaled", clip_scaled), ("scaled_out", scaled_out),
                    ("scaled_in", scaled_in)):
        _req(t, torch.float32, name)
    _req(old_denoised, torch.float32, "old_denoised")
    _req(noise, torch.float32, "noise")
    _req(denoised_out, torch.float32, "denoised_out")
    _req(eps_out, torch.float32, "eps_out")
    if not eps.is_cuda or not eps.is_contiguous() or eps.dtype not in DTYPE_CODE:
        raise RuntimeError("eps must be a contiguous CUDA tensor of dtype fp32/fp16/bf16")
    n_images = x.shape[0]
    if n_images == 0:  # empty batch: nothing to launch
        return x
    L = x[0].numel()
    R = 1 + n_sub
    if eps.numel() != n_images * R * L:
        raise RuntimeError(f"eps has {eps.numel()} elements, expected {n_images}*{R}*{L}")
    if n_sub < 0 or n_sub > CPD_MAX_SUBPROMPTS:  # 0: eps already is e_t, one row per image
        raise RuntimeError(f"n_sub={n_sub} out of range")
    p = StepParams()
    p.eps = eps.data_ptr()
    p.eps_dtype = DTYPE_CODE[eps.dtype]
    p.eps_image_stride = R * L
    p.eps_row_stride = L
    p.x = x.data_ptr()
    p.old_denoised = old_denoised.data_ptr() if old_denoised is not None else None
    p.noise = noise.data_ptr() if noise is not None else None
    p.denoised_out = denoised_out.data_ptr() if denoised_out is not None else None
    p.eps_out = eps_out.data_ptr() if eps_out is not None else None
    p.n_images, p.n_sub, p.hw = n_images, n_sub, L // 4
    for k in range(n_sub):
        p.weights[k] = float(weights[k])
        p.mask_scalar[k] = float(mask_scalars[k])
        m = masks[k] if masks is not None else None
        if m is not None:
            _req(m, torch.float32, f"masks[{k}]")
            if m.numel() != L // 4:
                raise RuntimeError(f"masks[{k}] must have {L // 4} elements")
            p.masks[k] = m.data_ptr()
        else:
            p.masks[k] = None
    p.guidance = float(guidance)
    p.sampler, p.pred_type = int(sampler), int(pred_type)
    p.sigma_hat, p.v_c_eps, p.v_c_x_div, p.dt, p.sigma_up = float(sigma_hat), float(v_c_eps), float(v_c_x_div), float(dt), float(sigma_up)
    p.dpm_ratio, p.dpm_expm1, p.dpm_c1, p.dpm_c2 = float(dpm_ratio), float(dpm_expm1), float(dpm_c1), float(dpm_c2)
    p.dpm_first, p.write_old = int(dpm_first), int(write_old)
    p.x_base = x_base.data_ptr() if x_base is not None else None
    p.x_out = x_out.data_ptr() if x_out is not None else None
    p.d_out = d_out.data_ptr() if d_out is not None else None
    for k in range(3):
        if k < len(d_prev) and d_prev[k] is not None:
            _req(d_prev[k], torch.float32, f"d_prev[{k}]")
            p.d_prev[k] = d_prev[k].data_ptr()
        else:
            p.d_prev[k] = None
    for k in range(4):
        p.lms_coeff[k] = float(lms_coeff[k]) if k < len(lms_coeff) else 0.0
    p.lms_order = len(lms_coeff)
    p.noise_mul = float(noise_mul)
    p.clip_scaled = clip_scaled.data_ptr() if clip_scaled is not None else None
    p.scaled_out = scaled_out.data_ptr() if scaled_out is not None else None
    p.scaled_in = scaled_in.data_ptr() if scaled_in is not None else None
    p.dyn = dyn.data_ptr() if dyn is not None else None  # device row of per-step scalars (cpd_step_select) overriding the by-value ones
    with _Prof("sampler_step", 0.0):
        check(load().cpd_sampler_step(C.byref(p), stream_ptr()), "cpd_sampler_step")
    _count()
    return x


def add_noise(x, noise, *, noise_mul, scale):
    """x += (noise * noise_mul) * scale in place (stochastic churn before the denoiser call), separately rounded fp32 ops."""
    _req(x, torch.float32, "x")
    _req(noise, torch.float32, "noise")
    if noise.numel() != x.numel():
        raise RuntimeError("noise must have the shape of x")
    with _Prof("small", 0.0):
        check(load().cpd_add_noise(ptr(x), ptr(noise), float(noise_mul), float(scale), int(x.numel()), stream_ptr()), "cpd_add_noise")
    _count()
    return x


def threshold(x, bound, *, alg, threshold, clamp_inplace=True):
    """Thresholding extension on the device (cpd_threshold): bound[b] = max(percentile_q(|x_b|), 1) or the static bound;
    optionally x <- half(clamp(x, -bound[b], bound[b])) in place.  x: [n_images, ...] fp32, bound: [n_images] fp32."""
    _req(x, torch.float32, "x")
    _req(bound, torch.float32, "bound")
    n = x.shape[0]
    if bound.numel() < n:
        raise RuntimeError("bound must hold one fp32 value per image")
    L = x[0].numel() if n else 4
    with _Prof("threshold", 0.0):
        check(load().cpd_threshold(ptr(x), n, L, int(alg), float(threshold), int(bool(clamp_inplace)), ptr(bound), stream_ptr()),
              "cpd_threshold")
    _count(2 if clamp_inplace else 1)
    return bound


def threshold_ex(x, bound, *, alg, threshold):
    """The non-clamp thresholding extensions in place on x: [n_images, C, h, w] fp32 (cpd_threshold_ex; threshold.py:87-286):
    min-max rescaling, torch.quantile / np.percentile bounds, RMS and per-pixel channel-RMS rescaling, all per image on the
    device; the result holds fp16-rounded values.  bound: [n_images] fp32, receives the per-image s."""
    _req(x, torch.float32, "x")
    _req(bound, torch.float32, "bound")
    if x.ndim != 4:
        raise RuntimeError(f"threshold_ex wants [n_images, C, h, w], got {tuple(x.shape)}")
    n = x.shape[0]
    if bound.numel() < n:
        raise RuntimeError("bound must hold one fp32 value per image")
    with _Prof("threshold", 0.0):
        check(load().cpd_threshold_ex(ptr(x), n, x.shape[1], x.shape[2] * x.shape[3], int(alg), float(threshold), ptr(bound),
                                      stream_ptr()), "cpd_threshold_ex")
    _count()
    return x


# Tile-variant selection of cpd_gemm_conv lives in the library (csrc/gemm_tune.cu): with variant 0 an unseen problem shape is timed on
# every applicable variant the first time it is launched outside stream capture (CPD_GEMM_AUTOTUNE=0 switches that off) - this
# module only lends the scratch output and the split-K workspace.
_SPLITK_WS = {}      # device index -> fp32 scratch for the split-K partial tiles
_SPLITK_FLOATS = 32 * 1024 * 1024
_TUNE_SCRATCH = {}   # device index -> byte scratch at least as large as the largest GEMM output seen


def _splitk_ws(device):
    ws = _SPLITK_WS.get(device.index)
    if ws is None:
        ws = _SPLITK_WS[device.index] = torch.empty(_SPLITK_FLOATS, dtype=torch.float32, device=device)
    return ws


def _tune_scratch(device, nbytes):
    t = _TUNE_SCRATCH.get(device.index)
    if t is None or t.numel() < nbytes:
        if torch.cuda.is_current_stream_capturing():
            return t
        t = _TUNE_SCRATCH[device.index] = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return t


def gemm_tune_table():
    """The library's per-shape variant table as text (cpd_gemm_tune_export): broadcast it from rank 0 and `gemm_tune_import` it on
    the other ranks when every process must pick identical tile variants (bit-identical results across ranks)."""
    lib = load()
    n = lib.cpd_gemm_tune_export(None, 0)
    buf = C.create_string_buffer(int(n))
    lib.cpd_gemm_tune_export(buf, n)
    return buf.value.decode()


def gemm_tune_import(text):
    n = load().cpd_gemm_tune_import(text.encode())
    if n < 0:
        raise RuntimeError("cpd_gemm_tune_import: malformed table")
    return n


def gemm_conv(a0, wt, out, *, n_img, h, w, c0, n_out, a1=None, c1=0, ksize=1, stride=1, bias=None, rowvec=None,
              rowvec_stride=0, residual=None, ld_res=0, ldd=None, epilogue=CPD_EPI_NONE, variant=0, m_valid=0, geglu_block=128,
              ln_sums_out=None, ln_sums=None, ln_parts=0, ln_g=None, ln_eps=1e-5, d_t=None, dt_col0=0, gn_sums_out=None):
    """a0/a1 and wt are 16-bit (fp16 or bf16, independently); out/residual share one 16-bit dtype.
    Folded LayerNorm (cpd_gemm_params.ln_*): a producer passes ln_sums_out = fp32 [max_parts, rows, 2] and gets (out, number of
    parts written) back; a consumer passes ln_sums, ln_parts,
    ln_g (and wt = gamma . W, bias = W beta + b).  d_t [n_out - dt_col0, rows]: columns from dt_col0 on are stored transposed.
    gn_sums_out: int64 [n_img, n_out, 2], zeroed by the caller: fixed-point GroupNorm statistics of the output (groupnorm_apply)."""
    a_f16 = _act(a0, "a0")
    _act(a1, "a1", like=a0.dtype)
    b_f16 = _act(wt, "wt")
    o_f16 = _act(out, "out")
    _act(residual, "residual", like=out.dtype)
    _req(bias, torch.float32, "bias")
    _req(rowvec, torch.float32, "rowvec")
    p = GemmParams()
    p.a0, p.a1, p.c0, p.c1 = a0.data_ptr(), (a1.data_ptr() if a1 is not None else None), c0, c1
    p.n_img, p.h_in, p.w_in, p.ksize, p.stride = n_img, h, w, ksize, stride
    p.wt, p.n_out = wt.data_ptr(), n_out
    p.bias = bias.data_ptr() if bias is not None else None
    p.rowvec = rowvec.data_ptr() if rowvec is not None else None
    p.rowvec_stride = rowvec_stride
    p.residual = residual.data_ptr() if residual is not None else None
    p.ld_res = ld_res
    p.d = out.data_ptr()
    p.ldd = ldd if ldd is not None else (n_out // 2 if epilogue == CPD_EPI_GEGLU else n_out)
    p.epilogue, p.variant, p.m_valid = epilogue, variant, m_valid
    p.a_fp16, p.b_fp16, p.out_fp16 = a_f16, b_f16, o_f16
    ws = _splitk_ws(out.device)
    p.splitk_ws, p.splitk_ws_floats = ws.data_ptr(), ws.numel()
    p.geglu_block = geglu_block if epilogue == CPD_EPI_GEGLU else 0
    if variant == 0:
        sc = _tune_scratch(out.device, out.numel() * out.element_size())
        if sc is not None:
            p.tune_scratch, p.tune_scratch_bytes = sc.data_ptr(), sc.numel()
    parts_out = C.c_int(0)
    if ln_sums_out is not None:
        _req(ln_sums_out, torch.float32, "ln_sums_out")
        p.ln_sums_out, p.ln_parts_out, p.ln_ld = ln_sums_out.data_ptr(), C.pointer(parts_out), ln_sums_out.shape[1]
    if ln_sums is not None:
        _req(ln_sums, torch.float32, "ln_sums")
        _req(ln_g, torch.float32, "ln_g")
        p.ln_sums, p.ln_parts, p.ln_ld, p.ln_g = ln_sums.data_ptr(), ln_parts, ln_sums.shape[1], ln_g.data_ptr()
        p.ln_c, p.ln_eps = ksize * ksize * (c0 + c1), ln_eps
    if gn_sums_out is not None:
        _req(gn_sums_out, torch.int64, "gn_sums_out")
        p.gn_sums_out = gn_sums_out.data_ptr()
    if d_t is not None:
        _act(d_t, "d_t", like=out.dtype)
        p.d_t, p.dt_col0, p.ldd_t = d_t.data_ptr(), dt_col0, d_t.shape[1]
    flops = 2.0 * n_img * (h // stride) * (w // stride) * n_out * ksize * ksize * (c0 + c1)
    label = f"M={n_img * (h // stride) * (w // stride)} N={n_out} K={ksize * ksize * (c0 + c1)}" + (" geglu" if epilogue else "")
    with _Prof("gemm_conv", flops, label):
        check(load().cpd_gemm_conv(C.byref(p), stream_ptr()), "cpd_gemm_conv")
    _count()
    if ln_sums_out is not None:
        return out, parts_out.value
    return out


def groupnorm(a0, gamma, beta, out, stats, *, n_img, hw, c0, a1=None, c1=0, eps=1e-5, silu=True):
    f16 = _act(a0, "a0")
    _act(a1, "a1", like=a0.dtype)
    _act(out, "out", like=a0.dtype)
    _req(gamma, torch.float32, "gamma")
    _req(beta, torch.float32, "beta")
    _req(stats, torch.float64, "stats")
    if stats.numel() < n_img * 64 * GN_MAX_CHUNKS:
        raise RuntimeError(f"stats scratch too small: need n_img * 64 * {GN_MAX_CHUNKS} doubles")
    with _Prof("groupnorm", 0.0, f"n={n_img} hw={hw} C={c0 + c1}"):
        check(load().cpd_groupnorm(ptr(a0), ptr(a1), c0, c1, n_img, hw, ptr(gamma), ptr(beta), float(eps), int(silu), f16, ptr(stats),
                                   ptr(out), stream_ptr()), "cpd_groupnorm")
    _count(load().cpd_groupnorm_launches(c0 + c1, n_img, hw))  # 1: shared-memory slab kernel, 2: statistics + apply
    return out


def groupnorm_apply(x, gamma, beta, out, chan_sums, *, n_img, hw, c, eps=1e-5, silu=True):
    """GroupNorm(32) from the fixed-point statistics a producing GEMM emitted (gemm_conv(gn_sums_out=...))."""
    f16 = _act(x, "x")
    _act(out, "out", like=x.dtype)
    _req(gamma, torch.float32, "gamma")
    _req(beta, torch.float32, "beta")
    _req(chan_sums, torch.int64, "chan_sums")
    with _Prof("groupnorm", 0.0, f"apply n={n_img} hw={hw} C={c}"):
        check(load().cpd_groupnorm_apply(ptr(x), c, n_img, hw, ptr(gamma), ptr(beta), float(eps), int(silu), f16, ptr(chan_sums), ptr(out),
                                         stream_ptr()), "cpd_groupnorm_apply")
    _count()
    return out


GN_MAX_CHUNKS = 64  # CPD_GN_MAX_CHUNKS of include/cpd_b200.h


def layernorm(x, gamma, beta, out, *, rows, c, eps=1e-5):
    f16 = _act(x, "x")
    _act(out, "out", like=x.dtype)
    _req(gamma, torch.float32, "gamma")
    _req(beta, torch.float32, "beta")
    with _Prof("layernorm", 0.0, f"rows={rows} C={c}"):
        check(load().cpd_layernorm(ptr(x), rows, c, ptr(gamma), ptr(beta), float(eps), f16, ptr(out), stream_ptr()), "cpd_layernorm")
    _count()
    return out


def timestep_embedding(t, out, *, dim, round_t_bf16=True):
    _req(t, torch.float32, "t")
    _req(out, torch.bfloat16, "out")
    with _Prof("small", 0.0):
        check(load().cpd_timestep_embedding(ptr(t), t.numel(), dim, int(round_t_bf16), ptr(out), stream_ptr()), "cpd_timestep_embedding")
    _count()
    return out


def small_linear(x, w, b, *, m, k, n, silu_in=False, out_f32=None, out_bf16=None, ld_out=None):
    _req(x, torch.bfloat16, "x")
    _req(w, torch.bfloat16, "w")
    _req(b, torch.float32, "b")
    _req(out_f32, torch.float32, "out_f32")
    _req(out_bf16, torch.bfloat16, "out_bf16")
    with _Prof("small", 0.0):
        check(load().cpd_small_linear(ptr(x), m, k, ptr(w), ptr(b), n, int(silu_in), ptr(out_f32), ptr(out_bf16),
                                      ld_out if ld_out is not None else n, stream_ptr()), "cpd_small_linear")
    _count()


def conv_in(x, wt, bias, out, *, n, cin, h, w, cout, scale=1.0, rows_per_image=1, scale_dev=None):
    _req(x, torch.float32, "x")
    _req(scale_dev, torch.float32, "scale_dev")
    _req(wt, torch.bfloat16, "wt")
    _req(bias, torch.float32, "bias")
    f16 = _act(out, "out")
    with _Prof("conv_in", 0.0):
        check(load().cpd_conv_in(ptr(x), n, cin, h, w, ptr(wt), ptr(bias), cout, float(scale), ptr(scale_dev), int(rows_per_image), f16, ptr(out),
                                 stream_ptr()), "cpd_conv_in")
    _count()
    return out


def conv_out(a, wt, bias, out, *, n, h, w, cin, cout):
    f16 = _act(a, "a")
    _req(wt, torch.bfloat16, "wt")
    _req(bias, torch.float32, "bias")
    if out.dtype not in (torch.bfloat16, torch.float32) or not out.is_cuda:
        raise RuntimeError("out must be a CUDA bf16/fp32 tensor")
    with _Prof("conv_out", 0.0):
        check(load().cpd_conv_out(ptr(a), n, h, w, cin, ptr(wt), ptr(bias), cout, ptr(out), DTYPE_CODE[out.dtype], f16, stream_ptr()),
              "cpd_conv_out")
    _count()
    return out


def upsample2x(a, out, *, n, h, w, c):
    _act(a, "a")
    _act(out, "out", like=a.dtype)
    with _Prof("upsample", 0.0):
        check(load().cpd_upsample2x(ptr(a), n, h, w, c, ptr(out), stream_ptr()), "cpd_upsample2x")
    _count()
    return out


def attention(q, k, vt, o, *, ldq, ldk, ldvt, ldo, batch, heads, nq, nk, nk_pad, dpad, scale, kv_batch=0, d_head=0):
    f16 = _act(q, "q")
    for t, nme in ((k, "k"), (vt, "vt"), (o, "o")):
        _act(t, nme, like=q.dtype)
    p = AttnParams()
    p.q, p.ldq, p.k, p.ldk, p.vt, p.ldvt, p.o, p.ldo = q.data_ptr(), ldq, k.data_ptr(), ldk, vt.data_ptr(), ldvt, o.data_ptr(), ldo
    p.batch, p.heads, p.nq, p.nk, p.nk_pad, p.dpad, p.scale = batch, heads, nq, nk, nk_pad, dpad, float(scale)
    p.kv_batch, p.act_fp16, p.d_head = kv_batch, f16, d_head
    d_alg = d_head if d_head > 0 else dpad  # algorithmic work counts the REAL head dim, not the padded one
    with _Prof("attention", 4.0 * batch * heads * nq * nk * d_alg, f"B={batch} H={heads} nq={nq} nk={nk} d={d_alg}"):
        check(load().cpd_attention(C.byref(p), stream_ptr()), "cpd_attention")
    _count()
    return o


def softmax_rows(x, out, *, rows, cols, scale, ld=None, ldo=None):
    """out[r][c] = softmax_c(scale * x[r][c]) over a 16-bit row-major matrix (VAE AttnBlock weights); out may alias x."""
    f16 = _act(x, "x")
    _act(out, "out", like=x.dtype)
    with _Prof("softmax_rows", 0.0, f"rows={rows} cols={cols}"):
        check(load().cpd_softmax_rows(ptr(x), rows, cols, int(ld if ld is not None else cols), float(scale), f16, ptr(out),
                                      int(ldo if ldo is not None else cols), stream_ptr()), "cpd_softmax_rows")
    _count()
    return out


def pointwise_small(x, w, b, out, *, n, cin, cout, hw, scale=1.0):
    """1x1 conv of an fp32 NCHW tensor with <= 8 channels (post_quant_conv), input multiplied by `scale` first."""
    for name, t in (("x", x), ("w", w), ("b", b), ("out", out)):
        _req(t, torch.float32, name)
    with _Prof("small", 0.0):
        check(load().cpd_pointwise_small(ptr(x), n, cin, cout, int(hw), ptr(w), ptr(b), float(scale), ptr(out), stream_ptr()),
              "cpd_pointwise_small")
    _count()
    return out


def images_to_uint8(x, out, *, ld_c=None):
    """out[n, h, w, c] = uint8(clamp((x[n, c, h, w] + 1) / 2, 0, 1) * 255) (prompts.py:472-475); x fp32 NCHW whose images are
    `ld_c` channels apart (the decoder's 4-channel buffer holding 3-channel images)."""
    if not x.is_cuda or x.dtype != torch.float32:
        raise RuntimeError("x must be a CUDA fp32 tensor")
    if out.dtype != torch.uint8 or not out.is_cuda or not out.is_contiguous():
        raise RuntimeError("out must be a contiguous CUDA uint8 tensor")
    n, c, h, w = x.shape
    ld_c = c if ld_c is None else ld_c
    if x.stride() != (ld_c * h * w, h * w, w, 1):
        raise RuntimeError("x must be NCHW with an image stride of ld_c * h * w")
    with _Prof("small", 0.0):
        check(load().cpd_images_to_uint8(ptr(x), n, c, h * w, ld_c, ptr(out), stream_ptr()), "cpd_images_to_uint8")
    _count()
    return out


# ---- guidance branches of the Denoiser (csrc/guidance.cu) ---------------------------------------------------------------
def gaussian_taps(kernel_size, sigma):
    """The 1-D taps of torchvision's _get_gaussian_kernel1d (fp32 torch ops on the host, as the reference evaluates them)."""
    half = (kernel_size - 1) * 0.5
    xs = torch.linspace(-half, half, steps=kernel_size, dtype=torch.float32)
    pdf = torch.exp(-0.5 * (xs / sigma).pow(2))
    return (pdf / pdf.sum()).contiguous()


def gaussian_blur(src, dst, *, kernel_size, sigma, planes_per_image=None, img_stride_src=None, img_stride_dst=None):
    """dst = GaussianBlur(kernel_size, sigma)(src) per plane, reflect padding.  src / dst: fp32 [n, planes, h, w]-shaped memory
    whose images may be strided (img_stride_* in elements; default: contiguous)."""
    if src.dtype != torch.float32 or dst.dtype != torch.float32 or not src.is_cuda or not dst.is_cuda:
        raise RuntimeError("gaussian_blur works on CUDA fp32 tensors")
    n, planes, h, w = src.shape
    taps = gaussian_taps(kernel_size, sigma)
    with _Prof("guidance", 0.0):
        check(load().cpd_gaussian_blur(ptr(src), ptr(dst), n, planes_per_image or planes, int(img_stride_src or planes * h * w),
                                       int(img_stride_dst or planes * h * w), h, w, C.c_void_p(taps.data_ptr()), kernel_size, stream_ptr()),
              "cpd_gaussian_blur")
    _count()
    return dst


def channel_mean(a, out, *, pixels, c):
    f16 = _act(a, "a")
    _req(out, torch.float32, "out")
    with _Prof("guidance", 0.0):
        check(load().cpd_channel_mean(ptr(a), int(pixels), int(c), f16, ptr(out), stream_ptr()), "cpd_channel_mean")
    _count()
    return out


def percentile(x, out, *, q):
    _req(x, torch.float32, "x")
    _req(out, torch.float32, "out")
    with _Prof("guidance", 0.0):
        check(load().cpd_percentile(ptr(x), int(x.numel()), float(q), ptr(out), stream_ptr()), "cpd_percentile")
    _count()
    return out


def attn_guide(stage, out, *, n_images, hw, x=None, eps_u=None, eps_stride=0, mask_mean=None, mask_img_stride=0, pct=None, blur=None,
               sum16=None, e_attn=None, sigma_hat=0.0, c_in=1.0, mode=2, scale=1.0, guidance=1.0):
    src = eps_u if eps_u is not None else e_attn
    with _Prof("guidance", 0.0):
        check(load().cpd_attn_guide(int(stage), int(n_images), int(hw), ptr(x), ptr(eps_u), DTYPE_CODE[src.dtype], int(eps_stride), ptr(mask_mean),
                                    int(mask_img_stride), ptr(pct), ptr(blur), ptr(sum16), ptr(e_attn), float(sigma_hat), float(c_in), int(mode),
                                    float(scale), float(guidance), ptr(out), stream_ptr()), "cpd_attn_guide")
    _count()
    return out
