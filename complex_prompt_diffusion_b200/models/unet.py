"""UNetModel: the Stable-Diffusion UNet forward on B200, same call signature as the reference
`cpd/models/unet.py:415-831` (`forward(x, timesteps, context, y=None, return_attn=..., ...)`) and the same
`state_dict` parameter names.  This class is a BINDING: the executor - weight packing, the text-context K / V^T cache,
every workspace buffer, the sequencing of the ~850 kernel launches of an evaluation and the CUDA graph that replays
them - lives behind the plan-level C ABI of libcpd_b200.so (`cpd_unet_plan_create`, `cpd_pack_weights`,
`cpd_cache_context_kv`, `cpd_unet_forward`; include/cpd_b200.h, csrc/unet_plan.cu).  PyTorch supplies device memory
for the inputs / outputs and the CUDA stream.

  3x3 / 1x1 convs and every Linear  -> cpd_gemm_conv   (tcgen05 implicit GEMM, TMA-fed, fused bias /
                                       time-embedding / residual / GEGLU epilogues)
  self- and cross-attention          -> cpd_attention   (fused flash-style kernels, tcgen05 + TMEM)
  GroupNorm(+SiLU), LayerNorm        -> cpd_groupnorm / cpd_layernorm
  timestep embedding + emb MLPs      -> cpd_timestep_embedding / cpd_small_linear
  4->C input conv, C->4 output conv  -> cpd_conv_in / cpd_conv_out

Weights are bf16 (the model dtype of BASELINE.json's bf16 configs); GroupNorm statistics are fp32
(models/util.py:103-105), accumulation is fp32 in TMEM.  There is no CPU or PyTorch fallback.
"""
import ctypes as C

import torch

from .. import ops
from .._lib import CPD_BF16, CPD_F32, DTYPE_CODE, UNetConfig, UNetIO, check, device_tensor, load, stream_ptr


def _enumerate_blocks(cfg):
    """Block structure built by the reference constructor (unet.py:545-727); host-side copy used by the synthetic-model
    fixtures (parameter shapes, FLOP enumerator).  The executor's own copy is csrc/unet_plan.cu:enumerate_blocks."""
    mc = cfg["model_channels"]
    inputs = [[("conv_in", cfg["in_channels"], mc)]]
    chans = [mc]
    ch, ds = mc, 1
    mult_list = list(cfg["channel_mult"])
    td = cfg["transformer_depth"]

    def depth(level):  # int, or one depth per level (SDXL extension; the middle block uses the last level's)
        return td if isinstance(td, int) else list(td)[min(level, len(td) - 1)]

    for level, mult in enumerate(mult_list):
        for _ in range(cfg["num_res_blocks"]):
            layers = [("res", ch, mult * mc)]
            ch = mult * mc
            if ds in cfg["attention_resolutions"]:
                layers.append(("attn", ch, depth(level)))
            inputs.append(layers)
            chans.append(ch)
        if level != len(mult_list) - 1:
            inputs.append([("down", ch)])
            chans.append(ch)
            ds *= 2
    middle = [("res", ch, ch), ("attn", ch, depth(len(mult_list) - 1)), ("res", ch, ch)]
    outputs = []
    for level, mult in list(enumerate(mult_list))[::-1]:
        for i in range(cfg["num_res_blocks"] + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, mc * mult, ch, ich)]
            ch = mc * mult
            if ds in cfg["attention_resolutions"]:
                layers.append(("attn", ch, depth(level)))
            if level and i == cfg["num_res_blocks"]:
                layers.append(("up", ch))
                ds //= 2
            outputs.append(layers)
    return inputs, middle, outputs


class UNetModel:
    DEFAULTS = dict(image_size=32, in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2,
                    attention_resolutions=(4, 2, 1), channel_mult=(1, 2, 4, 4), num_heads=8, num_head_channels=-1,
                    transformer_depth=1, context_dim=768, use_linear_in_transformer=False, use_spatial_transformer=True,
                    legacy=False, adm_in_channels=0)

    def __init__(self, state_dict=None, device="cuda", act_dtype=torch.float16, eps_dtype=torch.float32, use_cuda_graph=True,
                 **config):
        cfg = dict(self.DEFAULTS)
        cfg.update({k: v for k, v in config.items() if k in self.DEFAULTS})
        if not cfg["use_spatial_transformer"] or cfg["legacy"]:
            raise NotImplementedError("only the SpatialTransformer (legacy=False) UNet of the SD configs is supported")
        if config.get("num_classes") not in (None, "sequential"):
            raise NotImplementedError("class-conditional UNets are not on the hot path (only the SDXL vector conditioning is)")
        if (config.get("num_classes") == "sequential") != bool(cfg["adm_in_channels"]):
            raise ValueError("vector conditioning needs num_classes='sequential' together with adm_in_channels > 0")
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("UNetModel runs on CUDA only: there is no CPU fallback for the hot path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype = torch.bfloat16  # model (weight) dtype
        if act_dtype not in (torch.float16, torch.bfloat16):
            raise ValueError("act_dtype must be torch.float16 or torch.bfloat16")
        if eps_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("eps_dtype must be torch.float32 or torch.bfloat16")
        # Inter-kernel activations are 16-bit.  fp16 (default) has 3 more mantissa bits than bf16; tcgen05 kind::f16 runs
        # both at the same rate ("precision").
        self.act_dtype = act_dtype
        self.eps_dtype = eps_dtype
        self.model_channels = cfg["model_channels"]
        self.adm = int(cfg["adm_in_channels"])
        self.use_cuda_graph = bool(use_cuda_graph)
        self.inputs, self.middle, self.outputs = _enumerate_blocks(cfg)
        self._lib = load()
        c = UNetConfig()
        c.in_channels, c.out_channels, c.model_channels = cfg["in_channels"], cfg["out_channels"], cfg["model_channels"]
        c.num_res_blocks = cfg["num_res_blocks"]
        mult = list(cfg["channel_mult"])
        c.n_levels = len(mult)
        td = cfg["transformer_depth"]
        for i, m in enumerate(mult):
            c.channel_mult[i] = int(m)
            c.transformer_depth[i] = int(td if isinstance(td, int) else list(td)[min(i, len(td) - 1)])
        ar = list(cfg["attention_resolutions"])
        c.n_attention_resolutions = len(ar)
        for i, a in enumerate(ar):
            c.attention_resolutions[i] = int(a)
        c.num_heads, c.num_head_channels = int(cfg["num_heads"]), int(cfg["num_head_channels"])
        c.context_dim, c.use_linear_in_transformer = int(cfg["context_dim"]), int(bool(cfg["use_linear_in_transformer"]))
        c.adm_in_channels = self.adm
        c.act_fp16 = int(act_dtype == torch.float16)
        c.eps_dtype = DTYPE_CODE[eps_dtype]
        c.use_cuda_graph = int(self.use_cuda_graph)
        self._plan = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self._lib.cpd_unet_plan_create(C.byref(c), C.byref(self._plan)), "cpd_unet_plan_create")
        self._static = {}   # (shape, rows) -> static input / scalar / output buffers of the graph-replayed fast path
        self._ctx_key = None
        self._ctx_shape = None
        self._y_key = None
        self._y_rows = 0
        self._probe = torch.zeros(1, dtype=self.dtype, device=self.device)
        if state_dict is not None:
            self.load_state_dict(state_dict)

    def __del__(self):
        plan, lib = getattr(self, "_plan", None), getattr(self, "_lib", None)
        if plan is not None and plan.value and lib is not None:
            lib.cpd_unet_plan_destroy(plan)
            self._plan = None

    # ---- nn.Module-like surface used by the Denoiser ------------------------------------------------
    def parameters(self):
        yield self._probe

    def heads(self, ch):
        if self.cfg["num_head_channels"] == -1:
            return self.cfg["num_heads"], ch // self.cfg["num_heads"]
        return ch // self.cfg["num_head_channels"], self.cfg["num_head_channels"]

    def _stream(self):
        return stream_ptr(self.device)

    # ---- weights: packed once by the library ---------------------------------------------------------
    def load_state_dict(self, sd, strict=True):
        """`sd`: the reference's state_dict (names of cpd/models/unet.py / attention.py).  Every entry is handed to
        cpd_pack_weights under its reference name; entries the UNet configuration does not have are rejected when `strict`."""
        with torch.cuda.device(self.device):
            for name, t in sd.items():
                t = t.detach()
                if t.dtype not in DTYPE_CODE:
                    t = t.float()
                t = t.contiguous()
                rc = self._lib.cpd_pack_weights(self._plan, name.encode(), C.c_void_p(t.data_ptr()), DTYPE_CODE[t.dtype], t.numel(),
                                                int(t.is_cuda))
                if rc != 0:
                    msg = self._lib.cpd_last_error().decode("utf-8", "replace")
                    if "is not a parameter" in msg and not strict:
                        continue
                    raise RuntimeError(f"cpd_pack_weights({name}) failed (status {rc}): {msg}")
            missing = self._lib.cpd_unet_plan_missing_weights(self._plan)
            if missing and strict:
                raise KeyError(f"{missing} state_dict entries are missing, first: "
                               f"{self._lib.cpd_last_error().decode('utf-8', 'replace')}")
        self._ctx_key = None  # the cached K / V^T were projected with the old weights
        return self

    # ---- text context: step-invariant K / V^T of every cross-attention layer ---------------------------
    def set_context(self, context):
        """context: [Rc, tokens, D]; query row b uses context row (b % Rc)."""
        # The cache is keyed on the tensor OBJECT (kept alive here, so its storage cannot be recycled for another prompt)
        # and its version counter - never on data_ptr(): the caching allocator hands the address of a freed context to
        # the next one of the same shape.
        if self._ctx_key is not None and self._ctx_key[0] is context and self._ctx_key[1] == context._version:
            return
        ctx = context.detach().to(self.device)
        if ctx.dtype not in DTYPE_CODE:
            ctx = ctx.float()
        ctx = ctx.contiguous()
        if ctx.ndim != 3 or ctx.shape[2] != self.cfg["context_dim"]:
            raise ValueError(f"context must be [rows, tokens, {self.cfg['context_dim']}], got {tuple(ctx.shape)}")
        with torch.cuda.device(self.device):
            check(self._lib.cpd_cache_context_kv(self._plan, C.c_void_p(ctx.data_ptr()), DTYPE_CODE[ctx.dtype], ctx.shape[0], ctx.shape[1],
                                                 self._stream()), "cpd_cache_context_kv")
        ops.LAUNCHES += int(self._lib.cpd_unet_plan_launches(self._plan))
        self._ctx_key = (context, context._version)
        self._ctx_shape = (ctx.shape[0], ctx.shape[1])
        self._keep_ctx = ctx  # read asynchronously by the pack kernel

    def set_vector(self, y):
        """Vector conditioning of the SDXL extension: y [Ry, adm_in_channels]; UNet row b uses y row (b % Ry).  Kept by the plan in a
        persistent bf16 buffer (the model dtype), so captured graphs stay valid when the prompt changes."""
        if not self.adm:
            if y is not None:
                raise ValueError("this UNet has no vector conditioning (adm_in_channels = 0)")
            return
        if y is None:
            raise ValueError(f"this UNet needs vector conditioning y [rows, {self.adm}]")
        if self._y_key is not None and self._y_key[0] is y and self._y_key[1] == y._version:
            return
        if y.ndim != 2 or y.shape[1] != self.adm:
            raise ValueError(f"y must be [rows, {self.adm}], got {tuple(y.shape)}")
        yd = y.detach().to(self.device)
        if yd.dtype not in DTYPE_CODE:
            yd = yd.float()
        yd = yd.contiguous()
        with torch.cuda.device(self.device):
            check(self._lib.cpd_unet_set_vector(self._plan, C.c_void_p(yd.data_ptr()), DTYPE_CODE[yd.dtype], yd.shape[0], self._stream()),
                  "cpd_unet_set_vector")
        ops.LAUNCHES += 1
        self._y_key, self._y_rows, self._keep_y = (y, y._version), yd.shape[0], yd

    # ---- plan buffers as torch tensors (zero-copy views of library-owned memory) -------------------------
    def plan_buffer(self, name, dtype=None):
        """A named buffer of the plan (packed weights, workspace, "<block>.out" activations) as a flat torch view."""
        ptr, n = C.c_void_p(), C.c_int64()
        check(self._lib.cpd_unet_plan_buffer(self._plan, name.encode(), C.byref(ptr), C.byref(n)), "cpd_unet_plan_buffer")
        return device_tensor(ptr.value, n.value, self.act_dtype if dtype is None else dtype, self.device)

    def _tap(self, which, index, rows):
        """Output of a block of the last forward as an NCHW view [rows, c, h, w] of the plan's NHWC buffer."""
        ptr, c, h, w = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
        check(self._lib.cpd_unet_plan_tap(self._plan, which, index, C.byref(ptr), C.byref(c), C.byref(h), C.byref(w)), "cpd_unet_plan_tap")
        t = device_tensor(ptr.value, rows * c.value * h.value * w.value, self.act_dtype, self.device)
        return t.view(rows, h.value, w.value, c.value).permute(0, 3, 1, 2)

    def profile(self, on):
        """Per-launch CUDA-event timing of the following eager forwards (bench.py's roofline leg)."""
        self._lib.cpd_unet_plan_set_profile(self._plan, int(bool(on)))

    def profile_records(self):
        """[(kind, label, microseconds, flops)] of the launches recorded since profile(True)."""
        n = self._lib.cpd_unet_plan_profile_dump(self._plan, None, 0)
        buf = C.create_string_buffer(int(n))
        self._lib.cpd_unet_plan_profile_dump(self._plan, buf, n)
        out = []
        for line in buf.value.decode().splitlines():
            kind, label, us, fl = line.split("\t")
            out.append((kind, label, float(us), float(fl)))
        return out

    # ---- forward ---------------------------------------------------------------------------------------
    def _nhwc(self, t, R, c, h, w):
        """A caller-provided NCHW tensor [R, c, h, w] -> NHWC activation tensor (feature / skip injection, unet.py:806-813)."""
        t = torch.as_tensor(t)
        if tuple(t.shape) != (R, c, h, w):
            raise ValueError(f"injected tensor has shape {tuple(t.shape)}, expected {(R, c, h, w)}")
        return t.to(self.device, self.act_dtype).permute(0, 2, 3, 1).contiguous()

    def _run(self, x, rows_per_image, c_in_dev, t_dev, t_count, eps, inject=None, no_graph=False):
        """One cpd_unet_forward call.  x: fp32 [B, C, h, w] device tensor; c_in_dev / t_dev: device scalars; eps: output tensor."""
        B, _cin, h, w = x.shape
        R = B * rows_per_image
        io = UNetIO()
        io.x, io.n_images, io.h, io.w, io.rows_per_image = x.data_ptr(), B, h, w, rows_per_image
        io.c_in = c_in_dev.data_ptr() if c_in_dev is not None else None
        io.t, io.t_count = t_dev.data_ptr(), t_count
        io.eps = eps.data_ptr()
        io.no_graph = int(bool(no_graph))
        keep = []
        if inject:
            # shapes of the popped skip tensors / decoder inputs come from the plan's own record of the last forward: run the
            # geometry from the block structure instead (channels of the input blocks, spatial size per level)
            n_out = len(self.outputs)
            skips = (C.c_void_p * n_out)()
            feats = (C.c_void_p * n_out)()
            geo = self._decoder_geometry(h, w)
            for i in range(n_out):
                (sc, sh, sw), (fc, fh, fw) = geo[i]
                if inject.get("attns") is not None and inject.get("attns_stop", 10) > i:  # unet.py:806-809: replace the skip tensor
                    tns = self._nhwc(inject["attns"][i], R, sc, sh, sw)
                    keep.append(tns)
                    skips[i] = tns.data_ptr()
                if inject.get("feats") is not None and inject.get("feats_stop", 10) > i:  # unet.py:810-813: replace h
                    tns = self._nhwc(inject["feats"][i], R, fc, fh, fw)
                    keep.append(tns)
                    feats[i] = tns.data_ptr()
            io.inject_skips = C.cast(skips, C.POINTER(C.c_void_p))
            io.inject_feats = C.cast(feats, C.POINTER(C.c_void_p))
            keep += [skips, feats]
        with torch.cuda.device(self.device):
            check(self._lib.cpd_unet_forward(self._plan, C.byref(io), self._stream()), "cpd_unet_forward")
        ops.LAUNCHES += int(self._lib.cpd_unet_plan_launches(self._plan))
        self._keep_inject = keep  # read asynchronously
        return eps

    def _decoder_geometry(self, h, w):
        """Per output block: (channels, h, w) of the popped skip tensor and of the decoder state entering the block."""
        mc = self.model_channels
        hs = [(mc, h, w)]
        ch, ch_h, ch_w = mc, h, w
        for layers in self.inputs[1:]:
            for l in layers:
                if l[0] == "res":
                    ch = l[2]
                elif l[0] == "down":
                    ch_h, ch_w = ch_h // 2, ch_w // 2
            hs.append((ch, ch_h, ch_w))
        geo = []
        for layers in self.outputs:
            s = hs.pop()
            geo.append((s, (ch, ch_h, ch_w)))
            for l in layers:
                if l[0] == "res":
                    ch = l[2]
                elif l[0] == "up":
                    ch_h, ch_w = 2 * ch_h, 2 * ch_w
        return geo

    @torch.no_grad()
    def forward_rows(self, x, c_in, t, rows_per_image, inject=None):
        """Fast path used by the Denoiser: x [B,4,h,w] fp32 (unscaled), every image is evaluated on
        `rows_per_image` conditioning rows sharing x * c_in and the timestep t (denoiser.py:383-393).
        Returns eps rows [B*rows_per_image, 4, h, w] (eps_dtype), image-major.  The result lives in a buffer that the next
        call with the same shape overwrites."""
        if self._ctx_key is None:
            raise RuntimeError("UNetModel: no text context set (call set_context or pass `context`)")
        if self.adm and not self._y_rows:
            raise RuntimeError("UNetModel: no vector conditioning set (call set_vector or pass `y`)")
        key = (tuple(x.shape), rows_per_image)
        st = self._static.get(key)
        if st is None:
            # static buffers: the plan's CUDA graph of this shape is captured on THESE addresses and replayed per step; the
            # per-step scalars (c_in, t) live in a 2-float device buffer
            B, cout = x.shape[0], self.cfg["out_channels"]
            st = dict(x=torch.empty(tuple(x.shape), dtype=torch.float32, device=self.device),
                      sc=torch.empty(2, dtype=torch.float32, device=self.device),
                      eps=torch.empty(B * rows_per_image, cout, x.shape[2], x.shape[3], dtype=self.eps_dtype, device=self.device))
            self._static[key] = st
        st["x"].copy_(x, non_blocking=True)
        st["sc"][0:1].fill_(float(c_in))  # scalars travel by value in the fill kernels' parameters (no host buffer to race on)
        st["sc"][1:2].fill_(float(t))
        return self._run(st["x"], rows_per_image, st["sc"][0:1], st["sc"][1:2], 1, st["eps"], inject=inject,
                         no_graph=not self.use_cuda_graph)

    @torch.no_grad()
    def forward(self, x, timesteps=None, context=None, y=None, **kwargs):
        """Reference call signature (unet.py:765): x [N,4,h,w], timesteps [N], context [N or Rc, tokens, D].
        Returns out [N,4,h,w] (eps_dtype) or (out, skips) when return_attn=True (the 12 skip tensors, unet.py:802-804);
        with return_feat=True the per-output-block features are appended (unet.py:816-817).  Results are owned tensors."""
        if y is not None or self.adm:
            self.set_vector(y)
        inject = None
        if kwargs.get("inject_feats") is not None or kwargs.get("inject_attns") is not None:
            inject = dict(feats=kwargs.get("inject_feats"), feats_stop=kwargs.get("inject_feats_stop", 10),
                          attns=kwargs.get("inject_attns"), attns_stop=kwargs.get("inject_attns_stop", 10))
        if context is not None:
            self.set_context(context)
        if self._ctx_key is None:
            raise RuntimeError("UNetModel: no text context set (call set_context or pass `context`)")
        n = x.shape[0]
        x = x.to(self.device, torch.float32).contiguous()
        t_rows = torch.as_tensor(timesteps).to(self.device, torch.float32).reshape(-1).contiguous()
        if t_rows.numel() != n:
            raise ValueError("timesteps must have one entry per row of x")
        if n > 32:
            raise NotImplementedError("more than 32 rows with distinct timesteps: use forward_rows")
        out = torch.empty(n, self.cfg["out_channels"], x.shape[2], x.shape[3], dtype=self.eps_dtype, device=self.device)
        self._run(x, 1, None, t_rows, n, out, inject=inject, no_graph=True)
        want_skips, want_feats = bool(kwargs.get("return_attn", False)), bool(kwargs.get("return_feat", False))
        if not (want_skips or want_feats):
            return out
        # the reference pops the skip tensors in decoder order (unet.py:801-804): last input block first
        skips = [self._tap(0, i, n).clone() for i in reversed(range(len(self.inputs)))] if want_skips else None
        feats = [self._tap(2, i, n).float().clone() for i in range(len(self.outputs))] if want_feats else None
        if want_feats:
            return (out, skips, feats) if want_skips else (out, feats)
        return out, skips

    __call__ = forward
