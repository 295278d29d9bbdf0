"""Oracle (test infrastructure): functional restatement of the reference UNet forward.

Follows cpd/models/unet.py:169-280,415-831, cpd/models/attention.py:89-118,171-190,280-348,443-537
and cpd/models/util.py:65-105 (paths relative to /root/reference).  Parameter names are the
reference module's ``state_dict`` keys, so one weight dict drives the reference module (when it can
be shimmed), this oracle and the CUDA path.

Numeric modes: ``dtype=torch.float32`` is the reference fp32 path verbatim.  ``torch.bfloat16`` /
``torch.float16`` restate the product mode of the reference (weights ``.half()`` + autocast,
manager.py:25-36, prompts.py:374): matmul/conv inputs in the model dtype, GroupNorm in fp32
(models/util.py:103-105), softmax in q's dtype (attention.py:337), attention output buffer fp32
(attention.py:299) cast back to the model dtype for ``to_out``.
"""
import math
from dataclasses import dataclass, field
from typing import List

import torch
import torch.nn.functional as F


@dataclass
class UNetConfig:
    """cpd/config/config-1.49.yaml:27-42 (SD-1.x) by default; sd21() = cpd/config/v2-inference.yaml:20-37."""
    in_channels: int = 4
    out_channels: int = 4
    model_channels: int = 320
    num_res_blocks: int = 2
    attention_resolutions: List[int] = field(default_factory=lambda: [4, 2, 1])
    channel_mult: List[int] = field(default_factory=lambda: [1, 2, 4, 4])
    num_heads: int = 8
    num_head_channels: int = -1
    transformer_depth: object = 1  # int, or one depth per level (SDXL: [1, 2, 10]); the middle block uses the last
    context_dim: int = 768
    use_linear_in_transformer: bool = False
    adm_in_channels: int = 0  # > 0: vector conditioning y -> label_emb MLP added to the time embedding (SDXL: 2816)

    @staticmethod
    def sd15():
        return UNetConfig()

    @staticmethod
    def sd21():
        return UNetConfig(num_heads=-1, num_head_channels=64, context_dim=1024, use_linear_in_transformer=True)

    @staticmethod
    def sdxl():
        """SDXL-base UNet (BASELINE.json configs[3]).  NOT expressible by the reference constructor (scalar
        transformer_depth unet.py:467,593; num_classes only int / "continuous" :536-543; no yaml): an extension of the same
        block grammar - channel_mult [1, 2, 4], attention at ds 2 and 4 with 2 and 10 transformer blocks (10 in the middle),
        64-wide heads, linear projections, 2048-d context (two text encoders concatenated on the feature axis) and a
        2816-d vector conditioning through label_emb = Linear -> SiLU -> Linear added to the time embedding."""
        return UNetConfig(channel_mult=[1, 2, 4], attention_resolutions=[4, 2], num_heads=-1, num_head_channels=64,
                          transformer_depth=[1, 2, 10], context_dim=2048, use_linear_in_transformer=True, adm_in_channels=2816)

    @staticmethod
    def tiny_xl(context_dim=128):
        """Small config with the SDXL topology (per-level depths, label_emb) for fast tests."""
        return UNetConfig(model_channels=64, channel_mult=[1, 2, 4], attention_resolutions=[4, 2], num_heads=-1,
                          num_head_channels=32, transformer_depth=[1, 1, 2], context_dim=context_dim, num_res_blocks=1,
                          use_linear_in_transformer=True, adm_in_channels=80)

    def depth(self, level):
        td = self.transformer_depth
        return td if isinstance(td, int) else td[min(level, len(td) - 1)]

    @staticmethod
    def tiny(context_dim=64):
        """A small config with the same topology (for fast tests)."""
        return UNetConfig(model_channels=64, channel_mult=[1, 2], attention_resolutions=[1, 2], num_heads=2,
                          context_dim=context_dim, num_res_blocks=1)

    def heads(self, ch):
        """unet.py:571-578 with legacy=False."""
        if self.num_head_channels == -1:
            return self.num_heads, ch // self.num_heads
        return ch // self.num_head_channels, self.num_head_channels


def enumerate_blocks(cfg: UNetConfig):
    """Block structure exactly as built by UNetModel.__init__ (unet.py:545-727).

    Returns (input_blocks, middle, output_blocks); each block is a list of layer tuples:
      ("conv_in", cin, cout) | ("res", cin, cout) | ("attn", ch, depth) | ("down", ch) | ("up", ch)
    """
    mc = cfg.model_channels
    inputs = [[("conv_in", cfg.in_channels, mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            layers = [("res", ch, mult * mc)]
            ch = mult * mc
            if ds in cfg.attention_resolutions:
                layers.append(("attn", ch, cfg.depth(level)))
            inputs.append(layers)
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            inputs.append([("down", ch)])
            chans.append(ch)
            ds *= 2
    middle = [("res", ch, ch), ("attn", ch, cfg.depth(len(cfg.channel_mult) - 1)), ("res", ch, ch)]
    outputs = []
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, mc * mult)]
            ch = mc * mult
            if ds in cfg.attention_resolutions:
                layers.append(("attn", ch, cfg.depth(level)))
            if level and i == cfg.num_res_blocks:
                layers.append(("up", ch))
                ds //= 2
            outputs.append(layers)
    return inputs, middle, outputs


def param_shapes(cfg: UNetConfig):
    """name -> shape for every parameter, using the reference state_dict keys."""
    shapes = {}
    ted = cfg.model_channels * 4
    shapes["time_embed.0.weight"] = (ted, cfg.model_channels)
    shapes["time_embed.0.bias"] = (ted,)
    shapes["time_embed.2.weight"] = (ted, ted)
    shapes["time_embed.2.bias"] = (ted,)
    if cfg.adm_in_channels:
        shapes["label_emb.0.0.weight"] = (ted, cfg.adm_in_channels)
        shapes["label_emb.0.0.bias"] = (ted,)
        shapes["label_emb.0.2.weight"] = (ted, ted)
        shapes["label_emb.0.2.bias"] = (ted,)

    def res(p, cin, cout):
        shapes[p + "in_layers.0.weight"] = (cin,)
        shapes[p + "in_layers.0.bias"] = (cin,)
        shapes[p + "in_layers.2.weight"] = (cout, cin, 3, 3)
        shapes[p + "in_layers.2.bias"] = (cout,)
        shapes[p + "emb_layers.1.weight"] = (cout, ted)
        shapes[p + "emb_layers.1.bias"] = (cout,)
        shapes[p + "out_layers.0.weight"] = (cout,)
        shapes[p + "out_layers.0.bias"] = (cout,)
        shapes[p + "out_layers.3.weight"] = (cout, cout, 3, 3)
        shapes[p + "out_layers.3.bias"] = (cout,)
        if cin != cout:
            shapes[p + "skip_connection.weight"] = (cout, cin, 1, 1)
            shapes[p + "skip_connection.bias"] = (cout,)

    def attn(p, ch, depth):
        nh, dh = cfg.heads(ch)
        inner = nh * dh
        shapes[p + "norm.weight"] = (ch,)
        shapes[p + "norm.bias"] = (ch,)
        if cfg.use_linear_in_transformer:
            shapes[p + "proj_in.weight"] = (inner, ch)
            shapes[p + "proj_out.weight"] = (ch, inner)
        else:
            shapes[p + "proj_in.weight"] = (inner, ch, 1, 1)
            shapes[p + "proj_out.weight"] = (ch, inner, 1, 1)
        shapes[p + "proj_in.bias"] = (inner,)
        shapes[p + "proj_out.bias"] = (ch,)
        for d in range(depth):
            b = p + f"transformer_blocks.{d}."
            for a, cdim in (("attn1", inner), ("attn2", cfg.context_dim)):
                shapes[b + a + ".to_q.weight"] = (inner, inner)
                shapes[b + a + ".to_k.weight"] = (inner, cdim)
                shapes[b + a + ".to_v.weight"] = (inner, cdim)
                shapes[b + a + ".to_out.0.weight"] = (inner, inner)
                shapes[b + a + ".to_out.0.bias"] = (inner,)
            shapes[b + "ff.net.0.proj.weight"] = (inner * 8, inner)
            shapes[b + "ff.net.0.proj.bias"] = (inner * 8,)
            shapes[b + "ff.net.2.weight"] = (inner, inner * 4)
            shapes[b + "ff.net.2.bias"] = (inner,)
            for n in ("norm1", "norm2", "norm3"):
                shapes[b + n + ".weight"] = (inner,)
                shapes[b + n + ".bias"] = (inner,)

    def block(prefix, layers):
        for j, l in enumerate(layers):
            p = f"{prefix}{j}."
            if l[0] == "conv_in":
                shapes[p + "weight"] = (l[2], l[1], 3, 3)
                shapes[p + "bias"] = (l[2],)
            elif l[0] == "res":
                res(p, l[1], l[2])
            elif l[0] == "attn":
                attn(p, l[1], l[2])
            elif l[0] == "down":
                shapes[p + "op.weight"] = (l[1], l[1], 3, 3)
                shapes[p + "op.bias"] = (l[1],)
            elif l[0] == "up":
                shapes[p + "conv.weight"] = (l[1], l[1], 3, 3)
                shapes[p + "conv.bias"] = (l[1],)

    inputs, middle, outputs = enumerate_blocks(cfg)
    for i, layers in enumerate(inputs):
        block(f"input_blocks.{i}.", layers)
    block("middle_block.", middle)
    for i, layers in enumerate(outputs):
        block(f"output_blocks.{i}.", layers)
    shapes["out.0.weight"] = (cfg.model_channels,)
    shapes["out.0.bias"] = (cfg.model_channels,)
    shapes["out.2.weight"] = (cfg.out_channels, cfg.model_channels, 3, 3)
    shapes["out.2.bias"] = (cfg.out_channels,)
    return shapes


def make_weights(cfg: UNetConfig, seed=0, dtype=torch.float32):
    """Seeded NON-ZERO weight fixture shared by the oracle and the CUDA path.

    The reference zero-initialises 187 tensors (unet.py:235-237,732; attention.py:520-524) which
    makes eps == 0 at init; here every matrix is ~N(0, 1/fan_in) (so activations keep O(1) scale
    through the network), norm gains ~1 +- 0.1 and biases ~N(0, 0.02^2).
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(cfg).items():
        if len(shape) == 1:
            is_gain = name.endswith("weight")
            w = torch.randn(shape, generator=g)
            w = 1.0 + 0.1 * w if is_gain else 0.02 * w
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            w = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
        sd[name] = w.to(dtype)
    return sd


def timestep_embedding(timesteps, dim, max_period=10000):
    """models/util.py:65-85 (cos first, fp32)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def group_norm32(x, w, b, eps):
    """models/util.py:103-105: computed in fp32, cast back."""
    return F.group_norm(x.float(), 32, w.float(), b.float(), eps).type(x.dtype)


class OracleUNet:
    def __init__(self, cfg: UNetConfig, sd: dict, dtype=torch.float32):
        self.cfg = cfg
        self.dtype = dtype
        self.sd = {k: v.to(dtype) for k, v in sd.items()}
        self.inputs, self.middle, self.outputs = enumerate_blocks(cfg)
        self.taps = None  # optional dict name -> tensor of intermediate activations (for layer-wise parity)

    def _tap(self, name, t):
        if self.taps is not None:
            self.taps[name] = t.detach().float().clone()

    # --- layers --------------------------------------------------------------------------------
    def _res(self, p, x, emb):
        """unet.py:260-280 (no scale-shift, no up/down)."""
        sd = self.sd
        h = F.silu(group_norm32(x, sd[p + "in_layers.0.weight"], sd[p + "in_layers.0.bias"], 1e-5))
        h = F.conv2d(h, sd[p + "in_layers.2.weight"], sd[p + "in_layers.2.bias"], padding=1)
        e = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"]).type(h.dtype)
        h = h + e[:, :, None, None]
        h = F.silu(group_norm32(h, sd[p + "out_layers.0.weight"], sd[p + "out_layers.0.bias"], 1e-5))
        h = F.conv2d(h, sd[p + "out_layers.3.weight"], sd[p + "out_layers.3.bias"], padding=1)
        if (p + "skip_connection.weight") in sd:
            x = F.conv2d(x, sd[p + "skip_connection.weight"], sd[p + "skip_connection.bias"])
        return x + h

    def _cross_attention(self, p, x, context, nh):
        """attention.py:280-348 with a single slice (D5)."""
        sd = self.sd
        ctx = x if context is None else context
        q = F.linear(x, sd[p + "to_q.weight"])
        k = F.linear(ctx, sd[p + "to_k.weight"])
        v = F.linear(ctx, sd[p + "to_v.weight"])
        b, n, inner = q.shape
        dh = inner // nh

        def split(t):
            return t.reshape(b, t.shape[1], nh, dh).permute(0, 2, 1, 3).reshape(b * nh, t.shape[1], dh)

        q, k, v = split(q), split(k), split(v)
        s1 = torch.einsum("bid,bjd->bij", q, k) * (dh ** -0.5)
        s2 = s1.softmax(dim=-1, dtype=q.dtype)
        r1 = torch.zeros(q.shape[0], q.shape[1], v.shape[2])  # fp32 buffer (attention.py:299)
        r1[:, :] = torch.einsum("bij,bjd->bid", s2, v)
        r2 = r1.reshape(b, nh, n, dh).permute(0, 2, 1, 3).reshape(b, n, inner).to(x.dtype)
        return F.linear(r2, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])

    def _attn(self, p, x, context, depth=1):
        """SpatialTransformer.forward, attention.py:526-537 (+ BasicTransformerBlock :485-490, GEGLU :92-100)."""
        sd, cfg = self.sd, self.cfg
        b, c, hh, ww = x.shape
        nh, _ = cfg.heads(c)
        x_in = x
        x = F.group_norm(x, 32, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
        if cfg.use_linear_in_transformer:
            x = x.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
            x = F.linear(x, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])
        else:
            x = F.conv2d(x, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])
            x = x.permute(0, 2, 3, 1).reshape(b, hh * ww, -1)
        for d in range(depth):
            bp = p + f"transformer_blocks.{d}."
            dim = x.shape[-1]
            x = self._cross_attention(bp + "attn1.", F.layer_norm(x, (dim,), sd[bp + "norm1.weight"], sd[bp + "norm1.bias"], 1e-5), None, nh) + x
            self._tap(bp + "attn1", x)
            x = self._cross_attention(bp + "attn2.", F.layer_norm(x, (dim,), sd[bp + "norm2.weight"], sd[bp + "norm2.bias"], 1e-5), context, nh) + x
            self._tap(bp + "attn2", x)
            y = F.layer_norm(x, (dim,), sd[bp + "norm3.weight"], sd[bp + "norm3.bias"], 1e-5)
            y = F.linear(y, sd[bp + "ff.net.0.proj.weight"], sd[bp + "ff.net.0.proj.bias"])
            a, gate = y.chunk(2, dim=-1)
            y = a * F.gelu(gate)
            x = F.linear(y, sd[bp + "ff.net.2.weight"], sd[bp + "ff.net.2.bias"]) + x
            self._tap(bp + "ff", x)
        if cfg.use_linear_in_transformer:
            x = F.linear(x, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
            x = x.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
        else:
            x = x.reshape(b, hh, ww, -1).permute(0, 3, 1, 2)
            x = F.conv2d(x, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
        return x + x_in

    def _block(self, prefix, layers, h, emb, context):
        sd = self.sd
        for j, l in enumerate(layers):
            p = f"{prefix}{j}."
            if l[0] == "conv_in":
                h = F.conv2d(h, sd[p + "weight"], sd[p + "bias"], padding=1)
            elif l[0] == "res":
                h = self._res(p, h, emb)
            elif l[0] == "attn":
                h = self._attn(p, h, context, l[2])
            elif l[0] == "down":
                h = F.conv2d(h, sd[p + "op.weight"], sd[p + "op.bias"], stride=2, padding=1)  # unet.py:151-160
            elif l[0] == "up":
                h = F.interpolate(h, scale_factor=2, mode="nearest")  # unet.py:116
                h = F.conv2d(h, sd[p + "conv.weight"], sd[p + "conv.bias"], padding=1)
            self._tap(p[:-1], h)
        return h

    # --- forward -------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, x, timesteps, context, return_attn=False, y=None, return_feat=False, inject_feats=None, inject_feats_stop=10,
                 inject_attns=None, inject_attns_stop=10, **_):
        """UNetModel.forward, unet.py:765-831.  x:[R,4,h,w], timesteps:[R], context:[R,77,D]."""
        sd, cfg = self.sd, self.cfg
        t_emb = timestep_embedding(timesteps, cfg.model_channels).to(self.dtype)
        emb = F.linear(F.silu(F.linear(t_emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])),
                       sd["time_embed.2.weight"], sd["time_embed.2.bias"])
        if cfg.adm_in_channels:  # vector conditioning (SDXL extension): emb = emb + label_emb(y)
            assert y is not None and y.shape == (x.shape[0], cfg.adm_in_channels), "this UNet needs y [rows, adm_in_channels]"
            emb = emb + F.linear(F.silu(F.linear(y.to(self.dtype), sd["label_emb.0.0.weight"], sd["label_emb.0.0.bias"])),
                                 sd["label_emb.0.2.weight"], sd["label_emb.0.2.bias"])
        context = context.to(self.dtype)
        hs = []
        h = x.type(self.dtype)
        for i, layers in enumerate(self.inputs):
            h = self._block(f"input_blocks.{i}.", layers, h, emb, context)
            hs.append(h)
        h = self._block("middle_block.", self.middle, h, emb, context)
        skips, feats = [], []
        for i, layers in enumerate(self.outputs):
            skip = hs.pop()
            skips.append(skip)
            if inject_attns is not None and inject_attns_stop > i:  # unet.py:806-809: replace the skip tensor
                assert inject_attns[i].shape == skip.shape
                skip = inject_attns[i].to(self.dtype)
            if inject_feats is not None and inject_feats_stop > i:  # unet.py:810-813: replace the running feature map
                assert inject_feats[i].shape == h.shape
                h = inject_feats[i].to(self.dtype)
            h = torch.cat([h, skip], dim=1)
            h = self._block(f"output_blocks.{i}.", layers, h, emb, context)
            feats.append(h)  # unet.py:816-817 (return_feat)
        # unet.py:818 casts h back to x.dtype before self.out; under the product's autocast the final
        # conv then runs (and returns) in the model dtype.
        h = F.silu(group_norm32(h, sd["out.0.weight"], sd["out.0.bias"], 1e-5))
        out = F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)
        self._tap("out", out)
        if return_attn:  # unet.py:822-831
            return (out, skips, feats) if return_feat else (out, skips)
        return (out, feats) if return_feat else out

    def parameters(self):
        return iter(self.sd.values())


def count_flops(cfg: UNetConfig, h, w, ctx_len=77):
    """Algorithmic FLOPs per UNet row-evaluation: sum 2*M*N*K over conv/linear + 4*heads*d*Nq*Nk per
    attention (the FlopCounterMode convention of BASELINE.md section 3)."""
    inputs, middle, outputs = enumerate_blocks(cfg)
    ted = cfg.model_channels * 4
    conv = lin = att = 0
    lin += 2 * cfg.model_channels * ted + 2 * ted * ted
    if cfg.adm_in_channels:
        lin += 2 * cfg.adm_in_channels * ted + 2 * ted * ted
    hw = [h, w]

    def layer(l):
        nonlocal conv, lin, att
        px = hw[0] * hw[1]
        if l[0] == "conv_in":
            conv += 2 * px * l[2] * l[1] * 9
        elif l[0] == "res":
            conv += 2 * px * l[2] * l[1] * 9 + 2 * px * l[2] * l[2] * 9
            lin += 2 * ted * l[2]
            if l[1] != l[2]:
                conv += 2 * px * l[1] * l[2]
        elif l[0] == "attn":
            ch = l[1]
            nh, dh = cfg.heads(ch)
            inner = nh * dh
            pr = 2 * px * ch * inner * 2
            if cfg.use_linear_in_transformer:
                lin += pr
            else:
                conv += pr
            for _ in range(l[2]):
                lin += 2 * px * inner * inner * 4  # attn1 q,k,v,out
                lin += 2 * px * inner * inner * 2 + 2 * ctx_len * cfg.context_dim * inner * 2  # attn2
                lin += 2 * px * inner * inner * 8 + 2 * px * inner * 4 * inner  # ff
                att += 4 * nh * dh * px * px + 4 * nh * dh * px * ctx_len
        elif l[0] == "down":
            hw[0] //= 2
            hw[1] //= 2
            conv += 2 * hw[0] * hw[1] * l[1] * l[1] * 9
        elif l[0] == "up":
            hw[0] *= 2
            hw[1] *= 2
            conv += 2 * hw[0] * hw[1] * l[1] * l[1] * 9

    for layers in inputs:
        for l in layers:
            layer(l)
    for l in middle:
        layer(l)
    for layers in outputs:
        for l in layers:
            layer(l)
    conv += 2 * hw[0] * hw[1] * cfg.out_channels * cfg.model_channels * 9
    return {"conv": conv, "linear": lin, "attention": att, "total": conv + lin + att}
