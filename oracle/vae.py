"""Oracle (test infrastructure): CPU restatement of the first-stage DECODER - `AutoencoderKL.decode`
(cpd/models/autoencoder.py:825-828) = post_quant_conv (:800) + `Decoder.forward` (:453-509) with `ResnetBlock` (:153-179,
temb = None), `AttnBlock` (:214-274), `Upsample` (:87-91), `Normalize` = GroupNorm(32, eps 1e-6) (:73-74) - the step right
after the denoising loop (callers `prompts.py:324-334,459,472-480`).

Defect repaired like D5: `AttnBlock.forward` probes `torch.cuda.memory_stats` (:232-245) to pick a slice count and so cannot
run on the CPU; the arithmetic restated here is the single-slice path (w = softmax(q k / sqrt(c)); h = v w^T; proj_out; + x).
Pinned against the shimmed reference Decoder in tests/golden/ref_vae.npz (oracle/make_golden.py).
"""
from dataclasses import dataclass, field
from typing import List

import math
import torch
import torch.nn.functional as F


@dataclass
class VAEConfig:
    """ddconfig of cpd/config/config-1.49.yaml:50-64 (the SD-1.x / 2.x first stage) by default."""
    ch: int = 128
    out_ch: int = 3
    ch_mult: List[int] = field(default_factory=lambda: [1, 2, 4, 4])
    num_res_blocks: int = 2
    z_channels: int = 4
    embed_dim: int = 4

    @staticmethod
    def sd():
        return VAEConfig()

    @staticmethod
    def tiny():
        return VAEConfig(ch=64, ch_mult=[1, 2], num_res_blocks=1)


def decoder_blocks(cfg: VAEConfig):
    """Decoder.__init__ (autoencoder.py:398-449): returns (block_in at the bottom, levels) where levels is a list, in
    EXECUTION order (lowest resolution first), of (i_level, [(cin, cout) ...], has_upsample)."""
    nres = len(cfg.ch_mult)
    block_in = cfg.ch * cfg.ch_mult[nres - 1]
    bottom = block_in
    levels = []
    for i_level in reversed(range(nres)):
        block_out = cfg.ch * cfg.ch_mult[i_level]
        blocks = []
        for _ in range(cfg.num_res_blocks + 1):
            blocks.append((block_in, block_out))
            block_in = block_out
        levels.append((i_level, blocks, i_level != 0))
    return bottom, levels, block_in


def param_shapes(cfg: VAEConfig):
    """name -> shape with the AutoencoderKL state_dict keys of the decode path."""
    sh = {"post_quant_conv.weight": (cfg.z_channels, cfg.embed_dim, 1, 1), "post_quant_conv.bias": (cfg.z_channels,)}
    bottom, levels, last = decoder_blocks(cfg)
    d = "decoder."
    sh[d + "conv_in.weight"] = (bottom, cfg.z_channels, 3, 3)
    sh[d + "conv_in.bias"] = (bottom,)

    def res(p, cin, cout):
        sh[p + "norm1.weight"] = (cin,)
        sh[p + "norm1.bias"] = (cin,)
        sh[p + "conv1.weight"] = (cout, cin, 3, 3)
        sh[p + "conv1.bias"] = (cout,)
        sh[p + "norm2.weight"] = (cout,)
        sh[p + "norm2.bias"] = (cout,)
        sh[p + "conv2.weight"] = (cout, cout, 3, 3)
        sh[p + "conv2.bias"] = (cout,)
        if cin != cout:
            sh[p + "nin_shortcut.weight"] = (cout, cin, 1, 1)
            sh[p + "nin_shortcut.bias"] = (cout,)

    res(d + "mid.block_1.", bottom, bottom)
    a = d + "mid.attn_1."
    sh[a + "norm.weight"] = (bottom,)
    sh[a + "norm.bias"] = (bottom,)
    for n in ("q", "k", "v", "proj_out"):
        sh[a + n + ".weight"] = (bottom, bottom, 1, 1)
        sh[a + n + ".bias"] = (bottom,)
    res(d + "mid.block_2.", bottom, bottom)
    for i_level, blocks, has_up in levels:
        for i_block, (cin, cout) in enumerate(blocks):
            res(d + f"up.{i_level}.block.{i_block}.", cin, cout)
        if has_up:
            c = blocks[-1][1]
            sh[d + f"up.{i_level}.upsample.conv.weight"] = (c, c, 3, 3)
            sh[d + f"up.{i_level}.upsample.conv.bias"] = (c,)
    sh[d + "norm_out.weight"] = (last,)
    sh[d + "norm_out.bias"] = (last,)
    sh[d + "conv_out.weight"] = (cfg.out_ch, last, 3, 3)
    sh[d + "conv_out.bias"] = (cfg.out_ch,)
    return sh


def make_weights(cfg: VAEConfig, seed=0, dtype=torch.float32):
    """Seeded weight fixture shared by the oracle and the CUDA path (matrices ~ N(0, 1/fan_in), gains 1 +- 0.1)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(cfg).items():
        if len(shape) == 1:
            w = torch.randn(shape, generator=g)
            w = 1.0 + 0.1 * w if name.endswith("weight") else 0.02 * w
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            w = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
        sd[name] = w.to(dtype)
    return sd


class OracleVAEDecoder:
    def __init__(self, cfg: VAEConfig, sd: dict, dtype=torch.float32):
        self.cfg, self.dtype = cfg, dtype
        self.sd = {k: v.to(dtype) for k, v in sd.items()}
        self.taps = None

    def _tap(self, name, t):
        if self.taps is not None:
            self.taps[name] = t.detach().float().clone()

    def _norm(self, p, x):
        return F.group_norm(x, 32, self.sd[p + "weight"], self.sd[p + "bias"], 1e-6)  # Normalize, :73-74

    def _res(self, p, x):
        """ResnetBlock.forward with temb = None (:153-179)."""
        sd = self.sd
        h = F.conv2d(F.silu(self._norm(p + "norm1.", x)), sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
        h = F.conv2d(F.silu(self._norm(p + "norm2.", h)), sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
        if (p + "nin_shortcut.weight") in sd:
            x = F.conv2d(x, sd[p + "nin_shortcut.weight"], sd[p + "nin_shortcut.bias"])
        return x + h

    def _attn(self, p, x):
        """AttnBlock.forward (:214-274), single slice."""
        sd = self.sd
        h_ = self._norm(p + "norm.", x)
        q = F.conv2d(h_, sd[p + "q.weight"], sd[p + "q.bias"])
        k = F.conv2d(h_, sd[p + "k.weight"], sd[p + "k.bias"])
        v = F.conv2d(h_, sd[p + "v.weight"], sd[p + "v.bias"])
        b, c, hh, ww = q.shape
        q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
        k = k.reshape(b, c, hh * ww)
        w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
        w_ = F.softmax(w_, dim=2)
        h2 = torch.bmm(v.reshape(b, c, hh * ww), w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
        return F.conv2d(h2, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"]) + x

    @torch.no_grad()
    def __call__(self, z):
        """AutoencoderKL.decode (:825-828): z [B, z_channels, h, w] -> image [B, out_ch, 8h.., 8w..]."""
        sd, cfg = self.sd, self.cfg
        d = "decoder."
        z = F.conv2d(z.to(self.dtype), sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
        h = F.conv2d(z, sd[d + "conv_in.weight"], sd[d + "conv_in.bias"], padding=1)
        self._tap("conv_in", h)
        h = self._res(d + "mid.block_1.", h)
        h = self._attn(d + "mid.attn_1.", h)
        self._tap("mid.attn_1", h)
        h = self._res(d + "mid.block_2.", h)
        _, levels, _ = decoder_blocks(cfg)
        for i_level, blocks, has_up in levels:
            for i_block in range(len(blocks)):
                h = self._res(d + f"up.{i_level}.block.{i_block}.", h)
            if has_up:
                h = F.interpolate(h, scale_factor=2.0, mode="nearest")  # Upsample.forward :87-91
                h = F.conv2d(h, sd[d + f"up.{i_level}.upsample.conv.weight"], sd[d + f"up.{i_level}.upsample.conv.bias"], padding=1)
            self._tap(f"up.{i_level}", h)
        h = F.silu(self._norm(d + "norm_out.", h))
        return F.conv2d(h, sd[d + "conv_out.weight"], sd[d + "conv_out.bias"], padding=1)


def count_flops(cfg: VAEConfig, h, w):
    """2*M*N*K over every conv + 4*c*T*T for the mid attention, per image, for a latent of h x w."""
    bottom, levels, last = decoder_blocks(cfg)
    px = h * w
    fl = 2 * px * cfg.z_channels * cfg.embed_dim + 2 * px * bottom * cfg.z_channels * 9

    def res(cin, cout, px):
        f = 2 * px * cout * cin * 9 + 2 * px * cout * cout * 9
        return f + (2 * px * cin * cout if cin != cout else 0)

    fl += 2 * res(bottom, bottom, px) + 4 * 2 * px * bottom * bottom + 4 * bottom * px * px
    for _, blocks, has_up in levels:
        for cin, cout in blocks:
            fl += res(cin, cout, px)
        if has_up:
            px *= 4
            fl += 2 * px * blocks[-1][1] * blocks[-1][1] * 9
    return fl + 2 * px * cfg.out_ch * last * 9
