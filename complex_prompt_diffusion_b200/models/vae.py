"""VAEDecoder: the first-stage decoder on B200 - `AutoencoderKL.decode` of the reference (cpd/models/autoencoder.py:825-828:
post_quant_conv + `Decoder.forward` :453-509), the step right after the denoising loop (callers
`prompts.py:324-334,459,472-480`).  Same `state_dict` names (`post_quant_conv.*`, `decoder.*`), executed as a static plan of the
same hand-written sm_100a kernels as the UNet over NHWC 16-bit activations:

  3x3 / 1x1 convs, nin_shortcut, q / k / v / proj_out  -> cpd_gemm_conv (tcgen05 implicit GEMM, residual fused)
  GroupNorm(32, eps 1e-6) (+ SiLU)                      -> cpd_groupnorm
  AttnBlock (one head of width C = 512, 4096 tokens)    -> Q K^T and P V as two cpd_gemm_conv GEMMs per image around
                                                           cpd_softmax_rows (the head is too wide for the flash kernel's TMEM
                                                           layout, and at 1 head x 4096^2 the score matrix is 32 MB)
  nearest 2x upsample                                   -> cpd_upsample2x
  post_quant_conv, z -> C input conv, C -> 3 output conv -> cpd_pointwise_small, cpd_conv_in, cpd_conv_out

There is no CPU or PyTorch fallback.
"""
import torch

from .. import ops


def _decoder_blocks(cfg):
    """Decoder.__init__ (autoencoder.py:398-449) in execution order: (bottom channels, [(i_level, [(cin, cout)...], has_up)], last)."""
    mult = list(cfg["ch_mult"])
    block_in = cfg["ch"] * mult[-1]
    bottom = block_in
    levels = []
    for i_level in reversed(range(len(mult))):
        block_out = cfg["ch"] * mult[i_level]
        blocks = []
        for _ in range(cfg["num_res_blocks"] + 1):
            blocks.append((block_in, block_out))
            block_in = block_out
        levels.append((i_level, blocks, i_level != 0))
    return bottom, levels, block_in


class VAEDecoder:
    DEFAULTS = dict(ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4, embed_dim=4, attn_resolutions=(),
                    scale_factor=0.18215)

    def __init__(self, state_dict=None, device="cuda", act_dtype=torch.float16, **config):
        cfg = dict(self.DEFAULTS)
        cfg.update({k: v for k, v in config.items() if k in self.DEFAULTS})
        if list(cfg["attn_resolutions"]):
            raise NotImplementedError("attention inside the up blocks (attn_resolutions != []) is not used by the SD first stage")
        if cfg["out_ch"] > 4 or cfg["z_channels"] > 8:
            raise NotImplementedError("out_ch <= 4 and z_channels <= 8 are supported")
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("VAEDecoder runs on CUDA only: there is no CPU fallback")
        self.act_dtype = act_dtype
        self.bottom, self.levels, self.last = _decoder_blocks(cfg)
        self.w, self._ws = {}, {}
        if state_dict is not None:
            self.load_state_dict(state_dict)

    def parameters(self):
        for v in self.w.values():
            yield v

    def load_state_dict(self, sd, strict=True):
        dev, W = self.device, {}

        def bf(t):
            return t.detach().to(torch.bfloat16)

        def f32(name):
            return bf(sd[name]).float().contiguous().to(dev)

        def conv3(name):  # [Cout, Cin, 3, 3] -> [Cout, 3, 3, Cin] in the activation dtype (tensor-core operand)
            return bf(sd[name]).permute(0, 2, 3, 1).contiguous().to(dev).to(self.act_dtype)

        def mat(name):
            t = bf(sd[name])
            return t.reshape(t.shape[0], t.shape[1]).contiguous().to(dev).to(self.act_dtype)

        def res(p, cin, cout):
            W[p + "gn1.g"], W[p + "gn1.b"] = f32(p + "norm1.weight"), f32(p + "norm1.bias")
            W[p + "conv1.w"], W[p + "conv1.b"] = conv3(p + "conv1.weight"), f32(p + "conv1.bias")
            W[p + "gn2.g"], W[p + "gn2.b"] = f32(p + "norm2.weight"), f32(p + "norm2.bias")
            W[p + "conv2.w"], W[p + "conv2.b"] = conv3(p + "conv2.weight"), f32(p + "conv2.bias")
            if cin != cout:
                W[p + "skip.w"], W[p + "skip.b"] = mat(p + "nin_shortcut.weight"), f32(p + "nin_shortcut.bias")

        d = "decoder."
        W["pqc.w"] = bf(sd["post_quant_conv.weight"]).float().reshape(self.cfg["z_channels"], self.cfg["embed_dim"]).contiguous().to(dev)
        W["pqc.b"] = f32("post_quant_conv.bias")
        W["conv_in.w"] = bf(sd[d + "conv_in.weight"]).permute(0, 2, 3, 1).contiguous().to(dev)  # bf16 [C][3][3][z] (CUDA-core kernel)
        W["conv_in.b"] = f32(d + "conv_in.bias")
        res(d + "mid.block_1.", self.bottom, self.bottom)
        a = d + "mid.attn_1."
        W[a + "norm.g"], W[a + "norm.b"] = f32(a + "norm.weight"), f32(a + "norm.bias")
        for n in ("q", "k", "v", "proj_out"):
            W[a + n + ".w"], W[a + n + ".b"] = mat(a + n + ".weight"), f32(a + n + ".bias")
        res(d + "mid.block_2.", self.bottom, self.bottom)
        for i_level, blocks, has_up in self.levels:
            for i_block, (cin, cout) in enumerate(blocks):
                res(d + f"up.{i_level}.block.{i_block}.", cin, cout)
            if has_up:
                p = d + f"up.{i_level}.upsample."
                W[p + "w"], W[p + "b"] = conv3(p + "conv.weight"), f32(p + "conv.bias")
        W["out.gn.g"], W["out.gn.b"] = f32(d + "norm_out.weight"), f32(d + "norm_out.bias")
        wo = bf(sd[d + "conv_out.weight"]).permute(0, 2, 3, 1)  # [out_ch][3][3][C] -> padded to 4 output channels
        wo4 = torch.zeros(4, 3, 3, self.last, dtype=torch.bfloat16)
        wo4[: self.cfg["out_ch"]] = wo
        W["out.w"] = wo4.contiguous().to(dev)
        bo = torch.zeros(4)
        bo[: self.cfg["out_ch"]] = bf(sd[d + "conv_out.bias"]).float()
        W["out.b"] = bo.to(dev)
        self.w = W
        return self

    def _buf(self, name, numel, dtype=None):
        dtype = self.act_dtype if dtype is None else dtype
        key = (name, numel, dtype)
        t = self._ws.get(key)
        if t is None:
            t = torch.empty(numel, dtype=dtype, device=self.device)
            self._ws[key] = t
        return t

    # ---- layers ---------------------------------------------------------------------------------------
    def _res(self, p, x, cin, cout, B, h, w, stats, tag):
        """ResnetBlock.forward, temb = None (autoencoder.py:153-179)."""
        W, hw = self.w, h * w
        gn = self._buf("gn", B * hw * max(cin, cout))
        ops.groupnorm(x, W[p + "gn1.g"], W[p + "gn1.b"], gn, stats, n_img=B, hw=hw, c0=cin, eps=1e-6, silu=True)
        h1 = self._buf("h1", B * hw * cout)
        ops.gemm_conv(gn, W[p + "conv1.w"], h1, n_img=B, h=h, w=w, c0=cin, n_out=cout, ksize=3, bias=W[p + "conv1.b"])
        ops.groupnorm(h1, W[p + "gn2.g"], W[p + "gn2.b"], gn, stats, n_img=B, hw=hw, c0=cout, eps=1e-6, silu=True)
        if cin != cout:
            skip = self._buf("skip", B * hw * cout)
            ops.gemm_conv(x, W[p + "skip.w"], skip, n_img=B, h=h, w=w, c0=cin, n_out=cout, ksize=1, bias=W[p + "skip.b"])
        else:
            skip = x
        out = self._buf(tag, B * hw * cout)
        ops.gemm_conv(gn, W[p + "conv2.w"], out, n_img=B, h=h, w=w, c0=cout, n_out=cout, ksize=3, bias=W[p + "conv2.b"],
                      residual=skip, ld_res=cout)
        return out

    def _attn(self, p, x, C, B, h, w, stats):
        """AttnBlock.forward (autoencoder.py:214-274): one head of width C over T = h * w tokens."""
        W, T = self.w, h * w
        gn = self._buf("gn", B * T * C)
        ops.groupnorm(x, W[p + "norm.g"], W[p + "norm.b"], gn, stats, n_img=B, hw=T, c0=C, eps=1e-6, silu=False)
        q = self._buf("at.q", B * T * C)
        k = self._buf("at.k", B * T * C)
        ops.gemm_conv(gn, W[p + "q.w"], q, n_img=1, h=1, w=B * T, c0=C, n_out=C, bias=W[p + "q.b"])
        ops.gemm_conv(gn, W[p + "k.w"], k, n_img=1, h=1, w=B * T, c0=C, n_out=C, bias=W[p + "k.b"])
        o = self._buf("at.o", B * T * C)
        s = self._buf("at.s", T * T)
        vt = self._buf("at.vt", C * T)
        for b in range(B):
            rows = slice(b * T * C, (b + 1) * T * C)
            # V^T = W_v X^T (the bias of v is added after P V: softmax rows sum to one)
            ops.gemm_conv(W[p + "v.w"], gn[rows], vt, n_img=1, h=1, w=C, c0=C, n_out=T)
            ops.gemm_conv(q[rows], k[rows], s, n_img=1, h=1, w=T, c0=C, n_out=T)                    # S = Q K^T
            ops.softmax_rows(s, s, rows=T, cols=T, scale=float(int(C) ** (-0.5)))                    # :250-255
            ops.gemm_conv(s, vt, o[rows], n_img=1, h=1, w=T, c0=T, n_out=C, bias=W[p + "v.b"])       # O = P V (+ b_v)
        out = self._buf("mid.attn.out", B * T * C)
        ops.gemm_conv(o, W[p + "proj_out.w"], out, n_img=1, h=1, w=B * T, c0=C, n_out=C, bias=W[p + "proj_out.b"], residual=x, ld_res=C)
        return out

    @torch.no_grad()
    def decode(self, z, unscale=False):
        """z [B, z_channels, h, w] fp32 -> image [B, out_ch, H, W] fp32 (H = h * 2^(levels - 1)).  `unscale=True` applies the
        1 / scale_factor of `decode_first_stage` (z = z / 0.18215) inside the first kernel."""
        W, cfg = self.w, self.cfg
        z = z.to(self.device, torch.float32).contiguous()
        B, zc, h, w = z.shape
        if zc != cfg["embed_dim"]:
            raise ValueError(f"z must have {cfg['embed_dim']} channels")
        stats = self._buf("gn.stats", B * 64 * ops.GN_MAX_CHUNKS, torch.float64)
        z2 = self._buf("z2", B * cfg["z_channels"] * h * w, torch.float32).view(B, cfg["z_channels"], h, w)
        ops.pointwise_small(z, W["pqc.w"], W["pqc.b"], z2, n=B, cin=zc, cout=cfg["z_channels"], hw=h * w,
                            scale=(1.0 / cfg["scale_factor"]) if unscale else 1.0)
        C = self.bottom
        hcur = self._buf("conv_in.out", B * h * w * C)
        ops.conv_in(z2, W["conv_in.w"], W["conv_in.b"], hcur, n=B, cin=cfg["z_channels"], h=h, w=w, cout=C, scale=1.0, rows_per_image=1)
        d = "decoder."
        hcur = self._res(d + "mid.block_1.", hcur, C, C, B, h, w, stats, "mid.b1.out")
        hcur = self._attn(d + "mid.attn_1.", hcur, C, B, h, w, stats)
        hcur = self._res(d + "mid.block_2.", hcur, C, C, B, h, w, stats, "mid.b2.out")
        for i_level, blocks, has_up in self.levels:
            for i_block, (cin, cout) in enumerate(blocks):
                hcur = self._res(d + f"up.{i_level}.block.{i_block}.", hcur, cin, cout, B, h, w, stats, f"up.{i_level}.{i_block % 2}.out")
                C = cout
            if has_up:
                up = self._buf("up", B * 4 * h * w * C)
                ops.upsample2x(hcur, up, n=B, h=h, w=w, c=C)
                h, w = 2 * h, 2 * w
                p = d + f"up.{i_level}.upsample."
                hcur = self._buf(f"up.{i_level}.us.out", B * h * w * C)
                ops.gemm_conv(up, W[p + "w"], hcur, n_img=B, h=h, w=w, c0=C, n_out=C, ksize=3, bias=W[p + "b"])
        gn = self._buf("gn", B * h * w * C)
        ops.groupnorm(hcur, W["out.gn.g"], W["out.gn.b"], gn, stats, n_img=B, hw=h * w, c0=C, eps=1e-6, silu=True)
        out4 = self._buf("image", B * 4 * h * w, torch.float32).view(B, 4, h, w)
        ops.conv_out(gn, W["out.w"], W["out.b"], out4, n=B, h=h, w=w, cin=C, cout=4)
        return out4[:, : cfg["out_ch"]]

    def decode_to_uint8(self, z, unscale=True):
        """Latents -> display images: decode, then the reference's tail `clamp((x + 1) / 2, 0, 1).permute(0, 2, 3, 1).mul(255)
        .to(uint8)` (prompts.py:324-334,472-475) as one kernel.  Returns uint8 [B, H, W, out_ch] on the device."""
        img = self.decode(z, unscale=unscale)  # a view of the 4-channel fp32 buffer
        B, c, H, Wd = img.shape
        out = self._buf("image.u8", B * H * Wd * c, torch.uint8).view(B, H, Wd, c)
        ops.images_to_uint8(img, out, ld_c=4)
        return out

    __call__ = decode
