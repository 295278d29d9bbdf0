"""Synthetic-model fixtures for benchmarks and tools: the architecture presets of BASELINE.json's configs, seeded random-init
`state_dict`s with the reference's parameter names and shapes (there is no network for checkpoints), and the algorithmic FLOP
enumerators used for the roofline figures.  Nothing here touches the oracle: `oracle/` stays test infrastructure that only the
tests, smoke() and the CPU baseline leg of bench.py use."""
import math

import torch

from .unet import _enumerate_blocks
from .vae import _decoder_blocks

UNET_PRESETS = {
    # cpd/config/config-1.49.yaml:27-42
    "sd15": dict(model_channels=320, channel_mult=(1, 2, 4, 4), attention_resolutions=(4, 2, 1), num_res_blocks=2, num_heads=8,
                 num_head_channels=-1, transformer_depth=1, context_dim=768, use_linear_in_transformer=False, adm_in_channels=0),
    # cpd/config/v2-inference.yaml:20-37
    "sd21": dict(model_channels=320, channel_mult=(1, 2, 4, 4), attention_resolutions=(4, 2, 1), num_res_blocks=2, num_heads=-1,
                 num_head_channels=64, transformer_depth=1, context_dim=1024, use_linear_in_transformer=True, adm_in_channels=0),
    # SDXL-base (extension of the block grammar)
    "sdxl": dict(model_channels=320, channel_mult=(1, 2, 4), attention_resolutions=(4, 2), num_res_blocks=2, num_heads=-1,
                 num_head_channels=64, transformer_depth=(1, 2, 10), context_dim=2048, use_linear_in_transformer=True,
                 adm_in_channels=2816),
    "tiny": dict(model_channels=64, channel_mult=(1, 2), attention_resolutions=(1, 2), num_res_blocks=1, num_heads=2,
                 num_head_channels=-1, transformer_depth=1, context_dim=64, use_linear_in_transformer=False, adm_in_channels=0),
    "tiny_xl": dict(model_channels=64, channel_mult=(1, 2, 4), attention_resolutions=(4, 2), num_res_blocks=1, num_heads=-1,
                    num_head_channels=32, transformer_depth=(1, 1, 2), context_dim=128, use_linear_in_transformer=True,
                    adm_in_channels=80),
}
VAE_PRESETS = {"sd": dict(ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4, embed_dim=4),
               "tiny": dict(ch=64, out_ch=3, ch_mult=(1, 2), num_res_blocks=1, z_channels=4, embed_dim=4)}


def unet_kwargs(name):
    """Constructor kwargs of UNetModel for a preset."""
    kw = dict(UNET_PRESETS[name])
    kw["num_classes"] = "sequential" if kw["adm_in_channels"] else None
    return kw


def _cfg(preset):
    cfg = dict(in_channels=4, out_channels=4)
    cfg.update(preset)
    return cfg


def _heads(cfg, ch):
    if cfg["num_head_channels"] == -1:
        return cfg["num_heads"], ch // cfg["num_heads"]
    return ch // cfg["num_head_channels"], cfg["num_head_channels"]


def unet_param_shapes(preset):
    """name -> shape with the state_dict keys of cpd/models/unet.py / attention.py."""
    cfg = _cfg(preset)
    mc = cfg["model_channels"]
    ted = mc * 4
    sh = {"time_embed.0.weight": (ted, mc), "time_embed.0.bias": (ted,), "time_embed.2.weight": (ted, ted), "time_embed.2.bias": (ted,)}
    if cfg["adm_in_channels"]:
        sh.update({"label_emb.0.0.weight": (ted, cfg["adm_in_channels"]), "label_emb.0.0.bias": (ted,),
                   "label_emb.0.2.weight": (ted, ted), "label_emb.0.2.bias": (ted,)})
    lin = cfg["use_linear_in_transformer"]

    def block(prefix, layers):
        for j, l in enumerate(layers):
            p = f"{prefix}{j}."
            if l[0] == "conv_in":
                sh[p + "weight"], sh[p + "bias"] = (l[2], l[1], 3, 3), (l[2],)
            elif l[0] == "res":
                cin, cout = l[1], l[2]
                sh.update({p + "in_layers.0.weight": (cin,), p + "in_layers.0.bias": (cin,), p + "in_layers.2.weight": (cout, cin, 3, 3),
                           p + "in_layers.2.bias": (cout,), p + "emb_layers.1.weight": (cout, ted), p + "emb_layers.1.bias": (cout,),
                           p + "out_layers.0.weight": (cout,), p + "out_layers.0.bias": (cout,),
                           p + "out_layers.3.weight": (cout, cout, 3, 3), p + "out_layers.3.bias": (cout,)})
                if cin != cout:
                    sh[p + "skip_connection.weight"], sh[p + "skip_connection.bias"] = (cout, cin, 1, 1), (cout,)
            elif l[0] == "attn":
                ch, depth = l[1], l[2]
                nh, dh = _heads(cfg, ch)
                inner = nh * dh
                sh[p + "norm.weight"], sh[p + "norm.bias"] = (ch,), (ch,)
                sh[p + "proj_in.weight"] = (inner, ch) if lin else (inner, ch, 1, 1)
                sh[p + "proj_out.weight"] = (ch, inner) if lin else (ch, inner, 1, 1)
                sh[p + "proj_in.bias"], sh[p + "proj_out.bias"] = (inner,), (ch,)
                for d in range(depth):
                    b = p + f"transformer_blocks.{d}."
                    for a, cdim in (("attn1", inner), ("attn2", cfg["context_dim"])):
                        sh.update({b + a + ".to_q.weight": (inner, inner), b + a + ".to_k.weight": (inner, cdim),
                                   b + a + ".to_v.weight": (inner, cdim), b + a + ".to_out.0.weight": (inner, inner),
                                   b + a + ".to_out.0.bias": (inner,)})
                    sh.update({b + "ff.net.0.proj.weight": (inner * 8, inner), b + "ff.net.0.proj.bias": (inner * 8,),
                               b + "ff.net.2.weight": (inner, inner * 4), b + "ff.net.2.bias": (inner,)})
                    for n in ("norm1", "norm2", "norm3"):
                        sh[b + n + ".weight"], sh[b + n + ".bias"] = (inner,), (inner,)
            elif l[0] == "down":
                sh[p + "op.weight"], sh[p + "op.bias"] = (l[1], l[1], 3, 3), (l[1],)
            elif l[0] == "up":
                sh[p + "conv.weight"], sh[p + "conv.bias"] = (l[1], l[1], 3, 3), (l[1],)

    inputs, middle, outputs = _enumerate_blocks(cfg)
    for i, layers in enumerate(inputs):
        block(f"input_blocks.{i}.", layers)
    block("middle_block.", middle)
    for i, layers in enumerate(outputs):
        block(f"output_blocks.{i}.", layers)
    sh.update({"out.0.weight": (mc,), "out.0.bias": (mc,), "out.2.weight": (cfg["out_channels"], mc, 3, 3), "out.2.bias": (cfg["out_channels"],)})
    return sh


def vae_param_shapes(preset):
    """name -> shape with the AutoencoderKL state_dict keys of the decode path (cpd/models/autoencoder.py)."""
    cfg = dict(preset)
    bottom, levels, last = _decoder_blocks(cfg)
    sh = {"post_quant_conv.weight": (cfg["z_channels"], cfg["embed_dim"], 1, 1), "post_quant_conv.bias": (cfg["z_channels"],)}
    d = "decoder."
    sh[d + "conv_in.weight"], sh[d + "conv_in.bias"] = (bottom, cfg["z_channels"], 3, 3), (bottom,)

    def res(p, cin, cout):
        sh.update({p + "norm1.weight": (cin,), p + "norm1.bias": (cin,), p + "conv1.weight": (cout, cin, 3, 3), p + "conv1.bias": (cout,),
                   p + "norm2.weight": (cout,), p + "norm2.bias": (cout,), p + "conv2.weight": (cout, cout, 3, 3), p + "conv2.bias": (cout,)})
        if cin != cout:
            sh[p + "nin_shortcut.weight"], sh[p + "nin_shortcut.bias"] = (cout, cin, 1, 1), (cout,)

    res(d + "mid.block_1.", bottom, bottom)
    a = d + "mid.attn_1."
    sh[a + "norm.weight"], sh[a + "norm.bias"] = (bottom,), (bottom,)
    for n in ("q", "k", "v", "proj_out"):
        sh[a + n + ".weight"], sh[a + n + ".bias"] = (bottom, bottom, 1, 1), (bottom,)
    res(d + "mid.block_2.", bottom, bottom)
    for i_level, blocks, has_up in levels:
        for i_block, (cin, cout) in enumerate(blocks):
            res(d + f"up.{i_level}.block.{i_block}.", cin, cout)
        if has_up:
            c = blocks[-1][1]
            sh[d + f"up.{i_level}.upsample.conv.weight"], sh[d + f"up.{i_level}.upsample.conv.bias"] = (c, c, 3, 3), (c,)
    sh.update({d + "norm_out.weight": (last,), d + "norm_out.bias": (last,), d + "conv_out.weight": (cfg["out_ch"], last, 3, 3),
               d + "conv_out.bias": (cfg["out_ch"],)})
    return sh


def random_state_dict(shapes, seed=0):
    """Seeded random-init weights: matrices ~ N(0, 1 / fan_in) (activations keep O(1) scale through the network), norm gains
    1 +- 0.1, biases ~ N(0, 0.02^2).  (The reference zero-initialises 187 tensors, which would make eps == 0.)"""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in shapes.items():
        if len(shape) == 1:
            w = torch.randn(shape, generator=g)
            w = 1.0 + 0.1 * w if name.endswith("weight") else 0.02 * w
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            w = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
        sd[name] = w
    return sd


def unet_flops(preset, h, w, ctx_len=77):
    """Algorithmic FLOPs per UNet row-evaluation: sum 2*M*N*K over every conv / linear + 4*heads*d*Nq*Nk per attention
    (softmax / norm / elementwise excluded)."""
    cfg = _cfg(preset)
    inputs, middle, outputs = _enumerate_blocks(cfg)
    mc = cfg["model_channels"]
    ted = mc * 4
    fl = 2 * mc * ted + 2 * ted * ted
    if cfg["adm_in_channels"]:
        fl += 2 * cfg["adm_in_channels"] * ted + 2 * ted * ted
    hw = [h, w]

    def layer(l):
        nonlocal fl
        px = hw[0] * hw[1]
        if l[0] == "conv_in":
            fl += 2 * px * l[2] * l[1] * 9
        elif l[0] == "res":
            fl += 2 * px * l[2] * l[1] * 9 + 2 * px * l[2] * l[2] * 9 + 2 * ted * l[2]
            if l[1] != l[2]:
                fl += 2 * px * l[1] * l[2]
        elif l[0] == "attn":
            ch = l[1]
            nh, dh = _heads(cfg, ch)
            inner = nh * dh
            fl += 2 * px * ch * inner * 2
            for _ in range(l[2]):
                fl += 2 * px * inner * inner * 4
                fl += 2 * px * inner * inner * 2 + 2 * ctx_len * cfg["context_dim"] * inner * 2
                fl += 2 * px * inner * inner * 8 + 2 * px * inner * 4 * inner
                fl += 4 * nh * dh * px * px + 4 * nh * dh * px * ctx_len
        elif l[0] == "down":
            hw[0] //= 2
            hw[1] //= 2
            fl += 2 * hw[0] * hw[1] * l[1] * l[1] * 9
        elif l[0] == "up":
            hw[0] *= 2
            hw[1] *= 2
            fl += 2 * hw[0] * hw[1] * l[1] * l[1] * 9

    for layers in inputs:
        for l in layers:
            layer(l)
    for l in middle:
        layer(l)
    for layers in outputs:
        for l in layers:
            layer(l)
    return fl + 2 * hw[0] * hw[1] * cfg["out_channels"] * mc * 9


def vae_flops(preset, h, w):
    """2*M*N*K over every conv + 4*C*T*T for the mid attention, per image, for a latent of h x w."""
    cfg = dict(preset)
    bottom, levels, last = _decoder_blocks(cfg)
    px = h * w
    fl = 2 * px * cfg["z_channels"] * cfg["embed_dim"] + 2 * px * bottom * cfg["z_channels"] * 9

    def res(cin, cout, px):
        return 2 * px * cout * cin * 9 + 2 * px * cout * cout * 9 + (2 * px * cin * cout if cin != cout else 0)

    fl += 2 * res(bottom, bottom, px) + 4 * 2 * px * bottom * bottom + 4 * bottom * px * px
    for _, blocks, has_up in levels:
        for cin, cout in blocks:
            fl += res(cin, cout, px)
        if has_up:
            px *= 4
            fl += 2 * px * blocks[-1][1] * blocks[-1][1] * 9
    return fl + 2 * px * cfg["out_ch"] * last * 9
