"""Import shims that let the UNMODIFIED reference modules load in the build container.

TEST INFRASTRUCTURE ONLY.  This file is only used by ``oracle/make_golden.py`` (run once, in the
container that has ``/root/reference``) to produce the committed fixtures under ``tests/golden``.
Nothing here is imported by the product package, the ``-m gpu`` tests, ``smoke()`` or ``bench.py``.

Why shims are needed:
  * ``cpd/__init__.py:1`` imports ``cpd.manager`` -> ``open_clip`` (absent) and ``cpd/vram.py:3`` has the
    ``Ordereddict`` typo, so ``import cpd`` fails on any interpreter.  We register ``cpd`` as a bare
    namespace module (its ``__init__`` never runs) and provide a stub ``cpd.vram`` that exposes the one
    symbol the path uses (``device_lookup``, ``cpd/vram.py:12-19``) with every entry mapped to the CPU.
  * ``IPython``, ``matplotlib``, ``skimage`` are absent here; they are only used for notebook display.
No reference source is copied: the reference files are imported from where they lie.
"""
import sys
import types

REF_ROOT = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    import torch

    if "cpd" in sys.modules and getattr(sys.modules["cpd"], "_shimmed", False):
        return sys.modules["cpd"]
    cpd = types.ModuleType("cpd")
    cpd.__path__ = [REF_ROOT + "/cpd"]
    cpd._shimmed = True
    sys.modules["cpd"] = cpd
    cpu = torch.device("cpu")
    _stub("cpd.vram", device_lookup={k: cpu for k in ("c", "cpu", "g", "gpu", "cuda", "device")})
    disp = _stub("IPython.display", clear_output=lambda **k: None, display=lambda *a, **k: None)
    _stub("IPython", display=disp)
    plt = _stub("matplotlib.pyplot")
    _stub("matplotlib", pyplot=plt)
    meas = _stub("skimage.measure")
    expo = _stub("skimage.exposure")
    _stub("skimage", measure=meas, exposure=expo)
    lc = _stub("omegaconf.listconfig", ListConfig=type("ListConfig", (list,), {}))
    _stub("omegaconf", listconfig=lc, OmegaConf=object)
    # D6: hard-coded .cuda() calls (denoiser.py:362,384-392; k_diffusion.py:68,73) -> identity on CPU
    torch.Tensor.cuda = lambda self, *a, **k: self
    # D5: CrossAttention.forward probes CUDA memory (attention.py:301-306); report "plenty" so steps == 1
    torch.cuda.memory_stats = lambda *a, **k: {"active_bytes.all.current": 0, "reserved_bytes.all.current": 1 << 50}
    torch.cuda.mem_get_info = lambda *a, **k: (1 << 50, 1 << 50)
    torch.cuda.current_device = lambda: 0
    # cpd/models/autoencoder.py:9 imports taming's VectorQuantizer (absent here; only VQModel uses it, not the KL decoder)
    for name in ("taming", "taming.modules", "taming.modules.vqvae"):
        _stub(name)
    _stub("taming.modules.vqvae.quantize", VectorQuantizer2=object)
    import cpd.samplers  # noqa: F401  (first, to dodge the scheduler<->samplers import cycle)
    # D4: unet.py:592-596,649-653,703-707 pass use_linear=/use_checkpoint= which
    # SpatialTransformer.__init__ (attention.py:500-502) does not accept -> drop them.
    import cpd.models.attention as RA
    _orig = RA.SpatialTransformer.__init__

    def _init(self, *a, use_linear=False, use_checkpoint=False, **k):
        _orig(self, *a, **k)

    RA.SpatialTransformer.__init__ = _init
    return cpd


def build_reference_unet(cfg):
    """Instantiate the reference UNetModel (cpd/models/unet.py:415) for an oracle UNetConfig."""
    import cpd.models.unet as RU

    return RU.UNetModel(image_size=32, in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                        model_channels=cfg.model_channels, attention_resolutions=cfg.attention_resolutions,
                        num_res_blocks=cfg.num_res_blocks, channel_mult=cfg.channel_mult,
                        num_heads=cfg.num_heads, num_head_channels=cfg.num_head_channels,
                        use_spatial_transformer=True, transformer_depth=cfg.transformer_depth,
                        context_dim=cfg.context_dim, use_checkpoint=False, legacy=False)


def build_reference_vae_decode(cfg, sd):
    """AutoencoderKL.decode (cpd/models/autoencoder.py:825-828) from its two parts - post_quant_conv (:800) and Decoder (:380) -
    without the loss / encoder the full class would instantiate.  Returns a callable z -> image."""
    import torch
    import cpd.models.autoencoder as AE

    dec = AE.Decoder(ch=cfg.ch, out_ch=cfg.out_ch, ch_mult=tuple(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks,
                     attn_resolutions=[], in_channels=3, resolution=256, z_channels=cfg.z_channels)
    pqc = torch.nn.Conv2d(cfg.embed_dim, cfg.z_channels, 1)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}, strict=True)
    pqc.load_state_dict({"weight": sd["post_quant_conv.weight"], "bias": sd["post_quant_conv.bias"]})
    dec.eval()

    @torch.no_grad()
    def decode(z):
        return dec(pqc(z))
    return decode
