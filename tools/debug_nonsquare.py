import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from complex_prompt_diffusion_b200.models.unet import UNetModel
from oracle.unet import UNetConfig, OracleUNet, make_weights
cfg = UNetConfig.tiny()
sd = make_weights(cfg, seed=0)
oracle = OracleUNet(cfg, {k: v.to(torch.bfloat16).float() for k, v in sd.items()})
gpu = UNetModel(sd, device="cuda", model_channels=cfg.model_channels, channel_mult=tuple(cfg.channel_mult),
                attention_resolutions=tuple(cfg.attention_resolutions), num_res_blocks=cfg.num_res_blocks,
                num_heads=cfg.num_heads, context_dim=cfg.context_dim, use_cuda_graph=False)
g = torch.Generator().manual_seed(77)
for (h, w) in ((16, 32), (32, 16), (8, 64), (64, 8), (24, 40)):
    x = torch.randn(2, 4, h, w, generator=g)
    t = torch.tensor([700.0, 30.5]).to(torch.bfloat16).float()
    ctx = torch.randn(2, 77, cfg.context_dim, generator=g)
    oracle.taps = {}
    ref = oracle(x, t, ctx.to(torch.bfloat16).float())
    try:
        out = gpu(x.cuda(), t.cuda(), ctx.cuda())
        torch.cuda.synchronize()
    except Exception as e:
        print(h, w, "ERROR", str(e)[:150]); continue
    rel = lambda a, b: ((a.float().cpu() - b.float().cpu()).norm() / b.float().cpu().norm()).item()
    print(f"== {h}x{w}: final rel {rel(out, ref):.3e}")
    for name, r in oracle.taps.items():
        for key, buf in gpu._ws.items():
            if key[0] == name + ".out" and buf.numel() == r.numel():
                got = buf.view(2, r.shape[2], r.shape[3], r.shape[1]).permute(0, 3, 1, 2)
                e = rel(got, r)
                if e > 5e-3 or name.endswith("blocks.0.0"):
                    print(f"   {name:30s} {tuple(r.shape)} rel {e:.3e}")
