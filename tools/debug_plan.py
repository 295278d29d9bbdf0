import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from complex_prompt_diffusion_b200.models.unet import UNetModel
from complex_prompt_diffusion_b200.models import fixtures
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_sampling import _unet_pair, rel
DEV = "cuda:0"
cfg, oracle, gpu = _unet_pair("tiny", torch.float32)
g = torch.Generator().manual_seed(3)
n, hw = 3, 16
x = torch.randn(n, 4, hw, hw, generator=g).to(torch.bfloat16).float()
t = torch.tensor([937.93, 11.278, 500.5]).to(torch.bfloat16).float()
ctx = torch.randn(n, 77, cfg.context_dim, generator=g).to(torch.bfloat16).float()
oracle.sd = {k: v.to(torch.bfloat16).float() for k, v in oracle.sd.items()}
oracle.taps = {}
ref = oracle(x, t, ctx)
out = gpu(x.to(DEV), t.to(DEV), ctx.to(DEV))
torch.cuda.synchronize()
for r in range(n):
    print("row", r, rel(out[r], ref[r]))
for name, tref in oracle.taps.items():
    try:
        buf = gpu.plan_buffer(name + ".out")
    except RuntimeError:
        continue
    if buf.numel() == tref.numel():
        got = buf.view(n, tref.shape[2], tref.shape[3], tref.shape[1]).permute(0, 3, 1, 2)
        print(f"{name:40s}", " ".join(f"{rel(got[r], tref[r]):.2e}" for r in range(n)))
