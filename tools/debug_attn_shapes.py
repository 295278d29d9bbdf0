import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from complex_prompt_diffusion_b200 import ops
def run(B, H, Nq, Nk, d):
    g = torch.Generator().manual_seed(B * Nq + d)
    dpad = (d + 15) // 16 * 16
    nk_pad = (Nk + 15) // 16 * 16
    bf = lambda t: t.to(torch.bfloat16)
    q, k, v = bf(torch.randn(B, Nq, H, d, generator=g)), bf(torch.randn(B, Nk, H, d, generator=g)), bf(torch.randn(B, Nk, H, d, generator=g))
    qp = torch.zeros(B, Nq, H, dpad, dtype=torch.bfloat16); qp[..., :d] = q
    kp = torch.zeros(B, nk_pad, H, dpad, dtype=torch.bfloat16); kp[:, :Nk, :, :d] = k
    vp = torch.zeros(B, nk_pad, H, dpad, dtype=torch.bfloat16); vp[:, :Nk, :, :d] = v
    vt = vp.permute(2, 3, 0, 1).reshape(H * dpad, B * nk_pad).contiguous().cuda()
    o = torch.full((B, Nq, H, dpad), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.attention(qp.cuda(), kp.cuda(), vt, o, ldq=H * dpad, ldk=H * dpad, ldvt=B * nk_pad, ldo=H * dpad, batch=B, heads=H,
                  nq=Nq, nk=Nk, nk_pad=nk_pad, dpad=dpad, scale=d ** -0.5, d_head=d)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().cuda().permute(0, 2, 1, 3) for t in (q, k, v))
    ref = torch.softmax(qf @ kf.transpose(-1, -2) * d ** -0.5, dim=-1) @ vf
    got = o[..., :d].permute(0, 2, 1, 3).float()
    return ((got - ref).norm() / ref.norm()).item()
for shp in [(2, 2, 512, 512, 32), (2, 2, 512, 512, 40), (1, 1, 256, 512, 32), (1, 1, 512, 512, 16), (2, 2, 512, 640, 32), (2, 2, 512, 384, 32),
            (2, 2, 512, 1024, 32), (2, 2, 1024, 1024, 32), (1, 8, 4096, 4096, 40), (2, 2, 960, 960, 32)]:
    print(shp, f"rel {run(*shp):.3e}")
