import sys, os, torch
sys.path.insert(0, "/root/repo")
from complex_prompt_diffusion_b200 import ops
ops.AUTOTUNE = False
for (M, N, K, v) in [(256, 160, 64, 160), (256, 160, 640, 160), (16384, 640, 640, 160), (16384, 640, 640, 2160), (16384, 640, 640, 224), (256,160,64,1)]:
    a = torch.randn(M, K, device="cuda").half(); w = torch.randn(N, K, device="cuda").half(); o = torch.empty(M, N, device="cuda", dtype=torch.float16)
    for _ in range(3):
        ops.gemm_conv(a, w, o, n_img=1, h=1, w=M, c0=K, n_out=N, variant=v)
    torch.cuda.synchronize()
