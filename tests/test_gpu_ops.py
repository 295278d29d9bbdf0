"""GPU parity of every C-ABI kernel against a plain torch fp32 reference of the same op.
Run on a B200 with:  python -m pytest tests -m gpu"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from complex_prompt_diffusion_b200 import ops as o
    o.load()
    return o


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def bf(t):
    return t.to(torch.bfloat16)


DEV = "cuda"


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K,variant", [(128, 128, 64, 1), (256, 256, 128, 2), (384, 320, 320, 1), (1000, 1280, 640, 2),
                                           (77 * 4, 384, 768, 1), (4096, 640, 2560, 0), (128, 128, 64, 2),
                                           # persistent CTA-pair kernel: auto tile width, forced widths, ragged M / N,
                                           # more tiles than SM pairs (several tiles per cluster, both TMEM stages)
                                           (128, 128, 64, 0), (256, 256, 128, 0), (384, 320, 320, 0), (1000, 1280, 640, 0),
                                           (77 * 4, 384, 768, 0), (65536, 320, 320, 0), (16384, 640, 640, 160), (4096, 1280, 1280, 256),
                                           (8192, 768, 320, 64), (5000, 328, 192, 96), (20000, 640, 128, 224), (300, 2560, 320, 192),
                                           # (the two-pair multicast cluster, variant 1000 + BN, is an EXPERIMENTAL=1 build only)
                                           (70000, 320, 2880, 0), (40000, 2560, 320, 256),
                                           # wide tiles: two sub-tiles per tile, single TMEM accumulator stage (variant 2000 + BN)
                                           (65536, 320, 320, 2160), (16384, 640, 640, 2160), (4096, 1280, 1280, 2160), (30000, 512, 192, 2256),
                                           (1000, 328, 128, 2096), (50000, 1280, 256, 2192),
                                           # split-K (+ 10000 * S): partial sums reduced through the fp32 workspace
                                           (1024, 1280, 1280, 20160), (1000, 328, 640, 30128), (4096, 1280, 5120, 42160), (256, 128, 2048, 40064)])
def test_gemm_plain(ops, M, N, K, variant):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = bf(torch.randn(M, K, generator=g)).to(DEV)
    w = bf(torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = bf(torch.randn(M, N, generator=g)).to(DEV)
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.gemm_conv(a, w, out, n_img=1, h=1, w=M, c0=K, n_out=N, bias=bias, residual=res, ld_res=N, variant=variant)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias + res.float()
    r = rel(out, ref)
    print(f"gemm {M}x{N}x{K} v{variant}: rel {r:.3e}")
    assert torch.isfinite(out.float()).all()
    assert r < 4e-3


def test_gemm_no_epilogue_exactness(ops):
    # small integers: products/sums exactly representable -> tcgen05 result must be EXACT
    M, N, K = 256, 128, 192
    g = torch.Generator().manual_seed(5)
    a = torch.randint(-3, 4, (M, K), generator=g).float()
    w = torch.randint(-3, 4, (N, K), generator=g).float()
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    for variant in (1, 0, 64):
        out.fill_(float("nan"))
        ops.gemm_conv(bf(a).to(DEV), bf(w).to(DEV), out, n_img=1, h=1, w=M, c0=K, n_out=N, variant=variant)
        torch.cuda.synchronize()
        ref = a @ w.t()
        bad = (out.float().cpu() != bf(ref).float()).nonzero()
        print(f"variant {variant} mismatches:", bad.shape[0], bad[:10].tolist())
        assert bad.shape[0] == 0


@pytest.mark.parametrize("M,C,block,variant", [(512, 320, 128, 1), (512, 320, 128, 0), (512, 320, 256, 0), (40000, 320, 256, 0),
                                               (4096, 640, 256, 0)])
def test_gemm_geglu(ops, M, C, block, variant):
    g = torch.Generator().manual_seed(11)
    a = bf(torch.randn(M, C, generator=g)).to(DEV)
    w = bf(torch.randn(8 * C, C, generator=g) / math.sqrt(C))
    b = bf(torch.randn(8 * C, generator=g) * 0.1).float()
    inner4 = 4 * C
    hb = block // 2
    wp = torch.cat([w[:inner4].reshape(-1, hb, C), w[inner4:].reshape(-1, hb, C)], dim=1).reshape(8 * C, C).contiguous().to(DEV)
    bp = torch.cat([b[:inner4].reshape(-1, hb), b[inner4:].reshape(-1, hb)], dim=1).reshape(8 * C).contiguous().to(DEV)
    out = torch.empty(M, inner4, dtype=torch.bfloat16, device=DEV)
    ops.gemm_conv(a, wp, out, n_img=1, h=1, w=M, c0=C, n_out=8 * C, bias=bp, epilogue=ops.CPD_EPI_GEGLU, variant=variant,
                  geglu_block=block)
    torch.cuda.synchronize()
    y = a.float() @ w.float().to(DEV).t() + b.to(DEV)
    ref = y[:, :inner4] * F.gelu(y[:, inner4:])
    r = rel(out, ref)
    print(f"geglu M{M} C{C} block{block} v{variant} rel {r:.3e}")
    assert r < 6e-3


def _ln_fold_operands(g, K, N, act):
    """LayerNorm parameters and a projection, plus the folded forms the consumer GEMM takes (unet_plan.cu: ln_fold_kernel)."""
    gamma = bf(1.0 + 0.3 * torch.randn(K, generator=g)).float()
    beta = bf(0.2 * torch.randn(K, generator=g)).float()
    w = bf(torch.randn(N, K, generator=g) / math.sqrt(K)).float()
    b = bf(0.1 * torch.randn(N, generator=g)).float()
    wf = (w * gamma).to(act)
    return gamma, beta, w, b, wf, wf.float().sum(1), w @ beta + b


@pytest.mark.parametrize("M,K,N,pv,cv,act", [(4096, 320, 384, 0, 0, torch.float16), (65536, 320, 1152, 160, 192, torch.float16),
                                             (5000, 640, 1920, 2160, 128, torch.float16), (4096, 1280, 3840, 64, 256, torch.bfloat16),
                                             (16384, 640, 640, 2128, 2160, torch.float16), (304, 320, 320, 96, 0, torch.float16),
                                             (128, 320, 320, 160, 2160, torch.float16), (640, 320, 1152, 0, 160, torch.float16)])
def test_gemm_folded_layernorm_and_transposed_tail(ops, M, K, N, pv, cv, act):
    """x = a W0^T + b0 + res (producer, emits the rows' partial sums); then LN(x) W^T + b through the folded consumer:
    rstd * (x (gamma . W)^T - mean * g) + (W beta + b), its last third stored transposed (the fused Q | K | V projection of
    attention.py:476-487).  Reference: torch LayerNorm on the stored x."""
    g = torch.Generator().manual_seed(M + K + N)
    a = torch.randn(M, K, generator=g).to(act).to(DEV)
    w0 = (torch.randn(K, K, generator=g) / math.sqrt(K)).to(act).to(DEV)
    b0 = torch.randn(K, generator=g).to(DEV)
    res = (2.0 * torch.randn(M, K, generator=g) + 0.7).to(act).to(DEV)  # non-zero row means
    x = torch.full((M, K), float("nan"), dtype=act, device=DEV)
    max_parts = 2 * ((K + 63) // 64)
    sums = torch.full((max_parts, M + 8, 2), float("nan"), device=DEV)
    _, parts = ops.gemm_conv(a, w0, x, n_img=1, h=1, w=M, c0=K, n_out=K, bias=b0, residual=res, ld_res=K, variant=pv, ln_sums_out=sums)
    torch.cuda.synchronize()
    assert 0 < parts <= max_parts
    xf = x.float()
    s = sums[:parts, :M].sum(0)
    assert rel(s[:, 0], xf.sum(1)) < 2e-3 and rel(s[:, 1], (xf * xf).sum(1)) < 2e-3
    gamma, beta, w, b, wf, gvec, bfold = _ln_fold_operands(g, K, N, act)
    ref = F.layer_norm(xf, (K,), gamma.to(DEV), beta.to(DEV), 1e-5) @ w.to(DEV).t() + b.to(DEV)
    n_t = N // 3 // 32 * 32
    out = torch.full((M, N - n_t), float("nan"), dtype=act, device=DEV)
    out_t = torch.full((n_t, (M + 15) // 8 * 8), float("nan"), dtype=act, device=DEV)
    ops.gemm_conv(x, wf.to(DEV), out, n_img=1, h=1, w=M, c0=K, n_out=N, ldd=N - n_t, bias=bfold.to(DEV), variant=cv, ln_sums=sums,
                  ln_parts=parts, ln_g=gvec.to(DEV), d_t=out_t, dt_col0=N - n_t)
    torch.cuda.synchronize()
    r0, r1 = rel(out, ref[:, :N - n_t]), rel(out_t[:, :M].t(), ref[:, N - n_t:])
    print(f"folded LN {M}x{N}x{K} producer v{pv} ({parts} parts) consumer v{cv}: rel {r0:.3e}, transposed tail {r1:.3e}")
    assert torch.isnan(out_t[:, M:].float()).all()  # nothing written beyond the valid rows
    tol = 2e-2 if act == torch.bfloat16 else 4e-3
    assert r0 < tol and r1 < tol


@pytest.mark.parametrize("n,h,w,cin,cout,ks,res,variant,act", [(16, 64, 64, 320, 320, 3, True, 0, torch.float16), (3, 32, 32, 64, 640, 3, False, 160, torch.float16),
                                                              (5, 16, 16, 128, 1280, 3, True, 256, torch.bfloat16), (2, 8, 8, 64, 128, 3, False, 0, torch.float16),
                                                              (4, 32, 32, 640, 640, 1, True, 2160, torch.float16), (3, 16, 24, 64, 96, 3, False, 96, torch.float16)])
def test_gemm_groupnorm_statistics_and_apply(ops, n, h, w, cin, cout, ks, res, variant, act):
    """A conv / GEMM emits the GroupNorm statistics of its output (fixed-point integer atomics); cpd_groupnorm_apply then
    normalises in one pass.  Reference: F.group_norm (+ SiLU) on the stored output (models/util.py:95-105)."""
    g = torch.Generator().manual_seed(n * h + cin + cout)
    x = torch.randn(n, h, w, cin, generator=g).to(act).to(DEV)
    wt = (torch.randn(cout, ks * ks * cin, generator=g) / math.sqrt(ks * ks * cin)).to(act).to(DEV)
    bias = (torch.randn(cout, generator=g) + 0.5).to(DEV)
    r = (torch.randn(n * h * w, cout, generator=g)).to(act).to(DEV) if res else None
    out = torch.empty(n * h * w, cout, dtype=act, device=DEV)
    sums = torch.zeros(n, cout, 2, dtype=torch.int64, device=DEV)
    ops.gemm_conv(x, wt, out, n_img=n, h=h, w=w, c0=cin, n_out=cout, ksize=ks, bias=bias, residual=r, ld_res=cout if res else 0,
                  variant=variant, gn_sums_out=sums)
    if variant == 0:  # the tuner re-zeroes the accumulators in front of the real launch; a second (table) launch adds again
        sums.zero_()
        ops.gemm_conv(x, wt, out, n_img=n, h=h, w=w, c0=cin, n_out=cout, ksize=ks, bias=bias, residual=r, ld_res=cout if res else 0,
                      variant=variant, gn_sums_out=sums)
    torch.cuda.synchronize()
    of = out.float().reshape(n, h * w, cout)
    s1 = sums[:, :, 0].double() / 2 ** 24
    s2 = sums[:, :, 1].double() / 2 ** 12
    e1 = (s1 - of.double().sum(1)).abs().max().item() / max(1.0, of.double().sum(1).abs().max().item())
    e2 = (s2 - (of.double() ** 2).sum(1)).abs().max().item() / (of.double() ** 2).sum(1).max().item()
    print(f"GN statistics {n}x{h}x{w} {cin}->{cout} k{ks} v{variant}: sum err {e1:.2e}, sum-of-squares err {e2:.2e}")
    assert e1 < 2e-3 and e2 < 2e-3  # (the sums are of the fp32 values in front of the 16-bit rounding)
    gamma = (1.0 + 0.2 * torch.randn(cout, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(cout, generator=g)).to(DEV)
    y = torch.empty_like(out)
    ops.groupnorm_apply(out, gamma, beta, y, sums, n_img=n, hw=h * w, c=cout, eps=1e-5, silu=True)
    torch.cuda.synchronize()
    ref = F.silu(F.group_norm(of.permute(0, 2, 1), 32, gamma, beta, 1e-5)).permute(0, 2, 1).reshape(n * h * w, cout)
    rr = rel(y, ref)
    print(f"   groupnorm_apply rel {rr:.3e}")
    assert rr < (2e-2 if act == torch.bfloat16 else 4e-3)


@pytest.mark.parametrize("M,C", [(4096, 320), (16384, 640)])
def test_gemm_geglu_folded_layernorm(ops, M, C):
    g = torch.Generator().manual_seed(M + C)
    act = torch.float16
    x = (1.5 * torch.randn(M, C, generator=g) - 0.4).to(act).to(DEV)
    xf = x.float()
    parts = 3
    sums = torch.zeros(parts, M, 2, device=DEV)  # the producer's layout: any split of the row sums over the parts
    cuts = [0, 96, 224, C]
    for i in range(parts):
        sums[i, :, 0] = xf[:, cuts[i]:cuts[i + 1]].sum(1)
        sums[i, :, 1] = (xf[:, cuts[i]:cuts[i + 1]] ** 2).sum(1)
    gamma, beta, w, b, wf, gvec, bfold = _ln_fold_operands(g, C, 8 * C, act)
    inner4, hb = 4 * C, 128

    def inter(t):  # [128 value rows | 128 gate rows] per 256-column tile
        return torch.cat([t[:inner4].reshape(-1, hb, *t.shape[1:]), t[inner4:].reshape(-1, hb, *t.shape[1:])], dim=1).reshape(t.shape).contiguous()
    out = torch.empty(M, inner4, dtype=act, device=DEV)
    ops.gemm_conv(x, inter(wf).to(DEV), out, n_img=1, h=1, w=M, c0=C, n_out=8 * C, bias=inter(bfold).to(DEV), epilogue=ops.CPD_EPI_GEGLU,
                  geglu_block=256, ln_sums=sums, ln_parts=parts, ln_g=inter(gvec).to(DEV))
    torch.cuda.synchronize()
    y = F.layer_norm(xf, (C,), gamma.to(DEV), beta.to(DEV), 1e-5) @ w.to(DEV).t() + b.to(DEV)
    ref = y[:, :inner4] * F.gelu(y[:, inner4:])
    r = rel(out, ref)
    print(f"geglu + folded LN M{M} C{C}: rel {r:.3e}")
    assert r < 4e-3


@pytest.mark.parametrize("n,h,w,c0,c1,cout,stride,variant", [
    (2, 16, 16, 64, 0, 128, 1, 1), (1, 64, 64, 320, 0, 320, 1, 0), (3, 8, 8, 128, 64, 256, 1, 2), (2, 32, 32, 64, 0, 64, 2, 1),
    (4, 8, 8, 1280, 1280, 1280, 1, 0), (2, 16, 16, 128, 0, 128, 2, 2), (1, 24, 24, 64, 0, 64, 1, 1), (2, 12, 12, 64, 64, 128, 1, 1),
    # CTA-pair kernel: every level of the SD UNet (64^2 .. 8^2), stride 2, skip concat, tiny 2x2 / 4x4 images (boxes
    # spanning several images), odd image counts, forced tile widths
    (2, 16, 16, 64, 0, 128, 1, 0), (3, 8, 8, 128, 64, 256, 1, 0), (2, 32, 32, 64, 0, 64, 2, 0), (2, 16, 16, 128, 0, 128, 2, 0),
    (1, 24, 24, 64, 0, 64, 1, 0), (2, 12, 12, 64, 64, 128, 1, 0), (16, 64, 64, 320, 0, 320, 1, 0), (16, 32, 32, 640, 320, 640, 1, 160),
    (16, 16, 16, 1280, 0, 1280, 1, 256), (16, 8, 8, 1280, 1280, 1280, 1, 128), (5, 2, 2, 128, 0, 128, 1, 0), (3, 4, 4, 64, 64, 192, 1, 0),
    (6, 4, 4, 128, 0, 64, 2, 0), (16, 64, 64, 320, 0, 320, 2, 0), (7, 16, 16, 64, 0, 320, 1, 96),
    # (the two-pair multicast cluster, variant 1000 + BN, is an EXPERIMENTAL=1 build only); stride 2 + concat at forced widths instead
    (6, 4, 4, 128, 0, 128, 1, 64), (3, 32, 32, 64, 64, 256, 1, 128), (5, 16, 16, 128, 0, 512, 1, 192),
    # wide 256 x 320 tiles (two sub-tiles, one accumulator stage)
    (16, 64, 64, 320, 0, 320, 1, 2160), (16, 32, 32, 640, 320, 640, 1, 2160), (16, 16, 16, 1280, 0, 1280, 1, 2160),
    (16, 8, 8, 1280, 1280, 1280, 1, 2160), (16, 64, 64, 320, 0, 320, 2, 2160), (3, 32, 32, 64, 64, 256, 1, 2128),
    # split-K on the small-M levels (8x8 / 4x4 images)
    (16, 8, 8, 1280, 0, 1280, 1, 20160), (16, 8, 8, 1280, 1280, 1280, 1, 42160), (5, 4, 4, 128, 64, 192, 1, 30096), (16, 16, 16, 640, 0, 1280, 2, 20160)])
def test_conv3x3(ops, n, h, w, c0, c1, cout, stride, variant):
    g = torch.Generator().manual_seed(n * h + cout)
    cin = c0 + c1
    x = bf(torch.randn(n, cin, h, w, generator=g))
    wt = bf(torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin))
    bias = torch.randn(cout, generator=g)
    rowvec = torch.randn(n, cout, generator=g)
    ho, wo = h // stride, w // stride
    res = bf(torch.randn(n, cout, ho, wo, generator=g))
    nhwc = x.permute(0, 2, 3, 1).contiguous()
    a0 = nhwc[..., :c0].contiguous().to(DEV)
    a1 = nhwc[..., c0:].contiguous().to(DEV) if c1 else None
    wp = wt.permute(0, 2, 3, 1).contiguous().to(DEV)
    out = torch.full((n, ho, wo, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.gemm_conv(a0, wp, out, n_img=n, h=h, w=w, c0=c0, a1=a1, c1=c1, n_out=cout, ksize=3, stride=stride,
                  bias=bias.to(DEV), rowvec=rowvec.to(DEV).contiguous(), rowvec_stride=cout,
                  residual=res.permute(0, 2, 3, 1).contiguous().to(DEV), ld_res=cout, variant=variant)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().to(DEV), wt.float().to(DEV), bias.to(DEV), stride=stride, padding=1)
    ref = ref + rowvec.to(DEV)[:, :, None, None] + res.float().to(DEV)
    r = rel(out.permute(0, 3, 1, 2), ref)
    print(f"conv3x3 n{n} {h}x{w} {c0}+{c1}->{cout} s{stride} v{variant}: rel {r:.3e}")
    assert torch.isfinite(out.float()).all()
    assert r < 4e-3


def test_conv1x1_two_sources(ops):
    n, h, w, c0, c1, cout = 2, 16, 16, 128, 64, 320
    g = torch.Generator().manual_seed(3)
    x = bf(torch.randn(n, h, w, c0 + c1, generator=g))
    wt = bf(torch.randn(cout, c0 + c1, generator=g) / math.sqrt(c0 + c1)).to(DEV)
    out = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device=DEV)
    ops.gemm_conv(x[..., :c0].contiguous().to(DEV), wt, out, n_img=n, h=h, w=w, c0=c0, a1=x[..., c0:].contiguous().to(DEV), c1=c1,
                  n_out=cout, ksize=1)
    torch.cuda.synchronize()
    ref = x.float().to(DEV) @ wt.float().t()
    assert rel(out, ref) < 4e-3


# ------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("B,H,Nq,Nk,d", [(2, 8, 256, 256, 40), (1, 4, 1024, 1024, 80), (2, 2, 64, 64, 160), (3, 8, 256, 77, 40),
                                         (2, 5, 576, 576, 64), (1, 2, 64, 64, 32), (4, 8, 4096, 77, 40), (1, 8, 4096, 4096, 40),
                                         (2, 3, 320, 200, 40), (1, 2, 256, 1000, 96),
                                         # split-row kernel (attention_umma4.cu: d <= 63, more than two key blocks), ragged shapes
                                         (2, 3, 300, 700, 40), (1, 2, 1000, 330, 48), (1, 1, 129, 257, 16), (2, 2, 512, 384, 63)])
@pytest.mark.parametrize("two_tile", [False, True])
def test_attention(ops, B, H, Nq, Nk, d, two_tile):
    """two_tile=True passes d_head, which opens the shape dispatch of cpd_attention (one key block -> attention_umma5.cu,
    d <= 63 with many key blocks -> attention_umma4.cu, otherwise the persistent two-tile kernel attention_umma3.cu);
    d_head = 0 and shapes outside those domains run the one-tile kernel."""
    g = torch.Generator().manual_seed(B * Nq + d)
    dpad = (d + 15) // 16 * 16
    nk_pad = Nk if Nk % 16 == 0 else (Nk + 15) // 16 * 16
    q = bf(torch.randn(B, Nq, H, d, generator=g))
    k = bf(torch.randn(B, Nk, H, d, generator=g))
    v = bf(torch.randn(B, Nk, H, d, generator=g))
    qp = torch.zeros(B, Nq, H, dpad, dtype=torch.bfloat16); qp[..., :d] = q
    kp = torch.zeros(B, nk_pad, H, dpad, dtype=torch.bfloat16); kp[:, :Nk, :, :d] = k
    vp = torch.zeros(B, nk_pad, H, dpad, dtype=torch.bfloat16); vp[:, :Nk, :, :d] = v
    vt = vp.permute(2, 3, 0, 1).reshape(H * dpad, B * nk_pad).contiguous().to(DEV)
    o = torch.full((B, Nq, H, dpad), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.attention(qp.to(DEV), kp.to(DEV), vt, o, ldq=H * dpad, ldk=H * dpad, ldvt=B * nk_pad, ldo=H * dpad, batch=B, heads=H,
                  nq=Nq, nk=Nk, nk_pad=nk_pad, dpad=dpad, scale=d ** -0.5, d_head=d if two_tile else 0)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().to(DEV).permute(0, 2, 1, 3) for t in (q, k, v))
    ref = torch.softmax(qf @ kf.transpose(-1, -2) * d ** -0.5, dim=-1) @ vf  # [B,H,Nq,d]
    got = o[..., :d].permute(0, 2, 1, 3)
    r = rel(got, ref)
    print(f"attention B{B} H{H} {Nq}x{Nk} d{d} two_tile={two_tile}: rel {r:.3e}")
    assert torch.isfinite(o.float()).all()
    assert r < 1e-2
    if dpad != d:
        assert (o[..., d:] == 0).all()


@pytest.mark.parametrize("B,kvb,H,Nq,d,dtype", [(4, 2, 8, 512, 40, torch.float16), (8, 4, 5, 1024, 64, torch.float16), (4, 4, 8, 1024, 80, torch.bfloat16),
                                                (6, 3, 8, 384, 40, torch.float16), (2, 1, 10, 256, 64, torch.bfloat16), (4, 2, 8, 320, 40, torch.float16),
                                                (16, 4, 8, 4096, 40, torch.float16)])
def test_cross_attention_shared_context(ops, B, kvb, H, Nq, d, dtype):
    """The cross-attention kernel (attention_umma5.cu): 77 context tokens, K / V^T of `kvb` context rows shared by the B query
    batches (query batch b reads context row b % kvb, denoiser.py:385), fp16 and bf16 activations, the two-tile (d = 40) and
    one-tile (d = 64 with 128-byte swizzled staging rows, d = 80) instantiations, and a shape that falls back (nq = 320)."""
    g = torch.Generator().manual_seed(B * Nq + d + kvb)
    Nk, nk_pad = 77, 80
    dpad = (d + 15) // 16 * 16
    q = torch.randn(B, Nq, H, d, generator=g).to(dtype)
    k = torch.randn(kvb, Nk, H, d, generator=g).to(dtype)
    v = torch.randn(kvb, Nk, H, d, generator=g).to(dtype)
    qp = torch.zeros(B, Nq, H, dpad, dtype=dtype); qp[..., :d] = q
    kp = torch.zeros(kvb, nk_pad, H, dpad, dtype=dtype); kp[:, :Nk, :, :d] = k
    vp = torch.zeros(kvb, nk_pad, H, dpad, dtype=dtype); vp[:, :Nk, :, :d] = v
    vt = vp.permute(2, 3, 0, 1).reshape(H * dpad, kvb * nk_pad).contiguous().to(DEV)
    o = torch.full((B, Nq, H, dpad), float("nan"), dtype=dtype, device=DEV)
    ops.attention(qp.to(DEV), kp.to(DEV), vt, o, ldq=H * dpad, ldk=H * dpad, ldvt=kvb * nk_pad, ldo=H * dpad, batch=B, heads=H,
                  nq=Nq, nk=Nk, nk_pad=nk_pad, dpad=dpad, scale=d ** -0.5, d_head=d, kv_batch=kvb)
    torch.cuda.synchronize()
    idx = torch.arange(B) % kvb
    qf = q.float().to(DEV).permute(0, 2, 1, 3)
    kf, vf = k.float().to(DEV)[idx].permute(0, 2, 1, 3), v.float().to(DEV)[idx].permute(0, 2, 1, 3)
    ref = torch.softmax(qf @ kf.transpose(-1, -2) * d ** -0.5, dim=-1) @ vf
    r = rel(o[..., :d].permute(0, 2, 1, 3), ref)
    print(f"cross-attention B{B} kv{kvb} H{H} {Nq}x77 d{d} {dtype}: rel {r:.3e}")
    assert torch.isfinite(o.float()).all()
    assert r < (4e-3 if dtype == torch.float16 else 1e-2)
    if dpad != d:
        assert (o[..., d:] == 0).all()


# ------------------------------------------------------------------------------------------------ norms / small ops
@pytest.mark.parametrize("n,hw,c0,c1,silu,eps", [(2, 4096, 320, 0, True, 1e-5), (3, 64, 1280, 1280, True, 1e-5), (1, 1024, 640, 320, False, 1e-6),
                                                  (2, 256, 64, 0, True, 1e-5), (16, 4096, 320, 320, True, 1e-5),
                                                  # shared-memory slab kernel: cluster sizes 1 / 2 / 4 / 8, ragged pixel ranges, slabs that
                                                  # straddle the two inputs, 4 / 16 / 30 / 60 channels per group, the one-block-per-SM size
                                                  (2, 9216, 320, 0, True, 1e-5), (2, 576, 640, 640, True, 1e-5), (4, 144, 1280, 0, False, 1e-5),
                                                  (1, 1024, 1280, 640, True, 1e-5), (2, 4096, 640, 320, True, 1e-5), (2, 256, 128, 0, True, 1e-6),
                                                  (1, 100, 512, 0, False, 1e-6), (3, 7, 320, 0, True, 1e-5), (2, 2304, 320, 0, True, 1e-5)])
def test_groupnorm(ops, n, hw, c0, c1, silu, eps):
    g = torch.Generator().manual_seed(hw + c0)
    C = c0 + c1
    x = bf(torch.randn(n, hw, C, generator=g) * 2 + 0.5)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    out = torch.empty(n, hw, C, dtype=torch.bfloat16, device=DEV)
    stats = torch.empty(n * 64 * ops.GN_MAX_CHUNKS, dtype=torch.float64, device=DEV)
    ops.groupnorm(x[..., :c0].contiguous().to(DEV), gamma.to(DEV), beta.to(DEV), out, stats, n_img=n, hw=hw, c0=c0,
                  a1=x[..., c0:].contiguous().to(DEV) if c1 else None, c1=c1, eps=eps, silu=silu)
    torch.cuda.synchronize()
    ref = F.group_norm(x.float().to(DEV).permute(0, 2, 1), 32, gamma.to(DEV), beta.to(DEV), eps)
    if silu:
        ref = F.silu(ref)
    r = rel(out.permute(0, 2, 1), ref)
    print(f"groupnorm n{n} hw{hw} C{C}: rel {r:.3e}")
    assert r < 4e-3


@pytest.mark.parametrize("rows,c", [(4096, 320), (1000, 640), (64, 1280), (5, 64), (300, 2048)])
def test_layernorm(ops, rows, c):
    g = torch.Generator().manual_seed(rows + c)
    x = bf(torch.randn(rows, c, generator=g) * 3 + 1).to(DEV)
    gamma, beta = torch.randn(c, generator=g).to(DEV), torch.randn(c, generator=g).to(DEV)
    out = torch.empty_like(x)
    ops.layernorm(x, gamma, beta, out, rows=rows, c=c)
    torch.cuda.synchronize()
    ref = F.layer_norm(x.float(), (c,), gamma, beta, 1e-5)
    assert rel(out, ref) < 4e-3


def test_timestep_embedding_and_small_linear(ops):
    t = torch.tensor([937.93, 11.278, 500.0], device=DEV)
    out = torch.empty(3, 320, dtype=torch.bfloat16, device=DEV)
    ops.timestep_embedding(t, out, dim=320, round_t_bf16=True)
    tt = t.to(torch.bfloat16).float()
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=DEV) / half)
    args = tt[:, None] * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    assert (out.float() - ref).abs().max().item() < 1e-2
    g = torch.Generator().manual_seed(0)
    for m, k, n, silu in [(3, 320, 1280, False), (16, 1280, 4000, True), (1, 1280, 1280, True), (20, 64, 256, True)]:
        x = bf(torch.randn(m, k, generator=g)).to(DEV)
        w = bf(torch.randn(n, k, generator=g) / math.sqrt(k)).to(DEV)
        b = torch.randn(n, generator=g).to(DEV)
        o32 = torch.empty(m, n, device=DEV)
        o16 = torch.empty(m, n, dtype=torch.bfloat16, device=DEV)
        ops.small_linear(x, w, b, m=m, k=k, n=n, silu_in=silu, out_f32=o32, out_bf16=o16)
        xin = F.silu(x.float()) if silu else x.float()
        ref = xin @ w.float().t() + b
        assert rel(o32, ref) < 6e-3 and rel(o16, ref) < 6e-3


def test_conv_in_out_upsample(ops):
    g = torch.Generator().manual_seed(1)
    n, h, w, cout = 2, 16, 16, 64
    x = torch.randn(n, 4, h, w, generator=g)
    wt = bf(torch.randn(cout, 4, 3, 3, generator=g) / 6)
    bias = torch.randn(cout, generator=g)
    out = torch.empty(n * 3, h, w, cout, dtype=torch.bfloat16, device=DEV)
    ops.conv_in(x.to(DEV), wt.permute(0, 2, 3, 1).contiguous().to(DEV), bias.to(DEV), out, n=n, cin=4, h=h, w=w, cout=cout,
                scale=0.37, rows_per_image=3)
    ref = F.conv2d(bf(x * 0.37).float().to(DEV), wt.float().to(DEV), bias.to(DEV), padding=1).permute(0, 2, 3, 1)
    for r in range(3):
        assert rel(out.view(n, 3, h, w, cout)[:, r], ref) < 4e-3
    cin = 320
    a = bf(torch.randn(n, h, w, cin, generator=g)).to(DEV)
    w2 = bf(torch.randn(4, cin, 3, 3, generator=g) / math.sqrt(9 * cin))
    b2 = torch.randn(4, generator=g)
    o2 = torch.empty(n, 4, h, w, dtype=torch.bfloat16, device=DEV)
    ops.conv_out(a, w2.permute(0, 2, 3, 1).contiguous().to(DEV), b2.to(DEV), o2, n=n, h=h, w=w, cin=cin, cout=4)
    ref2 = F.conv2d(a.float().permute(0, 3, 1, 2), w2.float().to(DEV), b2.to(DEV), padding=1)
    assert rel(o2, ref2) < 4e-3
    up = torch.empty(n, 2 * h, 2 * w, cin, dtype=torch.bfloat16, device=DEV)
    ops.upsample2x(a, up, n=n, h=h, w=w, c=cin)
    ref3 = F.interpolate(a.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), ref3)


@pytest.mark.parametrize("n,h,w,cout,rpi,act", [(4, 64, 64, 320, 4, torch.float16), (2, 8, 12, 64, 3, torch.bfloat16), (1, 6, 6, 64, 2, torch.float16),
                                                 (1, 128, 128, 320, 2, torch.bfloat16), (3, 5, 4, 128, 1, torch.float16)])
def test_conv_in_vs_torch(ops, n, h, w, cout, rpi, act):
    """cpd_conv_in (unet.py:548 on x * c_in, denoiser.py:390-391): the four-pixels-per-thread kernel (w % 4 == 0) and the per-pixel
    one (other widths), every conditioning-row copy, against torch's conv2d on the same rounded operands."""
    g = torch.Generator().manual_seed(h * w + cout)
    x = torch.randn(n, 4, h, w, generator=g)
    wt = bf(torch.randn(cout, 4, 3, 3, generator=g) / 6)
    bias = torch.randn(cout, generator=g)
    out = torch.full((n * rpi, h, w, cout), float("nan"), dtype=act, device=DEV)
    ops.conv_in(x.to(DEV), wt.permute(0, 2, 3, 1).contiguous().to(DEV), bias.to(DEV), out, n=n, cin=4, h=h, w=w, cout=cout, scale=0.37,
                rows_per_image=rpi)
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d((x * 0.37).to(act).float().to(DEV), wt.float().to(DEV), bias.to(DEV), padding=1).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    o5 = out.view(n, rpi, h, w, cout)
    assert torch.isfinite(o5.float()).all()
    for r in range(rpi):
        assert rel(o5[:, r], ref) < (4e-3 if act == torch.bfloat16 else 5e-4)
        assert torch.equal(o5[:, r], o5[:, 0])


@pytest.mark.parametrize("n,h,w,cin,act,odt", [(3, 64, 64, 320, torch.float16, torch.float32), (2, 24, 40, 320, torch.bfloat16, torch.float32),
                                                 (1, 96, 96, 320, torch.float16, torch.bfloat16), (2, 8, 8, 128, torch.bfloat16, torch.float32),
                                                 (1, 70, 130, 64, torch.float16, torch.float32), (1, 128, 128, 320, torch.float16, torch.float32)])
def test_conv_out_tiled_vs_torch(ops, n, h, w, cin, act, odt):
    """cpd_conv_out (unet.py:729-733, NHWC -> NCHW, cout 4): the tiled kernel (strips of 64 pixels, transposed staging,
    broadcast fp32 weights) on full / partial / multiple strips per row, both activation formats and output dtypes, against
    torch's conv2d on the same rounded operands - and against the one-warp-per-pixel kernel it replaced."""
    g = torch.Generator().manual_seed(h * w + cin)
    a = torch.randn(n, h, w, cin, generator=g).to(act).to(DEV)
    wt = bf(torch.randn(4, cin, 3, 3, generator=g) / math.sqrt(9 * cin))
    b = torch.randn(4, generator=g)
    out = torch.full((n, 4, h, w), float("nan"), dtype=odt, device=DEV)
    ops.conv_out(a, wt.permute(0, 2, 3, 1).contiguous().to(DEV), b.to(DEV), out, n=n, h=h, w=w, cin=cin, cout=4)
    ref = F.conv2d(a.float().permute(0, 3, 1, 2), wt.float().to(DEV), b.to(DEV), padding=1)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert rel(out, ref) < (4e-3 if odt == torch.bfloat16 else 2e-5)


# ------------------------------------------------------------------------------------------------ fp16 activations x bf16 weights
def test_gemm_rejects_mixed_operand_formats(ops):
    """tcgen05 kind::f16 traps (illegal instruction) when A is fp16 and B is bf16: the C ABI refuses it up front."""
    a = torch.zeros(128, 64, dtype=torch.float16, device=DEV)
    w = torch.zeros(128, 64, dtype=torch.bfloat16, device=DEV)
    out = torch.zeros(128, 128, dtype=torch.float16, device=DEV)
    with pytest.raises(RuntimeError, match="same 16-bit format"):
        ops.gemm_conv(a, w, out, n_img=1, h=1, w=128, c0=64, n_out=128)


@pytest.mark.parametrize("a_dt,b_dt,o_dt", [(torch.float16, torch.float16, torch.float16), (torch.bfloat16, torch.bfloat16, torch.float16),
                                            (torch.float16, torch.float16, torch.bfloat16)])
def test_gemm_operand_and_output_formats(ops, a_dt, b_dt, o_dt):
    """fp16 or bf16 operands (same format for A and B), independent 16-bit output format."""
    M, N, K = 512, 384, 320
    g = torch.Generator().manual_seed(9)
    a = torch.randn(M, K, generator=g).to(a_dt).to(DEV)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(b_dt).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(o_dt).to(DEV)
    out = torch.empty(M, N, dtype=o_dt, device=DEV)
    ops.gemm_conv(a, w, out, n_img=1, h=1, w=M, c0=K, n_out=N, bias=bias, residual=res, ld_res=N)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias + res.float()
    r = rel(out, ref)
    print(f"gemm a={a_dt} b={b_dt} out={o_dt}: rel {r:.3e}")
    assert r < (6e-4 if o_dt == torch.float16 else 4e-3)


def test_conv_attention_norms_fp16_activations(ops):
    g = torch.Generator().manual_seed(21)
    n, h, w, cin, cout = 2, 16, 16, 128, 128
    x = torch.randn(n, cin, h, w, generator=g).half()
    wt = bf(torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)).half()  # bf16 values, stored as fp16 (exact)
    out = torch.empty(n, h, w, cout, dtype=torch.float16, device=DEV)
    ops.gemm_conv(x.permute(0, 2, 3, 1).contiguous().to(DEV), wt.permute(0, 2, 3, 1).contiguous().to(DEV), out, n_img=n, h=h, w=w,
                  c0=cin, n_out=cout, ksize=3)
    ref = F.conv2d(x.float().to(DEV), wt.float().to(DEV), padding=1)
    assert rel(out.permute(0, 3, 1, 2), ref) < 6e-4
    # attention
    B, H, Nq, d = 2, 8, 256, 40
    dpad = 48
    q, k, v = (torch.randn(B, Nq, H, d, generator=g).half() for _ in range(3))
    qp, kp, vp = (torch.zeros(B, Nq, H, dpad, dtype=torch.float16) for _ in range(3))
    qp[..., :d], kp[..., :d], vp[..., :d] = q, k, v
    vt = vp.permute(2, 3, 0, 1).reshape(H * dpad, B * Nq).contiguous().to(DEV)
    o = torch.empty(B, Nq, H, dpad, dtype=torch.float16, device=DEV)
    ops.attention(qp.to(DEV), kp.to(DEV), vt, o, ldq=H * dpad, ldk=H * dpad, ldvt=B * Nq, ldo=H * dpad, batch=B, heads=H, nq=Nq,
                  nk=Nq, nk_pad=Nq, dpad=dpad, scale=d ** -0.5)
    qf, kf, vf = (t.float().to(DEV).permute(0, 2, 1, 3) for t in (q, k, v))
    ref = torch.softmax(qf @ kf.transpose(-1, -2) * d ** -0.5, dim=-1) @ vf
    r = rel(o[..., :d].permute(0, 2, 1, 3), ref)
    print(f"attention fp16: rel {r:.3e}")
    assert r < 1.5e-3
    # norms
    C, hw = 320, 1024
    xx = (torch.randn(n, hw, C, generator=g) * 2 + 0.5).half().to(DEV)
    gamma, beta = torch.randn(C, generator=g).to(DEV), torch.randn(C, generator=g).to(DEV)
    og = torch.empty_like(xx)
    stats = torch.empty(n * 64 * ops.GN_MAX_CHUNKS, dtype=torch.float64, device=DEV)
    ops.groupnorm(xx, gamma, beta, og, stats, n_img=n, hw=hw, c0=C, silu=True)
    refg = F.silu(F.group_norm(xx.float().permute(0, 2, 1), 32, gamma, beta, 1e-5))
    assert rel(og.permute(0, 2, 1), refg) < 6e-4
    ol = torch.empty_like(xx)
    ops.layernorm(xx.view(-1, C), gamma, beta, ol.view(-1, C), rows=n * hw, c=C)
    assert rel(ol, F.layer_norm(xx.float(), (C,), gamma, beta, 1e-5)) < 6e-4
