"""The plan-level C ABI (cpd_unet_plan_create / cpd_pack_weights / cpd_cache_context_kv / cpd_unet_forward)
driven through ctypes alone - the way a non-Python host binds it - against the oracle and against
the Python binding `UNetModel`."""
import ctypes as C
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _plan_from_ctypes(L, lib, cfg, sd, eps_dtype=0, use_graph=1):
    c = L.UNetConfig()
    c.in_channels = c.out_channels = 4
    c.model_channels, c.num_res_blocks = cfg.model_channels, cfg.num_res_blocks
    c.n_levels = len(cfg.channel_mult)
    td = cfg.transformer_depth
    for i, m in enumerate(cfg.channel_mult):
        c.channel_mult[i] = m
        c.transformer_depth[i] = td if isinstance(td, int) else list(td)[min(i, len(td) - 1)]
    c.n_attention_resolutions = len(cfg.attention_resolutions)
    for i, a in enumerate(cfg.attention_resolutions):
        c.attention_resolutions[i] = a
    c.num_heads, c.num_head_channels = cfg.num_heads, cfg.num_head_channels
    c.context_dim, c.use_linear_in_transformer, c.adm_in_channels = cfg.context_dim, int(cfg.use_linear_in_transformer), cfg.adm_in_channels
    c.act_fp16, c.eps_dtype, c.use_cuda_graph = 1, eps_dtype, use_graph
    plan = C.c_void_p()
    assert lib.cpd_unet_plan_create(C.byref(c), C.byref(plan)) == 0, lib.cpd_last_error()
    assert lib.cpd_unet_plan_missing_weights(plan) == len(sd)
    for k, (name, t) in enumerate(sd.items()):  # host pointers and device pointers, fp32 and bf16 sources
        t = t.contiguous()
        if k % 3 == 1:
            t = t.to(torch.bfloat16)
        if k % 2:
            t = t.to(DEV)
        assert lib.cpd_pack_weights(plan, name.encode(), C.c_void_p(t.data_ptr()), L.DTYPE_CODE[t.dtype], t.numel(), int(t.is_cuda)) == 0, \
            lib.cpd_last_error()
    assert lib.cpd_unet_plan_missing_weights(plan) == 0
    return plan


@pytest.mark.parametrize("cfg_name,hw,B,rpi", [("tiny", 16, 2, 3), ("tiny_xl", 32, 1, 2), ("sd15", 32, 1, 2)])
def test_plan_c_abi_vs_oracle_and_python_binding(cfg_name, hw, B, rpi):
    from complex_prompt_diffusion_b200 import _lib as L
    from complex_prompt_diffusion_b200.models.unet import UNetModel
    from oracle.unet import UNetConfig, OracleUNet, make_weights
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    lib = L.load()
    cfg = getattr(UNetConfig, cfg_name)()
    sd = make_weights(cfg, seed=0)
    plan = _plan_from_ctypes(L, lib, cfg, sd)
    g = torch.Generator().manual_seed(hw + B)
    R = B * rpi
    x = torch.randn(B, 4, hw, hw, generator=g)
    ctx = torch.randn(rpi, 77, cfg.context_dim, generator=g)
    y = torch.randn(rpi, cfg.adm_in_channels, generator=g) if cfg.adm_in_channels else None
    c_in, t = 0.3125, 500.0  # both exactly representable in bf16
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ctx_d = ctx.to(DEV)
    assert lib.cpd_cache_context_kv(plan, C.c_void_p(ctx_d.data_ptr()), L.CPD_F32, rpi, 77, stream) == 0, lib.cpd_last_error()
    if y is not None:
        y_d = y.to(DEV)
        assert lib.cpd_unet_set_vector(plan, C.c_void_p(y_d.data_ptr()), L.CPD_F32, rpi, stream) == 0, lib.cpd_last_error()
    x_d = x.to(DEV)
    sc = torch.tensor([c_in, t], device=DEV)
    eps = torch.full((R, 4, hw, hw), float("nan"), device=DEV)
    io = L.UNetIO()
    io.x, io.n_images, io.h, io.w, io.rows_per_image = x_d.data_ptr(), B, hw, hw, rpi
    io.c_in, io.t, io.t_count, io.eps = sc[0:1].data_ptr(), sc[1:2].data_ptr(), 1, eps.data_ptr()
    # 1. first call: eager warm-up + graph capture + replay; 2. pure replay; 3. eager (no_graph): all bit-identical
    outs = []
    for no_graph in (0, 0, 1):
        io.no_graph = no_graph
        eps.fill_(float("nan"))
        assert lib.cpd_unet_forward(plan, C.byref(io), stream) == 0, lib.cpd_last_error()
        torch.cuda.synchronize()
        assert lib.cpd_unet_plan_launches(plan) > 50
        outs.append(eps.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), "graph replay differs from the eager evaluation"
    # the oracle on bf16-rounded weights / inputs in fp32 arithmetic
    oracle = OracleUNet(cfg, {k: v.to(torch.bfloat16).float() for k, v in sd.items()})
    x_in = (x * c_in).repeat_interleave(rpi, dim=0)
    ctx_rows = ctx.to(torch.bfloat16).float().repeat(B, 1, 1)
    kw = {} if y is None else {"y": y.to(torch.bfloat16).float().repeat(B, 1)}
    ref = oracle(x_in.to(torch.bfloat16).float(), torch.full((R,), t), ctx_rows, **kw)
    r = rel(outs[0], ref)
    print(f"plan {cfg_name} {hw}x{hw} B{B} rows{rpi}: eps rel-L2 vs oracle {r:.3e}")
    assert r < 1e-2
    # the Python binding drives the same entry points: bit-identical
    gpu = UNetModel(sd, device=DEV, model_channels=cfg.model_channels, channel_mult=tuple(cfg.channel_mult),
                    attention_resolutions=tuple(cfg.attention_resolutions), num_res_blocks=cfg.num_res_blocks, num_heads=cfg.num_heads,
                    num_head_channels=cfg.num_head_channels, context_dim=cfg.context_dim,
                    use_linear_in_transformer=cfg.use_linear_in_transformer, transformer_depth=cfg.transformer_depth,
                    adm_in_channels=cfg.adm_in_channels, num_classes="sequential" if cfg.adm_in_channels else None)
    gpu.set_context(ctx_d)
    if y is not None:
        gpu.set_vector(y_d)
    got = gpu.forward_rows(x_d, c_in, t, rpi).clone()
    torch.cuda.synchronize()
    assert torch.equal(got, outs[0]), "UNetModel.forward_rows differs from the raw C ABI call"
    # a named buffer of the plan and a tap of the last forward
    ptr, n = C.c_void_p(), C.c_int64()
    assert lib.cpd_unet_plan_buffer(plan, b"input_blocks.0.0.out", C.byref(ptr), C.byref(n)) == 0 and n.value == R * hw * hw * cfg.model_channels
    ch, th, tw = C.c_int(), C.c_int(), C.c_int()
    assert lib.cpd_unet_plan_tap(plan, 0, 0, C.byref(ptr), C.byref(ch), C.byref(th), C.byref(tw)) == 0
    assert (ch.value, th.value, tw.value) == (cfg.model_channels, hw, hw)
    assert lib.cpd_unet_plan_buffer(plan, b"no.such.buffer", C.byref(ptr), C.byref(n)) != 0
    # errors: unknown parameter name, wrong size, forward without a usable shape
    assert lib.cpd_pack_weights(plan, b"not.a.parameter", C.c_void_p(x_d.data_ptr()), 0, 4, 1) != 0
    assert b"not a parameter" in lib.cpd_last_error()
    assert lib.cpd_pack_weights(plan, b"out.2.bias", C.c_void_p(x_d.data_ptr()), 0, 5, 1) != 0
    io.h = hw + 1
    assert lib.cpd_unet_forward(plan, C.byref(io), stream) != 0
    lib.cpd_unet_plan_destroy(plan)


def test_plan_per_row_timesteps_and_profile_records():
    """t_count = rows (the reference call signature: one timestep per row) and the per-launch profile records."""
    from complex_prompt_diffusion_b200 import _lib as L
    from oracle.unet import UNetConfig, OracleUNet, make_weights
    lib = L.load()
    cfg = UNetConfig.tiny()
    sd = make_weights(cfg, seed=0)
    plan = _plan_from_ctypes(L, lib, cfg, sd, use_graph=0)
    g = torch.Generator().manual_seed(3)
    n, hw = 3, 16
    x = torch.randn(n, 4, hw, hw, generator=g).to(torch.bfloat16).float()
    t = torch.tensor([937.93, 11.278, 500.5]).to(torch.bfloat16).float()
    ctx = torch.randn(n, 77, cfg.context_dim, generator=g).to(torch.bfloat16).float()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ctx_d, x_d, t_d = ctx.to(DEV), x.to(DEV), t.to(DEV)
    assert lib.cpd_cache_context_kv(plan, C.c_void_p(ctx_d.data_ptr()), L.CPD_F32, n, 77, stream) == 0, lib.cpd_last_error()
    eps = torch.empty(n, 4, hw, hw, device=DEV)
    io = L.UNetIO()
    io.x, io.n_images, io.h, io.w, io.rows_per_image = x_d.data_ptr(), n, hw, hw, 1
    io.t, io.t_count, io.eps = t_d.data_ptr(), n, eps.data_ptr()
    lib.cpd_unet_plan_set_profile(plan, 1)
    assert lib.cpd_unet_forward(plan, C.byref(io), stream) == 0, lib.cpd_last_error()
    need = lib.cpd_unet_plan_profile_dump(plan, None, 0)
    buf = C.create_string_buffer(int(need))
    lib.cpd_unet_plan_profile_dump(plan, buf, need)
    lib.cpd_unet_plan_set_profile(plan, 0)
    lines = buf.value.decode().splitlines()
    kinds = {ln.split("\t")[0] for ln in lines}
    assert {"gemm_conv", "attention", "groupnorm", "layernorm", "conv_in", "conv_out"} <= kinds and len(lines) > 50
    oracle = OracleUNet(cfg, {k: v.to(torch.bfloat16).float() for k, v in sd.items()})
    r = rel(eps, oracle(x, t, ctx))
    print(f"plan per-row timesteps: rel {r:.3e}")
    assert r < 1e-2
    lib.cpd_unet_plan_destroy(plan)
