"""CPU-side tests: host logic of the drop-in package and the C ABI surface (no compute calls without a GPU)."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    path = os.path.join(ROOT, "complex_prompt_diffusion_b200", "libcpd_b200.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", os.path.join(ROOT, "complex_prompt_diffusion_b200", "csrc"), "-j8"], check=True)
    return path


def test_c_abi_exports_every_declared_symbol(lib_path):
    header = open(os.path.join(ROOT, "include", "cpd_b200.h")).read()
    declared = set(re.findall(r"\b(cpd_[a-z0-9_]+)\s*\(", header))
    assert {"cpd_sampler_step", "cpd_gemm_conv", "cpd_attention", "cpd_groupnorm", "cpd_last_error"} <= declared
    lib = ctypes.CDLL(lib_path)
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/cpd_b200.h but not exported"
    from complex_prompt_diffusion_b200 import _lib
    assert set(_lib.EXPORTED_SYMBOLS) == declared
    assert _lib.load().cpd_abi_version() == 1


def test_argument_validation_without_gpu(lib_path):
    """Invalid arguments are rejected with a status code and a message before any CUDA call."""
    from complex_prompt_diffusion_b200 import _lib
    lib = _lib.load()
    p = _lib.StepParams()
    assert lib.cpd_sampler_step(ctypes.byref(p), None) != 0
    assert b"eps and x" in lib.cpd_last_error()
    gp = _lib.GemmParams()
    gp.a0 = gp.wt = gp.d = 16
    gp.c0 = 100
    assert lib.cpd_gemm_conv(ctypes.byref(gp), None) != 0
    assert b"multiples of 64" in lib.cpd_last_error()
    # thresholding entry points: null pointers, unknown algorithms, percentiles / quantiles outside their range
    assert lib.cpd_threshold(None, 1, 64, _lib.CPD_THRESH_DYNAMIC, 99.0, 1, None, None) != 0
    assert b"null pointer" in lib.cpd_last_error()
    assert lib.cpd_threshold(16, 1, 64, 9, 99.0, 1, 16, None) != 0
    assert b"unknown algorithm" in lib.cpd_last_error()
    assert lib.cpd_threshold(16, 1, 64, _lib.CPD_THRESH_DYNAMIC, 101.0, 1, 16, None) != 0
    assert b"outside [0, 100]" in lib.cpd_last_error()
    assert lib.cpd_threshold(16, 1, 66, _lib.CPD_THRESH_STATIC, 1.0, 1, 16, None) != 0
    assert b"multiple of 4" in lib.cpd_last_error()
    assert lib.cpd_threshold_ex(16, 1, 4, 64, _lib.CPD_THRESH_STATIC, 1.0, 16, None) != 0  # the clamp variants live in cpd_threshold
    assert b"unknown algorithm" in lib.cpd_last_error()
    assert lib.cpd_threshold_ex(16, 1, 4, 64, _lib.CPD_THRESH_RENORM, 250.0, 16, None) != 0
    assert b"outside [0, 1]" in lib.cpd_last_error()
    assert lib.cpd_threshold_ex(16, 1, 0, 64, _lib.CPD_THRESH_RENORM, 0.9, 16, None) != 0
    # plan-level entry points: configuration errors are caught before anything touches the device
    assert lib.cpd_unet_plan_create(None, None) != 0
    assert b"null argument" in lib.cpd_last_error()
    cfg, plan = _lib.UNetConfig(), ctypes.c_void_p()
    cfg.n_levels, cfg.model_channels = 2, 100
    assert lib.cpd_unet_plan_create(ctypes.byref(cfg), ctypes.byref(plan)) != 0 and not plan.value
    assert b"multiple of 64" in lib.cpd_last_error()
    assert lib.cpd_unet_forward(None, None, None) != 0
    assert lib.cpd_pack_weights(None, b"x", None, 0, 0, 0) != 0
    # the tuned-variant table round-trips through its text form
    assert lib.cpd_gemm_tune_import(b"1,1,4096,320,0,320,1,1,0,0,0,0,1,0=160\n") == 1
    n = lib.cpd_gemm_tune_export(None, 0)
    buf = ctypes.create_string_buffer(int(n))
    lib.cpd_gemm_tune_export(buf, n)
    assert b"1,1,4096,320,0,320,1,1,0,0,0,0,1,0=160" in buf.value
    assert lib.cpd_gemm_tune_import(b"garbage") == -1
    p2 = _lib.StepParams()
    p2.eps = p2.x = 16
    p2.n_sub, p2.hw = 17, 64
    assert lib.cpd_sampler_step(ctypes.byref(p2), None) != 0
    assert b"out of range [0,16]" in lib.cpd_last_error()
    assert lib.cpd_threshold(16, 0, 64, _lib.CPD_THRESH_DYNAMIC, 99.0, 1, 16, None) == 0  # empty batch: nothing launched
    assert lib.cpd_threshold_ex(16, 0, 4, 64, _lib.CPD_THRESH_RENORM, 0.9, 16, None) == 0


def test_groupnorm_launch_plan_without_gpu(lib_path):
    """cpd_groupnorm_launches is pure host logic: every GroupNorm shape of the SD-1.5 / SD-2.1 / SDXL UNets and of the first-stage
    decoder's 64 x 64 level takes the one-launch shared-memory kernel (1); shapes whose (image, group slab) cannot be split into
    resident blocks, group sizes a vector of 8 channels cannot be split on, and bad arguments do not (2 / 0)."""
    lib = ctypes.CDLL(lib_path)
    f = lib.cpd_groupnorm_launches
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_int] * 3
    for hw in (64, 144, 256, 576, 1024, 2304, 4096, 9216):
        for c in (320, 640, 1280, 2560):
            if (c, hw) != (2560, 9216):  # 80 channels per group at 96 x 96: a 1.5 MB slab (no UNet has it)
                assert f(c, 16, hw) == 1, (c, hw)
    assert f(2560, 16, 9216) == 2
    for c, hw in ((960, 1024), (1920, 1024), (1920, 256), (2560, 64), (512, 4096), (128, 1024)):
        assert f(c, 4, hw) == 1, (c, hw)
    assert f(960, 16, 4096) == 2      # a 960 KB slab: not even a cluster of 8 blocks holds it
    assert f(128, 1, 512 * 512) == 2  # first-stage decoder at full resolution
    assert f(96, 2, 256) == 2 and f(480, 2, 256) == 2  # 3 / 15 channels per group
    assert f(320, 70000, 64) == 2     # more images than the grid's z dimension
    assert f(100, 1, 64) == 0 and f(320, 0, 64) == 0 and f(320, 1, 0) == 0


def test_registry_and_wrapper_surface():
    from complex_prompt_diffusion_b200 import samplers
    assert {"Euler", "Euler Ancestral", "DPM++ 2m"} <= set(samplers.lookup)
    with pytest.raises(KeyError):
        samplers.make({"name": "nope", "args": {}})
    with pytest.raises(ValueError):
        samplers.create(3)
    # DiffusionSamplerWrapper (diffusion.py:51-111): defaults, to_json layout, shape = [C, W // 8, H // 8] (W before H)
    from complex_prompt_diffusion_b200.samplers.diffusion import DiffusionSamplerWrapper
    calls = []

    class Inner:
        def __init__(self, model):
            self.model = model

        def sample(self, **kw):
            calls.append(kw)
            return ("latents", "aux")

    wr = DiffusionSamplerWrapper("Euler", constructor=Inner, model={"unet": None}, width=768, height=512, steps=30, scale=5.0)
    assert wr.to_json() == {"name": "Euler", "args": {"batch_size": 1, "width": 768, "height": 512, "z_channels": 4, "scale": 5.0,
                                                       "use_start_code": False, "steps": 30, "eta": 0, "temperature": 1,
                                                       "denoising_strength": 0.0}}
    assert wr.sample(conditioning="c", scheduler="karras") == "latents"
    kw = calls[-1]
    assert kw["shape"] == [4, 96, 64] and kw["steps"] == 30 and kw["batch_size"] == 1 and kw["conditioning"] == "c"
    assert kw["unconditional_guidance_scale"] == 5.0 and kw["eta"] == 0 and kw["temperature"] == 1 and kw["x_T"] is None
    wr2 = DiffusionSamplerWrapper("Euler", constructor=Inner, model=None, use_start_code=True, batch_size=2)
    wr2.sample()
    assert tuple(calls[-1]["x_T"].shape) == (2, 4, 64, 64)


def test_product_has_no_cpu_fallback():
    from complex_prompt_diffusion_b200.samplers.extension.denoiser import Denoiser

    class CpuUNet:
        def parameters(self):
            return iter([torch.zeros(1)])

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Denoiser(CpuUNet())
    from complex_prompt_diffusion_b200.models.unet import UNetModel
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        UNetModel(device="cpu")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "complex_prompt_diffusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports the oracle"
    # developer tools measure the product, never the oracle
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            src = open(os.path.join(ROOT, "tools", f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"tools/{f} imports the oracle"
    # bench.py: only the CPU legs (cpu_sample / oracle_cfg, i.e. cpu_baseline and --impl reference) may touch oracle/
    src = open(os.path.join(ROOT, "bench.py")).read()
    gpu_arm = src[src.index("def run_b200"):src.index("def sampler_roofline")]
    assert not re.search(r"^\s*(from|import)\s+oracle", gpu_arm, re.M), "the GPU arm of bench.py imports the oracle"


def test_fixtures_agree_with_the_oracle_enumerators():
    """The product-side presets / parameter shapes / FLOP enumerators (used by bench.py's GPU arm) describe the same
    architectures as the oracle's."""
    from complex_prompt_diffusion_b200.models import fixtures as F
    from oracle.unet import UNetConfig, param_shapes, count_flops
    from oracle.vae import VAEConfig, param_shapes as vae_shapes, count_flops as vae_fl
    for n in ("sd15", "sd21", "sdxl", "tiny", "tiny_xl"):
        assert F.unet_param_shapes(F.UNET_PRESETS[n]) == param_shapes(getattr(UNetConfig, n)())
        assert F.unet_flops(F.UNET_PRESETS[n], 64, 64) == count_flops(getattr(UNetConfig, n)(), 64, 64)["total"]
    for n in ("sd", "tiny"):
        assert F.vae_param_shapes(F.VAE_PRESETS[n]) == vae_shapes(getattr(VAEConfig, n)())
        assert F.vae_flops(F.VAE_PRESETS[n], 64, 64) == vae_fl(getattr(VAEConfig, n)(), 64, 64)
    assert abs(F.unet_flops(F.UNET_PRESETS["sd15"], 64, 64) / 1e12 - 0.8033) < 1e-3


def test_schedule_matches_oracle_and_kats(golden_dir):
    import json
    from complex_prompt_diffusion_b200.scheduler import SigmaScheduler
    from oracle.schedule import OracleSchedule
    kat = json.load(open(os.path.join(golden_dir, "schedule_kat.json")))
    s, o = SigmaScheduler(), OracleSchedule()
    assert torch.equal(s.sigmas, o.sigmas)
    for key, gs in kat["get_sigmas"].items():
        alg, n = key.split("_")
        sig = s.get_sigmas(alg, int(n))
        assert sig.view(torch.int32 if sig.dtype == torch.float32 else torch.int64).tolist() == gs["bits"]
        low, high = s.sigma_to_idx(sig[:-1])
        assert low.tolist() == kat["sigma_to_t"][key]["low_idx"] and high.tolist() == kat["sigma_to_t"][key]["high_idx"]
        assert [float(v) for v in s.sigma_to_t(sig[:-1])] == kat["sigma_to_t"][key]["t"]
    # binary search == top-k on random, boundary and out-of-range sigmas (integer contract, bit-exact)
    g = torch.Generator().manual_seed(0)
    sig = torch.cat([torch.rand(2000, generator=g) * 16, s.sigmas[:50].float(), s.sigmas[-50:].float(),
                     ((s.sigmas[:-1] + s.sigmas[1:]) / 2)[::37].float(), torch.tensor([0.0, 1e-6, 14.26, 20.0, 100.0])])
    t_o, low_o, high_o = o.sigma_to_t_idx(sig)
    low, high = s.sigma_to_idx(sig)
    assert torch.equal(low, low_o) and torch.equal(high, high_o)
    assert torch.equal(s.sigma_to_t(sig), t_o)


def test_step_scalars_match_oracle_expressions():
    from complex_prompt_diffusion_b200.samplers.dpmpp import dpmpp_2m_scalars
    from complex_prompt_diffusion_b200.samplers.euler import get_ancestral_step
    from oracle import samplers as OS
    from oracle.schedule import OracleSchedule
    sig = OracleSchedule().get_sigmas("karras", 8)
    for i in range(8):
        a, b = get_ancestral_step(sig[i], sig[i + 1])
        c, d = OS.get_ancestral_step(sig[i], sig[i + 1])
        assert float(a) == float(c) and float(b) == float(d)
        ratio, em, c1, c2, first = dpmpp_2m_scalars(sig, i, have_history=i > 0)
        t, tn = sig[i].log().neg(), sig[i + 1].log().neg()
        assert ratio == float(tn.neg().exp() / t.neg().exp()) and em == float((-(tn - t)).expm1())
        assert first == int(i == 0 or i == 7)


def test_guidance_decay_matches_oracle():
    from complex_prompt_diffusion_b200.samplers.extension.denoiser import Denoiser
    from oracle.denoiser import guidance_scale
    for t_idx in (0, 3, 5, 19):
        a = Denoiser.guidance_scale(unconditional_guidance_scale=7.5, t_idx=t_idx, total_steps=21, decaying_uc_scale=True)
        assert a == float(guidance_scale(7.5, t_idx, 21, True))


def test_extension_registry_mirrors_the_reference_names():
    """samplers/extension/registry.py + threshold.py:7-286: the same names resolve here, with the reference's constructor
    (threshold_x, threshold_e) and methods; nothing runs on the CPU (the product has no CPU path: it must say so)."""
    import torch
    from complex_prompt_diffusion_b200.samplers.extension import create, lookup, make
    names = {"none", "static_thresholding", "dynamic_thresholding", "dynanormic_thresholding", "scaled_dynamic_perc_thresholding",
             "renorm_thresholding", "norm_thresholding", "scaled_norm_thresholding", "spatial_norm_thresholding",
             "scaled_spatial_norm_thresholding"}
    assert names <= set(lookup)
    ext = make({"name": "dynamic_thresholding", "args": {"threshold_x": None, "threshold_e": 99.0}})  # manager.py:84-90
    assert ext.threshold_x is None and ext.threshold_e == 99.0
    for attr in ("apply", "modify_score", "_apply", "__call__"):
        assert callable(getattr(ext, attr))
    x = torch.zeros(1, 4, 8, 8)
    assert create("none")(x) is x  # the identity extension
    assert create("dynamic_thresholding").modify_score(x, x, 0, None) is x  # no threshold_e: e_t passes through
    with pytest.raises(RuntimeError, match="runs on the device"):
        create("static_thresholding")(x, threshold=0.5)
    with pytest.raises(NotImplementedError):
        create("norm_thresholding")(x, threshold=50.0)
    with pytest.raises(KeyError):
        create("no_such_extension")


def test_prompt_and_mask_parsing_match_the_reference(golden_dir):
    """WeightedPrompt._parse_prompt (prompts.py:546-589) and CompositionalPrompt._parse_mask_style
    (:737-856) against outputs of the reference's own methods (tests/golden/ref_prompts.npz, 12 strings, 950 mask styles x
    latent shapes incl. every alias), and the composition dict of _build_embeddings (:622-648)."""
    import json
    import numpy as np
    import torch
    from complex_prompt_diffusion_b200.embeddings import CompositionalConditioning, parse_mask_style, parse_weighted_prompt
    g = np.load(os.path.join(golden_dir, "ref_prompts.npz"))
    parsed = json.loads(str(g["prompt_parsed"]))
    for text, (prompts, weights) in zip(g["prompt_strings"].tolist(), parsed):
        got_p, got_w = parse_weighted_prompt(text)
        assert got_p == prompts and got_w == weights, text
    for case, line, axis in zip(g["mask_cases"].tolist(), g["mask_lines"], g["mask_axes"]):
        shape, style = case.split("|")
        H, W = (int(v) for v in shape.split("x"))
        m = parse_mask_style(style, H, W)
        assert m.dtype == torch.uint8 and tuple(m.shape) == (1, H // 8, W // 8), case
        n = m.shape[int(axis)]
        ref = torch.from_numpy(line[:n].copy())
        ref = ref.view(1, 1, n).expand_as(m) if axis == 2 else ref.view(1, n, 1).expand_as(m)
        assert torch.equal(m, ref), case
    for bad in ("perspective",):
        with pytest.raises(NotImplementedError):
            parse_mask_style(bad, 512, 512)
    for bad in ("diagonal_half", "left_1_valid", "left_half_maybe"):
        with pytest.raises(ValueError):
            parse_mask_style(bad, 512, 512)
    e = [torch.randn(1, 77, 16) for _ in range(4)]
    comp = (CompositionalConditioning(e[0], scale=1.2, height=256, width=320)
            .add_filter(e[1], strength=0.6).add_filter(e[2], strength=-0.4).add_filter(e[3], strength=0)
            .add_masked_filter(e[3], "left_half_valid", strength=0.8).build())
    assert [c[0] for c in comp["and"]] == [1.2, 0.6, 0.8] and [c[0] for c in comp["not"]] == [0.4]
    assert comp["and"][0][1] is e[0] and comp["and"][0][3] == 1 and comp["not"][0][1] is e[2]
    assert tuple(comp["and"][2][3].shape) == (1, 1, 32, 40) and int(comp["and"][2][3].sum()) == 32 * 20
    with pytest.raises(ValueError, match="embedder"):
        CompositionalConditioning("a text prompt")
    texts = []
    comp2 = CompositionalConditioning("base", embedder=lambda t: (texts.append(t), torch.zeros(1, 77, 8))[1]).add_weighted("cat:1.5 dog:-0.5 tree")
    assert texts == ["base", "cat", "dog", "tree"] and [c[0] for c in comp2.build()["and"]] == [1, 1.5, 1.0] and comp2.build()["not"][0][0] == 0.5


def test_schedules_match_the_discrete_scheduler_of_the_reference(golden_dir):
    """SigmaScheduler.get_sigmas_{karras, exponential, quad, vp, sigmoid} of cpd/scheduler/discrete.py:21-85
    (tests/golden/schedule_kat2.json, n = 1 .. 30, default and non-default sigma_min / sigma_max / rho): the oracle and the
    product scheduler return the same float32 bit patterns (get_sigmas appends the final 0)."""
    import json
    from complex_prompt_diffusion_b200.scheduler import SigmaScheduler
    from oracle.schedule import OracleSchedule
    kat = json.load(open(os.path.join(golden_dir, "schedule_kat2.json")))
    s, o = SigmaScheduler(), OracleSchedule()
    wide = {"sigma_min": 0.03, "sigma_max": 14.6, "rho": 5.0}
    for key, ref in kat.items():
        alg, n, tag = key.split("|")
        kw = dict(wide) if tag == "wide" else {}
        for impl in (s, o):
            sig = impl.get_sigmas(alg, int(n), **kw)
            assert sig.dtype == torch.float32 and float(sig[-1]) == 0.0, key
            assert sig[:-1].view(torch.int32).tolist() == ref["bits"], (key, type(impl).__name__)


def test_noise_generator_matches_the_reference(golden_dir):
    """cpd/noise.py NoiseGenerator (seed property under iter / constant / loop / random, sample(),
    noise.py:34-46,86-93) - seeds and drawn tensors identical to the reference's (tests/golden/ref_noise.npz); the oracle's
    restricted generator agrees on the modes it restates."""
    import random
    import numpy as np
    from complex_prompt_diffusion_b200.noise import NoiseGenerator, build_cycle_mod
    from oracle.samplers import OracleNoiseGenerator
    g = np.load(os.path.join(golden_dir, "ref_noise.npz"))
    assert build_cycle_mod(5) == g["build_cycle_mod_5"].tolist() and build_cycle_mod(3) == g["build_cycle_mod_3"].tolist()
    cases = sorted({k.rsplit("|", 1)[0] for k in g.files if k.endswith("|draws")})
    assert len(cases) == 6
    for case in cases:
        mode, seed, cyc = case.split("|")
        random.seed(2024)
        ng = NoiseGenerator((1, 4, 4, 4), "cpu", seed=int(seed), seed_mode=mode, cycle_size=int(cyc))
        og = OracleNoiseGenerator((1, 4, 4, 4), "cpu", seed=int(seed), seed_mode=mode) if mode in ("iter", "constant", "c") else None
        for k in range(6):
            x = ng.sample()
            assert ng.last_seed == int(g[case + "|seeds"][k]), (case, k)
            assert torch.equal(x, torch.from_numpy(g[case + "|draws"][k])), (case, k)
            if og is not None:
                assert torch.equal(og.sample(), x)
    ng = NoiseGenerator((2, 4, 8, 8), "cpu", seed=3)
    assert torch.equal(ng.sample(seed=99), torch.from_numpy(g["explicit|99"])) and ng.last_seed == int(g["explicit|after"][0])
