"""Pin the CPU oracle against fixtures produced by the shimmed, unmodified reference
(oracle/make_golden.py) and against the known-answer values of SURVEY.md section 8-c."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.schedule import OracleSchedule
from oracle.unet import UNetConfig, OracleUNet, make_weights, count_flops
from oracle.denoiser import OracleDenoiser
from oracle import samplers as OS


@pytest.fixture(scope="module")
def kat(golden_dir):
    with open(os.path.join(golden_dir, "schedule_kat.json")) as f:
        return json.load(f)


def test_training_tables(kat):
    s = OracleSchedule()
    assert int(torch.unique(s.betas).numel()) == kat["n_distinct_betas"] == 113
    for i, v in kat["betas"].items():
        assert float(s.betas[int(i)]) == v
    for i, v in kat["alphas_cumprod"].items():
        assert float(s.alphas_cumprod[int(i)]) == v
    for i, v in kat["sigmas_table"].items():
        assert float(s.sigmas[int(i)]) == v
    assert float(s.sigmas.sum()) == kat["sigmas_table_sum"]
    # SURVEY 8-c known answers
    assert float(s.sigmas[0]) == 0.028295591748715043
    assert float(s.sigmas[999]) == 14.259975589529
    assert float(s.alphas_cumprod[999]) == 0.004893639107497308


@pytest.mark.parametrize("key", ["karras_10", "karras_20", "karras_30", "linear_10", "linear_20", "exp_10", "quad_10", "vp_10"])
def test_get_sigmas_and_indices_bit_exact(kat, key):
    alg, n = key.split("_")
    s = OracleSchedule()
    sig = s.get_sigmas(alg, int(n))
    g = kat["get_sigmas"][key]
    assert str(sig.dtype) == g["dtype"]
    bits = sig.view(torch.int32 if sig.dtype == torch.float32 else torch.int64).tolist()
    assert bits == g["bits"]
    t, low, high = s.sigma_to_t_idx(sig[:-1])
    assert low.tolist() == kat["sigma_to_t"][key]["low_idx"]
    assert high.tolist() == kat["sigma_to_t"][key]["high_idx"]
    assert [float(v) for v in t] == kat["sigma_to_t"][key]["t"]


def test_survey_known_answers():
    s = OracleSchedule()
    sig = s.get_sigmas("karras", 10)
    _, low, high = s.sigma_to_t_idx(sig[:-1])
    assert low.tolist() == [937, 864, 778, 673, 545, 396, 243, 117, 42, 11]
    assert (high - low).tolist() == [1] * 10
    _, low20, _ = s.sigma_to_t_idx(s.get_sigmas("karras", 20)[:-1])
    assert low20.tolist() == [937, 904, 869, 830, 788, 741, 691, 635, 574, 508, 437, 364, 290, 220, 158, 106, 67, 39, 21, 11]
    t = s.sigma_to_t(sig[:-1])
    assert float(t[0]) == 937.9314529721662 and float(t[-1]) == 11.278044358791906


def test_t_to_sigma(kat):
    s = OracleSchedule()
    got = s.t_to_sigma(torch.tensor(kat["t_to_sigma"]["t"]))
    assert [float(v) for v in got] == kat["t_to_sigma"]["sigma"]


def test_flop_enumerator_matches_baseline():
    # BASELINE.md section 3 (FlopCounterMode on the reference module)
    assert abs(count_flops(UNetConfig.sd15(), 32, 32)["total"] / 1e12 - 0.1801) < 1e-4
    assert abs(count_flops(UNetConfig.sd15(), 64, 64)["total"] / 1e12 - 0.8033) < 1e-4
    assert abs(count_flops(UNetConfig.sd21(), 96, 96)["total"] / 1e12 - 2.1491) < 1e-4


def _load_case(golden_dir):
    z = np.load(os.path.join(golden_dir, "ref_sampling.npz"))
    embs = torch.from_numpy(z["embs"])
    mask = torch.from_numpy(z["mask"])
    sc = z["scales"]
    c = {"and": [(float(sc[0]), embs[0:1], None, 1), (float(sc[1]), embs[1:2], None, mask)],
         "not": [(float(sc[2]), embs[2:3], None, 1)]}
    return z, c


@pytest.mark.parametrize("name,sched,pred", [("Euler", "karras", "epsilon"), ("DPM++ 2m", "karras", "epsilon"),
                                             ("Euler Ancestral", "karras", "epsilon"), ("Euler", "exp", "velocity"),
                                             ("DPM++ 2m", "linear", "velocity")])
def test_oracle_sampling_end_to_end_close_to_reference(golden_dir, name, sched, pred):
    z, c = _load_case(golden_dir)
    cfg = UNetConfig.tiny()
    den = OracleDenoiser(OracleUNet(cfg, make_weights(cfg, seed=0)))
    key = f"{name}|{sched}|{pred}".replace(" ", "_")
    noises = list(torch.from_numpy(z[key + "|noise"])) if (key + "|noise") in z.files else []
    dens = []
    out = OS.sample(den, name, int(z["steps"]), torch.from_numpy(z["x_T"]).clone(),
                    noise_sampler=(lambda x: noises.pop(0)) if noises else None,
                    callback=lambda d: dens.append(d["eps"].clone()),
                    conditioning=c, unconditional_conditioning=torch.from_numpy(z["uc"]),
                    unconditional_guidance_scale=float(z["guidance"]), scheduler=sched, pred_type=pred)
    ref_final = torch.from_numpy(z[key + "|final"])
    ref_den = torch.from_numpy(z[key + "|denoised"])
    for i, d in enumerate(dens):
        rel = ((d - ref_den[i]).norm() / ref_den[i].norm()).item()
        assert rel < 5e-3, (i, rel)  # fp16-delta rounding flips amplify the ~1e-6 UNet difference (bit-exact test above)
    rel = ((out - ref_final).norm() / ref_final.norm()).item()
    assert rel < 5e-3, rel


CASES = [("Euler", "karras", "epsilon"), ("DPM++ 2m", "karras", "epsilon"), ("Euler Ancestral", "karras", "epsilon"),
         ("Euler", "exp", "velocity"), ("DPM++ 2m", "linear", "velocity")]


class _ReplayUNet:
    """Returns the reference UNet's recorded outputs, so the denoiser/sampler restatement is checked BIT-EXACTLY."""

    def __init__(self, outs, xs, ts):
        self.outs, self.xs, self.ts, self.i = outs, xs, ts, 0

    def parameters(self):
        return iter([torch.zeros(1)])

    def __call__(self, x, t, ctx, **kw):
        assert torch.equal(x, self.xs[self.i]), "UNet input x differs from the reference's"
        assert torch.equal(t, self.ts[self.i]), "UNet input t differs from the reference's"
        o = self.outs[self.i]
        self.i += 1
        return o, [o] * 12


@pytest.mark.parametrize("name,sched,pred", CASES)
def test_oracle_denoiser_and_samplers_bit_exact_on_replayed_unet(golden_dir, name, sched, pred):
    z, c = _load_case(golden_dir)
    key = f"{name}|{sched}|{pred}".replace(" ", "_")
    unet = _ReplayUNet(torch.from_numpy(z[key + "|unet_out"]), torch.from_numpy(z[key + "|unet_x"]), torch.from_numpy(z[key + "|unet_t"]))
    den = OracleDenoiser(unet, dtype=torch.float32)
    noises = list(torch.from_numpy(z[key + "|noise"])) if (key + "|noise") in z.files else []
    dens = []
    out = OS.sample(den, name, int(z["steps"]), torch.from_numpy(z["x_T"]).clone(),
                    noise_sampler=(lambda x: noises.pop(0)) if noises else None,
                    callback=lambda d: dens.append(d["eps"].clone()),
                    conditioning=c, unconditional_conditioning=torch.from_numpy(z["uc"]),
                    unconditional_guidance_scale=float(z["guidance"]), scheduler=sched, pred_type=pred)
    assert torch.equal(torch.stack(dens), torch.from_numpy(z[key + "|denoised"]))
    assert torch.equal(out, torch.from_numpy(z[key + "|final"]))


def test_oracle_unet_matches_reference_unet(golden_dir):
    z, _ = _load_case(golden_dir)
    cfg = UNetConfig.tiny()
    unet = OracleUNet(cfg, make_weights(cfg, seed=0))
    key = "Euler|karras|epsilon"
    ctx = torch.cat([torch.from_numpy(z["uc"]), torch.from_numpy(z["embs"])])
    for i in (0, 3, 5):
        out = unet(torch.from_numpy(z[key + "|unet_x"][i]), torch.from_numpy(z[key + "|unet_t"][i]), ctx)
        ref = torch.from_numpy(z[key + "|unet_out"][i])
        assert ((out - ref).norm() / ref.norm()).item() < 1e-5


def test_fp16_delta_differs_from_exact_combine():
    """the reference computes the weighted delta in fp16 (denoiser.py:450-460); the
    implementation must match THAT, which differs from the exact fp32 combine by ~2e-3 rel-L2."""
    from oracle.denoiser import combine_fp16
    g = torch.Generator().manual_seed(0)
    e_u = torch.randn(1, 4, 64, 64, generator=g)
    e_k = [e_u + 0.3 * torch.randn(1, 4, 64, 64, generator=g) for _ in range(3)]
    w = [torch.tensor([1.0]), torch.tensor([0.6]), torch.tensor([-0.4])]
    m = [torch.tensor([1.0])] * 3
    got = e_u + 7.5 * combine_fp16(e_k, e_u, w, m)
    exact = e_u + 7.5 * sum(wk * (ek - e_u) for wk, ek in zip(w, e_k))
    rel = ((got - exact).norm() / exact.norm()).item()
    assert 1e-4 < rel < 1e-2


# ---- remaining k-diffusion samplers + Denoiser scale clip (tests/golden/ref_sampling2.npz, oracle/make_golden.py) ----
MORE_CASES = [("Huen", "karras", "epsilon", {}), ("DPM2", "karras", "epsilon", {}), ("DPM2 Ancestral", "karras", "epsilon", {}),
              ("DPM++ 2s Ancestral", "karras", "epsilon", {}), ("LMS", "karras", "epsilon", {}), ("Huen", "exp", "velocity", {}),
              ("Euler", "karras", "epsilon", {"scaled_clip": True, "scaled_clip_threshold": 97.0}),
              ("DPM++ 2m", "karras", "epsilon", {"scaled_clip": True, "scaled_clip_alg": "static_thresholding",
                                                 "scaled_clip_threshold": 0.5})]


def more_key(name, sched, pred, extra):
    return f"{name}|{sched}|{pred}".replace(" ", "_") + ("|" + "|".join(f"{k}={v}" for k, v in extra.items()) if extra else "")


@pytest.mark.parametrize("name,sched,pred,extra", MORE_CASES)
def test_oracle_more_samplers_bit_exact_on_replayed_unet(golden_dir, name, sched, pred, extra):
    z, c = _load_case(golden_dir)
    z2 = np.load(os.path.join(golden_dir, "ref_sampling2.npz"))
    key = more_key(name, sched, pred, extra)
    unet = _ReplayUNet(torch.from_numpy(z2[key + "|unet_out"]), torch.from_numpy(z2[key + "|unet_x"]), torch.from_numpy(z2[key + "|unet_t"]))
    den = OracleDenoiser(unet, dtype=torch.float32)
    noises = list(torch.from_numpy(z2[key + "|noise"])) if (key + "|noise") in z2.files else []
    dens = []
    out = OS.sample(den, name, int(z["steps"]), torch.from_numpy(z["x_T"]).clone(),
                    noise_sampler=(lambda x: noises.pop(0)) if noises else None,
                    callback=lambda d: dens.append(d["eps"].clone()),
                    conditioning=c, unconditional_conditioning=torch.from_numpy(z["uc"]),
                    unconditional_guidance_scale=float(z["guidance"]), scheduler=sched, pred_type=pred, **extra)
    assert unet.i == len(unet.outs), "the restatement made a different number of UNet evaluations than the reference"
    assert torch.equal(torch.stack(dens), torch.from_numpy(z2[key + "|denoised"]))
    assert torch.equal(out, torch.from_numpy(z2[key + "|final"]))


def test_oracle_vae_decoder_matches_reference(golden_dir):
    """the first-stage decoder restatement (oracle/vae.py) against the shimmed reference Decoder +
    post_quant_conv on the same seeded weights and latent (tests/golden/ref_vae.npz)."""
    from oracle.vae import VAEConfig, OracleVAEDecoder, make_weights as vae_weights
    z = np.load(os.path.join(golden_dir, "ref_vae.npz"))
    cfg = VAEConfig.tiny()
    out = OracleVAEDecoder(cfg, vae_weights(cfg, seed=0))(torch.from_numpy(z["z"]))
    ref = torch.from_numpy(z["image"])
    rel = ((out - ref).norm() / ref.norm()).item()
    assert out.shape == ref.shape and rel < 1e-5, rel


CHURN_CASES = [("Euler", "karras", "epsilon", {"s_churn": 4.0, "s_noise": 1.003}), ("Huen", "karras", "epsilon", {"s_churn": 4.0}),
               ("DPM2", "karras", "epsilon", {"s_churn": 9.0, "s_tmin": 0.5, "s_tmax": 6.0, "s_noise": 0.99})]


@pytest.mark.parametrize("name,sched,pred,extra", CHURN_CASES)
def test_oracle_stochastic_churn_bit_exact_on_replayed_unet(golden_dir, name, sched, pred, extra):
    """s_churn > 0 (gamma > 0: noise added before the denoiser call, sigma_hat > sigma) against runs of the shimmed reference
    (tests/golden/ref_sampling3.npz)."""
    z, c = _load_case(golden_dir)
    z3 = np.load(os.path.join(golden_dir, "ref_sampling3.npz"))
    key = more_key(name, sched, pred, extra)
    unet = _ReplayUNet(torch.from_numpy(z3[key + "|unet_out"]), torch.from_numpy(z3[key + "|unet_x"]), torch.from_numpy(z3[key + "|unet_t"]))
    den = OracleDenoiser(unet, dtype=torch.float32)
    noises = list(torch.from_numpy(z3[key + "|noise"]))
    dens = []
    out = OS.sample(den, name, int(z["steps"]), torch.from_numpy(z["x_T"]).clone(), noise_sampler=lambda x: noises.pop(0),
                    callback=lambda d: dens.append(d["eps"].clone()), conditioning=c, unconditional_conditioning=torch.from_numpy(z["uc"]),
                    unconditional_guidance_scale=float(z["guidance"]), scheduler=sched, pred_type=pred, **extra)
    assert torch.equal(torch.stack(dens), torch.from_numpy(z3[key + "|denoised"]))
    assert torch.equal(out, torch.from_numpy(z3[key + "|final"]))


def test_oracle_thresholding_extensions_bit_exact_vs_reference(golden_dir):
    """Every runnable registered extension of samplers/extension/threshold.py (oracle/make_golden.py --threshold-only ran
    the reference's own classes): the restatement reproduces the reference's fp16 output bit for bit."""
    from oracle.make_golden import THRESHOLD_CASES
    from oracle.samplers import threshold_apply
    g = np.load(os.path.join(golden_dir, "ref_threshold.npz"))
    for j in range(2):
        x = torch.from_numpy(g[f"x{j}"])
        for k, (name, thr) in enumerate(THRESHOLD_CASES):
            y = threshold_apply(x, name, thr)
            assert torch.equal(y, torch.from_numpy(g[f"y{j}_{k}"]).float()), (j, name, thr)
    with pytest.raises(NotImplementedError):
        threshold_apply(x, "norm_thresholding", 50.0)  # D12: NameError in the reference


CORRECTOR_CASES = [("Euler", "karras", "epsilon", {"score_corrector": ("static_thresholding", 1.5, 0.9)}),
                   ("DPM++ 2m", "karras", "epsilon", {"score_corrector": ("dynamic_thresholding", 95.0, 97.0)}),
                   ("Euler Ancestral", "karras", "epsilon", {"score_corrector": ("renorm_thresholding", None, 96.0)}),
                   ("Huen", "karras", "epsilon", {"score_corrector": ("scaled_dynamic_perc_thresholding", 90.0, 95.0),
                                                  "scaled_clip": True, "scaled_clip_alg": "dynanormic_thresholding",
                                                  "scaled_clip_threshold": 99.0})]


@pytest.mark.parametrize("name,sched,pred,extra", CORRECTOR_CASES)
def test_oracle_score_corrector_bit_exact_on_replayed_unet(golden_dir, name, sched, pred, extra):
    """The score_corrector hook (denoiser.py:517-518) and a non-clamp scaled_clip_alg against runs of the shimmed reference
    with its own extension classes (tests/golden/ref_sampling4.npz)."""
    z, c = _load_case(golden_dir)
    z4 = np.load(os.path.join(golden_dir, "ref_sampling4.npz"))
    key = more_key(name, sched, pred, extra)
    extra = dict(extra)
    nm, tx, te = extra["score_corrector"]
    extra["score_corrector"] = OS.OracleScoreCorrector(nm, tx, te)
    unet = _ReplayUNet(torch.from_numpy(z4[key + "|unet_out"]), torch.from_numpy(z4[key + "|unet_x"]), torch.from_numpy(z4[key + "|unet_t"]))
    den = OracleDenoiser(unet, dtype=torch.float32)
    noises = list(torch.from_numpy(z4[key + "|noise"])) if (key + "|noise") in z4.files else []
    dens = []
    out = OS.sample(den, name, int(z["steps"]), torch.from_numpy(z["x_T"]).clone(),
                    noise_sampler=(lambda x: noises.pop(0)) if noises else None,
                    callback=lambda d: dens.append(d["eps"].clone()),
                    conditioning=c, unconditional_conditioning=torch.from_numpy(z["uc"]),
                    unconditional_guidance_scale=float(z["guidance"]), scheduler=sched, pred_type=pred, **extra)
    assert unet.i == len(unet.outs)
    assert torch.equal(torch.stack(dens), torch.from_numpy(z4[key + "|denoised"]))
    assert torch.equal(out, torch.from_numpy(z4[key + "|final"]))


IMG2IMG_CASES = [("Euler", "karras", "epsilon", {"decode": True, "denoising_strength": 0.6}),
                 ("DPM++ 2m", "karras", "epsilon", {"decode": True, "denoising_strength": 0.35}),
                 ("Euler Ancestral", "exp", "epsilon", {"decode": True, "denoising_strength": 1.0})]


@pytest.mark.parametrize("name,sched,pred,extra", IMG2IMG_CASES)
def test_oracle_img2img_branch_bit_exact_on_replayed_unet(golden_dir, name, sched, pred, extra):
    """The decode=True branch of KDiffusionSampler.sample (k_diffusion.py:64-70: schedule truncated by denoising_strength,
    strength clamped at 0.999, x = x_T + randn * sigmas[0]) against runs of the shimmed reference (ref_sampling5.npz); the
    initial noise is the first draw of the global CPU generator after torch.manual_seed(77), as in the recorded run."""
    z, c = _load_case(golden_dir)
    z5 = np.load(os.path.join(golden_dir, "ref_sampling5.npz"))
    key = more_key(name, sched, pred, extra)
    unet = _ReplayUNet(torch.from_numpy(z5[key + "|unet_out"]), torch.from_numpy(z5[key + "|unet_x"]), torch.from_numpy(z5[key + "|unet_t"]))
    den = OracleDenoiser(unet, dtype=torch.float32)
    noises = list(torch.from_numpy(z5[key + "|noise"])) if (key + "|noise") in z5.files else []
    dens = []
    torch.manual_seed(77)
    out = OS.sample(den, name, int(z["steps"]), torch.from_numpy(z["x_T"]).clone(),
                    noise_sampler=(lambda x: noises.pop(0)) if noises else None,
                    callback=lambda d: dens.append(d["eps"].clone()),
                    conditioning=c, unconditional_conditioning=torch.from_numpy(z["uc"]),
                    unconditional_guidance_scale=float(z["guidance"]), scheduler=sched, pred_type=pred, **extra)
    assert unet.i == len(unet.outs) and unet.i < int(z["steps"]), "truncated schedule: fewer evaluations than steps"
    assert torch.equal(torch.stack(dens), torch.from_numpy(z5[key + "|denoised"]))
    assert torch.equal(out, torch.from_numpy(z5[key + "|final"]))


DECAY_CASES = [("Euler", "karras", "epsilon", {"decaying_uc_scale": True}),
               ("DPM++ 2m", "karras", "epsilon", {"decaying_uc_scale": True, "decaying_uc_scale_start": 0, "decaying_uc_scale_min": 3}),
               ("Huen", "exp", "epsilon", {"decaying_uc_scale": True, "decaying_uc_scale_start": 2, "decaying_uc_scale_min": 0.5})]


@pytest.mark.parametrize("name,sched,pred,extra", DECAY_CASES)
def test_oracle_guidance_decay_bit_exact_on_replayed_unet(golden_dir, name, sched, pred, extra):
    """Decaying guidance scale (denoiser.py:477-494) against runs of the shimmed reference (tests/golden/ref_sampling6.npz)."""
    z, c = _load_case(golden_dir)
    z6 = np.load(os.path.join(golden_dir, "ref_sampling6.npz"))
    key = more_key(name, sched, pred, extra)
    unet = _ReplayUNet(torch.from_numpy(z6[key + "|unet_out"]), torch.from_numpy(z6[key + "|unet_x"]), torch.from_numpy(z6[key + "|unet_t"]))
    den = OracleDenoiser(unet, dtype=torch.float32)
    dens = []
    out = OS.sample(den, name, int(z["steps"]), torch.from_numpy(z["x_T"]).clone(), callback=lambda d: dens.append(d["eps"].clone()),
                    conditioning=c, unconditional_conditioning=torch.from_numpy(z["uc"]),
                    unconditional_guidance_scale=float(z["guidance"]), scheduler=sched, pred_type=pred, **extra)
    assert unet.i == len(unet.outs)
    assert torch.equal(torch.stack(dens), torch.from_numpy(z6[key + "|denoised"]))
    assert torch.equal(out, torch.from_numpy(z6[key + "|final"]))


def test_oracle_unet_injection_matches_reference_unet(golden_dir):
    """unet.py:774-813: return_attn (the 12 skip tensors, deepest first), return_feat, inject_attns / inject_feats with their
    *_stop indices - the oracle UNet against the shimmed reference UNetModel on the same seeded weights
    (tests/golden/ref_unet_inject.npz).  The injected tensors are rebuilt from the oracle's own forward pass."""
    z = np.load(os.path.join(golden_dir, "ref_unet_inject.npz"))
    cfg = UNetConfig.tiny()
    unet = OracleUNet(cfg, make_weights(cfg, seed=0))
    x, t, ctx = (torch.from_numpy(z[k]) for k in ("x", "t", "ctx"))

    def close(a, b):
        b = torch.from_numpy(b)
        return a.shape == b.shape and ((a - b).norm() / b.norm()).item() < 1e-5

    out, skips, feats = unet(x, t, ctx, return_attn=True, return_feat=True)
    assert len(skips) == int(z["n_skips"]) and len(feats) == int(z["n_feats"])
    assert [list(s.shape) for s in skips] == z["skip_shapes"].tolist() and [list(f.shape) for f in feats] == z["feat_shapes"].tolist()
    assert close(out, z["out"]) and close(skips[0], z["skip0"]) and close(skips[-1], z["skip_last"]) and close(feats[-1], z["feat_last"])
    inj_a = [s * 0.5 for s in skips]
    inj_f = [skips[0] * 0.3] + [f * 0.7 for f in feats[:-1]]
    out_a = unet(x, t, ctx, inject_attns=inj_a, inject_attns_stop=5)
    out_f = unet(x, t, ctx, inject_feats=inj_f, inject_feats_stop=3)
    out_af, skips_af = unet(x, t, ctx, return_attn=True, inject_attns=inj_a, inject_attns_stop=12, inject_feats=inj_f, inject_feats_stop=7)
    assert close(out_a, z["out_a"]) and close(out_f, z["out_f"]) and close(out_af, z["out_af"])
    assert close(skips_af[0], z["returned_skip0_af"])  # return_attn reports the ORIGINAL skip, not the injected one (:802-808)
    assert not close(out_a, z["out"]) and not close(out_f, z["out"])


# ------------------------------------------------------------------------------ guidance branches
def _guidance_cases():
    from oracle.make_golden import GUIDANCE_CASES
    return list(GUIDANCE_CASES)


@pytest.mark.parametrize("name,sched,pred,extra", _guidance_cases())
def test_oracle_guidance_branches_bit_exact_on_replayed_unet(golden_dir, name, sched, pred, extra):
    """The oracle's unconditional-blur / attention-guidance / depth-mask restatement (oracle/denoiser.py) against runs of the
    shimmed reference (tests/golden/ref_sampling7.npz): every UNet input the reference produced (x * c_in incl. the depth
    channel, the timestep, the guided latent of the extra evaluation) and the final latent are reproduced bit for bit when the
    recorded UNet outputs and saliency sources are replayed.  The blur's random sigma comes from the global torch RNG, seeded
    like the generator script (the reference's Euler also draws its unused churn noise from it every step)."""
    z = np.load(os.path.join(golden_dir, "ref_sampling7.npz"))
    embs, mask, sc = torch.from_numpy(z["embs"]), torch.from_numpy(z["mask"]), z["scales"]
    c = {"and": [(float(sc[0]), embs[0:1], None, 1), (float(sc[1]), embs[1:2], None, mask)], "not": [(float(sc[2]), embs[2:3], None, 1)]}
    key = f"{name}|{sched}|{pred}".replace(" ", "_") + "|" + "|".join(f"{k}={v}" for k, v in extra.items())

    class Replay:
        i = 0

        def parameters(self):
            return iter([torch.zeros(1)])

        def __call__(self, x, t, ctx, **kw):
            k = f"{key}|call{self.i}"
            assert torch.equal(x, torch.from_numpy(z[k + "|x"])), f"UNet input x of call {self.i} differs from the reference's"
            assert torch.equal(t.double(), torch.from_numpy(z[k + "|t"])), f"UNet input t of call {self.i} differs from the reference's"
            o = torch.from_numpy(z[k + "|out"])
            sk = torch.from_numpy(z[k + "|skip"]) if (k + "|skip") in z.files else o
            self.i += 1
            return o, [sk] * 12

    ex = dict(extra)
    if ex.get("depth_mask"):
        ex["depth_mask"] = torch.from_numpy(z["depth_mask"])
    den = OracleDenoiser(Replay())
    torch.manual_seed(77)
    out = OS.sample(den, name, int(z["steps"]), torch.from_numpy(z["x_T"]).clone(),
                    noise_sampler=(lambda x: torch.randn_like(x)) if name == "Euler" else None, conditioning=c,
                    unconditional_conditioning=torch.from_numpy(z["uc"]), unconditional_guidance_scale=float(z["guidance"]),
                    scheduler=sched, pred_type=pred, **ex)
    assert den.unet.i == int(z[key + "|n_calls"])
    assert torch.equal(out, torch.from_numpy(z[key + "|final"]))
