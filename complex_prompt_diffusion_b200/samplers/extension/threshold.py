"""Thresholding / score-corrector extensions (interface of cpd/samplers/extension/threshold.py:7-286), on the device.

Same registered names, constructor (`threshold_x`, `threshold_e`) and methods (`apply`, `modify_score`, `__call__`,
`_apply`) as the reference's classes, so `create("dynamic_thresholding", threshold_e=99.0)` or the manager's
`make({"name": ..., "args": {"threshold_x": ..., "threshold_e": ...}})` (manager.py:84-90) keep working.  The arithmetic runs
in `csrc/threshold.cu` (cpd_threshold / cpd_threshold_ex): no np.percentile on a CPU copy.  Differences, all stated in
DESIGN.md: the result is an fp32 tensor holding fp16-rounded values (repair D10: the reference returns `.half()`), images of
a batch are independent (D7), `norm_thresholding` raises (D12: NameError in the reference).
"""
import torch

from .registry import register


@register("none")
class ScoreCorrector:
    """threshold.py:7-45: the identity extension and the base of the others."""
    alg_name = "none"
    default_threshold = None

    def __init__(self, threshold_x=None, threshold_e=None):
        self.threshold_x = threshold_x
        self.threshold_e = threshold_e

    def apply(self, x, t, **kwargs):
        return self._apply(x, **kwargs)

    def modify_score(self, e_t, x, t, c, **kwargs):
        """threshold.py:17-31.  The reference also thresholds `x` when threshold_x is set but drops the result (the
        Denoiser hands it a clone, denoiser.py:524), so only e_t is processed here."""
        if self.threshold_e:
            kwargs = dict(kwargs, name="e_t", threshold=self.threshold_e)
            e_t = self._apply(e_t, **kwargs)
        return e_t

    def __call__(self, x, **kwargs):
        if isinstance(x, dict):
            if "t" not in x and "sigma" in x:
                x["t"] = x["sigma"]
            return self._apply(**x)
        return self._apply(x, **kwargs)

    def _apply(self, x, **kwargs):
        if self.alg_name == "none":
            return x
        from .denoiser import apply_threshold
        if not x.is_cuda:
            raise RuntimeError(f"{self.alg_name}: the extension runs on the device; got a {x.device} tensor")
        squeeze = x.ndim == 3
        y = (x.unsqueeze(0) if squeeze else x).float().contiguous().clone()
        bound = torch.empty(y.shape[0], dtype=torch.float32, device=y.device)
        apply_threshold(y, bound, self.alg_name, kwargs.get("threshold", self.default_threshold))
        return y.squeeze(0) if squeeze else y


def _variant(name, default):
    cls = type("".join(w.capitalize() for w in name.split("_")), (ScoreCorrector,), {"alg_name": name, "default_threshold": default,
                                                                                      "__doc__": f"`{name}` of threshold.py."})
    return register(name)(cls)


StaticThresholding = _variant("static_thresholding", 1)                               # threshold.py:47-62
DynamicThresholding = _variant("dynamic_thresholding", 99.66)                         # :63-85
DynanormicThresholding = _variant("dynanormic_thresholding", 99.66)                   # :87-116
ScaledDynamicPercThresholding = _variant("scaled_dynamic_perc_thresholding", 99.66)   # :118-146
RenormThresholding = _variant("renorm_thresholding", 99.66)                           # :148-180
ScaledNormThresholding = _variant("scaled_norm_thresholding", 99.66)                  # :207-237
SpatialNormThresholding = _variant("spatial_norm_thresholding", 99.66)                # :239-254
ScaledSpatialNormThresholding = _variant("scaled_spatial_norm_thresholding", 99.66)   # :256-286


@register("norm_thresholding")
class NormThresholding(ScoreCorrector):
    """threshold.py:182-205 reads `x_max` before assigning it: NameError in the reference (D12), nothing to reproduce."""
    alg_name = "norm_thresholding"

    def _apply(self, x, **kwargs):
        raise NotImplementedError("norm_thresholding cannot run in the reference (threshold.py:194 uses an undefined x_max)")
