"""DiffusionSamplerWrapper: configuration holder around an inner sampler
(interface of cpd/samplers/diffusion.py:51-127)."""


class DiffusionSamplerWrapper:
    _FIELDS = (("batch_size", 1), ("width", 512), ("height", 512), ("z_channels", 4), ("scale", 7.5),
               ("use_start_code", False), ("steps", 50), ("eta", 0), ("temperature", 1), ("denoising_strength", 0.0))

    def __init__(self, name: str, **kwargs):
        constructor = kwargs.get("constructor")
        if constructor is None:
            raise ValueError("DiffusionSamplerWrapper needs a `constructor` for the inner sampler")
        self.sampler = constructor(kwargs.get("model"))
        self.name = name
        for key, default in self._FIELDS:
            setattr(self, key, kwargs.get(key, default))

    def to_json(self):
        return {"name": self.name, "args": {key: getattr(self, key) for key, _ in self._FIELDS}}

    def sample(self, conditioning=None, **kwargs):
        """diffusion.py:84-111.  Two repairs: the reference overwrites a caller's `x_T` with its (always None) start code and
        raises UnboundLocalError when `use_start_code` is set (`start_code` is read before assignment, :91-93); here a given
        `x_T` is honoured and `use_start_code` draws one."""
        # shape is [C, width // 8, height // 8] - W before H, as diffusion.py:89 has it
        shape = [self.z_channels, self.width // 8, self.height // 8]
        kwargs["unconditional_guidance_scale"] = self.scale
        kwargs["eta"] = self.eta
        kwargs["temperature"] = self.temperature
        if self.use_start_code and kwargs.get("x_T") is None:
            import torch
            kwargs["x_T"] = torch.randn([self.batch_size] + shape)
        kwargs.setdefault("x_T", None)
        result = self.sampler.sample(steps=self.steps, conditioning=conditioning, batch_size=self.batch_size,
                                     shape=shape, **kwargs)
        return result[0] if isinstance(result, tuple) else result
