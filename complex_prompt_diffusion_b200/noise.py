"""NoiseGenerator: seeded noise inputs for the denoising loop (interface of cpd/noise.py:12-96).

Only the parts the hot path uses are mirrored: the `seed` property with its seed modes (iter / constant / loop /
random, noise.py:34-46) and `sample(seed=None)` = `torch.manual_seed(seed); torch.randn(shape, device)`
(noise.py:86-93).  Note the reference's "iter" mode increments BEFORE use, so the first draw uses seed0 + 1.
Histogram matching / exemplar sequences (skimage) are image-space utilities outside the path.
"""
import random

import torch


def build_cycle_mod(n=5):
    up = list(range(1, n))
    return up + [-v for v in up][::-1]


class NoiseGenerator:
    def __init__(self, shape, device, seed=0, torch_generator=None, seed_mode="iter", cycle_size=5, logger=print):
        self._log = logger
        self._seed = seed
        self.seed_mode = seed_mode
        self.generator = torch_generator
        self.shape, self.device = tuple(shape), device
        self._seed_list = build_cycle_mod(n=cycle_size)
        self._seed_idx = 0

    @property
    def seed(self):
        mode = self.seed_mode
        if mode == "iter":
            self._seed += 1
        elif mode in ("constant", "const", "c"):
            pass
        elif mode in ("loop", "l"):
            self._seed = self._seed_list[self._seed_idx % len(self._seed_list)]
        else:
            self._seed = random.randint(0, 10000)
        return self._seed

    @property
    def last_seed(self):
        return self._seed

    def sample(self, seed=None, match_noise=None):
        if match_noise is not None:
            raise NotImplementedError("histogram matching (skimage) is outside the hot-path scope")
        if seed is None:
            seed = self.seed
        torch.manual_seed(seed)
        return torch.randn(self.shape, device=self.device)

    def sampler(self):
        """A `noise_sampler(x)` for the ancestral samplers that draws successive seeded tensors on the CPU
        (device-independent values) and moves them to x's device."""
        cpu = NoiseGenerator(self.shape, "cpu", self._seed, seed_mode=self.seed_mode)

        def draw(x):
            n = cpu.sample()
            self._seed = cpu._seed
            return n.to(x.device)
        return draw
