"""The conditioning contract of the hot path: how the reference turns weighted / masked sub-prompts
into the dict the Denoiser consumes,

    {"and": [(scale, emb [1, 77, D], guide_emb, mask), ...], "not": [(scale, emb, guide_emb, mask), ...]}

(cpd/embeddings/prompts.py:622-648 `CompositionalPrompt._build_embeddings`; first "and" entry = the base prompt with its own
scale; masks are 1 or uint8 [1, 1, H/8, W/8]).  Host code, once per prompt.  The text encoders are outside the hot path
(SURVEY.md section 8): embeddings come in as tensors, or through any `embedder(text) -> [1, 77, D]` callable.

  parse_weighted_prompt  - WeightedPrompt._parse_prompt (prompts.py:546-589): "a cat:1.5 a dog:0.5 trees" -> sub-prompts / weights
  parse_mask_style       - CompositionalPrompt._parse_mask_style (prompts.py:737-856): "left_third_hidden" -> uint8 mask
  CompositionalConditioning - add_conjunction / add_negation / add_filter / add_masked_filter (prompts.py:668-735) + build()
"""
import math

import torch

_SIZES = {2: ("2", "half"), 3: ("3", "third"), 4: ("4", "quarter", "fourth"), 5: ("5", "fifrth"), 6: ("6", "sixth"),
          7: ("7", "seventh"), 8: ("8", "eigth"), 9: ("9", "ninth"), 10: ("10", "tenth")}  # spellings as registered, prompts.py:739-750
_DIRECTIONS = {"top": ("top", "t", "north"), "bottom": ("bottom", "bot", "b", "south"), "left": ("left", "l", "west"),
               "right": ("right", "r", "east")}
_VALID, _HIDDEN = ("valid", "visible", "show", "v"), ("hidden", "hide", "h")


def parse_weighted_prompt(text):
    """prompts.py:546-589: text up to the first ':' is a sub-prompt, the token after ':' (up to the next space) its weight
    (1.0 when empty or not a number); repeat; a remainder without ':' is a last sub-prompt of weight 1.  Returns
    (prompts, weights)."""
    prompts, weights = [], []
    while text:
        if ":" not in text:
            prompts.append(text)
            weights.append(1.0)
            break
        head, text = text.split(":", 1)
        token, _, text = text.partition(" ")
        try:
            weight = float(token) if token else 1.0
        except ValueError:
            weight = 1.0  # the reference warns ("are you missing a space?") and falls back to 1
        prompts.append(head)
        weights.append(weight)
    return prompts, weights


def parse_mask_style(style, height, width):
    """prompts.py:737-856: "<direction>_<size>_<minority>" -> uint8 mask [1, height // 8, width // 8] (1 = the sub-prompt
    applies).  direction: left / right / top / bottom (and the aliases l, west, ...); size: half .. tenth or 2 .. 10 (default
    half); minority: whether that strip is `valid` (default) or `hidden`.  The strip is floor(n * ratio) wide on the valid side
    and ceil(n * ratio) on the hidden side, in the reference's own float arithmetic; a style whose two parts do not add up to
    the latent size fails there with an assertion and here with ValueError.  "perspective" builds a [h, h] matrix that fails
    the reference's shape assertion (prompts.py:779-781,855): not reproducible, NotImplementedError."""
    if style == "perspective":
        raise NotImplementedError("mask style 'perspective' fails the reference's own shape assertion (prompts.py:855)")
    parts = style.split("_")
    direction = next((k for k, names in _DIRECTIONS.items() if parts[0] in names), None)
    if direction is None:
        raise ValueError(f"mask style {style!r}: unknown direction {parts[0]!r}")
    size = parts[1] if len(parts) > 1 else "half"
    denom = next((k for k, names in _SIZES.items() if size in names), None)
    if denom is None:
        raise ValueError(f"mask style {style!r}: unknown size {size!r}")
    minority = parts[2] if len(parts) > 2 else "valid"
    if minority not in _VALID + _HIDDEN:
        raise ValueError(f"mask style {style!r}: unknown minority {minority!r}")
    minor, major = 1 / denom, (denom - 1) / denom
    strip_valid = minority in _VALID  # the named strip is the valid part
    valid_ratio, hidden_ratio = (minor, major) if strip_valid else (major, minor)
    h, w = height // 8, width // 8
    n = w if direction in ("left", "right") else h
    n_valid, n_hidden = int(math.floor(n * valid_ratio)), int(math.ceil(n * hidden_ratio))
    if n_valid + n_hidden != n:
        raise ValueError(f"mask style {style!r} does not tile {n} latent cells ({n_valid} valid + {n_hidden} hidden)")
    # which part comes first: the smaller part sits on the named side; on a tie the minority keyword decides (:784-853)
    near = direction in ("left", "top")
    if n_valid != n_hidden:
        valid_first = (n_valid < n_hidden) == near
    else:
        valid_first = strip_valid == near
    line = torch.zeros(n, dtype=torch.uint8)
    if valid_first:
        line[:n_valid] = 1
    else:
        line[n_hidden:] = 1
    if direction in ("left", "right"):
        return line.view(1, 1, w).expand(1, h, w).contiguous()
    return line.view(1, h, 1).expand(1, h, w).contiguous()


class CompositionalConditioning:
    """The composition half of CompositionalPrompt (prompts.py:591-735) over ready-made embeddings: a base prompt plus
    conjunctions ("and") and negations ("not"), each with a scale and a mask.  `prompt` arguments are embeddings [1, 77, D] or,
    when an `embedder` callable was given, strings.  build() returns the dict of _build_embeddings (:622-648)."""

    def __init__(self, base, scale=1, mask=1, guide=None, embedder=None, height=512, width=512):
        self.embedder = embedder
        self.height, self.width = height, width
        self.base = (scale, self._embed(base), guide, mask)
        self._conjunctions, self._negations = [], []

    def _embed(self, prompt):
        if isinstance(prompt, str):
            if self.embedder is None:
                raise ValueError("a text prompt needs an `embedder` (the text encoders are outside this package)")
            prompt = self.embedder(prompt)
        if not torch.is_tensor(prompt) or prompt.ndim != 3 or prompt.shape[0] != 1:
            raise ValueError(f"an embedding must be a [1, tokens, D] tensor, got {getattr(prompt, 'shape', type(prompt))}")
        return prompt

    def add_conjunction(self, prompt, scale=1, mask=1, guide=None):
        self._conjunctions.append((1 if scale is None else scale, self._embed(prompt), guide, 1 if mask is None else mask))
        return self

    def add_negation(self, prompt, scale=1, mask=1, guide=None):
        self._negations.append((1 if scale is None else scale, self._embed(prompt), guide, 1 if mask is None else mask))
        return self

    def add_filter(self, prompt, strength=1.0, mask=1):
        """prompts.py:706-712: the sign of `strength` picks conjunction or negation; 0 adds nothing."""
        if strength == 0:
            return self
        if strength > 0:
            return self.add_conjunction(prompt, scale=strength, mask=mask)
        return self.add_negation(prompt, scale=abs(strength), mask=mask)

    def add_masked_filter(self, prompt, mask, strength=1.0):
        """prompts.py:714-735: `mask` is a tensor / array or a "<direction>_<size>_<minority>" style string."""
        if isinstance(mask, str):
            mask = parse_mask_style(mask, self.height, self.width)
        mask = torch.as_tensor(mask)
        if mask.ndim < 4:
            mask = mask.reshape(1, 1, mask.shape[-2], mask.shape[-1])
        return self.add_filter(prompt, strength=strength, mask=mask)

    def add_weighted(self, text):
        """A "sub:weight sub:weight ..." string (WeightedPrompt): every sub-prompt becomes a filter with its weight."""
        for sub, weight in zip(*parse_weighted_prompt(text)):
            self.add_filter(sub, strength=weight)
        return self

    def build(self):
        return {"and": [self.base] + list(self._conjunctions), "not": list(self._negations)}
