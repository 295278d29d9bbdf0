from .prompts import CompositionalConditioning, parse_mask_style, parse_weighted_prompt

__all__ = ["CompositionalConditioning", "parse_mask_style", "parse_weighted_prompt"]
