"""Registry for denoiser/sampler extensions (interface of cpd/samplers/extension/registry.py).
Thresholding extensions are SURVEY.md 8-f 'next' rows; the registry exists so callers can probe it."""
import copy

lookup = {}


def register(name):
    def decorator(cls):
        lookup[name] = cls
        return cls
    return decorator


def make(spec, args=None):
    spec_args = copy.copy(spec.get("args", {}))
    if args is not None:
        spec_args.update(args)
    if spec["name"] not in lookup:
        raise KeyError(f"no extension registered under {spec['name']!r}; known: {sorted(lookup)}")
    return lookup[spec["name"]](**spec_args)


def create(name, **kwargs):
    if not isinstance(name, str):
        raise ValueError(f"`create` needs the name of a registered extension, got {name!r}")
    return make({"name": name, "args": kwargs})
