"""Name -> class plugin registry for samplers (interface of cpd/samplers/registry.py:3-28).

`register(name)` decorates a wrapper class, `make({"name", "args"}, extra_args)` instantiates it as
`cls(name=..., **args)`, `create(name, **kw)` is the keyword form.  Unlike the reference, unknown names
raise KeyError instead of being passed to `eval`.
"""
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
        raise KeyError(f"no sampler registered under {spec['name']!r}; known: {sorted(lookup)}")
    return lookup[spec["name"]](name=spec["name"], **spec_args)


def create(name, **kwargs):
    if not isinstance(name, str):
        raise ValueError(f"`create` needs the name of a registered sampler, got {name!r}")
    return make({"name": name, "args": kwargs})
