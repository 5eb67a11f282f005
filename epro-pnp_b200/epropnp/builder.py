"""Config-driven construction, as the detection variant of the reference uses it
(EPro-PnP-Det/epropnp_det/ops/pnp/builder.py:6-19: `build_pnp`, `build_camera`, `build_cost_fun` over three
registries; its classes are registered with `@PNP.register_module()` etc. and nested `solver=` / `init_solver=`
entries of a config are themselves dicts with a `type` key).

The reference takes `Registry` / `build_from_cfg` from mmcv; this module provides the small part of that contract the
path needs, without the dependency, so detection-style configs

    build_pnp(dict(type='EProPnP4DoF', mc_samples=512, num_iter=4,
                   solver=dict(type='LMSolver', dof=4, num_iter=5,
                               init_solver=dict(type='RSLMSolver', dof=4, num_points=16, num_proposals=64, num_iter=3))))

work unchanged, while instances (the canonical / 6DoF style) are passed through as they are."""
import inspect


class Registry:
    """Name -> class table with the decorator interface configs rely on."""

    def __init__(self, name):
        self.name = name
        self._table = {}

    def __len__(self):
        return len(self._table)

    def __contains__(self, key):
        return key in self._table

    def __repr__(self):
        return f"Registry(name={self.name!r}, items={sorted(self._table)})"

    @property
    def module_dict(self):
        return dict(self._table)

    def get(self, key):
        return self._table.get(key)

    def _add(self, cls, name, force):
        if not inspect.isclass(cls):
            raise TypeError(f"only classes can be registered in {self.name!r}, got {type(cls)}")
        for key in ([cls.__name__] if name is None else ([name] if isinstance(name, str) else list(name))):
            if key in self._table and not force:
                raise KeyError(f"{key!r} is already registered in {self.name!r}")
            self._table[key] = cls

    def register_module(self, name=None, force=False, module=None):
        """`@R.register_module()`, `@R.register_module(name='alias')` or `R.register_module(module=Cls)`."""
        if module is not None:
            self._add(module, name, force)
            return module

        def decorate(cls):
            self._add(cls, name, force)
            return cls
        return decorate

    def build(self, cfg, **default_args):
        return build_from_cfg(cfg, self, default_args)


def build_from_cfg(cfg, registry, default_args=None):
    """Instantiate `cfg['type']` (a registered name or a class) with the remaining keys; `default_args` fill in keys the
    config does not set.  Anything that is not a dict is returned unchanged, so callers may hand over ready objects."""
    if cfg is None or not isinstance(cfg, dict):
        return cfg
    if "type" not in cfg and not (default_args and "type" in default_args):
        raise KeyError(f"a config for {registry.name!r} needs a 'type' key, got {sorted(cfg)}")
    args = dict(cfg)
    for k, v in (default_args or {}).items():
        args.setdefault(k, v)
    kind = args.pop("type")
    if isinstance(kind, str):
        cls = registry.get(kind)
        if cls is None:
            raise KeyError(f"{kind!r} is not in the {registry.name!r} registry ({sorted(registry.module_dict)})")
    elif inspect.isclass(kind):
        cls = kind
    else:
        raise TypeError(f"'type' must be a name or a class, got {type(kind)}")
    try:
        return cls(**args)
    except Exception as e:
        raise type(e)(f"{cls.__name__}: {e}")


PNP = Registry("pnp")
CAMERA = Registry("camera")
COSTFUN = Registry("cost_fun")


def build_pnp(cfg, **default_args):
    return build_from_cfg(cfg, PNP, default_args)


def build_camera(cfg, **default_args):
    return build_from_cfg(cfg, CAMERA, default_args)


def build_cost_fun(cfg, **default_args):
    return build_from_cfg(cfg, COSTFUN, default_args)
