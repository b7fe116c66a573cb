"""Registration boundary (reference: embodiedscan/registry.py:11-13, necks/__init__.py:4-6).

If EmbodiedScan/mmengine are importable the module registers itself into ``embodiedscan.registry.MODELS`` (so
``MODELS.build(cfg.model.preshape)`` in ``detectors/sparse_featfusion_grounder_preshape.py:95`` picks it up from the
unchanged ``configs/grounding/*.py``); otherwise into a local registry with the same two calls the reference uses:
``register_module()`` and ``build(dict(type=..., **kwargs))``.
"""
from __future__ import annotations


class Registry:
    """Minimal stand-in for mmengine.Registry: name -> class, ``build`` pops ``type``."""

    def __init__(self, name: str):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force: bool = False, module=None):
        def deco(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force and self.module_dict[key] is not cls:
                raise KeyError(f"{key} is already registered in {self.name}")
            self.module_dict[key] = cls
            return cls
        return deco(module) if module is not None else deco

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg):
        cfg = dict(cfg)
        if "type" not in cfg:
            raise KeyError("cfg must contain the key 'type'")
        typ = cfg.pop("type")
        cls = self.module_dict.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry")
        return cls(**cfg)


def _resolve():
    try:  # the real thing, when the host project is installed
        from embodiedscan.registry import MODELS as ES_MODELS  # type: ignore
        return ES_MODELS, True
    except Exception:
        return Registry("model"), False


MODELS, USING_EMBODIEDSCAN = _resolve()
