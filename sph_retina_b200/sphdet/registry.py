"""Registries the hot-path classes plug into.

With mmdet 2.x importable the real registries are used, so that ``MaxIoUAssigner`` /
``build_loss`` find the classes by name exactly as with the reference
(mmdet/core/bbox/iou_calculators/builder.py:4-9, mmdet/models/builder.py).  Without mmdet (this
image) a small stand-in with the same ``register_module`` / ``build`` surface is provided."""
from __future__ import annotations


class _Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force:
                raise KeyError("%s is already registered in %s" % (key, self.name))
            self.module_dict[key] = cls
            return cls
        if module is not None:
            return deco(module)
        if callable(name) and not isinstance(name, str):   # used as @REG.register_module without ()
            cls, name = name, None
            return deco(cls)
        return deco

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, default_args=None):
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        typ = args.pop("type")
        cls = self.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError("%s is not in the %s registry" % (typ, self.name))
        return cls(**args)


def _resolve(path, attr, fallback_name):
    try:
        mod = __import__(path, fromlist=[attr])
        return getattr(mod, attr), True
    except Exception:
        return _Registry(fallback_name), False


IOU_CALCULATORS, IOU_CALCULATORS_IS_MMDET = _resolve("mmdet.core.bbox.iou_calculators.builder", "IOU_CALCULATORS",
                                                      "IoU calculator")
LOSSES, LOSSES_IS_MMDET = _resolve("mmdet.models.builder", "LOSSES", "loss")
BBOX_CODERS, BBOX_CODERS_IS_MMDET = _resolve("mmdet.core.bbox.builder", "BBOX_CODERS", "bbox_coder")


def build_iou_calculator(cfg, default_args=None):
    """mmdet/core/bbox/iou_calculators/builder.py:7-9."""
    if IOU_CALCULATORS_IS_MMDET:
        from mmcv.utils import build_from_cfg
        return build_from_cfg(cfg, IOU_CALCULATORS, default_args)
    return IOU_CALCULATORS.build(cfg, default_args)


def build_loss(cfg):
    if LOSSES_IS_MMDET:
        from mmdet.models.builder import build_loss as _b
        return _b(cfg)
    return LOSSES.build(cfg)


def build_bbox_coder(cfg, default_args=None):
    """mmdet/core/bbox/builder.py: build_bbox_coder."""
    if BBOX_CODERS_IS_MMDET:
        from mmcv.utils import build_from_cfg
        return build_from_cfg(cfg, BBOX_CODERS, default_args)
    return BBOX_CODERS.build(cfg, default_args)
