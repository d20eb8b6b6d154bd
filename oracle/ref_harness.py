"""TEST INFRASTRUCTURE ONLY -- container-side loader for the *real* reference.

This module imports the reference's own Python modules from ``/root/reference``
(read-only, present only in the build container, never on the GPU box) so that

  * ``oracle/make_golden.py`` can freeze golden input/output vectors under
    ``tests/golden/`` (the reference ships none), and
  * ``tests/test_oracle_vs_reference.py`` can pin the torch restatement in
    ``oracle/sph_oracle.py`` against the reference itself whenever the
    reference tree is present.

The reference cannot be imported as shipped: ``sphdet/iou/sph_iou_api.py:2``
imports ``mmcv.ops`` and ``sphdet/iou/sph_iou_calculator.py:1`` imports the
mmdet registry, neither of which is installable offline.  We therefore inject
stub modules:

  * ``mmcv.ops.box_iou_rotated`` / ``diff_iou_rotated_2d``  ->  the reference's
    OWN vendored pure-torch implementation
    ``sphdet/iou/diff_iou_rotated.py:325-343`` (what the authors call the
    bug-fixed version of the mmcv op).  The mmcv CUDA kernel itself is absent:
    parity at that exact boundary is UNPINNED.
  * registries -> pass-through decorators.

Nothing in the product package may import this file.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SPH_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "sphdet", "iou", "sph_iou_api.py"))


class _PassThroughRegistry:
    def __init__(self, name):
        self.name = name
        self.modules = {}

    def register_module(self, *args, **kwargs):
        def deco(cls):
            self.modules[cls.__name__] = cls
            return cls
        if len(args) == 1 and callable(args[0]) and not kwargs:
            return deco(args[0])
        return deco


def _load_file(modname: str, relpath: str):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


_LOADED = None


def load_reference():
    """Returns a namespace with the reference's hot-path callables."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import torch

    # the vendored rotated IoU is free of third-party imports: load it first, by path
    diff_mod = _load_file("_ref_diff_iou_rotated", "sphdet/iou/diff_iou_rotated.py")

    def box_iou_rotated(b1, b2, mode="iou", aligned=False, clockwise=True):
        # stand-in for mmcv.ops.box_iou_rotated (call site sph_iou_api.py:79 passes aligned=True, clockwise=True;
        # naive_iou, :194, passes (b1, b2, mode, is_aligned))
        assert clockwise
        if not aligned:
            R, C = b1.size(0), b2.size(0)
            return box_iou_rotated(b1.repeat_interleave(C, 0), b2.repeat(R, 1), mode, True).view(R, C)
        corners1 = diff_mod.box2corners(b1.unsqueeze(0))
        corners2 = diff_mod.box2corners(b2.unsqueeze(0))
        inter, _ = diff_mod.oriented_box_intersection_2d(corners1, corners2)
        inter = inter.squeeze(0)
        a1 = b1[:, 2] * b1[:, 3]
        a2 = b2[:, 2] * b2[:, 3]
        if mode == "iou":
            return inter / (a1 + a2 - inter)
        return inter / a1

    def _unavailable(*a, **k):
        raise NotImplementedError("planar mmcv op not available in the oracle harness")

    mmcv = types.ModuleType("mmcv")
    mmcv.__path__ = []
    mmcv_ops = types.ModuleType("mmcv.ops")
    mmcv_ops.box_iou_rotated = box_iou_rotated
    mmcv_ops.diff_iou_rotated_2d = diff_mod.diff_iou_rotated_2d
    # mmcv.ops.bbox_overlaps (planar, used by naive_iou for BFoV): the restatement of the published mmcv 1.6.0 kernel
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import sph_oracle as _so
    mmcv_ops.bbox_overlaps = _so.mmcv_bbox_overlaps
    mmcv_ops.batched_nms = _so.mmcv_batched_nms      # published mmcv 1.6.0 algorithm restated (unpinned at that boundary)
    mmcv.ops = mmcv_ops
    mmcv.jit = lambda *a, **k: (lambda f: f)
    sys.modules.setdefault("mmcv", mmcv)
    sys.modules.setdefault("mmcv.ops", mmcv_ops)

    reg = _PassThroughRegistry("IoU calculator")
    for name in ("mmdet", "mmdet.core", "mmdet.core.bbox", "mmdet.core.bbox.iou_calculators"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    builder = types.ModuleType("mmdet.core.bbox.iou_calculators.builder")
    builder.IOU_CALCULATORS = reg
    sys.modules["mmdet.core.bbox.iou_calculators.builder"] = builder

    # import the reference package tree (sphdet/__init__ pulls heavy deps: bypass it)
    sphdet = types.ModuleType("sphdet")
    sphdet.__path__ = [os.path.join(REFERENCE_ROOT, "sphdet")]
    sys.modules["sphdet"] = sphdet
    bbox = types.ModuleType("sphdet.bbox")
    bbox.__path__ = [os.path.join(REFERENCE_ROOT, "sphdet", "bbox")]
    sys.modules["sphdet.bbox"] = bbox
    # box_formator imports scipy-side helpers only lazily enough for our use
    try:
        box_formator = importlib.import_module("sphdet.bbox.box_formator")
    except Exception:  # pragma: no cover - kent helpers missing deps
        box_formator = types.ModuleType("sphdet.bbox.box_formator")
        box_formator.Sph2PlanarBoxTransform = None
        box_formator.Planar2KentTransform = None
        sys.modules["sphdet.bbox.box_formator"] = box_formator
    iou_pkg = types.ModuleType("sphdet.iou")
    iou_pkg.__path__ = [os.path.join(REFERENCE_ROOT, "sphdet", "iou")]
    sys.modules["sphdet.iou"] = iou_pkg
    try:
        importlib.import_module("sphdet.iou.kent_iou_calculator")
    except Exception:
        kent = types.ModuleType("sphdet.iou.kent_iou_calculator")
        kent.kent_iou_calculator = _unavailable
        sys.modules["sphdet.iou.kent_iou_calculator"] = kent
    api = importlib.import_module("sphdet.iou.sph_iou_api")
    calc = importlib.import_module("sphdet.iou.sph_iou_calculator")
    for n in ("unbiased_iou", "sph2pob_standard_iou", "sph2pob_legacy_iou", "sph2pob_efficient_iou",
              "naive_iou", "fov_iou", "sph_iou"):
        setattr(iou_pkg, n, getattr(api, n))
    iou_pkg.SphOverlaps2D = calc.SphOverlaps2D
    iou_pkg.sph_overlaps = calc.sph_overlaps
    std = importlib.import_module("sphdet.iou.sph2pob_standard")
    eff = importlib.import_module("sphdet.iou.sph2pob_efficient")

    # NMS (sphdet/bbox/nms/__init__ also imports the planar NMS, which needs mmcv.ops.batched_nms: stubbed)
    nms_pkg = types.ModuleType("sphdet.bbox.nms")
    nms_pkg.__path__ = [os.path.join(REFERENCE_ROOT, "sphdet", "bbox", "nms")]
    sys.modules["sphdet.bbox.nms"] = nms_pkg
    sph_nms = importlib.import_module("sphdet.bbox.nms.sph_nms")
    try:
        planar_nms = importlib.import_module("sphdet.bbox.nms.planar_nms")
    except Exception:
        planar_nms = None
    try:
        nms_utils = importlib.import_module("sphdet.bbox.nms.utils")       # multiclass_nms (R-CNN heads), utils.py:6-95
    except Exception:
        nms_utils = None

    # losses: exec obb_iou_loss / OBBIoULoss with the in-tree weighted_loss (mmdet/models/losses/utils.py)
    losses_utils = _load_mmdet_loss_utils()
    mm_models = types.ModuleType("mmdet.models")
    mm_models.__path__ = []
    mm_builder = types.ModuleType("mmdet.models.builder")
    mm_builder.LOSSES = _PassThroughRegistry("loss")
    mm_losses = types.ModuleType("mmdet.models.losses")
    mm_losses.__path__ = []
    mm_losses.weighted_loss = losses_utils.weighted_loss
    sys.modules["mmdet.models"] = mm_models
    sys.modules["mmdet.models.builder"] = mm_builder
    sys.modules["mmdet.models.losses"] = mm_losses
    sys.modules["mmdet.models.losses.utils"] = losses_utils
    # the reference's vendored mmdet L1Loss (mmdet/models/losses/smooth_l1_loss.py:107-146), loaded under its package
    # name so that its relative imports (..builder, .utils) resolve to the modules above
    smooth_l1 = _load_file("mmdet.models.losses.smooth_l1_loss", "mmdet/models/losses/smooth_l1_loss.py")
    mm_losses.L1Loss = smooth_l1.L1Loss
    mmrot = types.ModuleType("mmrotate")
    mmrot.__path__ = []
    mmrot_models = types.ModuleType("mmrotate.models")
    mmrot_models.__path__ = []
    mmrot_losses = types.ModuleType("mmrotate.models.losses")

    class _Dummy(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, *a, **k):
            raise NotImplementedError
    mmrot_losses.RotatedIoULoss = _Dummy
    # mmrotate 0.3.2 is absent: GDLoss / KFLoss are the restatement in oracle/mmrotate_losses.py (parity UNPINNED at that
    # boundary); the reference's own subclasses and its Sph2PobTransfrom decorator run on top of them unchanged
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import mmrotate_losses as _mmr
    mmrot_losses.GDLoss = _mmr.GDLoss
    mmrot_losses.RotatedIoULoss = _mmr.RotatedIoULoss
    mmrot_losses.KFLoss = _mmr.KFLoss
    sys.modules["mmrotate"] = mmrot
    sys.modules["mmrotate.models"] = mmrot_models
    sys.modules["mmrotate.models.losses"] = mmrot_losses
    losses_pkg = types.ModuleType("sphdet.losses")
    losses_pkg.__path__ = [os.path.join(REFERENCE_ROOT, "sphdet", "losses")]
    sys.modules["sphdet.losses"] = losses_pkg
    iou_loss = importlib.import_module("sphdet.losses.sph2pob_iou_loss")
    transform = importlib.import_module("sphdet.losses.sph2pob_transform")
    gd_loss = importlib.import_module("sphdet.losses.sph2pob_gd_loss")
    kf_loss = importlib.import_module("sphdet.losses.sph2pob_kf_loss")
    l1_loss = importlib.import_module("sphdet.losses.sph2pob_l1_loss")
    # sph2pob_l1_loss.py:25 leaves a pdb.set_trace() in the constructor: make it a no-op for this module only
    l1_loss.pdb = types.SimpleNamespace(set_trace=lambda *a, **k: None)

    # bbox coders (sphdet/bbox/coder/delta_xywh*_sph_bbox_coder.py): need only the mmdet base class and registry
    for name in ("mmdet.core.bbox.coder",):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    base_coder = types.ModuleType("mmdet.core.bbox.coder.base_bbox_coder")

    class BaseBBoxCoder:                                  # mmdet/core/bbox/coder/base_bbox_coder.py: an ABC with no state
        def __init__(self, **kwargs):
            pass
    base_coder.BaseBBoxCoder = BaseBBoxCoder
    sys.modules["mmdet.core.bbox.coder.base_bbox_coder"] = base_coder
    bbox_builder = types.ModuleType("mmdet.core.bbox.builder")
    bbox_builder.BBOX_CODERS = _PassThroughRegistry("bbox_coder")
    sys.modules["mmdet.core.bbox.builder"] = bbox_builder
    coder4 = _load_file("_ref_coder_bfov", "sphdet/bbox/coder/delta_xywh_sph_bbox_coder.py")
    coder5 = _load_file("_ref_coder_rbfov", "sphdet/bbox/coder/delta_xywha_rsph_bbox_coder.py")

    # the anchor-free heads' coder imports ..box_formator relatively: load it under its package name (the package's
    # __init__ pulls in the Kent coder and scipy, so the file is loaded by path)
    dist_coder = None
    if getattr(box_formator, "__file__", None):
        dist_coder = _load_file("sphdet.bbox.coder.distance_point_sph_bbox_coder",
                                "sphdet/bbox/coder/distance_point_sph_bbox_coder.py")

    gen = _load_file("_ref_generate_data", "tests/utils/generate_data.py")

    ns = types.SimpleNamespace(
        api=api, calc=calc, std=std, eff=eff, diff=diff_mod, nms=sph_nms,
        iou_loss=iou_loss, transform=transform, box_formator=box_formator,
        generate_boxes=gen.generate_boxes,
        sph2pob_efficient_iou=api.sph2pob_efficient_iou,
        sph2pob_standard_iou=api.sph2pob_standard_iou,
        sph2pob_legacy_iou=api.sph2pob_legacy_iou,
        sph_iou=api.sph_iou, fov_iou=api.fov_iou, naive_iou=api.naive_iou, unbiased_iou=api.unbiased_iou,
        SphOverlaps2D=calc.SphOverlaps2D, SphNMS=sph_nms.SphNMS, PlanarNMS=getattr(planar_nms, 'PlanarNMS', None),
        multiclass_nms=getattr(nms_utils, 'multiclass_nms', None),
        Sph2PobIoULoss=iou_loss.Sph2PobIoULoss, SphIoULossLegacy=iou_loss.SphIoULossLegacy,
        Sph2PobGDLoss=gd_loss.Sph2PobGDLoss, Sph2PobKFLoss=kf_loss.Sph2PobKFLoss, Sph2PobL1Loss=l1_loss.Sph2PobL1Loss,
        jiter_spherical_bboxes=api.jiter_spherical_bboxes,
        jiter_rotated_bboxes=api.jiter_rotated_bboxes,
        DeltaXYWHSphBBoxCoder=coder4.DeltaXYWHSphBBoxCoder, DeltaXYWHASphBBoxCoder=coder5.DeltaXYWHASphBBoxCoder,
        coder4=coder4, coder5=coder5, dist_coder=dist_coder,
        DistancePointSphBBoxCoder=getattr(dist_coder, 'DistancePointSphBBoxCoder', None),
    )
    _LOADED = ns
    return ns


def _load_mmdet_loss_utils():
    """mmdet/models/losses/utils.py with its ``mmcv.jit`` decorator neutralised."""
    return _load_file("_ref_mmdet_loss_utils", "mmdet/models/losses/utils.py")


class float64_mode:
    """Run the reference in double precision ("the truth"): the
    reference allocates some temporaries with the default dtype
    (sph2pob_standard.py:293), hence the global switch."""

    def __enter__(self):
        import torch
        self._old = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)
        return self

    def __exit__(self, *exc):
        import torch
        torch.set_default_dtype(self._old)
        return False
