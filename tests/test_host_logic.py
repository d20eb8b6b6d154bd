"""Host-side logic that needs no GPU: registries, the drop-in signatures, loud failure on CPU tensors,
key packing and shard arithmetic of the multi-GPU path."""
import inspect

import numpy as np
import pytest
import torch


def test_registry_drop_in_names():
    from sph_retina_b200.sphdet import registry
    from sph_retina_b200.sphdet.iou import SphOverlaps2D
    from sph_retina_b200.sphdet.losses import Sph2PobIoULoss
    calc = registry.build_iou_calculator(dict(type='SphOverlaps2D', backend='sph2pob_efficient_iou', box_version=5))
    assert isinstance(calc, SphOverlaps2D) and calc.backend == 'sph2pob_efficient_iou' and calc.box_version == 5
    loss = registry.build_loss(dict(type='Sph2PobIoULoss', mode='ciou', loss_weight=2.0))
    assert isinstance(loss, Sph2PobIoULoss) and loss.mode == 'ciou' and loss.loss_weight == 2.0
    assert repr(calc) == 'SphOverlaps2D()'


def test_signatures_match_the_reference():
    """Argument names/defaults of sphdet/iou/sph_iou_api.py:94-98,130,156, sph_iou_calculator.py:12-20,58,
    sph2pob_iou_loss.py:17-31 and sph_nms.py:8,18."""
    from sph_retina_b200.sphdet.bbox.nms import SphNMS
    from sph_retina_b200.sphdet.iou import (SphOverlaps2D, fov_iou, sph2pob_efficient_iou, sph2pob_standard_iou,
                                            sph_iou, sph_overlaps)
    from sph_retina_b200.sphdet.losses import Sph2PobIoULoss

    def sig(f):
        return [(p.name, p.default) for p in inspect.signature(f).parameters.values()]
    E = inspect.Parameter.empty
    want = [('bboxes1', E), ('bboxes2', E), ('mode', 'iou'), ('is_aligned', False), ('calculator', 'common'),
            ('rbb_edge', 'arc'), ('rbb_angle', 'equator')]
    assert sig(sph2pob_efficient_iou) == want and sig(sph2pob_standard_iou) == want
    want = [('bboxes1', E), ('bboxes2', E), ('mode', 'iou'), ('is_aligned', False), ('calculator', 'diff')]
    assert sig(sph_iou) == want and sig(fov_iou) == want
    assert sig(sph_overlaps) == [('bboxes1', E), ('bboxes2', E), ('mode', 'iou'), ('is_aligned', False),
                                 ('backend', 'unbiased_iou')]
    assert sig(SphOverlaps2D.__init__)[1:] == [('backend', 'unbiased_iou'), ('box_version', 4)]
    assert sig(SphOverlaps2D.__call__)[1:] == [('bboxes1', E), ('bboxes2', E), ('mode', 'iou'), ('is_aligned', False)]
    assert sig(Sph2PobIoULoss.__init__)[1:] == [('mode', 'iou'), ('eps', 1e-6), ('reduction', 'mean'), ('loss_weight', 1.0)]
    assert [n for n, _ in sig(Sph2PobIoULoss.forward)][1:6] == ['pred', 'target', 'weight', 'avg_factor', 'reduction_override']
    assert sig(SphNMS.__init__)[1:] == [('iou_calculator', 'sph2pob_efficient')]
    assert sig(SphNMS.__call__)[1:] == [('boxes', E), ('scores', E), ('idxs', E), ('nms_cfg', E), ('class_agnostic', False)]


def test_empty_inputs_and_errors_follow_the_reference():
    from sph_retina_b200.sphdet.iou import SphOverlaps2D, sph2pob_efficient_iou, sph_overlaps
    e, b = torch.zeros(0, 4), torch.rand(3, 4)
    assert sph2pob_efficient_iou(e, b).shape == (0, 3)                  # sph_iou_api.py:56-57
    assert sph2pob_efficient_iou(b, e).shape == (3, 0)
    assert sph2pob_efficient_iou(e, e, is_aligned=True).shape == (0, 1)
    assert SphOverlaps2D('sph2pob_efficient_iou')(e, b).shape == (0, 3)
    with pytest.raises(AssertionError):
        sph_overlaps(b, b, mode='giou', backend='sph2pob_efficient_iou')   # :75
    with pytest.raises(AssertionError):
        sph_overlaps(b, b, backend='no_such_iou')                          # :76
    with pytest.raises(AssertionError):
        SphOverlaps2D('sph2pob_efficient_iou')(torch.rand(3, 7), b)        # :41
    with pytest.raises(AssertionError):
        sph2pob_efficient_iou(b, b, rbb_edge='diagonal')                   # sph_iou_api.py:51
    with pytest.raises(NotImplementedError):
        sph_overlaps(b, b, backend='kent_iou')                             # outside the path: refused, not faked
    with pytest.raises(ValueError):
        sph_overlaps(torch.zeros(2, 5), torch.zeros(2, 5), backend='sph2pob_legacy_iou')   # BFoV only (sph2pob_legacy.py:52-53)


def test_no_cpu_fallback():
    """A CPU tensor must raise -- never silently compute somewhere else."""
    from sph_retina_b200 import _native
    from sph_retina_b200.sphdet.bbox.nms import SphNMS
    from sph_retina_b200.sphdet.iou import sph2pob_efficient_iou, sph_iou
    from sph_retina_b200.sphdet.losses import Sph2PobIoULoss
    b = torch.rand(8, 4) * 50 + 10
    with pytest.raises(_native.SphkError):
        sph2pob_efficient_iou(b, b)
    with pytest.raises(_native.SphkError):
        sph_iou(b, b, is_aligned=True)
    with pytest.raises(_native.SphkError):
        Sph2PobIoULoss()(b.clone().requires_grad_(True), b)
    with pytest.raises(_native.SphkError):
        SphNMS()(b, torch.rand(8), torch.zeros(8, dtype=torch.long), dict(iou_threshold=0.5))


def test_product_never_imports_the_oracle():
    import os
    import re
    from conftest import ROOT
    pkg = os.path.join(ROOT, "sph_retina_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|sph_oracle|ref_harness)", text, flags=re.M), f
                assert "oracle/" not in text.replace("the oracle", ""), f


def test_key_packing_orders_by_value_then_lowest_index():
    from sph_retina_b200.sharded import pack_keys
    from test_sharded_gloo import unpack_gathered_reference
    v = torch.tensor([0.0, 0.5, 0.5, 1.0, 0.25])
    i = torch.tensor([7, 9, 3, 4000000000, 0])
    k = pack_keys(v, i)
    assert k.dtype == torch.int64 and bool((k >= 0).all())
    vv, ii, _, _ = unpack_gathered_reference(k, 1, 5, 0, 5)
    assert torch.equal(vv, v) and ii.tolist() == i.tolist()
    assert k[2] > k[1] > k[4] > k[0] and k[3] == k.max()      # ties -> lowest index wins
    # two ranks' views of one ground truth (block = [0 anchor slots | 1 GT key]): the maximum is (0.5, index 3)
    _, _, vals, idx = unpack_gathered_reference(torch.stack([k[1:2], k[2:3]]), 2, 0, 1, 0)
    assert vals.item() == 0.5 and idx.item() == 3


def test_block_layout_of_the_sharded_exchange():
    """key_block / block_capacity / exchange_blocks without a process group (world = 1) and the layout arithmetic."""
    from sph_retina_b200.sharded import block_capacity, exchange_blocks, key_block, shard_bounds
    for n, w in ((0, 1), (1, 3), (7, 2), (1 << 20, 8), (1000003, 8)):
        cap = block_capacity(n, w)
        assert cap == max(shard_bounds(n, w, r)[1] - shard_bounds(n, w, r)[0] for r in range(w))
    blk = key_block(10, 4, 1, "cpu", fresh=True)
    assert blk.shape == (14,) and blk.dtype == torch.int64 and not blk.any()
    assert key_block(10, 4, 1, "cpu") is key_block(10, 4, 1, "cpu")          # cached per (device, shape)
    g = exchange_blocks(blk)
    assert g.shape == (1, 14) and g.data_ptr() == blk.data_ptr()             # world = 1: no copy, no collective


def test_shard_bounds_cover_and_balance():
    from sph_retina_b200.sharded import shard_bounds
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[r][1] == b[r + 1][0] for r in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_call_validates_its_arguments_before_any_launch():
    """sharded_max_overlaps refuses a shard that is not the rank's contiguous slice and an unknown exchange route; the
    push route's partial-array count follows the kernel's 256-column tiles (one key array per column tile)."""
    from sph_retina_b200 import _native
    from sph_retina_b200.sharded import MAX_PUSH_PARTS, sharded_max_overlaps
    a, g = torch.zeros(10, 5), torch.zeros(4, 5)
    with pytest.raises(ValueError):
        sharded_max_overlaps(a, g, 10, 3)                          # world = 1: the shard must start at row 0
    with pytest.raises(ValueError):
        sharded_max_overlaps(a, g, 12, 0)                          # ... and hold all rows
    with pytest.raises(ValueError):
        sharded_max_overlaps(a, g, 10, 0, exchange='smoke-signals')
    assert [_native.key_push_parts(c) for c in (0, 1, 256, 257, 1024, 2048, 2049)] == [1, 1, 1, 2, 4, 8, 9]
    assert MAX_PUSH_PARTS == 8


def test_weight_reduce_matches_mmdet_rules():
    from sph_retina_b200.sphdet.losses.sph2pob_iou_loss import _weight_reduce_loss
    loss, w = torch.tensor([1.0, 2.0, 3.0]), torch.tensor([1.0, 0.0, 1.0])
    assert _weight_reduce_loss(loss).item() == 2.0
    assert _weight_reduce_loss(loss, w).item() == pytest.approx(4 / 3)
    assert _weight_reduce_loss(loss, w, 'sum').item() == 4.0
    assert _weight_reduce_loss(loss, w, 'mean', avg_factor=2).item() == pytest.approx(2.0, rel=1e-6)
    assert _weight_reduce_loss(loss, w, 'none', avg_factor=2).tolist() == [1.0, 0.0, 3.0]
    with pytest.raises(ValueError):
        _weight_reduce_loss(loss, w, 'sum', avg_factor=2)


def test_vectorised_assign_wrt_overlaps_equals_the_reference_loop():
    """SphMaxIoUAssigner.assign_wrt_overlaps (no Python loop) against the literal restatement of
    mmdet/core/bbox/assigners/max_iou_assigner.py:135-220, on matrices full of ties and zeros."""
    import os
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sph_oracle as O
    from sph_retina_b200.sphdet.assigners import SphMaxIoUAssigner
    torch.manual_seed(0)
    for trial in range(40):
        K, N = int(torch.randint(0, 12, (1,))), int(torch.randint(0, 300, (1,)))
        ov = (torch.rand(K, N) * torch.rand(K, N)).round(decimals=2)
        ov[torch.rand(K, N) < 0.5] = 0
        if trial % 5 == 0 and N:
            ov[:, torch.rand(N) < 0.2] = -1          # an ignore region
        lab = torch.randint(0, 5, (K,))
        for all_ in (True, False):
            for mlq in (True, False):
                for neg in (0.4, (0.1, 0.4)):
                    a = SphMaxIoUAssigner(0.5, neg, min_pos_iou=0.0 if trial % 2 else 0.2, gt_max_assign_all=all_,
                                          match_low_quality=mlq)
                    r = a.assign_wrt_overlaps(ov, lab)
                    g, m, l = O.assign_wrt_overlaps(ov, lab, 0.5, neg, a.min_pos_iou, all_, mlq)
                    assert r.num_gts == K and torch.equal(r.gt_inds, g) and torch.equal(r.max_overlaps, m)
                    assert torch.equal(r.labels, l)


def test_naive_iou_and_nms_calculators_host_contract():
    """sph_iou_api.py:181-198 / sph_nms.py:8-16: names, defaults and refusals."""
    from sph_retina_b200.sphdet.bbox.nms import SphNMS
    from sph_retina_b200.sphdet.iou import SphOverlaps2D, naive_iou
    E = inspect.Parameter.empty
    assert [(p.name, p.default) for p in inspect.signature(naive_iou).parameters.values()] == [
        ('bboxes1', E), ('bboxes2', E), ('mode', 'iou'), ('is_aligned', False), ('box_formator', 'sph2pix')]
    b = torch.rand(3, 4) * 50 + 10
    with pytest.raises(AssertionError):
        naive_iou(b, b, mode='iof')
    with pytest.raises(NotImplementedError):
        naive_iou(b, b, box_formator='sph2tan')
    assert naive_iou(b[:0], b).shape == (0, 3) and SphOverlaps2D('naive_iou')(b, b[:0]).shape == (3, 0)
    assert SphNMS('naive_iou').iou_calculator == 'naive_iou' and SphNMS().iou_calculator == 'sph2pob_efficient'
    assert SphNMS('unbiased_iou').iou_calculator == 'unbiased_iou'
    with pytest.raises(NotImplementedError):
        SphNMS('planar')


def test_other_sph2pob_losses_host_contract():
    """Sph2PobGDLoss / Sph2PobKFLoss / Sph2PobL1Loss (sphdet/losses/__init__.py:3-8): registry names, constructor
    checks of the mmrotate / reference classes, and the loud failure on CPU tensors."""
    from sph_retina_b200 import _native
    from sph_retina_b200.sphdet import registry
    from sph_retina_b200.sphdet.losses import Sph2PobGDLoss, Sph2PobKFLoss, Sph2PobL1Loss
    gd = registry.build_loss(dict(type='Sph2PobGDLoss', loss_type='kld', fun='log1p', tau=1.0, loss_weight=5.0, sqrt=False))
    assert isinstance(gd, Sph2PobGDLoss) and gd.tau == 1.0 and gd.kwargs == dict(sqrt=False) and gd.reduction == 'mean'
    assert isinstance(registry.build_loss(dict(type='Sph2PobKFLoss', fun='ln')), Sph2PobKFLoss)
    l1 = registry.build_loss(dict(type='Sph2PobL1Loss', angle_modifier='modulus', swap=True))
    assert isinstance(l1, Sph2PobL1Loss) and l1.encode and l1.swap
    for bad in (dict(loss_type='bogus'), dict(loss_type='gwd', fun='exp'), dict(loss_type='gwd', reduction='avg')):
        with pytest.raises(AssertionError):
            Sph2PobGDLoss(**bad)
    with pytest.raises(AssertionError):
        Sph2PobKFLoss(fun='log1p')
    with pytest.raises(AssertionError):
        Sph2PobL1Loss(angle_modifier='wrap')
    with pytest.raises(TypeError):                      # gwd_loss() takes `normalize`, not `sqrt`
        Sph2PobGDLoss('gwd', sqrt=True)(torch.zeros(0, 4), torch.zeros(0, 4))
    b = torch.rand(4, 5) * 50 + 10
    for loss in (Sph2PobGDLoss('gwd'), Sph2PobKFLoss(), Sph2PobL1Loss()):
        with pytest.raises(_native.SphkError):
            loss(b, b.clone())
        with pytest.raises(AssertionError):
            loss(b, b.clone(), reduction_override='avg')


def test_legacy_iou_loss_host_contract():
    """SphIoULossLegacy: constructor of mmrotate's RotatedIoULoss (linear overrides mode), registered in LOSSES."""
    from sph_retina_b200.sphdet.losses import SphIoULossLegacy
    from sph_retina_b200.sphdet.registry import LOSSES
    L = SphIoULossLegacy()
    assert (L.mode, L.eps, L.reduction, L.loss_weight) == ('log', 1e-6, 'mean', 1.0)
    assert SphIoULossLegacy(linear=True, mode='square').mode == 'linear'
    with pytest.raises(AssertionError):
        SphIoULossLegacy(mode='iou')
    assert LOSSES.get('SphIoULossLegacy') is SphIoULossLegacy
    with pytest.raises(AssertionError):
        L(torch.zeros(2, 4), torch.zeros(2, 4), reduction_override='max')


def test_planar_nms_host_contract():
    from sph_retina_b200.sphdet.bbox.nms import PlanarNMS
    assert PlanarNMS().box_formator == 'sph2pix'
    with pytest.raises(NotImplementedError):
        PlanarNMS('sph2tan')
    b = torch.rand(4, 5)
    with pytest.raises(NotImplementedError):
        PlanarNMS()(b, torch.rand(4), torch.zeros(4, dtype=torch.long), dict(type='nms', iou_threshold=0.5))      # RBFoV
    with pytest.raises(NotImplementedError):
        PlanarNMS()(b[:, :4], torch.rand(4), torch.zeros(4, dtype=torch.long), dict(type='soft_nms'))
    dets, keep = PlanarNMS()(b[:, :4], torch.tensor([0.1, 0.9, 0.5, 0.3]), torch.zeros(4, dtype=torch.long), None)
    assert keep.tolist() == [1, 2, 3, 0] and dets.shape == (4, 5)


def test_sph_l1_loss_alias():
    """sphdet/losses/__init__.py:1: SphL1Loss = mmdet's L1Loss (|pred - target|, weight, avg_factor, reductions)."""
    from sph_retina_b200.sphdet.losses import SphL1Loss
    torch.manual_seed(0)
    p, t, w = torch.randn(7, 4, requires_grad=True), torch.randn(7, 4), torch.rand(7, 4)
    assert torch.allclose(SphL1Loss()(p, t), (p - t).abs().mean())
    assert torch.allclose(SphL1Loss(reduction='sum', loss_weight=2.0)(p, t, w), 2.0 * ((p - t).abs() * w).sum())
    assert torch.allclose(SphL1Loss()(p, t, w, avg_factor=3.0), ((p - t).abs() * w).sum() / (3.0 + torch.finfo(torch.float32).eps))
    assert SphL1Loss()(p, t, reduction_override='none').shape == (7, 4)
    assert SphL1Loss()(p[:0], t[:0]).item() == 0.0
    SphL1Loss()(p, t, w).backward()
    assert torch.allclose(p.grad, torch.sign(p.detach() - t) * w / 28)
