"""Row-sharded N x M overlaps for large anchor x GT sweeps (BASELINE config #5).

One process per GPU.  The long (anchor) axis is split contiguously over the ranks, the short (GT)
set is replicated; every rank runs the fused max/argmax kernel on its shard with GLOBAL index
offsets, and NCCL is used only for
  * one all_gather of the packed per-anchor (max, argmax over GT)   -- 8 B per anchor,
  * one all_reduce(MAX) of the packed per-GT (max, argmax over anchors) -- 8 B per GT.
Packing: int64 key = float32 bits << 32 | (0xFFFFFFFF - index).  IoU >= 0, so integer order is
(value, then LOWEST index); MAX over ranks therefore equals the single-device tie rule.

The reference has no multi-GPU awareness on this path; the single-device contract
being reproduced is MaxIoUAssigner's overlaps.max(dim=0/1)
(mmdet/core/bbox/assigners/max_iou_assigner.py:173-176)."""
from __future__ import annotations

import torch
import torch.distributed as dist

_IDX_MASK = 0xFFFFFFFF


def pack_keys(values: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """(float32 >= 0, index < 2**32) -> sortable int64 keys."""
    bits = values.detach().to(torch.float32).contiguous().view(torch.int32).to(torch.int64)
    return (bits << 32) | (_IDX_MASK - indices.to(torch.int64))


def unpack_keys(keys: torch.Tensor):
    """keys -> (values, indices); the kernel's 0 key ("no positive overlap") reads as (0.0, index 0)."""
    vals = (keys >> 32).to(torch.int32).view(torch.float32)
    idx = torch.where(keys == 0, torch.zeros_like(keys), _IDX_MASK - (keys & _IDX_MASK))
    return vals, idx


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous, balanced split of n rows: the first n % world ranks hold one extra row."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_assignment(anchor_keys_local: torch.Tensor, gt_keys_local: torch.Tensor, n_anchors: int, group=None):
    """Collective step: local packed keys -> global (anchor_max, anchor_arg, gt_max, gt_arg) on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        a_keys, g_keys = anchor_keys_local, gt_keys_local
    else:
        rank = dist.get_rank(group)
        sizes = [shard_bounds(n_anchors, world, r) for r in range(world)]
        cap = max(hi - lo for lo, hi in sizes)
        pad = anchor_keys_local.new_zeros(cap)
        pad[:anchor_keys_local.numel()] = anchor_keys_local
        gathered = anchor_keys_local.new_empty(world * cap)
        dist.all_gather_into_tensor(gathered, pad, group=group)
        a_keys = torch.cat([gathered[r * cap:r * cap + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])
        g_keys = gt_keys_local.clone()
        dist.all_reduce(g_keys, op=dist.ReduceOp.MAX, group=group)
        assert sizes[rank][1] - sizes[rank][0] == anchor_keys_local.numel()
    a_max, a_arg = unpack_keys(a_keys)
    g_max, g_arg = unpack_keys(g_keys)
    return a_max, a_arg, g_max, g_arg


def sharded_max_overlaps(anchors_local, gts, n_anchors, anchor_offset, backend='sph2pob_efficient_iou', mode='iou',
                         anchors_are='bboxes1', group=None):
    """overlaps.max over both axes of the logical [n_anchors x n_gt] (or transposed) matrix.

    anchors_local : this rank's contiguous shard [n_local, D] starting at global row `anchor_offset`
    gts           : the replicated short set [G, D]
    anchors_are   : 'bboxes1' -> IoU(anchor, gt) (config #5 call), 'bboxes2' -> IoU(gt, anchor)
                    (the assigner's orientation); the jitters are role-asymmetric, so this matters.
    Returns (anchor_max[n_anchors], anchor_arg -> gt index, gt_max[G], gt_arg -> global anchor index)."""
    from . import _native
    kind = {'sph2pob_standard_iou': 'sph2pob_standard', 'sph2pob_efficient_iou': 'sph2pob_efficient'}[backend]
    # the kernel's packed keys go straight into the collectives: no unpack / repack round trip
    with torch.no_grad():
        if anchors_are == 'bboxes1':
            a_keys, g_keys = _native.iou_pairwise_keys(kind, anchors_local, gts, mode, row_base=anchor_offset)
        else:
            g_keys, a_keys = _native.iou_pairwise_keys(kind, gts, anchors_local, mode, col_base=anchor_offset)
    return gather_assignment(a_keys, g_keys, n_anchors, group)
