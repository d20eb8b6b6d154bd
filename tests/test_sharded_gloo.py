"""The N>1 host logic of the row-sharded overlaps (sph_retina_b200/sharded.py) on CPU: world_size 2
and 3 over gloo.  The per-shard IoU here comes from the oracle and the unpacking of the gathered blocks from
`unpack_gathered_reference` below (test infrastructure: a torch restatement of the one-launch kernel
sphk_unpack_gathered_keys, which the GPU suite checks against this very function); on the GPU box
tests/test_gpu_parity.py runs the same sharding with the CUDA kernels producing and unpacking the keys."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def unpack_gathered_reference(gathered, world, n_long, n_short, cap):
    """What sphk_unpack_gathered_keys computes (include/sphk.h), in plain torch on whatever device `gathered` is on."""
    from sph_retina_b200.sharded import shard_bounds
    gathered = gathered.view(world, cap + n_short)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_long, world, r)
        parts.append(gathered[r, :hi - lo])
    a_keys = torch.cat(parts) if parts else gathered.new_zeros(0)
    g_keys = gathered[:, cap:].max(dim=0)[0]

    def unpack(keys):
        vals = (keys >> 32).to(torch.int32).view(torch.float32)
        idx = torch.where(keys == 0, torch.zeros_like(keys), 0xFFFFFFFF - (keys & 0xFFFFFFFF))
        return vals, idx
    return unpack(a_keys) + unpack(g_keys)


def _plant_ties(anchors, gts):
    """exact duplicates in different shards: the gathered argmax must be the LOWEST global index"""
    if anchors.size(0) > 8:
        anchors[5] = gts[2]
        anchors[anchors.size(0) - 3] = gts[2]


def _worker(rank, world, port, n_anchors, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import sph_oracle as O
    from test_sharded_gloo import _plant_ties
    from test_sharded_gloo import unpack_gathered_reference
    from sph_retina_b200.sharded import block_capacity, exchange_blocks, key_block, pack_keys, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)        # the eager oracle must round identically here and in the parent (same chunking)
    try:
        anchors = O.generate_boxes(n_anchors, alpha_range=(5, 100), beta_range=(5, 100), box="rbfov", seed=3)
        gts = O.generate_boxes(24, alpha_range=(5, 100), beta_range=(5, 100), box="rbfov", seed=4)
        _plant_ties(anchors, gts)
        lo, hi = shard_bounds(n_anchors, world, rank)
        cap, G = block_capacity(n_anchors, world), gts.size(0)
        block = key_block(n_anchors, G, world, "cpu")
        assert block.numel() == cap + G and not block.any()
        if hi > lo:
            iou = O.sph2pob_iou(anchors[lo:hi], gts, "efficient").float()        # [n_local, G]
            a_max, a_arg = iou.max(dim=1)
            g_max, g_arg = iou.max(dim=0)
            # what the kernel leaves in the block: key 0 where nothing overlaps, GLOBAL anchor indices for the GT keys
            block[:hi - lo] = torch.where(a_max > 0, pack_keys(a_max, a_arg), torch.zeros_like(a_arg))
            block[cap:] = torch.where(g_max > 0, pack_keys(g_max, g_arg + lo), torch.zeros_like(g_arg))
        gathered = exchange_blocks(block)
        assert gathered.shape == (world, cap + G)
        res = unpack_gathered_reference(gathered, world, n_anchors, G, cap)
        torch.save([r.clone() for r in res], os.path.join(out_dir, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_anchors", [(2, 301), (3, 2)])
def test_sharded_assignment_equals_single_process(tmp_path, world, n_anchors):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sph_oracle as O
    mp.spawn(_worker, args=(world, _free_port(), n_anchors, str(tmp_path)), nprocs=world, join=True)
    anchors = O.generate_boxes(n_anchors, alpha_range=(5, 100), beta_range=(5, 100), box="rbfov", seed=3)
    gts = O.generate_boxes(24, alpha_range=(5, 100), beta_range=(5, 100), box="rbfov", seed=4)
    _plant_ties(anchors, gts)
    # the oracle evaluated shard by shard, single-threaded, exactly as the workers did (vectorised libm paths may
    # round a tail element differently from a full vector: what is under test is the exchange, not the oracle)
    from sph_retina_b200.sharded import shard_bounds
    nt = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        parts = [O.sph2pob_iou(anchors[lo:hi], gts, "efficient").float() for lo, hi in
                 (shard_bounds(n_anchors, world, r) for r in range(world)) if hi > lo]
    finally:
        torch.set_num_threads(nt)
    iou = torch.cat(parts)
    a_max, a_arg = iou.max(dim=1)
    g_max = iou.max(dim=0)[0]
    # lowest-index tie rule, computed independently of torch.max's convention
    g_arg = torch.tensor([int(torch.nonzero(iou[:, j] == g_max[j])[0]) if g_max[j] > 0 else 0 for j in range(gts.size(0))])
    a_arg_low = torch.tensor([int(torch.nonzero(iou[i] == a_max[i])[0]) if a_max[i] > 0 else 0 for i in range(n_anchors)])
    for rank in range(world):
        ra_max, ra_arg, rg_max, rg_arg = torch.load(os.path.join(str(tmp_path), "rank%d.pt" % rank))
        assert torch.equal(ra_max, a_max) and torch.equal(ra_arg, a_arg_low)
        assert torch.equal(rg_max, g_max) and torch.equal(rg_arg, g_arg)
