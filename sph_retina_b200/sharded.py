"""Row-sharded N x M overlaps for large anchor x GT sweeps (BASELINE configs[4]).

One process per GPU.  The long (anchor) axis is split contiguously over the ranks, the short (GT) set is
replicated.  Every rank runs the fused max/argmax kernel on its shard with GLOBAL index offsets; the kernel
writes its packed keys straight into the rank's communication block

    block = [ keys of the rank's anchors (cap = ceil(n / world) slots) | the rank's keys of the G ground truths ]

and the exchange is ONE kernel launch that writes the per-anchor (max, argmax) of the whole anchor set in global order
and reduces the per-GT keys over the ranks.  Three routes feed it:
  * 'peer' (default where available): the blocks live in symmetric memory; the COMPUTE kernel stores the finished keys of
    its anchors into every peer's buffer while it is still running (``sphk_iou_pairwise_keys_push``: the transfer hides
    under the math), and the unpack launch does an in-kernel flag handshake, fetches the 8 B x G per peer and reads the
    anchors' keys locally (``sphk_unpack_peer_keys``) -- no collective at all;
  * 'peer-pull': the compute kernel writes locally only and the unpack launch reads every shard's keys from its owner
    over NVLink after the handshake (the first version of the route; kept for A/B timing);
  * 'nccl': ONE ``all_gather_into_tensor`` of the blocks (8 B per anchor + 8 B x G per rank), then
    ``sphk_unpack_gathered_keys``.
No padding copy, no concatenation, no eager unpacking on either route.
Packing: int64 key = float32 bits << 32 | (0xFFFFFFFF - index).  IoU >= 0, so integer order is (value, then LOWEST
index); the maximum over ranks therefore equals the single-device tie rule.

The reference has no multi-GPU awareness on this path; the single-device contract being reproduced is
MaxIoUAssigner's overlaps.max(dim=0/1) (mmdet/core/bbox/assigners/max_iou_assigner.py:173-176)."""
from __future__ import annotations

import torch
import torch.distributed as dist

_IDX_MASK = 0xFFFFFFFF
_KINDS = {'sph2pob_standard_iou': 'sph2pob_standard', 'sph2pob_efficient_iou': 'sph2pob_efficient'}


def pack_keys(values: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """(float32 >= 0, index < 2**32) -> the kernel's sortable int64 keys."""
    bits = values.detach().to(torch.float32).contiguous().view(torch.int32).to(torch.int64)
    return (bits << 32) | (_IDX_MASK - indices.to(torch.int64))


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous, balanced split of n rows: the first n % world ranks hold one extra row."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def block_capacity(n_anchors: int, world: int) -> int:
    """Anchor-key slots of one rank's block: the largest shard."""
    return -(-n_anchors // world)


def _world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


_blocks = {}


def key_block(n_anchors: int, n_gt: int, world: int, device, fresh: bool = False) -> torch.Tensor:
    """The rank's communication block [cap + n_gt] int64.  The slots past the rank's own shard are padding that no
    kernel writes: they are zeroed once here.  One block per (device, shape) is kept and re-used: the kernel that fills
    it, the collective that reads it and the unpack that follows are ordered on the caller's stream."""
    key = (str(device), n_anchors, n_gt, world)
    blk = None if fresh else _blocks.get(key)
    if blk is None:
        blk = torch.zeros(block_capacity(n_anchors, world) + n_gt, dtype=torch.int64, device=device)
        if not fresh:
            _blocks[key] = blk
    return blk


def exchange_blocks(block: torch.Tensor, group=None) -> torch.Tensor:
    """all-gather of the ranks' blocks -> [world, cap + n_gt]; the only collective of the sharded path."""
    world = _world(group)
    if world == 1:
        return block.view(1, -1)
    gathered = torch.empty((world, block.numel()), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(gathered.view(-1), block, group=group)
    return gathered


def gather_assignment(block: torch.Tensor, n_anchors: int, n_gt: int, group=None):
    """Collective step (NCCL route): the rank's filled block -> global (anchor_max, anchor_arg, gt_max, gt_arg) on every
    rank (argmax int64 as torch.max returns it).  One all_gather + one kernel launch."""
    from . import _native
    world = _world(group)
    gathered = exchange_blocks(block, group)
    return _native.unpack_gathered_keys(gathered, world, n_anchors, n_gt, block_capacity(n_anchors, world))


MAX_PUSH_PARTS = 8      # push route: at most this many partial key arrays per anchor (n_gt <= 2048), else the pull route


class PeerExchange:
    """The exchange without a collective: every rank keeps the key blocks of ALL ranks in a symmetric buffer
    (torch.distributed._symmetric_memory: the same allocation mapped into every process of the node).  Layout per rank, in
    int64 elements:  [ even steps: world slots | odd steps: world slots | flags ];  slot s = block of rank s =
    [ parts x cap anchor keys | n_gt ].
    parts > 1 or push=True ('peer'): the compute kernel of step t stores the keys of every tile into slot `rank` of parity
    (t & 1) in EVERY rank's buffer while it runs (``sphk_iou_pairwise_keys_push``; parts = column tiles of the kernel);
    the unpack launch (``sphk_unpack_peer_keys``) then needs the flag handshake, 8 B x n_gt per peer over NVLink, and
    local reads.  push=False ('peer-pull', parts = 1): the kernel writes the rank's own slot only and the unpack launch
    reads every slot from its owner over NVLink."""
    FLAG_SLOTS = 32

    def __init__(self, n_anchors: int, n_gt: int, device, group=None, parts: int = 1, push: bool = False):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 16:
            raise RuntimeError("PeerExchange: at most 16 ranks (one node)")
        self.n_anchors, self.n_gt, self.parts, self.push = n_anchors, n_gt, parts, push
        self.cap = block_capacity(n_anchors, self.world)
        self.per = parts * self.cap + n_gt
        self.area = self.world * self.per                      # one parity: the blocks of all ranks
        self.buf = symm_mem.empty(2 * self.area + self.FLAG_SLOTS, dtype=torch.int64, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.buf.zero_()                       # padding slots and flags start at zero ...
        torch.cuda.synchronize(device)
        self.hdl.barrier()                     # ... on every rank before anybody raises a flag or pushes a key
        self.ptrs_dev = int(self.hdl.buffer_ptrs_dev)
        self.step = 0

    def next_block(self) -> torch.Tensor:
        """The rank's own slot of the next step ([parts * cap + n_gt] int64 view of the symmetric buffer)."""
        self.step += 1
        off = (self.step & 1) * self.area + self.rank * self.per
        return self.buf[off:off + self.per]

    def compute(self, kind, anchors_local, gts, mode, anchor_offset: int):
        """Push route: the step's fused max / argmax kernel (anchors = bboxes1); its anchors' keys go to every rank."""
        from . import _native
        block = self.next_block()
        off = (self.step & 1) * self.area + self.rank * self.per
        _native.iou_pairwise_keys_push(kind, anchors_local, gts, block[self.parts * self.cap:], self.ptrs_dev, self.world, off,
                                       self.cap, mode=mode, row_base=anchor_offset)

    def finish(self, out=None):
        """-> (anchor_max, anchor_arg, gt_max, gt_arg) of the step whose block was handed out last."""
        from . import _native
        return _native.unpack_peer_keys(self.ptrs_dev, self.rank, self.world, self.step, (self.step & 1) * self.area, 2 * self.area,
                                        self.n_anchors, self.n_gt, self.cap, self.buf.device, out=out, long_parts=self.parts,
                                        long_pushed=self.push)


_peer_exchanges = {}
_peer_failed = []


def _all_ranks_ok(ok: bool, device, group) -> bool:
    """Collective AND of a per-rank success flag: the ranks must all take the same route."""
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()))


def peer_exchange(n_anchors: int, n_gt: int, device, group=None, parts: int = 1, push: bool = False):
    """The cached PeerExchange of (shape, device, route), or None where symmetric memory is not available on EVERY rank
    (then: NCCL route).  Collective on first use: each step of the set-up is agreed on by all ranks before the next one."""
    key = (str(device), n_anchors, n_gt, _world(group), parts, push)
    ex = _peer_exchanges.get(key)
    if ex is None and not _peer_failed:
        why = None
        try:
            import torch.distributed._symmetric_memory as symm_mem      # noqa: F401  (API present in this torch?)
            probe = symm_mem.empty(16, dtype=torch.int64, device=device)
            del probe
        except Exception as e:                # pragma: no cover
            why = repr(e)
        if not _all_ranks_ok(why is None, device, group):
            why = why or "symmetric memory unavailable on another rank"
        else:
            try:
                ex = PeerExchange(n_anchors, n_gt, device, group, parts, push)
            except Exception as e:            # pragma: no cover
                why = repr(e)
            if not _all_ranks_ok(ex is not None, device, group):
                ex, why = None, why or "symmetric-memory rendezvous failed on another rank"
        if ex is None:
            import warnings
            _peer_failed.append(why)
            warnings.warn("sph_retina_b200.sharded: symmetric-memory exchange unavailable (%s); using the NCCL all_gather route" % why)
        else:
            _peer_exchanges[key] = ex
    return ex


def exchange_route(device=None, group=None) -> str:
    """'peer' | 'nccl' | 'single' -- which route sharded_max_overlaps(exchange='auto') takes in this process."""
    if _world(group) == 1:
        return "single"
    return "nccl" if _peer_failed else "peer"


def sharded_max_overlaps(anchors_local, gts, n_anchors, anchor_offset, backend='sph2pob_efficient_iou', mode='iou',
                         anchors_are='bboxes1', group=None, exchange='auto'):
    """overlaps.max over both axes of the logical [n_anchors x n_gt] (or transposed) matrix.

    anchors_local : this rank's contiguous shard [n_local, D] starting at global row `anchor_offset`
                    (= shard_bounds(n_anchors, world, rank))
    gts           : the replicated short set [G, D]
    anchors_are   : 'bboxes1' -> IoU(anchor, gt) (config #5 call), 'bboxes2' -> IoU(gt, anchor)
                    (the assigner's orientation); the jitters are role-asymmetric, so this matters.
    exchange      : 'peer' -- no collective: the compute kernel stores the anchors' keys into every peer's symmetric buffer
                    while it runs, the unpack launch does the flag handshake and reads locally; 'peer-pull' -- the
                    compute kernel writes locally only, the unpack launch reads the keys from their owners over NVLink;
                    'nccl' -- one all_gather_into_tensor + the unpack launch; 'auto' -- 'peer' where available.
                    Every rank must make the same choice.
    Returns (anchor_max[n_anchors], anchor_arg -> gt index, gt_max[G], gt_arg -> global anchor index)."""
    from . import _native
    kind = _KINDS[backend]
    world = _world(group)
    n_local, n_gt = anchors_local.size(0), gts.size(0)
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_bounds(n_anchors, world, rank)
    if (lo, hi - lo) != (anchor_offset, n_local):
        raise ValueError("rank %d of %d must hold anchors [%d, %d) of %d, got offset %d and %d rows"
                         % (rank, world, lo, hi, n_anchors, anchor_offset, n_local))
    cap = block_capacity(n_anchors, world)
    if exchange not in ('auto', 'peer', 'peer-pull', 'nccl'):
        raise ValueError("exchange must be 'auto', 'peer', 'peer-pull' or 'nccl'")
    ex = None
    if world > 1 and exchange != 'nccl':
        from . import _native as _nat
        # push: the anchors must be the kernel's rows (one partial key array per 256 ground truths); else pull
        parts = _nat.key_push_parts(n_gt)
        push = exchange != 'peer-pull' and anchors_are == 'bboxes1' and parts <= MAX_PUSH_PARTS
        ex = peer_exchange(n_anchors, n_gt, anchors_local.device, group, parts if push else 1, push)
    if exchange in ('peer', 'peer-pull') and world > 1 and ex is None:
        raise RuntimeError("sharded_max_overlaps(exchange=%r): symmetric memory is not available: %s" % (exchange, _peer_failed))
    if ex is not None and ex.push:
        with torch.no_grad():
            ex.compute(kind, anchors_local, gts, mode, anchor_offset)
        return ex.finish()
    block = ex.next_block() if ex is not None else key_block(n_anchors, n_gt, world, anchors_local.device)
    a_out, g_out = block[:n_local], block[cap:]
    # the kernel's packed keys go straight into the communication block: no unpack / repack / copy
    with torch.no_grad():
        if anchors_are == 'bboxes1':
            _native.iou_pairwise_keys(kind, anchors_local, gts, mode, row_base=anchor_offset, row_keys_out=a_out,
                                      col_keys_out=g_out)
        else:
            _native.iou_pairwise_keys(kind, gts, anchors_local, mode, col_base=anchor_offset, row_keys_out=g_out,
                                      col_keys_out=a_out)
    if ex is not None:
        return ex.finish()
    return gather_assignment(block, n_anchors, n_gt, group)


class HostSweep:
    """The sharded sweep for callers whose boxes and results live in HOST memory (pinned): copies and kernels of a step
    are pipelined over row chunks instead of running one after the other.

        hs = HostSweep(n_anchors, n_local, n_gt, D, device)            # once: staging buffers, a copy stream, events
        hs(anchors_pinned, gts_pinned, anchor_offset, out_anchor_max, out_anchor_arg, out_gt_max, out_gt_arg)

    Per chunk of the rank's anchors: H2D on the copy stream -> the fused max/argmax kernel on the caller's stream (its
    per-GT keys accumulate over the chunks: keep_col_keys) -> unpack of the chunk's per-anchor keys (final: every chunk
    sees all ground truths) -> D2H on the copy stream, under the next chunk's kernel.  Then the per-GT keys are reduced
    over the ranks (the same 'peer' / 'nccl' routes as sharded_max_overlaps, on 8 B x G per rank) and copied back.
    Every rank ends up with the per-anchor result of ITS anchors and the global per-GT result -- the data-parallel
    consumer's view; the all-ranks per-anchor gather of sharded_max_overlaps is not part of it.  All calls are
    asynchronous: the outputs are valid once the caller's stream has been synchronised (``hs.done`` is recorded last)."""

    def __init__(self, n_anchors, n_local, n_gt, D, device, group=None, min_chunk_rows=65536, max_chunks=8, exchange='auto'):
        self.n_anchors, self.n_local, self.n_gt, self.D = n_anchors, n_local, n_gt, D
        self.device, self.group = torch.device(device), group
        self.world = _world(group)
        # chunk plan: what stays exposed is the H2D of the FIRST chunk and the D2H of the LAST one, while every extra
        # launch costs one more kernel tail (~15 us): a short first and last chunk (1/8 of the rows, at most 32,768)
        # around one or two long ones; equal chunks when the caller fixes their number (min_chunk_rows = 1)
        if min_chunk_rows <= 1:
            self.chunks = max(1, min(max_chunks, n_local))
            self.bounds = [shard_bounds(n_local, self.chunks, j) for j in range(self.chunks)]
        elif n_local < 2 * min_chunk_rows:
            self.chunks, self.bounds = 1, [(0, n_local)]
        else:
            edge = min(32768, n_local // 8)
            cuts = [0, edge] + ([n_local // 2] if n_local >= 4 * min_chunk_rows else []) + [n_local - edge, n_local]
            self.bounds = list(zip(cuts[:-1], cuts[1:]))
            self.chunks = len(self.bounds)
        dev = self.device
        self.a_d = torch.empty((n_local, D), dtype=torch.float32, device=dev)
        self.g_d = torch.empty((n_gt, D), dtype=torch.float32, device=dev)
        self.akeys = torch.zeros(n_local, dtype=torch.int64, device=dev)
        self.amax_d = torch.empty(n_local, dtype=torch.float32, device=dev)
        self.aarg_d = torch.empty(n_local, dtype=torch.int64, device=dev)
        self.gmax_d = torch.empty(n_gt, dtype=torch.float32, device=dev)
        self.garg_d = torch.empty(n_gt, dtype=torch.int64, device=dev)
        self.none_f = torch.empty(0, dtype=torch.float32, device=dev)
        self.none_i = torch.empty(0, dtype=torch.int64, device=dev)
        self.copy_stream = torch.cuda.Stream(dev)
        self.ev_in = [torch.cuda.Event() for _ in range(self.chunks)]
        self.ev_out = [torch.cuda.Event() for _ in range(self.chunks)]
        self.done = torch.cuda.Event()
        # the per-GT keys of this rank: one block [0 anchor slots | n_gt], in symmetric memory where that route is available
        self.ex = peer_exchange(0, n_gt, dev, group) if (self.world > 1 and exchange != 'nccl') else None
        self.gblock = None if self.ex is not None else torch.zeros(n_gt, dtype=torch.int64, device=dev)

    def __call__(self, anchors_host, gts_host, anchor_offset, out_anchor_max, out_anchor_arg, out_gt_max, out_gt_arg,
                 backend='sph2pob_efficient_iou', mode='iou'):
        from . import _native
        kind = _KINDS[backend]
        cur, cp = torch.cuda.current_stream(self.device), self.copy_stream
        cp.wait_stream(cur)                                   # the previous step's consumers of the staging buffers are done
        with torch.cuda.stream(cp):
            for j, (lo, hi) in enumerate(self.bounds):
                self.a_d[lo:hi].copy_(anchors_host[lo:hi], non_blocking=True)
                self.ev_in[j].record(cp)
        self.g_d.copy_(gts_host, non_blocking=True)
        gkeys = self.ex.next_block() if self.ex is not None else self.gblock
        with torch.no_grad():
            for j, (lo, hi) in enumerate(self.bounds):
                cur.wait_event(self.ev_in[j])
                _native.iou_pairwise_keys(kind, self.a_d[lo:hi], self.g_d, mode, row_base=anchor_offset + lo,
                                          row_keys_out=self.akeys[lo:hi], col_keys_out=gkeys, keep_col_keys=j > 0)
                _native.unpack_gathered_keys(self.akeys[lo:hi], 1, hi - lo, 0, hi - lo,
                                             out=(self.amax_d[lo:hi], self.aarg_d[lo:hi], self.none_f, self.none_i))
                self.ev_out[j].record(cur)
                with torch.cuda.stream(cp):
                    cp.wait_event(self.ev_out[j])
                    out_anchor_max[lo:hi].copy_(self.amax_d[lo:hi], non_blocking=True)
                    out_anchor_arg[lo:hi].copy_(self.aarg_d[lo:hi], non_blocking=True)
        res = (self.none_f, self.none_i, self.gmax_d, self.garg_d)
        if self.ex is not None:
            self.ex.finish(out=res)
        else:
            _native.unpack_gathered_keys(exchange_blocks(gkeys, self.group).view(-1), self.world, 0, self.n_gt, 0, out=res)
        out_gt_max.copy_(self.gmax_d, non_blocking=True)
        out_gt_arg.copy_(self.garg_d, non_blocking=True)
        cur.wait_stream(cp)                                   # the chunk results are on their way: the step ends when they have landed
        self.done.record(cur)
        return out_anchor_max, out_anchor_arg, out_gt_max, out_gt_arg
