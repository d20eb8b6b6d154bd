#!/usr/bin/env python
"""Where the sharded sweep's step goes on every rank (run under torchrun, one rank per GPU):

    per step and rank:  flush | [align] | e0 | compute kernel | e_mid | exchange + unpack launch | e1

k_r = e_mid - e0 (the rank's own kernel), w_r = e1 - e_mid (handshake wait + peer reads + result writes).  The rank that
arrives last at the handshake waits for nobody: min over ranks of w_r = the cost of the exchange itself, the rest of a
rank's w_r is waiting for slower / later peers.  --align runs the symmetric-memory barrier between the flush and e0
(ranks leave the L2 flush at different times; in a pipeline without flushes the previous step's handshake aligns them).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29711 \
        tools/exchange_probe.py [--align] [--steps 30]"""
import argparse
import json
import os
import statistics
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sph_retina_b200 import _native, synthetic as S  # noqa: E402
from sph_retina_b200.sharded import block_capacity, peer_exchange, shard_bounds  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--align", action="store_true")
    ap.add_argument("--route", default="push", choices=["push", "pull", "local"],
                    help="push: keys stored into the peers' buffers by the compute kernel; pull: read by the unpack launch; "
                         "local: the compute kernel alone into an ordinary (not symmetric) buffer, no exchange")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // world
        if per >= 1:
            os.sched_setaffinity(0, cores[local * per:(local + 1) * per])
    except (AttributeError, OSError, ValueError):
        pass
    dist.init_process_group("nccl", device_id=dev)
    n, g = 1 << 20, 1024
    A = S.generate_boxes(n, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=0)
    G = S.generate_boxes(g, alpha_range=(1, 100), beta_range=(1, 100), box="rbfov", seed=1).to(dev)
    lo, hi = shard_bounds(n, world, rank)
    a = A[lo:hi].contiguous().to(dev)
    cap = block_capacity(n, world)
    parts = _native.key_push_parts(g) if args.route == "push" else 1
    ex = peer_exchange(n, g, dev, None, parts, args.route == "push")
    assert ex is not None, "symmetric memory unavailable"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = None
    evs = []
    plain = torch.zeros(cap + g, dtype=torch.int64, device=dev)
    for it in range(args.steps + 5):
        flush.zero_()
        if args.align:
            ex.hdl.barrier()
        e0, em, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        if args.route == "push":
            ex.compute("sph2pob_efficient", a, G, "iou", lo)
        else:
            blk = ex.next_block() if args.route == "pull" else plain
            _native.iou_pairwise_keys("sph2pob_efficient", a, G, row_base=lo, row_keys_out=blk[:hi - lo], col_keys_out=blk[cap:])
        em.record()
        if args.route != "local":
            out = ex.finish(out=out)
        e1.record()
        if it >= 5:
            evs.append((e0, em, e1))
    torch.cuda.synchronize()
    k = torch.tensor([x.elapsed_time(y) for x, y, _ in evs], dtype=torch.float64, device=dev)
    w = torch.tensor([y.elapsed_time(z) for _, y, z in evs], dtype=torch.float64, device=dev)
    ks = [torch.empty_like(k) for _ in range(world)]
    ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ks, k)
    dist.all_gather(ws, w)
    if rank == 0:
        K, W = torch.stack(ks).cpu(), torch.stack(ws).cpu()          # [world, steps]
        T = K + W
        rec = {"world": world, "route": args.route, "align": args.align, "steps": args.steps,
               "kernel_ms_mean_per_rank": [round(float(x), 4) for x in K.mean(1)],
               "tail_ms_mean_per_rank": [round(float(x), 4) for x in W.mean(1)],
               "step_ms_mean_per_rank": [round(float(x), 4) for x in T.mean(1)],
               "step_ms_max_over_ranks_of_sums": round(float(T.sum(1).max()) / args.steps, 4),
               "exchange_cost_ms (mean over steps of min over ranks of the tail)": round(float(W.min(0).values.mean()), 4),
               "kernel_ms_max_over_ranks_mean": round(float(K.max(0).values.mean()), 4),
               "kernel_ms_min_over_ranks_mean": round(float(K.min(0).values.mean()), 4),
               "tail_ms_max_over_ranks_mean": round(float(W.max(0).values.mean()), 4),
               "kernel_ms_median_all": round(float(statistics.median(K.flatten().tolist())), 4)}
        print(json.dumps(rec))
        if args.tag:
            os.makedirs("gpurun_out", exist_ok=True)
            with open("gpurun_out/exchange_probe_%s.json" % args.tag, "w") as f:
                json.dump(rec, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
