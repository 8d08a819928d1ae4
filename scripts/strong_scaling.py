#!/usr/bin/env python
"""scripts/strong_scaling.py -- BASELINE.json configs[2] as written: ONE 2^nv commit sharded by row range over N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/strong_scaling.py [--nv 24] [--steps 20] [--warmup 5]

Every rank commits its contiguous row range (device-resident evaluations) and the 32-byte roots are all-gathered with
NCCL inside the timed region (the only collective of the path).  Device-timed, max over ranks.  Prints
one JSON line on rank 0.  (bench.py measures weak scaling: one whole commit per GPU.)
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nv", type=int, default=24)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--p2p", action="store_true",
                    help="gather the roots with the peer-memory kernel (zipgpu_peer_roots_allgather) instead of NCCL")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from helpers import KECCAK_SEEDS, shape_for
    from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices
    from zinc_b200 import _native as nat
    from zinc_b200.dist import PeerRoots, shard_range

    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    L = nat.lib()
    ctx = Context(local)
    nv = args.nv
    row_len, num_rows, cw = shape_for(nv)
    depth = cw.bit_length() - 1
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, KECCAK_SEEDS[0]),
                                     shuffle_seeded_indices(cw, KECCAK_SEEDS[1]))
    h = code.native(ctx, 1, 4)
    begin, count = shard_range(num_rows, rank, world)
    evals = np.random.Generator(np.random.PCG64(0x21C0 + nv)).integers(0, 1 << 64, size=1 << nv, dtype=np.uint64)
    d_ev = torch.from_numpy(evals[begin * row_len:(begin + count) * row_len].view(np.int64)).to(dev)
    d_rows = torch.empty(count * cw * 4, dtype=torch.int64, device=dev)
    d_lay = torch.empty(count * ((2 << depth) - 2) * 32, dtype=torch.uint8, device=dev)
    d_roots_all = torch.empty(num_rows * 32, dtype=torch.uint8, device=dev)
    mine = d_roots_all[begin * 32:(begin + count) * 32]
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)

    peer = PeerRoots(ctx, num_rows) if (args.p2p and world > 1) else None
    gathered = [0]

    def step():
        nat.check(L.zipgpu_commit_device(h, count, d_ev.data_ptr(), d_rows.data_ptr(), d_lay.data_ptr(),
                                         mine.data_ptr(), sptr))
        if peer is not None:
            gathered[0] = peer.allgather(begin, count, mine.data_ptr(), sptr)
        elif world > 1:
            dist.all_gather_into_tensor(d_roots_all, mine)  # equal shards (power-of-two rows and ranks)

    assert num_rows % world == 0, "strong-scaling script expects the rows to divide evenly"
    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    if peer is not None:  # bring the peer-gathered roots where the check below looks
        d_roots_all.copy_(peer.tensor(gathered[0]))
        torch.cuda.synchronize()
    # every rank holds all roots: compare with a single-GPU commit of the whole polynomial on rank 0
    ok = None
    if rank == 0:
        d_all = torch.from_numpy(evals.view(np.int64)).to(dev)
        r_all = torch.empty(num_rows * 32, dtype=torch.uint8, device=dev)
        nat.check(L.zipgpu_commit_device(h, num_rows, d_all.data_ptr(), None, None, r_all.data_ptr(), sptr))
        torch.cuda.synchronize()
        ok = bool(torch.equal(r_all, d_roots_all))
        print(json.dumps({"metric": "zip_commit_evals_per_sec", "scaling": "strong", "n_gpus": world, "nv": nv,
                          "ms_per_commit": ms, "value": (1 << nv) / (ms * 1e-3), "unit": "evals/s",
                          "rows_per_gpu": count, "collective": ("peer-memory kernel (zipgpu_peer_roots_allgather)" if peer is not None else "ncclAllGather") +
                                        " of the 32-byte roots inside the timed region",
                          "roots_equal_single_gpu_commit": ok}), flush=True)
    if peer is not None:
        dist.barrier()
        peer.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
