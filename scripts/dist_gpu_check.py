#!/usr/bin/env python
"""scripts/dist_gpu_check.py -- N-GPU check of the sharded commit paths (run under torchrun on a GPU box).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/dist_gpu_check.py [--nv 16]

Every rank commits its row range (commit) / its polynomials (batch_commit) on its own B200 through libzipgpu; the
roots are all-gathered with NCCL; rank 0 compares everything with the CPU oracle (the checker).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nv", type=int, default=16)
    ap.add_argument("--polys", type=int, default=5)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from helpers import KECCAK_SEEDS, shape_for
    from oracle import cbind
    from zinc_b200 import Context, DenseMultilinearExtension, MultilinearZipParams, RaaCode, ZipTypes
    from zinc_b200.dist import sharded_batch_commit, sharded_commit, sharded_open_columns

    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    nv = args.nv
    row_len, num_rows, cw = shape_for(nv)
    cbind.build()
    p1, p2 = cbind.perm_from_seed(cw, KECCAK_SEEDS[0]), cbind.perm_from_seed(cw, KECCAK_SEEDS[1])
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2)
    pp = MultilinearZipParams.new(nv, num_rows, code)
    evals = np.random.default_rng(99).integers(0, 1 << 64, size=1 << nv, dtype=np.uint64)  # same on every rank
    poly = DenseMultilinearExtension.from_evaluations_vec(nv, evals)

    data, begin, count, comm = sharded_commit(pp, poly, ctx)
    rc, rows, layers, roots = cbind.commit_mt(evals, num_rows, row_len, 2, 0, 0, p1, p2, threads=8, faithful=False)
    ok = rc == 0 and b"".join(comm.roots) == roots.tobytes()
    mine = data.rows().reshape(-1) if data is not None else np.empty(0, dtype=np.uint64)
    ok &= np.array_equal(mine, rows[begin * cw * 4:(begin + count) * cw * 4])
    per = ((2 * cw) - 2) * 32
    mine_l = data.layers().reshape(-1) if data is not None else np.empty(0, dtype=np.uint8)
    ok &= np.array_equal(mine_l, layers[begin * per:(begin + count) * per])

    # row -> column redistribution for `open`: 16 columns, every rank serves its resident rows, NCCL all-gather
    cols = np.random.default_rng(7).integers(0, cw, size=16).astype(np.uint32)
    got_v, got_p = sharded_open_columns(data, cols, num_rows)
    depth = cw.bit_length() - 1
    for ci, col in enumerate(cols):
        exp = rows.reshape(num_rows, cw, 4)[:, col, :]
        ok &= np.array_equal(got_v[ci], exp)
        off = 0
        for lvl in range(depth):  # path level lvl = sibling of (col >> lvl) in that level of every row's layers
            sib = (int(col) >> lvl) ^ 1
            exp_p = layers.reshape(num_rows, per // 32, 32)[:, off + sib, :]
            ok &= np.array_equal(got_p[ci, :, lvl, :], exp_p)
            off += cw >> lvl

    polys = [DenseMultilinearExtension.from_evaluations_vec(
        nv, np.random.default_rng(500 + k).integers(0, 1 << 64, size=1 << nv, dtype=np.uint64))
        for k in range(args.polys)]
    local_data, comms = sharded_batch_commit(pp, polys, ctx)
    ok &= sorted(local_data) == [p for p in range(args.polys) if p % world == rank]
    for k, pk in enumerate(polys):
        rc, _, _, roots_k = cbind.commit_mt(pk.evaluations.reshape(-1), num_rows, row_len, 2, 0, 0, p1, p2,
                                            threads=8, faithful=False, want_rows=False, want_layers=False)
        ok &= rc == 0 and b"".join(comms[k].roots) == roots_k.tobytes()

    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"dist_gpu_check world={world} nv={nv}: {'OK' if int(flag.item()) else 'MISMATCH'} "
              f"(rows {begin}..{begin + count} on rank 0, launches={ctx.launch_count})", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
