"""zinc_b200/dist.py -- multi-GPU commit: one process per GPU (torch.distributed), rows sharded, roots gathered.

Every row of the evaluation matrix is encoded and Merkle-hashed independently (commit.rs:71-81: one tree per
row, the commitment is the list of row roots), so a commit shards by contiguous ROW RANGE with no data-path
exchange; the only collective is an all-gather of the 32-byte roots.  `batch_commit`
(commit.rs:134-142) shards whole polynomials instead.  The prover data (rows, layers) stays on the GPU that
produced it.
"""
from __future__ import annotations

from typing import Callable

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous balanced partition of range(n): -> (begin, count)"""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def _all_gather_bytes(local: np.ndarray, counts: list[int], group=None) -> np.ndarray:
    """all-gather variable-length uint8 blocks (NCCL: through device memory; gloo: host tensors)"""
    import torch
    import torch.distributed as dist

    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mx = max(counts)
    buf = torch.zeros(mx, dtype=torch.uint8, device=dev)
    buf[: local.size] = torch.from_numpy(np.array(local, dtype=np.uint8, copy=True).reshape(-1)).to(dev)
    out = [torch.empty(mx, dtype=torch.uint8, device=dev) for _ in counts]
    dist.all_gather(out, buf, group=group)
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, counts)])


def sharded_commit(pp, poly, ctx=None, group=None, commit_rows: Callable | None = None):
    """Row-range sharded MultilinearZip::commit.

    Returns (local_data, begin, count, MultilinearZipCommitment with ALL roots).  `commit_rows(pp, evals_slice,
    num_rows_local)` -> (local_data, roots uint8[num_rows_local, 32]) defaults to the GPU path; the CPU tests
    inject the oracle here to exercise the sharding/gather logic without a GPU.
    """
    import torch.distributed as dist

    from .zip import MultilinearZip, MultilinearZipCommitment, MultilinearZipParams, DenseMultilinearExtension, \
        _validate_input, as_limbs

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    _validate_input("commit", pp.num_vars, [poly])
    lc = pp.linear_code
    row_len = lc.row_len()
    expected = pp.num_rows * row_len
    ev = as_limbs(poly.evaluations, lc.zt.N)
    assert ev.shape[0] == expected, (
        f"Polynomial has an incorrect number of evaluations ({ev.shape[0]}) for the expected matrix size ({expected})")
    begin, count = shard_range(pp.num_rows, rank, world)
    local_evals = ev[begin * row_len:(begin + count) * row_len]
    if commit_rows is None:
        def commit_rows(pp_, evals_, n_):
            sub = MultilinearZipParams(pp_.num_vars, n_, pp_.linear_code)
            data, comm = MultilinearZip.commit_resident(
                sub, DenseMultilinearExtension(np.ascontiguousarray(evals_), pp_.num_vars), ctx)
            return data, np.frombuffer(b"".join(comm.roots), dtype=np.uint8).reshape(n_, 32)
    if count:
        local_data, local_roots = commit_rows(pp, local_evals, count)
    else:
        local_data, local_roots = None, np.empty((0, 32), dtype=np.uint8)
    if world > 1:
        counts = [shard_range(pp.num_rows, r, world)[1] * 32 for r in range(world)]
        roots = _all_gather_bytes(local_roots.reshape(-1), counts, group).reshape(pp.num_rows, 32)
    else:
        roots = local_roots
    return local_data, begin, count, MultilinearZipCommitment([roots[i].tobytes() for i in range(pp.num_rows)])


def sharded_batch_commit(pp, polys, ctx=None, group=None, commit_poly: Callable | None = None):
    """batch_commit with polynomial p handled by rank p % world; every rank gets all commitments.

    Returns (dict poly_index -> local prover data, list of MultilinearZipCommitment for every polynomial)."""
    import torch.distributed as dist

    from .zip import MultilinearZip, MultilinearZipCommitment

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if commit_poly is None:
        def commit_poly(pp_, poly_):
            data, comm = MultilinearZip.commit_resident(pp_, poly_, ctx)
            return data, np.frombuffer(b"".join(comm.roots), dtype=np.uint8).reshape(pp_.num_rows, 32)
    mine = [p for p in range(len(polys)) if p % world == rank]
    local, blocks = {}, []
    for p in mine:
        data, roots = commit_poly(pp, polys[p])
        local[p] = data
        blocks.append(roots.reshape(-1))
    local_bytes = np.concatenate(blocks) if blocks else np.empty(0, dtype=np.uint8)
    per = pp.num_rows * 32
    if world > 1:
        counts = [len([p for p in range(len(polys)) if p % world == r]) * per for r in range(world)]
        allb = _all_gather_bytes(local_bytes, counts, group)
        offs = np.cumsum([0] + counts)
        comms = [None] * len(polys)
        for r in range(world):
            for k, p in enumerate([p for p in range(len(polys)) if p % world == r]):
                blk = allb[offs[r] + k * per: offs[r] + (k + 1) * per].reshape(pp.num_rows, 32)
                comms[p] = MultilinearZipCommitment([blk[i].tobytes() for i in range(pp.num_rows)])
    else:
        comms = [MultilinearZipCommitment([b.reshape(-1, 32)[i].tobytes() for i in range(pp.num_rows)])
                 for b in blocks]
    return local, comms
