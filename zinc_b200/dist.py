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


def sharded_commit(pp, poly, ctx=None, group=None, commit_rows: Callable | None = None, peer=None):
    """Row-range sharded MultilinearZip::commit.

    `peer` (a PeerRoots built for pp.num_rows rows): gather the roots with the peer-memory kernel straight from the
    resident commit's device buffer instead of a torch.distributed all-gather of host bytes.

    Returns (local_data, begin, count, MultilinearZipCommitment with ALL roots).  `commit_rows(pp, evals_slice,
    num_rows_local)` -> (local_data, roots uint8[num_rows_local, 32]) defaults to the GPU path; the CPU tests
    inject a host function here to exercise the sharding/gather logic without a GPU.
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
    if world > 1 and peer is not None and commit_rows is None:
        # the GPU path proper: H2D of this rank's slice, encode + hash, and the roots exchange inside the kernel that
        # produces the roots (zipgpu_commit_resident_sharded); every rank ends up with all roots
        local_data, roots = peer.commit_resident(lc, pp, np.ascontiguousarray(local_evals), begin, count, ctx)
        return local_data, begin, count, MultilinearZipCommitment([roots[i].tobytes() for i in range(pp.num_rows)])
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


def sharded_open_columns(local_data, columns, num_rows: int, group=None, open_local: Callable | None = None):
    """Column openings of a row-sharded commitment (open_z.rs:124-143): every rank extracts, from ITS resident rows, the
    entries of the requested columns and their Merkle paths (zipgpu_data_open_columns), and the row ranges are
    all-gathered in rank order -- the row -> column redistribution of the opening phase.  No codeword or layer data
    other than the opened columns ever leaves a GPU.

    Returns (values uint64 [ncols, num_rows, K], paths uint8 [ncols, num_rows, depth, 32]) on every rank.
    `open_local(local_data, columns)` defaults to ResidentZipData.open_columns (the CPU tests inject a host function)."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    cols = np.ascontiguousarray(columns, dtype=np.uint32)
    if open_local is None:
        open_local = lambda d, c: d.open_columns(c)
    if local_data is not None:
        vals, paths = open_local(local_data, cols)  # [ncols, local_rows, K], [ncols, local_rows, depth, 32]
    else:
        vals, paths = None, None
    if world == 1:
        return vals, paths
    # shapes are known from rank-independent quantities except K and depth: take them from a rank that has rows
    meta = np.zeros(2, dtype=np.int64)
    if vals is not None:
        meta[:] = (vals.shape[2], paths.shape[2])
    import torch

    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mt = torch.from_numpy(meta).to(dev)
    dist.all_reduce(mt, op=dist.ReduceOp.MAX, group=group)
    k, depth = int(mt[0].item()), int(mt[1].item())
    ncols = cols.size
    counts_rows = [shard_range(num_rows, r, world)[1] for r in range(world)]
    v_local = (vals if vals is not None else np.empty((ncols, 0, k), dtype=np.uint64))
    p_local = (paths if paths is not None else np.empty((ncols, 0, depth, 32), dtype=np.uint8))
    v_all = _all_gather_bytes(np.ascontiguousarray(v_local).view(np.uint8).reshape(-1),
                              [ncols * c * k * 8 for c in counts_rows], group)
    p_all = _all_gather_bytes(np.ascontiguousarray(p_local).reshape(-1),
                              [ncols * c * depth * 32 for c in counts_rows], group)
    out_v = np.empty((ncols, num_rows, k), dtype=np.uint64)
    out_p = np.empty((ncols, num_rows, depth, 32), dtype=np.uint8)
    ov = op = r0 = 0
    for c in counts_rows:
        nv_, np_ = ncols * c * k * 8, ncols * c * depth * 32
        out_v[:, r0:r0 + c] = v_all[ov:ov + nv_].view(np.uint64).reshape(ncols, c, k)
        out_p[:, r0:r0 + c] = p_all[op:op + np_].reshape(ncols, c, depth, 32)
        ov, op, r0 = ov + nv_, op + np_, r0 + c
    return out_v, out_p


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


class PeerRoots:
    """All-gather of the row roots over NVLink peer memory (include/zipgpu.h, csrc/peer_roots.cu): one kernel per GPU
    and step stores the local roots into every peer's buffer, signals and waits -- the only exchange step of a
    row-sharded commit without NCCL on the data path.  One process per GPU on one node; construction exchanges the IPC
    descriptors once through torch.distributed (any backend)."""

    def __init__(self, ctx, total_rows: int, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _native as nat

        self._nat, self._C = nat, C
        self._ctx = ctx
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.total_rows = total_rows
        self.handle = C.c_void_p()
        mine = np.zeros(64, dtype=np.uint8)
        nat.check(nat.lib().zipgpu_peer_roots_create(ctx.handle, total_rows, self.rank, self.world, C.byref(self.handle),
                                                     nat.ptr(mine)))
        if self.world > 1:
            backend = dist.get_backend(group)
            dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
            t = torch.from_numpy(mine).to(dev)
            out = torch.empty(64 * self.world, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(out, t, group=group)
            everyone = np.ascontiguousarray(out.cpu().numpy())
            nat.check(nat.lib().zipgpu_peer_roots_connect(self.handle, nat.ptr(everyone)))
            dist.barrier(group)  # nobody stores into a peer before that peer's buffers exist and are mapped
        else:
            nat.check(nat.lib().zipgpu_peer_roots_connect(self.handle, nat.ptr(mine)))

    def allgather(self, row_begin: int, count: int, d_local_roots: int, stream=None) -> int:
        """d_local_roots: device pointer to count*32 bytes.  Enqueues on `stream` (None = the context's stream) and
        returns the device pointer that holds all total_rows*32 bytes once the stream has passed this point."""
        C = self._C
        out = C.c_void_p()
        self._nat.check(self._nat.lib().zipgpu_peer_roots_allgather(
            self.handle, row_begin, count, C.c_void_p(d_local_roots) if d_local_roots else None, stream, C.byref(out)))
        return out.value

    def commit_device(self, hcode, row_begin: int, count: int, d_evals: int, d_rows: int | None, d_layers: int | None,
                      stream=None) -> int:
        """zipgpu_commit_device_sharded: commit of this rank's row range from device-resident evaluations with the
        roots exchange fused into the kernel that produces the roots.  Returns the device pointer holding ALL roots
        once the stream has passed this point (valid until the call after next)."""
        C = self._C
        out = C.c_void_p()
        self._nat.check(self._nat.lib().zipgpu_commit_device_sharded(
            hcode, self.handle, row_begin, count, C.c_void_p(d_evals) if d_evals else None,
            C.c_void_p(d_rows) if d_rows else None, C.c_void_p(d_layers) if d_layers else None, stream, C.byref(out)))
        return out.value

    def commit_resident(self, code, pp, local_evals: np.ndarray, row_begin: int, count: int, ctx):
        """zipgpu_commit_resident_sharded: host evaluations of this rank's rows in, ALL roots of the commitment out
        (uint8 [num_rows, 32]); the prover data of the local rows stays on this GPU."""
        from .zip import ResidentZipData

        C = self._C
        zt = code.zt
        roots = np.empty((self.total_rows, 32), dtype=np.uint8)
        h = C.c_void_p()
        self._nat.check(self._nat.lib().zipgpu_commit_resident_sharded(
            code.native(ctx, zt.N, zt.K), self.handle, row_begin, count, self._nat.ptr(local_evals) if count else None,
            self._nat.ptr(roots), C.byref(h)))
        cw = code.codeword_len()
        data = ResidentZipData(h, count, cw, zt.K, cw.bit_length() - 1, code.row_len(), ctx) if h else None
        return data, roots

    def status(self) -> None:
        """raises ZipGpuError(ERR_PEER_TIMEOUT) if a kernel of this exchange gave up waiting for a peer"""
        self._nat.check(self._nat.lib().zipgpu_peer_roots_status(self.handle))

    def sync(self) -> None:
        """wait for the context's stream (where the exchange enqueues by default) and check the exchange's status"""
        self._ctx.sync()
        self.status()

    def tensor(self, ptr: int):
        """a torch uint8 view of the gathered roots behind the pointer `allgather` returned"""
        import torch

        class _Raw:
            pass

        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (self.total_rows * 32,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        return torch.as_tensor(raw, device=torch.device("cuda", torch.cuda.current_device()))

    def close(self) -> None:
        if self.handle:
            self._nat.lib().zipgpu_peer_roots_destroy(self.handle)
            self.handle = None
