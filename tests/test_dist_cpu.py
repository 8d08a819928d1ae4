"""CPU, world_size 2 over gloo: the N>1 path (row-range sharding, all-gather of roots, batch sharding by
polynomial).  The per-rank compute is injected from the oracle so that the host-side logic runs without a GPU."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nv, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from oracle import cbind
    from zinc_b200 import DenseMultilinearExtension, MultilinearZipParams, RaaCode, ZipTypes
    from zinc_b200.dist import sharded_batch_commit, sharded_commit, sharded_open_columns
    from helpers import shape_for

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    row_len, num_rows, cw = shape_for(nv)
    p1, p2 = cbind.perm_from_seed(cw, 1), cbind.perm_from_seed(cw, 2)
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2)
    pp = MultilinearZipParams.new(nv, num_rows, code)
    rng = np.random.default_rng(1234)  # same data on every rank
    evals = rng.integers(0, 1 << 64, size=1 << nv, dtype=np.uint64)
    poly = DenseMultilinearExtension.from_evaluations_vec(nv, evals)

    def commit_rows(pp_, ev_, n_):
        rc, rows, layers, roots = cbind.commit(np.ascontiguousarray(ev_).reshape(-1), n_, row_len, 2, p1, p2)
        assert rc == 0
        return (rows, layers), roots.reshape(n_, 32)

    data, begin, count, comm = sharded_commit(pp, poly, commit_rows=commit_rows)
    rc, rows, layers, roots = cbind.commit(evals, num_rows, row_len, 2, p1, p2)
    ok = b"".join(comm.roots) == roots.tobytes()
    ok &= np.array_equal(data[0], rows[begin * cw * 4:(begin + count) * cw * 4])

    # column openings of the sharded commitment: each rank serves its rows, the ranges are all-gathered
    import ctypes as C
    depth = cw.bit_length() - 1
    per = (2 << depth) - 2
    cols = np.array([0, 3, cw - 1, cw // 2], dtype=np.uint32)
    p8 = C.POINTER(C.c_uint8)

    def open_rows(rows_, layers_, n_rows):
        vals = np.empty((cols.size, n_rows, 4), dtype=np.uint64)
        paths = np.empty((cols.size, n_rows, depth, 32), dtype=np.uint8)
        path = np.zeros(depth * 32, dtype=np.uint8)
        for ci, col in enumerate(cols):
            for r in range(n_rows):
                vals[ci, r] = rows_[(r * cw + col) * 4:(r * cw + col) * 4 + 4]
                lay = np.ascontiguousarray(layers_[r * per * 32:(r + 1) * per * 32])
                cbind.lib().zo_merkle_create_proof(depth, lay.ctypes.data_as(p8), int(col), path.ctypes.data_as(p8))
                paths[ci, r] = path.reshape(depth, 32)
        return vals, paths

    got_v, got_p = sharded_open_columns(data, cols, num_rows, open_local=lambda d, c: open_rows(d[0], d[1], count))
    exp_v, exp_p = open_rows(rows, layers, num_rows)
    ok &= np.array_equal(got_v, exp_v) and np.array_equal(got_p, exp_p)

    polys = [DenseMultilinearExtension.from_evaluations_vec(
        nv, np.random.default_rng(50 + k).integers(0, 1 << 64, size=1 << nv, dtype=np.uint64)) for k in range(5)]

    def commit_poly(pp_, poly_):
        rc, rows, layers, roots = cbind.commit(poly_.evaluations.reshape(-1), num_rows, row_len, 2, p1, p2)
        return (rows, layers), roots.reshape(num_rows, 32)

    local, comms = sharded_batch_commit(pp, polys, commit_poly=commit_poly)
    ok &= sorted(local) == [p for p in range(5) if p % world == rank]
    for k, poly_k in enumerate(polys):
        rc, _, _, roots_k = cbind.commit(poly_k.evaluations.reshape(-1), num_rows, row_len, 2, p1, p2)
        ok &= b"".join(comms[k].roots) == roots_k.tobytes()
    open(os.path.join(out_dir, f"rank{rank}.ok"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.parametrize("nv", [6, 9])
def test_sharded_commit_world2_gloo(nv, tmp_path, oracle):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, nv, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"rank{r}.ok").read() == "1"
