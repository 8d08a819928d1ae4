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
    from zinc_b200.dist import sharded_batch_commit, sharded_commit
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
