"""Multi-GPU commit paths against the single-GPU result and the oracle.

  * the roots exchange fused into the roots-producing kernel (zipgpu_commit_device_sharded / _resident_sharded),
    with world = 1 so that the fused code path runs on a one-GPU box as well;
  * zipgpu_mgpu_* (ONE process, n devices) through the host mirror's MultiContext: commit, commit_no_merkle,
    batch_commit, commit_resident + open_columns / wire / combine_rows in row order.  On a one-GPU box the same code
    runs with one device and with two ranks placed on the same device (test knob); with >= 2 GPUs over all of them.
"""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import KECCAK_SEEDS, MOCK_SEEDS, shape_for

pytestmark = pytest.mark.gpu


def _code(nv, seeds, oracle):
    from zinc_b200 import RaaCode, ZipTypes

    row_len, num_rows, cw = shape_for(nv)
    p1, p2 = oracle.perm_from_seed(cw, seeds[0]), oracle.perm_from_seed(cw, seeds[1])
    return RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2), row_len, num_rows, cw, p1, p2


def _device_count():
    from zinc_b200 import _native as nat

    n = C.c_int()
    nat.lib().zipgpu_device_count(C.byref(n))
    return n.value


@pytest.mark.parametrize("nv", [6, 8, 12, 16, 20, 21, 22])
def test_fused_exchange_world1_device(nv, oracle, ctx):
    """zipgpu_commit_device_sharded with a one-rank exchange: the launch that produces the roots also stores them into
    the result buffer and runs the publish/wait handshake with itself; roots == zipgpu_commit_device == oracle"""
    import torch

    from zinc_b200 import _native as nat
    from zinc_b200.dist import PeerRoots

    L = nat.lib()
    code, row_len, num_rows, cw, p1, p2 = _code(nv, KECCAK_SEEDS, oracle)
    h = code.native(ctx, 1, 4)
    evals = np.random.default_rng(nv).integers(0, 1 << 64, size=1 << nv, dtype=np.uint64)
    dev = torch.device("cuda", ctx.device)
    d_ev = torch.from_numpy(evals.view(np.int64)).to(dev)
    d_roots = torch.empty(num_rows * 32, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    nat.check(L.zipgpu_commit_device(h, num_rows, d_ev.data_ptr(), None, None, d_roots.data_ptr(), None))
    ctx.sync()
    peer = PeerRoots(ctx, num_rows)
    launches0 = ctx.launch_count
    for step in range(3):  # both parities of the double buffer
        ptr = peer.commit_device(h, 0, num_rows, d_ev.data_ptr(), None, None)
        peer.sync()
        got = peer.tensor(ptr).cpu().numpy()
        assert np.array_equal(got, d_roots.cpu().numpy()), f"step {step}"
    # fused: no stand-alone exchange kernel was launched -- same launch count per commit as the plain call
    per_commit = (ctx.launch_count - launches0) // 3
    l1 = ctx.launch_count
    nat.check(L.zipgpu_commit_device(h, num_rows, d_ev.data_ptr(), None, None, d_roots.data_ptr(), None))
    ctx.sync()
    assert per_commit == ctx.launch_count - l1
    rc, _, _, roots = oracle.commit_mt(evals, num_rows, row_len, 2, 0, 0, p1, p2, threads=8, faithful=False,
                                       want_rows=False, want_layers=False)
    assert rc == 0 and got.tobytes() == roots.tobytes()
    peer.close()


@pytest.mark.parametrize("units", ["1", "2"])
def test_fused_exchange_in_the_commit_kernel(units, oracle, ctx, monkeypatch):
    """a shard whose trees are finished by the warp-specialised commit kernel itself (tops epilogue): that ONE launch also
    carries the roots exchange, with whole-row units and with half-row units (roots of split rows made by the CTA that
    arrives second)"""
    import torch

    from zinc_b200.dist import PeerRoots

    code, row_len, num_rows, cw, p1, p2 = _code(22, KECCAK_SEEDS, oracle)
    num_rows = 333
    h = code.native(ctx, 1, 4)
    evals = np.random.default_rng(5).integers(0, 1 << 64, size=num_rows * row_len, dtype=np.uint64)
    dev = torch.device("cuda", ctx.device)
    d_ev = torch.from_numpy(evals.view(np.int64)).to(dev)
    depth = cw.bit_length() - 1
    d_rows = torch.empty(num_rows * cw * 4, dtype=torch.int64, device=dev)
    d_lay = torch.empty(num_rows * ((2 << depth) - 2) * 32, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    monkeypatch.setenv("ZIPGPU_WS_UNITS", units)
    monkeypatch.setenv("ZIPGPU_WS_TOPS", "1")
    total = num_rows + 100  # the local rows are [100, 433) of a larger commitment
    peer = PeerRoots(ctx, total)
    rc, _, _, roots = oracle.commit_mt(evals, num_rows, row_len, 2, 0, 0, p1, p2, threads=8, faithful=False,
                                       want_rows=False, want_layers=False)
    assert rc == 0
    for step in range(3):
        l0 = ctx.launch_count
        ptr = peer.commit_device(h, 100, num_rows, d_ev.data_ptr(), d_rows.data_ptr(), d_lay.data_ptr())
        peer.sync()
        assert ctx.launch_count - l0 == 1
        got = peer.tensor(ptr).cpu().numpy()
        assert got[100 * 32:].tobytes() == roots.tobytes(), step
    peer.close()


def test_exchange_standalone_and_empty_rank(oracle, ctx):
    """the stand-alone exchange kernel (roots already in device memory) and a rank without rows (count = 0)"""
    import torch

    from zinc_b200.dist import PeerRoots

    dev = torch.device("cuda", ctx.device)
    total = 96
    local = torch.arange(total * 32, dtype=torch.int64, device=dev).to(torch.uint8)
    peer = PeerRoots(ctx, total)
    ptr = peer.allgather(0, total, local.data_ptr())
    peer.sync()
    assert torch.equal(peer.tensor(ptr), local)
    ptr = peer.allgather(0, 0, 0)  # nothing to contribute: handshake only
    peer.sync()
    peer.close()


def _check_multi(mctx, ctx, oracle, nv=14):
    from zinc_b200 import DenseMultilinearExtension, MultilinearZip, MultilinearZipParams
    import oracle.pyoracle as po

    code, row_len, num_rows, cw, p1, p2 = _code(nv, MOCK_SEEDS, oracle)
    depth = cw.bit_length() - 1
    pp = MultilinearZipParams.new(nv, num_rows, code)
    poly = DenseMultilinearExtension.rand(nv, np.random.default_rng(7 + nv))
    rc, rows, layers, roots = oracle.commit_mt(poly.evaluations.reshape(-1), num_rows, row_len, 2, 0, 0, p1, p2,
                                               threads=8, faithful=False)
    assert rc == 0
    # commit with every output on the host
    data, comm = MultilinearZip.commit(pp, poly, mctx)
    assert np.array_equal(data.rows.reshape(-1), rows)
    assert np.array_equal(np.concatenate([t.layers.reshape(-1) for t in data.rows_merkle_trees]), layers)
    assert b"".join(comm.roots) == roots.tobytes()
    # commit_no_merkle
    d2, c2 = MultilinearZip.commit_no_merkle(pp, poly, mctx)
    assert np.array_equal(d2.rows.reshape(-1), rows) and c2.roots == []
    # resident: roots through the in-kernel exchange, data sharded over the devices
    for rep in range(3):
        res, comm_r = MultilinearZip.commit_resident(pp, poly, mctx)
        assert b"".join(comm_r.roots) == roots.tobytes(), rep
        if rep < 2:
            res.free()
    assert np.array_equal(res.rows().reshape(-1), rows)
    assert np.array_equal(res.layers().reshape(-1), layers)
    assert np.array_equal(res.rows(3, 5).reshape(-1), rows.reshape(num_rows, -1)[3:8].reshape(-1))
    single, _ = MultilinearZip.commit_resident(pp, poly, ctx)
    cols = np.array([1, 0, cw - 1, cw // 3, 1], dtype=np.uint32)
    v_s, p_s = single.open_columns(cols)
    v_m, p_m = res.open_columns(cols)
    assert np.array_equal(v_s, v_m) and np.array_equal(p_s, p_m)
    assert single.open_columns_wire(cols) == res.open_columns_wire(cols)
    co = np.random.default_rng(5).integers(-(1 << 63), (1 << 63) - 1, size=num_rows, dtype=np.int64)
    assert np.array_equal(single.combine_rows(co, 8), res.combine_rows(co, 8))
    # every device that owns rows holds ALL roots after the exchange
    import torch
    from zinc_b200 import _native as nat

    for g in range(mctx.num_devices):
        ptr = nat.lib().zipgpu_mgpu_data_roots_device(res.handle, g)
        if not ptr:
            continue
        buf = (C.c_uint8 * (num_rows * 32))()
        dev_id = nat.lib().zipgpu_ctx_device(C.c_void_p(nat.lib().zipgpu_mgpu_ctx(mctx.handle, g)))
        with torch.cuda.device(dev_id):
            raw = type("R", (), {})()
            raw.__cuda_array_interface__ = {"shape": (num_rows * 32,), "typestr": "|u1", "data": (ptr, False), "version": 2}
            got = torch.as_tensor(raw, device=torch.device("cuda", dev_id)).cpu().numpy()
        assert got.tobytes() == roots.tobytes(), f"device {g}"
    single.free()
    res.free()
    # batch_commit: polynomial p -> device p mod n, results in input order
    polys = [DenseMultilinearExtension.rand(nv, np.random.default_rng(50 + i)) for i in range(5)]
    outs = MultilinearZip.batch_commit(pp, polys, mctx)
    for p_, (d_, c_) in zip(polys, outs):
        rc, rows_p, _, roots_p = oracle.commit_mt(p_.evaluations.reshape(-1), num_rows, row_len, 2, 0, 0, p1, p2,
                                                  threads=8, faithful=False)
        assert rc == 0 and np.array_equal(d_.rows.reshape(-1), rows_p) and b"".join(c_.roots) == roots_p.tobytes()
    # uneven split: 3 rows of a tiny code over the devices (some ranks may get nothing)
    from zinc_b200 import RaaCode, ZipTypes

    tcode = RaaCode.with_permutations(ZipTypes(), 8, 2, oracle.perm_from_seed(16, 1), oracle.perm_from_seed(16, 2))
    tpp = MultilinearZipParams.new(5, 3, tcode)
    ev = np.arange(1, 25, dtype=np.int64)
    tpoly = DenseMultilinearExtension(ev.view(np.uint64).reshape(-1, 1), 5)
    tres, tcomm = MultilinearZip.commit_resident(tpp, tpoly, mctx)
    rc, _, _, troots = oracle.commit(ev.view(np.uint64), 3, 8, 2, oracle.perm_from_seed(16, 1), oracle.perm_from_seed(16, 2))
    assert rc == 0 and b"".join(tcomm.roots) == troots.tobytes()
    tres.free()


def test_mgpu_one_device(oracle, ctx):
    from zinc_b200 import MultiContext

    m = MultiContext(n=1)
    assert m.num_devices == 1
    try:
        _check_multi(m, ctx, oracle)
        assert m.launch_count > 0
    finally:
        m.close()


def test_mgpu_two_ranks_on_one_device(oracle, ctx, monkeypatch):
    """the whole multi-rank path (row shards, worker threads, the in-kernel roots exchange between two contexts and
    their publish/wait handshake) on a one-GPU box: both ranks live on device 0"""
    from zinc_b200 import MultiContext

    monkeypatch.setenv("ZIPGPU_MGPU_ALLOW_DUPLICATE", "1")
    monkeypatch.setenv("ZIPGPU_PEER_TIMEOUT_MS", "10000")
    m = MultiContext(devices=[ctx.device, ctx.device])
    assert m.num_devices == 2
    try:
        _check_multi(m, ctx, oracle)
    finally:
        m.close()


def test_mgpu_all_devices(oracle, ctx):
    if _device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from zinc_b200 import MultiContext

    m = MultiContext()
    assert m.num_devices == _device_count()
    try:
        _check_multi(m, ctx, oracle, nv=16)
        _check_multi(m, ctx, oracle, nv=20)
    finally:
        m.close()


def test_peer_timeout_is_an_error_not_a_trap(oracle, ctx, monkeypatch):
    """a rank whose peer never shows up gets ZIPGPU_ERR_PEER_TIMEOUT after the bounded wait and the context stays
    usable (round 1 trapped, which poisons the context)"""
    import torch

    from zinc_b200 import _native as nat

    monkeypatch.setenv("ZIPGPU_PEER_TIMEOUT_MS", "200")
    L = nat.lib()
    # rank 0 of a two-rank exchange whose rank 1 lives in the same process but never runs a step
    a, b = C.c_void_p(), C.c_void_p()
    from zinc_b200 import Context

    ctx2 = Context(ctx.device)
    nat.check(L.zipgpu_peer_roots_create(ctx.handle, 64, 0, 2, C.byref(a), None))
    nat.check(L.zipgpu_peer_roots_create(ctx2.handle, 64, 1, 2, C.byref(b), None))
    arr = (C.c_void_p * 2)(a, b)
    nat.check(L.zipgpu_peer_roots_connect_local(arr, 2))
    local = torch.zeros(32 * 32, dtype=torch.uint8, device=torch.device("cuda", ctx.device))
    out = C.c_void_p()
    nat.check(L.zipgpu_peer_roots_allgather(a, 0, 32, C.c_void_p(local.data_ptr()), None, C.byref(out)))
    ctx.sync()
    assert L.zipgpu_peer_roots_status(a) == nat.ERR_PEER_TIMEOUT
    assert b"rank 1" in L.zipgpu_last_error()
    # the context is still alive
    code, row_len, num_rows, cw, p1, p2 = _code(8, MOCK_SEEDS, oracle)
    from zinc_b200 import DenseMultilinearExtension, MultilinearZip, MultilinearZipParams

    pp = MultilinearZipParams.new(8, num_rows, code)
    MultilinearZip.commit(pp, DenseMultilinearExtension.rand(8, np.random.default_rng(1)), ctx)
    L.zipgpu_peer_roots_destroy(a)
    L.zipgpu_peer_roots_destroy(b)
    ctx2.close()


def test_context_close_releases_handles(oracle):
    """codes and resident data are owned by the context and die with it; using them afterwards raises
    instead of touching freed device memory; a new context never sees a stale cached handle"""
    from zinc_b200 import Context, DenseMultilinearExtension, MultilinearZip, MultilinearZipParams
    from zinc_b200.zip import ZipGpuClosed

    code, row_len, num_rows, cw, p1, p2 = _code(8, MOCK_SEEDS, oracle)
    pp = MultilinearZipParams.new(8, num_rows, code)
    poly = DenseMultilinearExtension.rand(8, np.random.default_rng(2))
    c1 = Context(0)
    res, comm1 = MultilinearZip.commit_resident(pp, poly, c1)
    c1.close()
    with pytest.raises(ZipGpuClosed):
        res.rows()
    with pytest.raises(ZipGpuClosed):
        MultilinearZip.commit(pp, poly, c1)
    c2 = Context(0)
    _, comm2 = MultilinearZip.commit(pp, poly, c2)  # builds a fresh native code in c2
    assert comm1.roots == comm2.roots
    c2.close()
