"""GPU parity at BASELINE.json's sizes and configs: nv = 20..24 commits against the multithreaded C oracle, the batched
64 x 2^18 config, prover-flow shapes with live-transcript-like seeds, and the proximity-test row combination."""
import ctypes as C

import numpy as np
import pytest

from helpers import I64_MAX, I64_MIN, KECCAK_SEEDS, shape_for

pytestmark = pytest.mark.gpu


def _code(nv, seeds, oracle):
    from zinc_b200 import RaaCode, ZipTypes

    row_len, num_rows, cw = shape_for(nv)
    p1, p2 = oracle.perm_from_seed(cw, seeds[0]), oracle.perm_from_seed(cw, seeds[1])
    return RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2), row_len, num_rows, cw, p1, p2


@pytest.mark.parametrize("nv", [22, 24])
def test_commit_resident_full_size_matches_oracle(nv, oracle, ctx):
    """BASELINE configs[1..2] shape rule at nv = 22 / 24: every root, plus the codewords and layers of sampled row
    ranges read back from the device-resident prover data, bit for bit against the oracle."""
    from zinc_b200 import DenseMultilinearExtension, MultilinearZip, MultilinearZipParams

    code, row_len, num_rows, cw, p1, p2 = _code(nv, KECCAK_SEEDS, oracle)
    pp = MultilinearZipParams.new(nv, num_rows, code)
    evals = np.random.default_rng(0x21C0 + nv).integers(0, 1 << 64, size=1 << nv, dtype=np.uint64)
    poly = DenseMultilinearExtension.from_evaluations_vec(nv, evals)
    res, comm = MultilinearZip.commit_resident(pp, poly, ctx)
    rc, _, _, roots = oracle.commit_mt(evals, num_rows, row_len, 2, 0, 0, p1, p2, threads=16, faithful=False,
                                       want_rows=False, want_layers=False)
    assert rc == 0
    assert b"".join(comm.roots) == roots.tobytes(), "roots differ"
    per = (2 * cw - 2) * 32
    for r0, n in ((0, 3), (num_rows // 2 - 1, 2), (num_rows - 2, 2)):
        rc, rows, layers, _ = oracle.commit(evals[r0 * row_len:(r0 + n) * row_len], n, row_len, 2, p1, p2)
        assert rc == 0
        assert np.array_equal(res.rows(r0, n).reshape(-1), rows), f"codewords differ at rows {r0}..{r0 + n}"
        assert np.array_equal(res.layers(r0, n).reshape(-1), layers), f"layers differ at rows {r0}..{r0 + n}"
    res.free()


def test_commit_linearity_at_scale(oracle, ctx):
    """commit.rs:559-583 at nv = 20: encode(3*r1 + 5*r2) == 3*encode(r1) + 5*encode(r2), checked on sampled entries
    with Python integers (a size-independent property of the linear code)"""
    from oracle import pyoracle as po
    from zinc_b200 import DenseMultilinearExtension, MultilinearZip, MultilinearZipParams

    nv = 20
    code, row_len, num_rows, cw, p1, p2 = _code(nv, KECCAK_SEEDS, oracle)
    pp = MultilinearZipParams.new(nv, num_rows, code)
    rng = np.random.default_rng(5)
    a = rng.integers(-(1 << 59), 1 << 59, size=1 << nv, dtype=np.int64)
    b = rng.integers(-(1 << 59), 1 << 59, size=1 << nv, dtype=np.int64)
    enc = lambda v: MultilinearZip.commit_no_merkle(pp, DenseMultilinearExtension.from_evaluations_vec(nv, v), ctx)[0].rows
    ra, rb, rc_ = enc(a), enc(b), enc(3 * a + 5 * b)
    idx = rng.integers(0, num_rows * cw, size=2000)
    for i in idx:
        va, vb, vc = (po.to_signed([int(w) for w in r[i]]) for r in (ra, rb, rc_))
        assert vc == 3 * va + 5 * vb


def test_batched_commit_64_polys_nv18(oracle, ctx):
    """BASELINE configs[3]: 64 independent 2^18 MLEs sharing one pp (commit.rs:134-142), roots only, one submission"""
    from zinc_b200 import _native as nat

    nv, k = 18, 64
    code, row_len, num_rows, cw, p1, p2 = _code(nv, KECCAK_SEEDS, oracle)
    h = code.native(ctx, 1, 4)
    rng = np.random.default_rng(18)
    evals = [rng.integers(0, 1 << 64, size=1 << nv, dtype=np.uint64) for _ in range(k)]
    roots = [np.zeros(num_rows * 32, dtype=np.uint8) for _ in range(k)]
    arr = lambda xs: (C.c_void_p * k)(*[nat.ptr(x) for x in xs])
    nat.check(nat.lib().zipgpu_batch_commit(h, k, num_rows, arr(evals), None, None, arr(roots)))
    for p in range(k):
        rc, _, _, exp = oracle.commit_mt(evals[p], num_rows, row_len, 2, 0, 0, p1, p2, threads=16, faithful=False,
                                         want_rows=False, want_layers=False)
        assert rc == 0 and np.array_equal(roots[p], exp), f"poly {p}"


@pytest.mark.parametrize("nv", [3, 12, 13, 14, 15, 16])
def test_prover_flow_shapes_with_live_seeds(nv, oracle, ctx):
    """BASELINE configs[4]: the commit inside Prover::prove (zinc/prover.rs:305-328) draws its seeds from the live
    Fiat-Shamir transcript, so they differ per proof: arbitrary 64-bit seeds, z-vector sizes 2^12..2^16 and the
    8-entry z of examples/simple_r1cs.rs."""
    from zinc_b200 import DenseMultilinearExtension, MultilinearZip, MultilinearZipParams

    rng = np.random.default_rng(1000 + nv)
    seeds = tuple(int(x) for x in rng.integers(0, 1 << 64, size=2, dtype=np.uint64))
    code, row_len, num_rows, cw, p1, p2 = _code(nv, seeds, oracle)
    pp = MultilinearZipParams.new(nv, num_rows, code)
    evals = rng.integers(0, 1 << 64, size=1 << nv, dtype=np.uint64)
    data, comm = MultilinearZip.commit(pp, DenseMultilinearExtension.from_evaluations_vec(nv, evals), ctx)
    rc, rows, layers, roots = oracle.commit(evals, num_rows, row_len, 2, p1, p2)
    assert rc == 0 and np.array_equal(data.rows.reshape(-1), rows)
    assert np.array_equal(np.concatenate([t.layers.reshape(-1) for t in data.rows_merkle_trees]), layers)
    assert b"".join(comm.roots) == roots.tobytes()


@pytest.mark.parametrize("nv", [4, 9, 12, 16])
def test_combine_rows_matches_bigint(nv, oracle, ctx):
    """open_z.rs:100-113 / zip/utils.rs:94-127: u' = sum_i coeff_i * row_i in Int<8>, against Python integers"""
    from oracle import pyoracle as po
    from zinc_b200 import DenseMultilinearExtension, MultilinearZip, MultilinearZipParams

    code, row_len, num_rows, cw, p1, p2 = _code(nv, KECCAK_SEEDS, oracle)
    pp = MultilinearZipParams.new(nv, num_rows, code)
    rng = np.random.default_rng(77 + nv)
    for case in ("random", "extremes"):
        if case == "random":
            ev = rng.integers(I64_MIN, I64_MAX, size=1 << nv, dtype=np.int64, endpoint=True)
            co = rng.integers(I64_MIN, I64_MAX, size=num_rows, dtype=np.int64, endpoint=True)
        else:  # |sum| as large as it gets: i64::MIN * i64::MIN in every term, and alternating signs
            ev = np.full(1 << nv, I64_MIN, dtype=np.int64)
            ev[1::3] = I64_MAX
            co = np.full(num_rows, I64_MIN, dtype=np.int64)
            co[::2] = I64_MAX if nv % 2 else I64_MIN
        res, _ = MultilinearZip.commit_resident(pp, DenseMultilinearExtension.from_evaluations_vec(nv, ev), ctx)
        got = res.combine_rows(co, 8)
        exp = po.combine_rows([int(c) for c in co], [int(e) for e in ev], row_len, 8)
        assert [po.to_signed([int(w) for w in g]) for g in got] == exp, f"nv={nv} {case}"
        res.free()


def test_combine_rows_rejects_narrow_output(ctx):
    from zinc_b200 import _native as nat

    assert nat.lib().zipgpu_combine_rows_device(ctx.handle, 4, 4, 1, 1, 2, 1, None) == nat.ERR_WIDTH


def _commit_device(ctx, code, num_rows, cw, evals):
    """zipgpu_commit_device on torch device buffers -> (rows, layers, roots) as numpy"""
    import torch

    from zinc_b200 import _native as nat

    depth = cw.bit_length() - 1
    dev = torch.device("cuda", ctx.device)
    d_ev = torch.from_numpy(evals.view(np.int64)).to(dev)
    d_rows = torch.empty(num_rows * cw * 4, dtype=torch.int64, device=dev)
    d_lay = torch.empty(num_rows * ((2 << depth) - 2) * 32, dtype=torch.uint8, device=dev)
    d_roots = torch.empty(num_rows * 32, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    nat.check(nat.lib().zipgpu_commit_device(code.native(ctx, 1, 4), num_rows, d_ev.data_ptr(), d_rows.data_ptr(),
                                             d_lay.data_ptr(), d_roots.data_ptr(), None))
    ctx.sync()
    return d_rows.cpu().numpy().view(np.uint64), d_lay.cpu().numpy(), d_roots.cpu().numpy()


@pytest.mark.parametrize("nv", [16, 18, 20])
def test_fused_commit_kernel_forced_at_small_shapes(nv, oracle, ctx, monkeypatch):
    """the fused commit kernel (encode + leaf hashes + the lowest tree levels in one launch) for every entries-per-
    thread variant (E = 4 at nv 16, E = 8 at nv 18/20), forced on, against the oracle and the two-kernel path"""
    code, row_len, num_rows, cw, p1, p2 = _code(nv, KECCAK_SEEDS, oracle)
    evals = np.random.default_rng(nv).integers(0, 1 << 64, size=1 << nv, dtype=np.uint64)
    rc, rows, layers, roots = oracle.commit_mt(evals, num_rows, row_len, 2, 0, 0, p1, p2, threads=8, faithful=False)
    assert rc == 0
    monkeypatch.setenv("ZIPGPU_FUSE_MIN_ROWS", "1")
    launches0 = ctx.launch_count
    f_rows, f_lay, f_roots = _commit_device(ctx, code, num_rows, cw, evals)
    fused_launches = ctx.launch_count - launches0
    monkeypatch.setenv("ZIPGPU_NO_FUSE", "1")
    monkeypatch.setenv("ZIPGPU_NO_CTA_TREE", "1")  # compare with the subtree-pass kernels (small trees default to the CTA-tree path)
    launches0 = ctx.launch_count
    u_rows, u_lay, u_roots = _commit_device(ctx, code, num_rows, cw, evals)
    assert fused_launches <= ctx.launch_count - launches0  # the leaf pass is folded into the encoder launch
    for got in ((f_rows, f_lay, f_roots), (u_rows, u_lay, u_roots)):
        assert np.array_equal(got[0], rows) and np.array_equal(got[1], layers) and np.array_equal(got[2], roots)


def test_fused_commit_kernel_default_at_nv22(oracle, ctx):
    """zipgpu_commit_device takes the fused kernel by itself from 10 x SMs rows: rows, every layer and the roots"""
    nv = 22
    code, row_len, num_rows, cw, p1, p2 = _code(nv, KECCAK_SEEDS, oracle)
    evals = np.random.default_rng(nv).integers(0, 1 << 64, size=1 << nv, dtype=np.uint64)
    rc, rows, layers, roots = oracle.commit_mt(evals, num_rows, row_len, 2, 0, 0, p1, p2, threads=16, faithful=False)
    assert rc == 0
    g_rows, g_lay, g_roots = _commit_device(ctx, code, num_rows, cw, evals)
    assert np.array_equal(g_roots, roots)
    assert np.array_equal(g_rows, rows)
    assert np.array_equal(g_lay, layers)


@pytest.mark.parametrize("row_len", [64, 1024, 4096, 8192])
def test_encode_wide_to_M_matches_oracle(row_len, oracle, ctx):
    """LinearCode::encode = encode_wide::<N, M> (code.rs:35-37; the verifier re-encodes the combined row in Int<8>,
    verify_z.rs:74-78): 16-word outputs take the generic (run-time width, CTA-wide write-out) encoder variant"""
    from zinc_b200 import RaaCode, ZipTypes

    cw = 2 * row_len
    p1, p2 = oracle.perm_from_seed(cw, 21), oracle.perm_from_seed(cw, 22)
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2)
    row = np.random.default_rng(row_len).integers(0, 1 << 64, size=row_len, dtype=np.uint64)
    rc, exp = oracle.encode_rows(row, 1, row_len, 2, p1, p2, in_limbs=1, out_limbs=8)
    assert rc == 0
    assert np.array_equal(code.encode(row, ctx).reshape(-1), exp)


@pytest.mark.parametrize("nv", [22, 23])
def test_zero_copy_commit_opt_in(nv, oracle, ctx, monkeypatch):
    """ZIPGPU_ZEROCOPY=1: the fused commit kernel reads pinned host evaluations in place over PCIe and keeps a copy in
    HBM for the opening phase; same roots, rows, layers, and the proximity row combination sees the copied evals"""
    import torch

    from oracle import pyoracle as po
    from zinc_b200 import DenseMultilinearExtension, MultilinearZip, MultilinearZipParams

    code, row_len, num_rows, cw, p1, p2 = _code(nv, KECCAK_SEEDS, oracle)
    pp = MultilinearZipParams.new(nv, num_rows, code)
    pinned = torch.empty((1 << nv, 1), dtype=torch.int64).pin_memory()
    evals = np.random.default_rng(3).integers(0, 1 << 64, size=1 << nv, dtype=np.uint64)
    pinned.numpy().view(np.uint64)[:, 0] = evals
    monkeypatch.setenv("ZIPGPU_ZEROCOPY", "1")
    poly = DenseMultilinearExtension(pinned.numpy().view(np.uint64), nv)
    res, comm = MultilinearZip.commit_resident(pp, poly, ctx)
    rc, _, _, roots = oracle.commit_mt(evals, num_rows, row_len, 2, 0, 0, p1, p2, threads=16, faithful=False,
                                       want_rows=False, want_layers=False)
    assert rc == 0 and b"".join(comm.roots) == roots.tobytes()
    rc, rows, layers, _ = oracle.commit(evals[5 * row_len:7 * row_len], 2, row_len, 2, p1, p2)
    assert np.array_equal(res.rows(5, 2).reshape(-1), rows) and np.array_equal(res.layers(5, 2).reshape(-1), layers)
    co = np.zeros(num_rows, dtype=np.int64)
    co[[0, 7, num_rows - 1]] = (3, -5, 11)  # sparse coefficients keep the big-int check cheap
    got = res.combine_rows(co, 8)
    ev = evals.view(np.int64)
    for col in (0, 1, row_len - 1):
        exp = 3 * int(ev[col]) - 5 * int(ev[7 * row_len + col]) + 11 * int(ev[(num_rows - 1) * row_len + col])
        assert po.to_signed([int(w) for w in got[col]]) == exp
    res.free()


def test_two_host_threads_share_one_context(oracle, ctx):
    """encoding_is_consistent_across_threads (commit.rs:439-470) in the GPU setting: the C ABI is thread-safe per
    context, so concurrent commits from two host threads on ONE context give the single-threaded results"""
    import threading

    from zinc_b200 import DenseMultilinearExtension, MultilinearZip, MultilinearZipParams

    nv = 16
    code, row_len, num_rows, cw, p1, p2 = _code(nv, KECCAK_SEEDS, oracle)
    pp = MultilinearZipParams.new(nv, num_rows, code)
    code.native(ctx, 1, 4)  # create the per-pp state once, before the threads start
    polys = [DenseMultilinearExtension.rand(nv, np.random.default_rng(100 + i)) for i in range(4)]
    expect = []
    for p in polys:
        rc, rows, layers, roots = oracle.commit(p.evaluations.reshape(-1), num_rows, row_len, 2, p1, p2)
        assert rc == 0
        expect.append((rows, roots))
    results, errors = {}, []

    def worker(tid):
        try:
            for rep in range(6):
                for i in range(tid, len(polys), 2):
                    data, comm = MultilinearZip.commit(pp, polys[i], ctx)
                    results[(i, rep)] = (data.rows.reshape(-1).copy(), b"".join(comm.roots))
        except Exception as ex:  # surfaced below
            errors.append(ex)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for (i, rep), (rows, roots) in results.items():
        assert np.array_equal(rows, expect[i][0]) and roots == expect[i][1].tobytes(), (i, rep)


def test_open_columns_wire_format(oracle, ctx):
    """the proof-stream bytes of the column openings: write_integers (LE u64 limbs) then per row be64(depth) || path
    (pcs_transcript.rs:115-135,198-211), assembled here from the structured openings and from the oracle's layers"""
    from zinc_b200 import DenseMultilinearExtension, MultilinearZip, MultilinearZipParams

    nv = 10
    code, row_len, num_rows, cw, p1, p2 = _code(nv, KECCAK_SEEDS, oracle)
    depth = cw.bit_length() - 1
    pp = MultilinearZipParams.new(nv, num_rows, code)
    poly = DenseMultilinearExtension.rand(nv, np.random.default_rng(4))
    res, _ = MultilinearZip.commit_resident(pp, poly, ctx)
    cols = np.array([5, 0, cw - 1, 5], dtype=np.uint32)
    vals, paths = res.open_columns(cols)
    rc, rows, layers, _ = oracle.commit(poly.evaluations.reshape(-1), num_rows, row_len, 2, p1, p2)
    assert rc == 0 and np.array_equal(vals[0, :, :].reshape(-1), rows.reshape(num_rows, cw, 4)[:, 5, :].reshape(-1))
    expect = b""
    for ci in range(cols.size):
        expect += vals[ci].astype("<u8").tobytes()
        for r in range(num_rows):
            expect += depth.to_bytes(8, "big") + paths[ci, r].tobytes()
    got = res.open_columns_wire(cols)
    assert len(got) == len(expect) and got == expect
    # the same bytes as a view of the context's pinned proof-stream buffer (reused and grown across calls)
    view = res.open_columns_wire_view(cols[:2])
    assert view.tobytes() == expect[:len(expect) // 2]
    view = res.open_columns_wire_view(np.concatenate([cols] * 300))
    assert view.size == 300 * len(expect) and view[-len(expect):].tobytes() == expect
    assert res.open_columns_wire_view(np.zeros(0, dtype=np.uint32)).size == 0
    res.free()


@pytest.mark.parametrize("fused", [False, True])
def test_longest_codeword_shape(fused, oracle, ctx, monkeypatch):
    """cw = 16384 (row_len 8192: the nv = 25 / 26 shape, one 1024-thread CTA per SM with 16 entries per thread): a few
    hundred rows through both the two-kernel and the fused commit path against the oracle"""
    row_len, cw, num_rows = 8192, 16384, 300
    from zinc_b200 import RaaCode, ZipTypes

    p1, p2 = oracle.perm_from_seed(cw, KECCAK_SEEDS[0]), oracle.perm_from_seed(cw, KECCAK_SEEDS[1])
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2)
    evals = np.random.default_rng(26).integers(0, 1 << 64, size=num_rows * row_len, dtype=np.uint64)
    rc, rows, layers, roots = oracle.commit_mt(evals, num_rows, row_len, 2, 0, 0, p1, p2, threads=16, faithful=False)
    assert rc == 0
    if fused:
        monkeypatch.setenv("ZIPGPU_FUSE_MIN_ROWS", "1")
    else:
        monkeypatch.setenv("ZIPGPU_NO_FUSE", "1")
    g_rows, g_lay, g_roots = _commit_device(ctx, code, num_rows, cw, evals)
    assert np.array_equal(g_roots, roots) and np.array_equal(g_rows, rows) and np.array_equal(g_lay, layers)


@pytest.mark.parametrize("row_len,num_rows", [(4096, 148), (4096, 1000), (4096, 1333), (2048, 100), (2048, 777), (2048, 2048),
                                              (1024, 200), (1024, 1025), (1024, 3000), (512, 300), (512, 4097), (256, 256), (256, 1000)])
def test_warp_specialised_commit_kernel(row_len, num_rows, oracle, ctx, monkeypatch):
    """cw = 8192 / 4096 / 2048 (nv = 19 .. 24) take the warp-specialised commit kernel (per CTA one thread group encodes
    into alternating plane sets while the other hashes): codewords, every layer and the roots against the oracle, for row
    counts below / above the dynamic-claiming threshold and not a multiple of the grid; and against the other paths"""
    from zinc_b200 import RaaCode, ZipTypes

    cw = 2 * row_len
    p1, p2 = oracle.perm_from_seed(cw, KECCAK_SEEDS[0]), oracle.perm_from_seed(cw, KECCAK_SEEDS[1])
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2)
    evals = np.random.default_rng(num_rows).integers(0, 1 << 64, size=num_rows * row_len, dtype=np.uint64)
    rc, rows, layers, roots = oracle.commit_mt(evals, num_rows, row_len, 2, 0, 0, p1, p2, threads=16, faithful=False)
    assert rc == 0
    monkeypatch.setenv("ZIPGPU_FUSE_MIN_ROWS", "1")
    monkeypatch.setenv("ZIPGPU_WS16K_MIN_ROWS", "1")  # the product uses this kernel from 2048 rows
    for knob in (None, "ZIPGPU_NO_WS", "ZIPGPU_NO_FUSE"):
        if knob:
            monkeypatch.setenv(knob, "1")
        g_rows, g_lay, g_roots = _commit_device(ctx, code, num_rows, cw, evals)
        assert np.array_equal(g_roots, roots), knob
        assert np.array_equal(g_rows, rows), knob
        assert np.array_equal(g_lay, layers), knob


@pytest.mark.parametrize("num_rows", [1, 2, 73, 74, 75, 149, 500, 1211])
def test_cw16384_cluster_commit_kernel(num_rows, oracle, ctx, monkeypatch):
    """cw = 16384 (nv = 25 / 26) through the opt-in commit_wsc_kernel -- a 2-CTA cluster per row, the two plane sets split over the
    shared memories of the SM pair, the encoder's gathers and scan totals crossing over through distributed shared
    memory -- for row counts below, at and above one row per cluster and with dynamic claiming; rows, layers and roots
    against the oracle, twice"""
    from zinc_b200 import RaaCode, ZipTypes

    row_len, cw = 8192, 16384
    p1, p2 = oracle.perm_from_seed(cw, KECCAK_SEEDS[0]), oracle.perm_from_seed(cw, KECCAK_SEEDS[1])
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2)
    evals = np.random.default_rng(num_rows + 7).integers(0, 1 << 64, size=num_rows * row_len, dtype=np.uint64)
    if num_rows >= 2:
        evals[:row_len] = np.uint64((1 << 63) - 1)          # i64::MAX row
        evals[row_len:2 * row_len] = np.uint64(1 << 63)      # i64::MIN row
    rc, rows, layers, roots = oracle.commit_mt(evals, num_rows, row_len, 2, 0, 0, p1, p2, threads=16, faithful=False)
    assert rc == 0
    monkeypatch.setenv("ZIPGPU_FUSE_MIN_ROWS", "1")
    monkeypatch.setenv("ZIPGPU_WSC", "1")  # opt-in: measured slower than the single-SM forms
    monkeypatch.setenv("ZIPGPU_WSC_MIN_ROWS", "1")
    for rep in range(2):
        g_rows, g_lay, g_roots = _commit_device(ctx, code, num_rows, cw, evals)
        assert np.array_equal(g_rows, rows), rep
        assert np.array_equal(g_lay, layers), rep
        assert np.array_equal(g_roots, roots), rep


@pytest.mark.parametrize("units", [1, 2])
@pytest.mark.parametrize("row_len,num_rows", [(4096, 37), (4096, 391), (2048, 523), (1024, 97), (1024, 1500), (512, 611),
                                              (256, 59), (256, 1777)])
def test_warp_specialised_commit_kernel_sub_row_units(row_len, num_rows, units, oracle, ctx, monkeypatch):
    """the sub-row work units of the warp-specialised commit kernel (a row hashed as 2 units, the units -- not
    the rows -- split statically and evenly over the CTAs; rows shared by two CTAs are encoded by both and each writes
    its part of the codeword): every unit count forced at row counts where CTAs get less than one, exactly one and
    several units, with shares that start and end in the middle of a row"""
    from zinc_b200 import RaaCode, ZipTypes

    cw = 2 * row_len
    p1, p2 = oracle.perm_from_seed(cw, KECCAK_SEEDS[0]), oracle.perm_from_seed(cw, KECCAK_SEEDS[1])
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2)
    evals = np.random.default_rng(num_rows + units).integers(0, 1 << 64, size=num_rows * row_len, dtype=np.uint64)
    rc, rows, layers, roots = oracle.commit_mt(evals, num_rows, row_len, 2, 0, 0, p1, p2, threads=16, faithful=False)
    assert rc == 0
    monkeypatch.setenv("ZIPGPU_FUSE_MIN_ROWS", "1")
    monkeypatch.setenv("ZIPGPU_WS_UNITS", str(units))
    g_rows, g_lay, g_roots = _commit_device(ctx, code, num_rows, cw, evals)
    assert np.array_equal(g_roots, roots)
    assert np.array_equal(g_rows, rows)
    assert np.array_equal(g_lay, layers)


@pytest.mark.parametrize("units", [1, 2])
@pytest.mark.parametrize("row_len,num_rows", [(4096, 1), (4096, 149), (4096, 517), (4096, 1400), (2048, 523), (2048, 2300),
                                              (1024, 97), (1024, 3000), (512, 611), (256, 59), (256, 4500)])
def test_warp_specialised_commit_kernel_tree_tops(row_len, num_rows, units, oracle, ctx, monkeypatch):
    """whole trees in the one launch: the epilogue of the warp-specialised commit kernel that finishes the trees of the
    units a CTA hashed (4 nodes per thread in registers, then level by level through shared memory; rows whose two
    units were hashed by two CTAs are joined by whichever arrives second) -- forced for every shape and unit count, at
    row counts with less than one, one and several batches of 8 units per CTA and split rows at the CTA boundaries;
    layers, rows and roots against the oracle, twice (the boundary flags re-arm themselves), and ONE launch each"""
    from zinc_b200 import RaaCode, ZipTypes

    cw = 2 * row_len
    p1, p2 = oracle.perm_from_seed(cw, KECCAK_SEEDS[0]), oracle.perm_from_seed(cw, KECCAK_SEEDS[1])
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2)
    evals = np.random.default_rng(num_rows * 3 + units).integers(0, 1 << 64, size=num_rows * row_len, dtype=np.uint64)
    rc, rows, layers, roots = oracle.commit_mt(evals, num_rows, row_len, 2, 0, 0, p1, p2, threads=16, faithful=False)
    assert rc == 0
    monkeypatch.setenv("ZIPGPU_FUSE_MIN_ROWS", "1")
    monkeypatch.setenv("ZIPGPU_WS_UNITS", str(units))
    monkeypatch.setenv("ZIPGPU_WS_TOPS", "1")
    for rep in range(2):
        l0 = ctx.launch_count
        g_rows, g_lay, g_roots = _commit_device(ctx, code, num_rows, cw, evals)
        assert ctx.launch_count - l0 == 1, "the commit must be one launch"
        assert np.array_equal(g_roots, roots), rep
        assert np.array_equal(g_rows, rows), rep
        assert np.array_equal(g_lay, layers), rep
    monkeypatch.setenv("ZIPGPU_WS_TOPS", "0")
    l0 = ctx.launch_count
    g_rows, g_lay, g_roots = _commit_device(ctx, code, num_rows, cw, evals)
    assert ctx.launch_count - l0 >= 2
    assert np.array_equal(g_roots, roots) and np.array_equal(g_lay, layers)


def test_host_alloc_huge_page_backed_pinned_memory(oracle, ctx):
    """zipgpu_host_alloc: page-locked memory on 2 MiB pages from 2 MiB up (cudaHostAlloc below); usable as input and
    output of the host-pointer entry points, freed with zipgpu_host_free (NULL is fine)"""
    from zinc_b200 import _native as nat

    L = nat.lib()
    code, row_len, num_rows, cw, p1, p2 = _code(20, KECCAK_SEEDS, oracle)
    n = num_rows * row_len
    blocks = []
    for nbytes in (n * 8, num_rows * 32, (2 << 20) + 4096, 1):
        p = C.c_void_p()
        nat.check(L.zipgpu_host_alloc(nbytes, C.byref(p)))
        assert p.value and p.value % 64 == 0
        blocks.append(p)
    assert blocks[0].value % (2 << 20) == 0  # the large block starts on a 2 MiB boundary
    ev = np.frombuffer((C.c_uint64 * n).from_address(blocks[0].value), dtype=np.uint64)
    ev[:] = np.random.default_rng(9).integers(0, 1 << 64, size=n, dtype=np.uint64)
    roots = np.frombuffer((C.c_uint8 * (num_rows * 32)).from_address(blocks[1].value), dtype=np.uint8)
    hd = C.c_void_p()
    nat.check(L.zipgpu_commit_resident(code.native(ctx, 1, 4), num_rows, blocks[0], blocks[1], C.byref(hd)))
    L.zipgpu_data_free(hd)
    rc, _, _, want = oracle.commit_mt(ev.copy(), num_rows, row_len, 2, 0, 0, p1, p2, threads=8, faithful=False,
                                      want_rows=False, want_layers=False)
    assert rc == 0 and roots.tobytes() == want.tobytes()
    del ev, roots
    for p in blocks:
        nat.check(L.zipgpu_host_free(p))
    nat.check(L.zipgpu_host_free(None))


def test_peer_roots_allgather_single_rank(ctx):
    """the peer-memory roots exchange degenerates to a copy + self-signal on one GPU (N > 1: scripts/strong_scaling.py
    --p2p and test_peer_roots_two_gpus below); two steps exercise the double buffering"""
    import torch

    from zinc_b200.dist import PeerRoots

    dev = torch.device("cuda:0")
    pr = PeerRoots(ctx, 512)
    for step in range(3):
        local = torch.randint(0, 256, (512 * 32,), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        ptr = pr.allgather(0, 512, local.data_ptr())
        ctx.sync()
        assert torch.equal(pr.tensor(ptr), local), step
    pr.close()


def _peer_worker(rank, world, port, q):
    import os

    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    import torch
    import torch.distributed as dist

    from zinc_b200 import Context
    from zinc_b200.dist import PeerRoots, shard_range

    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = Context(rank)
    total = 1024
    pr = PeerRoots(ctx, total)
    begin, count = shard_range(total, rank, world)
    ok = True
    for step in range(4):
        gen = torch.Generator().manual_seed(100 + step)
        everything = torch.randint(0, 256, (total * 32,), dtype=torch.uint8, generator=gen)
        local = everything[begin * 32:(begin + count) * 32].cuda()
        torch.cuda.synchronize()
        ptr = pr.allgather(begin, count, local.data_ptr())
        ctx.sync()
        ok = ok and bool(torch.equal(pr.tensor(ptr).cpu(), everything))
    dist.barrier()
    pr.close()
    # a row-sharded commit whose roots are gathered by the peer-memory kernel == the single-GPU commit
    import numpy as np

    from zinc_b200 import DenseMultilinearExtension, MultilinearZip, RaaCode, ZipTypes, shuffle_seeded_indices
    from zinc_b200.dist import sharded_commit

    nv, row_len, num_rows, cw = 14, 128, 128, 256
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, 1), shuffle_seeded_indices(cw, 2))
    pp = MultilinearZip.setup(1 << nv, code)
    poly = DenseMultilinearExtension.rand(nv, np.random.default_rng(14))
    pr2 = PeerRoots(ctx, pp.num_rows)
    _, begin, count, comm = sharded_commit(pp, poly, ctx, peer=pr2)
    _, full = MultilinearZip.commit(pp, poly, ctx)
    ok = ok and comm.roots == full.roots and (begin, count) == shard_range(pp.num_rows, rank, world)
    dist.barrier()
    pr2.close()
    q.put((rank, ok))
    dist.destroy_process_group()


def test_peer_roots_two_gpus():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs of one node")
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_peer_worker, args=(r, 2, 29731, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    res = sorted(q.get(timeout=5) for _ in range(2))
    assert res == [(0, True), (1, True)], res


@pytest.mark.parametrize("chunked", [False, True, "fused"])
@pytest.mark.parametrize("row_len,num_rows,limbs", [(16384, 37, 1), (16384, 301, 1), (32768, 5, 1), (65536, 3, 1), (16384, 6, 2)])
def test_codewords_longer_than_shared_memory(row_len, num_rows, limbs, chunked, oracle, ctx, monkeypatch):
    """cw = 32768 / 65536 / 131072 (nv = 27 ... 34 row shapes; code_raa.rs:42-43 and structs.rs:79-90 put no bound on nv):
    the encoders of raa_big.cu (raw permutations, intermediate vectors in global scratch) -- the row-per-CTA form (Int<1>:
    one persistent CTA per SM, rows in segments of 16384 positions, stores through per-warp tiles for 3-limb scans; 301
    rows = more than one row per CTA) and the three chunked launches (forced, and what Int<2> inputs use) -- codewords,
    every Merkle layer and the roots of a row sample against the oracle; cw = 131072 needs 98-bit entries (4-limb scans),
    Int<2> inputs 5-limb scans"""
    if chunked == "fused":  # the commit form of the row-per-CTA encoder (leaves + tree levels 1..4 in the same launch;
        monkeypatch.setenv("ZIPGPU_FUSE_MIN_ROWS", "1")  # taken from 888 rows by default), where the shape has one
    elif chunked:
        monkeypatch.setenv("ZIPGPU_BIG_CHUNKED", "1")
    from zinc_b200 import DenseMultilinearExtension, MultilinearZip, MultilinearZipParams, RaaCode, RandomFieldZipTypes

    cw = 2 * row_len
    p1, p2 = oracle.perm_from_seed(cw, KECCAK_SEEDS[0]), oracle.perm_from_seed(cw, KECCAK_SEEDS[1])
    zt = RandomFieldZipTypes(limbs)
    code = RaaCode.with_permutations(zt, row_len, 2, p1, p2)
    evals = np.random.default_rng(row_len + limbs).integers(0, 1 << 64, size=num_rows * row_len * limbs, dtype=np.uint64)
    rc, rows, layers, roots = oracle.commit(evals, num_rows, row_len, 2, p1, p2, in_limbs=limbs, out_limbs=4 * limbs)
    assert rc == 0
    nv = (num_rows * row_len - 1).bit_length()
    pp = MultilinearZipParams.new(nv, num_rows, code)
    poly = DenseMultilinearExtension(evals.reshape(-1, limbs), nv)
    data, comm = MultilinearZip.commit(pp, poly, ctx)
    assert np.array_equal(data.rows.reshape(-1), rows)
    assert np.array_equal(np.concatenate([t.layers.reshape(-1) for t in data.rows_merkle_trees]), layers)
    assert b"".join(comm.roots) == roots.tobytes()


def test_encode_rows_long_non_power_of_two_codeword(oracle, ctx):
    """encode_rows / commit_no_merkle only need the code, not a power-of-two codeword: cw = 20000 (3 chunks, the last
    one partial) through the chunked encoder"""
    from zinc_b200 import MultilinearZip, MultilinearZipParams, RaaCode, ZipTypes

    row_len, num_rows = 10000, 9
    cw = 2 * row_len
    p1, p2 = oracle.perm_from_seed(cw, 1), oracle.perm_from_seed(cw, 2)
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2)
    evals = np.random.default_rng(3).integers(0, 1 << 64, size=num_rows * row_len, dtype=np.uint64)
    rc, rows = oracle.encode_rows(evals, num_rows, row_len, 2, p1, p2)
    assert rc == 0
    pp = MultilinearZipParams.new(17, num_rows, code)
    got = MultilinearZip.encode_rows(pp, cw, row_len, evals, ctx)
    assert np.array_equal(got.reshape(-1), rows)


@pytest.mark.parametrize("num_rows", [1, 97, 148, 500, 1211])
def test_cw16384_warp_specialised_commit_kernel(num_rows, oracle, ctx, monkeypatch):
    """cw = 16384 (nv = 25 / 26) through commit_ws16k_kernel -- one plane set, the ENC group encodes in two half-passes
    and stores the codeword from registers, the HASH group hashes it back from global memory -- for row counts below,
    at and above one row per SM and above the dynamic-claiming threshold; against the oracle and the other two paths"""
    from zinc_b200 import RaaCode, ZipTypes

    row_len, cw = 8192, 16384
    p1, p2 = oracle.perm_from_seed(cw, KECCAK_SEEDS[0]), oracle.perm_from_seed(cw, KECCAK_SEEDS[1])
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2)
    evals = np.random.default_rng(num_rows).integers(0, 1 << 64, size=num_rows * row_len, dtype=np.uint64)
    if num_rows >= 2:
        evals[:row_len] = np.uint64((1 << 63) - 1)          # i64::MAX row
        evals[row_len:2 * row_len] = np.uint64(1 << 63)      # i64::MIN row
    rc, rows, layers, roots = oracle.commit_mt(evals, num_rows, row_len, 2, 0, 0, p1, p2, threads=16, faithful=False)
    assert rc == 0
    monkeypatch.setenv("ZIPGPU_FUSE_MIN_ROWS", "1")
    monkeypatch.setenv("ZIPGPU_WS16K_MIN_ROWS", "1")  # the product uses this kernel from 2048 rows
    for knob in (None, "ZIPGPU_NO_WS", "ZIPGPU_NO_FUSE"):
        if knob:
            monkeypatch.setenv(knob, "1")
        g_rows, g_lay, g_roots = _commit_device(ctx, code, num_rows, cw, evals)
        assert np.array_equal(g_roots, roots), knob
        assert np.array_equal(g_rows, rows), knob
        assert np.array_equal(g_lay, layers), knob
