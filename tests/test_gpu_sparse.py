"""GPU parity for ZipLinearCode, the sparse code (zip/code.rs:77-215), through the C ABI / host mirror: the generic
kernel (any i64 coefficient, any shape) and the tensor-core kernel (0..255 coefficients, 128-multiples), against the
C oracle and the golden fixtures, bit for bit."""
import json
import os

import numpy as np
import pytest

from helpers import I64_MAX, I64_MIN

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def random_matrix(rng, n, m, d, lo, hi):
    from zinc_b200 import SparseMatrixZ

    cols = np.empty((n, d), dtype=np.uint32)
    for i in range(n):
        cols[i] = np.sort(rng.choice(m, size=d, replace=False))
    return SparseMatrixZ(n, m, d, cols, rng.integers(lo, hi, size=n * d, endpoint=True))


def check_against_oracle(code, evals_u64, num_rows, oracle, ctx, in_limbs=1, out_limbs=4, threads=8):
    import ctypes as C

    from zinc_b200 import _native as nat

    a, b = code.a, code.b
    cw = code.codeword_len()
    depth = cw.bit_length() - 1
    h = code.native(ctx, in_limbs, out_limbs)
    rows = np.empty(num_rows * cw * out_limbs, dtype=np.uint64)
    layers = np.empty(num_rows * ((2 << depth) - 2) * 32, dtype=np.uint8)
    roots = np.empty(num_rows * 32, dtype=np.uint8)
    nat.check(nat.lib().zipgpu_commit(h, num_rows, nat.ptr(evals_u64), nat.ptr(rows), nat.ptr(layers), nat.ptr(roots)))
    rc, orows, olayers, oroots = oracle.sparse_commit(evals_u64, num_rows, code.row_len(), a.n, a.d, a.cols, a.coef,
                                                      b.cols, b.coef, in_limbs, out_limbs, threads=threads)
    assert rc == 0
    assert np.array_equal(rows, orows), "codewords differ"
    assert np.array_equal(layers, olayers), "layers differ"
    assert np.array_equal(roots, oroots), "roots differ"
    return rows


def test_sparse_golden_fixtures(ctx):
    import blake3

    from zinc_b200 import (DenseMultilinearExtension, MultilinearZip, MultilinearZipParams, SparseMatrixZ, ZipLinearCode,
                           ZipTypes)

    with open(os.path.join(GOLD, "sparse_vectors.json")) as f:
        fixtures = json.load(f)
    for fx in fixtures:
        n, m, d = fx["cw"] // 2, fx["row_len"], fx["cells_per_row"]
        code = ZipLinearCode.with_matrices(ZipTypes(), m, fx["cw"], SparseMatrixZ(n, m, d, fx["cols_a"], fx["coef_a"]),
                                           SparseMatrixZ(n, m, d, fx["cols_b"], fx["coef_b"]))
        pp = MultilinearZipParams.new(fx["nv"], fx["num_rows"], code)
        poly = DenseMultilinearExtension.from_evaluations_vec(fx["nv"], [int(v) for v in fx["evals"]])
        data, comm = MultilinearZip.commit(pp, poly, ctx)
        assert blake3.blake3(data.rows.tobytes()).hexdigest() == fx["rows_blake3"], (fx["nv"], fx["transcript"])
        layers = b"".join(t.layers.tobytes() for t in data.rows_merkle_trees)
        assert blake3.blake3(layers).hexdigest() == fx["layers_blake3"]
        assert [r.hex() for r in comm.roots] == fx["roots"]


def test_reference_commit_tests_with_zip_linear_code(ctx):
    """commit.rs:218-300: setup_test_params over MockTranscript + ZipLinearCode; rejects too many variables, is
    deterministic, separates different polynomials, succeeds for 2 and 4 variables"""
    from zinc_b200 import (DefaultLinearCodeSpec, DenseMultilinearExtension, InvalidPcsParam, MockTranscript, MultilinearZip,
                           MultilinearZipParams, ZipLinearCode)

    def setup(num_vars):
        code = ZipLinearCode.new(DefaultLinearCodeSpec(), 1 << num_vars, MockTranscript())
        pp = MultilinearZipParams.new(num_vars, 1 << -(-num_vars // 2), code)
        return pp, DenseMultilinearExtension.from_evaluations_vec(num_vars, list(range(1, (1 << num_vars) + 1)))

    pp, _ = setup(3)
    with pytest.raises(InvalidPcsParam):
        MultilinearZip.commit(pp, DenseMultilinearExtension.from_evaluations_vec(4, list(range(1, 17))), ctx)
    pp, poly = setup(3)
    assert MultilinearZip.commit(pp, poly, ctx)[1].roots == MultilinearZip.commit(pp, poly, ctx)[1].roots
    c1 = MultilinearZip.commit(pp, DenseMultilinearExtension.from_evaluations_vec(3, [1] * 8), ctx)[1]
    c2 = MultilinearZip.commit(pp, DenseMultilinearExtension.from_evaluations_vec(3, [2] * 8), ctx)[1]
    assert c1.roots != c2.roots
    for nv, evals in ((4, [42] * 16), (2, [1, 2, 3, 4])):
        code = ZipLinearCode.new(DefaultLinearCodeSpec(), 1 << nv, MockTranscript())
        pp = MultilinearZipParams.new(nv, 1 << (nv // 2), code)
        data, comm = MultilinearZip.commit(pp, DenseMultilinearExtension.from_evaluations_vec(nv, evals), ctx)
        assert len(comm.roots) == 1 << (nv // 2)


@pytest.mark.parametrize("row_len,num_rows,lo,hi", [(8, 5, -(1 << 62), 1 << 62), (64, 33, -1000, 1000), (128, 16, -5, 5),
                                                     (256, 7, 0, 1)])
def test_generic_kernel_matches_oracle(row_len, num_rows, lo, hi, oracle, ctx, monkeypatch):
    from zinc_b200 import ZipLinearCode, ZipTypes

    monkeypatch.setenv("ZIPGPU_SPARSE_GENERIC", "1")
    rng = np.random.default_rng(row_len)
    cw, d = 2 * row_len, row_len // 2
    code = ZipLinearCode.with_matrices(ZipTypes(), row_len, cw, random_matrix(rng, cw // 2, row_len, d, lo, hi),
                                       random_matrix(rng, cw // 2, row_len, d, lo, hi))
    assert code.kernel_kind(ctx) == "generic"
    evals = rng.integers(I64_MIN, I64_MAX, size=num_rows * row_len, dtype=np.int64, endpoint=True).view(np.uint64)
    if lo < -(1 << 32):
        evals = (evals.view(np.int64) >> 8).view(np.uint64)  # keep |sum| below 2^255
    check_against_oracle(code, evals, num_rows, oracle, ctx)


@pytest.mark.parametrize("kernel", ["tcgen05", "mma_sync"])
@pytest.mark.parametrize("row_len,num_rows,cmax", [(128, 16, 1), (128, 37, 1), (256, 100, 1), (512, 64, 255), (1024, 19, 7)])
def test_tensor_kernel_matches_oracle(row_len, num_rows, cmax, kernel, oracle, ctx, monkeypatch):
    """both tensor-core kernels: tcgen05.mma.kind::i8 (default) and the mma.sync version of the same product"""
    from zinc_b200 import ZipLinearCode, ZipTypes

    if kernel == "mma_sync":
        monkeypatch.setenv("ZIPGPU_SPARSE_MMA_SYNC", "1")
    rng = np.random.default_rng(1000 + row_len + num_rows)
    cw, d = 2 * row_len, row_len // 2
    code = ZipLinearCode.with_matrices(ZipTypes(), row_len, cw, random_matrix(rng, cw // 2, row_len, d, 0, cmax),
                                       random_matrix(rng, cw // 2, row_len, d, 0, cmax))
    assert code.kernel_kind(ctx) == "tensor"
    evals = rng.integers(I64_MIN, I64_MAX, size=num_rows * row_len, dtype=np.int64, endpoint=True)
    evals[:row_len] = I64_MIN          # the extremes of the byte-plane bias
    evals[row_len:2 * row_len] = I64_MAX
    evals[2 * row_len:3 * row_len] = 0
    evals[3 * row_len:4 * row_len] = -1
    check_against_oracle(code, evals.view(np.uint64), num_rows, oracle, ctx)


def test_tensor_kernel_equals_generic_kernel_int2(oracle, ctx, monkeypatch):
    """INT_LIMBS = 2: Int<2> evaluations, Int<8> codewords (16 byte planes per entry)"""
    from zinc_b200 import RandomFieldZipTypes, ZipLinearCode

    rng = np.random.default_rng(5)
    row_len, num_rows = 256, 21
    cw, d = 2 * row_len, row_len // 2
    zt = RandomFieldZipTypes(2)
    mats = random_matrix(rng, cw // 2, row_len, d, 0, 1), random_matrix(rng, cw // 2, row_len, d, 0, 1)
    code = ZipLinearCode.with_matrices(zt, row_len, cw, *mats)
    assert code.kernel_kind(ctx, 2, 8) == "tensor"
    evals = rng.integers(0, 1 << 64, size=num_rows * row_len * 2, dtype=np.uint64)
    evals[:2 * row_len:2] = 0
    evals[1:2 * row_len:2] = 1 << 63  # Int<2>::MIN
    got = check_against_oracle(code, evals, num_rows, oracle, ctx, in_limbs=2, out_limbs=8)
    monkeypatch.setenv("ZIPGPU_SPARSE_GENERIC", "1")
    code2 = ZipLinearCode.with_matrices(zt, row_len, cw, *mats)
    assert code2.kernel_kind(ctx, 2, 8) == "generic"
    assert np.array_equal(got, check_against_oracle(code2, evals, num_rows, oracle, ctx, in_limbs=2, out_limbs=8))


def test_keccak_sampled_code_nv12(oracle, ctx):
    """ZipLinearCode::new over a fresh KeccakTranscript (what a prover builds), 2^12 evaluations"""
    from zinc_b200 import DefaultLinearCodeSpec, KeccakTranscript, ZipLinearCode

    code = ZipLinearCode.new(DefaultLinearCodeSpec(), 1 << 12, KeccakTranscript())
    assert (code.row_len(), code.codeword_len()) == (64, 128)
    rng = np.random.default_rng(12)
    evals = rng.integers(I64_MIN, I64_MAX, size=1 << 12, dtype=np.int64, endpoint=True).view(np.uint64)
    check_against_oracle(code, evals, 64, oracle, ctx)


def test_sparse_nv20_full_size(oracle, ctx):
    """2^20 evaluations: 1024 rows x 1024, cw 2048, 512 cells per matrix row; resident commit + column openings"""
    from zinc_b200 import DenseMultilinearExtension, MultilinearZip, MultilinearZipParams, ZipLinearCode, ZipTypes

    rng = np.random.default_rng(20)
    row_len, num_rows, cw = 1024, 1024, 2048
    code = ZipLinearCode.with_matrices(ZipTypes(), row_len, cw, random_matrix(rng, 1024, row_len, 512, 0, 1),
                                       random_matrix(rng, 1024, row_len, 512, 0, 1))
    evals = rng.integers(I64_MIN, I64_MAX, size=1 << 20, dtype=np.int64, endpoint=True)
    rows = check_against_oracle(code, evals.view(np.uint64), num_rows, oracle, ctx)
    pp = MultilinearZipParams.new(20, num_rows, code)
    data, comm = MultilinearZip.commit_resident(pp, DenseMultilinearExtension.from_evaluations_vec(20, evals), ctx)
    vals, paths = data.open_columns([0, 1, 777, 2047])
    want = rows.reshape(num_rows, cw, 4)
    assert np.array_equal(vals[2], want[:, 777, :])
    data.free()


def test_sparse_linearity_nv22(ctx):
    """size-independent property at 2^22 evaluations (2048 x 2048 -> cw 4096): encode(a) + encode(b) == encode(a + b)"""
    from zinc_b200 import DenseMultilinearExtension, MultilinearZip, MultilinearZipParams, ZipLinearCode, ZipTypes

    rng = np.random.default_rng(22)
    row_len, num_rows, cw = 2048, 2048, 4096
    code = ZipLinearCode.with_matrices(ZipTypes(), row_len, cw, random_matrix(rng, 2048, row_len, 1024, 0, 1),
                                       random_matrix(rng, 2048, row_len, 1024, 0, 1))
    assert code.kernel_kind(ctx) == "tensor"
    pp = MultilinearZipParams.new(22, num_rows, code)
    a = rng.integers(-(1 << 62), 1 << 62, size=1 << 22, dtype=np.int64)
    b = rng.integers(-(1 << 62), 1 << 62, size=1 << 22, dtype=np.int64)
    enc = lambda v: MultilinearZip.commit_no_merkle(pp, DenseMultilinearExtension.from_evaluations_vec(22, v), ctx)[0].rows
    ra, rb, rs = enc(a), enc(b), enc(a + b)
    # Int<4> adds: the sums stay far below 2^128, so two u64 limbs with carry decide equality
    lo = ra[..., 0] + rb[..., 0]
    carry = (lo < ra[..., 0]).astype(np.uint64)
    assert np.array_equal(lo, rs[..., 0])
    assert np.array_equal(ra[..., 1] + rb[..., 1] + carry, rs[..., 1])


@pytest.mark.parametrize("num_rows", [1024, 1500, 4096])
def test_device_commit_overlapped_equals_serial(num_rows, ctx, monkeypatch):
    """zipgpu_commit_device on large matrices (>= 2^26 codeword entries, forced here) encodes row chunks on the
    high-priority stream while the previous chunks are hashed; the result (rows, every layer, roots) must equal the serial schedule's, which the tests above pin to the
    oracle"""
    import torch

    from zinc_b200 import ZipLinearCode, ZipTypes, _native as nat

    rng = np.random.default_rng(num_rows)
    row_len, cw, depth = 512, 1024, 10
    code = ZipLinearCode.with_matrices(ZipTypes(), row_len, cw, random_matrix(rng, 512, row_len, 256, 0, 1),
                                       random_matrix(rng, 512, row_len, 256, 0, 1))
    h = code.native(ctx, 1, 4)
    dev = torch.device("cuda:0")
    evals = torch.from_numpy(rng.integers(I64_MIN, I64_MAX, size=num_rows * row_len, dtype=np.int64, endpoint=True)).to(dev)

    def run():
        rows = torch.zeros(num_rows * cw * 4, dtype=torch.int64, device=dev)
        layers = torch.zeros(num_rows * ((2 << depth) - 2) * 32, dtype=torch.uint8, device=dev)
        roots = torch.zeros(num_rows * 32, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        nat.check(nat.lib().zipgpu_commit_device(h, num_rows, evals.data_ptr(), rows.data_ptr(), layers.data_ptr(),
                                                 roots.data_ptr(), None))
        ctx.sync()
        return rows.cpu().numpy(), layers.cpu().numpy(), roots.cpu().numpy()

    monkeypatch.setenv("ZIPGPU_SPARSE_OVERLAP", "1")
    launches0 = ctx.launch_count
    overlapped = run()
    assert ctx.launch_count - launches0 >= 8 * 3  # 8 row chunks: plane split + GEMM + >= 1 tree pass each
    monkeypatch.setenv("ZIPGPU_SPARSE_OVERLAP", "0")
    serial = run()
    for a, b, name in zip(overlapped, serial, ("rows", "layers", "roots")):
        assert np.array_equal(a, b), name
    # and the serial device path against the host API (oracle-checked above) on the first rows
    nr = 64
    rows_h = np.empty(nr * cw * 4, dtype=np.uint64)
    roots_h = np.empty(nr * 32, dtype=np.uint8)
    ev_h = evals[: nr * row_len].cpu().numpy().view(np.uint64)
    nat.check(nat.lib().zipgpu_commit(h, nr, nat.ptr(ev_h), nat.ptr(rows_h), None, nat.ptr(roots_h)))
    assert np.array_equal(rows_h, serial[0][: nr * cw * 4].view(np.uint64))
    assert np.array_equal(roots_h, serial[2][: nr * 32])


def test_sparse_code_create_rejects_bad_input(ctx):
    """error behaviour of zipgpu_sparse_code_create: negative status + message, never a crash"""
    import ctypes as C

    from zinc_b200 import _native as nat

    L = nat.lib()
    cols = np.array([0, 1, 0, 1], dtype=np.uint32)
    coef = np.ones(4, dtype=np.int64)

    def create(row_len, cw, d, in_limbs, out_limbs, ca=cols, fa=coef):
        h = C.c_void_p()
        rc = L.zipgpu_sparse_code_create(ctx.handle, row_len, cw, d, in_limbs, out_limbs, nat.ptr(ca), nat.ptr(fa),
                                         nat.ptr(ca), nat.ptr(fa), C.byref(h))
        if rc == 0:
            L.zipgpu_code_destroy(h)
        return rc, L.zipgpu_last_error().decode()

    assert create(2, 4, 2, 1, 4)[0] == 0
    rc, msg = create(2, 5, 2, 1, 4)
    assert rc < 0 and "even" in msg
    rc, msg = create(2, 4, 3, 1, 4)
    assert rc < 0 and "cells_per_row" in msg
    rc, msg = create(2, 4, 2, 2, 1)
    assert rc < 0 and "in_limbs" in msg
    bad = np.array([0, 2, 0, 1], dtype=np.uint32)  # column 2 of a 2-column matrix
    rc, msg = create(2, 4, 2, 1, 4, ca=bad)
    assert rc < 0 and "out of range" in msg
