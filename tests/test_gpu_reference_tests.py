"""GPU: the reference's own commit-path tests (src/zip/pcs/commit.rs:240-775, src/zip/code_raa.rs:198-342)
re-expressed against the drop-in API, plus the edge cases of the boundary (other limb widths, depth 0, columns)."""
import numpy as np
import pytest

from helpers import I64_MAX, KECCAK_SEEDS, shape_for

pytestmark = pytest.mark.gpu


def to_int(limbs_row):
    v = 0
    for i, w in enumerate(limbs_row):
        v |= int(w) << (64 * i)
    bits = 64 * len(limbs_row)
    return v - (1 << bits) if v >> (bits - 1) else v


def setup_test_params(num_vars, transcript=None):
    """commit.rs:220-238 with RaaCode (the code in scope) instead of ZipLinearCode"""
    from zinc_b200 import (DefaultLinearCodeSpec, DenseMultilinearExtension, MockTranscript, MultilinearZip, RaaCode)

    poly_size = 1 << num_vars
    code = RaaCode.new(DefaultLinearCodeSpec(), poly_size, transcript or MockTranscript())
    pp = MultilinearZip.setup(poly_size, code)
    poly = DenseMultilinearExtension.from_evaluations_vec(num_vars, np.arange(1, poly_size + 1, dtype=np.int64))
    return pp, poly


def test_commit_is_deterministic(ctx):
    from zinc_b200 import MultilinearZip

    pp, poly = setup_test_params(3)
    r1 = MultilinearZip.commit(pp, poly, ctx)
    r2 = MultilinearZip.commit(pp, poly, ctx)
    assert r1[1].roots == r2[1].roots


def test_different_polynomials_produce_different_commitments(ctx):
    from zinc_b200 import DenseMultilinearExtension, MultilinearZip

    pp, _ = setup_test_params(3)
    c1 = MultilinearZip.commit(pp, DenseMultilinearExtension.from_evaluations_vec(3, np.full(8, 1, np.int64)), ctx)[1]
    c2 = MultilinearZip.commit(pp, DenseMultilinearExtension.from_evaluations_vec(3, np.full(8, 2, np.int64)), ctx)[1]
    assert c1.roots != c2.roots


@pytest.mark.parametrize("nv,vals", [(4, [42] * 16), (2, [1, 2, 3, 4])])
def test_commit_succeeds_for_small_polynomials(nv, vals, ctx):
    from zinc_b200 import DenseMultilinearExtension, MultilinearZip

    pp, _ = setup_test_params(nv)
    data, comm = MultilinearZip.commit(pp, DenseMultilinearExtension.from_evaluations_vec(nv, np.array(vals, np.int64)), ctx)
    assert len(comm.roots) == pp.num_rows


def test_merkle_tree_depth_count_and_sizes(ctx):
    from zinc_b200 import MultilinearZip

    pp, poly = setup_test_params(3)
    data, comm = MultilinearZip.commit(pp, poly, ctx)
    cw = pp.linear_code.codeword_len()
    for tree in data.rows_merkle_trees:
        assert tree.depth == cw.bit_length() - 1
        assert tree.layers.shape == ((2 << tree.depth) - 2, 32)
    assert len(data.rows_merkle_trees) == pp.num_rows and data.rows.shape == (pp.num_rows * cw, 4)


def test_commit_no_merkle_produces_empty_trees(ctx):
    from zinc_b200 import MultilinearZip

    pp, poly = setup_test_params(3)
    data, comm = MultilinearZip.commit_no_merkle(pp, poly, ctx)
    assert data.rows.shape[0] == pp.num_rows * pp.linear_code.codeword_len()
    assert data.rows_merkle_trees == [] and comm.roots == []
    full, _ = MultilinearZip.commit(pp, poly, ctx)
    assert np.array_equal(full.rows, data.rows)


def test_encoded_rows_match_linear_code_definition(ctx):
    """commit.rs:356-380"""
    from zinc_b200 import MultilinearZip

    for nv in (3, 6, 10):
        pp, poly = setup_test_params(nv)
        lc = pp.linear_code
        enc = MultilinearZip.encode_rows(pp, lc.codeword_len(), lc.row_len(), poly.evaluations, ctx)
        assert enc.shape[0] == pp.num_rows * lc.codeword_len()
        assert np.count_nonzero(enc) > 0  # commit.rs:415-428
        for i in range(min(pp.num_rows, 4)):
            row = poly.evaluations[i * lc.row_len():(i + 1) * lc.row_len()]
            assert np.array_equal(enc[i * lc.codeword_len():(i + 1) * lc.codeword_len()], lc.encode_wide(row, ctx=ctx))


def test_corrupted_encoding_changes_merkle_root(ctx):
    """commit.rs:383-398"""
    from zinc_b200 import MerkleTree, MultilinearZip

    pp, poly = setup_test_params(3)
    data, comm = MultilinearZip.commit(pp, poly, ctx)
    cw = pp.linear_code.codeword_len()
    row0 = data.rows[:cw].copy()
    assert MerkleTree.new(data.rows_merkle_trees[0].depth, row0, 4, ctx).root == comm.roots[0]
    row0[0] = [999999, 0, 0, 0]
    assert MerkleTree.new(data.rows_merkle_trees[0].depth, row0, 4, ctx).root != comm.roots[0]


def test_batch_commit(ctx):
    """commit.rs:324-339, 400-413"""
    from zinc_b200 import DenseMultilinearExtension, MultilinearZip

    pp, poly = setup_test_params(3)
    polys = [DenseMultilinearExtension.from_evaluations_vec(3, np.arange(1, 9, dtype=np.int64)),
             DenseMultilinearExtension.from_evaluations_vec(3, np.arange(9, 17, dtype=np.int64))]
    outs = MultilinearZip.batch_commit(pp, polys, ctx)
    assert len(outs) == 2 and outs[0][1].roots != outs[1][1].roots
    single = MultilinearZip.commit(pp, poly, ctx)
    batch = MultilinearZip.batch_commit(pp, [poly], ctx)[0]
    assert batch[1].roots == single[1].roots and np.array_equal(batch[0].rows, single[0].rows)
    for (d, c), p in zip(outs, polys):
        s = MultilinearZip.commit(pp, p, ctx)
        assert c.roots == s[1].roots and np.array_equal(d.rows, s[0].rows)
        assert all(np.array_equal(a.layers, b.layers) for a, b in zip(d.rows_merkle_trees, s[0].rows_merkle_trees))


def test_zero_alternating_and_large_values(ctx):
    """commit.rs:472-493, 617-632"""
    from zinc_b200 import DenseMultilinearExtension, MultilinearZip

    pp, _ = setup_test_params(3)
    data, comm = MultilinearZip.commit(pp, DenseMultilinearExtension.from_evaluations_vec(3, np.zeros(8, np.int64)), ctx)
    assert len(comm.roots) == pp.num_rows and not data.rows.any()
    alt = np.where(np.arange(8) % 2 == 0, 1, -1).astype(np.int64)
    assert len(MultilinearZip.commit(pp, DenseMultilinearExtension.from_evaluations_vec(3, alt), ctx)[1].roots) == pp.num_rows
    mx = DenseMultilinearExtension.from_evaluations_vec(3, np.full(8, I64_MAX, np.int64))
    lc = pp.linear_code
    enc = MultilinearZip.encode_rows(pp, lc.codeword_len(), lc.row_len(), mx.evaluations, ctx)
    assert enc.shape[0] == pp.num_rows * lc.codeword_len()
    # the last entry of a row is the sum of all prefix sums: strictly positive and above 2^63
    assert to_int(enc[lc.codeword_len() - 1]) > (1 << 63)


def test_encode_rows_succeeds_for_single_row(ctx):
    """commit.rs:504-518"""
    from zinc_b200 import DefaultLinearCodeSpec, MockTranscript, MultilinearZip, MultilinearZipParams, RaaCode

    code = RaaCode.new(DefaultLinearCodeSpec(), 4, MockTranscript())
    pp = MultilinearZipParams.new(2, 1, code)
    enc = MultilinearZip.encode_rows(pp, code.codeword_len(), code.row_len(), np.full(4, 5, np.int64), ctx)
    assert enc.shape[0] == code.codeword_len()


def test_merkle_root_integrity_is_maintained(ctx):
    """commit.rs:520-535"""
    from zinc_b200 import DenseMultilinearExtension, MerkleTree, MultilinearZip

    pp, _ = setup_test_params(3)
    data, comm = MultilinearZip.commit(pp, DenseMultilinearExtension.from_evaluations_vec(3, np.full(8, 42, np.int64)), ctx)
    cw = pp.linear_code.codeword_len()
    for i, tree in enumerate(data.rows_merkle_trees):
        ind = MerkleTree.new(tree.depth, data.rows[i * cw:(i + 1) * cw], 4, ctx)
        assert tree.root == ind.root == comm.roots[i]


def test_matrix_dimensions_and_many_variables(ctx):
    """commit.rs:537-548, 594-615"""
    from zinc_b200 import MultilinearZip

    for nv, rows in ((2, 2), (4, 4), (6, 8), (16, 256)):
        pp, poly = setup_test_params(nv)
        assert pp.num_rows == rows == 1 << (nv // 2)
        data, comm = MultilinearZip.commit(pp, poly, ctx)
        assert len(comm.roots) == pp.num_rows == len(data.rows_merkle_trees)


def test_linear_code_preserves_linearity(ctx):
    """commit.rs:558-583 (3*r1 + 5*r2) and code_raa.rs:278-315 (encode::<N, M>, zero -> zero)"""
    from zinc_b200 import DefaultLinearCodeSpec, MockTranscript, MultilinearZip, RaaCode

    pp, poly = setup_test_params(4)
    lc = pp.linear_code
    rl, cw = lc.row_len(), lc.codeword_len()
    enc = MultilinearZip.encode_rows(pp, cw, rl, poly.evaluations, ctx)
    ev = poly.evaluations.reshape(-1).astype(np.int64)
    comb = 3 * ev[:rl] + 5 * ev[rl:2 * rl]
    comb_enc = lc.encode_wide(comb, ctx=ctx)
    for i in range(cw):
        assert to_int(comb_enc[i]) == 3 * to_int(enc[i]) + 5 * to_int(enc[cw + i])
    code = RaaCode.new(DefaultLinearCodeSpec(), 16, MockTranscript())
    a, b = np.arange(1, 5, dtype=np.int64), np.arange(5, 9, dtype=np.int64)
    ea, eb, es = code.encode(a, ctx), code.encode(b, ctx), code.encode(a + b, ctx)
    assert ea.shape == (8, 8)  # M = Int<8>
    assert [to_int(x) for x in es] == [to_int(x) + to_int(y) for x, y in zip(ea, eb)]
    assert not code.encode(np.zeros(4, np.int64), ctx).any()
    with pytest.raises(AssertionError, match="Row length must match the code's row length"):  # code_raa.rs:334-342
        code.encode(np.array([1, 2, 3], np.int64), ctx)


# ---- boundary edge cases beyond the reference's tests ----------------------------------------------------
@pytest.mark.parametrize("limbs", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("depth", [0, 1, 2, 3, 4, 7, 11])
def test_merkle_rows_all_leaf_widths(limbs, depth, oracle, ctx):
    import ctypes as C

    from zinc_b200 import _native as nat

    rng = np.random.default_rng(100 * limbs + depth)
    num_rows = 3
    leaves = rng.integers(0, 1 << 64, size=(num_rows << depth) * limbs, dtype=np.uint64)
    per = (2 << depth) - 2
    layers = np.zeros(num_rows * per * 32, dtype=np.uint8)
    roots = np.zeros(num_rows * 32, dtype=np.uint8)
    nat.check(nat.lib().zipgpu_merkle_rows(ctx.handle, num_rows, depth, limbs, nat.ptr(leaves),
                                           nat.ptr(layers) if per else None, nat.ptr(roots)))
    for r in range(num_rows):
        rc, l, root = oracle.merkle_tree(depth, leaves[(r << depth) * limbs:((r + 1) << depth) * limbs], limbs)
        assert rc == 0 and np.array_equal(root, roots[32 * r:32 * r + 32])
        assert np.array_equal(l, layers[r * per * 32:(r + 1) * per * 32])
    roots2 = np.zeros_like(roots)  # layers_out = NULL
    nat.check(nat.lib().zipgpu_merkle_rows(ctx.handle, num_rows, depth, limbs, nat.ptr(leaves), None, nat.ptr(roots2)))
    assert np.array_equal(roots, roots2)


@pytest.mark.parametrize("nv", [4, 8, 12])
def test_two_limb_inputs(nv, oracle, ctx):
    """INT_LIMBS = 2: N = Int<2> -> K = Int<8> (64-byte leaves, still one BLAKE3 block) -- SURVEY.md 8f-4"""
    from zinc_b200 import (DenseMultilinearExtension, MultilinearZip, MultilinearZipParams, RaaCode,
                           RandomFieldZipTypes)

    row_len, num_rows, cw = shape_for(nv)
    p1, p2 = oracle.perm_from_seed(cw, KECCAK_SEEDS[0]), oracle.perm_from_seed(cw, KECCAK_SEEDS[1])
    code = RaaCode.with_permutations(RandomFieldZipTypes(2), row_len, 2, p1, p2)
    pp = MultilinearZipParams.new(nv, num_rows, code)
    poly = DenseMultilinearExtension.rand(nv, np.random.default_rng(nv), limbs=2)
    data, comm = MultilinearZip.commit(pp, poly, ctx)
    rc, rows, layers, roots = oracle.commit(poly.evaluations.reshape(-1), num_rows, row_len, 2, p1, p2, in_limbs=2, out_limbs=8)
    assert rc == 0 and np.array_equal(data.rows.reshape(-1), rows)
    assert np.array_equal(np.concatenate([t.layers.reshape(-1) for t in data.rows_merkle_trees]), layers)
    assert b"".join(comm.roots) == roots.tobytes()


def test_resident_data_and_column_openings(oracle, ctx):
    """zipgpu_commit_resident + open_z.rs:124-143 / pcs/utils.rs:163-176 column openings from resident data"""
    import ctypes as C

    from zinc_b200 import DenseMultilinearExtension, MultilinearZip, MultilinearZipParams, RaaCode, ZipTypes

    nv = 10
    row_len, num_rows, cw = shape_for(nv)
    depth = cw.bit_length() - 1
    p1, p2 = oracle.perm_from_seed(cw, 1), oracle.perm_from_seed(cw, 2)
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2)
    pp = MultilinearZipParams.new(nv, num_rows, code)
    poly = DenseMultilinearExtension.rand(nv, np.random.default_rng(3))
    res, comm = MultilinearZip.commit_resident(pp, poly, ctx)
    rc, rows, layers, roots = oracle.commit(poly.evaluations.reshape(-1), num_rows, row_len, 2, p1, p2)
    assert b"".join(comm.roots) == roots.tobytes()
    assert np.array_equal(res.rows().reshape(-1), rows)
    assert np.array_equal(res.layers().reshape(-1), layers)
    assert np.array_equal(res.rows(5, 2).reshape(-1), rows[5 * cw * 4:7 * cw * 4])
    cols = np.array([0, 1, cw - 1, 17, 17, cw // 2], dtype=np.uint32)
    vals, paths = res.open_columns(cols)
    L = oracle.lib()
    per = (2 << depth) - 2
    path = np.zeros(depth * 32, dtype=np.uint8)
    p8 = C.POINTER(C.c_uint8)
    for ci, col in enumerate(cols):
        for r in range(num_rows):
            assert np.array_equal(vals[ci, r], rows[(r * cw + col) * 4:(r * cw + col) * 4 + 4])
            lay = layers[r * per * 32:(r + 1) * per * 32]
            L.zo_merkle_create_proof(depth, lay.ctypes.data_as(p8), int(col), path.ctypes.data_as(p8))
            assert np.array_equal(paths[ci, r].reshape(-1), path)
            if r % 7 == 0:
                root = roots[32 * r:32 * r + 32].copy()
                leaf = vals[ci, r].copy()
                assert L.zo_merkle_verify(depth, path.ctypes.data_as(p8), root.ctypes.data_as(p8),
                                          leaf.ctypes.data_as(C.POINTER(C.c_uint64)), 4, int(col)) == 0
    res.free()


def test_code_create_rejects_bad_arguments(ctx):
    import ctypes as C

    from zinc_b200 import _native as nat

    L = nat.lib()
    h = C.c_void_p()
    ident = np.arange(8, dtype=np.uint32)
    dup = ident.copy(); dup[3] = 2
    assert L.zipgpu_code_create(ctx.handle, 4, 2, 1, 4, nat.ptr(ident), nat.ptr(dup), C.byref(h)) == nat.ERR_INVALID
    assert b"permutation" in L.zipgpu_last_error()
    assert L.zipgpu_code_create(ctx.handle, 4, 2, 1, 1, nat.ptr(ident), nat.ptr(ident), C.byref(h)) == nat.ERR_WIDTH
    assert b"Cannot fit 70-bit wide codeword entries in 64 bits integers" in L.zipgpu_last_error()
    assert L.zipgpu_code_create(ctx.handle, 0, 2, 1, 4, nat.ptr(ident), nat.ptr(ident), C.byref(h)) == nat.ERR_INVALID
    # non power-of-two codeword (rep = 3): encode works, commit refuses like MerkleTree::new (pcs/utils.rs:75)
    p = np.arange(12, dtype=np.uint32)[::-1].copy()
    assert L.zipgpu_code_create(ctx.handle, 4, 3, 1, 4, nat.ptr(p), nat.ptr(p), C.byref(h)) == 0
    ev = np.arange(8, dtype=np.uint64)
    out = np.zeros(2 * 12 * 4, dtype=np.uint64)
    roots = np.zeros(64, dtype=np.uint8)
    assert L.zipgpu_encode_rows(h, 2, nat.ptr(ev), nat.ptr(out)) == 0
    assert L.zipgpu_commit(h, 2, nat.ptr(ev), None, None, nat.ptr(roots)) == nat.ERR_INVALID
    assert b"is_power_of_two" in L.zipgpu_last_error()
    L.zipgpu_code_destroy(h)


def test_rep3_encode_matches_oracle(oracle, ctx):
    """repetition factors other than 2 (LinearCodeSpec::repetition_factor, code.rs:217-226)"""
    from zinc_b200 import RaaCode, ZipTypes

    rng = np.random.default_rng(8)
    for row_len, rep in ((4, 3), (16, 4), (10, 3), (64, 8)):
        cw = row_len * rep
        p1, p2 = oracle.perm_from_seed(cw, 11), oracle.perm_from_seed(cw, 12)
        code = RaaCode.with_permutations(ZipTypes(), row_len, rep, p1, p2)
        row = rng.integers(0, 1 << 64, size=row_len, dtype=np.uint64)
        rc, exp = oracle.encode_rows(row, 1, row_len, rep, p1, p2)
        assert rc == 0 and np.array_equal(code.encode_wide(row, ctx=ctx).reshape(-1), exp)
