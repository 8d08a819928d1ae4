"""GPU parity proper: the CUDA path (through the C ABI / host mirror) against the CPU oracle, bit for bit."""
import numpy as np
import pytest

from helpers import KECCAK_SEEDS, MOCK_SEEDS, input_patterns, shape_for

pytestmark = pytest.mark.gpu


def make_code(nv, seeds, oracle):
    from zinc_b200 import RaaCode, ZipTypes

    row_len, num_rows, cw = shape_for(nv)
    p1, p2 = oracle.perm_from_seed(cw, seeds[0]), oracle.perm_from_seed(cw, seeds[1])
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2)
    return code, row_len, num_rows, cw, p1, p2


@pytest.mark.parametrize("nv", [2, 3, 4, 5, 6, 8, 10, 11, 12, 14, 16])
@pytest.mark.parametrize("seeds", [MOCK_SEEDS, KECCAK_SEEDS], ids=["mock", "keccak"])
def test_commit_matches_oracle(nv, seeds, oracle, ctx):
    from zinc_b200 import DenseMultilinearExtension, MultilinearZip, MultilinearZipParams

    code, row_len, num_rows, cw, p1, p2 = make_code(nv, seeds, oracle)
    pp = MultilinearZipParams.new(nv, num_rows, code)
    for name, evals in input_patterns(nv):
        poly = DenseMultilinearExtension.from_evaluations_vec(nv, evals)
        data, comm = MultilinearZip.commit(pp, poly, ctx)
        rc, rows, layers, roots = oracle.commit(evals.view(np.uint64), num_rows, row_len, 2, p1, p2)
        assert rc == 0, name
        assert np.array_equal(data.rows.reshape(-1), rows), f"codewords differ: nv={nv} {name}"
        got_layers = np.concatenate([t.layers.reshape(-1) for t in data.rows_merkle_trees])
        assert np.array_equal(got_layers, layers), f"layers differ: nv={nv} {name}"
        assert b"".join(comm.roots) == roots.tobytes(), f"roots differ: nv={nv} {name}"
        if name == "random" and nv >= 10:
            break  # the edge patterns are covered at the small sizes; keep the big ones quick


@pytest.mark.parametrize("nv", [18, 20])
def test_commit_large_matches_oracle(nv, oracle, ctx):
    from zinc_b200 import DenseMultilinearExtension, MultilinearZip, MultilinearZipParams

    code, row_len, num_rows, cw, p1, p2 = make_code(nv, KECCAK_SEEDS, oracle)
    pp = MultilinearZipParams.new(nv, num_rows, code)
    name, evals = next(input_patterns(nv))
    poly = DenseMultilinearExtension.from_evaluations_vec(nv, evals)
    data, comm = MultilinearZip.commit(pp, poly, ctx)
    rc, rows, layers, roots = oracle.commit_mt(evals.view(np.uint64), num_rows, row_len, 2, 0, 0, p1, p2,
                                               threads=8, faithful=False)
    assert rc == 0
    assert np.array_equal(data.rows.reshape(-1), rows)
    assert np.array_equal(np.concatenate([t.layers.reshape(-1) for t in data.rows_merkle_trees]), layers)
    assert b"".join(comm.roots) == roots.tobytes()


def test_merkle_tree_int3_proofs(oracle, ctx):
    """pcs/utils.rs:340-363: 1024 random Int<3> leaves; every proof verifies (checked with the oracle's verify)"""
    import ctypes as C

    from zinc_b200 import MerkleProof, MerkleTree

    rng = np.random.default_rng(7)
    leaves = rng.integers(0, 1 << 64, size=(1024, 3), dtype=np.uint64)
    tree = MerkleTree.new(10, leaves, 3, ctx)
    rc, layers, root = oracle.merkle_tree(10, leaves.reshape(-1), 3)
    assert rc == 0 and tree.root == root.tobytes()
    assert np.array_equal(tree.layers.reshape(-1), layers)
    L = oracle.lib()
    for i in range(0, 1024, 37):
        proof = MerkleProof.create_proof(tree, i)
        path = np.frombuffer(b"".join(proof.merkle_path), dtype=np.uint8).copy()
        rootb = np.frombuffer(tree.root, dtype=np.uint8).copy()
        leaf = leaves[i].copy()
        assert L.zo_merkle_verify(10, path.ctypes.data_as(C.POINTER(C.c_uint8)),
                                  rootb.ctypes.data_as(C.POINTER(C.c_uint8)),
                                  leaf.ctypes.data_as(C.POINTER(C.c_uint64)), 3, i) == 0


@pytest.mark.parametrize("num_rows,depth,limbs", [(1, 16, 4), (1, 12, 4), (37, 9, 4), (256, 10, 4), (3, 13, 8), (5, 1, 4),
                                                   (2, 11, 3), (64, 7, 2)])
def test_cta_tree_latency_path_matches_oracle_and_pass_path(num_rows, depth, limbs, oracle, ctx, monkeypatch):
    """small whole trees (<= 2^18 leaves) take the CTA-per-subtree kernel; same layers and roots as the oracle
    (pcs/utils.rs:74-118, the MerkleRoot shapes of zip_benches.rs:80-98) and as the subtree-pass kernels"""
    from zinc_b200 import _native as nat

    rng = np.random.default_rng(depth * 100 + num_rows)
    leaves = rng.integers(0, 1 << 64, size=num_rows * (1 << depth) * limbs, dtype=np.uint64)
    per_row = ((2 << depth) - 2) * 32

    def run():
        layers = np.zeros(num_rows * per_row, dtype=np.uint8)
        roots = np.zeros(num_rows * 32, dtype=np.uint8)
        n0 = ctx.launch_count
        nat.check(nat.lib().zipgpu_merkle_rows(ctx.handle, num_rows, depth, limbs, nat.ptr(leaves), nat.ptr(layers), nat.ptr(roots)))
        return layers, roots, ctx.launch_count - n0

    lay_c, roots_c, n_cta = run()
    monkeypatch.setenv("ZIPGPU_NO_CTA_TREE", "1")
    lay_p, roots_p, n_pass = run()
    assert n_cta == -(-depth // 10) and n_cta <= n_pass
    assert np.array_equal(lay_c, lay_p) and np.array_equal(roots_c, roots_p)
    for r in range(num_rows):
        rc, olay, oroot = oracle.merkle_tree(depth, leaves[r * (1 << depth) * limbs:(r + 1) * (1 << depth) * limbs], limbs)
        assert rc == 0
        assert np.array_equal(lay_c[r * per_row:(r + 1) * per_row], olay), r
        assert np.array_equal(roots_c[r * 32:(r + 1) * 32], oroot), r
