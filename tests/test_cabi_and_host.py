"""CPU: the C-ABI library loads and exports every symbol include/zipgpu.h declares; host logic of the mirror API
(geometry, validation, error behaviour) works without a GPU; the product never touches oracle/."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "zipgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zipgpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from zinc_b200 import _native as nat

    L = C.CDLL(nat.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 38
    for n in names:
        assert hasattr(L, n), f"libzipgpu.so does not export {n}"
    assert set(names) == set(nat.SIGNATURES), set(names) ^ set(nat.SIGNATURES)
    assert b"sm_100a" in nat.lib().zipgpu_version()


def test_no_cpu_fallback_without_gpu():
    """on a box without a GPU the product must fail loudly (ZIPGPU_ERR_NO_DEVICE), never compute on the CPU"""
    from zinc_b200 import _native as nat

    cnt = C.c_int(-1)
    rc = nat.lib().zipgpu_device_count(C.byref(cnt))
    if rc == 0 and cnt.value > 0:
        pytest.skip("a GPU is visible here")
    h = C.c_void_p()
    assert nat.lib().zipgpu_ctx_create(0, C.byref(h)) == nat.ERR_NO_DEVICE
    assert b"no CPU fallback" in nat.lib().zipgpu_last_error()
    from zinc_b200 import Context

    with pytest.raises(nat.ZipGpuError):
        Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "zinc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                bad = re.search(r"(from\s+oracle|import\s+oracle|zip_oracle|libzip_oracle|oracle/|oracle\.)", txt)
                assert not bad, f"{f} references the checker ({bad.group(0)}): the product must not route through it"


def test_geometry_matches_reference_rules():
    from helpers import shape_for
    from zinc_b200 import _native as nat

    L = nat.lib()
    for nv in range(1, 31):
        row_len, num_rows, _ = shape_for(nv)
        assert L.zipgpu_raa_row_len(1 << nv) == row_len
        assert L.zipgpu_num_rows(1 << nv, row_len) == num_rows
    assert L.zipgpu_raa_codeword_width_bits(1, 1 << 30, 2) == 96   # code_raa.rs:318-332
    assert L.zipgpu_raa_codeword_width_bits(1, 1 << 24, 2) == 90
    assert L.zipgpu_raa_codeword_width_bits(1, 1 << 25, 2) == 92   # num_vars_even
    assert L.zipgpu_raa_codeword_width_bits(2, 1 << 16, 4) == 128 + 16 + 4


def test_product_perm_from_seed_equals_oracle(oracle):
    from zinc_b200 import shuffle_seeded_indices

    for seed in (1, 2, 12345, 0xB9736F582676E7E8, 0xD7397E6260CE9C3E):
        for n in (1, 2, 3, 4, 12, 13, 14, 100, 512, 8192, 16384):
            assert np.array_equal(shuffle_seeded_indices(n, seed), oracle.perm_from_seed(n, seed)), (n, seed)


def test_raa_code_new_mirrors_reference():
    """code_raa.rs:35-86 + zip_benches.rs:100-106"""
    from zinc_b200 import DefaultLinearCodeSpec, KeccakTranscript, MockTranscript, MultilinearZip, RaaCode

    code = RaaCode.new(DefaultLinearCodeSpec(), 1 << 16, KeccakTranscript())
    assert (code.row_len(), code.codeword_len()) == (256, 512)
    assert (code.perm_1_seed, code.perm_2_seed) == (0xB9736F582676E7E8, 0xD7397E6260CE9C3E)
    assert code.num_column_opening() == 1000 and code.num_proximity_testing() == 1
    pp = MultilinearZip.setup(1 << 16, code)
    assert (pp.num_vars, pp.num_rows) == (16, 256)
    code = RaaCode.new(DefaultLinearCodeSpec(), 16, MockTranscript())
    assert (code.perm_1_seed, code.perm_2_seed, code.row_len()) == (1, 2, 4)
    p1, p2 = code.permutations()
    assert sorted(p1.tolist()) == list(range(8)) and sorted(p2.tolist()) == list(range(8))
    with pytest.raises(AssertionError, match="is_power_of_two"):
        MultilinearZip.setup(12, code)


def test_constructor_panics_on_insufficient_codeword_width():
    """code_raa.rs:317-332"""
    from zinc_b200 import DefaultLinearCodeSpec, MockTranscript, RaaCode, ZipTypes

    with pytest.raises(AssertionError, match="Cannot fit 96-bit wide codeword entries in 64 bits integers"):
        RaaCode.new(DefaultLinearCodeSpec(), 1 << 30, MockTranscript(), ZipTypes(N=1, L=2, K=1, M=4))


def test_commit_validation_happens_before_the_gpu():
    """commit.rs:54-63: Err for too many variables, panic for a wrong evaluation count -- both host side"""
    from zinc_b200 import (DefaultLinearCodeSpec, DenseMultilinearExtension, InvalidPcsParam, MockTranscript,
                           MultilinearZip, MultilinearZipParams, RaaCode)

    code = RaaCode.new(DefaultLinearCodeSpec(), 8, MockTranscript())
    pp = MultilinearZip.setup(8, code)
    poly4 = DenseMultilinearExtension.from_evaluations_vec(4, np.arange(1, 17, dtype=np.int64))
    with pytest.raises(InvalidPcsParam, match="Too many variates of poly to commit"):  # commit.rs:241-250
        MultilinearZip.commit(pp, poly4)
    with pytest.raises(InvalidPcsParam):
        MultilinearZip.batch_commit(pp, [poly4])
    poly3 = DenseMultilinearExtension.from_evaluations_vec(3, np.arange(1, 9, dtype=np.int64))
    bad_pp = MultilinearZipParams.new(3, 3, code)  # commit.rs:550-556 reject_incompatible_dimensions
    with pytest.raises(AssertionError, match="incorrect number of evaluations"):
        MultilinearZip.commit(bad_pp, poly3)
    poly3.evaluations = poly3.evaluations[:7]  # commit.rs:585-592
    with pytest.raises(AssertionError, match="incorrect number of evaluations"):
        MultilinearZip.commit(pp, poly3)
    assert MultilinearZip.batch_commit(pp, []) == []  # commit.rs:495-502


def test_dense_mle_semantics():
    """poly_z/mle/dense.rs:43-64"""
    from zinc_b200 import DenseMultilinearExtension

    p = DenseMultilinearExtension.from_evaluations_vec(3, np.array([1, -2, 3], dtype=np.int64))
    assert p.evaluations.shape == (8, 1) and p.evaluations[3:].sum() == 0
    assert p.evaluations[1, 0] == np.uint64(2**64 - 2)
    with pytest.raises(AssertionError, match="should not exceed"):
        DenseMultilinearExtension.from_evaluations_vec(2, np.arange(5, dtype=np.int64))
    q = DenseMultilinearExtension.from_evaluations_vec(1, np.array([-1, 5], dtype=np.int64), limbs=2)
    assert q.evaluations.tolist() == [[2**64 - 1, 2**64 - 1], [5, 0]]  # sign-extended limbs


def test_merkle_tree_new_panics_on_non_power_of_two_leaves():
    """commit.rs:634-640"""
    from zinc_b200 import MerkleTree

    with pytest.raises(AssertionError, match=r"leaves.len\(\).is_power_of_two\(\)"):
        MerkleTree.new(3, np.arange(7, dtype=np.int64))


def test_shard_range_partitions():
    from zinc_b200.dist import shard_range

    for n in (1, 2, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_product_transcript_equals_oracle_transcript():
    """the host mirror's incremental Keccak sponge against the oracle's one-shot Keccak (pinned to hashlib there)"""
    from oracle import pyoracle as po
    from zinc_b200 import KeccakTranscript
    from zinc_b200.transcript import _Sponge, keccak256

    for n in (0, 1, 135, 136, 137, 271, 272, 273, 1000):
        data = bytes((7 * i + 1) % 256 for i in range(n))
        sp = _Sponge()
        sp.update(data[:n // 3])
        sp.update(data[n // 3:])
        assert sp.digest() == keccak256(data) == po.keccak256(data), n
    a, b = KeccakTranscript(), po.KeccakTranscript()
    a.absorb(b"zinc" * 50)
    b.absorb(b"zinc" * 50)
    for _ in range(40):
        assert a.get_usize_in_range(3, 67) == b.get_usize_in_range(3, 67)
        assert a.get_encoding_element() == b.get_encoding_element()
    assert a.get_u64() == b.get_u64()


def test_zip_linear_code_new_mirrors_reference():
    """ZipLinearCode::new (code.rs:100-147) over both transcripts: shapes and sampled cells equal the oracle's"""
    from oracle import pyoracle as po
    from zinc_b200 import DefaultLinearCodeSpec, KeccakTranscript, MockTranscript, ZipLinearCode

    for poly_size, T, OT in ((16, MockTranscript, po.MockTranscript), (1 << 7, KeccakTranscript, po.KeccakTranscript)):
        code = ZipLinearCode.new(DefaultLinearCodeSpec(), poly_size, T())
        row_len, cw, a, b = po.zip_linear_code_new(poly_size, OT())
        assert (code.row_len(), code.codeword_len()) == (row_len, cw)
        assert code.num_column_opening() == 1000 and code.num_proximity_testing() == 1
        assert code.a.cols.tolist() == a[0] and code.a.coef.tolist() == a[1]
        assert code.b.cols.tolist() == b[0] and code.b.coef.tolist() == b[1]
        assert code.a.d == row_len // 2 and code.a.n == cw // 2
    with pytest.raises(AssertionError):
        ZipLinearCode.new(DefaultLinearCodeSpec(), 12, MockTranscript())  # code.rs:105 poly_size.is_power_of_two()


def test_multi_gpu_context_has_no_cpu_fallback_either():
    """zipgpu_mgpu_create (ONE process, n GPUs) refuses to exist without a device, like zipgpu_ctx_create"""
    from zinc_b200 import MultiContext
    from zinc_b200 import _native as nat

    cnt = C.c_int(-1)
    rc = nat.lib().zipgpu_device_count(C.byref(cnt))
    if rc == 0 and cnt.value > 0:
        pytest.skip("a GPU is visible here")
    h = C.c_void_p()
    assert nat.lib().zipgpu_mgpu_create(None, 0, C.byref(h)) == nat.ERR_NO_DEVICE
    with pytest.raises(nat.ZipGpuError):
        MultiContext()
    # the NULL-argument paths of the sharded entry points answer without touching a device
    assert nat.lib().zipgpu_commit_resident_sharded(None, None, 0, 0, None, None, None) == nat.ERR_INVALID
    assert nat.lib().zipgpu_peer_roots_status(None) == nat.ERR_INVALID
    assert nat.lib().zipgpu_encode_f(None, 0, 1, None, None, None) == nat.ERR_INVALID
    assert nat.lib().zipgpu_encode_wide(None, 0, 8, 8, None, None) == nat.ERR_INVALID


def test_as_limbs_accepts_signed_limb_arrays():
    """a 2-D int64 array with limbs > 1 (multi-limb Int<2> inputs given as signed words) used to recurse for
    ever; it is a reinterpretation of the two's-complement words, while 1-D int64 values are sign-extended"""
    from zinc_b200.zip import as_limbs

    a = np.array([[1, -1], [-(1 << 63), (1 << 63) - 1]], dtype=np.int64)
    got = as_limbs(a, 2)
    assert got.dtype == np.uint64 and got.shape == (2, 2)
    assert got.tolist() == [[1, 0xFFFFFFFFFFFFFFFF], [1 << 63, (1 << 63) - 1]]
    assert as_limbs(np.zeros((4, 2), np.int64), 2).shape == (4, 2)
    ext = as_limbs(np.array([-2, 3], dtype=np.int64), 3)  # values: sign-extended into the upper limbs
    assert ext.tolist() == [[0xFFFFFFFFFFFFFFFE, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF], [3, 0, 0]]
    assert as_limbs(np.array([5, 6], dtype=np.int32), 1).tolist() == [[5], [6]]


def test_bench_and_library_agree_on_the_row_partition():
    """bench.py, dist.py and mgpu.cu (balanced contiguous row ranges) must cut a commitment the same way"""
    import importlib.util

    from zinc_b200.dist import shard_range

    src = open(os.path.join(ROOT, "bench.py")).read()
    ns = {}
    start = src.index("def shard_range")
    end = src.index("def gen_evals")
    exec(src[start:end], ns)
    for n in (0, 1, 3, 512, 4096, 4099):
        for world in (1, 2, 3, 8):
            for r in range(world):
                assert ns["shard_range"](n, r, world) == shard_range(n, r, world)
