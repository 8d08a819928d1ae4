"""CPU: pins the oracle (C restatement + Python restatement) to every vector the domain offers:
the reference's own unit-test examples, the real `blake3` crate (PyPI binding), Keccak KATs and the golden
fixtures of tests/golden/.  No GPU needed."""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import KECCAK_SEEDS, MOCK_SEEDS

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def u64(vals, limbs=1):
    from oracle import pyoracle as po

    return np.array([w for v in vals for w in po.to_words(v, limbs)], dtype=np.uint64)


# ---- BLAKE3 ----------------------------------------------------------------------------------------------
def test_blake3_official_empty_vector(oracle):
    assert oracle.blake3(b"").hex() == "af1349b9f5f9a1a6a0404dea36dcc9499bcb25c9adc112b7cc9a93cae41f3262"


def test_blake3_matches_crate_binding(oracle):
    blake3 = pytest.importorskip("blake3")
    rng = np.random.default_rng(1)
    for n in [0, 1, 31, 32, 33, 63, 64, 65, 127, 128, 129, 1023, 1024, 1025, 2047, 2048, 2049, 3073, 4096, 8193]:
        data = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert oracle.blake3(data) == blake3.blake3(data).digest(), n


def test_blake3_portable_equals_simd(oracle):
    rng = np.random.default_rng(2)
    datas = [rng.integers(0, 256, size=n, dtype=np.uint8).tobytes() for n in (0, 32, 64, 100, 1500)]
    simd = [oracle.blake3(d) for d in datas]
    oracle.lib().zo_blake3_force_portable(1)
    try:
        assert [oracle.blake3(d) for d in datas] == simd
    finally:
        oracle.lib().zo_blake3_force_portable(0)


def test_hash_kats_from_golden(oracle):
    from oracle import pyoracle as po

    kats = json.load(open(os.path.join(GOLD, "hash_kats.json")))
    for k in kats["leaf"]:
        v, limbs = int(k["value"]), k["limbs"]
        assert po.int_to_bytes(v, limbs).hex() == k["bytes"]
        assert oracle.blake3(bytes.fromhex(k["bytes"])).hex() == k["digest"]
        rc, layers, root = oracle.merkle_tree(0, u64([v], limbs), limbs)  # depth 0: root == leaf hash
        assert rc == 0 and root.tobytes().hex() == k["digest"]
    for k in kats["node"]:
        assert oracle.blake3(bytes.fromhex(k["left"]) + bytes.fromhex(k["right"])).hex() == k["digest"]
    for k in kats["blake3"]:
        data = bytes((i * 7 + 3) % 251 for i in range(k["len"]))
        assert oracle.blake3(data).hex() == k["digest"]


def test_survey_leaf_kats(oracle):
    """SURVEY.md 8c known answers (produced with the blake3 crate binding)"""
    from oracle import pyoracle as po

    kat = {0: "2ada83c1819a5372dae1238fc1ded123c8104fdaa15862aaee69428a1820fcda",
           1: "19506519c965ba8cf8168f6eee5c80f0bf4fab8df1619f99b16f02bbef2b000e",
           -1: "9b34f060fbc0f0aa11f150e26519deff613277b60656f0f8356ed2261505f5c5"}
    for v, d in kat.items():
        assert oracle.blake3(po.int_to_bytes(v, 4)).hex() == d
    assert po.int_to_bytes(1, 4) == bytes(7) + b"\x01" + bytes(24)  # int.rs:201-210: limb 0 first, big-endian
    z = bytes.fromhex(kat[0])
    assert oracle.blake3(z + z).hex() == "969c821452711f40e3d4023d9a09c3ae62ba1192d8884526c03d53688a498055"


# ---- Int<N>, repeat, accumulate (reference unit-test examples) --------------------------------------------
def test_repeat_examples(oracle):
    """code_raa.rs:199-221"""
    import ctypes as C

    L = oracle.lib()
    inp = u64([10, 20])
    for rep, exp in ((3, [10, 20, 10, 20, 10, 20]), (1, [10, 20])):
        out = np.zeros(2 * rep, dtype=np.uint64)
        L.zo_repeat(inp.ctypes.data_as(C.POINTER(C.c_uint64)), 2, 1, rep, out.ctypes.data_as(C.POINTER(C.c_uint64)), 1)
        assert out.tolist() == exp


def test_accumulate_examples(oracle):
    """code_raa.rs:224-244 incl. negatives"""
    import ctypes as C

    from oracle import pyoracle as po

    L = oracle.lib()
    for inp, exp in (([1, 2, 3, 4], [1, 3, 6, 10]), ([5, 0, 2, 0], [5, 5, 7, 7]), ([-1, 5, -10, 2], [-1, 4, -6, -4])):
        for limbs in (1, 4):
            a = u64(inp, limbs)
            assert L.zo_accumulate(a.ctypes.data_as(C.POINTER(C.c_uint64)), 4, limbs) == 0
            got = [po.to_signed(a[i * limbs:(i + 1) * limbs].tolist()) for i in range(4)]
            assert got == exp
        assert po.accumulate(inp) == exp


def test_widening_sign_extends(oracle):
    """zip/utils.rs:163-234"""
    import ctypes as C

    L = oracle.lib()
    p64 = C.POINTER(C.c_uint64)
    a = np.array([1, 2], dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    L.zo_widen(a.ctypes.data_as(p64), 2, out.ctypes.data_as(p64), 4)
    assert out.tolist() == [1, 2, 0, 0]  # test_expand_normal
    a = np.array([123], dtype=np.uint64)
    out3 = np.zeros(3, dtype=np.uint64)
    L.zo_widen(a.ctypes.data_as(p64), 1, out3.ctypes.data_as(p64), 3)
    assert out3.tolist() == [123, 0, 0]  # test_expand_zero_padding
    m1 = np.array([2**64 - 1, 2**64 - 1], dtype=np.uint64)
    L.zo_widen(m1.ctypes.data_as(p64), 2, out.ctypes.data_as(p64), 4)
    assert out.tolist() == [2**64 - 1] * 4  # negative numbers sign-extend


def test_accumulate_overflow_is_reported(oracle):
    import ctypes as C

    a = u64([(1 << 63) - 1, 1])
    assert oracle.lib().zo_accumulate(a.ctypes.data_as(C.POINTER(C.c_uint64)), 2, 1) == 1


# ---- shape rules -----------------------------------------------------------------------------------------
def test_shape_rules(oracle):
    """code_raa.rs:42-43, structs.rs:79-90, commit.rs:538-548 (matrix_dimensions_are_invariant)"""
    from oracle import pyoracle as po

    L = oracle.lib()
    from helpers import shape_for

    for nv in range(1, 31):
        row_len, num_rows, _ = shape_for(nv)
        # closed form 2^ceil(nv/2) x 2^floor(nv/2), except nv = 1 and 3 where isqrt(2^nv) is itself a power of two
        if nv not in (1, 3):
            assert (row_len, num_rows) == (1 << ((nv + 1) // 2), 1 << (nv // 2))
        assert L.zo_raa_row_len(1 << nv) == row_len == po.raa_row_len(1 << nv), nv
        assert L.zo_num_rows(1 << nv, row_len) == num_rows == po.num_rows_for(1 << nv, row_len)
    assert shape_for(3) == (2, 4, 4) and shape_for(1) == (1, 2, 2)
    # code_raa.rs:318-332: Int<1> -> Int<1> at 2^30 needs 96 bits
    assert po.raa_width_bits(1, 1 << 30, 2) == 96
    assert L.zo_raa_width_ok(1, 1, 1 << 30, 2) == 0 and L.zo_raa_width_ok(1, 4, 1 << 30, 2) == 1


# ---- Keccak transcript -> seeds --------------------------------------------------------------------------
def test_keccak_vs_hashlib_and_kats():
    from oracle import pyoracle as po
    from zinc_b200.transcript import keccak256

    rng = np.random.default_rng(3)
    for n in [0, 1, 135, 136, 137, 271, 272, 1000]:
        d = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert po.keccak256(d, 0x06) == hashlib.sha3_256(d).digest()  # same permutation/padding rule, SHA-3 suffix
        assert po.keccak256(d) == keccak256(d)
    assert po.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"


def test_transcript_reference_kat():
    """transcript.rs:213-234: get_challenge after absorbing "This is a test string!" (the reference's own KAT)"""
    from oracle import pyoracle as po

    h = po.keccak256(b"This is a test string!")
    lo, hi = int.from_bytes(h[:16], "big"), int.from_bytes(h[16:], "big")
    modulus = 3618502788666131213697322783095070105623107215331596699973092056135872020481
    nbits = modulus.bit_length() - 1
    val = (lo + (hi & ((1 << (nbits - 128)) - 1)) * (1 << 128)) % modulus
    assert val == 693058076479703886486101269644733982722902192016595549603371045888466087870


def test_fresh_keccak_transcript_seeds():
    from oracle import pyoracle as po
    from zinc_b200.transcript import KeccakTranscript, MockTranscript

    for T in (po.KeccakTranscript, KeccakTranscript):
        t = T()
        assert (t.get_u64(), t.get_u64()) == KECCAK_SEEDS
    m = MockTranscript()
    assert (m.get_u64(), m.get_u64()) == MOCK_SEEDS


# ---- rand 0.9.2 restatement (unpinned against the crate; internal consistency only) -----------------------
def test_chacha_core_against_openssl():
    crypto = pytest.importorskip("cryptography.hazmat.primitives.ciphers")
    from oracle import pyoracle as po

    key = bytes(range(32))
    kw = [int(x) for x in np.frombuffer(key, dtype="<u4")]
    ks = crypto.Cipher(crypto.algorithms.ChaCha20(key, bytes(16)), mode=None).encryptor().update(bytes(192))
    blocks = sum((po.chacha_block(kw, c, 0, 20) for c in range(3)), [])
    assert np.array(blocks, dtype="<u4").tobytes() == ks


CHACHA12_TC1 = ("9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f"
                "0564f879d27ae3c02ce82834acfa8c793a629f2ca0de6919610be82f411326be")
CHACHA20_TC1 = ("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
                "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586")


def test_chacha12_published_vector(oracle):
    """StdRng = ChaCha12: the published ChaCha12 test vector (zero 256-bit key, zero nonce, block 0), and the ChaCha20
    one for the same core, for ALL three restatements: oracle C, pyoracle, and the product's csrc/rand_compat.cpp"""
    import ctypes as C

    from oracle import pyoracle as po
    from zinc_b200 import _native as nat

    zero = np.zeros(8, dtype=np.uint32)
    for rounds, expect in ((12, CHACHA12_TC1), (20, CHACHA20_TC1)):
        out = np.empty(16, dtype=np.uint32)
        oracle.lib().zo_chacha_block(zero.ctypes.data_as(C.POINTER(C.c_uint32)), 0, 0, rounds,
                                     out.ctypes.data_as(C.POINTER(C.c_uint32)))
        assert out.astype("<u4").tobytes().hex() == expect, f"oracle C, {rounds} rounds"
        assert np.array(po.chacha_block([0] * 8, 0, 0, rounds), dtype="<u4").tobytes().hex() == expect
        out2 = np.empty(16, dtype=np.uint32)
        nat.check(nat.lib().zipgpu_chacha_block(nat.ptr(zero), 0, rounds, nat.ptr(out2)))
        assert out2.astype("<u4").tobytes().hex() == expect, f"rand_compat.cpp, {rounds} rounds"
    # a non-trivial key and counter: the three agree with OpenSSL's ChaCha20 (covers key/counter word placement)
    crypto = pytest.importorskip("cryptography.hazmat.primitives.ciphers")
    key = bytes(range(32))
    kw = np.frombuffer(key, dtype="<u4").astype(np.uint32)
    ks = crypto.Cipher(crypto.algorithms.ChaCha20(key, (5).to_bytes(8, "little") + bytes(8)), mode=None).encryptor().update(bytes(64))
    out3 = np.empty(16, dtype=np.uint32)
    nat.check(nat.lib().zipgpu_chacha_block(nat.ptr(kw), 5, 20, nat.ptr(out3)))
    assert out3.astype("<u4").tobytes() == ks


def test_rand_python_equals_c(oracle):
    from oracle import pyoracle as po

    for seed in (0, 1, 2, 12345, 54321) + KECCAK_SEEDS:
        r = po.StdRng(seed)
        assert [r.next_u32() for _ in range(130)] == oracle.stdrng_words(seed, 130).tolist()
        for n in (1, 2, 3, 10, 13, 14, 100, 1000):
            assert po.perm_from_seed(n, seed) == oracle.perm_from_seed(n, seed).tolist()


def test_rand_snapshot(oracle):
    snap = json.load(open(os.path.join(GOLD, "rand_snapshot.json")))
    for s, perm in snap["perm16"].items():
        assert oracle.perm_from_seed(16, int(s)).tolist() == perm


def test_shuffle_properties(oracle):
    """code_raa.rs:247-276: deterministic per seed, different across seeds, a true permutation"""
    a, b, c = oracle.perm_from_seed(10, 12345), oracle.perm_from_seed(10, 12345), oracle.perm_from_seed(10, 54321)
    assert a.tolist() == b.tolist() and a.tolist() != c.tolist()
    assert a.tolist() != list(range(10)) and c.tolist() != list(range(10))
    for n in (2, 16, 512, 8192):
        assert sorted(oracle.perm_from_seed(n, 7).tolist()) == list(range(n))


def test_seeded_encode_equals_gather_form(oracle):
    """the literal repeat/shuffle/accumulate chain (code_raa.rs:98-102) == the permutation-array form"""
    rng = np.random.default_rng(5)
    for row_len in (2, 4, 16, 64):
        row = rng.integers(0, 1 << 64, size=row_len, dtype=np.uint64)
        for s1, s2 in (MOCK_SEEDS, KECCAK_SEEDS):
            rc, lit = oracle.encode_row_seeded(row, 2, s1, s2)
            p1, p2 = oracle.perm_from_seed(2 * row_len, s1), oracle.perm_from_seed(2 * row_len, s2)
            rc2, gat = oracle.encode_rows(row, 1, row_len, 2, p1, p2)
            assert rc == 0 and rc2 == 0 and np.array_equal(lit, gat)


# ---- whole-path: C oracle vs Python oracle vs golden ------------------------------------------------------
def test_c_oracle_matches_golden_vectors(oracle):
    blake3 = pytest.importorskip("blake3")
    vecs = json.load(open(os.path.join(GOLD, "commit_vectors.json")))
    assert len(vecs) >= 60
    for fx in vecs:
        nv, cw, nr, rl = fx["nv"], fx["cw"], fx["num_rows"], fx["row_len"]
        p1, p2 = oracle.perm_from_seed(cw, fx["seeds"][0]), oracle.perm_from_seed(cw, fx["seeds"][1])
        assert blake3.blake3(p1.tobytes()).hexdigest() == fx["perm1_blake3"]
        if "evals" in fx:
            evals = u64([int(v) for v in fx["evals"]])
            assert p1.tolist() == fx["perm1"] and p2.tolist() == fx["perm2"]
        else:
            from golden.make_golden import pattern

            evals = u64(pattern(fx["pattern"], 1 << nv, nv))
        rc, rows, layers, roots = oracle.commit(evals, nr, rl, 2, p1, p2)
        assert rc == 0
        assert blake3.blake3(rows.tobytes()).hexdigest() == fx["rows_blake3"], (nv, fx["pattern"])
        assert blake3.blake3(layers.tobytes()).hexdigest() == fx["layers_blake3"], (nv, fx["pattern"])
        assert [roots[32 * i:32 * i + 32].tobytes().hex() for i in range(nr)] == fx["roots"]
        if "rows" in fx:
            assert rows.tobytes().hex() == "".join(fx["rows"])
            assert layers.tobytes().hex() == "".join("".join(l) for l in fx["layers"])


def test_commit_mt_equals_single_thread(oracle):
    rng = np.random.default_rng(9)
    nv, rl, nr, cw = 10, 32, 32, 64
    evals = rng.integers(0, 1 << 64, size=1 << nv, dtype=np.uint64)
    s1, s2 = KECCAK_SEEDS
    p1, p2 = oracle.perm_from_seed(cw, s1), oracle.perm_from_seed(cw, s2)
    rc, rows, layers, roots = oracle.commit(evals, nr, rl, 2, p1, p2)
    for threads in (1, 3, 8):
        for faithful in (False, True):
            rc2, r2, l2, ro2 = oracle.commit_mt(evals, nr, rl, 2, s1, s2, p1, p2, threads, faithful)
            assert rc2 == 0 and np.array_equal(rows, r2) and np.array_equal(layers, l2) and np.array_equal(roots, ro2)


def test_merkle_proofs_verify(oracle):
    """pcs/utils.rs:340-363"""
    import ctypes as C

    rng = np.random.default_rng(11)
    leaves = rng.integers(0, 1 << 64, size=1024 * 3, dtype=np.uint64)
    rc, layers, root = oracle.merkle_tree(10, leaves, 3)
    assert rc == 0
    L = oracle.lib()
    p8, p64 = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
    path = np.zeros(10 * 32, dtype=np.uint8)
    for i in range(0, 1024, 17):
        L.zo_merkle_create_proof(10, layers.ctypes.data_as(p8), i, path.ctypes.data_as(p8))
        leaf = leaves[3 * i:3 * i + 3].copy()
        assert L.zo_merkle_verify(10, path.ctypes.data_as(p8), root.ctypes.data_as(p8), leaf.ctypes.data_as(p64), 3, i) == 0
        bad = leaf.copy(); bad[0] ^= 1
        assert L.zo_merkle_verify(10, path.ctypes.data_as(p8), root.ctypes.data_as(p8), bad.ctypes.data_as(p64), 3, i) != 0


def test_merkle_rejects_non_power_of_two(oracle):
    """commit.rs:634-640"""
    rc, _, _ = oracle.merkle_tree(3, np.arange(7, dtype=np.uint64), 1)
    assert rc != 0


def test_linearity_and_zero(oracle):
    """code_raa.rs:279-315, commit.rs:559-583"""
    from oracle import pyoracle as po

    p1, p2 = po.perm_from_seed(8, 1), po.perm_from_seed(8, 2)
    a, b = [1, 2, 3, 4], [5, 6, 7, 8]
    ea, eb = po.encode_row_perm(a, 2, p1, p2), po.encode_row_perm(b, 2, p1, p2)
    assert po.encode_row_perm([3 * x + 5 * y for x, y in zip(a, b)], 2, p1, p2) == [3 * x + 5 * y for x, y in zip(ea, eb)]
    assert po.encode_row_perm([0] * 4, 2, p1, p2) == [0] * 8


# ---- ZipLinearCode, the sparse code (zip/code.rs:77-215) --------------------------------------------------
def _random_sparse(rng, n, m, d, lo, hi):
    cols = np.stack([np.sort(rng.choice(m, size=d, replace=False)) for _ in range(n)]).astype(np.uint32).reshape(-1)
    coef = rng.integers(lo, hi, size=n * d, endpoint=True).astype(np.int64)
    return cols, coef


@pytest.mark.parametrize("in_limbs,out_limbs", [(1, 4), (2, 8), (1, 2)])
def test_sparse_python_equals_c(oracle, in_limbs, out_limbs):
    from oracle import pyoracle as po

    rng = np.random.default_rng(11 + in_limbs)
    row_len, n, d, num_rows = 16, 16, 8, 3
    a = _random_sparse(rng, n, row_len, d, -(1 << 40), 1 << 40)
    b = _random_sparse(rng, n, row_len, d, -3, 3)
    evals = [int(x) for x in rng.integers(-(1 << 63), (1 << 63) - 1, size=num_rows * row_len)]
    if in_limbs == 2:
        evals = [v * ((1 << 62) + 12345) for v in evals]
    if out_limbs == 2:
        a = (a[0], (a[1] % 7).astype(np.int64))  # keep the sums inside 128 bits
    rc, rows, _, roots = oracle.sparse_commit(u64(evals, in_limbs), num_rows, row_len, n, d, a[0], a[1], b[0], b[1],
                                              in_limbs, out_limbs, want_layers=False, threads=2)
    assert rc == 0
    want = []
    for r in range(num_rows):
        want += po.sparse_encode_row(evals[r * row_len:(r + 1) * row_len], n, d,
                                     (a[0].tolist(), a[1].tolist()), (b[0].tolist(), b[1].tolist()))
    assert np.array_equal(rows, u64(want, out_limbs))


def test_sparse_golden_vectors(oracle):
    blake3 = pytest.importorskip("blake3")
    with open(os.path.join(GOLD, "sparse_vectors.json")) as f:
        fixtures = json.load(f)
    assert len(fixtures) == 9
    for fx in fixtures:
        evals = u64([int(v) for v in fx["evals"]])
        rc, rows, layers, roots = oracle.sparse_commit(evals, fx["num_rows"], fx["row_len"], fx["cw"] // 2, fx["cells_per_row"],
                                                       fx["cols_a"], fx["coef_a"], fx["cols_b"], fx["coef_b"])
        assert rc == 0
        assert blake3.blake3(rows.tobytes()).hexdigest() == fx["rows_blake3"], (fx["nv"], fx["transcript"])
        assert blake3.blake3(layers.tobytes()).hexdigest() == fx["layers_blake3"]
        assert roots.tobytes().hex() == "".join(fx["roots"])


def test_mock_transcript_sparse_matrix_shape():
    """pcs/tests.rs:24-56 by hand: every sample_unique_columns bumps the counter once and takes the first d columns;
    every encoding element is the next counter value."""
    from oracle import pyoracle as po

    row_len, cw, a, b = po.zip_linear_code_new(16, po.MockTranscript())
    assert (row_len, cw) == (4, 8)
    assert a[0] == [0, 1] * 4 and a[1] == [2, 3, 5, 6, 8, 9, 11, 12]
    assert b[0] == [0, 1] * 4 and b[1] == [14, 15, 17, 18, 20, 21, 23, 24]


def test_keccak_encoding_elements_repeat_within_a_row():
    """transcript.rs:176-181: get_encoding_element does not advance the sponge, so all cells of a matrix row carry
    the same bit -- restated faithfully, not `fixed`."""
    from oracle import pyoracle as po

    _, _, a, b = po.zip_linear_code_new(1 << 8, po.KeccakTranscript())
    d = 8
    for coef in (a[1], b[1]):
        for i in range(0, len(coef), d):
            assert len(set(coef[i:i + d])) == 1
        assert set(coef) == {0, 1}
