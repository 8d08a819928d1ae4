"""f-3: RaaCode::encode_f (the code over field elements, code_raa.rs:133-138) against the Python big-int oracle, and the
PcsTranscript wire format (pcs_transcript.rs:68-128,146-211) of what the commit / open path produces."""
import numpy as np
import pytest

from helpers import KECCAK_SEEDS, MOCK_SEEDS

# moduli: the 251-bit prime of the reference's transcript KAT (transcript.rs:213-234), a 64-bit and a 128-bit prime, and a
# 192-bit prime WITHOUT a spare top bit (2^192 - 237), which exercises the carry branch of reduce_modulus
P251 = 3618502788666131213697322783095070105623107215331596699973092056135872020481
MODULI = [(P251, 4), (18446744069414584321, 1), ((1 << 127) - 1, 2), ((1 << 192) - 237, 3)]


def _to_limbs(vals, limbs):
    out = np.empty((len(vals), limbs), dtype=np.uint64)
    for i, v in enumerate(vals):
        for l in range(limbs):
            out[i, l] = (v >> (64 * l)) & 0xFFFFFFFFFFFFFFFF
    return out


def test_field_add_oracle_matches_modular_addition():
    """the reference's add (wrapping add, then one conditional subtraction) is (a + b) mod p for reduced inputs, with and
    without a spare bit in the modulus"""
    from oracle import pyoracle as po
    import random

    rnd = random.Random(7)
    for p, limbs in MODULI:
        for _ in range(200):
            a, b = rnd.randrange(p), rnd.randrange(p)
            assert po.field_add(a, b, p, limbs) == (a + b) % p
        assert po.field_add(p - 1, p - 1, p, limbs) == (2 * p - 2) % p
        assert po.field_add(0, 0, p, limbs) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("row_len", [4, 128, 1000, 4096])
@pytest.mark.parametrize("modulus,limbs", MODULI)
def test_encode_f_matches_oracle(row_len, modulus, limbs, oracle, ctx):
    import random

    from oracle import pyoracle as po
    from zinc_b200 import RaaCode, ZipTypes

    cw = 2 * row_len
    p1, p2 = oracle.perm_from_seed(cw, MOCK_SEEDS[0]), oracle.perm_from_seed(cw, MOCK_SEEDS[1])
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2)
    rnd = random.Random(row_len * 31 + limbs)
    row = [rnd.randrange(modulus) for _ in range(row_len)]
    row[0], row[-1] = modulus - 1, modulus - 1  # extremes
    expect = po.encode_f_row(row, 2, p1.tolist(), p2.tolist(), modulus, limbs)
    got = code.encode_f(_to_limbs(row, limbs), modulus, limbs, ctx)
    assert np.array_equal(got, _to_limbs(expect, limbs))


@pytest.mark.gpu
def test_encode_f_linearity_and_zero(oracle, ctx):
    """code_raa.rs:279-315 in the field setting: encode_f(a) + encode_f(b) == encode_f(a + b), zero -> zero"""
    import random

    from zinc_b200 import RaaCode, ZipTypes

    row_len, (p, limbs) = 256, MODULI[0]
    cw = 2 * row_len
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, oracle.perm_from_seed(cw, KECCAK_SEEDS[0]),
                                     oracle.perm_from_seed(cw, KECCAK_SEEDS[1]))
    rnd = random.Random(1)
    a = [rnd.randrange(p) for _ in range(row_len)]
    b = [rnd.randrange(p) for _ in range(row_len)]
    ea, eb = code.encode_f(_to_limbs(a, limbs), p, limbs, ctx), code.encode_f(_to_limbs(b, limbs), p, limbs, ctx)
    eab = code.encode_f(_to_limbs([(x + y) % p for x, y in zip(a, b)], limbs), p, limbs, ctx)

    def ints(arr):
        return [sum(int(arr[i, l]) << (64 * l) for l in range(limbs)) for i in range(arr.shape[0])]

    assert [(x + y) % p for x, y in zip(ints(ea), ints(eb))] == ints(eab)
    assert not code.encode_f(np.zeros((row_len, limbs), dtype=np.uint64), p, limbs, ctx).any()


@pytest.mark.gpu
def test_encode_f_rejects_unreduced_input(oracle, ctx):
    from zinc_b200 import RaaCode, ZipTypes
    from zinc_b200._native import ERR_INVALID, ZipGpuError

    code = RaaCode.with_permutations(ZipTypes(), 8, 2, oracle.perm_from_seed(16, 1), oracle.perm_from_seed(16, 2))
    row = np.zeros((8, 1), dtype=np.uint64)
    row[3, 0] = 97
    with pytest.raises(ZipGpuError) as ei:
        code.encode_f(row, 97, 1, ctx)
    assert ei.value.code == ERR_INVALID and "not reduced" in ei.value.message


def test_pcs_stream_round_trip():
    """test_read_write! of pcs_transcript.rs:213-280 for the integer / commitment / proof calls"""
    from zinc_b200 import PcsStream

    rng = np.random.default_rng(0)
    ints = rng.integers(0, 1 << 64, size=(5, 4), dtype=np.uint64)
    roots = rng.integers(0, 256, size=(3, 32), dtype=np.uint8)
    path = [bytes(rng.integers(0, 256, size=32, dtype=np.uint8)) for _ in range(7)]
    w = PcsStream()
    w.write_integers(ints)
    w.write_commitments(roots)
    w.write_merkle_proof(path)
    w.write_integer(ints[0])
    data = w.into_bytes()
    assert len(data) == 5 * 32 + 3 * 32 + 8 + 7 * 32 + 32
    # write_integer: little-endian u64 limbs, LSW first (pcs_transcript.rs:108-116)
    assert data[:8] == int(ints[0, 0]).to_bytes(8, "little")
    # write_merkle_proof: big-endian u64 length prefix (pcs_transcript.rs:198-211)
    assert data[5 * 32 + 3 * 32:5 * 32 + 3 * 32 + 8] == (7).to_bytes(8, "big")
    r = PcsStream(data)
    assert np.array_equal(r.read_integers(5, 4), ints)
    assert r.read_commitments(3) == [bytes(x) for x in roots]
    assert r.read_merkle_proof() == path
    assert np.array_equal(r.read_integer(4), ints[0])
    with pytest.raises(EOFError):
        r.read_commitment()


@pytest.mark.gpu
def test_gpu_outputs_are_wire_bytes(oracle, ctx):
    """the roots a commit returns are write_commitments(roots), the opened columns' wire bytes parse back with
    read_integers / read_merkle_proof into the structured openings, and a combined row is write_integers of its values"""
    from zinc_b200 import (DenseMultilinearExtension, MultilinearZip, MultilinearZipParams, PcsStream, RaaCode, ZipTypes)

    nv, row_len, num_rows = 10, 32, 32
    cw = 2 * row_len
    depth = 6
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, oracle.perm_from_seed(cw, 1), oracle.perm_from_seed(cw, 2))
    pp = MultilinearZipParams.new(nv, num_rows, code)
    poly = DenseMultilinearExtension.rand(nv, np.random.default_rng(3))
    data, comm = MultilinearZip.commit_resident(pp, poly, ctx)
    w = PcsStream()
    w.write_commitments(comm.roots)
    assert PcsStream(w.into_bytes()).read_commitments(num_rows) == comm.roots
    cols = np.array([5, 63, 0], dtype=np.uint32)
    vals, paths = data.open_columns(cols)
    r = PcsStream(data.open_columns_wire(cols))
    for ci in range(cols.size):
        assert np.array_equal(r.read_integers(num_rows, 4), vals[ci])
        for row in range(num_rows):
            assert r.read_merkle_proof() == [bytes(p) for p in paths[ci, row]]
    assert r.pos == len(r.stream)
    co = np.random.default_rng(1).integers(0, 1 << 64, size=num_rows, dtype=np.uint64)
    comb = data.combine_rows(co, 8)
    w2 = PcsStream()
    w2.write_integers(comb)
    assert w2.into_bytes() == comb.astype("<u8").tobytes() and len(w2.into_bytes()) == row_len * 64
    data.free()


def _signed_limbs(vals, limbs):
    return _to_limbs([v & ((1 << (64 * limbs)) - 1) for v in vals], limbs)


@pytest.mark.gpu
@pytest.mark.parametrize("row_len", [4, 128, 1000, 4096])
@pytest.mark.parametrize("in_limbs,out_limbs,bits", [(8, 8, 200), (4, 8, 250), (3, 3, 150), (1, 8, 63), (5, 7, 300), (8, 8, 480)])
def test_encode_wide_any_widths_matches_oracle(row_len, in_limbs, out_limbs, bits, oracle, ctx):
    """RaaCode::encode_wide::<In, Out> (code_raa.rs:125-131) at the widths the tiled prover encoder does not carry -- above
    all the verifier's encode_wide::<M, M> of a combined row (In = Out = Int<8>, verify_z.rs:74-78): signed entries of
    `bits` bits (sums of coeff x evaluation products are ~140 bits), sign extension on load, against the C oracle's
    repeat -> permute -> accumulate -> permute -> accumulate at the same widths, and the extremes"""
    import random

    from zinc_b200 import RaaCode, ZipTypes

    cw = 2 * row_len
    p1, p2 = oracle.perm_from_seed(cw, KECCAK_SEEDS[0]), oracle.perm_from_seed(cw, KECCAK_SEEDS[1])
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2)
    rnd = random.Random(row_len * 131 + in_limbs * 8 + out_limbs)
    bits = min(bits, 64 * in_limbs - 1)
    row = [rnd.randrange(-(1 << bits), 1 << bits) for _ in range(row_len)]
    row[0], row[-1] = (1 << bits) - 1, -(1 << bits)
    r = _signed_limbs(row, in_limbs)
    rc, expect = oracle.encode_rows(r.reshape(-1), 1, row_len, 2, p1, p2, in_limbs=in_limbs, out_limbs=out_limbs)
    got = code.encode_wide(r, in_limbs, out_limbs, ctx)
    assert np.array_equal(got.reshape(-1), expect)
    if bits + 2 * cw.bit_length() < 64 * out_limbs:
        assert rc == 0  # no overflow: the reference's checked adds would not have fired


@pytest.mark.gpu
def test_encode_wide_m_to_m_against_python_bigints(oracle, ctx):
    """the verifier's shape end to end on a small code: combine_rows on the device (Int<8> combined row), then
    encode_wide::<M, M> of it, against Python big-int arithmetic (oracle/pyoracle.py) -- and linearity:
    encode(sum_i c_i row_i) == sum_i c_i encode(row_i), the identity verify_column_testing checks per opened column"""
    import random

    from oracle import pyoracle as po
    from zinc_b200 import RaaCode, ZipTypes

    row_len, num_rows = 64, 8
    cw = 2 * row_len
    p1, p2 = oracle.perm_from_seed(cw, MOCK_SEEDS[0]), oracle.perm_from_seed(cw, MOCK_SEEDS[1])
    code = RaaCode.with_permutations(ZipTypes(), row_len, 2, p1, p2)
    rnd = random.Random(3)
    evals = [rnd.randrange(-(1 << 63), 1 << 63) for _ in range(num_rows * row_len)]
    coeffs = [rnd.randrange(-(1 << 63), 1 << 63) for _ in range(num_rows)]
    combined = [sum(coeffs[i] * evals[i * row_len + c] for i in range(num_rows)) for c in range(row_len)]
    enc_combined = po.encode_row_perm(combined, 2, p1.tolist(), p2.tolist())
    got = code.encode_wide(_signed_limbs(combined, 8), 8, 8, ctx)
    assert np.array_equal(got, _signed_limbs(enc_combined, 8))
    per_row = [po.encode_row_perm(evals[i * row_len:(i + 1) * row_len], 2, p1.tolist(), p2.tolist()) for i in range(num_rows)]
    lin = [sum(coeffs[i] * per_row[i][j] for i in range(num_rows)) for j in range(cw)]
    assert lin == enc_combined
