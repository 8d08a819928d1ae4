"""oracle/pyoracle.py -- TEST INFRASTRUCTURE ONLY.

Second, independent CPU restatement of zinc's Zip commit path in pure Python (arbitrary-precision ints)
on top of the `blake3` PyPI package, which wraps the same Rust `blake3` crate the reference depends on
(Cargo.toml:30).  It pins the C oracle (oracle/zip_oracle.c) and generates tests/golden/*.json.  Only
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import anything under oracle/.

Citations are file:line relative to /root/reference.
"""
from __future__ import annotations

import struct

MASK64 = (1 << 64) - 1


# ----------------------------------------------------------------------------------------------
# Int<N> model (field/int.rs): two's complement, N u64 limbs, least-significant limb first
# ----------------------------------------------------------------------------------------------
def to_signed(words: list[int]) -> int:
    """limbs (LSW first, int.rs:230-232) -> Python int"""
    v = 0
    for i, w in enumerate(words):
        v |= (w & MASK64) << (64 * i)
    bits = 64 * len(words)
    return v - (1 << bits) if v >> (bits - 1) else v


def to_words(v: int, limbs: int) -> list[int]:
    """Python int -> limbs; sign-extending like `From<&Int<M>> for Int<N>` (int.rs:194-199)"""
    bits = 64 * limbs
    assert -(1 << (bits - 1)) <= v < (1 << (bits - 1)), "value does not fit (crypto-bigint add would panic)"
    v &= (1 << bits) - 1
    return [(v >> (64 * i)) & MASK64 for i in range(limbs)]


def int_to_bytes(v: int, limbs: int) -> bytes:
    """ToBytes for Int<N> (int.rs:201-210): limbs LSW->MSW, each limb big-endian."""
    return b"".join(struct.pack(">Q", w) for w in to_words(v, limbs))


# ----------------------------------------------------------------------------------------------
# RAA code (zip/code_raa.rs)
# ----------------------------------------------------------------------------------------------
def repeat(row: list[int], rep: int) -> list[int]:
    """code_raa.rs:142-152"""
    return [row[j % len(row)] for j in range(len(row) * rep)]


def accumulate(v: list[int]) -> list[int]:
    """code_raa.rs:164-171"""
    out = list(v)
    for i in range(1, len(out)):
        out[i] += out[i - 1]
    return out


def raa_row_len(poly_size: int) -> int:
    """code_raa.rs:42-43"""
    import math

    num_vars = poly_size.bit_length() - 1
    r = math.isqrt(1 << num_vars)
    return 1 << (r - 1).bit_length() if r > 1 else 1


def num_rows_for(poly_size: int, row_len: int) -> int:
    """pcs/structs.rs:79-90"""
    num_vars = poly_size.bit_length() - 1
    q = (1 << num_vars) // row_len
    return 1 << (q - 1).bit_length() if q > 1 else 1


def raa_width_bits(in_limbs: int, poly_size: int, rep: int) -> int:
    """code_raa.rs:53-67"""
    num_vars = poly_size.bit_length() - 1
    nv_even = num_vars if num_vars % 2 == 0 else num_vars + 1
    rep_log = (rep - 1).bit_length() if rep > 1 else 0
    return 64 * in_limbs + nv_even + 2 * rep_log


# ----------------------------------------------------------------------------------------------
# shuffle_seeded (zip/utils.rs:139-142) -- rand 0.9.2; PARITY UNPINNED against the real crate
# ----------------------------------------------------------------------------------------------
def _rotl32(x, n):
    return ((x << n) | (x >> (32 - n))) & 0xFFFFFFFF


def chacha_block(key: list[int], counter: int, stream: int, rounds: int) -> list[int]:
    s = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key) + [
        counter & 0xFFFFFFFF, (counter >> 32) & 0xFFFFFFFF, stream & 0xFFFFFFFF, (stream >> 32) & 0xFFFFFFFF]
    x = list(s)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl32(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl32(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl32(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl32(x[b] ^ x[c], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(x[i] + s[i]) & 0xFFFFFFFF for i in range(16)]


def seed_from_u64(state: int) -> list[int]:
    """rand_core 0.9 SeedableRng::seed_from_u64 (PCG32 expansion to 8 LE words)"""
    key = []
    for _ in range(8):
        state = (state * 6364136223846793005 + 11634580027462260723) & MASK64
        xorshifted = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
        rot = state >> 59
        key.append(((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & 0xFFFFFFFF)
    return key


class StdRng:
    """rand 0.9 StdRng = ChaCha12, 64-bit counter, stream 0; words handed out in order."""

    def __init__(self, seed: int):
        self.key = seed_from_u64(seed)
        self.counter = 0
        self.buf: list[int] = []

    def next_u32(self) -> int:
        if not self.buf:
            self.buf = chacha_block(self.key, self.counter, 0, 12)
            self.counter += 1
        return self.buf.pop(0)

    def below(self, bound: int) -> int:
        """UniformInt<u32>::sample_single_inclusive, Canon's method (rand 0.9, not `unbiased`)"""
        m = self.next_u32() * bound
        hi, lo = m >> 32, m & 0xFFFFFFFF
        if lo > ((-bound) & 0xFFFFFFFF):
            new_hi = (self.next_u32() * bound) >> 32
            if lo + new_hi > 0xFFFFFFFF:
                hi += 1
        return hi


def _calc_bound(m: int):
    product, current = m, m + 1
    while product * current <= 0xFFFFFFFF:
        product *= current
        current += 1
    return product, current - m


def shuffle_seeded(v: list, seed: int) -> list:
    """zip/utils.rs:139-142 with rand 0.9 SliceRandom::shuffle (IncreasingUniform chooser)."""
    v = list(v)
    if len(v) <= 1:
        return v
    rng = StdRng(seed)
    n, chunk, remaining = 0, 0, 1
    for i in range(len(v)):
        next_n = n + 1
        if remaining > 0:
            next_remaining = remaining - 1
        else:
            bound, cnt = _calc_bound(next_n)
            chunk = rng.below(bound)
            next_remaining = cnt - 1
        if next_remaining == 0:
            j = chunk
        else:
            j = chunk % next_n
            chunk //= next_n
        remaining, n = next_remaining, next_n
        v[i], v[j] = v[j], v[i]
    return v


def perm_from_seed(n: int, seed: int) -> list[int]:
    return shuffle_seeded(list(range(n)), seed)


# ----------------------------------------------------------------------------------------------
# encode (code_raa.rs:89-105) / encode_rows (commit.rs:158-183)
# ----------------------------------------------------------------------------------------------
def encode_row_seeded(row: list[int], rep: int, seed1: int, seed2: int) -> list[int]:
    r = repeat(row, rep)
    r = accumulate(shuffle_seeded(r, seed1))
    return accumulate(shuffle_seeded(r, seed2))


def encode_row_perm(row: list[int], rep: int, perm1: list[int], perm2: list[int]) -> list[int]:
    r = repeat(row, rep)
    r = accumulate([r[p] for p in perm1])
    return accumulate([r[p] for p in perm2])


def field_add(a: int, b: int, modulus: int, limbs: int) -> int:
    """RandomField AddAssign (field/arithmetic.rs:66-77 -> FieldConfig::add_assign / reduce_modulus, field/config.rs:53-76):
    s = a + b wrapping at 2^(64 limbs) with carry c; if c or s >= modulus: s -= modulus (wrapping)"""
    full = a + b
    carry, s = full >> (64 * limbs), full & ((1 << (64 * limbs)) - 1)
    if carry or s >= modulus:
        s = (s - modulus) & ((1 << (64 * limbs)) - 1)
    return s


def encode_f_row(row: list[int], rep: int, perm1, perm2, modulus: int, limbs: int) -> list[int]:
    """RaaCode::encode_f (code_raa.rs:133-138) = encode_inner (code_raa.rs:89-105) with Out = F: repeat (142-152, a clone
    of every element), shuffle, accumulate (164-171) with the field's +=, shuffle, accumulate.  Elements are the stored
    residues of the RandomField values."""
    n = len(row)
    y = [row[perm1[i] % n] for i in range(n * rep)]  # repeat o shuffle_seeded in gather form
    for i in range(1, len(y)):
        y[i] = field_add(y[i], y[i - 1], modulus, limbs)
    y = [y[perm2[i]] for i in range(len(y))]
    for i in range(1, len(y)):
        y[i] = field_add(y[i], y[i - 1], modulus, limbs)
    return y


def encode_rows(evals: list[int], num_rows: int, row_len: int, rep: int, perm1, perm2) -> list[int]:
    out: list[int] = []
    for i in range(num_rows):
        out += encode_row_perm(evals[i * row_len:(i + 1) * row_len], rep, perm1, perm2)
    return out


# ----------------------------------------------------------------------------------------------
# MerkleTree (pcs/utils.rs:66-118)
# ----------------------------------------------------------------------------------------------
def combine_rows(coeffs: list[int], evals: list[int], row_len: int, out_limbs: int) -> list[int]:
    """zip/utils.rs:94-127 as used by the proximity test (open_z.rs:100-113): combined[col] = sum_i coeff_i *
    evals[i*row_len + col] in M = Int<out_limbs> after `expand` (zip/utils.rs:129-137).  crypto-bigint's checked
    mul/add would panic on overflow; `to_words` asserts the same range."""
    out = []
    for col in range(row_len):
        acc = 0
        for i, c in enumerate(coeffs):
            acc += c * evals[i * row_len + col]
        to_words(acc, out_limbs)  # range check
        out.append(acc)
    return out


def merkle_tree(depth: int, leaves: list[int], leaf_limbs: int):
    """returns (root: bytes, layers: list[bytes]) with layers laid out as in utils.rs:77-85"""
    import blake3

    n = len(leaves)
    assert n & (n - 1) == 0 and n == 1 << depth  # utils.rs:75-76
    layers = [blake3.blake3(int_to_bytes(v, leaf_limbs)).digest() for v in leaves]  # utils.rs:87-93
    offset = 0
    for d in range(depth, 0, -1):  # utils.rs:98
        width = 1 << d
        cur = layers[offset:offset + width]
        for k in range(width // 2):
            layers.append(blake3.blake3(cur[2 * k] + cur[2 * k + 1]).digest())  # utils.rs:107-111
        offset += width
    root = layers.pop()
    return root, layers


def commit(evals: list[int], num_rows: int, row_len: int, rep: int, perm1, perm2, out_limbs: int):
    """commit.rs:50-87 -> (rows, layers per row, roots)"""
    cw = row_len * rep
    depth = (cw - 1).bit_length() if cw > 1 else 0  # commit.rs:67
    rows = encode_rows(evals, num_rows, row_len, rep, perm1, perm2)
    roots, all_layers = [], []
    for i in range(num_rows):
        root, layers = merkle_tree(depth, rows[i * cw:(i + 1) * cw], out_limbs)
        roots.append(root)
        all_layers.append(layers)
    return rows, all_layers, roots


# ----------------------------------------------------------------------------------------------
# Keccak-256 + KeccakTranscript::get_u64 (transcript.rs:41-55,142-155,183-185): source of the RAA seeds
# ----------------------------------------------------------------------------------------------
_KRC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
        0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
        0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
        0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
        0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_KROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]


def _keccak_f(a):
    rol = lambda x, n: ((x << n) | (x >> (64 - n))) & MASK64 if n else x
    for rc in _KRC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = rol(a[x][y], _KROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    return a


def keccak256(data: bytes, pad: int = 0x01) -> bytes:
    """Keccak-256 (pad=0x01, what sha3::Keccak256 computes); pad=0x06 gives SHA3-256 for cross-checking."""
    rate = 136
    msg = bytearray(data) + bytes([pad]) + bytes((-len(data) - 1) % rate)
    msg[-1] |= 0x80
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            a[i % 5][i // 5] ^= int.from_bytes(msg[off + 8 * i:off + 8 * i + 8], "little")
        a = _keccak_f(a)
    return b"".join(a[i % 5][i // 5].to_bytes(8, "little") for i in range(4))


class KeccakTranscript:
    """transcript.rs:14-55,142-155,183-185 (only what RaaCode::new consumes)"""

    def __init__(self):
        self.buf = b""

    def absorb(self, v: bytes):
        self.buf += v

    def get_random_bytes(self, length: int) -> bytes:
        out, counter = b"", 0
        while len(out) < length:
            out += keccak256(self.buf + struct.pack(">i", counter))
            counter += 1
        return out[:length]

    def get_u64(self) -> int:
        ch = self.get_random_bytes(8)
        self.buf += b"\x12" + ch + b"\x34"
        return int.from_bytes(ch, "little")

    # --- the draws ZipLinearCode::new makes (zip/code.rs:271-296) ---
    def get_usize_in_range(self, start: int, end: int) -> int:
        """transcript.rs:161-172"""
        ch = keccak256(self.buf)
        self.buf += b"\x88" + ch + b"\x11"
        return start + int.from_bytes(ch[:8], "little") % (end - start)

    def get_encoding_element(self) -> int:
        """transcript.rs:176-181: the LSB of one random byte; the state is NOT advanced"""
        return self.get_random_bytes(1)[0] & 1

    def sample_unique_columns(self, start: int, end: int, columns: set, count: int) -> int:
        """transcript.rs:187-201"""
        added = 0
        while added < count:
            c = self.get_usize_in_range(start, end)
            if c not in columns:
                columns.add(c)
                added += 1
        return added


class MockTranscript:
    """zip/pcs/tests.rs:24-56"""

    def __init__(self):
        self.counter = 0

    def get_encoding_element(self) -> int:
        self.counter += 1
        return self.counter

    def get_u64(self) -> int:
        self.counter += 1
        return self.counter

    def sample_unique_columns(self, start: int, end: int, columns: set, count: int) -> int:
        self.counter += 1
        inserted = 0
        for i in range(start, end):
            if i not in columns:
                columns.add(i)
                inserted += 1
                if inserted == count:
                    break
        return inserted


def sparse_matrix_sample_new(n: int, m: int, d: int, transcript):
    """SparseMatrixZ::sample_new (code.rs:271-296): per matrix row draw d unique columns into a BTreeSet, then one
    encoding element per column in ascending column order.  -> (cols[n*d], coef[n*d])"""
    cols, coef = [], []
    for _ in range(n):
        chosen: set = set()
        transcript.sample_unique_columns(0, m, chosen, d)
        for c in sorted(chosen):
            cols.append(c)
            coef.append(transcript.get_encoding_element())
    return cols, coef


def zip_linear_code_new(poly_size: int, transcript, rep: int = 2):
    """ZipLinearCode::new / new_multilinear (code.rs:100-147): row_len, codeword_len and the matrices a, b of
    dimension (codeword_len/2) x row_len with row_len/2 cells per row (code.rs:134)."""
    assert poly_size & (poly_size - 1) == 0
    num_vars = poly_size.bit_length() - 1
    n_0 = min(20, (1 << num_vars) - 1)
    assert (1 << num_vars) > n_0
    row_len = raa_row_len(poly_size)  # the same isqrt().next_power_of_two() (code.rs:127)
    cw = row_len * rep
    a = sparse_matrix_sample_new(cw // 2, row_len, row_len // 2, transcript)
    b = sparse_matrix_sample_new(cw // 2, row_len, row_len // 2, transcript)
    return row_len, cw, a, b


def sparse_mat_vec(n: int, d: int, cols, coef, vec: list[int]) -> list[int]:
    """SparseMatrixZ::mat_vec_mul (code.rs:299-321), exact Python ints"""
    return [sum(coef[i * d + k] * vec[cols[i * d + k]] for k in range(d)) for i in range(n)]


def sparse_encode_row(row: list[int], n: int, d: int, a, b) -> list[int]:
    """ZipLinearCode::encode_wide (code.rs:186-201)"""
    return sparse_mat_vec(n, d, a[0], a[1], row) + sparse_mat_vec(n, d, b[0], b[1], row)
