"""shared helpers for the parity tests"""
import numpy as np

KECCAK_SEEDS = (0xB9736F582676E7E8, 0xD7397E6260CE9C3E)  # fresh KeccakTranscript (zip_benches.rs:102-106)
MOCK_SEEDS = (1, 2)  # MockTranscript (pcs/tests.rs:24-56)

I64_MAX = (1 << 63) - 1
I64_MIN = -(1 << 63)


def shape_for(nv: int):
    """row_len / num_rows / cw for a 2^nv MLE: row_len = isqrt(2^nv).next_power_of_two() (code_raa.rs:42-43),
    num_rows = (2^nv / row_len).next_power_of_two() (structs.rs:82), cw = 2 * row_len (rep = 2)"""
    import math

    r = math.isqrt(1 << nv)
    row_len = 1 << (r - 1).bit_length() if r > 1 else 1
    q = (1 << nv) // row_len
    num_rows = 1 << (q - 1).bit_length() if q > 1 else 1
    return row_len, num_rows, 2 * row_len


def input_patterns(nv: int, seed: int = 0):
    """the input families the reference's tests use (commit.rs:234,266-267,281,475,487-489,620-621) + random"""
    n = 1 << nv
    rng = np.random.default_rng(0x21C0 + nv + seed)
    yield "random", rng.integers(I64_MIN, I64_MAX, size=n, dtype=np.int64, endpoint=True)
    yield "one_to_n", np.arange(1, n + 1, dtype=np.int64)
    yield "zeros", np.zeros(n, dtype=np.int64)
    yield "i64_max", np.full(n, I64_MAX, dtype=np.int64)
    yield "i64_min", np.full(n, I64_MIN, dtype=np.int64)
    yield "alternating", np.where(np.arange(n) % 2 == 0, 1, -1).astype(np.int64)
    yield "const42", np.full(n, 42, dtype=np.int64)
