"""zinc_b200/transcript.py -- the `ZipTranscript::get_u64` sources RaaCode::new draws its seeds from.

Host logic, off the hot path.  Mirrors:
  KeccakTranscript   src/transcript.rs:14-55,142-155,183-185   (Keccak-256 Fiat-Shamir transcript)
  MockTranscript     src/zip/pcs/tests.rs:24-56                (counter: get_u64 -> 1, 2, ...)
"""
from __future__ import annotations

_RC = (
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
    0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
    0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
    0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
)
_M = (1 << 64) - 1


def _permute(st: list[int]) -> None:
    """Keccak-f[1600] on 25 lanes, lane (x, y) at st[x + 5*y]; rho/pi walked as the usual 24-step cycle."""
    for rc in _RC:
        col = [st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20] for x in range(5)]
        for x in range(5):
            d = col[(x + 4) % 5] ^ (((col[(x + 1) % 5] << 1) | (col[(x + 1) % 5] >> 63)) & _M)
            for y in range(0, 25, 5):
                st[x + y] ^= d
        x, y, cur = 1, 0, st[1]
        for t in range(24):
            x, y = y, (2 * x + 3 * y) % 5
            r = ((t + 1) * (t + 2) // 2) % 64
            st[x + 5 * y], cur = ((cur << r) | (cur >> (64 - r))) & _M if r else cur, st[x + 5 * y]
        for y in range(0, 25, 5):
            row = st[y:y + 5]
            for x in range(5):
                st[x + y] = row[x] ^ (~row[(x + 1) % 5] & _M & row[(x + 2) % 5])
        st[0] ^= rc


def keccak256(data: bytes) -> bytes:
    """Keccak-256 with the original 0x01 padding (what the `sha3::Keccak256` of transcript.rs:2 computes)."""
    rate = 136
    buf = bytearray(data)
    buf.append(0x01)
    buf.extend(b"\x00" * (-len(buf) % rate))
    buf[-1] |= 0x80
    st = [0] * 25
    for off in range(0, len(buf), rate):
        for i in range(rate // 8):
            st[i] ^= int.from_bytes(buf[off + 8 * i:off + 8 * i + 8], "little")
        _permute(st)
    return b"".join(st[i].to_bytes(8, "little") for i in range(4))


class KeccakTranscript:
    """src/transcript.rs:14-55 -- only the part RaaCode::new consumes (absorb / get_u64)."""

    def __init__(self) -> None:
        self._absorbed = bytearray()

    def absorb(self, v: bytes) -> None:  # transcript.rs:35-37
        self._absorbed += v

    def get_random_bytes(self, length: int) -> bytes:  # transcript.rs:41-55
        out = bytearray()
        counter = 0
        while len(out) < length:
            out += keccak256(bytes(self._absorbed) + counter.to_bytes(4, "big", signed=True))
            counter += 1
        return bytes(out[:length])

    def get_u64(self) -> int:  # transcript.rs:183-185 via get_integer_challenge::<Int<1>> (142-155)
        challenge = self.get_random_bytes(8)
        self._absorbed += b"\x12" + challenge + b"\x34"
        return int.from_bytes(challenge, "little")


class MockTranscript:
    """src/zip/pcs/tests.rs:24-56: a counter; RaaCode::new over it gets seeds 1 and 2."""

    def __init__(self) -> None:
        self.counter = 0

    def get_u64(self) -> int:
        self.counter += 1
        return self.counter
