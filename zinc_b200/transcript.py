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


class _Sponge:
    """Incremental Keccak-256 (rate 136): `update`, `copy`, `digest` -- what the reference does with
    `hasher.update`, `hasher.clone()` and `finalize()` (transcript.rs:35-55)."""

    RATE = 136

    def __init__(self) -> None:
        self.st = [0] * 25
        self.buf = bytearray()

    def copy(self) -> "_Sponge":
        c = _Sponge.__new__(_Sponge)
        c.st = list(self.st)
        c.buf = bytearray(self.buf)
        return c

    def _block(self, blk) -> None:
        st = self.st
        for i in range(self.RATE // 8):
            st[i] ^= int.from_bytes(blk[8 * i:8 * i + 8], "little")
        _permute(st)

    def update(self, data: bytes) -> None:
        self.buf += data
        while len(self.buf) >= self.RATE:
            self._block(self.buf[:self.RATE])
            del self.buf[:self.RATE]

    def digest(self) -> bytes:
        c = self.copy()
        blk = bytearray(c.buf) + b"\x01" + bytes(self.RATE - len(c.buf) - 1)
        blk[-1] |= 0x80
        c._block(blk)
        return b"".join(c.st[i].to_bytes(8, "little") for i in range(4))


class KeccakTranscript:
    """src/transcript.rs:14-55,142-201 -- the draws RaaCode::new and ZipLinearCode::new make."""

    def __init__(self) -> None:
        self._hasher = _Sponge()

    def absorb(self, v: bytes) -> None:  # transcript.rs:35-37
        self._hasher.update(bytes(v))

    def get_random_bytes(self, length: int) -> bytes:  # transcript.rs:41-55
        out = bytearray()
        counter = 0
        while len(out) < length:
            h = self._hasher.copy()
            h.update(counter.to_bytes(4, "big", signed=True))
            out += h.digest()
            counter += 1
        return bytes(out[:length])

    def get_u64(self) -> int:  # transcript.rs:183-185 via get_integer_challenge::<Int<1>> (142-155)
        challenge = self.get_random_bytes(8)
        self._hasher.update(b"\x12" + challenge + b"\x34")
        return int.from_bytes(challenge, "little")

    def get_usize_in_range(self, start: int, end: int) -> int:  # transcript.rs:161-172
        challenge = self._hasher.digest()
        self._hasher.update(b"\x88" + challenge + b"\x11")
        return start + int.from_bytes(challenge[:8], "little") % (end - start)

    def get_encoding_element(self) -> int:  # transcript.rs:176-181 (0 or 1; does not advance the state)
        return self.get_random_bytes(1)[0] & 1

    def sample_unique_columns(self, start: int, end: int, columns: set, count: int) -> int:  # transcript.rs:187-201
        added = 0
        while added < count:
            candidate = self.get_usize_in_range(start, end)
            if candidate not in columns:
                columns.add(candidate)
                added += 1
        return added


class MockTranscript:
    """src/zip/pcs/tests.rs:24-56: a counter; RaaCode::new over it gets seeds 1 and 2."""

    def __init__(self) -> None:
        self.counter = 0

    def get_u64(self) -> int:
        self.counter += 1
        return self.counter

    def get_encoding_element(self) -> int:
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
