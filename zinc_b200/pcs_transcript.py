"""zinc_b200/pcs_transcript.py -- the proof-stream wire format of zinc's PcsTranscript for the data this path produces.

The reference's prover appends to `PcsTranscript::stream` (src/zip/pcs_transcript.rs); these are the byte layouts of
the calls the commit / open path makes, so that GPU outputs can be spliced into a proof stream unchanged:

    write_commitment(s)   pcs_transcript.rs:68-73,146-155   32 raw digest bytes each
    write_integer(s)      pcs_transcript.rs:108-128         every u64 limb little-endian, least significant limb first
    write_merkle_proof    pcs_transcript.rs:198-211         be64(path length) then the path digests
    read_*                pcs_transcript.rs:76-98,130-196   the inverses

`roots` as returned by zipgpu_commit* ARE write_commitments(roots); `rows` / combine_rows outputs ARE write_integers of
their values; zipgpu_data_open_columns_wire emits write_integers(column) + write_merkle_proof per row on the GPU.
(Field elements -- write_field_element, big-endian -- belong to the sumcheck side and are not produced here.)
"""
from __future__ import annotations

import numpy as np


class PcsStream:
    """A byte stream with the reference's write_* / read_* calls for integers, commitments and Merkle proofs."""

    def __init__(self, data: bytes = b""):
        self.stream = bytearray(data)
        self.pos = 0

    # ---- writing (prover) ----
    def write_commitment(self, digest: bytes) -> None:
        assert len(digest) == 32
        self.stream += digest

    def write_commitments(self, digests) -> None:
        """digests: iterable of 32-byte values, or a uint8 array [n, 32] as the GPU returns the roots"""
        if isinstance(digests, np.ndarray):
            assert digests.dtype == np.uint8 and digests.size % 32 == 0
            self.stream += digests.tobytes()
        else:
            for d in digests:
                self.write_commitment(d)

    def write_integer(self, limbs) -> None:
        """one Int<n>: n u64 limbs, least significant first (Integer::as_words), each little-endian"""
        self.stream += np.ascontiguousarray(limbs, dtype="<u8").tobytes()

    def write_integers(self, values: np.ndarray) -> None:
        """values: uint64 [count, n] (the layout every zipgpu call uses) -> count * n * 8 bytes"""
        self.stream += np.ascontiguousarray(values, dtype="<u8").tobytes()

    def write_merkle_proof(self, path) -> None:
        """path: list of 32-byte digests or uint8 [depth, 32]"""
        path = [bytes(p) for p in path]
        self.stream += len(path).to_bytes(8, "big")
        for p in path:
            self.write_commitment(p)

    # ---- reading (verifier) ----
    def _take(self, n: int) -> bytes:
        if self.pos + n > len(self.stream):
            raise EOFError("failed to fill whole buffer")  # io::ErrorKind::UnexpectedEof in the reference
        out = bytes(self.stream[self.pos:self.pos + n])
        self.pos += n
        return out

    def read_commitment(self) -> bytes:
        return self._take(32)

    def read_commitments(self, n: int) -> list[bytes]:
        return [self.read_commitment() for _ in range(n)]

    def read_integer(self, limbs: int) -> np.ndarray:
        return np.frombuffer(self._take(8 * limbs), dtype="<u8").astype(np.uint64)

    def read_integers(self, count: int, limbs: int) -> np.ndarray:
        return np.frombuffer(self._take(8 * limbs * count), dtype="<u8").astype(np.uint64).reshape(count, limbs)

    def read_merkle_proof(self) -> list[bytes]:
        n = int.from_bytes(self._take(8), "big")
        return [self.read_commitment() for _ in range(n)]

    def into_bytes(self) -> bytes:
        return bytes(self.stream)
