"""zinc_b200/zip.py -- host-side mirror of zinc's Zip PCS *commit* API on top of libzipgpu (C ABI).

The reference is a Rust crate and no Rust toolchain exists in this image, so the host layer a Rust maintainer
would write is mirrored here with the same names, argument meaning and error behaviour, so
that the parity tests read like the reference's own tests:

    RaaCode.new / row_len / codeword_len / encode_wide       src/zip/code_raa.rs:35-131
    DefaultLinearCodeSpec                                    src/zip/code.rs:229-242
    MultilinearZip.setup                                     src/zip/pcs/structs.rs:79-90
    MultilinearZip.commit / commit_no_merkle / batch_commit / encode_rows
                                                             src/zip/pcs/commit.rs:50-183
    MerkleTree.new, MerkleProof.create_proof                 src/zip/pcs/utils.rs:66-118,163-176
    DenseMultilinearExtension                                src/poly_z/mle/dense.rs:22-64

`Int<N>` values are numpy uint64 arrays whose last axis holds the N limbs, least significant first
(field/int.rs:230-232); a plain 1-D int64/uint64 array is accepted for N = 1.  Rust `Err` results are raised
as `Error`, Rust panics (assert!/assert_eq!) as AssertionError carrying the reference's message.
All arithmetic and hashing happens on the GPU; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass, field

import numpy as np

from . import _native as nat


class Error(Exception):
    """zip::Error (src/zip/zip.rs:12-24)"""


class InvalidPcsParam(Error):
    pass


# ----------------------------------------------------------------------------------------------------------
# context
# ----------------------------------------------------------------------------------------------------------
_uid_counter = [0]


def _next_uid() -> int:
    _uid_counter[0] += 1
    return _uid_counter[0]


class Context:
    """One zipgpu context = one GPU (one process per GPU under torch.distributed).

    Native handles that live inside the context (per-pp codes, resident prover data, roots exchanges) are registered
    here, keyed by the owning Python object's uid -- never by id(), which Python reuses -- and are destroyed by close()
    BEFORE the context, so that nothing can point into a destroyed context afterwards."""

    multi = False

    def __init__(self, device: int | None = None):
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        nat.check(nat.lib().zipgpu_ctx_create(device, C.byref(h)))
        self.handle = h
        self.device = device
        self._codes: dict[tuple[int, int, int], C.c_void_p] = {}   # (code uid, in_limbs, out_limbs) -> zipgpu_code*
        self._live: dict[int, weakref.ref] = {}                    # uid -> weak ref to an object with _release()
        self._stream_buf = None                                    # (ptr, capacity): pinned proof-stream buffer

    def proof_stream_buffer(self, nbytes: int) -> np.ndarray:
        """a uint8 view of `nbytes` of this context's PINNED proof-stream buffer (zipgpu_host_alloc; grown on demand, freed
        with the context).  What `open` appends to the transcript stream (pcs_transcript.rs:115-135,198-211) is tens of
        MB per proof: into fresh pageable memory the page faults and the staged copy cost 20x the gather + PCIe time.
        The view is valid until the next call."""
        self._check_open()
        if self._stream_buf is None or self._stream_buf[1] < nbytes:
            if self._stream_buf is not None:
                nat.lib().zipgpu_host_free(self._stream_buf[0])
                self._stream_buf = None
            cap = max(int(nbytes), 1 << 20)
            p = C.c_void_p()
            nat.check(nat.lib().zipgpu_host_alloc(cap, C.byref(p)))
            self._stream_buf = (p, cap)
        buf = (C.c_uint8 * nbytes).from_address(self._stream_buf[0].value)
        return np.frombuffer(buf, dtype=np.uint8, count=nbytes)

    def _check_open(self) -> None:
        if not self.handle:
            raise ZipGpuClosed("this zipgpu context has been closed")

    def sync(self) -> None:
        self._check_open()
        nat.check(nat.lib().zipgpu_ctx_sync(self.handle))

    @property
    def launch_count(self) -> int:
        self._check_open()
        return int(nat.lib().zipgpu_ctx_launch_count(self.handle))

    def _destroy_code(self, h) -> None:
        nat.lib().zipgpu_code_destroy(h)

    def drop_code(self, code) -> None:
        """destroy the native handles of one code object (RaaCode / ZipLinearCode) in this context"""
        for key in [k for k in self._codes if k[0] == code._uid]:
            self._destroy_code(self._codes.pop(key))

    def close(self) -> None:
        if self.handle:
            for ref in list(self._live.values()):
                obj = ref()
                if obj is not None:
                    obj._release()
            self._live.clear()
            for h in self._codes.values():
                self._destroy_code(h)
            self._codes.clear()
            if self._stream_buf is not None:
                nat.lib().zipgpu_host_free(self._stream_buf[0])
                self._stream_buf = None
            self._destroy_ctx()
            self.handle = None

    def _destroy_ctx(self) -> None:
        nat.lib().zipgpu_ctx_destroy(self.handle)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ZipGpuClosed(RuntimeError):
    pass


class MultiContext(Context):
    """zipgpu_mgpu: ONE process driving several GPUs behind the same commit API (include/zipgpu.h).  Pass it wherever
    a Context goes: MultilinearZip.commit / commit_no_merkle / batch_commit / commit_resident shard rows (or whole
    polynomials) over the devices and return results in the reference's order; the roots are exchanged between the
    GPUs by the kernel that produces them."""

    multi = True

    def __init__(self, devices: list[int] | None = None, n: int = 0):
        h = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            nat.check(nat.lib().zipgpu_mgpu_create(arr, len(devices), C.byref(h)))
        else:
            nat.check(nat.lib().zipgpu_mgpu_create(None, n, C.byref(h)))
        self.handle = h
        self.num_devices = int(nat.lib().zipgpu_mgpu_num_devices(h))
        self.device = None
        self._codes = {}
        self._live = {}
        self._stream_buf = None

    def sync(self) -> None:
        self._check_open()
        for g in range(self.num_devices):
            nat.check(nat.lib().zipgpu_ctx_sync(C.c_void_p(nat.lib().zipgpu_mgpu_ctx(self.handle, g))))

    @property
    def launch_count(self) -> int:
        self._check_open()
        return int(nat.lib().zipgpu_mgpu_launch_count(self.handle))

    def _destroy_code(self, h) -> None:
        nat.lib().zipgpu_mgpu_code_destroy(h)

    def _destroy_ctx(self) -> None:
        nat.lib().zipgpu_mgpu_destroy(self.handle)


_default_ctx: Context | None = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


# ----------------------------------------------------------------------------------------------------------
# types
# ----------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ZipTypes:
    """traits/types.rs:202-217: limb counts of the four integer types of a Zip instantiation."""

    N: int = 1
    L: int = 2
    K: int = 4
    M: int = 8


def RandomFieldZipTypes(int_limbs: int = 1) -> ZipTypes:
    """field/int.rs:276-289: N = Int<n>, L = Int<2n>, K = Int<4n>, M = Int<8n>."""
    return ZipTypes(int_limbs, 2 * int_limbs, 4 * int_limbs, 8 * int_limbs)


def as_limbs(values, limbs: int) -> np.ndarray:
    """-> contiguous uint64 array [n, limbs] (two's complement, LSW first); int64 input is sign-extended."""
    a = np.asarray(values)
    if a.dtype == np.int64 and (a.ndim == 1 or limbs == 1 and a.ndim == 2 and a.shape[1] == 1):
        a = a.reshape(-1)
        out = np.empty((a.size, limbs), dtype=np.uint64)
        out[:, 0] = a.view(np.uint64)
        if limbs > 1:
            out[:, 1:] = np.where(a < 0, np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64(0))[:, None]
        return out
    if a.dtype == np.int64:  # already limb-shaped: [n, limbs] (or flat, n * limbs) two's-complement words
        a = np.ascontiguousarray(a).view(np.uint64)
    elif a.dtype.kind in "iu" and a.dtype != np.uint64:
        return as_limbs(a.astype(np.int64), limbs)
    if a.dtype == object:  # python ints
        flat = [int(v) for v in a.reshape(-1)]
        out = np.empty((len(flat), limbs), dtype=np.uint64)
        for i, v in enumerate(flat):
            v &= (1 << (64 * limbs)) - 1
            for l in range(limbs):
                out[i, l] = (v >> (64 * l)) & 0xFFFFFFFFFFFFFFFF
        return out
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim == 1:
        assert limbs == 1 or a.size % limbs == 0
        return a.reshape(-1, limbs) if limbs > 1 else a.reshape(-1, 1)
    assert a.shape[-1] == limbs, f"expected {limbs} limbs, got {a.shape[-1]}"
    return a.reshape(-1, limbs)


@dataclass
class DenseMultilinearExtension:
    """poly_z/mle/dense.rs:22-28"""

    evaluations: np.ndarray  # uint64 [2^num_vars, N]
    num_vars: int

    @staticmethod
    def from_evaluations_vec(num_vars: int, evaluations, limbs: int = 1) -> "DenseMultilinearExtension":
        ev = as_limbs(evaluations, limbs)
        # dense.rs:45-49
        assert ev.shape[0] <= (1 << num_vars), (
            f"The size of evaluations should not exceed 2^num_vars. \n eval len: {ev.shape[0]}. num vars: {num_vars}")
        if ev.shape[0] != (1 << num_vars):  # dense.rs:51-58: zero-pad
            pad = np.zeros(((1 << num_vars) - ev.shape[0], limbs), dtype=np.uint64)
            ev = np.concatenate([ev, pad], axis=0)
        return DenseMultilinearExtension(np.ascontiguousarray(ev), num_vars)

    from_evaluations_slice = from_evaluations_vec

    @staticmethod
    def rand(num_vars: int, rng: np.random.Generator, limbs: int = 1) -> "DenseMultilinearExtension":
        """dense.rs:140-145: every limb uniform (Int::random, int.rs:187-192)"""
        ev = rng.integers(0, 1 << 64, size=((1 << num_vars), limbs), dtype=np.uint64)
        return DenseMultilinearExtension(ev, num_vars)


class DefaultLinearCodeSpec:
    """code.rs:229-242"""

    def num_column_opening(self) -> int:
        return 1000

    def repetition_factor(self) -> int:
        return 2

    def num_proximity_testing(self, _log2_q: int, _n: int, _n_0: int) -> int:
        return 1


def shuffle_seeded_indices(n: int, seed: int) -> np.ndarray:
    """The index form of zip/utils.rs:139-142: `shuffle_seeded` applied to [0..n)."""
    out = np.empty(n, dtype=np.uint32)
    nat.check(nat.lib().zipgpu_perm_from_seed(seed & 0xFFFFFFFFFFFFFFFF, n, nat.ptr(out)))
    return out


class RaaCode:
    """code_raa.rs:16-139.  The permutations are materialised once per code (they depend only on the seeds and
    the codeword length) and uploaded to the GPU on first use."""

    def __init__(self, zt: ZipTypes, row_len: int, repetition_factor: int, num_column_opening: int,
                 num_proximity_testing: int, perm_1_seed: int | None, perm_2_seed: int | None,
                 perms: tuple[np.ndarray, np.ndarray] | None = None):
        self.zt = zt
        self._row_len = row_len
        self.repetition_factor = repetition_factor
        self._num_column_opening = num_column_opening
        self._num_proximity_testing = num_proximity_testing
        self.perm_1_seed = perm_1_seed
        self.perm_2_seed = perm_2_seed
        self._perms = perms
        self._uid = _next_uid()

    @staticmethod
    def new(spec, poly_size: int, transcript, zt: ZipTypes = ZipTypes()) -> "RaaCode":
        """code_raa.rs:35-86"""
        num_vars = poly_size.bit_length() - 1  # ilog2
        row_len = int(nat.lib().zipgpu_raa_row_len(1 << num_vars))
        repetition_factor = spec.repetition_factor()
        num_column_opening = spec.num_column_opening()
        log2_q = zt.N
        n_0 = min(20, (1 << num_vars) - 1)
        num_proximity_testing = spec.num_proximity_testing(log2_q, row_len, n_0)
        width = int(nat.lib().zipgpu_raa_codeword_width_bits(zt.N, 1 << num_vars, repetition_factor))
        assert 64 * zt.K >= width, f"Cannot fit {width}-bit wide codeword entries in {64 * zt.K} bits integers"
        perm_1_seed = transcript.get_u64()
        perm_2_seed = transcript.get_u64()
        return RaaCode(zt, row_len, repetition_factor, num_column_opening, num_proximity_testing, perm_1_seed,
                       perm_2_seed)

    @staticmethod
    def with_permutations(zt: ZipTypes, row_len: int, repetition_factor: int, perm1, perm2) -> "RaaCode":
        """What a Rust host does: hand over the arrays produced by the real `shuffle_seeded`."""
        p1 = np.ascontiguousarray(perm1, dtype=np.uint32)
        p2 = np.ascontiguousarray(perm2, dtype=np.uint32)
        return RaaCode(zt, row_len, repetition_factor, 1000, 1, None, None, (p1, p2))

    def row_len(self) -> int:
        return self._row_len

    def codeword_len(self) -> int:
        return self._row_len * self.repetition_factor

    def num_column_opening(self) -> int:
        return self._num_column_opening

    def num_proximity_testing(self) -> int:
        return self._num_proximity_testing

    def permutations(self) -> tuple[np.ndarray, np.ndarray]:
        if self._perms is None:
            cw = self.codeword_len()
            self._perms = (shuffle_seeded_indices(cw, self.perm_1_seed), shuffle_seeded_indices(cw, self.perm_2_seed))
        return self._perms

    def native(self, ctx: Context, in_limbs: int, out_limbs: int) -> C.c_void_p:
        """the per-pp device state in `ctx` (zipgpu_code, or zipgpu_mgpu_code for a MultiContext); owned by the context"""
        ctx._check_open()
        key = (self._uid, in_limbs, out_limbs)
        h = ctx._codes.get(key)
        if h is None:
            p1, p2 = self.permutations()
            h = C.c_void_p()
            create = nat.lib().zipgpu_mgpu_code_create if ctx.multi else nat.lib().zipgpu_code_create
            rc = create(ctx.handle, self._row_len, self.repetition_factor, in_limbs, out_limbs,
                        nat.ptr(p1), nat.ptr(p2), C.byref(h))
            if rc == nat.ERR_WIDTH:
                raise AssertionError(nat.lib().zipgpu_last_error().decode())
            nat.check(rc)
            ctx._codes[key] = h
        return h

    def encode_wide(self, row, in_limbs: int | None = None, out_limbs: int | None = None,
                    ctx: Context | None = None) -> np.ndarray:
        """code_raa.rs:125-131 -> [cw, out_limbs] uint64"""
        in_limbs = in_limbs or self.zt.N
        out_limbs = out_limbs or self.zt.K
        r = as_limbs(row, in_limbs)
        assert r.shape[0] == self._row_len, "Row length must match the code's row length"  # code_raa.rs:93-97
        ctx = ctx or default_context()
        out = np.empty((self.codeword_len(), out_limbs), dtype=np.uint64)
        if in_limbs > 2:
            # wider inputs than the prover's evaluations: the verifier's encode_wide::<M, M> of a combined row
            # (verify_z.rs:74-78) -- the latency encoder, permutations of the same code
            if ctx.multi:
                raise Error("encode_wide of Int<%d> rows runs on one GPU: pass a Context" % in_limbs)
            nat.check(nat.lib().zipgpu_encode_wide(self.native(ctx, self.zt.N, self.zt.K), 1, in_limbs, out_limbs, nat.ptr(r),
                                                   nat.ptr(out)))
            return out
        enc = nat.lib().zipgpu_mgpu_encode_rows if ctx.multi else nat.lib().zipgpu_encode_rows
        nat.check(enc(self.native(ctx, in_limbs, out_limbs), 1, nat.ptr(r), nat.ptr(out)))
        return out

    def encode(self, row, ctx: Context | None = None) -> np.ndarray:
        """code.rs:35-37: encode == encode_wide::<N, M>"""
        return self.encode_wide(row, self.zt.N, self.zt.M, ctx)

    def encode_f(self, row, modulus: int, limbs: int, ctx: Context | None = None) -> np.ndarray:
        """code_raa.rs:133-138: the same code over field elements, every += the field's addition.  `row`: the stored
        residues of the RandomField<limbs> values ([row_len, limbs] uint64, python ints also accepted), all < modulus;
        -> [cw, limbs] uint64.  (The verifier's re-encoding of the combined row, verify_z.rs:141-142.)"""
        r = as_limbs(np.asarray(row, dtype=object) if not isinstance(row, np.ndarray) else row, limbs)
        assert r.shape[0] == self._row_len, "Row length must match the code's row length"  # code_raa.rs:93-97
        ctx = ctx or default_context()
        if ctx.multi:
            raise Error("encode_f runs on one GPU: pass a Context")
        mod = as_limbs(np.array([modulus], dtype=object), limbs).reshape(-1)
        out = np.empty((self.codeword_len(), limbs), dtype=np.uint64)
        nat.check(nat.lib().zipgpu_encode_f(self.native(ctx, self.zt.N, self.zt.K), 1, limbs, nat.ptr(mod), nat.ptr(r),
                                            nat.ptr(out)))
        return out


class SparseMatrixZ:
    """zip/code.rs:265-336: `n` rows of `d` (column, coefficient) cells over `m` columns; cells in row order."""

    def __init__(self, n: int, m: int, d: int, cols, coef):
        self.n, self.m, self.d = n, m, d
        self.cols = np.ascontiguousarray(cols, dtype=np.uint32).reshape(n * d)
        self.coef = np.ascontiguousarray(coef, dtype=np.int64).reshape(n * d)

    @staticmethod
    def sample_new(n: int, m: int, d: int, transcript) -> "SparseMatrixZ":
        """code.rs:274-296: per row, d unique columns into a BTreeSet, then one encoding element per column in
        ascending column order."""
        cols = np.empty(n * d, dtype=np.uint32)
        coef = np.empty(n * d, dtype=np.int64)
        for i in range(n):
            chosen: set = set()
            transcript.sample_unique_columns(0, m, chosen, d)
            for k, c in enumerate(sorted(chosen)):
                cols[i * d + k] = c
                coef[i * d + k] = transcript.get_encoding_element()
        return SparseMatrixZ(n, m, d, cols, coef)

    def to_dense(self) -> np.ndarray:
        out = np.zeros((self.n, self.m), dtype=np.int64)
        for i in range(self.n):
            for k in range(self.d):
                out[i, self.cols[i * self.d + k]] = self.coef[i * self.d + k]
        return out


class ZipLinearCode:
    """zip/code.rs:77-215, the sparse code: encode_wide(row) = a.mat_vec_mul(row) ‖ b.mat_vec_mul(row).  The two
    matrices are sampled on the host (they depend only on the transcript and the shape) and uploaded on first use."""

    def __init__(self, zt: ZipTypes, row_len: int, codeword_len: int, num_column_opening: int,
                 num_proximity_testing: int, a: SparseMatrixZ, b: SparseMatrixZ):
        assert a.n == b.n == codeword_len // 2 and a.m == b.m == row_len and a.d == b.d
        self.zt = zt
        self._row_len = row_len
        self._codeword_len = codeword_len
        self.repetition_factor = codeword_len // row_len
        self._num_column_opening = num_column_opening
        self._num_proximity_testing = num_proximity_testing
        self.a, self.b = a, b
        self._uid = _next_uid()

    @staticmethod
    def new(spec, poly_size: int, transcript, zt: ZipTypes = ZipTypes()) -> "ZipLinearCode":
        """code.rs:100-147"""
        assert poly_size & (poly_size - 1) == 0 and poly_size > 0
        num_vars = poly_size.bit_length() - 1
        n_0 = min(20, (1 << num_vars) - 1)
        assert (1 << num_vars) > n_0
        row_len = int(nat.lib().zipgpu_raa_row_len(1 << num_vars))  # the same isqrt().next_power_of_two(), code.rs:127
        codeword_len = row_len * spec.repetition_factor()
        num_proximity_testing = spec.num_proximity_testing(zt.N, row_len, n_0)
        a = SparseMatrixZ.sample_new(codeword_len // 2, row_len, row_len // 2, transcript)  # code.rs:134,150-161
        b = SparseMatrixZ.sample_new(codeword_len // 2, row_len, row_len // 2, transcript)
        return ZipLinearCode(zt, row_len, codeword_len, spec.num_column_opening(), num_proximity_testing, a, b)

    @staticmethod
    def with_matrices(zt: ZipTypes, row_len: int, codeword_len: int, a: SparseMatrixZ, b: SparseMatrixZ) -> "ZipLinearCode":
        """What a Rust host does: hand over the cells of the matrices `ZipLinearCode::new` sampled."""
        return ZipLinearCode(zt, row_len, codeword_len, 1000, 1, a, b)

    def row_len(self) -> int:
        return self._row_len

    def codeword_len(self) -> int:
        return self._codeword_len

    def num_column_opening(self) -> int:
        return self._num_column_opening

    def num_proximity_testing(self) -> int:
        return self._num_proximity_testing

    def native(self, ctx: Context, in_limbs: int, out_limbs: int) -> C.c_void_p:
        ctx._check_open()
        if ctx.multi:
            raise Error("the sparse ZipLinearCode is single-GPU (the multi-GPU context takes RaaCode)")
        key = (self._uid, in_limbs, out_limbs)
        h = ctx._codes.get(key)
        if h is None:
            h = C.c_void_p()
            nat.check(nat.lib().zipgpu_sparse_code_create(ctx.handle, self._row_len, self._codeword_len, self.a.d, in_limbs,
                                                          out_limbs, nat.ptr(self.a.cols), nat.ptr(self.a.coef),
                                                          nat.ptr(self.b.cols), nat.ptr(self.b.coef), C.byref(h)))
            ctx._codes[key] = h
        return h

    def kernel_kind(self, ctx: Context | None = None, in_limbs: int | None = None, out_limbs: int | None = None) -> str:
        ctx = ctx or default_context()
        k = nat.lib().zipgpu_code_sparse_kind(self.native(ctx, in_limbs or self.zt.N, out_limbs or self.zt.K))
        return {1: "tensor", 0: "generic"}[k]

    def encode_wide(self, row, in_limbs: int | None = None, out_limbs: int | None = None,
                    ctx: Context | None = None) -> np.ndarray:
        """code.rs:186-201 -> [cw, out_limbs] uint64"""
        in_limbs = in_limbs or self.zt.N
        out_limbs = out_limbs or self.zt.K
        r = as_limbs(row, in_limbs)
        assert r.shape[0] == self._row_len, "Row length must match the code's row length"  # code.rs:191-195
        ctx = ctx or default_context()
        out = np.empty((self._codeword_len, out_limbs), dtype=np.uint64)
        nat.check(nat.lib().zipgpu_encode_rows(self.native(ctx, in_limbs, out_limbs), 1, nat.ptr(r), nat.ptr(out)))
        return out

    def encode(self, row, ctx: Context | None = None) -> np.ndarray:
        """code.rs:35-37: encode == encode_wide::<N, M>"""
        return self.encode_wide(row, self.zt.N, self.zt.M, ctx)


@dataclass
class MultilinearZipParams:
    """pcs/structs.rs:11-29"""

    num_vars: int
    num_rows: int
    linear_code: RaaCode

    @staticmethod
    def new(num_vars: int, num_rows: int, linear_code: RaaCode) -> "MultilinearZipParams":
        return MultilinearZipParams(num_vars, num_rows, linear_code)


@dataclass
class MerkleTree:
    """pcs/utils.rs:66-71"""

    root: bytes
    depth: int
    layers: np.ndarray  # uint8 [(2 << depth) - 2, 32]

    @staticmethod
    def new(depth: int, leaves, leaf_limbs: int | None = None, ctx: Context | None = None) -> "MerkleTree":
        """pcs/utils.rs:74-85"""
        a = np.asarray(leaves)
        if leaf_limbs is None:
            leaf_limbs = a.shape[-1] if a.ndim == 2 else 1
        lv = as_limbs(a, leaf_limbs)
        n = lv.shape[0]
        assert n != 0 and n & (n - 1) == 0, "assertion failed: leaves.len().is_power_of_two()"  # utils.rs:75
        assert n == 1 << depth, f"assertion `left == right` failed\n  left: {n}\n right: {1 << depth}"  # utils.rs:76
        ctx = ctx or default_context()
        layers = np.empty(((2 << depth) - 2, 32), dtype=np.uint8)
        root = np.empty(32, dtype=np.uint8)
        nat.check(nat.lib().zipgpu_merkle_rows(ctx.handle, 1, depth, leaf_limbs, nat.ptr(lv),
                                               nat.ptr(layers) if layers.size else None, nat.ptr(root)))
        return MerkleTree(root.tobytes(), depth, layers)


@dataclass
class MerkleProof:
    """pcs/utils.rs:131-176"""

    merkle_path: list[bytes] = field(default_factory=list)

    @staticmethod
    def create_proof(merkle_tree: MerkleTree, leaf: int) -> "MerkleProof":
        """pcs/utils.rs:163-176 (index arithmetic over the layers the GPU produced)"""
        offset, path = 0, []
        for depth in range(merkle_tree.depth, 0, -1):
            width = 1 << depth
            idx = (leaf >> (merkle_tree.depth - depth)) ^ 1
            path.append(merkle_tree.layers[offset + idx].tobytes())
            offset += width
        return MerkleProof(path)


@dataclass
class MultilinearZipData:
    """pcs/structs.rs:31-65"""

    rows: np.ndarray  # uint64 [num_rows * cw, K]
    rows_merkle_trees: list[MerkleTree]

    def roots(self) -> list[bytes]:
        return [t.root for t in self.rows_merkle_trees]

    def root_at_index(self, index: int) -> bytes:
        return self.rows_merkle_trees[index].root


@dataclass
class MultilinearZipCommitment:
    """pcs/structs.rs:40-45"""

    roots: list[bytes]


class ResidentZipData:
    """Device-resident MultilinearZipData (rows + layers stay in HBM for `open`); see zipgpu_commit_resident.
    With a MultiContext the data stays sharded by row range over the GPUs that produced it (zipgpu_mgpu_data); every
    accessor returns results in row order."""

    def __init__(self, handle: C.c_void_p, num_rows: int, cw: int, out_limbs: int, depth: int, row_len: int,
                 ctx: "Context | None" = None):
        self.handle, self.num_rows, self.cw, self.out_limbs, self.depth = handle, num_rows, cw, out_limbs, depth
        self.row_len = row_len
        self._ctx = ctx
        self._multi = bool(ctx is not None and ctx.multi)
        self._uid = _next_uid()
        if ctx is not None:
            ctx._live[self._uid] = weakref.ref(self)

    def _h(self):
        if not self.handle:
            raise ZipGpuClosed("this prover data has been freed (or its context closed)")
        return self.handle

    def shards(self):
        """[(zipgpu_data* or None, row_begin, row_count)] -- one entry for a single-GPU handle"""
        if not self._multi:
            return [(self._h(), 0, self.num_rows)]
        out = []
        for g in range(self._ctx.num_devices):
            b, n = C.c_size_t(), C.c_size_t()
            h = nat.lib().zipgpu_mgpu_data_shard(self._h(), g, C.byref(b), C.byref(n))
            out.append((C.c_void_p(h) if h else None, b.value, n.value))
        return out

    def _read(self, fn, per_row_shape, dtype, row_begin, row_count):
        row_count = self.num_rows - row_begin if row_count is None else row_count
        out = np.empty((row_count,) + per_row_shape, dtype=dtype)
        for h, b, n in self.shards():
            lo, hi = max(b, row_begin), min(b + n, row_begin + row_count)
            if h is None or lo >= hi:
                continue
            nat.check(fn(h, lo - b, hi - lo, nat.ptr(out[lo - row_begin:hi - row_begin])))
        return out

    def rows(self, row_begin: int = 0, row_count: int | None = None) -> np.ndarray:
        out = self._read(nat.lib().zipgpu_data_read_rows, (self.cw, self.out_limbs), np.uint64, row_begin, row_count)
        return out.reshape(-1, self.out_limbs)

    def layers(self, row_begin: int = 0, row_count: int | None = None) -> np.ndarray:
        return self._read(nat.lib().zipgpu_data_read_layers, ((2 << self.depth) - 2, 32), np.uint8, row_begin, row_count)

    def open_columns(self, columns) -> tuple[np.ndarray, np.ndarray]:
        """open_z.rs:124-143: -> (values [ncols, num_rows, K], paths [ncols, num_rows, depth, 32])"""
        cols = np.ascontiguousarray(columns, dtype=np.uint32)
        vals = np.empty((cols.size, self.num_rows, self.out_limbs), dtype=np.uint64)
        paths = np.empty((cols.size, self.num_rows, self.depth, 32), dtype=np.uint8)
        fn = nat.lib().zipgpu_mgpu_data_open_columns if self._multi else nat.lib().zipgpu_data_open_columns
        nat.check(fn(self._h(), cols.size, nat.ptr(cols), nat.ptr(vals), nat.ptr(paths) if paths.size else None))
        return vals, paths

    def open_columns_wire(self, columns) -> bytes:
        """open_z.rs:124-143 as proof-stream bytes (PcsTranscript::write_integers + write_merkle_proof,
        pcs_transcript.rs:115-135,198-211): what `open` appends to the transcript stream for these columns"""
        return self.open_columns_wire_view(columns).tobytes()

    def open_columns_wire_view(self, columns) -> np.ndarray:
        """the same bytes as a uint8 view of the context's pinned proof-stream buffer (no page faults, no staged copy, no
        extra host copy: the D2H runs at PCIe rate); valid until the next call on this context"""
        cols = np.ascontiguousarray(columns, dtype=np.uint32)
        L = nat.lib()
        per = int((L.zipgpu_mgpu_data_open_columns_wire_bytes if self._multi else L.zipgpu_data_open_columns_wire_bytes)(self._h()))
        out = self._ctx.proof_stream_buffer(cols.size * per)
        fn = L.zipgpu_mgpu_data_open_columns_wire if self._multi else L.zipgpu_data_open_columns_wire
        if cols.size:
            nat.check(fn(self._h(), cols.size, nat.ptr(cols), out.ctypes.data_as(C.c_void_p)))
        return out

    def combine_rows(self, coeffs, out_limbs: int = 8) -> np.ndarray:
        """open_z.rs:100-113 / zip/utils.rs:94-127: u' = sum_i coeffs[i] * row_i over the unencoded evaluations,
        operands expanded N -> M; -> [row_len, out_limbs] uint64"""
        c = as_limbs(coeffs, 1).reshape(-1)
        assert c.size == self.num_rows, "one coefficient per row"
        row_len = self.row_len
        out = np.empty((row_len, out_limbs), dtype=np.uint64)
        fn = nat.lib().zipgpu_mgpu_data_combine_rows if self._multi else nat.lib().zipgpu_data_combine_rows
        nat.check(fn(self._h(), nat.ptr(c), out_limbs, nat.ptr(out)))
        return out

    def _release(self) -> None:
        if self.handle:
            (nat.lib().zipgpu_mgpu_data_free if self._multi else nat.lib().zipgpu_data_free)(self.handle)
            self.handle = None

    def free(self) -> None:
        self._release()
        if self._ctx is not None:
            self._ctx._live.pop(self._uid, None)

    def __del__(self):
        try:
            if self._ctx is None or self._ctx.handle:
                self.free()
        except Exception:
            pass


def _validate_input(function: str, param_num_vars: int, polys) -> None:
    """pcs/utils.rs:24-58 (commit passes no points)"""
    for poly in polys:
        if param_num_vars < poly.num_vars:
            raise InvalidPcsParam(
                f"Too many variates of poly to {function} (param supports variates up to {param_num_vars} "
                f"but got {poly.num_vars})")


class MultilinearZip:
    """pcs/structs.rs:9, commit.rs:18-184"""

    @staticmethod
    def setup(poly_size: int, linear_code: RaaCode) -> MultilinearZipParams:
        """pcs/structs.rs:79-90"""
        assert poly_size != 0 and poly_size & (poly_size - 1) == 0, "assertion failed: poly_size.is_power_of_two()"
        num_vars = poly_size.bit_length() - 1
        num_rows = int(nat.lib().zipgpu_num_rows(poly_size, linear_code.row_len()))
        return MultilinearZipParams(num_vars, num_rows, linear_code)

    @staticmethod
    def encode_rows(pp: MultilinearZipParams, codeword_len: int, row_len: int, evals,
                    ctx: Context | None = None) -> np.ndarray:
        """commit.rs:158-183 -> [num_rows * codeword_len, K] uint64"""
        lc = pp.linear_code
        zt = lc.zt
        ev = as_limbs(evals, zt.N)
        assert codeword_len == lc.codeword_len() and row_len == lc.row_len()
        # chunks_exact semantics of commit.rs:168-170 need at least num_rows * row_len evaluations
        assert ev.shape[0] >= pp.num_rows * row_len, "not enough evaluations for num_rows rows"
        ctx = ctx or default_context()
        out = np.empty((pp.num_rows * codeword_len, zt.K), dtype=np.uint64)
        enc = nat.lib().zipgpu_mgpu_encode_rows if ctx.multi else nat.lib().zipgpu_encode_rows
        nat.check(enc(lc.native(ctx, zt.N, zt.K), pp.num_rows, nat.ptr(ev), nat.ptr(out)))
        return out

    @staticmethod
    def commit(pp: MultilinearZipParams, poly: DenseMultilinearExtension, ctx: Context | None = None
               ) -> tuple[MultilinearZipData, MultilinearZipCommitment]:
        """commit.rs:50-87"""
        _validate_input("commit", pp.num_vars, [poly])
        lc = pp.linear_code
        zt = lc.zt
        expected_num_evals = pp.num_rows * lc.row_len()
        n = poly.evaluations.shape[0]
        assert n == expected_num_evals, (  # commit.rs:56-63
            f"Polynomial has an incorrect number of evaluations ({n}) for the expected matrix size "
            f"({expected_num_evals})")
        cw = lc.codeword_len()
        assert cw & (cw - 1) == 0, "assertion failed: leaves.len().is_power_of_two()"  # utils.rs:75 via commit.rs:73
        depth = cw.bit_length() - 1  # commit.rs:67
        ctx = ctx or default_context()
        ev = as_limbs(poly.evaluations, zt.N)
        rows = np.empty((pp.num_rows * cw, zt.K), dtype=np.uint64)
        layers = np.empty((pp.num_rows, (2 << depth) - 2, 32), dtype=np.uint8)
        roots = np.empty((pp.num_rows, 32), dtype=np.uint8)
        com = nat.lib().zipgpu_mgpu_commit if ctx.multi else nat.lib().zipgpu_commit
        nat.check(com(lc.native(ctx, zt.N, zt.K), pp.num_rows, nat.ptr(ev), nat.ptr(rows),
                      nat.ptr(layers) if layers.size else None, nat.ptr(roots)))
        trees = [MerkleTree(roots[i].tobytes(), depth, layers[i]) for i in range(pp.num_rows)]
        assert len(trees) == pp.num_rows  # commit.rs:76
        return MultilinearZipData(rows, trees), MultilinearZipCommitment([t.root for t in trees])

    @staticmethod
    def commit_no_merkle(pp: MultilinearZipParams, poly: DenseMultilinearExtension, ctx: Context | None = None
                         ) -> tuple[MultilinearZipData, MultilinearZipCommitment]:
        """commit.rs:104-119"""
        _validate_input("commit", pp.num_vars, [poly])
        lc = pp.linear_code
        rows = MultilinearZip.encode_rows(pp, lc.codeword_len(), lc.row_len(), poly.evaluations, ctx)
        return MultilinearZipData(rows, []), MultilinearZipCommitment([])

    @staticmethod
    def batch_commit(pp: MultilinearZipParams, polys: list[DenseMultilinearExtension], ctx: Context | None = None
                     ) -> list[tuple[MultilinearZipData, MultilinearZipCommitment]]:
        """commit.rs:134-142, all polynomials in one pipelined GPU submission"""
        if not polys:
            return []
        _validate_input("commit", pp.num_vars, polys)
        lc = pp.linear_code
        zt = lc.zt
        cw = lc.codeword_len()
        expected = pp.num_rows * lc.row_len()
        for poly in polys:
            n = poly.evaluations.shape[0]
            assert n == expected, (
                f"Polynomial has an incorrect number of evaluations ({n}) for the expected matrix size ({expected})")
        assert cw & (cw - 1) == 0, "assertion failed: leaves.len().is_power_of_two()"
        depth = cw.bit_length() - 1
        ctx = ctx or default_context()
        k = len(polys)
        evs = [as_limbs(p.evaluations, zt.N) for p in polys]
        rows = [np.empty((pp.num_rows * cw, zt.K), dtype=np.uint64) for _ in range(k)]
        layers = [np.empty((pp.num_rows, (2 << depth) - 2, 32), dtype=np.uint8) for _ in range(k)]
        roots = [np.empty((pp.num_rows, 32), dtype=np.uint8) for _ in range(k)]
        arr = lambda xs: (C.c_void_p * k)(*[nat.ptr(x) for x in xs])
        bc = nat.lib().zipgpu_mgpu_batch_commit if ctx.multi else nat.lib().zipgpu_batch_commit
        nat.check(bc(lc.native(ctx, zt.N, zt.K), k, pp.num_rows, arr(evs), arr(rows),
                     arr(layers) if layers[0].size else None, arr(roots)))
        out = []
        for p in range(k):
            trees = [MerkleTree(roots[p][i].tobytes(), depth, layers[p][i]) for i in range(pp.num_rows)]
            out.append((MultilinearZipData(rows[p], trees), MultilinearZipCommitment([t.root for t in trees])))
        return out

    @staticmethod
    def commit_resident(pp: MultilinearZipParams, poly: DenseMultilinearExtension, ctx: Context | None = None
                        ) -> tuple[ResidentZipData, MultilinearZipCommitment]:
        """commit with the prover data (rows, layers) left on the GPU behind a handle (extension)"""
        _validate_input("commit", pp.num_vars, [poly])
        lc = pp.linear_code
        zt = lc.zt
        expected = pp.num_rows * lc.row_len()
        n = poly.evaluations.shape[0]
        assert n == expected, (
            f"Polynomial has an incorrect number of evaluations ({n}) for the expected matrix size ({expected})")
        cw = lc.codeword_len()
        assert cw & (cw - 1) == 0, "assertion failed: leaves.len().is_power_of_two()"
        depth = cw.bit_length() - 1
        ctx = ctx or default_context()
        ev = as_limbs(poly.evaluations, zt.N)
        roots = np.empty((pp.num_rows, 32), dtype=np.uint8)
        h = C.c_void_p()
        cr = nat.lib().zipgpu_mgpu_commit_resident if ctx.multi else nat.lib().zipgpu_commit_resident
        nat.check(cr(lc.native(ctx, zt.N, zt.K), pp.num_rows, nat.ptr(ev), nat.ptr(roots), C.byref(h)))
        return (ResidentZipData(h, pp.num_rows, cw, zt.K, depth, lc.row_len(), ctx),
                MultilinearZipCommitment([roots[i].tobytes() for i in range(pp.num_rows)]))
