"""oracle/cbind.py -- TEST INFRASTRUCTURE ONLY: ctypes bindings of oracle/libzip_oracle.so (zip_oracle.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libzip_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "zip_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libzip_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        u64p, u32p, u8p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
        sz, i, u64 = C.c_size_t, C.c_int, C.c_uint64
        L.zo_shuffle_seeded.argtypes = [C.c_void_p, sz, sz, u64]
        L.zo_perm_from_seed.argtypes = [u32p, sz, u64]
        L.zo_stdrng_words.argtypes = [u64, u32p, sz]
        L.zo_chacha_block.argtypes = [u32p, u64, u64, i, u32p]
        L.zo_seed_from_u64.argtypes = [u64, u32p]
        L.zo_encode_row_seeded.argtypes = [u64p, sz, i, sz, u64, u64, u64p, i]
        L.zo_encode_rows_perm.argtypes = [u64p, sz, sz, i, sz, u32p, u32p, u64p, i]
        L.zo_blake3.argtypes = [u8p, sz, u8p]
        L.zo_blake3_force_portable.argtypes = [i]
        L.zo_merkle_tree_new.argtypes = [sz, u64p, sz, i, u8p, u8p]
        L.zo_merkle_create_proof.argtypes = [sz, u8p, sz, u8p]
        L.zo_merkle_verify.argtypes = [sz, u8p, u8p, u64p, i, sz]
        L.zo_commit_perm.argtypes = [u64p, sz, sz, i, sz, u32p, u32p, i, u64p, u8p, u8p]
        L.zo_commit_mt.argtypes = [u64p, sz, sz, i, sz, u64, u64, u32p, u32p, i, u64p, u8p, u8p, i, i]
        i64p = C.POINTER(C.c_int64)
        L.zo_sparse_mat_vec.argtypes = [sz, sz, sz, u32p, i64p, u64p, i, u64p, i]
        L.zo_sparse_commit_mt.argtypes = [u64p, sz, sz, i, sz, sz, u32p, i64p, u32p, i64p, i, u64p, u8p, u8p, i]
        L.zo_raa_row_len.argtypes = [sz]; L.zo_raa_row_len.restype = sz
        L.zo_num_rows.argtypes = [sz, sz]; L.zo_num_rows.restype = sz
        L.zo_raa_width_ok.argtypes = [i, i, sz, sz]
        L.zo_accumulate.argtypes = [u64p, sz, i]
        L.zo_repeat.argtypes = [u64p, sz, i, sz, u64p, i]
        L.zo_widen.argtypes = [u64p, i, u64p, i]
        L.zo_int_to_bytes.argtypes = [u64p, i, u8p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def perm_from_seed(n: int, seed: int) -> np.ndarray:
    idx = np.empty(n, dtype=np.uint32)
    lib().zo_perm_from_seed(_p(idx, C.c_uint32), n, seed)
    return idx


def stdrng_words(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint32)
    lib().zo_stdrng_words(seed, _p(out, C.c_uint32), n)
    return out


def blake3(data: bytes) -> bytes:
    buf = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, dtype=np.uint8)
    out = np.empty(32, dtype=np.uint8)
    lib().zo_blake3(_p(np.ascontiguousarray(buf), C.c_uint8), len(data), _p(out, C.c_uint8))
    return out.tobytes()


def encode_rows(evals: np.ndarray, num_rows: int, row_len: int, rep: int, perm1, perm2,
                in_limbs: int = 1, out_limbs: int = 4):
    """evals: uint64[num_rows*row_len*in_limbs] -> (rc, uint64[num_rows*cw*out_limbs])"""
    evals = np.ascontiguousarray(evals, dtype=np.uint64)
    out = np.empty(num_rows * row_len * rep * out_limbs, dtype=np.uint64)
    rc = lib().zo_encode_rows_perm(_p(evals, C.c_uint64), num_rows, row_len, in_limbs, rep,
                                   _p(perm1, C.c_uint32), _p(perm2, C.c_uint32), _p(out, C.c_uint64), out_limbs)
    return rc, out


def encode_row_seeded(row: np.ndarray, rep: int, seed1: int, seed2: int, in_limbs: int = 1, out_limbs: int = 4):
    row = np.ascontiguousarray(row, dtype=np.uint64)
    row_len = row.size // in_limbs
    out = np.empty(row_len * rep * out_limbs, dtype=np.uint64)
    rc = lib().zo_encode_row_seeded(_p(row, C.c_uint64), row_len, in_limbs, rep, seed1, seed2,
                                    _p(out, C.c_uint64), out_limbs)
    return rc, out


def merkle_tree(depth: int, leaves: np.ndarray, leaf_limbs: int = 4):
    """-> (rc, layers uint8[((2<<depth)-2)*32], root uint8[32])"""
    leaves = np.ascontiguousarray(leaves, dtype=np.uint64)
    n = leaves.size // leaf_limbs
    layers = np.empty(max(((2 << depth) - 2), 0) * 32, dtype=np.uint8)
    root = np.empty(32, dtype=np.uint8)
    rc = lib().zo_merkle_tree_new(depth, _p(leaves, C.c_uint64), n, leaf_limbs,
                                  _p(layers, C.c_uint8) if layers.size else None, _p(root, C.c_uint8))
    return rc, layers, root


def commit(evals: np.ndarray, num_rows: int, row_len: int, rep: int, perm1, perm2,
           in_limbs: int = 1, out_limbs: int = 4, want_layers: bool = True):
    """-> (rc, rows uint64, layers uint8 or None, roots uint8[num_rows*32])"""
    evals = np.ascontiguousarray(evals, dtype=np.uint64)
    cw = row_len * rep
    depth = (cw - 1).bit_length() if cw > 1 else 0
    rows = np.empty(num_rows * cw * out_limbs, dtype=np.uint64)
    layers = np.empty(num_rows * ((2 << depth) - 2) * 32, dtype=np.uint8) if want_layers else None
    roots = np.empty(num_rows * 32, dtype=np.uint8)
    rc = lib().zo_commit_perm(_p(evals, C.c_uint64), num_rows, row_len, in_limbs, rep,
                              _p(perm1, C.c_uint32), _p(perm2, C.c_uint32), out_limbs,
                              _p(rows, C.c_uint64), _p(layers, C.c_uint8) if want_layers and layers.size else None,
                              _p(roots, C.c_uint8))
    return rc, rows, layers, roots


def commit_mt(evals: np.ndarray, num_rows: int, row_len: int, rep: int, seed1: int, seed2: int, perm1, perm2,
              threads: int, faithful: bool, in_limbs: int = 1, out_limbs: int = 4,
              want_rows: bool = True, want_layers: bool = True):
    evals = np.ascontiguousarray(evals, dtype=np.uint64)
    cw = row_len * rep
    depth = (cw - 1).bit_length() if cw > 1 else 0
    rows = np.empty(num_rows * cw * out_limbs, dtype=np.uint64) if want_rows else None
    layers = np.empty(num_rows * ((2 << depth) - 2) * 32, dtype=np.uint8) if want_layers else None
    roots = np.empty(num_rows * 32, dtype=np.uint8)
    rc = lib().zo_commit_mt(_p(evals, C.c_uint64), num_rows, row_len, in_limbs, rep, seed1, seed2,
                            _p(perm1, C.c_uint32), _p(perm2, C.c_uint32), out_limbs,
                            _p(rows, C.c_uint64), _p(layers, C.c_uint8), _p(roots, C.c_uint8), threads, int(faithful))
    return rc, rows, layers, roots


def sparse_commit(evals: np.ndarray, num_rows: int, row_len: int, n: int, d: int, cols_a, coef_a, cols_b, coef_b,
                  in_limbs: int = 1, out_limbs: int = 4, want_layers: bool = True, want_roots: bool = True,
                  threads: int = 1):
    """ZipLinearCode as the code (zip/code.rs:77-215): -> (rc, rows, layers or None, roots or None)"""
    evals = np.ascontiguousarray(evals, dtype=np.uint64)
    cols_a, cols_b = (np.ascontiguousarray(c, dtype=np.uint32) for c in (cols_a, cols_b))
    coef_a, coef_b = (np.ascontiguousarray(c, dtype=np.int64) for c in (coef_a, coef_b))
    cw = 2 * n
    depth = (cw - 1).bit_length() if cw > 1 else 0
    rows = np.empty(num_rows * cw * out_limbs, dtype=np.uint64)
    layers = np.empty(num_rows * ((2 << depth) - 2) * 32, dtype=np.uint8) if want_layers and want_roots else None
    roots = np.empty(num_rows * 32, dtype=np.uint8) if want_roots else None
    rc = lib().zo_sparse_commit_mt(_p(evals, C.c_uint64), num_rows, row_len, in_limbs, n, d,
                                   _p(cols_a, C.c_uint32), _p(coef_a, C.c_int64), _p(cols_b, C.c_uint32),
                                   _p(coef_b, C.c_int64), out_limbs, _p(rows, C.c_uint64),
                                   _p(layers, C.c_uint8) if layers is not None and layers.size else None,
                                   _p(roots, C.c_uint8), threads)
    return rc, rows, layers, roots
