"""zinc_b200/_native.py -- ctypes binding of libzipgpu.so (include/zipgpu.h).

There is no CPU fallback: if the library is missing or no sm_100 GPU is visible, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZIPGPU_LIB") or os.path.join(_HERE, "libzipgpu.so")  # ZIPGPU_LIB: A/B builds (build.py --variant)

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_WIDTH, ERR_PEER_TIMEOUT = 0, -1, -2, -3, -4, -5, -6, -7

u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
vp, sz, i32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64

# name -> (restype, argtypes): every symbol include/zipgpu.h declares
SIGNATURES = {
    "zipgpu_version": (C.c_char_p, []),
    "zipgpu_last_error": (C.c_char_p, []),
    "zipgpu_device_count": (i32, [C.POINTER(i32)]),
    "zipgpu_ctx_create": (i32, [i32, C.POINTER(vp)]),
    "zipgpu_ctx_destroy": (None, [vp]),
    "zipgpu_ctx_device": (i32, [vp]),
    "zipgpu_ctx_sync": (i32, [vp]),
    "zipgpu_ctx_launch_count": (u64, [vp]),
    "zipgpu_host_alloc": (i32, [sz, C.POINTER(vp)]),
    "zipgpu_host_free": (i32, [vp]),
    "zipgpu_host_register": (i32, [vp, sz]),
    "zipgpu_host_unregister": (i32, [vp]),
    "zipgpu_perm_from_seed": (i32, [u64, C.c_uint32, vp]),
    "zipgpu_chacha_block": (i32, [vp, u64, i32, vp]),
    "zipgpu_raa_row_len": (sz, [sz]),
    "zipgpu_num_rows": (sz, [sz, sz]),
    "zipgpu_raa_codeword_width_bits": (i32, [i32, sz, sz]),
    "zipgpu_code_create": (i32, [vp, sz, sz, i32, i32, vp, vp, C.POINTER(vp)]),
    "zipgpu_sparse_code_create": (i32, [vp, sz, sz, sz, i32, i32, vp, vp, vp, vp, C.POINTER(vp)]),
    "zipgpu_code_sparse_kind": (i32, [vp]),
    "zipgpu_code_destroy": (None, [vp]),
    "zipgpu_peer_roots_create": (i32, [vp, sz, i32, i32, C.POINTER(vp), vp]),
    "zipgpu_peer_roots_connect": (i32, [vp, vp]),
    "zipgpu_peer_roots_allgather": (i32, [vp, sz, sz, vp, vp, C.POINTER(vp)]),
    "zipgpu_peer_roots_destroy": (None, [vp]),
    "zipgpu_peer_roots_connect_local": (i32, [C.POINTER(vp), i32]),
    "zipgpu_peer_roots_status": (i32, [vp]),
    "zipgpu_commit_device_sharded": (i32, [vp, vp, sz, sz, vp, vp, vp, vp, C.POINTER(vp)]),
    "zipgpu_commit_resident_sharded": (i32, [vp, vp, sz, sz, vp, vp, C.POINTER(vp)]),
    "zipgpu_mgpu_create": (i32, [C.POINTER(i32), i32, C.POINTER(vp)]),
    "zipgpu_mgpu_destroy": (None, [vp]),
    "zipgpu_mgpu_num_devices": (i32, [vp]),
    "zipgpu_mgpu_ctx": (vp, [vp, i32]),
    "zipgpu_mgpu_launch_count": (u64, [vp]),
    "zipgpu_mgpu_code_create": (i32, [vp, sz, sz, i32, i32, vp, vp, C.POINTER(vp)]),
    "zipgpu_mgpu_code_destroy": (None, [vp]),
    "zipgpu_mgpu_code_device": (vp, [vp, i32]),
    "zipgpu_mgpu_encode_rows": (i32, [vp, sz, vp, vp]),
    "zipgpu_mgpu_commit": (i32, [vp, sz, vp, vp, vp, vp]),
    "zipgpu_mgpu_batch_commit": (i32, [vp, sz, sz, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    "zipgpu_mgpu_commit_resident": (i32, [vp, sz, vp, vp, C.POINTER(vp)]),
    "zipgpu_mgpu_data_free": (None, [vp]),
    "zipgpu_mgpu_data_num_rows": (sz, [vp]),
    "zipgpu_mgpu_data_shard": (vp, [vp, i32, C.POINTER(sz), C.POINTER(sz)]),
    "zipgpu_mgpu_data_roots_device": (vp, [vp, i32]),
    "zipgpu_mgpu_data_open_columns": (i32, [vp, sz, vp, vp, vp]),
    "zipgpu_mgpu_data_open_columns_wire_bytes": (sz, [vp]),
    "zipgpu_mgpu_data_open_columns_wire": (i32, [vp, sz, vp, vp]),
    "zipgpu_mgpu_data_combine_rows": (i32, [vp, vp, i32, vp]),
    "zipgpu_data_all_roots_device": (vp, [vp]),
    "zipgpu_data_open_columns_strided": (i32, [vp, sz, vp, vp, vp, sz, sz]),
    "zipgpu_code_row_len": (sz, [vp]),
    "zipgpu_code_codeword_len": (sz, [vp]),
    "zipgpu_code_merkle_depth": (i32, [vp]),
    "zipgpu_encode_rows": (i32, [vp, sz, vp, vp]),
    "zipgpu_encode_rows_device": (i32, [vp, sz, vp, vp, vp]),
    "zipgpu_encode_f": (i32, [vp, sz, i32, vp, vp, vp]),
    "zipgpu_encode_wide": (i32, [vp, sz, i32, i32, vp, vp]),
    "zipgpu_merkle_rows": (i32, [vp, sz, i32, i32, vp, vp, vp]),
    "zipgpu_merkle_rows_device": (i32, [vp, sz, i32, i32, vp, vp, vp, vp]),
    "zipgpu_commit": (i32, [vp, sz, vp, vp, vp, vp]),
    "zipgpu_commit_device": (i32, [vp, sz, vp, vp, vp, vp, vp]),
    "zipgpu_batch_commit": (i32, [vp, sz, sz, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    "zipgpu_commit_resident": (i32, [vp, sz, vp, vp, C.POINTER(vp)]),
    "zipgpu_data_free": (None, [vp]),
    "zipgpu_data_num_rows": (sz, [vp]),
    "zipgpu_data_rows_device": (vp, [vp]),
    "zipgpu_data_layers_device": (vp, [vp]),
    "zipgpu_data_roots_device": (vp, [vp]),
    "zipgpu_data_read_rows": (i32, [vp, sz, sz, vp]),
    "zipgpu_data_read_layers": (i32, [vp, sz, sz, vp]),
    "zipgpu_data_open_columns": (i32, [vp, sz, vp, vp, vp]),
    "zipgpu_data_open_columns_wire_bytes": (sz, [vp]),
    "zipgpu_data_open_columns_wire": (i32, [vp, sz, vp, vp]),
    "zipgpu_data_combine_rows": (i32, [vp, vp, i32, vp]),
    "zipgpu_combine_rows_device": (i32, [vp, sz, sz, vp, vp, i32, vp, vp]),
    "zipgpu_profile_enable": (i32, [vp, i32]),
    "zipgpu_profile_read": (i32, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(u64), i32]),
    "zipgpu_microbench_int32": (i32, [vp, i32, i32, C.POINTER(C.c_double)]),
}


class ZipGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"zipgpu error {code}: {msg}")
        self.code = code
        self.message = msg


_lib = None


def lib() -> C.CDLL:
    """Load libzipgpu.so (built in-tree by zinc_b200.build / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -m zinc_b200.build` (nvcc, sm_100a). "
                "zinc_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise ZipGpuError(rc, lib().zipgpu_last_error().decode("utf-8", "replace"))


def ptr(a) -> int | None:
    """address of a numpy array / torch tensor / int / None"""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data
