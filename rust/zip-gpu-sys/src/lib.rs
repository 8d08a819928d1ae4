// UNCOMPILED (no Rust toolchain in this image) -- see ../../README.md
//
// extern "C" surface of include/zipgpu.h, one declaration per symbol the commit path binds, plus a small safe
// wrapper (Ctx / Code / Data with Drop).  Integers cross the boundary as the in-memory layout of `[Int<n>]`:
// n little-endian u64 limbs, least significant first (zinc: src/field/int.rs:23-25,230-232).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)] pub struct zipgpu_ctx { _p: [u8; 0] }
#[repr(C)] pub struct zipgpu_peer_roots { _private: [u8; 0] }
#[repr(C)]
pub struct zipgpu_code { _p: [u8; 0] }
#[repr(C)] pub struct zipgpu_data { _p: [u8; 0] }

pub const ZIPGPU_OK: c_int = 0;
pub const ZIPGPU_ERR_INVALID: c_int = -1;
pub const ZIPGPU_ERR_CUDA: c_int = -2;
pub const ZIPGPU_ERR_NOMEM: c_int = -3;
pub const ZIPGPU_ERR_UNSUPPORTED: c_int = -4;
pub const ZIPGPU_ERR_NO_DEVICE: c_int = -5;
pub const ZIPGPU_ERR_WIDTH: c_int = -6;

unsafe extern "C" {
    pub fn zipgpu_last_error() -> *const c_char;
    pub fn zipgpu_device_count(count: *mut c_int) -> c_int;
    pub fn zipgpu_ctx_create(device: c_int, out: *mut *mut zipgpu_ctx) -> c_int;
    pub fn zipgpu_ctx_destroy(ctx: *mut zipgpu_ctx);
    pub fn zipgpu_host_register(p: *mut c_void, bytes: usize) -> c_int;
    pub fn zipgpu_host_unregister(p: *mut c_void) -> c_int;
    // RaaCode (code_raa.rs:35-86): per-pp state, permutations uploaded once
    pub fn zipgpu_code_create(ctx: *mut zipgpu_ctx, row_len: usize, repetition_factor: usize, in_limbs: c_int,
                              out_limbs: c_int, perm1: *const u32, perm2: *const u32,
                              out: *mut *mut zipgpu_code) -> c_int;
    // ZipLinearCode (code.rs:100-147): cells of the two sampled sparse matrices in, same handle type out
    pub fn zipgpu_sparse_code_create(ctx: *mut zipgpu_ctx, row_len: usize, codeword_len: usize, cells_per_row: usize,
                                     in_limbs: c_int, out_limbs: c_int, cols_a: *const u32, coef_a: *const i64,
                                     cols_b: *const u32, coef_b: *const i64, out: *mut *mut zipgpu_code) -> c_int;
    pub fn zipgpu_code_sparse_kind(code: *const zipgpu_code) -> c_int;
    // multi-GPU: all-gather of the row roots over NVLink peer memory (one process per GPU, one node)
    pub fn zipgpu_peer_roots_create(ctx: *mut zipgpu_ctx, total_rows: usize, rank: c_int, world: c_int,
                                    out: *mut *mut zipgpu_peer_roots, ipc_out: *mut u8) -> c_int;
    pub fn zipgpu_peer_roots_connect(pr: *mut zipgpu_peer_roots, ipc_all: *const u8) -> c_int;
    pub fn zipgpu_peer_roots_allgather(pr: *mut zipgpu_peer_roots, row_begin: usize, count: usize,
                                       d_local_roots: *const u8, stream: *mut c_void, d_all_out: *mut *mut u8) -> c_int;
    pub fn zipgpu_peer_roots_destroy(pr: *mut zipgpu_peer_roots);
    pub fn zipgpu_code_destroy(code: *mut zipgpu_code);
    // encode_rows / commit_no_merkle (commit.rs:104-119,158-183)
    pub fn zipgpu_encode_rows(code: *mut zipgpu_code, num_rows: usize, evals: *const u64, rows_out: *mut u64) -> c_int;
    // MerkleTree::new batched (pcs/utils.rs:74-118)
    pub fn zipgpu_merkle_rows(ctx: *mut zipgpu_ctx, num_rows: usize, depth: c_int, leaf_limbs: c_int,
                              leaves: *const u64, layers_out: *mut u8, roots_out: *mut u8) -> c_int;
    // commit (commit.rs:50-87)
    pub fn zipgpu_commit(code: *mut zipgpu_code, num_rows: usize, evals: *const u64, rows_out: *mut u64,
                         layers_out: *mut u8, roots_out: *mut u8) -> c_int;
    // batch_commit (commit.rs:134-142)
    pub fn zipgpu_batch_commit(code: *mut zipgpu_code, num_polys: usize, num_rows: usize,
                               evals: *const *const u64, rows_out: *const *mut u64, layers_out: *const *mut u8,
                               roots_out: *const *mut u8) -> c_int;
    // device-resident MultilinearZipData for `open` (structs.rs:33-38; open_z.rs:124-143)
    pub fn zipgpu_commit_resident(code: *mut zipgpu_code, num_rows: usize, evals: *const u64, roots_out: *mut u8,
                                  handle: *mut *mut zipgpu_data) -> c_int;
    pub fn zipgpu_data_free(data: *mut zipgpu_data);
    pub fn zipgpu_data_read_rows(data: *const zipgpu_data, row_begin: usize, row_count: usize, rows_out: *mut u64) -> c_int;
    pub fn zipgpu_data_read_layers(data: *const zipgpu_data, row_begin: usize, row_count: usize, layers_out: *mut u8) -> c_int;
    pub fn zipgpu_data_open_columns(data: *const zipgpu_data, num_cols: usize, columns: *const u32,
                                    col_values_out: *mut u64, paths_out: *mut u8) -> c_int;
}

/// `rc != 0` -> the library's thread-local message (never unwinds across the FFI)
pub fn check(rc: c_int) -> Result<(), String> {
    if rc == ZIPGPU_OK { return Ok(()); }
    let msg = unsafe { core::ffi::CStr::from_ptr(zipgpu_last_error()) }.to_string_lossy().into_owned();
    Err(format!("zipgpu error {rc}: {msg}"))
}

pub struct Ctx(pub *mut zipgpu_ctx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {} // the library is thread-safe per context
impl Ctx {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut p = core::ptr::null_mut();
        check(unsafe { zipgpu_ctx_create(device, &mut p) })?;
        Ok(Ctx(p))
    }
}
impl Drop for Ctx { fn drop(&mut self) { unsafe { zipgpu_ctx_destroy(self.0) } } }

pub struct Code(pub *mut zipgpu_code);
unsafe impl Send for Code {}
unsafe impl Sync for Code {}
impl Code {
    /// `perm1`/`perm2`: `shuffle_seeded` (zip/utils.rs:139-142) applied to `0..codeword_len` with the two seeds
    pub fn new(ctx: &Ctx, row_len: usize, rep: usize, in_limbs: usize, out_limbs: usize, perm1: &[u32], perm2: &[u32])
               -> Result<Self, String> {
        assert_eq!(perm1.len(), row_len * rep);
        assert_eq!(perm2.len(), row_len * rep);
        let mut p = core::ptr::null_mut();
        check(unsafe { zipgpu_code_create(ctx.0, row_len, rep, in_limbs as c_int, out_limbs as c_int,
                                          perm1.as_ptr(), perm2.as_ptr(), &mut p) })?;
        Ok(Code(p))
    }
}
impl Drop for Code { fn drop(&mut self) { unsafe { zipgpu_code_destroy(self.0) } } }
