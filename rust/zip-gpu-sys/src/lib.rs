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
#[repr(C)] pub struct zipgpu_mgpu { _p: [u8; 0] }
#[repr(C)] pub struct zipgpu_mgpu_code { _p: [u8; 0] }
#[repr(C)] pub struct zipgpu_mgpu_data { _p: [u8; 0] }

pub const ZIPGPU_OK: c_int = 0;
pub const ZIPGPU_ERR_INVALID: c_int = -1;
pub const ZIPGPU_ERR_CUDA: c_int = -2;
pub const ZIPGPU_ERR_NOMEM: c_int = -3;
pub const ZIPGPU_ERR_UNSUPPORTED: c_int = -4;
pub const ZIPGPU_ERR_NO_DEVICE: c_int = -5;
pub const ZIPGPU_ERR_WIDTH: c_int = -6;
pub const ZIPGPU_ERR_PEER_TIMEOUT: c_int = -7;

unsafe extern "C" {
    pub fn zipgpu_last_error() -> *const c_char;
    pub fn zipgpu_device_count(count: *mut c_int) -> c_int;
    pub fn zipgpu_ctx_create(device: c_int, out: *mut *mut zipgpu_ctx) -> c_int;
    pub fn zipgpu_ctx_destroy(ctx: *mut zipgpu_ctx);
    // pinned, huge-page-backed host memory for evaluations / proof streams (include/zipgpu.h)
    pub fn zipgpu_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn zipgpu_host_free(p: *mut c_void) -> c_int;
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
    // multi-GPU, one process (or thread) per GPU: the roots exchange fused into the roots-producing kernel
    pub fn zipgpu_peer_roots_create(ctx: *mut zipgpu_ctx, total_rows: usize, rank: c_int, world: c_int,
                                    out: *mut *mut zipgpu_peer_roots, ipc_out: *mut u8) -> c_int;
    pub fn zipgpu_peer_roots_connect(pr: *mut zipgpu_peer_roots, ipc_all: *const u8) -> c_int;
    pub fn zipgpu_peer_roots_connect_local(all: *const *mut zipgpu_peer_roots, n: c_int) -> c_int;
    pub fn zipgpu_peer_roots_allgather(pr: *mut zipgpu_peer_roots, row_begin: usize, count: usize,
                                       d_local_roots: *const u8, stream: *mut c_void, d_all_out: *mut *mut u8) -> c_int;
    pub fn zipgpu_peer_roots_status(pr: *mut zipgpu_peer_roots) -> c_int;
    pub fn zipgpu_peer_roots_destroy(pr: *mut zipgpu_peer_roots);
    pub fn zipgpu_commit_device_sharded(code: *mut zipgpu_code, pr: *mut zipgpu_peer_roots, row_begin: usize, count: usize,
                                        d_evals: *const u64, d_rows_out: *mut u64, d_layers_out: *mut u8,
                                        stream: *mut c_void, d_all_roots_out: *mut *mut u8) -> c_int;
    pub fn zipgpu_commit_resident_sharded(code: *mut zipgpu_code, pr: *mut zipgpu_peer_roots, row_begin: usize, count: usize,
                                          evals: *const u64, roots_all_out: *mut u8, handle: *mut *mut zipgpu_data) -> c_int;
    // multi-GPU, ONE process: what ZincProver needs (zinc/prover.rs:313-315) -- device list once, same calls as single-GPU
    pub fn zipgpu_mgpu_create(devices: *const c_int, n: c_int, out: *mut *mut zipgpu_mgpu) -> c_int;
    pub fn zipgpu_mgpu_destroy(m: *mut zipgpu_mgpu);
    pub fn zipgpu_mgpu_num_devices(m: *const zipgpu_mgpu) -> c_int;
    pub fn zipgpu_mgpu_code_create(m: *mut zipgpu_mgpu, row_len: usize, repetition_factor: usize, in_limbs: c_int,
                                   out_limbs: c_int, perm1: *const u32, perm2: *const u32,
                                   out: *mut *mut zipgpu_mgpu_code) -> c_int;
    pub fn zipgpu_mgpu_code_destroy(code: *mut zipgpu_mgpu_code);
    pub fn zipgpu_mgpu_encode_rows(code: *mut zipgpu_mgpu_code, num_rows: usize, evals: *const u64, rows_out: *mut u64) -> c_int;
    pub fn zipgpu_mgpu_commit(code: *mut zipgpu_mgpu_code, num_rows: usize, evals: *const u64, rows_out: *mut u64,
                              layers_out: *mut u8, roots_out: *mut u8) -> c_int;
    pub fn zipgpu_mgpu_batch_commit(code: *mut zipgpu_mgpu_code, num_polys: usize, num_rows: usize,
                                    evals: *const *const u64, rows_out: *const *mut u64, layers_out: *const *mut u8,
                                    roots_out: *const *mut u8) -> c_int;
    pub fn zipgpu_mgpu_commit_resident(code: *mut zipgpu_mgpu_code, num_rows: usize, evals: *const u64, roots_out: *mut u8,
                                       handle: *mut *mut zipgpu_mgpu_data) -> c_int;
    pub fn zipgpu_mgpu_data_free(data: *mut zipgpu_mgpu_data);
    pub fn zipgpu_mgpu_data_open_columns(data: *const zipgpu_mgpu_data, num_cols: usize, columns: *const u32,
                                         col_values_out: *mut u64, paths_out: *mut u8) -> c_int;
    pub fn zipgpu_mgpu_data_open_columns_wire_bytes(data: *const zipgpu_mgpu_data) -> usize;
    pub fn zipgpu_mgpu_data_open_columns_wire(data: *const zipgpu_mgpu_data, num_cols: usize, columns: *const u32,
                                              stream_out: *mut u8) -> c_int;
    pub fn zipgpu_mgpu_data_combine_rows(data: *const zipgpu_mgpu_data, coeffs: *const u64, out_limbs: c_int,
                                         combined_out: *mut u64) -> c_int;
    // encode_f (code_raa.rs:133-138): the code over field elements (stored residues, `limbs` u64 words each)
    pub fn zipgpu_encode_f(code: *mut zipgpu_code, num_rows: usize, limbs: c_int, modulus: *const u64, rows: *const u64,
                           out: *mut u64) -> c_int;
    // encode_wide::<In, Out> for any widths (verifier: In = Out = M, verify_z.rs:74-78)
    pub fn zipgpu_encode_wide(code: *mut zipgpu_code, num_rows: usize, in_limbs: c_int, out_limbs: c_int, rows: *const u64,
                              out: *mut u64) -> c_int;
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

/// Per-pp device state.  Borrows its context: a `Code` cannot outlive the `Ctx` whose device memory holds its tables
/// (zipgpu_ctx_destroy frees everything the context's allocator handed out).
pub struct Code<'ctx>(pub *mut zipgpu_code, core::marker::PhantomData<&'ctx Ctx>);
unsafe impl Send for Code<'_> {}
unsafe impl Sync for Code<'_> {}
impl<'ctx> Code<'ctx> {
    /// `perm1`/`perm2`: `shuffle_seeded` (zip/utils.rs:139-142) applied to `0..codeword_len` with the two seeds
    pub fn new(ctx: &'ctx Ctx, row_len: usize, rep: usize, in_limbs: usize, out_limbs: usize, perm1: &[u32], perm2: &[u32])
               -> Result<Self, String> {
        assert_eq!(perm1.len(), row_len * rep);
        assert_eq!(perm2.len(), row_len * rep);
        let mut p = core::ptr::null_mut();
        check(unsafe { zipgpu_code_create(ctx.0, row_len, rep, in_limbs as c_int, out_limbs as c_int,
                                          perm1.as_ptr(), perm2.as_ptr(), &mut p) })?;
        Ok(Code(p, core::marker::PhantomData))
    }
}
impl Drop for Code<'_> { fn drop(&mut self) { unsafe { zipgpu_code_destroy(self.0) } } }

/// ONE process, n GPUs (zipgpu_mgpu_*): same calls, rows / polynomials sharded over the devices.
pub struct MultiCtx(pub *mut zipgpu_mgpu);
unsafe impl Send for MultiCtx {}
unsafe impl Sync for MultiCtx {}
impl MultiCtx {
    /// `devices`: CUDA ordinals; empty = all visible devices
    pub fn new(devices: &[i32]) -> Result<Self, String> {
        let mut p = core::ptr::null_mut();
        let ptr = if devices.is_empty() { core::ptr::null() } else { devices.as_ptr() };
        check(unsafe { zipgpu_mgpu_create(ptr, devices.len() as c_int, &mut p) })?;
        Ok(MultiCtx(p))
    }
    pub fn num_devices(&self) -> usize { unsafe { zipgpu_mgpu_num_devices(self.0) as usize } }
}
impl Drop for MultiCtx { fn drop(&mut self) { unsafe { zipgpu_mgpu_destroy(self.0) } } }

pub struct MultiCode<'ctx>(pub *mut zipgpu_mgpu_code, core::marker::PhantomData<&'ctx MultiCtx>);
unsafe impl Send for MultiCode<'_> {}
unsafe impl Sync for MultiCode<'_> {}
impl<'ctx> MultiCode<'ctx> {
    pub fn new(ctx: &'ctx MultiCtx, row_len: usize, rep: usize, in_limbs: usize, out_limbs: usize, perm1: &[u32],
               perm2: &[u32]) -> Result<Self, String> {
        assert_eq!(perm1.len(), row_len * rep);
        assert_eq!(perm2.len(), row_len * rep);
        let mut p = core::ptr::null_mut();
        check(unsafe { zipgpu_mgpu_code_create(ctx.0, row_len, rep, in_limbs as c_int, out_limbs as c_int,
                                               perm1.as_ptr(), perm2.as_ptr(), &mut p) })?;
        Ok(MultiCode(p, core::marker::PhantomData))
    }
}
impl Drop for MultiCode<'_> { fn drop(&mut self) { unsafe { zipgpu_mgpu_code_destroy(self.0) } } }

/// Device-resident prover data of a multi-GPU commit (rows + layers stay sharded over the GPUs that produced them).
pub struct MultiData<'ctx>(pub *mut zipgpu_mgpu_data, core::marker::PhantomData<&'ctx MultiCtx>);
unsafe impl Send for MultiData<'_> {}
impl Drop for MultiData<'_> { fn drop(&mut self) { unsafe { zipgpu_mgpu_data_free(self.0) } } }
