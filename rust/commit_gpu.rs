// UNCOMPILED (no Rust toolchain in this image) -- see README.md in this directory.
//
// Drop-in for zinc: save as src/zip/pcs/commit_gpu.rs, add `#[cfg(feature = "gpu")] mod commit_gpu;` to
// src/zip/pcs.rs, `gpu = ["dep:zip-gpu-sys"]` to [features] and the path dependency to Cargo.toml.
// With the feature on, `MultilinearZip::<ZT, RaaCode<ZT>>::commit` et al. (commit.rs:50-183) forward here;
// callers (zinc/prover.rs:313-315, benches/zip_benches.rs) are unchanged.
//
// Needs one accessor in src/zip/code_raa.rs because RaaCode's seeds are private (code_raa.rs:16-32):
//
//     impl<ZT: ZipTypes> RaaCode<ZT> {
//         /// gather form of the two shuffles: shuffled[i] == original[perm[i]]
//         pub(crate) fn permutations(&self) -> (Vec<u32>, Vec<u32>) {
//             let n = self.codeword_len();
//             let mut p1: Vec<u32> = (0..n as u32).collect();
//             let mut p2 = p1.clone();
//             shuffle_seeded(&mut p1, self.perm_1_seed);   // the real rand 0.9.2: parity stays on this side
//             shuffle_seeded(&mut p2, self.perm_2_seed);
//             (p1, p2)
//         }
//         pub(crate) fn repetition_factor(&self) -> usize { self.repetition_factor }
//     }
use std::sync::OnceLock;

use zip_gpu_sys as sys;

use super::{
    structs::{MultilinearZip, MultilinearZipCommitment, MultilinearZipData, MultilinearZipParams},
    utils::{validate_input, MerkleTree},
};
use crate::{
    poly_z::mle::DenseMultilinearExtension,
    traits::{Field, Integer, PrimitiveConversion, Words, ZipTypes},
    zip::{code::LinearCode, code_raa::RaaCode, Error},
};

/// The process-wide multi-GPU context: `ZIPGPU_DEVICES=0,1,2,3` picks the devices, default = all visible B200s.
/// One device is simply the n = 1 case; rows (commit) and polynomials (batch_commit) are sharded inside the library.
fn ctx() -> &'static sys::MultiCtx {
    static CTX: OnceLock<sys::MultiCtx> = OnceLock::new();
    CTX.get_or_init(|| {
        let devs: Vec<i32> = std::env::var("ZIPGPU_DEVICES").ok()
            .map(|s| s.split(',').filter_map(|x| x.trim().parse().ok()).collect()).unwrap_or_default();
        sys::MultiCtx::new(&devs).expect("zipgpu: no B200 visible (the gpu feature has no CPU fallback)")
    })
}

/// `[Int<n>]` -> flat little-endian limbs.  Goes through `Integer::as_words` (traits/types.rs:183) instead of
/// transmuting: crypto_bigint::Int is not guaranteed `repr(transparent)` all the way down.
fn limbs<I: Integer>(xs: &[I]) -> Vec<u64> {
    let mut out = Vec::with_capacity(xs.len() * I::W::num_words());
    for x in xs { out.extend_from_slice(x.as_words()); }
    out
}
fn from_limbs<K: Integer>(flat: &[u64]) -> Vec<K> {
    // Words is `Default + IndexMut<usize>` (traits/types.rs:125-138); its Word converts from u64 through
    // PrimitiveConversion (traits/types.rs:225-227,270-284)
    let n = K::W::num_words();
    flat.chunks_exact(n)
        .map(|chunk| {
            let mut w = K::W::default();
            for (i, x) in chunk.iter().enumerate() {
                w[i] = <<K::W as Words>::Word as PrimitiveConversion<u64>>::from_primitive(*x);
            }
            K::from_words(w)
        })
        .collect()
}
fn hashes(bytes: &[u8]) -> Vec<blake3::Hash> {
    bytes.chunks_exact(32).map(|c| blake3::Hash::from_bytes(c.try_into().unwrap())).collect()
}

impl<ZT: ZipTypes> MultilinearZip<ZT, RaaCode<ZT>> {
    fn gpu_code(pp: &MultilinearZipParams<ZT, RaaCode<ZT>>) -> sys::MultiCode<'static> {
        // a production version caches this per pp (the tables depend only on the seeds and cw); the prover builds one
        // code per proof anyway (zinc/prover.rs:313), and zipgpu_mgpu_code_create takes its buffers from a cache
        let lc = &pp.linear_code;
        let (p1, p2) = lc.permutations();
        sys::MultiCode::new(ctx(), lc.row_len(), lc.repetition_factor(), <ZT::N as Integer>::W::num_words(),
                            <ZT::K as Integer>::W::num_words(), &p1, &p2)
            .unwrap_or_else(|e| panic!("{e}"))   // ZIPGPU_ERR_WIDTH carries code_raa.rs:68-72's message
    }

    /// the reference's checks of commit.rs:54-63, host-side, BEFORE the FFI call, same Err / same panic text
    fn check_poly<F: Field>(pp: &MultilinearZipParams<ZT, RaaCode<ZT>>, poly: &DenseMultilinearExtension<ZT::N>)
        -> Result<(), Error> {
        validate_input("commit", pp.num_vars, [poly], None::<&[F]>)?;          // Err(InvalidPcsParam) as upstream
        let expected = pp.num_rows * pp.linear_code.row_len();
        assert_eq!(poly.evaluations.len(), expected,                             // commit.rs:56-63
            "Polynomial has an incorrect number of evaluations ({}) for the expected matrix size ({})",
            poly.evaluations.len(), expected);
        Ok(())
    }

    fn assemble(pp: &MultilinearZipParams<ZT, RaaCode<ZT>>, depth: usize, rows: &[u64], layers: &[u8], roots: &[u8])
        -> (MultilinearZipData<ZT::K>, MultilinearZipCommitment) {
        let per_row = (2usize << depth) - 2;
        let roots = hashes(roots);
        let trees: Vec<MerkleTree> = layers.chunks_exact(per_row * 32).zip(&roots)
            .map(|(l, r)| MerkleTree { root: *r, depth, layers: hashes(l) }).collect();
        assert_eq!(trees.len(), pp.num_rows);                                    // commit.rs:76
        (MultilinearZipData::new(from_limbs(rows), trees), MultilinearZipCommitment { roots })
    }

    /// commit.rs:50-87
    pub fn commit_gpu<F: Field>(
        pp: &MultilinearZipParams<ZT, RaaCode<ZT>>,
        poly: &DenseMultilinearExtension<ZT::N>,
    ) -> Result<(MultilinearZipData<ZT::K>, MultilinearZipCommitment), Error> {
        Self::check_poly::<F>(pp, poly)?;
        let cw = pp.linear_code.codeword_len();
        assert!(cw.is_power_of_two());                                           // utils.rs:75 via commit.rs:73
        let depth = cw.ilog2() as usize;                                         // commit.rs:67
        let k = <ZT::K as Integer>::W::num_words();
        let code = Self::gpu_code(pp);
        let evals = limbs(&poly.evaluations);
        let mut rows = vec![0u64; pp.num_rows * cw * k];
        let mut layers = vec![0u8; pp.num_rows * ((2usize << depth) - 2) * 32];
        let mut roots = vec![0u8; pp.num_rows * 32];
        sys::check(unsafe { sys::zipgpu_mgpu_commit(code.0, pp.num_rows, evals.as_ptr(), rows.as_mut_ptr(),
                                                    layers.as_mut_ptr(), roots.as_mut_ptr()) })
            .map_err(Error::InvalidPcsParam)?;
        Ok(Self::assemble(pp, depth, &rows, &layers, &roots))
    }

    /// commit.rs:104-119: encode only, empty trees and roots
    pub fn commit_no_merkle_gpu<F: Field>(
        pp: &MultilinearZipParams<ZT, RaaCode<ZT>>,
        poly: &DenseMultilinearExtension<ZT::N>,
    ) -> Result<(MultilinearZipData<ZT::K>, MultilinearZipCommitment), Error> {
        Self::check_poly::<F>(pp, poly)?;
        let rows = Self::encode_rows_gpu(pp, pp.linear_code.codeword_len(), pp.linear_code.row_len(), &poly.evaluations);
        Ok((MultilinearZipData::new(rows, vec![]), MultilinearZipCommitment { roots: vec![] }))
    }

    /// commit.rs:158-183
    pub fn encode_rows_gpu(pp: &MultilinearZipParams<ZT, RaaCode<ZT>>, codeword_len: usize, _row_len: usize,
                           evals: &[ZT::N]) -> Vec<ZT::K> {
        let k = <ZT::K as Integer>::W::num_words();
        let code = Self::gpu_code(pp);
        let flat = limbs(evals);
        let mut rows = vec![0u64; pp.num_rows * codeword_len * k];
        sys::check(unsafe { sys::zipgpu_mgpu_encode_rows(code.0, pp.num_rows, flat.as_ptr(), rows.as_mut_ptr()) })
            .unwrap_or_else(|e| panic!("{e}"));
        from_limbs(&rows)
    }

    /// commit.rs:134-142: all polynomials in ONE submission -- polynomial p goes to device p mod n, and on each device
    /// the batch runs as one matrix (H2D of poly p+1 overlaps the kernels of p)
    pub fn batch_commit_gpu<F: Field>(
        pp: &MultilinearZipParams<ZT, RaaCode<ZT>>,
        polys: &[DenseMultilinearExtension<ZT::N>],
    ) -> Result<Vec<(MultilinearZipData<ZT::K>, MultilinearZipCommitment)>, Error> {
        for poly in polys { Self::check_poly::<F>(pp, poly)?; }
        if polys.is_empty() { return Ok(vec![]); }
        let cw = pp.linear_code.codeword_len();
        assert!(cw.is_power_of_two());
        let depth = cw.ilog2() as usize;
        let k = <ZT::K as Integer>::W::num_words();
        let code = Self::gpu_code(pp);
        let evals: Vec<Vec<u64>> = polys.iter().map(|p| limbs(&p.evaluations)).collect();
        let mut rows: Vec<Vec<u64>> = polys.iter().map(|_| vec![0u64; pp.num_rows * cw * k]).collect();
        let mut layers: Vec<Vec<u8>> = polys.iter().map(|_| vec![0u8; pp.num_rows * ((2usize << depth) - 2) * 32]).collect();
        let mut roots: Vec<Vec<u8>> = polys.iter().map(|_| vec![0u8; pp.num_rows * 32]).collect();
        let ev_ptrs: Vec<*const u64> = evals.iter().map(|v| v.as_ptr()).collect();
        let row_ptrs: Vec<*mut u64> = rows.iter_mut().map(|v| v.as_mut_ptr()).collect();
        let lay_ptrs: Vec<*mut u8> = layers.iter_mut().map(|v| v.as_mut_ptr()).collect();
        let root_ptrs: Vec<*mut u8> = roots.iter_mut().map(|v| v.as_mut_ptr()).collect();
        sys::check(unsafe { sys::zipgpu_mgpu_batch_commit(code.0, polys.len(), pp.num_rows, ev_ptrs.as_ptr(),
                                                          row_ptrs.as_ptr(), lay_ptrs.as_ptr(), root_ptrs.as_ptr()) })
            .map_err(Error::InvalidPcsParam)?;
        Ok((0..polys.len()).map(|p| Self::assemble(pp, depth, &rows[p], &layers[p], &roots[p])).collect())
    }

    /// For provers that open right away (zinc/prover.rs:315-320): the rows and layers stay on the GPUs that produced them
    /// (sharded by row range), only the roots come back; `open` then pulls the 1000 opened columns with
    /// zipgpu_mgpu_data_open_columns_wire and the proximity row with zipgpu_mgpu_data_combine_rows.
    pub fn commit_resident_gpu<F: Field>(
        pp: &MultilinearZipParams<ZT, RaaCode<ZT>>,
        poly: &DenseMultilinearExtension<ZT::N>,
    ) -> Result<(sys::MultiData<'static>, MultilinearZipCommitment), Error> {
        Self::check_poly::<F>(pp, poly)?;
        let code = Self::gpu_code(pp);
        let evals = limbs(&poly.evaluations);
        let mut roots = vec![0u8; pp.num_rows * 32];
        let mut h = core::ptr::null_mut();
        sys::check(unsafe { sys::zipgpu_mgpu_commit_resident(code.0, pp.num_rows, evals.as_ptr(), roots.as_mut_ptr(), &mut h) })
            .map_err(Error::InvalidPcsParam)?;
        Ok((sys::MultiData(h, core::marker::PhantomData), MultilinearZipCommitment { roots: hashes(&roots) }))
    }
}
