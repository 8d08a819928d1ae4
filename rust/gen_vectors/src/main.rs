// UNCOMPILED (no Rust toolchain in this image) -- see ../../README.md
//
// Golden vectors from the real reference, for tests/test_rust_vectors.py:
//   shuffle_seeded (zinc src/zip/utils.rs:139-142, rand 0.9.2) applied to 0..n for n in {16, 512, 8192} and the seeds
//     {1, 2} (MockTranscript, src/zip/pcs/tests.rs:24-56) and {0xb9736f582676e7e8, 0xd7397e6260ce9c3e} (the seeds a
//     fresh KeccakTranscript yields, src/transcript.rs:183-185);
//   StdRng::seed_from_u64(seed) first 8 u32 words, to localise a mismatch (PCG seeding / ChaCha12 / word order);
//   MultilinearZip::commit roots (src/zip/pcs/commit.rs:50-87) for nv in {4, 8, 12}, evaluations 1..=2^nv
//     (commit.rs:234), with MockTranscript and with a fresh KeccakTranscript; plus rows[0..4] of the first codeword and
//     the first leaf hash, to localise a mismatch (encode vs to_bytes vs tree).
//
// `shuffle_seeded`, `MockTranscript` are crate-private in zinc: run this as a `#[test]` inside the crate
// (copy `mod gen` below into src/zip/pcs/tests.rs and `cargo test gen_vectors -- --nocapture`) or make them `pub`.
use std::fmt::Write;

use rand::{RngCore, SeedableRng, rngs::StdRng};
use zinc::{
    field::Int,
    poly_z::mle::DenseMultilinearExtension,
    transcript::KeccakTranscript,
    zip::{
        code::{DefaultLinearCodeSpec, LinearCode},
        code_raa::RaaCode,
        pcs::structs::{MultilinearZip, ZipTranscript},
        utils::shuffle_seeded,
    },
    define_random_field_zip_types, implement_random_field_zip_types,
};

define_random_field_zip_types!();
implement_random_field_zip_types!(1);
type ZT = RandomFieldZipTypes<1>;

/// counter transcript of src/zip/pcs/tests.rs:24-56: get_u64 returns 1, 2, ...
struct Mock(u64);
impl ZipTranscript<Int<2>> for Mock {
    fn get_encoding_element(&mut self) -> Int<2> { self.0 += 1; Int::from(self.0 as i64) }
    fn get_u64(&mut self) -> u64 { self.0 += 1; self.0 }
    fn sample_unique_columns(&mut self, range: std::ops::Range<usize>, columns: &mut std::collections::BTreeSet<usize>,
                             count: usize) -> usize {
        self.0 += 1;
        let mut inserted = 0;
        for i in range.clone() { if columns.insert(i) { inserted += 1; if inserted == count { break; } } }
        inserted
    }
}

fn hex(b: &[u8]) -> String { b.iter().map(|x| format!("{x:02x}")).collect() }

fn main() {
    let mut out = String::from("{\n \"generator\": \"rust/gen_vectors (zinc + rand 0.9.2 + blake3 1.8.2)\",\n");
    // ---- shuffle_seeded ----
    let seeds: [u64; 4] = [1, 2, 0xb9736f582676e7e8, 0xd7397e6260ce9c3e];
    out.push_str(" \"stdrng_first_words\": {");
    for (k, s) in seeds.iter().enumerate() {
        let mut rng = StdRng::seed_from_u64(*s);
        let w: Vec<String> = (0..8).map(|_| rng.next_u32().to_string()).collect();
        write!(out, "{}\"{}\": [{}]", if k > 0 { ", " } else { "" }, s, w.join(", ")).unwrap();
    }
    out.push_str("},\n \"shuffle_seeded\": {");
    let mut first = true;
    for n in [16usize, 512, 8192] {
        for s in seeds {
            let mut v: Vec<u32> = (0..n as u32).collect();
            shuffle_seeded(&mut v, s);
            let items: Vec<String> = v.iter().map(|x| x.to_string()).collect();
            write!(out, "{}\"{}:{}\": [{}]", if first { "" } else { ", " }, n, s, items.join(", ")).unwrap();
            first = false;
        }
    }
    out.push_str("},\n \"commit\": {");
    // ---- commit roots ----
    first = true;
    for nv in [4usize, 8, 12] {
        for tr in ["mock", "keccak"] {
            let code: RaaCode<ZT> = if tr == "mock" {
                RaaCode::new(&DefaultLinearCodeSpec, 1 << nv, &mut Mock(0))
            } else {
                RaaCode::new(&DefaultLinearCodeSpec, 1 << nv, &mut KeccakTranscript::new())
            };
            let pp = MultilinearZip::<ZT, _>::setup(1 << nv, code);
            let evals: Vec<Int<1>> = (1..=(1i64 << nv)).map(Int::from).collect();
            let poly = DenseMultilinearExtension::from_evaluations_vec(nv, evals);
            let (data, comm) = MultilinearZip::<ZT, _>::commit::<zinc::field::RandomField<4>>(&pp, &poly).unwrap();
            let roots: Vec<String> = comm.roots.iter().map(|h| format!("\"{}\"", hex(h.as_bytes()))).collect();
            let head: Vec<String> = data.rows[..4].iter()
                .map(|x| format!("[{}]", x.as_words().iter().map(|w| w.to_string()).collect::<Vec<_>>().join(", ")))
                .collect();
            write!(out, "{}\"{}:{}\": {{\"num_rows\": {}, \"row_len\": {}, \"rows_head\": [{}], \"roots\": [{}]}}",
                   if first { "" } else { ", " }, nv, tr, pp.num_rows, pp.linear_code.row_len(), head.join(", "),
                   roots.join(", ")).unwrap();
            first = false;
        }
    }
    out.push_str("}\n}\n");
    print!("{out}");
}
