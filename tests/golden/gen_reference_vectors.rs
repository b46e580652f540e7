// gen_reference_vectors.rs -- ciphertext-level golden vectors from the UNMODIFIED reference crate.
//
// This repository's oracle (oracle/exacto_oracle.c) is a restatement of the reference; the reference itself ships no
// ciphertext-level vectors and cannot be built in the image this repository was developed in (no cargo).  This file
// lets a maintainer WITH cargo close that gap without touching either code base:
//
//   cp tests/golden/gen_reference_vectors.rs  <exacto checkout>/tests/
//   cd <exacto checkout>
//   EXACTO_VECTOR_DIR=/tmp/exacto_vectors cargo test --release --test gen_reference_vectors -- --nocapture
//   cp /tmp/exacto_vectors/*.u64  <this repository>/tests/golden/reference_vectors/
//   python -m pytest tests/test_reference_vectors.py          # oracle (CPU) and, with -m gpu, the CUDA path
//
// Every file is raw little-endian u64.  Inputs and outputs are in the COEFFICIENT domain (canonical residues in
// [0, q)): the negacyclic ring element is unique there, whereas NTT-domain words depend on concrete-ntt's root and
// ordering (ring/ntt.rs:42-67), which the two code bases do not share.  For each case `<name>`:
//
//   <name>.meta.u64   [n, q, num_aux, aux.., plain_modulus, gadget_base, gadget_digits, dbfv_base, d, dbfv_plain_modulus]
//   <name>.in1.u64    ct1  [d][2][n]      <name>.in2.u64   ct2  [d][2][n]
//   <name>.rlk.u64    key  [G][2][n]      <name>.out.u64   dbfv_mul(ct1, ct2, rlk)  [d][2][n]
//
// Cases: BASELINE.json configs 1 (compact_bfv, d = 1: out = bfv_mul_and_relin), 2 (compact_dbfv), 3' (README custom
// set + the u64 profile's aux primes, d = 2) and 4 (u64 profile, d = 8).  Inputs are uniform residues from a
// SplitMix64 stream -- exactly what the multiplication sees for real ciphertexts -- so no key generation is involved.

use std::io::Write;
use std::sync::Arc;

use exacto::bfv::eval::bfv_mul_and_relin;
use exacto::bfv::keygen::RelinKey;
use exacto::bfv::BfvCiphertext;
use exacto::dbfv::ciphertext::DbfvCiphertext;
use exacto::dbfv::eval::dbfv_mul;
use exacto::params::presets::{compact_bfv, compact_dbfv, u64_dbfv};
use exacto::params::{BfvParams, BfvParamsBuilder, DbfvParams};
use exacto::ring::poly::CoeffPoly;
use exacto::ring::rns::RnsPoly;

struct SplitMix64(u64);
impl SplitMix64 {
    fn next(&mut self) -> u64 {
        self.0 = self.0.wrapping_add(0x9E37_79B9_7F4A_7C15);
        let mut z = self.0;
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
        z ^ (z >> 31)
    }
    fn poly(&mut self, n: usize, q: u64) -> CoeffPoly {
        CoeffPoly { coeffs: (0..n).map(|_| self.next() % q).collect(), modulus: q }
    }
}

fn dump(dir: &str, name: &str, what: &str, words: &[u64]) {
    let mut f = std::fs::File::create(format!("{dir}/{name}.{what}.u64")).unwrap();
    for w in words {
        f.write_all(&w.to_le_bytes()).unwrap();
    }
}

fn rns(p: &CoeffPoly, bfv: &BfvParams) -> RnsPoly {
    RnsPoly::from_coeff_poly(p, &bfv.ct_basis).unwrap()
}

fn run_case(dir: &str, name: &str, bfv: &Arc<BfvParams>, dbfv: Option<&Arc<DbfvParams>>, seed: u64) {
    let n = bfv.ring_degree;
    let q = bfv.ct_basis.moduli[0];
    let d = dbfv.map(|p| p.num_digits).unwrap_or(1);
    let mut rng = SplitMix64(seed);
    let mut in1 = Vec::new();
    let mut in2 = Vec::new();
    let mut limbs1 = Vec::new();
    let mut limbs2 = Vec::new();
    for (words, limbs) in [(&mut in1, &mut limbs1), (&mut in2, &mut limbs2)] {
        for _ in 0..d {
            let c0 = rng.poly(n, q);
            let c1 = rng.poly(n, q);
            words.extend_from_slice(&c0.coeffs);
            words.extend_from_slice(&c1.coeffs);
            limbs.push(BfvCiphertext { c: vec![rns(&c0, bfv), rns(&c1, bfv)], params: bfv.clone() });
        }
    }
    let mut rlk_words = Vec::new();
    let mut keys = Vec::new();
    for _ in 0..bfv.gadget_digits {
        let k0 = rng.poly(n, q);
        let k1 = rng.poly(n, q);
        rlk_words.extend_from_slice(&k0.coeffs);
        rlk_words.extend_from_slice(&k1.coeffs);
        keys.push((rns(&k0, bfv), rns(&k1, bfv)));
    }
    let rlk = RelinKey { keys, params: bfv.clone() };
    let out_limbs: Vec<BfvCiphertext> = match dbfv {
        None => vec![bfv_mul_and_relin(&limbs1[0], &limbs2[0], &rlk).unwrap()],
        Some(p) => {
            let a = DbfvCiphertext { limbs: limbs1, degree: d, mul_depth: 0, params: p.clone() };
            let b = DbfvCiphertext { limbs: limbs2, degree: d, mul_depth: 0, params: p.clone() };
            let prod = dbfv_mul(&a, &b, &rlk).unwrap();
            assert_eq!(prod.limbs.len(), d);
            assert_eq!(prod.mul_depth, 1);
            prod.limbs
        }
    };
    let mut out = Vec::new();
    for limb in &out_limbs {
        assert_eq!(limb.c.len(), 2);
        for comp in &limb.c {
            out.extend_from_slice(&comp.to_coeff_poly(&bfv.ct_basis).coeffs);
        }
    }
    let aux: Vec<u64> = bfv.aux_basis.as_ref().map(|b| b.moduli.clone()).unwrap_or_default();
    let mut meta = vec![n as u64, q, aux.len() as u64];
    meta.extend_from_slice(&aux);
    meta.extend_from_slice(&[bfv.plain_modulus, bfv.gadget_base, bfv.gadget_digits as u64]);
    match dbfv {
        None => meta.extend_from_slice(&[2, 1, 2]),
        Some(p) => meta.extend_from_slice(&[p.base, p.num_digits as u64, p.plain_modulus]),
    }
    dump(dir, name, "meta", &meta);
    dump(dir, name, "in1", &in1);
    dump(dir, name, "in2", &in2);
    dump(dir, name, "rlk", &rlk_words);
    dump(dir, name, "out", &out);
    eprintln!("{name}: n={n} q={q} d={d} G={} -> {} output words", bfv.gadget_digits, out.len());
}

#[test]
fn gen_reference_vectors() {
    let dir = std::env::var("EXACTO_VECTOR_DIR").unwrap_or_else(|_| "target/exacto_vectors".into());
    std::fs::create_dir_all(&dir).unwrap();

    // config 1: compact_bfv, bfv_mul_and_relin (src/params/presets.rs:24-35)
    let c1 = compact_bfv().unwrap();
    run_case(&dir, "cfg1_compact_bfv", &c1, None, 0xE8AC_7001);

    // config 2: compact_dbfv, d = 2 (src/params/presets.rs:86-98)
    let c2 = compact_dbfv().unwrap();
    run_case(&dir, "cfg2_compact_dbfv", &c2.bfv_params.clone(), Some(&c2), 0xE8AC_7002);

    // config 3': README custom set (README.md:109-119) + the u64 profile's aux primes (without them the reference
    // itself refuses: schoolbook overflow guard, src/bfv/eval.rs:426-431), wrapped as dBFV p = 65536, b = 256, d = 2
    let bfv3 = BfvParamsBuilder::new()
        .ring_degree(4096)
        .plain_modulus(65537)
        .ct_moduli(vec![576460752308273153])
        .aux_moduli(vec![18014398509998081, 36028797018972161])
        .sigma(3.2)
        .build()
        .unwrap();
    let c3 = DbfvParams::new(bfv3.clone(), 256, 2, 65536).unwrap();
    run_case(&dir, "cfg3p_custom_dbfv", &bfv3, Some(&c3), 0xE8AC_7003);

    // config 4: paper_repro u64 profile, d = 8 (src/params/presets.rs:61-75)
    let c4 = u64_dbfv().unwrap();
    run_case(&dir, "cfg4_u64_dbfv", &c4.bfv_params.clone(), Some(&c4), 0xE8AC_7004);
}
