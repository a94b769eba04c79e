//! Golden-vector dumper for the lasso_b200 parity contract (see integration/rust/README.md of lasso_b200).
//! Lives inside the crate because `utils`, `poly` and `msm` are private modules.  Uses only the public
//! constructors the benchmarks and e2e tests use (src/benches/bench.rs:36-73, src/e2e_test.rs:17-62).
//!
//!   DETERMINISTIC_TEST_RNG=1 LASSO_DUMP_DIR=/tmp/lasso_vectors \
//!       cargo +nightly test --release dump_vectors -- --nocapture --test-threads=1
use std::fs::{create_dir_all, File};
use std::io::Write;
use std::path::PathBuf;

use ark_curve25519::{EdwardsProjective as G, Fr};
use ark_ec::CurveGroup;
use ark_ff::UniformRand;
use ark_serialize::CanonicalSerialize;
use ark_std::test_rng;
use merlin::Transcript;
use rand_chacha::rand_core::RngCore;

use crate::{
  lasso::{
    densified::DensifiedRepresentation,
    surge::{SparsePolyCommitmentGens, SparsePolynomialEvaluationProof},
  },
  poly::dense_mlpoly::PolyCommitmentGens,
  subtables::{
    and::AndSubtableStrategy, lt::LTSubtableStrategy, or::OrSubtableStrategy,
    range_check::RangeCheckSubtableStrategy, xor::XorSubtableStrategy, SubtableStrategy,
  },
  utils::math::Math,
  utils::random::RandomTape,
};

fn out_dir(case: &str) -> PathBuf {
  let root = std::env::var("LASSO_DUMP_DIR").unwrap_or_else(|_| "lasso_vectors".to_string());
  let p = PathBuf::from(root).join(case);
  create_dir_all(&p).unwrap();
  p
}

/// the in-memory Montgomery limbs of an Fr (what a `&[Fr]` looks like through a `*const u64`)
fn fr_limbs(x: &Fr, out: &mut Vec<u8>) {
  for l in x.0 .0.iter() {
    out.extend_from_slice(&l.to_le_bytes());
  }
}

/// gens_n.G || gens_1.G[0] || h as affine points, 64 B each (Montgomery limbs of x, then y)
fn gens_stream(g: &PolyCommitmentGens<G>, out: &mut Vec<u8>) -> usize {
  let mut pts: Vec<G> = g.gens.gens_n.G.clone();
  pts.push(g.gens.gens_1.G[0]);
  pts.push(g.gens.gens_n.h);
  for p in pts.iter() {
    let a = p.into_affine();
    for l in a.x.0 .0.iter() {
      out.extend_from_slice(&l.to_le_bytes());
    }
    for l in a.y.0 .0.iter() {
      out.extend_from_slice(&l.to_le_bytes());
    }
  }
  pts.len()
}

macro_rules! dump_case {
  ($test_name:ident, $case:expr, $kind:expr, $Strategy:ty, $C:expr, $M:expr, $log_r:expr, $sparsity:expr, $same_index:expr) => {
    #[test]
    fn $test_name() {
      const C: usize = $C;
      const M: usize = $M;
      type S = $Strategy;
      const NUM_MEMORIES: usize = <S as SubtableStrategy<Fr, C, M>>::NUM_MEMORIES;
      let log_m: usize = M.log_2();
      let s: usize = $sparsity;
      let log_s: usize = s.next_power_of_two().log_2();

      // inputs, drawn like src/benches/bench.rs:13-34 (one index repeated in all dimensions) or independently
      let mut rng = test_rng();
      let mut nz: Vec<[usize; C]> = Vec::with_capacity(s);
      for _ in 0..s {
        if $same_index {
          nz.push([rng.next_u64() as usize % M; C]);
        } else {
          let mut row = [0usize; C];
          for d in 0..C {
            row[d] = rng.next_u64() as usize % M;
          }
          nz.push(row);
        }
      }
      let r: Vec<Fr> = (0..log_s).map(|_| Fr::rand(&mut rng)).collect();
      // RandomTape::new draws its seed from a FRESH test_rng() (src/utils/random.rs:15-30): the same draw here
      let tape_seed = Fr::rand(&mut test_rng());

      let mut dense: DensifiedRepresentation<Fr, C> = DensifiedRepresentation::from_lookup_indices(&nz, log_m);
      let gens = SparsePolyCommitmentGens::<G>::new(b"gens_sparse_poly", C, s.next_power_of_two(), NUM_MEMORIES, log_m);
      let commitment = dense.commit::<G>(&gens);
      let mut random_tape = RandomTape::<G>::new(b"proof");
      let mut prover_transcript = Transcript::new(b"example");
      let proof = SparsePolynomialEvaluationProof::<G, C, M, S>::prove(
        &mut dense,
        &r,
        &gens,
        &mut prover_transcript,
        &mut random_tape,
      );
      let mut verify_transcript = Transcript::new(b"example");
      proof
        .verify(&commitment, &r, &gens, &mut verify_transcript)
        .expect("should verify");

      let dir = out_dir($case);
      let mut buf: Vec<u8> = Vec::new();
      for row in nz.iter() {
        for v in row.iter() {
          buf.extend_from_slice(&(*v as u64).to_le_bytes());
        }
      }
      File::create(dir.join("indices.u64")).unwrap().write_all(&buf).unwrap();
      buf.clear();
      for x in r.iter() {
        fr_limbs(x, &mut buf);
      }
      File::create(dir.join("r.fr")).unwrap().write_all(&buf).unwrap();
      buf.clear();
      fr_limbs(&tape_seed, &mut buf);
      File::create(dir.join("tape_seed.fr")).unwrap().write_all(&buf).unwrap();
      // the three PolyCommitmentGens share one label: they are prefixes of one generator stream; dump the widest
      buf.clear();
      let widest = [&gens.gens_combined_l_variate, &gens.gens_combined_log_m_variate, &gens.gens_derefs]
        .into_iter()
        .max_by_key(|g| g.gens.gens_n.n)
        .unwrap();
      let n_points = gens_stream(widest, &mut buf);
      File::create(dir.join("gens.aff")).unwrap().write_all(&buf).unwrap();
      buf.clear();
      commitment.serialize_compressed(&mut buf).unwrap();
      File::create(dir.join("commitment.bin")).unwrap().write_all(&buf).unwrap();
      buf.clear();
      proof.serialize_compressed(&mut buf).unwrap();
      File::create(dir.join("proof.bin")).unwrap().write_all(&buf).unwrap();
      let manifest = format!(
        "{{\"case\": \"{}\", \"kind\": {}, \"C\": {}, \"log_m\": {}, \"log_r\": {}, \"lookups\": {}, \"num_memories\": {}, \
         \"n_generators\": {}, \"generator_label\": \"gens_sparse_poly\", \"transcript_label\": \"example\", \
         \"tape_label\": \"proof\", \"deterministic_test_rng\": {}}}\n",
        $case,
        $kind,
        C,
        log_m,
        $log_r,
        s,
        NUM_MEMORIES,
        n_points,
        std::env::var("DETERMINISTIC_TEST_RNG").is_ok()
      );
      File::create(dir.join("manifest.json")).unwrap().write_all(manifest.as_bytes()).unwrap();
      println!("dumped {} -> {}", $case, dir.display());
    }
  };
}

// kinds: 0 AND, 1 OR, 2 XOR, 3 LT, 4 RANGE_CHECK (include/lasso_b200.h)
dump_case!(dump_vectors_and_c1_s1024, "and_c1_s1024", 0, AndSubtableStrategy, 1, 65536, 0, 1 << 10, true); // BASELINE configs[0]
dump_case!(dump_vectors_or_c2_s700, "or_c2_s700", 1, OrSubtableStrategy, 2, 256, 0, 700, false); // ragged: padded with address 0
dump_case!(dump_vectors_xor_c4_s4096, "xor_c4_s4096", 2, XorSubtableStrategy, 4, 65536, 0, 1 << 12, true); // the headline shape, small
dump_case!(dump_vectors_lt_c4_s128, "lt_c4_s128", 3, LTSubtableStrategy, 4, 16, 0, 128, false); // e2e_test.rs prove_4d_lt_big_s
dump_case!(dump_vectors_range40_c3_s16, "range40_c3_s16", 4, RangeCheckSubtableStrategy::<40>, 3, 256, 40, 16, false); // e2e_test.rs prove_3d_range
dump_case!(dump_vectors_xor_c4_s2p20, "xor_c4_s2p20", 2, XorSubtableStrategy, 4, 65536, 0, 1 << 20, true); // BASELINE configs[1] (minutes)
