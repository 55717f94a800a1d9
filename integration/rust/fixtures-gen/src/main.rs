//! fixtures-gen - known answers produced by the REAL reference stack, for tests/test_ark_fixtures.py.
//!
//! Three sub-commands (see Cargo.toml for the invocation):
//!
//! `seeded <in_dir> <out_dir>`  byte-level pin of the prover.  For every `<name>.r1cs` / `<name>.cases` pair written
//!     by `python tools/ark_fixtures.py inputs` (this repo's R1CS matrices of libzkp's circuits, full assignments z
//!     and prover randomness r, s) it runs ark-groth16's own setup (seeded StdRng) and
//!     `Groth16::<Bn254>::create_proof_with_reduction(circuit, &pk, r, s)` on a circuit that replays those matrices,
//!     and writes `<name>_ark_pk.bin`, `<name>_ark_vk.bin` (ark-serialize uncompressed) and `<name>_ark_proofs.bin`
//!     (256 B per case).  The GPU engine and the CPU oracle must reproduce every proof byte for byte from that pk.
//!
//! `reference <out_dir>`  pin of the circuits and key files.  Through libzkp's PUBLIC API only (its circuits are
//!     private): `set_snark_key_dir(out_dir)`, then `SnarkBackend::prove_equality_zk` / `prove_membership_zk`, which
//!     makes the reference generate and persist `equality_mimc_{pk,vk}.bin` / `membership_mimc_{pk,vk}.bin`
//!     (src/backend/snark.rs:72-139) and yields proofs under OsRng.  Writes `reference_proofs.bin`.  This repo must
//!     load those key files, verify those proofs, and prove from those keys (`tools/ark_fixtures.py ours`).
//!
//! `verify <dir>`  closes the loop: `ours_proofs.bin` (proofs made by the GPU engine from the reference's key files)
//!     go through `SnarkBackend::verify_equality_zk` / `verify_membership_zk` - "ark-groth16's verifier accepts".
//!
//! File formats are little-endian and documented in tools/ark_fixtures.py.
use ark_bn254::{Bn254, Fr};
use ark_ff::PrimeField;
use ark_groth16::Groth16;
use ark_relations::lc;
use ark_relations::r1cs::{ConstraintSynthesizer, ConstraintSystemRef, LinearCombination, SynthesisError, Variable};
use ark_serialize::CanonicalSerialize;
use ark_std::rand::{rngs::StdRng, SeedableRng};
use libzkp::backend::snark::{fr_to_commitment, mimc_hash_native, set_snark_key_dir, SnarkBackend};
use std::fs;
use std::path::Path;

struct Reader<'a> {
    b: &'a [u8],
    p: usize,
}
impl<'a> Reader<'a> {
    fn u32(&mut self) -> u32 {
        let v = u32::from_le_bytes(self.b[self.p..self.p + 4].try_into().unwrap());
        self.p += 4;
        v
    }
    fn u64(&mut self) -> u64 {
        let v = u64::from_le_bytes(self.b[self.p..self.p + 8].try_into().unwrap());
        self.p += 8;
        v
    }
    fn fr(&mut self) -> Fr {
        let v = Fr::from_le_bytes_mod_order(&self.b[self.p..self.p + 32]);
        self.p += 32;
        v
    }
    fn bytes(&mut self, n: usize) -> &'a [u8] {
        let v = &self.b[self.p..self.p + n];
        self.p += n;
        v
    }
}

/// Sparse rows of one matrix: (coefficient, column); column 0 = One, 1..n_inst = instance, the rest = witness.
type Rows = Vec<Vec<(Fr, usize)>>;

/// A circuit that enforces exactly the rows it is given - the matrices come from this repo, the proving from arkworks.
#[derive(Clone)]
struct MatrixCircuit {
    n_inst: usize,
    n_wit: usize,
    a: Rows,
    b: Rows,
    c: Rows,
    z: Option<Vec<Fr>>, // full assignment (z[0] = 1); None during setup
}

impl ConstraintSynthesizer<Fr> for MatrixCircuit {
    fn generate_constraints(self, cs: ConstraintSystemRef<Fr>) -> Result<(), SynthesisError> {
        let mut vars: Vec<Variable> = Vec::with_capacity(self.n_inst + self.n_wit);
        vars.push(Variable::One);
        for i in 1..self.n_inst {
            let v = self.z.as_ref().map(|z| z[i]);
            vars.push(cs.new_input_variable(|| v.ok_or(SynthesisError::AssignmentMissing))?);
        }
        for j in 0..self.n_wit {
            let v = self.z.as_ref().map(|z| z[self.n_inst + j]);
            vars.push(cs.new_witness_variable(|| v.ok_or(SynthesisError::AssignmentMissing))?);
        }
        let build = |row: &Vec<(Fr, usize)>| -> LinearCombination<Fr> {
            let mut l = lc!();
            for (coeff, col) in row {
                l = l + (*coeff, vars[*col]);
            }
            l
        };
        for i in 0..self.a.len() {
            cs.enforce_constraint(build(&self.a[i]), build(&self.b[i]), build(&self.c[i]))?;
        }
        Ok(())
    }
}

fn read_matrix(r: &mut Reader, m: usize) -> Rows {
    let nnz = r.u32() as usize;
    let rowptr: Vec<usize> = (0..=m).map(|_| r.u32() as usize).collect();
    let cols: Vec<usize> = (0..nnz).map(|_| r.u32() as usize).collect();
    let vals: Vec<Fr> = (0..nnz).map(|_| r.fr()).collect();
    (0..m).map(|i| (rowptr[i]..rowptr[i + 1]).map(|k| (vals[k], cols[k])).collect()).collect()
}

fn seeded(in_dir: &Path, out_dir: &Path) {
    fs::create_dir_all(out_dir).unwrap();
    for entry in fs::read_dir(in_dir).unwrap() {
        let path = entry.unwrap().path();
        if path.extension().map(|e| e != "r1cs").unwrap_or(true) {
            continue;
        }
        let name = path.file_stem().unwrap().to_str().unwrap().to_string();
        let raw = fs::read(&path).unwrap();
        let mut r = Reader { b: &raw, p: 0 };
        assert_eq!(r.bytes(4), b"LZR1");
        let (m, n_inst, n_wit) = (r.u32() as usize, r.u32() as usize, r.u32() as usize);
        let (a, b, c) = (read_matrix(&mut r, m), read_matrix(&mut r, m), read_matrix(&mut r, m));
        let shape = MatrixCircuit { n_inst, n_wit, a, b, c, z: None };
        // circuit_specific_setup as the reference calls it (snark.rs:318,337), but with a seeded RNG
        let mut rng = StdRng::seed_from_u64(0x6c7a6b70); // "lzkp"
        let pk = Groth16::<Bn254>::generate_random_parameters_with_reduction(shape.clone(), &mut rng).unwrap();
        let mut pk_bytes = Vec::new();
        pk.serialize_uncompressed(&mut pk_bytes).unwrap();
        let mut vk_bytes = Vec::new();
        pk.vk.serialize_uncompressed(&mut vk_bytes).unwrap();
        fs::write(out_dir.join(format!("{name}_ark_pk.bin")), &pk_bytes).unwrap();
        fs::write(out_dir.join(format!("{name}_ark_vk.bin")), &vk_bytes).unwrap();

        let raw = fs::read(in_dir.join(format!("{name}.cases"))).unwrap();
        let mut r = Reader { b: &raw, p: 0 };
        assert_eq!(r.bytes(4), b"LZCS");
        let (count, n_vars) = (r.u32() as usize, r.u32() as usize);
        assert_eq!(n_vars, n_inst + n_wit);
        let mut proofs = Vec::with_capacity(256 * count);
        for _ in 0..count {
            let z: Vec<Fr> = (0..n_vars).map(|_| r.fr()).collect();
            let (rr, ss) = (r.fr(), r.fr());
            let mut circ = shape.clone();
            circ.z = Some(z);
            let proof = Groth16::<Bn254>::create_proof_with_reduction(circ, &pk, rr, ss).unwrap();
            proof.serialize_uncompressed(&mut proofs).unwrap();
        }
        assert_eq!(proofs.len(), 256 * count);
        fs::write(out_dir.join(format!("{name}_ark_proofs.bin")), &proofs).unwrap();
        println!("{name}: m = {m}, {count} proofs, pk {} bytes", pk_bytes.len());
    }
}

/// reference_proofs.bin: u32 count, then per record  u32 kind (0 equality, 1 membership), u64 value, u32 set_len,
/// u64[set_len] set, [u8; 32] commitment, u32 proof_len, proof bytes.
fn reference(out_dir: &Path) {
    fs::create_dir_all(out_dir).unwrap();
    set_snark_key_dir(out_dir.to_str().unwrap()).expect("key dir");
    let mut out: Vec<u8> = Vec::new();
    let eq: [u64; 4] = [5, 42, 0, u64::MAX];
    let mb: [(u64, Vec<u64>); 3] = [(2, vec![1, 2, 3]), (25, vec![10, 20, 25, 30, 40]), (9, (0..64).collect())];
    out.extend(((eq.len() + mb.len()) as u32).to_le_bytes());
    for v in eq {
        let cm = fr_to_commitment(mimc_hash_native(v));
        let proof = SnarkBackend::prove_equality_zk(v, v, cm);
        assert!(!proof.is_empty() && SnarkBackend::verify_equality_zk(&proof, &cm));
        out.extend(0u32.to_le_bytes());
        out.extend(v.to_le_bytes());
        out.extend(0u32.to_le_bytes());
        out.extend(cm);
        out.extend((proof.len() as u32).to_le_bytes());
        out.extend(proof);
    }
    for (v, set) in mb {
        let cm = fr_to_commitment(mimc_hash_native(v));
        let proof = SnarkBackend::prove_membership_zk(v, set.clone(), cm);
        assert!(!proof.is_empty() && SnarkBackend::verify_membership_zk(&proof, &set, &cm));
        out.extend(1u32.to_le_bytes());
        out.extend(v.to_le_bytes());
        out.extend((set.len() as u32).to_le_bytes());
        for x in &set {
            out.extend(x.to_le_bytes());
        }
        out.extend(cm);
        out.extend((proof.len() as u32).to_le_bytes());
        out.extend(proof);
    }
    fs::write(out_dir.join("reference_proofs.bin"), out).unwrap();
    println!("wrote key files and reference_proofs.bin to {}", out_dir.display());
}

/// ours_proofs.bin has the layout of reference_proofs.bin; the keys in <dir> are the reference's own.
fn verify(dir: &Path) {
    set_snark_key_dir(dir.to_str().unwrap()).expect("key dir");
    let raw = fs::read(dir.join("ours_proofs.bin")).expect("run `python tools/ark_fixtures.py ours` first");
    let mut r = Reader { b: &raw, p: 0 };
    let count = r.u32();
    let mut rejected = 0;
    for i in 0..count {
        let kind = r.u32();
        let v = r.u64();
        let n = r.u32() as usize;
        let set: Vec<u64> = (0..n).map(|_| r.u64()).collect();
        let cm = r.bytes(32).to_vec();
        let len = r.u32() as usize;
        let proof = r.bytes(len);
        let ok = if kind == 0 { SnarkBackend::verify_equality_zk(proof, &cm) } else { SnarkBackend::verify_membership_zk(proof, &set, &cm) };
        println!("proof {i} (kind {kind}, value {v}): {}", if ok { "accepted" } else { "REJECTED" });
        rejected += (!ok) as u32;
    }
    std::process::exit(if rejected == 0 { 0 } else { 1 });
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    match args.get(1).map(|s| s.as_str()) {
        Some("seeded") if args.len() == 4 => seeded(Path::new(&args[2]), Path::new(&args[3])),
        Some("reference") if args.len() == 3 => reference(Path::new(&args[2])),
        Some("verify") if args.len() == 3 => verify(Path::new(&args[2])),
        _ => eprintln!("usage: fixtures-gen seeded <in_dir> <out_dir> | reference <out_dir> | verify <dir>"),
    }
}
