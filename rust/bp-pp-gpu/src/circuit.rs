//! `circuit::{PartitionType, Proof, SerializableProof, Witness}` -- the data models of src/circuit.rs:14-91, field for field.
//! `ArithmeticCircuit` itself is bound through `bppp_circuit_{commit,prove,verify}` (dense W_m / W_l, tabulated partition,
//! fresh `Transcript::new(label)`): see INTEGRATION.md section 2 for the descriptor a shim fills from the reference struct.
use k256::{AffinePoint, ProjectivePoint, Scalar};
use serde::{Deserialize, Serialize};

#[derive(Clone, Debug, Copy, PartialEq)]
pub enum PartitionType { LO, LL, LR, NO }

#[derive(Clone, Debug)]
pub struct Proof {
    pub c_l: ProjectivePoint,
    pub c_r: ProjectivePoint,
    pub c_o: ProjectivePoint,
    pub c_s: ProjectivePoint,
    pub r: Vec<ProjectivePoint>,
    pub x: Vec<ProjectivePoint>,
    pub l: Vec<Scalar>,
    pub n: Vec<Scalar>,
}

#[derive(Serialize, Deserialize, Clone, Debug)]
pub struct SerializableProof {
    pub c_l: AffinePoint,
    pub c_r: AffinePoint,
    pub c_o: AffinePoint,
    pub c_s: AffinePoint,
    pub r: Vec<AffinePoint>,
    pub x: Vec<AffinePoint>,
    pub l: Vec<Scalar>,
    pub n: Vec<Scalar>,
}

fn lift(v: &[AffinePoint]) -> Vec<ProjectivePoint> { v.iter().map(ProjectivePoint::from).collect() }
fn lower(v: &[ProjectivePoint]) -> Vec<AffinePoint> { v.iter().map(|p| p.to_affine()).collect() }

impl From<&SerializableProof> for Proof {
    fn from(s: &SerializableProof) -> Self {
        Proof { c_l: (&s.c_l).into(), c_r: (&s.c_r).into(), c_o: (&s.c_o).into(), c_s: (&s.c_s).into(), r: lift(&s.r), x: lift(&s.x), l: s.l.clone(), n: s.n.clone() }
    }
}
impl From<&Proof> for SerializableProof {
    fn from(p: &Proof) -> Self {
        SerializableProof { c_l: p.c_l.to_affine(), c_r: p.c_r.to_affine(), c_o: p.c_o.to_affine(), c_s: p.c_s.to_affine(), r: lower(&p.r), x: lower(&p.x), l: p.l.clone(), n: p.n.clone() }
    }
}

/// Arithmetic circuit witness (circuit.rs:80-91).
#[derive(Clone, Debug)]
pub struct Witness {
    pub v: Vec<Vec<Scalar>>,
    pub s_v: Vec<Scalar>,
    pub w_l: Vec<Scalar>,
    pub w_r: Vec<Scalar>,
    pub w_o: Vec<Scalar>,
}
