//! Dumps what the unmodified reference computes for a fixed input, so the oracle (`oracle/`) and the CUDA path can be
//! pinned against it by `compare.py`.  Test infrastructure only.  Public API used:
//!   ferreus_rbf_utils::{FmmTree, KernelParams, KernelType}   (ferreus_rbf_utils/src/utils.rs:383-493,
//!                                                             kernel_helpers.rs:17-81)
//!   ferreus_bbfmm::{FmmParams, M2LCompressionType}           (ferreus_bbfmm/src/bbfmm.rs:62-108)
//!   ferreus_rbf::{RBFInterpolator, interpolant_config::*}    (ferreus_rbf/src/rbf.rs:267-333,
//!                                                             interpolant_config.rs:37-225)
//! All matrices travel as raw little-endian f64, row-major, no header.
//!
//!   matvec <pts> <n> <dim> <w> <nrhs> <kernel 0..9> <order> <adaptive 0|1> <sparse 0|1> <max_pts>
//!          <compression 0 none|1 svd|2 aca> <eps> <out>
//!       set_weights(w) + evaluate(w, pts)  ->  out (n x nrhs)
//!   fit <pts> <n> <dim> <values> <kernel 0 linear|1 tps|2 cubic> <out_point_coeffs> <out_poly_coeffs>
//!       RBFInterpolator::builder(..).build() with default settings -> coefficients
use std::fs;

use faer::Mat;
use ferreus_bbfmm::{FmmParams, M2LCompressionType};
use ferreus_rbf::interpolant_config::{InterpolantSettings, RBFKernelType};
use ferreus_rbf::RBFInterpolator;
use ferreus_rbf_utils::{FmmTree, KernelParams, KernelType};

fn read_mat(path: &str, rows: usize, cols: usize) -> Mat<f64> {
    let bytes = fs::read(path).expect("cannot read input");
    assert_eq!(bytes.len(), rows * cols * 8, "unexpected size of {path}");
    let vals: Vec<f64> = bytes
        .chunks_exact(8)
        .map(|c| f64::from_le_bytes(c.try_into().unwrap()))
        .collect();
    Mat::from_fn(rows, cols, |i, j| vals[i * cols + j])
}

fn write_mat(path: &str, m: &Mat<f64>) {
    let mut out = Vec::with_capacity(m.nrows() * m.ncols() * 8);
    for i in 0..m.nrows() {
        for j in 0..m.ncols() {
            out.extend_from_slice(&m[(i, j)].to_le_bytes());
        }
    }
    fs::write(path, out).expect("cannot write output");
}

fn kernel_type(index: usize) -> KernelType {
    // registry order, ferreus_rbf_utils/src/utils.rs:558-571
    match index {
        0 => KernelType::LinearRbf,
        1 => KernelType::ThinPlateSplineRbf,
        2 => KernelType::CubicRbf,
        3 => KernelType::Spheroidal3Rbf,
        4 => KernelType::Spheroidal5Rbf,
        5 => KernelType::Spheroidal7Rbf,
        6 => KernelType::Spheroidal9Rbf,
        7 => KernelType::Laplacian,
        8 => KernelType::OneOverR2,
        9 => KernelType::OneOverR4,
        _ => panic!("kernel index 0..9"),
    }
}

fn main() {
    let a: Vec<String> = std::env::args().collect();
    let num = |i: usize| -> usize { a[i].parse().expect("integer argument") };
    match a.get(1).map(|s| s.as_str()) {
        Some("matvec") => {
            let (n, dim, nrhs) = (num(3), num(4), num(6));
            let pts = read_mat(&a[2], n, dim);
            let w = read_mat(&a[5], n, nrhs);
            let order = num(8);
            let compression = match num(12) {
                0 => M2LCompressionType::None,
                1 => M2LCompressionType::SVD,
                _ => M2LCompressionType::ACA,
            };
            let params = FmmParams {
                max_points_per_cell: num(11),
                compression_type: compression,
                epsilon: a[13].parse().expect("epsilon"),
                eval_chunk_size: 1024,
            };
            let kp = KernelParams::builder(kernel_type(num(7))).build();
            let mut tree = FmmTree::new(pts.clone(), order, kp, num(9) != 0, num(10) != 0, None, Some(params));
            tree.set_weights(&w.as_ref());
            let out = tree.evaluate(&w.as_ref(), &pts).expect("evaluate failed");
            write_mat(&a[14], &out);
        }
        Some("fit") => {
            let (n, dim) = (num(3), num(4));
            let pts = read_mat(&a[2], n, dim);
            let vals = read_mat(&a[5], n, 1);
            let kernel = match num(6) {
                0 => RBFKernelType::Linear,
                1 => RBFKernelType::ThinPlateSpline,
                _ => RBFKernelType::Cubic,
            };
            let model = RBFInterpolator::builder(pts, vals, InterpolantSettings::builder(kernel).build()).build();
            write_mat(&a[7], &model.coefficients.point_coefficients);
            match &model.coefficients.poly_coefficients {
                Some(p) => write_mat(&a[8], p),
                None => fs::write(&a[8], []).unwrap(),
            }
        }
        _ => eprintln!("usage: see the header of src/main.rs"),
    }
}
