//! `svdlibrs` API on top of libsvdlas2_b200 (NVIDIA B200, sm_100a).
//!
//! Same names, argument meaning, defaults and result types as svdlibrs 0.5.1; the LAS2 Lanczos recurrence, the
//! sparse operator, reorthogonalisation and the Ritz back-transform run on the GPU.  UNBUILT in this repository
//! (no Rust toolchain in the build image) -- see rust_shim/README.md.
#![allow(non_snake_case)]

extern crate ndarray;
use ndarray::prelude::*;
use std::ffi::CStr;

pub mod error;
mod ffi;
use error::SvdLibError;

/// Raw arrays of a sparse matrix in the layout the C ABI takes (usize == u64 on every supported target).
#[doc(hidden)]
pub enum RawSparse<'a> {
    Csr { offsets: &'a [usize], indices: &'a [usize], values: &'a [f64] },
    Csc { offsets: &'a [usize], indices: &'a [usize], values: &'a [f64] },
    Coo { rows: Vec<usize>, cols: Vec<usize>, values: Vec<f64> },
}

/// Sparse matrix
pub trait SMat {
    fn nrows(&self) -> usize;
    fn ncols(&self) -> usize;
    fn nnz(&self) -> usize;
    fn svd_opa(&self, x: &[f64], y: &mut [f64], transposed: bool); // y = A*x
    /// Hands the matrix to the GPU library without conversion.  Implemented for the nalgebra-sparse types.
    #[doc(hidden)]
    fn raw(&self) -> Option<RawSparse<'_>> {
        None
    }
}

/// Singular Value Decomposition Components: `ut` is d x nrows(A), `vt` is d x ncols(A)
#[derive(Debug, Clone, PartialEq)]
pub struct SvdRec {
    pub d: usize,
    pub ut: Array2<f64>,
    pub s: Array1<f64>,
    pub vt: Array2<f64>,
    pub diagnostics: Diagnostics,
}

/// Computational Diagnostics
#[derive(Debug, Clone, PartialEq)]
pub struct Diagnostics {
    pub non_zero: usize,
    pub dimensions: usize,
    pub iterations: usize,
    pub transposed: bool,
    pub lanczos_steps: usize,
    pub ritz_values_stabilized: usize,
    pub significant_values: usize,
    pub singular_values: usize,
    pub end_interval: [f64; 2],
    pub kappa: f64,
    pub random_seed: u32,
}

/// SVD at full dimensionality: `svdLAS2(A, 0, 0, &[-1.0e-30, 1.0e-30], 1.0e-6, 0)`
pub fn svd(A: &dyn SMat) -> Result<SvdRec, SvdLibError> {
    svdLAS2(A, 0, 0, &[-1.0e-30, 1.0e-30], 1.0e-6, 0)
}

/// SVD at desired dimensionality
pub fn svd_dim(A: &dyn SMat, dimensions: usize) -> Result<SvdRec, SvdLibError> {
    svdLAS2(A, dimensions, 0, &[-1.0e-30, 1.0e-30], 1.0e-6, 0)
}

/// SVD at desired dimensionality with supplied seed (`> 0`; otherwise a seed is generated)
pub fn svd_dim_seed(A: &dyn SMat, dimensions: usize, random_seed: u32) -> Result<SvdRec, SvdLibError> {
    svdLAS2(A, dimensions, 0, &[-1.0e-30, 1.0e-30], 1.0e-6, random_seed)
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::svdlas2_last_error()).to_string_lossy().into_owned() }
}

/// Compute a singular value decomposition (parameters exactly as in svdlibrs::svdLAS2)
pub fn svdLAS2(
    A: &dyn SMat,
    dimensions: usize,
    iterations: usize,
    end_interval: &[f64; 2],
    kappa: f64,
    random_seed: u32,
) -> Result<SvdRec, SvdLibError> {
    let raw = A.raw().ok_or_else(|| {
        SvdLibError::Las2Error("unsupported SMat implementor: the B200 path needs the raw CSR/CSC/COO arrays".into())
    })?;
    let (entry, a, b, values): (ffi::Entry, &[usize], &[usize], &[f64]) = match &raw {
        RawSparse::Csr { offsets, indices, values } => (ffi::CSR, offsets, indices, values),
        RawSparse::Csc { offsets, indices, values } => (ffi::CSC, offsets, indices, values),
        RawSparse::Coo { rows, cols, values } => (ffi::COO, rows, cols, values),
    };
    let mut out: *mut ffi::svdlas2_result = std::ptr::null_mut();
    // usize and u64 have the same layout on 64-bit targets: the index arrays are borrowed, not copied
    let rc = unsafe {
        entry(
            A.nrows() as u64, A.ncols() as u64, values.len() as u64, a.as_ptr() as *const u64,
            b.as_ptr() as *const u64, values.as_ptr(), dimensions as u64, iterations as u64, end_interval.as_ptr(),
            kappa, random_seed, &mut out,
        )
    };
    match rc {
        ffi::OK => {}
        ffi::ERR_IMTQLB => return Err(SvdLibError::ImtqlbError(last_error())),
        ffi::ERR_STARTV => return Err(SvdLibError::StartvError(last_error())),
        ffi::ERR_STPONE => return Err(SvdLibError::StponeError(last_error())),
        ffi::ERR_IMTQL2 => return Err(SvdLibError::Imtql2Error(last_error())),
        ffi::ERR_LAS2 => return Err(SvdLibError::Las2Error(last_error())),
        ffi::ERR_REFERENCE_PANIC => panic!("{}", last_error()),
        other => return Err(SvdLibError::Las2Error(format!("svdlas2_b200 status {other}: {}", last_error()))),
    }
    let r = unsafe { &*out };
    let d = r.d as usize;
    let take = |p: *const f64, n: usize| -> Vec<f64> {
        if n == 0 { Vec::new() } else { unsafe { std::slice::from_raw_parts(p, n) }.to_vec() }
    };
    let g = &r.diagnostics;
    let rec = SvdRec {
        d,
        ut: Array::from_shape_vec((d, r.ut_cols as usize), take(r.ut, d * r.ut_cols as usize)),
        s: Array::from_shape_vec(d, take(r.s, d)),
        vt: Array::from_shape_vec((d, r.vt_cols as usize), take(r.vt, d * r.vt_cols as usize)),
        diagnostics: Diagnostics {
            non_zero: g.non_zero as usize,
            dimensions: g.dimensions as usize,
            iterations: g.iterations as usize,
            transposed: g.transposed != 0,
            lanczos_steps: g.lanczos_steps as usize,
            ritz_values_stabilized: g.ritz_values_stabilized as usize,
            significant_values: g.significant_values as usize,
            singular_values: g.singular_values as usize,
            end_interval: g.end_interval,
            kappa: g.kappa,
            random_seed: g.random_seed,
        },
    };
    unsafe { ffi::svdlas2_free(out) };
    Ok(SvdRec { d: rec.d, ut: rec.ut?, s: rec.s?, vt: rec.vt?, diagnostics: rec.diagnostics })
}

impl SvdRec {
    /// `ut' * diag(s) * vt`
    pub fn recompose(&self) -> Array2<f64> {
        let sdiag = Array2::from_diag(&self.s);
        self.ut.t().dot(&sdiag).dot(&self.vt)
    }
}

// ---- the three nalgebra-sparse input types ----------------------------------------------------------------

fn gather_lanes(offsets: &[usize], indices: &[usize], values: &[f64], x: &[f64], y: &mut [f64]) {
    for (lane, yv) in y.iter_mut().enumerate() {
        let mut acc = 0.0;
        for k in offsets[lane]..offsets[lane + 1] {
            acc += values[k] * x[indices[k]];
        }
        *yv = acc;
    }
}

fn scatter_lanes(offsets: &[usize], indices: &[usize], values: &[f64], x: &[f64], y: &mut [f64]) {
    y.fill(0.0);
    for (lane, xv) in x.iter().enumerate() {
        for k in offsets[lane]..offsets[lane + 1] {
            y[indices[k]] += values[k] * xv;
        }
    }
}

impl SMat for nalgebra_sparse::csr::CsrMatrix<f64> {
    fn nrows(&self) -> usize { self.nrows() }
    fn ncols(&self) -> usize { self.ncols() }
    fn nnz(&self) -> usize { self.nnz() }
    fn svd_opa(&self, x: &[f64], y: &mut [f64], transposed: bool) {
        let (o, i, v) = self.csr_data();
        if transposed { scatter_lanes(o, i, v, x, y) } else { gather_lanes(o, i, v, x, y) }
    }
    fn raw(&self) -> Option<RawSparse<'_>> {
        let (offsets, indices, values) = self.csr_data();
        Some(RawSparse::Csr { offsets, indices, values })
    }
}

impl SMat for nalgebra_sparse::csc::CscMatrix<f64> {
    fn nrows(&self) -> usize { self.nrows() }
    fn ncols(&self) -> usize { self.ncols() }
    fn nnz(&self) -> usize { self.nnz() }
    fn svd_opa(&self, x: &[f64], y: &mut [f64], transposed: bool) {
        let (o, i, v) = self.csc_data();
        if transposed { gather_lanes(o, i, v, x, y) } else { scatter_lanes(o, i, v, x, y) }
    }
    fn raw(&self) -> Option<RawSparse<'_>> {
        let (offsets, indices, values) = self.csc_data();
        Some(RawSparse::Csc { offsets, indices, values })
    }
}

impl SMat for nalgebra_sparse::coo::CooMatrix<f64> {
    fn nrows(&self) -> usize { self.nrows() }
    fn ncols(&self) -> usize { self.ncols() }
    fn nnz(&self) -> usize { self.nnz() }
    fn svd_opa(&self, x: &[f64], y: &mut [f64], transposed: bool) {
        y.fill(0.0);
        for (i, j, v) in self.triplet_iter() {
            if transposed { y[j] += v * x[i] } else { y[i] += v * x[j] }
        }
    }
    fn raw(&self) -> Option<RawSparse<'_>> {
        Some(RawSparse::Coo {
            rows: self.row_indices().to_vec(),
            cols: self.col_indices().to_vec(),
            values: self.values().to_vec(),
        })
    }
}

/// Not part of the reference crate: gives the registered host buffers and the device scratch the CUDA library keeps
/// cached between calls back to the operating system / the driver.
pub fn trim_caches() {
    unsafe { ffi::svdlas2_trim_caches() };
}
