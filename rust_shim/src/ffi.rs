//! Raw binding of include/svdlas2.h (field order and types must match the header exactly).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct svdlas2_diagnostics {
    pub non_zero: u64,
    pub dimensions: u64,
    pub iterations: u64,
    pub transposed: i32,
    pub lanczos_steps: u64,
    pub ritz_values_stabilized: u64,
    pub significant_values: u64,
    pub singular_values: u64,
    pub end_interval: [f64; 2],
    pub kappa: f64,
    pub random_seed: u32,
}

/// side-channel timings / counters (9 doubles, 11 u64); not surfaced by the crate API
#[repr(C)]
pub struct svdlas2_profile {
    pub seconds: [f64; 9],
    pub counters: [u64; 11],
}

#[repr(C)]
pub struct svdlas2_result {
    pub d: u64,
    pub ut_cols: u64,
    pub vt_cols: u64,
    pub ut: *mut f64,
    pub s: *mut f64,
    pub vt: *mut f64,
    pub diagnostics: svdlas2_diagnostics,
    pub profile: svdlas2_profile,
    pub trace_len: u64,
    pub alf: *mut f64,
    pub bet: *mut f64,
    pub ritz: *mut f64,
    pub bnd: *mut f64,
    pub nside_row_begin: u64,
    pub nside_row_count: u64,
}

pub const OK: c_int = 0;
pub const ERR_IMTQLB: c_int = 1;
pub const ERR_STARTV: c_int = 2;
pub const ERR_STPONE: c_int = 3;
pub const ERR_IMTQL2: c_int = 4;
pub const ERR_LAS2: c_int = 5;
pub const ERR_REFERENCE_PANIC: c_int = 7;

type OneShot = unsafe extern "C" fn(
    nrows: u64, ncols: u64, nnz: u64, a: *const u64, b: *const u64, values: *const f64, dimensions: u64,
    iterations: u64, end_interval: *const f64, kappa: f64, random_seed: u32, out: *mut *mut svdlas2_result,
) -> c_int;

extern "C" {
    pub fn svdlas2_csr(nrows: u64, ncols: u64, nnz: u64, row_offsets: *const u64, col_indices: *const u64,
        values: *const f64, dimensions: u64, iterations: u64, end_interval: *const f64, kappa: f64,
        random_seed: u32, out: *mut *mut svdlas2_result) -> c_int;
    pub fn svdlas2_csc(nrows: u64, ncols: u64, nnz: u64, col_offsets: *const u64, row_indices: *const u64,
        values: *const f64, dimensions: u64, iterations: u64, end_interval: *const f64, kappa: f64,
        random_seed: u32, out: *mut *mut svdlas2_result) -> c_int;
    pub fn svdlas2_coo(nrows: u64, ncols: u64, nnz: u64, row_indices: *const u64, col_indices: *const u64,
        values: *const f64, dimensions: u64, iterations: u64, end_interval: *const f64, kappa: f64,
        random_seed: u32, out: *mut *mut svdlas2_result) -> c_int;
    pub fn svdlas2_free(result: *mut svdlas2_result);
    pub fn svdlas2_last_error() -> *const c_char;
    /// hands the library's cached result buffers and device scratch back (include/svdlas2.h)
    pub fn svdlas2_trim_caches() -> c_int;
}

pub const CSR: OneShot = svdlas2_csr;
pub const CSC: OneShot = svdlas2_csc;
pub const COO: OneShot = svdlas2_coo;
pub type Entry = OneShot;
