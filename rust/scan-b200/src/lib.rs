//! scan-b200 -- the normalize -> PCA hot path of scan-rs on B200 GPUs, behind the reference's own entry points.
//!
//! Call sites change from (tools/src/bin/cmd.rs:67-81)
//! ```ignore
//! let norm_mat = normalize(matrix.view(), normalization);
//! let (u, d, v) = BkSvd::new().run_pca(&norm_mat, num_pcs)?;
//! ```
//! to
//! ```ignore
//! let ctx = scan_b200::Context::new(0)?;
//! let dev = scan_b200::DeviceCountMatrix::upload(&ctx, &matrix.view())?;
//! let norm_mat = dev.normalize(normalization, None)?;              // impl DataMat
//! let (u, d, v) = scan_b200::GpuBkSvd::new().run_pca(&norm_mat, num_pcs)?;   // the same PcaResult
//! ```
//! The normalized matrix of the reference is `LowRankOffset<D, impl MatrixMap<u32, f64>>` (normalization.rs:46): the map is an
//! opaque closure chain a backend cannot introspect, so the device path owns both steps and takes the UN-normalized matrix.
//!
//! NOT COMPILED where it was written (no rustc in that image).  The C ABI underneath is exercised by the C++ layer
//! (include/scanb200.hpp, tests/cpp/host_api_test.cpp) and the Python layer (scan_rs_b200/*.py); `src/ffi.rs` is generated from
//! include/scanb200.h.
pub mod ffi;

use anyhow::{format_err, Error};
use ndarray::{Array1, Array2};
use scan_rs::dim_red::{DataMat, Pca, PcaResult};
use scan_rs::normalization::Normalization;
use snoop::{CancelProgress, CancellationError};
use sqz::{AdaptiveMat, AdaptiveVec, MatrixMap};
use std::ffi::{c_void, CStr};
use std::marker::PhantomData;
use std::ops::Deref;

fn check(rc: i32) -> Result<(), Error> {
    if rc == ffi::SB_OK {
        return Ok(());
    }
    if rc == ffi::SB_ERR_CANCELLED {
        return Err(CancellationError.into()); // snoop/src/lib.rs:5-18
    }
    let msg = unsafe { CStr::from_ptr(ffi::sb_last_error()) }.to_string_lossy().into_owned();
    Err(format_err!("{msg}")) // "The input matrix must be at least 2x2." / "invalid k" verbatim (bk_svd.rs:74,78)
}

/// One GPU + stream.  `!Send`: a context is driven from the thread that uses it.
pub struct Context {
    h: *mut ffi::sb_ctx,
    owned: bool,
}

impl Context {
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::sb_init(device, &mut h) })?; // no CPU fallback: SB_ERR_CUDA without a device
        Ok(Self { h, owned: true })
    }
    pub fn set_option(&self, name: &str, value: f64) -> Result<(), Error> {
        let c = std::ffi::CString::new(name)?;
        check(unsafe { ffi::sb_set_option(self.h, c.as_ptr(), value) })
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        if self.owned {
            unsafe { ffi::sb_shutdown(self.h) }
        }
    }
}

/// `AdaptiveMat<u32>` resident on the device (cell-major + the layouts both products stream).
pub struct DeviceCountMatrix<'c> {
    h: *mut ffi::sb_mat,
    rows: usize,
    cols: usize,
    _ctx: PhantomData<&'c Context>,
}

impl<'c> DeviceCountMatrix<'c> {
    /// Decode every AdaptiveVec once (`foreach`, vec.rs:1230-1273) into u64 indptr / u32 index / u32 count and upload.
    pub fn upload<D, M>(ctx: &'c Context, mat: &AdaptiveMat<u32, D, M>) -> Result<Self, Error>
    where
        D: Deref<Target = [AdaptiveVec]>,
        M: MatrixMap<u32, u32>,
    {
        let (mut indptr, mut idx, mut val) = (vec![0u64], Vec::<u32>::new(), Vec::<u32>::new());
        for (i, v) in mat.iter().enumerate() {
            // rows if CSR, columns if CSC (mat.rs:178-187)
            v.foreach(|j, x| {
                let x = mat.get_map().map(x, i, j);
                if x != 0 {
                    idx.push(j as u32);
                    val.push(x);
                }
            });
            indptr.push(idx.len() as u64);
        }
        let major = if mat.is_csr() { ffi::SB_GENE_MAJOR } else { ffi::SB_CELL_MAJOR };
        let mut h = std::ptr::null_mut();
        check(unsafe {
            ffi::sb_upload(ctx.h, major, mat.rows() as u32, mat.cols() as u64, indptr.as_ptr(), idx.as_ptr(), val.as_ptr(), &mut h)
        })?;
        Ok(Self { h, rows: mat.rows(), cols: mat.cols(), _ctx: PhantomData })
    }

    /// The same walk into the narrow host form (cell-major, at most 65,536 genes): u16 gene + u8 count, counts >= 255 in a side
    /// list -- 3 bytes per entry cross PCIe instead of 8, and the copy is the largest part of an end-to-end call.
    pub fn upload_compact<D, M>(ctx: &'c Context, mat: &AdaptiveMat<u32, D, M>) -> Result<Self, Error>
    where
        D: Deref<Target = [AdaptiveVec]>,
        M: MatrixMap<u32, u32>,
    {
        assert!(!mat.is_csr() && mat.rows() <= 65536);
        let (mut indptr, mut idx, mut cnt) = (vec![0u64], Vec::<u16>::new(), Vec::<u8>::new());
        let (mut big_pos, mut big_cnt) = (Vec::<u64>::new(), Vec::<u32>::new());
        for (c, v) in mat.iter().enumerate() {
            v.foreach(|g, x| {
                let x = mat.get_map().map(x, g, c);
                if x != 0 {
                    if x >= 255 {
                        big_pos.push(idx.len() as u64);
                        big_cnt.push(x);
                    }
                    idx.push(g as u16);
                    cnt.push(x.min(255) as u8);
                }
            });
            indptr.push(idx.len() as u64);
        }
        let mut h = std::ptr::null_mut();
        check(unsafe {
            ffi::sb_upload_compact(ctx.h, mat.rows() as u32, mat.cols() as u64, indptr.as_ptr(), idx.as_ptr(), cnt.as_ptr(),
                                   big_pos.len() as u64, big_pos.as_ptr(), big_cnt.as_ptr(), &mut h)
        })?;
        Ok(Self { h, rows: mat.rows(), cols: mat.cols(), _ctx: PhantomData })
    }

    /// The same walk into the packed host form (cell-major, any gene count): one byte of gene delta + one nibble of count per
    /// entry, escapes in two side lists -- about 1.6 bytes per entry cross PCIe (the compact form: 3, the plain one: 8).  The
    /// layout is stated beside `sb_upload_packed` in include/scanb200.h; `sb_pack_csc_fill` builds it from plain arrays.
    pub fn upload_packed<D, M>(ctx: &'c Context, mat: &AdaptiveMat<u32, D, M>) -> Result<Self, Error>
    where
        D: Deref<Target = [AdaptiveVec]>,
        M: MatrixMap<u32, u32>,
    {
        assert!(!mat.is_csr());
        let (mut indptr, mut dgene, mut cnt4) = (vec![0u64], Vec::<u8>::new(), Vec::<u8>::new());
        let (mut esc_pos, mut esc_gene) = (Vec::<u64>::new(), Vec::<u32>::new());
        let (mut big_pos, mut big_cnt) = (Vec::<u64>::new(), Vec::<u32>::new());
        for (c, v) in mat.iter().enumerate() {
            let mut prev: i64 = -1;
            v.foreach(|g, x| {
                let x = mat.get_map().map(x, g, c);
                if x != 0 {
                    let pos = dgene.len() as u64;
                    let d = g as i64 - prev;
                    if d >= 1 && d <= 255 {
                        dgene.push(d as u8);
                    } else {
                        dgene.push(0);
                        esc_pos.push(pos);
                        esc_gene.push(g as u32);
                    }
                    let nib = if x >= 15 {
                        big_pos.push(pos);
                        big_cnt.push(x);
                        15u8
                    } else {
                        x as u8
                    };
                    if pos % 2 == 0 {
                        cnt4.push(nib);
                    } else {
                        *cnt4.last_mut().unwrap() |= nib << 4;
                    }
                    prev = g as i64;
                }
            });
            indptr.push(dgene.len() as u64);
        }
        let mut h = std::ptr::null_mut();
        check(unsafe {
            ffi::sb_upload_packed(ctx.h, mat.rows() as u32, mat.cols() as u64, indptr.as_ptr(), dgene.as_ptr(), cnt4.as_ptr(),
                                  esc_pos.len() as u64, esc_pos.as_ptr(), esc_gene.as_ptr(),
                                  big_pos.len() as u64, big_pos.as_ptr(), big_cnt.as_ptr(), &mut h)
        })?;
        Ok(Self { h, rows: mat.rows(), cols: mat.cols(), _ctx: PhantomData })
    }

    pub fn shape(&self) -> [usize; 2] {
        [self.rows, self.cols]
    }

    /// sum_axis::<u32>(Axis(0)) (mat.rs:377-406): per-cell UMI totals
    pub fn cell_totals(&self) -> Result<Array1<u32>, Error> {
        let mut out = Array1::<u32>::zeros(self.cols);
        check(unsafe { ffi::sb_cell_totals(self.h, out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// normalize / normalize_with_size_factor (normalization.rs:46-102); the enum order is SB_NORM_*
    pub fn normalize(&self, norm: Normalization, size_factors: Option<&Array1<u32>>) -> Result<DeviceNormalized<'_>, Error> {
        let mut h = std::ptr::null_mut();
        let sf = size_factors.map_or(std::ptr::null(), |a| a.as_ptr());
        check(unsafe { ffi::sb_normalize(self.h, norm as i32, sf, &mut h) })?;
        Ok(DeviceNormalized { h, mat: self })
    }

    /// mean_var_axis(Axis(1)) of the SizeNormalized view (diff-exp/src/diff_exp.rs:458-472)
    pub fn mean_var_genes(&self, size_factors: Option<&[f64]>) -> Result<(Array1<f64>, Array1<f64>), Error> {
        let (mut mean, mut var) = (Array1::<f64>::zeros(self.rows), Array1::<f64>::zeros(self.rows));
        let sf = size_factors.map_or(std::ptr::null(), |a| a.as_ptr());
        check(unsafe { ffi::sb_mean_var_axis(self.h, 1, sf, mean.as_mut_ptr(), var.as_mut_ptr()) })?;
        Ok((mean, var))
    }
}
impl Drop for DeviceCountMatrix<'_> {
    fn drop(&mut self) {
        unsafe { ffi::sb_free_mat(self.h) }
    }
}

/// `LowRankOffset<D, impl MatrixMap<u32, f64>>` on the device: counts + fused map + rank-1 offset, never materialised.
pub struct DeviceNormalized<'a> {
    h: *mut ffi::sb_nmat,
    mat: &'a DeviceCountMatrix<'a>,
}
impl DataMat for DeviceNormalized<'_> {
    fn shape(&self) -> [usize; 2] {
        self.mat.shape()
    }
}
impl Drop for DeviceNormalized<'_> {
    fn drop(&mut self) {
        unsafe { ffi::sb_free_nmat(self.h) } // before the matrix it borrows: guaranteed by the lifetime
    }
}

unsafe extern "C" fn progress_trampoline<S: CancelProgress>(fraction: f64, user: *mut c_void) -> i32 {
    let s = &mut *(user as *mut S);
    s.set_progress_check(fraction).is_err() as i32 // Err(CancellationError) -> stop at this milestone
}

/// A local struct (not `BkSvd`): the blanket `impl<T> Pca<T, f64> for BkSvd` (bk_svd.rs:41-47) must not be overlapped.
pub struct GpuBkSvd {
    pub k_multiplier: f64,
    pub n_iter: usize,
}
impl GpuBkSvd {
    pub fn new() -> Self {
        Self { k_multiplier: 2.0, n_iter: 5 } // BkSvd::new (bk_svd.rs:24-31)
    }
}
impl Default for GpuBkSvd {
    fn default() -> Self {
        Self::new()
    }
}
impl<'a> Pca<DeviceNormalized<'a>, f64> for GpuBkSvd {
    fn run_pca_cancellable(&self, a: &DeviceNormalized<'a>, k: usize, mut snoop: impl CancelProgress) -> Result<PcaResult, Error> {
        let [m, n] = a.shape();
        let (mut u, mut s, mut v) = (Array2::<f64>::zeros((m, k)), Array1::<f64>::zeros(k), Array2::<f64>::zeros((n, k)));
        check(unsafe {
            ffi::sb_bksvd_run_pca(a.h, k as u32, self.k_multiplier, self.n_iter as u32, Some(progress_trampoline::<_>),
                                  &mut snoop as *mut _ as *mut c_void, u.as_mut_ptr(), s.as_mut_ptr(), v.as_mut_ptr())
        })?;
        Ok((u, s, v)) // v is already vt.reversed_axes() (bk_svd.rs:51)
    }
}

/// RandSvd (rand_svd.rs:13-50): ignores the snoop, as the reference does.
pub struct GpuRandSvd {
    pub l_multiplier: f64,
    pub n_iter: usize,
}
impl GpuRandSvd {
    pub fn new() -> Self {
        Self { l_multiplier: 10.0, n_iter: 2 }
    }
}
impl<'a> Pca<DeviceNormalized<'a>, f64> for GpuRandSvd {
    fn run_pca_cancellable(&self, a: &DeviceNormalized<'a>, k: usize, _snoop: impl CancelProgress) -> Result<PcaResult, Error> {
        let [m, n] = a.shape();
        let (mut u, mut s, mut v) = (Array2::<f64>::zeros((m, k)), Array1::<f64>::zeros(k), Array2::<f64>::zeros((n, k)));
        check(unsafe {
            ffi::sb_randsvd_run_pca(a.h, k as u32, self.l_multiplier, self.n_iter as u32, u.as_mut_ptr(), s.as_mut_ptr(), v.as_mut_ptr())
        })?;
        Ok((u, s, v))
    }
}

/// Irlba (irlba.rs:36-69).  The reference cannot run IRLBA on a LowRankOffset (no 1-D `Dot`); the device type can.
pub struct GpuIrlba {
    pub tol: f64,
    pub max_iter: usize,
}
impl GpuIrlba {
    pub fn new() -> Self {
        Self { tol: 0.0001, max_iter: 50 }
    }
}
impl<'a> Pca<DeviceNormalized<'a>, f64> for GpuIrlba {
    fn run_pca_cancellable(&self, a: &DeviceNormalized<'a>, k: usize, mut snoop: impl CancelProgress) -> Result<PcaResult, Error> {
        let [m, n] = a.shape();
        let (mut u, mut s, mut v) = (Array2::<f64>::zeros((m, k)), Array1::<f64>::zeros(k), Array2::<f64>::zeros((n, k)));
        check(unsafe {
            ffi::sb_irlba(a.h, k as u32, self.tol, self.max_iter as u32, std::ptr::null(), Some(progress_trampoline::<_>),
                          &mut snoop as *mut _ as *mut c_void, u.as_mut_ptr(), s.as_mut_ptr(), v.as_mut_ptr(),
                          std::ptr::null_mut(), std::ptr::null_mut())
        })?;
        Ok((u, s, v))
    }
}

/// Every GPU of the box from one caller thread (the reference's only caller is a single process).  `run` executes the
/// closure once per rank, concurrently, each on its own library worker thread with its own `Context`; inside it a rank uploads
/// its contiguous cell shard and makes the ordinary calls -- the collectives meet across the workers.
pub struct MultiGpu {
    h: *mut ffi::sb_multi,
    n: usize,
}
impl MultiGpu {
    pub fn new(devices: &[i32]) -> Result<Self, Error> {
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::sb_multi_init(devices.len() as i32, devices.as_ptr(), &mut h) })?;
        Ok(Self { h, n: devices.len() })
    }
    pub fn size(&self) -> usize {
        self.n
    }
    pub fn run<F>(&self, f: F) -> Result<(), Error>
    where
        F: Fn(usize, &Context) -> Result<(), Error> + Sync,
    {
        unsafe extern "C" fn tramp<F: Fn(usize, &Context) -> Result<(), Error> + Sync>(rank: i32, ctx: *mut ffi::sb_ctx, user: *mut c_void) -> i32 {
            let f = &*(user as *const F);
            let c = Context { h: ctx, owned: false };
            match std::panic::catch_unwind(std::panic::AssertUnwindSafe(|| f(rank as usize, &c))) {
                Ok(Ok(())) => ffi::SB_OK,
                Ok(Err(e)) => if e.is::<CancellationError>() { ffi::SB_ERR_CANCELLED } else { ffi::SB_ERR_INVALID_ARG },
                Err(_) => ffi::SB_ERR_INVALID_ARG, // no panic crosses the ABI
            }
        }
        check(unsafe { ffi::sb_multi_run(self.h, Some(tramp::<F>), &f as *const F as *mut c_void) })
    }
}
impl Drop for MultiGpu {
    fn drop(&mut self) {
        unsafe { ffi::sb_multi_shutdown(self.h) }
    }
}
