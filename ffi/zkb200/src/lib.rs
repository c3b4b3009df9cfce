//! Thin shim keeping the reference's module paths and signatures on top of `libzkb200.so`.
//!
//! NOT COMPILED in the build container (no Rust toolchain there): it is the binding a
//! maintainer adds, kept deliberately thin; the same call sequence is exercised from Python
//! (`zk-research-implementations_b200/*.py`) by the parity tests.
//!
//! `F` must be one of the three 4-limb Montgomery fields the library instantiates.  The element <-> limb
//! cast relies on ark-ff 0.5's layout `Fp<MontBackend<C, 4>, 4>(BigInt<4>([u64; 4]), PhantomData)`.
pub mod field;
pub mod fiat_shamir;
pub mod gkr;
pub mod multilinear_polynomial;
pub mod sum_check_protocol;
pub mod univariate_polynomial;

use std::cell::RefCell;
use zkb200_sys as sys;

/// One device context per thread and field (the library is not internally synchronised).
pub struct Ctx(pub *mut sys::zkb_ctx);
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { sys::zkb_ctx_destroy(self.0) };
    }
}
thread_local! { static CTX: RefCell<Vec<(i32, i32, std::rc::Rc<Ctx>)>> = RefCell::new(Vec::new()); }

pub fn ctx(field_id: i32, mode: i32) -> std::rc::Rc<Ctx> {
    CTX.with(|c| {
        let mut c = c.borrow_mut();
        if let Some((_, _, x)) = c.iter().find(|(f, m, _)| *f == field_id && *m == mode) {
            return x.clone();
        }
        let mut raw = std::ptr::null_mut();
        let device = std::env::var("ZKB200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        check(std::ptr::null_mut(), unsafe { sys::zkb_ctx_create(field_id, device, mode, &mut raw) });
        let x = std::rc::Rc::new(Ctx(raw));
        c.push((field_id, mode, x.clone()));
        x
    })
}

/// Status -> the reference's panic strings (nothing unwinds across the C boundary).
pub fn check(ctx: *mut sys::zkb_ctx, status: i32) {
    if status == sys::ZKB_OK {
        return;
    }
    let msg = unsafe { std::ffi::CStr::from_ptr(sys::zkb_strerror(status)) }.to_string_lossy().into_owned();
    match status {
        // "Invalid evaluations", "Invalid number of values", "all evaluations must have same length",
        // "all product polys must have same degree": byte-identical to the reference's panic!() messages
        sys::ZKB_ERR_NOT_POW2 | sys::ZKB_ERR_ARITY | sys::ZKB_ERR_LENGTH_MISMATCH | sys::ZKB_ERR_DEGREE_MISMATCH => panic!("{msg}"),
        _ => {
            let detail = if ctx.is_null() {
                String::new()
            } else {
                unsafe { std::ffi::CStr::from_ptr(sys::zkb_ctx_last_error(ctx)) }.to_string_lossy().into_owned()
            };
            panic!("zkb200: {msg} {detail}")
        }
    }
}
