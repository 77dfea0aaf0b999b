//! Same module paths / names as `huff_coding::prelude` for `L = u8`, backed by the CUDA library.
//! Reference panics are reproduced with the reference's messages; `Err` types carry the same data.
//! NOTE: written against include/huffb200.h but never compiled in this repository (no rustc available).
use huffb200_sys as sys;
use std::collections::HashMap;

thread_local! {
    // one hb_ctx per calling thread (an hb_ctx is not thread-safe)
    static CTX: *mut sys::hb_ctx = {
        let mut p: *mut sys::hb_ctx = std::ptr::null_mut();
        let rc = unsafe { sys::hb_ctx_create(0, &mut p) };
        assert!(rc == sys::HB_OK, "libhuffb200: no usable CUDA device (there is no CPU fallback)");
        p
    };
}
fn ctx() -> *mut sys::hb_ctx { CTX.with(|c| *c) }

/// `huff_coding::weights::build_weights_map` for u8 (weights.rs:82-84)
pub fn build_weights_map(letters: &[u8]) -> HashMap<u8, usize> {
    let mut w = [0u64; 256];
    let rc = unsafe { sys::hb_histogram_u8(ctx(), letters.as_ptr(), letters.len(), w.as_mut_ptr()) };
    assert!(rc == sys::HB_OK);
    (0..256usize).filter(|&b| w[b] != 0).map(|b| (b as u8, w[b] as usize)).collect()
}

/// `huff_coding::tree::HuffTree<u8>` (tree_inner.rs:193-196)
#[derive(Clone)]
pub struct HuffTree { raw: Box<sys::hb_tree> }

impl HuffTree {
    /// tree_inner.rs:281-320.  Leaves are inserted in ascending letter order (the canonical order; the reference
    /// inserts in HashMap iteration order, which is random per process).
    pub fn from_weights(weights: HashMap<u8, usize>) -> Self {
        let mut w = [0u64; 256];
        for (l, f) in weights { w[l as usize] = f as u64; }
        let mut raw: Box<sys::hb_tree> = unsafe { Box::new(std::mem::zeroed()) };
        let rc = unsafe { sys::hb_tree_from_weights(w.as_ptr(), sys::HB_ORDER_ASC, &mut *raw) };
        if rc == sys::HB_ERR_EMPTY_WEIGHTS { panic!("provided empty weights") }
        assert!(rc == sys::HB_OK);
        HuffTree { raw }
    }
}

/// `huff_coding::comp::CompressData<u8>` (comp.rs:41-89)
pub struct CompressData { comp_bytes: Vec<u8>, padding_bits: u8, huff_tree: HuffTree }

impl CompressData {
    pub fn new(comp_bytes: Vec<u8>, padding_bits: u8, huff_tree: HuffTree) -> Self {
        if comp_bytes.is_empty() { panic!("provided comp_bytes are empty") }
        if padding_bits > 7 { panic!("padding bits cannot be larger than 7") }
        CompressData { comp_bytes, padding_bits, huff_tree }
    }
    pub fn comp_bytes(&self) -> &[u8] { &self.comp_bytes }
    pub fn padding_bits(&self) -> u8 { self.padding_bits }
    pub fn huff_tree(&self) -> &HuffTree { &self.huff_tree }
    pub fn into_inner(self) -> (Vec<u8>, u8, HuffTree) { (self.comp_bytes, self.padding_bits, self.huff_tree) }
}

/// comp.rs:561-590
#[derive(Debug, Clone)]
pub struct CompressError { message: &'static str, missing_letter: u8 }
impl CompressError {
    pub fn message(&self) -> &str { self.message }
    pub fn missing_letter(&self) -> &u8 { &self.missing_letter }
}

unsafe fn take(ptr: *mut u8, n: usize) -> Vec<u8> {
    let v = std::slice::from_raw_parts(ptr, n).to_vec();
    sys::hb_free(ptr as *mut _);
    v
}

/// comp.rs:353-356
pub fn compress(letters: &[u8]) -> CompressData {
    let mut raw: Box<sys::hb_tree> = unsafe { Box::new(std::mem::zeroed()) };
    let (mut p, mut n, mut pad) = (std::ptr::null_mut(), 0usize, 0u8);
    let rc = unsafe { sys::hb_compress_u8(ctx(), letters.as_ptr(), letters.len(), sys::HB_ORDER_ASC, &mut *raw, &mut p, &mut n, &mut pad) };
    if rc == sys::HB_ERR_EMPTY_WEIGHTS { panic!("provided empty weights") }
    assert!(rc == sys::HB_OK);
    CompressData::new(unsafe { take(p, n) }, pad, HuffTree { raw })
}

/// comp.rs:419-451
pub fn compress_with_tree(letters: &[u8], huff_tree: HuffTree) -> Result<CompressData, CompressError> {
    let (mut p, mut n, mut pad, mut missing) = (std::ptr::null_mut(), 0usize, 0u8, 0u8);
    let rc = unsafe { sys::hb_compress_with_tree_u8(ctx(), letters.as_ptr(), letters.len(), &*huff_tree.raw, &mut p, &mut n, &mut pad, &mut missing) };
    match rc {
        sys::HB_OK => Ok(CompressData::new(unsafe { take(p, n) }, pad, huff_tree)),
        sys::HB_ERR_MISSING_LETTER => Err(CompressError { message: "letter not found in codes", missing_letter: missing }),
        sys::HB_ERR_EMPTY_COMP => panic!("provided comp_bytes are empty"),
        _ => panic!("libhuffb200 status {}", rc),
    }
}

/// comp.rs:487-519
pub fn decompress(comp_data: &CompressData) -> Vec<u8> {
    let (mut p, mut n) = (std::ptr::null_mut(), 0usize);
    let rc = unsafe { sys::hb_decompress_u8(ctx(), comp_data.comp_bytes.as_ptr(), comp_data.comp_bytes.len(), comp_data.padding_bits, &*comp_data.huff_tree.raw, &mut p, &mut n) };
    assert!(rc == sys::HB_OK);
    unsafe { take(p, n) }
}
