//! `GpuFmIndex<'a, P, B, E>`: the reference's `FmIndex` API (sview-fmindex/src/lib.rs:15-28) over libsvfm.so.
//! NOT COMPILED IN THE BUILD IMAGE (no Rust toolchain there).
use core::marker::PhantomData;
use sview_fmindex::{blocks::*, text_encoders::*, Block, LoadError, Position, TextEncoder};
use svfm_sys as sys;

/// Plane count of a block type (`Block2`..`Block6`); BLOCK_LEN comes from the reference's `Block` trait.
pub trait GpuBlock: Block { const PLANES: u32; }
impl<V: Vector> GpuBlock for Block2<V> { const PLANES: u32 = 2; }
impl<V: Vector> GpuBlock for Block3<V> { const PLANES: u32 = 3; }
impl<V: Vector> GpuBlock for Block4<V> { const PLANES: u32 = 4; }
impl<V: Vector> GpuBlock for Block5<V> { const PLANES: u32 = 5; }
impl<V: Vector> GpuBlock for Block6<V> { const PLANES: u32 = 6; }
pub trait GpuEncoder: TextEncoder { const KIND: u32; }
impl GpuEncoder for PassThrough { const KIND: u32 = 0; }
impl GpuEncoder for EncodingTable { const KIND: u32 = 1; }

pub struct GpuFmIndex<'a, P: Position, B: GpuBlock, E: GpuEncoder> {
    handle: *mut sys::svfm_index,
    source_blob: &'a [u8],
    _p: PhantomData<(P, B, E)>,
}
unsafe impl<P: Position, B: GpuBlock, E: GpuEncoder> Send for GpuFmIndex<'_, P, B, E> {}
unsafe impl<P: Position, B: GpuBlock, E: GpuEncoder> Sync for GpuFmIndex<'_, P, B, E> {}

impl<'a, P: Position, B: GpuBlock, E: GpuEncoder> GpuFmIndex<'a, P, B, E> {
    /// `FmIndex::load` (load_from_blob.rs:28-85)
    pub fn load(blob: &'a [u8]) -> Result<Self, LoadError> {
        let t = sys::svfm_type { pos_bits: P::BITS, planes: B::PLANES, vec_bits: B::BLOCK_LEN, encoder: E::KIND };
        let mut handle = core::ptr::null_mut();
        let mut detail = [0u64; 2];
        match unsafe { sys::svfm_load(blob.as_ptr(), blob.len(), t, 0, &mut handle, detail.as_mut_ptr()) } {
            sys::SVFM_OK => Ok(Self { handle, source_blob: blob, _p: PhantomData }),
            sys::SVFM_ERR_BLOB_SIZE => Err(LoadError::MismatchedBlobSize(detail[0] as usize, detail[1] as usize)),
            sys::SVFM_ERR_INVALID_FORMAT => Err(LoadError::InvalidFormat),
            rc => panic!("svfm_load failed with code {rc}"),
        }
    }
    /// `blob()` (reference_to_source_blob.rs:9)
    pub fn blob(&self) -> &'a [u8] { self.source_blob }

    fn pack(patterns: &[&[u8]]) -> (Vec<u8>, Vec<u64>) {
        let mut data = Vec::new();
        let mut offs = vec![0u64];
        for p in patterns { data.extend_from_slice(p); offs.push(data.len() as u64); }
        (data, offs)
    }
    fn count_flags(&self, patterns: &[&[u8]], flags: u32) -> Vec<P> {
        let (data, offs) = Self::pack(patterns);
        let mut out = vec![P::ZERO; patterns.len()];
        let rc = unsafe { sys::svfm_count_batch(self.handle, data.as_ptr(), offs.as_ptr(), patterns.len() as u64, 0, flags,
                                                out.as_mut_ptr() as *mut _) };
        assert!(rc == sys::SVFM_OK, "svfm_count_batch: {rc}"); // empty pattern: the reference panics too
        out
    }
    fn locate_flags(&self, patterns: &[&[u8]], flags: u32) -> (Vec<u64>, Vec<P>) {
        let (data, offs) = Self::pack(patterns);
        let mut out_offs = vec![0u64; patterns.len() + 1];
        let (mut pos, mut total) = (core::ptr::null_mut(), 0u64);
        let rc = unsafe { sys::svfm_locate_batch_alloc(self.handle, data.as_ptr(), offs.as_ptr(), patterns.len() as u64, 0,
                                                       flags, out_offs.as_mut_ptr() as *mut _, &mut pos, &mut total) };
        assert!(rc == sys::SVFM_OK, "svfm_locate_batch: {rc}");
        // no occurrence at all: the library returns a null pointer, which slice::from_raw_parts must never see
        if total == 0 || pos.is_null() { return (out_offs, Vec::new()); }
        let v = unsafe { core::slice::from_raw_parts(pos as *const P, total as usize) }.to_vec();
        unsafe { sys::svfm_free_positions(pos) };
        (out_offs, v)
    }
    /// `count` (locate/with_slice.rs:5-8)
    pub fn count(&self, pattern: &[u8]) -> P { self.count_flags(&[pattern], 0)[0] }
    /// `locate` (locate/with_slice.rs:10-13): SA-row order, like the reference
    pub fn locate(&self, pattern: &[u8]) -> Vec<P> { self.locate_flags(&[pattern], 0).1 }
    /// `locate_to_buffer` (locate/with_slice.rs:15-18): appends
    pub fn locate_to_buffer(&self, pattern: &[u8], buffer: &mut Vec<P>) { buffer.extend(self.locate(pattern)) }
    /// rev-iterator twins (locate/with_rev_iter.rs:5-18)
    pub fn count_rev_iter<I: Iterator<Item = u8>>(&self, it: I) -> P {
        let rev: Vec<u8> = it.collect();
        self.count_flags(&[&rev], sys::SVFM_REVERSED)[0]
    }
    pub fn locate_rev_iter<I: Iterator<Item = u8>>(&self, it: I) -> Vec<P> {
        let rev: Vec<u8> = it.collect();
        self.locate_flags(&[&rev], sys::SVFM_REVERSED).1
    }
    pub fn locate_rev_iter_to_buffer<I: Iterator<Item = u8>>(&self, it: I, buffer: &mut Vec<P>) {
        buffer.extend(self.locate_rev_iter(it))
    }
    /// batched entry points (new): results in input order
    pub fn count_batch(&self, patterns: &[&[u8]]) -> Vec<P> { self.count_flags(patterns, 0) }
    /// CSR: pattern i owns positions[offs[i]..offs[i+1]]
    pub fn locate_batch(&self, patterns: &[&[u8]], sorted: bool) -> (Vec<u64>, Vec<P>) {
        self.locate_flags(patterns, if sorted { sys::SVFM_SORTED } else { 0 })
    }
}
impl<P: Position, B: GpuBlock, E: GpuEncoder> Drop for GpuFmIndex<'_, P, B, E> {
    fn drop(&mut self) { unsafe { sys::svfm_free(self.handle) } }
}
