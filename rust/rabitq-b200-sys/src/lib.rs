//! Raw bindings to `include/rbq.h` plus a thin safe wrapper with the shape of
//! `rabitq_rs::IvfRabitqIndex::{load_from_path, search, batch_search}` (reference src/ivf.rs:1477-1752).
//! Source only -- not compiled in the build image (no cargo/rustc).
use std::ffi::{c_char, c_int, c_void, CStr, CString};

#[repr(C)]
pub struct rbq_index { _private: [u8; 0] }

extern "C" {
    pub fn rbq_last_error() -> *const c_char;
    pub fn rbq_index_load(path: *const c_char, device: c_int, shard_rank: c_int, shard_count: c_int, out: *mut *mut rbq_index) -> c_int;
    pub fn rbq_index_free(ix: *mut rbq_index);
    pub fn rbq_index_len(ix: *const rbq_index) -> usize;
    pub fn rbq_index_dim(ix: *const rbq_index) -> usize;
    pub fn rbq_index_cluster_count(ix: *const rbq_index) -> usize;
    pub fn rbq_search_batch(ix: *const rbq_index, queries: *const f32, nq: usize, dim: usize, top_k: usize, nprobe: usize,
                            ids: *mut u64, scores: *mut f32, counts: *mut u32) -> c_int;
    pub fn rbq_search_batch_filtered(ix: *const rbq_index, queries: *const f32, nq: usize, dim: usize, top_k: usize, nprobe: usize,
                                     filter_bits: *const u64, filter_nbits: usize, ids: *mut u64, scores: *mut f32, counts: *mut u32) -> c_int;
    pub fn rbq_index_save(ix: *const rbq_index, path: *const c_char) -> c_int;
    pub fn rbq_search_batch_device(ix: *const rbq_index, d_queries: *const f32, nq: usize, dim: usize, top_k: usize, nprobe: usize,
                                   d_filter: *const u64, filter_nbits: usize, d_ids: *mut u64, d_scores: *mut f32, d_counts: *mut u32,
                                   stream: *mut c_void) -> c_int;
    pub fn rbq_fetch_embedding(ix: *const rbq_index, vector_id: u64, out: *mut f32, found: *mut c_int) -> c_int;
    // multi-GPU, three phases (include/rbq.h): the caller runs the NCCL exchanges in between
    pub fn rbq_dist_front(ix: *const rbq_index, d_queries: *const f32, nq: usize, dim: usize, top_k: usize, nprobe: usize,
                          q_begin: usize, q_count: usize, d_probes: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn rbq_dist_head(ix: *const rbq_index, nq: usize, top_k: usize, nprobe: usize, d_probes: *const c_void, d_tau: *mut f32,
                         d_ids: *mut u64, d_scores: *mut f32, d_counts: *mut u32, stream: *mut c_void) -> c_int;
    pub fn rbq_dist_tail(ix: *const rbq_index, nq: usize, top_k: usize, nprobe: usize, d_tau: *const f32, d_ids: *mut u64,
                         d_scores: *mut f32, d_counts: *mut u32, stream: *mut c_void) -> c_int;
    pub fn rbq_merge_topk_packed_device(ix: *const rbq_index, nshards: c_int, nq: usize, top_k: usize, packed: *const c_void,
                                        chunk_bytes: usize, out_ids: *mut u64, out_scores: *mut f32, out_counts: *mut u32,
                                        stream: *mut c_void) -> c_int;
}

/// Mirrors `rabitq_rs::RabitqError` (reference src/lib.rs:39-57); codes are `rbq_status`.
#[derive(Debug)]
pub enum RabitqError {
    DimensionMismatch(String), InvalidConfig(String), EmptyIndex, Io(String), InvalidPersistence(String), Cuda(String),
}
fn check(rc: c_int) -> Result<(), RabitqError> {
    if rc == 0 { return Ok(()); }
    let msg = unsafe { CStr::from_ptr(rbq_last_error()) }.to_string_lossy().into_owned();
    Err(match rc { 1 => RabitqError::DimensionMismatch(msg), 2 => RabitqError::InvalidConfig(msg), 3 => RabitqError::EmptyIndex,
                   4 => RabitqError::Io(msg), 5 => RabitqError::InvalidPersistence(msg), _ => RabitqError::Cuda(msg) })
}

#[derive(Debug, Clone, Copy)]
pub struct SearchParams { pub top_k: usize, pub nprobe: usize }
#[derive(Debug, Clone, PartialEq)]
pub struct SearchResult { pub id: usize, pub score: f32 }

/// Device-resident index; same surface as the CPU `IvfRabitqIndex` for the search path.
pub struct IvfRabitqIndex { h: *mut rbq_index }
unsafe impl Send for IvfRabitqIndex {}
unsafe impl Sync for IvfRabitqIndex {} // calls on one handle are serialised inside librbq

impl IvfRabitqIndex {
    pub fn load_from_path<P: AsRef<std::path::Path>>(path: P) -> Result<Self, RabitqError> {
        let c = CString::new(path.as_ref().to_string_lossy().as_bytes()).unwrap();
        let mut h = std::ptr::null_mut();
        check(unsafe { rbq_index_load(c.as_ptr(), 0, 0, 1, &mut h) })?;
        Ok(Self { h })
    }
    pub fn len(&self) -> usize { unsafe { rbq_index_len(self.h) } }
    pub fn is_empty(&self) -> bool { self.len() == 0 }
    pub fn cluster_count(&self) -> usize { unsafe { rbq_index_cluster_count(self.h) } }

    /// `IvfRabitqIndex::fetch_embedding` (reference src/ivf.rs:1247-1307).
    pub fn fetch_embedding(&self, vector_id: usize) -> Option<Vec<f32>> {
        let dim = unsafe { rbq_index_dim(self.h) };
        let (mut out, mut found) = (vec![0.0f32; dim], 0 as c_int);
        let rc = unsafe { rbq_fetch_embedding(self.h, vector_id as u64, out.as_mut_ptr(), &mut found) };
        if rc == 0 && found != 0 { Some(out) } else { None }
    }

    pub fn batch_search(&self, queries: &[&[f32]], params: SearchParams) -> Vec<Result<Vec<SearchResult>, RabitqError>> {
        let dim = unsafe { rbq_index_dim(self.h) };
        if let Some(bad) = queries.iter().find(|q| q.len() != dim) {
            let msg = format!("dimension mismatch: expected {dim}, got {}", bad.len());
            return queries.iter().map(|_| Err(RabitqError::DimensionMismatch(msg.clone()))).collect();
        }
        let (nq, k) = (queries.len(), params.top_k);
        let flat: Vec<f32> = queries.iter().flat_map(|q| q.iter().copied()).collect();
        let (mut ids, mut scores, mut counts) = (vec![0u64; nq * k.max(1)], vec![0f32; nq * k.max(1)], vec![0u32; nq]);
        if let Err(e) = check(unsafe { rbq_search_batch(self.h, flat.as_ptr(), nq, dim, k, params.nprobe,
                                                        ids.as_mut_ptr(), scores.as_mut_ptr(), counts.as_mut_ptr()) }) {
            let msg = format!("{e:?}");
            return (0..nq).map(|_| Err(RabitqError::Cuda(msg.clone()))).collect();
        }
        (0..nq).map(|q| Ok((0..counts[q] as usize).map(|i| SearchResult { id: ids[q * k + i] as usize, score: scores[q * k + i] }).collect())).collect()
    }
    pub fn search(&self, query: &[f32], params: SearchParams) -> Result<Vec<SearchResult>, RabitqError> {
        self.batch_search(&[query], params).pop().unwrap()
    }
}
impl Drop for IvfRabitqIndex { fn drop(&mut self) { unsafe { rbq_index_free(self.h) } } }
