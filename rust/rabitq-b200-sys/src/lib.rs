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
    // multi-GPU as ONE call: librbq owns the NCCL communicator (include/rbq.h)
    pub fn rbq_comm_unique_id(id_out: *mut u8) -> c_int;
    pub fn rbq_comm_init(ix: *mut rbq_index, id: *const u8, rank: c_int, world: c_int) -> c_int;
    pub fn rbq_comm_destroy(ix: *mut rbq_index) -> c_int;
    pub fn rbq_set_exact_merge(ix: *mut rbq_index, on: c_int) -> c_int;
    pub fn rbq_search_batch_sharded(ix: *const rbq_index, queries: *const f32, nq: usize, dim: usize, top_k: usize, nprobe: usize,
                                    ids: *mut u64, scores: *mut f32, counts: *mut u32) -> c_int;
    pub fn rbq_search_batch_sharded_device(ix: *const rbq_index, d_queries: *const f32, nq: usize, dim: usize, top_k: usize,
                                           nprobe: usize, d_ids: *mut u64, d_scores: *mut f32, d_counts: *mut u32,
                                           stream: *mut c_void) -> c_int;
}
pub const RBQ_COMM_ID_BYTES: usize = 128;

/// Mirrors `rabitq_rs::RabitqError` (reference src/lib.rs:39-57); codes are `rbq_status`.
#[derive(Debug)]
pub enum RabitqError {
    DimensionMismatch(String), InvalidConfig(String), EmptyIndex, Io(String), InvalidPersistence(String), Cuda(String),
}
impl Clone for RabitqError {
    fn clone(&self) -> Self {
        match self {
            Self::DimensionMismatch(m) => Self::DimensionMismatch(m.clone()), Self::InvalidConfig(m) => Self::InvalidConfig(m.clone()),
            Self::EmptyIndex => Self::EmptyIndex, Self::Io(m) => Self::Io(m.clone()),
            Self::InvalidPersistence(m) => Self::InvalidPersistence(m.clone()), Self::Cuda(m) => Self::Cuda(m.clone()),
        }
    }
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
pub struct IvfRabitqIndex { h: *mut rbq_index, sharded: bool }
unsafe impl Send for IvfRabitqIndex {}
unsafe impl Sync for IvfRabitqIndex {} // calls on one handle are serialised inside librbq (mutex on the host, an event on the device)

impl IvfRabitqIndex {
    pub fn load_from_path<P: AsRef<std::path::Path>>(path: P) -> Result<Self, RabitqError> {
        let c = CString::new(path.as_ref().to_string_lossy().as_bytes()).unwrap();
        let mut h = std::ptr::null_mut();
        check(unsafe { rbq_index_load(c.as_ptr(), 0, 0, 1, &mut h) })?;
        Ok(Self { h, sharded: false })
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

    /// One shard of the index on `device` (lists assigned size-balanced; centroids and rotator replicated).
    pub fn load_shard<P: AsRef<std::path::Path>>(path: P, device: i32, shard_rank: i32, shard_count: i32) -> Result<Self, RabitqError> {
        let c = CString::new(path.as_ref().to_string_lossy().as_bytes()).unwrap();
        let mut h = std::ptr::null_mut();
        check(unsafe { rbq_index_load(c.as_ptr(), device, shard_rank, shard_count, &mut h) })?;
        Ok(Self { h, sharded: false })
    }
    /// Rank 0: the rendezvous id to ship to the other ranks (any transport).
    pub fn comm_unique_id() -> Result<[u8; RBQ_COMM_ID_BYTES], RabitqError> {
        let mut id = [0u8; RBQ_COMM_ID_BYTES];
        check(unsafe { rbq_comm_unique_id(id.as_mut_ptr()) })?;
        Ok(id)
    }
    /// Collective over all shards; afterwards `batch_search` runs the sharded search (every rank passes the same batch and
    /// receives the merged result).  `exact` = bit-identical to the single-GPU answer.
    pub fn comm_init(&mut self, id: &[u8; RBQ_COMM_ID_BYTES], rank: i32, world: i32, exact: bool) -> Result<(), RabitqError> {
        check(unsafe { rbq_comm_init(self.h, id.as_ptr(), rank, world) })?;
        check(unsafe { rbq_set_exact_merge(self.h, exact as c_int) })?;
        self.sharded = true;
        Ok(())
    }

    /// `batch_search` (reference src/ivf.rs:1743-1752): one Result per query, like the reference's per-query `search` calls --
    /// a query of the wrong length fails alone (`DimensionMismatch`), the others are searched; a failing device call
    /// gives every searched query a copy of the error the C ABI reported (same variant as `rbq_status`).
    pub fn batch_search(&self, queries: &[&[f32]], params: SearchParams) -> Vec<Result<Vec<SearchResult>, RabitqError>> {
        let dim = unsafe { rbq_index_dim(self.h) };
        let good: Vec<usize> = (0..queries.len()).filter(|&i| queries[i].len() == dim).collect();
        let mut out: Vec<Result<Vec<SearchResult>, RabitqError>> = queries.iter()
            .map(|q| Err(RabitqError::DimensionMismatch(format!("expected {dim}, got {}", q.len())))).collect();
        if good.is_empty() { return out; }
        let (nq, k) = (good.len(), params.top_k);
        let flat: Vec<f32> = good.iter().flat_map(|&i| queries[i].iter().copied()).collect();
        let (mut ids, mut scores, mut counts) = (vec![0u64; nq * k.max(1)], vec![0f32; nq * k.max(1)], vec![0u32; nq]);
        let rc = unsafe {
            if self.sharded {
                rbq_search_batch_sharded(self.h, flat.as_ptr(), nq, dim, k, params.nprobe, ids.as_mut_ptr(), scores.as_mut_ptr(), counts.as_mut_ptr())
            } else {
                rbq_search_batch(self.h, flat.as_ptr(), nq, dim, k, params.nprobe, ids.as_mut_ptr(), scores.as_mut_ptr(), counts.as_mut_ptr())
            }
        };
        match check(rc) {
            Err(e) => for &i in &good { out[i] = Err(e.clone()); },
            Ok(()) => for (j, &i) in good.iter().enumerate() {
                out[i] = Ok((0..counts[j] as usize).map(|t| SearchResult { id: ids[j * k + t] as usize, score: scores[j * k + t] }).collect());
            },
        }
        out
    }
    /// `search_filtered` (reference src/ivf.rs:1723-1730): `allowed` = the ids of the RoaringBitmap (`filter.iter()`).
    pub fn search_filtered(&self, query: &[f32], params: SearchParams, allowed: impl Iterator<Item = u32>) -> Result<Vec<SearchResult>, RabitqError> {
        let dim = unsafe { rbq_index_dim(self.h) };
        if query.len() != dim { return Err(RabitqError::DimensionMismatch(format!("expected {dim}, got {}", query.len()))); }
        let mut words: Vec<u64> = Vec::new();
        let mut nbits = 0usize;
        for id in allowed {
            let w = id as usize / 64;
            if w >= words.len() { words.resize(w + 1, 0); }
            words[w] |= 1u64 << (id % 64);
            nbits = nbits.max(id as usize + 1);
        }
        let k = params.top_k;
        let (mut ids, mut scores, mut counts) = (vec![0u64; k.max(1)], vec![0f32; k.max(1)], vec![0u32; 1]);
        // an empty bitmap is still a filter (admits nothing): pass a non-null pointer with zero bits
        let dummy = [0u64; 1];
        let wp = if words.is_empty() { dummy.as_ptr() } else { words.as_ptr() };
        check(unsafe { rbq_search_batch_filtered(self.h, query.as_ptr(), 1, dim, k, params.nprobe, wp, nbits,
                                                 ids.as_mut_ptr(), scores.as_mut_ptr(), counts.as_mut_ptr()) })?;
        Ok((0..counts[0] as usize).map(|t| SearchResult { id: ids[t] as usize, score: scores[t] }).collect())
    }
    pub fn search(&self, query: &[f32], params: SearchParams) -> Result<Vec<SearchResult>, RabitqError> {
        self.batch_search(&[query], params).pop().unwrap()
    }
}
impl Drop for IvfRabitqIndex { fn drop(&mut self) { unsafe { rbq_index_free(self.h) } } }
