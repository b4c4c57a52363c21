//! Safe handles over the C ABI of `include/fac.h`.
//!
//! This is the layer the reference crate (`fuzzy-aho-corasick` 0.5.0) would hold behind a `gpu` cargo
//! feature: `FuzzyAhoCorasickBuilder::build` (src/builder.rs:181-484) ends by constructing a [`GpuEngine`]
//! from the values it already has, `FuzzyAhoCorasick::search` (src/query.rs:30-38) turns the
//! [`RawMatch`] records of [`GpuEngine::search`] back into `FuzzyMatch { text, pattern, .. }` borrows, and
//! the stream functions (src/stream.rs:319-638) forward their `Read` / `Write` / closure arguments to
//! [`GpuEngine::search_stream`] / [`GpuEngine::replace_stream`].  INTEGRATION.md shows those call sites.
//!
//! There is no CPU fallback: every failure other than `HaystackTooLarge` surfaces as [`GpuError`].
use fuzzy_aho_corasick_gpu_sys as sys;
use std::ffi::CStr;
use std::io::{Read, Write};
use std::os::raw::{c_int, c_void};
use std::ptr;

pub use sys::fac_match as RawMatch;
pub use sys::fac_stream_stats as StreamStats;

#[derive(Debug, Clone, PartialEq, Eq)]
pub enum GpuError {
    /// `SearchError::HaystackTooLarge { graphemes }` (src/error.rs:9-17)
    HaystackTooLarge { graphemes: u64 },
    InvalidUtf8,
    Cuda(String),
    OutOfMemory(String),
    InvalidArgument(String),
    Unsupported(String),
    Io(String),
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::fac_last_error_string()) }.to_string_lossy().into_owned()
}

fn check(status: c_int) -> Result<(), GpuError> {
    match status {
        sys::FAC_OK => Ok(()),
        sys::FAC_HAYSTACK_TOO_LARGE => Err(GpuError::HaystackTooLarge { graphemes: unsafe { sys::fac_last_haystack_graphemes() } }),
        sys::FAC_INVALID_UTF8 => Err(GpuError::InvalidUtf8),
        sys::FAC_OOM => Err(GpuError::OutOfMemory(last_error())),
        sys::FAC_INVALID_ARGUMENT => Err(GpuError::InvalidArgument(last_error())),
        sys::FAC_UNSUPPORTED => Err(GpuError::Unsupported(last_error())),
        sys::FAC_IO_ERROR => Err(GpuError::Io(last_error())),
        _ => Err(GpuError::Cuda(last_error())),
    }
}

/// `FuzzyLimits` (src/structs.rs:292-363); `None` is encoded as -1 on the wire.
#[derive(Debug, Clone, Copy, Default)]
pub struct Limits {
    pub insertions: Option<u8>,
    pub deletions: Option<u8>,
    pub substitutions: Option<u8>,
    pub swaps: Option<u8>,
    pub edits: Option<u8>,
}
impl Limits {
    fn raw(&self) -> sys::fac_limits {
        let f = |v: Option<u8>| v.map(|x| x as i16).unwrap_or(-1);
        sys::fac_limits { insertions: f(self.insertions), deletions: f(self.deletions), substitutions: f(self.substitutions), swaps: f(self.swaps), edits: f(self.edits) }
    }
}

/// `Pattern` (src/structs.rs:597-610)
#[derive(Debug, Clone)]
pub struct PatternSpec<'a> {
    pub text: &'a str,
    pub weight: f32,
    pub limits: Option<Limits>,
    pub custom_unique_id: Option<u32>,
}

/// The builder state `build()` has collected (src/builder.rs:22-143).
#[derive(Debug, Clone, Default)]
pub struct EngineConfig<'a> {
    pub case_insensitive: bool,
    pub limits: Option<Limits>,
    /// (insertion, deletion, substitution, swap); `None` = `FuzzyPenalties::default()`
    pub penalties: Option<(f32, f32, f32, f32)>,
    pub beam_width: Option<u64>,
    pub auto_beam: Option<(u64, u64)>,
    pub min_symbol_similarity: f32,
    /// `Similarity::from_map` entries; `None` = DEFAULT_SIMILARITY (src/builder.rs:492-526)
    pub similarity: Option<&'a [(char, char, f32)]>,
    /// `mapping_scored(a, b, score)` (src/builder.rs:116-132)
    pub mappings: &'a [(&'a str, &'a str, f32)],
    /// CUDA devices to place the automaton on (empty = the current device).  More than one device lets the stream
    /// functions deal their windows round-robin.
    pub devices: &'a [i32],
}

pub struct GpuEngine {
    raw: *mut sys::fac_engine,
}
// The handle is immutable after creation and re-entrant (src/structs.rs:522-529, src/stream.rs:395-402).
unsafe impl Send for GpuEngine {}
unsafe impl Sync for GpuEngine {}

impl Drop for GpuEngine {
    fn drop(&mut self) {
        unsafe { sys::fac_engine_free(self.raw) }
    }
}

/// A match list owned by the library (pinned host memory) until dropped.
pub struct GpuMatches {
    raw: *mut sys::fac_matches,
}
unsafe impl Send for GpuMatches {}
impl Drop for GpuMatches {
    fn drop(&mut self) {
        unsafe { sys::fac_matches_free(self.raw) }
    }
}
impl GpuMatches {
    pub fn as_slice(&self) -> &[RawMatch] {
        let n = unsafe { sys::fac_matches_len(self.raw) };
        let p = unsafe { sys::fac_matches_data(self.raw) };
        if n == 0 || p.is_null() { &[] } else { unsafe { std::slice::from_raw_parts(p, n) } }
    }
    pub fn states_pushed(&self) -> u64 {
        unsafe { sys::fac_matches_states_pushed(self.raw) }
    }
    pub fn device_ms(&self) -> f64 {
        unsafe { sys::fac_matches_device_ms(self.raw) }
    }
}

#[derive(Debug, Clone, Copy, PartialEq, Eq)]
#[repr(i32)]
pub enum Order { Unsorted = 0, Default = 1, Greedy = 2, CoverageWeighted = 3 }
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
#[repr(i32)]
pub enum Overlap { Keep = 0, NonOverlapping = 1, NonOverlappingUnique = 2 }

impl GpuEngine {
    pub fn new(cfg: &EngineConfig<'_>, patterns: &[PatternSpec<'_>]) -> Result<Self, GpuError> {
        let sim: Vec<sys::fac_sim_pair> = cfg.similarity.unwrap_or(&[]).iter()
            .map(|&(a, b, s)| sys::fac_sim_pair { a: a as u32, b: b as u32, similarity: s }).collect();
        let maps: Vec<sys::fac_mapping> = cfg.mappings.iter()
            .map(|&(a, b, s)| sys::fac_mapping { a: a.as_ptr().cast(), a_len: a.len(), b: b.as_ptr().cast(), b_len: b.len(), score: s }).collect();
        let none = Limits::default().raw();
        let (pi, pd, ps, pw) = cfg.penalties.unwrap_or((0.0, 0.0, 0.0, 0.0));
        let raw_cfg = sys::fac_config {
            case_insensitive: cfg.case_insensitive as i32,
            has_limits: cfg.limits.is_some() as i32,
            limits: cfg.limits.map(|l| l.raw()).unwrap_or(none),
            has_penalties: cfg.penalties.is_some() as i32,
            penalty_insertion: pi, penalty_deletion: pd, penalty_substitution: ps, penalty_swap: pw,
            beam_width: cfg.beam_width.unwrap_or(0),
            has_auto_beam: cfg.auto_beam.is_some() as i32,
            auto_beam_budget: cfg.auto_beam.map(|x| x.0).unwrap_or(0),
            auto_beam_width: cfg.auto_beam.map(|x| x.1).unwrap_or(0),
            min_symbol_similarity: cfg.min_symbol_similarity,
            has_similarity: cfg.similarity.is_some() as i32,
            similarity: if sim.is_empty() { ptr::null() } else { sim.as_ptr() },
            n_similarity: sim.len(),
            mappings: if maps.is_empty() { ptr::null() } else { maps.as_ptr() },
            n_mappings: maps.len(),
        };
        let pats: Vec<sys::fac_pattern> = patterns.iter().map(|p| sys::fac_pattern {
            text: p.text.as_ptr().cast(), len: p.text.len(), weight: p.weight,
            has_limits: p.limits.is_some() as i32, limits: p.limits.map(|l| l.raw()).unwrap_or(none),
            unique_id: p.custom_unique_id.map(|x| x as i64).unwrap_or(-1),
        }).collect();
        let mut raw = ptr::null_mut();
        let st = unsafe {
            if cfg.devices.is_empty() { sys::fac_engine_create(&raw_cfg, pats.as_ptr(), pats.len(), &mut raw) }
            else { sys::fac_engine_create_multi(cfg.devices.as_ptr(), cfg.devices.len(), &raw_cfg, pats.as_ptr(), pats.len(), &mut raw) }
        };
        check(st)?;
        Ok(GpuEngine { raw })
    }

    /// `max_match_graphemes` (src/stream.rs:213-253)
    pub fn max_match_graphemes(&self) -> usize {
        unsafe { sys::fac_engine_max_match_graphemes(self.raw) }
    }
    /// `Prefiltered::is_active` (src/prefilter.rs:125-127)
    pub fn prefilter_active(&self) -> bool {
        unsafe { sys::fac_engine_prefilter_active(self.raw) != 0 }
    }

    /// `engine.search(haystack, &opts)` (src/query.rs:30-38); `use_prefilter` = `Prefiltered::search` (src/prefilter.rs:135).
    pub fn search(&self, haystack: &str, threshold: f32, order: Order, overlap: Overlap, use_prefilter: bool) -> Result<GpuMatches, GpuError> {
        let mut out = ptr::null_mut();
        check(unsafe { sys::fac_search(self.raw, haystack.as_ptr(), haystack.len(), threshold, order as c_int, overlap as c_int, use_prefilter as c_int, &mut out) })?;
        Ok(GpuMatches { raw: out })
    }

    /// `search_stream` / `search_stream_parallel` (src/stream.rs:319-429): `on_match` sees absolute offsets in stream order.
    pub fn search_stream<R: Read, F: FnMut(&RawMatch)>(&self, reader: R, threshold: f32, on_match: F) -> Result<StreamStats, GpuError> {
        let mut rd = ReadCtx { r: reader, err: None };
        let mut cb = on_match;
        let mut stats = StreamStats::default();
        let st = unsafe {
            sys::fac_search_stream_stats(self.raw, read_tramp::<R>, (&mut rd as *mut ReadCtx<R>).cast(), threshold,
                                         Some(match_tramp::<F>), (&mut cb as *mut F).cast(), &mut stats)
        };
        if let Some(e) = rd.err.take() { return Err(GpuError::Io(e.to_string())); }
        check(st)?;
        Ok(stats)
    }

    /// `FuzzyReplacer::replace_stream` (src/replacer.rs:35-46): replacement text by pattern index.
    pub fn replace_stream<R: Read, W: Write>(&self, reader: R, writer: W, threshold: f32, replacements: &[&str]) -> Result<StreamStats, GpuError> {
        let mut rd = ReadCtx { r: reader, err: None };
        let mut wr = WriteCtx { w: writer, err: None };
        let ptrs: Vec<*const u8> = replacements.iter().map(|s| s.as_ptr()).collect();
        let lens: Vec<usize> = replacements.iter().map(|s| s.len()).collect();
        let mut stats = StreamStats::default();
        let st = unsafe {
            sys::fac_replace_stream_table(self.raw, read_tramp::<R>, (&mut rd as *mut ReadCtx<R>).cast(), write_tramp::<W>,
                                          (&mut wr as *mut WriteCtx<W>).cast(), threshold, ptrs.as_ptr(), lens.as_ptr(), ptrs.len(), &mut stats)
        };
        if let Some(e) = rd.err.take() { return Err(GpuError::Io(e.to_string())); }
        if let Some(e) = wr.err.take() { return Err(GpuError::Io(e.to_string())); }
        check(st)?;
        Ok(stats)
    }
}

struct ReadCtx<R> { r: R, err: Option<std::io::Error> }
struct WriteCtx<W> { w: W, err: Option<std::io::Error> }

unsafe extern "C" fn read_tramp<R: Read>(user: *mut c_void, buf: *mut u8, cap: usize) -> i64 {
    let ctx = &mut *(user as *mut ReadCtx<R>);
    let dst = std::slice::from_raw_parts_mut(buf, cap);
    loop {
        match ctx.r.read(dst) {
            Ok(n) => return n as i64,
            Err(e) if e.kind() == std::io::ErrorKind::Interrupted => continue,
            Err(e) => { ctx.err = Some(e); return -1; }
        }
    }
}
unsafe extern "C" fn write_tramp<W: Write>(user: *mut c_void, buf: *const u8, len: usize) -> c_int {
    let ctx = &mut *(user as *mut WriteCtx<W>);
    match ctx.w.write_all(std::slice::from_raw_parts(buf, len)) {
        Ok(()) => 0,
        Err(e) => { ctx.err = Some(e); 1 }
    }
}
unsafe extern "C" fn match_tramp<F: FnMut(&RawMatch)>(user: *mut c_void, m: *const RawMatch) {
    (*(user as *mut F))(&*m)
}
