//! Raw bindings to `libfacgpu.so` -- one item per declaration of `include/fac.h` (FAC_ABI_VERSION 2).
//!
//! `tests/test_rust_bindings.py` of the repository diffs this `extern "C"` block against the header
//! (names, parameter counts, struct field lists), so the two cannot drift apart unnoticed.  No Rust
//! toolchain exists in the image the library is developed in; the crate is source that a maintainer
//! compiles on a machine with cargo + nvcc.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const FAC_ABI_VERSION: c_int = 2;

// fac_status
pub const FAC_OK: c_int = 0;
pub const FAC_HAYSTACK_TOO_LARGE: c_int = 1;
pub const FAC_INVALID_UTF8: c_int = 2;
pub const FAC_CUDA_ERROR: c_int = 3;
pub const FAC_OOM: c_int = 4;
pub const FAC_INVALID_ARGUMENT: c_int = 5;
pub const FAC_UNSUPPORTED: c_int = 6;
pub const FAC_IO_ERROR: c_int = 7;

// fac_order / fac_overlap (src/options.rs:10-36 of the reference)
pub const FAC_ORDER_UNSORTED: c_int = 0;
pub const FAC_ORDER_DEFAULT: c_int = 1;
pub const FAC_ORDER_GREEDY: c_int = 2;
pub const FAC_ORDER_COVERAGE_WEIGHTED: c_int = 3;
pub const FAC_OVERLAP_KEEP: c_int = 0;
pub const FAC_OVERLAP_NON_OVERLAPPING: c_int = 1;
pub const FAC_OVERLAP_NON_OVERLAPPING_UNIQUE: c_int = 2;

// flags of fac_search_args.flags / fac_matches_apply_device
pub const FAC_HAYSTACK_ON_DEVICE: u32 = 1;
pub const FAC_RESULT_ON_DEVICE: u32 = 2;
pub const FAC_TEXT_IS_UNICODE: u32 = 4;
pub const FAC_APPLY_PRESORTED: u32 = 8;

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct fac_limits {
    pub insertions: i16,
    pub deletions: i16,
    pub substitutions: i16,
    pub swaps: i16,
    pub edits: i16,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct fac_pattern {
    pub text: *const c_char,
    pub len: usize,
    pub weight: f32,
    pub has_limits: i32,
    pub limits: fac_limits,
    pub unique_id: i64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct fac_sim_pair {
    pub a: u32,
    pub b: u32,
    pub similarity: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct fac_mapping {
    pub a: *const c_char,
    pub a_len: usize,
    pub b: *const c_char,
    pub b_len: usize,
    pub score: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct fac_config {
    pub case_insensitive: i32,
    pub has_limits: i32,
    pub limits: fac_limits,
    pub has_penalties: i32,
    pub penalty_insertion: f32,
    pub penalty_deletion: f32,
    pub penalty_substitution: f32,
    pub penalty_swap: f32,
    pub beam_width: u64,
    pub has_auto_beam: i32,
    pub auto_beam_budget: u64,
    pub auto_beam_width: u64,
    pub min_symbol_similarity: f32,
    pub has_similarity: i32,
    pub similarity: *const fac_sim_pair,
    pub n_similarity: usize,
    pub mappings: *const fac_mapping,
    pub n_mappings: usize,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq)]
pub struct fac_match {
    pub start: u64,
    pub end: u64,
    pub pattern_index: u32,
    pub similarity: f32,
    pub insertions: u8,
    pub deletions: u8,
    pub substitutions: u8,
    pub swaps: u8,
    pub edits: u8,
    pub pad_: [u8; 3],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct fac_shard {
    pub own_begin: usize,
    pub own_end: usize,
    pub read_end: usize,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct fac_search_args {
    pub haystack: *const u8,
    pub len: usize,
    pub own_begin: usize,
    pub own_end: usize,
    pub base: u64,
    pub threshold: f32,
    pub order: c_int,
    pub overlap: c_int,
    pub use_prefilter: i32,
    pub flags: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct fac_window {
    pub text: *const u8,
    pub len: usize,
    pub base: u64,
    pub commit: usize,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct fac_stream_stats {
    pub bytes_read: u64,
    pub bytes_written: u64,
    pub windows: u64,
    pub matches: u64,
    pub states: u64,
    pub device_ms: f64,
    pub expand_ms: f64,
    pub kernel_launches: u32,
    pub devices: u32,
}

#[repr(C)]
pub struct fac_engine {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct fac_matches {
    _opaque: [u8; 0],
}

pub type fac_read_fn = unsafe extern "C" fn(user: *mut c_void, buf: *mut u8, cap: usize) -> i64;
pub type fac_write_fn = unsafe extern "C" fn(user: *mut c_void, buf: *const u8, len: usize) -> c_int;
pub type fac_match_fn = unsafe extern "C" fn(user: *mut c_void, m: *const fac_match);
pub type fac_replace_fn = unsafe extern "C" fn(
    user: *mut c_void,
    m: *const fac_match,
    window_base: u64,
    text: *const u8,
    text_len: usize,
    repl: *mut *const u8,
    repl_len: *mut usize,
) -> c_int;

extern "C" {
    pub fn fac_last_error_string() -> *const c_char;
    pub fn fac_abi_version() -> c_int;
    pub fn fac_build_source_hash() -> *const c_char;
    pub fn fac_engine_create(cfg: *const fac_config, patterns: *const fac_pattern, n_patterns: usize, out: *mut *mut fac_engine) -> c_int;
    pub fn fac_engine_create_on(device: c_int, cfg: *const fac_config, patterns: *const fac_pattern, n_patterns: usize, out: *mut *mut fac_engine) -> c_int;
    pub fn fac_engine_create_multi(devices: *const c_int, n_devices: usize, cfg: *const fac_config, patterns: *const fac_pattern, n_patterns: usize, out: *mut *mut fac_engine) -> c_int;
    pub fn fac_engine_free(engine: *mut fac_engine);
    pub fn fac_engine_max_match_graphemes(engine: *const fac_engine) -> usize;
    pub fn fac_engine_prefilter_active(engine: *const fac_engine) -> c_int;
    pub fn fac_engine_num_nodes(engine: *const fac_engine) -> usize;
    pub fn fac_engine_num_patterns(engine: *const fac_engine) -> usize;
    pub fn fac_engine_device(engine: *const fac_engine) -> c_int;
    pub fn fac_engine_num_devices(engine: *const fac_engine) -> usize;
    pub fn fac_search(engine: *const fac_engine, haystack: *const u8, len: usize, threshold: f32, order: c_int, overlap: c_int, use_prefilter: c_int, out: *mut *mut fac_matches) -> c_int;
    pub fn fac_search_device(engine: *const fac_engine, d_haystack: *const u8, len: usize, threshold: f32, order: c_int, overlap: c_int, use_prefilter: c_int, out: *mut *mut fac_matches) -> c_int;
    pub fn fac_last_haystack_graphemes() -> u64;
    pub fn fac_search_shard(engine: *const fac_engine, haystack: *const u8, len: usize, own_begin: usize, own_end: usize, base: u64, threshold: f32, on_device: c_int, out: *mut *mut fac_matches) -> c_int;
    pub fn fac_matches_apply(engine: *const fac_engine, input: *const fac_match, n: usize, order: c_int, overlap: c_int, out: *mut *mut fac_matches) -> c_int;
    pub fn fac_plan_shards(max_match_graphemes: usize, haystack: *const u8, len: usize, n_shards: usize, out: *mut fac_shard) -> c_int;
    pub fn fac_search_ex(engine: *const fac_engine, args: *const fac_search_args, out: *mut *mut fac_matches) -> c_int;
    pub fn fac_matches_apply_device(engine: *const fac_engine, d_in: *const fac_match, n: usize, order: c_int, overlap: c_int, flags: u32, out: *mut *mut fac_matches) -> c_int;
    pub fn fac_matches_device_data(m: *const fac_matches) -> *const fac_match;
    pub fn fac_search_windows(engine: *const fac_engine, windows: *const fac_window, n_windows: usize, threshold: f32, out: *mut *mut fac_matches) -> c_int;
    pub fn fac_matches_data(m: *const fac_matches) -> *const fac_match;
    pub fn fac_matches_len(m: *const fac_matches) -> usize;
    pub fn fac_matches_states_pushed(m: *const fac_matches) -> u64;
    pub fn fac_matches_device_ms(m: *const fac_matches) -> f64;
    pub fn fac_matches_expand_ms(m: *const fac_matches) -> f64;
    pub fn fac_matches_kernel_launches(m: *const fac_matches) -> u32;
    pub fn fac_matches_free(m: *mut fac_matches);
    pub fn fac_search_stream(engine: *const fac_engine, read: fac_read_fn, read_user: *mut c_void, threshold: f32, on_match: Option<fac_match_fn>, match_user: *mut c_void, bytes_read: *mut u64) -> c_int;
    pub fn fac_replace_stream(engine: *const fac_engine, read: fac_read_fn, read_user: *mut c_void, write: fac_write_fn, write_user: *mut c_void, threshold: f32, replace: fac_replace_fn, replace_user: *mut c_void, bytes_written: *mut u64) -> c_int;
    pub fn fac_replace_stream_table(engine: *const fac_engine, read: fac_read_fn, read_user: *mut c_void, write: fac_write_fn, write_user: *mut c_void, threshold: f32, replacements: *const *const u8, replacement_lens: *const usize, n_replacements: usize, stats: *mut fac_stream_stats) -> c_int;
    pub fn fac_search_stream_stats(engine: *const fac_engine, read: fac_read_fn, read_user: *mut c_void, threshold: f32, on_match: Option<fac_match_fn>, match_user: *mut c_void, stats: *mut fac_stream_stats) -> c_int;
}
