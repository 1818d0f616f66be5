//! `extern "C"` declarations of include/felics_b200.h -- the whole drop-in boundary, nothing else.
//! Checked against the header by tests/test_rust_facade.py (names, argument count, argument and return types).
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub const FELICS_OK: c_int = 0;
pub const FELICS_ERR_IO: c_int = -1;
pub const FELICS_ERR_INVALID_VALUE: c_int = -2;
pub const FELICS_ERR_VALUE_OVERFLOW: c_int = -3;
pub const FELICS_ERR_INVALID_DIMENSIONS: c_int = -4;
pub const FELICS_ERR_INVALID_COLOR_TYPE: c_int = -5;
pub const FELICS_ERR_INVALID_PIXEL_DEPTH: c_int = -6;
pub const FELICS_ERR_INVALID_SIGNATURE: c_int = -7;
pub const FELICS_ERR_BUFFER_TOO_SMALL: c_int = -8;
pub const FELICS_ERR_CUDA: c_int = -9;
pub const FELICS_ERR_UNSUPPORTED: c_int = -10;
pub const FELICS_ERR_CORRUPT: c_int = -11;
pub const FELICS_ERR_INVALID_ARGUMENT: c_int = -12;
pub const FELICS_HEADER_BYTES: usize = 14;

/// `struct felics_header` (format.rs:44-49)
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct felics_header {
    pub color_type: u8,
    pub pixel_depth: u8,
    pub width: u32,
    pub height: u32,
}

/// Opaque context: one device, one stream, its scratch memory.
#[repr(C)]
pub struct felics_ctx {
    _private: [u8; 0],
}

extern "C" {
    pub fn felics_ctx_create(device: c_int, out: *mut *mut felics_ctx) -> c_int;
    pub fn felics_ctx_destroy(ctx: *mut felics_ctx);
    pub fn felics_ctx_set_stream(ctx: *mut felics_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn felics_last_error() -> *const c_char;
    pub fn felics_read_header(buf: *const u8, len: usize, out: *mut felics_header) -> c_int;
    pub fn felics_write_header(hdr: *const felics_header, out14: *mut u8) -> c_int;
    pub fn felics_pixel_bytes(hdr: *const felics_header) -> usize;
    pub fn felics_compress_bound(hdr: *const felics_header) -> usize;
    pub fn felics_compress(ctx: *mut felics_ctx, pixels: *const c_void, hdr: *const felics_header, out: *mut u8, cap: usize, out_len: *mut usize) -> c_int;
    pub fn felics_decompress(ctx: *mut felics_ctx, fel: *const u8, len: usize, pixels_out: *mut c_void, cap: usize, hdr_out: *mut felics_header) -> c_int;
    pub fn felics_compress_device(ctx: *mut felics_ctx, d_pixels: *const c_void, hdr: *const felics_header, d_out: *mut u8, cap: usize, out_len: *mut usize) -> c_int;
    pub fn felics_decompress_device(ctx: *mut felics_ctx, d_fel: *const u8, len: usize, d_pixels_out: *mut c_void, cap: usize, hdr_out: *mut felics_header) -> c_int;
    pub fn felics_compress_batch(ctx: *mut felics_ctx, n: usize, pixels: *const c_void, hdr: *const felics_header, arena: *mut u8, arena_cap: usize, offsets: *mut u64) -> c_int;
    pub fn felics_compress_batch_device(ctx: *mut felics_ctx, n: usize, d_pixels: *const c_void, hdr: *const felics_header, d_arena: *mut u8, arena_cap: usize, offsets: *mut u64) -> c_int;
    pub fn felics_compress_batch_v(ctx: *mut felics_ctx, n: usize, pixels: *const *const c_void, hdrs: *const felics_header, arena: *mut u8, arena_cap: usize, offsets: *mut u64) -> c_int;
    pub fn felics_decompress_batch(ctx: *mut felics_ctx, n: usize, arena: *const u8, offsets: *const u64, hdr: *const felics_header, pixels_out: *mut c_void, status: *mut c_int) -> c_int;
    pub fn felics_decompress_batch_device(ctx: *mut felics_ctx, n: usize, d_arena: *const u8, offsets: *const u64, hdr: *const felics_header, d_pixels_out: *mut c_void, status: *mut c_int) -> c_int;
    pub fn felics_decompress_batch_v(ctx: *mut felics_ctx, n: usize, arena: *const u8, offsets: *const u64, pixels_out: *const *mut c_void, caps: *const usize, hdrs_out: *mut felics_header, status: *mut c_int) -> c_int;
    pub fn felics_sidecar_build(ctx: *mut felics_ctx, band_rows: u32, sidecar_out: *mut u8, cap: usize, out_len: *mut usize) -> c_int;
    pub fn felics_decompress_sidecar(ctx: *mut felics_ctx, fel: *const u8, len: usize, sidecar: *const u8, sidecar_len: usize, pixels_out: *mut c_void, cap: usize, hdr_out: *mut felics_header) -> c_int;
    pub fn felics_profile_enable(ctx: *mut felics_ctx, on: c_int) -> c_int;
    pub fn felics_profile_reset(ctx: *mut felics_ctx) -> c_int;
    pub fn felics_profile_stage_count() -> c_int;
    pub fn felics_profile_stage_name(stage: c_int) -> *const c_char;
    pub fn felics_profile_stage_ms(ctx: *mut felics_ctx, stage: c_int) -> f64;
    pub fn felics_profile_stage_launches(ctx: *mut felics_ctx, stage: c_int) -> u64;
    pub fn felics_profile_total_launches(ctx: *mut felics_ctx) -> u64;
    pub fn felics_version() -> *const c_char;
}
