//! Raw `extern "C"` bindings of `include/ofdm_engine.h` (ABI version 1) -- exactly the symbols the reference crate's
//! modem path needs. Field order and types mirror the C structs one to one.
//! NOT compiled in this repository's build image (no rustc); kept in sync with the header by
//! tests/test_abi_host.py::test_rust_shim_declares_every_symbol.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct ofdm_engine {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct ofdm_fc32 {
    pub re: f32,
    pub im: f32,
}

pub const OFDM_MOD_BPSK: u32 = 0;
pub const OFDM_MOD_QPSK: u32 = 1;
pub const OFDM_MOD_QAM64: u32 = 2;
pub const OFDM_SYNC_REFERENCE: u32 = 0;
pub const OFDM_SYNC_SCHMIDL_COX: u32 = 1;
pub const OFDM_CFO_REFERENCE: u32 = 0;
pub const OFDM_CFO_ANGLE_OF_SUM: u32 = 1;
pub const OFDM_PHASE_REFERENCE: u32 = 0;
pub const OFDM_PHASE_ANGLE_OF_SUM: u32 = 1;
pub const OFDM_MEM_HOST: c_int = 0;
pub const OFDM_MEM_DEVICE: c_int = 1;
pub const OFDM_OK: i32 = 0;
pub const OFDM_TOO_SHORT: i32 = 1;
pub const OFDM_NO_SYNC: i32 = 2;
pub const OFDM_BAD_HEADER: i32 = 3;
pub const OFDM_NEG_OFFSET: i32 = 4;

#[repr(C)]
pub struct ofdm_cfg {
    pub struct_size: u32,
    pub nfft: u32,
    pub cp: u32,
    pub modulation: u32,
    pub guard_bands: u32,
    pub fec: u32,
    pub sync_mode: u32,
    pub cfo_mode: u32,
    pub phase_mode: u32,
    pub sync_window: u32,
    pub locking: *const ofdm_fc32,
    pub preamble: *const ofdm_fc32,
    pub training: *const ofdm_fc32,
}

#[repr(C)]
pub struct ofdm_rx_diag {
    pub offset: *mut i32,
    pub f_delta: *mut f32,
    pub h_k: *mut ofdm_fc32,
    pub n_data_syms: *mut u32,
    pub points: *mut ofdm_fc32,
    pub points_stride: u32,
}

#[repr(C)]
pub struct ofdm_channel_params {
    pub snr_db: f32,
    pub cfo_max: f32,
    pub lead_min: u32,
    pub lead_max: u32,
    pub multipath: u32,
    pub noise_mode: u32,
    pub seed: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct ofdm_frame_info {
    pub offset: u64,
    pub f_delta: f32,
    pub metric: f32,
    pub status: i32,
    pub out_len: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct ofdm_peak {
    pub offset: u64,
    pub f_delta: f32,
    pub metric: f32,
}

extern "C" {
    pub fn ofdm_abi_version() -> u32;
    pub fn ofdm_cfg_default(cfg: *mut ofdm_cfg);
    pub fn ofdm_status_name(status: i32) -> *const c_char;
    pub fn ofdm_engine_create(cfg: *const ofdm_cfg, device: c_int, out: *mut *mut ofdm_engine) -> c_int;
    pub fn ofdm_engine_destroy(h: *mut ofdm_engine);
    pub fn ofdm_last_error(h: *const ofdm_engine) -> *const c_char;
    pub fn ofdm_engine_reserve(h: *mut ofdm_engine, max_streams: u32, max_capture_samples: u64) -> c_int;
    pub fn ofdm_get_tables(h: *const ofdm_engine, locking80: *mut ofdm_fc32, preamble80: *mut ofdm_fc32, training64: *mut ofdm_fc32) -> c_int;
    pub fn ofdm_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn ofdm_host_free(p: *mut c_void);
    pub fn ofdm_coded_len(cfg: *const ofdm_cfg, payload_len: u32) -> u32;
    pub fn ofdm_frame_data_syms(cfg: *const ofdm_cfg, payload_len: u32) -> u32;
    pub fn ofdm_frame_len(cfg: *const ofdm_cfg, payload_len: u32) -> u32;
    pub fn ofdm_max_payload(cfg: *const ofdm_cfg, n_data_syms: u32) -> u32;
    pub fn ofdm_tx_encode_batch(h: *mut ofdm_engine, payload: *const u8, payload_len: *const u32, payload_stride: u32, n_streams: u32,
                                iq_out: *mut ofdm_fc32, iq_stride: u32, frame_len_out: *mut u32, mem: c_int, stream: *mut c_void) -> c_int;
    pub fn ofdm_rx_decode_batch(h: *mut ofdm_engine, iq: *const ofdm_fc32, n_samples: *const u32, n_streams: u32, iq_stride: u32,
                                max_n_samples: u32, out: *mut u8, out_stride: u32, out_len: *mut u32, status: *mut i32,
                                diag: *const ofdm_rx_diag, mem: c_int, stream: *mut c_void) -> c_int;
    pub fn ofdm_channel_apply_batch(h: *mut ofdm_engine, tx: *const ofdm_fc32, tx_len: *const u32, tx_stride: u32, n_streams: u32,
                                    p: *const ofdm_channel_params, rx: *mut ofdm_fc32, rx_stride: u32, rx_len: *mut u32,
                                    lead_out: *mut u32, cfo_out: *mut f32, mem: c_int, stream: *mut c_void) -> c_int;
    pub fn ofdm_ber_accumulate(h: *mut ofdm_engine, reference: *const u8, ref_len: *const u32, ref_stride: u32, got: *const u8,
                               got_len: *const u32, got_stride: u32, status: *const i32, n_streams: u32, counters: *mut u64,
                               mem: c_int, stream: *mut c_void) -> c_int;
    pub fn ofdm_sync_search(h: *mut ofdm_engine, iq: *const ofdm_fc32, n_samples: u64, peaks: *mut ofdm_peak, max_peaks: u32,
                            n_peaks: *mut u32, mem: c_int, stream: *mut c_void) -> c_int;
    pub fn ofdm_sync_counts(h: *mut ofdm_engine, counts: *mut u32, stream: *mut c_void) -> c_int;
    pub fn ofdm_rx_decode_capture(h: *mut ofdm_engine, iq: *const ofdm_fc32, n_samples: u64, peaks: *const ofdm_peak, n_frames: u32,
                                  max_frame_samples: u32, out: *mut u8, out_stride: u32, out_len: *mut u32, status: *mut i32,
                                  mem: c_int, stream: *mut c_void) -> c_int;
    pub fn ofdm_rx_decode_file(h: *mut ofdm_engine, path: *const c_char, start: u64, stop: u64, chunk_samples: u32,
                               max_frame_samples: u32, out: *mut u8, out_stride: u32, frames: *mut ofdm_frame_info,
                               max_frames: u32, n_frames: *mut u32) -> c_int;
    pub fn ofdm_stats_allreduce(h: *mut ofdm_engine, counters: *mut u64, nccl_comm: *mut c_void, mem: c_int, stream: *mut c_void) -> c_int;
    pub fn ofdm_rs_encoded_len(data_len: usize) -> usize;
    pub fn ofdm_rs_decoded_len(coded_len: usize) -> usize;
    pub fn ofdm_rs_encode_batch(h: *mut ofdm_engine, data: *const u8, data_len: *const u32, n_streams: u32, data_stride: u32,
                                coded: *mut u8, coded_stride: u32, coded_len: *mut u32, mem: c_int, stream: *mut c_void) -> c_int;
    pub fn ofdm_rs_decode_batch(h: *mut ofdm_engine, coded: *const u8, coded_len: *const u32, n_streams: u32, coded_stride: u32,
                                data: *mut u8, data_stride: u32, data_len: *mut u32, n_corrected: *mut u32, n_failed: *mut u32,
                                mem: c_int, stream: *mut c_void) -> c_int;
    pub fn ofdm_profile_begin(h: *mut ofdm_engine, max_calls: u32) -> c_int;
    pub fn ofdm_profile_read(h: *mut ofdm_engine, acquire_ms: *mut f32, decode_ms: *mut f32, n_calls: *mut u32) -> c_int;
    pub fn ofdm_kernel_launches(h: *const ofdm_engine) -> u64;
    pub fn ofdm_last_h2d_bytes(h: *const ofdm_engine) -> u64;
}
