//! Drop-in for the reference crate's modem entry points (src/transmitter.rs:10-15, src/receiver.rs:8-13): same
//! signatures, engine behind them. The reference's `pub use transmitter::*; pub use receiver::*;` (src/lib.rs:14-21)
//! is replaced by `pub use ofdm_gpu::{encode, decode, ModulationScheme};` -- examples and tests compile unchanged.
//! NOT compiled in this repository's build image (no rustc); see INTEGRATION.md.
use anyhow::{anyhow, Result};
use num::complex::Complex64;
use ofdm_sys as sys;
use std::{ffi::CStr, ptr};

#[derive(Clone, Copy, Debug, PartialEq)]
pub enum ModulationScheme {
    Bpsk,
    Qpsk,
    Qam,
}

pub struct Engine {
    h: *mut sys::ofdm_engine,
    cfg: sys::ofdm_cfg,
}

impl Engine {
    pub fn new(guard_bands: bool, modulation: ModulationScheme, fec: bool, device: i32) -> Result<Self> {
        unsafe {
            let mut cfg: sys::ofdm_cfg = std::mem::zeroed();
            sys::ofdm_cfg_default(&mut cfg);
            cfg.guard_bands = guard_bands as u32;
            cfg.modulation = modulation as u32;
            cfg.fec = fec as u32;
            let mut h = ptr::null_mut();
            if sys::ofdm_engine_create(&cfg, device, &mut h) != 0 {
                return Err(anyhow!("{}", CStr::from_ptr(sys::ofdm_last_error(ptr::null())).to_string_lossy()));
            }
            Ok(Engine { h, cfg })
        }
    }

    /// Batched `encode`: one frame per payload.
    pub fn encode_batch(&mut self, payloads: &[&[u8]]) -> Result<Vec<Vec<Complex64>>> {
        let n = payloads.len() as u32;
        let lens: Vec<u32> = payloads.iter().map(|p| p.len() as u32).collect();
        let pstride = lens.iter().copied().max().unwrap_or(1).max(1);
        let istride = lens.iter().map(|&l| unsafe { sys::ofdm_frame_len(&self.cfg, l) }).max().unwrap_or(880);
        let mut pay = vec![0u8; (n * pstride) as usize];
        for (i, p) in payloads.iter().enumerate() {
            pay[i * pstride as usize..i * pstride as usize + p.len()].copy_from_slice(p);
        }
        let mut iq = vec![sys::ofdm_fc32::default(); (n * istride) as usize];
        let mut flen = vec![0u32; n as usize];
        let rc = unsafe {
            sys::ofdm_tx_encode_batch(self.h, pay.as_ptr(), lens.as_ptr(), pstride, n, iq.as_mut_ptr(), istride, flen.as_mut_ptr(),
                                      sys::OFDM_MEM_HOST, ptr::null_mut())
        };
        self.check(rc)?;
        Ok((0..n as usize)
            .map(|i| iq[i * istride as usize..i * istride as usize + flen[i] as usize].iter().map(|s| Complex64::new(s.re as f64, s.im as f64)).collect())
            .collect())
    }

    /// Batched `decode`: `Err` per stream where the reference returns Err / panics.
    pub fn decode_batch(&mut self, captures: &[&[Complex64]]) -> Result<Vec<Result<Vec<u8>>>> {
        let n = captures.len() as u32;
        let ns: Vec<u32> = captures.iter().map(|c| c.len() as u32).collect();
        let istride = ns.iter().copied().max().unwrap_or(1).max(1);
        let mut iq = vec![sys::ofdm_fc32::default(); (n * istride) as usize];
        for (i, c) in captures.iter().enumerate() {
            for (k, s) in c.iter().enumerate() {
                iq[i * istride as usize + k] = sys::ofdm_fc32 { re: s.re as f32, im: s.im as f32 }; // sig_to_bytes, src/utils.rs:228-236
            }
        }
        let ostride = (istride / 80 + 1) * 48 + 16;
        let mut out = vec![0u8; (n * ostride) as usize];
        let mut olen = vec![0u32; n as usize];
        let mut status = vec![0i32; n as usize];
        let rc = unsafe {
            sys::ofdm_rx_decode_batch(self.h, iq.as_ptr(), ns.as_ptr(), n, istride, istride, out.as_mut_ptr(), ostride, olen.as_mut_ptr(),
                                      status.as_mut_ptr(), ptr::null(), sys::OFDM_MEM_HOST, ptr::null_mut())
        };
        self.check(rc)?;
        Ok((0..n as usize)
            .map(|i| match status[i] {
                sys::OFDM_OK => Ok(out[i * ostride as usize..i * ostride as usize + olen[i] as usize].to_vec()),
                sys::OFDM_TOO_SHORT => Err(anyhow!("Input not long enough, bailing early")), // src/receiver.rs:28
                s => Err(anyhow!("{}", unsafe { CStr::from_ptr(sys::ofdm_status_name(s)) }.to_string_lossy())),
            })
            .collect())
    }

    /// `create_transmission_bytes` (src/utils.rs:97-137): RS(255,223) blocks, the tail block always emitted.
    pub fn rs_encode(&mut self, data: &[u8]) -> Result<Vec<u8>> {
        let n = [data.len() as u32];
        let cap = unsafe { sys::ofdm_rs_encoded_len(data.len()) };
        let mut coded = vec![0u8; cap];
        let mut clen = [0u32];
        let rc = unsafe {
            sys::ofdm_rs_encode_batch(self.h, data.as_ptr(), n.as_ptr(), 1, data.len().max(1) as u32, coded.as_mut_ptr(), cap as u32,
                                      clen.as_mut_ptr(), sys::OFDM_MEM_HOST, ptr::null_mut())
        };
        self.check(rc)?;
        coded.truncate(clen[0] as usize);
        Ok(coded)
    }

    /// `decipher_transmission_bytes` (src/utils.rs:152-180): `None` when a block is beyond repair.
    pub fn rs_decode(&mut self, coded: &[u8]) -> Result<Option<Vec<u8>>> {
        let n = [coded.len() as u32];
        let cap = unsafe { sys::ofdm_rs_decoded_len(coded.len()) };
        let mut data = vec![0u8; cap];
        let (mut dlen, mut fixed, mut failed) = ([0u32], [0u32], [0u32]);
        let rc = unsafe {
            sys::ofdm_rs_decode_batch(self.h, coded.as_ptr(), n.as_ptr(), 1, coded.len().max(1) as u32, data.as_mut_ptr(), cap as u32,
                                      dlen.as_mut_ptr(), fixed.as_mut_ptr(), failed.as_mut_ptr(), sys::OFDM_MEM_HOST, ptr::null_mut())
        };
        self.check(rc)?;
        data.truncate(dlen[0] as usize);
        Ok(if failed[0] == 0 { Some(data) } else { None })
    }

    fn check(&self, rc: i32) -> Result<()> {
        if rc == 0 {
            Ok(())
        } else {
            Err(anyhow!("{}", unsafe { CStr::from_ptr(sys::ofdm_last_error(self.h)) }.to_string_lossy()))
        }
    }
}

impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { sys::ofdm_engine_destroy(self.h) }
    }
}

thread_local! {
    /// One engine per (guard_bands, modulation) and thread: the reference's call-per-frame style (`encode!` / `decode!` in every
    /// example) must not pay for cudaMalloc + stream + event set-up on each call. Handles are not thread-safe, hence thread-local.
    static ENGINES: std::cell::RefCell<std::collections::HashMap<(bool, u32), Engine>> = std::cell::RefCell::new(std::collections::HashMap::new());
}

fn with_engine<T>(guard_bands: Option<bool>, modulation: Option<ModulationScheme>, f: impl FnOnce(&mut Engine) -> Result<T>) -> Result<T> {
    let (g, m) = (guard_bands.unwrap_or(false), modulation.unwrap_or(ModulationScheme::Bpsk));
    ENGINES.with(|cell| {
        let mut map = cell.borrow_mut();
        if !map.contains_key(&(g, m as u32)) {
            map.insert((g, m as u32), Engine::new(g, m, false, 0)?);
        }
        f(map.get_mut(&(g, m as u32)).unwrap())
    })
}

/// src/transmitter.rs:10-15 -- same signature (the `#[optargs::optfn]` attribute can be kept on this function).
pub fn encode(data: &[u8], guard_bands: Option<bool>, modulation: Option<ModulationScheme>) -> Vec<Complex64> {
    with_engine(guard_bands, modulation, |e| Ok(e.encode_batch(&[data])?.pop().unwrap())).expect("encode")
}

/// src/receiver.rs:8-13 -- same signature and error behaviour.
pub fn decode(samples: Vec<Complex64>, guard_bands: Option<bool>, modulation: Option<ModulationScheme>) -> Result<Vec<u8>> {
    with_engine(guard_bands, modulation, |e| e.decode_batch(&[&samples])?.pop().unwrap())
}
