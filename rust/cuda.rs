//! `src/cuda.rs` — the FFI shim a pfv-rs maintainer adds behind `--features cuda`.
//!
//! UNCOMPILED: the build image for this engine has no `cargo`/`rustc`; this file is the source a maintainer
//! would drop into the crate (INTEGRATION.md walks through the call sites).  It binds exactly the entry points
//! declared in `include/pfv_b200.h` and keeps the `pfv_rs::dec::Decoder` / `pfv_rs::enc::Encoder` API unchanged.
#![allow(non_camel_case_types)]

use std::ffi::CStr;
use std::io;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct pfv_mbhdr {
    pub mx: i8,
    pub my: i8,
    pub has_coeff: u8,
    pub reserved: u8,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct pfv_geometry {
    pub width: u32, pub height: u32,
    pub cwidth: u32, pub cheight: u32,
    pub pw: u32, pub ph: u32,
    pub cpw: u32, pub cph: u32,
    pub nb_y: u32, pub nb_c: u32, pub nb: u32,
    pub frame_bytes: u32,
}

#[repr(C)]
pub struct pfv_decode_job {
    pub kind: u32,
    pub flags: u32,
    pub dst_slot: u32,
    pub ref_slot: u32,
    pub qidx: [u8; 3],
    pub reserved: u8,
    pub hdr: *const pfv_mbhdr,
    pub coeff: *const i16,
    pub out_y: *mut u8,
    pub out_u: *mut u8,
    pub out_v: *mut u8,
}

#[repr(C)]
pub struct pfv_encode_job {
    pub kind: u32,
    pub flags: u32,
    pub dst_slot: u32,
    pub ref_slot: u32,
    pub px_err: f32,
    pub reserved: u32,
    pub src_y: *const u8,
    pub src_u: *const u8,
    pub src_v: *const u8,
    pub hdr_out: *mut pfv_mbhdr,
    pub coeff_out: *mut i16,
}

/// pfv_decode_job_sparse: what the entropy loops of dec.rs:261-296 / :378-417 produce before the dense scatter
#[repr(C)]
pub struct pfv_decode_job_sparse {
    pub kind: u32,
    pub flags: u32,
    pub dst_slot: u32,
    pub ref_slot: u32,
    pub qidx: [u8; 3],
    pub reserved: u8,
    pub hdr: *const pfv_mbhdr,
    pub mb_off: *const u32,      // nb + 1 offsets into tok
    pub tok: *const u32,         // (position << 16) | (value as u16)
    pub ntok: u32,
    pub reserved2: u32,
    pub out_y: *mut u8,
    pub out_u: *mut u8,
    pub out_v: *mut u8,
}

/// pfv_encode_job_sparse: the frame's RLE sequence (what rle_encode, rle.rs:9-39, pushes) instead of dense coefficients
#[repr(C)]
pub struct pfv_encode_job_sparse {
    pub kind: u32,
    pub flags: u32,
    pub dst_slot: u32,
    pub ref_slot: u32,
    pub px_err: f32,
    pub tok_cap: u32,
    pub src_y: *const u8,
    pub src_u: *const u8,
    pub src_v: *const u8,
    pub hdr_out: *mut pfv_mbhdr,
    pub mb_off_out: *mut u32,    // optional, nb + 1
    pub tok_out: *mut u32,       // pinned: run | size << 4 | (coeff as u16) << 16
    pub stats_out: *mut u32,     // pinned: PFV_TOKSTATS_WORDS
}

pub const PFV_TOKSTATS_WORDS: usize = 36;
pub const PFV_TOKSTATS_NTOK: usize = 32;
pub const PFV_TOKSTATS_FLAGS: usize = 33;

pub enum pfv_ctx {}

pub const PFV_FRAME_I: u32 = 1;
pub const PFV_FRAME_P: u32 = 2;

#[link(name = "pfv_b200")]
extern "C" {
    pub fn pfv_abi_version() -> c_int;
    pub fn pfv_last_error() -> *const c_char;
    pub fn pfv_ctx_create(device: c_int, width: u32, height: u32, qtables: *const [i32; 64], nq: u32,
                          nslots: u32, max_jobs: u32, ext_stream: *mut c_void, out: *mut *mut pfv_ctx) -> c_int;
    pub fn pfv_ctx_destroy(ctx: *mut pfv_ctx);
    pub fn pfv_ctx_geometry(ctx: *const pfv_ctx, out: *mut pfv_geometry) -> c_int;
    pub fn pfv_sync(ctx: *mut pfv_ctx) -> c_int;
    pub fn pfv_slot_reset(ctx: *mut pfv_ctx, slot: u32) -> c_int;
    pub fn pfv_slot_read_visible(ctx: *mut pfv_ctx, slot: u32, y: *mut u8, u: *mut u8, v: *mut u8) -> c_int;
    pub fn pfv_decode_submit(ctx: *mut pfv_ctx, jobs: *const pfv_decode_job, njobs: u32) -> c_int;
    pub fn pfv_decode_submit_sparse(ctx: *mut pfv_ctx, jobs: *const pfv_decode_job_sparse, njobs: u32) -> c_int;
    pub fn pfv_ctx_last_submit_id(ctx: *const pfv_ctx) -> u64;
    pub fn pfv_ctx_wait_submit(ctx: *mut pfv_ctx, submit_id: u64) -> c_int;
    pub fn pfv_encode_submit(ctx: *mut pfv_ctx, jobs: *const pfv_encode_job, njobs: u32) -> c_int;
    pub fn pfv_encode_submit_sparse(ctx: *mut pfv_ctx, jobs: *const pfv_encode_job_sparse, njobs: u32) -> c_int;
    pub fn pfv_host_alloc(out: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn pfv_host_free(p: *mut c_void);
}

fn check(rc: c_int) -> io::Result<()> {
    if rc == 0 {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(pfv_last_error()) }.to_string_lossy().into_owned();
    Err(io::Error::new(io::ErrorKind::Other, format!("pfv_b200 error {}: {}", rc, msg)))
}

/// Device-side replacement of `Decoder.framebuffer` / `Encoder.prev_frame`: two frame slots that ping-pong,
/// which is how the two-phase update of `VideoPlane::decode_plane_delta_into` (common.rs:498-521) is kept.
pub struct CudaPlanes {
    ctx: *mut pfv_ctx,
    cur: u32,
}

impl CudaPlanes {
    pub fn new(width: usize, height: usize, qtables: &[[i32; 64]]) -> io::Result<CudaPlanes> {
        let mut ctx: *mut pfv_ctx = std::ptr::null_mut();
        check(unsafe {
            pfv_ctx_create(0, width as u32, height as u32, qtables.as_ptr(), qtables.len() as u32, 2, 1,
                           std::ptr::null_mut(), &mut ctx)
        })?;
        Ok(CudaPlanes { ctx, cur: 0 })
    }

    /// replaces the three `deserialize_plane` calls of dec.rs:303-310
    pub fn decode_iframe(&mut self, coeff: &[i16], qidx: [u8; 3]) -> io::Result<()> {
        let job = pfv_decode_job {
            kind: PFV_FRAME_I, flags: 0, dst_slot: self.cur ^ 1, ref_slot: 0, qidx, reserved: 0,
            hdr: std::ptr::null(), coeff: coeff.as_ptr(),
            out_y: std::ptr::null_mut(), out_u: std::ptr::null_mut(), out_v: std::ptr::null_mut(),
        };
        check(unsafe { pfv_decode_submit(self.ctx, &job, 1) })?;
        self.cur ^= 1;
        check(unsafe { pfv_sync(self.ctx) })   // `coeff` is a borrowed Vec: it must outlive the copy
    }

    /// replaces the three `deserialize_plane_delta` calls of dec.rs:425-432
    pub fn decode_pframe(&mut self, hdr: &[pfv_mbhdr], coeff: &[i16], qidx: [u8; 3]) -> io::Result<()> {
        let job = pfv_decode_job {
            kind: PFV_FRAME_P, flags: 0, dst_slot: self.cur ^ 1, ref_slot: self.cur, qidx, reserved: 0,
            hdr: hdr.as_ptr(), coeff: coeff.as_ptr(),
            out_y: std::ptr::null_mut(), out_u: std::ptr::null_mut(), out_v: std::ptr::null_mut(),
        };
        check(unsafe { pfv_decode_submit(self.ctx, &job, 1) })?;
        self.cur ^= 1;
        check(unsafe { pfv_sync(self.ctx) })
    }

    /// Sparse variant of `decode_iframe` / `decode_pframe`: the caller's entropy loop pushes
    /// `((out_idx & 255) << 16) | (coeff as u16 as u32)` into `tok` and bumps `mb_off[(out_idx >> 8) + 1..]`
    /// instead of writing `coefficients[out_idx] = coeff` (dec.rs:288, :410); 10-20x fewer bytes cross PCIe.
    pub fn decode_sparse(&mut self, kind: u32, hdr: &[pfv_mbhdr], mb_off: &[u32], tok: &[u32], qidx: [u8; 3]) -> io::Result<()> {
        let job = pfv_decode_job_sparse {
            kind, flags: 0, dst_slot: self.cur ^ 1, ref_slot: self.cur, qidx, reserved: 0,
            hdr: if kind == PFV_FRAME_P { hdr.as_ptr() } else { std::ptr::null() },
            mb_off: mb_off.as_ptr(), tok: tok.as_ptr(), ntok: tok.len() as u32, reserved2: 0,
            out_y: std::ptr::null_mut(), out_u: std::ptr::null_mut(), out_v: std::ptr::null_mut(),
        };
        check(unsafe { pfv_decode_submit_sparse(self.ctx, &job, 1) })?;
        self.cur ^= 1;
        check(unsafe { pfv_sync(self.ctx) })
    }

    /// replaces the retframe blits of dec.rs:195-197 (tight visible planes)
    pub fn read_visible(&mut self, y: &mut [u8], u: &mut [u8], v: &mut [u8]) -> io::Result<()> {
        check(unsafe { pfv_slot_read_visible(self.ctx, self.cur, y.as_mut_ptr(), u.as_mut_ptr(), v.as_mut_ptr()) })?;
        check(unsafe { pfv_sync(self.ctx) })
    }

    /// replaces enc.rs:84-97 (kind = I) / enc.rs:134-147 (kind = P): encode + closed-loop recon into prev_frame
    pub fn encode(&mut self, kind: u32, y: &[u8], u: &[u8], v: &[u8], px_err: f32,
                  hdr_out: &mut [pfv_mbhdr], coeff_out: &mut [i16]) -> io::Result<()> {
        let job = pfv_encode_job {
            kind, flags: 0, dst_slot: self.cur ^ 1, ref_slot: self.cur, px_err, reserved: 0,
            src_y: y.as_ptr(), src_u: u.as_ptr(), src_v: v.as_ptr(),
            hdr_out: hdr_out.as_mut_ptr(), coeff_out: coeff_out.as_mut_ptr(),
        };
        check(unsafe { pfv_encode_submit(self.ctx, &job, 1) })?;
        self.cur ^= 1;
        check(unsafe { pfv_sync(self.ctx) })
    }
}

/// Pinned output of the sparse encode seam; the device stores into it directly.
pub struct PinnedTokens {
    ptr: *mut u32,
    cap: usize,
}

impl PinnedTokens {
    pub fn new(nb: usize) -> io::Result<PinnedTokens> {
        let cap = nb * 256;                                // an RLE entry consumes at least one coefficient
        let mut p: *mut c_void = std::ptr::null_mut();
        check(unsafe { pfv_host_alloc(&mut p, (cap + PFV_TOKSTATS_WORDS) * 4) })?;
        Ok(PinnedTokens { ptr: p as *mut u32, cap })
    }
    pub fn stats(&self) -> &[u32] { unsafe { std::slice::from_raw_parts(self.ptr.add(self.cap), PFV_TOKSTATS_WORDS) } }
    /// (num_zeroes, coeff_size, coeff) of every RLESequence of the frame, in stream order
    pub fn entries(&self) -> impl Iterator<Item = (u8, u8, i16)> + '_ {
        let n = self.stats()[PFV_TOKSTATS_NTOK] as usize;
        unsafe { std::slice::from_raw_parts(self.ptr, n) }.iter().map(|t| ((t & 15) as u8, ((t >> 4) & 15) as u8, (t >> 16) as u16 as i16))
    }
}

impl Drop for PinnedTokens {
    fn drop(&mut self) { unsafe { pfv_host_free(self.ptr as *mut c_void) } }
}

impl CudaPlanes {
    /// Sparse variant of `encode`: replaces enc.rs:84-97 / :134-147 AND the rle_encode + update_table loops of
    /// write_iframe_packet / write_pframe_packet (enc.rs:256-262, :264-266 / :370-378, :380-382).  After the call
    /// `out.stats()[0..16] + out.stats()[16..32]` is the `symbol_table` handed to rle_create_huffman and `out.entries()` is the
    /// concatenation of `block_coeff` in macroblock order.
    pub fn encode_sparse(&mut self, kind: u32, y: &[u8], u: &[u8], v: &[u8], px_err: f32,
                         hdr_out: &mut [pfv_mbhdr], out: &mut PinnedTokens) -> io::Result<()> {
        let job = pfv_encode_job_sparse {
            kind, flags: 0, dst_slot: self.cur ^ 1, ref_slot: self.cur, px_err, tok_cap: out.cap as u32,
            src_y: y.as_ptr(), src_u: u.as_ptr(), src_v: v.as_ptr(),
            hdr_out: hdr_out.as_mut_ptr(), mb_off_out: std::ptr::null_mut(),
            tok_out: out.ptr, stats_out: unsafe { out.ptr.add(out.cap) },
        };
        check(unsafe { pfv_encode_submit_sparse(self.ctx, &job, 1) })?;
        self.cur ^= 1;
        check(unsafe { pfv_sync(self.ctx) })
    }
}

impl Drop for CudaPlanes {
    fn drop(&mut self) {
        unsafe { pfv_ctx_destroy(self.ctx) }
    }
}
