/*
 * pfv_b200.h — C ABI of the B200 (sm_100a) macroblock engine for Pretty Fast Video.
 *
 * This is the drop-in boundary for the per-macroblock hot path of pfv-rs 0.2.2
 * (codec 2.1.1).  The reference has no FFI/plugin interface; the seam this ABI
 * replaces is where `pfv_rs::dec::Decoder` / `pfv_rs::enc::Encoder` hand dense i16
 * coefficients, macroblock headers and u8 planes to the rayon macroblock loops of
 * `VideoPlane` (reference file:line cited per entry point, paths relative to the
 * reference root).  INTEGRATION.md shows the Rust `extern "C"` block and the
 * `#[cfg(feature = "cuda")]` call sites a maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns PFV_OK (0) or a negative
 *     pfv_status, and never unwinds.  pfv_last_error() gives a text for the calling
 *     thread's last failure.
 *   - a context belongs to ONE device and is NOT thread safe (the reference's
 *     Encoder/Decoder methods take `&mut self`); distinct contexts are independent.
 *   - the context owns device memory: a pool of padded frame slots (Y|U|V, geometry
 *     of VideoFrame::new_padded, src/frame.rs:28-49), coefficient/header staging,
 *     derived quantiser tables, three CUDA streams (H2D, compute, D2H).
 *   - all *_submit calls are asynchronous with respect to the host.  Host buffers
 *     passed to them must stay valid until pfv_sync() returns, and should come from
 *     pfv_host_alloc() (pinned) — pageable memory works but serialises the copies.
 *   - dense coefficient layout (src/dec.rs:258,376,450-517): NB*256 int16, macroblocks
 *     of Y then U then V in row-major order, each macroblock = sub-blocks
 *     (0,0),(8,0),(0,8),(8,8) (src/common.rs:145-149), each sub-block = 64
 *     coefficients in zig-zag scan order (src/dct.rs:88-99).
 */
#ifndef PFV_B200_H
#define PFV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFV_B200_ABI_VERSION 1

typedef enum pfv_status {
    PFV_OK = 0,
    PFV_ERR_BAD_ARG = -1,      /* reference: assert!/panic! on bad dims/state (src/enc.rs:38,76-80) */
    PFV_ERR_CUDA = -2,         /* a CUDA runtime/driver call failed; see pfv_last_error()            */
    PFV_ERR_NO_DEVICE = -3,    /* no usable sm_100 device: the engine has NO CPU fallback            */
    PFV_ERR_BAD_MV = -4,       /* motion vector leaves the padded plane (src/common.rs:258-259)     */
    PFV_ERR_NOMEM = -5,
    PFV_ERR_BAD_STREAM = -6,   /* container level: DecodeError::FormatError (src/dec.rs:31-35)      */
    PFV_ERR_BAD_VERSION = -7,  /* container level: DecodeError::VersionError                         */
    PFV_ERR_IO = -8,           /* container level: truncated stream (io::Error / DecodeError::IOError) */
    PFV_ERR_STATE = -9         /* e.g. encode after finish (src/enc.rs:80,130,176,183)               */
} pfv_status;

/* src/dec.rs:9-13 DeltaBlockHeader {mvec_x:i8, mvec_y:i8, has_coeff:bool}, padded to 4 bytes. */
typedef struct pfv_mbhdr {
    int8_t  mx;
    int8_t  my;
    uint8_t has_coeff;
    uint8_t reserved;          /* must be 0 on input; written 0 on output */
} pfv_mbhdr;

/* src/frame.rs:28-49 */
typedef struct pfv_geometry {
    uint32_t width, height;    /* visible luma                    */
    uint32_t cwidth, cheight;  /* visible chroma (w/2, h/2)       */
    uint32_t pw, ph;           /* padded luma plane               */
    uint32_t cpw, cph;         /* padded chroma plane             */
    uint32_t nb_y, nb_c, nb;   /* macroblocks: luma, one chroma plane, whole frame */
    uint32_t frame_bytes;      /* pw*ph + 2*cpw*cph               */
} pfv_geometry;

typedef struct pfv_ctx pfv_ctx;

enum { PFV_FRAME_I = 1, PFV_FRAME_P = 2 };

/* job.flags */
enum {
    PFV_JOB_DEVICE_PTRS = 1u   /* hdr/coeff/src pointers are DEVICE pointers already resident in HBM
                                  (no H2D is issued; out pointers, if non-NULL, are still host)       */
};

/*
 * One frame of decode work.  Replaces, for one frame, the three `deserialize_plane` calls of
 * Decoder::decode_iframe (src/dec.rs:303-310 -> VideoPlane::decode_plane_into, src/common.rs:477-496)
 * or the three `deserialize_plane_delta` calls of Decoder::decode_pframe (src/dec.rs:425-432 ->
 * VideoPlane::decode_plane_delta_into, src/common.rs:498-521).  The two-phase "read every block of
 * the old plane, then write" rule of common.rs:498-521 is kept by requiring dst_slot != ref_slot.
 */
typedef struct pfv_decode_job {
    uint32_t kind;             /* PFV_FRAME_I or PFV_FRAME_P                                         */
    uint32_t flags;
    uint32_t dst_slot;         /* frame slot that receives the decoded padded frame                  */
    uint32_t ref_slot;         /* P: slot holding the previous frame (Decoder.framebuffer)           */
    uint8_t  qidx[3];          /* q-table index for Y,U,V (src/dec.rs:244-246, 346-348)              */
    uint8_t  reserved;
    const pfv_mbhdr *hdr;      /* P: nb headers; ignored for I                                       */
    const int16_t   *coeff;    /* nb*256 coefficients (skipped P macroblocks are never read)         */
    uint8_t *out_y, *out_u, *out_v; /* optional host destinations for the visible crop
                                  (retframe blit, src/dec.rs:195-197): tight w*h, w/2*h/2, w/2*h/2.
                                  All three NULL = leave the frame on the device only.               */
} pfv_decode_job;

/*
 * One frame of encode work.  Replaces the per-plane encode + closed-loop decode + blit of
 * Encoder::encode_iframe (src/enc.rs:84-97: VideoPlane::encode_plane src/common.rs:351-386, then
 * decode_plane src/common.rs:423-446) or Encoder::encode_pframe (src/enc.rs:134-147:
 * encode_plane_delta src/common.rs:388-421 incl. block_search :154-204, then decode_plane_delta
 * :448-475).  q-tables used are fixed by kind as in the reference: I = tables 0,1,1; P = 2,3,3.
 */
typedef struct pfv_encode_job {
    uint32_t kind;             /* PFV_FRAME_I or PFV_FRAME_P                                         */
    uint32_t flags;
    uint32_t dst_slot;         /* receives the reconstructed frame (new Encoder.prev_frame)          */
    uint32_t ref_slot;         /* P: slot holding Encoder.prev_frame; must differ from dst_slot      */
    float    px_err;           /* P: quality*1.5 (src/enc.rs:41); skip iff SSD <= px_err^2*256       */
    uint32_t reserved;
    const uint8_t *src_y, *src_u, *src_v;  /* tight planes w*h, w/2*h/2, w/2*h/2 (VideoFrame)        */
    pfv_mbhdr *hdr_out;        /* P: nb headers (host, or device with PFV_JOB_DEVICE_PTRS); may be NULL for I */
    int16_t   *coeff_out;      /* nb*256 coefficients.  P: entries of skipped macroblocks are not written */
} pfv_encode_job;

/* ---- library level --------------------------------------------------------------------------- */
int         pfv_abi_version(void);
const char *pfv_last_error(void);
int         pfv_device_count(void);                         /* <0 on error */
void        pfv_geometry_for(uint32_t width, uint32_t height, pfv_geometry *out);   /* src/frame.rs:28-49 */
/* Encoder::new q-table derivation (src/enc.rs:40-51): out[0..3] = intra_l, intra_c, inter_l, inter_c */
int         pfv_make_qtables(int quality, int32_t out[4][64], float *px_err_out);

/* pinned host memory for job buffers */
int  pfv_host_alloc(void **out, size_t bytes);
void pfv_host_free(void *p);

/* ---- context --------------------------------------------------------------------------------- */
/*
 * Creates the engine state for one stream geometry on `device`.
 *   qtables/nq : the stream's q-tables in raster order (header tables, src/dec.rs:96-111; for an
 *                encoder the four tables of src/enc.rs:48-51 in header order, src/enc.rs:202-216).
 *   nslots     : frame slots in the pool (>= 2).  Every slot starts as VideoFrame::new_padded:
 *                Y = 0, U = V = 128 (src/frame.rs:38-43).
 *   max_jobs   : largest number of jobs one *_submit call will carry (sizes the staging rings).
 *   ext_stream : NULL, or a cudaStream_t the compute work is launched on (lets a host program time
 *                or order against the kernels); copies still use the context's own copy streams
 *                unless PFV_JOB_DEVICE_PTRS makes them unnecessary.
 */
int  pfv_ctx_create(int device, uint32_t width, uint32_t height,
                    const int32_t (*qtables)[64], uint32_t nq,
                    uint32_t nslots, uint32_t max_jobs, void *ext_stream, pfv_ctx **out);
void pfv_ctx_destroy(pfv_ctx *ctx);
int  pfv_ctx_geometry(const pfv_ctx *ctx, pfv_geometry *out);
int  pfv_sync(pfv_ctx *ctx);                                /* waits for all submitted work; returns the
                                                               first deferred error (e.g. PFV_ERR_BAD_MV) */
int  pfv_slot_reset(pfv_ctx *ctx, uint32_t slot);           /* back to Y=0, U=V=128                         */
/* whole padded frame (Y|U|V, geometry.frame_bytes) to/from a slot; synchronous.  Test and debug aid.        */
int  pfv_slot_read(pfv_ctx *ctx, uint32_t slot, uint8_t *frame_out);
int  pfv_slot_write(pfv_ctx *ctx, uint32_t slot, const uint8_t *frame_in);
/* visible crop of a slot into tight planes (src/dec.rs:195-197); asynchronous (D2H stream).                 */
int  pfv_slot_read_visible(pfv_ctx *ctx, uint32_t slot, uint8_t *y, uint8_t *u, uint8_t *v);
/* device address of a slot's padded frame, for consumers that keep frames on the GPU                        */
int  pfv_slot_device_ptr(pfv_ctx *ctx, uint32_t slot, void **out);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* Jobs of one call must be independent (no job's ref_slot is another job's dst_slot); calls are
 * ordered with respect to each other, so frame k+1 of a GOP goes in a later call than frame k. */
int  pfv_decode_submit(pfv_ctx *ctx, const pfv_decode_job *jobs, uint32_t njobs);
int  pfv_encode_submit(pfv_ctx *ctx, const pfv_encode_job *jobs, uint32_t njobs);

/* number of kernel launches this context has issued (bench.py's gpu_launches) */
uint64_t pfv_ctx_launch_count(const pfv_ctx *ctx);
/* device time in ms of the compute-stream work between the first and last kernel of the most recent
 * *_submit call (CUDA events recorded on the compute stream around the launches); valid after pfv_sync. */
int  pfv_ctx_last_kernel_ms(pfv_ctx *ctx, float *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* PFV_B200_H */
