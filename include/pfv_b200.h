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
    PFV_JOB_DEVICE_PTRS = 1u,  /* hdr/coeff/src pointers are DEVICE pointers already resident in HBM
                                  (no H2D is issued; out pointers, if non-NULL, are still host)       */
    PFV_JOB_SRC_RGB = 2u,      /* encode jobs: src_y points at packed RGB8 (w*h*3 bytes), src_u/src_v are
                                  ignored; the engine converts on the device exactly like the reference's
                                  load_frame + VideoFrame::from_planes (src/lib.rs:337-363, src/frame.rs:51-60,
                                  reduce src/common.rs:523-536)                                        */
    PFV_JOB_DENSE = 4u         /* decode key-frame jobs, a HINT: most 8x8 sub-blocks of this frame carry AC
                                  coefficients (the caller's entropy decoder knows: more than ~6 non-zero
                                  coefficients per sub-block).  Such frames skip the classify/compact staging that
                                  pays on ordinary streams.  Results never depend on the hint; sparse jobs set it
                                  themselves from their token count.                                   */
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
                                  All three NULL = leave the frame on the device only.  When the
                                  plane WIDTHS need no padding (w % 16 == 0 and (w/2) % 16 == 0, e.g.
                                  1920, 3840) and out_u = out_y + pw*ph, out_v = out_u + cpw*cph (the
                                  padded plane sizes of pfv_geometry), the picture travels as ONE copy
                                  instead of three - ~10 % more pictures per second over PCIe.       */
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
/* Colour / format steps next to the path: the visible crop of a slot (src/dec.rs:195-197) as packed
 * RGB8, w*h*3 bytes, converted on the device exactly like the reference's save_frame (src/lib.rs:365-395: chroma
 * `double`d, src/common.rs:538-556; JPEG YCbCr in f32; `as u8`).  pfv_slot_read_rgb copies it to host memory
 * (asynchronous, D2H stream, like pfv_slot_read_visible); pfv_slot_convert_rgb leaves it in device memory the caller
 * owns ("decode straight into a game-ready texture", README.md:20), ordered on the compute stream. */
int  pfv_slot_read_rgb(pfv_ctx *ctx, uint32_t slot, uint8_t *rgb_host);
int  pfv_slot_convert_rgb(pfv_ctx *ctx, uint32_t slot, void *rgb_device);
/* the same for n slots in one launch per 64 pictures: picture i goes to rgb_device + i * stride (stride >= w*h*3) */
int  pfv_slots_convert_rgb(pfv_ctx *ctx, const uint32_t *slots, uint32_t n, void *rgb_device, size_t stride);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* Jobs of one call must be independent (no job's ref_slot is another job's dst_slot); calls are
 * ordered with respect to each other, so frame k+1 of a GOP goes in a later call than frame k. */
int  pfv_decode_submit(pfv_ctx *ctx, const pfv_decode_job *jobs, uint32_t njobs);
int  pfv_encode_submit(pfv_ctx *ctx, const pfv_encode_job *jobs, uint32_t njobs);

/*
 * Sparse coefficient transport.  What the entropy decoder's symbol loop produces
 * (src/dec.rs:261-296, :378-417) before it is scattered into the dense Vec<i16>: for every macroblock the
 * list of its non-zero coefficients.  tok[i] = (position << 16) | uint16(value), position = index 0..255 inside
 * the macroblock (sub-block * 64 + scan position); the tokens of macroblock m are tok[mb_off[m] .. mb_off[m+1]).
 * The engine copies only the tokens over PCIe and expands them on the device (expand_tokens_kernel) into the same
 * dense layout pfv_decode_submit takes, so results are identical by construction.  A P macroblock with
 * has_coeff = 0 must own no tokens.  Positions may repeat; the last token of a position wins (src/dec.rs:288).
 */
typedef struct pfv_decode_job_sparse {
    uint32_t kind;             /* PFV_FRAME_I or PFV_FRAME_P                                         */
    uint32_t flags;            /* must be 0 (host pointers only)                                     */
    uint32_t dst_slot, ref_slot;
    uint8_t  qidx[3];
    uint8_t  reserved;
    const pfv_mbhdr *hdr;      /* P: nb headers                                                      */
    const uint32_t  *mb_off;   /* nb + 1 offsets into tok, mb_off[0] = 0, mb_off[nb] = ntok          */
    const uint32_t  *tok;      /* ntok tokens                                                        */
    uint32_t ntok;
    uint32_t reserved2;
    uint8_t *out_y, *out_u, *out_v;
} pfv_decode_job_sparse;
int  pfv_decode_submit_sparse(pfv_ctx *ctx, const pfv_decode_job_sparse *jobs, uint32_t njobs);

/*
 * Sparse ENCODE transport (SURVEY §8 f2 "on encode" + f4 "GPU-side RLE").  Same work as pfv_encode_submit, but what
 * comes back is the frame's RLE sequence instead of nb*256 dense coefficients: the run-length pass of
 * rle_encode (src/rle.rs:9-39, called per macroblock by src/enc.rs:256-262 / :370-378) and the symbol count of
 * update_table (src/rle.rs:41-47) run on the device.  tok_out[i] = run | size << 4 | uint16(value) << 16 is one
 * RLESequence {num_zeroes, coeff_size, coeff} (src/rle.rs:3-7), in stream order (macroblocks in frame order, P
 * macroblocks without coefficients contribute nothing).  stats_out receives PFV_TOKSTATS_WORDS words:
 *   [0..15]  how often each num_zeroes symbol occurs      [16..31] how often each coeff_size symbol occurs
 *   [PFV_TOKSTATS_NTOK]  entries the frame produced       [PFV_TOKSTATS_FLAGS]  PFV_TOKFLAG_* bits
 * which is everything rle_create_huffman (src/rle.rs:49-66) and the bit writer need: pfv_packet_encode_tokens turns
 * (hdr, tok, stats) into the packet pfv_packet_encode would have produced from the dense coefficients.
 * tok_out, stats_out and mb_off_out (optional: nb+1 offsets of each macroblock's entries) are written by the device
 * itself with exactly the bytes the frame needs, so they must be device-accessible: pinned host memory
 * (pfv_host_alloc, cudaHostAlloc, cudaHostRegister) or device memory.  Pageable memory is refused (PFV_ERR_BAD_ARG).
 * nb*256 entries always suffice; a smaller tok_cap that overflows stores tok_cap entries and sets PFV_TOKFLAG_OVERFLOW.
 */
#define PFV_TOKSTATS_WORDS   36u
#define PFV_TOKSTATS_NTOK    32u
#define PFV_TOKSTATS_FLAGS   33u
#define PFV_TOKFLAG_OVERFLOW 1u     /* the frame produced more than tok_cap entries                                  */
#define PFV_TOKFLAG_RANGE    2u     /* a coefficient needs more than 15 bits: not representable (src/rle.rs:24, :43) */
typedef struct pfv_encode_job_sparse {
    uint32_t kind;             /* PFV_FRAME_I or PFV_FRAME_P                                         */
    uint32_t flags;            /* PFV_JOB_SRC_RGB, PFV_JOB_DEVICE_PTRS (sources and hdr_out on the device) */
    uint32_t dst_slot, ref_slot;
    float    px_err;
    uint32_t tok_cap;          /* capacity of tok_out in entries                                     */
    const uint8_t *src_y, *src_u, *src_v;
    pfv_mbhdr *hdr_out;        /* P: nb headers (any host memory, or device with PFV_JOB_DEVICE_PTRS) */
    uint32_t  *mb_off_out;     /* optional, nb + 1                                                   */
    uint32_t  *tok_out;        /* tok_cap entries                                                    */
    uint32_t  *stats_out;      /* PFV_TOKSTATS_WORDS words                                           */
} pfv_encode_job_sparse;
int  pfv_encode_submit_sparse(pfv_ctx *ctx, const pfv_encode_job_sparse *jobs, uint32_t njobs);

/* Every *_submit (and pfv_slot_read_visible) call takes the next submit id (1, 2, ...).  pfv_ctx_wait_submit blocks
 * until the device-to-host copies of that submit have landed (it must be one of the 64 most recent ids); unlike
 * pfv_sync it does not wait for later submits, which is what a read-ahead decoder needs. */
uint64_t pfv_ctx_last_submit_id(const pfv_ctx *ctx);
int      pfv_ctx_wait_submit(pfv_ctx *ctx, uint64_t submit_id);

/* number of kernel launches this context has issued (bench.py's gpu_launches) */
uint64_t pfv_ctx_launch_count(const pfv_ctx *ctx);
/* device time in ms of the compute-stream work between the first and last kernel of the most recent
 * *_submit call (CUDA events recorded on the compute stream around the launches); valid after pfv_sync.  Timing costs
 * two driver calls per submit, so it is off until this function is called once: that first call switches it on and
 * returns PFV_ERR_STATE. */
int  pfv_ctx_last_kernel_ms(pfv_ctx *ctx, float *ms_out);


/* ---- host side of the codec: container + entropy layer + the reference's Encoder / Decoder ---------------- */
/*
 * Everything below runs on the host (north_star keeps src/huffman.rs and src/rle.rs there) and is built on the
 * hot-path entry points above.  pfv_decoder / pfv_encoder mirror pfv_rs::dec::Decoder (src/dec.rs:15-224) and
 * pfv_rs::enc::Encoder (src/enc.rs:12-188): same constructor arguments, same packet semantics, same error
 * classes (PFV_ERR_BAD_STREAM = DecodeError::FormatError, PFV_ERR_BAD_VERSION = VersionError, PFV_ERR_IO =
 * io::Error / IOError).  The reader R / writer W of the Rust generics are an in-memory byte range / a growable
 * byte buffer by default, or the caller's callbacks (pfv_decoder_open_reader, pfv_encoder_set_writer).
 * encode_iframe / encode_pframe return once the planes have been copied; the GPU submit, the entropy coding and the
 * packet append happen on internal threads, and an error there comes back from the next call on the encoder.
 */
typedef struct pfv_stream_info {            /* file header, src/dec.rs:38-118 / src/enc.rs:190-219 */
    uint32_t version;                        /* 211 */
    uint32_t width, height, framerate;
    uint32_t num_qtables;
    uint64_t first_packet;                   /* byte offset of the first packet = Decoder.reset_pos */
} pfv_stream_info;

typedef struct pfv_packet {                 /* one packet of the container, src/dec.rs:179-220 */
    uint8_t  type;                           /* 0 EOF, 1 I-frame (len 0 = drop frame), 2 P-frame, other = skipped */
    uint8_t  reserved[3];
    uint32_t len;                            /* payload bytes */
    uint64_t payload;                        /* byte offset of the payload in the stream */
} pfv_packet;

/* Header parse (Decoder::new).  qtables_out may be NULL; otherwise it receives min(num_qtables, qcap) tables. */
int pfv_stream_parse_header(const uint8_t *data, size_t len, pfv_stream_info *info,
                            int32_t (*qtables_out)[64], uint32_t qcap);
/* Packet scan from `offset` (no entropy decoding: u8 type + u32 len per packet).  Stops after the EOF packet, at
 * the end of the data, or when `cap` packets are written; *n_out = packets written, *truncated_out = 1 if the data
 * ends inside a packet (the reference would fail with an io::Error when it reaches it).  This is what a host uses
 * to find GOP boundaries for frame-parallel sharding. */
int pfv_stream_index(const uint8_t *data, size_t len, uint64_t offset, pfv_packet *out, uint32_t cap,
                     uint32_t *n_out, int *truncated_out);

/* Entropy decode of one frame payload into the sparse seam (Decoder::decode_iframe src/dec.rs:226-296 /
 * decode_pframe :328-417, without the macroblock loops).  kind = PFV_FRAME_I / PFV_FRAME_P.  hdr_out: nb headers
 * (P only, may be NULL for I).  mb_off_out: nb+1.  tok_out: capacity tok_cap tokens (nb*256 always suffices). */
int pfv_packet_decode(const pfv_geometry *g, uint32_t kind, const uint8_t *payload, size_t len, uint8_t qidx_out[3],
                      pfv_mbhdr *hdr_out, uint32_t *mb_off_out, uint32_t *tok_out, uint32_t tok_cap, uint32_t *ntok_out);
/* Entropy encode of one frame from the dense seam (Encoder::write_iframe_packet src/enc.rs:237-330 /
 * write_pframe_packet :332-481): writes the 5-byte packet header + payload to out (capacity cap), *len_out = bytes. */
int pfv_packet_encode(const pfv_geometry *g, uint32_t kind, const pfv_mbhdr *hdr, const int16_t *coeff,
                      uint8_t *out, size_t cap, size_t *len_out);
size_t pfv_packet_encode_bound(const pfv_geometry *g);   /* a capacity that always suffices */
/* The same packet from the sparse encode seam: tok/stats as pfv_encode_submit_sparse delivers them (the Huffman tree
 * comes from the two histograms, src/rle.rs:49-66; the payload size is known before the first bit is written). */
int pfv_packet_encode_tokens(const pfv_geometry *g, uint32_t kind, const pfv_mbhdr *hdr, const uint32_t *tok,
                             const uint32_t *stats, uint8_t *out, size_t cap, size_t *len_out);
/* Host restatement of the device tokenizer (the run-length pass alone): dense coefficients -> tok/stats/mb_off in
 * the format above.  What a host without the sparse seam runs before pfv_packet_encode_tokens; mb_off_out may be NULL. */
int pfv_packet_tokenize(const pfv_geometry *g, uint32_t kind, const pfv_mbhdr *hdr, const int16_t *coeff,
                        uint32_t *tok_out, uint32_t tok_cap, uint32_t *mb_off_out, uint32_t *stats_out);
/* the largest number of tokens pfv_packet_decode can emit for this payload (<= nb*256): sizes tok_out */
uint32_t pfv_packet_token_bound(const pfv_geometry *g, const uint8_t *payload, size_t len);

/* -- Decoder (src/dec.rs) -------------------------------------------------------------------------------------- */
typedef struct pfv_decoder pfv_decoder;
/* Decoder::new(reader, num_threads) (src/dec.rs:38-134).  `data` stays owned by the caller and must outlive the
 * decoder.  num_threads sizes the host entropy-decode pool (the reference's rayon pool is replaced by the GPU).
 * read_ahead = how many frames may be in flight (entropy decoded / on the GPU) ahead of the one being returned;
 * 0 picks a default.  The decoded pictures are identical for any value. */
int  pfv_decoder_open(const uint8_t *data, size_t len, int device, uint32_t num_threads, uint32_t read_ahead,
                      pfv_decoder **out);
/* The same for a caller that holds a reader (Decoder<R: Read + Seek>, src/dec.rs:15-28) instead of a byte range: `read`
 * is called until it returns 0 (end of stream; < 0 = error -> PFV_ERR_IO) and the decoder keeps the bytes. */
typedef long long (*pfv_read_fn)(void *user, uint8_t *buf, size_t cap);
int  pfv_decoder_open_reader(pfv_read_fn read, void *user, int device, uint32_t num_threads, uint32_t read_ahead,
                             pfv_decoder **out);
void pfv_decoder_close(pfv_decoder *d);
uint32_t pfv_decoder_width(const pfv_decoder *d);        /* src/dec.rs:136 */
uint32_t pfv_decoder_height(const pfv_decoder *d);       /* src/dec.rs:140 */
uint32_t pfv_decoder_framerate(const pfv_decoder *d);    /* src/dec.rs:144 */
int  pfv_decoder_reset(pfv_decoder *d);                  /* src/dec.rs:148-152: rewinds; the framebuffer is NOT cleared */
/* Decoder::advance_frame (src/dec.rs:169-224): consumes packets up to and including the next frame or drop frame.
 * Returns 1 = Ok(true), 0 = Ok(false) (EOF packet reached), <0 = error.  *got_frame = 1 when a picture was
 * produced (the reference's onvideo callback fired); then *y,*u,*v point at the tight visible planes (retframe,
 * src/dec.rs:195-197) in pinned host memory owned by the decoder, valid until the next call on this decoder. */
int  pfv_decoder_advance_frame(pfv_decoder *d, int *got_frame, const uint8_t **y, const uint8_t **u, const uint8_t **v);
/* Decoder::advance_delta (src/dec.rs:154-167): onvideo is called once per produced frame. */
typedef void (*pfv_onvideo_fn)(void *user, const uint8_t *y, const uint8_t *u, const uint8_t *v);
int  pfv_decoder_advance_delta(pfv_decoder *d, double delta, pfv_onvideo_fn onvideo, void *user);
/* the engine context under the decoder (e.g. to read the padded framebuffer slot in tests) and the slot that
 * currently holds Decoder.framebuffer */
pfv_ctx *pfv_decoder_ctx(pfv_decoder *d);
uint32_t pfv_decoder_framebuffer_slot(const pfv_decoder *d);

/* -- Encoder (src/enc.rs) -------------------------------------------------------------------------------------- */
typedef struct pfv_encoder pfv_encoder;
/* Encoder::new(writer, width, height, framerate, quality, num_threads) (src/enc.rs:37-73); the header is written
 * at once.  num_threads sizes the host entropy-coding pool. */
int  pfv_encoder_open(uint32_t width, uint32_t height, uint32_t framerate, int quality, uint32_t num_threads,
                      int device, pfv_encoder **out);
void pfv_encoder_close(pfv_encoder *e);                  /* Drop: finishes the stream if finish() was not called */
/* Encoder<W: Write> (src/enc.rs:12-26): hand the stream to a writer instead of keeping it in memory.  Call right after
 * pfv_encoder_open: the header goes out at once (from the calling thread), every packet as soon as it is finished, in
 * stream order, one call at a time, from an internal writer thread.  The callback returns 0, or non-zero for an I/O error
 * (-> PFV_ERR_IO from the next call on the encoder).  With a writer set pfv_encoder_bytes returns an empty range. */
typedef int (*pfv_write_fn)(void *user, const uint8_t *data, size_t len);
int  pfv_encoder_set_writer(pfv_encoder *e, pfv_write_fn writer, void *user);
/* src/enc.rs:75-123 / :125-173.  y,u,v: tight planes w*h, w/2*h/2, w/2*h/2 (VideoFrame); read before return. */
int  pfv_encoder_encode_iframe(pfv_encoder *e, const uint8_t *y, const uint8_t *u, const uint8_t *v);
int  pfv_encoder_encode_pframe(pfv_encoder *e, const uint8_t *y, const uint8_t *u, const uint8_t *v);
int  pfv_encoder_encode_dropframe(pfv_encoder *e);       /* src/enc.rs:175-180 */
int  pfv_encoder_finish(pfv_encoder *e);                 /* src/enc.rs:182-188 */
/* the writer: everything written so far (all frames submitted are flushed first).  Pointer valid until the next
 * call on this encoder. */
int  pfv_encoder_bytes(pfv_encoder *e, const uint8_t **data, size_t *len);
pfv_ctx *pfv_encoder_ctx(pfv_encoder *e);
uint32_t pfv_encoder_prev_frame_slot(const pfv_encoder *e);

#ifdef __cplusplus
}
#endif
#endif /* PFV_B200_H */
