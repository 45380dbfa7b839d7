/*
 * pfv_oracle.h — CPU oracle for the Pretty-Fast-Video macroblock hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (pretty_fast_video_b200/,
 * include/) may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * PARITY UNPINNED: the reference (pfv-rs 0.2.2, Rust) cannot be compiled in this
 * environment (no cargo/rustc) and every fixture in it is a git-LFS pointer stub,
 * and its tests hold no golden outputs for this path.  This file is
 * a plain-C restatement of the reference algorithm, function by function, with the
 * reference file:line each one follows given in pfv_oracle.c.  It is cross-checked
 * against an independent pure-Python restatement (oracle/pfv_ref.py) and the derived
 * known-answer vectors of SURVEY.md Appendix C (tests/test_oracle_kat.py).
 */
#ifndef PFV_ORACLE_H
#define PFV_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* dec.rs:9-13 DeltaBlockHeader, padded to 4 bytes so it matches the product's pfv_mbhdr. */
typedef struct {
    int8_t  mx;
    int8_t  my;
    uint8_t has_coeff;
    uint8_t reserved;
} pfvo_mbhdr;

/* frame.rs:28-49 geometry of a padded frame. */
typedef struct {
    uint32_t width, height;          /* visible luma size                     */
    uint32_t cwidth, cheight;        /* visible chroma size (w/2, h/2)        */
    uint32_t pw, ph;                 /* padded luma plane                     */
    uint32_t cpw, cph;               /* padded chroma plane                   */
    uint32_t nb_y, nb_c;             /* macroblocks per luma / chroma plane   */
    uint32_t nb;                     /* nb_y + 2*nb_c                         */
} pfvo_geometry;

void pfvo_geometry_for(uint32_t width, uint32_t height, pfvo_geometry *g);

/* --- tables (dct.rs:1-47) ------------------------------------------------------ */
const int32_t *pfvo_dct_scale_factor(void);
const int32_t *pfvo_q_table_intra(void);
const int32_t *pfvo_q_table_inter(void);
const uint8_t *pfvo_zigzag(void);
const uint8_t *pfvo_inv_zigzag(void);

/* enc.rs:40-51: out[0]=intra_l out[1]=intra_c out[2]=inter_l out[3]=inter_c (header order enc.rs:202-216) */
void  pfvo_make_qtables(int quality, int32_t out[4][64]);
float pfvo_px_err(int quality);

/* --- 1-D / 8x8 primitives -------------------------------------------------------- */
void pfvo_fdct8(int32_t v[8]);                                         /* dct.rs:176-239 */
void pfvo_idct8(int32_t v[8]);                                         /* dct.rs:241-293 */
void pfvo_quant_encode(const int32_t m[64], const int32_t q[64], int16_t out[64]);   /* dct.rs:88-99 */
void pfvo_quant_decode(const int16_t c[64], const int32_t q[64], int32_t m[64]);     /* dct.rs:75-86 */
void pfvo_encode_subblock(const uint8_t px[64], const int32_t q[64], int16_t out[64]);        /* common.rs:287-298 */
void pfvo_encode_subblock_delta(const int16_t d[64], const int32_t q[64], int16_t out[64]);   /* common.rs:300-311 */
void pfvo_decode_subblock(const int16_t c[64], const int32_t q[64], uint8_t out[64]);         /* common.rs:313-325 */

/* --- macroblock level ------------------------------------------------------------- */
void pfvo_encode_block(const uint8_t px[256], const int32_t q[64], int16_t out[256]);         /* common.rs:141-152 */
void pfvo_decode_block(const int16_t c[256], const int32_t q[64], uint8_t out[256]);          /* common.rs:238-252 */
/* common.rs:154-204; ref is a padded plane rw x rh. Returns best error; writes dx,dy,best 16x16. */
float pfvo_block_search(const uint8_t src[256], const uint8_t *ref, int rw, int rh,
                        int cx, int cy, int stepsize, int *dx, int *dy, uint8_t best[256]);
/* common.rs:206-236 */
void pfvo_encode_block_delta(const uint8_t src[256], const uint8_t *ref, int rw, int rh,
                             int bx, int by, const int32_t q[64], float px_err,
                             pfvo_mbhdr *hdr, int16_t out[256]);
/* common.rs:254-285 */
void pfvo_decode_block_delta(const pfvo_mbhdr *hdr, const int16_t c[256], const uint8_t *ref, int rw, int rh,
                             int bx, int by, const int32_t q[64], uint8_t out[256]);

/* --- plane level (MB loops; nthreads mirrors the rayon pool size) ----------------- */
/* common.rs:351-386: src is tight w x h; coeff_out holds blocks*256 i16 (row-major MBs). */
void pfvo_encode_plane(const uint8_t *src, int w, int h, const int32_t q[64], uint8_t clear_color,
                       int16_t *coeff_out, int nthreads);
/* common.rs:388-421: ref is the padded plane (pw x ph). */
void pfvo_encode_plane_delta(const uint8_t *src, int w, int h, const uint8_t *ref,
                             const int32_t q[64], float px_err, uint8_t clear_color,
                             pfvo_mbhdr *hdr_out, int16_t *coeff_out, int nthreads);
/* common.rs:477-496: target is padded pw x ph. */
void pfvo_decode_plane_into(const int16_t *coeff, int pw, int ph, const int32_t q[64],
                            uint8_t *target, int nthreads);
/* common.rs:498-521: refplane is read (old) then overwritten (two-phase). */
void pfvo_decode_plane_delta_into(const pfvo_mbhdr *hdr, const int16_t *coeff, int pw, int ph,
                                  const int32_t q[64], uint8_t *refplane, int nthreads);

/* --- frame level: the seam (dense coefficients <-> planes), no entropy coding ----- */
/* State frame = padded Y | U | V, contiguous; initial state Y=0, U=V=128 (frame.rs:38-43). */
size_t pfvo_frame_bytes(const pfvo_geometry *g);
void   pfvo_frame_init(const pfvo_geometry *g, uint8_t *frame);
/* dec.rs:298-310: coeff = nb*256 i16 (Y MBs, U MBs, V MBs). */
void   pfvo_decode_iframe_coeffs(const pfvo_geometry *g, const int32_t (*qtables)[64], const uint8_t qidx[3],
                                 const int16_t *coeff, uint8_t *frame, int nthreads);
/* dec.rs:419-432 */
void   pfvo_decode_pframe_coeffs(const pfvo_geometry *g, const int32_t (*qtables)[64], const uint8_t qidx[3],
                                 const pfvo_mbhdr *hdr, const int16_t *coeff, uint8_t *frame, int nthreads);
/* enc.rs:84-97: y/u/v tight; writes coefficients and updates prev_frame (closed-loop recon). */
void   pfvo_encode_iframe_coeffs(const pfvo_geometry *g, const int32_t (*qtables)[64],
                                 const uint8_t *y, const uint8_t *u, const uint8_t *v,
                                 int16_t *coeff_out, uint8_t *prev_frame, int nthreads);
/* enc.rs:134-147 */
void   pfvo_encode_pframe_coeffs(const pfvo_geometry *g, const int32_t (*qtables)[64], float px_err,
                                 const uint8_t *y, const uint8_t *u, const uint8_t *v,
                                 pfvo_mbhdr *hdr_out, int16_t *coeff_out, uint8_t *prev_frame, int nthreads);
/* dec.rs:195-197: crop the padded state into tight planes. */
void   pfvo_crop_frame(const pfvo_geometry *g, const uint8_t *frame, uint8_t *y, uint8_t *u, uint8_t *v);

/* --- colour / format helpers next to the path ------------------------ */
void pfvo_plane_reduce(const uint8_t *src, int w, int h, uint8_t *dst);                       /* common.rs:523-536 */
void pfvo_plane_double(const uint8_t *src, int w, int h, uint8_t *dst);                       /* common.rs:538-556 */
/* rgb: packed w*h*3; y: w*h; u,v: (w/2)*(h/2) */
void pfvo_rgb_to_yuv420(const uint8_t *rgb, int w, int h, uint8_t *y, uint8_t *u, uint8_t *v);   /* lib.rs:337-363 + frame.rs:51-60 */
void pfvo_yuv420_to_rgb(const uint8_t *y, const uint8_t *u, const uint8_t *v, int w, int h, uint8_t *rgb);   /* lib.rs:365-395 */

/* --- entropy layer (rle.rs, huffman.rs) and container (enc.rs:190-481, dec.rs) ----- */
typedef struct pfvo_encoder pfvo_encoder;
typedef struct pfvo_decoder pfvo_decoder;

/* enc.rs:37-73 Encoder::new. Returns NULL on bad arguments. */
pfvo_encoder *pfvo_encoder_new(int width, int height, int framerate, int quality, int nthreads);
int  pfvo_encoder_encode_iframe(pfvo_encoder *e, const uint8_t *y, const uint8_t *u, const uint8_t *v);   /* enc.rs:75 */
int  pfvo_encoder_encode_pframe(pfvo_encoder *e, const uint8_t *y, const uint8_t *u, const uint8_t *v);   /* enc.rs:125 */
int  pfvo_encoder_encode_dropframe(pfvo_encoder *e);                                                     /* enc.rs:175 */
int  pfvo_encoder_finish(pfvo_encoder *e);                                                               /* enc.rs:182 */
const uint8_t *pfvo_encoder_bytes(const pfvo_encoder *e, size_t *len);
/* last encoded frame's seam data (for kernel-level parity tests) */
const int16_t    *pfvo_encoder_last_coeffs(const pfvo_encoder *e);
const pfvo_mbhdr *pfvo_encoder_last_headers(const pfvo_encoder *e);
const uint8_t    *pfvo_encoder_prev_frame(const pfvo_encoder *e);
void pfvo_encoder_free(pfvo_encoder *e);

/* dec.rs:38-134 Decoder::new over an in-memory stream.  *err: 0 ok, 1 FormatError, 2 VersionError, 3 IOError */
pfvo_decoder *pfvo_decoder_new(const uint8_t *data, size_t len, int nthreads, int *err);
int  pfvo_decoder_width(const pfvo_decoder *d);
int  pfvo_decoder_height(const pfvo_decoder *d);
int  pfvo_decoder_framerate(const pfvo_decoder *d);
int  pfvo_decoder_reset(pfvo_decoder *d);                                                                /* dec.rs:148 */
/* dec.rs:169-224. Returns 1 = more data (frame_out says whether a frame was produced), 0 = EOF, <0 = IO error.
 * y/u/v (tight) are written when a frame is produced. */
int  pfvo_decoder_advance_frame(pfvo_decoder *d, uint8_t *y, uint8_t *u, uint8_t *v, int *frame_out);
/* last decoded packet's seam data: kind 1 = I, 2 = P */
int  pfvo_decoder_last_kind(const pfvo_decoder *d);
const int16_t    *pfvo_decoder_last_coeffs(const pfvo_decoder *d);
const pfvo_mbhdr *pfvo_decoder_last_headers(const pfvo_decoder *d);
const uint8_t    *pfvo_decoder_last_qidx(const pfvo_decoder *d);
const uint8_t    *pfvo_decoder_framebuffer(const pfvo_decoder *d);
const int32_t    *pfvo_decoder_qtables(const pfvo_decoder *d, int *nq);
void pfvo_decoder_free(pfvo_decoder *d);

/* test_entropy / test_entropy_2 (lib.rs:96-239): RLE + Huffman + bit-pack a coefficient run and read it back.
 * Returns number of bytes the run coded to, or <0 on mismatch. `decoded` receives the round trip. */
long pfvo_entropy_roundtrip(const int16_t *data, size_t n, int16_t *decoded);

/* rle_encode (rle.rs:9-39) + update_table (rle.rs:41-47) of one coefficient slice: returns the sequence length, writes at
 * most cap entries, adds the symbol counts to table[16] (may be NULL). */
size_t pfvo_rle_encode(const int16_t *data, size_t n, uint8_t *num_zeroes, uint8_t *coeff_size, int16_t *coeff, size_t cap,
                       int32_t table[16]);

#ifdef __cplusplus
}
#endif
#endif
