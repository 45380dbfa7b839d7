// pfv_codec.cpp — host side of the codec: container, entropy layer and the reference's Decoder / Encoder objects,
// built on the hot-path C ABI (pfv_decode_submit_sparse / pfv_encode_submit).  No device code here.
//
// What it mirrors (paths relative to the reference root):
//   container      src/enc.rs:190-235 (header, EOF, drop packet), src/dec.rs:38-134 (header parse), :169-224 (packets)
//   I/P payloads   src/enc.rs:237-481 (write), src/dec.rs:226-296, :328-417 (read)
//   RLE            src/rle.rs:9-66          Huffman  src/huffman.rs:71-119, :156-217
//   bit order      bitstream-io 1.6 LittleEndian: value bits LSB first, write_signed(n) = n-bit two's complement
//   objects        pfv_rs::dec::Decoder src/dec.rs:15-224, pfv_rs::enc::Encoder src/enc.rs:12-188
//
// Design (not the reference's): the reference decodes one symbol at a time through a seekable BitReader
// (position_in_bits + read + seek_bits per symbol, src/huffman.rs:156-197) on the caller's thread and then runs the
// macroblock loops on a rayon pool.  Here the macroblock loops are GPU kernels, so the host's job is to keep them
// fed: packets are length-prefixed and their entropy state is per frame, so a pool of host threads entropy-decodes
// frames AHEAD of the one being returned (64-bit window, 11-bit code LUT, one refill per token), straight into the
// sparse token form that crosses PCIe; frames are submitted to the engine in stream order (P frames chain on the
// previous picture) and handed back in order from a ring of pinned pictures.  The encoder is the mirror image:
// kernels emit dense coefficients + headers into pinned rings, a pool entropy-codes finished frames while the GPU
// works on the next ones, packets are appended in order.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "pfv_internal.h"
#include "pfv_pool.h"

using pfv::set_error;
using pfv::Pool;

namespace {

const uint8_t kMagic[8] = {'P', 'F', 'V', 'I', 'D', 'E', 'O', 0};   // src/common.rs:1
const uint32_t kVersion = 211;                                      // src/common.rs:2

inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline void wr16(std::vector<uint8_t> &v, uint32_t x) { v.push_back((uint8_t)x); v.push_back((uint8_t)(x >> 8)); }
inline void wr32(uint8_t *p, uint32_t x) { p[0] = (uint8_t)x; p[1] = (uint8_t)(x >> 8); p[2] = (uint8_t)(x >> 16); p[3] = (uint8_t)(x >> 24); }

// ---------------------------------------------------------------------------------------------------
// Huffman code of one frame (src/huffman.rs:71-119): 16 symbols, weights from the packet
// ---------------------------------------------------------------------------------------------------
constexpr int LUT_BITS = 11;

struct Huffman {
    uint32_t val[16];          // code bits, first bit of the code in bit 0 (Code::append, src/huffman.rs:30-32)
    uint8_t  len[16];          // 0 = symbol unused (or the only symbol: zero-length code)
    int      nsym = 0;         // symbols with weight > 0
    int      only = -1;        // the single symbol when nsym == 1
    // tree for codes longer than LUT_BITS: node i has child[i][bit]; leaves carry sym >= 0
    int16_t  child[31][2];
    int8_t   sym[31];
    int      root = -1;
    uint8_t  lut[1 << LUT_BITS];   // (len << 4) | symbol, 0 = longer than LUT_BITS
    // a token is two symbols back to back (run length, value size: src/dec.rs:262-275); when both codes fit the
    // window one lookup resolves the pair: run | size << 4 | total length << 8, 0 = take the two-step path
    uint16_t pair[1 << LUT_BITS];

    void build_pairs()
    {
        for (uint32_t b = 0; b < (1u << LUT_BITS); b++) {
            pair[b] = 0;
            const uint8_t e1 = lut[b];
            if (!e1) continue;
            const uint32_t l1 = e1 >> 4;
            const uint8_t e2 = lut[b >> l1];
            if (!e2) continue;
            const uint32_t l2 = e2 >> 4;
            if (l1 + l2 > (uint32_t)LUT_BITS) continue;              // the second code would need bits the index does not hold
            pair[b] = (uint16_t)((e1 & 15u) | ((e2 & 15u) << 4) | ((l1 + l2) << 8));
        }
    }

    void build(const uint8_t table[16])
    {
        memset(val, 0, sizeof(val));
        memset(len, 0, sizeof(len));
        memset(lut, 0, sizeof(lut));
        nsym = 0; only = -1; root = -1;
        struct N { uint32_t freq; int id; };
        N list[16];
        int n = 0, nodes = 0;
        uint32_t freq[31];
        for (int ch = 0; ch < 16; ch++)
            if (table[ch] > 0) {
                sym[nodes] = (int8_t)ch; child[nodes][0] = child[nodes][1] = -1; freq[nodes] = table[ch];
                list[n++] = {table[ch], nodes++};
            }
        nsym = n;
        if (n == 0) return;
        std::stable_sort(list, list + n, [](const N &a, const N &b) { return a.freq > b.freq; });   // descending, stable
        while (n > 1) {
            const N a = list[--n], b = list[--n];                    // a = last (smallest), b = the one before it
            sym[nodes] = -1; child[nodes][0] = (int16_t)a.id; child[nodes][1] = (int16_t)b.id;   // left = a, right = b
            freq[nodes] = a.freq + b.freq;
            int pos = n;
            for (int i = 0; i < n; i++) if (freq[nodes] > list[i].freq) { pos = i; break; }      // get_insert_index
            for (int i = n; i > pos; i--) list[i] = list[i - 1];
            list[pos] = {freq[nodes], nodes};
            n++; nodes++;
        }
        root = list[0].id;
        if (nsym == 1) { only = sym[root]; return; }                // zero-length code (legal, src/huffman.rs:125-131)
        // assign_codes (src/huffman.rs:204-217), iteratively
        struct S { int node; uint32_t v; uint8_t l; };
        S stack[32];
        int sp = 0;
        stack[sp++] = {root, 0u, 0};
        while (sp) {
            const S s = stack[--sp];
            if (sym[s.node] >= 0) {
                val[sym[s.node]] = s.v; len[sym[s.node]] = s.l;
                if (s.l <= LUT_BITS)
                    for (uint32_t v = s.v; v < (1u << LUT_BITS); v += 1u << s.l) lut[v] = (uint8_t)((s.l << 4) | sym[s.node]);
                continue;
            }
            stack[sp++] = {child[s.node][1], s.v | (1u << s.l), (uint8_t)(s.l + 1)};
            stack[sp++] = {child[s.node][0], s.v, (uint8_t)(s.l + 1)};
        }
    }
};

// ---------------------------------------------------------------------------------------------------
// LSB-first bit reader over one payload
// ---------------------------------------------------------------------------------------------------
struct BitReader {
    const uint8_t *p;
    size_t   nbytes;
    uint64_t total;            // bits
    uint64_t pos = 0;          // bits consumed

    BitReader(const uint8_t *data, size_t n) : p(data), nbytes(n), total((uint64_t)n * 8) {}
    // at least 57 valid bits starting at the current position (zeros past the end)
    inline uint64_t peek() const
    {
        const size_t b = (size_t)(pos >> 3);
        uint64_t w = 0;
        if (b + 8 <= nbytes) memcpy(&w, p + b, 8);
        else for (size_t i = 0; b + i < nbytes && i < 8; i++) w |= (uint64_t)p[b + i] << (8 * i);
        return w >> (pos & 7);
    }
    inline bool overrun() const { return pos > total; }
    inline uint32_t take(int n) { const uint32_t v = (uint32_t)(peek() & ((1ull << n) - 1)); pos += (uint64_t)n; return v; }
};

inline int32_t sign_extend(uint32_t raw, int bits) { return (int32_t)(raw << (32 - bits)) >> (32 - bits); }

// one Huffman symbol out of window w; returns its length (bits consumed) or -1
inline int huff_symbol(const Huffman &h, uint64_t w, int &symbol)
{
    const uint8_t e = h.lut[w & ((1u << LUT_BITS) - 1)];
    if (e) { symbol = e & 15; return e >> 4; }
    int node = h.root, l = 0;                                        // read_slow, src/huffman.rs:125-154
    while (h.sym[node] < 0) {
        node = h.child[node][(w >> l) & 1];
        if (++l > 16) return -1;
    }
    symbol = h.sym[node];
    return l;
}

// Upper bound of the tokens a payload can emit.  A token that carries a value costs its two codes plus `size` >= 1
// value bits: at least 3 bits when the tree has two or more symbols (every code is then >= 1 bit long).  A tree with
// a single symbol s has zero-length codes: every token is (run s, size s) and costs s bits.
uint64_t token_bound(const uint8_t *payload, size_t len, uint32_t nb)
{
    const uint64_t all = (uint64_t)nb * 256;
    if (len < 19) return 1;
    int nsym = 0, only = 0;
    for (int i = 0; i < 16; i++) if (payload[i]) { nsym++; only = i; }
    const uint64_t bits = (uint64_t)(len - 19) * 8;
    const uint64_t bound = nsym >= 2 ? bits / 3 + 1 : (nsym == 1 && only > 0 ? bits / (uint64_t)only + 1 : 1);
    return std::min(all, bound);
}

struct PlaneDims { uint32_t pw, ph, bw, bh; };

void plane_dims(const pfv_geometry &g, PlaneDims out[3])
{
    out[0] = {g.pw, g.ph, g.pw / 16, g.ph / 16};
    out[1] = out[2] = {g.cpw, g.cph, g.cpw / 16, g.cph / 16};
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// container
// ---------------------------------------------------------------------------------------------------
extern "C" int pfv_stream_parse_header(const uint8_t *data, size_t len, pfv_stream_info *info, int32_t (*qtables_out)[64], uint32_t qcap)
{
    if (!data || !info) return set_error(PFV_ERR_BAD_ARG, "NULL argument");
    memset(info, 0, sizeof(*info));
    if (len < 8) return set_error(PFV_ERR_IO, "stream ends inside the magic (src/dec.rs:41-46)");
    if (memcmp(data, kMagic, 8) != 0) return set_error(PFV_ERR_BAD_STREAM, "bad magic: not a PFV stream (src/dec.rs:48-52)");
    if (len < 12) return set_error(PFV_ERR_IO, "stream ends inside the version field");
    info->version = rd32(data + 8);
    if (info->version != kVersion)
        return set_error(PFV_ERR_BAD_VERSION, "codec version %u, this decoder reads %u (src/dec.rs:55-60)", info->version, kVersion);
    if (len < 20) return set_error(PFV_ERR_IO, "stream ends inside the header");
    info->width = rd16(data + 12);
    info->height = rd16(data + 14);
    info->framerate = rd16(data + 16);
    info->num_qtables = rd16(data + 18);
    const size_t need = 20 + (size_t)info->num_qtables * 128;
    if (len < need) return set_error(PFV_ERR_IO, "stream ends inside the q-tables (src/dec.rs:96-111)");
    if (qtables_out)
        for (uint32_t t = 0; t < info->num_qtables && t < qcap; t++)
            for (int i = 0; i < 64; i++) qtables_out[t][i] = rd16(data + 20 + (size_t)t * 128 + i * 2);
    info->first_packet = need;
    return PFV_OK;
}

extern "C" int pfv_stream_index(const uint8_t *data, size_t len, uint64_t offset, pfv_packet *out, uint32_t cap, uint32_t *n_out,
                                int *truncated_out)
{
    if (!data || !n_out || (!out && cap)) return set_error(PFV_ERR_BAD_ARG, "NULL argument");
    uint32_t n = 0;
    int trunc = 0;
    uint64_t off = offset;
    while (n < cap) {
        if (off == len) { trunc = 1; break; }                       // no EOF packet: read_u8 would fail (src/dec.rs:179)
        if (off + 5 > len) { trunc = 1; break; }
        pfv_packet pk;
        memset(&pk, 0, sizeof(pk));
        pk.type = data[off];
        pk.len = rd32(data + off + 1);
        pk.payload = off + 5;
        const bool has_payload = pk.type != 0;
        if (has_payload && pk.payload + pk.len > len) { trunc = 1; break; }
        out[n++] = pk;
        if (pk.type == 0) break;                                    // EOF marker (src/dec.rs:183-187)
        off = pk.payload + pk.len;
    }
    *n_out = n;
    if (truncated_out) *truncated_out = trunc;
    return PFV_OK;
}

// ---------------------------------------------------------------------------------------------------
// entropy decode of one payload -> sparse seam
// ---------------------------------------------------------------------------------------------------
extern "C" int pfv_packet_decode(const pfv_geometry *g, uint32_t kind, const uint8_t *payload, size_t len, uint8_t qidx_out[3],
                                 pfv_mbhdr *hdr_out, uint32_t *mb_off_out, uint32_t *tok_out, uint32_t tok_cap, uint32_t *ntok_out)
{
    if (!g || !payload || !qidx_out || !mb_off_out || !ntok_out || (!tok_out && tok_cap))
        return set_error(PFV_ERR_BAD_ARG, "NULL argument");
    if (kind != PFV_FRAME_I && kind != PFV_FRAME_P) return set_error(PFV_ERR_BAD_ARG, "bad kind %u", kind);
    if (kind == PFV_FRAME_P && !hdr_out) return set_error(PFV_ERR_BAD_ARG, "P frames need hdr_out");
    if (len < 19) return set_error(PFV_ERR_IO, "payload of %zu bytes ends inside the symbol table / q-table indices", len);
    Huffman h;
    h.build(payload);                                               // src/dec.rs:234-241
    qidx_out[0] = payload[16]; qidx_out[1] = payload[17]; qidx_out[2] = payload[18];   // src/dec.rs:244-246
    BitReader br(payload, len);
    br.pos = 19 * 8;
    const uint32_t nb = g->nb;
    uint32_t ntok = 0;

    h.build_pairs();
    // one (run, size[, value]) token out of ONE 57-bit window; returns false on a malformed stream
    auto token = [&](uint32_t &run, int &size, int32_t &value) -> bool {
        uint64_t w = br.peek();
        uint32_t used;
        const uint16_t pe = h.pair[w & ((1u << LUT_BITS) - 1)];
        if (pe) {
            run = pe & 15u; size = (pe >> 4) & 15; used = pe >> 8;
        } else if (h.only >= 0) {                                   // zero-length codes
            if (h.only == 0) return false;                           // (0, 0) forever: the reference would never return
            run = (uint32_t)h.only; size = h.only; used = 0;
        } else {
            if (h.nsym == 0) return false;                          // empty tree: reference hits unreachable!()
            int s1, s2;
            const int l1 = huff_symbol(h, w, s1);
            if (l1 < 0) return false;
            const int l2 = huff_symbol(h, w >> l1, s2);
            if (l2 < 0) return false;
            run = (uint32_t)s1; size = s2; used = (uint32_t)(l1 + l2);
        }
        value = size > 0 ? sign_extend((uint32_t)(w >> used) & ((1u << size) - 1u), size) : 0;   // read_signed::<i16>(size), src/dec.rs:286
        br.pos += used + (uint32_t)size;                            // <= 15 + 15 + 15 bits: inside the window
        return !br.overrun();
    };

    if (kind == PFV_FRAME_I) {
        // src/dec.rs:258-296: one continuous run over nb*256 coefficients
        const uint64_t total = (uint64_t)nb * 256;
        uint64_t out_idx = 0;
        uint32_t cur_mb = 0;
        mb_off_out[0] = 0;
        while (out_idx < total) {
            uint32_t run; int size; int32_t value;
            if (!token(run, size, value))
                return set_error(br.overrun() ? PFV_ERR_IO : PFV_ERR_BAD_STREAM, "I-frame payload: %s at coefficient %llu",
                                 br.overrun() ? "bit stream ends" : "undecodable symbol", (unsigned long long)out_idx);
            out_idx += run;
            if (size > 0) {
                if (out_idx >= total) return set_error(PFV_ERR_BAD_STREAM, "I-frame payload: coefficient index past the frame (src/dec.rs:288 would panic)");
                const uint32_t mb = (uint32_t)(out_idx >> 8);
                while (cur_mb < mb) mb_off_out[++cur_mb] = ntok;
                if (ntok >= tok_cap) return set_error(PFV_ERR_NOMEM, "token buffer of %u entries is too small", tok_cap);
                tok_out[ntok++] = ((uint32_t)(out_idx & 255) << 16) | (uint32_t)(uint16_t)value;
                out_idx++;
            }
        }
        while (cur_mb < nb) mb_off_out[++cur_mb] = ntok;
    } else {
        // src/dec.rs:359-372 headers, then :376-417 the coded macroblocks
        PlaneDims pd[3];
        plane_dims(*g, pd);
        uint32_t m = 0;
        for (int p = 0; p < 3; p++)
            for (uint32_t by = 0; by < pd[p].bh; by++)
                for (uint32_t bx = 0; bx < pd[p].bw; bx++, m++) {
                    const uint64_t w = br.peek();
                    pfv_mbhdr hd;
                    hd.mx = 0; hd.my = 0; hd.reserved = 0;
                    hd.has_coeff = (uint8_t)((w >> 1) & 1);
                    if (w & 1) {
                        hd.mx = (int8_t)sign_extend((uint32_t)(w >> 2) & 127u, 7);
                        hd.my = (int8_t)sign_extend((uint32_t)(w >> 9) & 127u, 7);
                        br.pos += 16;
                    } else {
                        br.pos += 2;
                    }
                    if (br.overrun()) return set_error(PFV_ERR_IO, "P-frame payload ends inside the macroblock headers");
                    const int sx = (int)bx * 16 + hd.mx, sy = (int)by * 16 + hd.my;
                    if (sx < 0 || sy < 0 || sx > (int)pd[p].pw - 16 || sy > (int)pd[p].ph - 16)
                        return set_error(PFV_ERR_BAD_MV, "macroblock %u: motion vector (%d,%d) leaves the padded plane (src/common.rs:258-259)",
                                         m, hd.mx, hd.my);
                    hdr_out[m] = hd;
                }
        mb_off_out[0] = 0;
        for (m = 0; m < nb; m++) {
            if (hdr_out[m].has_coeff) {
                uint32_t out_idx = 0;
                while (out_idx < 256) {
                    uint32_t run; int size; int32_t value;
                    if (!token(run, size, value))
                        return set_error(br.overrun() ? PFV_ERR_IO : PFV_ERR_BAD_STREAM, "P-frame payload: %s in macroblock %u",
                                         br.overrun() ? "bit stream ends" : "undecodable symbol", m);
                    out_idx += run;
                    if (size > 0) {
                        if (out_idx >= 256) return set_error(PFV_ERR_BAD_STREAM, "P-frame payload: coefficient index past the macroblock (src/dec.rs:410 would panic)");
                        if (ntok >= tok_cap) return set_error(PFV_ERR_NOMEM, "token buffer of %u entries is too small", tok_cap);
                        tok_out[ntok++] = (out_idx << 16) | (uint32_t)(uint16_t)value;
                        out_idx++;
                    }
                }
            }
            mb_off_out[m + 1] = ntok;
        }
    }
    *ntok_out = ntok;
    return PFV_OK;
}

// ---------------------------------------------------------------------------------------------------
// entropy encode of one frame from the dense seam
// ---------------------------------------------------------------------------------------------------
namespace {

// RLE tokens of one macroblock (src/rle.rs:9-39): packed (run) | (size << 4) | (uint16 value << 16)
inline uint32_t rle_macroblock(const int16_t *c, uint32_t *out, uint32_t hist[16])
{
    uint32_t n = 0, run = 0;
    for (int i = 0; i < 256; i += 4) {
        uint64_t four;
        memcpy(&four, c + i, 8);
        if (four == 0) { run += 4; continue; }
        for (int k = 0; k < 4; k++) {
            const int16_t v = c[i + k];
            if (v == 0) { run++; continue; }
            while (run > 15) { out[n++] = 15u; hist[15]++; hist[0]++; run -= 15; }
            const uint32_t a = (uint16_t)(v < 0 ? -v : v);           // i16::abs wraps for -32768 -> 32768 as u16
            const uint32_t size = (32u - (uint32_t)__builtin_clz(a)) + 1u;   // bit length of |v| plus the sign bit, src/rle.rs:23-24
            out[n++] = run | (size << 4) | ((uint32_t)(uint16_t)v << 16);
            hist[run]++; hist[size & 15u]++;
            run = 0;
        }
    }
    while (run > 15) { out[n++] = 15u; hist[15]++; hist[0]++; run -= 15; }
    if (run > 0) { out[n++] = run; hist[run]++; hist[0]++; }
    return n;
}

struct BitWriter {
    uint8_t *p;
    uint64_t acc = 0;
    int      n = 0;
    explicit BitWriter(uint8_t *dst) : p(dst) {}
    inline void put(uint32_t v, int bits)                           // bits <= 32, v < 2^bits
    {
        acc |= (uint64_t)v << n;
        n += bits;
        if (n >= 32) { const uint32_t w = (uint32_t)acc; memcpy(p, &w, 4); p += 4; acc >>= 32; n -= 32; }
    }
    inline uint8_t *finish()                                        // byte_align: zero padding
    {
        while (n > 0) { *p++ = (uint8_t)acc; acc >>= 8; n -= 8; }
        n = 0;
        return p;
    }
};

// The packet from one frame's RLE sequence and symbol statistics (the sparse encode seam, pfv_encode_submit_sparse):
// tree from the histograms, exact size first, then the bits.
int encode_packet_tokens(const pfv_geometry &g, uint32_t kind, const pfv_mbhdr *hdr, const uint32_t *tok, const uint32_t *stats,
                         std::vector<uint8_t> &out)
{
    const uint32_t nb = g.nb;
    const size_t ntok = stats[PFV_TOKSTATS_NTOK];
    if (stats[PFV_TOKSTATS_FLAGS] & PFV_TOKFLAG_OVERFLOW)
        return set_error(PFV_ERR_BAD_ARG, "the frame produced %zu RLE entries, more than the token buffer holds", ntok);
    if (stats[PFV_TOKSTATS_FLAGS] & PFV_TOKFLAG_RANGE)
        return set_error(PFV_ERR_BAD_ARG, "a coefficient needs more than 15 bits: not representable (src/rle.rs:24, :43)");
    // update_table counts both symbols of an entry into one table (src/rle.rs:41-47)
    uint32_t hist[16];
    for (int i = 0; i < 16; i++) hist[i] = stats[i] + stats[16 + i];
    // rle_create_huffman (src/rle.rs:49-66): weights normalised to u8, i32 arithmetic
    int32_t mx = 0;
    for (int i = 0; i < 16; i++) mx = std::max(mx, (int32_t)hist[i]);
    uint8_t table[16];
    for (int i = 0; i < 16; i++) {
        if (hist[i] > 0) {
            const int32_t prod = (int32_t)((uint32_t)hist[i] * 255u);   // wrapping like release-mode Rust
            int32_t v = prod / mx;
            if (v < 1) v = 1;
            table[i] = (uint8_t)v;
        } else table[i] = 0;
    }
    Huffman h;
    h.build(table);
    // exact payload size: every symbol costs its code, every entry its `size` value bits
    uint64_t bits = 19 * 8;
    if (kind == PFV_FRAME_P)
        for (uint32_t m = 0; m < nb; m++) bits += (hdr[m].mx != 0 || hdr[m].my != 0) ? 16 : 2;
    for (int i = 0; i < 16; i++) bits += (uint64_t)hist[i] * h.len[i] + (uint64_t)stats[16 + i] * (uint32_t)i;
    // one table lookup per entry: both codes concatenated, indexed by the entry's low byte (run | size << 4)
    uint32_t pair_val[256];
    uint8_t  pair_len[256], pair_bits[256];
    for (uint32_t b = 0; b < 256; b++) {
        const uint32_t run = b & 15u, size = b >> 4;
        pair_val[b] = h.val[run] | (h.val[size] << h.len[run]);
        pair_len[b] = (uint8_t)(h.len[run] + h.len[size]);
        pair_bits[b] = (uint8_t)(pair_len[b] + size);
    }
    {   // the statistics must describe this very sequence: the buffer below is sized from them
        uint64_t seq_bits = 0, sym_bits = 0;
        for (size_t i = 0; i < ntok; i++) seq_bits += pair_bits[tok[i] & 255u];
        for (int i = 0; i < 16; i++) sym_bits += (uint64_t)hist[i] * h.len[i] + (uint64_t)stats[16 + i] * (uint32_t)i;
        if (seq_bits != sym_bits)
            return set_error(PFV_ERR_BAD_ARG, "RLE entries and their statistics disagree (%llu bits against %llu)",
                             (unsigned long long)seq_bits, (unsigned long long)sym_bits);
    }
    const size_t payload = (size_t)((bits + 7) / 8);
    out.resize(5 + payload + 8);                                     // +8: the writer stores 4 bytes at a time
    out[0] = (uint8_t)(kind == PFV_FRAME_I ? 1 : 2);                 // src/enc.rs:324, :475
    wr32(&out[1], (uint32_t)payload);
    uint8_t *p = &out[5];
    memcpy(p, table, 16);                                            // src/enc.rs:290-292
    if (kind == PFV_FRAME_I) { p[16] = 0; p[17] = 1; p[18] = 1; }    // src/enc.rs:296-298
    else { p[16] = 2; p[17] = 3; p[18] = 3; }                        // src/enc.rs:409-411
    BitWriter bw(p + 19);
    if (kind == PFV_FRAME_P)
        for (uint32_t m = 0; m < nb; m++) {                         // src/enc.rs:414-452
            const bool has_mvec = hdr[m].mx != 0 || hdr[m].my != 0;
            bw.put((has_mvec ? 1u : 0u) | (hdr[m].has_coeff ? 2u : 0u), 2);
            if (has_mvec) bw.put(((uint32_t)hdr[m].mx & 127u) | (((uint32_t)hdr[m].my & 127u) << 7), 14);
        }
    for (size_t i = 0; i < ntok; i++) {                              // src/enc.rs:301-315, :455-466
        const uint32_t t = tok[i], b = t & 255u, size = b >> 4;
        // code pair (<= 30 bits) and value bits (<= 15) leave the 64-bit window room: one put each
        bw.put(pair_val[b], pair_len[b]);
        if (size) bw.put((t >> 16) & ((1u << size) - 1u), (int)size);
    }
    uint8_t *end = bw.finish();
    if ((size_t)(end - &out[5]) != payload) return set_error(PFV_ERR_STATE, "internal: packet size mismatch");
    out.resize(5 + payload);
    return PFV_OK;
}

// Run-length pass on the host: what the device tokenizer (pfv_kernels_tok.cu) emits, from dense coefficients.
int tokenize_frame(const pfv_geometry &g, uint32_t kind, const pfv_mbhdr *hdr, const int16_t *coeff, uint32_t *tok, size_t cap,
                   uint32_t *mb_off, uint32_t *stats)
{
    const uint32_t nb = g.nb;
    uint32_t hist[16] = {0}, sizes[16] = {0};
    size_t ntok = 0;
    uint32_t flags = 0;
    uint32_t local[256 + 32];
    for (uint32_t m = 0; m < nb; m++) {
        if (mb_off) mb_off[m] = (uint32_t)ntok;
        if (kind == PFV_FRAME_P && !hdr[m].has_coeff) continue;      // subblocks: None (src/enc.rs:357-358)
        const uint32_t n = rle_macroblock(coeff + (size_t)m * 256, local, hist);
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t size = (local[i] >> 4) & 31u;
            if (size > 15) flags |= PFV_TOKFLAG_RANGE;
            sizes[size & 15u]++;
            if (ntok + i < cap) tok[ntok + i] = local[i];
        }
        ntok += n;
    }
    if (mb_off) mb_off[nb] = (uint32_t)ntok;
    if (ntok > cap) flags |= PFV_TOKFLAG_OVERFLOW;
    memset(stats, 0, PFV_TOKSTATS_WORDS * sizeof(uint32_t));
    for (int i = 0; i < 16; i++) { stats[16 + i] = sizes[i]; stats[i] = hist[i] - sizes[i]; }   // rle_macroblock counts both symbols into hist
    stats[PFV_TOKSTATS_NTOK] = (uint32_t)ntok;
    stats[PFV_TOKSTATS_FLAGS] = flags;
    return PFV_OK;
}

// Encodes into `out` (resized); returns PFV_OK or an error.  scratch is reused across calls by the caller.
int encode_packet(const pfv_geometry &g, uint32_t kind, const pfv_mbhdr *hdr, const int16_t *coeff, std::vector<uint32_t> &scratch,
                  std::vector<uint8_t> &out)
{
    scratch.resize((size_t)g.nb * 256);
    uint32_t stats[PFV_TOKSTATS_WORDS];
    int rc = tokenize_frame(g, kind, hdr, coeff, scratch.data(), scratch.size(), nullptr, stats);
    if (rc) return rc;
    if (stats[PFV_TOKSTATS_FLAGS] & PFV_TOKFLAG_RANGE) {
        for (size_t i = 0; i < stats[PFV_TOKSTATS_NTOK]; i++)
            if (((scratch[i] >> 4) & 31u) > 15)
                return set_error(PFV_ERR_BAD_ARG, "coefficient %d needs %u bits: not representable (src/rle.rs:24, :43)",
                                 (int)(int16_t)(scratch[i] >> 16), (scratch[i] >> 4) & 31u);
    }
    return encode_packet_tokens(g, kind, hdr, scratch.data(), stats, out);
}

}  // namespace

extern "C" uint32_t pfv_packet_token_bound(const pfv_geometry *g, const uint8_t *payload, size_t len)
{
    if (!g || !payload) return 0;
    return (uint32_t)token_bound(payload, len, g->nb);
}

extern "C" size_t pfv_packet_encode_bound(const pfv_geometry *g)
{
    if (!g) return 0;
    return 5 + 19 + (size_t)g->nb * 2 + (size_t)g->nb * 256 * 6 + 16;
}

extern "C" int pfv_packet_encode(const pfv_geometry *g, uint32_t kind, const pfv_mbhdr *hdr, const int16_t *coeff, uint8_t *out,
                                 size_t cap, size_t *len_out)
{
    if (!g || !coeff || !out || !len_out) return set_error(PFV_ERR_BAD_ARG, "NULL argument");
    if (kind != PFV_FRAME_I && kind != PFV_FRAME_P) return set_error(PFV_ERR_BAD_ARG, "bad kind %u", kind);
    if (kind == PFV_FRAME_P && !hdr) return set_error(PFV_ERR_BAD_ARG, "P frames need headers");
    std::vector<uint32_t> scratch;
    std::vector<uint8_t> pkt;
    int rc = encode_packet(*g, kind, hdr, coeff, scratch, pkt);
    if (rc) return rc;
    if (pkt.size() > cap) return set_error(PFV_ERR_NOMEM, "packet of %zu bytes does not fit %zu", pkt.size(), cap);
    memcpy(out, pkt.data(), pkt.size());
    *len_out = pkt.size();
    return PFV_OK;
}

extern "C" int pfv_packet_encode_tokens(const pfv_geometry *g, uint32_t kind, const pfv_mbhdr *hdr, const uint32_t *tok,
                                        const uint32_t *stats, uint8_t *out, size_t cap, size_t *len_out)
{
    if (!g || !tok || !stats || !out || !len_out) return set_error(PFV_ERR_BAD_ARG, "NULL argument");
    if (kind != PFV_FRAME_I && kind != PFV_FRAME_P) return set_error(PFV_ERR_BAD_ARG, "bad kind %u", kind);
    if (kind == PFV_FRAME_P && !hdr) return set_error(PFV_ERR_BAD_ARG, "P frames need headers");
    std::vector<uint8_t> pkt;
    int rc = encode_packet_tokens(*g, kind, hdr, tok, stats, pkt);
    if (rc) return rc;
    if (pkt.size() > cap) return set_error(PFV_ERR_NOMEM, "packet of %zu bytes does not fit %zu", pkt.size(), cap);
    memcpy(out, pkt.data(), pkt.size());
    *len_out = pkt.size();
    return PFV_OK;
}

extern "C" int pfv_packet_tokenize(const pfv_geometry *g, uint32_t kind, const pfv_mbhdr *hdr, const int16_t *coeff,
                                   uint32_t *tok_out, uint32_t tok_cap, uint32_t *mb_off_out, uint32_t *stats_out)
{
    if (!g || !coeff || !tok_out || !stats_out) return set_error(PFV_ERR_BAD_ARG, "NULL argument");
    if (kind != PFV_FRAME_I && kind != PFV_FRAME_P) return set_error(PFV_ERR_BAD_ARG, "bad kind %u", kind);
    if (kind == PFV_FRAME_P && !hdr) return set_error(PFV_ERR_BAD_ARG, "P frames need headers");
    return tokenize_frame(*g, kind, hdr, coeff, tok_out, tok_cap, mb_off_out, stats_out);
}

// ---------------------------------------------------------------------------------------------------
// a small fixed pool of host threads
// ---------------------------------------------------------------------------------------------------
namespace {

struct Pinned {
    void  *p = nullptr;
    size_t bytes = 0;
    int reserve(size_t n)
    {
        if (n <= bytes) return PFV_OK;
        if (p) pfv_host_free(p);
        p = nullptr; bytes = 0;
        int rc = pfv_host_alloc(&p, n);
        if (rc) return rc;
        bytes = n;
        return PFV_OK;
    }
    ~Pinned() { if (p) pfv_host_free(p); }
};

}  // namespace

// ---------------------------------------------------------------------------------------------------
// Decoder
// ---------------------------------------------------------------------------------------------------
namespace {

enum { W_FREE = 0, W_QUEUED, W_READY, W_SUBMITTED };

struct DecWork {
    Pinned   meta, out;            // mb_off | headers | tokens (adjacent: one H2D copy per frame) ; decoded planes
    uint32_t ntok = 0, kind = 0;
    uint8_t  qidx[3] = {0, 0, 0};
    int      state = W_FREE;       // guarded by pfv_decoder::m
    int      status = PFV_OK;
    char     err[256] = "";
    uint64_t submit_id = 0;
    uint32_t slot = 0;             // the frame slot this work item owns (fixed: 1 + its index)
    uint32_t ref_slot = 0;         // P: slot of the previous frame of the stream (Decoder.framebuffer at that point)
    uint32_t chain = 0;            // frames between two key frames form a chain; different chains are independent
    uint32_t packet = 0;           // index into packets
    uint64_t epoch = 0;
};

}  // namespace

struct pfv_decoder {
    const uint8_t *data = nullptr;
    size_t len = 0;
    std::vector<uint8_t> owned;                       // pfv_decoder_open_reader: the stream as it came out of the caller's reader
    pfv_stream_info info{};
    pfv_geometry geo{};
    std::vector<pfv_packet> packets;
    bool truncated = false;
    pfv_ctx *ctx = nullptr;
    std::unique_ptr<Pool> pool;
    uint32_t depth = 0, nslots = 0;
    std::vector<std::unique_ptr<DecWork>> work;
    std::mutex m;
    std::condition_variable cv;
    // stream position
    uint32_t cursor = 0;           // next packet advance_frame looks at
    uint32_t sched = 0;            // next packet to consider for read-ahead
    std::deque<DecWork *> inflight;   // frames scheduled (entropy queued or later), in stream order
    DecWork *delivered = nullptr;  // the picture handed out by the previous call
    uint32_t fb_slot = 0;          // slot that holds Decoder.framebuffer for the next frame to be SCHEDULED
    uint32_t delivered_slot = 0;   // slot of the last picture handed out (slot 0 = the blank initial framebuffer)
    uint32_t chain = 0;            // chain number of the last scheduled frame
    uint32_t lanes = 4;            // jobs per submit: frames of up to this many chains (GOPs) go out together
    bool eof = false;
    double delta_accum = 0.0;
    uint64_t epoch = 0;            // bumped by reset(): results of older entropy jobs are dropped
    size_t ysz = 0, csz = 0;
    size_t off_u = 0, off_v = 0;   // where the U and V pictures sit in a work item's pinned picture buffer
    double t_prof[7] = {0, 0, 0, 0, 0, 0, 0};   // PFV_TRACE: seconds in schedule / wait entropy+submit / wait GPU / refill / entropy jobs / cv wait / submit call
    uint64_t n_prof = 0;
};

static inline double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void decoder_entropy_job(pfv_decoder *d, DecWork *w)
{
    const pfv_packet &pk = d->packets[w->packet];
    const uint32_t nb = d->geo.nb;
    uint32_t *mb_off = static_cast<uint32_t *>(w->meta.p);
    pfv_mbhdr *hdr = reinterpret_cast<pfv_mbhdr *>(mb_off + nb + 1);
    uint32_t *tok = mb_off + (nb + 1) + nb;
    uint32_t ntok = 0;
    const double t0 = now_s();
    int rc = pfv_packet_decode(&d->geo, w->kind, d->data + pk.payload, pk.len, w->qidx, hdr, mb_off, tok,
                               (uint32_t)(w->meta.bytes / 4 - (2 * (size_t)nb + 1)), &ntok);
    const double t1 = now_s();
    if (rc) snprintf(w->err, sizeof(w->err), "%s", pfv_last_error());
    {
        std::lock_guard<std::mutex> l(d->m);
        d->t_prof[4] += t1 - t0;
        w->status = rc;
        w->ntok = ntok;
        w->state = W_READY;
    }
    d->cv.notify_all();
}

// schedule entropy decoding of frames ahead of the cursor (caller thread only)
static int decoder_schedule(pfv_decoder *d)
{
    while (d->inflight.size() < d->depth && d->sched < d->packets.size()) {
        const pfv_packet &pk = d->packets[d->sched];
        if (pk.type == 0) break;
        const bool frame = (pk.type == 1 && pk.len > 0) || pk.type == 2;
        if (!frame) { d->sched++; continue; }
        DecWork *w = nullptr;
        for (auto &x : d->work) if (x->state == W_FREE && x.get() != d->delivered) { w = x.get(); break; }
        if (!w) break;
        // token capacity: an emitted token costs at least 3 bits when the tree has two or more symbols, and a
        // one-symbol tree emits none
        const uint64_t cap = token_bound(d->data + pk.payload, pk.len, d->geo.nb);
        int rc = w->meta.reserve((2 * (size_t)d->geo.nb + 1 + (size_t)cap) * 4);
        if (rc) return rc;
        rc = w->out.reserve(d->geo.frame_bytes);
        if (rc) return rc;
        w->kind = pk.type == 1 ? PFV_FRAME_I : PFV_FRAME_P;
        if (w->kind == PFV_FRAME_I) d->chain++;                      // a key frame depends on nothing: new chain
        w->chain = d->chain;
        w->ref_slot = d->fb_slot;                                   // the previous frame of the stream (src/dec.rs:425-432)
        d->fb_slot = w->slot;
        w->packet = d->sched;
        w->status = PFV_OK;
        w->err[0] = 0;
        w->epoch = d->epoch;
        { std::lock_guard<std::mutex> l(d->m); w->state = W_QUEUED; }
        d->inflight.push_back(w);
        d->pool->post([d, w] { decoder_entropy_job(d, w); });
        d->sched++;
    }
    return PFV_OK;
}

// Submit scheduled frames whose entropy decode is done.  Frames of one chain (key frame + the P frames behind it) go
// out in stream order, one per submit; frames of DIFFERENT chains are independent, so one submit carries the next
// frame of up to `lanes` chains - that divides the per-submit cost (a dozen CUDA calls) by the number of GOPs in
// flight.  `must` (the next picture to hand out) is waited for; everything else only goes if it is ready.
static int decoder_submit_ready(pfv_decoder *d, DecWork *must)
{
    const uint32_t nb = d->geo.nb;
    for (int round = 0; round < 64; round++) {
        std::vector<DecWork *> batch;
        std::vector<uint32_t> closed;                                // chains that cannot contribute (another) frame to this batch
        auto is_closed = [&](uint32_t c) { return std::find(closed.begin(), closed.end(), c) != closed.end(); };
        for (DecWork *w : d->inflight) {
            if (w->state == W_SUBMITTED) {
                if (w == must) must = nullptr;                      // went out with an earlier refill: nothing to wait for
                continue;
            }
            if (is_closed(w->chain)) continue;
            closed.push_back(w->chain);                             // at most one frame per chain and submit, in order
            {
                std::unique_lock<std::mutex> l(d->m);
                if (w->state != W_READY) {
                    if (w != must) continue;
                    const double tw = now_s();
                    d->cv.wait(l, [w] { return w->state == W_READY; });
                    d->t_prof[5] += now_s() - tw;
                }
            }
            if (w->status != PFV_OK) {
                if (w == must) return set_error(w->status, "%s", w->err);
                continue;                                           // reported when the cursor reaches it
            }
            for (int p = 0; p < 3; p++)
                if (w->qidx[p] >= d->info.num_qtables) {
                    w->status = PFV_ERR_BAD_STREAM;
                    snprintf(w->err, sizeof(w->err), "q-table index %u >= %u (src/dec.rs:244-246 would panic)", w->qidx[p], d->info.num_qtables);
                    break;
                }
            if (w->status != PFV_OK) {
                if (w == must) return set_error(w->status, "%s", w->err);
                continue;
            }
            batch.push_back(w);
            if (batch.size() == d->lanes) break;
        }
        if (batch.empty()) break;
        pfv_decode_job_sparse jobs[8];
        for (size_t i = 0; i < batch.size(); i++) {
            DecWork *w = batch[i];
            pfv_decode_job_sparse &j = jobs[i];
            memset(&j, 0, sizeof(j));
            j.kind = w->kind;
            j.ref_slot = w->ref_slot;
            j.dst_slot = w->slot;
            memcpy(j.qidx, w->qidx, 3);
            j.mb_off = static_cast<const uint32_t *>(w->meta.p);
            j.hdr = reinterpret_cast<const pfv_mbhdr *>(j.mb_off + nb + 1);
            j.tok = j.mb_off + (nb + 1) + nb;
            j.ntok = w->ntok;
            j.out_y = static_cast<uint8_t *>(w->out.p);
            j.out_u = j.out_y + d->off_u;
            j.out_v = j.out_y + d->off_v;
        }
        const double ts = now_s();
        int rc = pfv_decode_submit_sparse_trusted(d->ctx, jobs, (uint32_t)batch.size());
        d->t_prof[6] += now_s() - ts;
        if (rc) return rc;
        const uint64_t id = pfv_ctx_last_submit_id(d->ctx);
        for (DecWork *w : batch) {
            w->submit_id = id;
            std::lock_guard<std::mutex> l(d->m);
            w->state = W_SUBMITTED;
            if (w == must) must = nullptr;
        }
        if (!must) break;                                           // one batch per call unless the needed picture is still behind
    }
    return PFV_OK;
}

// Decoder<R: Read + Seek> (src/dec.rs:15-28) for a caller that has a reader instead of a byte range: the reader is drained
// once (the packet index and the read-ahead pipeline work on the whole stream) and the decoder owns the bytes.
extern "C" int pfv_decoder_open_reader(pfv_read_fn read, void *user, int device, uint32_t num_threads, uint32_t read_ahead, pfv_decoder **out)
{
    if (!out) return set_error(PFV_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    if (!read) return set_error(PFV_ERR_BAD_ARG, "read is NULL");
    std::vector<uint8_t> buf;
    size_t have = 0;
    for (;;) {
        if (buf.size() < have + ((size_t)1 << 20)) buf.resize(std::max(buf.size() * 2, have + ((size_t)1 << 20)));
        const long long n = read(user, buf.data() + have, buf.size() - have);
        if (n < 0) return set_error(PFV_ERR_IO, "the reader callback failed (io::Error of R::read)");
        if (n == 0) break;
        have += (size_t)n;
    }
    buf.resize(have);
    int rc = pfv_decoder_open(buf.data(), buf.size(), device, num_threads, read_ahead, out);
    if (rc) return rc;
    // std::vector's move keeps the heap block: the decoder's data pointer stays valid
    (*out)->owned = std::move(buf);
    (*out)->data = (*out)->owned.data();
    return PFV_OK;
}

extern "C" int pfv_decoder_open(const uint8_t *data, size_t len, int device, uint32_t num_threads, uint32_t read_ahead, pfv_decoder **out)
{
    if (!out) return set_error(PFV_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    if (!data) return set_error(PFV_ERR_BAD_ARG, "data is NULL");
    std::unique_ptr<pfv_decoder> d(new (std::nothrow) pfv_decoder());
    if (!d) return set_error(PFV_ERR_NOMEM, "out of host memory");
    int rc = pfv_stream_parse_header(data, len, &d->info, nullptr, 0);
    if (rc) return rc;
    if (d->info.width == 0 || d->info.height == 0 || (d->info.width & 1) || (d->info.height & 1))
        return set_error(PFV_ERR_BAD_STREAM, "frame size %ux%u must be even and non-zero (src/frame.rs:13)", d->info.width, d->info.height);
    if (d->info.num_qtables == 0) return set_error(PFV_ERR_BAD_STREAM, "stream carries no q-tables");
    std::vector<int32_t> qt((size_t)d->info.num_qtables * 64);
    rc = pfv_stream_parse_header(data, len, &d->info, reinterpret_cast<int32_t(*)[64]>(qt.data()), d->info.num_qtables);
    if (rc) return rc;
    d->data = data;
    d->len = len;
    pfv_geometry_for(d->info.width, d->info.height, &d->geo);
    d->ysz = (size_t)d->geo.width * d->geo.height;
    d->csz = (size_t)d->geo.cwidth * d->geo.cheight;
    // When no row is padded the pictures are kept in the slot's own layout (Y | U | V with their padding ROWS in
    // between): the engine then moves all three with one copy.  Otherwise tight planes back to back.
    if (d->geo.pw == d->geo.width && d->geo.cpw == d->geo.cwidth) {
        d->off_u = (size_t)d->geo.pw * d->geo.ph;
        d->off_v = d->off_u + (size_t)d->geo.cpw * d->geo.cph;
    } else {
        d->off_u = d->ysz;
        d->off_v = d->ysz + d->csz;
    }
    // packet index: O(1) per packet, no entropy decoding (src/dec.rs:179-180)
    {
        uint32_t cap = 1024, n = 0;
        int trunc = 0;
        for (;;) {
            d->packets.resize(cap);
            rc = pfv_stream_index(data, len, d->info.first_packet, d->packets.data(), cap, &n, &trunc);
            if (rc) return rc;
            if (n < cap) break;
            cap *= 4;
        }
        d->packets.resize(n);
        d->truncated = trunc != 0;
    }
    // default: enough frames in flight to keep every entropy thread busy plus a few on the GPU
    d->depth = read_ahead ? read_ahead : std::max<uint32_t>(6, 2 * (num_threads ? num_threads : 1) + 4);
    if (d->depth > 48) d->depth = 48;                                // pfv_ctx_wait_submit reaches 64 submits back
    d->nslots = d->depth + 2;                                        // slot 0: the blank initial framebuffer; then one per work item
    rc = pfv_ctx_create(device, d->info.width, d->info.height, reinterpret_cast<const int32_t(*)[64]>(qt.data()),
                        // q-table indices are u8 in the packets (src/dec.rs:244-246): tables beyond 256 can never be addressed;
                        // the reference's Decoder::new accepts such a header, so does this one
                        d->info.num_qtables > 256 ? 256u : d->info.num_qtables, d->nslots, d->lanes, nullptr, &d->ctx);
    if (rc) return rc;
    if ((rc = pfv_ctx_reserve_staging(d->ctx, true, false)) != PFV_OK) { pfv_ctx_destroy(d->ctx); return rc; }   // (not inside the first advance_frame)
    // pinned buffers are sized once, from the largest frame packet of the stream (cudaHostAlloc costs milliseconds:
    // never on the per-frame path).  Token capacity: an emitted token costs at least 3 bits when the tree has two or
    // more symbols, and a one-symbol tree emits none.
    uint64_t tok_cap = 1;
    for (const pfv_packet &pk : d->packets)
        if ((pk.type == 1 || pk.type == 2) && pk.len > 0)
            tok_cap = std::max(tok_cap, token_bound(data + pk.payload, pk.len, d->geo.nb));
    for (uint32_t i = 0; i < d->depth + 1; i++) {
        std::unique_ptr<DecWork> w(new DecWork());
        if ((rc = w->meta.reserve((2 * (size_t)d->geo.nb + 1 + (size_t)tok_cap) * 4)) ||
            (rc = w->out.reserve(d->geo.frame_bytes))) {
            pfv_ctx_destroy(d->ctx);
            return rc;
        }
        w->slot = i + 1;
        d->work.push_back(std::move(w));
    }
    d->pool.reset(new Pool(num_threads ? num_threads : 1));
    *out = d.release();
    return PFV_OK;
}

static void decoder_drain(pfv_decoder *d)
{
    // wait for entropy jobs still running (they hold pointers into the work items)
    for (DecWork *w : d->inflight) {
        std::unique_lock<std::mutex> l(d->m);
        d->cv.wait(l, [w] { return w->state != W_QUEUED; });
    }
    if (d->ctx) pfv_sync(d->ctx);
    for (DecWork *w : d->inflight) { std::lock_guard<std::mutex> l(d->m); w->state = W_FREE; }
    d->inflight.clear();
}

extern "C" void pfv_decoder_close(pfv_decoder *d)
{
    if (!d) return;
    decoder_drain(d);
    if (getenv("PFV_TRACE") && d->n_prof)
        fprintf(stderr, "[pfv_decoder] %llu frames; per frame us: schedule %.0f, wait entropy + submit %.0f, wait GPU %.0f, refill %.0f; entropy job %.0f; of which cv wait %.0f, submit call %.0f\n",
                (unsigned long long)d->n_prof, 1e6 * d->t_prof[0] / d->n_prof, 1e6 * d->t_prof[1] / d->n_prof,
                1e6 * d->t_prof[2] / d->n_prof, 1e6 * d->t_prof[3] / d->n_prof, 1e6 * d->t_prof[4] / d->n_prof,
                1e6 * d->t_prof[5] / d->n_prof, 1e6 * d->t_prof[6] / d->n_prof);
    d->pool.reset();
    if (d->ctx) pfv_ctx_destroy(d->ctx);
    delete d;
}

extern "C" uint32_t pfv_decoder_width(const pfv_decoder *d) { return d ? d->info.width : 0; }
extern "C" uint32_t pfv_decoder_height(const pfv_decoder *d) { return d ? d->info.height : 0; }
extern "C" uint32_t pfv_decoder_framerate(const pfv_decoder *d) { return d ? d->info.framerate : 0; }
extern "C" pfv_ctx *pfv_decoder_ctx(pfv_decoder *d) { return d ? d->ctx : nullptr; }
extern "C" uint32_t pfv_decoder_framebuffer_slot(const pfv_decoder *d) { return d ? d->delivered_slot : 0; }

extern "C" int pfv_decoder_reset(pfv_decoder *d)
{
    if (!d) return set_error(PFV_ERR_BAD_ARG, "NULL decoder");
    // frames decoded ahead of the cursor are dropped; Decoder.framebuffer stays what the last returned picture
    // left there (src/dec.rs:148-152 does not clear it)
    decoder_drain(d);
    d->epoch++;
    d->fb_slot = d->delivered_slot;
    d->cursor = d->sched = 0;
    d->eof = false;
    return PFV_OK;
}

extern "C" int pfv_decoder_advance_frame(pfv_decoder *d, int *got_frame, const uint8_t **y, const uint8_t **u, const uint8_t **v)
{
    if (!d) return set_error(PFV_ERR_BAD_ARG, "NULL decoder");
    if (got_frame) *got_frame = 0;
    if (d->eof) return 0;                                            // src/dec.rs:171-173
    for (;;) {
        if (d->cursor >= d->packets.size())
            return set_error(PFV_ERR_IO, "stream ends without an EOF packet%s (src/dec.rs:179-180 read fails)",
                             d->truncated ? " (last packet is cut short)" : "");
        const pfv_packet &pk = d->packets[d->cursor];
        if (pk.type == 0) { d->eof = true; return 0; }               // src/dec.rs:183-187
        if (pk.type == 1 && pk.len == 0) { d->cursor++; if (d->sched < d->cursor) d->sched = d->cursor; return 1; }   // drop frame, src/dec.rs:190-201
        if (pk.type != 1 && pk.type != 2) { d->cursor++; if (d->sched < d->cursor) d->sched = d->cursor; continue; }  // src/dec.rs:216-219
        break;
    }
    const double t0 = now_s();
    int rc = decoder_schedule(d);
    if (rc) return rc;
    const double t1 = now_s();
    if (d->inflight.empty() || d->inflight.front()->packet != d->cursor)
        return set_error(PFV_ERR_STATE, "internal: read-ahead queue out of step with the cursor");
    DecWork *w = d->inflight.front();
    rc = decoder_submit_ready(d, w);
    if (rc) {
        // the failing frame is consumed, like a reference decode that returned Err after reading the packet; the
        // framebuffer stays what the last returned picture left there
        decoder_drain(d);
        d->fb_slot = d->delivered_slot;
        d->cursor++;
        d->sched = d->cursor;
        return rc;
    }
    if (d->delivered) {
        // The previous picture's buffers may be reused now - not earlier: its slot is the reference of the frame that
        // has just been submitted.
        std::lock_guard<std::mutex> l(d->m);
        d->delivered->state = W_FREE;
        d->delivered = nullptr;
    }
    const double t2 = now_s();
    rc = pfv_ctx_wait_submit_polling(d->ctx, w->submit_id, 300e-6);
    if (rc) {
        // same recovery as a failed submit: the frame is consumed, everything decoded ahead is dropped and the framebuffer
        // stays what the last returned picture left there (the reference reports an error only for the frame that fails)
        decoder_drain(d);
        d->fb_slot = d->delivered_slot;
        d->cursor++;
        d->sched = d->cursor;
        return rc;
    }
    const double t3 = now_s();
    d->inflight.pop_front();
    d->delivered = w;
    d->delivered_slot = w->slot;
    d->cursor++;
    // keep the pipeline full for the next call.  A failure here belongs to a LATER frame: this call's picture is decoded
    // and is handed out; the frame that failed reports its error when its turn comes (its submit is retried then)
    if (decoder_schedule(d) == PFV_OK) (void)decoder_submit_ready(d, nullptr);
    const double t4 = now_s();
    d->t_prof[0] += t1 - t0; d->t_prof[1] += t2 - t1; d->t_prof[2] += t3 - t2; d->t_prof[3] += t4 - t3; d->n_prof++;
    if (got_frame) *got_frame = 1;
    const uint8_t *base = static_cast<const uint8_t *>(w->out.p);
    if (y) *y = base;
    if (u) *u = base + d->off_u;
    if (v) *v = base + d->off_v;
    return 1;
}

extern "C" int pfv_decoder_advance_delta(pfv_decoder *d, double delta, pfv_onvideo_fn onvideo, void *user)
{
    if (!d) return set_error(PFV_ERR_BAD_ARG, "NULL decoder");
    d->delta_accum += delta;                                         // src/dec.rs:156
    const double per_frame = 1.0 / (double)d->info.framerate;
    while (d->delta_accum >= per_frame) {
        int got = 0;
        const uint8_t *y, *u, *v;
        const int rc = pfv_decoder_advance_frame(d, &got, &y, &u, &v);
        if (rc < 0) return rc;
        if (got && onvideo) onvideo(user, y, u, v);
        if (rc == 0) return 0;                                       // src/dec.rs:160-162
        d->delta_accum -= per_frame;
    }
    return 1;
}

// ---------------------------------------------------------------------------------------------------
// Encoder
// ---------------------------------------------------------------------------------------------------
namespace {

struct EncWork {
    Pinned   src, coeff, hdr;      // coeff: the RLE sequence + statistics (sparse seam, default) or nb*256 dense coefficients
    std::vector<uint8_t> packet;   // the finished packet; the buffer is reused frame after frame (a fresh 100+ KB vector per frame
                                   // means an mmap/munmap pair per frame, and every munmap interrupts all threads of the process)
    uint32_t kind = 0;
    uint64_t submit_id = 0;
    bool     busy = false;         // guarded by pfv_encoder::m; released once the packet has been appended to the stream
};

struct OutPacket {
    EncWork *work = nullptr;       // frame packets: bytes are in work->packet
    std::vector<uint8_t> bytes;    // literal packets (drop frame, eof)
    bool ready = false;
    int  status = PFV_OK;
    char err[256] = "";
};

}  // namespace

// The Encoder's own output buffer (its `W` when no writer callback is set): bytes are only ever appended.  As a std::vector it cost the
// writer thread - the one serial stage behind the entropy coders - 170-230 us per 1080p frame (PFV_TRACE, 150 KB packets): every
// doubling moved the stream into fresh memory, so each byte of a 70 MB stream was page-faulted in twice, 4 KB at a time, and copied
// once.  Here the buffer is an anonymous mapping that grows in place (mremap moves page tables, not bytes): every page is faulted in once.
struct GrowBuf {
    uint8_t *p = nullptr;
    size_t   n = 0, cap = 0;
    GrowBuf() = default;
    GrowBuf(const GrowBuf &) = delete;
    GrowBuf &operator=(const GrowBuf &) = delete;
    ~GrowBuf() { if (p) munmap(p, cap); }
    bool reserve(size_t want)
    {
        if (want <= cap) return true;
        size_t nc = cap ? cap : ((size_t)8 << 20);
        while (nc < want) nc *= 2;
        void *q = p ? mremap(p, cap, nc, MREMAP_MAYMOVE) : mmap(nullptr, nc, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (q == MAP_FAILED) return false;
#ifdef MADV_HUGEPAGE
        // Huge pages make the faults 512 times fewer, but where the host's transparent_hugepage/defrag is "madvise" (the default
        // of most distributions) an advised region compacts memory ON the fault and can stall the writer for milliseconds: opt-in.
        static const bool huge = getenv("PFV_STREAM_HUGEPAGES") && atoi(getenv("PFV_STREAM_HUGEPAGES")) != 0;
        if (huge) madvise(q, nc, MADV_HUGEPAGE);
#endif
        p = static_cast<uint8_t *>(q);
        cap = nc;
        return true;
    }
    bool append(const void *src, size_t k)
    {
        if (!k) return true;
        if (!reserve(n + k)) return false;
        memcpy(p + n, src, k);
        n += k;
        return true;
    }
    const uint8_t *data() const { return p; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    void clear() { n = 0; }
};

struct pfv_encoder {
    uint32_t width = 0, height = 0, framerate = 0;
    float px_err = 0.f;
    pfv_geometry geo{};
    pfv_ctx *ctx = nullptr;
    std::unique_ptr<Pool> pool;
    std::unique_ptr<Pool> copy_pool;                  // helpers for the copy of the caller's planes into pinned memory
    unsigned copy_helpers = 0;
    std::vector<std::unique_ptr<EncWork>> work;
    std::mutex m;
    std::condition_variable cv;
    std::deque<std::shared_ptr<OutPacket>> pending;   // packets in stream order, not yet appended to `stream`
    // The calling thread only copies the planes into a pinned work item (the reference borrows them for the call) and queues it.
    // Two internal threads do the rest: the SUBMITTER hands queued frames to the GPU in order (~75 us of driver calls per frame)
    // and posts their entropy coding to the pool; the WRITER appends finished packets in stream order (or feeds the writer
    // callback) and releases their work items.  Errors that happen there are sticky and come back from the next call.
    std::deque<std::pair<EncWork *, std::shared_ptr<OutPacket>>> submit_q;
    std::thread submitter, writer_thread;
    bool stop = false;
    int  async_rc = PFV_OK;
    char async_err[256] = "";
    GrowBuf stream;                                   // the writer W (when no callback is set)
    pfv_write_fn writer = nullptr;                    // pfv_encoder_set_writer: packets leave through it as soon as they are finished
    void *writer_user = nullptr;
    uint32_t prev_slot = 0;
    bool finished = false;
    bool dense = false;            // PFV_ENCODER_DENSE=1: the dense seam (pfv_encode_submit + host run-length pass)
    bool trace = false;            // PFV_TRACE=1: where the calling thread spends its time, printed at close
    double t_flush = 0, t_wait = 0, t_copy = 0, t_submit = 0, t_gpuwait = 0, t_entropy = 0, t_write = 0;   // (the last three under e->m)
    uint64_t n_frames = 0;
    size_t ysz = 0, csz = 0;
};

// The caller's planes may be reused as soon as encode_*frame returns (the reference borrows them for the call only), so they are
// copied into pinned memory first.  One thread moves ~8-10 GB/s, which at 1080p (3.1 MB) is the slowest step of the whole
// call; large frames are therefore copied in 256 KB pieces by the calling thread and a few helpers.
static void encoder_copy_planes(pfv_encoder *e, uint8_t *dst, const uint8_t *y, const uint8_t *u, const uint8_t *v)
{
    struct Seg { uint8_t *d; const uint8_t *s; size_t n; };
    const Seg planes[3] = {{dst, y, e->ysz}, {dst + e->ysz, u, e->csz}, {dst + e->ysz + e->csz, v, e->csz}};
    const size_t total = e->ysz + 2 * e->csz;
    if (!e->copy_pool || total < ((size_t)1 << 20)) {
        for (const Seg &p : planes) memcpy(p.d, p.s, p.n);
        return;
    }
    constexpr size_t PIECE = (size_t)256 << 10;
    std::vector<Seg> segs;
    for (const Seg &p : planes)
        for (size_t o = 0; o < p.n; o += PIECE) segs.push_back({p.d + o, p.s + o, std::min(PIECE, p.n - o)});
    std::atomic<size_t> next{0};
    std::atomic<unsigned> done{0};
    auto work = [&] {
        for (;;) {
            const size_t i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= segs.size()) break;
            memcpy(segs[i].d, segs[i].s, segs[i].n);
        }
    };
    const unsigned helpers = e->copy_helpers;
    for (unsigned i = 0; i < helpers; i++)
        e->copy_pool->post([&] { work(); done.fetch_add(1, std::memory_order_release); });
    work();
    while (done.load(std::memory_order_acquire) != helpers) std::this_thread::yield();
}

// the sticky error of the internal threads, if any
static int encoder_async_error(pfv_encoder *e)
{
    std::lock_guard<std::mutex> l(e->m);
    return e->async_rc == PFV_OK ? PFV_OK : set_error(e->async_rc, "%s", e->async_err);
}

// block until everything queued so far has been appended to `stream` (or handed to the writer callback)
static int encoder_flush(pfv_encoder *e)
{
    {
        std::unique_lock<std::mutex> l(e->m);
        e->cv.wait(l, [e] { return e->pending.empty() && e->submit_q.empty(); });
    }
    return encoder_async_error(e);
}

static void encoder_set_async_error(pfv_encoder *e, int rc, const char *msg)     // e->m held
{
    if (e->async_rc == PFV_OK) {
        e->async_rc = rc;
        snprintf(e->async_err, sizeof(e->async_err), "%s", msg);
    }
}

// WRITER thread: packets leave in stream order as soon as they are finished
static void encoder_writer_main(pfv_encoder *e)
{
    for (;;) {
        std::shared_ptr<OutPacket> p;
        {
            std::unique_lock<std::mutex> l(e->m);
            e->cv.wait(l, [e] { return e->stop || (!e->pending.empty() && e->pending.front()->ready); });
            if (e->pending.empty() || !e->pending.front()->ready) return;          // stop
            p = e->pending.front();                                                // stays at the front until it is written
        }
        int rc = p->status;
        const char *msg = p->err;
        const double tw0 = e->trace ? now_s() : 0;
        if (rc == PFV_OK) {
            const std::vector<uint8_t> &b = p->work ? p->work->packet : p->bytes;
            if (e->writer) {
                if (!b.empty() && e->writer(e->writer_user, b.data(), b.size()) != 0) {
                    rc = PFV_ERR_IO;
                    msg = "the writer callback failed (io::Error of W::write_all)";
                }
            } else {
                if (!e->stream.append(b.data(), b.size())) {
                    rc = PFV_ERR_NOMEM;
                    msg = "out of host memory for the encoded stream";
                }
            }
        }
        {
            std::lock_guard<std::mutex> l(e->m);
            if (rc != PFV_OK) encoder_set_async_error(e, rc, msg);
            if (p->work) p->work->busy = false;
            e->pending.pop_front();
            if (e->trace) e->t_write += now_s() - tw0;
        }
        e->cv.notify_all();
    }
}

static void encoder_submit_one(pfv_encoder *e, EncWork *w, const std::shared_ptr<OutPacket> &pkt);

// SUBMITTER thread: the only thread that drives the engine context
static void encoder_submitter_main(pfv_encoder *e)
{
    for (;;) {
        std::pair<EncWork *, std::shared_ptr<OutPacket>> item;
        {
            std::unique_lock<std::mutex> l(e->m);
            e->cv.wait(l, [e] { return e->stop || !e->submit_q.empty(); });
            if (e->submit_q.empty()) return;                                       // stop
            item = e->submit_q.front();
        }
        encoder_submit_one(e, item.first, item.second);
        {
            std::lock_guard<std::mutex> l(e->m);
            e->submit_q.pop_front();
        }
        e->cv.notify_all();
    }
}

extern "C" int pfv_encoder_open(uint32_t width, uint32_t height, uint32_t framerate, int quality, uint32_t num_threads, int device,
                                pfv_encoder **out)
{
    if (!out) return set_error(PFV_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    int32_t qt[4][64];
    float px_err = 0.f;
    int rc = pfv_make_qtables(quality, qt, &px_err);                 // assert!(quality >= 0 && quality <= 10), src/enc.rs:38
    if (rc) return rc;
    if (framerate > 65535) return set_error(PFV_ERR_BAD_ARG, "framerate %u does not fit the u16 header field (src/enc.rs:197)", framerate);
    std::unique_ptr<pfv_encoder> e(new (std::nothrow) pfv_encoder());
    if (!e) return set_error(PFV_ERR_NOMEM, "out of host memory");
    e->width = width; e->height = height; e->framerate = framerate; e->px_err = px_err;
    rc = pfv_ctx_create(device, width, height, qt, 4, 2, 1, nullptr, &e->ctx);
    if (rc) return rc;
    if ((rc = pfv_ctx_reserve_staging(e->ctx, false, true)) != PFV_OK) { pfv_ctx_destroy(e->ctx); return rc; }
    pfv_ctx_geometry(e->ctx, &e->geo);
    e->ysz = (size_t)e->geo.width * e->geo.height;
    e->csz = (size_t)e->geo.cwidth * e->geo.cheight;
    // frames in flight: one per entropy thread, two on the GPU, two finished and waiting to be appended (a work item is held
    // until then); pfv_ctx_wait_submit reaches 64 submits back
    const uint32_t depth = std::min<uint32_t>(std::max<uint32_t>(num_threads, 2) + 4, 32);
    {
        const char *env = getenv("PFV_ENCODER_DENSE");
        e->dense = env && atoi(env) != 0;
        env = getenv("PFV_TRACE");
        e->trace = env && atoi(env) != 0;
    }
    // sparse seam: nb*256 entries always suffice (an entry consumes at least one coefficient), + the statistics block
    const size_t out_bytes = e->dense ? (size_t)e->geo.nb * 512 : ((size_t)e->geo.nb * 256 + PFV_TOKSTATS_WORDS) * sizeof(uint32_t);
    for (uint32_t i = 0; i < depth; i++) {
        std::unique_ptr<EncWork> w(new EncWork());
        if ((rc = w->src.reserve(e->ysz + 2 * e->csz)) || (rc = w->coeff.reserve(out_bytes)) ||
            (rc = w->hdr.reserve((size_t)e->geo.nb * sizeof(pfv_mbhdr)))) {
            pfv_ctx_destroy(e->ctx);
            return rc;
        }
        e->work.push_back(std::move(w));
    }
    e->pool.reset(new Pool(num_threads ? std::min<uint32_t>(num_threads, depth) : 1));
    e->copy_helpers = num_threads >= 2 ? std::min<uint32_t>(num_threads - 1, 5) : 0;   // (measured on a 16-thread host, frames/s: 2: 5.4-6.0 k, 3: 4.8-5.5 k, 5: 6.1-7.4 k, 7: 6.0-6.8 k)
    if (const char *env = getenv("PFV_ENCODER_COPY_HELPERS")) {      // tuning aid
        const int v = atoi(env);
        if (v >= 0 && v <= 15) e->copy_helpers = (unsigned)v;
    }
    if (e->copy_helpers) e->copy_pool.reset(new Pool(e->copy_helpers));
    // write_header, src/enc.rs:190-219
    std::vector<uint8_t> s;
    s.insert(s.end(), kMagic, kMagic + 8);
    s.resize(12);
    wr32(&s[8], kVersion);
    wr16(s, width); wr16(s, height); wr16(s, framerate);
    wr16(s, 4);
    for (int t = 0; t < 4; t++)                                      // intra_l, intra_c, inter_l, inter_c
        for (int i = 0; i < 64; i++) wr16(s, (uint32_t)qt[t][i]);
    if (!e->stream.append(s.data(), s.size())) {
        pfv_ctx_destroy(e->ctx);
        return set_error(PFV_ERR_NOMEM, "out of host memory");
    }
    e->submitter = std::thread(encoder_submitter_main, e.get());
    e->writer_thread = std::thread(encoder_writer_main, e.get());
    *out = e.release();
    return PFV_OK;
}

static int encoder_frame(pfv_encoder *e, uint32_t kind, const uint8_t *y, const uint8_t *u, const uint8_t *v)
{
    if (!e) return set_error(PFV_ERR_BAD_ARG, "NULL encoder");
    if (e->finished) return set_error(PFV_ERR_STATE, "encoder is finished (assert!(!self.finished), src/enc.rs:80,130)");
    if (!y || !u || !v) return set_error(PFV_ERR_BAD_ARG, "NULL plane");
    int rc = encoder_async_error(e);
    if (rc) return rc;
    const double t0 = e->trace ? now_s() : 0;
    EncWork *w = nullptr;
    {
        // a work item is free again once its packet has been written (the writer thread releases it)
        std::unique_lock<std::mutex> l(e->m);
        e->cv.wait(l, [&] {
            for (auto &x : e->work) if (!x->busy) { w = x.get(); return true; }
            return false;
        });
        w->busy = true;
    }
    const double t1 = e->trace ? now_s() : 0;
    encoder_copy_planes(e, static_cast<uint8_t *>(w->src.p), y, u, v);
    w->kind = kind;
    std::shared_ptr<OutPacket> pkt(new OutPacket());
    pkt->work = w;
    {
        std::lock_guard<std::mutex> l(e->m);
        e->pending.push_back(pkt);                                   // its place in the stream is fixed here
        e->submit_q.emplace_back(w, pkt);
    }
    e->cv.notify_all();
    if (e->trace) {
        e->t_wait += t1 - t0; e->t_copy += now_s() - t1;
        e->n_frames++;
    }
    return PFV_OK;
}

// (submitter thread) one frame: submit to the engine, then post its entropy coding
static void encoder_submit_one(pfv_encoder *e, EncWork *w, const std::shared_ptr<OutPacket> &pkt)
{
    const double t0 = e->trace ? now_s() : 0;
    uint8_t *src = static_cast<uint8_t *>(w->src.p);
    const uint32_t kind = w->kind;
    const uint32_t dst_slot = e->prev_slot ^ 1u;
    uint32_t *tok = static_cast<uint32_t *>(w->coeff.p), *stats = tok + (size_t)e->geo.nb * 256;
    int rc;
    if (e->dense) {
        pfv_encode_job j;
        memset(&j, 0, sizeof(j));
        j.kind = kind;
        j.ref_slot = e->prev_slot;
        j.dst_slot = dst_slot;
        j.px_err = e->px_err;
        j.src_y = src; j.src_u = src + e->ysz; j.src_v = src + e->ysz + e->csz;
        j.hdr_out = static_cast<pfv_mbhdr *>(w->hdr.p);
        j.coeff_out = static_cast<int16_t *>(w->coeff.p);
        rc = pfv_encode_submit(e->ctx, &j, 1);
    } else {
        pfv_encode_job_sparse j;
        memset(&j, 0, sizeof(j));
        j.kind = kind;
        j.ref_slot = e->prev_slot;
        j.dst_slot = dst_slot;
        j.px_err = e->px_err;
        j.src_y = src; j.src_u = src + e->ysz; j.src_v = src + e->ysz + e->csz;
        j.hdr_out = static_cast<pfv_mbhdr *>(w->hdr.p);
        j.tok_out = tok;
        j.tok_cap = e->geo.nb * 256u;
        j.stats_out = stats;
        rc = pfv_encode_submit_sparse(e->ctx, &j, 1);
    }
    if (rc) {
        std::lock_guard<std::mutex> l(e->m);
        snprintf(pkt->err, sizeof(pkt->err), "%s", pfv_last_error());
        pkt->status = rc;
        pkt->ready = true;                                           // the writer turns it into the sticky error and frees the item
        return;
    }
    {
        std::lock_guard<std::mutex> l(e->m);
        e->prev_slot = dst_slot;                                     // src/enc.rs:95-97, :145-147
    }
    w->submit_id = pfv_ctx_last_submit_id(e->ctx);
    if (e->trace) e->t_submit += now_s() - t0;
    e->pool->post([e, w, pkt, tok, stats] {
        const double ta = e->trace ? now_s() : 0;
        int rc2 = pfv_ctx_wait_submit(e->ctx, w->submit_id);
        const double tb = e->trace ? now_s() : 0;
        if (rc2 == PFV_OK) {
            static thread_local std::vector<uint32_t> scratch;
            const pfv_mbhdr *hdr = static_cast<const pfv_mbhdr *>(w->hdr.p);
            rc2 = e->dense ? encode_packet(e->geo, w->kind, hdr, static_cast<const int16_t *>(w->coeff.p), scratch, w->packet)
                           : encode_packet_tokens(e->geo, w->kind, hdr, tok, stats, w->packet);
        }
        if (rc2) snprintf(pkt->err, sizeof(pkt->err), "%s", pfv_last_error());
        const double tc = e->trace ? now_s() : 0;
        {
            std::lock_guard<std::mutex> l(e->m);
            pkt->status = rc2;
            pkt->ready = true;
            e->t_gpuwait += tb - ta; e->t_entropy += tc - tb;
        }
        e->cv.notify_all();
    });
}

extern "C" int pfv_encoder_encode_iframe(pfv_encoder *e, const uint8_t *y, const uint8_t *u, const uint8_t *v)
{
    return encoder_frame(e, PFV_FRAME_I, y, u, v);
}

extern "C" int pfv_encoder_encode_pframe(pfv_encoder *e, const uint8_t *y, const uint8_t *u, const uint8_t *v)
{
    return encoder_frame(e, PFV_FRAME_P, y, u, v);
}

static int encoder_literal_packet(pfv_encoder *e, uint8_t type)
{
    std::shared_ptr<OutPacket> pkt(new OutPacket());
    pkt->bytes.assign(5, 0);
    pkt->bytes[0] = type;                                            // u32 length 0 follows
    pkt->ready = true;
    {
        std::lock_guard<std::mutex> l(e->m);
        e->pending.push_back(pkt);
    }
    e->cv.notify_all();
    return PFV_OK;
}

extern "C" int pfv_encoder_encode_dropframe(pfv_encoder *e)
{
    if (!e) return set_error(PFV_ERR_BAD_ARG, "NULL encoder");
    if (e->finished) return set_error(PFV_ERR_STATE, "encoder is finished (src/enc.rs:176)");
    return encoder_literal_packet(e, 1);                             // write_drop_packet, src/enc.rs:229-235
}

extern "C" int pfv_encoder_finish(pfv_encoder *e)
{
    if (!e) return set_error(PFV_ERR_BAD_ARG, "NULL encoder");
    if (e->finished) return set_error(PFV_ERR_STATE, "finish() called twice (assert!(!self.finished), src/enc.rs:183)");
    e->finished = true;
    encoder_literal_packet(e, 0);                                    // write_eof, src/enc.rs:221-227
    return encoder_flush(e);
}

// Encoder<W: Write> (src/enc.rs:12-26): what has been written so far (the header) goes out at once, every later packet as
// soon as it is finished and in order.
extern "C" int pfv_encoder_set_writer(pfv_encoder *e, pfv_write_fn writer, void *user)
{
    if (!e || !writer) return set_error(PFV_ERR_BAD_ARG, "NULL argument");
    int rc = encoder_flush(e);
    if (rc) return rc;
    if (!e->stream.empty() && writer(user, e->stream.data(), e->stream.size()) != 0)
        return set_error(PFV_ERR_IO, "the writer callback failed (io::Error of W::write_all)");
    std::lock_guard<std::mutex> l(e->m);                             // (the writer thread is idle: nothing is pending)
    e->stream.clear();
    e->writer = writer;
    e->writer_user = user;
    return PFV_OK;
}

extern "C" int pfv_encoder_bytes(pfv_encoder *e, const uint8_t **data, size_t *len)
{
    if (!e || !data || !len) return set_error(PFV_ERR_BAD_ARG, "NULL argument");
    int rc = encoder_flush(e);
    if (rc) return rc;
    *data = e->stream.data();
    *len = e->stream.size();
    return PFV_OK;
}

extern "C" pfv_ctx *pfv_encoder_ctx(pfv_encoder *e) { return e ? e->ctx : nullptr; }
extern "C" uint32_t pfv_encoder_prev_frame_slot(const pfv_encoder *e) { return e ? e->prev_slot : 0; }

extern "C" void pfv_encoder_close(pfv_encoder *e)
{
    if (!e) return;
    if (!e->finished) pfv_encoder_finish(e);                         // Drop, src/enc.rs:28-34
    else encoder_flush(e);
    {
        std::lock_guard<std::mutex> l(e->m);
        e->stop = true;
    }
    e->cv.notify_all();
    if (e->submitter.joinable()) e->submitter.join();
    if (e->writer_thread.joinable()) e->writer_thread.join();
    if (e->trace && e->n_frames)
        fprintf(stderr, "[pfv encoder] %llu frames, per frame: calling thread waits for a free work item %.1f us, copies the planes "
                        "%.1f us; submitter thread %.1f us; a pool thread waits for the frame's GPU work %.1f us and codes it in %.1f us; "
                        "writer thread %.1f us\n", (unsigned long long)e->n_frames,
                1e6 * e->t_wait / e->n_frames, 1e6 * e->t_copy / e->n_frames, 1e6 * e->t_submit / e->n_frames,
                1e6 * e->t_gpuwait / e->n_frames, 1e6 * e->t_entropy / e->n_frames, 1e6 * e->t_write / e->n_frames);
    e->pool.reset();
    e->copy_pool.reset();
    if (e->ctx) { pfv_sync(e->ctx); pfv_ctx_destroy(e->ctx); }
    delete e;
}
