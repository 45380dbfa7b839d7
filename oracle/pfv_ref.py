"""Second, independent restatement of the reference's macroblock arithmetic in pure Python.

TEST INFRASTRUCTURE ONLY (imported by tests/ alone).  PARITY UNPINNED by the reference's own tests
(SURVEY.md §8c): this file exists so that two restatements written separately from the Rust source
(this one and oracle/pfv_oracle.c) can be checked against each other and against the derived KATs.
Written from /root/reference/src/{dct,common}.rs; i32 wrapping emulated explicitly; Rust `/` on
integers truncates toward zero.
"""

FP_BITS = 8  # dct.rs:1

DCT_SCALE_FACTOR = [  # dct.rs:4-13
    32, 37, 34, 26, 32, 26, 34, 37, 37, 43, 39, 31, 37, 31, 39, 43,
    34, 39, 35, 28, 34, 28, 35, 39, 26, 31, 28, 22, 26, 22, 28, 31,
    32, 37, 34, 26, 32, 26, 34, 37, 26, 31, 28, 22, 26, 22, 28, 31,
    34, 39, 35, 28, 34, 28, 35, 39, 37, 43, 39, 31, 37, 31, 39, 43]
Q_TABLE_INTRA = [  # dct.rs:16-25
    8, 16, 19, 22, 26, 27, 29, 34, 16, 16, 22, 24, 27, 29, 34, 37,
    19, 22, 26, 27, 29, 34, 34, 38, 22, 22, 26, 27, 29, 34, 37, 40,
    22, 26, 27, 29, 32, 35, 40, 48, 26, 27, 29, 32, 35, 40, 48, 58,
    26, 27, 29, 34, 38, 46, 56, 69, 27, 29, 35, 38, 46, 56, 69, 83]
Q_TABLE_INTER = [16] * 64  # dct.rs:28-37
INV_ZIGZAG_TABLE = [  # dct.rs:39-42
    0, 1, 5, 6, 14, 15, 27, 28, 2, 4, 7, 13, 16, 26, 29, 42, 3, 8, 12, 17, 25, 30, 41, 43, 9, 11, 18, 24, 31,
    40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61,
    35, 36, 48, 49, 57, 58, 62, 63]
ZIGZAG_TABLE = [  # dct.rs:44-47
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7,
    14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39,
    46, 53, 60, 61, 54, 47, 55, 62, 63]


def w32(x):
    """wrap to i32 (release-mode Rust)"""
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def tdiv(a, b):
    """Rust integer division: truncation toward zero"""
    q = abs(a) // abs(b)
    return -q if (a < 0) != (b < 0) else q


def fdct(v):  # dct.rs:176-239
    i0, i1, i2, i3, i4, i5, i6, i7 = v
    a0, a1, a2, a3 = w32(i0 + i7), w32(i1 + i6), w32(i2 + i5), w32(i3 + i4)
    a4, a5, a6, a7 = w32(i0 - i7), w32(i1 - i6), w32(i2 - i5), w32(i3 - i4)
    b0, b1, b2, b3 = w32(a0 + a3), w32(a1 + a2), w32(a0 - a3), w32(a1 - a2)
    c0, c1 = w32(b0 + b1), w32(b0 - b1)
    c2 = w32(b2 + tdiv(b2, 4) + tdiv(b3, 2))
    c3 = w32(tdiv(b2, 2) - b3 - tdiv(b3, 4))
    b4 = w32(tdiv(a7, 4) + a4 + tdiv(a4, 4) - tdiv(a4, 16))
    b7 = w32(tdiv(a4, 4) - a7 - tdiv(a7, 4) + tdiv(a7, 16))
    b5 = w32(a5 + a6 - tdiv(a6, 4) - tdiv(a6, 16))
    b6 = w32(a6 - a5 + tdiv(a5, 4) + tdiv(a5, 16))
    c4, c5, c6, c7 = w32(b4 + b5), w32(b4 - b5), w32(b6 + b7), w32(b6 - b7)
    d4, d5, d6, d7 = c4, w32(c5 + c7), w32(c5 - c7), c6
    return [c0, d4, c2, d6, c1, d5, c3, d7]


def idct(v):  # dct.rs:241-293
    c0, d4, c2, d6, c1, d5, c3, d7 = v
    c4, c5, c7, c6 = d4, w32(d5 + d6), w32(d5 - d6), d7
    b4, b5, b6, b7 = w32(c4 + c5), w32(c4 - c5), w32(c6 + c7), w32(c6 - c7)
    b0, b1 = w32(c0 + c1), w32(c0 - c1)
    b2 = w32(c2 + tdiv(c2, 4) + tdiv(c3, 2))
    b3 = w32(tdiv(c2, 2) - c3 - tdiv(c3, 4))
    a4 = w32(tdiv(b7, 4) + b4 + tdiv(b4, 4) - tdiv(b4, 16))
    a7 = w32(tdiv(b4, 4) - b7 - tdiv(b7, 4) + tdiv(b7, 16))
    a5 = w32(b5 - b6 + tdiv(b6, 4) + tdiv(b6, 16))
    a6 = w32(b6 + b5 - tdiv(b5, 4) - tdiv(b5, 16))
    a0, a1, a2, a3 = w32(b0 + b2), w32(b1 + b3), w32(b1 - b3), w32(b0 - b2)
    return [w32(a0 + a4), w32(a1 + a5), w32(a2 + a6), w32(a3 + a7),
            w32(a3 - a7), w32(a2 - a6), w32(a1 - a5), w32(a0 - a4)]


def _rows(m, f):  # dct.rs:139-145, 157-163
    out = list(m)
    for r in range(8):
        out[r * 8:r * 8 + 8] = f(out[r * 8:r * 8 + 8])
    return out


def _cols(m, f):  # dct.rs:148-154, 166-172
    out = list(m)
    for c in range(8):
        col = f([out[c + r * 8] for r in range(8)])
        for r in range(8):
            out[c + r * 8] = col[r]
    return out


def quant_encode(m, q):  # dct.rs:88-99
    out = []
    for i, idx in enumerate(ZIGZAG_TABLE):
        n = w32(m[idx] * DCT_SCALE_FACTOR[idx]) >> (FP_BITS * 2)
        v = tdiv(n, q[idx]) & 0xFFFF
        out.append(v - 0x10000 if v & 0x8000 else v)
    return out


def quant_decode(c, q):  # dct.rs:75-86 (tables indexed by the scan position)
    out = [0] * 64
    for i, idx in enumerate(INV_ZIGZAG_TABLE):
        n = w32(c[idx] * DCT_SCALE_FACTOR[idx])
        out[i] = w32(n * q[idx])
    return out


def encode_subblock(px, q):  # common.rs:287-298
    m = [(p - 128) << FP_BITS for p in px]
    return quant_encode(_cols(_rows(m, fdct), fdct), q)


def encode_subblock_delta(d, q):  # common.rs:300-311
    m = [tdiv(x, 2) << FP_BITS for x in d]
    return quant_encode(_cols(_rows(m, fdct), fdct), q)


def decode_subblock(c, q):  # common.rs:313-325
    m = _rows(_cols(quant_decode(c, q), idct), idct)
    return [min(255, max(0, (x >> FP_BITS) + 128)) for x in m]


def apply_residuals(delta_px, prev_px):  # common.rs:98-104
    return [min(255, max(0, p + (d - 128) * 2)) for d, p in zip(delta_px, prev_px)]


def make_qtables(quality):  # enc.rs:40-51, f32 arithmetic
    import struct

    def f32(x):
        return struct.unpack("f", struct.pack("f", x))[0]

    qscale = f32(quality * 0.25)

    def tab(base, half):
        out = []
        for x in base:
            a = f32(f32(float(x)) * qscale)
            if half:
                a = f32(a * 0.5)
            out.append(int(max(a, 1.0)))
        return out

    return [tab(Q_TABLE_INTRA, True), tab(Q_TABLE_INTRA, False), tab(Q_TABLE_INTER, True), tab(Q_TABLE_INTER, False)]


def block_search(src, ref, rw, rh, cx, cy, step):  # common.rs:154-204; src = 256 px, ref = padded plane
    def ssd(x, y, limit):  # common.rs:125-139
        s = 0.0
        for r in range(16):
            for c in range(16):
                d = float(src[r * 16 + c]) - float(ref[(y + r) * rw + x + c])
                s += d * d
                if s >= limit:
                    return s
        return s

    best_dx = best_dy = 0
    best = ssd(cx, cy, float("inf"))
    for my in (-1, 0, 1):
        oy = cy + my * step
        if oy < 0 or oy > rh - 16:
            continue
        for mx in (-1, 0, 1):
            if mx == 0 and my == 0:
                continue
            ox = cx + mx * step
            if ox < 0 or ox > rw - 16:
                continue
            e = ssd(ox, oy, best)
            if e < best:
                best, best_dx, best_dy = e, mx * step, my * step
    if step > 1:
        dx2, dy2, e2 = block_search(src, ref, rw, rh, cx + best_dx, cy + best_dy, step // 2)
        return best_dx + dx2, best_dy + dy2, e2
    return best_dx, best_dy, best


# --- colour / format helpers; numpy float32 keeps the reference's f32 evaluation order -------------
def f32_as_u8(f):  # Rust `as u8`: truncation toward zero, saturating
    import numpy as np
    return np.clip(np.trunc(f), 0, 255).astype(np.uint8)


def rgb_to_yuv420(rgb):  # lib.rs:337-363 + frame.rs:51-60 (reduce: common.rs:523-536); rgb = uint8[h, w, 3]
    import numpy as np
    f = np.float32
    r, g, b = (rgb[..., i].astype(f) for i in range(3))
    y = (f(0.299) * r) + (f(0.587) * g) + (f(0.114) * b)
    u = f(128.0) - (f(0.168736) * r) - (f(0.331264) * g) + (f(0.5) * b)
    v = f(128.0) + (f(0.5) * r) - (f(0.418688) * g) - (f(0.081312) * b)
    return f32_as_u8(y), f32_as_u8(u)[::2, ::2].copy(), f32_as_u8(v)[::2, ::2].copy()


def yuv420_to_rgb(y, u, v):  # lib.rs:365-395 (double: common.rs:538-556)
    import numpy as np
    f = np.float32
    uf = np.repeat(np.repeat(u, 2, 0), 2, 1).astype(f) - f(128.0)
    vf = np.repeat(np.repeat(v, 2, 0), 2, 1).astype(f) - f(128.0)
    fy = y.astype(f)
    r = fy + (f(1.402) * vf)
    g = fy - (f(0.344136) * uf) - (f(0.714136) * vf)
    b = fy + (f(1.772) * uf)
    return np.stack([f32_as_u8(r), f32_as_u8(g), f32_as_u8(b)], -1)


def rle_encode(data):  # rle.rs:9-39 -> [(num_zeroes, coeff_size, coeff)]
    out, run = [], 0
    for val in data:
        val = int(val)
        if val == 0:
            run += 1
            continue
        while run > 15:
            out.append((15, 0, 0))
            run -= 15
        c = (-val if val < 0 else val) & 0xffff          # val.abs() as u16 (wraps for i16::MIN)
        out.append((run, c.bit_length() + 1, val))       # (16 - leading_zeros) + 1
        run = 0
    while run > 15:
        out.append((15, 0, 0))
        run -= 15
    if run > 0:
        out.append((run, 0, 0))
    return out


def update_table(table, seq):  # rle.rs:41-47
    for z, s, _ in seq:
        table[z] += 1
        table[s] += 1
    return table
