// decode_speed.cpp — the reference's own speed test against the C ABI, in C++ (no Python, no torch).
//
// pfv-rs measures itself with `test_decode_speed_2` (src/lib.rs:310-335): read test2.pfv into memory, then 50 times
// { Decoder::new(Cursor::new(&data), 6); while advance_frame(|frame| black_box(frame)) {} } and print
// "Decoded {} frames in {} ms".  This program is that test with pfv_decoder_* in place of pfv_rs::dec::Decoder:
//
//   decode_speed <stream.pfv> [runs=50] [threads=6] [device=0]
//
// and, because the reference's fixtures are git-LFS stubs, it can make its own input first with pfv_encoder_*
// (the mirror of test_encode_2, src/lib.rs:271-292: fps 30, quality 2, a key frame every 60 frames):
//
//   decode_speed --make <out.pfv> <width> <height> <frames>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <vector>

#include "../include/pfv_b200.h"

static int die(const char *what, int rc)
{
    fprintf(stderr, "%s: status %d: %s\n", what, rc, pfv_last_error());
    return 1;
}

static int make_stream(const char *path, uint32_t w, uint32_t h, uint32_t frames)
{
    pfv_encoder *enc = nullptr;
    int rc = pfv_encoder_open(w, h, 30, 2, 6, 0, &enc);            // src/lib.rs:274-277
    if (rc) return die("pfv_encoder_open", rc);
    std::vector<uint8_t> y((size_t)w * h), u((size_t)(w / 2) * (h / 2)), v(u.size());
    for (uint32_t t = 0; t < frames; t++) {
        // a drifting gradient with a moving bright square: key frames, skipped blocks, motion and residuals
        for (uint32_t r = 0; r < h; r++)
            for (uint32_t c = 0; c < w; c++) {
                const uint32_t sq = (c + 3 * t) % w < 64 && (r + 2 * t) % h < 64 ? 90 : 0;
                y[(size_t)r * w + c] = (uint8_t)((c / 4 + r / 3 + sq + ((c * 7 + r * 13 + t) % 5)) & 255);
            }
        for (size_t i = 0; i < u.size(); i++) { u[i] = (uint8_t)(118 + (i + t) % 17); v[i] = (uint8_t)(140 - (i + 2 * t) % 23); }
        rc = t % 60 == 0 ? pfv_encoder_encode_iframe(enc, y.data(), u.data(), v.data())     // src/lib.rs:281-288
                         : pfv_encoder_encode_pframe(enc, y.data(), u.data(), v.data());
        if (rc) return die("encode", rc);
    }
    if ((rc = pfv_encoder_finish(enc))) return die("pfv_encoder_finish", rc);
    const uint8_t *data; size_t len;
    if ((rc = pfv_encoder_bytes(enc, &data, &len))) return die("pfv_encoder_bytes", rc);
    FILE *f = fopen(path, "wb");
    if (!f || fwrite(data, 1, len, f) != len) { perror(path); return 1; }
    fclose(f);
    printf("Wrote %s: %u frames %ux%u, %zu bytes\n", path, frames, w, h, len);
    pfv_encoder_close(enc);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 6 && strcmp(argv[1], "--make") == 0)
        return make_stream(argv[2], (uint32_t)atoi(argv[3]), (uint32_t)atoi(argv[4]), (uint32_t)atoi(argv[5]));
    if (argc < 2) { fprintf(stderr, "usage: %s <stream.pfv> [runs] [threads] [device] | --make <out.pfv> <w> <h> <frames>\n", argv[0]); return 2; }
    const int runs = argc > 2 ? atoi(argv[2]) : 50, threads = argc > 3 ? atoi(argv[3]) : 6, device = argc > 4 ? atoi(argv[4]) : 0;
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    std::vector<uint8_t> data;
    uint8_t buf[1 << 16];
    for (size_t n; (n = fread(buf, 1, sizeof(buf), f)) > 0;) data.insert(data.end(), buf, buf + n);   // src/lib.rs:312-316
    fclose(f);
    unsigned long long checksum = 0;
    for (int run = 0; run < runs; run++) {                          // src/lib.rs:320
        pfv_decoder *dec = nullptr;
        int rc = pfv_decoder_open(data.data(), data.size(), device, (uint32_t)threads, 0, &dec);
        if (rc) return die("pfv_decoder_open", rc);
        const size_t ysz = (size_t)pfv_decoder_width(dec) * pfv_decoder_height(dec);
        int outframes = 0;
        const auto t0 = std::chrono::steady_clock::now();           // src/lib.rs:325
        for (;;) {
            int got = 0;
            const uint8_t *y, *u, *v;
            rc = pfv_decoder_advance_frame(dec, &got, &y, &u, &v);  // src/lib.rs:327
            if (rc < 0) return die("pfv_decoder_advance_frame", rc);
            if (got) { outframes++; checksum += y[0] + y[ysz - 1] + u[0] + v[0]; }   // black_box(frame)
            if (rc == 0) break;
        }
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        printf("Decoded %d frames in %.0f ms\n", outframes, ms);    // src/lib.rs:332
        pfv_decoder_close(dec);
    }
    printf("checksum %llu\n", checksum);
    return 0;
}
