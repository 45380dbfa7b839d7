#!/bin/bash
# sparse encode seam: its tests, the encode bench line and the Encoder stream leg: tools/gpu_tok.sh <tag>
TAG=${1:-tok}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_codec.py tests/test_gpu_parity.py -x -q -k "sparse_encode or encoder or encode or round_trip" 2>&1 | tail -15 | tee $OUT/pytest_$TAG.txt
timeout 600 python bench.py --workload encode_p_1080p --steps 6 --warmup 3 --extras 0 --cpu-budget 2 2> $OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print('value %.0f frac %.3f e2e %s' % (d['value'], d['roofline']['frac'], json.dumps(d['e2e'])))
"
tail -3 $OUT/bench_$TAG.err
timeout 600 python - <<'PY' 2>&1 | tail -8
import time, os, numpy as np
from pretty_fast_video_b200 import codec
from pretty_fast_video_b200.synth import SynthVideo
w, h, gop = 1920, 1080, 15
sv = SynthVideo(w, h, 0x50465602)
src = [sv.frame(t) for t in range(gop + 9)]
def enc_pass(n, threads, dense):
    os.environ["PFV_ENCODER_DENSE"] = dense
    with codec.Encoder(w, h, 30, 5, num_threads=threads) as enc:
        t0 = time.perf_counter()
        for t in range(n):
            (enc.encode_iframe if t % gop == 0 else enc.encode_pframe)(src[t % gop + ((t // gop) % 4) * 3])
        enc.finish()
        dt = time.perf_counter() - t0
        return n / dt, len(enc.bytes())
for threads in (4, 16):
    for dense in ("1", "0"):
        enc_pass(60, threads, dense)
        r = [enc_pass(240, threads, dense) for _ in range(3)]
        print("Encoder threads %2d dense=%s: %.0f frames/s (bytes %d)" % (threads, dense, max(x[0] for x in r), r[0][1]))
PY
