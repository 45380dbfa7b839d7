#!/bin/bash
# Round-end GPU visit: parity tests, smoke, the default bench line, ncu launch lists of the bench command and one
# `--set full` capture per hot kernel.  gpurun --timeout 2400 -- bash tools/gpu_final.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_$TAG.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee $OUT/smoke_$TAG.txt
echo "== bench"; timeout 1200 python bench.py 2> $OUT/bench_$TAG.err > $OUT/bench_$TAG.json; tail -c 600 $OUT/bench_$TAG.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tee $OUT/bench_ref_$TAG.json | cut -c1-300
echo "== ncu launch list of the default bench command (no extras)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 --extras 0 --cpu-budget 0.2 --e2e 0 > $OUT/ncu_launch_$TAG.log 2>&1
for spec in "decode_i:decode_i_stream:decode_i_1080p:3:2" "decode_p:mc_copy4|residual_sb2:decode_p_1080p:12:4" "encode_p:encode_p_kernel:encode_p_1080p:6:1"; do
  IFS=: read NAME K WL SK CNT <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SK -c $CNT -f -o $OUT/prof_${NAME}_$TAG \
      python bench.py --workload $WL --steps 2 --warmup 3 --extras 0 --cpu-budget 0.1 --e2e 0 > $OUT/ncu_full_${NAME}_$TAG.log 2>&1
  tail -1 $OUT/ncu_full_${NAME}_$TAG.log
done
echo "== ncu launch list of the Encoder object (sparse encode seam: encode kernels + tokenizer + store), 1080p, 16 frames"
cat > /tmp/enc_ll.py <<'PY'
from pretty_fast_video_b200 import codec
from pretty_fast_video_b200.synth import SynthVideo
sv = SynthVideo(1920, 1080, 0x50465602)
src = [sv.frame(t) for t in range(8)]
with codec.Encoder(1920, 1080, 30, 5, num_threads=4) as enc:
    for t in range(16):
        (enc.encode_iframe if t % 8 == 0 else enc.encode_pframe)(src[t % 8])
    enc.finish()
    print(len(enc.bytes()))
PY
PYTHONPATH=$PWD timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_enc_$TAG.csv \
    python /tmp/enc_ll.py > $OUT/ncu_launch_enc_$TAG.log 2>&1
tail -1 $OUT/ncu_launch_enc_$TAG.log
echo "== device time of the sparse encode seam next to the dense one (32 x 1080p per submit, resident)"
timeout 300 python tools/exp/tok_cost.py 2>&1 | tail -6 | tee $OUT/tok_cost_$TAG.txt
ls $OUT | tr '\n' ' '

