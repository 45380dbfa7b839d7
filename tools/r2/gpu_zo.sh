#!/bin/bash
# round 2, visit zo: the Encoder's stream buffer as a growing anonymous mapping (the writer thread was the serial stage)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zo; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_codec.py -m gpu -q -x -k "encoder or round_trip or writer" > $O/t.log 2>&1
echo "rc=$?" >> $O/t.log
cat /sys/kernel/mm/transparent_hugepage/enabled > $O/trace.txt
timeout 300 python tools/exp/enc_trace.py >> $O/trace.txt 2>&1
for hh in 5 7 2; do echo "copy helpers $hh" >> $O/trace.txt; PFV_ENCODER_COPY_HELPERS=$hh timeout 300 python tools/exp/enc_trace.py >> $O/trace.txt 2>&1; done
tail -n 3 $O/t.log; grep "480 frames\|always\|helpers" $O/trace.txt
