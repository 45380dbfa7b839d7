#!/bin/bash
# round 2, visit zw: the Encoder's stream buffer with and without huge pages advised (run-to-run spread over six passes each)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zw; mkdir -p $O
cat /sys/kernel/mm/transparent_hugepage/defrag > $O/enc.txt
echo "4 KB pages" >> $O/enc.txt; timeout 300 python tools/exp/enc_trace.py >> $O/enc.txt 2>&1
echo "huge pages advised" >> $O/enc.txt; PFV_STREAM_HUGEPAGES=1 timeout 300 python tools/exp/enc_trace.py >> $O/enc.txt 2>&1
echo "4 KB pages" >> $O/enc.txt; timeout 300 python tools/exp/enc_trace.py >> $O/enc.txt 2>&1
grep "frames/s\|pages\|madvise\|writer" $O/enc.txt | sed 's/.*writer thread/writer thread/' | cut -c1-200
