#!/bin/bash
# round 2, last 2-GPU visit: the final code's bench line under torchrun at N = 2, both arms; the two-device parity test
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zz_n2; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
echo "rc=$?" >> $O/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "two_devices" > $O/t_2dev.log 2>&1
ls -la $O; tail -n 3 $O/bench_n2.err; tail -n 2 $O/t_2dev.log
