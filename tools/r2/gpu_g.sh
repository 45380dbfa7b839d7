#!/bin/bash
# round 2, seventh visit (2 GPUs): the N > 1 bench line with its P-stream extras, the PCIe ceiling at N = 1 and 2, encode-I residency
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2g; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
for r in 1 2 3; do
  PFV_ENCODE_I_RESIDENT=$r $B --workload encode_i_1080p > $O/ei_res$r.json 2> $O/ei_res$r.err
done
timeout 300 python tools/pcie_ceiling.py > $O/pcie_n1.json 2> $O/pcie_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/pcie_ceiling.py > $O/pcie_n2.json 2> $O/pcie_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
echo "rc=$?" >> $O/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 1 --impl reference > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "two_devices" > $O/t_2dev.log 2>&1
ls -la $O
