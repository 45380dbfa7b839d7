#!/bin/bash
# round 2: 8 GPUs of one box - the PCIe ceiling with all ranks copying, then the bench line with its P-stream extras
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2n8; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
nproc > $O/nproc.txt; free -g >> $O/nproc.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/pcie_ceiling.py > $O/pcie_n8.json 2> $O/pcie_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 tools/pcie_ceiling.py > $O/pcie_n4.json 2> $O/pcie_n4.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err
echo "rc=$?" >> $O/bench_n8.err
ls -la $O
