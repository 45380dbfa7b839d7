#!/usr/bin/env python
"""What the box gives the end-to-end legs: N processes (one per GPU) copy pinned host <-> device in BOTH directions at once with
plain cudaMemcpyAsync, chunk sizes as in bench.py's e2e legs (8 x 1080p frames of coefficients up, 8 pictures down).

  python tools/pcie_ceiling.py                                       # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/pcie_ceiling.py

Prints one JSON line: per-rank and summed GB/s each way, the GPU's NUMA node and local CPUs (sysfs).  bench.py runs the same
measurement inside every e2e leg (`e2e.pcie_ceiling_gbs`, `e2e.frac_of_pcie_ceiling`)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    dist = bench.Dist(int(os.environ.get("WORLD_SIZE", "1")))
    torch.cuda.set_device(dist.local)
    binding = bench.bind_to_gpu_numa_node(torch, dist.local)
    nb = 12240
    out = {}
    for name, h2d, d2h in (("dense_decode", 8 * nb * 512, 8 * 3110400), ("sparse_decode", 8 * 300000, 8 * 3110400),
                           ("sparse_encode", 8 * 3110400, 8 * 450000)):
        up, down = bench.pcie_ceiling(torch, dist, h2d, d2h, seconds=0.5)
        out[name] = {"h2d_gbs_rank0": up, "d2h_gbs_rank0": down, "h2d_gbs_sum": dist.sum(up), "d2h_gbs_sum": dist.sum(down),
                     "h2d_chunk_bytes": h2d, "d2h_chunk_bytes": d2h}
    if dist.rank == 0:
        print(json.dumps({"n_gpus": dist.world, "host_binding_rank0": binding, "host_cpus": os.cpu_count(), "pcie": out}))
    dist.close()


if __name__ == "__main__":
    main()
