#!/usr/bin/env python
"""bench.py — throughput of the macroblock hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME] [--extras 0|1]

A "step" is one pass of the hot path over one batch of synthetic frames.  Workloads (BASELINE.json configs):
  decode_i_1080p  configs[1]  64 independent 1080p key frames per GPU, one launch per step           (default)
  decode_p_1080p  configs[2]  32 GOPs x 15 frames (1 key frame / 15), frame k of every GOP per launch
  encode_p_1080p  configs[3]  the same GOPs encoded (full SSD block search), frame k of every GOP per launch
  decode_p_4k     configs[4]  16 GOPs x 15 frames of 3840x2160 per GPU, GOPs sharded over the ranks
  encode_i_1080p   64 independent 1080p key frames encoded (+ closed-loop reconstruction) per launch

`value`  : whole-job frames/s with the inputs resident in HBM (CUDA events on the launching stream).
`e2e`    : the same through the C ABI with pinned HOST buffers, H2D of every token/header/source byte and D2H of
           every decoded plane (or RLE entry) inside the timed region.  `e2e.value` is the seam the product's own
           Decoder / Encoder use (sparse: pfv_decode_submit_sparse / pfv_encode_submit_sparse - the entropy decoder's
           tokens go up, the device's RLE sequence comes down); `e2e.dense` is the same leg with the reference's dense
           Vec<i16> seam (pfv_decode_submit / pfv_encode_submit).  `e2e.pcie_ceiling_gbs` is what concurrent pinned
           cudaMemcpyAsync in both directions reaches on this box with all ranks copying at once.
`verified`: after every timed leg (outside the timed region) the first and the last job's final frame is read back
           and compared with the oracle on the same inputs; a mismatch fails the run.
`roofline`: algorithmic bytes per launch / mean launch time, against MEASURED_PEAKS.json's HBM copy bandwidth.
`cpu_baseline`: the oracle (plain-C restatement of the reference algorithm, OpenMP over macroblocks like the
           reference's rayon par_iter) on this box's host cores, on a bounded sample.  The Rust reference itself
           cannot be built here (no cargo/rustc), so kind = "port".

Multi-GPU: one process per GPU (torchrun), frames/GOPs sharded over ranks, NO collective on the data path
(weak scaling: per-GPU work fixed); barrier + max-over-ranks timing through torch.distributed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MB_BYTES_DEC_I = 768          # 512 R coeff + 256 W          
MB_BYTES_DEC_P_CODED = 1027   # 3 hdr + 512 coeff + 256 ref + 256 W
MB_BYTES_DEC_P_SKIP = 515
MB_BYTES_ENC_P_CODED = 1283   # 256 src + 256 ref + 3 + 512 + 256
MB_BYTES_ENC_P_SKIP = 771
MB_BYTES_ENC_I = 1024

WORKLOADS = {
    # key-frame workloads: one launch takes 64 frames (0.1 - 0.25 ms); a step is PASSES launches over the resident batch so that the
    # time the first launch of the timed region waits for its host submit (~50 us, once) does not weigh on a 2 ms measurement
    "decode_i_1080p": dict(w=1920, h=1080, frames=64, gops=0, gop=1, quality=5, seed=0x50465601),
    "decode_p_1080p": dict(w=1920, h=1080, frames=0, gops=32, gop=15, quality=5, seed=0x50465602),
    # the same stream with 64 GOPs side by side (as many macroblocks per launch as decode_p_4k has): what a deeper batch is worth
    "decode_p_1080p_64": dict(w=1920, h=1080, frames=0, gops=64, gop=15, quality=5, seed=0x50465602),
    "encode_p_1080p": dict(w=1920, h=1080, frames=0, gops=32, gop=15, quality=5, seed=0x50465602),
    "decode_p_4k": dict(w=3840, h=2160, frames=0, gops=16, gop=15, quality=5, seed=0x50465603),
    "encode_i_1080p": dict(w=1920, h=1080, frames=64, gops=0, gop=1, quality=5, seed=0x50465601),
    # stress stream of SURVEY 8d: uniform-random pixels, every sub-block dense (worst case for the decode kernels)
    "decode_i_1080p_dense": dict(w=1920, h=1080, frames=64, gops=0, gop=1, quality=5, seed=0x50465604, kind="random"),
}


# ----------------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


PASSES = 4                      # launches per step of the key-frame (gop = 1) workloads, see WORKLOADS
SEARCH_CANDIDATES_PER_MB = 33   # block_search: the centre + 4 levels x 8 neighbours (src/common.rs:154-204), out-of-plane ones skipped
DISTINCT_SEQUENCES = 8          # GOP workloads: lane g plays sequence g mod 8 (building 32 distinct 1080p sequences in numpy costs more
                                # than the whole bench; every lane still has its own buffers in HBM)


def config_of(workload):
    """The `config` object, identical in both arms (the driver compares them)."""
    cfg = WORKLOADS[workload]
    per_step = cfg["frames"] * PASSES if cfg["gop"] == 1 else cfg["gops"] * cfg["gop"]
    return {"workload": workload, "width": cfg["w"], "height": cfg["h"], "quality": cfg["quality"],
            "frames_per_step_per_gpu": per_step, "frames_per_launch": cfg["frames"] if cfg["gop"] == 1 else cfg["gops"], "gop": cfg["gop"],
            "l2": "inputs+outputs per step exceed the 126 MB L2 (no flush needed)",
            "sharding": "frames/GOPs split over ranks, no collective on the data path"}


def bind_to_gpu_numa_node(torch, index):
    """Run this rank's host threads (and, by first touch, its pinned arenas) on the CPUs next to its GPU."""
    try:
        pr = torch.cuda.get_device_properties(index)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(path + "/numa_node") as f:
            node = int(f.read().strip())
        with open(path + "/local_cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "local_cpulist": cpulist, "bound_cpus": len(cpus) if cpus else len(allowed),
                "host_cpus": len(allowed)}
    except Exception as ex:
        return {"error": repr(ex)}


def pcie_ceiling(torch, dist, h2d_bytes, d2h_bytes, seconds=0.4):
    """What the box gives an e2e leg: every rank at once copies pinned host <-> device in both directions with plain
    cudaMemcpyAsync on two streams, one up-chunk and one down-chunk per "submit" of the sizes the leg uses.  The two directions
    advance in lockstep (pair k starts when pair k-1 has started both ways), so the direction that moves more bytes runs
    flat out WITH the other direction's traffic beside it, as in the leg itself; host buffers rotate over 256 MB each way
    (a single re-used buffer stays in the host's last-level cache and flatters the number).  Returns this rank's GB/s
    (h2d, d2h); the smaller direction's figure is paced by the larger one's and is not a ceiling."""
    dev = torch.device("cuda", torch.cuda.current_device())
    m_up = max(1, min(64, (256 << 20) // max(h2d_bytes, 1)))
    m_dn = max(1, min(64, (256 << 20) // max(d2h_bytes, 1)))
    hs = torch.empty((m_up, h2d_bytes), dtype=torch.uint8).pin_memory()
    hd = torch.empty((m_dn, d2h_bytes), dtype=torch.uint8).pin_memory()
    ds = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    dd = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def burst(n):
        prev1 = prev2 = None
        for k in range(n):
            e1, e2 = torch.cuda.Event(), torch.cuda.Event()
            with torch.cuda.stream(s1):
                if prev2 is not None:
                    s1.wait_event(prev2)
                e1.record(s1)
                ds.copy_(hs[k % m_up], non_blocking=True)
            with torch.cuda.stream(s2):
                if prev1 is not None:
                    s2.wait_event(prev1)
                e2.record(s2)
                hd[k % m_dn].copy_(dd, non_blocking=True)
            prev1, prev2 = e1, e2
    burst(4)
    torch.cuda.synchronize()
    dist.barrier()
    n = max(8, min(2000, int(seconds * 45e9 / max(h2d_bytes, d2h_bytes, 1))))
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(s1); e[2].record(s2)
    burst(n)
    e[1].record(s1); e[3].record(s2)
    torch.cuda.synchronize()
    dist.barrier()
    return n * h2d_bytes / e[0].elapsed_time(e[1]) / 1e6, n * d2h_bytes / e[2].elapsed_time(e[3]) / 1e6


def frac_of_ceiling(h2d_bytes, d2h_bytes, ms, ceil_h2d, ceil_d2h):
    """achieved / ceiling of the direction that moves more bytes (the other one is paced by it, see pcie_ceiling)"""
    if d2h_bytes >= h2d_bytes:
        return d2h_bytes / ms / 1e6 / ceil_d2h
    return h2d_bytes / ms / 1e6 / ceil_h2d


def ncu_traffic(kernel_key):
    """dram bytes per launch from the committed ncu summary (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(kernel_key)
    return None


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


class Dist:
    def __init__(self, want_gpus):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.torch = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        if want_gpus != self.world and self.rank == 0 and want_gpus > 1:
            print(f"[bench] --gpus {want_gpus} but WORLD_SIZE={self.world}: launch with torchrun "
                  f"--nproc-per-node {want_gpus}", file=sys.stderr)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
# building the synthetic streams ON THE GPU with the engine's own encoder (no oracle on this path)
# ----------------------------------------------------------------------------------------------------
class Streams:
    """Dense coefficients + macroblock headers for `lanes` independent GOPs of `gop` frames each, resident in
    HBM (torch tensors) and, lazily, in pinned host memory."""

    def __init__(self, torch, cfg, rank, stream):
        from pretty_fast_video_b200 import PFV_FRAME_I, PFV_FRAME_P, Engine, make_qtables
        from pretty_fast_video_b200.engine import EncodeJob
        from pretty_fast_video_b200.synth import SynthVideo
        self.torch, self.cfg = torch, cfg
        w, h = cfg["w"], cfg["h"]
        self.lanes = cfg["frames"] if cfg["gop"] == 1 else cfg["gops"]
        self.gop = cfg["gop"]
        self.qt, self.px_err = make_qtables(cfg["quality"])
        eng = Engine(w, h, self.qt, nslots=2 * self.lanes, max_jobs=self.lanes, device=torch.cuda.current_device(),
                     stream=stream.cuda_stream)
        g = eng.geometry
        self.geo = g
        self.nb = g.nb
        n = self.lanes * self.gop
        dev = torch.device("cuda", torch.cuda.current_device())
        self.d_coeff = torch.zeros((self.gop, self.lanes, g.nb * 256), dtype=torch.int16, device=dev)
        self.d_hdr = torch.zeros((self.gop, self.lanes, g.nb, 4), dtype=torch.uint8, device=dev)
        ysz, csz = w * h, (w // 2) * (h // 2)
        self.src_bytes = ysz + 2 * csz
        self.d_src = torch.zeros((self.gop, self.lanes, self.src_bytes), dtype=torch.uint8, device=dev)
        # I-only workload: every frame of one moving sequence; GOP workloads: lane g plays sequence g mod DISTINCT_SEQUENCES
        cache = {}
        for lane in range(self.lanes):
            if self.gop == 1:
                sv = SynthVideo(w, h, cfg["seed"] + 1000 * rank, kind=cfg.get("kind", "moving")) if lane == 0 else sv
                frames = [np.concatenate([p.ravel() for p in sv.frame(lane)])]
            else:
                key = lane % DISTINCT_SEQUENCES
                if key not in cache:
                    sv = SynthVideo(w, h, cfg["seed"] + 1000 * rank + key, kind=cfg.get("kind", "moving"))
                    cache[key] = [np.concatenate([p.ravel() for p in sv.frame(t)]) for t in range(self.gop)]
                frames = cache[key]
            for k, buf in enumerate(frames):
                self.d_src[k, lane].copy_(torch.from_numpy(buf))
        del cache
        torch.cuda.synchronize()
        cur = [2 * i for i in range(self.lanes)]
        for k in range(self.gop):
            jobs = []
            for lane in range(self.lanes):
                base = self.d_src[k, lane].data_ptr()
                dst = cur[lane] ^ 1
                jobs.append(EncodeJob(PFV_FRAME_I if k == 0 else PFV_FRAME_P, dst,
                                      (base, base + ysz, base + ysz + csz), self.d_coeff[k, lane].data_ptr(),
                                      ref_slot=cur[lane], px_err=self.px_err, hdr_out=self.d_hdr[k, lane].data_ptr(),
                                      device_ptrs=True))
                cur[lane] = dst
            with torch.cuda.stream(stream):
                eng.encode_submit(jobs)
        eng.sync()
        eng.close()
        hdr = self.d_hdr.cpu().numpy()
        self.coded = hdr[..., 2] != 0                       # [gop, lanes, nb]
        if self.gop > 1:
            self.coded[0] = True
            # dec.rs:376: skipped macroblocks carry zero coefficients in the dense array
            mask = torch.from_numpy(~self.coded[1:]).to(dev)
            self.d_coeff[1:].view(self.gop - 1, self.lanes, g.nb, 256)[mask] = 0
        else:
            self.coded[:] = True
        self.mv_nonzero = float(((hdr[1:, ..., 0] != 0) | (hdr[1:, ..., 1] != 0)).mean()) if self.gop > 1 else 0.0
        self.coded_frac = float(self.coded[1:].mean()) if self.gop > 1 else 1.0
        # share of sub-blocks with any AC coefficient (the ones that take the full transform in the decode kernels)
        self.general_frac = float((self.d_coeff.view(-1, 64)[:, 1:] != 0).any(dim=1).float().mean().item())
        self._host = None

    def host(self):
        """pinned copies (coefficients, headers, source planes)"""
        if self._host is None:
            from pretty_fast_video_b200 import PinnedArena
            n = self.d_coeff.numel() * 2 + self.d_hdr.numel() + self.d_src.numel() + 3 * 4096
            arena = PinnedArena(n)
            hc = arena.take(tuple(self.d_coeff.shape), np.int16)
            hh = arena.take(tuple(self.d_hdr.shape), np.uint8)
            hs = arena.take(tuple(self.d_src.shape), np.uint8)
            hc[...] = self.d_coeff.cpu().numpy()
            hh[...] = self.d_hdr.cpu().numpy()
            hs[...] = self.d_src.cpu().numpy()
            self._host = (arena, hc, hh, hs)
        return self._host


# ----------------------------------------------------------------------------------------------------
# the timed legs
# ----------------------------------------------------------------------------------------------------
def time_steps(torch, dist, stream, step_fn, sync_fn, steps, warmup, wall=False):
    """W untimed + exactly K timed steps, barrier + synchronize on both sides; returns (ms_per_step max over
    ranks, this rank's ms_per_step)."""
    for _ in range(warmup):
        step_fn()
    sync_fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    if wall:
        t0 = time.perf_counter()
        for _ in range(steps):
            step_fn()
        sync_fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / steps
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                step_fn()
            e1.record(stream)
        sync_fn()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    dist.barrier()
    return dist.max(ms), ms


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import pfvo
    return pfvo


def verify_decode(torch, st: Streams, eng, final_slot_of_lane):
    """Outside the timed region: the first and the last lane's framebuffer after the last frame of the step against
    the oracle decoding the same coefficients (the whole GOP chain for P workloads)."""
    pfvo = _oracle()
    og = pfvo.geometry_for(st.cfg["w"], st.cfg["h"])
    nt = min(16, os.cpu_count() or 1)
    ok = True
    for lane in sorted({0, st.lanes - 1}):
        frame = pfvo.frame_init(og)
        for k in range(st.gop):
            c = st.d_coeff[k, lane].cpu().numpy()
            if k == 0:
                pfvo.decode_iframe_coeffs(og, st.qt, (0, 1, 1), c, frame, nt)
            else:
                pfvo.decode_pframe_coeffs(og, st.qt, (2, 3, 3), st.d_hdr[k, lane].cpu().numpy(), c, frame, nt)
        ok &= bool(np.array_equal(eng.slot_read(final_slot_of_lane(lane)), frame))
    return ok


def verify_encode(torch, st: Streams, eng, final_slot_of_lane, d_c, d_h):
    """The first and the last lane: headers, coefficients of coded macroblocks and reconstruction of the step's last
    frame against the oracle encoding the same source planes (the whole GOP chain for P workloads)."""
    pfvo = _oracle()
    w, h = st.cfg["w"], st.cfg["h"]
    og = pfvo.geometry_for(w, h)
    nt = min(16, os.cpu_count() or 1)
    ysz, csz = w * h, (w // 2) * (h // 2)
    ok = True
    for lane in sorted({0, st.lanes - 1}):
        prev = pfvo.frame_init(og)
        hd = None
        for k in range(st.gop):
            b = st.d_src[k, lane].cpu().numpy()
            y, u, v = b[:ysz].reshape(h, w), b[ysz:ysz + csz].reshape(h // 2, w // 2), b[ysz + csz:].reshape(h // 2, w // 2)
            if k == 0:
                c = pfvo.encode_iframe_coeffs(og, st.qt, y, u, v, prev, nt)
            else:
                hd, c = pfvo.encode_pframe_coeffs(og, st.qt, st.px_err, y, u, v, prev, nt)
        got_c = d_c[lane].cpu().numpy().reshape(-1, 256)
        if st.gop == 1:
            ok &= bool(np.array_equal(got_c, c.reshape(-1, 256)))
        else:
            coded = hd[:, 2] != 0
            ok &= bool(np.array_equal(d_h[lane].cpu().numpy(), hd))
            ok &= bool(np.array_equal(got_c[coded], c.reshape(-1, 256)[coded]))
        ok &= bool(np.array_equal(eng.slot_read(final_slot_of_lane(lane)), prev))
    return ok


def run_decode(torch, dist, stream, st: Streams, steps, warmup, do_e2e=True):
    from pretty_fast_video_b200 import PFV_FRAME_I, PFV_FRAME_P, Engine
    from pretty_fast_video_b200.engine import DecodeJob
    g, L, G = st.geo, st.lanes, st.gop
    w, h = st.cfg["w"], st.cfg["h"]
    dev_index = torch.cuda.current_device()

    def make_jobs(eng, k, cur, coeff_ptr, hdr_ptr, device, outs=None):
        jobs = []
        for lane in range(L):
            dst = cur[lane] ^ 1
            jobs.append(DecodeJob(PFV_FRAME_I if k == 0 else PFV_FRAME_P, dst, coeff_ptr(k, lane),
                                  (0, 1, 1) if k == 0 else (2, 3, 3), ref_slot=cur[lane],
                                  hdr=hdr_ptr(k, lane) if k else None, device_ptrs=device,
                                  out=outs(k, lane) if outs else None,
                                  dense_hint=(k == 0 and st.general_frac > 0.5)))   # the caller knows its content (PFV_JOB_DENSE)
            cur[lane] = dst
        return eng.build_decode_jobs(jobs), jobs

    # ---- device-resident ------------------------------------------------------------------------
    eng = Engine(w, h, st.qt, nslots=2 * L, max_jobs=L, device=dev_index, stream=stream.cuda_stream)
    cur = [2 * i for i in range(L)]
    # two passes over the GOP bring every lane's ping-pong state back to where it started only for even G;
    # prebuild 2*G steps' worth of job tables so the slot pattern is exact for any G
    tables = []
    for rep in range(2):
        for k in range(G):
            tables.append(make_jobs(eng, k, cur, lambda k, l: st.d_coeff[k, l].data_ptr(),
                                    lambda k, l: st.d_hdr[k, l].data_ptr(), True))
    phase = [0]

    passes = PASSES if G == 1 else 1

    def step_dev():
        for _ in range(passes):
            base = (phase[0] % 2) * G
            for k in range(G):
                arr, jobs = tables[base + k]
                eng.decode_submit(jobs, prebuilt=arr)
            phase[0] += 1

    l0 = eng.launch_count
    max_ms, my_ms = time_steps(torch, dist, stream, step_dev, eng.sync, steps, warmup)
    launches_per_step = (eng.launch_count - l0) // (steps + warmup)
    last_jobs = tables[((phase[0] - 1) % 2) * G + G - 1][1]
    verified = verify_decode(torch, st, eng, lambda lane: last_jobs[lane].dst_slot)
    eng.close()

    # ---- end to end through host buffers ------------------------------------------------------------
    e2e = None
    if do_e2e:
        from pretty_fast_video_b200 import PinnedArena
        arena, hc, hh, hs = st.host()
        chunk = min(L, 8)                                   # jobs per submit: H2D of chunk i+1 overlaps chunk i
        ysz, csz = w * h, (w // 2) * (h // 2)
        # The caller's picture buffers keep the planes at the PADDED plane strides (y | u | v at 0, pw*ph, pw*ph + cpw*cph, as the
        # product's Decoder lays its own out): when rows are not padded (1080p, 4K) the engine then moves a picture with ONE
        # copy - the few padding rows between the planes ride along - instead of three (round 2: 42 -> ~50 GB/s D2H).
        ny_p, nc_p = g.pw * g.ph, g.cpw * g.cph
        one_copy = g.pw == w and g.cpw == w // 2
        pic_bytes = ny_p + nc_p + csz if one_copy else ysz + 2 * csz      # what one picture's copy (or copies) moves
        out_stride = ny_p + 2 * nc_p
        out_arena = PinnedArena(G * L * out_stride + 4096)
        outb = out_arena.take((G, L, out_stride), np.uint8)
        nframes = G * L
        d2h = int(pic_bytes * nframes)

        def outs(k, lane):
            b = outb[k, lane].ctypes.data
            return (b, b + ny_p, b + ny_p + nc_p)

        def check_outputs():
            """the pictures that came back over PCIe (first / last lane, last frame of the GOP) against the oracle"""
            pfvo = _oracle()
            og = pfvo.geometry_for(w, h)
            nt = min(16, os.cpu_count() or 1)
            ok = True
            for lane in sorted({0, L - 1}):
                frame = pfvo.frame_init(og)
                for k in range(G):
                    if k == 0:
                        pfvo.decode_iframe_coeffs(og, st.qt, (0, 1, 1), hc[k, lane], frame, nt)
                    else:
                        pfvo.decode_pframe_coeffs(og, st.qt, (2, 3, 3), hh[k, lane], hc[k, lane], frame, nt)
                y, u, v = pfvo.crop_frame(og, frame)
                got = outb[G - 1, lane]
                ok &= bool(np.array_equal(got[:ysz], y.ravel()) and np.array_equal(got[ny_p:ny_p + csz], u.ravel()) and
                           np.array_equal(got[ny_p + nc_p:ny_p + nc_p + csz], v.ravel()))
            return ok

        # ---- sparse coefficient transport (pfv_decode_submit_sparse): what the product's Decoder hands over ----
        from pretty_fast_video_b200 import codec
        from pretty_fast_video_b200.engine import SparseDecodeJob
        toks = [[codec.dense_to_tokens(hc[k, l], st.nb) for l in range(L)] for k in range(G)]
        ntok_total = sum(t[1].size for row in toks for t in row)
        tarena = PinnedArena(4 * ntok_total + G * L * (4 * (st.nb + 1) + 512) + 4096)
        ptoks = []
        for k in range(G):
            row = []
            for l in range(L):
                mo = tarena.take((st.nb + 1,), np.uint32); mo[...] = toks[k][l][0]
                tk = tarena.take((max(1, toks[k][l][1].size),), np.uint32); tk[:toks[k][l][1].size] = toks[k][l][1]
                row.append((mo, tk[:toks[k][l][1].size]))
            ptoks.append(row)
        del toks
        eng3 = Engine(w, h, st.qt, nslots=2 * L, max_jobs=chunk, device=dev_index, stream=stream.cuda_stream)
        cur3 = [2 * i for i in range(L)]
        tabs3 = []
        for rep in range(2):
            for k in range(G):
                jobs = []
                for lane in range(L):
                    dst = cur3[lane] ^ 1
                    jobs.append(SparseDecodeJob(PFV_FRAME_I if k == 0 else PFV_FRAME_P, dst, ptoks[k][lane][0], ptoks[k][lane][1],
                                                (0, 1, 1) if k == 0 else (2, 3, 3), ref_slot=cur3[lane],
                                                hdr=hh[k, lane] if k else None, out=outs(k, lane)))
                    cur3[lane] = dst
                tabs3.append([(eng3.build_sparse_decode_jobs(jobs[i:i + chunk]), jobs[i:i + chunk]) for i in range(0, L, chunk)])
        ph3 = [0]

        def step_sparse():
            base = (ph3[0] % 2) * G
            for k in range(G):
                for arr, jobs in tabs3[base + k]:
                    eng3.decode_submit_sparse(jobs, prebuilt=arr)
            ph3[0] += 1

        h2d_s = int(4 * ntok_total + nframes * 4 * (st.nb + 1) + (G - 1) * L * st.nb * 4)
        # the box's ceiling for THIS leg's traffic mix: per submit, tokens + headers up and `chunk` pictures down
        ceil_h2d, ceil_d2h = pcie_ceiling(torch, dist, max(4096, h2d_s * chunk // nframes), chunk * pic_bytes)
        outb[...] = 0
        s_max, s_my = time_steps(torch, dist, stream, step_sparse, eng3.sync, max(2, steps // 2), 2, wall=True)
        e2e = {"value": nframes * dist.world / (s_max * 1e-3), "unit": "frames/s", "ms_per_step": s_max, "frames_per_step": nframes,
               "h2d_bytes_per_step": h2d_s, "d2h_bytes_per_step": d2h, "jobs_per_submit": chunk,
               "seam": "sparse: pfv_decode_submit_sparse - (position,value) tokens over PCIe, dense layout rebuilt on the GPU, pictures back",
               "nonzero_coefficients_per_frame": ntok_total / nframes,
               "pcie_gbs_each_way": [h2d_s / s_my / 1e6, d2h / s_my / 1e6],
               "pcie_ceiling_gbs": [ceil_h2d, ceil_d2h],
               "frac_of_pcie_ceiling": frac_of_ceiling(h2d_s, d2h, s_my, ceil_h2d, ceil_d2h),
               "verified": check_outputs()}
        eng3.close()
        tarena.close()

        # ---- the reference's dense Vec<i16> seam (pfv_decode_submit) ----
        eng2 = Engine(w, h, st.qt, nslots=2 * L, max_jobs=chunk, device=dev_index, stream=stream.cuda_stream)
        cur2 = [2 * i for i in range(L)]
        tabs2 = []
        for rep in range(2):
            for k in range(G):
                full, jobs = make_jobs(eng2, k, cur2, lambda k, l: hc[k, l].ctypes.data, lambda k, l: hh[k, l].ctypes.data,
                                       False, outs)
                tabs2.append([(eng2.build_decode_jobs(jobs[i:i + chunk]), jobs[i:i + chunk]) for i in range(0, L, chunk)])
        ph2 = [0]

        def step_e2e():
            base = (ph2[0] % 2) * G
            for k in range(G):
                for arr, jobs in tabs2[base + k]:
                    eng2.decode_submit(jobs, prebuilt=arr)
            ph2[0] += 1

        outb[...] = 0
        ceil_h2d, ceil_d2h = pcie_ceiling(torch, dist, chunk * st.nb * 512, chunk * pic_bytes)
        e_max, e_my = time_steps(torch, dist, stream, step_e2e, eng2.sync, max(2, steps // 2), 2, wall=True)
        h2d = int(st.nb * 512 * nframes + (G - 1) * L * st.nb * 4)
        e2e["dense"] = {"value": nframes * dist.world / (e_max * 1e-3), "unit": "frames/s", "ms_per_step": e_max, "frames_per_step": nframes,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "pcie_gbs_each_way": [h2d / e_my / 1e6, d2h / e_my / 1e6], "pcie_ceiling_gbs": [ceil_h2d, ceil_d2h],
                        "frac_of_pcie_ceiling": frac_of_ceiling(h2d, d2h, e_my, ceil_h2d, ceil_d2h),
                        "seam": "dense: pfv_decode_submit - the reference's Vec<i16> of nb*256 coefficients over PCIe",
                        "verified": check_outputs()}
        eng2.close()
        out_arena.close()

    nframes = G * L
    if G == 1:
        alg = nframes * st.nb * MB_BYTES_DEC_I
    else:
        coded = st.coded[1:]
        alg = L * st.nb * MB_BYTES_DEC_I + int(coded.sum()) * MB_BYTES_DEC_P_CODED + int((~coded).sum()) * MB_BYTES_DEC_P_SKIP
    # (the e2e legs above move one pass of the batch per step; the resident leg makes `passes` of them)
    return dict(max_ms=max_ms, my_ms=my_ms, frames=nframes * passes, alg_bytes=alg * passes, launches_per_step=launches_per_step,
                e2e=e2e, verified=verified)


def run_encode(torch, dist, stream, st: Streams, steps, warmup, do_e2e=True):
    from pretty_fast_video_b200 import PFV_FRAME_I, PFV_FRAME_P, Engine, PinnedArena
    from pretty_fast_video_b200.engine import EncodeJob
    g, L, G = st.geo, st.lanes, st.gop
    w, h = st.cfg["w"], st.cfg["h"]
    ysz, csz = w * h, (w // 2) * (h // 2)
    dev_index = torch.cuda.current_device()
    dev = torch.device("cuda", dev_index)
    d_c = torch.zeros((L, st.nb * 256), dtype=torch.int16, device=dev)
    d_h = torch.zeros((L, st.nb, 4), dtype=torch.uint8, device=dev)

    def make_tables(eng, src_ptr, c_ptr, h_ptr, device, chunk):
        cur = [2 * i for i in range(L)]
        tabs = []
        for rep in range(2):
            for k in range(G):
                jobs = []
                for lane in range(L):
                    b = src_ptr(k, lane)
                    dst = cur[lane] ^ 1
                    jobs.append(EncodeJob(PFV_FRAME_I if k == 0 else PFV_FRAME_P, dst, (b, b + ysz, b + ysz + csz),
                                          c_ptr(k, lane), ref_slot=cur[lane], px_err=st.px_err, hdr_out=h_ptr(k, lane),
                                          device_ptrs=device))
                    cur[lane] = dst
                tabs.append([(eng.build_encode_jobs(jobs[i:i + chunk]), jobs[i:i + chunk]) for i in range(0, L, chunk)])
        return tabs

    eng = Engine(w, h, st.qt, nslots=2 * L, max_jobs=L, device=dev_index, stream=stream.cuda_stream)
    tabs = make_tables(eng, lambda k, l: st.d_src[k, l].data_ptr(), lambda k, l: d_c[l].data_ptr(),
                       lambda k, l: d_h[l].data_ptr(), True, L)
    ph = [0]

    passes = PASSES if G == 1 else 1

    def step_dev():
        for _ in range(passes):
            base = (ph[0] % 2) * G
            for k in range(G):
                for arr, jobs in tabs[base + k]:
                    eng.encode_submit(jobs, prebuilt=arr)
            ph[0] += 1

    l0 = eng.launch_count
    max_ms, my_ms = time_steps(torch, dist, stream, step_dev, eng.sync, steps, warmup)
    launches_per_step = (eng.launch_count - l0) // (steps + warmup)
    last_jobs = tabs[((ph[0] - 1) % 2) * G + G - 1][0][1]
    verified = verify_encode(torch, st, eng, lambda lane: last_jobs[lane].dst_slot, d_c, d_h)
    eng.close()

    e2e = None
    if do_e2e:
        arena, hc, hh, hs = st.host()
        chunk = min(L, 8)
        nframes = G * L
        # ---- the sparse encode seam (pfv_encode_submit_sparse): what the product's Encoder uses.  The run-length pass runs
        # on the device and the device itself stores each frame's RLE sequence into pinned host memory ----
        from pretty_fast_video_b200 import _native as N
        from pretty_fast_video_b200.engine import SparseEncodeJob
        cap = st.nb * 64                                            # entries per frame; the overflow flag is checked below
        eng3 = Engine(w, h, st.qt, nslots=2 * L, max_jobs=chunk, device=dev_index, stream=stream.cuda_stream)
        sa = PinnedArena(G * L * (cap * 4 + st.nb * 4 + 1024) + 8192)
        ot = sa.take((G, L, cap), np.uint32)
        os_ = sa.take((G, L, 64), np.uint32)
        oh3 = sa.take((G, L, st.nb, 4), np.uint8)
        cur = [2 * i for i in range(L)]
        tabs3 = []
        for rep in range(2):
            for k in range(G):
                jobs = []
                for lane in range(L):
                    b = hs[k, lane].ctypes.data
                    dst = cur[lane] ^ 1
                    jobs.append(SparseEncodeJob(PFV_FRAME_I if k == 0 else PFV_FRAME_P, dst, (b, b + ysz, b + ysz + csz), ot[k, lane],
                                                os_[k, lane], tok_cap=cap, ref_slot=cur[lane], px_err=st.px_err, hdr_out=oh3[k, lane]))
                    cur[lane] = dst
                tabs3.append([(eng3.build_sparse_encode_jobs(jobs[i:i + chunk]), jobs[i:i + chunk]) for i in range(0, L, chunk)])
        ph3 = [0]

        def step_sparse():
            base = (ph3[0] % 2) * G
            for k in range(G):
                for arr, jobs in tabs3[base + k]:
                    eng3.encode_submit_sparse(jobs, prebuilt=arr)
            ph3[0] += 1

        # the box's ceiling for this leg's traffic mix: `chunk` source frames up, ~0.1 of the dense coefficient bytes down
        ceil_h2d, ceil_d2h = pcie_ceiling(torch, dist, chunk * (ysz + 2 * csz), chunk * st.nb * 48)
        s_max, s_my = time_steps(torch, dist, stream, step_sparse, eng3.sync, max(2, steps // 2), 2, wall=True)
        ntok = os_[:, :, N.PFV_TOKSTATS_NTOK].astype(np.int64)
        assert not (os_[:, :, N.PFV_TOKSTATS_FLAGS] != 0).any(), "token buffer overflow in the sparse encode leg"
        # the RLE sequence that came back over PCIe (first / last lane, last frame) against rle_encode of the oracle's coefficients
        pfvo = _oracle()
        og = pfvo.geometry_for(w, h)
        nt = min(16, os.cpu_count() or 1)
        ok = True
        for lane in sorted({0, L - 1}):
            prev = pfvo.frame_init(og)
            hd = None
            for k in range(G):
                b = hs[k, lane]
                y, u, v = b[:ysz].reshape(h, w), b[ysz:ysz + csz].reshape(h // 2, w // 2), b[ysz + csz:].reshape(h // 2, w // 2)
                if k == 0:
                    c = pfvo.encode_iframe_coeffs(og, st.qt, y, u, v, prev, nt)
                else:
                    hd, c = pfvo.encode_pframe_coeffs(og, st.qt, st.px_err, y, u, v, prev, nt)
            o_tok, o_table, _ = pfvo.rle_frame(c, None if G == 1 else hd[:, 2])
            n = int(os_[G - 1, lane, N.PFV_TOKSTATS_NTOK])
            ok &= bool(n == o_tok.size and np.array_equal(ot[G - 1, lane, :n], o_tok))
            ok &= bool(np.array_equal(os_[G - 1, lane, :16].astype(np.int64) + os_[G - 1, lane, 16:32], o_table))
            if G > 1:
                ok &= bool(np.array_equal(oh3[G - 1, lane], hd))
        h2d = int(nframes * (ysz + 2 * csz))
        d2h_s = int(ntok.sum() * 4 + nframes * N.PFV_TOKSTATS_WORDS * 4 + (G - 1) * L * st.nb * 4)
        e2e = {"value": nframes * dist.world / (s_max * 1e-3), "unit": "frames/s", "ms_per_step": s_max, "frames_per_step": nframes,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_s, "jobs_per_submit": chunk,
               "rle_entries_per_frame": float(ntok.mean()),
               "seam": "sparse: pfv_encode_submit_sparse - source planes up, run-length pass on the GPU, RLE sequence stored by the device into pinned host memory",
               "pcie_gbs_each_way": [h2d / s_my / 1e6, d2h_s / s_my / 1e6], "pcie_ceiling_gbs": [ceil_h2d, ceil_d2h],
               "frac_of_pcie_ceiling": frac_of_ceiling(h2d, d2h_s, s_my, ceil_h2d, ceil_d2h),
               "verified": ok}
        eng3.close()
        sa.close()

        # ---- the reference's dense seam (pfv_encode_submit): all coefficients come back ----
        eng2 = Engine(w, h, st.qt, nslots=2 * L, max_jobs=chunk, device=dev_index, stream=stream.cuda_stream)
        oa = PinnedArena(G * L * (st.nb * 512 + st.nb * 4) + 8192)
        oc = oa.take((G, L, st.nb * 256), np.int16)
        oh = oa.take((G, L, st.nb, 4), np.uint8)
        tabs2 = make_tables(eng2, lambda k, l: hs[k, l].ctypes.data, lambda k, l: oc[k, l].ctypes.data,
                            lambda k, l: oh[k, l].ctypes.data, False, chunk)
        ph2 = [0]

        def step_e2e():
            base = (ph2[0] % 2) * G
            for k in range(G):
                for arr, jobs in tabs2[base + k]:
                    eng2.encode_submit(jobs, prebuilt=arr)
            ph2[0] += 1

        ceil_h2d, ceil_d2h = pcie_ceiling(torch, dist, chunk * (ysz + 2 * csz), chunk * st.nb * 512)
        e_max, e_my = time_steps(torch, dist, stream, step_e2e, eng2.sync, max(2, steps // 2), 2, wall=True)
        d2h = int(nframes * st.nb * 512 + (G - 1) * L * st.nb * 4)
        e2e["dense"] = {"value": nframes * dist.world / (e_max * 1e-3), "unit": "frames/s", "ms_per_step": e_max, "frames_per_step": nframes,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "pcie_gbs_each_way": [h2d / e_my / 1e6, d2h / e_my / 1e6], "pcie_ceiling_gbs": [ceil_h2d, ceil_d2h],
                        "frac_of_pcie_ceiling": frac_of_ceiling(h2d, d2h, e_my, ceil_h2d, ceil_d2h),
                        "seam": "dense: pfv_encode_submit - nb*256 int16 coefficients per frame back over PCIe"}
        eng2.close()
        oa.close()
    if G == 1:
        alg = L * st.nb * MB_BYTES_ENC_I
    else:
        coded = st.coded[1:]
        alg = L * st.nb * MB_BYTES_ENC_I + int(coded.sum()) * MB_BYTES_ENC_P_CODED + int((~coded).sum()) * MB_BYTES_ENC_P_SKIP
    return dict(max_ms=max_ms, my_ms=my_ms, frames=G * L * passes, alg_bytes=alg * passes, launches_per_step=launches_per_step,
                e2e=e2e, verified=verified)


def run_decoder_stream(torch, dist, budget_s, nthreads):
    """Stream level, the reference's own public API: pfv_rs::dec::Decoder::advance_frame over an in-memory .pfv
    (what src/lib.rs:310-335 test_decode_speed_2 times).  The stream is made with the engine's Encoder on the GPU;
    the timed region is container parse + entropy decode (host pool) + H2D tokens + kernels + D2H pictures."""
    from pretty_fast_video_b200 import codec
    from pretty_fast_video_b200.synth import SynthVideo
    w, h, gop, ngop = 1920, 1080, 15, 16
    sv = SynthVideo(w, h, 0x50465602)
    src = [sv.frame(t) for t in range(gop + 3 * 3)]              # GOP g starts 3 frames later in the sequence (g mod 4)
    with codec.Encoder(w, h, 30, 5, num_threads=nthreads, device=torch.cuda.current_device()) as enc:
        for t in range(gop * ngop):
            (enc.encode_iframe if t % gop == 0 else enc.encode_pframe)(src[t % gop + ((t // gop) % 4) * 3])
        enc.finish()
        data = enc.bytes()
    nfr = gop * ngop

    def one_pass():
        n = 0
        with codec.Decoder(data, num_threads=nthreads, device=torch.cuda.current_device()) as dec:
            t0 = time.perf_counter()
            while dec.advance_frame(lambda fr: None):
                n += 1
            dt = time.perf_counter() - t0
        assert n == nfr
        return dt

    one_pass()
    times = [one_pass() for _ in range(3)]
    gpu_fps = nfr / min(times)
    # the mirror image: Encoder.encode_iframe / encode_pframe over the same source frames -> .pfv bytes
    def enc_pass(n):
        with codec.Encoder(w, h, 30, 5, num_threads=nthreads, device=torch.cuda.current_device()) as enc:
            t0 = time.perf_counter()
            for t in range(n):
                (enc.encode_iframe if t % gop == 0 else enc.encode_pframe)(src[t % gop + ((t // gop) % 4) * 3])
            enc.finish()                                             # every packet is in the Encoder's stream buffer (its `W`)
            dt = time.perf_counter() - t0
            assert len(enc.bytes()) > 0                              # (a Python-side copy of the whole stream: outside the timed region)
            return dt
    enc_pass(gop)
    # 32 GOPs (480 frames) per pass: with 8 the fill and the drain of the Encoder's pipeline (a few frames' worth of work in flight
    # when finish() is called) were a tenth of a 24 ms pass
    enc_n = 32 * gop
    enc_times = [enc_pass(enc_n) for _ in range(3)]
    enc_fps = enc_n / min(enc_times)                                 # best of 3; the spread is reported beside it
    # the oracle's Decoder on the same bytes (entropy + MB loops, nthreads OpenMP threads for the MB loops)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import pfvo
    done, t_used = 0, 0.0
    while t_used < budget_s:
        dec = pfvo.Decoder(data, nthreads=nthreads)
        t0 = time.perf_counter()
        k = 0
        while k < 15:
            more, fr = dec.advance_frame()
            if not more:
                break
            k += 1
        t_used += time.perf_counter() - t0
        done += k
        dec.close()
    og = pfvo.geometry_for(w, h)
    e_done, e_used = 0, 0.0
    while e_used < budget_s / 2:
        oenc = pfvo.Encoder(w, h, 30, 5, nthreads=nthreads)
        t0 = time.perf_counter()
        for t in range(6):
            y_, u_, v_ = src[t]
            (oenc.encode_iframe if t == 0 else oenc.encode_pframe)(y_, u_, v_)
        e_used += time.perf_counter() - t0
        e_done += 6
        oenc.close()
    return {"value": gpu_fps, "unit": "frames/s", "frames": nfr, "stream_bytes": len(data), "host_threads": nthreads,
            "encoder": {"value": enc_fps, "unit": "frames/s", "frames_per_pass": enc_n, "spread": [enc_n / max(enc_times), enc_n / min(enc_times)],
                        "note": "Encoder.encode_iframe/encode_pframe (1 key frame / 15) to .pfv bytes: the calling thread copies the planes into pinned memory, "
                                "a submitter thread drives the GPU, a writer thread appends the packets; "
                        "planes H2D, kernels (full block search), run-length pass on the GPU, RLE sequence stored into pinned host memory, Huffman + bit packing on the host pool",
                        "cpu_baseline": {"value": e_done / e_used, "unit": "frames/s", "cores": nthreads, "kind": "port",
                                         "sample": f"{e_done} frames (1 key + 5 P) through the oracle Encoder in {e_used:.1f} s"}},
            "note": "Decoder.advance_frame over an in-memory 1920x1080 .pfv (1 key frame / 15): entropy decode on the host pool, "
                    "sparse tokens H2D, kernels, pictures D2H",
            "cpu_baseline": {"value": done / t_used, "unit": "frames/s", "cores": nthreads, "kind": "port",
                             "sample": f"{done} frames of the same stream through the oracle Decoder (entropy + MB loops) in {t_used:.1f} s"}}


def run_config1_stream(torch, dist, budget_s, nthreads):
    """BASELINE.json configs[0] restated (the LFS fixture test2.pfv is a pointer stub): a 512x384, 161-frame,
    quality-2 stream with a key frame every 60 (the parameters of src/lib.rs:271-292), decoded frame by frame through
    Decoder.advance_frame as src/lib.rs:310-335 (test_decode_speed_2) does; the oracle Decoder on ONE thread beside it.
    Small frames: a frame's kernels take a few microseconds, the per-frame host work (entropy decode, driver calls) is
    what is timed here."""
    from pretty_fast_video_b200 import codec
    from pretty_fast_video_b200.synth import SynthVideo
    pfvo = _oracle()
    w, h, n, key = 512, 384, 161, 60
    sv = SynthVideo(w, h, 0x50465600)
    with codec.Encoder(w, h, 30, 2, num_threads=nthreads, device=torch.cuda.current_device()) as enc:
        for t in range(n):
            (enc.encode_iframe if t % key == 0 else enc.encode_pframe)(sv.frame(t))
        enc.finish()
        data = enc.bytes()

    def one_pass(check=None):
        k = 0
        with codec.Decoder(data, num_threads=nthreads, device=torch.cuda.current_device()) as dec:
            t0 = time.perf_counter()
            if check is None:
                while dec.advance_frame(lambda fr: None):
                    k += 1
            else:
                while dec.advance_frame(lambda fr: check.append(tuple(p.copy() for p in fr))):
                    k += 1
            dt = time.perf_counter() - t0
        assert k == n
        return dt

    got = []
    one_pass(got)
    times = [one_pass() for _ in range(5)]
    # the oracle's Decoder on the same bytes, one thread (configs[0]: "single thread")
    odec = pfvo.Decoder(data, nthreads=1)
    t0 = time.perf_counter()
    want = []
    while True:
        more, fr = odec.advance_frame()
        if fr is not None:
            want.append(fr)
        if not more:
            break
    t_cpu = time.perf_counter() - t0
    odec.close()
    ok = len(want) == len(got) == n and all(np.array_equal(a, b) for fa, fb in zip(got, want) for a, b in zip(fa, fb))
    # the drop-in call with PAGEABLE buffers, one frame per call (INTEGRATION.md section 4 hands over plain Vec<i16>s):
    # pfv_decode_submit + pfv_sync per frame on 1080p key frames
    return {"value": n / min(times), "unit": "frames/s", "frames": n, "width": w, "height": h, "quality": 2, "key_every": key,
            "stream_bytes": len(data), "host_threads": nthreads, "verified": bool(ok),
            "spread": [n / max(times), n / min(times)],
            "note": "Decoder.advance_frame, frame by frame (what src/lib.rs:310-335 times); every picture compared with the oracle Decoder's",
            "cpu_baseline": {"value": n / t_cpu, "unit": "frames/s", "cores": 1, "kind": "port",
                             "sample": f"{n} frames of the same stream through the oracle Decoder (entropy + MB loops) on 1 thread in {t_cpu:.2f} s"}}


def run_pageable_drop_in(torch, dist, stream):
    """The reference's call sites hand over pageable Vec<i16> / Vec<u8> (src/dec.rs:258,376): pfv_decode_submit + pfv_sync per
    frame with PAGEABLE numpy buffers, one 1080p key frame per call, picture copied back into pageable planes."""
    from pretty_fast_video_b200 import PFV_FRAME_I, Engine, make_qtables
    from pretty_fast_video_b200.engine import DecodeJob
    from pretty_fast_video_b200.synth import SynthVideo
    pfvo = _oracle()
    w, h, n = 1920, 1080, 24
    qt, _ = make_qtables(5)
    og = pfvo.geometry_for(w, h)
    sv = SynthVideo(w, h, 0x50465601)
    coeffs = []
    for i in range(4):
        prev = pfvo.frame_init(og)
        coeffs.append(pfvo.encode_iframe_coeffs(og, qt, *sv.frame(i), prev, min(16, os.cpu_count() or 1)))
    y = np.empty((h, w), np.uint8); u = np.empty((h // 2, w // 2), np.uint8); v = np.empty((h // 2, w // 2), np.uint8)
    with Engine(w, h, qt, nslots=2, max_jobs=1, device=torch.cuda.current_device(), stream=stream.cuda_stream) as e:
        def one(i):
            e.decode_submit([DecodeJob(PFV_FRAME_I, i & 1, coeffs[i % 4], (0, 1, 1), out=(y, u, v))])
            e.sync()
        for i in range(4):
            one(i)
        t0 = time.perf_counter()
        for i in range(n):
            one(i)
        dt = time.perf_counter() - t0
        want = pfvo.frame_init(og)
        pfvo.decode_iframe_coeffs(og, qt, (0, 1, 1), coeffs[(n - 1) % 4], want)
        wy, wu, wv = pfvo.crop_frame(og, want)
        ok = np.array_equal(y, wy) and np.array_equal(u, wu) and np.array_equal(v, wv)
    return {"value": n / dt, "unit": "frames/s", "verified": bool(ok), "ms_per_frame": 1e3 * dt / n,
            "h2d_bytes_per_frame": int(og.nb * 512), "d2h_bytes_per_frame": int(w * h * 3 // 2),
            "note": "pfv_decode_submit + pfv_sync per 1080p key frame with pageable host buffers (dense seam): the unmodified "
                    "call-site patch of INTEGRATION.md; pinned buffers and batched submits are what the e2e legs measure"}


def run_format_steps(torch, dist, stream, steps):
    """the colour/format kernels next to the path, device resident: 64 decoded 1080p slots -> packed RGB
    (save_frame, src/lib.rs:365-395) through pfv_slot_convert_rgb, one launch per picture."""
    from pretty_fast_video_b200 import Engine, make_qtables
    w, h, n = 1920, 1080, 64
    qt, _ = make_qtables(5)
    dev = torch.device("cuda", torch.cuda.current_device())
    eng = Engine(w, h, qt, nslots=n, max_jobs=1, device=torch.cuda.current_device(), stream=stream.cuda_stream)
    fb = eng.geometry.frame_bytes
    rng = np.random.default_rng(3)
    for sl in range(n):
        eng.slot_write(sl, rng.integers(0, 256, fb, dtype=np.uint8))
    out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)

    def step():
        for sl in range(n):
            eng.slot_convert_rgb(sl, out[sl].data_ptr())

    ms, _ = time_steps(torch, dist, stream, step, eng.sync, steps, 3)
    slots = np.arange(n, dtype=np.uint32)

    def step_batch():
        eng.slots_convert_rgb(slots, out.data_ptr(), w * h * 3)

    ms_b, _ = time_steps(torch, dist, stream, step_batch, eng.sync, steps, 3)
    # both entry points give the same bytes (the single-picture one is compared with the oracle in tests/test_format_helpers.py)
    ref = out[n - 1].clone()
    eng.slot_convert_rgb(n - 1, ref.data_ptr())
    eng.sync()
    same = bool(torch.equal(ref, out[n - 1]))
    eng.close()
    alg = n * (w * h + 2 * (w // 2) * (h // 2) + w * h * 3)       # planes read once, RGB written once
    peak, _ = peaks()
    return {"value": n / (ms_b * 1e-3), "unit": "frames/s", "ms_per_step": ms_b, "launches_per_step": 1,
            "achieved_gbs": alg / (ms_b * 1e-3) / 1e9, "roofline_frac": alg / (ms_b * 1e-3) / 1e9 / peak, "verified": same,
            "note": "pfv_slots_convert_rgb: 64 decoded 1080p slots -> packed RGB in one launch",
            "one_launch_per_picture": {"value": n / (ms * 1e-3), "unit": "frames/s", "launches_per_step": n,
                                       "achieved_gbs": alg / (ms * 1e-3) / 1e9, "note": "pfv_slot_convert_rgb per picture (launch bound)"}}


# ----------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# ----------------------------------------------------------------------------------------------------
def cpu_port_leg(workload, budget_s, nthreads, reps=None):
    """Times the oracle's macroblock loops (dense coefficients <-> planes, no entropy coding: the same seam the
    GPU legs time) on this box's host cores.  Returns frames/s and a description of the sample."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import pfvo
    from pretty_fast_video_b200.synth import SynthVideo
    cfg = WORKLOADS[workload]
    w, h = cfg["w"], cfg["h"]
    og = pfvo.geometry_for(w, h)
    qt, px_err = pfvo.make_qtables(cfg["quality"])
    gop = cfg["gop"]
    nsrc = 4 if gop == 1 else min(gop, 6)
    sv = SynthVideo(w, h, cfg["seed"], kind=cfg.get("kind", "moving"))
    frames = [sv.frame(t) for t in range(nsrc)]
    # seam data from the oracle's own encoder (this is the CPU arm; the GPU arm makes its input on the GPU)
    prev = pfvo.frame_init(og)
    seam = []
    for t, (y, u, v) in enumerate(frames):
        if gop == 1 or t == 0:
            c = pfvo.encode_iframe_coeffs(og, qt, y, u, v, prev, nthreads)
            seam.append((1, None, c))
        else:
            hd, c = pfvo.encode_pframe_coeffs(og, qt, px_err, y, u, v, prev, nthreads)
            seam.append((2, hd, c))
    done, t_used = 0, 0.0
    state = pfvo.frame_init(og)
    while t_used < budget_s and (reps is None or done < reps * len(seam)):
        if workload.startswith("encode"):
            prev = pfvo.frame_init(og)
        t0 = time.perf_counter()
        for t, (kind, hd, c) in enumerate(seam):
            if workload.startswith("encode"):
                y, u, v = frames[t]
                if kind == 1:
                    pfvo.encode_iframe_coeffs(og, qt, y, u, v, prev, nthreads)
                else:
                    pfvo.encode_pframe_coeffs(og, qt, px_err, y, u, v, prev, nthreads)
            elif kind == 1:
                pfvo.decode_iframe_coeffs(og, qt, (0, 1, 1), c, state, nthreads)
            else:
                pfvo.decode_pframe_coeffs(og, qt, (2, 3, 3), hd, c, state, nthreads)
        t_used += time.perf_counter() - t0
        done += len(seam)
    fps = done / t_used
    sample = (f"{done} frames ({len(seam)} distinct {w}x{h} frames"
              f"{', 1 key + ' + str(len(seam) - 1) + ' P' if gop > 1 else ', all key'}; oracle MB loops only, "
              f"dense coefficients <-> planes) in {t_used:.1f} s on {nthreads} OpenMP threads")
    return fps, sample, og.nb


def metric_name(workload):
    return "1080p decode frames/sec" if workload.startswith("decode") and "1080p" in workload else f"{workload} frames/sec"


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nthreads = os.cpu_count() or 1
    cfg = WORKLOADS[args.workload]
    t_all = time.perf_counter()
    vals = []
    # one step = the same number of frames the GPU arm decodes per step (a bounded sample: the distinct frames
    # are cycled), capped so that the whole run stays within a few minutes
    per_step = cfg["frames"] * PASSES if cfg["gop"] == 1 else cfg["gops"] * cfg["gop"]
    distinct = 4 if cfg["gop"] == 1 else min(cfg["gop"], 6)
    reps = max(1, min(per_step // distinct, 64))
    for i in range(args.warmup + args.steps):
        fps, sample, nb = cpu_port_leg(args.workload, budget_s=1e9, nthreads=nthreads, reps=reps)
        if i >= args.warmup:
            vals.append(fps)
        if time.perf_counter() - t_all > 240:
            break
    v = float(np.mean(vals)) if vals else fps
    line = {
        "impl": "reference", "metric": metric_name(args.workload),
        "value": v, "unit": "frames/s", "mb_per_s": v * nb, "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
        "ms_per_step": 1e3 * reps * distinct / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i32", "data": "synthetic",
        "config": config_of(args.workload),
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": nthreads, "kind": "port", "sample": sample,
                         "frames_per_step": reps * distinct},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C restatement of the reference algorithm (oracle/); the Rust crate cannot be built here (no cargo/rustc)",
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="decode_i_1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--extras", type=int, default=1, help="also run the other workloads (N > 1: the resident P-stream legs only)")
    ap.add_argument("--cpu-budget", type=float, default=10.0)
    ap.add_argument("--e2e", type=int, default=1, help="0 skips the host-buffer leg (kernel tuning runs only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    from pretty_fast_video_b200 import lib
    lib()                                                   # fail loudly if the CUDA library is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the engine has no CPU fallback")
    dist = Dist(args.gpus)
    torch.cuda.set_device(dist.local)
    stream = torch.cuda.Stream()
    peak, peak_src = peaks()

    def run(workload, steps, warmup, do_e2e=True):
        cfg = WORKLOADS[workload]
        st = Streams(torch, cfg, dist.rank, stream)
        fn = run_encode if workload.startswith("encode") else run_decode
        r = fn(torch, dist, stream, st, steps, warmup, do_e2e)
        r["st"] = st
        return r

    binding = bind_to_gpu_numa_node(torch, dist.local)
    sampler = ClockSampler(dist.local)
    sampler.start()
    r = run(args.workload, args.steps, args.warmup, bool(args.e2e))
    clocks = sampler.stop()
    st, cfg = r["st"], WORKLOADS[args.workload]
    fps = r["frames"] * dist.world / (r["max_ms"] * 1e-3)
    launches = r["launches_per_step"]
    achieved = r["alg_bytes"] / (r["my_ms"] * 1e-3) / 1e9      # this rank's kernels
    kernel_of = {"decode_i_1080p": "decode_i_stream_kernel", "decode_i_1080p_dense": "decode_i_direct_kernel",   # PFV_JOB_DENSE
                 "decode_p_1080p": "decode_p_fused_kernel", "decode_p_4k": "decode_p_fused_kernel",
                 "decode_p_1080p_64": "decode_p_fused_kernel",
                 "encode_p_1080p": "encode_p2_kernel", "encode_i_1080p": "encode_i_persist_kernel"}
    kernel_key = kernel_of[args.workload]

    def stats_of(xs):
        return {"mb_per_frame": xs.nb, "coded_mb_fraction_p": xs.coded_frac, "nonzero_mv_fraction_p": xs.mv_nonzero,
                "ac_subblock_fraction": xs.general_frac,
                "distinct_sequences": 1 if xs.gop == 1 else min(xs.lanes, DISTINCT_SEQUENCES)}

    verified_all = bool(r["verified"]) and all(bool(x.get("verified", True)) for x in ((r["e2e"] or {}), (r["e2e"] or {}).get("dense", {})))
    line = {
        "metric": metric_name(args.workload),
        "value": fps, "unit": "frames/s", "mb_per_s": fps * st.nb,
        "n_gpus": dist.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["max_ms"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
        "config": config_of(args.workload),
        "workload_stats": stats_of(st),
        "verified": verified_all,
        "gpu_launches": launches * args.steps,
        "e2e": {k: v for k, v in (r["e2e"] or {}).items()},
        "roofline": {"bound": "hbm", "kernel": kernel_key, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src, "alg_bytes_per_step": r["alg_bytes"],
                     "launches_per_step": launches, "traffic": ncu_traffic(kernel_key + ":" + args.workload) or ncu_traffic(kernel_key),
                     "traffic_source": "profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel on this workload, per launch)"},
        "clocks": clocks,
        "host_binding": binding,
    }
    if args.workload.startswith("encode_p"):
        line["candidate_pixels_per_s"] = fps * st.nb * SEARCH_CANDIDATES_PER_MB * 256

    nthreads = os.cpu_count() or 1
    if dist.world == 1 and args.cpu_budget > 0:
        cfps, sample, _ = cpu_port_leg(args.workload, args.cpu_budget, nthreads)
        line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": nthreads, "kind": "port", "sample": sample}
        if nthreads > 1 and args.cpu_budget >= 1.0:
            # the same port on ONE thread (SURVEY 8d quotes the CPU path at 1 thread and at all threads)
            c1, s1, _ = cpu_port_leg(args.workload, min(args.cpu_budget / 4, 3.0), 1)
            line["cpu_baseline"]["one_thread"] = {"value": c1, "unit": "frames/s", "cores": 1, "sample": s1}
    if args.extras:
        extras = {}
        # N > 1 (the scaling run): the P-stream configs ride along, resident legs only, so that BASELINE.json configs[2]
        # and configs[4] (GOPs sharded over the ranks) sit on the driver's record at every N
        names = ("decode_i_1080p_dense", "decode_p_1080p", "decode_p_1080p_64", "encode_p_1080p", "encode_i_1080p", "decode_p_4k") if dist.world == 1 \
            else ("decode_p_1080p", "decode_p_4k")
        for wl in names:
            if wl == args.workload:
                continue
            try:
                x = run(wl, max(3, args.steps // 4), 3, do_e2e=(dist.world == 1 and wl not in ("decode_p_4k", "decode_p_1080p_64")))
                xs = x["st"]
                xfps = x["frames"] * dist.world / (x["max_ms"] * 1e-3)
                extras[wl] = {
                    "value": xfps, "unit": "frames/s", "mb_per_s": xfps * xs.nb, "ms_per_step": x["max_ms"], "n_gpus": dist.world,
                    "frames_per_step_per_gpu": x["frames"], "launches_per_step": x["launches_per_step"],
                    "workload_stats": stats_of(xs), "kernel": kernel_of[wl], "verified": bool(x["verified"]),
                    "roofline_frac": x["alg_bytes"] / (x["my_ms"] * 1e-3) / 1e9 / peak,
                    "achieved_gbs": x["alg_bytes"] / (x["my_ms"] * 1e-3) / 1e9,
                    "traffic": ncu_traffic(kernel_of[wl] + ":" + wl) or ncu_traffic(kernel_of[wl]),
                    "e2e": x["e2e"],
                }
                if wl.startswith("encode_p"):
                    extras[wl]["candidate_pixels_per_s"] = xfps * xs.nb * SEARCH_CANDIDATES_PER_MB * 256
                if dist.world == 1 and args.cpu_budget > 0 and wl != "decode_p_1080p_64":     # (the same stream as decode_p_1080p)
                    cf, cs, _ = cpu_port_leg(wl, min(args.cpu_budget, 3.0 if wl == "decode_p_4k" else 6.0), nthreads)
                    extras[wl]["cpu_baseline"] = {"value": cf, "unit": "frames/s", "cores": nthreads, "kind": "port", "sample": cs}
                verified_all &= bool(x["verified"]) and all(bool(y.get("verified", True)) for y in ((x["e2e"] or {}), (x["e2e"] or {}).get("dense", {})))
                del x, xs
            except Exception as ex:                      # an extra must never lose the headline line
                extras[wl] = {"error": repr(ex)}
                verified_all = False
        if dist.world == 1:
            try:
                extras["rgb_out_1080p"] = run_format_steps(torch, dist, stream, max(3, args.steps // 4))
            except Exception as ex:
                extras["rgb_out_1080p"] = {"error": repr(ex)}
            try:
                extras["decoder_stream_1080p"] = run_decoder_stream(torch, dist, min(args.cpu_budget, 6.0), nthreads)
            except Exception as ex:
                extras["decoder_stream_1080p"] = {"error": repr(ex)}
            try:
                extras["config1_stream_512x384"] = run_config1_stream(torch, dist, min(args.cpu_budget, 6.0), nthreads)
                verified_all &= bool(extras["config1_stream_512x384"]["verified"])
            except Exception as ex:
                extras["config1_stream_512x384"] = {"error": repr(ex)}
            try:
                extras["drop_in_pageable_1080p"] = run_pageable_drop_in(torch, dist, stream)
                verified_all &= bool(extras["drop_in_pageable_1080p"]["verified"])
            except Exception as ex:
                extras["drop_in_pageable_1080p"] = {"error": repr(ex)}
        line["extras"] = extras
        line["verified"] = verified_all
    line["verified"] = dist.sum(1.0 if line["verified"] else 0.0) == float(dist.world)     # every rank's legs
    if dist.rank == 0:
        print(json.dumps(line))
    dist.close()
    if not line["verified"]:
        raise SystemExit("bench.py: a timed leg's output differs from the oracle (see the \"verified\" keys)")


if __name__ == "__main__":
    main()
