#!/usr/bin/env python3
"""bench.py -- registrations/sec on synthetic 64-ring scan pairs (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle restatement) on host cores

A "step" is one pass of the whole hot path (voxel -> FPFH -> match -> TIM graph -> max clique ->
GNC-TLS + COTE) over one batch of `--pairs` synthetic 64-ring pairs per GPU (BASELINE configs[2]: 256
pairs on 1 GPU; 8 ranks x 256 = configs[3]'s 2048 pairs).  Pairs are independent, so ranks shard them
(weak scaling) and the only collective is one NCCL all_gather of the fixed-size result records.

  value : whole-job registrations/s with the raw scans already resident in HBM (device pointers through
          qb200_register_batch), CUDA events on the launching stream, max over ranks.
  e2e   : the same call with pinned HOST buffers -- H2D of every scan and D2H of the result records are
          inside the timed region.
  roofline      : the dominant kernel (tc_nn_kernel, the N_src x N_tgt x 33 contraction on tcgen05) from CUDA events
                  recorded around it inside the timed steps.
  cpu_baseline  : the CPU oracle (restatement of the reference path; the reference binary itself cannot be
                  built here) timed on this box's host cores on a bounded sample of the same pairs.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "registrations/sec (64-ring pair)"
UNIT = "registrations/s"


def _host_threads() -> int:
    """Threads for the CPU arm: usable cores (affinity, cgroup quota), one per physical core (SMT siblings only
    oversubscribe the OpenMP loops)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        quota, period = Path("/sys/fs/cgroup/cpu.max").read_text().split()
        if quota != "max":
            n = min(n, max(1, int(quota) // int(period)))
    except (OSError, ValueError):
        pass
    try:
        sib = Path("/sys/devices/system/cpu/cpu0/topology/thread_siblings_list").read_text().strip()
        smt = len(sib.replace("-", ",").split(",")) if sib else 1
        if "-" in sib:
            a, b = sib.split("-")[:2]
            smt = int(b) - int(a) + 1
        if smt > 1 and n >= (os.cpu_count() or n):
            n = max(1, n // smt)
    except (OSError, ValueError):
        pass
    return n


def load_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def gen_pairs(seeds):
    from quatro_b200 import synth
    synth._lib()  # build/load once before the threads start
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 8)) as ex:
        return list(ex.map(lambda s: synth.outdoor_pair(int(s))[:2], seeds))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank, world):
    """The reference's CPU path (oracle restatement) on the host cores; rank 0 only."""
    if rank != 0:
        return
    from oracle import Oracle
    from quatro_b200.capi import default_params
    o = Oracle()
    cores = o.set_num_threads(_host_threads())   # torchrun exports OMP_NUM_THREADS=1: ask for every usable host core explicitly
    p = default_params()
    per_step = args.ref_pairs_per_step
    pairs = gen_pairs(range(per_step))
    for _ in range(max(args.warmup, 1)):
        o.register_pair(pairs[0][0], pairs[0][1], p)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        for s, t in pairs:
            o.register_pair(s, t, p)
            done += 1
    dt = time.perf_counter() - t0
    val = done / dt
    sample = f"{per_step} pairs per step x {args.steps} steps, sequential pairs, OpenMP inside each stage"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64",
        "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference path (oracle/); the PCL/FLANN/pmc binary cannot be built in this image",
    }))


def workload_config(args, world):
    return {"workload": f"batch of {args.pairs} synthetic 64-ring pairs per GPU (BASELINE configs[2]; {world}x{args.pairs} global, 8 GPUs = configs[3])",
            "pairs_per_gpu": args.pairs, "global_pairs": args.pairs * world, "scan": "64 rings x 1800 azimuths, ~111k returns, ground flagged",
            "params": "config/params.yaml defaults (voxel 0.3, normal_r 0.5, fpfh_r 0.75, noise_bound 0.3, PMC_HEU, median COTE)",
            "l2": "inputs larger than L2 (~0.9 GB of raw scans per GPU per step vs 126 MB)",
            "parallelism": f"dp{world}: independent pairs sharded across ranks, one NCCL all_gather of result records per step"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=256, help="pairs per GPU per step")
    ap.add_argument("--graph-L", type=int, default=3000, help="correspondences per set of the K8 roofline pass (0 = skip)")
    ap.add_argument("--slots", type=int, default=64, help="pairs per device wave (two lanes of this size alternate)")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--ref-pairs-per-step", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from quatro_b200 import synth
    from quatro_b200.capi import Handle, Pair, default_params, RESULT_DTYPE, MEM_HOST, MEM_DEVICE

    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: whatever NCCL logs (its version banner under NCCL_DEBUG=VERSION) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    p = default_params()
    P = args.pairs

    # ---- synthetic inputs: pinned host copy (e2e) and device-resident copy (value) ----
    pairs = gen_pairs(range(rank * P, rank * P + P))
    total_pts = sum(len(s) + len(t) for s, t in pairs)
    host = torch.empty((total_pts, 4), dtype=torch.float32).pin_memory()
    hv = host.numpy()
    offs, o = [], 0
    for s, t in pairs:
        hv[o:o + len(s)] = s; offs.append((o, len(s))); o += len(s)
        hv[o:o + len(t)] = t; offs.append((o, len(t))); o += len(t)
    dvc = host.to(dev, non_blocking=False)
    pa_host, pa_dev = (Pair * P)(), (Pair * P)()
    for i in range(P):
        (so, sn), (to, tn) = offs[2 * i], offs[2 * i + 1]
        pa_host[i].src, pa_host[i].n_src, pa_host[i].tgt, pa_host[i].n_tgt = host.data_ptr() + so * 16, sn, host.data_ptr() + to * 16, tn
        pa_dev[i].src, pa_dev[i].n_src, pa_dev[i].tgt, pa_dev[i].n_tgt = dvc.data_ptr() + so * 16, sn, dvc.data_ptr() + to * 16, tn
    h2d_bytes = total_pts * 16
    d2h_bytes = P * RESULT_DTYPE.itemsize

    handle = Handle(device=local_rank, max_batch_slots=min(args.slots, P))
    stream = torch.cuda.current_stream(dev)
    handle.set_stream(stream.cuda_stream)
    out = np.zeros(P, RESULT_DTYPE)
    out_t = torch.zeros((P, RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    gathered = [torch.empty_like(out_t) for _ in range(world)] if world > 1 else None

    def step(pa, kind):
        handle.register_batch_raw(pa, P, p, kind, out)
        if world > 1:  # the path's only collective: gather the per-pair result records
            out_t.copy_(torch.from_numpy(out.view(np.uint8).reshape(P, -1)), non_blocking=True)
            dist.all_gather(gathered, out_t)

    def timed(pa, kind, steps, sampler=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = handle.launch_count()
        kms, kcalls, sms = np.zeros(2), np.zeros(2), np.zeros(8)
        e0.record(stream)
        for _ in range(steps):
            step(pa, kind)
            m, c = handle.kernel_ms()
            kms += m; kcalls += c; sms += handle.stage_ms()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), handle.launch_count() - l0, kms, kcalls, sms, clocks

    for _ in range(args.warmup):
        step(pa_dev, MEM_DEVICE)
    dev_ms, launches, kms, kcalls, sms, clocks = timed(pa_dev, MEM_DEVICE, args.steps, ClockSampler(local_rank) if rank == 0 else None)
    res_dev = out.copy()
    for _ in range(max(1, args.warmup // 2)):
        step(pa_host, MEM_HOST)
    e2e_ms, _, _, _, _, _ = timed(pa_host, MEM_HOST, args.steps)
    assert out.tobytes() == res_dev.tobytes(), "host-buffer and device-buffer runs disagree"

    value = world * P * args.steps / (dev_ms * 1e-3)
    e2e_value = world * P * args.steps / (e2e_ms * 1e-3)

    # BASELINE configs[1] (one pair on one GPU): latency of a single registration through the same call, host buffers
    single_ms = None
    if rank == 0:
        lat = []
        for i in range(12):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            handle.register_batch_raw(pa_host, 1, p, MEM_HOST, out)
            a1.record(stream)
            torch.cuda.synchronize(dev)
            if i >= 2:
                lat.append(a0.elapsed_time(a1))
        single_ms = float(np.median(lat))

    if rank == 0:
        peaks = load_peaks()
        nA, nB, L = res_dev["n_src_vox"].astype(np.float64), res_dev["n_tgt_vox"].astype(np.float64), res_dev["n_corr"].astype(np.float64)
        # K6: 66 flop per (src,tgt) descriptor pair (the 2*33 of the ||a||^2+||b||^2-2ab contraction)
        match_flops_step = float((66.0 * nA * nB).sum())
        match_ms_launch = kms[0] / max(kcalls[0], 1)
        launches_per_step = kcalls[0] / args.steps
        match_tflops = match_flops_step / launches_per_step / (match_ms_launch * 1e-3) / 1e12 if match_ms_launch > 0 else 0.0
        # K8: algorithmic bytes = 2*L*16 (matched points) + L*ceil(L/32)*4 (bit adjacency) + 4L (degrees)
        graph_bytes_step = float((2 * L * 16 + L * np.ceil(L / 32) * 4 + 4 * L).sum())
        graph_ms_launch = kms[1] / max(kcalls[1], 1)
        graph_gbs = graph_bytes_step / (kcalls[1] / args.steps) / (graph_ms_launch * 1e-3) / 1e9 if graph_ms_launch > 0 else 0.0
        step_ms = dev_ms / args.steps
        roofline = {"kernel": "tc_nn_kernel (K6: tcgen05 3xTF32 filter of the N_src x N_tgt x 33 distance matrix + in-kernel exact fp32 evaluation)",
                    "bound": "tensor", "achieved": match_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": match_tflops / peaks["bf16_tflops"],
                    "traffic": 3.978e8 * (match_flops_step / launches_per_step) / 2.05e11,  # ncu: 397.8 MB for a 64-pair launch (profiles/r01_ncu_summary.md), scaled by the launch's work
                    "traffic_unit": "bytes per launch (dram read + write, ncu --set full capture of the same launch geometry)",
                    "peak_source": peaks["source"] + ", burst bf16",
                    "launch_ms": match_ms_launch, "launches_per_step": launches_per_step, "share_of_step": float(kms[0] / args.steps / step_ms),
                    "flops_per_launch": match_flops_step / launches_per_step,
                    "note": "achieved = 66 flop per descriptor pair (algorithmic, all n_src x n_tgt pairs) / launch time measured with CUDA events inside the timed steps (other lanes' kernels share the SMs meanwhile); the kernel skips ~72% of the tiles by a norm lower bound and runs 3xTF32 on K padded to 40 on the rest; ncu: tensor pipe 15%, issue 43%, bound by the exact evaluation of the survivors (DESIGN.md 5.1)"}
        roofline_graph = {"kernel": "tim_graph_kernel (K8)", "bound": "hbm", "achieved": graph_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                          "frac": graph_gbs / peaks["hbm_gbs"], "traffic": None, "launch_ms": graph_ms_launch,
                          "bytes_per_launch": graph_bytes_step / max(kcalls[1] / args.steps, 1), "mean_L": float(L.mean()),
                          "note": "fp32-pipe bound at algorithmic-minimum bytes (SURVEY.md 8d): ~30 instr per pair test vs 0.27 B per pair"}
        # K8 at the size BASELINE's configs name (~3k correspondences per pair): 32 precomputed correspondence sets through
        # qb200_solve_batch (device-resident), tim_graph_kernel timed with CUDA events inside the call
        roofline_graph_3k = None
        if args.graph_L > 0:
            gsets, keep = [], []
            for i in range(32):
                a4, b4, _, _ = synth.matched_pairs(7000 + i, args.graph_L, inlier_ratio=0.03, noise=0.04)
                ta, tb = torch.from_numpy(np.ascontiguousarray(a4)).to(dev), torch.from_numpy(np.ascontiguousarray(b4)).to(dev)
                keep.append((ta, tb))
                gsets.append((ta.data_ptr(), tb.data_ptr(), len(a4)))
            torch.cuda.synchronize(dev)
            handle.solve_batch(gsets, p, kind=MEM_DEVICE)  # warm-up
            gms, gcalls = 0.0, 0
            for _ in range(5):
                rg = handle.solve_batch(gsets, p, kind=MEM_DEVICE)
                m, c = handle.kernel_ms()
                gms += float(m[1]); gcalls += int(c[1])
            Lg = float(args.graph_L)
            g_bytes = 32 * (2 * Lg * 16 + Lg * np.ceil(Lg / 32) * 4 + 4 * Lg)
            g_ms = gms / max(gcalls, 1)
            g_pairs = 32 * Lg * (Lg - 1) / 2
            roofline_graph_3k = {"kernel": "tim_graph_kernel (K8)", "bound": "hbm", "achieved": g_bytes / (g_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                                 "unit": "GB/s", "frac": g_bytes / (g_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": None, "launch_ms": g_ms,
                                 "L": int(Lg), "sets_per_launch": 32, "bytes_per_launch": g_bytes, "pair_tests_per_s": g_pairs / (g_ms * 1e-3),
                                 "valid_sets": int(rg["valid"].sum()),
                                 "note": "algorithmic-minimum bytes (0.27 B per pair test) against ~30 fp32 instructions per pair test: the kernel is bound by the fp32 pipe, pair_tests_per_s is the meaningful rate"}
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import Oracle
            o = Oracle()
            cores = o.set_num_threads(_host_threads())
            o.register_pair(pairs[0][0], pairs[0][1], p)  # warm-up
            t0 = time.perf_counter(); n = 0
            checked = 0
            while n < min(P, 64) and time.perf_counter() - t0 < args.cpu_baseline_seconds:
                r, _ = o.register_pair(pairs[n][0], pairs[n][1], p)
                g = res_dev[n]
                assert (r.n_corr, r.clique_size, r.n_edges) == (g["n_corr"], g["clique_size"], g["n_edges"]), f"pair {n}: GPU result differs from the CPU oracle"
                assert np.allclose(np.asarray(g["T"]), np.array(r.T[:]), atol=1e-9)
                checked += 1; n += 1
            dt = time.perf_counter() - t0
            cpu = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {n} pairs of rank 0's batch, sequential pairs, OpenMP({cores}) inside stages; all {checked} matched the GPU records"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (front end, match, graph filter) / f64 (graph boundary, GNC, COTE)",
            "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_graph": roofline_graph,
            "roofline_graph_3k": roofline_graph_3k, "cpu_baseline": cpu,
            "stages_ms_per_step": {k: float(v / args.steps) for k, v in zip(["h2d", "voxel", "fpfh", "match", "graph", "clique", "pose", "d2h"], sms)},
            "valid_pairs": int(res_dev["valid"].sum()), "mean_n_vox": float((nA.mean() + nB.mean()) / 2), "mean_L": float(L.mean()),
            "mean_clique": float(res_dev["clique_size"].mean()),
            "single_pair_latency_ms": single_ms,
        }
        print(json.dumps(line))
    handle.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
