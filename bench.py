#!/usr/bin/env python
"""Headline benchmark: 1-s clips/sec through features (mel+MFCC+delta) + CoughDetectorResidual.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch of 16 384 synthetic 1-s clips per GPU
(BASELINE.json configs[2]); clips shard across ranks with no data-path collective (weak scaling),
only the final (B, 2) logit gather.  Prints ONE JSON line on rank 0 (contract in the task brief):
value = device-resident throughput, e2e = the same through cd_e2e_forward_host with pinned HOST
buffers (H2D of the audio and D2H of the logits inside the timed region), roofline = the feature
kernel against the measured HBM peak, cpu_baseline = the oracle port on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 16384
AUDIO_BYTES = 64000           # 16000 fp32 samples
FEAT_BYTES = 36360            # 90 x 101 fp32
CNN_FLOPS = 42_865_600        # 2 x 21 432 800 MAC per clip
NCU_TRAFFIC_PER_CLIP = 100_154 # feat_kernel dram__bytes_read+write per clip, one ncu --set full capture (profiles/)
METRIC = "1-s clips/sec (mel+MFCC+CNN)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (same torchaudio/torch calls as the reference's per-clip loop)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, n = args
    import torch
    torch.set_num_threads(1)
    from cough_detector_b200 import synth
    from oracle import reference_port as rp
    torch.manual_seed(0)
    model = rp.PortResidual().eval()
    pp = rp.PortPreprocessor()
    clips = synth.synth_batch(n, seed=seed, device="cpu")
    t0 = time.perf_counter()
    out = rp.run_clips(pp, model, clips)
    dt = time.perf_counter() - t0
    return n, dt, float(out.sum())


def cpu_throughput(clips_per_worker: int, workers: int, seed0: int = 0):
    """clips/s of the port with `workers` single-threaded processes (intra-op threads hurt the per-clip path)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        pool.map(_cpu_worker, [(seed0 + 1000 + i, 2) for i in range(workers)])       # import + warm-up, untimed
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(seed0 + i, clips_per_worker) for i in range(workers)])
        wall = time.perf_counter() - t0
    n = sum(r[0] for r in res)
    return n / wall, n, wall, max(r[1] for r in res)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # pragma: no cover
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(host_cores(), 64)
    per_worker = 512
    vals = []
    for s in range(args.warmup + args.steps):
        v, n, wall, _ = cpu_throughput(per_worker, cores, seed0=s * 100)
        if s >= args.warmup:
            vals.append((v, n, wall))
    n_tot = sum(v[1] for v in vals)
    t_tot = sum(v[2] for v in vals)
    value = n_tot / t_tot
    sample = f"{cores} procs x {per_worker} clips per step, {args.steps} steps (of the {BATCH}-clip workload)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(1, len(vals)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "e2e features + CoughDetectorResidual, per-clip loop, CPU", "batch": BATCH},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler (profiling recipe's nvidia-smi line)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.t = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._pump, daemon=True)
        self.t.start()

    def _pump(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def stop(self, t0: float, t1: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, ln in self.rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    cpu = None
    if world == 1 and not args.no_cpu_baseline:       # before CUDA is touched; bounded to ~10-20 s
        cores = min(host_cores(), 64)
        v, n, wall, _ = cpu_throughput(768, cores)
        cpu = {"value": v, "unit": "clips/s", "cores": cores, "kind": "port",
               "sample": f"{n} clips of the same synthetic workload, {cores} single-thread processes, {wall:.1f} s"}

    import torch
    import torch.distributed as dist
    from cough_detector_b200 import _lib, synth, create_model
    from cough_detector_b200.model import _PREC, DEFAULT_PRECISION

    torch.cuda.set_device(local_rank)
    from cough_detector_b200.sharding import bind_to_gpu_numa
    numa_bound = bind_to_gpu_numa(local_rank)         # pinned host buffers on the GPU's own socket (e2e path is PCIe-bound)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("CD_BENCH_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    precision = args.precision or DEFAULT_PRECISION
    prec = _PREC[precision]
    h = _lib.get_handle(local_rank)
    torch.manual_seed(0)
    model = create_model("residual", n_mels=90, num_classes=2, in_channels=1).eval()
    sd = synth.randomized_state_dict(model, seed=1)
    h.load_state_dict(sd, key="bench")

    B = args.batch
    audio = synth.synth_batch(B, seed=100 + rank, device=f"cuda:{local_rank}")      # 1.05 GB > 126 MB L2
    gathered = torch.empty(world * B, 2, device=audio.device) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        logits = h.e2e_forward(audio, prec)
        if world > 1:
            dist.all_gather_into_tensor(gathered, logits)       # the only collective: final logit gather
        return logits

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = h.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        logits = step()
    ev1.record()
    barrier()
    w1 = time.perf_counter()
    launches = h.launches - l0
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=audio.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- per-stage timing for the roofline objects ----------------------------------------------------
    # feature kernel alone on BASELINE configs[1]: 65 536 resident clips (4.2 GB in, 2.4 GB out), fp32 features out
    FB = args.feat_batch
    faudio = synth.synth_batch(FB, seed=200 + rank, device=f"cuda:{local_rank}")
    ffeat = torch.empty(FB, 1, 90, 101, device=audio.device)
    for _ in range(2):
        h.lib.cd_feat_forward(h._h, faudio.data_ptr(), 16000, FB, ffeat.data_ptr(), 1, 0.97, h._stream())
    torch.cuda.synchronize()
    fe0, fe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nrep = max(3, min(args.steps, 10))
    fe0.record()
    for _ in range(nrep):
        h.lib.cd_feat_forward(h._h, faudio.data_ptr(), 16000, FB, ffeat.data_ptr(), 1, 0.97, h._stream())
    fe1.record()
    torch.cuda.synchronize()
    feat_ms = fe0.elapsed_time(fe1) / nrep
    del faudio, ffeat
    feat = h.feat_forward(audio)
    torch.cuda.synchronize()
    ce0, ce1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h.cnn_forward(feat, prec)
    torch.cuda.synchronize()
    ce0.record()
    for _ in range(nrep):
        h.cnn_forward(feat, prec)
    ce1.record()
    torch.cuda.synchronize()
    cnn_ms = ce0.elapsed_time(ce1) / nrep
    clocks = sampler.stop(w0, time.perf_counter()) if rank == 0 else None
    del feat

    # ---- end to end through the C ABI with HOST buffers ----------------------------------------------
    host = audio.cpu().pin_memory()
    out = torch.empty(B, 2).pin_memory()
    for _ in range(max(1, min(2, args.warmup))):
        h.e2e_forward_host(host, prec, out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h.e2e_forward_host(host, prec, out)         # synchronous: returns when `out` is complete
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=audio.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * B * args.steps / e2e_s
    same = bool(torch.equal(out, logits.cpu()))
    # what the link itself does: a bare pinned host->device copy of the same buffer, all ranks at once (the e2e
    # path's floor; on a multi-GPU box the ranks share the host's memory and PCIe fabric)
    barrier()
    dev_copy = torch.empty_like(audio)
    dev_copy.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(3):
        dev_copy.copy_(host, non_blocking=True)
    pe1.record()
    torch.cuda.synchronize()
    h2d_gbs = 3 * B * AUDIO_BYTES / (pe0.elapsed_time(pe1) * 1e-3) / 1e9
    del dev_copy

    # additive: the same call fed int16 PCM (the wav sample format) -- half the host->device bytes
    pcm = (host / host.abs().amax(dim=1, keepdim=True).clamp(min=1e-9) * 30000.0).round().to(torch.int16).pin_memory()
    out16 = torch.empty(B, 2).pin_memory()
    h.e2e_forward_host(pcm, prec, out16)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h.e2e_forward_host(pcm, prec, out16)
    torch.cuda.synchronize()
    i16_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([i16_s], device=audio.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        i16_s = float(t.item())
    e2e_i16 = world * B * args.steps / i16_s

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        tc_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        feat_gbs = FB * (AUDIO_BYTES + FEAT_BYTES) / (feat_ms * 1e-3) / 1e9
        cnn_tfs = B * CNN_FLOPS / (cnn_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"e2e features + CoughDetectorResidual {precision} inference, batch {B} clips per GPU "
                                   "(BASELINE configs[2])", "batch_per_gpu": B, "global_batch": world * B,
                       "parallelism": f"clip-sharded x{world}, no data-path collective, final logit all_gather",
                       "l2": "inputs (1.05 GB/GPU) larger than L2, no flush needed", "precision": precision},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": B * AUDIO_BYTES,
                    "d2h_bytes_per_step": B * 8, "bit_identical_to_device_path": same,
                    "h2d_gbs": e2e_value / world * AUDIO_BYTES / 1e9, "h2d_gbs_bare_copy": h2d_gbs,
                    "host_numa_bound": numa_bound,
                    "api": "cd_e2e_forward_host (pinned host audio in, host logits out)"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "feat_kernel", "bound": "hbm", "achieved": feat_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": feat_gbs / hbm_peak, "traffic": FB * NCU_TRAFFIC_PER_CLIP, "peak_source": peak_src,
                         "traffic_source": "ncu --set full of this launch (profiles/r1_ncu_feat_kernel_65536.csv): dram read + write 6.564 GB "
                                           "per 65 536-clip launch = 100 154 B per clip, scaled by the launch size",
                         "algorithmic_bytes_per_clip": AUDIO_BYTES + FEAT_BYTES, "launch_ms": feat_ms,
                         "clips_per_s": FB / (feat_ms * 1e-3), "clips_per_launch": FB,
                         "note": "BASELINE configs[1]: timed alone over 65 536 resident clips right after the e2e region"},
            "roofline_cnn": {"kernel": "cd_cnn_forward: fp32 features -> blocked bf16 map (1 launch) + stem + 2 residual blocks + head", "bound": "tensor", "achieved": cnn_tfs,
                             "peak": tc_peak, "unit": "TFLOP/s", "frac": cnn_tfs / tc_peak, "launch_ms": cnn_ms,
                             "algorithmic_flops_per_clip": CNN_FLOPS, "clips_per_s": B / (cnn_ms * 1e-3)},
            "e2e_int16": {"value": e2e_i16, "unit": "clips/s", "h2d_bytes_per_step": B * 32000, "d2h_bytes_per_step": B * 8,
                          "api": "cd_e2e_forward_host_i16 (pinned int16 PCM in; additive entry point, not in the reference)"},
            "cpu_baseline": cpu,
            "wall_s": w1 - w0,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--feat-batch", type=int, default=65536)
    ap.add_argument("--precision", default=None, choices=[None, "fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        if args.warmup < 3:
            args.warmup = 3                       # timing rule: W >= 3
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
