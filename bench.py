#!/usr/bin/env python
"""Headline benchmark: 1-s clips/sec through features (mel+MFCC+delta) + CoughDetectorResidual.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch of 16 384 synthetic 1-s clips per GPU
(BASELINE.json configs[2]); clips shard across ranks with no data-path collective (weak scaling),
only the final (B, 2) logit gather.  Prints ONE JSON line on rank 0 (contract in the task brief):
value = device-resident throughput, e2e = the same through cd_e2e_forward_host with pinned HOST
buffers (H2D of the audio and D2H of the logits inside the timed region), roofline = the feature
kernel against the measured HBM peak, cpu_baseline = the reference's own bytecode (oracle/_ref, else the oracle port) on the host cores, stream =
BASELINE configs[3] (1 h of audio, 100 ms hop, time-sharded), sweep = configs[4] (4 194 304 clips
generated per chunk on the device, strong scaling).  Timed regions last >= 1 s.
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
# feat_kernel dram__bytes_read + dram__bytes_write per clip: a CONSTANT taken from one ncu --set full capture of this launch
# (profiles/r2_ncu_feat_kernel_65536.csv: 4.252 GB + 2.351 GB per 65 536-clip launch), scaled by the launch size below
NCU_TRAFFIC_PER_CLIP = 100_757
NCU_TRAFFIC_SOURCE = ("ncu --set full of cd_feat_forward over 65 536 clips (profiles/r2_ncu_feat_kernel_65536.csv): dram read 4.252 GB + "
                      "write 2.351 GB = 100 757 B per clip, 1.004 x the algorithmic bytes (a constant in bench.py, scaled by the launch size)")
METRIC = "1-s clips/sec (mel+MFCC+CNN)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference from oracle/_ref when built (oracle/ref_build.py), else the oracle port
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One single-threaded worker of the CPU arm.  With oracle/_ref built (oracle/ref_build.py: the unmodified
    reference's own bytecode) it runs the reference itself, per clip as BASELINE.md §5 says; otherwise the port."""
    seed, n = args
    import torch
    torch.set_num_threads(1)
    from cough_detector_b200 import synth
    from oracle import ref_build
    ref = ref_build.load()
    torch.manual_seed(0)
    clips = synth.synth_batch(n, seed=seed, device="cpu")
    if ref is not None:
        pp = ref["preprocessing"].AudioPreprocessor(use_pcen=False, use_pre_emphasis=False, use_delta_delta=False,
                                                    use_spectral_contrast=False)
        model = ref["model"].create_model("residual", n_mels=90, num_classes=2, in_channels=1).eval()
        t0 = time.perf_counter()
        out = []
        with torch.no_grad():
            for w in clips:
                w = pp.pad_or_trim(pp.normalize(w.reshape(1, -1)))
                out.append(model(pp.extract_features(w).unsqueeze(0))[0])
        out = torch.stack(out)
        dt = time.perf_counter() - t0
        return n, dt, float(out.sum()), "reference"
    from oracle import reference_port as rp
    if n == 2 and seed == 1000:
        print(f"bench.py: CPU arm falls back to the port ({ref_build.last_error})", file=sys.stderr, flush=True)
    model = rp.PortResidual().eval()
    pp = rp.PortPreprocessor()
    t0 = time.perf_counter()
    out = rp.run_clips(pp, model, clips)
    dt = time.perf_counter() - t0
    return n, dt, float(out.sum()), "port"


def cpu_throughput(clips_per_worker: int, workers: int, seed0: int = 0, pool=None):
    """clips/s of the CPU arm with `workers` single-threaded processes (intra-op threads hurt the per-clip path).
    `pool`: an already warmed-up pool to reuse (the reference arm keeps one across its steps)."""
    own = pool is None
    if own:
        pool = cpu_pool(workers)
    try:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(seed0 + i, clips_per_worker) for i in range(workers)], chunksize=1)
        wall = time.perf_counter() - t0
    finally:
        if own:
            pool.terminate()
            pool.join()
    n = sum(r[0] for r in res)
    return n / wall, n, wall, max(r[1] for r in res), res[0][3]


def cpu_pool(workers: int):
    """`workers` spawned processes, each past its imports and a two-clip warm-up (untimed)."""
    import multiprocessing as mp
    pool = mp.get_context("spawn").Pool(workers)
    pool.map(_cpu_worker, [(1000 + i, 2) for i in range(workers)], chunksize=1)
    return pool


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
    pool = cpu_pool(cores)
    try:
        for s in range(args.warmup + args.steps):
            v, n, wall, _, kind = cpu_throughput(per_worker, cores, seed0=s * 100, pool=pool)
            if s >= args.warmup:
                vals.append((v, n, wall))
    finally:
        pool.terminate()
        pool.join()
    n_tot = sum(v[1] for v in vals)
    t_tot = sum(v[2] for v in vals)
    value = n_tot / t_tot
    sample = f"{cores} procs x {per_worker} clips per step, {args.steps} steps (of the {BATCH}-clip workload)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(1, len(vals)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "e2e features + CoughDetectorResidual, per-clip loop, CPU", "batch": BATCH},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": kind, "sample": sample},
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
STREAM_SAMPLES = 57_600_000     # BASELINE configs[3]: 1 h of 16 kHz audio
STREAM_HOP = 1600               # 100 ms hop -> 35 991 one-second windows
SWEEP_CLIPS = 4_194_304         # BASELINE configs[4]
SWEEP_CHUNK = 16384


def run_gpu_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    cpu = None
    if world == 1 and not args.no_cpu_baseline:       # before CUDA is touched; bounded to ~10-20 s
        cores = min(host_cores(), 64)
        v, n, wall, _, kind = cpu_throughput(768, cores)
        cpu = {"value": v, "unit": "clips/s", "cores": cores, "kind": kind,
               "sample": f"{n} clips of the same synthetic workload, {cores} single-thread processes, {wall:.1f} s"}

    import torch
    import torch.distributed as dist
    from cough_detector_b200 import _lib, synth, create_model, sharding
    from cough_detector_b200.model import _PREC, DEFAULT_PRECISION

    torch.cuda.set_device(local_rank)
    numa_bound = sharding.bind_to_gpu_numa(local_rank)   # pinned host buffers on the GPU's own socket (e2e path is PCIe-bound)
    if world > 1:
        # NCCL's environment (NCCL_DEBUG, ...) is left exactly as the launcher set it
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    precision = args.precision or DEFAULT_PRECISION
    prec = _PREC[precision]
    h = _lib.get_handle(local_rank)
    torch.manual_seed(0)
    model = create_model("residual", n_mels=90, num_classes=2, in_channels=1).eval()
    sd = synth.randomized_state_dict(model, seed=1)
    h.load_state_dict(sd, key="bench")

    B = args.batch
    audio = synth.synth_batch(B, seed=100 + rank, device=dev)      # 1.05 GB > 126 MB L2
    gathered = torch.empty(world * B, 2, device=dev) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return float(x)
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def one_pass():
        logits = h.e2e_forward(audio, prec)
        if world > 1:
            dist.all_gather_into_tensor(gathered, logits)       # the only collective: final logit gather
        return logits

    def timed_ms(fn, reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), r

    for _ in range(args.warmup):
        one_pass()
    barrier()
    # One step = `passes` back-to-back passes over the resident batch, sized (from a calibration pass) so that the timed
    # region lasts >= min_seconds: long enough for >= 10 clock samples and for the sustained power / clock regime.
    cal_ms, _ = timed_ms(one_pass, 3)
    per_pass = max_over_ranks(cal_ms / 3)
    passes = max(1, int(-(-args.min_seconds * 1e3 // (per_pass * args.steps))))
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = h.launches
    barrier()
    w0 = time.perf_counter()
    ms, logits = timed_ms(one_pass, args.steps * passes)
    barrier()
    w1 = time.perf_counter()
    launches = h.launches - l0
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    ms = max_over_ranks(ms)
    value = world * B * args.steps * passes / (ms * 1e-3)

    # ---- per-stage timing for the roofline objects ----------------------------------------------------
    # feature kernel alone on BASELINE configs[1]: 65 536 resident clips (4.2 GB in, 2.4 GB out), fp32 features out
    FB = args.feat_batch
    faudio = synth.synth_batch(FB, seed=200 + rank, device=dev)
    ffeat = torch.empty(FB, 1, 90, 101, device=dev)

    def feat_pass():
        h.lib.cd_feat_forward(h._h, faudio.data_ptr(), 16000, FB, ffeat.data_ptr(), 1, 0.97, h._stream())

    timed_ms(feat_pass, 2)
    nrep = max(3, min(args.steps, 10))
    fms, _ = timed_ms(feat_pass, nrep)
    feat_ms = fms / nrep
    del faudio, ffeat
    feat = h.feat_forward(audio)
    timed_ms(lambda: h.cnn_forward(feat, prec), 1)
    cms, _ = timed_ms(lambda: h.cnn_forward(feat, prec), nrep)
    cnn_ms = cms / nrep
    del feat

    # ---- end to end through the C ABI with HOST buffers ----------------------------------------------
    host = audio.cpu().pin_memory()
    out = torch.empty(B, 2).pin_memory()
    for _ in range(max(1, min(2, args.warmup))):
        h.e2e_forward_host(host, prec, out)
    e2e_reps = max(args.steps, int(args.min_seconds / 0.02))   # ~20 ms per call at PCIe speed -> >= min_seconds
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_reps):
        h.e2e_forward_host(host, prec, out)         # synchronous: returns when `out` is complete
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * B * e2e_reps / e2e_s
    same = bool(torch.equal(out, logits.cpu()))
    # what the link itself does: a bare pinned host->device copy of the same buffer, all ranks at once (the e2e
    # path's floor; on a multi-GPU box the ranks share the host's memory and PCIe fabric)
    barrier()
    dev_copy = torch.empty_like(audio)
    dev_copy.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    pms, _ = timed_ms(lambda: dev_copy.copy_(host, non_blocking=True), 5)
    h2d_gbs = 5 * B * AUDIO_BYTES / (max_over_ranks(pms) * 1e-3) / 1e9
    del dev_copy

    # additive: the same call fed int16 PCM (the wav sample format) -- half the host->device bytes
    pcm = (host / host.abs().amax(dim=1, keepdim=True).clamp(min=1e-9) * 30000.0).round().to(torch.int16).pin_memory()
    out16 = torch.empty(B, 2).pin_memory()
    h.e2e_forward_host(pcm, prec, out16)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_reps):
        h.e2e_forward_host(pcm, prec, out16)
    torch.cuda.synchronize()
    i16_s = max_over_ranks(time.perf_counter() - t0)
    e2e_i16 = world * B * e2e_reps / i16_s
    del host, pcm

    # ---- BASELINE configs[3]: 1 h stream, 1-s window, 100 ms hop, time-sharded over the ranks (shared STFT frames) ----
    stream_obj = None
    if not args.no_stream:
        w0s, nw, s0, s1 = sharding.shard_stream(STREAM_SAMPLES, STREAM_HOP, rank, world)
        shard = synth.synth_batch((s1 - s0 + 15999) // 16000, seed=900 + rank, device=dev).reshape(-1)[: s1 - s0]
        fn = lambda: h.e2e_forward_stream(shard, STREAM_HOP, prec)
        timed_ms(fn, 2)
        barrier()
        sms, lg = timed_ms(fn, 10)
        assert lg.shape[0] == nw
        sms = max_over_ranks(sms / 10)
        total = sharding.stream_windows(STREAM_SAMPLES, STREAM_HOP)
        if world > 1:
            counts = [sharding.shard_stream(STREAM_SAMPLES, STREAM_HOP, r, world)[1] for r in range(world)]
            lg = sharding.gather_rows(lg, counts)                  # probabilities of the whole hour, in time order
        assert lg.shape[0] == total
        stream_obj = {"workload": "BASELINE configs[3]: 57.6 M samples (1 h @ 16 kHz), 1-s window, 100 ms hop, audio resident, "
                                  f"time-sharded x{world} with 14 400-sample overlap, shared STFT frames, bf16 CNN",
                      "windows": total, "ms": sms, "windows_per_s": total / (sms * 1e-3), "x_realtime": 3600.0 / (sms * 1e-3),
                      "api": "cd_e2e_forward_stream per rank + one gather of the (windows, 2) logits"}
        del shard, lg

    # ---- BASELINE configs[4]: 4 194 304 clips end to end, strong scaling: every rank takes 1/N of the clips -----------
    # The clips (268 GB as fp32) cannot be resident: each 16 384-clip chunk is generated on the device from a
    # counter-based seed (seed = 7_000_000 + global chunk index), then run; only the hot path is timed (CUDA events).
    sweep_obj = None
    if not args.no_sweep:
        n_chunks = SWEEP_CLIPS // SWEEP_CHUNK
        c0, c1 = sharding.shard_range(n_chunks, rank, world)
        acc_ms, checksum = 0.0, 0.0
        barrier()
        ws0 = time.perf_counter()
        for c in range(c0, c1):
            chunk = synth.synth_batch(SWEEP_CHUNK, seed=7_000_000 + c, device=dev)
            m, lg = timed_ms(lambda: h.e2e_forward(chunk, prec), 1)
            acc_ms += m
            if c == c0:
                checksum = float(lg.double().sum().item())
        sweep_ms = max_over_ranks(acc_ms)
        wall = max_over_ranks(time.perf_counter() - ws0)
        sweep_obj = {"workload": f"BASELINE configs[4]: {SWEEP_CLIPS} clips end to end, {SWEEP_CHUNK}-clip chunks generated on the "
                                 "device from seed 7 000 000 + chunk index, strong scaling (1/N of the chunks per rank)",
                     "clips": SWEEP_CLIPS, "hot_path_ms": sweep_ms, "clips_per_s": SWEEP_CLIPS / (sweep_ms * 1e-3),
                     "wall_s_with_generation": wall, "scaling": "strong", "first_chunk_logit_checksum_rank0": checksum}

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
                       "passes_per_step": passes, "ms_per_pass": ms / (args.steps * passes),
                       "timed_region_s": ms * 1e-3,
                       "parallelism": f"clip-sharded x{world}, no data-path collective, final logit all_gather",
                       "l2": "inputs (1.05 GB/GPU) larger than L2, no flush needed", "precision": precision},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": B * AUDIO_BYTES,
                    "d2h_bytes_per_step": B * 8, "bit_identical_to_device_path": same, "calls_timed": e2e_reps,
                    "h2d_gbs": e2e_value / world * AUDIO_BYTES / 1e9, "h2d_gbs_bare_copy": h2d_gbs,
                    "host_numa_bound": numa_bound,
                    "api": "cd_e2e_forward_host (pinned host audio in, host logits out)"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "feat_kernel", "bound": "hbm", "achieved": feat_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": feat_gbs / hbm_peak, "traffic": FB * NCU_TRAFFIC_PER_CLIP, "peak_source": peak_src,
                         "traffic_source": NCU_TRAFFIC_SOURCE,
                         "algorithmic_bytes_per_clip": AUDIO_BYTES + FEAT_BYTES, "launch_ms": feat_ms,
                         "clips_per_s": FB / (feat_ms * 1e-3), "clips_per_launch": FB,
                         "note": "BASELINE configs[1]: timed alone over 65 536 resident clips right after the e2e region"},
            "roofline_cnn": {"kernel": "cd_cnn_forward: fp32 features -> blocked bf16 map (1 launch) + stem + 2 residual blocks + head", "bound": "tensor", "achieved": cnn_tfs,
                             "peak": tc_peak, "unit": "TFLOP/s", "frac": cnn_tfs / tc_peak, "launch_ms": cnn_ms,
                             "peak_note": "sustained bf16 peak: the e2e region runs >= 1 s back to back",
                             "algorithmic_flops_per_clip": CNN_FLOPS, "clips_per_s": B / (cnn_ms * 1e-3)},
            "e2e_int16": {"value": e2e_i16, "unit": "clips/s", "h2d_bytes_per_step": B * 32000, "d2h_bytes_per_step": B * 8,
                          "api": "cd_e2e_forward_host_i16 (pinned int16 PCM in; additive entry point, not in the reference)"},
            "stream": stream_obj,
            "sweep": sweep_obj,
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
    ap.add_argument("--no-stream", action="store_true", help="skip the configs[3] stream object")
    ap.add_argument("--no-sweep", action="store_true", help="skip the configs[4] 4 M-clip sweep object")
    ap.add_argument("--min-seconds", type=float, default=1.2, help="minimum length of the timed regions")
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
