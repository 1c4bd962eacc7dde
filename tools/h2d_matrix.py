"""Host -> device copy matrix behind the e2e (host-buffer) number: what the box gives a rank when 1 / 2 / 4 / 8 ranks copy at
once, for page-locked vs write-combined host memory and one vs two copy streams, and what cd_e2e_forward_host makes of it.

    python tools/h2d_matrix.py                                                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/h2d_matrix.py

Every rank copies its own 16 384-clip buffer (1.05 GB fp32) at the same time; the number reported is the SLOWEST rank's
GB/s (CUDA events per rank, max over ranks).  Prints one JSON object on rank 0.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cough_detector_b200 import _lib, create_model, sharding, synth  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
bound = sharding.bind_to_gpu_numa(local) if os.environ.get("CD_NO_NUMA_BIND") is None else False
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
B = 16384
h = _lib.get_handle(local)
torch.manual_seed(0)
m = create_model("residual").eval()
h.load_state_dict(synth.randomized_state_dict(m, seed=1), key="m")
src = synth.synth_batch(B, seed=3 + rank, device=dev)
dst = torch.empty_like(src)
out = torch.empty(B, 2).pin_memory()


def sync_all():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def worst(gbs: float) -> float:
    if world == 1:
        return gbs
    t = torch.tensor([gbs], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return float(t.item())


def bare_copy(host: torch.Tensor, streams: int, reps: int = 4) -> float:
    ss = [torch.cuda.Stream() for _ in range(streams)]
    parts = list(zip(host.chunk(streams), dst.chunk(streams)))
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for s, (hp, dp) in zip(ss, parts):
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                dp.copy_(hp, non_blocking=True)
        for s in ss:
            torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return worst(reps * host.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9)


def e2e(host: torch.Tensor, reps: int = 8) -> float:
    h.e2e_forward_host(host, 1, out)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(reps):
        h.e2e_forward_host(host, 1, out)
    dt = time.perf_counter() - t0
    return worst(reps * B * 64000 / dt / 1e9)


res = {"n_ranks": world, "numa_bound": bound, "clips_per_rank": B}
pinned = src.cpu().pin_memory()
wc = _lib.host_alloc(B * 16000, torch.float32, write_combined=True).view(B, 16000)
wc.copy_(pinned)
for name, host in (("pinned", pinned), ("write_combined", wc)):
    res[name] = {"bare_copy_1_stream_gbs": bare_copy(host, 1), "bare_copy_2_streams_gbs": bare_copy(host, 2),
                 "e2e_forward_host_gbs": e2e(host)}
    res[name]["e2e_clips_per_s_all_ranks"] = res[name]["e2e_forward_host_gbs"] * 1e9 / 64000 * world
pcm = (pinned / pinned.abs().amax(dim=1, keepdim=True).clamp(min=1e-9) * 30000.0).round().to(torch.int16).pin_memory()
t = e2e(pcm) / 2            # e2e() counts fp32 bytes; int16 moves half
res["pinned_int16"] = {"e2e_forward_host_gbs": t, "e2e_clips_per_s_all_ranks": t * 1e9 / 32000 * world}
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
