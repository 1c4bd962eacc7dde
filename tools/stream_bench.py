"""BASELINE config 4: 1 h of 16 kHz audio, 1-s window, 100 ms hop (35 991 windows), audio -> P(cough).

    python tools/stream_bench.py                                          # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tools/stream_bench.py                             # time-sharded over N GPUs

Under torchrun every rank takes a contiguous run of windows plus the samples they cover (neighbouring shards
overlap by window - hop samples, cough_detector_b200/sharding.py); no data-path collective, one gather of the
(windows, 2) logits at the end.  Times are CUDA events on each rank, max over ranks.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cough_detector_b200 import _lib, create_model, sharding, synth  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h = _lib.get_handle(local)
torch.manual_seed(0)
m = create_model("residual").eval()
h.load_state_dict(synth.randomized_state_dict(m, seed=1), key="s")
n = 57_600_000
x = synth.synth_batch(n // 16000, seed=9, device=f"cuda:{local}").reshape(-1)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        r = fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=x.device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return r, float(ms.item())


out = {"n_gpus": world}
for name, hp in (("shared_stft_hop1600", 1600), ("per_window_hop1599", 1599)):
    w0, nw, s0, s1 = sharding.shard_stream(n, hp, rank, world)
    mine = x[s0:s1]
    lg, ms = timed(lambda: h.e2e_forward_stream(mine, hp, 1))
    assert lg.shape[0] == nw
    total = sharding.stream_windows(n, hp)
    if world > 1:
        counts = [sharding.shard_stream(n, hp, r, world)[1] for r in range(world)]
        lg = sharding.gather_rows(lg, counts)
    assert lg.shape[0] == total
    out[name] = {"windows": total, "ms": ms, "windows_per_s": total / ms * 1e3, "x_realtime": 3600.0 / (ms * 1e-3)}
    if world == 1:
        sub = x[: 16000 + hp * 8191]
        _, fms = timed(lambda: h.feat_forward_stream(sub, hp), warm=3)
        out[name]["features_only_windows_per_s"] = 8192 / fms * 1e3
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
