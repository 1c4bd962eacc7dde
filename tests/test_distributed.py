"""N > 1 host logic on CPU: world_size-2 gloo run of the sharding + final gather (no GPU needed)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cough_detector_b200 import sharding


def test_shard_ranges_partition_everything():
    for n in (0, 1, 7, 16384, 35991):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_stream_shards_cover_every_window_once():
    n, hop = 57_600_000, 1600                         # BASELINE config 4: 1 h at 100 ms hop
    assert sharding.stream_windows(n, hop) == 35991
    seen = 0
    for r in range(8):
        w0, nw, s0, s1 = sharding.shard_stream(n, hop, r, 8)
        assert w0 == seen and s0 == w0 * hop and s1 <= n
        assert sharding.stream_windows(s1 - s0, hop) == nw       # the shard alone yields exactly its windows
        seen += nw
    assert seen == 35991
    assert sharding.shard_stream(15999, 1600, 0, 2) == (0, 0, 0, 0)


def _worker(rank, world, port, n, hop):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w0, nw, s0, s1 = sharding.shard_stream(n, hop, rank, world)
    # stand-in for the per-window logits of this shard: a function of the global window index and of the
    # shard's own samples, so mis-ordered or duplicated windows are caught
    stream = torch.arange(n, dtype=torch.float32)
    local = torch.stack([torch.stack([stream[s0 + j * hop], stream[s0 + j * hop + sharding.WINDOW - 1]]) for j in range(nw)]) \
        if nw else torch.zeros(0, 2)
    counts = [sharding.shard_stream(n, hop, r, world)[1] for r in range(world)]
    full = sharding.gather_rows(local, counts)
    want = torch.stack([torch.tensor([float(j * hop), float(j * hop + sharding.WINDOW - 1)])
                        for j in range(sharding.stream_windows(n, hop))])
    assert torch.equal(full, want), (rank, full.shape, want.shape)
    c0, c1 = sharding.shard_range(1001, rank, world)
    clips = torch.arange(c0, c1, dtype=torch.float32).view(-1, 1).repeat(1, 2)
    allc = sharding.gather_rows(clips, [e - s for s, e in (sharding.shard_range(1001, r, world) for r in range(world))])
    assert torch.equal(allc[:, 0], torch.arange(1001, dtype=torch.float32))
    dist.destroy_process_group()


def test_two_rank_gloo_gather_in_time_order():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 16000 + 1600 * 36 + 7, 1600), nprocs=2, join=True)
