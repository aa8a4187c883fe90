"""The N>1 path on CPU: world_size-2 gloo run of the sharding + statistics gather used by bench.py, with the
oracle standing in for the per-rank compute (streams are independent: no data-path collective)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ffmpeg_heaac_b200 import sharding, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_streams():
    for n in (1, 7, 8, 1024, 65536):
        for w in (1, 2, 3, 4, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S, T = 6, 4
    cfg, a = synthetic.make_workload("heaac_v1_stereo", S, T, seed=9, K=6)
    lo, hi = sharding.shard_range(S, rank, world)
    cfg.n_streams = hi - lo
    orc = pyoracle.Oracle(cfg)
    pcm, _ = orc.synth(a["coeffs"][:, lo:hi], a["ics"][:, lo:hi], None, a["sbr"][:, lo:hi], None)
    orc.close()
    csum = int(pcm.astype(np.int64).sum())
    audio = (hi - lo) * T * 2048 / 48000.0
    ms, total_audio, total_csum = sharding.gather_stats(dist, "cpu", 10.0 * (rank + 1), audio, csum)
    if rank == 0:
        q.put((ms, total_audio, total_csum))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_matches_single_process():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    S, T = 6, 4
    cfg, a = synthetic.make_workload("heaac_v1_stereo", S, T, seed=9, K=6)
    orc = pyoracle.Oracle(cfg)
    pcm, _ = orc.synth(a["coeffs"], a["ics"], None, a["sbr"], None)
    orc.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ms, audio, csum = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == 20.0                                    # max over ranks
    assert abs(audio - S * T * 2048 / 48000.0) < 1e-9    # whole-job audio seconds
    assert csum == float(pcm.astype(np.int64).sum())     # sharded result == unsharded result
