#!/usr/bin/env python3
"""Host <-> device copy bandwidth of the box, all GPUs at once (what bounds the packets -> PCM `e2e` leg when it is sharded over N
GPUs: every rank moves 268 MB of s16 PCM to the host per 1398 audio-seconds).
    python tools/pcie_probe.py                                    # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 tools/pcie_probe.py
Every rank copies 256 MiB pinned <-> device 10 times between two barriers; rank 0 prints per-rank and aggregate GB/s."""
import json, os, time
import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 << 20
h, h2 = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d, d2 = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    with torch.cuda.stream(s1):
        h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)


res = {}
for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True)), ("both_each_way", both)):
    fn(); torch.cuda.synchronize()
    if dist: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = torch.tensor([10 * n / dt / 1e9], device="cuda", dtype=torch.float64)
    if dist:
        allv = [torch.zeros_like(gbs) for _ in range(world)]
        dist.all_gather(allv, gbs)
        res[name] = [round(float(v.item()), 1) for v in allv]
    else:
        res[name] = [round(float(gbs.item()), 1)]
if rank == 0:
    out = {"n_gpus": world, "host_cpus": os.cpu_count()}
    for k, v in res.items():
        out[k] = {"per_rank_GBps": v, "aggregate_GBps": round(sum(v), 1)}
    # e2e ceiling: 268.4 MB of PCM per 1398.1 audio-seconds per rank
    out["e2e_ceiling_audio_s_per_s"] = round(out["d2h"]["aggregate_GBps"] * 1e9 / 268435456 * 1398.1)
    print(json.dumps(out))
if dist:
    dist.destroy_process_group()
