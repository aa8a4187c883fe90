#!/usr/bin/env python3
"""Small driver for profiling the device front end: decodes the bench packet fixture once or twice."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ffmpeg_heaac_b200 as hb  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
workload = sys.argv[4] if len(sys.argv) > 4 else "heaac_v1_stereo"
z = np.load(os.path.join(ROOT, "tests", "golden", "bench_%s_packets.npz" % workload))
sizes, blob = z["pkt_sizes"], z["pkt_blob"].tobytes()
K, N = sizes.shape
ofs = np.concatenate([[0], np.cumsum(sizes.ravel())])
pk = [[blob[ofs[k * N + i]:ofs[k * N + i + 1]] for i in range(N)] for k in range(K)]
batches = [hb.PacketBatch([[pk[s % K][(b * T + t) % N] for s in range(S)] for t in range(T)]) for b in range(max(1, N // T))]
dec = hb.Decoder(S, z["extradata"].tobytes(), int(z["core_rate"]), int(z["channels"]), max_frames=T, gpu_parse=True)
import torch
PCM_T = torch.empty(T * S * 2048 * 2 * 2, dtype=torch.uint8).pin_memory()
PCM = PCM_T.numpy().view(np.int16).reshape(T, S, 2048, 2)
PIPE = os.environ.get("PIPE") == "1"
if PIPE:
    PCM2_T = torch.empty(T * S * 2048 * 2 * 2, dtype=torch.uint8).pin_memory()
    PCM2 = PCM2_T.numpy().view(np.int16).reshape(T, S, 2048, 2)
    dec.decode(batches[0], pcm_out=PCM, out_shape=(2048, 2))
    t0 = time.perf_counter(); prev = None
    for i in range(steps):
        tk = dec.submit(batches[(i + 1) % len(batches)], [PCM, PCM2][i % 2])
        if prev is not None: dec.wait(prev)
        prev = tk
    dec.wait(prev)
    print("pipelined: %.2f ms/step" % ((time.perf_counter() - t0) * 1e3 / steps))
    steps = 0
for i in range(steps):
    t0 = time.perf_counter()
    pcm, _, cons = dec.decode(batches[i % len(batches)], pcm_out=PCM, out_shape=(2048, 2))
    print("step %d: %.2f ms, errors %d" % (i, (time.perf_counter() - t0) * 1e3, int((cons < 0).sum())))
print("k_parse ms/launch:", dec.info()["k_parse_ms"], "lanes", os.environ.get("HEAAC_PARSE_LANES", "1"))
dec.close()
