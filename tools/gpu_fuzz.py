#!/usr/bin/env python3
"""Corrupted packets through the device front end (run it under compute-sanitizer --tool memcheck):
every stream of the batch gets its own corruption of the golden packets; the call must complete, report errors
through consumed[] and leave the decoder usable."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ffmpeg_heaac_b200 as hb  # noqa: E402
from common import GOLDEN_CASES, golden_packets, load_golden  # noqa: E402

rng = np.random.default_rng(11)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for name in GOLDEN_CASES:
    d = load_golden(name)
    pk = golden_packets(d)
    ch = d["pcm_f32"].shape[1] if not d["ps_flag"] else 1
    dec = hb.Decoder(S, d["extradata"].tobytes(), int(d["core_rate"]), ch, max_frames=len(pk), gpu_parse=True)
    dec.decode(hb.PacketBatch([[p] * S for p in pk]))          # intact first: configures
    bad = tot = 0
    for rep in range(3):
        rows = []
        for p in pk:
            row = []
            for s in range(S):
                b = bytearray(p)
                mode = int(rng.integers(0, 4))
                if mode == 0:
                    for _ in range(int(rng.integers(1, 6))):
                        b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
                elif mode == 1:
                    b = b[:int(rng.integers(1, len(b)))]
                elif mode == 2:
                    b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
                row.append(bytes(b))
            rows.append(row)
        pcm, _, cons = dec.decode(hb.PacketBatch(rows))
        bad += int((cons < 0).sum()); tot += cons.size
    pcm, _, cons = dec.decode(hb.PacketBatch([[p] * S for p in pk]))
    assert (cons > 0).all()
    dec.close()
    print("%-22s %d of %d corrupted packets rejected, decoder still fine" % (name, bad, tot), flush=True)
