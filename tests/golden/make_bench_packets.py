#!/usr/bin/env python3
"""Generate tests/golden/bench_<workload>_packets.npz: the bitstreams bench.py's packets -> PCM leg decodes.

K distinct loopable streams (tests/heaac_streams.py over the reference's aacenc core)
of N frames each; bench.py replicates them over its batch of streams and walks them cyclically.  Only packets
are stored (no decoder output).  Needs oracle/_ref.  Run:  python tests/golden/make_bench_packets.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import heaac_streams as hs  # noqa: E402

K, N = 8, 64


def main():
    for name, stereo, ps in (("heaac_v1_stereo", True, False), ("heaac_v2_ps", False, True)):
        sizes, blobs, ed = [], [], None
        for k in range(K):
            pkts, ed, _ = hs.make_stream(5000 + k, N, stereo=stereo, ps=ps, loopable=True)
            sizes.append([len(p) for p in pkts])
            blobs.append(b"".join(pkts))
        path = os.path.join(HERE, "bench_%s_packets.npz" % name)
        np.savez_compressed(path, extradata=np.frombuffer(ed, np.uint8), pkt_sizes=np.asarray(sizes, np.int32),
                            pkt_blob=np.frombuffer(b"".join(blobs), np.uint8), core_rate=np.int32(24000),
                            channels=np.int32(2 if stereo else 1))
        print(name, "%d streams x %d frames, %.0f B/frame, %.0f KB" % (K, N, np.mean(sizes), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
