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
import lc_writer  # noqa: E402
import pyref  # noqa: E402

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


def lc_cases():
    """AAC-LC fixtures: config 4 (5.1, TNS, short blocks; own writer, no PNS so that frames stay on the parallel parse
    path) and config 1 batched (stereo 44.1 kHz, the reference's aacenc on noise).  LC frames carry no inter-frame
    parser state, so any stream loops."""
    out = {}
    sizes, blobs = [], []
    for k in range(K):
        pkts, ed, _ = lc_writer.make_lc_stream(7000 + k, N, "5.1", 3, pns=False)
        sizes.append([len(p) for p in pkts]); blobs.append(b"".join(pkts))
    out["lc_51_48k_tns"] = (sizes, blobs, ed, 48000, 6)
    sizes, blobs = [], []
    for k in range(K):
        pcm = pyref.lcg_noise(1024 * (N + 1), 2, seed=8000 + k)
        enc = pyref.Encoder(44100, 2, 128000)
        pkts, ed = enc.encode(pcm)[:N], enc.extradata
        enc.close()
        sizes.append([len(p) for p in pkts]); blobs.append(b"".join(pkts))
    out["lc_stereo_44k"] = (sizes, blobs, ed, 44100, 2)
    for name, (sizes, blobs, ed, rate, ch) in out.items():
        path = os.path.join(HERE, "bench_%s_packets.npz" % name)
        np.savez_compressed(path, extradata=np.frombuffer(ed, np.uint8), pkt_sizes=np.asarray(sizes, np.int32),
                            pkt_blob=np.frombuffer(b"".join(blobs), np.uint8), core_rate=np.int32(rate), channels=np.int32(ch))
        print(name, "%d streams x %d frames, %.0f B/frame, %.0f KB" % (K, N, np.mean(sizes), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    lc_cases()
    main()
