#!/usr/bin/env python3
"""Quick on-GPU parity probe (development aid): CUDA back end vs oracle vs reference on small cases."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import ffmpeg_heaac_b200 as hb
import pyref, pyoracle, heaac_streams as hs
abi = hb.abi

def run(mode, seeds, nfr, bias_mode=False):
    streams = []
    for seed in seeds:
        if mode == "lc":
            pcm = pyref.lcg_noise(1024 * (nfr + 1), 2, seed=seed).astype(np.int32)
            for st in (5000, 12000, 13500, 22000): pcm[st:st + 300] *= 8
            enc = pyref.Encoder(44100, 2, 128000); pk = enc.encode(np.clip(pcm, -32768, 32767).astype(np.int16))[:nfr]
            dec = pyref.Decoder(enc.extradata, 44100, 2, "stereo"); frames = [dec.decode(p) for p in pk]
            et = [abi.TYPE_CPE]; sbr = ps = False; d = dec
        else:
            stereo = mode == "v1"; ps = mode == "v2"; sbr = True
            pkts, ed, info = hs.make_stream(seed, nfr, stereo=stereo, ps=ps)
            frames, d = hs.decode_with_reference(pkts, ed, 24000, 2 if stereo else 1)
            et = [abi.TYPE_CPE] if stereo else [abi.TYPE_SCE]
        streams.append(frames)
    S = len(streams)
    st = lambda k: np.stack([np.stack([f[k] for f in fr]) for fr in streams], axis=1)   # [T,S,...]
    cfg = abi.make_config(S, et, d.sr_index, sbr=sbr, ps=ps, bias_mode=bias_mode, max_frames=nfr)
    co = st("coeffs") * (np.float32(1 / 32768) if bias_mode else np.float32(1))
    args = (co, st("ics"), st("tns"), st("sbr") if sbr else None, st("ps")[:, :, 0] if ps else None)
    orc = pyoracle.Oracle(cfg)
    opcm, of32 = orc.synth(*args)
    ctx = hb.Context(cfg)
    t0 = time.time()
    gpcm, gf32 = ctx.synth(*args, want_float=True)
    dt = time.time() - t0
    rp = np.stack([np.stack([f["pcm"] for f in fr]) for fr in streams], axis=1)
    print("%s S=%d T=%d bias=%d: gpu-vs-oracle float maxdiff %.6g s16 mismatch %d | gpu-vs-ref s16 maxdiff %d mismatch %d of %d | %.1f ms launches %d"
          % (mode, S, nfr, bias_mode, np.abs(gf32 - of32).max(), (gpcm != opcm).sum(),
             np.abs(gpcm.astype(int) - rp).max(), (gpcm != rp).sum(), rp.size, dt * 1e3, ctx.launch_count), flush=True)
    ctx.close()

if __name__ == "__main__":
    print(hb.lib().heaac_b200_version().decode())
    run("lc", [1, 2, 3], 12)
    run("lc", [1, 2, 3], 12, bias_mode=True)
    run("v1", [7, 8, 9, 10], 24)
    run("v1", [7, 8], 24, bias_mode=True)
    run("v2", [21, 26, 27, 28], 24)
