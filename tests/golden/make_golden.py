#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, needs /root/reference to build).

Each file holds, for a short synthetic stream: the raw_data_block packets and extradata, what the
reference's parsers hand to spectral_to_sample (aacdec.c:2078) packed into the C-ABI structs, and the
reference decoder's own output for them (s16 from the default C 'bias' path and full-precision float PCM
from its SIMD-scaling path).  Committed so that the oracle and the CUDA path can be pinned where the
reference is absent (the GPU box).  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyref  # noqa: E402
import heaac_streams as hs  # noqa: E402
import lc_writer  # noqa: E402

abi = pyref.abi


def lc_stream(seed, nfr, channels, rate, layout):
    pcm = pyref.lcg_noise(1024 * (nfr + 1), channels, seed=seed).astype(np.int32)
    for st in (3000, 7000, 8200, 15000):
        pcm[st:st + 300] *= 6          # attacks -> LONG_START / EIGHT_SHORT / LONG_STOP
    enc = pyref.Encoder(rate, channels, 64000 * channels)
    pk = enc.encode(np.clip(pcm, -32768, 32767).astype(np.int16))[:nfr]
    ed = enc.extradata
    enc.close()
    return pk, ed, rate, channels, layout


def capture(pkts, ed, rate, channels, layout):
    df = pyref.Decoder(ed, rate, channels, layout, float_mode=True)
    db = pyref.Decoder(ed, rate, channels, layout, float_mode=False)
    fr = [df.decode(p) for p in pkts]
    frb = [db.decode(p, capture=False) for p in pkts]
    out = {k: np.stack([f[k] for f in fr]) for k in ("coeffs", "ics", "tns", "sbr", "ps", "pcm_f32")}
    out["pcm_s16_simd_path"] = np.stack([f["pcm"] for f in fr])
    out["pcm_s16_c_path"] = np.stack([f["pcm"] for f in frb])
    out["sr_index"] = np.int32(df.sr_index)
    out["sbr_flag"] = np.int32(1 if df.sbr == 1 else 0)
    out["ps_flag"] = np.int32(1 if df.ps == 1 and channels == 1 else 0)
    out["elem_types"] = np.asarray(pyref.LAYOUTS[layout][0], np.int32)
    out["extradata"] = np.frombuffer(ed, np.uint8)
    out["pkt_sizes"] = np.asarray([len(p) for p in pkts], np.int32)
    out["pkt_blob"] = np.frombuffer(b"".join(pkts), np.uint8)
    out["core_rate"] = np.int32(rate)
    out["downsampled"] = np.int32(fr[0]["pcm_f32"].shape[-1] == 1024 and df.sbr == 1)
    df.close(); db.close()
    return out


def cce_stream(seed, n_frames):
    """AAC-LC 48 kHz with a program_config_element: SCE + CPE and two independently switched coupling channel elements
    (tests/lc_writer.py).  Returns (packets, extradata, (element types, element ids) in the reference's element order)."""
    et = [abi.TYPE_SCE, abi.TYPE_CPE, abi.TYPE_CCE, abi.TYPE_CCE]
    ed, ids, order = lc_writer.pce_config(3, et)
    w = lc_writer.LcWriter(seed, 3, et, tns_prob=0.4, short_bias=0.3)
    # CCE 0 -> the SCE and both channels of the CPE with separate gains; CCE 1 -> both CPE channels with one gain
    w.set_coupling({2: [(0, 2), (1, 3)], 3: [(1, 0)]})
    pk = [w.frame() for _ in range(n_frames)]
    return pk, ed, ([et[e] for e in order], [ids[e] for e in order])


def make_cce_golden():
    """Independent channel coupling (apply_independent_coupling, aacdec.c:1849-1862): boundary-level golden -- what the
    reference's parsers hand over (incl. the resolved coupling lists, oracle/ref_harness.c pack_cce_all) and its PCM."""
    pk, ed, lay = cce_stream(4242, 10)
    df = pyref.Decoder(ed, 48000, 3, lay, float_mode=True)
    fr = [df.decode(p) for p in pk]
    df.close()
    out = {k: np.stack([f[k] for f in fr]) for k in ("coeffs", "ics", "tns", "cce", "pcm_f32")}
    out["pcm_s16"] = np.stack([f["pcm"] for f in fr])
    out["elem_types"] = np.asarray(lay[0], np.int32)
    out["sr_index"] = np.int32(3)
    out["extradata"] = np.frombuffer(ed, np.uint8)           # AudioSpecificConfig with the program config element
    out["pkt_sizes"] = np.asarray([len(p) for p in pk], np.int32)
    out["pkt_blob"] = np.frombuffer(b"".join(pk), np.uint8)
    path = os.path.join(HERE, "lc_cce.npz")
    np.savez_compressed(path, **out)
    print("lc_cce", out["coeffs"].shape, out["pcm_f32"].shape, "%.0f KB" % (os.path.getsize(path) / 1024))


def main():
    cases = {
        "lc_stereo_44k": lc_stream(1234, 10, 2, 44100, "stereo"),
        "lc_51_48k": lc_stream(77, 6, 6, 48000, "5.1"),
    }
    for name, seed, stereo, ps in (("heaac_v1_stereo", 7, True, False), ("heaac_v2_ps", 26, False, True)):
        pkts, ed, info = hs.make_stream(seed, 12, stereo=stereo, ps=ps)
        cases[name] = (pkts, ed, 24000, 2 if stereo else 1, "stereo" if stereo else "mono")
    # own AAC-LC writer: TNS, sine + KBD windows, grouped short blocks, PNS, intensity, pulses, escapes (config 4)
    pkts, ed, w = lc_writer.make_lc_stream(4004, 8, "5.1", 3)
    cases["lc_51_tns_rich"] = (pkts, ed, 48000, 6, "5.1")
    print("lc_51_tns_rich content:", w.stats)
    # downsampled SBR: hierarchical explicit signalling (AOT 5) with the extension rate equal to the core rate
    pkts, ed, info = hs.make_stream(31, 8, stereo=True, ps=False)
    cases["heaac_v1_downsampled"] = (pkts, hs.asc_explicit(6, 2, 6), 24000, 2, "stereo")
    # HE-AAC 5.1: SBR on SCE + 2 CPE, the LFE upsampled only (aacdec.c:1923-1925)
    pkts, ed, info = hs.make_stream_51(51, 8)
    cases["heaac_v1_51"] = (pkts, ed, 24000, 6, "5.1")
    only = sys.argv[1:]
    if not only or "lc_cce" in only:
        make_cce_golden()
    for name, c in cases.items():
        if only and name not in only:
            continue
        d = capture(*c)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, d["coeffs"].shape, d["pcm_f32"].shape, "%.0f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
