"""Parity at BASELINE durations and at the config-5 shard size (-m gpu).

BASELINE.json quotes every config on 10 s of audio (430 AAC-LC frames at 44.1 kHz, 234 HE-AAC frames at a 24 kHz
core).  The recursive state of the path -- PS all-pass chains and transient smoothers (aacps.c:645-754), the chirp
factors bw_array (aacsbr.c:1316-1334), the gain / noise smoother history (aacsbr.c:1632-1647), stale H-matrix rows
(aacps.c:816-823), the noise / sine phase counters -- is compared here with the UNMODIFIED reference decoder
(oracle/_ref, run live on the GPU box) over that whole length, packets in -> PCM out through heaac_dec_decode
(aac_decode_frame, aacdec.c:1973-2108), and the place of the largest deviation is reported.
Tolerance (north star): float PCM <= 1e-5 of full scale, s16 within 1 LSB."""
import numpy as np
import pytest

import ffmpeg_heaac_b200 as hb
import pyoracle
from common import FULL_SCALE, TOL_FS, abi
from conftest import needs_ref

pytestmark = pytest.mark.gpu


def _decode_cuda(pk_streams, extradata, core_rate, channels, gpu_parse, max_frames=64):
    """pk_streams: list (streams) of lists (frames) of packets.  Decodes in calls of max_frames frames (state carries
    over between calls, like consecutive avcodec_decode_audio3 calls on one context)."""
    S, T = len(pk_streams), len(pk_streams[0])
    dec = hb.Decoder(S, extradata, core_rate, channels, max_frames=max_frames, want_float=True, gpu_parse=gpu_parse)
    pcm, f32 = [], []
    for t0 in range(0, T, max_frames):
        rows = [[pk_streams[s][t] for s in range(S)] for t in range(t0, min(T, t0 + max_frames))]
        p, f, consumed = dec.decode(hb.PacketBatch(rows))
        assert (consumed > 0).all()
        pcm.append(p); f32.append(f)
    info = dec.info()
    dec.close()
    assert info["parse_errors"] == 0 and info["gpu_launches"] > 0
    return np.concatenate(pcm), np.concatenate(f32)


def _report(what, f32, ref_f32, s16, ref_s16):
    d = np.abs(f32.astype(np.float64) - ref_f32)
    t, s = np.unravel_index(int(d.argmax()), d.shape)[:2]
    err = float(d.max()) / FULL_SCALE
    lsb = int(np.abs(s16.astype(np.int32) - ref_s16.astype(np.int32)).max())
    per_frame = d.reshape(d.shape[0], -1).max(axis=1) / FULL_SCALE
    print("%s: max float err %.3g FS at frame %d of %d (stream %d); last-quarter max %.3g FS; s16 max diff %d LSB, %.4f %% of samples differ"
          % (what, err, t, d.shape[0], s, float(per_frame[-len(per_frame) // 4:].max()), lsb,
             100.0 * float((s16 != ref_s16).mean())))
    assert err <= TOL_FS, "%s: float PCM error %.3g FS exceeds %.0e (frame %d)" % (what, err, TOL_FS, t)
    assert lsb <= 1, "%s: s16 differs by %d LSB" % (what, lsb)


@needs_ref
@pytest.mark.parametrize("mode", ["v1", "v2"])
@pytest.mark.parametrize("gpu_parse", [False, True])
def test_ten_second_heaac_streams_match_reference(mode, gpu_parse):
    """240 frames (10.2 s) of HE-AAC v1 stereo / v2 mono+PS: every SBR frame class, coupling, IPD/OPD and several
    20<->34-band PS switches occur; compared with the reference decoder frame by frame."""
    import heaac_streams as hs
    T = 240
    stereo, ps = mode == "v1", mode == "v2"
    ch = 2 if stereo else 1
    seeds = (9101, 9102, 9103) if stereo else (9201, 9202, 9203)
    streams, ref_f32, ref_s16 = [], [], []
    lead_var = trail_var = both_var = switches = ipdopd = coupled = 0
    ed = None
    for seed in seeds:
        pkts, ed, _ = hs.make_stream(seed, T, stereo=stereo, ps=ps)
        frames, d = hs.decode_with_reference(pkts, ed, 24000, ch)
        d.close()
        streams.append(pkts)
        ref_f32.append(np.stack([f["pcm_f32"] for f in frames]))
        ref_s16.append(np.stack([f["pcm"] for f in frames]))
        for f in frames:
            c0 = f["sbr"][0]["ch"][0]
            ne = int(c0["bs_num_env"])
            lv, tv = int(c0["t_env"][0]) != 0, int(c0["t_env"][ne]) != 16
            lead_var += lv; trail_var += tv; both_var += lv and tv
            coupled += int(f["sbr"][0]["bs_coupling"])
        if ps:
            p = np.stack([f["ps"][0] for f in frames])
            switches += int((np.diff(p["is34bands"].astype(int)) != 0).sum())
            ipdopd += int(p["enable_ipdopd"].sum())
    # all four frame classes: FIXFIX (neither), FIXVAR (trailing), VARFIX (leading), VARVAR (both)
    assert lead_var > both_var > 0 and trail_var > both_var and lead_var + trail_var - both_var < len(seeds) * T
    if ps:
        assert switches >= 2 and ipdopd > 0
    else:
        assert coupled > 0
    pcm, f32 = _decode_cuda(streams, ed, 24000, ch, gpu_parse)
    _report("HE-AAC %s, %d frames, %s parser" % (mode, T, "device" if gpu_parse else "host"),
            f32, np.stack(ref_f32, axis=1), pcm, np.stack(ref_s16, axis=1))


@needs_ref
def test_ten_second_lc_stereo_44k_matches_reference():
    """BASELINE config 1 as stated: AAC-LC stereo 44.1 kHz, 10 s (430 frames) of LCG noise (seed 1234) encoded by the
    reference's own aacenc, decoded by the reference on the CPU and by this repo's decoder."""
    import pyref
    T = 430
    pcm_in = pyref.lcg_noise(1024 * (T + 1), 2, seed=1234)
    enc = pyref.Encoder(44100, 2, 128000)
    pkts = enc.encode(pcm_in)[:T]
    ed = enc.extradata
    enc.close()
    d = pyref.Decoder(ed, 44100, 2, "stereo", float_mode=True)
    frames = [d.decode(p) for p in pkts]
    d.close()
    ref_f32 = np.stack([f["pcm_f32"] for f in frames])[:, None]
    ref_s16 = np.stack([f["pcm"] for f in frames])[:, None]
    for gpu_parse in (False, True):
        pcm, f32 = _decode_cuda([pkts], ed, 44100, 2, gpu_parse, max_frames=128)
        _report("AAC-LC stereo 44.1 kHz, %d frames, %s parser" % (T, "device" if gpu_parse else "host"), f32, ref_f32, pcm, ref_s16)


def test_config5_shard_8192_streams_matches_oracle():
    """One GPU's shard of BASELINE config 5 at 8 GPUs: 8192 HE-AAC v1 stereo streams (16,384 CTAs of k_chain, 22 waves), 32 frames.  A 32-stream sample spread over the batch is compared with the
    oracle over all 32 frames; identical streams give identical PCM wherever they sit in the batch."""
    from ffmpeg_heaac_b200 import synthetic
    S, T, K = 8192, 32, 32
    cfg, a = synthetic.make_workload("heaac_v1_stereo", S, T, seed=2025, K=K)
    rep = np.random.default_rng(3).standard_normal((T, K) + a["coeffs"].shape[2:]).astype(np.float32) * 150
    rep[..., 640:] = 0
    a["coeffs"] = np.ascontiguousarray(np.tile(rep, (1, S // K, 1, 1)))
    ctx = hb.Context(cfg)
    bufs = {k: ctx.alloc(a[k].nbytes).upload(a[k]) for k in ("coeffs", "ics", "sbr")}
    shape = (T, S, ctx.out_len, ctx.nout)
    out = ctx.alloc(int(np.prod(shape)) * 2)
    ctx.synth_device(T, bufs["coeffs"], bufs["ics"], out, sbr=bufs["sbr"])
    ctx.sync()
    pcm = out.download(np.int16, shape)
    assert ctx.launch_count > 0
    ctx.close()
    pick = np.arange(K) + K * ((np.arange(K) * 37) % (S // K))          # stream k of copy (37 k mod 256): all K patterns
    sub = abi.make_config(K, [abi.TYPE_CPE], cfg.sr_index, sbr=True, ps=False, max_frames=T)
    orc = pyoracle.Oracle(sub)
    opcm, _ = orc.synth(np.ascontiguousarray(a["coeffs"][:, :K]), np.ascontiguousarray(a["ics"][:, :K]), None,
                        np.ascontiguousarray(a["sbr"][:, :K]), None, n_threads=8, want_float=False)
    orc.close()
    diff = int(np.abs(pcm[:, pick].astype(np.int32) - opcm).max())
    print("config-5 shard: 8192 streams x 32 frames, 32-stream sample vs oracle: max s16 diff %d LSB, bit-exact: %s"
          % (diff, np.array_equal(pcm[:, pick], opcm)))
    assert diff == 0
    assert np.array_equal(pcm[:, :K], pcm[:, S - K:]), "stream results depend on batch position"
