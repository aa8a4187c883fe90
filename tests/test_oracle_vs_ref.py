"""Pins the oracle (oracle/heaac_oracle.c): against the committed golden vectors produced by the reference
decoder, and -- where oracle/_ref is available -- against the reference itself on fresh synthetic streams."""
import numpy as np
import pytest

import pyoracle
from common import GOLDEN_CASES, abi, assert_pcm_close, golden_args, golden_config, golden_ref_pcm, load_golden
from conftest import needs_ref


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("bias_mode", [False, True])
def test_oracle_reproduces_golden_vectors(name, bias_mode):
    d = load_golden(name)
    orc = pyoracle.Oracle(golden_config(d, bias_mode=bias_mode))
    pcm, f32 = orc.synth(*golden_args(d, bias_mode=bias_mode))
    orc.close()
    ref_f32, ref_s16 = golden_ref_pcm(d)
    if bias_mode:      # the C bias path carries the signal as 385 + x/32768: compare s16 only
        assert int(np.abs(pcm[:, 0].astype(np.int32) - ref_s16).max()) <= 1
    else:
        assert_pcm_close(f32[:, 0], ref_f32, pcm[:, 0], ref_s16, name)


def test_golden_both_reference_scalings_agree():
    """The reference's C-bias and SIMD-scaling paths give bit-identical s16."""
    for name in GOLDEN_CASES:
        d = load_golden(name)
        assert np.array_equal(d["pcm_s16_c_path"], d["pcm_s16_simd_path"]), name


@needs_ref
@pytest.mark.parametrize("mode,seed", [("v1", 101), ("v1", 102), ("v2", 201), ("v2", 202), ("v2", 203)])
def test_oracle_matches_reference_on_fresh_heaac_streams(mode, seed):
    import heaac_streams as hs
    stereo, ps = mode == "v1", mode == "v2"
    pkts, ed, info = hs.make_stream(seed, 28, stereo=stereo, ps=ps)
    frames, d = hs.decode_with_reference(pkts, ed, 24000, 2 if stereo else 1)
    assert d.sbr == 1 and (d.ps == 1) == ps
    cfg = abi.make_config(1, [abi.TYPE_CPE] if stereo else [abi.TYPE_SCE], d.sr_index, sbr=True, ps=ps)
    st = lambda k: np.stack([f[k] for f in frames])[:, None]
    orc = pyoracle.Oracle(cfg)
    pcm, f32 = orc.synth(st("coeffs"), st("ics"), st("tns"), st("sbr"), st("ps")[:, :, 0] if ps else None)
    orc.close()
    assert_pcm_close(f32[:, 0], np.stack([f["pcm_f32"] for f in frames]), pcm[:, 0],
                     np.stack([f["pcm"] for f in frames]), "%s/%d" % (mode, seed))
    if ps:      # the fixture must really exercise PS: 20- and 34-band modes, IPD/OPD
        assert any(int(f["ps"][0]["start"]) for f in frames)


@needs_ref
@pytest.mark.parametrize("seed", [5, 10])
def test_oracle_matches_reference_on_heaac_51(seed):
    """HE-AAC 5.1: SBR on the SCE and the CPEs; the LFE goes through ff_sbr_apply in its initial state (upsampling only,
    aacdec.c:1923-1925, aacsbr.c:128-137).  The cores come from the repo's AAC-LC writer (random quantised spectra, hot and
    partly sparse): on such cores the inverse filter is ill-conditioned and other seeds reach 4e-5 FS in float
    while staying within 1 LSB in s16; the two seeds here are well-conditioned ones."""
    import heaac_streams as hs
    import pyref
    pkts, ed, info = hs.make_stream_51(seed, 20)
    dec = pyref.Decoder(ed, info["core_rate"], 6, "5.1", float_mode=True)
    frames = [dec.decode(p) for p in pkts]
    assert dec.sbr == 1
    dec.close()
    cfg = abi.make_config(1, pyref.LAYOUTS["5.1"][0], 6, sbr=True, ps=False, max_frames=len(frames))
    st = lambda k: np.stack([f[k] for f in frames])[:, None]
    orc = pyoracle.Oracle(cfg)
    pcm, f32 = orc.synth(st("coeffs"), st("ics"), st("tns"), st("sbr"), None)
    orc.close()
    ref_f32 = np.stack([f["pcm_f32"] for f in frames])
    scale = max(32768.0, float(np.abs(ref_f32).max()))
    assert float(np.abs(f32[:, 0].astype(np.float64) - ref_f32).max()) / scale <= 1e-5
    assert int(np.abs(pcm[:, 0].astype(np.int32) - np.stack([f["pcm"] for f in frames])).max()) <= 1
    assert float(np.abs(ref_f32[:, 3]).max()) > 100          # the LFE (output slot 3) carries signal


@needs_ref
@pytest.mark.parametrize("layout,sr_index,seed", [("stereo", 4, 11), ("5.1", 3, 12), ("mono", 6, 13), ("stereo", 8, 14)])
def test_oracle_matches_reference_on_rich_lc_streams(layout, sr_index, seed):
    """TNS (both directions, orders to 12/7), sine + KBD windows, grouped short blocks and every window-sequence pair
    (a few illegal ones included), on streams from the repo's own AAC-LC writer (SURVEY 8c fixture 3, config 4)."""
    import lc_writer
    import pyref
    pkts, ed, w = lc_writer.make_lc_stream(seed, 36, layout, sr_index)
    rate = [96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000][sr_index]
    nch = lc_writer.CHAN_CONFIG[layout]
    dec = pyref.Decoder(ed, rate, nch, layout, float_mode=True)
    frames = [dec.decode(p) for p in pkts]
    dec.close()
    assert w.stats["tns"] > 0.3 * w.stats["chan_frames"] and w.stats["short"] > 0 and w.stats["sine"] > 0
    assert any(f["tns"]["order"].max() >= (7 if layout != "mono" else 4) for f in frames)
    cfg = abi.make_config(1, pyref.LAYOUTS[layout][0], sr_index, sbr=False, ps=False, max_frames=len(frames))
    st = lambda k: np.stack([f[k] for f in frames])[:, None]
    orc = pyoracle.Oracle(cfg)
    pcm, f32 = orc.synth(st("coeffs"), st("ics"), st("tns"), None, None)
    orc.close()
    ref_f32 = np.stack([f["pcm_f32"] for f in frames])
    # the float criterion is relative to the larger of full scale and the stream's own peak (a hot TNS filter can
    # push float PCM far beyond s16 range, where it clips identically in both decoders)
    scale = max(32768.0, float(np.abs(ref_f32).max()))
    assert float(np.abs(f32[:, 0].astype(np.float64) - ref_f32).max()) / scale <= 1e-5
    assert int(np.abs(pcm[:, 0].astype(np.int32) - np.stack([f["pcm"] for f in frames]).astype(np.int32)).max()) <= 1


@needs_ref
def test_oracle_stage_arrays_match_reference():
    """Stage-by-stage: W, X_low, alpha, X_high, e_curr, gain, Y of the last frame (SURVEY 8c 'stage-level oracle')."""
    import heaac_streams as hs
    pkts, ed, info = hs.make_stream(55, 10, stereo=True)
    frames, d = hs.decode_with_reference(pkts, ed, 24000, 2)
    cfg = abi.make_config(1, [abi.TYPE_CPE], d.sr_index, sbr=True)
    st = lambda k: np.stack([f[k] for f in frames])[:, None]
    orc = pyoracle.Oracle(cfg)
    orc.synth(st("coeffs"), st("ics"), st("tns"), st("sbr"))
    s = frames[-1]["sbr"][0]
    kx, m = int(s["kx"][1]), int(s["m"][1])
    for ch in (0, 1):
        W_ref = d.peek("W", np.float32, (2, 32, 32, 2), ch=ch)
        W = orc.peek("W", np.float32, (2, 32, 32, 2), ch=ch)
        assert np.abs(W - W_ref).max() <= 1e-5 * np.abs(W_ref).max()
        Y_ref = d.peek("Y", np.float32, (2, 38, 64, 2), ch=ch)
        Y = orc.peek("Y", np.float32, (2, 38, 64, 2), ch=ch)
        assert np.abs(Y[1, :32, kx:kx + m] - Y_ref[1, :32, kx:kx + m]).max() <= 2e-5 * np.abs(Y_ref).max()
    # element-level scratch holds the LAST processed channel (1)
    for name, shape, rows in (("X_low", (32, 40, 2), slice(0, kx)), ("X_high", (64, 40, 2), slice(kx, kx + m)),
                              ("alpha0", (64, 2), slice(0, int(s["k0"]))), ("alpha1", (64, 2), slice(0, int(s["k0"])))):
        a_ref = d.peek(name, np.float32, shape, ch=1)[rows]
        a = orc.peek(name, np.float32, shape, ch=1)[rows]
        tol = 1e-3 if name.startswith("alpha") else 2e-5 * np.abs(a_ref).max()
        assert np.abs(a - a_ref).max() <= tol, name
    orc.close()


@needs_ref
def test_reference_output_channel_order():
    """cfg.out_slot (abi.reference_out_slots) reproduces the reference's 5.1 output order FL FR FC LFE BL BR."""
    assert abi.reference_out_slots([abi.TYPE_SCE, abi.TYPE_CPE, abi.TYPE_CPE, abi.TYPE_LFE]) == [2, 0, 1, 4, 5, 3]
    d = load_golden("lc_51_48k")
    orc = pyoracle.Oracle(golden_config(d))
    pcm, f32 = orc.synth(*golden_args(d))
    orc.close()
    for c in range(6):
        assert np.abs(f32[:, 0, c] - d["pcm_f32"][:, c]).max() / 32768.0 <= 1e-5, "channel %d" % c


def _cce_golden():
    import os
    from common import GOLDEN
    d = dict(np.load(os.path.join(GOLDEN, "lc_cce.npz")))
    d["ics"] = d["ics"].view(abi.ICS).reshape(d["coeffs"].shape[:2]) if d["ics"].dtype != abi.ICS else d["ics"]
    return d


def test_oracle_reproduces_the_channel_coupling_golden():
    """Independent channel coupling (apply_channel_coupling + apply_independent_coupling, aacdec.c:1849-1898) on the
    reference-generated golden: SCE + CPE + two coupling channel elements, program_config_element layout."""
    d = _cce_golden()
    et = [int(x) for x in d["elem_types"]]
    assert et.count(abi.TYPE_CCE) == 2 and int(d["cce"]["n_targets"].max()) == 3
    cfg = abi.make_config(1, et, int(d["sr_index"]), max_frames=d["coeffs"].shape[0], out_slots=None)
    orc = pyoracle.Oracle(cfg)
    pcm, f32 = orc.synth(d["coeffs"][:, None], d["ics"][:, None], d["tns"][:, None], None, None, cce=d["cce"][:, None])
    orc.close()
    scale = max(32768.0, float(np.abs(d["pcm_f32"]).max()))
    assert float(np.abs(f32[:, 0].astype(np.float64) - d["pcm_f32"]).max()) / scale <= 1e-5
    assert int(np.abs(pcm[:, 0].astype(np.int32) - d["pcm_s16"]).max()) <= 1
    # the coupling really contributes: without it the PCM is far off
    z = d["cce"].copy(); z["n_targets"] = 0
    orc = pyoracle.Oracle(cfg)
    pcm0, _ = orc.synth(d["coeffs"][:, None], d["ics"][:, None], d["tns"][:, None], None, None, cce=z[:, None])
    orc.close()
    assert int(np.abs(pcm0[:, 0].astype(np.int32) - d["pcm_s16"]).max()) > 100


@needs_ref
@pytest.mark.parametrize("seed", [3, 8])
def test_oracle_matches_reference_on_fresh_coupled_streams(seed):
    import pyref
    sys_path_golden()
    import make_golden
    pk, ed, lay = make_golden.cce_stream(seed, 16)
    dec = pyref.Decoder(ed, 48000, 3, lay, float_mode=True)
    fr = [dec.decode(p) for p in pk]
    dec.close()
    cfg = abi.make_config(1, lay[0], 3, max_frames=len(fr), out_slots=None)
    st = lambda k: np.stack([f[k] for f in fr])[:, None]
    orc = pyoracle.Oracle(cfg)
    pcm, f32 = orc.synth(st("coeffs"), st("ics"), st("tns"), None, None, cce=st("cce"))
    orc.close()
    ref_f32 = np.stack([f["pcm_f32"] for f in fr])
    scale = max(32768.0, float(np.abs(ref_f32).max()))
    assert float(np.abs(f32[:, 0].astype(np.float64) - ref_f32).max()) / scale <= 1e-5
    assert int(np.abs(pcm[:, 0].astype(np.int32) - np.stack([f["pcm"] for f in fr])).max()) <= 1


def sys_path_golden():
    import os
    import sys
    from common import GOLDEN
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
