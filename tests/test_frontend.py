"""The repo's own host front end (heaac_fe_*, csrc/heaac_frontend.c): packets -> C-ABI frame structures.

Pinned (1) against the committed golden captures of what the REFERENCE's parsers handed to spectral_to_sample
(aacdec.c:2078) for the same packets, (2) where oracle/_ref is available, against the reference's parsers on
fresh streams, and (3) end to end: own front end -> oracle / CUDA back end -> PCM vs the reference decoder's
output.  With (3) the reference is not on the path at all."""
import numpy as np
import pytest

import ffmpeg_heaac_b200 as hb
import pyoracle
from common import (GOLDEN_CASES, abi, assert_pcm_close, diff_parsed, golden_config, golden_packets, golden_ref_pcm,
                    load_golden)
from conftest import needs_ref


def parse_golden(d):
    n_ch = d["coeffs"].shape[1]
    out_ch = d["pcm_f32"].shape[1]
    fe = hb.Frontend(d["extradata"].tobytes(), int(d["core_rate"]), out_ch if not d["ps_flag"] else 1)
    frames = [fe.parse(p) for p in golden_packets(d)]
    return fe, frames, n_ch


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_frontend_reproduces_reference_parser_on_golden_packets(name):
    d = load_golden(name)
    fe, frames, n_ch = parse_golden(d)
    assert fe.ncc == n_ch and fe.sr_index == int(d["sr_index"])
    assert fe.elem_types == [int(x) for x in d["elem_types"]]
    assert bool(fe.sbr) == bool(d["sbr_flag"]) and bool(fe.ps) == bool(d["ps_flag"])
    for t, (m, size) in enumerate(zip(frames, d["pkt_sizes"])):
        ref = {k: d[k][t] for k in ("coeffs", "ics", "tns", "sbr", "ps")}
        bad = diff_parsed(ref, m, n_ch, bool(d["sbr_flag"]), bool(d["ps_flag"]))
        assert not bad, "%s frame %d: %s" % (name, t, bad[:6])
        assert m["consumed"] == int(size)
    fe.close()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_frontend_plus_oracle_reproduces_reference_pcm(name):
    d = load_golden(name)
    fe, frames, n_ch = parse_golden(d)
    st = lambda k: np.ascontiguousarray(np.stack([f[k] for f in frames])[:, None])
    cfg = abi.make_config(1, fe.elem_types, fe.sr_index, sbr=bool(fe.sbr), ps=bool(fe.ps), max_frames=len(frames),
                          downsampled=fe.downsampled)
    orc = pyoracle.Oracle(cfg)
    pcm, f32 = orc.synth(st("coeffs"), st("ics"), st("tns"), st("sbr") if fe.sbr else None,
                         st("ps")[:, :, 0] if fe.ps else None)
    orc.close(); fe.close()
    ref_f32, ref_s16 = golden_ref_pcm(d)
    assert_pcm_close(f32[:, 0], ref_f32, pcm[:, 0], ref_s16, name)


def test_frontend_rejects_bad_input():
    fe = hb.Frontend(b"", 0, 0)                      # no extradata: the configuration must come from an ADTS header
    with pytest.raises(RuntimeError):
        fe.parse(bytes([0x21, 0x00, 0x49, 0x90]))    # ... and this is not one (aacdec.c:1994-2006)
    fe.close()
    fe = hb.Frontend(bytes([0x12, 0x10]), 44100, 2)  # AAC-LC 44.1 kHz stereo
    with pytest.raises(RuntimeError):
        fe.parse(b"")                                # empty packet
    with pytest.raises(RuntimeError):
        fe.parse(bytes([0xFF] * 7))                  # element id 7 == END with nothing decoded / overread
    fe.close()


def test_frontend_bias_mode_scales_coefficients():
    """bias_mode front end = the reference's C output path (sf_scale 1/32768 folded into the spectrum, aacdec.c:573-581)."""
    d = load_golden("lc_stereo_44k")
    pk = golden_packets(d)
    a = hb.Frontend(d["extradata"].tobytes(), 44100, 2, bias_mode=False)
    b = hb.Frontend(d["extradata"].tobytes(), 44100, 2, bias_mode=True)
    for p in pk[:4]:
        ca, cb = a.parse(p)["coeffs"], b.parse(p)["coeffs"]
        nz = ca != 0
        assert nz.any() and np.allclose(cb[nz] / ca[nz], 1.0 / 32768.0, rtol=1e-6)
    a.close(); b.close()


@needs_ref
@pytest.mark.parametrize("mode,seed", [("lc", 3), ("51", 4), ("v1", 111), ("v1", 112), ("v2", 211), ("v2", 212), ("v2", 213)])
def test_frontend_matches_reference_parser_on_fresh_streams(mode, seed):
    import heaac_streams as hs
    import pyref
    nfr = 24
    if mode in ("lc", "51"):
        ch, rate, layout = (2, 44100, "stereo") if mode == "lc" else (6, 48000, "5.1")
        pcm = pyref.lcg_noise(1024 * (nfr + 1), ch, seed=seed).astype(np.int32)
        for s0 in (5000, 12000, 13500, 22000):
            pcm[s0:s0 + 300] *= 8                     # attacks: block switching, TNS
        enc = pyref.Encoder(rate, ch, 64000 * ch)
        pkts, ed = enc.encode(np.clip(pcm, -32768, 32767).astype(np.int16))[:nfr], enc.extradata
        enc.close()
    else:
        stereo = mode == "v1"
        pkts, ed, _ = hs.make_stream(seed, nfr, stereo=stereo, ps=mode == "v2")
        ch, rate, layout = (2, 24000, "stereo") if stereo else (1, 24000, "mono")
    ref = pyref.Decoder(ed, rate, ch, layout)
    fe = hb.Frontend(ed, rate, ch)
    for t, p in enumerate(pkts):
        r, m = ref.decode(p), fe.parse(p)
        bad = diff_parsed(r, m, fe.ncc, ref.sbr == 1, ref.ps == 1 and ch == 1)
        assert not bad, "%s/%d frame %d: %s" % (mode, seed, t, bad[:6])
        assert r["consumed"] == m["consumed"]
    ref.close(); fe.close()


@needs_ref
@pytest.mark.parametrize("layout,sr_index,seed", [("stereo", 4, 21), ("5.1", 3, 22), ("mono", 6, 23), ("stereo", 11, 24), ("5.1", 0, 25)])
def test_frontend_matches_reference_parser_on_rich_lc_streams(layout, sr_index, seed):
    """PNS (shared LCG state), intensity stereo, M/S masks, pulse data, escape codes, TNS data, grouping."""
    import lc_writer
    import pyref
    pkts, ed, w = lc_writer.make_lc_stream(seed, 30, layout, sr_index)
    rate = [96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000][sr_index]
    nch = lc_writer.CHAN_CONFIG[layout]
    ref = pyref.Decoder(ed, rate, nch, layout)
    fe = hb.Frontend(ed, rate, nch)
    for t, p in enumerate(pkts):
        r, m = ref.decode(p), fe.parse(p)
        bad = diff_parsed(r, m, fe.ncc, False, False)
        assert not bad, "%s/%d frame %d: %s" % (layout, seed, t, bad[:6])
        assert r["consumed"] == m["consumed"]
    for k in ("pns", "esc", "pulse", "tns"):
        assert w.stats[k] > 0, k
    ref.close(); fe.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_frontend_plus_cuda_reproduces_reference_pcm(name):
    """packets -> own parser -> CUDA back end -> PCM, against the reference decoder's PCM for those packets."""
    d = load_golden(name)
    fe, frames, n_ch = parse_golden(d)
    S = 2
    st = lambda k: np.ascontiguousarray(np.repeat(np.stack([f[k] for f in frames])[:, None], S, axis=1))
    cfg = abi.make_config(S, fe.elem_types, fe.sr_index, sbr=bool(fe.sbr), ps=bool(fe.ps), max_frames=len(frames),
                          downsampled=fe.downsampled)
    ctx = hb.Context(cfg)
    pcm, f32 = ctx.synth(st("coeffs"), st("ics"), st("tns"), st("sbr") if fe.sbr else None,
                         st("ps")[:, :, 0] if fe.ps else None, want_float=True)
    assert ctx.launch_count > 0
    ctx.close(); fe.close()
    ref_f32, ref_s16 = golden_ref_pcm(d)
    for s in range(S):
        assert_pcm_close(f32[:, s], ref_f32, pcm[:, s], ref_s16, "%s stream %d" % (name, s))


def adts_wrap(pkt, sr_index, chan_config, profile=1):
    n = len(pkt) + 7
    bits = (0xFFF << 44) | (0 << 43) | (0 << 41) | (1 << 40) | (profile << 38) | (sr_index << 34) | (0 << 33) | \
           (chan_config << 30) | (0 << 26) | (n << 13) | (0x7FF << 2) | 0
    return bits.to_bytes(7, "big") + pkt


@pytest.mark.parametrize("name,chan_config", [("lc_stereo_44k", 2), ("lc_51_48k", 6)])
def test_frontend_adts_framing_equals_raw_with_extradata(name, chan_config):
    """ADTS-framed input with no extradata (aacdec.c:1994-2006, parse_adts_frame_header 1925-1960)."""
    d = load_golden(name)
    raw = hb.Frontend(d["extradata"].tobytes(), int(d["core_rate"]), chan_config)
    adts = hb.Frontend(b"", 0, 0)
    for p in golden_packets(d):
        a, b = raw.parse(p), adts.parse(adts_wrap(p, int(d["sr_index"]), chan_config))
        assert b["consumed"] == a["consumed"] + 7
        assert adts.ncc == raw.ncc and adts.elem_types == raw.elem_types and adts.sr_index == raw.sr_index
        assert not diff_parsed(a, b, raw.ncc, False, False)
    raw.close(); adts.close()


@pytest.mark.parametrize("name", GOLDEN_CASES + ("bench_heaac_v1_stereo_packets", "bench_heaac_v2_ps_packets",
                                                 "bench_lc_51_48k_tns_packets", "bench_lc_stereo_44k_packets"))
def test_two_phase_parse_equals_one_pass(name):
    """fe_phase_a + fe_phase_b (how the device front end spreads frames over lanes) == heaac_fe_parse, byte for byte,
    including streams with PNS (serial fallback), common_window transitions and looped bench fixtures."""
    import os
    from common import GOLDEN
    if name.startswith("bench_"):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        ofs = np.concatenate([[0], np.cumsum(z["pkt_sizes"].ravel())])
        blob = z["pkt_blob"].tobytes()
        n = z["pkt_sizes"].shape[1]
        pk = [blob[ofs[2 * n + i]:ofs[2 * n + i + 1]] for i in range(n)]        # stream 2 of the fixture
        pk = pk + pk[:16]                                                       # ... looped
        ed, rate, ch = z["extradata"].tobytes(), int(z["core_rate"]), int(z["channels"])
    else:
        d = load_golden(name)
        pk, ed, rate = golden_packets(d), d["extradata"].tobytes(), int(d["core_rate"])
        ch = d["pcm_f32"].shape[1] if not d["ps_flag"] else 1
    a, b = hb.Frontend(ed, rate, ch), hb.Frontend(ed, rate, ch)
    for t, p in enumerate(pk):
        x, y = a.parse(p), b.parse(p, two_phase=True)
        assert x["consumed"] == y["consumed"], (name, t)
        for k in ("coeffs", "ics", "sbr", "ps"):
            assert x[k].tobytes() == y[k].tobytes(), "%s frame %d: %s differs" % (name, t, k)
        assert not diff_parsed(x, y, a.ncc, bool(a.sbr), bool(a.ps)), (name, t)
    a.close(); b.close()


def dump_fuzz_cases(path):
    """cases.bin for tools/fuzz_frontend.c (the ASan / UBSan harness)."""
    import struct
    with open(path, "wb") as f:
        f.write(struct.pack("i", len(GOLDEN_CASES)))
        for name in GOLDEN_CASES:
            d = load_golden(name)
            pk = golden_packets(d)
            ch = d["pcm_f32"].shape[1] if not d["ps_flag"] else 1
            ed = d["extradata"].tobytes()
            f.write(struct.pack("iiii", len(ed), int(d["core_rate"]), ch, len(pk)) + ed)
            for p in pk:
                f.write(struct.pack("i", len(p)) + p)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_frontend_survives_corrupt_packets(name):
    """Byte overwrites, truncations and bit flips: the parser either returns a frame or an error code, and the next
    intact packet still parses (state stays usable).  tools/fuzz_frontend.c is the same loop under ASan/UBSan."""
    d = load_golden(name)
    pk = golden_packets(d)
    ch = d["pcm_f32"].shape[1] if not d["ps_flag"] else 1
    rng = np.random.default_rng(7)
    errors = 0
    for it in range(25):
        fe = hb.Frontend(d["extradata"].tobytes(), int(d["core_rate"]), ch)
        for i, p in enumerate(pk):
            b = bytearray(p)
            mode = int(rng.integers(0, 3))
            if mode == 0:
                for _ in range(int(rng.integers(1, 6))):
                    b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            elif mode == 1:
                b = b[:int(rng.integers(1, len(b)))]
            else:
                b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
            try:
                fe.parse(bytes(b), two_phase=bool(it & 1))
            except hb.HeaacError as e:
                assert e.code in (-22, -38)
                errors += 1
        r = fe.parse(pk[-1])                           # an intact packet afterwards
        assert r["consumed"] == len(pk[-1])
        fe.close()
    assert errors > 0


def test_adts_header_cannot_change_a_fixed_or_running_layout():
    """A batch sizes every stream's output slots for one layout (heaac_decoder.cpp, heaac_parse_gpu.cu).  A packet whose
    ADTS header announces another channel configuration or core rate -- a 5.1 stream in a stereo batch, or a crafted
    header -- must be rejected before anything is written: (a) by a stream whose layout was fixed at open time
    (heaac_fe_fix_layout), even as its very first packet, and (b) by a stream that has already decoded a frame."""
    import ctypes as C
    st, mc = load_golden("lc_stereo_44k"), load_golden("lc_51_48k")
    stereo = [adts_wrap(p, int(st["sr_index"]), 2) for p in golden_packets(st)]
    surround = [adts_wrap(p, int(mc["sr_index"]), 6) for p in golden_packets(mc)]
    probe = hb.Frontend(b"", 0, 0)
    probe.parse(stereo[0])
    assert probe.ncc == 2
    # (a) fixed before the first packet
    f = hb.Frontend(b"", 0, 0)
    assert hb.lib().heaac_fe_fix_layout(C.c_void_p(f.h), C.c_void_p(probe.h)) == 0
    guard = np.full((16, 1024), 7.0, np.float32)
    ics, tns = np.zeros(16, abi.ICS), np.zeros(16, abi.TNS)
    buf = np.frombuffer(surround[0], np.uint8)
    r = hb.lib().heaac_fe_parse(C.c_void_p(f.h), buf.ctypes.data, buf.size, guard.ctypes.data, ics.ctypes.data, tns.ctypes.data, None, None)
    assert r == -22 and (guard == 7.0).all(), "a 5.1 header reached the decode stage of a stereo slot"
    f._refresh()
    assert f.ncc == 2 and f.sr_index == int(st["sr_index"])
    # wrong sampling-rate index, right channel configuration
    bad_rate = adts_wrap(golden_packets(st)[0], 3, 2)
    buf = np.frombuffer(bad_rate, np.uint8)
    assert hb.lib().heaac_fe_parse(C.c_void_p(f.h), buf.ctypes.data, buf.size, guard.ctypes.data, ics.ctypes.data, tns.ctypes.data, None, None) == -22
    assert f.parse(stereo[0])["consumed"] == len(stereo[0])
    f.close()
    # (b) a running stream
    with pytest.raises(hb.HeaacError):
        probe.parse(surround[1])
    assert probe.ncc == 2
    probe.parse(stereo[1])
    probe.close()
    # without a fixed layout the first header still configures the stream (reference behaviour)
    g = hb.Frontend(b"", 0, 0)
    g.parse(surround[0])
    assert g.ncc == 6
    g.close()


# ---------------------------------------------------------------- program config elements and coupling channel elements
def _cce_golden():
    from test_oracle_vs_ref import _cce_golden as g
    return g()


def test_frontend_parses_a_pce_layout_with_coupling_elements():
    """AudioSpecificConfig with a program_config_element (decode_pce, aacdec.c:280-340) listing SCE + CPE + two coupling
    channel elements, and the CCEs themselves (decode_cce, aacdec.c:1503-1568), against the committed capture of what the
    reference's parsers left for spectral_to_sample on the same packets: coefficients, ICS and TNS of all five coded
    channels and the resolved coupling lists (target channel, gain) -- bit for bit."""
    d = _cce_golden()
    fe = hb.Frontend(d["extradata"].tobytes())
    assert fe.from_pce and fe.n_cce == 2 and fe.ncc == 5 and fe.sr_index == int(d["sr_index"])
    assert fe.elem_types == [int(x) for x in d["elem_types"]]
    for t, p in enumerate(golden_packets(d)):
        m = fe.parse(p)
        ref = {k: d[k][t] for k in ("coeffs", "ics", "tns")}
        bad = diff_parsed(ref, m, 5, False, False)
        assert not bad, "frame %d: %s" % (t, bad[:6])
        assert m["cce"].tobytes() == d["cce"][t].tobytes()
        assert m["consumed"] == int(d["pkt_sizes"][t])
    fe.close()


def test_frontend_pce_in_band_equals_pce_in_extradata():
    """ADTS with channel_configuration 0: the layout arrives as a PCE inside the first raw_data_block (aacdec.c:2012-2030).
    Same frames as with the PCE in the extradata; a PCE repeated in later frames is stepped over."""
    import lc_writer
    d = _cce_golden()
    et = [int(x) for x in d["elem_types"]]
    pce = lc_writer.pce_element(int(d["sr_index"]), et)
    a, b = hb.Frontend(d["extradata"].tobytes()), hb.Frontend(b"", 0, 0)
    assert b.ncc == 0
    for t, p in enumerate(golden_packets(d)):
        ma = a.parse(p)
        mb = b.parse(adts_wrap((pce if t % 3 == 0 else b"") + p, int(d["sr_index"]), 0))
        assert b.from_pce and b.elem_types == et
        for k in ("coeffs", "ics", "tns", "cce"):
            assert ma[k].tobytes() == mb[k].tobytes(), (t, k)
        assert mb["consumed"] == 7 + (len(pce) if t % 3 == 0 else 0) + len(p)
    a.close(); b.close()
    c = hb.Frontend(b"", 48000, 0)
    with pytest.raises(hb.HeaacError):                  # a channel element before any configuration
        c.parse(adts_wrap(golden_packets(d)[0], 3, 0))
    c.close()


def test_frontend_rejects_dependent_coupling():
    """Only independently switched coupling is inside the ABI: ind_sw_cce_flag = 0 -> HEAAC_E_UNSUPPORTED."""
    d = _cce_golden()
    fe = hb.Frontend(d["extradata"].tobytes())
    import lc_writer
    et = [int(x) for x in d["elem_types"]]
    w = lc_writer.LcWriter(1, 3, et)
    w.set_coupling({2: [(0, 2)], 3: [(1, 0)]})
    w.cce_ind_sw = 0
    with pytest.raises(hb.HeaacError) as e:
        fe.parse(w.frame())
    assert e.value.code == -38                        # HEAAC_E_UNSUPPORTED
    fe.close()


@needs_ref
@pytest.mark.parametrize("seed,layout,coupling", [
    (31, [0, 1, 3, 2], {3: [(1, 1), (0, 2)]}),                         # SCE CPE LFE + one CCE: right channel only, then the SCE
    (32, [1, 1, 2, 2, 2], {2: [(0, 3), (1, 3)], 3: [(0, 0)], 4: [(1, 2), (0, 1)]}),   # two CPEs, three CCEs, every ch_select
    (33, [0, 2], {1: [(0, 2)]}),
])
def test_frontend_matches_reference_parser_on_fresh_coupling_streams(seed, layout, coupling):
    """Other layouts and target selections, against the reference's parsers run here (oracle/_ref): every ch_select value
    (both channels with one gain, left only, right only, both with their own gains), several CCEs onto one target."""
    import lc_writer
    import pyref
    ed, ids, order = lc_writer.pce_config(3, layout)
    w = lc_writer.LcWriter(seed, 3, layout, tns_prob=0.3, short_bias=0.3)
    w.set_coupling(coupling)
    pk = [w.frame() for _ in range(12)]
    lay = ([layout[e] for e in order], [ids[e] for e in order])
    nout = sum(2 if t == abi.TYPE_CPE else 1 for t in layout if t != abi.TYPE_CCE)
    ref = pyref.Decoder(ed, 48000, nout, lay, float_mode=True)
    fe = hb.Frontend(ed)
    assert fe.elem_types == lay[0]
    for t, p in enumerate(pk):
        r, m = ref.decode(p), fe.parse(p)
        bad = diff_parsed(r, m, fe.ncc, False, False)
        assert not bad, "seed %d frame %d: %s" % (seed, t, bad[:6])
        assert r["cce"].tobytes() == m["cce"].tobytes(), (seed, t, r["cce"], m["cce"])
        assert r["consumed"] == m["consumed"]
    ref.close(); fe.close()


def test_frontend_plus_oracle_reproduces_reference_pcm_with_coupling():
    """End to end without the reference on the path: own front end (PCE layout, CCEs) -> oracle back end -> PCM, against the
    reference decoder's PCM of the same packets."""
    d = _cce_golden()
    fe = hb.Frontend(d["extradata"].tobytes())
    frames = [fe.parse(p) for p in golden_packets(d)]
    st = lambda k: np.ascontiguousarray(np.stack([f[k] for f in frames])[:, None])
    cfg = abi.make_config(1, fe.elem_types, fe.sr_index, max_frames=len(frames), out_slots=None)
    fe.close()
    orc = pyoracle.Oracle(cfg)
    pcm, f32 = orc.synth(st("coeffs"), st("ics"), st("tns"), None, None, cce=st("cce"))
    orc.close()
    scale = max(32768.0, float(np.abs(d["pcm_f32"]).max()))
    assert float(np.abs(f32[:, 0].astype(np.float64) - d["pcm_f32"]).max()) / scale <= 1e-5
    assert int(np.abs(pcm[:, 0].astype(np.int32) - d["pcm_s16"]).max()) <= 1


# ---------------------------------------------------------------- a chunk of frames the way the device front end treats it
def _fixture_or_golden(name, stream=2):
    import os
    from common import GOLDEN
    if name.startswith("bench_"):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        ofs = np.concatenate([[0], np.cumsum(z["pkt_sizes"].ravel())])
        blob = z["pkt_blob"].tobytes()
        n = z["pkt_sizes"].shape[1]
        pk = [blob[ofs[stream * n + i]:ofs[stream * n + i + 1]] for i in range(n)]
        return pk + pk[:16], z["extradata"].tobytes(), int(z["core_rate"]), int(z["channels"])
    d = load_golden(name)
    return golden_packets(d), d["extradata"].tobytes(), int(d["core_rate"]), (d["pcm_f32"].shape[1] if not d["ps_flag"] else 1)


def _chunk_vs_frames(pk, ed, rate, ch, chunk, what):
    """heaac_fe_parse_chunk over `chunk` frames at a time == heaac_fe_parse frame by frame (a failed frame: same error code,
    and the stream goes on identically)."""
    a, b = hb.Frontend(ed, rate, ch), hb.Frontend(ed, rate, ch)
    for t0 in range(0, len(pk), chunk):
        part = pk[t0:t0 + chunk]
        y = b.parse_chunk(part)
        for i, p in enumerate(part):
            try:
                x = a.parse(p)
            except hb.HeaacError as e:
                assert int(y["consumed"][i]) == e.code, (what, t0 + i)
                continue
            assert x["consumed"] == int(y["consumed"][i]), (what, t0 + i, x["consumed"], int(y["consumed"][i]))
            for k in ("coeffs", "ics"):
                assert x[k].tobytes() == y[k][i].tobytes(), "%s frame %d: %s differs" % (what, t0 + i, k)
            if a.sbr:
                assert x["sbr"].tobytes() == y["sbr"][i][:len(x["sbr"])].tobytes(), "%s frame %d: sbr differs" % (what, t0 + i)
            if a.ps:
                assert x["ps"].tobytes() == y["ps"][i:i + 1].tobytes(), "%s frame %d: ps differs" % (what, t0 + i)
            yi = {k: y[k][i] for k in ("coeffs", "ics", "tns", "sbr")}
            yi["ps"] = y["ps"][i:i + 1]
            assert not diff_parsed(x, yi, a.ncc, bool(a.sbr), bool(a.ps)), (what, t0 + i)
    a.close(); b.close()


@pytest.mark.parametrize("chunk", [32, 7])
@pytest.mark.parametrize("name", GOLDEN_CASES + ("bench_heaac_v1_stereo_packets", "bench_heaac_v2_ps_packets",
                                                 "bench_lc_51_48k_tns_packets"))
def test_chunk_parse_equals_frame_by_frame(name, chunk):
    """The SBR payloads of a chunk read side by side (first step for all frames against the element states at the chunk's
    start, re-read after a frame that changes them), then folded in frame order == the one-pass parser, byte for byte."""
    pk, ed, rate, ch = _fixture_or_golden(name)
    _chunk_vs_frames(pk, ed, rate, ch, chunk, name)


@needs_ref
@pytest.mark.parametrize("mode,seed", [("v1", 301), ("v1", 302), ("v2", 303), ("v2", 304), ("51", 305)])
def test_chunk_parse_on_long_generated_streams(mode, seed):
    """96 frames of the stream generator (all frame classes, headers that repeat and headers that change the tables in the
    middle of a chunk, coupling on and off, PS with 20 <-> 34 band switches), chunks of 32 and of 13."""
    import heaac_streams as hs
    if mode == "51":
        pk, ed, _ = hs.make_stream_51(seed, 48)
        rate, ch = 24000, 6
    else:
        pk, ed, _ = hs.make_stream(seed, 96, stereo=mode == "v1", ps=mode == "v2")
        rate, ch = 24000, 2 if mode == "v1" else 1
    for chunk in (32, 13):
        _chunk_vs_frames(pk, ed, rate, ch, chunk, "%s/%d/%d" % (mode, seed, chunk))


@pytest.mark.parametrize("name", ("heaac_v1_stereo", "heaac_v2_ps", "bench_heaac_v1_stereo_packets"))
def test_chunk_parse_with_corrupt_packets(name):
    """Bit flips, overwritten bytes and truncation inside a chunk: the same frames fail with the same codes, and every later
    frame comes out as from the one-pass parser (a failed payload must leave the element state the same way in both)."""
    pk, ed, rate, ch = _fixture_or_golden(name)
    rng = np.random.default_rng(11)
    bad = []
    for t, p in enumerate(pk):
        b = bytearray(p)
        mode = int(rng.integers(0, 5)) if t else 4
        if mode == 0:
            for _ in range(int(rng.integers(1, 6))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        elif mode == 1:
            b = b[:int(rng.integers(1, len(b)))]
        elif mode == 2:
            b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        bad.append(bytes(b))
    _chunk_vs_frames(bad, ed, rate, ch, 32, name + " corrupt")
    _chunk_vs_frames(bad, ed, rate, ch, 5, name + " corrupt")


@needs_ref
@pytest.mark.parametrize("stereo", [True, False])
def test_chunk_parse_when_a_header_changes_the_tables_inside_a_chunk(stereo):
    """Three generated streams with different SBR headers (start / stop frequency, crossover band, header period) spliced at
    frames 41 and 70: the frames behind a splice were read side by side with the OLD band counts, so they have to be read again
    once the frame with the new header has gone through the one-pass parser -- and the limiter-bands-only change (a header
    that differs in its second half only) must not trigger that."""
    import heaac_streams as hs
    a, ed, _ = hs.make_stream(401, 41, stereo=stereo, ps=not stereo, start_freq=5, stop_freq=9, header_period=8)
    b, _, _ = hs.make_stream(402, 29, stereo=stereo, ps=not stereo, start_freq=4, stop_freq=7, xover=1, header_period=5)
    c, _, _ = hs.make_stream(403, 30, stereo=stereo, ps=not stereo, start_freq=3, stop_freq=6, header_period=16)
    pk = a + b + c
    for chunk in (32, 16, 9):
        _chunk_vs_frames(pk, ed, 24000, 2 if stereo else 1, chunk, "spliced/%d" % chunk)


@needs_ref
def test_pulse_beyond_the_coded_bands_is_dropped_where_the_reference_uses_stale_state():
    """A documented departure on non-conformant input: pulse data positioned in a band behind max_sfb.
    The reference scales such a pulse by whatever an EARLIER frame left in the channel's band_type / sf arrays for that band
    (aacdec.c:1226: frame 0 below codes all bands, so frame 1's stray pulse picks up frame 0's scale factor and comes out
    non-zero); every parser here drops it.  Everything else of the frame -- all other coefficients, the side info, the bytes
    consumed -- is the reference's."""
    import lc_writer
    import pyref
    w = lc_writer.LcWriter(77, 4, [abi.TYPE_SCE], tns_prob=0.0, short_bias=0.0)
    w.pulses = False
    nswb = 49                                                    # 44.1 kHz long windows
    w.force_max_sfb = nswb
    p0 = w.frame()                                               # frame 0: every band coded
    w.force_max_sfb, w.overrun_pulse = 20, True
    p1 = w.frame()                                               # frame 1: 20 bands, one pulse in band 20
    at = w.overrun_at
    ed = lc_writer.audio_specific_config(4, 1)
    ref = pyref.Decoder(ed, 44100, 1, "mono")
    fe = hb.Frontend(ed, 44100, 1)
    r0, m0 = ref.decode(p0), fe.parse(p0)
    assert not diff_parsed(r0, m0, 1, False, False)
    r1, m1 = ref.decode(p1), fe.parse(p1)
    assert r1["consumed"] == m1["consumed"]
    assert m1["coeffs"][0, at] == 0.0                            # dropped here ...
    assert r1["coeffs"][0, at] != 0.0                            # ... scaled by frame 0's scale factor of that band in the reference
    rest = np.ones(1024, bool); rest[at] = False
    assert np.array_equal(r1["coeffs"][0, rest], m1["coeffs"][0, rest])
    assert r1["ics"].tobytes() == m1["ics"].tobytes()
    ref.close(); fe.close()


@pytest.mark.parametrize("name", ("bench_heaac_v1_stereo_packets", "bench_heaac_v2_ps_packets"))
def test_chunk_parse_takes_the_side_by_side_path_in_steady_state(name):
    """Performance semantics of the device front end's chunk routine, checked on the host: once a stream is under way, every
    chunk is ONE pass of the side-by-side step and no frame goes through the one-pass parser (repeated, unchanged SBR headers
    every 16 frames must not count as a change); only the chunk that holds the stream's very first header needs a second pass."""
    pk, ed, rate, ch = _fixture_or_golden(name)
    fe = hb.Frontend(ed, rate, ch)
    first = fe.parse_chunk(pk[:32])["stats"]
    assert first["one_pass_frames"] == 1 and first["passes"] <= 2, first
    for rep in range(3):
        for t0 in (32, 0):
            st = fe.parse_chunk(pk[t0:t0 + 32])["stats"]
            assert st == {"passes": 1, "one_pass_frames": 0, "stale": 0}, (rep, t0, st)
    fe.close()
