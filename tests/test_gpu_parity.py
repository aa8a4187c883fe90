"""Parity tests proper (-m gpu): the CUDA back end, called through the C-ABI, against
(1) the committed golden vectors of the reference decoder, (2) the oracle on the same seeded inputs,
(3) the reference itself (oracle/_ref travels to the GPU box as a built .so) on fresh streams, and
(4) size-independent properties at the BASELINE batch sizes.
Tolerance (north star): float PCM max abs error <= 1e-5 of full scale, s16 within +-1 LSB.  In the exact
(-fmad=false) build the CUDA path evaluates the oracle's IEEE operations in the same order, so the oracle
comparison is bit-exact; the tests assert the tolerance and report exactness."""
import ctypes
import os

import numpy as np
import pytest

import pyoracle
from common import (FULL_SCALE, GOLDEN_CASES, TOL_FS, abi, assert_pcm_close, golden_args, golden_config, golden_ref_pcm,
                    golden_packets, load_golden)
from conftest import needs_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import ffmpeg_heaac_b200 as m
    m.lib()
    return m


def run_cuda(hb, cfg, args, want_float=True):
    ctx = hb.Context(cfg)
    pcm, f32 = ctx.synth(*args, want_float=want_float)
    n = ctx.launch_count
    ctx.close()
    assert n > 0, "no kernel was launched"
    return pcm, f32


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("bias_mode", [False, True])
def test_cuda_reproduces_reference_golden_vectors(hb, name, bias_mode):
    d = load_golden(name)
    pcm, f32 = run_cuda(hb, golden_config(d, n_streams=3, bias_mode=bias_mode), golden_args(d, n_streams=3, bias_mode=bias_mode))
    ref_f32, ref_s16 = golden_ref_pcm(d)
    for s in range(3):
        if bias_mode:
            assert int(np.abs(pcm[:, s].astype(np.int32) - ref_s16).max()) <= 1
        else:
            assert_pcm_close(f32[:, s], ref_f32, pcm[:, s], ref_s16, "%s stream %d" % (name, s))


@pytest.mark.parametrize("name", ["lc_stereo_44k", "heaac_v1_stereo", "heaac_v2_ps", "lc_51_48k_tns"])
def test_cuda_matches_oracle_on_synthetic_workloads(hb, name):
    from ffmpeg_heaac_b200 import synthetic
    cfg, a = synthetic.make_workload(name, 24, 12, seed=41, K=24)
    args = (a["coeffs"], a["ics"], a["tns"], a["sbr"], a["ps"])
    pcm, f32 = run_cuda(hb, cfg, args)
    orc = pyoracle.Oracle(cfg)
    opcm, of32 = orc.synth(*args, n_threads=8)
    orc.close()
    err = float(np.abs(f32 - of32).max()) / FULL_SCALE
    print("%s: max float err %.3g FS, bit-exact s16: %s" % (name, err, np.array_equal(pcm, opcm)))
    # the CUDA path evaluates the oracle's IEEE operations in the oracle's order: not a tolerance, identity (a contracted
    # multiply-add anywhere on the path shows up here as 1e-7 FS)
    assert np.array_equal(f32, of32) and np.array_equal(pcm, opcm), "%s: max float err %.3g FS" % (name, err)


@needs_ref
@pytest.mark.parametrize("mode,seeds", [("v1", (301, 302, 303)), ("v2", (401, 402, 403, 404))])
def test_cuda_matches_reference_decoder_on_fresh_streams(hb, mode, seeds):
    """Unmodified reference front end + its own back end vs the same front end + the CUDA back end."""
    import heaac_streams as hs
    stereo, ps = mode == "v1", mode == "v2"
    per_stream = []
    for seed in seeds:
        pkts, ed, info = hs.make_stream(seed, 20, stereo=stereo, ps=ps)
        frames, d = hs.decode_with_reference(pkts, ed, 24000, 2 if stereo else 1)
        per_stream.append(frames)
    st = lambda k: np.stack([np.stack([f[k] for f in fr]) for fr in per_stream], axis=1)
    cfg = abi.make_config(len(seeds), [abi.TYPE_CPE] if stereo else [abi.TYPE_SCE], d.sr_index, sbr=True, ps=ps, max_frames=20)
    pcm, f32 = run_cuda(hb, cfg, (st("coeffs"), st("ics"), st("tns"), st("sbr"), st("ps")[:, :, 0] if ps else None))
    assert_pcm_close(f32, st("pcm_f32"), pcm, st("pcm"), mode)


def test_eight_short_and_window_shape_transitions(hb):
    """All four window sequences with both window shapes on consecutive frames."""
    from ffmpeg_heaac_b200 import synthetic
    cfg, a = synthetic.make_workload("lc_51_48k_tns", 16, 24, seed=77, K=16)
    seqs = set(a["ics"]["window_sequence"].ravel().tolist())
    assert seqs == {0, 1, 2, 3} and a["ics"]["tns_present"].mean() > 0.3
    args = (a["coeffs"], a["ics"], a["tns"], None, None)
    pcm, f32 = run_cuda(hb, cfg, args)
    orc = pyoracle.Oracle(cfg)
    opcm, of32 = orc.synth(*args, n_threads=8)
    orc.close()
    assert np.array_equal(f32, of32), "max float err %.3g FS" % (float(np.abs(f32 - of32).max()) / FULL_SCALE)      # identity, not a tolerance


def test_state_continuity_and_reset(hb):
    """T frames in one call == T single-frame calls; reset_state returns to the initial state."""
    from ffmpeg_heaac_b200 import synthetic
    cfg, a = synthetic.make_workload("heaac_v2_ps", 6, 8, seed=5, K=6)
    ctx = hb.Context(cfg)
    p_all, _ = ctx.synth(a["coeffs"], a["ics"], None, a["sbr"], a["ps"])
    ctx.reset_state()
    for t in range(8):
        p, _ = ctx.synth(a["coeffs"][t:t + 1], a["ics"][t:t + 1], None, a["sbr"][t:t + 1], a["ps"][t:t + 1])
        assert np.array_equal(p[0], p_all[t]), "frame %d" % t
    ctx.reset_state()
    p_again, _ = ctx.synth(a["coeffs"], a["ics"], None, a["sbr"], a["ps"])
    ctx.close()
    assert np.array_equal(p_again, p_all)


def test_device_resident_path_equals_host_path(hb):
    from ffmpeg_heaac_b200 import synthetic
    cfg, a = synthetic.make_workload("heaac_v1_stereo", 16, 6, seed=8, K=16)
    ctx = hb.Context(cfg)
    p_host, _ = ctx.synth(a["coeffs"], a["ics"], None, a["sbr"], None)
    ctx.reset_state()
    bufs = {k: ctx.alloc(a[k].nbytes).upload(a[k]) for k in ("coeffs", "ics", "sbr")}
    out = ctx.alloc(p_host.nbytes)
    ctx.synth_device(6, bufs["coeffs"], bufs["ics"], out, sbr=bufs["sbr"])
    ctx.sync()
    p_dev = out.download(np.int16, p_host.shape)
    ctx.close()
    assert np.array_equal(p_dev, p_host)


@pytest.mark.parametrize("name,S", [("heaac_v1_stereo", 1024), ("heaac_v2_ps", 4096), ("lc_51_48k_tns", 2048)])
def test_full_batch_properties(hb, name, S):
    """BASELINE batch sizes: (a) streams are independent -- a batch built by tiling K distinct streams gives
    identical PCM for identical streams wherever they sit in the batch; (b) a sample of streams matches the oracle;
    (c) determinism across two runs (checksum of checksums)."""
    from ffmpeg_heaac_b200 import synthetic
    K, T = 32, 4
    cfg, a = synthetic.make_workload(name, S, T, seed=99, K=K)
    rep = np.random.default_rng(1).standard_normal((T, K) + a["coeffs"].shape[2:]).astype(np.float32) * 150
    if cfg.sbr:
        rep[..., 640:] = 0
    a["coeffs"] = np.ascontiguousarray(np.tile(rep, (1, S // K, 1, 1)))      # identical spectra for tiled copies
    args = (a["coeffs"], a["ics"], a["tns"], a["sbr"], a["ps"])
    pcm, _ = run_cuda(hb, cfg, args, want_float=False)
    assert np.array_equal(pcm[:, :K], pcm[:, S - K:]), "stream results depend on batch position"
    c1 = int(np.bitwise_xor.reduce(pcm.astype(np.int64).sum(axis=(0, 2, 3))))
    pcm2, _ = run_cuda(hb, cfg, args, want_float=False)
    assert c1 == int(np.bitwise_xor.reduce(pcm2.astype(np.int64).sum(axis=(0, 2, 3)))), "non-deterministic"
    sub = abi.make_config(K, [cfg.elem_type[i] for i in range(cfg.n_elems)], cfg.sr_index, sbr=bool(cfg.sbr), ps=bool(cfg.ps))
    orc = pyoracle.Oracle(sub)
    cut = lambda x: None if x is None else np.ascontiguousarray(x[:, :K])
    opcm, _ = orc.synth(*[cut(x) for x in args], n_threads=8, want_float=False)
    orc.close()
    assert np.array_equal(pcm[:, :K], opcm)


def test_frames_per_launch_do_not_matter(hb):
    """k_chain walks through all frames of a call inside one launch (coefficients of frame f + 1 fetched by a bulk copy
    while frame f is processed, synthesis history kept in shared memory): T frames in one call, in calls of 5 + 7 frames
    and one frame at a time must give identical PCM for every stream."""
    from ffmpeg_heaac_b200 import synthetic
    T, S = 12, 300
    cfg, a = synthetic.make_workload("heaac_v1_stereo", S, T, seed=5, K=32)
    cut = lambda lo, hi: (a["coeffs"][lo:hi], a["ics"][lo:hi], None, a["sbr"][lo:hi], None)
    ctx = hb.Context(cfg)
    whole, _ = ctx.synth(*cut(0, T))
    for split in ([5, 7], [1] * T):
        ctx.reset_state()
        parts, lo = [], 0
        for n in split:
            p, _ = ctx.synth(*cut(lo, lo + n))
            parts.append(p); lo += n
        assert np.array_equal(np.concatenate(parts), whole), split
    ctx.close()
    assert whole.any()


def test_pure_upsampling_and_ps_off(hb):
    """sbr->start == 0 (bad header: 'pure upsampling mode', aacsbr.c:1029-1033) and ps->start == 0
    (R = copy of L, aacsbr.c:1751-1756)."""
    from ffmpeg_heaac_b200 import synthetic
    cfg, a = synthetic.make_workload("heaac_v2_ps", 4, 6, seed=2, K=4)
    a["sbr"]["start"] = 0
    a["sbr"]["kx"] = 32
    a["sbr"]["m"] = 0
    a["ps"]["start"] = 0
    args = (a["coeffs"], a["ics"], None, a["sbr"], a["ps"])
    pcm, f32 = run_cuda(hb, cfg, args)
    orc = pyoracle.Oracle(cfg)
    opcm, of32 = orc.synth(*args)
    orc.close()
    assert np.array_equal(f32, of32), "max float err %.3g FS" % (float(np.abs(f32 - of32).max()) / FULL_SCALE)      # identity, not a tolerance
    assert np.array_equal(pcm[..., 0], pcm[..., 1])


def test_synth_rejects_bad_batches(hb):
    cfg = abi.make_config(2, [abi.TYPE_CPE], 6, sbr=True)
    ctx = hb.Context(cfg)
    b = abi.Batch()
    assert hb.lib().heaac_b200_synth(ctx.h, ctypes.byref(b)) == -22          # no frames, no pointers
    b.n_frames = 1
    co = np.zeros((1, 2, 2, 1024), np.float32); ics = np.zeros((1, 2, 2), abi.ICS); out = np.zeros((1, 2, 2048, 2), np.int16)
    b.coeffs, b.ics, b.pcm = co.ctypes.data, ics.ctypes.data, out.ctypes.data
    assert hb.lib().heaac_b200_synth(ctx.h, ctypes.byref(b)) == -22          # SBR context without SBR parameters
    ctx.close()


@needs_ref
def test_imdct_shim_matches_ff_imdct_half(hb):
    """heaac_b200_mdct_init / imdct_half / mdct_end against the reference's ff_mdct_init / ff_imdct_half."""
    import pyref
    rng = np.random.default_rng(3)
    for nbits, scale in ((11, 1.0), (8, 1.0), (7, 1.0 / 64), (7, -2.0)):
        n = 1 << nbits
        x = rng.standard_normal(n // 2).astype(np.float32)
        m = pyref.lib().ref_mdct_init(nbits, 1, scale)
        ref = np.zeros(n // 2, np.float32)
        pyref.lib().ref_imdct_half(m, ref.ctypes.data_as(ctypes.c_void_p), x.ctypes.data_as(ctypes.c_void_p))
        pyref.lib().ref_mdct_end(m)
        t = hb.Mdct(nbits, 1, scale)
        got = t.imdct_half(x)
        t.close()
        assert np.abs(got - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())


class SbrShim:
    """ctypes wrapper of the ff_sbr_apply / ff_ps_apply shims (heaac_b200_sbr_*)."""

    def __init__(self, hb, id_aac, sr_index, ps):
        import ctypes as C
        self.L, self.C = hb.lib(), C
        h = C.c_void_p()
        r = self.L.heaac_b200_sbr_init(C.byref(h), id_aac, sr_index, int(ps), -1.0 / 1024.0, 0.0)
        assert r == 0, r
        self.h = h

    def sbr_apply(self, elem, psf, L, R):
        r = self.L.heaac_b200_sbr_apply(self.h, elem.ctypes.data, psf.ctypes.data if psf is not None else None,
                                        L.ctypes.data, R.ctypes.data if R is not None else None)
        assert r == 0, r

    def ps_apply(self, psf, L, R, top):
        r = self.L.heaac_b200_ps_apply(self.h, psf.ctypes.data, L.ctypes.data, R.ctypes.data, top)
        assert r == 0, r

    def close(self):
        self.L.heaac_b200_sbr_close(self.h)


@pytest.mark.parametrize("name", ["heaac_v1_stereo", "heaac_v2_ps"])
def test_sbr_apply_shim_matches_reference(hb, name):
    """heaac_b200_sbr_apply == ff_sbr_apply (aacsbr.c:1716-1771): fed with the core decoder's time-domain output and the
    frame's side info, frame by frame, it reproduces the reference decoder's float PCM and the batched path exactly."""
    d = load_golden(name)
    T = d["coeffs"].shape[0]
    et = [int(x) for x in d["elem_types"]]
    ps = bool(d["ps_flag"])
    lc = abi.make_config(1, et, int(d["sr_index"]), sbr=False, ps=False, max_frames=T)
    args = golden_args(d)
    _, ret = run_cuda(hb, lc, (args[0], args[1], args[2], None, None))       # [T, 1, ncc, 1024]: the elements' ret[]
    _, full = run_cuda(hb, golden_config(d), args)                           # batched path, float PCM [T, 1, nout, 2048]
    shim = SbrShim(hb, et[0], int(d["sr_index"]), ps)
    ref_f32, _ = golden_ref_pcm(d)
    for t in range(T):
        L = np.zeros(2048, np.float32); R = np.zeros(2048, np.float32)
        L[:1024] = ret[t, 0, 0]
        if ret.shape[2] > 1:
            R[:1024] = ret[t, 0, 1]
        shim.sbr_apply(np.ascontiguousarray(d["sbr"][t, 0:1]), np.ascontiguousarray(d["ps"][t, 0:1]) if ps else None, L, R)
        out = np.stack([L, R])
        assert np.array_equal(out, full[t, 0]), "frame %d: shim differs from the batched path" % t
        assert float(np.abs(out.astype(np.float64) - ref_f32[t]).max()) / FULL_SCALE <= TOL_FS
    shim.close()


def test_ps_apply_shim_matches_batched_path(hb):
    """heaac_b200_ps_apply == ff_ps_apply (aacps.c:973-992) on the reference's own array layout L/R[2][38][64]: its
    output equals what k_ps leaves in the batched path's X buffers, frame by frame."""
    d = load_golden("heaac_v2_ps")
    T = d["coeffs"].shape[0]
    XS, PLANE = 40, 64 * 40
    args = golden_args(d)
    pre = hb.Context(abi.make_config(1, [abi.TYPE_SCE], int(d["sr_index"]), sbr=True, ps=False, max_frames=1, keep_intermediates=True))
    post = hb.Context(golden_config(d, max_frames=1))
    shim = SbrShim(hb, abi.TYPE_SCE, int(d["sr_index"]), True)
    for t in range(T):
        one = [a[t:t + 1] if a is not None else None for a in args]
        pre.synth(one[0], one[1], one[2], one[3], None)
        post.synth(*one)
        x = pre.debug_read("X", np.float32, (1, 2, 64, XS))                 # before PS: [ch 0][plane][band][slot]
        y = post.debug_read("X", np.float32, (2, 2, 64, XS))                # after PS:  [ch][plane][band][slot]
        L = np.ascontiguousarray(x[0, :, :, :38].transpose(0, 2, 1))        # -> [plane][slot 38][band 64]
        R = np.zeros_like(L)
        e = d["sbr"][t, 0]
        shim.ps_apply(np.ascontiguousarray(d["ps"][t, 0:1]), L, R, int(e["kx"][1]) + int(e["m"][1]))
        for ch, got in ((0, L), (1, R)):
            want = y[ch, :, :, :32].transpose(0, 2, 1)
            assert np.array_equal(got[:, :32], want), "frame %d channel %d" % (t, ch)
    pre.close(); post.close(); shim.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_live_dropin_inside_the_reference_decoder(name):
    """the UNMODIFIED reference decoder (oracle/_ref: its own parsers, its own output stage) with the call
    `spectral_to_sample(ac)` (aacdec.c:2078) replaced, in process, by heaac_b200_synth on this GPU: the decoder's PCM for the
    golden packets must equal what the reference produced on its own (the golden file), frame by frame."""
    import pyref
    import ffmpeg_heaac_b200 as hb
    if not pyref.available():
        pytest.skip("oracle/_ref not built")
    d = load_golden(name)
    pk = golden_packets(d)
    et = [int(x) for x in d["elem_types"]]
    layout = {(1,): "stereo", (0,): "mono"}.get(tuple(et), "5.1")
    nch = sum(2 if t == 1 else 1 for t in et)
    ref_f32, ref_s16 = golden_ref_pcm(d)
    for float_mode in (True, False):
        dec = pyref.Decoder(d["extradata"].tobytes(), int(d["core_rate"]), nch, layout, float_mode=float_mode)
        dec.enable_dropin(hb.LIB_PATH)
        worst_f, worst_s = 0.0, 0
        for t, p in enumerate(pk):
            r = dec.decode(p, capture=False)
            worst_s = max(worst_s, int(np.abs(r["pcm"].astype(np.int32) - ref_s16[t].astype(np.int32)).max()))
            if float_mode:
                worst_f = max(worst_f, float(np.abs(r["pcm_f32"].astype(np.float64) - ref_f32[t]).max()) / 32768.0)
        assert dec.dropin_frames() == len(pk), "the back end did not take every frame"
        dec.close()
        print("%s drop-in (%s path): max float err %.3g FS, max s16 diff %d LSB" % (name, "SIMD-scaling" if float_mode else "C bias", worst_f, worst_s))
        assert worst_s <= 1 and worst_f <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("bias_mode", [False, True])
def test_channel_coupling_matches_oracle_and_reference(bias_mode):
    """Independent channel coupling (k_couple; apply_independent_coupling, aacdec.c:1849-1898) on the reference-generated
    golden: bit-exact against the oracle in both scalings, within tolerance of the reference's own PCM, batched."""
    import ffmpeg_heaac_b200 as hb
    from test_oracle_vs_ref import _cce_golden
    d = _cce_golden()
    et = [int(x) for x in d["elem_types"]]
    S = 5
    cfg = abi.make_config(S, et, int(d["sr_index"]), max_frames=d["coeffs"].shape[0], out_slots=None, bias_mode=bias_mode)
    rep = lambda a: np.ascontiguousarray(np.repeat(a[:, None], S, axis=1))
    co = rep(d["coeffs"]) * (np.float32(1.0 / 32768.0) if bias_mode else np.float32(1.0))
    ctx = hb.Context(cfg)
    pcm, f32 = ctx.synth(co, rep(d["ics"]), rep(d["tns"]), None, None, want_float=True, cce=rep(d["cce"]))
    assert ctx.nout == 3 and ctx.ncc == 5
    ctx.close()
    orc = pyoracle.Oracle(cfg)
    opcm, of32 = orc.synth(co, rep(d["ics"]), rep(d["tns"]), None, None, cce=rep(d["cce"]))
    orc.close()
    assert np.array_equal(pcm, opcm) and np.array_equal(f32, of32)
    if not bias_mode:
        scale = max(32768.0, float(np.abs(d["pcm_f32"]).max()))
        for s in range(S):
            assert float(np.abs(f32[:, s].astype(np.float64) - d["pcm_f32"]).max()) / scale <= 1e-5
            assert int(np.abs(pcm[:, s].astype(np.int32) - d["pcm_s16"]).max()) <= 1


@pytest.mark.gpu
def test_live_dropin_with_channel_coupling():
    """The drop-in run (reference front end, this back end, one process) on a program_config_element stream with coupling
    channel elements: the front end's ChannelCoupling is resolved by the harness binding and applied by k_couple."""
    import sys
    import pyref
    import ffmpeg_heaac_b200 as hb
    from common import GOLDEN
    if not pyref.available():
        pytest.skip("oracle/_ref not built")
    from test_oracle_vs_ref import _cce_golden
    d = _cce_golden()
    # the golden holds no packets: regenerate the stream the golden was made from (pure Python, deterministic)
    sys.path.insert(0, GOLDEN)
    import lc_writer
    et = [abi.TYPE_SCE, abi.TYPE_CPE, abi.TYPE_CCE, abi.TYPE_CCE]
    ed, ids, order = lc_writer.pce_config(3, et)
    w = lc_writer.LcWriter(4242, 3, et, tns_prob=0.4, short_bias=0.3)
    w.set_coupling({2: [(0, 2), (1, 3)], 3: [(1, 0)]})
    pk = [w.frame() for _ in range(d["coeffs"].shape[0])]
    dec = pyref.Decoder(ed, 48000, 3, ([et[e] for e in order], [ids[e] for e in order]), float_mode=True)
    dec.enable_dropin(hb.LIB_PATH)
    scale = max(32768.0, float(np.abs(d["pcm_f32"]).max()))
    for t, p in enumerate(pk):
        r = dec.decode(p, capture=False)
        assert int(np.abs(r["pcm"].astype(np.int32) - d["pcm_s16"][t]).max()) <= 1
        assert float(np.abs(r["pcm_f32"].astype(np.float64) - d["pcm_f32"][t]).max()) / scale <= 1e-5
    assert dec.dropin_frames() == len(pk)
    dec.close()
