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
                    load_golden)
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
    assert err <= TOL_FS and int(np.abs(pcm.astype(np.int32) - opcm).max()) <= 1
    print("%s: max float err %.3g FS, bit-exact s16: %s" % (name, err, np.array_equal(pcm, opcm)))


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
    assert float(np.abs(f32 - of32).max()) / FULL_SCALE <= TOL_FS


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
    assert int(np.abs(pcm[:, :K].astype(np.int32) - opcm).max()) <= 1


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
    assert float(np.abs(f32 - of32).max()) / FULL_SCALE <= TOL_FS
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
