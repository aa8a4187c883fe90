"""The batched packets -> PCM decoder (heaac_dec_*, csrc/heaac_decoder.cpp): the counterpart of the reference's
AVCodec entry points aac_decode_init / aac_decode_frame / aac_decode_close (aacdec.c:537, 1962, 2108) for a batch of
streams.  GPU tests compare its PCM with the reference decoder's PCM for the same packets (committed golden
captures); CPU tests cover argument handling and that nothing is decoded without a GPU."""
import ctypes

import numpy as np
import pytest

import ffmpeg_heaac_b200 as hb
from common import GOLDEN_CASES, abi, assert_pcm_close, golden_packets, golden_ref_pcm, load_golden
from test_frontend import adts_wrap


def open_for(d, S, **kw):
    ch = d["pcm_f32"].shape[1] if not d["ps_flag"] else 1
    return hb.Decoder(S, d["extradata"].tobytes(), int(d["core_rate"]), ch, max_frames=len(d["pkt_sizes"]), **kw)


def test_decoder_open_validates_arguments():
    L = hb.lib()
    h = ctypes.c_void_p()
    cfg = abi.DecoderConfig()
    assert L.heaac_dec_open(ctypes.byref(cfg), ctypes.byref(h)) == -22          # abi_version 0
    cfg.abi_version, cfg.n_streams, cfg.max_frames_per_call = abi.ABI_VERSION, 0, 4
    assert L.heaac_dec_open(ctypes.byref(cfg), ctypes.byref(h)) == -22          # no streams
    cfg.n_streams, cfg.extradata_size = 2, 2
    assert L.heaac_dec_open(ctypes.byref(cfg), ctypes.byref(h)) == -22          # size without pointer
    bad = np.asarray([0xF8, 0x00], np.uint8)                                    # AOT 31 -> escape, not AAC-LC
    cfg.extradata = bad.ctypes.data
    assert L.heaac_dec_open(ctypes.byref(cfg), ctypes.byref(h)) < 0
    ok = np.asarray([0x12, 0x10], np.uint8)
    cfg.extradata = ok.ctypes.data
    assert L.heaac_dec_open(ctypes.byref(cfg), ctypes.byref(h)) == 0
    v = ctypes.c_int(-7)
    assert L.heaac_dec_info(h, ctypes.byref(v), None, None, None, None, None) == 0 and v.value == -7   # not configured yet
    L.heaac_dec_close(h)


def test_decoder_has_no_cpu_path():
    """Without a usable GPU the first decode call fails with HEAAC_E_NODEV: there is no host fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d = load_golden("lc_stereo_44k")
    dec = open_for(d, 1)
    with pytest.raises(hb.HeaacError) as e:
        dec.decode(hb.PacketBatch([[p] for p in golden_packets(d)]))
    assert e.value.code in (-19, -5)
    dec.close()


def test_reference_out_slots_matches_python_mirror():
    L = hb.lib()
    T = abi
    for et in ([T.TYPE_SCE], [T.TYPE_CPE], [T.TYPE_SCE, T.TYPE_CPE], [T.TYPE_SCE, T.TYPE_CPE, T.TYPE_SCE],
               [T.TYPE_SCE, T.TYPE_CPE, T.TYPE_CPE], [T.TYPE_SCE, T.TYPE_CPE, T.TYPE_CPE, T.TYPE_LFE],
               [T.TYPE_SCE, T.TYPE_CPE, T.TYPE_CPE, T.TYPE_CPE, T.TYPE_LFE], [T.TYPE_CPE, T.TYPE_CPE]):
        a = np.asarray(et, np.uint8)
        out = np.zeros(16, np.uint8)
        n = L.heaac_b200_reference_out_slots(a.ctypes.data, len(et), out.ctypes.data)
        want = abi.reference_out_slots(et)
        assert n == len(want) and list(out[:n]) == want, et


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("threads", [1, 3, "gpu"])
def test_decoder_reproduces_reference_pcm(name, threads):
    """threads = host parser threads, or "gpu": entropy decoding by the device front end (k_parse)."""
    d = load_golden(name)
    S = 5
    dec = open_for(d, S, want_float=True, **({"gpu_parse": True} if threads == "gpu" else {"n_threads": threads}))
    pk = golden_packets(d)
    pcm, f32, consumed = dec.decode(hb.PacketBatch([[p] * S for p in pk]))
    i = dec.info()
    dec.close()
    ref_f32, ref_s16 = golden_ref_pcm(d)
    assert i["configured"] and i["parse_errors"] == 0 and i["gpu_launches"] > 0
    assert (i["frame_out_len"], i["n_out_channels"]) == ref_s16.shape[1:]
    assert np.array_equal(consumed, np.repeat(d["pkt_sizes"][:, None], S, axis=1))
    for s in range(S):
        assert_pcm_close(f32[:, s], ref_f32, pcm[:, s], ref_s16, "%s stream %d" % (name, s))


@pytest.mark.gpu
def test_decoder_calls_chain_and_adts_input():
    """Two calls of 5 + 7 frames == one call of 12 (state carries over); ADTS framing == raw + extradata."""
    d = load_golden("heaac_v1_stereo")
    pk = golden_packets(d)
    S = 4
    one = open_for(d, S)
    whole, _, _ = one.decode(hb.PacketBatch([[p] * S for p in pk]))
    one.close()
    two = open_for(d, S)
    a, _, _ = two.decode(hb.PacketBatch([[p] * S for p in pk[:5]]))
    b, _, _ = two.decode(hb.PacketBatch([[p] * S for p in pk[5:]]))
    two.close()
    assert np.array_equal(np.concatenate([a, b]), whole)
    adts = hb.Decoder(S, b"", 0, 0, max_frames=len(pk))
    c, _, consumed = adts.decode(hb.PacketBatch([[adts_wrap(p, int(d["sr_index"]), 2)] * S for p in pk]))
    adts.close()
    assert np.array_equal(c, whole) and np.array_equal(consumed[:, 0], d["pkt_sizes"] + 7)


@pytest.mark.gpu
@pytest.mark.parametrize("gpu_parse", [False, True])
def test_decoder_survives_a_corrupt_packet(gpu_parse):
    """A packet that fails to parse is reported through consumed[] and decoded as silence; the other streams of
    the batch are unaffected (the reference returns -1 for that packet only, aacdec.c:2064-2070)."""
    d = load_golden("lc_stereo_44k")
    pk = golden_packets(d)
    S = 3
    rows = [[p] * S for p in pk]
    rows[4] = [pk[4], b"\xff" * 40, pk[4]]
    dec = open_for(d, S, gpu_parse=gpu_parse)
    pcm, _, consumed = dec.decode(hb.PacketBatch(rows))
    i = dec.info()
    dec.close()
    assert consumed[4, 1] < 0 and i["parse_errors"] == 1 and (np.delete(consumed, 4, axis=0) > 0).all()
    _, ref_s16 = golden_ref_pcm(d)
    for s in (0, 2):
        assert int(np.abs(pcm[:, s].astype(np.int32) - ref_s16).max()) <= 1
    assert int(np.abs(pcm[:4, 1].astype(np.int32) - ref_s16[:4]).max()) <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("gpu_parse", [False, True])
def test_decoder_pipelined_submit_wait(gpu_parse):
    """Three calls in flight (heaac_dec_submit / heaac_dec_wait), each with its own output buffers, give the same
    PCM and consumed[] as one synchronous call."""
    d = load_golden("heaac_v2_ps")
    pk = golden_packets(d)
    S = 6
    one = open_for(d, S, gpu_parse=gpu_parse)
    whole, _, cons = one.decode(hb.PacketBatch([[p] * S for p in pk]))
    one.close()
    dec = open_for(d, S, gpu_parse=gpu_parse)
    parts = [pk[0:4], pk[4:8], pk[8:12]]
    first, _, _ = dec.decode(hb.PacketBatch([[p] * S for p in parts[0]]))      # configures; geometry known from here on
    outs = [np.zeros((4, S) + whole.shape[2:], np.int16) for _ in parts[1:]]
    cns = [np.zeros((4, S), np.int32) for _ in parts[1:]]
    tickets = [dec.submit(hb.PacketBatch([[p] * S for p in part]), o, c) for part, o, c in zip(parts[1:], outs, cns)]
    for tk in tickets:
        dec.wait(tk)
    assert dec.info()["parse_errors"] == 0
    dec.close()
    assert np.array_equal(np.concatenate([first] + outs), whole)
    assert np.array_equal(np.concatenate(cns), cons[4:])


@pytest.mark.gpu
def test_two_gpu_parse_decoders_share_the_device_tables():
    """Two decoders with the device front end on one GPU, closed in either order, keep working (the parser tables
    are one per-device copy reached through a __device__ pointer)."""
    d = load_golden("lc_stereo_44k")
    pk = golden_packets(d)
    batch = hb.PacketBatch([[p] * 2 for p in pk])
    a, b = open_for(d, 2, gpu_parse=True), open_for(d, 2, gpu_parse=True)
    pa, _, _ = a.decode(batch)
    pb, _, _ = b.decode(batch)
    b.close()                                           # the second one goes first ...
    c = open_for(d, 2, gpu_parse=True)
    pc, _, _ = c.decode(batch)                          # ... a third comes, the first is still usable
    a2 = open_for(d, 2, gpu_parse=True)
    a.close()
    pa2, _, _ = a2.decode(batch)
    c.close(); a2.close()
    assert np.array_equal(pa, pb) and np.array_equal(pa, pc) and np.array_equal(pa, pa2)
    _, ref_s16 = golden_ref_pcm(d)
    assert int(np.abs(pa[:, 0].astype(np.int32) - ref_s16).max()) <= 1


@pytest.mark.gpu
def test_decoder_full_batch_properties():
    """BASELINE-size batch (1024 HE-AAC v1 stereo streams x 32 frames) through the GPU-parse decoder: every packet
    consumed completely, replicas of a fixture stream give identical PCM, distinct streams differ, and the whole
    batch equals the host-parser decoder's output bit for bit."""
    import os
    from common import GOLDEN
    z = np.load(os.path.join(GOLDEN, "bench_heaac_v1_stereo_packets.npz"))
    sizes, blob = z["pkt_sizes"], z["pkt_blob"].tobytes()
    K, N = sizes.shape
    ofs = np.concatenate([[0], np.cumsum(sizes.ravel())])
    pk = [[blob[ofs[k * N + i]:ofs[k * N + i + 1]] for i in range(N)] for k in range(K)]
    S, T = 1024, 32
    batch = hb.PacketBatch([[pk[s % K][t] for s in range(S)] for t in range(T)])
    ed, rate, ch = z["extradata"].tobytes(), int(z["core_rate"]), int(z["channels"])
    gpu = hb.Decoder(S, ed, rate, ch, max_frames=T, gpu_parse=True)
    a, _, ca = gpu.decode(batch)
    ia = gpu.info()
    gpu.close()
    assert a.shape == (T, S, 2048, 2) and ia["parse_errors"] == 0
    assert np.array_equal(ca, np.asarray([[sizes[s % K][t] for s in range(S)] for t in range(T)]))
    for s in range(K, S, 97):
        assert np.array_equal(a[:, s], a[:, s % K]), "replica %d differs from its source stream" % s
    assert not np.array_equal(a[:, 0], a[:, 1])
    assert int(np.abs(a.astype(np.int32)).max()) > 1000            # real audio, not silence
    host = hb.Decoder(S, ed, rate, ch, max_frames=T)
    b, _, cb = host.decode(batch)
    host.close()
    assert np.array_equal(ca, cb) and np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("gpu_parse", [False, True])
def test_decoder_decodes_a_pce_stream_with_coupling_elements(gpu_parse):
    """Packets -> PCM for a stream whose layout is a program config element with two coupling channel elements (SCE + CPE
    coded, CCEs added in by k_couple): own front end (decode_pce / decode_cce) + CUDA back end against the reference
    decoder's PCM for the same packets.  Such layouts are parsed on the host whatever gpu_parse says."""
    from test_oracle_vs_ref import _cce_golden
    d = _cce_golden()
    S, T = 4, len(d["pkt_sizes"])
    dec = hb.Decoder(S, d["extradata"].tobytes(), 48000, 3, max_frames=T, want_float=True, gpu_parse=gpu_parse, n_threads=2)
    pk = golden_packets(d)
    pcm, f32, consumed = dec.decode(hb.PacketBatch([[p] * S for p in pk]))
    i = dec.info()
    dec.close()
    assert i["configured"] and i["parse_errors"] == 0 and i["gpu_launches"] > 0 and i["k_parse_ms"] == 0
    assert (i["frame_out_len"], i["n_out_channels"]) == (1024, 3)
    assert np.array_equal(consumed, np.repeat(d["pkt_sizes"][:, None], S, axis=1))
    scale = max(32768.0, float(np.abs(d["pcm_f32"]).max()))
    for s in range(S):
        assert float(np.abs(f32[:, s].astype(np.float64) - d["pcm_f32"]).max()) / scale <= 1e-5
        assert int(np.abs(pcm[:, s].astype(np.int32) - d["pcm_s16"]).max()) <= 1
