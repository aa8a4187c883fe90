"""Device build of the front end (csrc/heaac_parse_gpu.cu, FE_DEVICE): things that can differ between the host
and the device compilation of the shared parser source are checked against the host build here; end-to-end
PCM parity of the GPU-parse decoder lives in tests/test_decoder.py."""
import ctypes
import itertools
import os

import numpy as np
import pytest

import ffmpeg_heaac_b200 as hb
from common import GOLDEN, abi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("core_rate", [8000, 11025, 12000, 16000, 22050, 24000, 32000, 44100, 48000])
def test_device_sbr_tables_equal_host_over_the_whole_header_space(core_rate):
    """sbr_reset's table chain uses powf / log2f / lrintf (aacsbr.c:269-285, 303-488): the device libm must land on
    the same integers as the host's for EVERY header (start, stop, xover, freq_scale, alter_scale, noise_bands,
    limiter_bands) -- 262,144 headers per core rate."""
    L = hb.lib()
    ctx = hb.Context(abi.make_config(1, [abi.TYPE_CPE], 6, sbr=True))
    combos = np.asarray(list(itertools.product(range(16), range(16), range(8), range(4), range(2), range(4), range(4))), np.uint8)
    params = np.zeros((len(combos), 8), np.uint8)
    params[:, :7] = combos
    bad = accepted = 0
    for i0 in range(0, len(params), 32768):
        p = np.ascontiguousarray(params[i0:i0 + 32768])
        out = np.zeros(len(p), abi.SBR_ELEM)
        rc = np.zeros(len(p), np.int32)
        r = L.heaac_b200_sbr_tables_device(ctx.h, core_rate, len(p), p.ctypes.data, out.ctypes.data, rc.ctypes.data)
        assert r == 0
        host = np.zeros(1, abi.SBR_ELEM)
        for k in range(len(p)):
            host[...] = 0
            hr = L.heaac_b200_sbr_tables(core_rate, *[int(x) for x in p[k, :7]], host.ctypes.data)
            if hr != rc[k] or (hr == 0 and host.tobytes()[:140] != out[k:k + 1].tobytes()[:140]):
                bad += 1
            accepted += hr == 0
    ctx.close()
    assert bad == 0, "%d of %d headers differ between the device and host table derivation" % (bad, len(params))
    assert accepted > 1000


def test_gpu_parser_equals_host_parser_structures():
    """k_parse's coefficients, ICS, TNS and consumed[] for the bench fixture streams == heaac_fe_parse's (the PCM
    tests could hide a difference below the audio tolerance; this one cannot)."""
    z = np.load(os.path.join(GOLDEN, "bench_heaac_v1_stereo_packets.npz"))
    sizes, blob = z["pkt_sizes"], z["pkt_blob"].tobytes()
    K, N = sizes.shape
    ofs = np.concatenate([[0], np.cumsum(sizes.ravel())])
    pk = [[blob[ofs[k * N + i]:ofs[k * N + i + 1]] for i in range(N)] for k in range(K)]
    T = 24
    ed, rate, ch = z["extradata"].tobytes(), int(z["core_rate"]), int(z["channels"])
    batch = hb.PacketBatch([[pk[s][t] for s in range(K)] for t in range(T)])
    gpu = hb.Decoder(K, ed, rate, ch, max_frames=T, gpu_parse=True)
    host = hb.Decoder(K, ed, rate, ch, max_frames=T, n_threads=2)
    a, _, ca = gpu.decode(batch)
    b, _, cb = host.decode(batch)
    gpu.close(); host.close()
    assert np.array_equal(ca, cb) and (ca > 0).all()
    assert np.array_equal(a, b), "PCM of the GPU-parse and host-parse decoders differs"


@pytest.mark.parametrize("name", ["lc_51_tns_rich", "heaac_v1_stereo", "heaac_v2_ps"])
def test_device_front_end_survives_corrupt_packets(name):
    """Every stream of the batch gets its own corruption of the golden packets (overwrites, truncation, bit flips):
    the call completes, failures come back through consumed[], PCM equals the host-parser decoder's for the same
    corrupted input (both replace a failed frame by silence), and the decoder still decodes intact packets after.
    (The same parser source runs under ASan/UBSan on the host: tools/fuzz_frontend.c.)"""
    from common import golden_packets, load_golden
    d = load_golden(name)
    pk = golden_packets(d)
    ch = d["pcm_f32"].shape[1] if not d["ps_flag"] else 1
    S = 12
    rng = np.random.default_rng(5)
    rows = []
    for p in pk:
        row = []
        for s in range(S):
            b = bytearray(p)
            mode = int(rng.integers(0, 4))
            if mode == 0:
                for _ in range(int(rng.integers(1, 6))):
                    b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            elif mode == 1:
                b = b[:int(rng.integers(1, len(b)))]
            elif mode == 2:
                b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
            row.append(bytes(b))
        rows.append(row)
    rows[0] = [pk[0]] * S                                  # the first packet configures the decoder
    batch, good = hb.PacketBatch(rows), hb.PacketBatch([[p] * S for p in pk])
    out = []
    for gpu_parse in (True, False):
        dec = hb.Decoder(S, d["extradata"].tobytes(), int(d["core_rate"]), ch, max_frames=len(pk), gpu_parse=gpu_parse)
        pcm, _, cons = dec.decode(batch)
        pcm2, _, cons2 = dec.decode(good)
        dec.close()
        assert (cons2 > 0).all()
        out.append((pcm, cons, pcm2))
    assert (out[0][1] < 0).any()
    assert np.array_equal(out[0][1], out[1][1]), "consumed[] differs between the device and the host parser"
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][2], out[1][2])
