"""heaac_b200_sbr_tables (own host code) against the reference's sbr_reset (aacsbr.c:1023-1034)."""
import numpy as np
import pytest

import ffmpeg_heaac_b200 as hb
from conftest import needs_ref


def test_default_header_known_answer():
    """24 kHz core, bs_start_freq 5, bs_stop_freq 9, defaults -- values the reference derives."""
    t = hb.sbr_tables(24000, 5, 9)[0]
    assert (int(t["k0"]), int(t["kx"][1]), int(t["m"][1])) == (13, 13, 32)
    assert t["n"].tolist() == [8, 16] and int(t["n_q"]) == 4 and int(t["num_patches"]) == 3
    assert t["f_tablehigh"][:17].tolist() == [13, 14, 15, 16, 17, 18, 19, 20, 22, 24, 26, 28, 31, 34, 37, 41, 45]
    assert t["f_tablelow"][:9].tolist() == [13, 15, 17, 19, 22, 26, 31, 37, 45]
    assert t["f_tablenoise"][:5].tolist() == [13, 17, 22, 31, 45]
    assert t["f_tablelim"][:int(t["n_lim"]) + 1].tolist() == [13, 17, 24, 34, 45]
    assert t["patch_num_subbands"][:3].tolist() == [11, 10, 11] and t["patch_start_subband"][:3].tolist() == [1, 2, 2]


def test_rejects_what_the_reference_rejects():
    with pytest.raises(hb.HeaacError):
        hb.sbr_tables(24000, 0, 15)          # too many QMF subbands
    with pytest.raises(hb.HeaacError):
        hb.sbr_tables(12345, 5, 9)           # unsupported sample rate


@needs_ref
def test_against_reference_probes():
    import heaac_streams as hs
    import pyref
    enc = pyref.Encoder(24000, 2, 64000)
    core = enc.encode(pyref.lcg_noise(2048, 2))[0]
    rng = np.random.default_rng(0)
    accepted = 0
    pyref.lib().ref_set_log_level(-8)
    for _ in range(120):
        sf, st, xo = int(rng.integers(0, 16)), int(rng.integers(0, 16)), int(rng.integers(0, 8))
        e1 = (int(rng.integers(0, 4)), int(rng.integers(0, 2)), int(rng.integers(0, 4))) if rng.random() < 0.6 else None
        e2 = (int(rng.integers(0, 4)), int(rng.integers(0, 4)), 1, 1) if rng.random() < 0.6 else None
        hdr = hs.SbrHeader(rng, sf, st, xo, e1, e2)
        bw = hs.BitWriter(); bw.put(1, 1); hdr.write(bw); bw.put(0, 400)
        d = pyref.Decoder(enc.extradata, 24000, 2, "stereo")
        s = d.decode(hs.splice_fil(core, bw))["sbr"][0]
        d.close()
        fs, al, nb = e1 if e1 else (2, 1, 2)
        lb = e2[0] if e2 else 2
        try:
            mine = hb.sbr_tables(24000, sf, st, xo, fs, al, nb, lb)[0]
        except hb.HeaacError:
            assert s["start"] != 1, "own code rejected a header the reference accepts"
            continue
        assert s["start"] == 1, "own code accepted a header the reference rejects"
        accepted += 1
        for f in ("k0", "n_q", "n", "n_lim", "num_patches"):
            assert np.array_equal(mine[f], s[f]), f
        assert mine["kx"][1] == s["kx"][1] and mine["m"][1] == s["m"][1]
        for f, n in (("f_tablehigh", s["n"][1] + 1), ("f_tablelow", s["n"][0] + 1), ("f_tablenoise", s["n_q"] + 1),
                     ("f_tablelim", s["n_lim"] + 1), ("patch_num_subbands", s["num_patches"]),
                     ("patch_start_subband", s["num_patches"])):
            assert np.array_equal(mine[f][:n], s[f][:n]), f
    pyref.lib().ref_set_log_level(16)
    assert accepted > 20
