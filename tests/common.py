"""Shared helpers for the test-suite."""
import os
import numpy as np

import pyref
import pyoracle

abi = pyref.abi
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ("lc_stereo_44k", "lc_51_48k", "heaac_v1_stereo", "heaac_v2_ps")

# north_star tolerance: max abs error <= 1e-5 of full scale on float PCM, or +-1 LSB after s16 conversion
TOL_FS = 1e-5
FULL_SCALE = 32768.0


def load_golden(name):
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    d["ics"] = d["ics"].view(abi.ICS).reshape(d["coeffs"].shape[:2]) if d["ics"].dtype != abi.ICS else d["ics"]
    return d


def golden_config(d, n_streams=1, bias_mode=False, max_frames=None):
    et = [int(x) for x in d["elem_types"]]
    T = d["coeffs"].shape[0]
    return abi.make_config(n_streams, et, int(d["sr_index"]), sbr=bool(d["sbr_flag"]), ps=bool(d["ps_flag"]),
                           bias_mode=bias_mode, max_frames=max_frames or T)


def golden_args(d, n_streams=1, bias_mode=False):
    """Frame-major [T, S, ...] arrays for Context/Oracle.synth, the single golden stream tiled S times."""
    rep = lambda a: np.ascontiguousarray(np.repeat(a[:, None], n_streams, axis=1))
    co = d["coeffs"] * (np.float32(1.0 / 32768.0) if bias_mode else np.float32(1.0))
    sbr = rep(d["sbr"]) if d["sbr_flag"] else None
    ps = rep(d["ps"][:, 0]) if d["ps_flag"] else None
    return rep(co), rep(d["ics"]), rep(d["tns"]), sbr, ps


def golden_ref_pcm(d):
    """(float PCM [T, nout, len] in s16 units, s16 [T, len, nout]) as the reference decoder produced them."""
    return d["pcm_f32"], d["pcm_s16_c_path"]


def assert_pcm_close(f32, ref_f32, s16, ref_s16, what=""):
    err = float(np.abs(f32.astype(np.float64) - ref_f32).max()) / FULL_SCALE
    lsb = int(np.abs(s16.astype(np.int32) - ref_s16.astype(np.int32)).max())
    assert err <= TOL_FS, "%s: float PCM error %.3g FS exceeds %.0e" % (what, err, TOL_FS)
    assert lsb <= 1, "%s: s16 differs by %d LSB" % (what, lsb)
    return err, lsb
