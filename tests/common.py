"""Shared helpers for the test-suite."""
import os
import numpy as np

import pyref
import pyoracle

abi = pyref.abi
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ("lc_stereo_44k", "lc_51_48k", "lc_51_tns_rich", "heaac_v1_stereo", "heaac_v2_ps", "heaac_v1_downsampled", "heaac_v1_51")

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
                           bias_mode=bias_mode, max_frames=max_frames or T, downsampled=bool(d.get("downsampled", 0)))


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


def _diff_struct(a, b, path):
    if a.dtype.names:
        bad = []
        for n in a.dtype.names:
            if not n.startswith("reserved"):
                bad += _diff_struct(a[n], b[n], path + "." + n)
        return bad
    return [] if np.array_equal(a, b) else [path]


def diff_parsed(ref, mine, n_chan, sbr, ps):
    """Field-by-field comparison of two parsed frames (dicts of coeffs/ics/tns/sbr/ps as the C-ABI packs them).
    Only the entries the back end reads are compared: table tails past n/n_q/n_lim/num_patches, envelopes past
    bs_num_env and bands past the table length are don't-care (the reference leaves stale data there)."""
    bad = []
    if not np.array_equal(ref["coeffs"], mine["coeffs"]):
        bad.append("coeffs (max diff %g)" % np.abs(ref["coeffs"] - mine["coeffs"]).max())
    bad += _diff_struct(ref["ics"], mine["ics"], "ics")
    for c in range(len(ref["ics"])):          # TNS: only what tns_present / n_filt / order declare valid (stale beyond)
        if not ref["ics"][c]["tns_present"]:
            continue
        a, b = ref["tns"][c], mine["tns"][c]
        short = int(ref["ics"][c]["window_sequence"]) == 2
        nw = 8 if short else 1
        if not np.array_equal(a["n_filt"][:nw], b["n_filt"][:nw]):
            bad.append("tns%d.n_filt" % c)
            continue
        for slot in ([w for w in range(8) if a["n_filt"][w]] if short else range(int(a["n_filt"][0]))):
            o = int(a["order"][slot])
            if a["length"][slot] != b["length"][slot] or o != int(b["order"][slot]):
                bad.append("tns%d[%d].length/order" % (c, slot))
            elif o and (a["direction"][slot] != b["direction"][slot] or not np.array_equal(a["coef"][slot][:o], b["coef"][slot][:o])):
                bad.append("tns%d[%d].coef/direction" % (c, slot))
    if sbr:
        for se in range(len(ref["sbr"])):
            a, b = ref["sbr"][se], mine["sbr"][se]
            lim = {"f_tablelim": int(a["n_lim"]) + 1, "patch_num_subbands": int(a["num_patches"]),
                   "patch_start_subband": int(a["num_patches"]), "f_tablehigh": int(a["n"][1]) + 1,
                   "f_tablelow": int(a["n"][0]) + 1, "f_tablenoise": int(a["n_q"]) + 1}
            for n in a.dtype.names:
                if n in ("ch", "reserved"):
                    continue
                k = lim.get(n)
                if not np.array_equal(a[n][:k] if k is not None else a[n], b[n][:k] if k is not None else b[n]):
                    bad.append("sbr%d.%s" % (se, n))
            if not a["start"]:
                continue
            for c in range(2):
                ca, cb = a["ch"][c], b["ch"][c]
                ne, nn, nq = int(ca["bs_num_env"]), int(ca["bs_num_noise"]), int(a["n_q"])
                if ne == 0 and int(cb["bs_num_env"]) == 0:
                    continue                                   # second channel of a mono element
                pre = "sbr%d.ch%d." % (se, c)
                for n in ("bs_num_env", "bs_num_noise", "bs_amp_res", "bs_add_harmonic_flag", "e_a", "t_env_num_env_old"):
                    if not np.array_equal(ca[n], cb[n]):
                        bad.append(pre + n)
                for n, k in (("bs_freq_res", ne + 1), ("t_env", ne + 1), ("t_q", nn + 1)):
                    if not np.array_equal(ca[n][:k], cb[n][:k]):
                        bad.append(pre + n)
                if not np.array_equal(ca["bs_invf_mode"][:, :nq], cb["bs_invf_mode"][:, :nq]):
                    bad.append(pre + "bs_invf_mode")
                if ca["bs_add_harmonic_flag"] and not np.array_equal(ca["bs_add_harmonic"][:a["n"][1]], cb["bs_add_harmonic"][:a["n"][1]]):
                    bad.append(pre + "bs_add_harmonic")
                for e in range(ne):
                    nb = int(a["n"][ca["bs_freq_res"][e + 1]])
                    if not np.array_equal(ca["env_q"][e][:nb], cb["env_q"][e][:nb]):
                        bad.append(pre + "env_q[%d]" % e)
                for e in range(nn):
                    if not np.array_equal(ca["noise_q"][e][:nq], cb["noise_q"][e][:nq]):
                        bad.append(pre + "noise_q[%d]" % e)
    if ps:
        bad += _diff_struct(ref["ps"], mine["ps"], "ps")
    return bad


def golden_packets(d):
    ofs = np.concatenate([[0], np.cumsum(d["pkt_sizes"])])
    blob = d["pkt_blob"].tobytes()
    return [blob[ofs[i]:ofs[i + 1]] for i in range(len(d["pkt_sizes"]))]
