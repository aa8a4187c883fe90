"""Synthetic hot-path workloads: valid per-frame parameter sequences for the C-ABI without a bitstream.

bench.py and the full-size property tests need thousands of streams of *parsed* frames
(spectra + ICS/TNS/SBR/PS side info).  This module draws them directly in the parameter domain,
following the semantics the reference's parsers give each field (read_sbr_grid aacsbr.c:609-749,
read_sbr_envelope/noise 787-890, ff_ps_read_data aacps.c:150-279, decode_tns aacdec.c:854-887), with
the frequency tables derived by the library's own host code (heaac_b200_sbr_tables).  Levels follow
SURVEY 8d: dense noise spectra, moderate envelopes, so that SBR stays well conditioned.

K distinct streams are generated and tiled to the requested batch (different streams still carry
different spectra); all K sequences have the same length, so a batch is rectangular.
"""
import numpy as np

from . import abi

_WS_NEXT = {abi.ONLY_LONG: (abi.ONLY_LONG, abi.LONG_START), abi.LONG_START: (abi.EIGHT_SHORT,),
            abi.EIGHT_SHORT: (abi.EIGHT_SHORT, abi.LONG_STOP), abi.LONG_STOP: (abi.ONLY_LONG, abi.LONG_START)}
# tns_tmp2_map values are sin() of quantised angles (ISO 14496-3 4.6.9.3); any |r| < 1 is a legal reflection coefficient
_TNS_MAX_ORDER_LONG, _TNS_MAX_ORDER_SHORT = 12, 7


def coeffs(rng, T, S, ncc, sf_scale=-1.0 / 1024, level=150.0, bins=1024):
    """White spectra, flat up to `bins`, in the scaling convention the context was created with."""
    c = rng.standard_normal((T, S, ncc, 1024), dtype=np.float32) * np.float32(level)
    if bins < 1024:
        c[..., bins:] = 0
    return c * np.float32(sf_scale * -1024.0)


def ics_frames(rng, T, p_short=0.0, tns_prob=0.0, sr_index=3, lfe=False):
    """One channel's window-sequence walk (legal transitions only) + optional TNS filters."""
    ics = np.zeros(T, abi.ICS)
    tns = np.zeros(T, abi.TNS)
    ws = abi.ONLY_LONG
    for t in range(T):
        if t:
            nxt = _WS_NEXT[ws]
            if ws in (abi.ONLY_LONG, abi.LONG_STOP):
                ws = abi.LONG_START if (not lfe and rng.random() < p_short) else abi.ONLY_LONG
            elif ws == abi.EIGHT_SHORT:
                ws = abi.EIGHT_SHORT if rng.random() < 0.5 else abi.LONG_STOP
            else:
                ws = nxt[0]
        is8 = ws == abi.EIGHT_SHORT
        ics[t]["window_sequence"] = ws
        ics[t]["window_shape"] = int(rng.integers(0, 2))
        ics[t]["max_sfb"] = int(rng.integers(10, 15)) if is8 else int(rng.integers(30, 50))
        if rng.random() < tns_prob:
            ics[t]["tns_present"] = 1
            for w in range(8 if is8 else 1):
                nf = int(rng.integers(0, 2)) if is8 else int(rng.integers(1, 4))
                tns[t]["n_filt"][w] = nf
                for f in range(nf):
                    slot = w if is8 else f
                    tns[t]["length"][slot] = int(rng.integers(3, 12)) if is8 else int(rng.integers(8, 24))
                    order = int(rng.integers(1, (_TNS_MAX_ORDER_SHORT if is8 else _TNS_MAX_ORDER_LONG) + 1))
                    tns[t]["order"][slot] = order
                    tns[t]["direction"][slot] = int(rng.integers(0, 2))
                    ang = rng.uniform(-0.3, 0.3, order) * np.linspace(1.0, 0.2, order)
                    tns[t]["coef"][slot][:order] = np.sin(ang * np.pi / 2).astype(np.float32)
    return ics, tns


def _grid(rng, lead, allow_var=True):
    """A legal SBR time grid continuing from the previous frame's trailing border (`lead` = t_E[L] - 16)."""
    while True:
        if lead == 0 and (not allow_var or rng.random() < 0.6):
            fc = 0
        elif lead == 0:
            fc = int(rng.choice([1, 2, 3]))
        else:
            fc = 2 if rng.random() < 0.5 else 3
        ptr = 0
        if fc == 0:
            ne = int(rng.choice([1, 2, 4]))
            step = (16 + (ne >> 1)) // ne
            t = [i * step for i in range(ne)] + [16]
            res = [int(rng.integers(0, 2))] * ne
            trail = 0
        elif fc == 1:
            trail = int(rng.integers(0, 4)); nrel = int(rng.integers(0, 4)); ne = nrel + 1
            t = [0] * (ne + 1); t[ne] = 16 + trail
            for i in range(nrel):
                t[ne - 1 - i] = t[ne - i] - 2 * int(rng.integers(0, 2)) - 2
            res = [int(rng.integers(0, 2)) for _ in range(ne)]
            ptr = int(rng.integers(0, ne + 2))
        elif fc == 2:
            nrel = int(rng.integers(0, 4)); ne = nrel + 1
            t = [lead]
            for i in range(nrel):
                t.append(t[-1] + 2 * int(rng.integers(0, 2)) + 2)
            t.append(16)
            res = [int(rng.integers(0, 2)) for _ in range(ne)]
            ptr = int(rng.integers(0, ne + 2)); trail = 0
        else:
            trail = int(rng.integers(0, 4)); nl = int(rng.integers(0, 3)); nt = int(rng.integers(0, 3)); ne = nl + nt + 1
            t = [0] * (ne + 1); t[0] = lead; t[ne] = 16 + trail
            for i in range(nl):
                t[i + 1] = t[i] + 2 * int(rng.integers(0, 2)) + 2
            for i in range(nt):
                t[ne - 1 - i] = t[ne - i] - 2 * int(rng.integers(0, 2)) - 2
            res = [int(rng.integers(0, 2)) for _ in range(ne)]
            ptr = int(rng.integers(0, ne + 2))
        if all(t[i] < t[i + 1] for i in range(ne)):
            break
    # noise borders and transient envelope, aacsbr.c:720-746
    nn = 2 if ne > 1 else 1
    tq = [t[0], 0, 0]
    tq[nn] = t[ne]
    if nn > 1:
        if fc == 0:
            idx = ne >> 1
        elif fc & 1:
            idx = ne - max(ptr - 1, 1)
        else:
            idx = 1 if ptr == 0 else (ne - 1 if ptr == 1 else ptr - 1)
        tq[1] = t[idx]
    e_a1 = -1
    if (fc & 1) and ptr:
        e_a1 = ne + 1 - ptr
    elif fc == 2 and ptr > 1:
        e_a1 = ptr - 1
    return dict(fc=fc, ne=ne, t=t, res=res, nn=nn, tq=tq, e_a1=e_a1, trail=trail)


def sbr_frames(rng, T, stereo, core_rate=24000, start_freq=5, stop_freq=9, limiter_gains=None, interpol=None,
               smoothing=None, allow_var=True, header_period=16):
    """T consecutive HeaacSbrElem for one element (1 or 2 channels)."""
    from . import sbr_tables
    out = np.zeros(T, abi.SBR_ELEM)
    tab = sbr_tables(core_rate, start_freq, stop_freq)[0]
    lg = int(rng.integers(0, 4)) if limiter_gains is None else limiter_gains
    ip = int(rng.integers(0, 2)) if interpol is None else interpol
    sm = int(rng.integers(0, 2)) if smoothing is None else smoothing
    amp_hdr = int(rng.integers(0, 2))
    nch = 2 if stereo else 1
    prev = [dict(ne=0, t_end=0, res_last=0, e_a1=-1, invf=[0] * 5, trail=0) for _ in range(2)]
    n_low, n_high, n_q = int(tab["n"][0]), int(tab["n"][1]), int(tab["n_q"])
    for f in range(T):
        e = out[f]
        for k in ("k0", "n", "n_q", "n_lim", "f_tablelow", "f_tablehigh", "f_tablenoise", "f_tablelim", "num_patches",
                  "patch_num_subbands", "patch_start_subband"):
            e[k] = tab[k]
        e["start"] = 1
        e["reset"] = 1 if f == 0 else 0
        e["kx"] = (32 if f == 0 else tab["kx"][1], tab["kx"][1])      # kx' starts at 32, aacsbr.c:130
        e["m"] = (0 if f == 0 else tab["m"][1], tab["m"][1])
        e["bs_limiter_gains"], e["bs_interpol_freq"], e["bs_smoothing_mode"] = lg, ip, sm
        coupling = int(stereo and rng.random() < 0.4 and prev[0]["trail"] == prev[1]["trail"])
        e["bs_coupling"] = coupling
        grids = []
        for c in range(nch):
            grids.append(grids[0] if (c == 1 and coupling) else _grid(rng, prev[c]["trail"], allow_var))
        for c in range(nch):
            g, pc, ch = grids[c], prev[c], e["ch"][c]
            ne = g["ne"]
            ch["bs_num_env"], ch["bs_num_noise"] = ne, g["nn"]
            ch["bs_amp_res"] = 0 if (g["fc"] == 0 and ne == 1) else amp_hdr
            ch["t_env"][:ne + 1] = g["t"]
            ch["t_q"][:3] = g["tq"]
            ch["bs_freq_res"][0] = pc["res_last"]
            ch["bs_freq_res"][1:ne + 1] = g["res"]
            ch["t_env_num_env_old"] = pc["t_end"]
            ch["e_a"] = (-(1 if pc["e_a1"] != pc["ne"] else 0), g["e_a1"])
            invf = [int(rng.integers(0, 4)) for _ in range(5)] if not (c == 1 and coupling) else list(e["ch"][0]["bs_invf_mode"][0])
            ch["bs_invf_mode"][0] = invf
            ch["bs_invf_mode"][1] = pc["invf"]
            if rng.random() < 0.25:
                ch["bs_add_harmonic_flag"] = 1
                ch["bs_add_harmonic"][:n_high] = (rng.random(n_high) < 0.125)
            amp = int(ch["bs_amp_res"])
            bal = c == 1 and coupling
            if bal:                               # balance values are stored doubled (delta = 2, aacsbr.c:794)
                lo, hi = (3, 9) if amp else (6, 18)
            else:
                lo, hi = (4, 20) if amp else (8, 40)
            for i in range(ne):
                nb = n_high if g["res"][i] else n_low
                v = np.clip(int(rng.integers(lo, hi + 1)) + np.cumsum(rng.integers(-2, 3, nb)), lo, hi)
                ch["env_q"][i][:nb] = v * (2 if bal else 1)
            for i in range(g["nn"]):
                lo_n, hi_n = (3, 9) if bal else (4, 20)
                v = np.clip(int(rng.integers(lo_n, hi_n + 1)) + np.cumsum(rng.integers(-1, 2, n_q)), lo_n, hi_n)
                ch["noise_q"][i][:n_q] = v * (2 if bal else 1)
            prev[c] = dict(ne=ne, t_end=g["t"][ne], res_last=g["res"][ne - 1], e_a1=g["e_a1"], invf=invf, trail=g["trail"])
    return out


_NR_IIDICC = [10, 20, 34, 10, 20, 34]
_NR_IPDOPD = [5, 11, 17, 5, 11, 17]


def ps_frames(rng, T, header_period=8, ipdopd=True, modes=None):
    """T consecutive HeaacPsFrame for one stream (state carried as ff_ps_read_data would)."""
    out = np.zeros(T, abi.PS_FRAME)
    num_env_prev, is34_prev = 0, 0
    iid_mode = icc_mode = 0
    en_pd = 0
    last = {k: np.zeros(34, np.int8) for k in ("iid", "icc", "ipd", "opd")}
    for f in range(T):
        p = out[f]
        if f % header_period == 0:
            iid_mode, icc_mode = (int(rng.integers(0, 6)), int(rng.integers(0, 6))) if modes is None else modes
            en_pd = int(ipdopd and rng.random() < 0.6)
        nr_iid, nr_icc, nr_pd = _NR_IIDICC[iid_mode], _NR_IIDICC[icc_mode], _NR_IPDOPD[iid_mode]
        is34 = int(nr_iid == 34 or nr_icc == 34)
        p["start"] = 1
        p["iid_quant"], p["nr_iid_par"], p["nr_icc_par"], p["nr_ipdopd_par"] = int(iid_mode > 2), nr_iid, nr_icc, nr_pd
        p["icc_mode"], p["enable_ipdopd"] = icc_mode, en_pd
        p["is34bands"], p["is34bands_old"] = is34, is34_prev
        fc = int(rng.integers(0, 2))
        ne = [[0, 1, 2, 4], [1, 2, 3, 4]][fc][int(rng.integers(0, 4))]
        if fc:
            borders = sorted(rng.choice(np.arange(0, 32), size=ne, replace=False).tolist())
        else:
            borders = [(e * 32 >> {1: 0, 2: 1, 4: 2}.get(ne, 0)) - 1 for e in range(1, ne + 1)]
        lim = 7 + 8 * int(iid_mode > 2)
        last["iid"] = np.clip(last["iid"], -min(lim, 11), min(lim, 11)).astype(np.int8)   # quantiser may have changed
        rows = {k: [] for k in last}
        for e in range(ne):
            rows["iid"].append(np.clip(last["iid"] + rng.integers(-2, 3, 34), -min(lim, 11), min(lim, 11)))
            rows["icc"].append(np.clip(last["icc"] + rng.integers(-1, 2, 34), 0, 7))
            rows["ipd"].append((last["ipd"] + rng.integers(0, 8, 34)) & 7)
            rows["opd"].append((last["opd"] + rng.integers(0, 8, 34)) & 7)
            for k in last:
                last[k] = rows[k][-1].astype(np.int8)
        if ne == 0 or borders[-1] < 31:           # "fake envelope", aacps.c:235-252
            for k in last:
                rows[k].append(last[k])
            borders.append(31)
            ne += 1
        p["num_env"], p["num_env_old"] = ne, num_env_prev
        p["border_position"][0] = -1
        p["border_position"][1:ne + 1] = borders
        for e in range(ne):
            p["iid_par"][e][:nr_iid] = rows["iid"][e][:nr_iid]
            p["icc_par"][e][:nr_icc] = rows["icc"][e][:nr_icc]
            if en_pd:
                p["ipd_par"][e][:nr_pd] = rows["ipd"][e][:nr_pd]
                p["opd_par"][e][:nr_pd] = rows["opd"][e][:nr_pd]
        num_env_prev, is34_prev = ne, is34
    return out


LAYOUTS = {"mono": [abi.TYPE_SCE], "stereo": [abi.TYPE_CPE],
           "5.1": [abi.TYPE_SCE, abi.TYPE_CPE, abi.TYPE_CPE, abi.TYPE_LFE]}

WORKLOADS = {
    # BASELINE.json configs 1..4 as parameter-domain workloads
    "lc_stereo_44k":   dict(layout="stereo", sr_index=4, sbr=False, ps=False, p_short=0.05, tns=0.0),
    "heaac_v1_stereo": dict(layout="stereo", sr_index=6, sbr=True,  ps=False, p_short=0.02, tns=0.0),
    "heaac_v2_ps":     dict(layout="mono",   sr_index=6, sbr=True,  ps=True,  p_short=0.02, tns=0.0),
    "lc_51_48k_tns":   dict(layout="5.1",    sr_index=3, sbr=False, ps=False, p_short=0.35, tns=0.55),
    # HE-AAC 5.1: SBR on the SCE and the two CPEs, the LFE is upsampled only (aacdec.c:1923-1925)
    "heaac_v1_51":     dict(layout="5.1",    sr_index=6, sbr=True,  ps=False, p_short=0.05, tns=0.3),
}


def make_workload(name, S, T, seed=0, K=32, bias_mode=False):
    """Returns (cfg, arrays) with arrays = dict(coeffs [T,S,ncc,1024], ics, tns, sbr, ps) ready for Context.synth."""
    w = WORKLOADS[name]
    et = LAYOUTS[w["layout"]]
    ncc = abi.n_core_channels(et)
    rng = np.random.default_rng(seed)
    K = min(K, S)
    cfg = abi.make_config(S, et, w["sr_index"], sbr=w["sbr"], ps=w["ps"], bias_mode=bias_mode, max_frames=T)
    ics_k = np.zeros((T, K, ncc), abi.ICS)
    tns_k = np.zeros((T, K, ncc), abi.TNS)
    for k in range(K):
        ch = 0
        for t_ in et:
            nch = 2 if t_ == abi.TYPE_CPE else 1
            i_, n_ = ics_frames(rng, T, w["p_short"], w["tns"], w["sr_index"], lfe=t_ == abi.TYPE_LFE)
            for c in range(nch):
                if c == 1 and rng.random() < 0.5:          # independent window decisions half of the time
                    i2, n2 = ics_frames(rng, T, w["p_short"], w["tns"], w["sr_index"])
                    ics_k[:, k, ch + c], tns_k[:, k, ch + c] = i2, n2
                else:
                    ics_k[:, k, ch + c], tns_k[:, k, ch + c] = i_, n_
            ch += nch
    reps = -(-S // K)
    arrays = {"ics": np.ascontiguousarray(np.tile(ics_k, (1, reps, 1))[:, :S]),
              "tns": np.ascontiguousarray(np.tile(tns_k, (1, reps, 1))[:, :S]) if w["tns"] > 0 else None,
              "sbr": None, "ps": None}
    # spectra: independent noise per stream up to 4096 streams; larger batches (config 5's shards) tile 256 streams' worth --
    # the kernels' work does not depend on the values, and drawing 10^9 normals would dominate the run
    Sc = S if S <= 4096 else 256
    c = coeffs(rng, T, Sc, ncc, cfg.sf_scale, bins=1024 if not w["sbr"] else 640)
    arrays["coeffs"] = c if Sc == S else np.ascontiguousarray(np.tile(c, (1, -(-S // Sc), 1, 1))[:, :S])
    if w["sbr"]:
        nse = abi.n_sbr_elems(et)
        sbr_k = np.zeros((T, K, nse), abi.SBR_ELEM)
        for k in range(K):
            for e in range(nse):
                sbr_k[:, k, e] = sbr_frames(rng, T, [t for t in et if t != abi.TYPE_LFE][e] == abi.TYPE_CPE)
        arrays["sbr"] = np.ascontiguousarray(np.tile(sbr_k, (1, reps, 1))[:, :S])
    if w["ps"]:
        ps_k = np.zeros((T, K), abi.PS_FRAME)
        for k in range(K):
            ps_k[:, k] = ps_frames(rng, T)
        arrays["ps"] = np.ascontiguousarray(np.tile(ps_k, (1, reps))[:, :S])
    return cfg, arrays
