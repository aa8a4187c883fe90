"""Synthetic HE-AAC v1/v2 bitstream generator (test fixture code).

AAC-LC core frames come from the reference's own aacenc (through oracle/_ref); this module
splices a FIL element carrying a valid `sbr_extension_data()` (and, for v2, `ps_data()`)
with random parameters in front of each frame's END element.  Syntax follows what the
reference parser reads: ff_decode_sbr_extension / read_sbr_header / read_sbr_grid /
read_sbr_envelope / read_sbr_noise (aacsbr.c:207-259, 609-749, 787-890, 1044-1086),
decode_extension_payload (aacdec.c:1650-1690) and ff_ps_read_data (aacps.c:84-279).
SURVEY.md 8c/8d describe the fixture rules (dense noise core, SBR in the first frame,
SBR data in every frame, moderate envelope levels).
"""
import os
import sys
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(_ROOT, "oracle"))
import pyref  # noqa: E402
abi = pyref.abi


class BitWriter:
    def __init__(self):
        self.bits = []

    def put(self, val, n):
        assert 0 <= val < (1 << n) or n == 0, (val, n)
        for i in range(n - 1, -1, -1):
            self.bits.append((val >> i) & 1)

    def __len__(self):
        return len(self.bits)

    def tobytes(self):
        b = self.bits + [0] * (-len(self.bits) % 8)
        return np.packbits(np.asarray(b, np.uint8)).tobytes()


_HUFF = {}


def huff(name):
    """(codes, bits) of one of the reference's SBR/PS Huffman tables (oracle/ref_tables.c)."""
    if name not in _HUFF:
        bits = pyref.table(name + "_bits", np.uint8)
        raw = pyref.lib()
        import ctypes as C
        p = C.c_void_p()
        n = raw.ref_table((name + "_codes").encode(), C.byref(p))
        width = n // bits.size
        codes = pyref.table(name + "_codes", {1: np.uint8, 2: np.uint16, 4: np.uint32}[width])
        _HUFF[name] = (codes.astype(np.int64), bits.astype(np.int64))
    return _HUFF[name]


def put_huff(bw, table, sym):
    codes, bits = huff(table)
    assert 0 <= sym < codes.size, (table, sym)
    bw.put(int(codes[sym]), int(bits[sym]))


SBR_LAV = {"t_huffman_env_1_5dB": 60, "f_huffman_env_1_5dB": 60, "t_huffman_env_bal_1_5dB": 24,
           "f_huffman_env_bal_1_5dB": 24, "t_huffman_env_3_0dB": 31, "f_huffman_env_3_0dB": 31,
           "t_huffman_env_bal_3_0dB": 12, "f_huffman_env_bal_3_0dB": 12, "t_huffman_noise_3_0dB": 31,
           "t_huffman_noise_bal_3_0dB": 12}   # vlc_sbr_lav, aacsbr.c:73-74
CEIL_LOG2 = [0, 1, 2, 2, 3, 3]                # aacsbr.c:605-607


class SbrHeader:
    def __init__(self, rng, start_freq=5, stop_freq=9, xover=0, extra1=None, extra2=None, amp_res=None):
        self.amp_res = int(rng.integers(0, 2)) if amp_res is None else amp_res
        self.start_freq, self.stop_freq, self.xover = start_freq, stop_freq, xover
        self.extra1 = extra1   # None or (freq_scale, alter_scale, noise_bands)
        self.extra2 = extra2   # None or (limiter_bands, limiter_gains, interpol_freq, smoothing_mode)

    def write(self, bw):
        bw.put(self.amp_res, 1); bw.put(self.start_freq, 4); bw.put(self.stop_freq, 4); bw.put(self.xover, 3)
        bw.put(0, 2)
        bw.put(1 if self.extra1 else 0, 1); bw.put(1 if self.extra2 else 0, 1)
        if self.extra1:
            bw.put(self.extra1[0], 2); bw.put(self.extra1[1], 1); bw.put(self.extra1[2], 2)
        if self.extra2:
            bw.put(self.extra2[0], 2); bw.put(self.extra2[1], 2); bw.put(self.extra2[2], 1); bw.put(self.extra2[3], 1)


class SbrChanState:
    """Encoder-side mirror of what the decoder remembers between frames for delta coding."""

    def __init__(self):
        self.prev_env = None      # last envelope's values (list) and its freq_res
        self.prev_res = 0
        self.prev_noise = None
        self.prev_trail = 0       # abs_bord_trail - 16 of the previous frame


def _gen_grid(rng, st, allow_var, end_at_frame_border=False):
    """One SBR time/frequency grid (read_sbr_grid, aacsbr.c:609-749).
    Returns dict(fc, num_env, t_env, res, pointer, trail, fields=[(value, nbits)...], num_noise).
    end_at_frame_border: only classes whose last border is the frame border (FIXFIX / VARFIX), so that the
    stream may be followed by its own first frame (looped bench fixtures)."""
    lead = st.prev_trail
    while True:
        if lead == 0 and (not allow_var or rng.random() < 0.6):
            fc = 0
        elif end_at_frame_border:
            fc = 2
        elif lead == 0:
            fc = int(rng.choice([1, 2, 3]))
        else:
            fc = 2 if rng.random() < 0.5 else 3
        g = {"fc": fc}
        if fc == 0:                                   # FIXFIX
            ne = int(rng.choice([1, 2, 4]))
            res = [int(rng.integers(0, 2))] * ne
            t_env = [16 * i // ne for i in range(ne + 1)]
            ptr, trail = 0, 0
            fields = [(0, 2), ({1: 0, 2: 1, 4: 2}[ne], 2), (res[0], 1)]
        elif fc == 1:                                 # FIXVAR
            trail = int(rng.integers(0, 4)); nrel = int(rng.integers(0, 4))
            rel = [int(rng.integers(0, 2)) for _ in range(nrel)]
            ne = nrel + 1
            t_env = [0] * (ne + 1)
            t_env[ne] = 16 + trail
            for i in range(nrel):
                t_env[ne - 1 - i] = t_env[ne - i] - 2 * rel[i] - 2
            ptr = int(rng.integers(0, min(ne + 2, 1 << CEIL_LOG2[ne])))
            res = [int(rng.integers(0, 2)) for _ in range(ne)]
            fields = [(1, 2), (trail, 2), (nrel, 2)] + [(r, 2) for r in rel] + [(ptr, CEIL_LOG2[ne])] + \
                     [(res[ne - 1 - i], 1) for i in range(ne)]
        elif fc == 2:                                 # VARFIX
            nrel = int(rng.integers(0, 4)); rel = [int(rng.integers(0, 2)) for _ in range(nrel)]
            ne = nrel + 1
            t_env = [lead]
            for i in range(nrel):
                t_env.append(t_env[-1] + 2 * rel[i] + 2)
            t_env.append(16)
            ptr = int(rng.integers(0, min(ne + 2, 1 << CEIL_LOG2[ne])))
            res = [int(rng.integers(0, 2)) for _ in range(ne)]
            trail = 0
            fields = [(2, 2), (lead, 2), (nrel, 2)] + [(r, 2) for r in rel] + [(ptr, CEIL_LOG2[ne])] + \
                     [(r, 1) for r in res]
        else:                                         # VARVAR
            trail = int(rng.integers(0, 4))
            nl = int(rng.integers(0, 3)); nt = int(rng.integers(0, 3))
            rl = [int(rng.integers(0, 2)) for _ in range(nl)]; rt = [int(rng.integers(0, 2)) for _ in range(nt)]
            ne = nl + nt + 1
            t_env = [0] * (ne + 1)
            t_env[0] = lead; t_env[ne] = 16 + trail
            for i in range(nl):
                t_env[i + 1] = t_env[i] + 2 * rl[i] + 2
            for i in range(nt):
                t_env[ne - 1 - i] = t_env[ne - i] - 2 * rt[i] - 2
            ptr = int(rng.integers(0, min(ne + 2, 1 << CEIL_LOG2[ne])))
            res = [int(rng.integers(0, 2)) for _ in range(ne)]
            fields = [(3, 2), (lead, 2), (trail, 2), (nl, 2), (nt, 2)] + [(r, 2) for r in rl] + \
                     [(r, 2) for r in rt] + [(ptr, CEIL_LOG2[ne])] + [(r, 1) for r in res]
        if all(t_env[i] < t_env[i + 1] for i in range(ne)):
            break
    g.update(num_env=ne, res=res, pointer=ptr, trail=trail, t_env=t_env, fields=fields)
    g["num_noise"] = 2 if ne > 1 else 1
    return g


def _write_envelopes(bw, rng, st, grid, amp_res, n, coupled_bal, first):
    """read_sbr_envelope, aacsbr.c:787-854.  n = (N_low, N_high)."""
    ne = grid["num_env"]
    if coupled_bal:
        names = ("t_huffman_env_bal_3_0dB", "f_huffman_env_bal_3_0dB") if amp_res else \
                ("t_huffman_env_bal_1_5dB", "f_huffman_env_bal_1_5dB")
        start_bits = 5 if amp_res else 6
        lo, hi = (3, 9) if amp_res else (6, 18)        # transmitted balance index; decoder stores 2x (centre 12 / 24)
        step_max = 1
    else:
        names = ("t_huffman_env_3_0dB", "f_huffman_env_3_0dB") if amp_res else \
                ("t_huffman_env_1_5dB", "f_huffman_env_1_5dB")
        start_bits = 6 if amp_res else 7
        lo, hi = (4, 20) if amp_res else (8, 40)
        step_max = 2
    odd = n[1] & 1
    df = grid["df_env"]
    for e in range(ne):
        res = grid["res"][e]
        nb = n[res]
        if df[e] and st.prev_env is not None:
            if res == st.prev_res:
                base = list(st.prev_env)
            elif res:
                base = [st.prev_env[(j + odd) >> 1] for j in range(nb)]
            else:
                base = [st.prev_env[2 * j - odd if j else 0] for j in range(nb)]
            vals = []
            for j in range(nb):
                d = int(rng.integers(-step_max, step_max + 1))
                v = base[j] + d
                if v < lo: d += lo - v
                if v > hi: d -= v - hi
                put_huff(bw, names[0], d + SBR_LAV[names[0]])
                vals.append(base[j] + d)
        else:
            assert not df[e]
            v0 = int(rng.integers(lo, hi + 1))
            bw.put(v0, start_bits)
            vals = [v0]
            for j in range(1, nb):
                d = int(rng.integers(-step_max, step_max + 1))
                v = vals[-1] + d
                if v < lo: d += lo - v
                if v > hi: d -= v - hi
                put_huff(bw, names[1], d + SBR_LAV[names[1]])
                vals.append(vals[-1] + d)
        st.prev_env, st.prev_res = vals, res


def _write_noise(bw, rng, st, grid, n_q, coupled_bal):
    """read_sbr_noise, aacsbr.c:856-890."""
    if coupled_bal:
        tn, fn = "t_huffman_noise_bal_3_0dB", "f_huffman_env_bal_3_0dB"
        lo, hi = 3, 9
    else:
        tn, fn = "t_huffman_noise_3_0dB", "f_huffman_env_3_0dB"
        lo, hi = 4, 20
    df = grid["df_noise"]
    for e in range(grid["num_noise"]):
        if df[e] and st.prev_noise is not None:
            vals = []
            for j in range(n_q):
                d = int(rng.integers(-1, 2))
                v = st.prev_noise[j] + d
                if v < lo: d += lo - v
                if v > hi: d -= v - hi
                put_huff(bw, tn, d + SBR_LAV[tn])
                vals.append(st.prev_noise[j] + d)
        else:
            assert not df[e]
            v0 = int(rng.integers(lo, hi + 1))
            bw.put(v0, 5)
            vals = [v0]
            for j in range(1, n_q):
                d = int(rng.integers(-1, 2))
                v = vals[-1] + d
                if v < lo: d += lo - v
                if v > hi: d -= v - hi
                put_huff(bw, fn, d + SBR_LAV[fn])
                vals.append(vals[-1] + d)
        st.prev_noise = vals


class PsState:
    """Encoder-side mirror of the PSContext fields the parser keeps between frames (aacps.h:40-63)."""

    def __init__(self):
        self.enable_iid = 0; self.enable_icc = 0; self.enable_ext = 0
        self.nr_iid = 0; self.nr_icc = 0; self.nr_pd = 0; self.quant = 0     # as the decoder holds them
        self.is34 = 0; self.enable_ipdopd = 0
        self.prev = {"iid": None, "icc": None, "ipd": None, "opd": None}
        self.have_header = False


NR_IIDICC = [10, 20, 34, 10, 20, 34]
NR_IPDOPD = [5, 11, 17, 5, 11, 17]


def _ps_par(bw, rng, ps, kind, num, table_df, table_dt, offset, lo, hi, step, modulo):
    """One envelope of IID / ICC / IPD / OPD indices (READ_PAR_DATA, aacps.c:84-118)."""
    prev = ps.prev[kind]
    dt = 1 if (prev is not None and len(prev) == num and rng.random() < 0.35) else 0
    bw.put(dt, 1)
    vals = []
    if dt:
        for b in range(num):
            if modulo:
                d = int(rng.integers(0, 8))
                put_huff(bw, table_dt, d)
                vals.append((prev[b] + d) & 7)
            else:
                d = int(rng.integers(-step, step + 1))
                v = prev[b] + d
                if v < lo: d += lo - v
                if v > hi: d -= v - hi
                put_huff(bw, table_dt, d + offset)
                vals.append(prev[b] + d)
    else:
        cur = 0
        for b in range(num):
            if modulo:
                d = int(rng.integers(0, 8))
                put_huff(bw, table_df, d)
                cur = (cur + d) & 7
            else:
                target = int(rng.integers(lo, hi + 1)) if b == 0 else \
                    min(hi, max(lo, cur + int(rng.integers(-step, step + 1))))
                put_huff(bw, table_df, target - cur + offset)
                cur = target
            vals.append(cur)
    ps.prev[kind] = vals


def write_ps_data(bw, rng, ps, frame_idx, header_period=8, ipdopd=True, fixed_modes=None):
    """ps_data(), aacps.c:150-279.  Tracks the decoder's retained fields so that the stream never drives
    the reference into its one undefined corner: 20-band processing with nr_ipdopd_par == 17 and IPD/OPD
    enabled reads uninitialised stack (remap20 with full = 0 fills 11 of the 17 entries, aacps.c:845-848)."""
    header = (not ps.have_header) or (frame_idx % header_period == 0)
    bw.put(1 if header else 0, 1)
    if header:
        if fixed_modes is None:
            iid_mode = int(rng.integers(0, 6)); icc_mode = int(rng.integers(0, 6))
        else:
            iid_mode, icc_mode = fixed_modes
        ps.enable_iid = 1 if (rng.random() < 0.9 or not ps.have_header) else 0
        ps.enable_icc = 1 if rng.random() < 0.9 else 0
        bw.put(ps.enable_iid, 1)
        if ps.enable_iid:
            bw.put(iid_mode, 3)
            ps.nr_iid, ps.quant, ps.nr_pd = NR_IIDICC[iid_mode], int(iid_mode > 2), NR_IPDOPD[iid_mode]
        bw.put(ps.enable_icc, 1)
        if ps.enable_icc:
            bw.put(icc_mode, 3)
            ps.nr_icc = NR_IIDICC[icc_mode]
        if ps.enable_iid or ps.enable_icc:
            ps.is34 = int((ps.enable_iid and ps.nr_iid == 34) or (ps.enable_icc and ps.nr_icc == 34))
        hazard = (not ps.is34) and ps.nr_pd == 17
        # the extension is the only way to clear a stale enable_ipdopd: always send it when the hazard exists
        ps.enable_ext = 1 if (hazard or (ipdopd and rng.random() < 0.5)) else 0
        bw.put(ps.enable_ext, 1)
        ps.have_header = True
        ps.prev = {"iid": None, "icc": None, "ipd": None, "opd": None}
    frame_class = int(rng.integers(0, 2))
    bw.put(frame_class, 1)
    idx = int(rng.integers(0, 4))
    if header and frame_class == 0 and idx == 0:
        # num_env == 0 makes the decoder reuse the previous frame's last envelope (aacps.c:238-256); right after a header
        # that switched to the coarse IID quantiser those values (up to +-15) index the reference's H LUT out of bounds
        # (aacps.c:863): a second undefined corner the fixture must stay out of -- a header frame carries >= 1 envelope
        idx = int(rng.integers(1, 4))
    num_env = [[0, 1, 2, 4], [1, 2, 3, 4]][frame_class][idx]
    bw.put(idx, 2)
    if frame_class:
        borders = sorted(rng.choice(np.arange(0, 32), size=num_env, replace=False).tolist())
        for b in borders: bw.put(int(b), 5)
    if ps.enable_iid:
        lim = min(7 + 8 * ps.quant, 6 + 5 * ps.quant)
        for e in range(num_env):
            _ps_par(bw, rng, ps, "iid", ps.nr_iid, "huff_iid_df%d" % ps.quant, "huff_iid_dt%d" % ps.quant,
                    30 if ps.quant else 14, -lim, lim, 2, False)
    if ps.enable_icc:
        for e in range(num_env):
            _ps_par(bw, rng, ps, "icc", ps.nr_icc, "huff_icc_df", "huff_icc_dt", 7, 0, 7, 1, False)
    if ps.enable_ext:
        hazard = (not ps.is34) and ps.nr_pd == 17
        ext = BitWriter()
        ext.put(0, 2)                      # ps_extension_id 0 = IPD/OPD
        en = 1 if (rng.random() < 0.85 and not hazard and ipdopd) else 0
        ext.put(en, 1)
        ps.enable_ipdopd = en
        if en:
            for e in range(num_env):
                _ps_par(ext, rng, ps, "ipd", ps.nr_pd, "huff_ipd_df", "huff_ipd_dt", 0, 0, 7, 1, True)
                _ps_par(ext, rng, ps, "opd", ps.nr_pd, "huff_opd_df", "huff_opd_dt", 0, 0, 7, 1, True)
        ext.put(0, 1)                      # reserved_ps
        cnt = (len(ext) + 7) // 8
        if cnt < 15:
            bw.put(cnt, 4)
        else:
            bw.put(15, 4); bw.put(cnt - 15, 8)
        for b in ext.bits: bw.put(b, 1)
        bw.put(0, cnt * 8 - len(ext))
    elif ps.enable_ipdopd and (not ps.is34) and ps.nr_pd == 17:
        raise AssertionError("generator bug: stale IPD/OPD would hit the reference's uninitialised read")


class HeAacStreamGen:
    """Per-stream generator state: writes one sbr_extension_data() per core frame."""

    def __init__(self, seed, stereo, ps=False, header_period=16, allow_var_grid=True, start_freq=5, stop_freq=9,
                 xover=0, extra1=None, extra2="random", ps_modes=None, ps_ipdopd=True, tables=None):
        self.rng = np.random.default_rng(seed)
        self.stereo, self.ps = stereo, ps
        self.header_period = header_period
        self.allow_var = allow_var_grid
        if extra2 == "random":
            r = self.rng
            extra2 = (int(r.integers(0, 4)), int(r.integers(0, 4)), int(r.integers(0, 2)), int(r.integers(0, 2))) \
                if r.random() < 0.75 else None
        self.hdr = SbrHeader(self.rng, start_freq, stop_freq, xover, extra1, extra2)
        self.ch = [SbrChanState(), SbrChanState()]
        self.ps_state = PsState()
        self.ps_modes, self.ps_ipdopd = ps_modes, ps_ipdopd
        self.frame = 0
        self.loop_frames = None   # if set, frame loop_frames-1 ends at the frame border (stream can be looped)
        self.tables = tables      # dict(n=(N_low, N_high), n_q=..) from probe_tables()

    def sbr_payload(self):
        rng, bw = self.rng, BitWriter()
        n, n_q = self.tables["n"], self.tables["n_q"]
        send_header = self.frame % self.header_period == 0
        bw.put(1 if send_header else 0, 1)
        if send_header:
            self.hdr.write(bw)
        first = self.frame == 0
        nch = 2 if self.stereo else 1
        bw.put(0, 1)                                        # bs_data_extra
        coupling = 0
        if self.stereo:
            # a coupled frame shares one grid: only conformant when both channels' previous grids ended at
            # the same border (otherwise channel 1 would have QMF slots no envelope covers)
            coupling = 1 if (rng.random() < 0.4 and self.ch[0].prev_trail == self.ch[1].prev_trail) else 0
            bw.put(coupling, 1)
            if coupling != getattr(self, "_last_coupling", coupling):
                # channel 1 switches between level and balance domains: restart its delta-time coding
                self.ch[1].prev_env = None
                self.ch[1].prev_noise = None
        grids = []
        for c in range(nch):
            if c == 1 and coupling:
                g = dict(grids[0])
            else:
                g = _gen_grid(rng, self.ch[c], self.allow_var, self.loop_frames is not None and self.frame == self.loop_frames - 1)
                for v, nb in g["fields"]:
                    bw.put(v, nb)
            grids.append(g)
        if coupling:
            # copy_sbr_grid: channel 1 inherits channel 0's grid; keep the trail bookkeeping in step
            self.ch[1].prev_trail = self.ch[0].prev_trail
        amp = []
        for c in range(nch):
            g = grids[c] = dict(grids[c])
            a = self.hdr.amp_res
            if g["fc"] == 0 and g["num_env"] == 1:
                a = 0
            amp.append(a)
            st = self.ch[c]
            g["df_env"] = [0 if (st.prev_env is None and e == 0) else int(rng.random() < 0.3) for e in range(g["num_env"])]
            g["df_noise"] = [0 if (st.prev_noise is None and e == 0) else int(rng.random() < 0.3) for e in range(g["num_noise"])]
            for b in g["df_env"]: bw.put(b, 1)
            for b in g["df_noise"]: bw.put(b, 1)
        invf = [int(rng.integers(0, 4)) for _ in range(n_q)]
        for v in invf: bw.put(v, 2)
        if self.stereo and not coupling:
            for _ in range(n_q): bw.put(int(rng.integers(0, 4)), 2)
        if self.stereo and coupling:
            _write_envelopes(bw, rng, self.ch[0], grids[0], amp[0], n, False, first)
            _write_noise(bw, rng, self.ch[0], grids[0], n_q, False)
            _write_envelopes(bw, rng, self.ch[1], grids[1], amp[1], n, True, first)
            _write_noise(bw, rng, self.ch[1], grids[1], n_q, True)
        else:
            for c in range(nch):
                _write_envelopes(bw, rng, self.ch[c], grids[c], amp[c], n, False, first)
            for c in range(nch):
                _write_noise(bw, rng, self.ch[c], grids[c], n_q, False)
        for c in range(nch):
            flag = 1 if rng.random() < 0.25 else 0
            bw.put(flag, 1)
            if flag:
                for _ in range(n[1]): bw.put(1 if rng.random() < 0.125 else 0, 1)
        for c in range(nch):
            self.ch[c].prev_trail = grids[c]["trail"]
        # the coupled / uncoupled value domains differ; restart delta-time coding when the mode flips
        self._last_coupling = coupling
        if self.ps:
            ext = BitWriter()
            ext.put(2, 2)                                   # EXTENSION_ID_PS
            write_ps_data(ext, rng, self.ps_state, self.frame, ipdopd=self.ps_ipdopd, fixed_modes=self.ps_modes)
            size = (len(ext) + 7) // 8
            bw.put(1, 1)
            if size < 15:
                bw.put(size, 4)
            else:
                bw.put(15, 4); bw.put(size - 15, 8)
            for b in ext.bits: bw.put(b, 1)
            bw.put(0, size * 8 - len(ext))
        else:
            bw.put(0, 1)                                    # bs_extended_data
        self.frame += 1
        return bw


def splice_fil(pkt, payload_bits):
    """Insert FIL(EXT_SBR_DATA, payload) before the frame's END element.  aacenc pads with zero
    bits after END (= 0b111), so END starts at (last set bit - 2) (SURVEY 7 'Fixtures')."""
    bits = np.unpackbits(np.frombuffer(pkt, np.uint8))
    last = int(np.nonzero(bits)[0][-1])
    end_pos = last - 2
    assert bits[end_pos:end_pos + 3].tolist() == [1, 1, 1]
    bw = BitWriter()
    bw.bits = bits[:end_pos].tolist()
    nbits = 4 + len(payload_bits)
    cnt = (nbits + 7) // 8
    bw.put(6, 3)                               # TYPE_FIL
    if cnt < 15:
        bw.put(cnt, 4)
    else:
        assert cnt - 14 < 256, cnt
        bw.put(15, 4); bw.put(cnt - 14, 8)     # cnt = 15 + esc - 1, aacdec.c:2050-2051
    bw.put(0xD, 4)                             # EXT_SBR_DATA
    bw.bits += list(payload_bits.bits)
    bw.put(0, cnt * 8 - nbits)
    bw.put(7, 3)                               # TYPE_END
    return bw.tobytes()


def probe_tables(core_pkt, extradata, core_rate, channels, hdr):
    """Ask the reference which table sizes a header implies: decode one scratch frame whose SBR
    payload is `hdr` followed by zero bits and read N_low/N_high/N_Q from the captured element
    (the reference derives them in sbr_reset, aacsbr.c:1023-1034)."""
    bw = BitWriter()
    bw.put(1, 1)
    hdr.write(bw)
    bw.put(0, 400)
    pkt = splice_fil(core_pkt, bw)
    layout = "stereo" if channels == 2 else "mono"
    pyref.lib().ref_set_log_level(-8)
    d = pyref.Decoder(extradata, core_rate, channels, layout)
    r = d.decode(pkt)
    d.close()
    pyref.lib().ref_set_log_level(16)
    s = r["sbr"][0]
    assert s["start"] == 1, "probe header rejected by the reference"
    return {"n": (int(s["n"][0]), int(s["n"][1])), "n_q": int(s["n_q"])}


def core_noise(n_samples, channels, seed, div=8):
    return pyref.lcg_noise(n_samples, channels, seed=seed, div=div)


def make_stream(seed, n_frames, stereo=True, ps=False, core_rate=24000, bitrate_per_ch=32000, loopable=False, **gen_kw):
    """Returns (packets, extradata, info).  n_frames packets, each a raw_data_block with SBR.
    loopable: the last frame's grid ends at the frame border and n_frames should be a multiple of the header
    period, so frame 0 may follow frame n_frames-1."""
    ch = 2 if stereo else 1
    pcm = core_noise(1024 * (n_frames + 1), ch, seed)
    enc = pyref.Encoder(core_rate, ch, bitrate_per_ch * ch)
    core = enc.encode(pcm)[:n_frames]
    extradata = enc.extradata
    enc.close()
    gen = HeAacStreamGen(seed, stereo, ps=ps, **gen_kw)
    gen.tables = probe_tables(core[0], extradata, core_rate, ch, gen.hdr)
    if loopable:
        gen.loop_frames = n_frames
    pkts = [splice_fil(p, gen.sbr_payload()) for p in core]
    return pkts, extradata, {"tables": gen.tables, "channels": ch, "core_rate": core_rate}


def make_stream_51(seed, n_frames, core_sr_index=6, **gen_kw):
    """HE-AAC 5.1 (SCE CPE CPE LFE, implicit SBR signalling): AAC-LC frames from the repo's own writer (tests/lc_writer.py:
    it writes element by element, so the SBR fill elements can follow the SCE and the two CPEs) with one SBR generator per
    element; the LFE carries none -- the reference upsamples it (aacdec.c:1923-1925).  Returns (packets, extradata, info)."""
    import lc_writer
    rate = [96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000][core_sr_index]
    w = lc_writer.LcWriter(seed, core_sr_index, pyref.LAYOUTS["5.1"][0], tns_prob=0.3, short_bias=0.15)
    # table sizes of each element's header from the reference, probed on a stereo scratch frame of the same core rate
    enc = pyref.Encoder(rate, 2, 64000)
    probe_core = enc.encode(core_noise(2048, 2, seed))[0]
    probe_ed = enc.extradata
    enc.close()
    gens = {}
    for ei, t in enumerate(pyref.LAYOUTS["5.1"][0]):
        if t == abi.TYPE_LFE:
            continue
        g = HeAacStreamGen(seed * 7 + ei, t == abi.TYPE_CPE, **gen_kw)
        g.tables = probe_tables(probe_core, probe_ed, rate, 2, g.hdr)
        gens[ei] = g
    pkts = [w.frame(fil={ei: g.sbr_payload() for ei, g in gens.items()}) for _ in range(n_frames)]
    return pkts, lc_writer.audio_specific_config(core_sr_index, 6), {"core_rate": rate, "writer": w}


def asc_explicit(core_idx, chan_config, ext_idx, ps=False):
    """AudioSpecificConfig with hierarchical explicit SBR (AOT 5) / PS (AOT 29) signalling (mpeg4audio.c:77-130):
    ext_idx == core_idx asks for downsampled SBR."""
    bits = "{:05b}{:04b}{:04b}{:04b}{:05b}000".format(29 if ps else 5, core_idx, chan_config, ext_idx, 2)
    bits += "0" * (-len(bits) % 8)
    return int(bits, 2).to_bytes(len(bits) // 8, "big")


def decode_with_reference(pkts, extradata, core_rate, channels, float_mode=True):
    """Decode with the unmodified reference; returns list of per-frame capture dicts and the decoder."""
    layout = "stereo" if channels == 2 else "mono"
    d = pyref.Decoder(extradata, core_rate, channels, layout, float_mode=float_mode)
    frames = [d.decode(p) for p in pkts]
    return frames, d
