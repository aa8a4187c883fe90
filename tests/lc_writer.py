"""AAC-LC raw_data_block WRITER with random content (test fixture code, SURVEY.md 8c fixture (3) / config 4).

The reference's own encoder never sets the TNS bit, always uses KBD windows and never emits PNS, intensity
stereo, pulses or escape codes, so streams that exercise those decoder branches are written here directly:
random quantised spectra, section data, scalefactors, pulse data, TNS filters (both directions, orders up to
12 long / 7 short), both window shapes, block switching with grouping, M/S masks.  Syntax follows what the
reference's parser reads: decode_ics_info (aacdec.c:645-710), decode_band_types (720-756),
decode_scalefactors (768-822), decode_pulses (829-847), decode_tns (854-887),
decode_spectrum_and_dequant (988-1245), decode_cpe (1460-1494), element loop (1973-2076).
Huffman tables come from the reference (oracle/ref_tables.c) -- generation needs oracle/_ref; the generated
packets are committed inside tests/golden/*.npz.
"""
import numpy as np

import pyref
from heaac_streams import BitWriter

abi = pyref.abi
ONLY_LONG, LONG_START, EIGHT_SHORT, LONG_STOP = 0, 1, 2, 3
ZERO_BT, NOISE_BT, INTENSITY_BT2, INTENSITY_BT = 0, 13, 14, 15

_T = {}


def _tables():
    if not _T:
        _T["sf_code"] = pyref.table("scalefactor_code", np.uint32).astype(np.int64)
        _T["sf_bits"] = pyref.table("scalefactor_bits", np.uint8).astype(np.int64)
        for i in range(1, 12):
            _T["code%d" % i] = pyref.table("spectral_codes_%d" % (i - 1), np.uint16).astype(np.int64)
            _T["bits%d" % i] = pyref.table("spectral_bits_%d" % (i - 1), np.uint8).astype(np.int64)
        _T["swb1024"] = pyref.table("swb_offset_1024", np.uint16).reshape(13, -1)
        _T["swb128"] = pyref.table("swb_offset_128", np.uint16).reshape(13, -1)
        _T["nswb1024"] = pyref.table("num_swb_1024", np.uint8)
        _T["nswb128"] = pyref.table("num_swb_128", np.uint8)
    return _T


# book -> (dimension, range, signed, largest absolute value)
BOOKS = {1: (4, 3, True, 1), 2: (4, 3, True, 1), 3: (4, 3, False, 2), 4: (4, 3, False, 2), 5: (2, 9, True, 4),
         6: (2, 9, True, 4), 7: (2, 8, False, 7), 8: (2, 8, False, 7), 9: (2, 13, False, 12), 10: (2, 13, False, 12),
         11: (2, 17, False, 16)}


def audio_specific_config(sr_index, chan_config):
    v = (2 << 11) | (sr_index << 7) | (chan_config << 3)          # AOT 2 (LC), GASpecificConfig flags 0
    return bytes([v >> 8, v & 255])


class ChanState:
    def __init__(self):
        self.ws = ONLY_LONG


class LcWriter:
    """One stream.  frame() -> bytes of one raw_data_block."""

    def __init__(self, seed, sr_index, elem_types, tns_prob=0.6, short_bias=0.45, level=158, pns=True, intensity=True,
                 pulses=True, illegal_transitions=0.04):
        self.rng = np.random.default_rng(seed)
        self.sr_index, self.elem_types = sr_index, list(elem_types)
        self.tns_prob, self.short_bias, self.level = tns_prob, short_bias, level
        self.pns, self.intensity, self.pulses, self.illegal = pns, intensity, pulses, illegal_transitions
        self.ch = [ChanState() for _ in range(sum(2 if t == abi.TYPE_CPE else 1 for t in elem_types))]
        self.stats = {"frames": 0, "short": 0, "chan_frames": 0, "tns": 0, "sine": 0, "pns": 0, "is": 0, "esc": 0, "pulse": 0}
        _tables()

    # ---- side info ----------------------------------------------------------------------------------------
    def _next_ws(self, prev):
        r = self.rng
        if r.random() < self.illegal:
            return int(r.integers(0, 4))
        if prev in (ONLY_LONG, LONG_STOP):
            return LONG_START if r.random() < self.short_bias else ONLY_LONG
        return EIGHT_SHORT if r.random() < 0.6 else LONG_STOP      # after LONG_START / EIGHT_SHORT

    def _ics_info(self, bw, ws):
        r, t = self.rng, _T
        shape = int(r.random() < 0.5)
        bw.put(0, 1); bw.put(ws, 2); bw.put(shape, 1)
        info = {"ws": ws, "shape": shape}
        if ws == EIGHT_SHORT:
            nswb = int(t["nswb128"][self.sr_index])
            info["max_sfb"] = int(r.integers(0, nswb + 1)) if r.random() < 0.9 else nswb
            grouping = int(r.integers(0, 128)) if r.random() < 0.8 else int(r.choice([0, 127]))
            bw.put(info["max_sfb"], 4); bw.put(grouping, 7)
            glen = [1]
            for i in range(6, -1, -1):
                if (grouping >> i) & 1:
                    glen[-1] += 1
                else:
                    glen.append(1)
            info["group_len"], info["swb"] = glen, t["swb128"][self.sr_index]
        else:
            nswb = int(t["nswb1024"][self.sr_index])
            info["max_sfb"] = int(r.integers(nswb // 3, nswb + 1)) if r.random() < 0.95 else 0
            if getattr(self, "force_max_sfb", None) is not None:
                info["max_sfb"] = int(self.force_max_sfb)
            info["nswb"] = nswb
            bw.put(info["max_sfb"], 6); bw.put(0, 1)                # no prediction in LC
            info["group_len"], info["swb"] = [1], t["swb1024"][self.sr_index]
        self.stats["sine"] += shape == 0
        return info

    def _sections(self, bw, info, allow_is):
        """Random codebook runs per group; returns band_type[g][sfb]."""
        r = self.rng
        bits = 3 if info["ws"] == EIGHT_SHORT else 5
        esc = (1 << bits) - 1
        out = []
        for _ in info["group_len"]:
            bt, k = [], 0
            while k < info["max_sfb"]:
                u = r.random()
                if u < 0.15:
                    cb = ZERO_BT
                elif u < 0.23 and self.pns:
                    cb = NOISE_BT
                elif u < 0.33 and allow_is:
                    cb = int(r.choice([INTENSITY_BT, INTENSITY_BT2]))
                else:
                    cb = int(r.choice([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 11]))
                run = min(int(r.geometric(0.35)), info["max_sfb"] - k)
                bw.put(cb, 4)
                n = run
                while n >= esc:
                    bw.put(esc, bits); n -= esc
                bw.put(n, bits)
                bt += [cb] * run
                k += run
            out.append(bt)
        return out

    def _quantised(self, info, band_type):
        """Random quantised values q[g][sfb] = int array [group_len, band width] and the scalefactor target."""
        r = self.rng
        q, want = [], []
        swb = info["swb"]
        for g, glen in enumerate(info["group_len"]):
            qg, wg = [], []
            for i, cb in enumerate(band_type[g]):
                width = int(swb[i + 1] - swb[i])
                if cb in BOOKS:
                    dim, rng_, signed, lav = BOOKS[cb]
                    mag = np.minimum(r.geometric(0.5 if lav > 2 else 0.7, size=(glen, width)) - 1, lav)
                    if cb == 11 and r.random() < 0.5:                          # escapes: values 16 .. a few hundred
                        big = r.random((glen, width)) < 0.2
                        mag = np.where(big, r.integers(16, int(r.choice([40, 200, 1200])), size=(glen, width)), mag)
                        self.stats["esc"] += int((mag >= 16).sum())
                    sign = np.where(r.random((glen, width)) < 0.5, -1, 1)
                    v = mag * sign
                    qg.append(v)
                    peak = max(1, int(np.abs(v).max()))
                    wg.append(self.level - int(round(16.0 / 3.0 * np.log2(peak))) + int(r.integers(-6, 7)))
                else:
                    qg.append(None)
                    wg.append(None)
            q.append(qg); want.append(wg)
        return q, want

    def _scalefactors(self, bw, info, band_type, want, global_gain):
        r = self.rng
        code, bits = _T["sf_code"], _T["sf_bits"]
        off = [global_gain, global_gain - 90, 100]
        noise_first = True
        for g in range(len(info["group_len"])):
            for i, cb in enumerate(band_type[g]):
                if cb == ZERO_BT:
                    continue
                if cb in (INTENSITY_BT, INTENSITY_BT2):
                    tgt = int(np.clip(100 + r.integers(-12, 13), 0, 255))
                    d = int(np.clip(tgt - off[2], -60, 60)); off[2] += d
                    bw.put(int(code[d + 60]), int(bits[d + 60]))
                    self.stats["is"] += 1
                elif cb == NOISE_BT:
                    tgt = int(np.clip(global_gain - 90 + r.integers(-10, 11), 0, 255))
                    if noise_first:
                        noise_first = False
                        d = int(np.clip(tgt - off[1], -256, 255)); off[1] += d
                        bw.put(d + 256, 9)
                    else:
                        d = int(np.clip(tgt - off[1], -60, 60)); off[1] += d
                        bw.put(int(code[d + 60]), int(bits[d + 60]))
                    assert 0 <= off[1] <= 255
                    self.stats["pns"] += 1
                else:
                    tgt = int(np.clip(want[g][i], 0, 255))
                    d = int(np.clip(tgt - off[0], -60, 60)); off[0] += d
                    assert 0 <= off[0] <= 255
                    bw.put(int(code[d + 60]), int(bits[d + 60]))

    def _pulses(self, bw, info):
        r = self.rng
        if getattr(self, "overrun_pulse", False) and info["ws"] != EIGHT_SHORT and 0 < info["max_sfb"] < info["nswb"]:
            # NON-CONFORMANT on purpose: one pulse in the band right behind the last coded one (the reference accepts it:
            # decode_pulses only checks pulse_swb < num_swb and pos <= 1023, aacdec.c:826-852)
            bw.put(1, 1); bw.put(0, 2); bw.put(info["max_sfb"], 6)
            bw.put(1, 5); bw.put(9, 4)
            self.overrun_at = int(info["swb"][info["max_sfb"]]) + 1
            return
        if info["ws"] == EIGHT_SHORT or not self.pulses or r.random() > 0.2 or info["max_sfb"] == 0:
            bw.put(0, 1)
            return
        n = int(r.integers(1, 5))
        sfb = int(r.integers(0, info["max_sfb"]))
        pos = int(info["swb"][sfb])
        offs = []
        limit = int(info["swb"][info["max_sfb"]]) - 1     # conformant streams keep their pulses inside the coded bands
        for _ in range(n):
            o = int(r.integers(0, 32))
            if pos + o > limit:
                break
            pos += o
            offs.append(o)
        if not offs:
            bw.put(0, 1)
            return
        bw.put(1, 1); bw.put(len(offs) - 1, 2); bw.put(sfb, 6)
        for o in offs:
            bw.put(o, 5); bw.put(int(r.integers(0, 16)), 4)
        self.stats["pulse"] += 1

    def _tns(self, bw, info):
        r = self.rng
        if r.random() > self.tns_prob:
            bw.put(0, 1)
            return
        bw.put(1, 1)
        self.stats["tns"] += 1
        is8 = info["ws"] == EIGHT_SHORT
        for _ in range(8 if is8 else 1):
            nf = int(r.integers(0, 2)) if is8 else int(r.integers(0, 4))
            bw.put(nf, 1 if is8 else 2)
            if not nf:
                continue
            res = int(r.integers(0, 2))
            bw.put(res, 1)
            for _f in range(nf):
                bw.put(int(r.integers(0, 16 if is8 else 50)), 4 if is8 else 6)          # length in bands
                order = int(r.integers(0, 8)) if is8 else int(r.choice([0, 1, 2, 3, 4, 6, 8, 10, 12]))
                bw.put(order, 3 if is8 else 5)
                if not order:
                    continue
                bw.put(int(r.integers(0, 2)), 1)                                       # direction
                compress = int(r.integers(0, 2))
                bw.put(compress, 1)
                nb = res + 3 - compress
                for i in range(order):                   # mostly mild reflection coefficients: keeps the filter gain sane
                    if r.random() < 0.75:
                        c = int(r.choice([0, 1, (1 << nb) - 1]))
                    else:
                        c = int(r.integers(0, 1 << nb))
                    bw.put(c, nb)

    def _spectral(self, bw, info, band_type, q):
        for g in range(len(info["group_len"])):
            for i, cb in enumerate(band_type[g]):
                if cb not in BOOKS:
                    continue
                dim, rng_, signed, lav = BOOKS[cb]
                code, bits = _T["code%d" % cb], _T["bits%d" % cb]
                for row in q[g][i]:
                    for k in range(0, len(row), dim):
                        v = [int(x) for x in row[k:k + dim]]
                        idx = 0
                        for x in v:
                            idx = idx * rng_ + (x + lav if signed else min(abs(x), 16))
                        bw.put(int(code[idx]), int(bits[idx]))
                        if not signed:
                            for x in v:
                                if x:
                                    bw.put(int(x < 0), 1)
                            if cb == 11:
                                for x in v:
                                    a = abs(x)
                                    if a >= 16:
                                        n = a.bit_length() - 1
                                        bw.put((1 << (n - 4)) - 1, n - 4); bw.put(0, 1)
                                        bw.put(a - (1 << n), n)

    def _ics(self, bw, st, shared_info, allow_is):
        r = self.rng
        global_gain = int(np.clip(self.level + r.integers(-8, 9), 0, 255))
        bw.put(global_gain, 8)
        if shared_info is None:
            st.ws = self._next_ws(st.ws)
            info = self._ics_info(bw, st.ws)
        else:
            info = shared_info
        band_type = self._sections(bw, info, allow_is)
        q, want = self._quantised(info, band_type)
        self._scalefactors(bw, info, band_type, want, global_gain)
        self._pulses(bw, info)
        self._tns(bw, info)
        bw.put(0, 1)                                    # gain_control_data_present
        self._spectral(bw, info, band_type, q)
        self.stats["chan_frames"] += 1
        self.stats["short"] += info["ws"] == EIGHT_SHORT
        return info

    # ---- coupling channel element (ISO/IEC 14496-3 table 4.8; decode_cce, aacdec.c:1503-1565), independently switched ----
    def set_coupling(self, targets):
        """targets: {element index of a CCE: [(target element index, ch_select), ...]} with ch_select as in the bitstream
        (CPE targets: 0 = both channels with one gain, 1 = right only, 2 = left only, 3 = both with separate gains)."""
        self.coupling = targets
        self.cce_gain_lists = {}

    def _cce_head(self, bw, ei):
        r = self.rng
        tg = self.coupling[ei]
        tags = {}
        elem_tag = []
        for t in self.elem_types:
            elem_tag.append(tags.get(t, 0)); tags[t] = tags.get(t, 0) + 1
        bw.put(getattr(self, "cce_ind_sw", 1), 1)      # ind_sw_cce_flag
        bw.put(len(tg) - 1, 3)                          # num_coupled_elements
        lists = 0
        for te, sel in tg:
            is_cpe = self.elem_types[te] == abi.TYPE_CPE
            bw.put(int(is_cpe), 1); bw.put(elem_tag[te], 4)
            lists += 1
            if is_cpe:
                bw.put(sel, 2)
                lists += sel == 3
        bw.put(int(r.integers(0, 2)), 1)                # cc_domain (irrelevant once independently switched)
        bw.put(int(r.integers(0, 2)), 1)                # gain_element_sign
        bw.put(int(r.integers(0, 4)), 2)                # gain_element_scale
        self.cce_gain_lists[ei] = lists

    def _cce_gains(self, bw, ei):
        code, bits = _T["sf_code"], _T["sf_bits"]
        for _ in range(1, self.cce_gain_lists[ei]):     # the first list has gain 1; the others one common gain element each
            d = int(self.rng.integers(-6, 7))
            bw.put(int(code[d + 60]), int(bits[d + 60]))

    def frame(self, fil=None):
        """One raw_data_block.  fil: {element index: BitWriter with an sbr_extension_data() payload} -- a FIL element
        (EXT_SBR_DATA) is written right behind that element, where the reference attaches it to the element before it
        (aacdec.c:2042-2064)."""
        r = self.rng
        bw = BitWriter()
        ch = 0
        tags = {}
        for ei, t in enumerate(self.elem_types):
            tag = tags.get(t, 0)
            tags[t] = tag + 1
            bw.put(t, 3); bw.put(tag, 4)
            if t == abi.TYPE_CPE:
                a, b = self.ch[ch], self.ch[ch + 1]
                common = r.random() < 0.6
                bw.put(int(common), 1)
                shared = None
                if common:
                    a.ws = b.ws = self._next_ws(a.ws)
                    shared = self._ics_info(bw, a.ws)
                    ms = int(r.choice([0, 1, 1, 2]))
                    bw.put(ms, 2)
                    if ms == 1:
                        for _ in range(len(shared["group_len"]) * shared["max_sfb"]):
                            bw.put(int(r.integers(0, 2)), 1)
                self._ics(bw, a, shared, False)
                self._ics(bw, b, shared, self.intensity)
                ch += 2
            elif t == abi.TYPE_CCE:
                self._cce_head(bw, ei)
                self._ics(bw, self.ch[ch], None, False)
                self._cce_gains(bw, ei)
                ch += 1
            else:
                self._ics(bw, self.ch[ch], None, False)
                ch += 1
            if fil and ei in fil:
                payload = fil[ei]
                nbits = 4 + len(payload)
                cnt = (nbits + 7) // 8
                bw.put(6, 3)                            # TYPE_FIL
                if cnt < 15:
                    bw.put(cnt, 4)
                else:
                    bw.put(15, 4); bw.put(cnt - 14, 8)
                bw.put(0xD, 4)                          # EXT_SBR_DATA
                bw.bits += list(payload.bits)
                bw.put(0, cnt * 8 - nbits)
        bw.put(7, 3)                                    # TYPE_END
        self.stats["frames"] += 1
        return bw.tobytes()


CHAN_CONFIG = {"mono": 1, "stereo": 2, "5.1": 6}


def _pce_body(bw, sr_index, elem_types):
    """program_config_element() from element_instance_tag on (ISO/IEC 14496-3 table 4.2; decode_pce, aacdec.c:280-340):
    SCE / CPE as front channels, LFE, CCE as (independently switched) coupling channels.  Returns the instance tags."""
    tags, ids = {}, []
    for t in elem_types:
        ids.append(tags.get(t, 0)); tags[t] = tags.get(t, 0) + 1
    front = [(t, i) for t, i in zip(elem_types, ids) if t in (abi.TYPE_SCE, abi.TYPE_CPE)]
    lfe = [i for t, i in zip(elem_types, ids) if t == abi.TYPE_LFE]
    cc = [i for t, i in zip(elem_types, ids) if t == abi.TYPE_CCE]
    bw.put(0, 4)                                             # element_instance_tag
    bw.put(1, 2); bw.put(sr_index, 4)                        # object_type (LC), sampling_frequency_index
    bw.put(len(front), 4); bw.put(0, 4); bw.put(0, 4); bw.put(len(lfe), 2); bw.put(0, 3); bw.put(len(cc), 4)
    bw.put(0, 1); bw.put(0, 1); bw.put(0, 1)                 # no mixdown elements
    for t, i in front:
        bw.put(int(t == abi.TYPE_CPE), 1); bw.put(i, 4)
    for i in lfe:
        bw.put(i, 4)
    for i in cc:
        bw.put(1, 1); bw.put(i, 4)                           # cc_element_is_ind_sw, tag
    bw.put(0, (-len(bw)) % 8)                                # byte_alignment
    bw.put(0, 8)                                             # comment_field_bytes
    return ids


def pce_config(sr_index, elem_types):
    """AudioSpecificConfig with channel_configuration 0 and a program_config_element listing the elements.  Returns
    (extradata, elem_ids, order): the reference orders its elements -- and its output channels -- by instance tag, then
    type (output_configure, aacdec.c:245-262); `order` lists the element indices that way."""
    bw = BitWriter()
    bw.put(2, 5); bw.put(sr_index, 4); bw.put(0, 4)           # AOT 2, rate, channel_configuration 0
    bw.put(0, 3)                                             # frame_length_flag, depends_on_core_coder, extension_flag
    ids = _pce_body(bw, sr_index, elem_types)
    order = sorted(range(len(elem_types)), key=lambda e: (ids[e], elem_types[e]))
    return bw.tobytes(), ids, order


def pce_element(sr_index, elem_types):
    """The same program_config_element as a syntactic element of a raw_data_block (ID_PCE = 5): a whole number of bytes,
    to be put in front of a frame of a stream that has no other configuration (ADTS with channel_configuration 0)."""
    bw = BitWriter()
    bw.put(5, 3)
    _pce_body(bw, sr_index, elem_types)
    return bw.tobytes()


def make_lc_stream(seed, n_frames, layout="5.1", sr_index=3, **kw):
    """(packets, extradata, writer) for a random AAC-LC stream of the given channel layout."""
    w = LcWriter(seed, sr_index, pyref.LAYOUTS[layout][0], **kw)
    pkts = [w.frame() for _ in range(n_frames)]
    return pkts, audio_specific_config(sr_index, CHAN_CONFIG[layout]), w
