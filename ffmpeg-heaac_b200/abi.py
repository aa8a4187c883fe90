"""numpy mirrors of the C-ABI structs in include/heaac_b200.h.

Field names, order and sizes follow the header one to one (the header cites the
reference struct each field comes from: aac.h:135-162, sbr.h:59-156, aacps.h:40-63).
`check_sizes()` is asserted against the sizes the built library reports, so a drift
between this file and the header fails loudly.
"""
import numpy as np

TNS_MAX_ORDER = 20
TYPE_SCE, TYPE_CPE, TYPE_CCE, TYPE_LFE = 0, 1, 2, 3
ONLY_LONG, LONG_START, EIGHT_SHORT, LONG_STOP = 0, 1, 2, 3

ICS = np.dtype([
    ("window_sequence", "u1"), ("window_shape", "u1"), ("max_sfb", "u1"), ("tns_present", "u1"),
    ("prev_valid", "u1"), ("prev_window_sequence", "u1"), ("prev_window_shape", "u1"), ("reserved", "u1"),
], align=True)

TNS = np.dtype([
    ("n_filt", "u1", (8,)), ("length", "u1", (8,)), ("order", "u1", (8,)), ("direction", "u1", (8,)),
    ("coef", "f4", (8, TNS_MAX_ORDER)),
], align=True)

SBR_CHAN = np.dtype([
    ("bs_num_env", "u1"), ("bs_num_noise", "u1"), ("bs_amp_res", "u1"), ("bs_add_harmonic_flag", "u1"),
    ("e_a", "i1", (2,)), ("t_env_num_env_old", "u1"), ("reserved0", "u1"),
    ("bs_freq_res", "u1", (8,)), ("t_env", "u1", (8,)), ("t_q", "u1", (4,)),
    ("bs_invf_mode", "u1", (2, 5)), ("reserved1", "u1", (2,)),
    ("bs_add_harmonic", "u1", (48,)),
    ("env_q", "i2", (5, 48)), ("noise_q", "i2", (2, 5)), ("reserved2", "u1", (4,)),
], align=True)

SBR_ELEM = np.dtype([
    ("start", "u1"), ("reset", "u1"), ("bs_coupling", "u1"), ("bs_limiter_gains", "u1"),
    ("bs_interpol_freq", "u1"), ("bs_smoothing_mode", "u1"), ("num_patches", "u1"), ("n_q", "u1"),
    ("n", "u1", (2,)), ("n_lim", "u1"), ("k0", "u1"), ("kx", "u1", (2,)), ("m", "u1", (2,)),
    ("f_tablelow", "u1", (25,)), ("f_tablehigh", "u1", (49,)), ("f_tablenoise", "u1", (6,)),
    ("f_tablelim", "u1", (29,)),
    ("patch_num_subbands", "u1", (6,)), ("patch_start_subband", "u1", (6,)), ("reserved", "u1", (3,)),
    ("ch", SBR_CHAN, (2,)),
], align=True)

PS_FRAME = np.dtype([
    ("start", "u1"), ("num_env", "u1"), ("num_env_old", "u1"), ("iid_quant", "u1"),
    ("nr_iid_par", "u1"), ("nr_icc_par", "u1"), ("nr_ipdopd_par", "u1"), ("icc_mode", "u1"),
    ("enable_ipdopd", "u1"), ("is34bands", "u1"), ("is34bands_old", "u1"), ("reserved0", "u1"),
    ("border_position", "i1", (6,)), ("reserved1", "u1", (2,)),
    ("iid_par", "i1", (5, 34)), ("icc_par", "i1", (5, 34)),
    ("ipd_par", "i1", (5, 34)), ("opd_par", "i1", (5, 34)),
], align=True)

CCE = np.dtype([("coupling_point", "u1"), ("n_targets", "u1"), ("reserved", "u1", 2), ("target_ch", "u1", 16), ("gain", "f4", 16)])

SIZES = {"HeaacCce": CCE.itemsize, "HeaacIcs": ICS.itemsize, "HeaacTns": TNS.itemsize, "HeaacSbrChan": SBR_CHAN.itemsize,
         "HeaacSbrElem": SBR_ELEM.itemsize, "HeaacPsFrame": PS_FRAME.itemsize}

# ---- ctypes mirrors of HeaacConfig / HeaacBatch (include/heaac_b200.h) -------------
import ctypes as C

ABI_VERSION = 2
MAX_ELEMS = 8


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("device", C.c_int32), ("n_streams", C.c_int32), ("n_elems", C.c_int32),
        ("elem_type", C.c_uint8 * MAX_ELEMS), ("sr_index", C.c_int32), ("sbr", C.c_int32), ("ps", C.c_int32),
        ("add_bias", C.c_float), ("sf_scale", C.c_float), ("max_frames_per_call", C.c_int32),
        ("want_float_out", C.c_int32), ("keep_intermediates", C.c_int32), ("out_slot", C.c_uint8 * 16),
        ("chunk_frames", C.c_int32), ("sbr_downsampled", C.c_int32),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("n_frames", C.c_int32), ("on_device", C.c_int32),
        ("coeffs", C.c_void_p), ("ics", C.c_void_p), ("tns", C.c_void_p), ("sbr", C.c_void_p), ("ps", C.c_void_p),
        ("pcm", C.c_void_p), ("pcm_f32", C.c_void_p), ("cce", C.c_void_p),
    ]


class DecoderConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("device", C.c_int32), ("n_streams", C.c_int32), ("sample_rate", C.c_int32),
        ("channels", C.c_int32), ("max_frames_per_call", C.c_int32), ("n_threads", C.c_int32),
        ("want_float_out", C.c_int32), ("extradata", C.c_void_p), ("extradata_size", C.c_int32), ("gpu_parse", C.c_int32),
        ("chunk_frames", C.c_int32),
    ]


# aac_channel_layout_map (aacdectab.h:74-82): output order of the reference per channel configuration
_LAYOUT_MAP = {
    1: [(TYPE_SCE, 0)], 2: [(TYPE_CPE, 0)], 3: [(TYPE_CPE, 0), (TYPE_SCE, 0)],
    4: [(TYPE_CPE, 0), (TYPE_SCE, 0), (TYPE_SCE, 1)], 5: [(TYPE_CPE, 0), (TYPE_SCE, 0), (TYPE_CPE, 1)],
    6: [(TYPE_CPE, 0), (TYPE_SCE, 0), (TYPE_LFE, 0), (TYPE_CPE, 1)],
    7: [(TYPE_CPE, 0), (TYPE_SCE, 0), (TYPE_LFE, 0), (TYPE_CPE, 2), (TYPE_CPE, 1)],
}
_CHAN_CONFIG = {(TYPE_SCE,): 1, (TYPE_CPE,): 2, (TYPE_SCE, TYPE_CPE): 3, (TYPE_SCE, TYPE_CPE, TYPE_SCE): 4,
                (TYPE_SCE, TYPE_CPE, TYPE_CPE): 5, (TYPE_SCE, TYPE_CPE, TYPE_CPE, TYPE_LFE): 6,
                (TYPE_SCE, TYPE_CPE, TYPE_CPE, TYPE_CPE, TYPE_LFE): 7}


def reference_out_slots(elem_types):
    """Output slot of every core channel (bitstream element order) in the reference's output order
    (output_configure, aacdec.c:214-272).  Identity for layouts outside the default configurations."""
    cc = _CHAN_CONFIG.get(tuple(elem_types))
    ncc = n_core_channels(elem_types)
    if cc is None:
        return list(range(ncc))
    ids, cnt = [], {}
    for t in elem_types:
        ids.append((t, cnt.get(t, 0)))
        cnt[t] = cnt.get(t, 0) + 1
    first_slot, pos = {}, 0
    for key in _LAYOUT_MAP[cc]:
        first_slot[key] = pos
        pos += 2 if key[0] == TYPE_CPE else 1
    slots = []
    for key in ids:
        slots += [first_slot[key] + i for i in range(2 if key[0] == TYPE_CPE else 1)]
    return slots


def make_config(n_streams, elem_types, sr_index, sbr=False, ps=False, bias_mode=False, device=0,
                max_frames=1, want_float=True, keep_intermediates=False, out_slots="reference", downsampled=False):
    """Build a HeaacConfig.  bias_mode selects the reference's C 'bias' scaling (aacdec.c:573-576),
    otherwise the s16-unit scaling of its SIMD branch (aacdec.c:577-581)."""
    c = Config()
    c.abi_version = ABI_VERSION
    c.device = device
    c.n_streams = n_streams
    c.n_elems = len(elem_types)
    for i, t in enumerate(elem_types):
        c.elem_type[i] = t
    c.sr_index = sr_index
    c.sbr = int(bool(sbr))
    c.sbr_downsampled = int(bool(downsampled) and bool(sbr))
    c.ps = int(bool(ps))
    if bias_mode:
        c.add_bias = 385.0
        c.sf_scale = float(np.float32(1.0 / (-1024.0 * 32768.0)))
    else:
        c.add_bias = 0.0
        c.sf_scale = float(np.float32(1.0 / -1024.0))
    c.max_frames_per_call = max_frames
    c.want_float_out = int(bool(want_float))
    c.keep_intermediates = int(bool(keep_intermediates))
    if out_slots == "reference":
        out_slots = reference_out_slots(list(elem_types)) if not ps else [0]
    if out_slots is not None:
        for i, v in enumerate(out_slots):
            c.out_slot[i] = v
    return c


def n_core_channels(elem_types):
    return sum(2 if t == TYPE_CPE else 1 for t in elem_types)


def n_sbr_elems(elem_types):
    return sum(1 for t in elem_types if t != TYPE_LFE)
