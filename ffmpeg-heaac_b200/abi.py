"""numpy mirrors of the C-ABI structs in include/heaac_b200.h.

Field names, order and sizes follow the header one to one (the header cites the
reference struct each field comes from: aac.h:135-162, sbr.h:59-156, aacps.h:40-63).
`check_sizes()` is asserted against the sizes the built library reports, so a drift
between this file and the header fails loudly.
"""
import numpy as np

TNS_MAX_ORDER = 20
TYPE_SCE, TYPE_CPE, TYPE_LFE = 0, 1, 3
ONLY_LONG, LONG_START, EIGHT_SHORT, LONG_STOP = 0, 1, 2, 3

ICS = np.dtype([
    ("window_sequence", "u1"), ("window_shape", "u1"), ("max_sfb", "u1"), ("tns_present", "u1"),
    ("reserved", "u1", (4,)),
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

SIZES = {"HeaacIcs": ICS.itemsize, "HeaacTns": TNS.itemsize, "HeaacSbrChan": SBR_CHAN.itemsize,
         "HeaacSbrElem": SBR_ELEM.itemsize, "HeaacPsFrame": PS_FRAME.itemsize}
