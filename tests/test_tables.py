"""Every table the product computes on the host (csrc/heaac_tables.cpp) equals the reference's
initialised copy bit for bit, and the generated data header matches the reference's data."""
import ctypes as C

import numpy as np
import pytest

import pyemu
import pyoracle
import pyref
from conftest import needs_ref

TABLES = np.dtype([
    ("sine_1024", "f4", 1024), ("sine_128", "f4", 128), ("kbd_1024", "f4", 1024), ("kbd_128", "f4", 128),
    ("tcos_2048", "f4", 512), ("tsin_2048", "f4", 512), ("tcos_256", "f4", 64), ("tsin_256", "f4", 64),
    ("tcos_syn", "f4", 32), ("tsin_syn", "f4", 32), ("tcos_ana", "f4", 32), ("tsin_ana", "f4", 32),
    ("fft_wre", "f4", 256), ("fft_wim", "f4", 256), ("bitrev512", "u2", 512), ("bitrev32", "u1", 32),
    ("qmf_us", "f4", 640), ("qmf_ds", "f4", 320), ("noise", "f4", 1024),
    ("swb_offset_1024", "u2", (13, 64)), ("swb_offset_128", "u2", (13, 32)),
    ("num_swb_1024", "u1", 13), ("num_swb_128", "u1", 13), ("tns_max_bands_1024", "u1", 13), ("tns_max_bands_128", "u1", 13),
    ("pd_re_smooth", "f4", 512), ("pd_im_smooth", "f4", 512), ("HA", "f4", (46, 8, 4)), ("HB", "f4", (46, 8, 4)),
    ("f20_0_8", "f4", (8, 7, 2)), ("f34_0_12", "f4", (12, 7, 2)), ("f34_1_8", "f4", (8, 7, 2)), ("f34_2_4", "f4", (4, 7, 2)),
    ("Q_fract_allpass", "f4", (2, 50, 3, 2)), ("phi_fract", "f4", (2, 50, 2)), ("g1_Q2", "f4", 8),
    ("k_to_i_20", "i1", 72), ("k_to_i_34", "i1", 92), ("ps_inv_off", "u1", (2, 36)), ("ps_inv_k", "u1", (2, 92)),
], align=True)


@pytest.fixture(scope="module")
def product_tables():
    import ffmpeg_heaac_b200 as hb
    size = hb.lib().heaac_b200_sizeof(b"HeaacDevTables")
    assert size == TABLES.itemsize, "test mirror of HeaacDevTables is out of date (%d vs %d)" % (TABLES.itemsize, size)
    t = np.zeros(1, TABLES)
    pyemu.lib().heaac_emu_tables(t.ctypes.data_as(C.c_void_p))
    return t[0]


def test_windows_and_ps_tables_match_the_oracle(product_tables):
    for name in ("sine_1024", "sine_128", "kbd_1024", "kbd_128", "pd_re_smooth", "pd_im_smooth", "HA", "HB",
                 "f20_0_8", "f34_0_12", "f34_1_8", "f34_2_4", "Q_fract_allpass", "phi_fract"):
        a = pyoracle.table(name, np.float32)
        assert np.array_equal(a, product_tables[name].ravel()), name


@needs_ref
def test_tables_match_the_reference(product_tables):
    t = product_tables
    for name in ("pd_re_smooth", "pd_im_smooth", "HA", "HB", "f20_0_8", "f34_0_12", "f34_1_8", "f34_2_4",
                 "Q_fract_allpass", "phi_fract"):
        assert np.array_equal(pyref.table(name, np.float32), t[name].ravel()), name
    assert np.array_equal(pyref.table("sbr_qmf_window_us", np.float32), t["qmf_us"])
    assert np.array_equal(pyref.table("sbr_qmf_window_ds", np.float32), t["qmf_ds"])
    assert np.array_equal(pyref.table("sbr_noise_table", np.float32), t["noise"])
    assert np.array_equal(pyref.table("swb_offset_1024", np.uint16).reshape(13, 64), t["swb_offset_1024"])
    assert np.array_equal(pyref.table("swb_offset_128", np.uint16).reshape(13, 32), t["swb_offset_128"])
    for name in ("num_swb_1024", "num_swb_128", "tns_max_bands_1024", "tns_max_bands_128"):
        assert np.array_equal(pyref.table(name, np.uint8), t[name]), name
    assert np.array_equal(pyref.table("k_to_i_20", np.int8), t["k_to_i_20"][:71])
    assert np.array_equal(pyref.table("k_to_i_34", np.int8), t["k_to_i_34"][:91])
    assert np.array_equal(pyref.table("g1_Q2", np.float32), t["g1_Q2"][:7])
    k = np.zeros(1024, np.float32)
    pyref.lib().ref_kbd_window(k.ctypes.data_as(C.c_void_p), 4.0, 1024)
    assert np.array_equal(k, t["kbd_1024"])
    k = np.zeros(128, np.float32)
    pyref.lib().ref_kbd_window(k.ctypes.data_as(C.c_void_p), 6.0, 128)
    assert np.array_equal(k, t["kbd_128"])


@needs_ref
def test_mdct_twiddles_give_the_reference_transform():
    """ff_imdct_half known answers for the four MDCT instances of the path (mdct.c:124-159)."""
    rng = np.random.default_rng(5)
    for nbits, scale in ((11, 1.0), (8, 1.0), (7, 1.0 / 64), (7, -2.0)):
        n = 1 << nbits
        x = rng.standard_normal(n // 2).astype(np.float32)
        m = pyref.lib().ref_mdct_init(nbits, 1, scale)
        ref = np.zeros(n // 2, np.float32)
        pyref.lib().ref_imdct_half(m, ref.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p))
        pyref.lib().ref_mdct_end(m)
        got = pyoracle.imdct_half(nbits, scale, x)
        assert np.abs(got - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())


def test_imdct_matches_the_definition():
    """The O(N^2) definition the reference's own fft-test uses (fft-test.c:98-113):
    out[i] = -sum_k in[k] cos(pi (2i + 1 + N/2)(2k + 1) / (2N)); imdct_half is its middle half.
    The reference's tolerance there is 1e-3 (fft-test.c:187); we hold 1e-4 relative."""
    rng = np.random.default_rng(11)
    for nbits in (7, 8, 11):
        N = 1 << nbits
        x = rng.standard_normal(N // 2)
        i = np.arange(N)[:, None]
        k = np.arange(N // 2)[None, :]
        full = -(np.cos(np.pi * (2 * i + 1 + N / 2) * (2 * k + 1) / (2 * N)) * x[None, :]).sum(1)
        got = pyoracle.imdct_half(nbits, 1.0, x.astype(np.float32))
        assert np.abs(got - full[N // 4: N // 4 + N // 2]).max() <= 1e-4 * np.abs(full).max()
