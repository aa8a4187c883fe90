"""ADTS framing and MP4 esds helpers (csrc/heaac_container.c; adtsenc.c:84-147, aac_parser.c:29-67, isom.c)."""
import ctypes as C

import numpy as np
import pytest

import ffmpeg_heaac_b200 as hb
from common import golden_packets, load_golden
from test_frontend import adts_wrap


def adts_header(sr_index, chan_config, payload):
    out = np.zeros(7, np.uint8)
    assert hb.lib().heaac_adts_write_header(out.ctypes.data, sr_index, chan_config, payload) == 7
    return out.tobytes()


def test_adts_header_matches_the_bit_layout_and_round_trips():
    L = hb.lib()
    for sr, ch, n in ((4, 2, 371), (3, 6, 2000), (6, 1, 0), (11, 7, 8184)):
        hdr = adts_header(sr, ch, n)
        assert hdr == adts_wrap(b"", sr, ch)[:3] + hdr[3:]             # first bytes: sync, profile, rate, channel msb
        assert hdr == adts_wrap(b"\0" * n, sr, ch)[:7]                 # the whole header against the test suite's own writer
        v = [C.c_int() for _ in range(4)]
        b = np.frombuffer(hdr, np.uint8)
        assert L.heaac_adts_parse_header(b.ctypes.data, 7, *[C.byref(x) for x in v]) == 7
        assert [x.value for x in v] == [sr, ch, n + 7, 1]
    out = np.zeros(7, np.uint8)
    assert L.heaac_adts_write_header(out.ctypes.data, 13, 2, 10) == -22     # reserved sampling index
    assert L.heaac_adts_write_header(out.ctypes.data, 4, 2, 8186) == -22    # frame length field is 13 bits
    junk = np.asarray([0xFF, 0xE1, 0, 0, 0, 0, 0], np.uint8)
    assert L.heaac_adts_parse_header(junk.ctypes.data, 7, None, None, None, None) < 0


def test_adts_file_split_and_parse():
    """A .aac file image built from the golden LC packets: split it, then the front end parses every frame."""
    d = load_golden("lc_stereo_44k")
    pk = golden_packets(d)
    blob = b"".join(adts_header(int(d["sr_index"]), 2, len(p)) + p for p in pk) + b"\x12\x34"     # trailing junk
    buf = np.frombuffer(blob, np.uint8)
    off, sz = np.zeros(64, np.int32), np.zeros(64, np.int32)
    n = hb.lib().heaac_adts_split(buf.ctypes.data, buf.size, off.ctypes.data, sz.ctypes.data, 64)
    assert n == len(pk) and list(sz[:n]) == [len(p) + 7 for p in pk]
    raw, adts = hb.Frontend(d["extradata"].tobytes(), 44100, 2), hb.Frontend(b"", 0, 0)
    for i, p in enumerate(pk):
        a, b = raw.parse(p), adts.parse(blob[off[i]:off[i] + sz[i]])
        assert b["consumed"] == sz[i] and np.array_equal(a["coeffs"], b["coeffs"])
    raw.close(); adts.close()


def _descr(tag, payload, long_len=False):
    n = len(payload)
    ln = bytes([0x80, 0x80, 0x80, n]) if long_len else bytes([n])          # both length encodings occur in the wild
    return bytes([tag]) + ln + payload


@pytest.mark.parametrize("long_len", [False, True])
@pytest.mark.parametrize("with_version", [False, True])
def test_esds_yields_the_audio_specific_config(long_len, with_version):
    asc = bytes([0x13, 0x10, 0x56, 0xE5, 0x9D, 0x48, 0x80])                # LC 24 kHz stereo + explicit SBR + PS sync extensions
    dec_cfg = bytes([0x40, 0x15, 0, 0x18, 0, 0, 1, 0xF4, 0, 0, 1, 0xF4, 0]) + _descr(0x05, asc, long_len)
    es = bytes([0, 1, 0]) + _descr(0x04, dec_cfg, long_len) + _descr(0x06, b"\x02")
    esds = (b"\0\0\0\0" if with_version else b"") + _descr(0x03, es, long_len)
    buf = np.frombuffer(esds, np.uint8)
    p, n = C.c_void_p(), C.c_int()
    assert hb.lib().heaac_esds_find_asc(buf.ctypes.data, buf.size, C.byref(p), C.byref(n)) == 0
    got = bytes((C.c_uint8 * n.value).from_address(p.value))
    assert got == asc
    fe = hb.Frontend(got, 24000, 2)                                        # and the front end accepts it: explicit SBR
    assert fe.sbr == 1 and fe.sr_index == 6
    fe.close()
    bad = np.frombuffer(esds[:-8], np.uint8)                               # cuts into the AudioSpecificConfig
    assert hb.lib().heaac_esds_find_asc(bad.ctypes.data, bad.size, C.byref(p), C.byref(n)) < 0
