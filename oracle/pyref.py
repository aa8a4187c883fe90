"""ctypes bindings for oracle/_ref/libheaac_ref.so -- TEST INFRASTRUCTURE ONLY.

Drives the UNMODIFIED reference decoder/encoder (compiled from /root/reference by
oracle/Makefile) and captures, per frame, (a) what the reference's parsers hand to
spectral_to_sample (aacdec.c:2078) packed into the C-ABI structs, (b) its PCM output.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import ctypes as C
import os
import sys
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
sys.path.insert(0, _ROOT)
import ffmpeg_heaac_b200  # noqa: E402  (host-side ABI mirrors only; no kernels are called from here)
abi = ffmpeg_heaac_b200.abi

_LIB_PATH = os.path.join(_HERE, "_ref", "libheaac_ref.so")
PADDING = 8  # FF_INPUT_BUFFER_PADDING_SIZE (avcodec.h:3432)


def available():
    return os.path.exists(_LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_LIB_PATH)
        L.ref_dec_open.restype = C.c_void_p
        L.ref_dec_open.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_dec_close.argtypes = [C.c_void_p]
        L.ref_dec_set_capture.argtypes = [C.c_void_p] + [C.c_void_p] * 6 + [C.c_int]
        L.ref_dec_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        for f in ("ref_dec_channels", "ref_dec_sample_rate", "ref_dec_sbr", "ref_dec_ps", "ref_dec_sr_index"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.ref_dec_add_bias.argtypes = [C.c_void_p]; L.ref_dec_add_bias.restype = C.c_float
        L.ref_dec_sf_scale.argtypes = [C.c_void_p]; L.ref_dec_sf_scale.restype = C.c_float
        L.ref_dec_peek.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_int]
        L.ref_mdct_init.restype = C.c_void_p
        L.ref_mdct_init.argtypes = [C.c_int, C.c_int, C.c_double]
        L.ref_imdct_half.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_imdct_calc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_mdct_end.argtypes = [C.c_void_p]
        L.ref_kbd_window.argtypes = [C.c_void_p, C.c_float, C.c_int]
        L.ref_vector_fmul_window.argtypes = [C.c_void_p] * 4 + [C.c_float, C.c_int]
        L.ref_float_to_int16.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
        L.ref_enc_open.restype = C.c_void_p
        L.ref_enc_open.argtypes = [C.c_int, C.c_int, C.c_int]
        L.ref_enc_extradata.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_enc_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_enc_frame_bits.argtypes = [C.c_void_p]
        L.ref_enc_close.argtypes = [C.c_void_p]
        L.ref_table.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.ref_set_log_level.argtypes = [C.c_int]
        L.ref_dec_decode_many.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_dec_decode_many.restype = C.c_longlong
        _lib = L
    return _lib


def table(name, dtype, shape=None):
    """Copy one of the reference's data tables (oracle/ref_tables.c) into a numpy array."""
    p = C.c_void_p()
    n = lib().ref_table(name.encode(), C.byref(p))
    if n < 0:
        raise KeyError(name)
    buf = (C.c_char * n).from_address(p.value)
    a = np.frombuffer(bytes(buf), dtype=dtype).copy()
    return a.reshape(shape) if shape is not None else a


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


LAYOUTS = {
    # name -> (element types, element ids): bitstream element order of the reference's default
    # channel configurations (aacdec.c set_default_channel_config).
    "mono": ([abi.TYPE_SCE], [0]),
    "stereo": ([abi.TYPE_CPE], [0]),
    "5.1": ([abi.TYPE_SCE, abi.TYPE_CPE, abi.TYPE_CPE, abi.TYPE_LFE], [0, 0, 1, 0]),
}


def n_core_channels(elem_types):
    return sum(2 if t == abi.TYPE_CPE else 1 for t in elem_types)


class Encoder:
    """The reference's aacenc (AAC-LC only), used purely as a fixture generator."""

    def __init__(self, sample_rate, channels, bit_rate):
        self.h = lib().ref_enc_open(sample_rate, channels, bit_rate)
        if not self.h:
            raise RuntimeError("ref_enc_open failed")
        self.channels = channels
        ed = np.zeros(16, np.uint8)
        n = lib().ref_enc_extradata(self.h, _ptr(ed), 16)
        self.extradata = bytes(ed[:n])

    def encode(self, pcm):
        """pcm: int16 [n_samples, channels]; returns list of raw_data_block packets (bytes)."""
        pcm = np.ascontiguousarray(pcm, np.int16).reshape(-1, self.channels)
        nfr = pcm.shape[0] // 1024
        out = np.zeros(8192 * self.channels, np.uint8)
        pkts = []
        for i in range(nfr):
            fr = np.ascontiguousarray(pcm[i * 1024:(i + 1) * 1024])
            n = lib().ref_enc_encode(self.h, _ptr(fr), _ptr(out), out.size)
            if n < 0:
                raise RuntimeError("encode failed")
            if n > 0:
                pkts.append(bytes(out[:n]))
        return pkts

    def close(self):
        if self.h:
            lib().ref_enc_close(self.h)
            self.h = None


class Decoder:
    """One reference decoder context (AVCodec aac_decoder, aacdec.c:2128)."""

    def __init__(self, extradata, sample_rate, channels, layout, float_mode=True):
        et, ei = LAYOUTS[layout] if isinstance(layout, str) else layout
        self.elem_types, self.elem_ids = list(et), list(ei)
        ed = np.frombuffer(extradata, np.uint8).copy() if extradata else np.zeros(0, np.uint8)
        t = np.asarray(self.elem_types, np.int32)
        i = np.asarray(self.elem_ids, np.int32)
        self.h = lib().ref_dec_open(_ptr(ed) if ed.size else None, ed.size, sample_rate, channels,
                                    1 if float_mode else 0, len(et), _ptr(t), _ptr(i))
        if not self.h:
            raise RuntimeError("ref_dec_open failed")
        self.ncc = n_core_channels(self.elem_types)
        self.nse = sum(1 for x in self.elem_types if x != abi.TYPE_LFE)
        self.float_mode = float_mode
        self._out = np.zeros(2048 * 8 * 2 + 64, np.int16)

    def decode(self, pkt, capture=True):
        """Returns dict(pcm int16 [samples, ch], pcm_f32 [ch, samples] (float mode),
        coeffs/ics/tns/sbr/ps captured at the spectral_to_sample boundary)."""
        buf = np.zeros(len(pkt) + PADDING, np.uint8)
        buf[:len(pkt)] = np.frombuffer(pkt, np.uint8)
        r = {}
        if capture:
            r["coeffs"] = np.zeros((self.ncc, 1024), np.float32)
            r["ics"] = np.zeros(self.ncc, abi.ICS)
            r["tns"] = np.zeros(self.ncc, abi.TNS)
            r["sbr"] = np.zeros(max(self.nse, 1), abi.SBR_ELEM)
            r["ps"] = np.zeros(1, abi.PS_FRAME)
            n_cce = sum(1 for x in self.elem_types if x == abi.TYPE_CCE)
            if n_cce:
                r["cce"] = np.zeros(n_cce, abi.CCE)
                lib().ref_dec_set_capture_cce.argtypes = [C.c_void_p, C.c_void_p]
                lib().ref_dec_set_capture_cce(self.h, _ptr(r["cce"]))
        f32 = np.zeros(8 * 2048, np.float32)
        lib().ref_dec_set_capture(self.h, _ptr(r.get("coeffs")), _ptr(r.get("ics")), _ptr(r.get("tns")),
                                  _ptr(r.get("sbr")), _ptr(r.get("ps")), _ptr(f32), f32.size)
        nbytes = C.c_int(self._out.size * 2)
        ret = lib().ref_dec_decode(self.h, _ptr(buf), len(pkt), _ptr(self._out), C.byref(nbytes))
        if ret < 0:
            raise RuntimeError("reference decode failed (%d)" % ret)
        ch = lib().ref_dec_channels(self.h)
        ns = nbytes.value // 2 // max(ch, 1)
        r["pcm"] = self._out[:ns * ch].reshape(ns, ch).copy()
        if self.float_mode:
            r["pcm_f32"] = f32[:ns * ch].reshape(ch, ns).copy()
        r["consumed"] = ret
        return r

    def enable_dropin(self, libpath):
        """DROP-IN mode: from now on this decoder hands every frame to the product library (dlopen'ed `libpath`,
        heaac_b200_synth) where it would call spectral_to_sample (aacdec.c:2078) -- the unmodified front end and output
        stage with the B200 back end, in one process.  GPU box only."""
        L = lib()
        L.ref_dec_enable_dropin.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_dec_dropin_frames.restype = C.c_longlong
        L.ref_dec_dropin_frames.argtypes = [C.c_void_p]
        r = L.ref_dec_enable_dropin(self.h, libpath.encode())
        if r:
            raise RuntimeError("ref_dec_enable_dropin failed (%d)" % r)

    def dropin_frames(self):
        return int(lib().ref_dec_dropin_frames(self.h))

    def decode_many(self, pkts, repeat=1):
        """Decode the packet list `repeat` times in one C call (CPU-baseline timing); returns samples/channel."""
        sizes = np.asarray([len(p) for p in pkts], np.int32)
        offs = np.zeros(len(pkts), np.int32)
        blob = bytearray()
        for i, p in enumerate(pkts):
            offs[i] = len(blob)
            blob += p + bytes(PADDING)
        blob = np.frombuffer(bytes(blob), np.uint8)
        self._keep = (blob, offs, sizes)
        return lib().ref_dec_decode_many(self.h, _ptr(blob), _ptr(offs), _ptr(sizes), len(pkts), repeat)

    def peek(self, name, dtype, shape, elem=0, ch=0):
        a = np.zeros(shape, dtype)
        n = lib().ref_dec_peek(self.h, elem, ch, name.encode(), _ptr(a), a.nbytes)
        if n < 0:
            raise KeyError(name)
        return a

    @property
    def sbr(self):
        return lib().ref_dec_sbr(self.h)

    @property
    def ps(self):
        return lib().ref_dec_ps(self.h)

    @property
    def sr_index(self):
        return lib().ref_dec_sr_index(self.h)

    @property
    def channels(self):
        return lib().ref_dec_channels(self.h)

    @property
    def add_bias(self):
        return lib().ref_dec_add_bias(self.h)

    @property
    def sf_scale(self):
        return lib().ref_dec_sf_scale(self.h)

    def close(self):
        if self.h:
            lib().ref_dec_close(self.h)
            self.h = None


def lcg_noise(n, channels, seed=1234, div=16):
    """SURVEY 8d config 1 signal: int16 = ((lcg >> 16) - 32768) / div, x <- x*1664525 + 1013904223."""
    out = np.zeros((n, channels), np.int16)
    for c in range(channels):
        x = np.uint64((seed + 7919 * c) & 0xFFFFFFFF)
        v = np.zeros(n, np.int64)
        xs = int(x)
        for i in range(n):
            xs = (xs * 1664525 + 1013904223) & 0xFFFFFFFF
            v[i] = ((xs >> 16) - 32768) // div
        out[:, c] = v
    return out
