"""ctypes bindings for oracle/libheaac_oracle.so (the CPU restatement) -- TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import sys
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import ffmpeg_heaac_b200  # noqa: E402  (host-side ABI mirrors only; no kernels are called from here)
abi = ffmpeg_heaac_b200.abi

_LIB_PATH = os.path.join(_HERE, "libheaac_oracle.so")
_lib = None


def available():
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_LIB_PATH)
        L.heaac_oracle_create.restype = C.c_void_p
        L.heaac_oracle_create.argtypes = [C.POINTER(abi.Config)]
        L.heaac_oracle_destroy.argtypes = [C.c_void_p]
        L.heaac_oracle_synth.argtypes = [C.c_void_p, C.POINTER(abi.Batch), C.c_int]
        L.heaac_oracle_peek.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_int]
        L.heaac_oracle_imdct_half.argtypes = [C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.heaac_oracle_kbd_window.argtypes = [C.c_void_p, C.c_float, C.c_int]
        L.heaac_oracle_sine_window.argtypes = [C.c_void_p, C.c_int]
        L.heaac_oracle_table.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def table(name, dtype, shape=None):
    p = C.c_void_p()
    n = lib().heaac_oracle_table(name.encode(), C.byref(p))
    if n < 0:
        raise KeyError(name)
    a = np.frombuffer(bytes((C.c_char * n).from_address(p.value)), dtype=dtype).copy()
    return a.reshape(shape) if shape is not None else a


def imdct_half(nbits, scale, x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros((1 << nbits) // 2, np.float32)
    lib().heaac_oracle_imdct_half(nbits, scale, _ptr(out), _ptr(x))
    return out


class Oracle:
    """Same call shape as the product's HeaacContext: synth(frames) -> pcm."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.h = lib().heaac_oracle_create(C.byref(cfg))
        if not self.h:
            raise RuntimeError("heaac_oracle_create failed")
        et = [cfg.elem_type[i] for i in range(cfg.n_elems)]
        self.ncc = abi.n_core_channels(et)
        self.nse = abi.n_sbr_elems(et)
        self.nout = 2 if cfg.ps else self.ncc - sum(1 for t in et if t == abi.TYPE_CCE)
        self.out_len = 2048 if cfg.sbr and not cfg.sbr_downsampled else 1024

    def synth(self, coeffs, ics, tns=None, sbr=None, ps=None, n_threads=1, want_float=True, cce=None):
        """coeffs [T,S,ncc,1024] f32; ics [T,S,ncc]; tns [T,S,ncc]; sbr [T,S,nse]; ps [T,S].
        Returns (pcm int16 [T,S,out_len,nout], pcm_f32 [T,S,nout,out_len])."""
        S = self.cfg.n_streams
        coeffs = np.ascontiguousarray(coeffs, np.float32)
        T = coeffs.size // (S * self.ncc * 1024)
        ics = np.ascontiguousarray(ics, abi.ICS)
        keep = [coeffs, ics]
        b = abi.Batch()
        b.n_frames = T
        b.coeffs = coeffs.ctypes.data
        b.ics = ics.ctypes.data
        for name, arr, dt in (("tns", tns, abi.TNS), ("sbr", sbr, abi.SBR_ELEM), ("ps", ps, abi.PS_FRAME), ("cce", cce, abi.CCE)):
            if arr is not None:
                a = np.ascontiguousarray(arr, dt)
                keep.append(a)
                setattr(b, name, a.ctypes.data)
        pcm = np.zeros((T, S, self.out_len, self.nout), np.int16)
        b.pcm = pcm.ctypes.data
        f32 = None
        if want_float:
            f32 = np.zeros((T, S, self.nout, self.out_len), np.float32)
            b.pcm_f32 = f32.ctypes.data
        r = lib().heaac_oracle_synth(self.h, C.byref(b), n_threads)
        if r < 0:
            raise RuntimeError("heaac_oracle_synth failed (%d)" % r)
        return pcm, f32

    def peek(self, name, dtype, shape, stream=0, ch=0):
        a = np.zeros(shape, dtype)
        n = lib().heaac_oracle_peek(self.h, stream, ch, name.encode(), _ptr(a), a.nbytes)
        if n < 0:
            raise KeyError(name)
        return a

    def close(self):
        if self.h:
            lib().heaac_oracle_destroy(self.h)
            self.h = None
