"""ctypes binding for tests/emu/libheaac_emu.so -- TEST INFRASTRUCTURE ONLY (see heaac_emu.cpp)."""
import ctypes as C
import os
import subprocess
import sys
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
sys.path.insert(0, _ROOT)
import ffmpeg_heaac_b200  # noqa: E402  (host-side ABI mirrors only; no kernels are called from here)
abi = ffmpeg_heaac_b200.abi

_LIB = os.path.join(_HERE, "libheaac_emu.so")
_lib = None


def build():
    srcs = [os.path.join(_HERE, "heaac_emu.cpp"), os.path.join(_ROOT, "ffmpeg-heaac_b200", "csrc", "heaac_tables.cpp")]
    deps = srcs + [os.path.join(_ROOT, "ffmpeg-heaac_b200", "csrc", f) for f in ("heaac_kernels.cuh", "heaac_dev.cuh", "heaac_chain.cuh", "heaac_sbr.cuh", "heaac_regfft.cuh")]
    if os.path.exists(_LIB) and all(os.path.getmtime(_LIB) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-math-errno",
                           "-Wno-missing-braces", "-o", _LIB] + srcs)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.heaac_emu_create.restype = C.c_void_p
        L.heaac_emu_create.argtypes = [C.POINTER(abi.Config)]
        L.heaac_emu_destroy.argtypes = [C.c_void_p]
        L.heaac_emu_synth.argtypes = [C.c_void_p, C.POINTER(abi.Batch)]
        L.heaac_emu_tables.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class Emu:
    def __init__(self, cfg):
        self.cfg = cfg
        self.h = lib().heaac_emu_create(C.byref(cfg))
        et = [cfg.elem_type[i] for i in range(cfg.n_elems)]
        self.ncc = abi.n_core_channels(et)
        self.nout = 2 if cfg.ps else self.ncc - sum(1 for t in et if t == abi.TYPE_CCE)
        self.out_len = 2048 if cfg.sbr and not cfg.sbr_downsampled else 1024
        self.S = cfg.n_streams

    def synth(self, coeffs, ics, tns=None, sbr=None, ps=None, want_float=True, cce=None):
        coeffs = np.ascontiguousarray(coeffs, np.float32)
        T = coeffs.size // (self.S * self.ncc * 1024)
        keep = [coeffs, np.ascontiguousarray(ics, abi.ICS)]
        b = abi.Batch()
        b.n_frames = T
        b.coeffs, b.ics = keep[0].ctypes.data, keep[1].ctypes.data
        for name, arr, dt in (("tns", tns, abi.TNS), ("sbr", sbr, abi.SBR_ELEM), ("ps", ps, abi.PS_FRAME), ("cce", cce, abi.CCE)):
            if arr is not None:
                a = np.ascontiguousarray(arr, dt)
                keep.append(a)
                setattr(b, name, a.ctypes.data)
        pcm = np.zeros((T, self.S, self.out_len, self.nout), np.int16)
        b.pcm = pcm.ctypes.data
        f32 = np.zeros((T, self.S, self.nout, self.out_len), np.float32)
        b.pcm_f32 = f32.ctypes.data
        lib().heaac_emu_synth(self.h, C.byref(b))
        return pcm, f32

    def close(self):
        if self.h:
            lib().heaac_emu_destroy(self.h)
            self.h = None


def peek(emu, name, dtype, shape):
    L = lib()
    L.heaac_emu_peek.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
    a = np.zeros(shape, dtype)
    n = L.heaac_emu_peek(emu.h, name.encode(), a.ctypes.data, a.nbytes)
    if n < 0:
        raise KeyError(name)
    return a
