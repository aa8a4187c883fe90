"""B200-native HE-AAC synthesis back end: Python host mirror of the C-ABI (include/heaac_b200.h).

This package is a thin ctypes binding over `libheaac_b200.so` (hand-written sm_100a CUDA kernels +
C-ABI, built by `__graft_entry__.build()` / `csrc/Makefile`).  It mirrors the reference decoder's
interface for the one path it replaces -- `spectral_to_sample` + `float_to_int16_interleave`
(libavcodec/aacdec.c:1903-1933, 2096-2097) -- for a batch of independent streams.

There is NO CPU path: importing works anywhere, but `Context(...)` raises if the shared library is
missing or no sm_100 device is present.
"""
import ctypes as C
import os
import numpy as np

from . import abi  # noqa: F401
from .abi import (Batch, Config, make_config, ICS, TNS, SBR_ELEM, PS_FRAME,  # noqa: F401
                  TYPE_SCE, TYPE_CPE, TYPE_LFE, ONLY_LONG, LONG_START, EIGHT_SHORT, LONG_STOP)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libheaac_b200.so")
_lib = None

ERRORS = {-22: "HEAAC_E_INVAL", -12: "HEAAC_E_NOMEM", -19: "HEAAC_E_NODEV", -5: "HEAAC_E_LAUNCH", -38: "HEAAC_E_UNSUPPORTED"}

EXPORTS = [
    "heaac_b200_create", "heaac_b200_destroy", "heaac_b200_reset_state", "heaac_b200_n_core_channels",
    "heaac_b200_n_out_channels", "heaac_b200_frame_out_len", "heaac_b200_n_sbr_elems", "heaac_b200_synth",
    "heaac_b200_sync", "heaac_b200_launch_count", "heaac_b200_stream", "heaac_b200_dev_alloc", "heaac_b200_dev_free",
    "heaac_b200_dev_upload", "heaac_b200_dev_download", "heaac_b200_timer_start", "heaac_b200_timer_stop",
    "heaac_b200_debug_read", "heaac_b200_mdct_init", "heaac_b200_imdct_half", "heaac_b200_mdct_end",
    "heaac_b200_version", "heaac_b200_strerror", "heaac_b200_sizeof", "heaac_b200_sbr_tables",
    "heaac_b200_profile_enable", "heaac_b200_profile_read",
    "heaac_fe_open", "heaac_fe_close", "heaac_fe_parse", "heaac_fe_info",
    "heaac_b200_synth_submit", "heaac_b200_inputs_consumed", "heaac_b200_wait", "heaac_b200_chunk_frames",
    "heaac_b200_reference_out_slots", "heaac_b200_host_alloc", "heaac_b200_host_free",
    "heaac_dec_open", "heaac_dec_close", "heaac_dec_decode", "heaac_dec_info",
    "heaac_b200_parser_create", "heaac_b200_parser_destroy", "heaac_b200_parser_reset", "heaac_b200_parser_decode",
    "heaac_b200_parser_launch_count", "heaac_b200_parser_stats", "heaac_dec_parse_kernel_ms",
    "heaac_fe_parse_two_phase", "heaac_fe_copy_layout", "heaac_fe_fix_layout",
    "heaac_dec_submit", "heaac_dec_wait", "heaac_b200_parser_submit", "heaac_b200_parser_wait",
    "heaac_b200_ticket", "heaac_b200_ticket_wait", "heaac_b200_sbr_tables_device",
    "heaac_b200_sbr_init", "heaac_b200_sbr_apply", "heaac_b200_ps_apply", "heaac_b200_sbr_close",
    "heaac_fe_cce", "heaac_fe_layout_from_pce", "heaac_fe_parse_chunk", "heaac_fe_chunk_stats",
    "heaac_fe_sbr_downsampled", "heaac_adts_write_header", "heaac_adts_parse_header", "heaac_adts_split", "heaac_esds_find_asc",
]


class HeaacError(RuntimeError):
    def __init__(self, code, what):
        super().__init__("%s failed: %s (%d)" % (what, ERRORS.get(code, "error"), code))
        self.code = code


def lib():
    """Load libheaac_b200.so (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libheaac_b200.so not built (run __graft_entry__.build()); the product has no CPU path")
        L = C.CDLL(LIB_PATH)
        L.heaac_b200_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.heaac_b200_destroy.argtypes = [C.c_void_p]
        L.heaac_b200_destroy.restype = None
        L.heaac_b200_reset_state.argtypes = [C.c_void_p]
        for f in ("n_core_channels", "n_out_channels", "frame_out_len", "n_sbr_elems", "sync"):
            getattr(L, "heaac_b200_" + f).argtypes = [C.c_void_p]
        L.heaac_b200_synth.argtypes = [C.c_void_p, C.POINTER(Batch)]
        L.heaac_b200_launch_count.argtypes = [C.c_void_p]
        L.heaac_b200_launch_count.restype = C.c_longlong
        L.heaac_b200_stream.argtypes = [C.c_void_p]
        L.heaac_b200_stream.restype = C.c_size_t
        L.heaac_b200_dev_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.heaac_b200_dev_free.argtypes = [C.c_void_p, C.c_void_p]
        L.heaac_b200_dev_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.heaac_b200_dev_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.heaac_b200_timer_start.argtypes = [C.c_void_p]
        L.heaac_b200_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.heaac_b200_debug_read.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
        L.heaac_b200_mdct_init.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_double]
        L.heaac_b200_imdct_half.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.heaac_b200_imdct_half.restype = None
        L.heaac_b200_mdct_end.argtypes = [C.c_void_p]
        L.heaac_b200_mdct_end.restype = None
        L.heaac_b200_version.restype = C.c_char_p
        L.heaac_b200_strerror.restype = C.c_char_p
        L.heaac_b200_sizeof.argtypes = [C.c_char_p]
        L.heaac_b200_sbr_tables.argtypes = [C.c_int] * 8 + [C.c_void_p]
        L.heaac_fe_open.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.heaac_fe_close.argtypes = [C.c_void_p]
        L.heaac_fe_close.restype = None
        L.heaac_fe_parse.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.heaac_fe_parse_two_phase.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.heaac_fe_copy_layout.argtypes = [C.c_void_p, C.c_void_p]
        L.heaac_fe_fix_layout.argtypes = [C.c_void_p, C.c_void_p]
        L.heaac_fe_sbr_downsampled.argtypes = [C.c_void_p]
        L.heaac_fe_cce.argtypes = [C.c_void_p, C.c_void_p]
        L.heaac_fe_parse_chunk.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 8
        L.heaac_fe_chunk_stats.argtypes = [C.c_void_p]
        L.heaac_fe_chunk_stats.restype = None
        L.heaac_fe_layout_from_pce.argtypes = [C.c_void_p]
        L.heaac_fe_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 5 + [C.c_void_p]
        L.heaac_b200_synth_submit.argtypes = [C.c_void_p, C.POINTER(Batch)]
        L.heaac_b200_inputs_consumed.argtypes = [C.c_void_p]
        L.heaac_b200_wait.argtypes = [C.c_void_p]
        L.heaac_b200_chunk_frames.argtypes = [C.c_void_p]
        L.heaac_b200_reference_out_slots.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.heaac_b200_host_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.heaac_b200_host_free.argtypes = [C.c_void_p, C.c_void_p]
        L.heaac_dec_open.argtypes = [C.POINTER(abi.DecoderConfig), C.POINTER(C.c_void_p)]
        L.heaac_dec_close.argtypes = [C.c_void_p]
        L.heaac_dec_close.restype = None
        L.heaac_dec_decode.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.heaac_dec_submit.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.heaac_dec_submit.restype = C.c_longlong
        L.heaac_dec_wait.argtypes = [C.c_void_p, C.c_longlong]
        L.heaac_b200_ticket.argtypes = [C.c_void_p]
        L.heaac_b200_ticket.restype = C.c_longlong
        L.heaac_b200_ticket_wait.argtypes = [C.c_void_p, C.c_longlong]
        L.heaac_b200_sbr_tables_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.heaac_adts_write_header.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.heaac_adts_parse_header.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 4
        L.heaac_adts_split.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.heaac_esds_find_asc.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        L.heaac_b200_sbr_init.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
        L.heaac_b200_sbr_apply.argtypes = [C.c_void_p] * 5
        L.heaac_b200_ps_apply.argtypes = [C.c_void_p] * 4 + [C.c_int]
        L.heaac_b200_sbr_close.argtypes = [C.c_void_p]
        L.heaac_b200_sbr_close.restype = None
        L.heaac_dec_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4 + [C.POINTER(C.c_longlong)] * 2
        L.heaac_dec_parse_kernel_ms.argtypes = [C.c_void_p]
        L.heaac_dec_parse_kernel_ms.restype = C.c_double
        L.heaac_b200_profile_enable.argtypes = [C.c_void_p, C.c_int]
        L.heaac_b200_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_longlong)]
        for name, size in abi.SIZES.items():
            got = L.heaac_b200_sizeof(name.encode())
            if got != size:
                raise ImportError("ABI mismatch for %s: library %d, abi.py %d" % (name, got, size))
        if L.heaac_b200_sizeof(b"HeaacConfig") != C.sizeof(Config) or L.heaac_b200_sizeof(b"HeaacBatch") != C.sizeof(Batch) \
                or L.heaac_b200_sizeof(b"HeaacDecoderConfig") != C.sizeof(abi.DecoderConfig):
            raise ImportError("ABI mismatch for HeaacConfig/HeaacBatch/HeaacDecoderConfig")
        _lib = L
    return _lib


def _check(r, what):
    if r < 0:
        raise HeaacError(r, what)
    return r


class DeviceBuffer:
    """A device allocation owned by a Context (plain pointer + size, no torch types)."""

    def __init__(self, ctx, nbytes):
        self.ctx, self.nbytes = ctx, int(nbytes)
        p = C.c_void_p()
        _check(lib().heaac_b200_dev_alloc(ctx.h, self.nbytes, C.byref(p)), "dev_alloc")
        self.ptr = p.value

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        _check(lib().heaac_b200_dev_upload(self.ctx.h, self.ptr, arr.ctypes.data, arr.nbytes), "dev_upload")
        return self

    def download(self, dtype, shape):
        out = np.empty(shape, dtype)
        assert out.nbytes <= self.nbytes
        _check(lib().heaac_b200_dev_download(self.ctx.h, out.ctypes.data, self.ptr, out.nbytes), "dev_download")
        return out

    def free(self):
        if self.ptr:
            lib().heaac_b200_dev_free(self.ctx.h, self.ptr)
            self.ptr = None


class Context:
    """One batched decoder back end on one GPU (HeaacContext).

    synth(...) is the drop-in for the reference's per-frame `spectral_to_sample(ac)` +
    `float_to_int16_interleave` pair, for `n_frames` consecutive frames of `n_streams` streams.
    """

    def __init__(self, cfg):
        self.cfg = cfg
        h = C.c_void_p()
        _check(lib().heaac_b200_create(C.byref(cfg), C.byref(h)), "heaac_b200_create")
        self.h = h.value
        self.ncc = lib().heaac_b200_n_core_channels(self.h)
        self.nout = lib().heaac_b200_n_out_channels(self.h)
        self.out_len = lib().heaac_b200_frame_out_len(self.h)
        self.nse = lib().heaac_b200_n_sbr_elems(self.h)
        self.S = cfg.n_streams

    # -- host-pointer path: copies H2D / D2H inside the call (the e2e path) --
    def synth(self, coeffs, ics, tns=None, sbr=None, ps=None, want_float=False, pcm_out=None, cce=None):
        coeffs = np.ascontiguousarray(coeffs, np.float32)
        T = coeffs.size // (self.S * self.ncc * 1024)
        keep = [coeffs, np.ascontiguousarray(ics, ICS)]
        b = Batch()
        b.n_frames, b.on_device = T, 0
        b.coeffs, b.ics = keep[0].ctypes.data, keep[1].ctypes.data
        for name, arr, dt in (("tns", tns, TNS), ("sbr", sbr, SBR_ELEM), ("ps", ps, PS_FRAME), ("cce", cce, abi.CCE)):
            if arr is not None:
                a = np.ascontiguousarray(arr, dt)
                keep.append(a)
                setattr(b, name, a.ctypes.data)
        pcm = pcm_out if pcm_out is not None else np.empty((T, self.S, self.out_len, self.nout), np.int16)
        b.pcm = pcm.ctypes.data
        f32 = None
        if want_float:
            f32 = np.empty((T, self.S, self.nout, self.out_len), np.float32)
            b.pcm_f32 = f32.ctypes.data
        _check(lib().heaac_b200_synth(self.h, C.byref(b)), "heaac_b200_synth")
        return pcm, f32

    # -- device-resident path: pointers are DeviceBuffer.ptr; asynchronous --
    def synth_device(self, n_frames, coeffs, ics, pcm, tns=None, sbr=None, ps=None, pcm_f32=None):
        b = Batch()
        b.n_frames, b.on_device = n_frames, 1
        b.coeffs, b.ics, b.pcm = coeffs.ptr, ics.ptr, pcm.ptr if pcm is not None else None
        b.tns = tns.ptr if tns is not None else None
        b.sbr = sbr.ptr if sbr is not None else None
        b.ps = ps.ptr if ps is not None else None
        b.pcm_f32 = pcm_f32.ptr if pcm_f32 is not None else None
        _check(lib().heaac_b200_synth(self.h, C.byref(b)), "heaac_b200_synth")

    def alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def sync(self):
        _check(lib().heaac_b200_sync(self.h), "heaac_b200_sync")

    def reset_state(self):
        _check(lib().heaac_b200_reset_state(self.h), "heaac_b200_reset_state")

    def timer_start(self):
        _check(lib().heaac_b200_timer_start(self.h), "timer_start")

    def timer_stop(self):
        ms = C.c_float()
        _check(lib().heaac_b200_timer_stop(self.h, C.byref(ms)), "timer_stop")
        return ms.value

    def profile_enable(self, capacity):
        _check(lib().heaac_b200_profile_enable(self.h, capacity), "profile_enable")

    def profile_read(self):
        """{kernel: (total_ms, launches)} since profile_enable / the last read."""
        ms = (C.c_float * 4)()
        n = (C.c_longlong * 4)()
        _check(lib().heaac_b200_profile_read(self.h, ms, n), "profile_read")
        return {k: (ms[i], n[i]) for i, k in enumerate(("k_core", "k_chain", "k_ps", "k_synth"))}

    @property
    def launch_count(self):
        return lib().heaac_b200_launch_count(self.h)

    def debug_read(self, name, dtype, shape):
        out = np.zeros(shape, dtype)
        _check(lib().heaac_b200_debug_read(self.h, name.encode(), out.ctypes.data, out.nbytes), "debug_read")
        return out

    def close(self):
        if getattr(self, "h", None):
            lib().heaac_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sbr_tables(core_sample_rate, start_freq, stop_freq, xover=0, freq_scale=2, alter_scale=1, noise_bands=2,
               limiter_bands=2):
    """SBR frequency tables for a header (host code; the defaults are the bitstream defaults of
    read_sbr_header, aacsbr.c:233-252).  Returns a 1-element SBR_ELEM array with the table fields set."""
    e = np.zeros(1, SBR_ELEM)
    r = lib().heaac_b200_sbr_tables(core_sample_rate, start_freq, stop_freq, xover, freq_scale, alter_scale,
                                    noise_bands, limiter_bands, e.ctypes.data)
    _check(r, "heaac_b200_sbr_tables")
    return e


class Frontend:
    """The repo's own host parser for one stream (heaac_fe_*): packets -> per-frame C-ABI structures.
    Mirrors what the reference's aac_decode_frame does up to spectral_to_sample (aacdec.c:1973-2076)."""

    def __init__(self, extradata=b"", sample_rate=0, channels=0, bias_mode=False):
        h = C.c_void_p()
        ed = np.frombuffer(extradata, np.uint8).copy() if extradata else np.zeros(0, np.uint8)
        _check(lib().heaac_fe_open(C.byref(h), ed.ctypes.data if ed.size else None, ed.size, sample_rate, channels,
                                   int(bool(bias_mode))), "heaac_fe_open")
        self.h = h.value
        self._refresh()

    def _refresh(self):
        v = [C.c_int() for _ in range(5)]
        et = (C.c_uint8 * 8)()
        n = lib().heaac_fe_info(self.h, *[C.byref(x) for x in v], et)
        self.ncc, self.nse, self.sbr, self.ps, self.sr_index = [x.value for x in v]
        self.downsampled = bool(lib().heaac_fe_sbr_downsampled(self.h))
        self.elem_types = [int(et[i]) for i in range(max(n, 0))]
        self.n_cce = max(int(lib().heaac_fe_cce(self.h, None)), 0)
        self.from_pce = bool(lib().heaac_fe_layout_from_pce(self.h))

    def parse(self, pkt, two_phase=False):
        """Returns dict(coeffs [ncc,1024], ics [ncc], tns [ncc], sbr [nse], ps [1], consumed).
        two_phase: go through heaac_fe_parse_two_phase (the device front end's decomposition)."""
        buf = np.frombuffer(pkt, np.uint8)
        # sized for the largest configuration: an ADTS header inside this very packet may (re)configure the stream
        r = {"coeffs": np.zeros((16, 1024), np.float32), "ics": np.zeros(16, ICS), "tns": np.zeros(16, TNS),
             "sbr": np.zeros(8, SBR_ELEM), "ps": np.zeros(1, PS_FRAME)}
        fn = lib().heaac_fe_parse_two_phase if two_phase else lib().heaac_fe_parse
        n = fn(self.h, buf.ctypes.data, buf.size, r["coeffs"].ctypes.data, r["ics"].ctypes.data,
               r["tns"].ctypes.data, r["sbr"].ctypes.data, r["ps"].ctypes.data)
        _check(n, "heaac_fe_parse")
        self._refresh()
        for k in ("coeffs", "ics", "tns"):
            r[k] = r[k][:max(self.ncc, 1)]
        r["sbr"] = r["sbr"][:max(self.nse, 1)]
        r["consumed"] = n
        if self.n_cce:                                 # coupling side info of this frame (HeaacBatch.cce)
            r["cce"] = np.zeros(self.n_cce, abi.CCE)
            lib().heaac_fe_cce(self.h, r["cce"].ctypes.data)
        return r

    def parse_chunk(self, pkts):
        """Up to 32 consecutive packets the way the device front end treats a chunk (heaac_fe_parse_chunk).  Returns
        dict(coeffs [T,ncc,1024], ics [T,ncc], tns [T,ncc], sbr [T,nse], ps [T], consumed [T]); a failed frame has
        consumed < 0 and zeroed outputs.  The layout must be known (extradata)."""
        T, ncc, nse = len(pkts), self.ncc, max(self.nse, 1)
        bufs = [np.frombuffer(p, np.uint8) for p in pkts]
        ptrs = (C.c_void_p * T)(*[b.ctypes.data for b in bufs])
        sizes = np.asarray([b.size for b in bufs], np.int32)
        r = {"coeffs": np.zeros((T, ncc, 1024), np.float32), "ics": np.zeros((T, ncc), ICS), "tns": np.zeros((T, ncc), TNS),
             "sbr": np.zeros((T, nse), SBR_ELEM), "ps": np.zeros(T, PS_FRAME), "consumed": np.zeros(T, np.int32)}
        _check(lib().heaac_fe_parse_chunk(self.h, T, C.cast(ptrs, C.c_void_p), sizes.ctypes.data, r["coeffs"].ctypes.data,
                                          r["ics"].ctypes.data, r["tns"].ctypes.data, r["sbr"].ctypes.data,
                                          r["ps"].ctypes.data, r["consumed"].ctypes.data), "heaac_fe_parse_chunk")
        st = (C.c_int * 4)()
        lib().heaac_fe_chunk_stats(st)
        r["stats"] = {"passes": st[0], "one_pass_frames": st[1], "stale": st[2]}
        self._refresh()
        return r

    def close(self):
        if self.h:
            lib().heaac_fe_close(self.h)
            self.h = None


class PacketBatch:
    """n_frames x n_streams packets laid out the way heaac_dec_decode takes them (pointer + size arrays over one
    contiguous host blob).  Build it once; decoding does not copy the payload."""

    def __init__(self, packets):
        """packets[t][s] = bytes of stream s's frame t."""
        self.T, self.S = len(packets), len(packets[0])
        flat = [p for row in packets for p in row]
        assert len(flat) == self.T * self.S
        self.sizes = np.asarray([len(p) for p in flat], np.int32)
        pad = 8                                        # FF_INPUT_BUFFER_PADDING_SIZE-style zero tail per packet
        ofs = np.concatenate([[0], np.cumsum(self.sizes.astype(np.int64) + pad)])
        self.blob = np.zeros(int(ofs[-1]), np.uint8)
        for p, o in zip(flat, ofs[:-1]):
            self.blob[o:o + len(p)] = np.frombuffer(p, np.uint8)
        self.ptrs = (self.blob.ctypes.data + ofs[:-1]).astype(np.uint64)
        self.nbytes = int(self.sizes.sum())


class Decoder:
    """Batched packets -> PCM decoder (heaac_dec_*): S host front ends + the GPU back end, the counterpart of the
    reference's aac_decode_init / aac_decode_frame / aac_decode_close for n_streams streams."""

    def __init__(self, n_streams, extradata=b"", sample_rate=0, channels=0, max_frames=32, n_threads=0, device=0,
                 want_float=False, gpu_parse=False, chunk_frames=0):
        self._ed = np.frombuffer(extradata, np.uint8).copy() if extradata else np.zeros(0, np.uint8)
        cfg = abi.DecoderConfig()
        cfg.abi_version, cfg.device, cfg.n_streams = abi.ABI_VERSION, device, n_streams
        cfg.sample_rate, cfg.channels, cfg.max_frames_per_call = sample_rate, channels, max_frames
        cfg.n_threads, cfg.want_float_out, cfg.gpu_parse = n_threads, int(bool(want_float)), int(bool(gpu_parse))
        cfg.chunk_frames = int(chunk_frames)
        cfg.extradata, cfg.extradata_size = (self._ed.ctypes.data if self._ed.size else None), self._ed.size
        h = C.c_void_p()
        _check(lib().heaac_dec_open(C.byref(cfg), C.byref(h)), "heaac_dec_open")
        self.h, self.S, self.want_float = h.value, n_streams, bool(want_float)
        self.nout = self.out_len = None

    def info(self):
        v = [C.c_int() for _ in range(4)]
        w = [C.c_longlong() for _ in range(2)]
        ok = lib().heaac_dec_info(self.h, *[C.byref(x) for x in v], *[C.byref(x) for x in w])
        return {"configured": bool(ok), "n_out_channels": v[0].value, "frame_out_len": v[1].value, "sbr": v[2].value,
                "ps": v[3].value, "parse_errors": w[0].value, "gpu_launches": w[1].value,
                "k_parse_ms": lib().heaac_dec_parse_kernel_ms(self.h)}

    def decode(self, batch, pcm_out=None, out_shape=None):
        """batch: PacketBatch.  Returns (pcm s16 [T, S, out_len, nout], pcm_f32 or None, consumed [T, S]).
        Before the first call the output geometry is unknown: pass out_shape=(out_len, nout) or let the first call
        allocate for the largest case and trim."""
        T = batch.T
        assert batch.S == self.S
        consumed = np.zeros((T, self.S), np.int32)
        if self.nout is None and out_shape is None:
            cap = (2048, 16)
        else:
            cap = out_shape or (self.out_len, self.nout)
        pcm = pcm_out if pcm_out is not None else np.zeros((T, self.S, cap[0], cap[1]), np.int16)
        f32 = np.zeros((T, self.S, cap[1], cap[0]), np.float32) if self.want_float else None
        _check(lib().heaac_dec_decode(self.h, T, batch.ptrs.ctypes.data, batch.sizes.ctypes.data, pcm.ctypes.data,
                                      f32.ctypes.data if f32 is not None else None, consumed.ctypes.data), "heaac_dec_decode")
        i = self.info()
        self.nout, self.out_len = i["n_out_channels"], i["frame_out_len"]
        if pcm_out is None and cap != (self.out_len, self.nout):
            n = T * self.S * self.out_len * self.nout
            pcm = pcm.reshape(-1)[:n].reshape(T, self.S, self.out_len, self.nout).copy()
            if f32 is not None:
                f32 = f32.reshape(-1)[:n].reshape(T, self.S, self.nout, self.out_len).copy()
        return pcm, f32, consumed

    def submit(self, batch, pcm_out, consumed_out=None):
        """Asynchronous decode into caller-owned (page-locked) pcm_out; returns a ticket for wait()."""
        t = lib().heaac_dec_submit(self.h, batch.T, batch.ptrs.ctypes.data, batch.sizes.ctypes.data, pcm_out.ctypes.data, None,
                                   consumed_out.ctypes.data if consumed_out is not None else None)
        _check(t, "heaac_dec_submit")
        return t

    def wait(self, ticket):
        _check(lib().heaac_dec_wait(self.h, ticket), "heaac_dec_wait")

    def close(self):
        if self.h:
            lib().heaac_dec_close(self.h)
            self.h = None


class Mdct:
    """ff_mdct_init / ff_imdct_half / ff_mdct_end (mdct.c:61,124,226) as a batch-of-one device transform."""

    def __init__(self, nbits, inverse, scale):
        h = C.c_void_p()
        _check(lib().heaac_b200_mdct_init(C.byref(h), nbits, inverse, scale), "heaac_b200_mdct_init")
        self.h, self.n = h.value, 1 << nbits

    def imdct_half(self, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(self.n // 2, np.float32)
        lib().heaac_b200_imdct_half(self.h, out.ctypes.data, x.ctypes.data)
        return out

    def close(self):
        if self.h:
            lib().heaac_b200_mdct_end(self.h)
            self.h = None
