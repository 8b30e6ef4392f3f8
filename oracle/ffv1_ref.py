"""ctypes binding of oracle/_ref/libffv1ref.so = the UNMODIFIED reference FFV1 encoder/decoder compiled by
oracle/Makefile (TEST INFRASTRUCTURE ONLY).  available() is False when the library has not been built
(e.g. /root/reference absent and no prebuilt file travelled)."""
import ctypes, os, subprocess
import numpy as np
from . import pixfmt
from .ffv1_oracle import split_planes, _plane_args

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libffv1ref.so")
_LIB = None

def build(reference="/root/reference"):
    if os.path.isdir(reference):
        subprocess.check_call(["make", "-s", "-j8", "-C", _HERE, "ref", "REF=" + reference])

def available():
    return os.path.exists(_PATH)

def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(_PATH)
        L.ffv1ref_enc_open.restype = ctypes.c_void_p
        L.ffv1ref_enc_open.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p] + [ctypes.c_int] * 8
        L.ffv1ref_enc_extradata.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.ffv1ref_enc_frame.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int)] + \
            [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        L.ffv1ref_enc_close.argtypes = [ctypes.c_void_p]
        L.ffv1ref_dec_open.restype = ctypes.c_void_p
        L.ffv1ref_dec_open.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.ffv1ref_dec_packet.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                         ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]
        L.ffv1ref_dec_close.argtypes = [ctypes.c_void_p]
        L.ffv1ref_crc32_ieee.restype = ctypes.c_uint
        L.ffv1ref_crc32_ieee.argtypes = [ctypes.c_uint, ctypes.c_void_p, ctypes.c_int]
        _LIB = L
    return _LIB

class Encoder:
    def __init__(self, w, h, pix_fmt, gop=12, level=-1, coder=0, context=0, slices=0, slicecrc=-1, threads=1,
                 strict_experimental=0, two_pass=0, stats_in=None):
        """two_pass: 1 = first pass (AV_CODEC_FLAG_PASS1, statistics via stats_out()), 2 = second pass with stats_in"""
        self.h_ = None
        self.w, self.h, self.pix_fmt = w, h, pix_fmt
        if isinstance(stats_in, str):
            stats_in = stats_in.encode()
        L = lib()
        L.ffv1ref_enc_open_2pass.restype = ctypes.c_void_p
        L.ffv1ref_enc_open_2pass.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p] + [ctypes.c_int] * 9 + [ctypes.c_char_p]
        L.ffv1ref_enc_stats_out.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
        self.h_ = L.ffv1ref_enc_open_2pass(w, h, pix_fmt.encode(), gop, level, coder, context, slices, slicecrc, threads,
                                           strict_experimental, two_pass, stats_in)
        if not self.h_:
            raise ValueError("reference encoder refused these options")
        self.cap = 65536 + pixfmt.frame_bytes(pix_fmt, w, h) * 4
        self.buf = ctypes.create_string_buffer(self.cap)
    @property
    def extradata(self):
        b = ctypes.create_string_buffer(1 << 20)
        n = lib().ffv1ref_enc_extradata(self.h_, b, 1 << 20)
        return b.raw[:max(n, 0)]
    def stats_out(self):
        """flushes the encoder and returns AVCodecContext.stats_out (first pass)"""
        buf = ctypes.create_string_buffer(8 << 20)
        n = lib().ffv1ref_enc_stats_out(self.h_, buf, 8 << 20)
        if n < 0:
            raise RuntimeError("no statistics (%d)" % n)
        return buf.value.decode()
    def encode(self, frame, sar=(0, 1), interlaced=0, tff=0):
        planes = split_planes(np.ascontiguousarray(frame).view(np.uint8).reshape(-1), self.pix_fmt, self.w, self.h)
        ptrs, strides = _plane_args(planes)
        key = ctypes.c_int()
        n = lib().ffv1ref_enc_frame(self.h_, ptrs, strides, sar[0], sar[1], interlaced, tff, self.buf, self.cap, ctypes.byref(key))
        if n <= 0:
            raise RuntimeError("reference encode failed %d" % n)
        return self.buf.raw[:n], bool(key.value)
    def close(self):
        if self.h_:
            lib().ffv1ref_enc_close(self.h_); self.h_ = None
    def __del__(self):
        self.close()

class Decoder:
    def __init__(self, w, h, extradata=b"", threads=1, frame_threads=0):
        self.h_ = None
        self.w, self.h = w, h
        self.h_ = lib().ffv1ref_dec_open(w, h, extradata, len(extradata), threads, frame_threads)
        if not self.h_:
            raise ValueError("reference decoder refused extradata")
    def decode(self, pkt):
        cap = self.w * self.h * 8 + 64
        out = np.zeros(cap, np.uint8)
        name = ctypes.create_string_buffer(32); key = ctypes.c_int()
        n = lib().ffv1ref_dec_packet(self.h_, pkt, len(pkt), out.ctypes.data, cap, name, ctypes.byref(key))
        if n <= 0:
            raise RuntimeError("reference decode failed %d" % n)
        return out[:n].copy(), name.value.decode(), bool(key.value)
    def close(self):
        if self.h_:
            lib().ffv1ref_dec_close(self.h_); self.h_ = None
    def __del__(self):
        self.close()

def crc32(data, init=0):
    return lib().ffv1ref_crc32_ieee(init, data, len(data))

def fate_avi(raw, nframes, w, h, pix_fmt, level, slices):
    """encode + mux with the reference's AVI muxer exactly as FATE's enc_dec does; returns the AVI bytes"""
    L = lib()
    L.ffv1ref_fate_avi.restype = ctypes.c_int64
    L.ffv1ref_fate_avi.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
    out = np.zeros(len(raw) + (4 << 20), np.uint8)
    n = L.ffv1ref_fate_avi(raw.ctypes.data, nframes, w, h, pix_fmt.encode(), level, slices, out.ctypes.data, len(out))
    if n < 0:
        raise RuntimeError("fate_avi failed %d" % n)
    return out[:n].tobytes()

def mux(muxer, codec, raw, nframes, w, h, pix_fmt, level=-1, slices=0, batch=0):
    """encode with the named registered encoder and mux with the reference's "avi" or "nut" muxer, in memory"""
    L = lib()
    L.ffv1ref_mux_named.restype = ctypes.c_int64
    L.ffv1ref_mux_named.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
    raw = np.ascontiguousarray(raw)
    out = np.zeros(len(raw) + (4 << 20), np.uint8)
    n = L.ffv1ref_mux_named(muxer.encode(), codec.encode(), batch, raw.ctypes.data, nframes, w, h, pix_fmt.encode(), level, slices,
                            out.ctypes.data, len(out))
    if n < 0:
        raise RuntimeError("mux failed %d" % n)
    return out[:n].copy()

def nut_decode(decoder, data, max_bytes, opts=""):
    """the reference's NUT demuxer feeding the named registered decoder; returns (frames as one uint8 array, count, pix_fmt)"""
    L = lib()
    L.ffv1ref_nut_decode_named.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.POINTER(ctypes.c_int), ctypes.c_char_p]
    data = np.ascontiguousarray(data)
    dst = np.zeros(max_bytes + 1024, np.uint8)
    fb = ctypes.c_int(); name = ctypes.create_string_buffer(32)
    k = L.ffv1ref_nut_decode_named(decoder.encode(), opts.encode(), data.ctypes.data, len(data), dst.ctypes.data, len(dst), ctypes.byref(fb), name)
    if k < 0:
        raise RuntimeError("nut_decode failed %d" % k)
    return dst[:k * fb.value], k, name.value.decode()

def sws_convert(raw, w, h, dst_fmt, src_fmt="yuv420p"):
    """FATE's pixel-format conversion of a clip (ffmpeg's auto-inserted scale filter with -sws_flags
    neighbor+bitexact+accurate_rnd, tests/fate/vcodec.mak:119-127) by the reference's own libswscale"""
    L = lib()
    L.ffv1ref_sws_convert.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    ssz = L.ffv1ref_sws_convert(None, dst_fmt.encode(), None, src_fmt.encode(), w, h, 0)
    n = len(raw) // ssz
    dsz = L.ffv1ref_sws_convert(None, src_fmt.encode(), None, dst_fmt.encode(), w, h, 0)
    dst = np.zeros(dsz * n, np.uint8)
    r = L.ffv1ref_sws_convert(np.ascontiguousarray(raw).ctypes.data, src_fmt.encode(), dst.ctypes.data, dst_fmt.encode(), w, h, n)
    if r != dsz:
        raise RuntimeError("sws_convert failed %d" % r)
    return dst

def vsynth(name, reference="/root/reference"):
    """FATE's synthetic clips made by the reference's own generators (tests/videogen.c, tests/rotozoom.c,
    tests/Makefile:34-44), compiled into oracle/_ref/.  Returns (uint8 array, w, h)."""
    import tempfile
    d = os.path.join(_HERE, "_ref")
    for tool in ("videogen", "rotozoom"):
        exe = os.path.join(d, tool)
        if not os.path.exists(exe) and os.path.isdir(reference):
            subprocess.check_call(["gcc", "-O2", "-w", "-o", exe, os.path.join(reference, "tests", tool + ".c"), "-lm"])
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, name + ".yuv")
        if name == "vsynth1":
            subprocess.check_call([os.path.join(d, "videogen"), out]); w, h = 352, 288
        elif name == "vsynth3":
            subprocess.check_call([os.path.join(d, "videogen"), out, "34", "34"]); w, h = 34, 34
        elif name == "vsynth2":
            pre = os.path.join(d, "vsynth2.yuv")          # written by oracle/Makefile where the reference tree is mounted
            if os.path.exists(pre):
                return np.fromfile(pre, np.uint8), 352, 288
            subprocess.check_call([os.path.join(d, "rotozoom"), os.path.join(reference, "tests", "reference.pnm"), out]); w, h = 352, 288
        else:
            raise ValueError(name)
        return np.fromfile(out, np.uint8), w, h
