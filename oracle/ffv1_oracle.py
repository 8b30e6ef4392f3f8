"""ctypes binding of oracle/libffv1oracle.so (TEST INFRASTRUCTURE ONLY -- never imported by the product)."""
import ctypes, os, subprocess
import numpy as np
from . import pixfmt

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

class Params(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "width", "height", "version", "micro_version", "ac", "colorspace", "bits", "chroma_planes",
        "chroma_h_shift", "chroma_v_shift", "transparency", "layout", "packed_at_lsb", "context_model",
        "ec", "intra", "gop_size", "num_h_slices", "num_v_slices", "plane_count")] + [
        ("state_transition", ctypes.c_uint8 * 256),
        ("quant_tables", ctypes.c_int16 * (2 * 5 * 256)),
        ("context_count", ctypes.c_int * 2), ("force_pcm", ctypes.c_int)]

def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])

def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libffv1oracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.ffv1o_resolve.argtypes = [ctypes.POINTER(Params), ctypes.c_int, ctypes.c_int, ctypes.c_char_p] + [ctypes.c_int] * 6
        L.ffv1o_resolve_ex.argtypes = [ctypes.POINTER(Params), ctypes.c_int, ctypes.c_int, ctypes.c_char_p] + [ctypes.c_int] * 7
        L.ffv1o_write_extradata.argtypes = [ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_int]
        L.ffv1o_parse_extradata.argtypes = [ctypes.POINTER(Params), ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.ffv1o_encoder_new.restype = ctypes.c_void_p
        L.ffv1o_encoder_new.argtypes = [ctypes.POINTER(Params)]
        L.ffv1o_encode_frame.restype = ctypes.c_long
        L.ffv1o_encode_frame.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int),
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_long,
                                         ctypes.POINTER(ctypes.c_int)]
        L.ffv1o_encoder_decisions.restype = ctypes.c_uint64
        L.ffv1o_encoder_decisions.argtypes = [ctypes.c_void_p]
        L.ffv1o_encoder_free.argtypes = [ctypes.c_void_p]
        L.ffv1o_decoder_new.restype = ctypes.c_void_p
        L.ffv1o_decoder_new.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.ffv1o_decoder_params.restype = ctypes.POINTER(Params)
        L.ffv1o_decoder_params.argtypes = [ctypes.c_void_p]
        L.ffv1o_decode_frame.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.POINTER(ctypes.c_void_p),
                                         ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                         ctypes.POINTER(ctypes.c_uint64)]
        L.ffv1o_decoder_free.argtypes = [ctypes.c_void_p]
        L.ffv1o_slice_records.restype = ctypes.c_long
        L.ffv1o_slice_records.argtypes = [ctypes.POINTER(Params), ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int),
                                          ctypes.c_int, ctypes.c_void_p, ctypes.c_long]
        L.ffv1o_crc32.restype = ctypes.c_uint32
        L.ffv1o_crc32.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t]
        L.ffv1o_default_state_tables.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _LIB = L
    return _LIB

def split_planes(frame, pix_fmt, w, h):
    """frame: contiguous uint8 array of a tightly packed frame -> list of (2-D uint8 row views)"""
    out, off = [], 0
    for rows, rb in pixfmt.plane_shapes(pix_fmt, w, h):
        out.append(frame[off:off + rows * rb].reshape(rows, rb))
        off += rows * rb
    return out

def _plane_args(planes):
    ptrs = (ctypes.c_void_p * 4)(*([pl.ctypes.data for pl in planes] + [None] * (4 - len(planes))))
    strides = (ctypes.c_int * 4)(*([pl.strides[0] for pl in planes] + [0] * (4 - len(planes))))
    return ptrs, strides

def resolve(w, h, pix_fmt, gop=12, level=-1, coder=0, context=0, slices=0, slicecrc=-1, strict_experimental=0, force_pcm=0):
    """strict_experimental unlocks level 4; force_pcm (test knob) codes every version-4 slice in slice_coding_mode 1"""
    p = Params()
    r = lib().ffv1o_resolve_ex(ctypes.byref(p), w, h, pix_fmt.encode(), gop, level, coder, context, slices, slicecrc, strict_experimental)
    if r < 0:
        raise ValueError("ffv1o_resolve failed: %d" % r)
    p.force_pcm = force_pcm
    return p

def extradata(p):
    buf = ctypes.create_string_buffer(65536)
    n = lib().ffv1o_write_extradata(ctypes.byref(p), buf, 65536)
    if n < 0:
        raise ValueError(n)
    return buf.raw[:n]

class Encoder:
    def __init__(self, w, h, pix_fmt, **opts):
        self.h_ = None
        self.w, self.h, self.pix_fmt = w, h, pix_fmt
        self.params = resolve(w, h, pix_fmt, **opts)
        self.h_ = lib().ffv1o_encoder_new(ctypes.byref(self.params))
        self.cap = 65536 + pixfmt.frame_bytes(pix_fmt, w, h) * 4
        self.buf = ctypes.create_string_buffer(self.cap)
    @property
    def extradata(self):
        return extradata(self.params)
    def encode(self, frame, sar=(0, 1), picture_structure=3):
        planes = split_planes(np.ascontiguousarray(frame).view(np.uint8).reshape(-1), self.pix_fmt, self.w, self.h)
        ptrs, strides = _plane_args(planes)
        key = ctypes.c_int()
        n = lib().ffv1o_encode_frame(self.h_, ptrs, strides, sar[0], sar[1], picture_structure, self.buf, self.cap, ctypes.byref(key))
        if n < 0:
            raise RuntimeError("oracle encode failed %d" % n)
        return self.buf.raw[:n], bool(key.value)
    def set_force_pcm(self, on):
        """test knob: the following frames' slices are coded in slice_coding_mode 1 (version 4)"""
        lib().ffv1o_encoder_set_force_pcm.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib().ffv1o_encoder_set_force_pcm(self.h_, int(on))
    @property
    def decisions(self):
        return lib().ffv1o_encoder_decisions(self.h_)
    def close(self):
        if self.h_:
            lib().ffv1o_encoder_free(self.h_); self.h_ = None
    def __del__(self):
        self.close()

class Decoder:
    def __init__(self, w, h, pix_fmt, extradata=b""):
        self.h_ = None
        self.w, self.h, self.pix_fmt = w, h, pix_fmt
        self.h_ = lib().ffv1o_decoder_new(w, h, extradata, len(extradata))
        if not self.h_:
            raise ValueError("bad extradata")
    def decode(self, pkt, fill=0):
        """fill: value the output buffer holds before decoding (samples no slice covers keep it -- with subsampled
        chroma and slice edges off the chroma grid the reference leaves such samples uncoded)"""
        out = np.full(pixfmt.frame_bytes(self.pix_fmt, self.w, self.h), fill, np.uint8)
        planes = split_planes(out, self.pix_fmt, self.w, self.h)
        ptrs, strides = _plane_args(planes)
        key = ctypes.c_int(); dm = ctypes.c_uint64()
        r = lib().ffv1o_decode_frame(self.h_, pkt, len(pkt), ptrs, strides, ctypes.byref(key), ctypes.byref(dm))
        if r < 0:
            raise RuntimeError("oracle decode failed %d" % r)
        return out, bool(key.value), dm.value
    def close(self):
        if self.h_:
            lib().ffv1o_decoder_free(self.h_); self.h_ = None
    def __del__(self):
        self.close()

def slice_records(params, frame, pix_fmt, slice_index):
    w, h = params.width, params.height
    planes = split_planes(np.ascontiguousarray(frame).view(np.uint8).reshape(-1), pix_fmt, w, h)
    ptrs, strides = _plane_args(planes)
    cap = w * h * 4 + 16
    rec = np.zeros(cap, np.uint32)
    n = lib().ffv1o_slice_records(ctypes.byref(params), ptrs, strides, slice_index, rec.ctypes.data, cap)
    if n < 0:
        raise RuntimeError(n)
    return rec[:n]

def crc32(data, init=0):
    return lib().ffv1o_crc32(init, data, len(data))
