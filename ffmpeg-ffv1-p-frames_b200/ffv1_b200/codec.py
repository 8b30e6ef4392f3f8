"""ctypes binding of libffv1_b200.so.  See include/ffv1_b200.h for the C ABI."""
import ctypes, os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

class FFV1Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ffv1_b200 error %d: %s" % (code, msg))
        self.code = code

ERR_EINVAL, ERR_ENOMEM, ERR_ENOSYS = -22, -12, -38
ERR_INVALIDDATA, ERR_EXTERNAL, ERR_BUFFER_TOO_SMALL = -1094995529, -542398533, -1397118274

class _EncParams(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("pix_fmt", ctypes.c_char_p),
                ("gop_size", ctypes.c_int), ("level", ctypes.c_int), ("slices", ctypes.c_int),
                ("coder", ctypes.c_int), ("context", ctypes.c_int), ("slicecrc", ctypes.c_int),
                ("device", ctypes.c_int), ("max_batch_frames", ctypes.c_int), ("first_picture_number", ctypes.c_int64),
                ("flags", ctypes.c_int), ("stats_in", ctypes.c_char_p), ("strict_std_compliance", ctypes.c_int),
                ("bits_per_raw_sample", ctypes.c_int)]

FLAG_PASS1, FLAG_PASS2 = 1 << 9, 1 << 10      # AV_CODEC_FLAG_PASS1 / _PASS2

class _FrameProps(ctypes.Structure):
    _fields_ = [("sar_num", ctypes.c_int), ("sar_den", ctypes.c_int), ("picture_structure", ctypes.c_int)]

class Packet(ctypes.Structure):
    _fields_ = [("offset", ctypes.c_int64), ("size", ctypes.c_int32), ("flags", ctypes.c_int32), ("picture_number", ctypes.c_int64)]

class EncStats(ctypes.Structure):
    _fields_ = [("frames", ctypes.c_int), ("kernel_launches", ctypes.c_int), ("ms_total", ctypes.c_float),
                ("ms_pixel_kernel", ctypes.c_float), ("ms_model_kernel", ctypes.c_float), ("ms_coder_kernel", ctypes.c_float),
                ("ms_pack_kernel", ctypes.c_float), ("h2d_bytes", ctypes.c_int64), ("d2h_bytes", ctypes.c_int64),
                ("samples", ctypes.c_int64), ("decisions", ctypes.c_int64), ("packet_bytes", ctypes.c_int64), ("retries", ctypes.c_int)]

class _EncInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("version", "micro_version", "ac", "colorspace", "bits_per_raw_sample",
                "chroma_planes", "chroma_h_shift", "chroma_v_shift", "transparency", "num_h_slices", "num_v_slices",
                "slice_count", "ec", "intra", "context_count", "plane_count", "max_batch_frames")] + \
               [("samples_per_frame", ctypes.c_int64), ("frame_bytes", ctypes.c_int64)]

class _DecParams(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("extradata", ctypes.c_char_p),
                ("extradata_size", ctypes.c_int), ("device", ctypes.c_int), ("max_batch_frames", ctypes.c_int)]

class _DecInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("version", "micro_version", "ac", "colorspace", "bits_per_raw_sample",
                "chroma_planes", "chroma_h_shift", "chroma_v_shift", "transparency", "num_h_slices", "num_v_slices",
                "ec", "intra")] + [("pix_fmt", ctypes.c_char * 32), ("frame_bytes", ctypes.c_int64)]

class DecStats(ctypes.Structure):
    _fields_ = [("frames", ctypes.c_int), ("kernel_launches", ctypes.c_int), ("ms_total", ctypes.c_float),
                ("ms_decode_kernel", ctypes.c_float), ("h2d_bytes", ctypes.c_int64), ("d2h_bytes", ctypes.c_int64)]

def library_path():
    return os.path.join(os.path.dirname(_HERE), "libffv1_b200.so")

def lib():
    """Loads the CUDA library; raises if it was not built (no fallback implementation exists)."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise FFV1Error(ERR_EXTERNAL, "%s is missing: build it with `make -C %s` (or __graft_entry__.build()); "
                            "ffv1_b200 has no CPU fallback" % (path, os.path.dirname(path)))
        L = ctypes.CDLL(path)
        L.ffv1b200_version.restype = ctypes.c_char_p
        L.ffv1b200_strerror.restype = ctypes.c_char_p
        L.ffv1b200_strerror.argtypes = [ctypes.c_int]
        L.ffv1b200_last_error.restype = ctypes.c_char_p
        L.ffv1b200_enc_open.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(_EncParams)]
        L.ffv1b200_enc_close.argtypes = [ctypes.c_void_p]
        L.ffv1b200_enc_close.restype = None
        L.ffv1b200_enc_extradata.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int)]
        L.ffv1b200_enc_info.argtypes = [ctypes.c_void_p, ctypes.POINTER(_EncInfo)]
        L.ffv1b200_enc_set_frame_props.argtypes = [ctypes.c_void_p, ctypes.POINTER(_FrameProps)]
        L.ffv1b200_enc_set_frame_props.restype = None
        L.ffv1b200_enc_encode_host.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int),
                                               ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(Packet), ctypes.POINTER(ctypes.c_size_t)]
        L.ffv1b200_enc_submit_host.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int)]
        L.ffv1b200_enc_collect.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(Packet), ctypes.POINTER(ctypes.c_size_t)]
        L.ffv1b200_enc_collect_async.argtypes = L.ffv1b200_enc_collect.argtypes
        L.ffv1b200_enc_sync_output.argtypes = [ctypes.c_void_p]
        L.ffv1b200_bind_thread_to_device.argtypes = [ctypes.c_int]
        L.ffv1b200_enc_pending.argtypes = [ctypes.c_void_p]
        L.ffv1b200_enc_encode_device.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int),
                                                 ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(Packet), ctypes.POINTER(ctypes.c_size_t),
                                                 ctypes.c_void_p]
        L.ffv1b200_enc_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(EncStats)]
        L.ffv1b200_enc_stats_out.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        L.ffv1b200_enc_debug_records.restype = ctypes.c_int64
        L.ffv1b200_enc_debug_records.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
        L.ffv1b200_dec_open.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(_DecParams)]
        L.ffv1b200_dec_close.argtypes = [ctypes.c_void_p]
        L.ffv1b200_dec_close.restype = None
        L.ffv1b200_dec_info.argtypes = [ctypes.c_void_p, ctypes.POINTER(_DecInfo)]
        L.ffv1b200_dec_decode_host.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int),
                                               ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_uint64)]
        L.ffv1b200_dec_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(DecStats)]
        _LIB = L
    return _LIB

def _check(r):
    if r < 0:
        L = lib()
        msg = L.ffv1b200_last_error().decode() or L.ffv1b200_strerror(r).decode()
        raise FFV1Error(r, msg)
    return r

def device_count():
    return _check(lib().ffv1b200_device_count())

def bind_thread_to_device(device):
    """Pins the calling thread (and its future allocations) to the NUMA node of the GPU; returns the node or None."""
    r = lib().ffv1b200_bind_thread_to_device(device)
    return r if r >= 0 else None

# --- tightly packed frame geometry (what av_image_copy_to_buffer(align=1) produces) -------------------
def plane_shapes(pix_fmt, w, h):
    bps = 2 if pix_fmt.endswith("le") else 1
    if pix_fmt in ("bgr0", "bgra"):
        return [(h, w * 4)]
    if pix_fmt == "ya8":
        return [(h, w * 2)]
    if pix_fmt.startswith("gray"):
        return [(h, w * bps)]
    if pix_fmt.startswith("gbrp"):
        return [(h, w * bps)] * 3
    sub = {"420": (1, 1), "422": (1, 0), "444": (0, 0), "440": (0, 1), "411": (2, 0), "410": (2, 2)}
    for k, (hs, vs) in sub.items():
        if k in pix_fmt:
            c = (-((-h) >> vs), -((-w) >> hs) * bps)
            out = [(h, w * bps), c, c]
            if pix_fmt.startswith("yuva"):
                out.append((h, w * bps))
            return out
    raise ValueError("unknown pix_fmt " + pix_fmt)

def frame_bytes(pix_fmt, w, h):
    return sum(r * b for r, b in plane_shapes(pix_fmt, w, h))

def resolve_encoder(width, height, pix_fmt, g=12, level=-1, coder=0, context=0, slices=0, slicecrc=-1, flags=0, stats_in=None, strict=0,
                    bits_per_raw_sample=0):
    """encode_init's host half (no GPU needed): returns (info, extradata) or raises FFV1Error like FFV1Encoder would"""
    if isinstance(stats_in, str):
        stats_in = stats_in.encode()
    p = _EncParams(width, height, pix_fmt.encode(), g, level, slices, coder, context, slicecrc, 0, 0, 0, flags, stats_in, strict,
                   bits_per_raw_sample)
    info, n = _EncInfo(), ctypes.c_int()
    L = lib()
    L.ffv1b200_enc_resolve.argtypes = [ctypes.POINTER(_EncParams), ctypes.POINTER(_EncInfo), ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    buf = ctypes.create_string_buffer(1 << 20)
    _check(L.ffv1b200_enc_resolve(ctypes.byref(p), ctypes.byref(info), buf, 1 << 20, ctypes.byref(n)))
    return info, buf.raw[:n.value]

class FFV1Encoder:
    """Mirror of ff_ffv1_encoder (ffv1enc.c:1415-1444): __init__ = init, encode2/flush = encode2 (CAP_DELAY), close."""

    def __init__(self, width, height, pix_fmt, g=12, level=-1, coder=0, context=0, slices=0, slicecrc=-1,
                 device=0, max_batch_frames=64, first_picture_number=0, flags=0, stats_in=None, strict=0, bits_per_raw_sample=0):
        self._h = ctypes.c_void_p()
        self.width, self.height, self.pix_fmt = width, height, pix_fmt
        if isinstance(stats_in, str):
            stats_in = stats_in.encode()
        p = _EncParams(width, height, pix_fmt.encode(), g, level, slices, coder, context, slicecrc, device,
                       max_batch_frames, first_picture_number, flags, stats_in, strict, bits_per_raw_sample)
        _check(lib().ffv1b200_enc_open(ctypes.byref(self._h), ctypes.byref(p)))
        self.info = _EncInfo()
        _check(lib().ffv1b200_enc_info(self._h, ctypes.byref(self.info)))
        self.max_batch = self.info.max_batch_frames
        self._shapes = plane_shapes(pix_fmt, width, height)
        self._pending = []          # frames held back (AV_CODEC_CAP_DELAY)
        self._ready = []
        self._out = None

    @property
    def extradata(self):
        ptr, n = ctypes.c_void_p(), ctypes.c_int()
        _check(lib().ffv1b200_enc_extradata(self._h, ctypes.byref(ptr), ctypes.byref(n)))
        return ctypes.string_at(ptr, n.value) if n.value else b""

    def stats_out(self):
        """first pass (flags=FLAG_PASS1): AVCodecContext.stats_out as the reference leaves it when flushed (ffv1enc.c:1235-1277)"""
        need = ctypes.c_size_t()
        cap = 1 << 20
        while True:
            buf = ctypes.create_string_buffer(cap)
            r = lib().ffv1b200_enc_stats_out(self._h, buf, cap, ctypes.byref(need))
            if r == ERR_BUFFER_TOO_SMALL:
                cap = need.value
                continue
            _check(r)
            return buf.value.decode()

    def set_frame_props(self, sar=(0, 1), picture_structure=3):
        fp = _FrameProps(sar[0], sar[1], picture_structure)
        lib().ffv1b200_enc_set_frame_props(self._h, ctypes.byref(fp))

    def _out_buffer(self, nbytes):
        if self._out is None or self._out.nbytes < nbytes:
            self._out = np.empty(nbytes, np.uint8)
        return self._out

    def encode_batch(self, frames):
        """frames: sequence of contiguous uint8 arrays, each one tightly packed frame. Returns [(bytes, key)]."""
        out = []
        for i in range(0, len(frames), self.max_batch):
            out += self._encode_chunk(frames[i:i + self.max_batch])
        return out

    def _encode_chunk(self, frames):
        n = len(frames)
        fb = int(self.info.frame_bytes)
        planes = (ctypes.c_void_p * (4 * n))()
        ls = (ctypes.c_int * (4 * n))()
        keep = []
        for f, fr in enumerate(frames):
            a = np.ascontiguousarray(fr).view(np.uint8).reshape(-1)
            if a.nbytes != fb:
                raise ValueError("frame %d has %d bytes, expected %d" % (f, a.nbytes, fb))
            keep.append(a)
            off = 0
            for i, (rows, rb) in enumerate(self._shapes):
                planes[4 * f + i] = a.ctypes.data + off
                ls[4 * f + i] = rb
                off += rows * rb
        cap = n * (fb + fb // 4 + 65536)
        pk = (Packet * n)()
        needed = ctypes.c_size_t()
        while True:
            buf = self._out_buffer(cap)
            r = lib().ffv1b200_enc_encode_host(self._h, n, planes, ls, buf.ctypes.data, cap, pk, ctypes.byref(needed))
            if r == ERR_BUFFER_TOO_SMALL:
                cap = int(needed.value) + 4096
                continue
            _check(r)
            break
        return [(buf[pk[i].offset:pk[i].offset + pk[i].size].tobytes(), bool(pk[i].flags & 1)) for i in range(n)]

    # -- pipelined form: up to two batches in flight, copies of one overlap the kernels of the other -------------
    def _plane_table(self, frames):
        n = len(frames)
        fb = int(self.info.frame_bytes)
        planes = (ctypes.c_void_p * (4 * n))()
        ls = (ctypes.c_int * (4 * n))()
        keep = []
        for f, fr in enumerate(frames):
            a = np.ascontiguousarray(fr).view(np.uint8).reshape(-1)
            if a.nbytes != fb:
                raise ValueError("frame %d has %d bytes, expected %d" % (f, a.nbytes, fb))
            keep.append(a)
            off = 0
            for i, (rows, rb) in enumerate(self._shapes):
                planes[4 * f + i] = a.ctypes.data + off
                ls[4 * f + i] = rb
                off += rows * rb
        return planes, ls, keep

    def prepare(self, frames):
        """Build the plane-pointer table of a batch once (reusable across submits of the same host buffers)."""
        planes, ls, keep = self._plane_table(frames)
        return (planes, ls, keep, len(frames))

    def submit(self, frames):
        """Queue one batch (<= max_batch frames, or a prepare()d table); the frames must stay alive until collect()."""
        planes, ls, keep, n = frames if isinstance(frames, tuple) else self.prepare(frames)
        _check(lib().ffv1b200_enc_submit_host(self._h, n, planes, ls))
        self._inflight = getattr(self, "_inflight", [])
        self._inflight.append((keep, n))

    def collect(self, out=None, copy=True, wait_bytes=True):
        """Wait for the oldest submitted batch. Returns [(bytes, key)] (or the Packet array when copy=False).
        wait_bytes=False (needs out= and copy=False): return once the packets' device->host copy is queued; the bytes are
        complete after the next collect() / sync_output()."""
        if not wait_bytes and (out is None or copy):
            raise ValueError("wait_bytes=False needs a caller-owned out array and copy=False")
        fn = lib().ffv1b200_enc_collect if wait_bytes else lib().ffv1b200_enc_collect_async
        keep, n = self._inflight.pop(0)
        fb = int(self.info.frame_bytes)
        cap = n * (fb + fb // 4 + 65536) if out is None else out.nbytes
        pk = (Packet * n)()
        needed = ctypes.c_size_t()
        while True:
            buf = self._out_buffer(cap) if out is None else out
            r = fn(self._h, buf.ctypes.data, cap, pk, ctypes.byref(needed))
            if r == ERR_BUFFER_TOO_SMALL and out is None:
                cap = int(needed.value) + 4096
                continue
            _check(r)
            break
        if not copy:
            return pk
        return [(buf[pk[i].offset:pk[i].offset + pk[i].size].tobytes(), bool(pk[i].flags & 1)) for i in range(n)]

    def sync_output(self):
        _check(lib().ffv1b200_enc_sync_output(self._h))

    def pending(self):
        return _check(lib().ffv1b200_enc_pending(self._h))

    def encode_device(self, plane_ptrs, linesizes, d_out_ptr, d_out_cap, nframes, stream=None):
        """Frames already in device memory. plane_ptrs/linesizes: flat sequences of 4*nframes ints.
        Returns the Packet array (offsets into d_out)."""
        planes = plane_ptrs if isinstance(plane_ptrs, ctypes.Array) else (ctypes.c_void_p * (4 * nframes))(*plane_ptrs)
        ls = linesizes if isinstance(linesizes, ctypes.Array) else (ctypes.c_int * (4 * nframes))(*linesizes)
        pk = (Packet * nframes)()
        needed = ctypes.c_size_t()
        _check(lib().ffv1b200_enc_encode_device(self._h, nframes, planes, ls, d_out_ptr, d_out_cap, pk, ctypes.byref(needed), stream))
        return pk

    # -- AVCodec.encode2 semantics with AV_CODEC_CAP_DELAY: frames are held until a batch is full --------
    def encode2(self, frame):
        """Returns (packet_bytes, key) or None (got_packet = 0). Pass frame=None to drain (ffmpeg.c:1698-1770)."""
        if frame is not None:
            self._pending.append(np.array(frame, copy=True))
            if len(self._pending) >= self.max_batch:
                self._ready += self._encode_chunk(self._pending)
                self._pending = []
        elif self._pending:
            self._ready += self._encode_chunk(self._pending)
            self._pending = []
        return self._ready.pop(0) if self._ready else None

    def flush(self):
        out = []
        while True:
            p = self.encode2(None)
            if p is None:
                return out
            out.append(p)

    def stats(self):
        s = EncStats()
        _check(lib().ffv1b200_enc_stats(self._h, ctypes.byref(s)))
        return s

    def debug_records(self, frame_in_batch, slice_index):
        cap = self.width * self.height * 4 + 64
        rec = np.zeros(cap, np.uint32)
        n = _check(lib().ffv1b200_enc_debug_records(self._h, frame_in_batch, slice_index, rec.ctypes.data, cap))
        return rec[:n]

    def close(self):
        if getattr(self, "_h", None):
            lib().ffv1b200_enc_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

class FFV1Uploader:
    """On-GPU input preparation (hwupload_cuda + the format conversion of scale_npp, include/ffv1_b200.h): host frames in a
    capture / hardware-decoder layout -> device frames in one of the encoder's pix_fmts."""

    def __init__(self, width, height, src_pix_fmt, dst_pix_fmt=None, pool_frames=16, device=0):
        L = lib()
        L.ffv1b200_upload_open.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int]
        L.ffv1b200_upload_close.argtypes = [ctypes.c_void_p]
        L.ffv1b200_upload_close.restype = None
        L.ffv1b200_upload_pix_fmt.argtypes = [ctypes.c_void_p]
        L.ffv1b200_upload_pix_fmt.restype = ctypes.c_char_p
        L.ffv1b200_upload_frames.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int),
                                             ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int)]
        L.ffv1b200_enc_encode_cuda.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int),
                                               ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(Packet), ctypes.POINTER(ctypes.c_size_t)]
        self._h = ctypes.c_void_p()
        self.pool = pool_frames
        _check(L.ffv1b200_upload_open(ctypes.byref(self._h), src_pix_fmt.encode(), dst_pix_fmt.encode() if dst_pix_fmt else None,
                                      width, height, pool_frames, device))
        self.pix_fmt = L.ffv1b200_upload_pix_fmt(self._h).decode()

    def upload(self, frames, plane_rows_bytes):
        """frames: contiguous uint8 arrays (planes back to back, tightly packed); plane_rows_bytes: [(rows, row bytes)] of the
        source layout.  Returns (device plane pointer table, linesizes) for FFV1Encoder.encode_cuda."""
        n = len(frames)
        planes = (ctypes.c_void_p * (4 * n))()
        ls = (ctypes.c_int * (4 * n))()
        keep = []
        for f, fr in enumerate(frames):
            a = np.ascontiguousarray(fr).view(np.uint8).reshape(-1)
            keep.append(a)
            off = 0
            for i, (rows, rb) in enumerate(plane_rows_bytes):
                planes[4 * f + i] = a.ctypes.data + off
                ls[4 * f + i] = rb
                off += rows * rb
        dpl = (ctypes.c_void_p * (4 * n))()
        dls = (ctypes.c_int * 4)()
        _check(lib().ffv1b200_upload_frames(self._h, n, planes, ls, dpl, dls))
        return dpl, dls

    def close(self):
        if self._h:
            lib().ffv1b200_upload_close(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

def encode_cuda(enc, d_planes, d_linesizes, nframes):
    """AV_PIX_FMT_CUDA frames (device plane pointers, one linesize set) -> [(packet bytes, key)] (ffv1b200_enc_encode_cuda)"""
    fb = int(enc.info.frame_bytes)
    cap = nframes * (fb + fb // 4 + 65536)
    out = np.empty(cap, np.uint8)
    pk = (Packet * nframes)()
    need = ctypes.c_size_t()
    ls = (ctypes.c_int * (4 * nframes))(*([d_linesizes[i] for i in range(4)] * nframes))
    L = lib()
    L.ffv1b200_enc_encode_cuda.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int),
                                           ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(Packet), ctypes.POINTER(ctypes.c_size_t)]
    _check(L.ffv1b200_enc_encode_cuda(enc._h, nframes, d_planes, ls, out.ctypes.data, cap, pk, ctypes.byref(need)))
    return [(out[p.offset:p.offset + p.size].tobytes(), bool(p.flags & 1)) for p in pk]

class FFV1Decoder:
    """Mirror of ff_ffv1_decoder (ffv1dec.c:1139-1153): __init__ = decode_init, decode/decode_batch = decode_frame."""

    def __init__(self, width, height, extradata=b"", device=0, max_batch_frames=64):
        self._h = ctypes.c_void_p()
        self._extradata = bytes(extradata)
        p = _DecParams(width, height, self._extradata, len(self._extradata), device, max_batch_frames)
        _check(lib().ffv1b200_dec_open(ctypes.byref(self._h), ctypes.byref(p)))
        self.info = _DecInfo()
        _check(lib().ffv1b200_dec_info(self._h, ctypes.byref(self.info)))
        self.pix_fmt = self.info.pix_fmt.decode()
        self.width, self.height = width, height
        self.max_batch = max_batch_frames

    def decode_batch(self, packets, out=None):
        """packets: list of bytes. Returns [(uint8 frame array, key_frame, damaged_slice_mask)].
        out: optional uint8 array (e.g. pinned host memory) of len(packets) * frame_bytes the frames are written to."""
        res = []
        fb = int(self.info.frame_bytes)
        for i in range(0, len(packets), self.max_batch):
            o = out[i * fb:(i + self.max_batch) * fb] if (out is not None and fb) else None
            res += self._decode_chunk(packets[i:i + self.max_batch], o)
        return res

    def _decode_chunk(self, packets, out=None):
        n = len(packets)
        bufs = [np.frombuffer(p, np.uint8) for p in packets]
        ptrs = (ctypes.c_void_p * n)(*[b.ctypes.data for b in bufs])
        sizes = (ctypes.c_int * n)(*[len(p) for p in packets])
        if not self.info.frame_bytes:
            # FFV1 version 0/1: the stream parameters arrive with the first keyframe; a probe call parses them
            probe = np.empty(16, np.uint8)
            r = lib().ffv1b200_dec_decode_host(self._h, n, ptrs, sizes, probe.ctypes.data, 0, None, None)
            if r != ERR_BUFFER_TOO_SMALL:
                _check(r)
            _check(lib().ffv1b200_dec_info(self._h, ctypes.byref(self.info)))
            self.pix_fmt = self.info.pix_fmt.decode()
        fb = int(self.info.frame_bytes)
        if out is None:
            out = np.empty(n * fb, np.uint8)
        elif out.nbytes < n * fb or out.dtype != np.uint8:
            raise FFV1Error(ERR_BUFFER_TOO_SMALL, "output array too small")
        keys = (ctypes.c_int * n)()
        dmg = (ctypes.c_uint64 * n)()
        _check(lib().ffv1b200_dec_decode_host(self._h, n, ptrs, sizes, out.ctypes.data, out.nbytes, keys, dmg))
        return [(out[i * fb:(i + 1) * fb], bool(keys[i]), int(dmg[i])) for i in range(n)]

    def decode(self, packet):
        return self._decode_chunk([packet])[0]

    def stats(self):
        s = DecStats()
        _check(lib().ffv1b200_dec_stats(self._h, ctypes.byref(s)))
        return s

    def close(self):
        if getattr(self, "_h", None):
            lib().ffv1b200_dec_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
