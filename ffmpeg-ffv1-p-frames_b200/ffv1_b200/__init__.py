"""ffv1_b200 -- host-side Python mirror of the reference's FFV1 AVCodec interface, over the C ABI of
libffv1_b200.so (include/ffv1_b200.h).  The CUDA library is the only implementation: importing this module
without the built library, or calling it without a B200-class GPU, raises -- there is no CPU fallback.

    enc = FFV1Encoder(width, height, pix_fmt, g=16, level=3, coder=1, context=0, slices=24)   # = encode_init
    enc.extradata                                                                            # avctx->extradata
    pkts = enc.encode_batch(frames)            # list of (bytes, key_flag); frames: tightly packed uint8 arrays
    pkt  = enc.encode2(frame) / enc.flush()    # AVCodec.encode2 with AV_CODEC_CAP_DELAY semantics
    dec = FFV1Decoder(width, height, extradata); frames = dec.decode_batch([pkt, ...])

Option names and meaning follow the reference encoder (ffv1enc.c:1383-1399 and the generic -g/-level/-slices).
"""
from .codec import (FFV1Encoder, FFV1Decoder, FFV1Error, lib, library_path, device_count, frame_bytes, resolve_encoder,
                    FLAG_PASS1, FLAG_PASS2, FFV1Uploader, encode_cuda,
                    plane_shapes, EncStats, Packet, bind_thread_to_device)
from .partition import gop_aligned_ranges, reinterleave

__all__ = ["FFV1Encoder", "FFV1Decoder", "FFV1Error", "lib", "library_path", "device_count", "frame_bytes",
           "plane_shapes", "EncStats", "Packet", "bind_thread_to_device", "gop_aligned_ranges", "reinterleave"]
