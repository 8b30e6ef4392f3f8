/*
 * ffv1_b200.h -- C ABI of the B200-native FFV1 encode/decode hot path (libffv1_b200.so).
 *
 * This is the drop-in boundary: plain C, pointers and sizes only.  Each entry point replaces one
 * callback of the reference's AVCodec pair (file:line relative to the reference tree, an FFmpeg
 * 3.0.git checkout):
 *
 *   ffv1b200_enc_open            <- ff_ffv1_encoder.init    = encode_init   libavcodec/ffv1enc.c:669-1029
 *   ffv1b200_enc_extradata       <- avctx->extradata written by write_extradata       ffv1enc.c:545-619
 *   ffv1b200_enc_encode_* / _submit_host / _collect
 *                                <- ff_ffv1_encoder.encode2 = encode_frame  libavcodec/ffv1enc.c:1222-1373
 *                                   (batched: CAP_DELAY lets a codec hold frames, ffv1enc.c:1424)
 *   ffv1b200_enc_stats_out       <- avctx->stats_out at flush (first pass)      libavcodec/ffv1enc.c:1235-1277
 *   ffv1b200_enc_close           <- ff_ffv1_encoder.close   = encode_close  libavcodec/ffv1enc.c:1375-1379
 *   ffv1b200_dec_open            <- ff_ffv1_decoder.init    = decode_init   libavcodec/ffv1dec.c:876-893
 *   ffv1b200_dec_decode_*        <- ff_ffv1_decoder.decode  = decode_frame  libavcodec/ffv1dec.c:895-1035
 *   ffv1b200_dec_close           <- ff_ffv1_decoder.close   = ff_ffv1_close libavcodec/ffv1.c:205-243
 *
 * The AVCodec shim that binds these into libavcodec is integration/ffv1_b200_avcodec.c.
 * There is NO CPU fallback: every compute entry point fails with FFV1B200_ERR_EXTERNAL when no CUDA
 * device / sm_100 kernel image is usable.
 *
 * Return values follow libavcodec: 0 (or a non-negative count) on success, negative AVERROR-style code on
 * failure.  ffv1b200_strerror() maps them to text; ffv1b200_last_error() gives the detailed message of the
 * last failure on the calling thread (what the reference would have sent to av_log(AV_LOG_ERROR)).
 */
#ifndef FFV1_B200_H
#define FFV1_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFV1B200_ERR_EINVAL      (-22)           /* AVERROR(EINVAL)  */
#define FFV1B200_ERR_ENOMEM      (-12)           /* AVERROR(ENOMEM)  */
#define FFV1B200_ERR_ENOSYS      (-38)           /* AVERROR(ENOSYS)  */
#define FFV1B200_ERR_INVALIDDATA (-1094995529)   /* AVERROR_INVALIDDATA */
#define FFV1B200_ERR_EXTERNAL    (-542398533)    /* AVERROR_EXTERNAL: CUDA runtime / device failure */
#define FFV1B200_ERR_BUFFER_TOO_SMALL (-1397118274) /* AVERROR_BUFFER_TOO_SMALL: caller's output buffer */

#define FFV1B200_PKT_FLAG_KEY 1                  /* = AV_PKT_FLAG_KEY */

typedef struct FFV1B200Encoder FFV1B200Encoder;
typedef struct FFV1B200Decoder FFV1B200Decoder;

/* Encoder options: exactly what encode_init reads from AVCodecContext + the private AVOptions
 * (ffv1enc.c:1383-1399; SURVEY.md 8(b) "Inputs read from AVCodecContext"). */
typedef struct FFV1B200EncParams {
    int width, height;        /* AVCodecContext.width/height */
    const char *pix_fmt;      /* av_get_pix_fmt_name(avctx->pix_fmt), e.g. "yuv420p", "yuv422p10le", "gbrp14le", "bgr0" */
    int gop_size;             /* AVCodecContext.gop_size (-g); frame n is a keyframe iff gop_size==0 || n % gop_size == 0 */
    int level;                /* AVCodecContext.level: -1 unset, 0/1/3 = FFV1 version */
    int slices;               /* AVCodecContext.slices: 0 unset */
    int coder;                /* private option "coder":   0 rice, 1 ac (=2), 2 range_tab, -2 range_def */
    int context;              /* private option "context": 0 small (666 contexts), 1 large (7563 contexts) */
    int slicecrc;             /* private option "slicecrc": -1 auto, 0, 1 */
    int device;               /* CUDA device ordinal */
    int max_batch_frames;     /* capacity of one ffv1b200_enc_encode_* call (0 = default 64) */
    int64_t first_picture_number; /* picture_number of the first frame this instance will see (GOP-aligned multi-GPU
                                     shards continue the global count); normally 0 */
    int flags;                /* AVCodecContext.flags & (AV_CODEC_FLAG_PASS1 | AV_CODEC_FLAG_PASS2): two-pass coding
                                 (ffv1enc.c:680, 898-986, 1013-1027); 0 = single pass */
    const char *stats_in;     /* AVCodecContext.stats_in: the first pass's statistics text (ffv1b200_enc_stats_out), or NULL */
    int strict_std_compliance;/* AVCodecContext.strict_std_compliance; <= -2 (FF_COMPLIANCE_EXPERIMENTAL) unlocks level 4 */
    int bits_per_raw_sample;  /* AVCodecContext.bits_per_raw_sample: 0 = the pix_fmt's depth; for formats in 16-bit containers the
                                 depth that is coded instead (ffv1enc.c:728-748, 796-805), e.g. 12 for gbrp14le frames holding 12 bits */
} FFV1B200EncParams;

#define FFV1B200_FLAG_PASS1 (1 << 9)             /* = AV_CODEC_FLAG_PASS1 */
#define FFV1B200_FLAG_PASS2 (1 << 10)            /* = AV_CODEC_FLAG_PASS2 */

/* Per-frame side information that is coded into every slice header (ffv1enc.c:1044-1049). */
typedef struct FFV1B200FrameProps {
    int sar_num, sar_den;     /* AVFrame.sample_aspect_ratio; rawvideo default 0/1 */
    int picture_structure;    /* 3 progressive, 1 interlaced TFF, 2 interlaced BFF */
} FFV1B200FrameProps;

/* One encoded packet inside the caller's output buffer. */
typedef struct FFV1B200Packet {
    int64_t offset;           /* byte offset inside the output buffer */
    int32_t size;             /* pkt->size */
    int32_t flags;            /* FFV1B200_PKT_FLAG_KEY on keyframes (ffv1enc.c:1368) */
    int64_t picture_number;   /* coded picture number (pts order = coded order) */
} FFV1B200Packet;

/* Timing / counters of the last encode call (device times from CUDA events on the codec's stream). */
typedef struct FFV1B200EncStats {
    int   frames;
    int   kernel_launches;         /* launches of this library's kernels in the call */
    float ms_total;                /* first H2D (or first kernel) .. last D2H, device clock */
    float ms_pixel_kernel;         /* per-pixel pass (prediction/context/residual) */
    float ms_model_kernel;         /* adaptive-state replay  */
    float ms_coder_kernel;         /* range-coder interval arithmetic / golomb bit packing */
    float ms_pack_kernel;          /* packet assembly + CRC */
    int64_t h2d_bytes, d2h_bytes;
    int64_t samples;               /* coded samples */
    int64_t decisions;             /* binary range-coder decisions (0 in golomb mode) */
    int64_t packet_bytes;
    int   retries;                 /* internal re-runs after a scratch buffer had to grow */
} FFV1B200EncStats;

const char *ffv1b200_version(void);
const char *ffv1b200_strerror(int err);
const char *ffv1b200_last_error(void);
/* Number of usable CUDA devices with an sm_100 kernel image, or a negative error (never a CPU fallback). */
int ffv1b200_device_count(void);
/* Pins the calling thread to the CPUs of the NUMA node the device's PCIe link hangs off and prefers that node for the
 * thread's allocations (call it before allocating the pinned frame / packet buffers a device is fed from).  Returns the
 * node, or FFV1B200_ERR_ENOSYS when the topology is not visible (nothing is changed then).  Host-side plumbing of the
 * copy path the reference does not have (its frames never leave system memory). */
int ffv1b200_bind_thread_to_device(int device);
/* Pinned (page-locked) host memory for frames / packets that a device is fed from or writes to, placed on the device's
 * NUMA node where the topology is visible; NULL on failure.  Any host memory works with the *_host calls -- pinned memory
 * lets their copies run asynchronously at full link speed. */
void *ffv1b200_host_alloc(size_t bytes, int device);
void  ffv1b200_host_free(void *p);

/* ------------------------------------------------------------------ encoder */

int  ffv1b200_enc_open(FFV1B200Encoder **enc, const FFV1B200EncParams *params);
void ffv1b200_enc_close(FFV1B200Encoder *enc);

/* avctx->extradata: valid until close.  size 0 for FFV1 version 0/1 (in-band header). */
int  ffv1b200_enc_extradata(const FFV1B200Encoder *enc, const uint8_t **data, int *size);

/* Resolved configuration, for logging / INTEGRATION (what encode_init derived). */
typedef struct FFV1B200EncInfo {
    int version, micro_version, ac, colorspace, bits_per_raw_sample, chroma_planes, chroma_h_shift, chroma_v_shift,
        transparency, num_h_slices, num_v_slices, slice_count, ec, intra, context_count, plane_count, max_batch_frames;
    int64_t samples_per_frame;   /* coded samples (chroma overlap of unaligned slices included) */
    int64_t frame_bytes;         /* tightly packed input frame size */
} FFV1B200EncInfo;
int  ffv1b200_enc_info(const FFV1B200Encoder *enc, FFV1B200EncInfo *info);
/* The host half of encode_init alone (option resolution ffv1enc.c:676-1000, two-pass statistics 906-986, write_extradata
 * 545-619): needs no CUDA device.  Fills *info (may be NULL; samples_per_frame / max_batch_frames stay 0) and copies
 * the extradata the encoder would publish (extradata may be NULL to query *size).  Same errors as ffv1b200_enc_open. */
int  ffv1b200_enc_resolve(const FFV1B200EncParams *params, FFV1B200EncInfo *info, uint8_t *extradata, int cap, int *size);

void ffv1b200_enc_set_frame_props(FFV1B200Encoder *enc, const FFV1B200FrameProps *props);

/* Encode nframes (1..max_batch_frames) consecutive frames held in HOST memory.
 *   planes[f*4 + i], linesizes[f*4 + i] : AVFrame.data[i] / linesize[i] of frame f (unused planes NULL/0).
 *   out/out_cap    : host buffer receiving the packets back to back.
 *   pkts[nframes]  : filled with offset/size/flags per frame, in input order.
 * Blocking.  Returns nframes, or a negative error; FFV1B200_ERR_BUFFER_TOO_SMALL leaves the encoder state
 * unchanged (the call may be repeated with a larger buffer; *needed, if non-NULL, receives the size). */
int  ffv1b200_enc_encode_host(FFV1B200Encoder *enc, int nframes,
                              const uint8_t *const *planes, const int *linesizes,
                              uint8_t *out, size_t out_cap, FFV1B200Packet *pkts, size_t *needed);

/* Pipelined form of the same call (the AV_CODEC_CAP_DELAY way of using the codec, ffv1enc.c:1424): submit queues the
 * host->device copies and all kernels of a batch and returns; collect blocks until the OLDEST submitted batch is done and
 * copies its packets out.  Up to two batches may be in flight, so the copies of one batch overlap the kernels of the
 * other.  Frames passed to submit must stay valid until that batch has been collected.
 * collect returns the batch's frame count; FFV1B200_ERR_BUFFER_TOO_SMALL leaves the batch collectable (*needed = size). */
int  ffv1b200_enc_submit_host(FFV1B200Encoder *enc, int nframes, const uint8_t *const *planes, const int *linesizes);
int  ffv1b200_enc_collect(FFV1B200Encoder *enc, uint8_t *out, size_t out_cap, FFV1B200Packet *pkts, size_t *needed);
int  ffv1b200_enc_pending(const FFV1B200Encoder *enc);      /* batches submitted and not yet collected (0..2) */
/* collect without waiting for the packet BYTES: returns as soon as the batch is coded and its device->host copy has been
 * queued; pkts[] (offsets, sizes, flags) is valid on return, the bytes in `out` are complete once the next
 * ffv1b200_enc_collect* / ffv1b200_enc_sync_output / ffv1b200_enc_close call has returned.  With
 *     submit(k+1); collect_async(k); submit(k+2); collect_async(k+1); ...
 * the packets of batch k travel to the host while the frames of batch k+2 travel to the device (the link is full
 * duplex) and batch k+1 is being coded. */
int  ffv1b200_enc_collect_async(FFV1B200Encoder *enc, uint8_t *out, size_t out_cap, FFV1B200Packet *pkts, size_t *needed);
int  ffv1b200_enc_sync_output(FFV1B200Encoder *enc);

/* Same, for frames already resident in DEVICE memory (AV_PIX_FMT_CUDA AVFrames: data[i] are CUdeviceptr,
 * hwcontext_cuda.h:31-40).  Packets are produced in device memory (d_out, d_out_cap bytes); sizes/offsets are
 * returned in pkts (host).  If stream is non-NULL it is a cudaStream_t the work is ordered on. */
int  ffv1b200_enc_encode_device(FFV1B200Encoder *enc, int nframes,
                                const void *const *d_planes, const int *linesizes,
                                void *d_out, size_t d_out_cap, FFV1B200Packet *pkts, size_t *needed,
                                void *stream);

/* AV_PIX_FMT_CUDA frames in, packets out to HOST memory: what an AVCodec.encode2 needs when avctx->pix_fmt is
 * AV_PIX_FMT_CUDA (frame->data[i] = CUdeviceptr, sw_format from avctx->hw_frames_ctx; nvenc's use of the same frames:
 * libavcodec/nvenc.c:1162-1170).  Same packets as ffv1b200_enc_encode_host on the same pixels.
 * FFV1B200_ERR_BUFFER_TOO_SMALL (*needed = size): the batch IS coded; repeat the call with the same nframes and a larger
 * buffer to receive it (nothing is coded twice). */
int  ffv1b200_enc_encode_cuda(FFV1B200Encoder *enc, int nframes,
                              const void *const *d_planes, const int *linesizes,
                              uint8_t *out, size_t out_cap, FFV1B200Packet *pkts, size_t *needed);

int  ffv1b200_enc_stats(const FFV1B200Encoder *enc, FFV1B200EncStats *stats);

/* First pass of a two-pass encode (flags & FFV1B200_FLAG_PASS1): the text the reference leaves in
 * AVCodecContext.stats_out when the encoder is flushed (encode_frame with a NULL frame, ffv1enc.c:1235-1277), over all
 * frames coded so far.  Every batch must have been collected.  Writes a NUL-terminated string; returns its length, or
 * FFV1B200_ERR_BUFFER_TOO_SMALL with *needed = bytes required (terminator included).  The text is what a second encoder
 * takes as FFV1B200EncParams.stats_in. */
int  ffv1b200_enc_stats_out(FFV1B200Encoder *enc, char *buf, size_t cap, size_t *needed);

/* Test hook: run only the per-pixel pass on frames already uploaded by the last encode call and copy the
 * (context<<16 | diff&0xffff) records of one frame/slice (coding order) to host.  Returns the record count. */
int64_t ffv1b200_enc_debug_records(FFV1B200Encoder *enc, int frame_in_batch, int slice, uint32_t *dst, int64_t cap);

/* ------------------------------------------------------------------ on-GPU input preparation
 * The reference feeds a hardware encoder through its filter graph: vf_hwupload_cuda (libavfilter/vf_hwupload_cuda.c:144,
 * system-memory AVFrame -> AV_PIX_FMT_CUDA frame from a pool) and the pixel-format conversion of vf_scale_npp
 * (libavfilter/vf_scale_npp.c).  An uploader does both for this codec: frames in a capture / hardware-decoder layout
 *     nv12 -> yuv420p    p010le -> yuv420p10le    yuyv422, uyvy422 -> yuv422p    rgb24, bgr24 -> bgr0    rgba -> bgra
 * or already in one of the encoder's pix_fmts (upload only) become device frames the encoder takes
 * (ffv1b200_enc_encode_cuda / _encode_device).  Samples are re-arranged, never changed (P010's 10 bits move down). */
typedef struct FFV1B200Uploader FFV1B200Uploader;
/* dst_pix_fmt may be NULL (the format the source maps to); pool_frames = frames per call the pool holds */
int  ffv1b200_upload_open(FFV1B200Uploader **up, const char *src_pix_fmt, const char *dst_pix_fmt, int width, int height,
                          int pool_frames, int device);
void ffv1b200_upload_close(FFV1B200Uploader *up);
const char *ffv1b200_upload_pix_fmt(const FFV1B200Uploader *up);     /* pix_fmt of the device frames */
/* nframes (<= pool_frames) host frames -> device frames: d_planes[f*4+i] / d_linesizes[i] as the encoder's device entry
 * points take them; valid until the next call on this uploader.  Blocking.  Returns nframes or a negative error. */
int  ffv1b200_upload_frames(FFV1B200Uploader *up, int nframes, const uint8_t *const *planes, const int *linesizes,
                            void **d_planes, int *d_linesizes);
/* the conversion alone, for frames already in device memory (a hardware decoder's NV12 / P010 surfaces): d_src_planes[i],
 * src_linesizes[i] (multiples of 4) of frame 0, consecutive frames src_frame_stride bytes apart.  Ordered on `stream`
 * (a cudaStream_t) when that is non-NULL, else blocking. */
int  ffv1b200_convert_device(FFV1B200Uploader *up, int nframes, const void *const *d_src_planes, const int *src_linesizes,
                             long long src_frame_stride, void **d_planes, int *d_linesizes, void *stream);

/* ------------------------------------------------------------------ decoder */

typedef struct FFV1B200DecParams {
    int width, height;              /* AVCodecContext.width/height (container) */
    const uint8_t *extradata;       /* AVCodecContext.extradata (version >= 2 streams) */
    int extradata_size;
    int device;
    int max_batch_frames;           /* 0 = default 64 */
} FFV1B200DecParams;

typedef struct FFV1B200DecInfo {
    int version, micro_version, ac, colorspace, bits_per_raw_sample, chroma_planes, chroma_h_shift, chroma_v_shift,
        transparency, num_h_slices, num_v_slices, ec, intra;
    char pix_fmt[32];               /* the AVPixelFormat name read_header would select (ffv1dec.c:698-786) */
    int64_t frame_bytes;            /* tightly packed output frame size */
} FFV1B200DecInfo;

int  ffv1b200_dec_open(FFV1B200Decoder **dec, const FFV1B200DecParams *params);
void ffv1b200_dec_close(FFV1B200Decoder *dec);
int  ffv1b200_dec_info(const FFV1B200Decoder *dec, FFV1B200DecInfo *info);

/* Decode npackets consecutive packets (host memory) into tightly packed frames (host memory):
 *   pkt_data[i]/pkt_size[i]  : AVPacket.data/size
 *   out + i*frame_bytes      : decoded frame i (planes back to back, as av_image_copy_to_buffer(align=1))
 *   key_flags[i]             : AVFrame.key_frame;  damaged[i] : bit s set = slice s failed CRC/end check
 *                              (and was concealed from the previous frame, ffv1dec.c:998-1021)
 * The first packet of a shard must be a keyframe (ffv1dec.c:930-935) unless state carried over from the previous call.
 * `out` may be pageable memory (the pictures are staged on the device and copied behind the batch) or pinned memory that the
 * device can address (ffv1b200_host_alloc, cudaHostAlloc): the decode kernel then writes the pictures there itself. */
int  ffv1b200_dec_decode_host(FFV1B200Decoder *dec, int npackets,
                              const uint8_t *const *pkt_data, const int *pkt_size,
                              uint8_t *out, size_t out_cap, int *key_flags, uint64_t *damaged);

typedef struct FFV1B200DecStats {
    int frames, kernel_launches;
    float ms_total, ms_decode_kernel;
    int64_t h2d_bytes, d2h_bytes;
} FFV1B200DecStats;
int  ffv1b200_dec_stats(const FFV1B200Decoder *dec, FFV1B200DecStats *stats);

#ifdef __cplusplus
}
#endif
#endif /* FFV1_B200_H */
