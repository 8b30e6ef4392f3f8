/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin C wrapper that drives the UNMODIFIED reference FFV1 encoder/decoder
 * (/root/reference/libavcodec/ffv1enc.c, ffv1dec.c, compiled as they lie) through
 * the reference's public libavcodec API: avcodec_register / avcodec_open2 /
 * avcodec_encode_video2 / avcodec_decode_video2 (libavcodec/utils.c:178,1208,1922,2180).
 * It exists so that Python tests (ctypes) and bench.py's CPU-baseline leg can obtain the
 * reference's packets, extradata and decoded frames, and time the reference's
 * slice-threaded CPU path.  Built into oracle/_ref/libffv1ref.so by oracle/Makefile.
 */
#include <string.h>
#include <stdlib.h>
#include "libavcodec/avcodec.h"
#include "libavutil/opt.h"
#include "libavutil/pixdesc.h"
#include "libavutil/imgutils.h"

extern AVCodec ff_ffv1_encoder;
extern AVCodec ff_ffv1_decoder;

static int g_registered;
static void reg(void)
{
    if (!g_registered) {
        avcodec_register(&ff_ffv1_encoder);
        avcodec_register(&ff_ffv1_decoder);
        g_registered = 1;
        av_log_set_level(AV_LOG_ERROR);
    }
}

typedef struct RefEnc {
    AVCodecContext *ctx;
    AVFrame *frame;
    int64_t pts;
} RefEnc;

/* level<0 / slices==0 / slicecrc<0 mean "leave at the reference default". threads>1 enables the
 * reference's slice threading (pthread_slice.c). sar_num/den go to every AVFrame (they are coded in
 * each slice header, ffv1enc.c:1048-1049). */
static void *enc_open_impl(AVCodec *codec, int w, int h, const char *pix_fmt, int gop, int level, int coder, int context,
                           int slices, int slicecrc, int threads, int strict_experimental, int batch);
static int g_pass;
static const char *g_stats_in;
static int g_bits;                       /* AVCodecContext.bits_per_raw_sample of the next encoder that is opened */
void ffv1ref_set_next_bits_per_raw_sample(int bits) { g_bits = bits; }


void *ffv1ref_enc_open(int w, int h, const char *pix_fmt, int gop, int level, int coder, int context,
                       int slices, int slicecrc, int threads, int strict_experimental)
{
    reg();
    return enc_open_impl(&ff_ffv1_encoder, w, h, pix_fmt, gop, level, coder, context, slices, slicecrc, threads, strict_experimental, 0);
}

/* Two-pass coding: pass = 1 sets AV_CODEC_FLAG_PASS1 (statistics into stats_out), pass = 2 sets AV_CODEC_FLAG_PASS2;
 * stats_in (may be NULL) is handed to the encoder as AVCodecContext.stats_in (ffv1enc.c:906-986). */
void *ffv1ref_enc_open_2pass(int w, int h, const char *pix_fmt, int gop, int level, int coder, int context,
                             int slices, int slicecrc, int threads, int strict_experimental, int pass, const char *stats_in)
{
    void *r;
    reg();
    g_pass = pass; g_stats_in = stats_in;
    r = enc_open_impl(&ff_ffv1_encoder, w, h, pix_fmt, gop, level, coder, context, slices, slicecrc, threads, strict_experimental, 0);
    g_pass = 0; g_stats_in = NULL;
    return r;
}

/* flushes the encoder (frame = NULL: encode_frame then writes the statistics text, ffv1enc.c:1235-1277) and copies
 * AVCodecContext.stats_out; returns its length or <0 */
int ffv1ref_enc_stats_out(void *h, char *dst, int cap)
{
    RefEnc *e = h;
    AVPacket pkt;
    int got = 0, n;
    av_init_packet(&pkt); pkt.data = NULL; pkt.size = 0;
    if (avcodec_encode_video2(e->ctx, &pkt, NULL, &got) < 0) return -1;
    if (!e->ctx->stats_out) return -3;
    n = (int)strlen(e->ctx->stats_out);
    if (n + 1 > cap) return -2;
    memcpy(dst, e->ctx->stats_out, n + 1);
    return n;
}

/* Drop-in test support: register an out-of-tree AVCodec (the ffv1_b200 shim) with THIS libavcodec and open it by
 * name through the same public API path (avcodec_find_encoder_by_name, utils.c:3051). */
void ffv1ref_register_codec(AVCodec *codec) { reg(); avcodec_register(codec); }

void *ffv1ref_enc_open_named(const char *name, int w, int h, const char *pix_fmt, int gop, int level, int coder, int context,
                             int slices, int slicecrc, int batch)
{
    reg();
    AVCodec *codec = avcodec_find_encoder_by_name(name);
    if (!codec) return NULL;
    return enc_open_impl(codec, w, h, pix_fmt, gop, level, coder, context, slices, slicecrc, 1, 0, batch);
}

/* the same, with -strict experimental and the two-pass options (pass: 0 none, 1 = AV_CODEC_FLAG_PASS1, 2 = _PASS2) */
void *ffv1ref_enc_open_named_ex(const char *name, int w, int h, const char *pix_fmt, int gop, int level, int coder, int context,
                                int slices, int slicecrc, int batch, int strict_experimental, int pass, const char *stats_in)
{
    void *r;
    reg();
    AVCodec *codec = avcodec_find_encoder_by_name(name);
    if (!codec) return NULL;
    g_pass = pass; g_stats_in = stats_in;
    r = enc_open_impl(codec, w, h, pix_fmt, gop, level, coder, context, slices, slicecrc, 1, strict_experimental, batch);
    g_pass = 0; g_stats_in = NULL;
    return r;
}

static void *enc_open_impl(AVCodec *codec, int w, int h, const char *pix_fmt, int gop, int level, int coder, int context,
                           int slices, int slicecrc, int threads, int strict_experimental, int batch)
{
    enum AVPixelFormat pf = av_get_pix_fmt(pix_fmt);
    if (pf == AV_PIX_FMT_NONE) return NULL;
    RefEnc *e = calloc(1, sizeof(*e));
    e->ctx = avcodec_alloc_context3(codec);
    e->ctx->width = w; e->ctx->height = h; e->ctx->pix_fmt = pf;
    e->ctx->time_base = (AVRational){1, 25};
    e->ctx->gop_size = gop;
    e->ctx->level = level;
    e->ctx->slices = slices;
    e->ctx->flags |= AV_CODEC_FLAG_BITEXACT;
    if (strict_experimental) e->ctx->strict_std_compliance = FF_COMPLIANCE_EXPERIMENTAL;
    if (g_bits) { e->ctx->bits_per_raw_sample = g_bits; g_bits = 0; }
    if (g_pass == 1) e->ctx->flags |= AV_CODEC_FLAG_PASS1;
    if (g_pass == 2) e->ctx->flags |= AV_CODEC_FLAG_PASS2;
    if (g_stats_in) e->ctx->stats_in = av_strdup(g_stats_in);
    if (threads > 1) { e->ctx->thread_count = threads; e->ctx->thread_type = FF_THREAD_SLICE; }
    else e->ctx->thread_count = 1;
    av_opt_set_int(e->ctx->priv_data, "coder", coder, 0);
    av_opt_set_int(e->ctx->priv_data, "context", context, 0);
    av_opt_set_int(e->ctx->priv_data, "slicecrc", slicecrc, 0);
    if (batch > 0) av_opt_set_int(e->ctx->priv_data, "batch", batch, 0);
    if (avcodec_open2(e->ctx, codec, NULL) < 0) {
        avcodec_free_context(&e->ctx); free(e); return NULL;
    }
    e->frame = av_frame_alloc();
    return e;
}

int ffv1ref_enc_extradata(void *h, uint8_t *dst, int cap)
{
    RefEnc *e = h;
    if (e->ctx->extradata_size > cap) return -1;
    memcpy(dst, e->ctx->extradata, e->ctx->extradata_size);
    return e->ctx->extradata_size;
}

/* Encode one frame given plane pointers + linesizes; returns packet size (copied to dst), <0 on error.
 * *key receives AV_PKT_FLAG_KEY. */
int ffv1ref_enc_frame(void *h, uint8_t *const planes[4], const int linesize[4], int sar_num, int sar_den,
                      int interlaced, int tff, uint8_t *dst, int cap, int *key)
{
    RefEnc *e = h;
    AVPacket pkt;
    int got = 0, ret, i;
    av_init_packet(&pkt); pkt.data = NULL; pkt.size = 0;
    if (!planes) {                                  /* drain a delayed encoder (AV_CODEC_CAP_DELAY): frame = NULL */
        ret = avcodec_encode_video2(e->ctx, &pkt, NULL, &got);
        if (ret < 0) return ret;
        if (!got) return 0;
        if (pkt.size > cap) { av_packet_unref(&pkt); return -2; }
        memcpy(dst, pkt.data, pkt.size);
        ret = pkt.size;
        if (key) *key = !!(pkt.flags & AV_PKT_FLAG_KEY);
        av_packet_unref(&pkt);
        return ret;
    }
    /* hand the encoder a refcounted frame (as ffmpeg.c's rawvideo path does): copy the caller's rows in */
    av_frame_unref(e->frame);
    e->frame->format = e->ctx->pix_fmt; e->frame->width = e->ctx->width; e->frame->height = e->ctx->height;
    if ((ret = av_frame_get_buffer(e->frame, 32)) < 0) return ret;
    for (i = 0; i < 4 && planes[i]; i++) {
        int bw = av_image_get_linesize(e->ctx->pix_fmt, e->ctx->width, i);
        int hh = e->ctx->height;
        const AVPixFmtDescriptor *d = av_pix_fmt_desc_get(e->ctx->pix_fmt);
        if (i == 1 || i == 2) hh = -((-hh) >> d->log2_chroma_h);
        av_image_copy_plane(e->frame->data[i], e->frame->linesize[i], planes[i], linesize[i], bw, hh);
    }
    e->frame->pts = e->pts++;
    e->frame->sample_aspect_ratio = (AVRational){sar_num, sar_den};
    e->frame->interlaced_frame = interlaced;
    e->frame->top_field_first = tff;
    ret = avcodec_encode_video2(e->ctx, &pkt, e->frame, &got);
    if (ret < 0) return ret;
    if (!got) return 0;
    if (pkt.size > cap) { av_packet_unref(&pkt); return -2; }
    memcpy(dst, pkt.data, pkt.size);
    ret = pkt.size;
    if (key) *key = !!(pkt.flags & AV_PKT_FLAG_KEY);
    av_packet_unref(&pkt);
    return ret;
}

/* ---- AV_PIX_FMT_CUDA input (drop-in test of the shim's device-frame path).  The reference tree is built without
 * CONFIG_CUDA here, so the frames context is filled in by hand: a codec only reads sw_format / width / height from it
 * (nvenc.c:412-421), and the AVFrames carry caller-owned device pointers behind a dummy refcounted buffer. */
#include "libavutil/hwcontext.h"
void *ffv1ref_enc_open_named_cuda(const char *name, int w, int h, const char *sw_pix_fmt, int gop, int level, int coder, int context,
                                  int slices, int slicecrc, int batch)
{
    reg();
    AVCodec *codec = avcodec_find_encoder_by_name(name);
    enum AVPixelFormat sw = av_get_pix_fmt(sw_pix_fmt);
    if (!codec || sw == AV_PIX_FMT_NONE) return NULL;
    RefEnc *e = calloc(1, sizeof(*e));
    AVHWFramesContext *fc = av_mallocz(sizeof(*fc));
    e->ctx = avcodec_alloc_context3(codec);
    e->ctx->width = w; e->ctx->height = h; e->ctx->pix_fmt = AV_PIX_FMT_CUDA;
    fc->format = AV_PIX_FMT_CUDA; fc->sw_format = sw; fc->width = w; fc->height = h;
    e->ctx->hw_frames_ctx = av_buffer_create((uint8_t *)fc, sizeof(*fc), NULL, NULL, 0);
    e->ctx->time_base = (AVRational){1, 25};
    e->ctx->gop_size = gop; e->ctx->level = level; e->ctx->slices = slices;
    e->ctx->flags |= AV_CODEC_FLAG_BITEXACT;
    e->ctx->thread_count = 1;
    av_opt_set_int(e->ctx->priv_data, "coder", coder, 0);
    av_opt_set_int(e->ctx->priv_data, "context", context, 0);
    av_opt_set_int(e->ctx->priv_data, "slicecrc", slicecrc, 0);
    if (batch > 0) av_opt_set_int(e->ctx->priv_data, "batch", batch, 0);
    if (avcodec_open2(e->ctx, codec, NULL) < 0) { avcodec_free_context(&e->ctx); free(e); return NULL; }
    e->frame = av_frame_alloc();
    return e;
}

/* planes[] are DEVICE pointers that stay valid until the packet of this frame has been returned */
int ffv1ref_enc_frame_cuda(void *h, uint8_t *const planes[4], const int linesize[4], uint8_t *dst, int cap, int *key)
{
    RefEnc *e = h;
    AVPacket pkt;
    int got = 0, ret, i;
    av_init_packet(&pkt); pkt.data = NULL; pkt.size = 0;
    av_frame_unref(e->frame);
    e->frame->format = AV_PIX_FMT_CUDA; e->frame->width = e->ctx->width; e->frame->height = e->ctx->height;
    e->frame->buf[0] = av_buffer_alloc(1);                 /* refcounted: av_frame_clone() must not copy pixels */
    for (i = 0; i < 4; i++) { e->frame->data[i] = planes[i]; e->frame->linesize[i] = planes[i] ? linesize[i] : 0; }
    e->frame->pts = e->pts++;
    e->frame->sample_aspect_ratio = (AVRational){0, 1};
    ret = avcodec_encode_video2(e->ctx, &pkt, e->frame, &got);
    if (ret < 0) return ret;
    if (!got) return 0;
    if (pkt.size > cap) { av_packet_unref(&pkt); return -2; }
    memcpy(dst, pkt.data, pkt.size);
    ret = pkt.size;
    if (key) *key = !!(pkt.flags & AV_PKT_FLAG_KEY);
    av_packet_unref(&pkt);
    return ret;
}

void ffv1ref_enc_close(void *h)
{
    RefEnc *e = h;
    if (!e) return;
    av_frame_free(&e->frame);
    avcodec_close(e->ctx);
    avcodec_free_context(&e->ctx);
    free(e);
}

typedef struct RefDec {
    AVCodecContext *ctx;
    AVFrame *frame;
} RefDec;

static void *dec_open_impl(AVCodec *codec, int w, int h, const uint8_t *extradata, int extradata_size, int threads, int frame_threads);

void *ffv1ref_dec_open(int w, int h, const uint8_t *extradata, int extradata_size, int threads, int frame_threads)
{
    reg();
    return dec_open_impl(&ff_ffv1_decoder, w, h, extradata, extradata_size, threads, frame_threads);
}

void *ffv1ref_dec_open_named(const char *name, int w, int h, const uint8_t *extradata, int extradata_size)
{
    reg();
    AVCodec *codec = avcodec_find_decoder_by_name(name);
    if (!codec) return NULL;
    return dec_open_impl(codec, w, h, extradata, extradata_size, 1, 0);
}

/* private options of the named decoder ("batch=8:gpu=0") applied before avcodec_open2 */
static const char *g_dec_extra_opts;
void *ffv1ref_dec_open_named_opts(const char *name, int w, int h, const uint8_t *extradata, int extradata_size, const char *extra_opts)
{
    void *r;
    g_dec_extra_opts = extra_opts;
    r = ffv1ref_dec_open_named(name, w, h, extradata, extradata_size);
    g_dec_extra_opts = NULL;
    return r;
}

static void *dec_open_impl(AVCodec *codec, int w, int h, const uint8_t *extradata, int extradata_size, int threads, int frame_threads)
{
    RefDec *d = calloc(1, sizeof(*d));
    d->ctx = avcodec_alloc_context3(codec);
    d->ctx->width = w; d->ctx->height = h;
    if (extradata_size > 0) {
        d->ctx->extradata = av_mallocz(extradata_size + AV_INPUT_BUFFER_PADDING_SIZE);
        memcpy(d->ctx->extradata, extradata, extradata_size);
        d->ctx->extradata_size = extradata_size;
    }
    d->ctx->flags |= AV_CODEC_FLAG_BITEXACT;
    d->ctx->thread_count = threads > 1 ? threads : 1;
    if (threads > 1) d->ctx->thread_type = frame_threads ? FF_THREAD_FRAME : FF_THREAD_SLICE;
    if ((g_dec_extra_opts && *g_dec_extra_opts && av_set_options_string(d->ctx->priv_data, g_dec_extra_opts, "=", ":") < 0) ||
        avcodec_open2(d->ctx, codec, NULL) < 0) {
        avcodec_free_context(&d->ctx); free(d); return NULL;
    }
    d->frame = av_frame_alloc();
    return d;
}

/* Decode one packet. On success copies the planes tightly packed (plane after plane, width*bytes per row)
 * into dst, writes the pix_fmt name to fmt_name (cap 32) and returns total bytes; 0 = no frame; <0 error. */
int ffv1ref_dec_packet(void *h, const uint8_t *data, int size, uint8_t *dst, int cap, char *fmt_name, int *key)
{
    RefDec *d = h;
    AVPacket pkt;
    int got = 0, ret;
    uint8_t *buf = av_mallocz(size + AV_INPUT_BUFFER_PADDING_SIZE);
    memcpy(buf, data, size);
    av_init_packet(&pkt); pkt.data = buf; pkt.size = size;
    ret = avcodec_decode_video2(d->ctx, d->frame, &got, &pkt);
    av_free(buf);
    if (ret < 0) return ret;
    if (!got) return 0;
    ret = av_image_get_buffer_size(d->frame->format, d->frame->width, d->frame->height, 1);
    if (ret > cap) return -2;
    av_image_copy_to_buffer(dst, cap, (const uint8_t *const *)d->frame->data, d->frame->linesize,
                            d->frame->format, d->frame->width, d->frame->height, 1);
    if (fmt_name) { strncpy(fmt_name, av_get_pix_fmt_name(d->frame->format), 31); fmt_name[31] = 0; }
    if (key) *key = d->frame->key_frame;
    av_frame_unref(d->frame);
    return ret;
}

void ffv1ref_dec_close(void *h)
{
    RefDec *d = h;
    if (!d) return;
    av_frame_free(&d->frame);
    avcodec_close(d->ctx);
    avcodec_free_context(&d->ctx);
    free(d);
}

/* ---- throughput of an encoder through the public API (bench.py "e2e_avcodec"): `nframes` frames, taken round-robin from
 * `clip` (nclip tightly packed frames in pageable memory, wrapped as refcounted AVFrames WITHOUT copying, the way ffmpeg.c
 * hands rawvideo frames on), go through avcodec_encode_video2 of the named codec; the packets are dropped except the last
 * `nkeep`, which are copied to `keep` (sizes in keep_size, key flags in keep_key) for the caller's parity check.  Returns
 * the seconds from the first frame to the last drained packet (open/close excluded), < 0 on error. */
#include <time.h>
static void noop_free(void *opaque, uint8_t *data) { (void)opaque; (void)data; }
double ffv1ref_bench_encode(const char *name, int w, int h, const char *pix_fmt, int gop, int level, int coder, int context,
                            int slices, int slicecrc, int threads, const char *extra_opts,
                            const uint8_t *clip, int nclip, int nframes,
                            uint8_t *keep, int64_t keep_cap, int nkeep, int *keep_size, int *keep_key, int64_t *total_bytes)
{
    reg();
    AVCodec *codec = avcodec_find_encoder_by_name(name);
    if (!codec) return -1;
    RefEnc *e = enc_open_impl(codec, w, h, pix_fmt, gop, level, coder, context, slices, slicecrc, threads, 0, 0);
    /* (private options of the codec under test, "key=value:key=value", are applied before avcodec_open2 by reopening) */
    if (!e) return -2;
    if (extra_opts && *extra_opts) {
        avcodec_close(e->ctx); avcodec_free_context(&e->ctx); av_frame_free(&e->frame); free(e);
        enum AVPixelFormat pf = av_get_pix_fmt(pix_fmt);
        e = calloc(1, sizeof(*e));
        e->ctx = avcodec_alloc_context3(codec);
        e->ctx->width = w; e->ctx->height = h; e->ctx->pix_fmt = pf; e->ctx->time_base = (AVRational){1, 25};
        e->ctx->gop_size = gop; e->ctx->level = level; e->ctx->slices = slices; e->ctx->flags |= AV_CODEC_FLAG_BITEXACT;
        if (threads > 1) { e->ctx->thread_count = threads; e->ctx->thread_type = FF_THREAD_SLICE; } else e->ctx->thread_count = 1;
        av_opt_set_int(e->ctx->priv_data, "coder", coder, 0);
        av_opt_set_int(e->ctx->priv_data, "context", context, 0);
        av_opt_set_int(e->ctx->priv_data, "slicecrc", slicecrc, 0);
        if (av_set_options_string(e->ctx->priv_data, extra_opts, "=", ":") < 0 || avcodec_open2(e->ctx, codec, NULL) < 0) {
            avcodec_free_context(&e->ctx); free(e); return -3;
        }
        e->frame = av_frame_alloc();
    }
    const enum AVPixelFormat pf = e->ctx->pix_fmt;
    const int fb = av_image_get_buffer_size(pf, w, h, 1);
    int64_t bytes = 0, kept = 0;
    int npk = 0, i, ret = 0, got;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (i = 0; ; i++) {
        AVPacket pkt;
        AVFrame *f = NULL;
        av_init_packet(&pkt); pkt.data = NULL; pkt.size = 0;
        if (i < nframes) {
            f = e->frame;
            av_frame_unref(f);
            f->format = pf; f->width = w; f->height = h;
            av_image_fill_arrays(f->data, f->linesize, clip + (size_t)(i % nclip) * fb, pf, w, h, 1);
            f->buf[0] = av_buffer_create((uint8_t *)clip + (size_t)(i % nclip) * fb, fb, noop_free, NULL, AV_BUFFER_FLAG_READONLY);
            f->pts = i;
            f->sample_aspect_ratio = (AVRational){0, 1};
        }
        ret = avcodec_encode_video2(e->ctx, &pkt, f, &got);
        if (ret < 0) break;
        if (got) {
            if (npk >= nframes - nkeep && kept + pkt.size <= keep_cap) {
                const int k = npk - (nframes - nkeep);
                memcpy(keep + kept, pkt.data, pkt.size);
                keep_size[k] = pkt.size; keep_key[k] = !!(pkt.flags & AV_PKT_FLAG_KEY);
                kept += pkt.size;
            }
            bytes += pkt.size; npk++;
            av_packet_unref(&pkt);
        } else if (i >= nframes) break;                      /* drained */
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (total_bytes) *total_bytes = bytes;
    ffv1ref_enc_close(e);
    if (ret < 0) return ret;
    if (npk != nframes) return -4;
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* same for a decoder: npk packets (pkt[i], size[i]) through avcodec_decode_video2, pictures dropped except the last, which
 * is copied tightly packed to `last` (cap bytes).  Returns seconds, < 0 on error; *nframes_out = pictures received. */
double ffv1ref_bench_decode(const char *name, int w, int h, const uint8_t *extradata, int extradata_size, int threads,
                            int frame_threads, const char *extra_opts, const uint8_t *const *pkt, const int *size, int npk,
                            uint8_t *last, int cap, int *nframes_out)
{
    reg();
    AVCodec *codec = avcodec_find_decoder_by_name(name);
    if (!codec) return -1;
    RefDec *d = calloc(1, sizeof(*d));
    d->ctx = avcodec_alloc_context3(codec);
    d->ctx->width = w; d->ctx->height = h;
    if (extradata_size > 0) {
        d->ctx->extradata = av_mallocz(extradata_size + AV_INPUT_BUFFER_PADDING_SIZE);
        memcpy(d->ctx->extradata, extradata, extradata_size);
        d->ctx->extradata_size = extradata_size;
    }
    d->ctx->flags |= AV_CODEC_FLAG_BITEXACT;
    d->ctx->thread_count = threads > 1 ? threads : 1;
    if (threads > 1) d->ctx->thread_type = frame_threads ? FF_THREAD_FRAME : FF_THREAD_SLICE;
    if ((extra_opts && *extra_opts && av_set_options_string(d->ctx->priv_data, extra_opts, "=", ":") < 0) ||
        avcodec_open2(d->ctx, codec, NULL) < 0) { avcodec_free_context(&d->ctx); free(d); return -3; }
    d->frame = av_frame_alloc();
    /* packets with the padding lavc asks for, prepared outside the timed region */
    int maxsz = 0, i, ret = 0, got, nout = 0;
    for (i = 0; i < npk; i++) if (size[i] > maxsz) maxsz = size[i];
    uint8_t **padded = calloc(npk, sizeof(*padded));
    for (i = 0; i < npk; i++) { padded[i] = av_mallocz(size[i] + AV_INPUT_BUFFER_PADDING_SIZE); memcpy(padded[i], pkt[i], size[i]); }
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (i = 0; ; i++) {
        AVPacket p;
        av_init_packet(&p);
        p.data = i < npk ? padded[i] : NULL; p.size = i < npk ? size[i] : 0; p.pts = p.dts = i;
        ret = avcodec_decode_video2(d->ctx, d->frame, &got, &p);
        if (ret < 0) break;
        if (got) {
            nout++;
            if (nout == npk && last)
                av_image_copy_to_buffer(last, cap, (const uint8_t *const *)d->frame->data, d->frame->linesize,
                                        d->frame->format, d->frame->width, d->frame->height, 1);
            av_frame_unref(d->frame);
        } else if (i >= npk) break;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    for (i = 0; i < npk; i++) av_free(padded[i]);
    free(padded);
    if (nframes_out) *nframes_out = nout;
    ffv1ref_dec_close(d);
    if (ret < 0) return ret;
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* ---- FATE's input conversion (tests/fate/vcodec.mak:119-127: -pix_fmt yuv422p10 / yuv444p16 / bgr0 with
 * -sws_flags neighbor+bitexact, to which fate-run.sh:167 appends +accurate_rnd+bitexact): what ffmpeg's auto-inserted scale
 * filter does to the yuv420p clip (libavfilter/vf_scale.c:373-423: one SwsContext from the option API, MPEG-2 chroma
 * position 128 for a yuv420p side), frame by frame.  src/dst hold tightly packed frames.  Returns bytes per dst frame. */
#include "libswscale/swscale.h"
int ffv1ref_sws_convert(const uint8_t *src, const char *src_fmt, uint8_t *dst, const char *dst_fmt, int w, int h, int nframes)
{
    enum AVPixelFormat sf = av_get_pix_fmt(src_fmt), df = av_get_pix_fmt(dst_fmt);
    struct SwsContext *s;
    int ssize, dsize, i;
    if (sf == AV_PIX_FMT_NONE || df == AV_PIX_FMT_NONE) return -1;
    ssize = av_image_get_buffer_size(sf, w, h, 1);
    dsize = av_image_get_buffer_size(df, w, h, 1);
    if (!dst) return dsize;
    s = sws_alloc_context();
    av_opt_set_int(s, "srcw", w, 0); av_opt_set_int(s, "srch", h, 0); av_opt_set_int(s, "src_format", sf, 0);
    av_opt_set_int(s, "dstw", w, 0); av_opt_set_int(s, "dsth", h, 0); av_opt_set_int(s, "dst_format", df, 0);
    av_opt_set_int(s, "sws_flags", SWS_POINT | SWS_BITEXACT | SWS_ACCURATE_RND, 0);
    av_opt_set_int(s, "param0", SWS_PARAM_DEFAULT, 0); av_opt_set_int(s, "param1", SWS_PARAM_DEFAULT, 0);
    av_opt_set_int(s, "src_h_chr_pos", -513, 0); av_opt_set_int(s, "src_v_chr_pos", sf == AV_PIX_FMT_YUV420P ? 128 : -513, 0);
    av_opt_set_int(s, "dst_h_chr_pos", -513, 0); av_opt_set_int(s, "dst_v_chr_pos", df == AV_PIX_FMT_YUV420P ? 128 : -513, 0);
    if (sws_init_context(s, NULL, NULL) < 0) { sws_freeContext(s); return -2; }
    for (i = 0; i < nframes; i++) {
        uint8_t *sp[4], *dp[4]; int sl[4], dl[4];
        av_image_fill_arrays(sp, sl, src + (int64_t)i * ssize, sf, w, h, 1);
        av_image_fill_arrays(dp, dl, dst + (int64_t)i * dsize, df, w, h, 1);
        sws_scale(s, (const uint8_t *const *)sp, sl, 0, h, dp, dl);
    }
    sws_freeContext(s);
    return dsize;
}

/* CRC helper so tests can pin libavutil's AV_CRC_32_IEEE convention (crc.c:357-380). */
#include "libavutil/crc.h"
unsigned ffv1ref_crc32_ieee(unsigned init, const uint8_t *buf, int len)
{
    return av_crc(av_crc_get_table(AV_CRC_32_IEEE), init, buf, len);
}

/* ------------------------------------------------------------------------------------------------
 * FATE reproduction: encode a raw clip with the reference encoder and mux it with the reference's
 * own AVI muxer (libavformat/avienc.c) into memory, the way tests/fate-run.sh:171-193 (enc_dec)
 * drives ffmpeg.c, so that the MD5 + size in tests/ref/vsynth/vsynth*-ffv1* can be checked.
 * ---------------------------------------------------------------------------------------------- */
#include "libavformat/avformat.h"
#include "libavformat/avio.h"
#include "libavformat/internal.h"

extern AVOutputFormat ff_avi_muxer;
extern AVOutputFormat ff_nut_muxer;
extern AVInputFormat ff_nut_demuxer;
extern AVOutputFormat ff_matroska_muxer;
static AVOutputFormat *g_muxer = &ff_avi_muxer;          /* container fate_avi_impl writes */

typedef struct MemOut { uint8_t *buf; int64_t cap, pos, size; } MemOut;
static int mem_write(void *opaque, uint8_t *buf, int n)
{
    MemOut *m = opaque;
    if (m->pos + n > m->cap) return -1;
    memcpy(m->buf + m->pos, buf, n);
    m->pos += n;
    if (m->pos > m->size) m->size = m->pos;
    return n;
}
static int64_t mem_seek(void *opaque, int64_t off, int whence)
{
    MemOut *m = opaque;
    if (whence == AVSEEK_SIZE) return m->size;
    if (whence == SEEK_CUR) off += m->pos;
    else if (whence == SEEK_END) off += m->size;
    if (off < 0 || off > m->cap) return -1;
    m->pos = off;
    return off;
}

/* raw: nframes tightly packed frames in pix_fmt. Returns AVI size written into out, <0 on error. */
static int64_t fate_avi_impl(AVCodec *codec, int batch, const uint8_t *raw, int nframes, int w, int h, const char *pix_fmt,
                             int level, int slices, uint8_t *out, int64_t cap);

int64_t ffv1ref_fate_avi(const uint8_t *raw, int nframes, int w, int h, const char *pix_fmt,
                         int level, int slices, uint8_t *out, int64_t cap)
{
    reg();
    return fate_avi_impl(&ff_ffv1_encoder, 0, raw, nframes, w, h, pix_fmt, level, slices, out, cap);
}

/* the same FATE procedure with another registered encoder (the ffv1_b200 drop-in) */
int64_t ffv1ref_fate_avi_named(const char *name, int batch, const uint8_t *raw, int nframes, int w, int h, const char *pix_fmt,
                               int level, int slices, uint8_t *out, int64_t cap)
{
    reg();
    AVCodec *codec = avcodec_find_encoder_by_name(name);
    if (!codec) return -1;
    return fate_avi_impl(codec, batch, raw, nframes, w, h, pix_fmt, level, slices, out, cap);
}

static int64_t fate_avi_impl(AVCodec *codec, int batch, const uint8_t *raw, int nframes, int w, int h, const char *pix_fmt,
                             int level, int slices, uint8_t *out, int64_t cap)
{
    static int fmt_registered;
    enum AVPixelFormat pf = av_get_pix_fmt(pix_fmt);
    AVFormatContext *oc = NULL;
    AVStream *st;
    AVCodecContext *enc;
    AVFrame *frame;
    MemOut mo = { out, cap, 0, 0 };
    uint8_t *iobuf;
    int i, ret, fsize;
    reg();
    if (!fmt_registered) { av_register_output_format(&ff_avi_muxer); av_register_output_format(&ff_nut_muxer); av_register_output_format(&ff_matroska_muxer); fmt_registered = 1; }
    if (avformat_alloc_output_context2(&oc, g_muxer, NULL, NULL) < 0) return -1;
    iobuf = av_malloc(32768);
    oc->pb = avio_alloc_context(iobuf, 32768, 1, &mo, NULL, mem_write, mem_seek);
    oc->flags |= AVFMT_FLAG_BITEXACT;
    st = avformat_new_stream(oc, NULL);

    enc = avcodec_alloc_context3(codec);
    enc->width = w; enc->height = h; enc->pix_fmt = pf;
    enc->time_base = (AVRational){1, 25};
    enc->framerate = (AVRational){25, 1};
    enc->level = level;
    enc->slices = slices;
    enc->thread_count = 1;
    enc->flags |= AV_CODEC_FLAG_BITEXACT;
    enc->sample_aspect_ratio = (AVRational){0, 1};
    if (oc->oformat->flags & AVFMT_GLOBALHEADER) enc->flags |= AV_CODEC_FLAG_GLOBAL_HEADER;
    if (batch > 0) av_opt_set_int(enc->priv_data, "batch", batch, 0);
    if ((ret = avcodec_open2(enc, codec, NULL)) < 0) return ret;
    /* ffmpeg.c init_output_stream: avcodec_parameters_from_context + time base / frame rate hints */
    avcodec_parameters_from_context(st->codecpar, enc);
    avcodec_copy_context(st->codec, enc);
    st->time_base = enc->time_base;
    st->avg_frame_rate = (AVRational){25, 1};
    st->sample_aspect_ratio = enc->sample_aspect_ratio;
    av_dict_set(&st->metadata, "encoder", "Lavc ffv1", 0);
    if ((ret = avformat_write_header(oc, NULL)) < 0) return ret;

    frame = av_frame_alloc();
    fsize = av_image_get_buffer_size(pf, w, h, 1);
    for (i = 0; i < nframes; i++) {
        AVPacket pkt;
        int got = 0;
        av_init_packet(&pkt); pkt.data = NULL; pkt.size = 0;
        frame->format = pf; frame->width = w; frame->height = h;
        av_frame_get_buffer(frame, 32);
        {
            uint8_t *src[4]; int ls[4];
            av_image_fill_arrays(src, ls, raw + (int64_t)i * fsize, pf, w, h, 1);
            av_image_copy(frame->data, frame->linesize, (const uint8_t **)src, ls, pf, w, h);
        }
        frame->pts = i;
        frame->sample_aspect_ratio = (AVRational){0, 1};
        if ((ret = avcodec_encode_video2(enc, &pkt, frame, &got)) < 0) return ret;
        av_frame_unref(frame);
        if (got) {
            av_packet_rescale_ts(&pkt, enc->time_base, st->time_base);
            pkt.stream_index = 0;
            if ((ret = av_interleaved_write_frame(oc, &pkt)) < 0) return ret;
        }
    }
    for (;;) {                                       /* drain a delayed encoder (ffmpeg.c flush_encoders, 1698-1770) */
        AVPacket pkt;
        int got = 0;
        av_init_packet(&pkt); pkt.data = NULL; pkt.size = 0;
        if ((ret = avcodec_encode_video2(enc, &pkt, NULL, &got)) < 0) return ret;
        if (!got) break;
        av_packet_rescale_ts(&pkt, enc->time_base, st->time_base);
        pkt.stream_index = 0;
        if ((ret = av_interleaved_write_frame(oc, &pkt)) < 0) return ret;
    }
    av_write_trailer(oc);
    avio_flush(oc->pb);
    av_frame_free(&frame);
    avcodec_close(enc); avcodec_free_context(&enc);
    av_freep(&oc->pb->buffer); av_freep(&oc->pb);
    avformat_free_context(oc);
    return mo.size;
}

/* ---- container round trip: the same encode-and-mux procedure into NUT (libavformat/nutenc.c), and
 * the way back: the reference's NUT demuxer (nutdec.c) feeding a named decoder.  Everything in memory. */
int64_t ffv1ref_mux_named(const char *muxer, const char *name, int batch, const uint8_t *raw, int nframes, int w, int h,
                          const char *pix_fmt, int level, int slices, uint8_t *out, int64_t cap)
{
    int64_t r;
    reg();
    AVCodec *codec = avcodec_find_encoder_by_name(name);
    if (!codec) return -1;
    g_muxer = !strcmp(muxer, "nut") ? &ff_nut_muxer : (!strcmp(muxer, "matroska") ? &ff_matroska_muxer : &ff_avi_muxer);
    r = fate_avi_impl(codec, batch, raw, nframes, w, h, pix_fmt, level, slices, out, cap);
    g_muxer = &ff_avi_muxer;
    return r;
}

typedef struct MemIn { const uint8_t *buf; int64_t size, pos; } MemIn;
static int mem_read(void *opaque, uint8_t *buf, int n)
{
    MemIn *m = opaque;
    if (m->pos >= m->size) return AVERROR_EOF;
    if (n > m->size - m->pos) n = (int)(m->size - m->pos);
    memcpy(buf, m->buf + m->pos, n);
    m->pos += n;
    return n;
}
static int64_t mem_in_seek(void *opaque, int64_t off, int whence)
{
    MemIn *m = opaque;
    if (whence == AVSEEK_SIZE) return m->size;
    if (whence == SEEK_CUR) off += m->pos;
    else if (whence == SEEK_END) off += m->size;
    if (off < 0 || off > m->size) return -1;
    m->pos = off;
    return off;
}

/* NUT file in memory -> pictures of the named decoder, tightly packed one after the other in dst.  Returns the number of
 * pictures (the decoder is drained with empty packets at the end), < 0 on error; *frame_bytes = bytes per picture. */
int ffv1ref_nut_decode_named(const char *decoder, const char *dec_opts, const uint8_t *file, int64_t size,
                             uint8_t *dst, int64_t cap, int *frame_bytes, char *fmt_name)
{
    static int registered;
    AVFormatContext *ic = avformat_alloc_context();
    MemIn mi = { file, size, 0 };
    AVCodec *codec;
    AVCodecContext *dc;
    AVFrame *frame = av_frame_alloc();
    AVPacket pkt;
    uint8_t *iobuf = av_malloc(32768);
    int ret, n = 0, got, fb = 0, eof = 0;
    reg();
    if (!registered) { av_register_input_format(&ff_nut_demuxer); registered = 1; }
    ic->pb = avio_alloc_context(iobuf, 32768, 0, &mi, mem_read, NULL, mem_in_seek);
    if ((ret = avformat_open_input(&ic, NULL, &ff_nut_demuxer, NULL)) < 0) return ret;
    if (ic->nb_streams < 1) return -10;
    codec = avcodec_find_decoder_by_name(decoder);
    if (!codec) return -11;
    dc = avcodec_alloc_context3(codec);
    if ((ret = avcodec_parameters_to_context(dc, ic->streams[0]->codecpar)) < 0) return ret;
    dc->flags |= AV_CODEC_FLAG_BITEXACT;
    dc->thread_count = 1;
    if (dec_opts && *dec_opts && av_set_options_string(dc->priv_data, dec_opts, "=", ":") < 0) return -12;
    if ((ret = avcodec_open2(dc, codec, NULL)) < 0) return ret;
    for (;;) {
        av_init_packet(&pkt); pkt.data = NULL; pkt.size = 0;
        if (!eof && av_read_frame(ic, &pkt) < 0) eof = 1;
        ret = avcodec_decode_video2(dc, frame, &got, &pkt);
        if (!eof) av_packet_unref(&pkt);
        if (ret < 0) return ret;
        if (got) {
            fb = av_image_get_buffer_size(frame->format, frame->width, frame->height, 1);
            if ((int64_t)(n + 1) * fb > cap) return -13;
            av_image_copy_to_buffer(dst + (int64_t)n * fb, fb, (const uint8_t *const *)frame->data, frame->linesize,
                                    frame->format, frame->width, frame->height, 1);
            if (fmt_name) { strncpy(fmt_name, av_get_pix_fmt_name(frame->format), 31); fmt_name[31] = 0; }
            av_frame_unref(frame);
            n++;
        } else if (eof) break;
    }
    if (frame_bytes) *frame_bytes = fb;
    av_frame_free(&frame);
    avcodec_close(dc); avcodec_free_context(&dc);
    av_freep(&ic->pb->buffer); av_freep(&ic->pb);
    avformat_close_input(&ic);
    return n;
}

