/*
 * ffv1_b200_avcodec.c -- the libavcodec side of the drop-in: an AVCodec pair ("ffv1_b200") that exposes the
 * B200-native FFV1 core (libffv1_b200.so, include/ffv1_b200.h) behind exactly the interface the reference's own
 * ff_ffv1_encoder / ff_ffv1_decoder present (libavcodec/ffv1enc.c:1415-1444, ffv1dec.c:1139-1153):
 * same codec id, same pix_fmts, same private AVOptions (slicecrc / coder / context, ffv1enc.c:1383-1399), same
 * generic options read from AVCodecContext (gop_size, level, slices), extradata written at init, packets with
 * pts/dts/AV_PKT_FLAG_KEY as in ffv1enc.c:1365-1370.
 *
 * It is written against the reference tree's PUBLIC headers only (avcodec.h, opt.h, pixdesc.h, imgutils.h) and is
 * compiled by integration/Makefile where that tree is available.  Register it with avcodec_register()
 * (libavcodec/utils.c:178) or add REGISTER_ENCDEC(FFV1_B200, ffv1_b200) next to allcodecs.c:180.
 *
 * The encoder has AV_CODEC_CAP_DELAY like the reference's: frames are gathered into batches ("batch" private option,
 * default 64) so that the GPU sees many GOPs at once; the caller drains with frame == NULL as for any delayed codec
 * (ffmpeg.c:1698-1770).  There is no CPU fallback: init fails with AVERROR_EXTERNAL when no B200 is usable.
 */
#include <string.h>
#include "libavcodec/avcodec.h"
#include "libavutil/opt.h"
#include "libavutil/pixdesc.h"
#include "libavutil/imgutils.h"
#include "libavutil/mem.h"
#include "libavutil/log.h"
#include "libavutil/hwcontext.h"
#include "../include/ffv1_b200.h"

/* ------------------------------------------------------------------------------------------------ encoder */
typedef struct B200EncContext {
    const AVClass *class;              /* must be first (AVOptions) */
    int ec, ac, context_model;         /* same names / meaning as FFV1Context.ec/.ac/.context_model */
    int batch, device;
    int cuda_frames;                   /* avctx->pix_fmt == AV_PIX_FMT_CUDA: frame->data[] are device pointers */
    FFV1B200Encoder *enc;
    AVFrame **queue;                   /* frames waiting for a full batch (references) */
    int nqueued;
    uint8_t *outbuf;                   /* packets of the last batch */
    size_t outcap;
    FFV1B200Packet *pkts;
    int64_t *pts;
    int nready, next_ready;
    FFV1B200FrameProps props;
} B200EncContext;

static int b200_run_batch(AVCodecContext *avctx)
{
    B200EncContext *s = avctx->priv_data;
    const uint8_t *planes[4 * 1024];
    int linesizes[4 * 1024];
    size_t needed = 0;
    int i, p, ret;
    if (!s->nqueued)
        return 0;
    for (i = 0; i < s->nqueued; i++)
        for (p = 0; p < 4; p++) {
            planes[4 * i + p]    = s->queue[i]->data[p];
            linesizes[4 * i + p] = s->queue[i]->linesize[p];
        }
    for (;;) {
        if (s->cuda_frames)
            ret = ffv1b200_enc_encode_cuda(s->enc, s->nqueued, (const void *const *)planes, linesizes, s->outbuf, s->outcap, s->pkts, &needed);
        else
            ret = ffv1b200_enc_encode_host(s->enc, s->nqueued, planes, linesizes, s->outbuf, s->outcap, s->pkts, &needed);
        if (ret != FFV1B200_ERR_BUFFER_TOO_SMALL)
            break;
        av_freep(&s->outbuf);
        s->outcap = needed + 4096;
        if (!(s->outbuf = av_malloc(s->outcap)))
            return AVERROR(ENOMEM);
    }
    if (ret < 0) {
        av_log(avctx, AV_LOG_ERROR, "ffv1_b200: %s\n", ffv1b200_last_error());
        return ret;
    }
    for (i = 0; i < s->nqueued; i++) {
        s->pts[i] = s->queue[i]->pts;
        av_frame_free(&s->queue[i]);
    }
    s->nready = s->nqueued;
    s->next_ready = 0;
    s->nqueued = 0;
    return 0;
}

static av_cold int b200_encode_init(AVCodecContext *avctx)
{
    B200EncContext *s = avctx->priv_data;
    FFV1B200EncParams p;
    FFV1B200EncInfo info;
    const uint8_t *ed;
    int edsize, ret;

    enum AVPixelFormat sw_fmt = avctx->pix_fmt;
    if (avctx->pix_fmt == AV_PIX_FMT_CUDA) {
        /* frames live in device memory (hwcontext_cuda.h:31-40); the software format comes from the frames context,
         * as in nvenc.c:412-421 */
        AVHWFramesContext *fc;
        if (!avctx->hw_frames_ctx) {
            av_log(avctx, AV_LOG_ERROR, "AV_PIX_FMT_CUDA input needs avctx->hw_frames_ctx\n");
            return AVERROR(EINVAL);
        }
        fc = (AVHWFramesContext *)avctx->hw_frames_ctx->data;
        sw_fmt = fc->sw_format;
        s->cuda_frames = 1;
    }
    memset(&p, 0, sizeof(p));
    p.width = avctx->width; p.height = avctx->height;
    p.pix_fmt = av_get_pix_fmt_name(sw_fmt);
    p.gop_size = avctx->gop_size;
    p.level = avctx->level;
    p.slices = avctx->slices;
    p.coder = s->ac; p.context = s->context_model; p.slicecrc = s->ec;
    p.device = s->device;
    p.max_batch_frames = s->batch > 1024 ? 1024 : s->batch;
    p.first_picture_number = 0;
    if ((ret = ffv1b200_enc_open(&s->enc, &p)) < 0) {
        av_log(avctx, AV_LOG_ERROR, "ffv1_b200: %s\n", ffv1b200_last_error());
        return ret;
    }
    ffv1b200_enc_info(s->enc, &info);
    s->batch = info.max_batch_frames;
    avctx->bits_per_raw_sample = info.bits_per_raw_sample;
    ffv1b200_enc_extradata(s->enc, &ed, &edsize);
    if (edsize > 0) {                                   /* ffv1enc.c:556-558: owned by lavc, freed in avcodec_close */
        avctx->extradata = av_mallocz(edsize + AV_INPUT_BUFFER_PADDING_SIZE);
        if (!avctx->extradata)
            return AVERROR(ENOMEM);
        memcpy(avctx->extradata, ed, edsize);
        avctx->extradata_size = edsize;
    }
    s->queue = av_mallocz_array(s->batch, sizeof(*s->queue));
    s->pkts  = av_mallocz_array(s->batch, sizeof(*s->pkts));
    s->pts   = av_mallocz_array(s->batch, sizeof(*s->pts));
    s->outcap = (size_t)s->batch * (info.frame_bytes + info.frame_bytes / 4 + 65536);
    s->outbuf = av_malloc(s->outcap);
    if (!s->queue || !s->pkts || !s->pts || !s->outbuf)
        return AVERROR(ENOMEM);
    s->props.sar_num = 0; s->props.sar_den = 1; s->props.picture_structure = 3;
    return 0;
}

static int b200_encode_frame(AVCodecContext *avctx, AVPacket *pkt, const AVFrame *frame, int *got_packet)
{
    B200EncContext *s = avctx->priv_data;
    int ret;
    *got_packet = 0;
    if (frame) {
        /* sample aspect ratio and field order are coded into every slice header (ffv1enc.c:1044-1049) */
        FFV1B200FrameProps pr;
        pr.sar_num = frame->sample_aspect_ratio.num;
        pr.sar_den = frame->sample_aspect_ratio.den;
        pr.picture_structure = frame->interlaced_frame ? (frame->top_field_first ? 1 : 2) : 3;
        /* invariant: packets not yet returned + frames queued <= batch, and one packet leaves per call, so a full queue
         * means every packet of the previous batch has been handed out */
        if (s->nqueued == s->batch && (ret = b200_run_batch(avctx)) < 0)
            return ret;
        if (memcmp(&pr, &s->props, sizeof(pr))) {
            if (s->nqueued && s->nready == s->next_ready && (ret = b200_run_batch(avctx)) < 0)
                return ret;
            if (!s->nqueued) {                              /* takes effect at a batch boundary */
                s->props = pr;
                ffv1b200_enc_set_frame_props(s->enc, &pr);
            }
        }
        if (!(s->queue[s->nqueued] = av_frame_clone(frame)))
            return AVERROR(ENOMEM);
        s->nqueued++;
        if (s->nqueued == s->batch && s->nready == s->next_ready && (ret = b200_run_batch(avctx)) < 0)
            return ret;
    } else if (s->nready == s->next_ready && s->nqueued) {
        if ((ret = b200_run_batch(avctx)) < 0)           /* drain (AV_CODEC_CAP_DELAY) */
            return ret;
    }
    if (s->next_ready < s->nready) {
        const FFV1B200Packet *pk = &s->pkts[s->next_ready];
        if ((ret = av_new_packet(pkt, pk->size)) < 0)
            return ret;
        memcpy(pkt->data, s->outbuf + pk->offset, pk->size);
        pkt->pts = pkt->dts = s->pts[s->next_ready];
        if (pk->flags & FFV1B200_PKT_FLAG_KEY)
            pkt->flags |= AV_PKT_FLAG_KEY;
        s->next_ready++;
        *got_packet = 1;
    }
    return 0;
}

static av_cold int b200_encode_close(AVCodecContext *avctx)
{
    B200EncContext *s = avctx->priv_data;
    int i;
    for (i = 0; s->queue && i < s->nqueued; i++)
        av_frame_free(&s->queue[i]);
    av_freep(&s->queue); av_freep(&s->pkts); av_freep(&s->pts); av_freep(&s->outbuf);
    ffv1b200_enc_close(s->enc);
    s->enc = NULL;
    return 0;
}

#define OFFSET(x) offsetof(B200EncContext, x)
#define VE AV_OPT_FLAG_VIDEO_PARAM | AV_OPT_FLAG_ENCODING_PARAM
static const AVOption b200_enc_options[] = {
    /* the reference's private options, unchanged (ffv1enc.c:1383-1399) */
    { "slicecrc", "Protect slices with CRCs", OFFSET(ec), AV_OPT_TYPE_BOOL, { .i64 = -1 }, -1, 1, VE },
    { "coder", "Coder type", OFFSET(ac), AV_OPT_TYPE_INT, { .i64 = 0 }, -2, 2, VE, "coder" },
        { "rice", "Golomb rice", 0, AV_OPT_TYPE_CONST, { .i64 = 0 }, INT_MIN, INT_MAX, VE, "coder" },
        { "range_def", "Range with default table", 0, AV_OPT_TYPE_CONST, { .i64 = -2 }, INT_MIN, INT_MAX, VE, "coder" },
        { "range_tab", "Range with custom table", 0, AV_OPT_TYPE_CONST, { .i64 = 2 }, INT_MIN, INT_MAX, VE, "coder" },
        { "ac", "Range with custom table (the ac option exists for compatibility and is deprecated)", 0, AV_OPT_TYPE_CONST, { .i64 = 1 }, INT_MIN, INT_MAX, VE, "coder" },
    { "context", "Context model", OFFSET(context_model), AV_OPT_TYPE_INT, { .i64 = 0 }, 0, 1, VE },
    /* additions of the GPU codec */
    { "batch", "Frames gathered per GPU submission", OFFSET(batch), AV_OPT_TYPE_INT, { .i64 = 64 }, 1, 1024, VE },
    { "gpu", "CUDA device ordinal", OFFSET(device), AV_OPT_TYPE_INT, { .i64 = 0 }, 0, 64, VE },
    { NULL }
};

static const AVClass b200_enc_class = {
    .class_name = "ffv1_b200 encoder",
    .item_name  = av_default_item_name,
    .option     = b200_enc_options,
    .version    = LIBAVUTIL_VERSION_INT,
};

/* The reference also installs ffv1_defaults {"coder","-1"} (ffv1enc.c:1409-1412) to detect use of the DEPRECATED generic
 * AVCodecContext.coder_type; AVCodecDefault is an internal type, and this codec simply ignores coder_type. */

AVCodec ff_ffv1_b200_encoder = {
    .name           = "ffv1_b200",
    .long_name      = "FFmpeg video codec #1 (B200-native CUDA core)",
    .type           = AVMEDIA_TYPE_VIDEO,
    .id             = AV_CODEC_ID_FFV1,
    .priv_data_size = sizeof(B200EncContext),
    .init           = b200_encode_init,
    .encode2        = b200_encode_frame,
    .close          = b200_encode_close,
    .capabilities   = AV_CODEC_CAP_DELAY,
    .pix_fmts       = (const enum AVPixelFormat[]) {
        AV_PIX_FMT_YUV420P,   AV_PIX_FMT_YUVA420P,  AV_PIX_FMT_YUVA422P,  AV_PIX_FMT_YUV444P,
        AV_PIX_FMT_YUVA444P,  AV_PIX_FMT_YUV440P,   AV_PIX_FMT_YUV422P,   AV_PIX_FMT_YUV411P,
        AV_PIX_FMT_YUV410P,   AV_PIX_FMT_0RGB32,    AV_PIX_FMT_RGB32,     AV_PIX_FMT_YUV420P16,
        AV_PIX_FMT_YUV422P16, AV_PIX_FMT_YUV444P16, AV_PIX_FMT_YUV444P9,  AV_PIX_FMT_YUV422P9,
        AV_PIX_FMT_YUV420P9,  AV_PIX_FMT_YUV420P10, AV_PIX_FMT_YUV422P10, AV_PIX_FMT_YUV444P10,
        AV_PIX_FMT_YUVA444P16, AV_PIX_FMT_YUVA422P16, AV_PIX_FMT_YUVA420P16,
        AV_PIX_FMT_YUVA444P10, AV_PIX_FMT_YUVA422P10, AV_PIX_FMT_YUVA420P10,
        AV_PIX_FMT_YUVA444P9, AV_PIX_FMT_YUVA422P9, AV_PIX_FMT_YUVA420P9,
        AV_PIX_FMT_GRAY16,    AV_PIX_FMT_GRAY8,     AV_PIX_FMT_GBRP9,     AV_PIX_FMT_GBRP10,
        AV_PIX_FMT_GBRP12,    AV_PIX_FMT_GBRP14,    AV_PIX_FMT_YA8,
        AV_PIX_FMT_CUDA,      /* device-resident frames of any of the above (sw_format in avctx->hw_frames_ctx) */
        AV_PIX_FMT_NONE
    },
    .priv_class     = &b200_enc_class,
};

/* ------------------------------------------------------------------------------------------------ decoder */
typedef struct B200DecContext {
    FFV1B200Decoder *dec;
    FFV1B200DecInfo info;
    uint8_t *framebuf;
} B200DecContext;

static av_cold int b200_decode_init(AVCodecContext *avctx)
{
    B200DecContext *s = avctx->priv_data;
    FFV1B200DecParams p;
    int ret;
    memset(&p, 0, sizeof(p));
    p.width = avctx->width; p.height = avctx->height;
    p.extradata = avctx->extradata; p.extradata_size = avctx->extradata_size;
    p.max_batch_frames = 1;                              /* decode_frame returns one picture per packet (no delay) */
    if ((ret = ffv1b200_dec_open(&s->dec, &p)) < 0) {
        av_log(avctx, AV_LOG_ERROR, "ffv1_b200: %s\n", ffv1b200_last_error());
        return ret;
    }
    ffv1b200_dec_info(s->dec, &s->info);
    if (s->info.frame_bytes) {                           /* version >= 2: everything is known from the extradata */
        avctx->pix_fmt = av_get_pix_fmt(s->info.pix_fmt);
        avctx->bits_per_raw_sample = s->info.bits_per_raw_sample;
        if (!(s->framebuf = av_malloc(s->info.frame_bytes)))
            return AVERROR(ENOMEM);
    }
    return 0;
}

static int b200_decode_frame(AVCodecContext *avctx, void *data, int *got_frame, AVPacket *avpkt)
{
    B200DecContext *s = avctx->priv_data;
    AVFrame *frame = data;
    const uint8_t *pk = avpkt->data;
    const uint8_t *src[4];
    int ls[4], key = 0, size = avpkt->size, ret;
    uint64_t damaged = 0;
    if (!s->info.frame_bytes) {
        /* version 0/1: parameters are carried by the first keyframe (ffv1dec.c:646-696); a probe call parses them */
        uint8_t probe[16];
        ret = ffv1b200_dec_decode_host(s->dec, 1, &pk, &size, probe, 0, NULL, NULL);
        if (ret < 0 && ret != FFV1B200_ERR_BUFFER_TOO_SMALL) {
            av_log(avctx, AV_LOG_ERROR, "ffv1_b200: %s\n", ffv1b200_last_error());
            return ret;
        }
        ffv1b200_dec_info(s->dec, &s->info);
        avctx->pix_fmt = av_get_pix_fmt(s->info.pix_fmt);
        avctx->bits_per_raw_sample = s->info.bits_per_raw_sample;
        if (!(s->framebuf = av_malloc(s->info.frame_bytes)))
            return AVERROR(ENOMEM);
    }
    if ((ret = ffv1b200_dec_decode_host(s->dec, 1, &pk, &size, s->framebuf, s->info.frame_bytes, &key, &damaged)) < 0) {
        av_log(avctx, AV_LOG_ERROR, "ffv1_b200: %s\n", ffv1b200_last_error());
        return ret;
    }
    /* what ff_get_buffer() (internal) would fill in before calling the allocator (utils.c:890-960) */
    frame->width = avctx->width; frame->height = avctx->height; frame->format = avctx->pix_fmt;
    frame->sample_aspect_ratio = avctx->sample_aspect_ratio;
    if ((ret = avcodec_default_get_buffer2(avctx, frame, 0)) < 0)
        return ret;
    av_image_fill_arrays((uint8_t **)src, ls, s->framebuf, avctx->pix_fmt, avctx->width, avctx->height, 1);
    av_image_copy(frame->data, frame->linesize, src, ls, avctx->pix_fmt, avctx->width, avctx->height);
    frame->key_frame = key;
    frame->pict_type = AV_PICTURE_TYPE_I;                /* ffv1dec.c:913 */
    if (damaged)
        av_log(avctx, AV_LOG_ERROR, "slice(s) damaged: mask %#llx (concealed from the previous frame)\n", (unsigned long long)damaged);
    *got_frame = 1;
    return avpkt->size;                                  /* ffv1dec.c:1034 */
}

static av_cold int b200_decode_close(AVCodecContext *avctx)
{
    B200DecContext *s = avctx->priv_data;
    av_freep(&s->framebuf);
    ffv1b200_dec_close(s->dec);
    s->dec = NULL;
    return 0;
}

AVCodec ff_ffv1_b200_decoder = {
    .name           = "ffv1_b200",
    .long_name      = "FFmpeg video codec #1 (B200-native CUDA core)",
    .type           = AVMEDIA_TYPE_VIDEO,
    .id             = AV_CODEC_ID_FFV1,
    .priv_data_size = sizeof(B200DecContext),
    .init           = b200_decode_init,
    .close          = b200_decode_close,
    .decode         = b200_decode_frame,
    .capabilities   = AV_CODEC_CAP_DR1,
};
