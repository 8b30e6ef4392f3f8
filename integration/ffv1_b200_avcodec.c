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
 * Encoder (AV_CODEC_CAP_DELAY like the reference's, ffv1enc.c:1424): frames are gathered into batches ("batch" private
 * option) so that the GPU sees many GOPs at once, and the batches are PIPELINED:
 *
 *   caller thread   encode2(frame): reference the frame, hand it to the copy workers, return a finished packet if one waits
 *   copy workers    pageable AVFrame rows -> one of three pinned staging areas ("copy_threads" option), and finished packets
 *                   -> refcounted AVPacket buffers, both off the caller thread (libavcodec's encode wrapper wants every
 *                   packet in a buffer of its own, utils.c:1965-1995: one 1.2 MB copy per 1080p packet)
 *   batch full      ffv1b200_enc_submit_host(staging area): H2D copies + kernels are queued and the call returns; the
 *                   packets of the batch before last are collected (ffv1b200_enc_collect_async) into one of two pinned
 *                   output areas and handed out one per call
 *
 * so filling batch k+1, coding batch k and returning the packets of batch k-1 overlap.  Sample aspect ratio / field order
 * are per-frame properties coded into every slice header (ffv1enc.c:1044-1049): a change closes the current batch early
 * (batches carry their own properties through the library).  The caller drains with frame == NULL as for any delayed
 * codec (ffmpeg.c:1698-1770).
 *
 * Decoder: "batch" private option (default 1 = one picture per packet, no delay, like the reference).  With batch > 1 the
 * packets are gathered and decoded together (the GPU decoder's parallelism is GOPs x slices in flight), pictures come out
 * batch-1 calls later and the caller drains with empty packets (AV_CODEC_CAP_DELAY).
 *
 * There is no CPU fallback: init fails with AVERROR_EXTERNAL when no B200 is usable.
 */
#include <string.h>
#include <pthread.h>
#include "libavcodec/avcodec.h"
#include "libavutil/opt.h"
#include "libavutil/pixdesc.h"
#include "libavutil/imgutils.h"
#include "libavutil/mem.h"
#include "libavutil/log.h"
#include "libavutil/hwcontext.h"
#include "../include/ffv1_b200.h"

#define NSTAGE 3                       /* staging areas: one being filled + up to two batches in flight */
#define NOUT   2                       /* packet areas: one being handed out + one being written */

/* ------------------------------------------------------------------------------------------------ encoder */
struct ReadyPacket;
typedef struct CopyJob {
    AVFrame *src;                      /* frame job: reference held until the rows are copied */
    uint8_t *dst;                      /*            frame slot inside a staging area */
    int stage;
    struct ReadyPacket *rp;            /* packet job (src == NULL): bytes of a finished packet -> its own AVBufferRef */
} CopyJob;

typedef struct Batch {
    int stage, n;
    int64_t *pts;                      /* [batch] */
} Batch;

typedef struct ReadyPacket {
    const uint8_t *data;
    int size, flags, out;              /* out: packet area the bytes live in, -1 = own allocation (freed after hand-out) */
    int64_t pts;
    AVBufferRef *buf;                  /* the packet in a buffer of its own, made by a copy worker (else NULL) */
    int copied;                        /* the worker is done with this entry (or no worker was asked) */
} ReadyPacket;

typedef struct B200EncContext {
    const AVClass *class;              /* must be first (AVOptions) */
    int ec, ac, context_model;         /* same names / meaning as FFV1Context.ec/.ac/.context_model */
    int batch, device, copy_threads;
    int cuda_frames;                   /* avctx->pix_fmt == AV_PIX_FMT_CUDA: frame->data[] are device pointers */
    FFV1B200Encoder *enc;
    /* layout of one staged frame (planes back to back, 16-byte aligned pitches) */
    int nplanes, rows[4], rowbytes[4], pitch[4];
    size_t plane_off[4], frame_stride;
    uint8_t *stage[NSTAGE];
    int cur, nfill;                    /* staging area being filled, frames placed in it */
    int64_t *pts_fill;                 /* [batch] pts of the frames being gathered */
    FFV1B200FrameProps props;          /* of the frames in the batch being gathered */
    const uint8_t **planes; int *linesizes;     /* [4 * batch] argument tables of a submit */
    AVFrame **cuda_queue;              /* device-frame path: references held until the (blocking) batch call */
    Batch inflight[2]; int ninflight;  /* oldest first */
    uint8_t *out[NOUT]; size_t outcap[NOUT];
    int out_next, out_unsynced;        /* area the next collect writes; a device->host copy may still be running */
    FFV1B200Packet *pkts;              /* [batch] */
    ReadyPacket *ready; int ready_cap, ready_head, ready_count;
    /* copy workers */
    pthread_t *threads; int nthreads, quit;
    pthread_mutex_t lock; pthread_cond_t has_job, job_done;
    CopyJob *jobs; int job_cap, job_head, job_count;
    int outstanding[NSTAGE];           /* copies of a staging area not finished yet */
    int sync_init;
} B200EncContext;

static void copy_frame_rows(const B200EncContext *s, const AVFrame *f, uint8_t *dst)
{
    int p, y;
    for (p = 0; p < s->nplanes; p++) {
        uint8_t *d = dst + s->plane_off[p];
        const uint8_t *src = f->data[p];
        if (f->linesize[p] == s->pitch[p])
            memcpy(d, src, (size_t)s->pitch[p] * (s->rows[p] - 1) + s->rowbytes[p]);
        else
            for (y = 0; y < s->rows[p]; y++)
                memcpy(d + (size_t)y * s->pitch[p], src + (ptrdiff_t)y * f->linesize[p], s->rowbytes[p]);
    }
}

static void *copy_worker(void *arg)
{
    B200EncContext *s = arg;
    ffv1b200_bind_thread_to_device(s->device);         /* the staging areas live on the GPU's NUMA node */
    pthread_mutex_lock(&s->lock);
    for (;;) {
        CopyJob j;
        while (!s->job_count && !s->quit)
            pthread_cond_wait(&s->has_job, &s->lock);
        if (!s->job_count)
            break;
        j = s->jobs[s->job_head];
        s->job_head = (s->job_head + 1) % s->job_cap;
        s->job_count--;
        pthread_mutex_unlock(&s->lock);
        if (j.src) {
            copy_frame_rows(s, j.src, j.dst);
            av_frame_free(&j.src);
            pthread_mutex_lock(&s->lock);
            if (--s->outstanding[j.stage] == 0)
                pthread_cond_broadcast(&s->job_done);
        } else {
            /* the buffer av_new_packet() would make (reallocatable, padded), filled here instead of on the caller's thread */
            AVBufferRef *buf = NULL;
            if (av_buffer_realloc(&buf, j.rp->size + AV_INPUT_BUFFER_PADDING_SIZE) >= 0) {
                memcpy(buf->data, j.rp->data, j.rp->size);
                memset(buf->data + j.rp->size, 0, AV_INPUT_BUFFER_PADDING_SIZE);
            }
            pthread_mutex_lock(&s->lock);
            j.rp->buf = buf;                               /* NULL: out of memory, the caller copies itself */
            j.rp->copied = 1;
            pthread_cond_broadcast(&s->job_done);
        }
    }
    pthread_mutex_unlock(&s->lock);
    return NULL;
}

static int b200_err(AVCodecContext *avctx, int ret)
{
    av_log(avctx, AV_LOG_ERROR, "ffv1_b200: %s\n", ffv1b200_last_error());
    return ret;
}

static void ready_push(B200EncContext *s, const uint8_t *data, int size, int flags, int64_t pts, int out)
{
    ReadyPacket *r = &s->ready[(s->ready_head + s->ready_count) % s->ready_cap];
    r->data = data; r->size = size; r->flags = flags; r->pts = pts; r->out = out;
    r->buf = NULL; r->copied = 1;
    s->ready_count++;
}

/* hands the packets pushed last (the n newest entries, their bytes complete) to the copy workers */
static void ready_precopy(B200EncContext *s, int n)
{
    int i;
    if (!s->nthreads)
        return;
    pthread_mutex_lock(&s->lock);
    for (i = s->ready_count - n; i < s->ready_count; i++) {
        ReadyPacket *r = &s->ready[(s->ready_head + i) % s->ready_cap];
        CopyJob *j = &s->jobs[(s->job_head + s->job_count) % s->job_cap];
        r->copied = 0;
        j->src = NULL; j->dst = NULL; j->stage = 0; j->rp = r;
        s->job_count++;
    }
    pthread_cond_broadcast(&s->has_job);
    pthread_mutex_unlock(&s->lock);
}

/* the packet area `o` is about to be overwritten: packets of it that have not been handed out yet move to their own
 * allocations (only happens when batches of very different sizes follow each other, e.g. around a property change) */
static int spill_area(B200EncContext *s, int o)
{
    int i;
    if (s->nthreads) {                                     /* packets the workers are still copying out of this area */
        pthread_mutex_lock(&s->lock);
        for (i = 0; i < s->ready_count; i++) {
            ReadyPacket *r = &s->ready[(s->ready_head + i) % s->ready_cap];
            while (r->out == o && !r->copied)
                pthread_cond_wait(&s->job_done, &s->lock);
        }
        pthread_mutex_unlock(&s->lock);
    }
    for (i = 0; i < s->ready_count; i++) {
        ReadyPacket *r = &s->ready[(s->ready_head + i) % s->ready_cap];
        if (r->out == o && !r->buf) {
            uint8_t *c = av_malloc(r->size ? r->size : 1);
            if (!c)
                return AVERROR(ENOMEM);
            memcpy(c, r->data, r->size);
            r->data = c; r->out = -1;
        }
    }
    return 0;
}

static int sync_output(AVCodecContext *avctx)
{
    B200EncContext *s = avctx->priv_data;
    int ret;
    if (s->out_unsynced) {
        if ((ret = ffv1b200_enc_sync_output(s->enc)) < 0)
            return b200_err(avctx, ret);
        s->out_unsynced = 0;
    }
    return 0;
}

/* packets of the oldest batch in flight -> ready queue */
static int collect_oldest(AVCodecContext *avctx)
{
    B200EncContext *s = avctx->priv_data;
    Batch b = s->inflight[0];
    const int o = s->out_next;
    size_t needed = 0;
    int ret, i;
    if ((ret = sync_output(avctx)) < 0 || (ret = spill_area(s, o)) < 0)
        return ret;
    for (;;) {
        ret = ffv1b200_enc_collect_async(s->enc, s->out[o], s->outcap[o], s->pkts, &needed);
        if (ret != FFV1B200_ERR_BUFFER_TOO_SMALL)
            break;
        ffv1b200_host_free(s->out[o]);
        s->outcap[o] = needed + needed / 8 + 4096;
        if (!(s->out[o] = ffv1b200_host_alloc(s->outcap[o], s->device)))
            return AVERROR(ENOMEM);
    }
    if (ret < 0)
        return b200_err(avctx, ret);
    s->out_unsynced = 1;
    for (i = 0; i < b.n; i++)
        ready_push(s, s->out[o] + s->pkts[i].offset, s->pkts[i].size, s->pkts[i].flags, b.pts[i], o);
    if (s->nthreads) {
        /* the workers turn the batch into AVPacket buffers while the caller gathers the next frames; they need the bytes
         * now (the next batch's kernels are already queued on the device, so nothing idles while the copy lands) */
        if ((ret = sync_output(avctx)) < 0)
            return ret;
        ready_precopy(s, b.n);
    }
    s->inflight[0] = s->inflight[1];
    s->inflight[1] = b;                                  /* keeps its pts array for reuse */
    s->ninflight--;
    s->out_next = (o + 1) % NOUT;
    return 0;
}

/* the batch being gathered goes to the GPU */
static int submit_current(AVCodecContext *avctx)
{
    B200EncContext *s = avctx->priv_data;
    const int n = s->nfill, st = s->cur;
    int ret, i, p;
    if (!n)
        return 0;
    if (s->cuda_frames) {
        /* device frames: one blocking call per batch (no staging; the frames are already where the kernels read them) */
        const int o = s->out_next;
        size_t needed = 0;
        if ((ret = sync_output(avctx)) < 0 || (ret = spill_area(s, o)) < 0)
            return ret;
        for (i = 0; i < n; i++)
            for (p = 0; p < 4; p++) {
                s->planes[4 * i + p]    = s->cuda_queue[i]->data[p];
                s->linesizes[4 * i + p] = s->cuda_queue[i]->linesize[p];
            }
        ffv1b200_enc_set_frame_props(s->enc, &s->props);
        for (;;) {
            ret = ffv1b200_enc_encode_cuda(s->enc, n, (const void *const *)s->planes, s->linesizes, s->out[o], s->outcap[o], s->pkts, &needed);
            if (ret != FFV1B200_ERR_BUFFER_TOO_SMALL)
                break;
            /* (the batch is coded and the model state has moved on only when the packets fitted the library's own area;
             *  a too small host area is reported before that, see ffv1b200_enc_encode_cuda) */
            ffv1b200_host_free(s->out[o]);
            s->outcap[o] = needed + needed / 8 + 4096;
            if (!(s->out[o] = ffv1b200_host_alloc(s->outcap[o], s->device)))
                return AVERROR(ENOMEM);
        }
        if (ret < 0)
            return b200_err(avctx, ret);
        for (i = 0; i < n; i++) {
            ready_push(s, s->out[o] + s->pkts[i].offset, s->pkts[i].size, s->pkts[i].flags, s->pts_fill[i], o);
            av_frame_free(&s->cuda_queue[i]);
        }
        s->out_next = (o + 1) % NOUT;
        s->nfill = 0;
        return 0;
    }
    /* every row of the batch must have reached the staging area */
    pthread_mutex_lock(&s->lock);
    while (s->outstanding[st])
        pthread_cond_wait(&s->job_done, &s->lock);
    pthread_mutex_unlock(&s->lock);
    if (s->ninflight == 2 && (ret = collect_oldest(avctx)) < 0)
        return ret;
    for (i = 0; i < n; i++)
        for (p = 0; p < 4; p++) {
            s->planes[4 * i + p]    = p < s->nplanes ? s->stage[st] + (size_t)i * s->frame_stride + s->plane_off[p] : NULL;
            s->linesizes[4 * i + p] = p < s->nplanes ? s->pitch[p] : 0;
        }
    ffv1b200_enc_set_frame_props(s->enc, &s->props);
    if ((ret = ffv1b200_enc_submit_host(s->enc, n, s->planes, s->linesizes)) < 0)
        return b200_err(avctx, ret);
    {
        Batch *b = &s->inflight[s->ninflight++];
        int64_t *t = b->pts;
        b->stage = st; b->n = n;
        b->pts = s->pts_fill; s->pts_fill = t;           /* swap the pts arrays instead of copying */
    }
    /* next area: not the one just submitted, not the one of the other batch in flight */
    for (i = 0; i < NSTAGE; i++)
        if (i != st && !(s->ninflight == 2 && s->inflight[0].stage == i))
            break;
    s->cur = i;
    s->nfill = 0;
    return 0;
}

static av_cold int b200_encode_close(AVCodecContext *avctx)
{
    B200EncContext *s = avctx->priv_data;
    int i;
    av_freep(&avctx->stats_out);                        /* the codec owns it, as in ff_ffv1_close (ffv1.c:229) */
    if (s->threads) {
        pthread_mutex_lock(&s->lock);
        s->quit = 1;
        pthread_cond_broadcast(&s->has_job);
        pthread_mutex_unlock(&s->lock);
        for (i = 0; i < s->nthreads; i++)
            pthread_join(s->threads[i], NULL);
        av_freep(&s->threads);
    }
    for (i = 0; s->jobs && i < s->job_count; i++)
        av_frame_free(&s->jobs[(s->job_head + i) % s->job_cap].src);
    for (i = 0; s->cuda_queue && i < s->nfill; i++)
        av_frame_free(&s->cuda_queue[i]);
    for (i = 0; s->ready && i < s->ready_count; i++) {
        ReadyPacket *r = &s->ready[(s->ready_head + i) % s->ready_cap];
        if (r->out < 0)
            av_free((void *)r->data);
        av_buffer_unref(&r->buf);
    }
    ffv1b200_enc_close(s->enc);                          /* waits for everything in flight */
    s->enc = NULL;
    for (i = 0; i < NSTAGE; i++) { ffv1b200_host_free(s->stage[i]); s->stage[i] = NULL; }
    for (i = 0; i < NOUT; i++)   { ffv1b200_host_free(s->out[i]);   s->out[i] = NULL; }
    av_freep(&s->jobs); av_freep(&s->cuda_queue); av_freep(&s->ready); av_freep(&s->pkts);
    av_freep(&s->planes); av_freep(&s->linesizes); av_freep(&s->pts_fill);
    av_freep(&s->inflight[0].pts); av_freep(&s->inflight[1].pts);
    if (s->sync_init) {
        pthread_mutex_destroy(&s->lock); pthread_cond_destroy(&s->has_job); pthread_cond_destroy(&s->job_done);
        s->sync_init = 0;
    }
    return 0;
}

static av_cold int b200_encode_init(AVCodecContext *avctx)
{
    B200EncContext *s = avctx->priv_data;
    FFV1B200EncParams p;
    FFV1B200EncInfo info;
    const AVPixFmtDescriptor *desc;
    const uint8_t *ed;
    size_t off = 0;
    int edsize, ret, i;

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
    p.max_batch_frames = s->batch;
    p.first_picture_number = 0;
    /* two-pass coding and the experimental gate, as encode_init reads them (ffv1enc.c:680, 703, 906) */
    p.flags = avctx->flags & (AV_CODEC_FLAG_PASS1 | AV_CODEC_FLAG_PASS2);
    p.stats_in = avctx->stats_in;
    p.strict_std_compliance = avctx->strict_std_compliance;
    p.bits_per_raw_sample = avctx->bits_per_raw_sample;
    if ((ret = ffv1b200_enc_open(&s->enc, &p)) < 0)
        return b200_err(avctx, ret);
    /* from here on every failure goes through b200_encode_close (this codec cannot set the internal
     * FF_CODEC_CAP_INIT_CLEANUP flag the reference's decoder uses, internal.h:48) */
    ffv1b200_enc_info(s->enc, &info);
    s->batch = info.max_batch_frames;
    avctx->bits_per_raw_sample = info.bits_per_raw_sample;
    ffv1b200_enc_extradata(s->enc, &ed, &edsize);
    if (edsize > 0) {                                   /* ffv1enc.c:556-558: owned by lavc, freed in avcodec_close */
        avctx->extradata = av_mallocz(edsize + AV_INPUT_BUFFER_PADDING_SIZE);
        if (!avctx->extradata)
            goto nomem;
        memcpy(avctx->extradata, ed, edsize);
        avctx->extradata_size = edsize;
    }
    desc = av_pix_fmt_desc_get(sw_fmt);
    s->nplanes = av_pix_fmt_count_planes(sw_fmt);
    for (i = 0; i < s->nplanes; i++) {
        const int chroma = (i == 1 || i == 2) && !(desc->flags & AV_PIX_FMT_FLAG_RGB);
        s->rows[i]     = chroma ? -((-avctx->height) >> desc->log2_chroma_h) : avctx->height;
        s->rowbytes[i] = av_image_get_linesize(sw_fmt, avctx->width, i);
        s->pitch[i]    = (s->rowbytes[i] + 15) & ~15;
        s->plane_off[i] = off;
        off += (size_t)s->pitch[i] * s->rows[i];
    }
    s->frame_stride = (off + 255) & ~(size_t)255;
    s->ready_cap = 4 * s->batch + 8;
    s->planes    = av_mallocz_array(4 * (size_t)s->batch, sizeof(*s->planes));
    s->linesizes = av_mallocz_array(4 * (size_t)s->batch, sizeof(*s->linesizes));
    s->pkts      = av_mallocz_array(s->batch, sizeof(*s->pkts));
    s->pts_fill  = av_mallocz_array(s->batch, sizeof(int64_t));
    s->inflight[0].pts = av_mallocz_array(s->batch, sizeof(int64_t));
    s->inflight[1].pts = av_mallocz_array(s->batch, sizeof(int64_t));
    s->ready     = av_mallocz_array(s->ready_cap, sizeof(*s->ready));
    if (!s->planes || !s->linesizes || !s->pkts || !s->pts_fill || !s->inflight[0].pts || !s->inflight[1].pts || !s->ready)
        goto nomem;
    for (i = 0; i < NOUT; i++) {
        s->outcap[i] = (size_t)s->batch * (info.frame_bytes / 2 + 65536);
        if (!(s->out[i] = ffv1b200_host_alloc(s->outcap[i], s->device)))
            goto nomem;
    }
    if (s->cuda_frames) {
        if (!(s->cuda_queue = av_mallocz_array(s->batch, sizeof(*s->cuda_queue))))
            goto nomem;
    } else {
        for (i = 0; i < NSTAGE; i++)
            if (!(s->stage[i] = ffv1b200_host_alloc(s->frame_stride * s->batch, s->device)))
                goto nomem;
        s->job_cap = 4 * s->batch + 8;                   /* frames being gathered + packets of the batches handed out */
        if (!(s->jobs = av_mallocz_array(s->job_cap, sizeof(*s->jobs))))
            goto nomem;
        pthread_mutex_init(&s->lock, NULL); pthread_cond_init(&s->has_job, NULL); pthread_cond_init(&s->job_done, NULL);
        s->sync_init = 1;
        if (s->copy_threads > 0) {
            if (!(s->threads = av_mallocz_array(s->copy_threads, sizeof(*s->threads))))
                goto nomem;
            for (i = 0; i < s->copy_threads; i++) {
                if (pthread_create(&s->threads[i], NULL, copy_worker, s))
                    break;
                s->nthreads++;
            }
        }
    }
    s->props.sar_num = 0; s->props.sar_den = 1; s->props.picture_structure = 3;
    return 0;
nomem:
    b200_encode_close(avctx);
    return AVERROR(ENOMEM);
}

static int b200_encode_frame(AVCodecContext *avctx, AVPacket *pkt, const AVFrame *frame, int *got_packet)
{
    B200EncContext *s = avctx->priv_data;
    int ret;
    *got_packet = 0;
    if (frame) {
        /* sample aspect ratio and field order are coded into every slice header (ffv1enc.c:1044-1049): frames with other
         * properties than the batch being gathered start a new batch */
        FFV1B200FrameProps pr;
        AVFrame *ref;
        pr.sar_num = frame->sample_aspect_ratio.num;
        pr.sar_den = frame->sample_aspect_ratio.den;
        pr.picture_structure = frame->interlaced_frame ? (frame->top_field_first ? 1 : 2) : 3;
        if (memcmp(&pr, &s->props, sizeof(pr))) {
            if ((ret = submit_current(avctx)) < 0)
                return ret;
            s->props = pr;
        }
        if (!(ref = av_frame_clone(frame)))
            return AVERROR(ENOMEM);
        s->pts_fill[s->nfill] = frame->pts;
        if (s->cuda_frames) {
            s->cuda_queue[s->nfill] = ref;
        } else if (s->nthreads) {
            pthread_mutex_lock(&s->lock);
            CopyJob *j = &s->jobs[(s->job_head + s->job_count) % s->job_cap];
            j->src = ref; j->stage = s->cur; j->rp = NULL;
            j->dst = s->stage[s->cur] + (size_t)s->nfill * s->frame_stride;
            s->job_count++;
            s->outstanding[s->cur]++;
            pthread_cond_signal(&s->has_job);
            pthread_mutex_unlock(&s->lock);
        } else {
            copy_frame_rows(s, ref, s->stage[s->cur] + (size_t)s->nfill * s->frame_stride);
            av_frame_free(&ref);
        }
        s->nfill++;
        if (s->nfill == s->batch && (ret = submit_current(avctx)) < 0)
            return ret;
    } else {                                             /* drain (AV_CODEC_CAP_DELAY) */
        if ((ret = submit_current(avctx)) < 0)
            return ret;
        while (!s->ready_count && s->ninflight)
            if ((ret = collect_oldest(avctx)) < 0)
                return ret;
        /* first pass: once everything is coded the statistics go to stats_out, like the reference's flush call
         * (ffv1enc.c:1235-1277; freed in close like ffv1.c:229) */
        if ((avctx->flags & AV_CODEC_FLAG_PASS1) && !s->ninflight) {
            size_t need = 0;
            char probe[1];
            if (ffv1b200_enc_stats_out(s->enc, probe, 0, &need) == FFV1B200_ERR_BUFFER_TOO_SMALL && need) {
                av_freep(&avctx->stats_out);
                if (!(avctx->stats_out = av_malloc(need)))
                    return AVERROR(ENOMEM);
                if ((ret = ffv1b200_enc_stats_out(s->enc, avctx->stats_out, need, NULL)) < 0)
                    return b200_err(avctx, ret);
            }
        }
    }
    if (s->ready_count) {
        ReadyPacket *r = &s->ready[s->ready_head];
        if (s->nthreads) {
            pthread_mutex_lock(&s->lock);
            while (!r->copied)
                pthread_cond_wait(&s->job_done, &s->lock);
            pthread_mutex_unlock(&s->lock);
        }
        if (r->buf) {                                    /* made by a worker: the packet takes the buffer over */
            pkt->buf = r->buf; r->buf = NULL;
            pkt->data = pkt->buf->data; pkt->size = r->size;
        } else {
            if (r->out >= 0 && (ret = sync_output(avctx)) < 0)
                return ret;
            if ((ret = av_new_packet(pkt, r->size)) < 0)
                return ret;
            memcpy(pkt->data, r->data, r->size);
        }
        pkt->pts = pkt->dts = r->pts;
        if (r->flags & FFV1B200_PKT_FLAG_KEY)
            pkt->flags |= AV_PKT_FLAG_KEY;
        if (r->out < 0)
            av_free((void *)r->data);
        s->ready_head = (s->ready_head + 1) % s->ready_cap;
        s->ready_count--;
        *got_packet = 1;
    }
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
    { "batch", "Frames gathered per GPU submission", OFFSET(batch), AV_OPT_TYPE_INT, { .i64 = 256 }, 1, 4096, VE },
    { "gpu", "CUDA device ordinal", OFFSET(device), AV_OPT_TYPE_INT, { .i64 = 0 }, 0, 64, VE },
    { "copy_threads", "Threads moving AVFrame rows into pinned staging memory (0 = the calling thread)", OFFSET(copy_threads), AV_OPT_TYPE_INT, { .i64 = 4 }, 0, 64, VE },
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
    const AVClass *class;              /* must be first (AVOptions) */
    int batch, device;
    FFV1B200Decoder *dec;
    FFV1B200DecInfo info;
    uint8_t *framebuf;                 /* [batch] decoded pictures, tightly packed (pinned) */
    AVPacket *queue;                   /* [batch] packets gathered for the next batch (references) */
    const uint8_t **pkt_data; int *pkt_size, *key;
    uint64_t *damaged;
    int64_t *pts, *dts;
    int nqueued, nready, next_ready;
} B200DecContext;

static av_cold int b200_decode_close(AVCodecContext *avctx)
{
    B200DecContext *s = avctx->priv_data;
    int i;
    for (i = 0; s->queue && i < s->nqueued; i++)
        av_packet_unref(&s->queue[i]);
    av_freep(&s->queue); av_freep(&s->pkt_data); av_freep(&s->pkt_size); av_freep(&s->key);
    av_freep(&s->damaged); av_freep(&s->pts); av_freep(&s->dts);
    ffv1b200_host_free(s->framebuf);
    s->framebuf = NULL;
    ffv1b200_dec_close(s->dec);
    s->dec = NULL;
    return 0;
}

static int b200_dec_stream_known(AVCodecContext *avctx)
{
    B200DecContext *s = avctx->priv_data;
    ffv1b200_dec_info(s->dec, &s->info);
    if (!s->info.frame_bytes)
        return 0;
    avctx->pix_fmt = av_get_pix_fmt(s->info.pix_fmt);
    avctx->bits_per_raw_sample = s->info.bits_per_raw_sample;
    if (!s->framebuf && !(s->framebuf = ffv1b200_host_alloc((size_t)s->info.frame_bytes * s->batch, s->device)))
        return AVERROR(ENOMEM);
    return 1;
}

static av_cold int b200_decode_init(AVCodecContext *avctx)
{
    B200DecContext *s = avctx->priv_data;
    FFV1B200DecParams p;
    int ret;
    if (s->batch < 1)
        s->batch = 1;
    memset(&p, 0, sizeof(p));
    p.width = avctx->width; p.height = avctx->height;
    p.extradata = avctx->extradata; p.extradata_size = avctx->extradata_size;
    p.device = s->device;
    p.max_batch_frames = s->batch;
    if ((ret = ffv1b200_dec_open(&s->dec, &p)) < 0)
        return b200_err(avctx, ret);
    s->queue    = av_mallocz_array(s->batch, sizeof(*s->queue));
    s->pkt_data = av_mallocz_array(s->batch, sizeof(*s->pkt_data));
    s->pkt_size = av_mallocz_array(s->batch, sizeof(*s->pkt_size));
    s->key      = av_mallocz_array(s->batch, sizeof(*s->key));
    s->damaged  = av_mallocz_array(s->batch, sizeof(*s->damaged));
    s->pts      = av_mallocz_array(s->batch, sizeof(*s->pts));
    s->dts      = av_mallocz_array(s->batch, sizeof(*s->dts));
    if (!s->queue || !s->pkt_data || !s->pkt_size || !s->key || !s->damaged || !s->pts || !s->dts ||
        b200_dec_stream_known(avctx) < 0) {             /* version >= 2: everything is known from the extradata */
        b200_decode_close(avctx);
        return AVERROR(ENOMEM);
    }
    return 0;
}

/* the gathered packets -> pictures in framebuf */
static int b200_dec_run_batch(AVCodecContext *avctx)
{
    B200DecContext *s = avctx->priv_data;
    int i, ret;
    for (i = 0; i < s->nqueued; i++) {
        s->pkt_data[i] = s->queue[i].data; s->pkt_size[i] = s->queue[i].size;
        s->pts[i] = s->queue[i].pts; s->dts[i] = s->queue[i].dts;
    }
    if (!s->info.frame_bytes) {
        /* version 0/1: parameters are carried by the first keyframe (ffv1dec.c:646-696); a probe call parses them */
        uint8_t probe[16];
        ret = ffv1b200_dec_decode_host(s->dec, s->nqueued, s->pkt_data, s->pkt_size, probe, 0, NULL, NULL);
        if (ret < 0 && ret != FFV1B200_ERR_BUFFER_TOO_SMALL)
            return b200_err(avctx, ret);
        if ((ret = b200_dec_stream_known(avctx)) <= 0)
            return ret < 0 ? ret : AVERROR_INVALIDDATA;
    }
    ret = ffv1b200_dec_decode_host(s->dec, s->nqueued, s->pkt_data, s->pkt_size, s->framebuf,
                                   (size_t)s->info.frame_bytes * s->batch, s->key, s->damaged);
    for (i = 0; i < s->nqueued; i++)
        av_packet_unref(&s->queue[i]);
    if (ret < 0) {
        s->nqueued = 0;
        return b200_err(avctx, ret);
    }
    s->nready = s->nqueued; s->next_ready = 0; s->nqueued = 0;
    return 0;
}

static int b200_decode_frame(AVCodecContext *avctx, void *data, int *got_frame, AVPacket *avpkt)
{
    B200DecContext *s = avctx->priv_data;
    AVFrame *frame = data;
    const uint8_t *src[4];
    int ls[4], ret;
    *got_frame = 0;
    if (avpkt->size) {
        /* invariant: pictures not yet returned + packets queued <= batch (one picture leaves per call) */
        if ((ret = av_packet_ref(&s->queue[s->nqueued], avpkt)) < 0)
            return ret;
        s->nqueued++;
        if (s->nqueued == s->batch && s->next_ready == s->nready && (ret = b200_dec_run_batch(avctx)) < 0)
            return ret;
    } else if (s->next_ready == s->nready && s->nqueued) {
        if ((ret = b200_dec_run_batch(avctx)) < 0)       /* drain (AV_CODEC_CAP_DELAY): empty packets at the end */
            return ret;
    }
    if (s->next_ready < s->nready) {
        const int i = s->next_ready++;
        /* what ff_get_buffer() (internal) would fill in before calling the allocator (utils.c:890-960) */
        frame->width = avctx->width; frame->height = avctx->height; frame->format = avctx->pix_fmt;
        frame->sample_aspect_ratio = avctx->sample_aspect_ratio;
        if ((ret = avcodec_default_get_buffer2(avctx, frame, 0)) < 0)
            return ret;
        av_image_fill_arrays((uint8_t **)src, ls, s->framebuf + (size_t)i * s->info.frame_bytes, avctx->pix_fmt, avctx->width, avctx->height, 1);
        av_image_copy(frame->data, frame->linesize, src, ls, avctx->pix_fmt, avctx->width, avctx->height);
        frame->key_frame = s->key[i];
        frame->pict_type = AV_PICTURE_TYPE_I;            /* ffv1dec.c:913 */
        frame->pkt_pts = s->pts[i]; frame->pkt_dts = s->dts[i];
        if (s->damaged[i])
            av_log(avctx, AV_LOG_ERROR, "slice(s) damaged: mask %#llx (concealed from the previous frame)\n", (unsigned long long)s->damaged[i]);
        *got_frame = 1;
        /* a full queue that had to wait for the last picture of the previous batch */
        if (s->next_ready == s->nready && s->nqueued == s->batch && (ret = b200_dec_run_batch(avctx)) < 0)
            return ret;
    }
    return avpkt->size;                                  /* ffv1dec.c:1034 */
}

#define DOFFSET(x) offsetof(B200DecContext, x)
#define VD AV_OPT_FLAG_VIDEO_PARAM | AV_OPT_FLAG_DECODING_PARAM
static const AVOption b200_dec_options[] = {
    { "batch", "Packets gathered per GPU submission (1 = no delay)", DOFFSET(batch), AV_OPT_TYPE_INT, { .i64 = 1 }, 1, 4096, VD },
    { "gpu", "CUDA device ordinal", DOFFSET(device), AV_OPT_TYPE_INT, { .i64 = 0 }, 0, 64, VD },
    { NULL }
};

static const AVClass b200_dec_class = {
    .class_name = "ffv1_b200 decoder",
    .item_name  = av_default_item_name,
    .option     = b200_dec_options,
    .version    = LIBAVUTIL_VERSION_INT,
};

AVCodec ff_ffv1_b200_decoder = {
    .name           = "ffv1_b200",
    .long_name      = "FFmpeg video codec #1 (B200-native CUDA core)",
    .type           = AVMEDIA_TYPE_VIDEO,
    .id             = AV_CODEC_ID_FFV1,
    .priv_data_size = sizeof(B200DecContext),
    .init           = b200_decode_init,
    .close          = b200_decode_close,
    .decode         = b200_decode_frame,
    .capabilities   = AV_CODEC_CAP_DR1 | AV_CODEC_CAP_DELAY,
    .priv_class     = &b200_dec_class,
};
