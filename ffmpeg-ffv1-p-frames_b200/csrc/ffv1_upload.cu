// ffv1_upload.cu -- on-GPU input preparation: what the reference's filter graph does in front of a
// hardware encoder with vf_hwupload_cuda (libavfilter/vf_hwupload_cuda.c:144: system-memory AVFrame -> AV_PIX_FMT_CUDA frame
// from a pool) and with the pixel-format conversion half of vf_scale_npp (libavfilter/vf_scale_npp.c; no scaling here).
// Frames arrive in a capture / hardware-decoder layout (NV12, P010, packed 4:2:2, 24-bit RGB) or already in a layout
// the encoder takes, and leave as device frames in one of the encoder's pix_fmts, ready for ffv1b200_enc_encode_cuda /
// _encode_device.  The conversions are pure re-arrangements of samples (plus P010's shift): every output sample is an
// input sample, so the lossless property of the codec extends to the source layout.
#include "../../include/ffv1_b200.h"
#include "ffv1_model.h"
#include "ffv1_internal.h"
#include <cuda_runtime.h>
#include <cstring>
#include <string>
#include <memory>

using namespace ffv1;

namespace {

enum ConvKind { CV_COPY = 0, CV_NV12, CV_P010, CV_YUYV, CV_UYVY, CV_RGB24, CV_BGR24, CV_RGBA };

struct ConvArgs {
    const uint8_t *src[4];
    uint8_t *dst[4];
    int32_t sls[4], dls[4];
    int32_t w, h, kind;
    long long src_stride, dst_stride;      // bytes between consecutive frames
};

// one thread per pixel pair (4:2:x layouts) or per pixel (RGB); blockIdx.z = frame
__global__ void __launch_bounds__(256) k_convert(const ConvArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    const uint8_t *s0 = a.src[0] + (size_t)f * a.src_stride, *s1 = a.src[1] ? a.src[1] + (size_t)f * a.src_stride : nullptr;
    uint8_t *d0 = a.dst[0] + (size_t)f * a.dst_stride;
    uint8_t *d1 = a.dst[1] ? a.dst[1] + (size_t)f * a.dst_stride : nullptr;
    uint8_t *d2 = a.dst[2] ? a.dst[2] + (size_t)f * a.dst_stride : nullptr;
    const int w = a.w, h = a.h;
    switch (a.kind) {
    case CV_NV12: {                              // Y plane + interleaved UV plane -> yuv420p; x = chroma column
        const int cw = (w + 1) >> 1, ch = (h + 1) >> 1;
        if (x >= cw) return;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int px = 2 * x + k;
            if (px < w) d0[(size_t)y * a.dls[0] + px] = s0[(size_t)y * a.sls[0] + px];
        }
        if (y < ch) {
            const uchar2 uv = *reinterpret_cast<const uchar2 *>(s1 + (size_t)y * a.sls[1] + 2 * x);
            d1[(size_t)y * a.dls[1] + x] = uv.x;
            d2[(size_t)y * a.dls[2] + x] = uv.y;
        }
        break;
    }
    case CV_P010: {                              // 10 significant bits at the top of 16 -> yuv420p10le (LSB aligned)
        const int cw = (w + 1) >> 1, ch = (h + 1) >> 1;
        if (x >= cw) return;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int px = 2 * x + k;
            if (px < w)
                reinterpret_cast<uint16_t *>(d0 + (size_t)y * a.dls[0])[px] = reinterpret_cast<const uint16_t *>(s0 + (size_t)y * a.sls[0])[px] >> 6;
        }
        if (y < ch) {
            const ushort2 uv = *reinterpret_cast<const ushort2 *>(s1 + (size_t)y * a.sls[1] + 4 * x);
            reinterpret_cast<uint16_t *>(d1 + (size_t)y * a.dls[1])[x] = uv.x >> 6;
            reinterpret_cast<uint16_t *>(d2 + (size_t)y * a.dls[2])[x] = uv.y >> 6;
        }
        break;
    }
    case CV_YUYV: case CV_UYVY: {                // packed 4:2:2 -> yuv422p; x = chroma column
        const int cw = (w + 1) >> 1;
        if (x >= cw) return;
        const uchar4 q = *reinterpret_cast<const uchar4 *>(s0 + (size_t)y * a.sls[0] + 4 * x);
        const uint8_t y0 = a.kind == CV_YUYV ? q.x : q.y, u = a.kind == CV_YUYV ? q.y : q.x;
        const uint8_t y1 = a.kind == CV_YUYV ? q.z : q.w, v = a.kind == CV_YUYV ? q.w : q.z;
        d0[(size_t)y * a.dls[0] + 2 * x] = y0;
        if (2 * x + 1 < w) d0[(size_t)y * a.dls[0] + 2 * x + 1] = y1;
        d1[(size_t)y * a.dls[1] + x] = u;
        d2[(size_t)y * a.dls[2] + x] = v;
        break;
    }
    case CV_RGB24: case CV_BGR24: {              // 3 bytes per pixel -> bgr0 (B, G, R, 0 in memory)
        if (x >= w) return;
        const uint8_t *p = s0 + (size_t)y * a.sls[0] + 3 * x;
        const uint32_t c0 = p[0], c1 = p[1], c2 = p[2];
        const uint32_t b = a.kind == CV_RGB24 ? c2 : c0, r = a.kind == CV_RGB24 ? c0 : c2;
        reinterpret_cast<uint32_t *>(d0 + (size_t)y * a.dls[0])[x] = b | c1 << 8 | r << 16;
        break;
    }
    case CV_RGBA: {                              // R, G, B, A -> bgra
        if (x >= w) return;
        const uint32_t v = reinterpret_cast<const uint32_t *>(s0 + (size_t)y * a.sls[0])[x];
        reinterpret_cast<uint32_t *>(d0 + (size_t)y * a.dls[0])[x] = (v & 0xFF00FF00u) | ((v & 0xFFu) << 16) | ((v >> 16) & 0xFFu);
        break;
    }
    default: break;
    }
}

int ufail(int code, const std::string &msg) { set_last_error(msg); return code; }
#define CU_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return ufail(FFV1B200_ERR_EXTERNAL, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)

struct PlaneGeom { int rows[4], rowbytes[4], n; };

// source layouts that are not encoder formats
bool source_geometry(const std::string &fmt, int w, int h, PlaneGeom &g, int &kind, std::string &dst_fmt)
{
    const int cw = (w + 1) / 2, ch = (h + 1) / 2;
    memset(&g, 0, sizeof(g));
    if (fmt == "nv12")    { g.n = 2; g.rows[0] = h; g.rowbytes[0] = w; g.rows[1] = ch; g.rowbytes[1] = 2 * cw; kind = CV_NV12; dst_fmt = "yuv420p"; return true; }
    if (fmt == "p010le")  { g.n = 2; g.rows[0] = h; g.rowbytes[0] = 2 * w; g.rows[1] = ch; g.rowbytes[1] = 4 * cw; kind = CV_P010; dst_fmt = "yuv420p10le"; return true; }
    if (fmt == "yuyv422") { g.n = 1; g.rows[0] = h; g.rowbytes[0] = 4 * cw; kind = CV_YUYV; dst_fmt = "yuv422p"; return true; }
    if (fmt == "uyvy422") { g.n = 1; g.rows[0] = h; g.rowbytes[0] = 4 * cw; kind = CV_UYVY; dst_fmt = "yuv422p"; return true; }
    if (fmt == "rgb24")   { g.n = 1; g.rows[0] = h; g.rowbytes[0] = 3 * w; kind = CV_RGB24; dst_fmt = "bgr0"; return true; }
    if (fmt == "bgr24")   { g.n = 1; g.rows[0] = h; g.rowbytes[0] = 3 * w; kind = CV_BGR24; dst_fmt = "bgr0"; return true; }
    if (fmt == "rgba")    { g.n = 1; g.rows[0] = h; g.rowbytes[0] = 4 * w; kind = CV_RGBA; dst_fmt = "bgra"; return true; }
    return false;
}

} // namespace

struct FFV1B200Uploader {
    int device = 0, w = 0, h = 0, pool = 0, kind = CV_COPY;
    PlaneGeom src{}, dst{};
    int spitch[4] = {0, 0, 0, 0}, dpitch[4] = {0, 0, 0, 0};
    size_t soff[4] = {0, 0, 0, 0}, doff[4] = {0, 0, 0, 0}, sstride = 0, dstride = 0;
    DevBuf<uint8_t> d_src, d_dst;
    cudaStream_t stream = nullptr;
    std::string dst_fmt;
};

extern "C" {

int ffv1b200_upload_open(FFV1B200Uploader **out, const char *src_pix_fmt, const char *dst_pix_fmt, int width, int height,
                         int pool_frames, int device)
{
    if (!out || !src_pix_fmt) return ufail(FFV1B200_ERR_EINVAL, "null argument");
    *out = nullptr;
    if (width <= 0 || height <= 0 || pool_frames < 1) return ufail(FFV1B200_ERR_EINVAL, "invalid size");
    std::unique_ptr<FFV1B200Uploader> u(new FFV1B200Uploader());
    u->w = width; u->h = height; u->pool = pool_frames;
    std::string want = dst_pix_fmt ? dst_pix_fmt : "";
    if (!source_geometry(src_pix_fmt, width, height, u->src, u->kind, u->dst_fmt)) {
        u->kind = CV_COPY; u->dst_fmt = src_pix_fmt;                 // a layout the encoder takes as it is: upload only
    }
    if (!want.empty() && want != u->dst_fmt)
        return ufail(FFV1B200_ERR_ENOSYS, std::string("no conversion from ") + src_pix_fmt + " to " + want + " (it would be " + u->dst_fmt + ")");
    // geometry of the destination: the encoder's own view of the format (ffv1enc.c:720-820)
    Config c;
    std::string err;
    EncOptions o{width, height, u->dst_fmt, 1, -1, 0, 0, 0, -1};
    int r = resolve_encoder(o, c, err);
    if (r < 0) return ufail(r, err);
    u->dst.n = c.nb_src_planes;
    for (int i = 0; i < c.nb_src_planes; i++) c.plane_dims(i, &u->dst.rows[i], &u->dst.rowbytes[i]);
    if (u->kind == CV_COPY) u->src = u->dst;
    int ndev = ffv1b200_device_count();
    if (ndev < 0) return ndev;
    if (device < 0 || device >= ndev) return ufail(FFV1B200_ERR_EINVAL, "no such CUDA device");
    u->device = device;
    CU_TRY(cudaSetDevice(device));
    CU_TRY(cudaStreamCreateWithFlags(&u->stream, cudaStreamNonBlocking));
    size_t off = 0;
    for (int i = 0; i < u->dst.n; i++) {                             // 128-byte pitches: the TMA-staged per-pixel kernel applies
        u->dpitch[i] = (u->dst.rowbytes[i] + 127) & ~127;
        u->doff[i] = off; off += (size_t)u->dpitch[i] * u->dst.rows[i];
    }
    u->dstride = (off + 255) & ~(size_t)255;
    CU_TRY(u->d_dst.alloc(u->dstride * (size_t)pool_frames));
    if (u->kind != CV_COPY) {
        off = 0;
        for (int i = 0; i < u->src.n; i++) {
            u->spitch[i] = (u->src.rowbytes[i] + 127) & ~127;
            u->soff[i] = off; off += (size_t)u->spitch[i] * u->src.rows[i];
        }
        u->sstride = (off + 255) & ~(size_t)255;
        CU_TRY(u->d_src.alloc(u->sstride * (size_t)pool_frames));
    }
    *out = u.release();
    return 0;
}

void ffv1b200_upload_close(FFV1B200Uploader *u)
{
    if (!u) return;
    cudaSetDevice(u->device);
    if (u->stream) { cudaStreamSynchronize(u->stream); cudaStreamDestroy(u->stream); }
    delete u;
}

const char *ffv1b200_upload_pix_fmt(const FFV1B200Uploader *u) { return u ? u->dst_fmt.c_str() : nullptr; }

static int run_convert(FFV1B200Uploader *u, int n, const uint8_t *const *sp, const int *sls, long long sstride, cudaStream_t s)
{
    ConvArgs a{};
    for (int i = 0; i < 4; i++) {
        a.src[i] = i < u->src.n ? sp[i] : nullptr; a.sls[i] = i < u->src.n ? sls[i] : 0;
        a.dst[i] = i < u->dst.n ? u->d_dst.p + u->doff[i] : nullptr; a.dls[i] = u->dpitch[i];
    }
    a.w = u->w; a.h = u->h; a.kind = u->kind; a.src_stride = sstride; a.dst_stride = (long long)u->dstride;
    const int cols = (u->kind == CV_RGB24 || u->kind == CV_BGR24 || u->kind == CV_RGBA) ? u->w : (u->w + 1) / 2;
    dim3 grid((cols + 255) / 256, u->h, n);
    k_convert<<<grid, 256, 0, s>>>(a);
    CU_TRY(cudaGetLastError());
    return 0;
}

static void fill_out(const FFV1B200Uploader *u, int n, void **d_planes, int *d_linesizes)
{
    for (int f = 0; f < n; f++)
        for (int i = 0; i < 4; i++)
            d_planes[f * 4 + i] = i < u->dst.n ? (void *)(u->d_dst.p + (size_t)f * u->dstride + u->doff[i]) : nullptr;
    for (int i = 0; i < 4; i++) d_linesizes[i] = i < u->dst.n ? u->dpitch[i] : 0;
}

int ffv1b200_upload_frames(FFV1B200Uploader *u, int n, const uint8_t *const *planes, const int *linesizes,
                           void **d_planes, int *d_linesizes)
{
    if (!u || !planes || !linesizes || !d_planes || !d_linesizes) return ufail(FFV1B200_ERR_EINVAL, "null argument");
    if (n < 1 || n > u->pool) return ufail(FFV1B200_ERR_EINVAL, "nframes outside 1..pool_frames");
    CU_TRY(cudaSetDevice(u->device));
    cudaStream_t s = u->stream;
    const bool copy = u->kind == CV_COPY;
    for (int f = 0; f < n; f++)
        for (int i = 0; i < u->src.n; i++) {
            const uint8_t *src = planes[f * 4 + i];
            if (!src) return ufail(FFV1B200_ERR_EINVAL, "missing plane pointer");
            if (linesizes[f * 4 + i] < u->src.rowbytes[i]) return ufail(FFV1B200_ERR_EINVAL, "linesize smaller than a row");
            uint8_t *dst = copy ? u->d_dst.p + (size_t)f * u->dstride + u->doff[i] : u->d_src.p + (size_t)f * u->sstride + u->soff[i];
            CU_TRY(cudaMemcpy2DAsync(dst, copy ? u->dpitch[i] : u->spitch[i], src, linesizes[f * 4 + i], u->src.rowbytes[i], u->src.rows[i],
                                     cudaMemcpyHostToDevice, s));
        }
    if (!copy) {
        const uint8_t *sp[4]; int sl[4];
        for (int i = 0; i < 4; i++) { sp[i] = i < u->src.n ? u->d_src.p + u->soff[i] : nullptr; sl[i] = u->spitch[i]; }
        int r = run_convert(u, n, sp, sl, (long long)u->sstride, s);
        if (r < 0) return r;
    }
    CU_TRY(cudaStreamSynchronize(s));
    fill_out(u, n, d_planes, d_linesizes);
    return n;
}

int ffv1b200_convert_device(FFV1B200Uploader *u, int n, const void *const *d_src_planes, const int *src_linesizes, long long src_frame_stride,
                            void **d_planes, int *d_linesizes, void *stream)
{
    if (!u || !d_src_planes || !src_linesizes || !d_planes || !d_linesizes) return ufail(FFV1B200_ERR_EINVAL, "null argument");
    if (n < 1 || n > u->pool) return ufail(FFV1B200_ERR_EINVAL, "nframes outside 1..pool_frames");
    if (n > 1 && src_frame_stride <= 0) return ufail(FFV1B200_ERR_EINVAL, "frames of a batch must lie at a constant distance");
    CU_TRY(cudaSetDevice(u->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : u->stream;
    if (u->kind == CV_COPY) {
        for (int f = 0; f < n; f++)
            for (int i = 0; i < u->src.n; i++)
                CU_TRY(cudaMemcpy2DAsync(u->d_dst.p + (size_t)f * u->dstride + u->doff[i], u->dpitch[i],
                                         (const uint8_t *)d_src_planes[i] + (size_t)f * src_frame_stride, src_linesizes[i],
                                         u->src.rowbytes[i], u->src.rows[i], cudaMemcpyDeviceToDevice, s));
    } else {
        const uint8_t *sp[4]; int sl[4];
        for (int i = 0; i < 4; i++) { sp[i] = i < u->src.n ? (const uint8_t *)d_src_planes[i] : nullptr; sl[i] = i < u->src.n ? src_linesizes[i] : 0; }
        int r = run_convert(u, n, sp, sl, src_frame_stride, s);
        if (r < 0) return r;
    }
    if (!stream) CU_TRY(cudaStreamSynchronize(s));
    fill_out(u, n, d_planes, d_linesizes);
    return n;
}

} // extern "C"
