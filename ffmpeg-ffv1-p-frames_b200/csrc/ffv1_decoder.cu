// ffv1_decoder.cu -- C-ABI decoder entry points (include/ffv1_b200.h) and batch orchestration.
// Replaces decode_init / decode_frame / ff_ffv1_close of the reference (ffv1dec.c:876-893, 895-1035; ffv1.c:205-243).
#include "../../include/ffv1_b200.h"
#include "ffv1_model.h"
#include "ffv1_dec_kernels.cuh"
#include "ffv1_internal.h"
#include <cuda_runtime.h>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <thread>

using namespace ffv1;

struct FFV1B200Decoder {
    Config cfg;
    int device = 0, max_batch = 64, max_slices = 1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr};
    DecDeviceTables tab{};
    DevBuf<int16_t> d_quant; DevBuf<uint8_t> d_lut, d_model_init[2];
    DevBuf<uint8_t> d_pkt, d_out, d_state, d_prev, d_frame_key;
    PinnedBuf<uint8_t> h_pkt, h_frame_key;
    DevBuf<int16_t> d_ring;
    DevBuf<uint64_t> d_pkt_off; PinnedBuf<uint64_t> h_pkt_off;
    DevBuf<uint32_t> d_slice_start, d_slice_size, d_damaged; PinnedBuf<uint32_t> h_slice_start, h_slice_size, h_damaged;
    DevBuf<int32_t> d_slice_count, d_seg_first, d_seg_set; PinnedBuf<int32_t> h_slice_count, h_seg_first, h_seg_set;
    DevBuf<uint32_t> d_init_state; PinnedBuf<uint32_t> h_init_state;
    int nsets = 0, cur_set = 0, width = 0, height = 0;
    bool configured = false;          // version 0/1 streams are configured by their first keyframe
    bool key_frame_ok = false, have_prev = false;
    bool grid_covers = true;          // the slice grid's rectangles cover every sample of every plane (see configure)
    int slice_count = 0;
    int64_t frame_bytes = 0;
    FFV1B200DecStats stats{};
};

namespace {
int dfail(int code, const std::string &msg) { set_last_error(msg); return code; }
#define CU_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return dfail(FFV1B200_ERR_EXTERNAL, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)
}

// builds the device tables and buffers once the stream parameters are known
static int configure(FFV1B200Decoder *d)
{
    const Config &c = d->cfg;
    DecDeviceTables &t = d->tab;
    d->max_slices = c.slice_count();
    d->frame_bytes = c.frame_bytes();
    t.width = c.width; t.height = c.height; t.version = c.version; t.micro_version = c.micro_version; t.ac = c.ac;
    t.colorspace = c.colorspace; t.bits = c.bits;
    t.coded_bits = c.colorspace ? (c.bits <= 8 ? 8 : c.bits) + 1 : (c.bits <= 8 ? 8 : c.bits);
    t.chroma_planes = c.chroma_planes; t.hshift = c.chroma_h_shift; t.vshift = c.chroma_v_shift;
    t.transparency = c.transparency; t.packed_at_lsb = c.packed_at_lsb; t.ya8 = c.ya8; t.ec = c.ec;
    t.rgb32 = c.colorspace && c.bits <= 8;
    t.num_h_slices = c.num_h_slices; t.num_v_slices = c.num_v_slices; t.max_slices = d->max_slices; t.plane_count = c.plane_count;
    t.ctx_count[0] = c.context_count[0]; t.ctx_count[1] = c.context_count[1];
    int64_t off = 0;
    for (int i = 0; i < 4; i++) { t.plane_off[i] = 0; t.plane_pitch[i] = 0; }
    for (int i = 0; i < c.nb_src_planes; i++) {
        int rows, rb; c.plane_dims(i, &rows, &rb);
        t.plane_off[i] = (int32_t)off; t.plane_pitch[i] = rb;
        off += (int64_t)rows * rb;
    }
    t.frame_bytes = d->frame_bytes;
    const int maxctx = std::max(c.context_count[0], c.context_count[1]);
    t.state_stride = (int64_t)maxctx * 32;                    // golomb models need 8 bytes per context: fits as well
    // a slice header may announce any rectangle of the grid: rows are sized for the frame width
    t.ring_w = ((c.width + 2 * kDecRingPad + 31) / 32) * 32;
    // planar YUV: the model of the plane context being decoded (<= 24 KB: small context model, or Golomb-Rice states)
    // and the three-row line ring of a regular slice live in shared memory
    {
        // version 3 carries both quantisation table sets and every slice header picks one per plane: room for the
        // largest set that fits (a plane coded with a larger one keeps its model in global memory)
        int64_t model = 0;
        for (int i = 0; i < 2; i++) {
            const int64_t m = (int64_t)c.context_count[i] * (c.ac == AC_GOLOMB ? 8 : 32);
            if (m <= 24 * 1024 && m > model) model = m;
        }
        t.smem_model = (!c.colorspace && model > 0) ? (int32_t)((model + 15) & ~15) : 0;
        const int sw = (c.width + c.num_h_slices - 1) / c.num_h_slices + 1;
        const int rw = ((sw + 2 * kDecRingPad + 31) / 32) * 32;
        t.smem_ring_w = (!c.colorspace && rw * kDecSmemRingBytes <= 11 * 1024) ? rw : 0;
    }

    // Subsampled planes: slice i covers [x0 >> shift, (x0 >> shift) + ceil(w / 2^shift)), which can stop short of the next
    // slice's first chroma column (4:1:0 / 4:1:1 on odd grids): such samples are never coded, the reference leaves them as
    // the frame allocator gave them, this decoder returns 0 there -- from the staged copy's memset, so streams with such
    // holes do not take the direct write into pinned memory.
    {
        auto covers = [](int size, int nsl, int shift) {
            int reach = 0;
            for (int i = 0; i < nsl; i++) {
                const int a = size * i / nsl, b = size * (i + 1) / nsl;
                if ((a >> shift) > reach) return false;
                reach = std::max(reach, (a >> shift) + (((b - a) + (1 << shift) - 1) >> shift));
            }
            return reach >= ((size + (1 << shift) - 1) >> shift);
        };
        d->grid_covers = c.colorspace || !c.chroma_planes ||
                         (covers(c.width, c.num_h_slices, c.chroma_h_shift) && covers(c.height, c.num_v_slices, c.chroma_v_shift));
    }
    CU_TRY(d->d_quant.upload(&c.quant_tables[0][0][0], 2 * 5 * 256, d->stream));
    uint8_t lut[512];
    coder_state_tables(c, lut, lut + 256);
    CU_TRY(d->d_lut.upload(lut, 512, d->stream));
    t.quant = d->d_quant.p; t.lut = d->d_lut.p;
    for (int i = 0; i < 2; i++) {
        t.model_init[i] = nullptr;
        if (!c.initial_states[i].empty() && c.ac != AC_GOLOMB) {
            CU_TRY(d->d_model_init[i].upload(c.initial_states[i].data(), c.initial_states[i].size(), d->stream));
            t.model_init[i] = d->d_model_init[i].p;
        }
    }

    const size_t F = (size_t)d->max_batch;
    d->nsets = d->max_batch + 1;
    CU_TRY(d->d_state.alloc((size_t)d->nsets * d->max_slices * 3 * t.state_stride));
    CU_TRY(d->d_ring.alloc(F * d->max_slices * 4 * 3 * t.ring_w));
    CU_TRY(d->d_out.alloc(F * d->frame_bytes));
    CU_TRY(d->d_prev.alloc((size_t)d->frame_bytes));
    CU_TRY(d->d_pkt_off.alloc(F)); CU_TRY(d->h_pkt_off.alloc(F));
    CU_TRY(d->d_frame_key.alloc(F)); CU_TRY(d->h_frame_key.alloc(F));
    CU_TRY(d->d_slice_start.alloc(F * d->max_slices)); CU_TRY(d->h_slice_start.alloc(F * d->max_slices));
    CU_TRY(d->d_slice_size.alloc(F * d->max_slices)); CU_TRY(d->h_slice_size.alloc(F * d->max_slices));
    CU_TRY(d->d_damaged.alloc(F * d->max_slices)); CU_TRY(d->h_damaged.alloc(F * d->max_slices));
    CU_TRY(d->d_slice_count.alloc(F)); CU_TRY(d->h_slice_count.alloc(F));
    CU_TRY(d->d_seg_first.alloc(F + 1)); CU_TRY(d->h_seg_first.alloc(F + 1));
    CU_TRY(d->d_seg_set.alloc(F)); CU_TRY(d->h_seg_set.alloc(F));
    CU_TRY(d->d_init_state.alloc(F * 3)); CU_TRY(d->h_init_state.alloc(F * 3));
    CU_TRY(cudaStreamSynchronize(d->stream));
    d->configured = true;
    return 0;
}

extern "C" {

int ffv1b200_dec_open(FFV1B200Decoder **out, const FFV1B200DecParams *p)
{
    if (!out || !p) return dfail(FFV1B200_ERR_EINVAL, "null argument");
    *out = nullptr;
    if (p->width <= 0 || p->height <= 0) return dfail(FFV1B200_ERR_EINVAL, "invalid picture size");
    std::unique_ptr<FFV1B200Decoder> d(new FFV1B200Decoder());
    d->width = p->width; d->height = p->height;
    const bool inband = !p->extradata || p->extradata_size <= 0;     // FFV1 version 0/1: parameters come with every keyframe
    if (!inband) {
        std::string err;
        int r = parse_extradata(p->extradata, p->extradata_size, p->width, p->height, d->cfg, err);
        if (r < 0) return dfail(r, err);
    }
    int ndev = ffv1b200_device_count();
    if (ndev < 0) return ndev;
    if (p->device < 0 || p->device >= ndev) return dfail(FFV1B200_ERR_EINVAL, "no such CUDA device");
    d->device = p->device;
    d->max_batch = p->max_batch_frames > 0 ? p->max_batch_frames : 64;
    CU_TRY(cudaSetDevice(d->device));
    CU_TRY(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    for (auto &ev : d->ev) CU_TRY(cudaEventCreate(&ev));
    if (!inband) { int r = configure(d.get()); if (r < 0) return r; }
    *out = d.release();
    return 0;
}

void ffv1b200_dec_close(FFV1B200Decoder *d)
{
    if (!d) return;
    cudaSetDevice(d->device);
    if (d->stream) { cudaStreamSynchronize(d->stream); cudaStreamDestroy(d->stream); }
    for (auto &ev : d->ev) if (ev) cudaEventDestroy(ev);
    delete d;
}

int ffv1b200_dec_info(const FFV1B200Decoder *d, FFV1B200DecInfo *i)
{
    if (!d || !i) return FFV1B200_ERR_EINVAL;
    const Config &c = d->cfg;
    memset(i, 0, sizeof(*i));
    i->version = c.version; i->micro_version = c.micro_version; i->ac = c.ac; i->colorspace = c.colorspace;
    i->bits_per_raw_sample = c.bits; i->chroma_planes = c.chroma_planes; i->chroma_h_shift = c.chroma_h_shift;
    i->chroma_v_shift = c.chroma_v_shift; i->transparency = c.transparency; i->num_h_slices = c.num_h_slices;
    i->num_v_slices = c.num_v_slices; i->ec = c.ec; i->intra = c.intra;
    strncpy(i->pix_fmt, c.pix_fmt.c_str(), sizeof(i->pix_fmt) - 1);
    i->frame_bytes = d->frame_bytes;
    return 0;
}

int ffv1b200_dec_decode_host(FFV1B200Decoder *d, int n, const uint8_t *const *pkt_data, const int *pkt_size,
                             uint8_t *out, size_t out_cap, int *key_flags, uint64_t *damaged)
{
    if (!d || !pkt_data || !pkt_size || !out) return dfail(FFV1B200_ERR_EINVAL, "null argument");
    if (n < 1 || n > d->max_batch) return dfail(FFV1B200_ERR_EINVAL, "npackets outside 1..max_batch_frames");
    CU_TRY(cudaSetDevice(d->device));
    if (!d->configured) {
        // version 0/1: the first keyframe carries the parameters (ffv1dec.c:646-696)
        PrefixState ps; std::string err;
        int r = parse_frame_prefix_v01(pkt_data[0], pkt_size[0], d->width, d->height, d->cfg, false, ps, err);
        if (r < 0) return dfail(r, err);
        if ((r = configure(d)) < 0) return r;
    }
    if (out_cap < (size_t)n * d->frame_bytes) return dfail(FFV1B200_ERR_BUFFER_TOO_SMALL, "output buffer too small (see ffv1b200_dec_info)");
    cudaStream_t s = d->stream;
    const Config &c = d->cfg;
    const int ms = d->max_slices;
    const int trailer = 3 + 5 * (c.ec ? 1 : 0);
    const bool inband = c.version < 2;

    // ---- packet staging + slice tables (ffv1dec.c:924-937, 804-813, 948-989): host work is a few byte reads per slice
    size_t total = 0;
    for (int f = 0; f < n; f++) {
        if (!pkt_data[f] || pkt_size[f] < 2) return dfail(FFV1B200_ERR_INVALIDDATA, "empty packet");
        total += ((size_t)pkt_size[f] + 15) & ~(size_t)15;
    }
    if (d->h_pkt.n < total) CU_TRY(d->h_pkt.alloc(total + total / 4));
    if (d->d_pkt.n < total) CU_TRY(d->d_pkt.alloc(total + total / 4));
    // the packets move into the pinned staging area: a large batch (2.5 GB at 2048 1080p frames) is copied by several
    // threads, each a contiguous range of packets of about the same number of bytes
    {
        size_t o = 0;
        for (int f = 0; f < n; f++) { d->h_pkt_off.p[f] = o; o += ((size_t)pkt_size[f] + 15) & ~(size_t)15; }
        const int hw = (int)std::thread::hardware_concurrency();
        const int nthr = (int)std::max<size_t>(1, std::min<size_t>({(size_t)8, (size_t)std::max(1, hw / 2), total >> 25, (size_t)n}));
        auto copy_range = [&](int a, int b) { for (int f = a; f < b; f++) memcpy(d->h_pkt.p + d->h_pkt_off.p[f], pkt_data[f], (size_t)pkt_size[f]); };
        std::vector<int> cut(nthr + 1, n);
        cut[0] = 0;
        for (int t = 1, f = 0; t < nthr; t++) {
            while (f < n && d->h_pkt_off.p[f] < total / nthr * t) f++;
            cut[t] = f;
        }
        std::vector<std::thread> pool;
        int started = 1;                                     // ranges that have a thread (the caller's thread takes the first)
        try {
            for (; started < nthr; started++) pool.emplace_back(copy_range, cut[started], cut[started + 1]);
        } catch (...) {}                                     // no more threads to be had: the caller copies the rest itself
        copy_range(cut[0], cut[1]);
        if (started < nthr) copy_range(cut[started], n);
        for (auto &th : pool) th.join();
    }
    size_t off = 0;
    int nseg = 0;
    bool kfo = d->key_frame_ok;
    int scount = d->slice_count;
    for (int f = 0; f < n; f++) {
        const uint8_t *pk = pkt_data[f];
        const long size = pkt_size[f];
        off += ((size_t)size + 15) & ~(size_t)15;             // (staged above at h_pkt_off[f])
        // keyframe bit: get_rac on a fresh state 128 (range 0xFF00 -> range1 0x7F80)
        const unsigned low = (unsigned)pk[0] << 8 | pk[1];
        const bool key = low >= 0x7F80u;
        if (inband) {
            PrefixState ps; std::string err;
            int r = parse_frame_prefix_v01(pk, (int)size, d->width, d->height, d->cfg, true, ps, err);
            if (r < 0) return dfail(r, err);
            d->h_init_state.p[f * 3 + 0] = ps.low; d->h_init_state.p[f * 3 + 1] = ps.range; d->h_init_state.p[f * 3 + 2] = ps.pos;
            if (key) { scount = 1; kfo = true; }
            else if (!kfo) return dfail(FFV1B200_ERR_INVALIDDATA, "Cannot decode non-keyframe without valid keyframe");
            d->h_frame_key.p[f] = key ? 1 : 0;
            d->h_slice_count.p[f] = 1;
            d->h_slice_start.p[f * ms] = 0;
            d->h_slice_size.p[f * ms] = (uint32_t)size;
            if (f == 0 || key) d->h_seg_first.p[nseg++] = f;
            if (key_flags) key_flags[f] = key ? 1 : 0;
            continue;
        }
        if (key) {
            const uint8_t *q = pk + size;
            int cnt = 0;
            for (; cnt < kMaxSlices && trailer <= q - pk; cnt++) {      // (the reference tests 3 < q - pk and may read before the packet)
                const long sz = ((long)q[-trailer] << 16) | ((long)q[-trailer + 1] << 8) | q[-trailer + 2];
                if (sz + trailer > q - pk) break;
                q -= sz + trailer;
            }
            if (cnt <= 0 || cnt > ms) return dfail(FFV1B200_ERR_INVALIDDATA, "invalid slice count in keyframe");
            scount = cnt;
            kfo = true;
        } else if (!kfo) {
            return dfail(FFV1B200_ERR_INVALIDDATA, "Cannot decode non-keyframe without valid keyframe");
        }
        d->h_frame_key.p[f] = key ? 1 : 0;
        d->h_slice_count.p[f] = scount;
        const uint8_t *bp = pk + size;
        for (int si = scount - 1; si >= 0; si--) {
            if (bp - pk < trailer) return dfail(FFV1B200_ERR_INVALIDDATA, "Slice pointer chain broken");
            const long v = (((long)bp[-trailer] << 16) | ((long)bp[-trailer + 1] << 8) | bp[-trailer + 2]) + trailer;
            if (bp - pk < v) return dfail(FFV1B200_ERR_INVALIDDATA, "Slice pointer chain broken");
            bp -= v;
            d->h_slice_start.p[f * ms + si] = (uint32_t)(bp - pk);
            d->h_slice_size.p[f * ms + si] = (uint32_t)v;
        }
        for (int si = scount; si < ms; si++) { d->h_slice_start.p[f * ms + si] = 0; d->h_slice_size.p[f * ms + si] = 0; }
        if (f == 0 || key) d->h_seg_first.p[nseg++] = f;
        if (key_flags) key_flags[f] = key ? 1 : 0;
    }
    d->h_seg_first.p[nseg] = n;
    for (int i = 0; i < nseg; i++) d->h_seg_set.p[i] = (d->cur_set + i) % d->nsets;

    cudaEventRecord(d->ev[0], s);
    CU_TRY(cudaMemcpyAsync(d->d_pkt.p, d->h_pkt.p, off, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(d->d_pkt_off.p, d->h_pkt_off.p, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(d->d_frame_key.p, d->h_frame_key.p, n, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(d->d_slice_start.p, d->h_slice_start.p, sizeof(uint32_t) * n * ms, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(d->d_slice_size.p, d->h_slice_size.p, sizeof(uint32_t) * n * ms, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(d->d_slice_count.p, d->h_slice_count.p, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(d->d_seg_first.p, d->h_seg_first.p, sizeof(int32_t) * (nseg + 1), cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(d->d_seg_set.p, d->h_seg_set.p, sizeof(int32_t) * nseg, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemsetAsync(d->d_damaged.p, 0, sizeof(uint32_t) * n * ms, s));
    if (inband) CU_TRY(cudaMemcpyAsync(d->d_init_state.p, d->h_init_state.p, sizeof(uint32_t) * 3 * n, cudaMemcpyHostToDevice, s));
    // An output buffer in pinned, device-mapped memory (ffv1b200_host_alloc, cudaHostAlloc, torch pin_memory) is written by
    // the kernel directly -- the pictures cross the bus while the batch is still being decoded (4.5 GB/s of posted writes
    // at the decoder's pace) instead of in a copy of their own behind it; FFV1B200_DEC_ZEROCOPY=0 keeps the staged copy
    uint8_t *zc_out = nullptr;
    {
        const char *ev = getenv("FFV1B200_DEC_ZEROCOPY");
        cudaPointerAttributes at{};
        if (!(ev && atoi(ev) == 0) && d->grid_covers && cudaPointerGetAttributes(&at, out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
            // (the last byte must belong to the same mapping)
            cudaPointerAttributes ae{};
            const size_t span = (size_t)n * d->frame_bytes;
            if (cudaPointerGetAttributes(&ae, out + span - 1) == cudaSuccess && ae.type == cudaMemoryTypeHost &&
                ae.devicePointer == (uint8_t *)at.devicePointer + span - 1)
                zc_out = (uint8_t *)at.devicePointer;
        }
        cudaGetLastError();
    }
    // samples no slice covers (chroma columns cut off by a slice edge that is not on the chroma grid) stay 0
    if (!zc_out) CU_TRY(cudaMemsetAsync(d->d_out.p, 0, (size_t)n * d->frame_bytes, s));

    DecBatch b{};
    b.nframes = n; b.nseg = nseg;
    b.seg_first = d->d_seg_first.p; b.seg_set = d->d_seg_set.p; b.frame_key = d->d_frame_key.p;
    b.pkt = d->d_pkt.p; b.pkt_off = d->d_pkt_off.p; b.slice_start = d->d_slice_start.p; b.slice_size = d->d_slice_size.p;
    b.slice_count = d->d_slice_count.p; b.out = zc_out ? zc_out : d->d_out.p; b.prev_frame = d->have_prev ? d->d_prev.p : nullptr;
    b.state = d->d_state.p; b.ring = d->d_ring.p; b.damaged = d->d_damaged.p; b.init_state = d->d_init_state.p;
    b.zero_fill = zc_out ? 1 : 0;

    cudaEventRecord(d->ev[1], s);
    launch_dec_crc(d->tab, b, s);
    launch_decode(d->tab, b, s);
    cudaEventRecord(d->ev[2], s);
    d->stats.kernel_launches += c.ec ? 2 : 1;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(d->h_damaged.p, d->d_damaged.p, sizeof(uint32_t) * n * ms, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    for (int f = 0; f < n; f++) {
        uint64_t mask = 0;
        for (int si = 0; si < d->h_slice_count.p[f]; si++)
            if (d->h_damaged.p[f * ms + si] && si < 64) mask |= (uint64_t)1 << si;
        if (damaged) damaged[f] = mask;
        if (mask) { launch_conceal(d->tab, b, f, s); d->stats.kernel_launches++; }     // in frame order: a frame may need its predecessor's fix
    }
    if (!zc_out) CU_TRY(cudaMemcpyAsync(out, d->d_out.p, (size_t)n * d->frame_bytes, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(d->d_prev.p, b.out + (size_t)(n - 1) * d->frame_bytes, (size_t)d->frame_bytes, cudaMemcpyDefault, s));
    cudaEventRecord(d->ev[3], s);
    CU_TRY(cudaStreamSynchronize(s));
    float t0; cudaEventElapsedTime(&t0, d->ev[0], d->ev[3]); d->stats.ms_total += t0;
    cudaEventElapsedTime(&t0, d->ev[1], d->ev[2]); d->stats.ms_decode_kernel += t0;
    d->stats.frames += n;
    d->stats.h2d_bytes += (int64_t)off + (int64_t)n * (13 + 8 * ms) + 8 * nseg;
    d->stats.d2h_bytes += (int64_t)n * d->frame_bytes + 4 * (int64_t)n * ms;      // (zero-copy: the same bytes, written by the kernel)
    d->cur_set = d->h_seg_set.p[nseg - 1];
    d->key_frame_ok = kfo; d->slice_count = scount; d->have_prev = true;
    return n;
}

int ffv1b200_dec_stats(const FFV1B200Decoder *d, FFV1B200DecStats *s)
{
    if (!d || !s) return FFV1B200_ERR_EINVAL;
    *s = d->stats;
    return 0;
}

} // extern "C"
