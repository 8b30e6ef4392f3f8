// ffv1_encoder.cu -- C-ABI encoder entry points (include/ffv1_b200.h) and batch orchestration.
// Replaces encode_init / encode_frame / encode_close of the reference (ffv1enc.c:669-1029, 1222-1373, 1375-1379).
#include "../../include/ffv1_b200.h"
#include "ffv1_model.h"
#include "ffv1_enc_kernels.cuh"
#include "ffv1_internal.h"
#include <cuda_runtime.h>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <string>
#include <vector>
#include <algorithm>

using namespace ffv1;

struct FFV1B200Encoder {
    Config cfg;
    Tables tab;
    std::vector<uint8_t> extradata;
    int device = 0, max_batch = 64;
    int64_t picture_number = 0;
    FFV1B200FrameProps props{0, 1, 3};
    bool prefix_dirty = true;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {nullptr};

    // static device tables
    DevBuf<SliceGeom> d_slices; DevBuf<LineDesc> d_lines; DevBuf<int32_t> d_pc_lines; DevBuf<TileDesc> d_tiles;
    DevBuf<int16_t> d_quant; DevBuf<uint8_t> d_lut; DevBuf<uint16_t> d_prefix; DevBuf<int32_t> d_prefix_len;
    DevBuf<uint8_t> d_gprefix; DevBuf<int32_t> d_gprefix_len;
    // batch buffers
    DevBuf<uint8_t> d_in; size_t in_plane_off[4] = {0,0,0,0}; int in_pitch[4] = {0,0,0,0}; size_t in_frame_stride = 0;
    DevBuf<const uint8_t *> d_planes; PinnedBuf<const uint8_t *> h_planes;
    DevBuf<uint32_t> d_rec, d_run_cnt, d_slice_bytes, d_pkt_size;
    DevBuf<uint64_t> d_pkt_off;
    DevBuf<uint8_t> d_one_pow, d_run_pc;
    DevBuf<CtxTile> d_ctiles;
    DevBuf<int32_t> d_frame_seg; PinnedBuf<int32_t> h_frame_seg;
    DevBuf<uint32_t> d_line_pos, d_ctx_hist, d_list_start, d_list_count;
    DevBuf<uint16_t> d_list_order;
    DevBuf<uint2> d_lists;
    bool ctx_replay = false;          // context-decomposed replay (small context model) instead of one warp per chain
    DevBuf<uint16_t> d_dec;
    DevBuf<int32_t> d_seg_first; DevBuf<uint8_t> d_frame_key;
    PinnedBuf<int32_t> h_seg_first; PinnedBuf<uint8_t> h_frame_key;
    DevBuf<uint8_t> d_scratch, d_out, d_state_seg, d_carry[2];
    int carry_idx = 0;
    DevBuf<unsigned long long> d_status; PinnedBuf<unsigned long long> h_status;
    PinnedBuf<uint32_t> h_pkt_size; PinnedBuf<uint64_t> h_pkt_off;
    double dec_per_sample = 5.0;      // sizing of the decision stream (entries per sample), grows on demand
    double scratch_scale = 1.0;
    bool state_in_smem = true;
    bool fast_pixel = false;          // geometry allows the TMA-staged per-pixel kernel (pointer alignment is checked per call)
    int max_plane_width = 0, num_sms = 148;
    FFV1B200EncStats stats{};
    int last_nframes = 0;
};

namespace {

int fail(int code, const std::string &msg) { set_last_error(msg); return code; }

#define CU_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return fail(FFV1B200_ERR_EXTERNAL, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)

int upload_prefixes(FFV1B200Encoder *e)
{
    const int ns = e->cfg.slice_count();
    if (e->tab.layout.golomb) {
        std::vector<uint8_t> gp((size_t)ns * 2 * kMaxGolombPrefix, 0);
        std::vector<int32_t> gl((size_t)ns * 2, 0);
        for (int s = 0; s < ns; s++)
            for (int key = 0; key < 2; key++) {
                std::vector<uint8_t> b = slice_prefix_bytes(e->cfg, s, key != 0, e->props.sar_num, e->props.sar_den, e->props.picture_structure);
                if ((int)b.size() > kMaxGolombPrefix) return fail(FFV1B200_ERR_EINVAL, "slice header too long");
                std::copy(b.begin(), b.end(), gp.begin() + (size_t)(s * 2 + key) * kMaxGolombPrefix);
                gl[s * 2 + key] = (int32_t)b.size();
            }
        CU_TRY(e->d_gprefix.upload(gp.data(), gp.size(), e->stream));
        CU_TRY(e->d_gprefix_len.upload(gl.data(), gl.size(), e->stream));
        CU_TRY(cudaStreamSynchronize(e->stream));
        e->prefix_dirty = false;
        return 0;
    }
    std::vector<uint16_t> pre((size_t)(ns * 2 + 1) * kMaxPrefix, 0);
    pre[(size_t)ns * 2 * kMaxPrefix] = 129;                              // the decision that closes every slice (state 129, bit 0)
    std::vector<int32_t> len((size_t)ns * 2, 0);
    for (int s = 0; s < ns; s++)
        for (int key = 0; key < 2; key++) {
            std::vector<uint16_t> d = slice_prefix_decisions(e->cfg, s, key != 0, e->props.sar_num, e->props.sar_den, e->props.picture_structure);
            if ((int)d.size() > kMaxPrefix) return fail(FFV1B200_ERR_EINVAL, "slice header too long");
            std::copy(d.begin(), d.end(), pre.begin() + (size_t)(s * 2 + key) * kMaxPrefix);
            len[s * 2 + key] = (int32_t)d.size();
        }
    CU_TRY(e->d_prefix.upload(pre.data(), pre.size(), e->stream));
    CU_TRY(e->d_prefix_len.upload(len.data(), len.size(), e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    e->prefix_dirty = false;
    return 0;
}

int alloc_batch_buffers(FFV1B200Encoder *e)
{
    const Layout &L = e->tab.layout;
    const size_t F = (size_t)e->max_batch;
    // staged input frames: our own pitch (128 B multiples) per plane
    size_t off = 0;
    for (int i = 0; i < e->cfg.nb_src_planes; i++) {
        int rows, rb; e->cfg.plane_dims(i, &rows, &rb);
        e->in_pitch[i] = (rb + 127) & ~127;
        e->in_plane_off[i] = off;
        off += (size_t)e->in_pitch[i] * rows;
    }
    e->in_frame_stride = (off + 255) & ~(size_t)255;
    CU_TRY(e->d_in.alloc(e->in_frame_stride * F));
    CU_TRY(e->d_planes.alloc(F * 4)); CU_TRY(e->h_planes.alloc(F * 4));
    CU_TRY(e->d_rec.alloc((size_t)L.rec_per_frame * F));
    CU_TRY(e->d_run_cnt.alloc((size_t)L.runs_per_frame * F));
    CU_TRY(e->d_slice_bytes.alloc((size_t)L.nslices * F));
    CU_TRY(e->d_pkt_size.alloc(F)); CU_TRY(e->d_pkt_off.alloc(F + 1));
    CU_TRY(e->h_pkt_size.alloc(F)); CU_TRY(e->h_pkt_off.alloc(F + 1));
    CU_TRY(e->d_seg_first.alloc(F + 1)); CU_TRY(e->h_seg_first.alloc(F + 1));
    CU_TRY(e->d_frame_key.alloc(F)); CU_TRY(e->h_frame_key.alloc(F));
    CU_TRY(e->d_status.alloc(8)); CU_TRY(e->h_status.alloc(8));
    CU_TRY(e->d_scratch.alloc((size_t)L.scratch_per_frame * F));
    if (!L.golomb) CU_TRY(e->d_dec.alloc((size_t)L.dec_per_frame * F + 64));
    const size_t state_bytes = (size_t)L.nslices * L.npc * L.ctx_count * 32;
    for (int k = 0; k < 2; k++) {
        CU_TRY(e->d_carry[k].alloc(state_bytes));
        CU_TRY(cudaMemsetAsync(e->d_carry[k].p, 128, state_bytes, e->stream));
    }
    const size_t nseg_max = (F + (e->cfg.gop_size > 0 ? e->cfg.gop_size : 1) - 1) / (e->cfg.gop_size > 0 ? e->cfg.gop_size : 1) + 1;
    if (e->ctx_replay) {
        const size_t nchains = nseg_max * L.nslices * L.npc;
        CU_TRY(e->d_frame_seg.alloc(F)); CU_TRY(e->h_frame_seg.alloc(F));
        CU_TRY(e->d_line_pos.alloc((size_t)L.lines_per_frame * F));
        CU_TRY(e->d_ctx_hist.alloc((size_t)L.ctiles_per_frame * L.ctx_count * F));
        CU_TRY(e->d_list_start.alloc(nchains * L.ctx_count));
        CU_TRY(e->d_list_count.alloc(nchains * L.ctx_count));
        CU_TRY(e->d_list_order.alloc(nchains * L.ctx_count));
        CU_TRY(e->d_lists.alloc((size_t)L.samples_per_frame * F));
    }
    if ((!e->state_in_smem && !e->ctx_replay) || L.golomb) {
        const int g = e->cfg.gop_size > 0 ? e->cfg.gop_size : 1;
        CU_TRY(e->d_state_seg.alloc(state_bytes * ((F + g - 1) / g + 1)));               // one state set per GOP segment of a batch
    }
    return 0;
}

EncDeviceTables device_tables(FFV1B200Encoder *e)
{
    EncDeviceTables t;
    t.layout = e->tab.layout;
    t.slices = e->d_slices.p; t.lines = e->d_lines.p; t.pc_lines = e->d_pc_lines.p; t.tiles = e->d_tiles.p;
    t.ctiles = e->d_ctiles.p;
    t.quant = e->d_quant.p; t.trans_lut = e->d_lut.p; t.one_pow = e->d_one_pow.p; t.run_pc = e->d_run_pc.p; t.gprefix = e->d_gprefix.p; t.gprefix_len = e->d_gprefix_len.p; t.prefix = e->d_prefix.p; t.prefix_len = e->d_prefix_len.p;
    t.ec = e->cfg.ec; t.version = e->cfg.version; t.state_in_smem = e->state_in_smem ? 1 : 0;
    return t;
}

// runs the whole device pipeline for frames whose plane pointers are already in d_planes
int run_pipeline(FFV1B200Encoder *e, int nframes, const int linesizes[4], uint8_t *d_out, size_t d_out_cap, cudaStream_t s)
{
    const Layout &L = e->tab.layout;
    if (e->prefix_dirty) { int r = upload_prefixes(e); if (r < 0) return r; }

    // GOP segments of this batch (keyframe rule: ffv1enc.c:1299)
    int nseg = 0;
    for (int f = 0; f < nframes; f++) {
        const int64_t pn = e->picture_number + f;
        const bool key = e->cfg.gop_size == 0 || (pn % e->cfg.gop_size) == 0;
        e->h_frame_key.p[f] = key ? 1 : 0;
        if (f == 0 || key) e->h_seg_first.p[nseg++] = f;
        if (e->ctx_replay) e->h_frame_seg.p[f] = nseg - 1;
    }
    e->h_seg_first.p[nseg] = nframes;

    for (int attempt = 0; attempt < 6; attempt++) {
        CU_TRY(cudaMemcpyAsync(e->d_seg_first.p, e->h_seg_first.p, sizeof(int32_t) * (nseg + 1), cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(e->d_frame_key.p, e->h_frame_key.p, nframes, cudaMemcpyHostToDevice, s));
        if (e->ctx_replay) CU_TRY(cudaMemcpyAsync(e->d_frame_seg.p, e->h_frame_seg.p, sizeof(int32_t) * nframes, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemsetAsync(e->d_status.p, 0, sizeof(unsigned long long) * 8, s));

        EncDeviceTables t = device_tables(e);
        EncBatch b{};
        b.nframes = nframes; b.nseg = nseg;
        b.planes = e->d_planes.p;
        for (int i = 0; i < 4; i++) b.linesize[i] = linesizes[i];
        b.rec = e->d_rec.p; b.run_cnt = e->d_run_cnt.p; b.dec = e->d_dec.p;
        b.seg_first = e->d_seg_first.p; b.frame_key = e->d_frame_key.p;
        b.scratch = e->d_scratch.p; b.slice_bytes = e->d_slice_bytes.p;
        b.pkt_size = e->d_pkt_size.p; b.pkt_off = e->d_pkt_off.p;
        b.out = d_out; b.out_capacity = d_out_cap;
        b.state_seg = e->d_state_seg.p;
        b.carry_in = e->d_carry[e->carry_idx].p; b.carry_out = e->d_carry[e->carry_idx ^ 1].p;
        b.status = e->d_status.p;
        b.frame_seg = e->d_frame_seg.p; b.line_pos = e->d_line_pos.p; b.ctx_hist = e->d_ctx_hist.p;
        b.list_start = e->d_list_start.p; b.list_count = e->d_list_count.p; b.list_order = e->d_list_order.p; b.lists = e->d_lists.p;

        bool fast = e->fast_pixel;
        for (int i = 0; i < 4 && fast; i++) fast = (linesizes[i] & 15) == 0;
        for (int i = 0; i < nframes * 4 && fast; i++) fast = (reinterpret_cast<uintptr_t>(e->h_planes.p[i]) & 15) == 0;
        cudaEventRecord(e->ev[0], s);
        if (fast) launch_pixel_fast(t, b, e->max_plane_width, e->num_sms, s);
        else      launch_pixel(t, b, s);
        cudaEventRecord(e->ev[1], s);
        if (e->ctx_replay) launch_ctx_replay(t, b, s);
        else if (!L.golomb) launch_replay(t, b, s);
        cudaEventRecord(e->ev[2], s);
        if (!L.golomb) launch_rangecode(t, b, s); else launch_golomb(t, b, s);
        cudaEventRecord(e->ev[3], s);
        launch_pack(t, b, s);
        cudaEventRecord(e->ev[4], s);
        e->stats.kernel_launches += L.golomb ? 4 : (e->ctx_replay ? 9 : 5);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(e->h_status.p, e->d_status.p, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(e->h_pkt_size.p, e->d_pkt_size.p, sizeof(uint32_t) * nframes, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(e->h_pkt_off.p, e->d_pkt_off.p, sizeof(uint64_t) * (nframes + 1), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        e->stats.d2h_bytes += 64 + 12 * (int64_t)nframes + 8;

        const unsigned long long *st = e->h_status.p;
        if (st[0]) {            // a decision region overflowed: st[0] = needed entries per sample * 256; grow and re-run
            e->dec_per_sample = std::max(e->dec_per_sample * 1.25, (double)st[0] / 256.0 * 1.1);
            layout_decisions(e->tab, e->dec_per_sample);
            e->d_dec.release();
            CU_TRY(e->d_dec.alloc((size_t)e->tab.layout.dec_per_frame * e->max_batch + 64));
            CU_TRY(e->d_slices.upload(e->tab.slices.data(), e->tab.slices.size(), s));
            e->stats.retries++;
            continue;
        }
        if (st[1]) {            // a slice's coder output overflowed its scratch region: rebuild with larger regions
            double need = 1.0;
            for (auto &g : e->tab.slices) need = std::max(need, (double)st[1] / (double)g.scratch_cap);
            e->scratch_scale *= need * 1.25;
            uint32_t cur = 0;
            for (auto &g : e->tab.slices) {
                g.scratch_cap = (uint32_t)((((uint64_t)(g.scratch_cap * need * 1.25)) + 255) & ~255ull);
                g.scratch_off = cur; cur += g.scratch_cap;
            }
            e->tab.layout.scratch_per_frame = cur;
            e->d_scratch.release();
            CU_TRY(e->d_scratch.alloc((size_t)cur * e->max_batch));
            CU_TRY(e->d_slices.upload(e->tab.slices.data(), e->tab.slices.size(), s));
            e->stats.retries++;
            continue;
        }
        if (st[2]) {
            set_last_error("output buffer too small: need " + std::to_string(st[2]) + " bytes");
            return FFV1B200_ERR_BUFFER_TOO_SMALL;
        }
        float ms;
        cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]); e->stats.ms_pixel_kernel += ms;
        cudaEventElapsedTime(&ms, e->ev[1], e->ev[2]); e->stats.ms_model_kernel += ms;
        cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]); e->stats.ms_coder_kernel += ms;
        cudaEventElapsedTime(&ms, e->ev[3], e->ev[4]); e->stats.ms_pack_kernel += ms;
        e->stats.decisions += (int64_t)st[3];
        return 0;
    }
    return fail(FFV1B200_ERR_EXTERNAL, "scratch buffers kept overflowing");
}

void finish_batch(FFV1B200Encoder *e, int nframes, FFV1B200Packet *pkts)
{
    for (int f = 0; f < nframes; f++) {
        pkts[f].offset = (int64_t)e->h_pkt_off.p[f];
        pkts[f].size = (int32_t)e->h_pkt_size.p[f];
        pkts[f].flags = e->h_frame_key.p[f] ? FFV1B200_PKT_FLAG_KEY : 0;
        pkts[f].picture_number = e->picture_number + f;
    }
    e->stats.frames += nframes;
    e->stats.packet_bytes += (int64_t)e->h_pkt_off.p[nframes];
    uint64_t samples = 0;
    for (auto &g : e->tab.slices) samples += g.nsamples;
    e->stats.samples += (int64_t)samples * nframes;
    e->picture_number += nframes;
    e->carry_idx ^= 1;
    e->last_nframes = nframes;
}

} // namespace

extern "C" {

int ffv1b200_enc_open(FFV1B200Encoder **out, const FFV1B200EncParams *p)
{
    if (!out || !p || !p->pix_fmt) return fail(FFV1B200_ERR_EINVAL, "null argument");
    *out = nullptr;
    EncOptions o{p->width, p->height, p->pix_fmt, p->gop_size, p->level, p->slices, p->coder, p->context, p->slicecrc};
    std::unique_ptr<FFV1B200Encoder> e(new FFV1B200Encoder());
    std::string err;
    int r = resolve_encoder(o, e->cfg, err);
    if (r < 0) return fail(r, err);
    int ndev = ffv1b200_device_count();
    if (ndev < 0) return ndev;
    if (p->device < 0 || p->device >= ndev) return fail(FFV1B200_ERR_EINVAL, "no such CUDA device");
    e->device = p->device;
    e->max_batch = p->max_batch_frames > 0 ? p->max_batch_frames : 64;
    e->picture_number = p->first_picture_number;
    CU_TRY(cudaSetDevice(e->device));
    CU_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    for (auto &ev : e->ev) CU_TRY(cudaEventCreate(&ev));

    e->extradata = write_extradata(e->cfg);
    build_tables(e->cfg, e->tab);
    const Layout &L = e->tab.layout;
    e->state_in_smem = replay_smem_bytes(L) <= 227 * 1024;
    if (const char *v = getenv("FFV1B200_REPLAY_STATE")) { if (!strcmp(v, "global")) e->state_in_smem = false; }
    CU_TRY(configure_kernels(L));
    e->fast_pixel = pixel_fast_geometry_ok(L, e->tab.slices.data(), (int)e->tab.slices.size());
    if (const char *v = getenv("FFV1B200_PIXEL")) { if (!strcmp(v, "generic")) e->fast_pixel = false; }
    if (e->fast_pixel) CU_TRY(configure_pixel_fast(L));
    e->ctx_replay = ctx_replay_supported(L);
    if (const char *v = getenv("FFV1B200_REPLAY")) { if (!strcmp(v, "warp")) e->ctx_replay = false; }
    if (e->ctx_replay) CU_TRY(configure_ctx_replay(L));
    CU_TRY(e->d_ctiles.upload(e->tab.ctiles.data(), e->tab.ctiles.size(), e->stream));
    for (auto &g : e->tab.slices) for (int p = 0; p < L.nplanes; p++) e->max_plane_width = std::max(e->max_plane_width, g.pw[p]);
    CU_TRY(cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, e->device));

    CU_TRY(e->d_slices.upload(e->tab.slices.data(), e->tab.slices.size(), e->stream));
    CU_TRY(e->d_lines.upload(e->tab.lines.data(), e->tab.lines.size(), e->stream));
    CU_TRY(e->d_pc_lines.upload(e->tab.pc_lines.data(), e->tab.pc_lines.size(), e->stream));
    CU_TRY(e->d_tiles.upload(e->tab.tiles.data(), e->tab.tiles.size(), e->stream));
    CU_TRY(e->d_quant.upload(&e->cfg.quant_tables[e->cfg.context_model][0][0], 5 * 256, e->stream));
    uint8_t lut[512];
    coder_state_tables(e->cfg, lut, lut + 256);
    CU_TRY(e->d_lut.upload(lut, 512, e->stream));
    std::vector<uint8_t> one_pow(33 * 256);                      // one_state applied k times (k_replay's zero-run shortcut)
    for (int p = 0; p < 256; p++) one_pow[p] = (uint8_t)p;
    for (int k = 1; k <= 32; k++)
        for (int p = 0; p < 256; p++) one_pow[k * 256 + p] = lut[256 + one_pow[(k - 1) * 256 + p]];
    CU_TRY(e->d_one_pow.upload(one_pow.data(), one_pow.size(), e->stream));
    CU_TRY(e->d_run_pc.upload(e->tab.run_pc.data(), e->tab.run_pc.size(), e->stream));
    CU_TRY(e->d_prefix.alloc((size_t)(L.nslices * 2 + 1) * kMaxPrefix));
    CU_TRY(e->d_prefix_len.alloc((size_t)L.nslices * 2));
    r = alloc_batch_buffers(e.get());
    if (r < 0) return r;
    CU_TRY(cudaStreamSynchronize(e->stream));
    *out = e.release();
    return 0;
}

void ffv1b200_enc_close(FFV1B200Encoder *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) { cudaStreamSynchronize(e->stream); cudaStreamDestroy(e->stream); }
    for (auto &ev : e->ev) if (ev) cudaEventDestroy(ev);
    delete e;
}

int ffv1b200_enc_extradata(const FFV1B200Encoder *e, const uint8_t **data, int *size)
{
    if (!e || !data || !size) return FFV1B200_ERR_EINVAL;
    *data = e->extradata.empty() ? nullptr : e->extradata.data();
    *size = (int)e->extradata.size();
    return 0;
}

int ffv1b200_enc_info(const FFV1B200Encoder *e, FFV1B200EncInfo *i)
{
    if (!e || !i) return FFV1B200_ERR_EINVAL;
    const Config &c = e->cfg;
    i->version = c.version; i->micro_version = c.micro_version; i->ac = c.ac; i->colorspace = c.colorspace;
    i->bits_per_raw_sample = c.bits; i->chroma_planes = c.chroma_planes; i->chroma_h_shift = c.chroma_h_shift;
    i->chroma_v_shift = c.chroma_v_shift; i->transparency = c.transparency; i->num_h_slices = c.num_h_slices;
    i->num_v_slices = c.num_v_slices; i->slice_count = c.slice_count(); i->ec = c.ec; i->intra = c.intra;
    i->context_count = c.context_count[c.context_model]; i->plane_count = c.plane_count; i->max_batch_frames = e->max_batch;
    int64_t samples = 0;
    for (auto &g : e->tab.slices) samples += g.nsamples;
    i->samples_per_frame = samples;
    i->frame_bytes = c.frame_bytes();
    return 0;
}

void ffv1b200_enc_set_frame_props(FFV1B200Encoder *e, const FFV1B200FrameProps *p)
{
    if (!e || !p) return;
    if (p->sar_num != e->props.sar_num || p->sar_den != e->props.sar_den || p->picture_structure != e->props.picture_structure) {
        e->props = *p;
        e->prefix_dirty = true;
    }
}

int ffv1b200_enc_encode_host(FFV1B200Encoder *e, int nframes, const uint8_t *const *planes, const int *linesizes,
                             uint8_t *out, size_t out_cap, FFV1B200Packet *pkts, size_t *needed)
{
    if (!e || !planes || !linesizes || !out || !pkts) return fail(FFV1B200_ERR_EINVAL, "null argument");
    if (nframes < 1 || nframes > e->max_batch) return fail(FFV1B200_ERR_EINVAL, "nframes outside 1..max_batch_frames");
    CU_TRY(cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    cudaEventRecord(e->ev[5], s);
    // host -> device: every plane of every frame into the staging area (our pitch)
    int ls[4] = {0, 0, 0, 0};
    for (int f = 0; f < nframes; f++)
        for (int i = 0; i < 4; i++) {
            const uint8_t *dptr = nullptr;
            if (i < e->cfg.nb_src_planes) {
                int rows, rb; e->cfg.plane_dims(i, &rows, &rb);
                if (!planes[f * 4 + i]) return fail(FFV1B200_ERR_EINVAL, "missing plane pointer");
                uint8_t *dst = e->d_in.p + (size_t)f * e->in_frame_stride + e->in_plane_off[i];
                CU_TRY(cudaMemcpy2DAsync(dst, e->in_pitch[i], planes[f * 4 + i], linesizes[f * 4 + i], rb, rows, cudaMemcpyHostToDevice, s));
                e->stats.h2d_bytes += (int64_t)rb * rows;
                dptr = dst; ls[i] = e->in_pitch[i];
            }
            e->h_planes.p[f * 4 + i] = dptr;
        }
    CU_TRY(cudaMemcpyAsync(e->d_planes.p, e->h_planes.p, sizeof(void *) * 4 * nframes, cudaMemcpyHostToDevice, s));
    // packets are assembled in our device buffer, then copied out
    const size_t want = std::max<size_t>(out_cap, 1 << 20);
    if (e->d_out.n < want) { e->d_out.release(); CU_TRY(e->d_out.alloc(want)); }
    int r = run_pipeline(e, nframes, ls, e->d_out.p, std::min(e->d_out.n, out_cap), s);
    if (r < 0) {
        if (r == FFV1B200_ERR_BUFFER_TOO_SMALL && needed) *needed = (size_t)e->h_status.p[2];
        return r;
    }
    const size_t total = (size_t)e->h_pkt_off.p[nframes];
    CU_TRY(cudaMemcpyAsync(out, e->d_out.p, total, cudaMemcpyDeviceToHost, s));
    cudaEventRecord(e->ev[6], s);
    CU_TRY(cudaStreamSynchronize(s));
    e->stats.d2h_bytes += (int64_t)total;
    float ms; cudaEventElapsedTime(&ms, e->ev[5], e->ev[6]); e->stats.ms_total += ms;
    if (needed) *needed = total;
    finish_batch(e, nframes, pkts);
    return nframes;
}

int ffv1b200_enc_encode_device(FFV1B200Encoder *e, int nframes, const void *const *d_planes, const int *linesizes,
                               void *d_out, size_t d_out_cap, FFV1B200Packet *pkts, size_t *needed, void *stream)
{
    if (!e || !d_planes || !linesizes || !d_out || !pkts) return fail(FFV1B200_ERR_EINVAL, "null argument");
    if (nframes < 1 || nframes > e->max_batch) return fail(FFV1B200_ERR_EINVAL, "nframes outside 1..max_batch_frames");
    CU_TRY(cudaSetDevice(e->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    int ls[4] = {linesizes[0], linesizes[1], linesizes[2], linesizes[3]};
    for (int f = 0; f < nframes; f++)
        for (int i = 0; i < 4; i++) {
            if (linesizes[f * 4 + i] != ls[i]) return fail(FFV1B200_ERR_EINVAL, "all frames of a batch must share linesizes");
            e->h_planes.p[f * 4 + i] = (const uint8_t *)d_planes[f * 4 + i];
        }
    cudaEventRecord(e->ev[5], s);
    CU_TRY(cudaMemcpyAsync(e->d_planes.p, e->h_planes.p, sizeof(void *) * 4 * nframes, cudaMemcpyHostToDevice, s));
    int r = run_pipeline(e, nframes, ls, (uint8_t *)d_out, d_out_cap, s);
    if (r < 0) {
        if (r == FFV1B200_ERR_BUFFER_TOO_SMALL && needed) *needed = (size_t)e->h_status.p[2];
        return r;
    }
    cudaEventRecord(e->ev[6], s);
    CU_TRY(cudaStreamSynchronize(s));
    float ms; cudaEventElapsedTime(&ms, e->ev[5], e->ev[6]); e->stats.ms_total += ms;
    if (needed) *needed = (size_t)e->h_pkt_off.p[nframes];
    finish_batch(e, nframes, pkts);
    return nframes;
}

int ffv1b200_enc_stats(const FFV1B200Encoder *e, FFV1B200EncStats *s)
{
    if (!e || !s) return FFV1B200_ERR_EINVAL;
    *s = e->stats;
    return 0;
}

int64_t ffv1b200_enc_debug_records(FFV1B200Encoder *e, int frame, int slice, uint32_t *dst, int64_t cap)
{
    if (!e || !dst) return FFV1B200_ERR_EINVAL;
    if (frame < 0 || frame >= e->last_nframes || slice < 0 || slice >= e->cfg.slice_count()) return fail(FFV1B200_ERR_EINVAL, "bad frame/slice");
    cudaSetDevice(e->device);
    const Layout &L = e->tab.layout;
    const SliceGeom &g = e->tab.slices[slice];
    std::vector<uint32_t> tmp(g.rec_count);
    cudaError_t ce = cudaMemcpy(tmp.data(), e->d_rec.p + (size_t)frame * L.rec_per_frame + g.rec_first, sizeof(uint32_t) * g.rec_count, cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) return fail(FFV1B200_ERR_EXTERNAL, cudaGetErrorString(ce));
    int64_t n = 0;
    for (int li = 0; li < g.nlines; li++) {
        const LineDesc &ld = e->tab.lines[g.line_first + li];
        for (int x = 0; x < ld.w; x++) { if (n < cap) dst[n] = tmp[ld.rec_off + x]; n++; }
    }
    return n;
}

} // extern "C"
