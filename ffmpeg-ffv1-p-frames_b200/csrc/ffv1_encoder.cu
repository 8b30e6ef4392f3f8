// ffv1_encoder.cu -- C-ABI encoder entry points (include/ffv1_b200.h) and batch orchestration.
// Replaces encode_init / encode_frame / encode_close of the reference (ffv1enc.c:669-1029, 1222-1373, 1375-1379).
//
// A batch goes through three engines that run concurrently for consecutive batches:
//   copy-in stream : host frames -> device (merged into as few, as large copies as the host layout allows)
//   compute stream : per-pixel pass -> state replay -> range coder / Golomb coder -> packet assembly
//   copy-out stream: packets -> host
// Two "slots" hold what must be private to a batch in flight (input frames, packets, small control arrays); the large
// intermediates (records, decisions, context lists, coder scratch) are shared because the compute stream is serial.
// ffv1b200_enc_submit_host / ffv1b200_enc_collect expose the pipeline (AV_CODEC_CAP_DELAY semantics: packets of batch
// k come back while batch k+1 is being coded); ffv1b200_enc_encode_host is submit + collect.
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

namespace {
constexpr int kSlots = 2;       // batches in flight: H2D of batch k+2, kernels of k+1 and D2H of k overlap; in steady state that
                                // already runs at min(kernel rate, link rate), a third slot only costs 10 GB
constexpr int kCarry = 4;       // ring of model-state buffers: batch k reads [k % 4] and writes [(k+1) % 4]

struct Slot {
    DevBuf<uint8_t> d_in, d_out;
    DevBuf<const uint8_t *> d_planes; PinnedBuf<const uint8_t *> h_planes;
    DevBuf<int32_t> d_seg_first, d_frame_seg; PinnedBuf<int32_t> h_seg_first, h_frame_seg;
    DevBuf<uint8_t> d_frame_key; PinnedBuf<uint8_t> h_frame_key;
    DevBuf<unsigned long long> d_status; PinnedBuf<unsigned long long> h_status;
    DevBuf<uint32_t> d_pkt_size; PinnedBuf<uint32_t> h_pkt_size;
    DevBuf<uint64_t> d_pkt_off; PinnedBuf<uint64_t> h_pkt_off;
    // what every slice codes before its first sample (keyframe bit, slice header) depends on the frame properties the
    // batch was submitted with (ffv1enc.c:1044-1049): each slot keeps its own copy, so a batch in flight -- or re-run by
    // recover() -- is never coded with the properties of a later one
    DevBuf<uint16_t> d_prefix; DevBuf<int32_t> d_prefix_len;
    DevBuf<uint8_t> d_gprefix; DevBuf<int32_t> d_gprefix_len;
    FFV1B200FrameProps props{0, 1, 3};
    bool prefix_valid = false;
    cudaEvent_t ev_h2d = nullptr, ev_small = nullptr, ev_d2h = nullptr, ev[6] = {nullptr};
    bool d2h_pending = false;        // a packet copy from this slot's output area may still be running on the copy-out stream
    int nframes = 0, nseg = 0, carry_in = 0;
    int64_t first_pn = 0;
    int ls[4] = {0, 0, 0, 0};
    bool busy = false, fast = false;
    uint8_t *out = nullptr;          // where the packets are assembled (own d_out, or the caller's device buffer)
    size_t out_cap = 0;
    int64_t h2d_bytes = 0;
};
} // namespace

struct FFV1B200Encoder {
    Config cfg;
    Tables tab;
    std::vector<uint8_t> extradata;
    int device = 0, max_batch = 64;
    int64_t picture_number = 0;      // of the next frame to be SUBMITTED
    FFV1B200FrameProps props{0, 1, 3};          // of the frames submitted next
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;

    // static device tables
    DevBuf<SliceGeom> d_slices; DevBuf<LineDesc> d_lines; DevBuf<int32_t> d_pc_lines; DevBuf<TileDesc> d_tiles;
    DevBuf<CtxTile> d_ctiles;
    FastPlan fast_plan; DevBuf<FastItemDesc> d_fast_items;
    DevBuf<int16_t> d_quant; DevBuf<uint8_t> d_lut, d_one_pow, d_run_pc, d_init_state;
    // first pass of a two-pass encode: statistics accumulated on the device, keyframes coded so far
    DevBuf<unsigned long long> d_rc_stat, d_rc_stat2;
    int gob_count = 0;
    // shared intermediates
    DevBuf<uint32_t> d_rec, d_run_cnt, d_slice_bytes, d_line_pos, d_ctx_hist, d_list_start, d_list_count;
    DevBuf<uint16_t> d_dec, d_list_order;
    DevBuf<uint2> d_lists;
    DevBuf<uint8_t> d_scratch, d_state_seg, d_carry[kCarry];
    DevBuf<uint8_t> d_rct_idx;       // version-4 RGB: RCT coefficient pair chosen for every (frame, slice)
    Slot slot[kSlots];
    uint64_t submitted = 0, collected = 0;      // batch counters; slot of batch k = k % kSlots
    int carry_next = 0;                          // ring index holding the state after the last submitted batch
    double dec_per_sample = 5.0;
    bool state_in_smem = true, fast_pixel = false, ctx_replay = false, golomb_lists = false;
    int max_plane_width = 0, num_sms = 148, max_ctile_samples = 0;
    FFV1B200EncStats stats{};
    int last_slot = 0;
    int cuda_pending = 0;            // frames of a batch that ffv1b200_enc_encode_cuda has coded but could not hand out
};

namespace {

int fail(int code, const std::string &msg) { set_last_error(msg); return code; }

#define CU_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return fail(FFV1B200_ERR_EXTERNAL, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)

// the slot's header tables for the frame properties its batch is submitted with (stream-ordered on the stream the batch's
// kernels run on; the sources are pageable, so the copies are staged before the calls return)
int upload_prefixes(FFV1B200Encoder *e, Slot &sl, cudaStream_t s)
{
    const FFV1B200FrameProps &pr = e->props;
    if (sl.prefix_valid && sl.props.sar_num == pr.sar_num && sl.props.sar_den == pr.sar_den && sl.props.picture_structure == pr.picture_structure)
        return 0;
    const int ns = e->cfg.slice_count();
    const int nvar = rct_variants(e->cfg);                                   // version-4 RGB: one header per RCT coefficient pair
    if (e->tab.layout.golomb) {
        std::vector<uint8_t> gp((size_t)ns * 2 * nvar * kMaxGolombPrefix, 0);
        std::vector<int32_t> gl((size_t)ns * 2 * nvar, 0);
        for (int sx = 0; sx < ns; sx++)
            for (int key = 0; key < 2; key++)
                for (int v = 0; v < nvar; v++) {
                    std::vector<uint8_t> b = slice_prefix_bytes(e->cfg, sx, key != 0, pr.sar_num, pr.sar_den, pr.picture_structure, v);
                    if ((int)b.size() > kMaxGolombPrefix) return fail(FFV1B200_ERR_EINVAL, "slice header too long");
                    const size_t at = (size_t)(sx * 2 + key) * nvar + v;
                    std::copy(b.begin(), b.end(), gp.begin() + at * kMaxGolombPrefix);
                    gl[at] = (int32_t)b.size();
                }
        CU_TRY(sl.d_gprefix.upload(gp.data(), gp.size(), s));
        CU_TRY(sl.d_gprefix_len.upload(gl.data(), gl.size(), s));
    } else {
        std::vector<uint16_t> pre((size_t)(ns * 2 * nvar + 1) * kMaxPrefix, 0);
        pre[(size_t)ns * 2 * nvar * kMaxPrefix] = 129;                       // the decision that closes every slice (state 129, bit 0)
        std::vector<int32_t> len((size_t)ns * 2 * nvar, 0);
        for (int sx = 0; sx < ns; sx++)
            for (int key = 0; key < 2; key++)
                for (int v = 0; v < nvar; v++) {
                    std::vector<uint16_t> d = slice_prefix_decisions(e->cfg, sx, key != 0, pr.sar_num, pr.sar_den, pr.picture_structure, v);
                    if ((int)d.size() > kMaxPrefix) return fail(FFV1B200_ERR_EINVAL, "slice header too long");
                    const size_t at = (size_t)(sx * 2 + key) * nvar + v;
                    std::copy(d.begin(), d.end(), pre.begin() + at * kMaxPrefix);
                    len[at] = (int32_t)d.size();
                }
        CU_TRY(sl.d_prefix.upload(pre.data(), pre.size(), s));
        CU_TRY(sl.d_prefix_len.upload(len.data(), len.size(), s));
    }
    sl.props = pr;
    sl.prefix_valid = true;
    return 0;
}

int alloc_buffers(FFV1B200Encoder *e)
{
    const Layout &L = e->tab.layout;
    const size_t F = (size_t)e->max_batch;
    cudaStream_t s = e->s_comp;
    CU_TRY(e->d_rec.alloc((size_t)L.rec_per_frame * F));
    CU_TRY(e->d_run_cnt.alloc((size_t)L.runs_per_frame * F));
    CU_TRY(e->d_slice_bytes.alloc((size_t)L.nslices * F));
    if (rct_variants(e->cfg) > 1) CU_TRY(e->d_rct_idx.alloc((size_t)L.nslices * F));
    // coder output: bytes (Golomb-Rice) or k_rangecode's 16-bit entries, `scratch_cap` of them per slice
    CU_TRY(e->d_scratch.alloc((size_t)L.scratch_per_frame * F * (L.golomb ? 1 : 2)));
    if (!L.golomb || e->golomb_lists) CU_TRY(e->d_dec.alloc((size_t)L.dec_per_frame * F + 64));   // Golomb lists: the code words
    const size_t state_bytes = (size_t)L.nslices * L.npc * L.ctx_count * 32;
    for (int k = 0; k < kCarry; k++) {
        CU_TRY(e->d_carry[k].alloc(state_bytes));
        CU_TRY(cudaMemsetAsync(e->d_carry[k].p, 128, state_bytes, s));
    }
    const size_t g = e->cfg.gop_size > 0 ? e->cfg.gop_size : 1;
    const size_t nseg_max = (F + g - 1) / g + 1;
    if (e->ctx_replay || e->golomb_lists) {
        const size_t nchains = nseg_max * L.nslices * L.npc;
        CU_TRY(e->d_line_pos.alloc((size_t)L.lines_per_frame * F));
        CU_TRY(e->d_ctx_hist.alloc((size_t)L.ctiles_per_frame * L.ctx_count * F));
        CU_TRY(e->d_list_start.alloc(nchains * L.ctx_count));
        CU_TRY(e->d_list_count.alloc(nchains * L.ctx_count));
        CU_TRY(e->d_list_order.alloc(nchains * L.ctx_count));
        // 8-byte list entries; the tile-sorted lists hold 4-byte ones (12.4 instead of 24.9 MB per 1080p frame)
        CU_TRY(e->d_lists.alloc(L.tiled_lists ? ((size_t)L.samples_per_frame * F + 1) / 2 : (size_t)L.samples_per_frame * F));
    }
    if ((!e->state_in_smem && !e->ctx_replay) || (e->ctx_replay && ctx_replay_needs_global_state(L)) || (L.golomb && !e->golomb_lists))
        CU_TRY(e->d_state_seg.alloc(state_bytes * nseg_max));            // one state set per GOP segment of a batch
    for (Slot &sl : e->slot) {
        CU_TRY(sl.d_planes.alloc(F * 4)); CU_TRY(sl.h_planes.alloc(F * 4));
        CU_TRY(sl.d_seg_first.alloc(F + 1)); CU_TRY(sl.h_seg_first.alloc(F + 1));
        CU_TRY(sl.d_frame_seg.alloc(F)); CU_TRY(sl.h_frame_seg.alloc(F));
        CU_TRY(sl.d_frame_key.alloc(F)); CU_TRY(sl.h_frame_key.alloc(F));
        CU_TRY(sl.d_status.alloc(8)); CU_TRY(sl.h_status.alloc(8));
        CU_TRY(sl.d_pkt_size.alloc(F)); CU_TRY(sl.h_pkt_size.alloc(F));
        CU_TRY(sl.d_pkt_off.alloc(F + 1)); CU_TRY(sl.h_pkt_off.alloc(F + 1));
        CU_TRY(cudaEventCreate(&sl.ev_h2d)); CU_TRY(cudaEventCreate(&sl.ev_small));
        CU_TRY(cudaEventCreateWithFlags(&sl.ev_d2h, cudaEventDisableTiming));
        for (auto &ev : sl.ev) CU_TRY(cudaEventCreate(&ev));
    }
    return 0;
}

EncDeviceTables device_tables(FFV1B200Encoder *e, const Slot &sl)
{
    EncDeviceTables t;
    t.layout = e->tab.layout;
    t.slices = e->d_slices.p; t.lines = e->d_lines.p; t.pc_lines = e->d_pc_lines.p; t.tiles = e->d_tiles.p;
    t.ctiles = e->d_ctiles.p;
    t.quant = e->d_quant.p; t.trans_lut = e->d_lut.p; t.one_pow = e->d_one_pow.p; t.run_pc = e->d_run_pc.p;
    t.init_state = e->d_init_state.p;
    t.gprefix = sl.d_gprefix.p; t.gprefix_len = sl.d_gprefix_len.p; t.prefix = sl.d_prefix.p; t.prefix_len = sl.d_prefix_len.p;
    t.ec = e->cfg.ec; t.version = e->cfg.version; t.state_in_smem = e->state_in_smem ? 1 : 0;
    t.nvar = rct_variants(e->cfg);
    return t;
}

// keyframe flags and GOP segments of a batch that starts at picture number pn (keyframe rule: ffv1enc.c:1299)
void plan_batch(FFV1B200Encoder *e, Slot &sl, int nframes, int64_t pn)
{
    int nseg = 0;
    for (int f = 0; f < nframes; f++) {
        const bool key = e->cfg.gop_size == 0 || ((pn + f) % e->cfg.gop_size) == 0;
        sl.h_frame_key.p[f] = key ? 1 : 0;
        if (key) e->gob_count++;                              // ffv1enc.c:1302
        if (f == 0 || key) sl.h_seg_first.p[nseg++] = f;
        sl.h_frame_seg.p[f] = nseg - 1;
    }
    sl.h_seg_first.p[nseg] = nframes;
    sl.nframes = nframes; sl.nseg = nseg; sl.first_pn = pn;
}

// enqueues the whole device pipeline of a slot on stream s (plane pointers already in sl.h_planes)
int enqueue_kernels(FFV1B200Encoder *e, Slot &sl, cudaStream_t s)
{
    const Layout &L = e->tab.layout;
    const int nframes = sl.nframes;
    CU_TRY(cudaMemcpyAsync(sl.d_planes.p, sl.h_planes.p, sizeof(void *) * 4 * nframes, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(sl.d_seg_first.p, sl.h_seg_first.p, sizeof(int32_t) * (sl.nseg + 1), cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(sl.d_frame_key.p, sl.h_frame_key.p, nframes, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(sl.d_frame_seg.p, sl.h_frame_seg.p, sizeof(int32_t) * nframes, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemsetAsync(sl.d_status.p, 0, sizeof(unsigned long long) * 8, s));

    // the packet area of this slot may still be on its way to the host (ffv1b200_enc_collect_async)
    if (sl.d2h_pending) CU_TRY(cudaStreamWaitEvent(s, sl.ev_d2h, 0));
    EncDeviceTables t = device_tables(e, sl);
    EncBatch b{};
    b.nframes = nframes; b.nseg = sl.nseg;
    b.planes = sl.d_planes.p;
    for (int i = 0; i < 4; i++) b.linesize[i] = sl.ls[i];
    b.rec = e->d_rec.p; b.run_cnt = e->d_run_cnt.p; b.dec = e->d_dec.p;
    b.seg_first = sl.d_seg_first.p; b.frame_key = sl.d_frame_key.p; b.frame_seg = sl.d_frame_seg.p;
    b.scratch = e->d_scratch.p; b.slice_bytes = e->d_slice_bytes.p;
    b.pkt_size = sl.d_pkt_size.p; b.pkt_off = sl.d_pkt_off.p;
    b.out = sl.out; b.out_capacity = sl.out_cap;
    b.state_seg = e->d_state_seg.p;
    b.carry_in = e->d_carry[sl.carry_in].p; b.carry_out = e->d_carry[(sl.carry_in + 1) % kCarry].p;
    b.status = sl.d_status.p;
    b.line_pos = e->d_line_pos.p; b.ctx_hist = e->d_ctx_hist.p;
    b.list_start = e->d_list_start.p; b.list_count = e->d_list_count.p; b.list_order = e->d_list_order.p; b.lists = e->d_lists.p;
    b.rct_idx = e->d_rct_idx.p;
    if (L.tiled_lists) {        // the tile tables and tile decision counts live where the chain-wide lists keep their histograms
        b.tile_tab_pitch = (L.ctx_count + 8) & ~7;
        b.tile_tab = reinterpret_cast<uint16_t *>(e->d_ctx_hist.p);
        b.tile_nd = e->d_line_pos.p;
    }

    cudaEventRecord(sl.ev[0], s);
    if (t.nvar > 1) { launch_rct_search(t, b, s); e->stats.kernel_launches++; }      // choose_rct_params, before the RCT is applied
    if (sl.fast) {
        // tensor-map TMA needs the frames of the batch at a constant distance (true for the staged host path)
        long long fstride = nframes > 1 ? (long long)(sl.h_planes.p[4] - sl.h_planes.p[0]) : 0;
        for (int f = 1; f < nframes && fstride >= 0; f++)
            for (int i = 0; i < e->cfg.nb_src_planes; i++)
                if ((long long)(sl.h_planes.p[f * 4 + i] - sl.h_planes.p[(f - 1) * 4 + i]) != fstride) { fstride = -1; break; }
        launch_pixel_fast(t, b, e->fast_plan, e->d_fast_items.p, e->num_sms, s, sl.h_planes.p, fstride);
    }
    else         launch_pixel(t, b, s);
    cudaEventRecord(sl.ev[1], s);
    if (e->ctx_replay) {
        uint32_t max_dec_cap = 0;
        for (const SliceGeom &g : e->tab.slices) for (int pc = 0; pc < 3; pc++) max_dec_cap = std::max(max_dec_cap, g.dec_cap[pc]);
        launch_ctx_replay(t, b, e->max_ctile_samples, max_dec_cap, s);
    }
    else if (!L.golomb) launch_replay(t, b, s);
    cudaEventRecord(sl.ev[2], s);
    if (!L.golomb) launch_rangecode(t, b, s);
    else if (e->golomb_lists) { launch_golomb_lists(t, b, s); launch_golomb_coder(t, b, s); }
    else launch_golomb(t, b, s);
    cudaEventRecord(sl.ev[3], s);
    launch_pack(t, b, s);
    cudaEventRecord(sl.ev[4], s);
    if ((e->cfg.pass_flags & kPass1) && !L.golomb) {          // the reference gathers statistics in the range-coder modes only
        CU_TRY(launch_pass1_stats(t, b, e->d_rc_stat.p, e->d_rc_stat2.p, s));
        e->stats.kernel_launches++;
    }
    e->stats.kernel_launches += L.golomb ? (e->golomb_lists ? 9 : 5) : (e->ctx_replay ? (L.tiled_lists ? 8 : 10) : 6);     // (k_pack_init included)
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(sl.h_status.p, sl.d_status.p, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(sl.h_pkt_size.p, sl.d_pkt_size.p, sizeof(uint32_t) * nframes, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(sl.h_pkt_off.p, sl.d_pkt_off.p, sizeof(uint64_t) * (nframes + 1), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaEventRecord(sl.ev_small, s));
    e->stats.d2h_bytes += 64 + 12 * (int64_t)nframes + 8;
    return 0;
}

// A scratch area was too small: grow it and run every batch in flight again, oldest first (their inputs and the
// model state they start from are still intact: the state ring is deeper than the pipeline).
int recover(FFV1B200Encoder *e, bool own_out, cudaStream_t s)
{
    for (int attempt = 0; attempt < 8; attempt++) {
        CU_TRY(cudaStreamSynchronize(e->s_in));
        CU_TRY(cudaStreamSynchronize(s));
        bool again = false;
        for (uint64_t k = e->collected; k < e->submitted && !again; k++) {
            Slot &sl = e->slot[k % kSlots];
            const unsigned long long *st = sl.h_status.p;
            if (st[0]) {            // a decision region overflowed: st[0] = needed entries per sample * 256
                e->dec_per_sample = std::max(e->dec_per_sample * 1.25, (double)st[0] / 256.0 * 1.1);
                layout_decisions(e->tab, e->dec_per_sample);
                e->d_dec.release();
                CU_TRY(e->d_dec.alloc((size_t)e->tab.layout.dec_per_frame * e->max_batch + 64));
                CU_TRY(e->d_slices.upload(e->tab.slices.data(), e->tab.slices.size(), s));
                again = true;
            } else if (st[1]) {     // a slice's coder output overflowed its scratch region
                double need = 1.0;
                for (auto &g : e->tab.slices) need = std::max(need, (double)st[1] / (double)g.scratch_cap);
                uint32_t cur = 0;
                for (auto &g : e->tab.slices) {
                    g.scratch_cap = (uint32_t)((((uint64_t)(g.scratch_cap * need * 1.25)) + 255) & ~255ull);
                    g.scratch_off = cur; cur += g.scratch_cap;
                }
                e->tab.layout.scratch_per_frame = cur;
                e->d_scratch.release();
                CU_TRY(e->d_scratch.alloc((size_t)cur * e->max_batch * (e->tab.layout.golomb ? 1 : 2)));
                CU_TRY(e->d_slices.upload(e->tab.slices.data(), e->tab.slices.size(), s));
                again = true;
            } else if (st[2]) {     // the packet area was too small
                if (!own_out) {
                    set_last_error("output buffer too small: need " + std::to_string(st[2]) + " bytes");
                    return FFV1B200_ERR_BUFFER_TOO_SMALL;
                }
                sl.d_out.release();
                CU_TRY(sl.d_out.alloc((size_t)st[2] + (size_t)st[2] / 8 + 65536));
                sl.out = sl.d_out.p; sl.out_cap = sl.d_out.n;
                again = true;
            }
        }
        if (!again) return 0;
        e->stats.retries++;
        for (uint64_t k = e->collected; k < e->submitted; k++) {
            int r = enqueue_kernels(e, e->slot[k % kSlots], s);
            if (r < 0) return r;
        }
        for (uint64_t k = e->collected; k < e->submitted; k++) CU_TRY(cudaEventSynchronize(e->slot[k % kSlots].ev_small));
    }
    return fail(FFV1B200_ERR_EXTERNAL, "scratch buffers kept overflowing");
}

void account(FFV1B200Encoder *e, Slot &sl)
{
    float ms;
    cudaEventElapsedTime(&ms, sl.ev[0], sl.ev[1]); e->stats.ms_pixel_kernel += ms;
    cudaEventElapsedTime(&ms, sl.ev[1], sl.ev[2]); e->stats.ms_model_kernel += ms;
    cudaEventElapsedTime(&ms, sl.ev[2], sl.ev[3]); e->stats.ms_coder_kernel += ms;
    cudaEventElapsedTime(&ms, sl.ev[3], sl.ev[4]); e->stats.ms_pack_kernel += ms;
    cudaEventElapsedTime(&ms, sl.ev[0], sl.ev[4]); e->stats.ms_total += ms;
    e->stats.decisions += (int64_t)sl.h_status.p[3];
    e->stats.frames += sl.nframes;
    e->stats.packet_bytes += (int64_t)sl.h_pkt_off.p[sl.nframes];
    uint64_t samples = 0;
    for (auto &g : e->tab.slices) samples += g.nsamples;
    e->stats.samples += (int64_t)samples * sl.nframes;
    e->stats.h2d_bytes += sl.h2d_bytes;
}

void fill_packets(const Slot &sl, FFV1B200Packet *pkts)
{
    for (int f = 0; f < sl.nframes; f++) {
        pkts[f].offset = (int64_t)sl.h_pkt_off.p[f];
        pkts[f].size = (int32_t)sl.h_pkt_size.p[f];
        pkts[f].flags = sl.h_frame_key.p[f] ? FFV1B200_PKT_FLAG_KEY : 0;
        pkts[f].picture_number = sl.first_pn + f;
    }
}

bool pixel_fast_ok(const FFV1B200Encoder *e, const Slot &sl)
{
    bool fast = e->fast_pixel;
    for (int i = 0; i < 4 && fast; i++) fast = (sl.ls[i] & 15) == 0;
    for (int i = 0; i < sl.nframes * 4 && fast; i++) fast = (reinterpret_cast<uintptr_t>(sl.h_planes.p[i]) & 15) == 0;
    return fast;
}

// waits until no packet copy of an earlier ffv1b200_enc_collect_async is running any more
int finish_output(FFV1B200Encoder *e)
{
    for (Slot &sl : e->slot)
        if (sl.d2h_pending) {
            CU_TRY(cudaEventSynchronize(sl.ev_d2h));
            sl.d2h_pending = false;
        }
    return 0;
}

} // namespace

extern "C" {

static EncOptions options_of(const FFV1B200EncParams *p)
{
    EncOptions o{p->width, p->height, p->pix_fmt, p->gop_size, p->level, p->slices, p->coder, p->context, p->slicecrc};
    if (p->flags & FFV1B200_FLAG_PASS1) o.pass_flags |= kPass1;
    if (p->flags & FFV1B200_FLAG_PASS2) o.pass_flags |= kPass2;
    if (p->stats_in) o.stats_in = p->stats_in;
    o.strict_experimental = p->strict_std_compliance <= -2;
    o.bits_per_raw_sample = p->bits_per_raw_sample;
    return o;
}

int ffv1b200_enc_resolve(const FFV1B200EncParams *p, FFV1B200EncInfo *i, uint8_t *extradata, int cap, int *size)
{
    if (!p || !p->pix_fmt) return fail(FFV1B200_ERR_EINVAL, "null argument");
    Config c;
    std::string err;
    int r = resolve_encoder(options_of(p), c, err);
    if (r < 0) return fail(r, err);
    if (i) {
        memset(i, 0, sizeof(*i));
        i->version = c.version; i->micro_version = c.micro_version; i->ac = c.ac; i->colorspace = c.colorspace;
        i->bits_per_raw_sample = c.bits; i->chroma_planes = c.chroma_planes; i->chroma_h_shift = c.chroma_h_shift;
        i->chroma_v_shift = c.chroma_v_shift; i->transparency = c.transparency; i->num_h_slices = c.num_h_slices;
        i->num_v_slices = c.num_v_slices; i->slice_count = c.slice_count(); i->ec = c.ec; i->intra = c.intra;
        i->context_count = c.context_count[c.context_model]; i->plane_count = c.plane_count;
        i->frame_bytes = c.frame_bytes();
    }
    const std::vector<uint8_t> xd = write_extradata(c);
    if (size) *size = (int)xd.size();
    if (extradata) {
        if ((int)xd.size() > cap) return fail(FFV1B200_ERR_BUFFER_TOO_SMALL, "extradata buffer too small");
        if (!xd.empty()) memcpy(extradata, xd.data(), xd.size());
    }
    return 0;
}

int ffv1b200_enc_open(FFV1B200Encoder **out, const FFV1B200EncParams *p)
{
    if (!out || !p || !p->pix_fmt) return fail(FFV1B200_ERR_EINVAL, "null argument");
    *out = nullptr;
    const EncOptions o = options_of(p);
    std::unique_ptr<FFV1B200Encoder> e(new FFV1B200Encoder());
    std::string err;
    int r = resolve_encoder(o, e->cfg, err);
    if (r < 0) return fail(r, err);
    int ndev = ffv1b200_device_count();
    if (ndev < 0) return ndev;
    if (p->device < 0 || p->device >= ndev) return fail(FFV1B200_ERR_EINVAL, "no such CUDA device");
    e->device = p->device;
    e->max_batch = p->max_batch_frames > 0 ? p->max_batch_frames : 64;
    if (e->max_batch > 65535) return fail(FFV1B200_ERR_EINVAL, "max_batch_frames must be below 65536");
    e->picture_number = p->first_picture_number;
    CU_TRY(cudaSetDevice(e->device));
    CU_TRY(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&e->s_comp, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));

    e->extradata = write_extradata(e->cfg);
    // test hook: start with decision regions that are too small, so that the overflow -> grow -> re-run path is taken
    if (const char *v = getenv("FFV1B200_DEC_PER_SAMPLE")) { const double q = atof(v); if (q >= 0.25 && q <= 64.0) e->dec_per_sample = q; }
    build_tables(e->cfg, e->tab);
    if (e->dec_per_sample != 5.0) layout_decisions(e->tab, e->dec_per_sample);
    const Layout &L = e->tab.layout;
    e->state_in_smem = replay_smem_bytes(L) <= 227 * 1024;
    if (const char *v = getenv("FFV1B200_REPLAY_STATE")) { if (!strcmp(v, "global")) e->state_in_smem = false; }
    CU_TRY(configure_kernels(L));
    e->fast_pixel = pixel_fast_geometry_ok(L, e->tab.slices.data(), (int)e->tab.slices.size());
    if (const char *v = getenv("FFV1B200_PIXEL")) { if (!strcmp(v, "generic")) e->fast_pixel = false; }
    if (e->fast_pixel) {
        build_pixel_fast_plan(e->tab, e->fast_plan);
        CU_TRY(configure_pixel_fast(e->fast_plan));
        CU_TRY(e->d_fast_items.upload(e->fast_plan.items.data(), e->fast_plan.items.size(), e->s_comp));
    }
    e->ctx_replay = ctx_replay_supported(L);
    if (const char *v = getenv("FFV1B200_REPLAY")) { if (!strcmp(v, "warp")) e->ctx_replay = false; }
    if (e->ctx_replay) CU_TRY(configure_ctx_replay(L));
    for (auto &g : e->tab.slices) for (int pl = 0; pl < L.nplanes; pl++) e->max_plane_width = std::max(e->max_plane_width, g.pw[pl]);
    for (const CtxTile &ct : e->tab.ctiles) {
        const SliceGeom &g = e->tab.slices[ct.slice];
        int ns = 0;
        for (int i = 0; i < ct.nlines; i++) ns += e->tab.lines[g.line_first + e->tab.pc_lines[g.pc_line_first[ct.pc] + ct.first + i]].w;
        e->max_ctile_samples = std::max(e->max_ctile_samples, ns);
    }
    // Golomb-Rice mode decomposed by context (FFV1B200_GOLOMB=serial keeps the one-coder-per-chain kernel)
    e->golomb_lists = golomb_lists_supported(L, e->max_ctile_samples);
    if (const char *v = getenv("FFV1B200_GOLOMB")) { if (!strcmp(v, "serial")) e->golomb_lists = false; }
    if (e->golomb_lists) CU_TRY(configure_ctx_replay(L));
    CU_TRY(cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, e->device));

    cudaStream_t s = e->s_comp;
    CU_TRY(e->d_slices.upload(e->tab.slices.data(), e->tab.slices.size(), s));
    CU_TRY(e->d_lines.upload(e->tab.lines.data(), e->tab.lines.size(), s));
    CU_TRY(e->d_pc_lines.upload(e->tab.pc_lines.data(), e->tab.pc_lines.size(), s));
    CU_TRY(e->d_tiles.upload(e->tab.tiles.data(), e->tab.tiles.size(), s));
    CU_TRY(e->d_ctiles.upload(e->tab.ctiles.data(), e->tab.ctiles.size(), s));
    CU_TRY(e->d_quant.upload(&e->cfg.quant_tables[e->cfg.context_model][0][0], 5 * 256, s));
    uint8_t lut[512];
    coder_state_tables(e->cfg, lut, lut + 256);
    CU_TRY(e->d_lut.upload(lut, 512, s));
    std::vector<uint8_t> one_pow(33 * 256);                      // one_state applied k times (k_replay's zero-run shortcut)
    for (int q = 0; q < 256; q++) one_pow[q] = (uint8_t)q;
    for (int k = 1; k <= 32; k++)
        for (int q = 0; q < 256; q++) one_pow[k * 256 + q] = lut[256 + one_pow[(k - 1) * 256 + q]];
    CU_TRY(e->d_one_pow.upload(one_pow.data(), one_pow.size(), s));
    CU_TRY(e->d_run_pc.upload(e->tab.run_pc.data(), e->tab.run_pc.size(), s));
    {
        // second pass: the states a keyframe starts from (ffv1.c:188-190), for the table set in use
        const std::vector<uint8_t> &init = e->cfg.initial_states[e->cfg.context_model];
        bool coded = false;
        for (uint8_t v : init) if (v != 128) { coded = true; break; }
        if (coded && !L.golomb) CU_TRY(e->d_init_state.upload(init.data(), init.size(), s));
    }
    if (e->cfg.pass_flags & kPass1) {
        CU_TRY(e->d_rc_stat.alloc(512)); CU_TRY(e->d_rc_stat2.alloc((size_t)L.ctx_count * 64));
        CU_TRY(cudaMemsetAsync(e->d_rc_stat.p, 0, 512 * sizeof(unsigned long long), s));
        CU_TRY(cudaMemsetAsync(e->d_rc_stat2.p, 0, (size_t)L.ctx_count * 64 * sizeof(unsigned long long), s));
    }
    r = alloc_buffers(e.get());
    if (r < 0) return r;
    CU_TRY(cudaStreamSynchronize(s));
    *out = e.release();
    return 0;
}

void ffv1b200_enc_close(FFV1B200Encoder *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    finish_output(e);
    for (cudaStream_t s : {e->s_in, e->s_comp, e->s_out}) if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    for (Slot &sl : e->slot) {
        if (sl.ev_h2d) cudaEventDestroy(sl.ev_h2d);
        if (sl.ev_small) cudaEventDestroy(sl.ev_small);
        if (sl.ev_d2h) cudaEventDestroy(sl.ev_d2h);
        for (auto &ev : sl.ev) if (ev) cudaEventDestroy(ev);
    }
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
    e->props = *p;               // applies to the batches submitted from now on; batches in flight keep theirs
}

int ffv1b200_enc_submit_host(FFV1B200Encoder *e, int nframes, const uint8_t *const *planes, const int *linesizes)
{
    if (!e || !planes || !linesizes) return fail(FFV1B200_ERR_EINVAL, "null argument");
    if (nframes < 1 || nframes > e->max_batch) return fail(FFV1B200_ERR_EINVAL, "nframes outside 1..max_batch_frames");
    if (e->submitted - e->collected >= (uint64_t)kSlots) return fail(FFV1B200_ERR_EINVAL, "two batches are already in flight: collect one first");
    CU_TRY(cudaSetDevice(e->device));
    Slot &sl = e->slot[e->submitted % kSlots];
    { int r = upload_prefixes(e, sl, e->s_comp); if (r < 0) return r; }
    const Config &c = e->cfg;
    const int np = c.nb_src_planes;

    // ---- device layout of the staged frames: the host layout itself when its linesizes are 16-byte multiples (one
    //      contiguous copy per run of adjacent planes / frames), else rows are re-pitched to 128 bytes
    bool direct = true;
    for (int i = 0; i < np; i++) {
        if (!planes[i]) return fail(FFV1B200_ERR_EINVAL, "missing plane pointer");
        int rows, rb; c.plane_dims(i, &rows, &rb);
        if (linesizes[i] < rb) return fail(FFV1B200_ERR_EINVAL, "linesize smaller than a row");
        if (linesizes[i] & 15) direct = false;
    }
    for (int f = 1; f < nframes && direct; f++)
        for (int i = 0; i < np; i++) if (linesizes[f * 4 + i] != linesizes[i]) direct = false;
    size_t plane_off[4] = {0, 0, 0, 0}, off = 0;
    int pitch[4] = {0, 0, 0, 0};
    for (int i = 0; i < np; i++) {
        int rows, rb; c.plane_dims(i, &rows, &rb);
        pitch[i] = direct ? linesizes[i] : ((rb + 127) & ~127);
        plane_off[i] = off;
        off += (size_t)pitch[i] * rows;
    }
    const size_t stride = (off + 255) & ~(size_t)255;
    if (sl.d_in.n < stride * (size_t)e->max_batch) CU_TRY(sl.d_in.alloc(stride * (size_t)e->max_batch));   // the slot is idle

    cudaStream_t s = e->s_in;
    sl.h2d_bytes = 0;
    const uint8_t *run_src = nullptr; uint8_t *run_dst = nullptr; size_t run_len = 0;
    auto flush = [&]() -> cudaError_t {
        cudaError_t ce = cudaSuccess;
        if (run_len) ce = cudaMemcpyAsync(run_dst, run_src, run_len, cudaMemcpyHostToDevice, s);
        sl.h2d_bytes += (int64_t)run_len;
        run_len = 0;
        return ce;
    };
    for (int f = 0; f < nframes; f++)
        for (int i = 0; i < 4; i++) {
            const uint8_t *dptr = nullptr;
            if (i < np) {
                const uint8_t *src = planes[f * 4 + i];
                if (!src) return fail(FFV1B200_ERR_EINVAL, "missing plane pointer");
                int rows, rb; c.plane_dims(i, &rows, &rb);
                uint8_t *dst = sl.d_in.p + (size_t)f * stride + plane_off[i];
                if (direct) {
                    const size_t bytes = (size_t)pitch[i] * (rows - 1) + rb;      // never read past the last row's samples
                    const size_t span = (size_t)pitch[i] * rows;
                    if (run_len && src == run_src + run_len && dst == run_dst + run_len) run_len += bytes;
                    else { CU_TRY(flush()); run_src = src; run_dst = dst; run_len = bytes; }
                    // planes that really are adjacent on both sides continue the run across the row padding
                    const bool more = i + 1 < np ? planes[f * 4 + i + 1] == src + span
                                                 : (f + 1 < nframes && planes[(f + 1) * 4] == src + span && stride == off);
                    if (more) run_len = run_len - bytes + span;
                } else {
                    CU_TRY(flush());
                    CU_TRY(cudaMemcpy2DAsync(dst, pitch[i], src, linesizes[f * 4 + i], rb, rows, cudaMemcpyHostToDevice, s));
                    sl.h2d_bytes += (int64_t)rb * rows;
                }
                dptr = dst;
            }
            sl.h_planes.p[f * 4 + i] = dptr;
        }
    CU_TRY(flush());
    CU_TRY(cudaEventRecord(sl.ev_h2d, s));
    for (int i = 0; i < 4; i++) sl.ls[i] = pitch[i];

    plan_batch(e, sl, nframes, e->picture_number);
    sl.carry_in = e->carry_next;
    sl.fast = pixel_fast_ok(e, sl);
    if (!sl.d_out.p) CU_TRY(sl.d_out.alloc((size_t)e->max_batch * ((size_t)c.frame_bytes() / 2 + 65536)));
    sl.out = sl.d_out.p; sl.out_cap = sl.d_out.n;
    CU_TRY(cudaStreamWaitEvent(e->s_comp, sl.ev_h2d, 0));
    int r = enqueue_kernels(e, sl, e->s_comp);
    if (r < 0) return r;
    sl.busy = true;
    e->picture_number += nframes;
    e->carry_next = (e->carry_next + 1) % kCarry;
    e->submitted++;
    return nframes;
}

int ffv1b200_enc_sync_output(FFV1B200Encoder *e)
{
    if (!e) return fail(FFV1B200_ERR_EINVAL, "null argument");
    CU_TRY(cudaSetDevice(e->device));
    return finish_output(e);
}

int ffv1b200_enc_collect_async(FFV1B200Encoder *e, uint8_t *out, size_t out_cap, FFV1B200Packet *pkts, size_t *needed)
{
    if (!e || !out || !pkts) return fail(FFV1B200_ERR_EINVAL, "null argument");
    if (e->collected == e->submitted) return fail(FFV1B200_ERR_EINVAL, "no batch in flight");
    CU_TRY(cudaSetDevice(e->device));
    { int r = finish_output(e); if (r < 0) return r; }          // the previous batch's bytes are complete from here on
    Slot &sl = e->slot[e->collected % kSlots];
    CU_TRY(cudaEventSynchronize(sl.ev_small));
    if (sl.h_status.p[0] | sl.h_status.p[1] | sl.h_status.p[2]) {
        int r = recover(e, true, e->s_comp);
        if (r < 0) return r;
    }
    const size_t total = (size_t)sl.h_pkt_off.p[sl.nframes];
    if (needed) *needed = total;
    if (total > out_cap) {
        set_last_error("output buffer too small: need " + std::to_string(total) + " bytes");
        return FFV1B200_ERR_BUFFER_TOO_SMALL;                   // the batch stays collectable
    }
    // the copy runs on its own stream: the caller can submit the next batch (whose host->device copies use the other
    // direction of the link) while the packets are still on their way
    CU_TRY(cudaMemcpyAsync(out, sl.out, total, cudaMemcpyDeviceToHost, e->s_out));
    CU_TRY(cudaEventRecord(sl.ev_d2h, e->s_out));
    sl.d2h_pending = true;
    e->stats.d2h_bytes += (int64_t)total;
    account(e, sl);
    fill_packets(sl, pkts);
    sl.busy = false;
    e->last_slot = (int)(e->collected % kSlots);
    e->collected++;
    return sl.nframes;
}

int ffv1b200_enc_collect(FFV1B200Encoder *e, uint8_t *out, size_t out_cap, FFV1B200Packet *pkts, size_t *needed)
{
    const int r = ffv1b200_enc_collect_async(e, out, out_cap, pkts, needed);
    if (r < 0) return r;
    const int q = finish_output(e);
    return q < 0 ? q : r;
}

int ffv1b200_enc_pending(const FFV1B200Encoder *e)
{
    return e ? (int)(e->submitted - e->collected) : FFV1B200_ERR_EINVAL;
}

int ffv1b200_enc_encode_host(FFV1B200Encoder *e, int nframes, const uint8_t *const *planes, const int *linesizes,
                             uint8_t *out, size_t out_cap, FFV1B200Packet *pkts, size_t *needed)
{
    if (!e || !out || !pkts) return fail(FFV1B200_ERR_EINVAL, "null argument");
    // a previous call that failed with BUFFER_TOO_SMALL left its batch in flight: that one is handed out first
    if (e->submitted == e->collected) {
        int r = ffv1b200_enc_submit_host(e, nframes, planes, linesizes);
        if (r < 0) return r;
    }
    return ffv1b200_enc_collect(e, out, out_cap, pkts, needed);
}

int ffv1b200_enc_encode_device(FFV1B200Encoder *e, int nframes, const void *const *d_planes, const int *linesizes,
                               void *d_out, size_t d_out_cap, FFV1B200Packet *pkts, size_t *needed, void *stream)
{
    if (!e || !d_planes || !linesizes || !d_out || !pkts) return fail(FFV1B200_ERR_EINVAL, "null argument");
    if (nframes < 1 || nframes > e->max_batch) return fail(FFV1B200_ERR_EINVAL, "nframes outside 1..max_batch_frames");
    if (e->submitted != e->collected) return fail(FFV1B200_ERR_EINVAL, "collect the batches in flight first");
    CU_TRY(cudaSetDevice(e->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : e->s_comp;
    Slot &sl = e->slot[e->submitted % kSlots];
    { int r = upload_prefixes(e, sl, s); if (r < 0) return r; }
    for (int i = 0; i < 4; i++) sl.ls[i] = linesizes[i];
    for (int f = 0; f < nframes; f++)
        for (int i = 0; i < 4; i++) {
            if (linesizes[f * 4 + i] != sl.ls[i]) return fail(FFV1B200_ERR_EINVAL, "all frames of a batch must share linesizes");
            sl.h_planes.p[f * 4 + i] = (const uint8_t *)d_planes[f * 4 + i];
        }
    plan_batch(e, sl, nframes, e->picture_number);
    sl.carry_in = e->carry_next;
    sl.fast = pixel_fast_ok(e, sl);
    sl.out = (uint8_t *)d_out; sl.out_cap = d_out_cap;
    sl.h2d_bytes = 0;
    int r = enqueue_kernels(e, sl, s);
    if (r < 0) return r;
    e->submitted++;                                             // so that recover() sees the batch
    CU_TRY(cudaEventSynchronize(sl.ev_small));
    if (sl.h_status.p[0] | sl.h_status.p[1] | sl.h_status.p[2]) {
        r = recover(e, false, s);
        if (r < 0) {
            e->submitted--;                                     // state unchanged: the call may be repeated
            if (r == FFV1B200_ERR_BUFFER_TOO_SMALL && needed) *needed = (size_t)sl.h_status.p[2];
            return r;
        }
    }
    if (needed) *needed = (size_t)sl.h_pkt_off.p[nframes];
    account(e, sl);
    fill_packets(sl, pkts);
    e->last_slot = (int)((e->submitted - 1) % kSlots);
    e->collected++;
    e->picture_number += nframes;
    e->carry_next = (e->carry_next + 1) % kCarry;
    return nframes;
}

int ffv1b200_enc_encode_cuda(FFV1B200Encoder *e, int nframes, const void *const *d_planes, const int *linesizes,
                             uint8_t *out, size_t out_cap, FFV1B200Packet *pkts, size_t *needed)
{
    if (!e || !out || !pkts) return fail(FFV1B200_ERR_EINVAL, "null argument");
    if (e->submitted != e->collected) return fail(FFV1B200_ERR_EINVAL, "collect the batches in flight first");
    CU_TRY(cudaSetDevice(e->device));
    { int q = finish_output(e); if (q < 0) return q; }
    size_t total = 0;
    int r;
    if (e->cuda_pending) {
        // the previous call coded this batch but the caller's buffer was too small: only the hand-out is repeated
        if (nframes != e->cuda_pending) return fail(FFV1B200_ERR_EINVAL, "repeat the call that returned BUFFER_TOO_SMALL first");
        Slot &sl = e->slot[e->last_slot];
        total = (size_t)sl.h_pkt_off.p[sl.nframes];
        fill_packets(sl, pkts);
        r = nframes;
    } else {
        Slot &sl = e->slot[e->submitted % kSlots];
        if (!sl.d_out.p) CU_TRY(sl.d_out.alloc((size_t)e->max_batch * ((size_t)e->cfg.frame_bytes() / 2 + 65536)));
        for (;;) {                                              // the packets are assembled in the slot's own device buffer
            r = ffv1b200_enc_encode_device(e, nframes, d_planes, linesizes, sl.d_out.p, sl.d_out.n, pkts, &total, nullptr);
            if (r != FFV1B200_ERR_BUFFER_TOO_SMALL) break;
            CU_TRY(sl.d_out.alloc(total + 4096));
        }
        if (r < 0) return r;
    }
    Slot &sl = e->slot[e->last_slot];
    if (needed) *needed = total;
    if (total > out_cap) {
        e->cuda_pending = nframes;                              // coded; the same call with a larger buffer hands it out
        return fail(FFV1B200_ERR_BUFFER_TOO_SMALL, "output buffer too small: need " + std::to_string(total) + " bytes");
    }
    e->cuda_pending = 0;
    CU_TRY(cudaMemcpyAsync(out, sl.d_out.p, total, cudaMemcpyDeviceToHost, e->s_out));
    CU_TRY(cudaStreamSynchronize(e->s_out));
    e->stats.d2h_bytes += (int64_t)total;
    return r;
}

int ffv1b200_enc_stats(const FFV1B200Encoder *e, FFV1B200EncStats *s)
{
    if (!e || !s) return FFV1B200_ERR_EINVAL;
    *s = e->stats;
    return 0;
}

int ffv1b200_enc_stats_out(FFV1B200Encoder *e, char *buf, size_t cap, size_t *needed)
{
    if (!e || !buf) return fail(FFV1B200_ERR_EINVAL, "null argument");
    if (!(e->cfg.pass_flags & kPass1)) return fail(FFV1B200_ERR_EINVAL, "the encoder was not opened with FFV1B200_FLAG_PASS1");
    if (e->submitted != e->collected) return fail(FFV1B200_ERR_EINVAL, "collect the batches in flight first");
    CU_TRY(cudaSetDevice(e->device));
    CU_TRY(cudaStreamSynchronize(e->s_comp));
    PassStats st;
    const Config &c = e->cfg;
    for (int i = 0; i < 2; i++) st.rc_stat2[i].assign((size_t)c.context_count[i] * 64, 0);
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "counter width");
    CU_TRY(cudaMemcpy(st.rc_stat, e->d_rc_stat.p, sizeof(st.rc_stat), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(st.rc_stat2[c.context_model].data(), e->d_rc_stat2.p, st.rc_stat2[c.context_model].size() * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    st.gob_count = e->gob_count;
    const std::string text = format_stats(c, st);
    if (needed) *needed = text.size() + 1;
    if (text.size() + 1 > cap) return fail(FFV1B200_ERR_BUFFER_TOO_SMALL, "stats buffer too small: need " + std::to_string(text.size() + 1) + " bytes");
    memcpy(buf, text.c_str(), text.size() + 1);
    return (int)text.size();
}

int64_t ffv1b200_enc_debug_records(FFV1B200Encoder *e, int frame, int slice, uint32_t *dst, int64_t cap)
{
    if (!e || !dst) return FFV1B200_ERR_EINVAL;
    if (e->submitted != e->collected) return fail(FFV1B200_ERR_EINVAL, "collect the batches in flight first");
    if (frame < 0 || frame >= e->slot[e->last_slot].nframes || slice < 0 || slice >= e->cfg.slice_count()) return fail(FFV1B200_ERR_EINVAL, "bad frame/slice");
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
