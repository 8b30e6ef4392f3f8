// ffv1_replay_fused.cu -- adaptive-state replay in ONE kernel (small context model, range coder; sm_100a).
// EXPERIMENTAL (FFV1B200_REPLAY=fused): bit-exact on the whole test matrix, but at 168 ms per 1024 1080p frames it is
// slower than the list kernels of ffv1_ctx_replay.cu (124 ms): the tile-local lists are short (17 symbols on average),
// so the per-list and per-block set-up dominates and the kernel is bound by instruction issue (81 % issue-active,
// 146 G warp instructions; profiles/r01_replay_fused.txt).  Kept as the starting point for a version with larger tiles.
//
// Same job as the k_ctx_hist / k_ctx_scan / k_dec_layout / k_ctx_scatter / k_replay_* chain of ffv1_ctx_replay.cu --
// turn the (context, residual) records of k_pixel into the (probability state, bit) decision stream k_rangecode codes,
// i.e. the state half of put_symbol_inline / put_rac (ffv1enc.c:185-231, rangecoder.h:92-99) -- but without ever
// building per-context lists in global memory:
//
//   * one CTA per (GOP segment, slice, plane context) chain; the chain's model (ctx_count x 32 state bytes) stays in
//     shared memory for the whole chain;
//   * the chain's samples are walked in coding order, a TILE (<= 16 segments of <= 128 samples) at a time.  Inside a
//     tile the CTA does a stable counting sort by context in shared memory (per-segment histograms, one warp per
//     segment, ballot ranks inside a 32-sample group), so every context gets its symbols of the tile in coding order;
//   * then 16-lane groups replay the lists (longest first), lane = state slot as in k_replay_grp: a symbol with
//     exponent e <= 4 touches at most 15 slots, each lane keeps its slot's state in a register, and a symbol costs
//     ~10 instructions per group.  (One lane per context with all loads of a symbol in flight was measured first: a
//     lone warp issues an instruction every ~4 cycles, so ~135 instructions per symbol made the longest list of a
//     tile -- the critical path of the CTA -- 13 x slower than the 16-lane scheme.);
//   * the decisions of the tile are staged in shared memory at their final positions and written to the decision
//     region with coalesced 16-byte stores (a 2-byte scattered store per decision is what bounded the list kernels).
//
// The decision-region layout (lines in coding order, runs 16-byte aligned, run counts for k_rangecode) is produced on
// the fly: the CTA owns the region of its (frame, slice, plane context) and keeps the running position itself.
#include "ffv1_enc_kernels.cuh"
#include <algorithm>
#include <cstdlib>

namespace ffv1 {

constexpr int kFuThreads  = 512;
constexpr int kFuSegSym   = 128;                         // samples per segment (four 32-sample groups, one warp)
constexpr int kFuTileSegs = 16;                          // segments per tile (one per warp)
constexpr int kFuTileSym  = kFuSegSym * kFuTileSegs;
constexpr int kFuStage    = 15360;                       // decision entries staged per tile (7.5 per sample on average)
constexpr int kFuRow      = 36;                          // bytes per state row in shared memory (32 + padding against bank conflicts)
constexpr int kFuClasses  = 2;                           // lists of >= 32 symbols are dealt first
constexpr int kFuMaxCtx   = 1024;

__device__ __forceinline__ uint32_t fu_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t n = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

// lanes (among `active`) holding the same 10-bit key as the caller
__device__ __forceinline__ uint32_t fu_same_key(uint32_t key, uint32_t active)
{
    uint32_t grp = active;
#pragma unroll
    for (int b = 0; b < 10; b++) {
        const bool bit = (key >> b) & 1u;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, bit);
        grp &= bit ? m : ~m;
    }
    return grp;
}

__device__ __forceinline__ uint32_t fu_decisions_of(int d)
{
    return d ? (uint32_t)(2 * (31 - __clz((uint32_t)abs(d))) + 3) : 1u;       // put_symbol_inline: 1 or 2e+3
}

// put_symbol_inline (ffv1enc.c:185-231) decision by decision (any magnitude; the rare large residuals)
template <typename OUT>
__device__ __noinline__ void fu_symbol_serial(uint8_t *row, const uint8_t *lut, OUT *o, int d)
{
    const uint32_t a = (uint32_t)abs(d);
    const int ex = 31 - __clz(a);
    const int nd = 2 * ex + 3;
    for (int qi = 0; qi < nd; qi++) {
        int sl; uint32_t bit;
        if (qi == 0) { sl = 0; bit = 0u; }
        else if (qi <= ex) { sl = 1 + min(qi - 1, 9); bit = 0x100u; }
        else if (qi == ex + 1) { sl = 1 + min(ex, 9); bit = 0u; }
        else if (qi <= 2 * ex + 1) { const int i = ex - 1 - (qi - ex - 2); sl = 22 + min(i, 9); bit = ((a >> i) & 1u) << 8; }
        else { sl = 11 + min(ex, 10); bit = d < 0 ? 0x100u : 0u; }
        const uint32_t s = row[sl];
        o[qi] = (uint16_t)(s | bit);
        row[sl] = lut[bit + s];
    }
}

// roles of the 16 lanes of a group (EMAX = 4, |residual| < 32): lane 0 slot 0 ("is zero"), lanes 1..5 slots 1..5
// (exponent), lanes 6..9 slots 25..22 (mantissa bits 3..0), lanes 10..14 slots 11..15 (sign) -- the visiting order of
// put_symbol_inline (ffv1enc.c:202-229) is increasing in lane index
template <int EMAX>
__device__ __forceinline__ uint32_t fu_symbol_masks(int d, bool &slow)
{
    slow = false;
    if (d == 0) return 1u | (1u << 16);
    const uint32_t a = (uint32_t)abs(d);
    const int e = 31 - __clz(a);
    if (e > EMAX) { slow = true; return 0u; }
    const uint32_t ones_e = (1u << e) - 1u;
    const uint32_t mant = e ? (__brev(a & ones_e) >> (32 - e)) : 0u;       // bit t <- bit e-1-t of |d|
    const uint32_t visit = 1u | (((2u << e) - 1u) << 1) | (ones_e << (2 * EMAX + 2 - e)) | (1u << (2 * EMAX + 2 + e));
    const uint32_t bits = (ones_e << 1) | (mant << (2 * EMAX + 2 - e)) | ((d < 0 ? 1u : 0u) << (2 * EMAX + 2 + e));
    return visit | (bits << 16);
}

// P4 of k_replay_fused.  DIRECT: the tile has more decisions than the staging area holds, they go straight to the
// decision region (generic stores).
template <int EMAX, bool DIRECT>
__device__ __forceinline__ void fu_replay_lists(uint8_t *s_state, const uint8_t *s_lut, const uint16_t *s_cls, const uint16_t *s_cnt,
                                                const uint16_t *s_start, const uint32_t *s_ent, uint2 *s_blk_all, int *s_next,
                                                const int *s_ncls, const int nctx_pad, uint16_t *out)
{
    constexpr int G = 16;
    static_assert(3 * EMAX + 3 <= G, "roles must fit a group");
    const int tid = threadIdx.x, lane = tid & 31, g = lane & (G - 1);
    const uint32_t gmask = 0xFFFFu << (lane & 16);
    uint2 *s_blk = s_blk_all + (tid & ~(G - 1));
    const int slot = g == 0 ? 0 : (g <= EMAX + 1 ? g : (g <= 2 * EMAX + 1 ? 22 + (2 * EMAX + 1 - g) : (g <= 3 * EMAX + 2 ? 11 + g - (2 * EMAX + 2) : -1)));
    const bool has_slot = slot >= 0;
    const uint32_t lanebit = 1u << g, lt_mask = lanebit - 1u;
    const int rot = (g + 16 - 8) & 31;                               // rotr(visit | bits << 16, rot) puts my coded bit at bit 8
    const int n0 = s_ncls[0], nl = n0 + s_ncls[1];
    const uint32_t lut_s = (uint32_t)__cvta_generic_to_shared(s_lut);
    uint32_t n_left = 0u, st = 0u;
    const uint32_t *ep = s_ent;
    uint8_t *row = s_state;
    bool exhausted = false;
    for (;;) {
        if (n_left == 0u && !exhausted) {
                    int oi = 0;
            if (g == 0) oi = atomicAdd(s_next, 1);
            oi = __shfl_sync(gmask, oi, 0, G);
            if (oi >= nl) exhausted = true;
            else {
                const int c = oi < n0 ? s_cls[oi] : s_cls[nctx_pad + oi - n0];
                n_left = s_cnt[c];
                ep = s_ent + s_start[c];
                row = s_state + c * kFuRow;
                st = has_slot ? row[slot] : 0u;
            }
        }
        if (!__any_sync(0xFFFFFFFFu, n_left != 0u)) break;
        // ---- a block of up to G symbols per group: masks and positions, one symbol per lane
        const uint32_t m = min((uint32_t)G, n_left);
        uint32_t en = 0u, vb = 0u;
        bool slow = false;
        if ((uint32_t)g < m) { en = ep[g]; vb = fu_symbol_masks<EMAX>((int)en >> 20, slow); }
        __syncwarp();
        s_blk[g] = make_uint2(vb, (en & 0xFFFFFu) | (slow ? 0x80000000u : 0u));
        __syncwarp();
        const uint32_t mm = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, 16));
        if (!__any_sync(0xFFFFFFFFu, slow)) {
#pragma unroll 4
            for (uint32_t k = 0; k < mm; k++) {
                const uint2 q = s_blk[k];
                const uint32_t idx = q.y + __popc(q.x & lt_mask);
                const uint32_t val = (__funnelshift_r(q.x, q.x, rot) & 0x100u) | st;
                if (q.x & lanebit) {
                    out[idx] = (uint16_t)val;
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(st) : "r"(lut_s + val));
                }
            }
        } else {
            for (uint32_t k = 0; k < mm; k++) {
                const uint2 q = s_blk[k];
                if (q.y >> 31) {                                 // uniform inside a group: a large residual, decision by decision
                    if (has_slot) row[slot] = (uint8_t)st;
                    __syncwarp(gmask);
                    if (g == 0) fu_symbol_serial(row, s_lut, out + (q.y & 0xFFFFFu), (int)ep[k] >> 20);
                    __syncwarp(gmask);
                    if (has_slot) st = row[slot];
                } else if (q.x & lanebit) {
                    const uint32_t val = (__funnelshift_r(q.x, q.x, rot) & 0x100u) | st;
                    out[q.y + __popc(q.x & lt_mask)] = (uint16_t)val;
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(st) : "r"(lut_s + val));
                }
            }
        }
        ep += m; n_left -= m;
        if (n_left == 0u && m != 0u && has_slot) row[slot] = (uint8_t)st;      // the list is done: its state goes back to the model
    }
}

struct FusedParams {
    const FusedSeg *segs;
    const FusedSlice *slices;
};

template <int EMAX>
__global__ void __launch_bounds__(kFuThreads, 2) k_replay_fused(const EncDeviceTables T, const EncBatch B, const FusedParams P)
{
    extern __shared__ __align__(16) unsigned char fu_smem[];
    __shared__ uint8_t s_lut[512];
    __shared__ uint32_t s_segnd[kFuTileSegs], s_segstart[kFuTileSegs], s_wtot[kFuThreads / 32];
    __shared__ int s_ncls[kFuClasses];
    __shared__ int s_next;
    __shared__ uint32_t s_tile[4];               // aligned base, first entry, end of the tile's decisions; flags
    const Layout &L = T.layout;
    const int nctx = L.ctx_count, nctx_pad = (nctx + 7) & ~7;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint8_t *s_state = fu_smem;                                                        // [nctx][kFuRow] the chain's model
    uint16_t *s_start = reinterpret_cast<uint16_t *>(fu_smem + ((nctx * kFuRow + 15) & ~15));   // [nctx] list start inside s_ent
    uint16_t *s_cnt = s_start + nctx_pad;                                              // [nctx] symbols of the tile per context
    uint16_t *s_cls = s_cnt + nctx_pad;                                                // [classes][nctx] contexts by list length
    uint32_t *s_ent = reinterpret_cast<uint32_t *>(s_cls + kFuClasses * nctx_pad);     // [kFuTileSym] stage position | residual << 20
    uint16_t *s_stage = reinterpret_cast<uint16_t *>(s_ent + kFuTileSym);              // [kFuStage + 8] decisions of the tile
    uint32_t *s_wh = reinterpret_cast<uint32_t *>(s_stage);                            // [8][nctx] per-segment counts, 16 bits each (aliases s_stage)
    uint2 *s_blk_all = reinterpret_cast<uint2 *>(s_stage + kFuStage + 8);              // [threads] per group: 16 x (visit | bits << 16, position | slow << 31)

    for (int i = tid; i < 512; i += kFuThreads) s_lut[i] = T.trans_lut[i];
    if (B.status[0]) return;
    const int chain = blockIdx.x;
    const int pc = chain % L.npc, s = (chain / L.npc) % L.nslices, seg = chain / (L.npc * L.nslices);
    const SliceGeom &g = T.slices[s];
    const int f0 = B.seg_first[seg], f1 = B.seg_first[seg + 1];
    const bool key = B.frame_key[f0] != 0;
    const bool hand_over = f1 == B.nframes;
    const size_t coff = ((size_t)s * L.npc + pc) * ((size_t)nctx * 32);
    const int nsegs = P.slices[s].seg_count[pc];
    const FusedSeg *segs = P.segs + P.slices[s].seg_first[pc];
    if (nsegs == 0) return;
    // ---- the chain's model: 128 on keyframes (ffv1.c:177-202), else what the previous batch left
    {
        const uint32_t *cin = reinterpret_cast<const uint32_t *>(B.carry_in + coff);
        for (int i = tid; i < nctx * 8; i += kFuThreads)
            *reinterpret_cast<uint32_t *>(s_state + (i >> 3) * kFuRow + (i & 7) * 4) = key ? 0x80808080u : cin[i];
    }
    for (int i = tid; i < 8 * nctx; i += kFuThreads) s_wh[i] = 0u;
    if (tid < kFuClasses) s_ncls[tid] = 0;
    if (tid == 0) s_next = 0;
    __syncthreads();

    unsigned long long ndec_chain = 0;           // thread 0: real decisions of the chain
    for (int f = f0; f < f1; f++) {
        const uint32_t *rec_slice = B.rec + (size_t)f * L.rec_per_frame + g.rec_first;
        uint16_t *dec_pc = B.dec + (size_t)f * L.dec_per_frame + g.dec_off[pc];
        uint32_t *run_cnt = B.run_cnt + (size_t)f * L.runs_per_frame + g.run_first;
        // thread 0 keeps the layout of the (frame, slice, plane context) decision region
        uint32_t pos = 0, run_start = 0;
        int cur_run = segs[0].run;
        for (int t0 = 0; t0 < nsegs; t0 += kFuTileSegs) {
            const int nsg = min(kFuTileSegs, nsegs - t0);
            // ---- P1: one warp per segment: records -> registers, per-segment context histogram, decisions of the segment
            uint32_t rr[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
            int sw = 0;
            if (warp < nsg) {
                const FusedSeg sgm = segs[t0 + warp];
                sw = sgm.w;
                const uint32_t *recp = rec_slice + sgm.rec_off;
#pragma unroll
                for (int k = 0; k < 4; k++) if (k * 32 + lane < sw) rr[k] = recp[k * 32 + lane];
                uint32_t nd = 0;
                uint32_t *wh = s_wh + (warp >> 1) * nctx;
                const uint32_t one = (warp & 1) ? 0x10000u : 1u;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (k * 32 + lane < sw) {
                        nd += fu_decisions_of((int)(int16_t)(rr[k] & 0xFFFFu));
                        atomicAdd(&wh[rr[k] >> 16], one);
                    }
                nd = __reduce_add_sync(0xFFFFFFFFu, nd);
                if (lane == 0) s_segnd[warp] = nd;
            }
            __syncthreads();
            // ---- P2: per context (two per thread): counts of the segments -> exclusive offsets (in place) and the total;
            //          thread 0 lays out the tile's decisions
            uint32_t cc[2] = {0u, 0u};
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int c = 2 * tid + j;
                if (c < nctx) {
                    uint32_t run = 0;
#pragma unroll
                    for (int h = 0; h < 8; h++) {
                        const uint32_t v = s_wh[h * nctx + c];
                        const uint32_t lo = v & 0xFFFFu, hi = v >> 16;
                        s_wh[h * nctx + c] = run | ((run + lo) << 16);
                        run += lo + hi;
                    }
                    s_cnt[c] = (uint16_t)run;
                    cc[j] = run;
                }
            }
            const uint32_t pair = cc[0] + cc[1];
            const uint32_t incl = fu_incl_scan(pair, lane);
            if (lane == 31) s_wtot[warp] = incl;
            if (tid == 0) {
                uint32_t p = pos;
                for (int w = 0; w < nsg; w++) {
                    const int run = segs[t0 + w].run;
                    if (run != cur_run) {                    // a run of this plane context ended: the next one starts 16 B aligned
                        run_cnt[cur_run] = p - run_start;
                        p = (p + 7u) & ~7u;
                        run_start = p; cur_run = run;
                    }
                    s_segstart[w] = p;
                    p += s_segnd[w];
                    ndec_chain += s_segnd[w];
                }
                const uint32_t tstart = s_segstart[0];
                s_tile[0] = tstart & ~7u; s_tile[1] = tstart; s_tile[2] = p;
                s_tile[3] = (p - (tstart & ~7u) > (uint32_t)kFuStage ? 1u : 0u) | (p + 8u > g.dec_cap[pc] ? 2u : 0u);
                pos = p;
            }
            __syncthreads();
            if (s_tile[3] & 2u) {                            // the decision region is too small: the host grows it and retries
                if (tid == 0) {
                    const unsigned long long ns = g.pc_samples[pc];
                    atomicMax(&B.status[0], ((unsigned long long)g.dec_cap[pc] * 333ull + ns - 1) / ns + 1ull);   // 1.3 x the current entries per sample, x 256
                }
                return;
            }
            {
                uint32_t base = incl - pair;
                for (int w = 0; w < warp; w++) base += s_wtot[w];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int c = 2 * tid + j;
                    if (c < nctx) {
                        s_start[c] = (uint16_t)base;
                        const uint32_t n = cc[j];
                        if (n) {
                            const int cls = n >= 32u ? 0 : 1;
                            s_cls[cls * nctx_pad + atomicAdd(&s_ncls[cls], 1)] = (uint16_t)c;
                        }
                        base += n;
                    }
                }
            }
            __syncthreads();
            // ---- P3: stable placement: entry = position of the symbol's first decision | residual
            const uint32_t tile_base = s_tile[0];
            const bool direct = s_tile[3] & 1u;              // more decisions than the staging area holds: write them directly
            if (warp < nsg) {
                uint32_t run = s_segstart[warp] - tile_base;
                uint32_t *wh = s_wh + (warp >> 1) * nctx;
                const int sh = (warp & 1) * 16;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (k * 32 >= sw) break;
                    const bool act = k * 32 + lane < sw;
                    const uint32_t ctx = act ? rr[k] >> 16 : 0u;
                    const int d = (int)(int16_t)(rr[k] & 0xFFFFu);
                    const uint32_t nd = act ? fu_decisions_of(d) : 0u;
                    const uint32_t in = fu_incl_scan(nd, lane);
                    const uint32_t grp = fu_same_key(ctx, __ballot_sync(0xFFFFFFFFu, act));
                    const uint32_t rank = __popc(grp & lt_mask);
                    uint32_t off = 0u;
                    if (act && rank == 0u) off = (atomicAdd(&wh[ctx], (uint32_t)__popc(grp) << sh) >> sh) & 0xFFFFu;
                    off = __shfl_sync(0xFFFFFFFFu, off, (__ffs(grp) - 1) & 31);
                    if (act) s_ent[(uint32_t)s_start[ctx] + off + rank] = (run + in - nd) | ((uint32_t)d << 20);
                    run += __shfl_sync(0xFFFFFFFFu, in, 31);
                }
            }
            __syncthreads();
            // ---- P4: 16-lane groups take the lists (long ones first) from a shared counter; lane = state slot
            if (!direct) fu_replay_lists<EMAX, false>(s_state, s_lut, s_cls, s_cnt, s_start, s_ent, s_blk_all, &s_next, s_ncls, nctx_pad, s_stage);
            else         fu_replay_lists<EMAX, true>(s_state, s_lut, s_cls, s_cnt, s_start, s_ent, s_blk_all, &s_next, s_ncls, nctx_pad, dec_pc + tile_base);
            __syncthreads();
            // ---- P5: staged decisions -> decision region (16-byte vectors; ragged ends entry by entry)
            if (!direct) {
                const uint32_t tstart = s_tile[1], tend = s_tile[2];
                const uint32_t v0 = (tstart + 7u) & ~7u, v1 = tend & ~7u;
                for (uint32_t v = v0 + (uint32_t)tid * 8u; v < v1; v += kFuThreads * 8u)
                    *reinterpret_cast<uint4 *>(dec_pc + v) = *reinterpret_cast<const uint4 *>(s_stage + (v - tile_base));
                const uint32_t hend = min(v0, tend);
                if (tstart + (uint32_t)tid < hend) dec_pc[tstart + tid] = s_stage[tstart + tid - tile_base];
                if (v1 >= v0 && v1 + (uint32_t)tid < tend) dec_pc[v1 + tid] = s_stage[v1 + tid - tile_base];
            }
            __syncthreads();
            for (int i = tid; i < 8 * nctx; i += kFuThreads) s_wh[i] = 0u;
            if (tid < kFuClasses) s_ncls[tid] = 0;
            if (tid == 0) s_next = 0;
            __syncthreads();
        }
        if (tid == 0) run_cnt[cur_run] = pos - run_start;
    }
    if (tid == 0 && ndec_chain) atomicAdd(&B.status[3], ndec_chain);
    // ---- hand the model to the next batch when this segment runs to the end of the batch
    if (hand_over) {
        uint32_t *cout = reinterpret_cast<uint32_t *>(B.carry_out + coff);
        for (int i = tid; i < nctx * 8; i += kFuThreads)
            cout[i] = *reinterpret_cast<const uint32_t *>(s_state + (i >> 3) * kFuRow + (i & 7) * 4);
    }
}

// ------------------------------------------------------------------------------------------------ host side
bool fused_replay_supported(const Layout &L)
{
    // 8-bit content (residuals folded to <= 9 bits): exponents above 4 are rare enough for the serial path
    return !L.golomb && L.ctx_count <= kFuMaxCtx && L.coded_bits <= 9;
}

int fused_replay_smem_bytes(const Layout &L)
{
    const int nctx = L.ctx_count, nctx_pad = (nctx + 7) & ~7;
    return ((nctx * kFuRow + 15) & ~15) + (2 + kFuClasses) * nctx_pad * 2 + kFuTileSym * 4 + (kFuStage + 8) * 2 + kFuThreads * 8;
}

// segments of <= kFuSegSym samples, per (slice, plane context) in coding order
void build_fused_plan(const Tables &tab, FusedPlan &plan)
{
    const Layout &L = tab.layout;
    plan.segs.clear();
    plan.slices.assign(tab.slices.size(), FusedSlice());
    plan.ok = fused_replay_supported(L);
    for (size_t si = 0; si < tab.slices.size(); si++) {
        const SliceGeom &g = tab.slices[si];
        if (g.nruns > 65535) plan.ok = false;
        for (int pc = 0; pc < 3; pc++) {
            plan.slices[si].seg_first[pc] = (int32_t)plan.segs.size();
            for (int li = 0; li < g.pc_nlines[pc]; li++) {
                const LineDesc &ld = tab.lines[g.line_first + tab.pc_lines[g.pc_line_first[pc] + li]];
                for (int x0 = 0; x0 < ld.w; x0 += kFuSegSym) {
                    FusedSeg sg;
                    sg.rec_off = ld.rec_off + (uint32_t)x0;
                    sg.w = (uint16_t)std::min(kFuSegSym, (int)ld.w - x0);
                    sg.run = (uint16_t)ld.run;
                    plan.segs.push_back(sg);
                }
            }
            plan.slices[si].seg_count[pc] = (int32_t)plan.segs.size() - plan.slices[si].seg_first[pc];
        }
    }
    plan.smem_bytes = fused_replay_smem_bytes(L);
}

cudaError_t configure_fused_replay(const FusedPlan &plan)
{
    return cudaFuncSetAttribute(k_replay_fused<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes);
}

void launch_fused_replay(const EncDeviceTables &t, const EncBatch &b, const FusedPlan &plan, const FusedSeg *d_segs,
                         const FusedSlice *d_slices, cudaStream_t s)
{
    const Layout &L = t.layout;
    const int nchains = b.nseg * L.nslices * L.npc;
    FusedParams P;
    P.segs = d_segs; P.slices = d_slices;
    k_replay_fused<4><<<nchains, kFuThreads, plan.smem_bytes, s>>>(t, b, P);
}

} // namespace ffv1
