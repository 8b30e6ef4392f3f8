// ffv1_ctx_replay.cu -- adaptive-state replay decomposed by CONTEXT (small context model; sm_100a).
//
// put_rac's state update (rangecoder.h:92-99) depends only on the coded bit, and the 32-byte state row of a context is
// touched by nothing but the symbols coded in that context (ffv1enc.c:311-321).  Inside one
// (GOP segment, slice, plane context) chain the symbols of DIFFERENT contexts are therefore independent of each other:
// the only sequential dependency is along the list of symbols that share a context.  This file turns the record stream
// into those lists and replays every list with its own thread:
//
//   k_ctx_hist     per context tile (<= 16 lines of one (slice, plane context)): context histogram + decisions per line
//   k_ctx_scan     per chain: exclusive scan of the tile histograms along the chain (frame after frame, tile after
//                  tile) -> position of every tile's symbols inside the per-context lists; list starts, lengths and a
//                  longest-first processing order
//   k_dec_layout   per (frame, slice, plane context): where every line's decisions start in the decision region
//   k_ctx_scatter  stable scatter of (decision position, residual) into the per-context lists (coding order kept)
//   k_replay_ctx   one warp per (chain, context) list, longest lists first, lane = state slot: put_symbol_inline's
//                  binarisation (ffv1enc.c:185-231) with the context's 32-byte state in registers; writes p | bit<<8
//                  at the recorded positions of the decision stream that k_rangecode consumes
//   k_replay_grp   the same for 8-bit content, two lists per warp (16 lanes per symbol)
// Tile-sorted form of the lists (8-bit planar content, range coder and Golomb-Rice; Layout::tiled_lists): a context's
// list stays cut into one run per tile, so no histogram pass and no scan along the chains are needed:
//   k_tile_sort    counting sort of a 48-line tile by context in shared memory -> one contiguous block + run table
//   k_tile_layout  tiles' decision counts -> their places in the decision region
//   k_replay_grp<TILED> / k_gr_replay<TILED> (ffv1_enc_kernels.cu) walk a chain tile after tile
#include "ffv1_enc_kernels.cuh"
#include <cstdlib>
#include <cstring>

namespace ffv1 {

constexpr int kHistThreads = 256;
constexpr int kScanThreads = 1024;                       // one context per thread (<= kMaxListCtx)
constexpr int kScatterThreads = 128;                     // four tiles (one warp each) per CTA
constexpr int kCtxThreads = 1024;
constexpr int kMaxListCtx = 1024;                        // shared-memory sort / two-lists-per-warp replay / Golomb-Rice lists
constexpr int kMaxBigListCtx = 8192;                     // large context model (7563 contexts): direct scatter, model in global memory

__device__ __forceinline__ uint32_t cr_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t n = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

// lanes (among `active`) holding the same KEYBITS-bit key as the caller: one ballot per bit (match.any is far slower here)
template <int KEYBITS>
__device__ __forceinline__ uint32_t same_key_lanes(uint32_t key, uint32_t active)
{
    uint32_t grp = active;
#pragma unroll
    for (int b = 0; b < KEYBITS; b++) {
        const bool bit = (key >> b) & 1u;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, bit);
        grp &= bit ? m : ~m;
    }
    return grp;
}

__device__ __forceinline__ uint32_t decisions_of(int d)
{
    // put_symbol_inline: 1 decision for 0, else 2e+3 with e = floor(log2 |d|); branch-free: clz(0) = 32 gives 2*(-1)+3 = 1
    return (uint32_t)(65 - 2 * __clz((uint32_t)abs(d)));
}

// Golomb-Rice mode (ffv1enc.c:327-357): which of the 32 samples of a group are coded in run mode.  Run mode starts at a
// sample whose context is 0 and lasts up to and including the next non-zero residual:
//     M(i) = ctx(i)==0 || (M(i-1) && diff(i-1)==0)
// i.e. a carry chain with generate = zero & ctx0 and propagate = zero & ~ctx0, evaluated by one 64-bit addition.
// `carry` = the run continues into the next group of the line.  Samples with M && diff==0 are absorbed into the run
// (no VLC code, no state update); samples with M && diff!=0 end it and code diff-1 for positive residuals.
__device__ __forceinline__ uint32_t gr_run_members(uint32_t ctx0, uint32_t zero, uint32_t &carry)
{
    const uint32_t a = zero, b = zero & ctx0;
    const unsigned long long sum = (unsigned long long)a + b + carry;
    const uint32_t c = (uint32_t)sum ^ a ^ b;                 // carry INTO every bit = "a run reaches this sample"
    carry = (uint32_t)(sum >> 32);
    return ctx0 | c;
}

// ------------------------------------------------------------------------------------------------ k_ctx_hist
template <bool GOLOMB>
__global__ void __launch_bounds__(kHistThreads) k_ctx_hist(const EncDeviceTables T, const EncBatch B)
{
    extern __shared__ uint32_t s_hist[];                    // [ctx_count]
    const Layout &L = T.layout;
    const CtxTile ct = T.ctiles[blockIdx.x];
    const int f = blockIdx.y;
    const SliceGeom &g = T.slices[ct.slice];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < L.ctx_count; i += kHistThreads) s_hist[i] = 0;
    __syncthreads();
    const uint32_t *rec_slice = B.rec + (size_t)f * L.rec_per_frame + g.rec_first;
    const int32_t *my_lines = T.pc_lines + g.pc_line_first[ct.pc] + ct.first;
    for (int i = warp; i < ct.nlines; i += kHistThreads / 32) {
        const int line = my_lines[i];
        const LineDesc ld = T.lines[g.line_first + line];
        const uint32_t *recp = rec_slice + ld.rec_off;
        uint32_t nd = 0;
        if (GOLOMB) {
            // only the samples that get a VLC code count (run mode absorbs zero residuals); groups of a line in order
            uint32_t carry = 0u;
            for (int x0 = 0; x0 < ld.w; x0 += 32) {
                const bool act = x0 + lane < ld.w;
                const uint32_t r = act ? recp[x0 + lane] : 0xFFFF0001u;
                const uint32_t zero = __ballot_sync(0xFFFFFFFFu, act && (r & 0xFFFFu) == 0u);
                const uint32_t ctx0 = __ballot_sync(0xFFFFFFFFu, act && (r >> 16) == 0u);
                const uint32_t mem = gr_run_members(ctx0, zero, carry);
                if (act && !((mem & zero) >> lane & 1u)) atomicAdd(&s_hist[r >> 16], 1u);
            }
        } else
        for (int x = lane; x < ld.w; x += 32) {
            const uint32_t r = recp[x];
            nd += decisions_of((int)(int16_t)(r & 0xFFFFu));
            atomicAdd(&s_hist[r >> 16], 1u);
        }
        nd = __reduce_add_sync(0xFFFFFFFFu, nd);
        if (lane == 0) B.line_pos[(size_t)f * L.lines_per_frame + g.line_first + line] = nd;
    }
    __syncthreads();
    uint32_t *out = B.ctx_hist + ((size_t)f * L.ctiles_per_frame + blockIdx.x) * L.ctx_count;
    for (int i = tid; i < L.ctx_count; i += kHistThreads) out[i] = s_hist[i];
}

// ------------------------------------------------------------------------------------------------ k_ctx_scan
__global__ void __launch_bounds__(kScanThreads) k_ctx_scan(const EncDeviceTables T, const EncBatch B)
{
    extern __shared__ uint32_t s_scan[];                    // [ctx_count] list lengths, [ctx_count] list starts
    __shared__ uint32_t s_warp[kScanThreads / 32];
    const Layout &L = T.layout;
    const int chain = blockIdx.x;
    const int pc = chain % L.npc, s = (chain / L.npc) % L.nslices, seg = chain / (L.npc * L.nslices);
    const SliceGeom &g = T.slices[s];
    const int f0 = B.seg_first[seg], f1 = B.seg_first[seg + 1];
    const int nctx = L.ctx_count, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *s_total = s_scan, *s_start = s_scan + nctx;
    const int t0 = g.ct_first[pc], nt = g.ct_count[pc];
    for (int c = tid; c < nctx; c += kScanThreads) {
        uint32_t run = 0;
        for (int f = f0; f < f1; f++) {
            uint32_t *h = B.ctx_hist + ((size_t)f * L.ctiles_per_frame + t0) * nctx + c;
            // loads first, then the running sum and the stores: sixteen independent loads are in flight at a time
            for (int tb = 0; tb < nt; tb += 16) {
                uint32_t v[16];
#pragma unroll
                for (int k = 0; k < 16; k++) v[k] = tb + k < nt ? h[(size_t)(tb + k) * nctx] : 0u;
#pragma unroll
                for (int k = 0; k < 16; k++)
                    if (tb + k < nt) { h[(size_t)(tb + k) * nctx] = run; run += v[k]; }
            }
        }
        s_total[c] = run;
    }
    __syncthreads();
    // exclusive scan over the contexts: K consecutive contexts per thread (1 for the small model, 8 for 7563 contexts)
    {
        const int K = (nctx + kScanThreads - 1) / kScanThreads;
        uint32_t sum = 0;
        for (int k = 0; k < K; k++) { const int c = tid * K + k; if (c < nctx) sum += s_total[c]; }
        const uint32_t incl = cr_incl_scan(sum, lane);
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t base = incl - sum;
        for (int w = 0; w < warp; w++) base += s_warp[w];
        for (int k = 0; k < K; k++) { const int c = tid * K + k; if (c < nctx) { s_start[c] = base; base += s_total[c]; } }
    }
    __syncthreads();
    for (int c = tid; c < nctx; c += kScanThreads) {
        B.list_start[(size_t)chain * nctx + c] = s_start[c];
        B.list_count[(size_t)chain * nctx + c] = s_total[c];
        // longest list first (ties: lower context first)
        const uint32_t mine = s_total[c];
        int rank = 0;
        for (int o = 0; o < nctx; o++) {
            const uint32_t v = s_total[o];
            rank += (v > mine) || (v == mine && o < c);
        }
        B.list_order[(size_t)chain * nctx + rank] = (uint16_t)c;
    }
}

// ------------------------------------------------------------------------------------------------ k_dec_layout
__global__ void __launch_bounds__(128) k_dec_layout(const EncDeviceTables T, const EncBatch B)
{
    const Layout &L = T.layout;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long ndec = 0;
    if (idx < B.nframes * L.nslices * L.npc) {
        const int pc = idx % L.npc, s = (idx / L.npc) % L.nslices, f = idx / (L.npc * L.nslices);
        const SliceGeom &g = T.slices[s];
        const int nl = g.pc_nlines[pc];
        if (nl) {
            const int32_t *my_lines = T.pc_lines + g.pc_line_first[pc];
            uint32_t *line_pos = B.line_pos + (size_t)f * L.lines_per_frame + g.line_first;
            uint32_t *run_cnt = B.run_cnt + (size_t)f * L.runs_per_frame + g.run_first;
            uint32_t pos = 0, run_start = 0, cur_run = T.lines[g.line_first + my_lines[0]].run;
            for (int ib = 0; ib < nl; ib += 8) {
                // the loads of eight lines are issued together; only the running position is sequential
                int line[8]; uint32_t run[8], nd[8];
#pragma unroll
                for (int k = 0; k < 8; k++) line[k] = ib + k < nl ? my_lines[ib + k] : my_lines[nl - 1];
#pragma unroll
                for (int k = 0; k < 8; k++) { run[k] = T.lines[g.line_first + line[k]].run; nd[k] = line_pos[line[k]]; }
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (ib + k >= nl) break;
                    if (run[k] != cur_run) {                    // a run of this plane context ended: next one starts 16 B aligned
                        run_cnt[cur_run] = pos - run_start;
                        pos = (pos + 7u) & ~7u;
                        run_start = pos; cur_run = run[k];
                    }
                    line_pos[line[k]] = pos;
                    pos += nd[k];
                    ndec += nd[k];
                }
            }
            run_cnt[cur_run] = pos - run_start;
            if (pos + 8u > g.dec_cap[pc]) {
                const unsigned long long ns = g.pc_samples[pc];
                atomicMax(&B.status[0], ((unsigned long long)(pos + 8u) * 256ull + ns - 1) / ns + 1ull);
            }
        }
    }
    // one atomic per warp for the batch's decision count (tens of thousands of same-address atomics serialise)
#pragma unroll
    for (int d = 16; d; d >>= 1) ndec += __shfl_xor_sync(0xFFFFFFFFu, ndec, d);
    if ((threadIdx.x & 31) == 0 && ndec) atomicAdd(&B.status[3], ndec);
}

// List entry formats.  FMT 0: 8 bytes {position of the symbol's first decision, residual | frame << 16} (k_replay_ctx,
// Golomb-Rice).  FMT 1: 4 bytes, position | residual << 22 (k_replay_grp: 8-bit content, residuals fit 10 bits, a window
// lies inside one frame so the frame index is not needed; a (slice, plane context) decision region must stay below
// 2^22 entries).  Half the list traffic, and the staged scatter fits four CTAs per SM instead of three.
constexpr uint32_t kGrpPosBits = 22;

// ------------------------------------------------------------------------------------------------ k_ctx_scatter
// One warp per context tile; its lines are walked in coding order with one running list position per context in
// shared memory (initialised from k_ctx_scan's tile bases), so the scatter is stable by construction: inside a
// 32-sample group match.any ranks order the samples of a context, the leader advances the context's position.
template <int FMT, int KEYBITS>
__global__ void __launch_bounds__(kScatterThreads) k_ctx_scatter(const EncDeviceTables T, const EncBatch B)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Layout &L = T.layout;
    if (B.status[0]) return;
    const int nctx = L.ctx_count, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x * (kScatterThreads / 32) + warp;
    if (tile >= L.ctiles_per_frame) return;
    uint32_t *s_off = reinterpret_cast<uint32_t *>(smem_raw) + warp * nctx;       // [nctx] next list position per context
    const CtxTile ct = T.ctiles[tile];
    const int f = blockIdx.y;
    const SliceGeom &g = T.slices[ct.slice];
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int seg = B.frame_seg[f];
    const int f0 = B.seg_first[seg], seglen = B.seg_first[seg + 1] - f0;
    const int chain = (seg * L.nslices + ct.slice) * L.npc + ct.pc;
    {
        const uint32_t *tile_base = B.ctx_hist + ((size_t)f * L.ctiles_per_frame + tile) * nctx;
        const uint32_t *lstart = B.list_start + (size_t)chain * nctx;
        for (int c = lane; c < nctx; c += 32) s_off[c] = tile_base[c] + lstart[c];
    }
    __syncwarp();
    const uint32_t *rec_slice = B.rec + (size_t)f * L.rec_per_frame + g.rec_first;
    const int32_t *my_lines = T.pc_lines + g.pc_line_first[ct.pc] + ct.first;
    const size_t list_at = (size_t)f0 * L.samples_per_frame + (size_t)seglen * g.list_off[ct.pc];
    uint2 *list = B.lists + list_at;
    uint32_t *list32 = reinterpret_cast<uint32_t *>(B.lists) + list_at;
    const uint32_t *line_pos = B.line_pos + (size_t)f * L.lines_per_frame + g.line_first;
    for (int l = 0; l < ct.nlines; l++) {
        const int line = my_lines[l];
        const LineDesc ld = T.lines[g.line_first + line];
        const uint32_t *recp = rec_slice + ld.rec_off;
        uint32_t pos = line_pos[line];
        // 256 samples at a time: the eight loads of a chunk are in flight together (one load per 32-sample group
        // exposed the full memory latency every ~150 instructions)
        for (int xc = 0; xc < ld.w; xc += 256) {
            uint32_t rr[8];
#pragma unroll
            for (int k = 0; k < 8; k++) rr[k] = xc + k * 32 + lane < ld.w ? recp[xc + k * 32 + lane] : 0u;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int x0 = xc + k * 32;
                if (x0 >= ld.w) break;
                const bool act = x0 + lane < ld.w;
                const uint32_t r = rr[k];
                const uint32_t ctx = r >> 16;
                const uint32_t nd = act ? decisions_of((int)(int16_t)(r & 0xFFFFu)) : 0u;
                const uint32_t incl = cr_incl_scan(nd, lane);
                const uint32_t grp = same_key_lanes<KEYBITS>(ctx, __ballot_sync(0xFFFFFFFFu, act));   // contexts are < 2^KEYBITS
                const uint32_t rank = __popc(grp & lt_mask);
                if (act) {
                    if (FMT == 1) list32[s_off[ctx] + rank] = (pos + incl - nd) | (r << kGrpPosBits);
                    else list[s_off[ctx] + rank] = make_uint2(pos + incl - nd, (r & 0xFFFFu) | ((uint32_t)f << 16));
                }
                __syncwarp();
                if (act && rank == 0u) s_off[ctx] += __popc(grp);
                __syncwarp();
                pos += __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ k_ctx_scatter_sm
// Same result as k_ctx_scatter, but the tile is first sorted by context in SHARED memory (stable counting sort: one warp
// per line, per-line context histograms, ballot ranks inside a 32-sample group) and then written to the per-context
// lists as contiguous runs, one warp per run.  k_ctx_scatter's 8-byte stores land in ~20 different lists per
// 32-sample group; with thousands of tiles in flight the half-written sectors do not survive in L2 until their
// neighbours arrive, which costs 2x the DRAM traffic and stalls the warps on the store path.
constexpr int kScatterSmThreads = 512;
constexpr int kScatterSmMaxSamples = 5632;               // per tile (16 lines of <= 352 samples): 44 KB of entries

template <bool GOLOMB, int FMT>
__global__ void __launch_bounds__(kScatterSmThreads) k_ctx_scatter_sm(const EncDeviceTables T, const EncBatch B)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_wtot[kScatterSmThreads / 32];
    __shared__ int s_nne;                                               // contexts that occur in the tile
    const Layout &L = T.layout;
    if (B.status[0]) return;
    const int nctx = L.ctx_count, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *s_wh = reinterpret_cast<uint32_t *>(smem_raw);            // [8][nctx] per-line counts, two lines per word
    uint32_t *s_goff = s_wh + 8 * nctx;                                 // [nctx] where the tile's run of a context goes in its list
    uint16_t *s_start = reinterpret_cast<uint16_t *>(s_goff + nctx);    // [nctx] start of the run inside s_ent
    uint16_t *s_cnt = s_start + ((nctx + 7) & ~7);                      // [nctx]
    uint16_t *s_ne = s_cnt + ((nctx + 7) & ~7);                         // [nctx] the contexts that occur in the tile
    uint2 *s_ent = reinterpret_cast<uint2 *>(s_ne + ((nctx + 7) & ~7));
    uint32_t *s_ent32 = reinterpret_cast<uint32_t *>(s_ent);
    const int tile = blockIdx.x, f = blockIdx.y;
    const CtxTile ct = T.ctiles[tile];
    const SliceGeom &g = T.slices[ct.slice];
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int seg = B.frame_seg[f];
    const int f0 = B.seg_first[seg], seglen = B.seg_first[seg + 1] - f0;
    const int chain = (seg * L.nslices + ct.slice) * L.npc + ct.pc;
    for (int i = tid; i < 8 * nctx; i += kScatterSmThreads) s_wh[i] = 0u;
    if (tid == 0) s_nne = 0;
    {
        const uint32_t *tile_base = B.ctx_hist + ((size_t)f * L.ctiles_per_frame + tile) * nctx;
        const uint32_t *lstart = B.list_start + (size_t)chain * nctx;
        for (int c = tid; c < nctx; c += kScatterSmThreads) s_goff[c] = tile_base[c] + lstart[c];
    }
    __syncthreads();
    const uint32_t *rec_slice = B.rec + (size_t)f * L.rec_per_frame + g.rec_first;
    const int32_t *my_lines = T.pc_lines + g.pc_line_first[ct.pc] + ct.first;
    const uint32_t *recp = nullptr;
    int w = 0;
    uint32_t pos = 0;
    if (warp < ct.nlines) {
        const int line = my_lines[warp];
        const LineDesc ld = T.lines[g.line_first + line];
        recp = rec_slice + ld.rec_off;
        w = ld.w;
        pos = GOLOMB ? 0u : B.line_pos[(size_t)f * L.lines_per_frame + g.line_first + line];
        uint32_t *wh = s_wh + (warp >> 1) * nctx;
        const uint32_t one = (warp & 1) ? 0x10000u : 1u;
        if (GOLOMB) {
            uint32_t carry = 0u;
            for (int x0 = 0; x0 < w; x0 += 32) {
                const bool act = x0 + lane < w;
                const uint32_t r = act ? recp[x0 + lane] : 0xFFFF0001u;
                const uint32_t zero = __ballot_sync(0xFFFFFFFFu, act && (r & 0xFFFFu) == 0u);
                const uint32_t ctx0 = __ballot_sync(0xFFFFFFFFu, act && (r >> 16) == 0u);
                const uint32_t mem = gr_run_members(ctx0, zero, carry);
                if (act && !((mem & zero) >> lane & 1u)) atomicAdd(&wh[r >> 16], one);
            }
        } else
        for (int x = lane; x < w; x += 32) atomicAdd(&wh[recp[x] >> 16], one);
    }
    __syncthreads();
    // per context (two per thread): counts of the 16 lines -> exclusive offsets in place, total; block scan of the totals
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
    const uint32_t incl = cr_incl_scan(pair, lane);
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    {
        uint32_t base = incl - pair;
        for (int ww = 0; ww < warp; ww++) base += s_wtot[ww];
        if (2 * tid < nctx) s_start[2 * tid] = (uint16_t)base;
        if (2 * tid + 1 < nctx) s_start[2 * tid + 1] = (uint16_t)(base + cc[0]);
        if (cc[0]) s_ne[atomicAdd(&s_nne, 1)] = (uint16_t)(2 * tid);
        if (cc[1]) s_ne[atomicAdd(&s_nne, 1)] = (uint16_t)(2 * tid + 1);
    }
    __syncthreads();
    // stable placement: {position of the symbol's first decision, residual | frame << 16}
    if (warp < ct.nlines) {
        uint32_t *wh = s_wh + (warp >> 1) * nctx;
        const int sh = (warp & 1) * 16;
        uint32_t gcarry = 0u;
        const uint32_t rec_line = (uint32_t)(recp - rec_slice);         // Golomb-Rice mode: entries carry the record index
        for (int x0 = 0; x0 < w; x0 += 32) {
            bool act = x0 + lane < w;
            uint32_t r = act ? recp[x0 + lane] : (GOLOMB ? 0xFFFF0001u : 0u);
            if (GOLOMB) {
                const uint32_t zero = __ballot_sync(0xFFFFFFFFu, act && (r & 0xFFFFu) == 0u);
                const uint32_t ctx0 = __ballot_sync(0xFFFFFFFFu, act && (r >> 16) == 0u);
                const uint32_t mem = gr_run_members(ctx0, zero, gcarry);
                const bool inrun = (mem >> lane) & 1u;
                const int d = (int)(int16_t)(r & 0xFFFFu);
                if (inrun && d == 0) act = false;                         // absorbed into the run: no code, no list entry
                else if (inrun && d > 0) r = (r & 0xFFFF0000u) | (uint32_t)(d - 1);   // the residual that ends a run (ffv1enc.c:345-346)
            }
            const uint32_t ctx = r >> 16;
            const uint32_t nd = (act && !GOLOMB) ? decisions_of((int)(int16_t)(r & 0xFFFFu)) : 0u;
            const uint32_t in = GOLOMB ? 0u : cr_incl_scan(nd, lane);
            const uint32_t grp = __match_any_sync(0xFFFFFFFFu, act ? ctx : 0xFFFFFFFFu);
            const uint32_t rank = __popc(grp & lt_mask);
            uint32_t off = 0u;
            if (act && rank == 0u) off = (atomicAdd(&wh[ctx], (uint32_t)__popc(grp) << sh) >> sh) & 0xFFFFu;
            off = __shfl_sync(0xFFFFFFFFu, off, (__ffs(grp) - 1) & 31);
            if (act) {
                const uint32_t at = (uint32_t)s_start[ctx] + off + rank;
                if (FMT == 1) s_ent32[at] = (pos + in - nd) | (r << kGrpPosBits);
                else s_ent[at] = make_uint2(GOLOMB ? rec_line + (uint32_t)(x0 + lane) : pos + in - nd, (r & 0xFFFFu) | ((uint32_t)f << 16));
            }
            if (!GOLOMB) pos += __shfl_sync(0xFFFFFFFFu, in, 31);
        }
    }
    __syncthreads();
    // runs -> lists: one warp per context, 8 bytes per lane, contiguous
    const size_t list_at = (size_t)f0 * L.samples_per_frame + (size_t)seglen * g.list_off[ct.pc];
    const int nne = s_nne;
    for (int i = warp; i < nne; i += kScatterSmThreads / 32) {
        const int c = s_ne[i];
        const uint32_t n = s_cnt[c];
        if (FMT == 1) {
            uint32_t *dst = reinterpret_cast<uint32_t *>(B.lists) + list_at + s_goff[c];
            const uint32_t *src = s_ent32 + s_start[c];
            for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
        } else {
            uint2 *dst = B.lists + list_at + s_goff[c];
            const uint2 *src = s_ent + s_start[c];
            for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
        }
    }
}

// shared-memory accesses through 32-bit addresses kept in registers (the compiler otherwise rebuilds the shared-window
// address, cluster rank included, at every atomic)
__device__ __forceinline__ void sm_red_add(uint32_t addr, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t sm_atom_add(uint32_t addr, uint32_t v)
{
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void sm_st32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t sm_ld16(uint32_t addr)
{
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return v;
}

// ------------------------------------------------------------------------------------------------ k_tile_sort
// Tile-sorted lists (Layout::tiled_lists).  The chain-wide lists above need the position of every tile inside every
// list before a single entry can be placed: a histogram pass over all records and a scan along every chain.  Here a
// context's list stays cut into one run per tile: the tile is sorted by context in shared memory (as k_ctx_scatter_sm
// does) and leaves as ONE contiguous block at a place that only depends on the geometry, together with a table of where
// each context's run starts inside the block.  Decision positions are relative to the tile; k_tile_layout turns the
// tiles' decision counts into their places in the decision region.  k_replay_grp<TILED> walks a chain tile after tile.
// No histogram kernel, no scan kernel, no per-tile base tables to read back, and the block leaves with full-width
// coalesced stores (the per-context runs of the chain-wide lists average a dozen entries per store).
constexpr int kTileSortThreads = 32 * (kTiledLines / 2);  // two lines of the tile per warp
constexpr int kTileSortRows = kTileSortThreads / 64;      // per-warp context counts: two warps share a word

// GOLOMB: Golomb-Rice mode -- only the samples that get a VLC code are listed (run mode absorbs zero residuals, ffv1enc.c:
// 327-357; gr_run_members), an entry is the sample's record index inside the tile | its residual << 22 (the residual that
// ends a run is coded minus one), and no decisions are counted.
template <bool GOLOMB>
__global__ void __launch_bounds__(kTileSortThreads) k_tile_sort(const EncDeviceTables T, const EncBatch B)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_wtot[kTileSortThreads / 32];
    __shared__ uint32_t s_wnd[kTileSortThreads / 32];                    // decisions of every warp's lines, then their first position
    const Layout &L = T.layout;
    const int nctx = L.ctx_count, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *s_wh = reinterpret_cast<uint32_t *>(smem_raw);             // [rows][nctx] per-warp counts, two warps per word
    uint16_t *s_start = reinterpret_cast<uint16_t *>(s_wh + kTileSortRows * nctx);   // [pitch] start of every context's run; [nctx] = entries
    uint32_t *s_ent = reinterpret_cast<uint32_t *>(s_start + B.tile_tab_pitch);
    const int tile = blockIdx.x, f = blockIdx.y;
    const CtxTile ct = T.ctiles[tile];
    const SliceGeom &g = T.slices[ct.slice];
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int i = tid; i < kTileSortRows * nctx; i += kTileSortThreads) s_wh[i] = 0u;
    __syncthreads();
    const uint32_t *rec_slice = B.rec + (size_t)f * L.rec_per_frame + g.rec_first;
    const int32_t *my_lines = T.pc_lines + g.pc_line_first[ct.pc] + ct.first;
    // ---- pass 1: context counts and decisions of the warp's lines (2, or 4 in the tall tiles of narrow planes)
    const int lpw = (ct.nlines + kTileSortThreads / 32 - 1) / (kTileSortThreads / 32);
    uint32_t *wh = s_wh + (warp >> 1) * nctx;
    const int sh = (warp & 1) * 16;
    uint32_t wh_sa = (uint32_t)__cvta_generic_to_shared(wh), start_sa = (uint32_t)__cvta_generic_to_shared(s_start),
             ent_sa = (uint32_t)__cvta_generic_to_shared(s_ent);
    asm volatile("" : "+r"(wh_sa), "+r"(start_sa), "+r"(ent_sa));
    const uint32_t tile_rec0 = T.lines[g.line_first + my_lines[0]].rec_off;       // GOLOMB: record indices count from here
    {
        uint32_t nd = 0;
        for (int k = 0; k < lpw; k++) {
            const int li = lpw * warp + k;
            if (li < ct.nlines) {
                const LineDesc ld = T.lines[g.line_first + my_lines[li]];
                const uint32_t *recp = rec_slice + ld.rec_off;
                if (GOLOMB) {
                    uint32_t carry = 0u;
                    for (int x0 = 0; x0 < ld.w; x0 += 32) {
                        const bool act = x0 + lane < ld.w;
                        const uint32_t r = act ? recp[x0 + lane] : 0xFFFF0001u;
                        const uint32_t zero = __ballot_sync(0xFFFFFFFFu, act && (r & 0xFFFFu) == 0u);
                        const uint32_t ctx0 = __ballot_sync(0xFFFFFFFFu, act && (r >> 16) == 0u);
                        const uint32_t mem = gr_run_members(ctx0, zero, carry);
                        if (act && !((mem & zero) >> lane & 1u)) sm_red_add(wh_sa + 4u * (r >> 16), 1u << sh);
                    }
                } else
                for (int x = lane; x < ld.w; x += 32) {
                    const uint32_t r = recp[x];
                    nd += decisions_of((int)(int16_t)(r & 0xFFFFu));
                    sm_red_add(wh_sa + 4u * (r >> 16), 1u << sh);
                }
            }
        }
        nd = __reduce_add_sync(0xFFFFFFFFu, nd);
        if (lane == 0) s_wnd[warp] = nd;
    }
    __syncthreads();
    // ---- per context (two per thread): counts of the warps -> exclusive offsets in place; block scan of the totals
    uint32_t cc[2] = {0u, 0u};
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int c = 2 * tid + j;
        if (c < nctx) {
            uint32_t run = 0;
#pragma unroll
            for (int h = 0; h < kTileSortRows; h++) {
                const uint32_t v = s_wh[h * nctx + c];
                const uint32_t lo = v & 0xFFFFu, hi = v >> 16;
                s_wh[h * nctx + c] = run | ((run + lo) << 16);
                run += lo + hi;
            }
            cc[j] = run;
        }
    }
    const uint32_t pair = cc[0] + cc[1];
    const uint32_t incl = cr_incl_scan(pair, lane);
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    {
        uint32_t base = incl - pair;
        for (int ww = 0; ww < warp; ww++) base += s_wtot[ww];
        if (2 * tid < nctx) s_start[2 * tid] = (uint16_t)base;
        if (2 * tid + 1 < nctx) s_start[2 * tid + 1] = (uint16_t)(base + cc[0]);
        if (2 * tid == nctx || 2 * tid + 1 == nctx) s_start[nctx] = (uint16_t)(base + (2 * tid == nctx ? 0u : cc[0]));
        if (tid == 0) {                                                   // first decision of every warp's lines; the tile's total
            uint32_t run = 0;
            for (int ww = 0; ww < kTileSortThreads / 32; ww++) { const uint32_t v = s_wnd[ww]; s_wnd[ww] = run; run += v; }
            B.tile_nd[(size_t)f * L.ctiles_per_frame + tile] = run;
        }
    }
    __syncthreads();
    // ---- pass 2: stable placement, entry = position of the symbol's first decision inside the tile | residual << 22
    {
        uint32_t pos = s_wnd[warp];
        for (int k = 0; k < lpw; k++) {
            const int li = lpw * warp + k;
            if (li >= ct.nlines) break;
            const LineDesc ld = T.lines[g.line_first + my_lines[li]];
            const uint32_t *recp = rec_slice + ld.rec_off;
            const int w = ld.w;
            uint32_t gcarry = 0u;
            for (int x0 = 0; x0 < w; x0 += 32) {
                bool act = x0 + lane < w;
                uint32_t r = act ? recp[x0 + lane] : (GOLOMB ? 0xFFFF0001u : 0u);
                if (GOLOMB) {
                    const uint32_t zero = __ballot_sync(0xFFFFFFFFu, act && (r & 0xFFFFu) == 0u);
                    const uint32_t ctx0 = __ballot_sync(0xFFFFFFFFu, act && (r >> 16) == 0u);
                    const uint32_t mem = gr_run_members(ctx0, zero, gcarry);
                    const bool inrun = (mem >> lane) & 1u;
                    const int d = (int)(int16_t)(r & 0xFFFFu);
                    if (inrun && d == 0) act = false;                         // absorbed into the run: no code, no list entry
                    else if (inrun && d > 0) r = (r & 0xFFFF0000u) | (uint32_t)(d - 1);   // the residual that ends a run (ffv1enc.c:345-346)
                }
                const uint32_t ctx = r >> 16;
                const uint32_t nd = (act && !GOLOMB) ? decisions_of((int)(int16_t)(r & 0xFFFFu)) : 0u;
                const uint32_t in = GOLOMB ? 0u : cr_incl_scan(nd, lane);
                const uint32_t grp = __match_any_sync(0xFFFFFFFFu, act ? ctx : 0xFFFFFFFFu);
                const uint32_t rank = __popc(grp & lt_mask);
                uint32_t off = 0u;
                if (act && rank == 0u) off = (sm_atom_add(wh_sa + 4u * ctx, (uint32_t)__popc(grp) << sh) >> sh) & 0xFFFFu;
                off = __shfl_sync(0xFFFFFFFFu, off, (__ffs(grp) - 1) & 31);
                if (act) {
                    const uint32_t where = GOLOMB ? ld.rec_off - tile_rec0 + (uint32_t)(x0 + lane) : pos + in - nd;
                    sm_st32(ent_sa + 4u * (sm_ld16(start_sa + 2u * ctx) + off + rank), where | (r << kGrpPosBits));
                }
                if (!GOLOMB) pos += __shfl_sync(0xFFFFFFFFu, in, 31);
            }
        }
    }
    __syncthreads();
    // ---- the block and its table leave with coalesced stores
    const int seg = B.frame_seg[f];
    const int f0 = B.seg_first[seg], seglen = B.seg_first[seg + 1] - f0;
    uint32_t *dst = reinterpret_cast<uint32_t *>(B.lists) + (size_t)f0 * L.samples_per_frame + (size_t)seglen * g.list_off[ct.pc] +
                    (size_t)(f - f0) * g.pc_samples[ct.pc] + ct.sample_first;
    const uint32_t nent = GOLOMB ? (uint32_t)s_start[nctx] : ct.nsamples;
    for (uint32_t i = tid; i < nent; i += kTileSortThreads) dst[i] = s_ent[i];
    uint16_t *tab = B.tile_tab + ((size_t)f * L.ctiles_per_frame + tile) * B.tile_tab_pitch;
    for (int i = tid; i <= nctx; i += kTileSortThreads) tab[i] = s_start[i];
}

// ------------------------------------------------------------------------------------------------ k_tile_layout
// per (frame, slice, plane context): the tiles' decision counts -> their first decision inside the region (one run per
// plane context on this path), the run's length for k_rangecode, overflow detection, the batch's decision count
__global__ void __launch_bounds__(128) k_tile_layout(const EncDeviceTables T, const EncBatch B)
{
    const Layout &L = T.layout;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long ndec = 0;
    if (idx < B.nframes * L.nslices * L.npc) {
        const int pc = idx % L.npc, s = (idx / L.npc) % L.nslices, f = idx / (L.npc * L.nslices);
        const SliceGeom &g = T.slices[s];
        const int nt = g.ct_count[pc];
        if (nt) {
            uint32_t *nd = B.tile_nd + (size_t)f * L.ctiles_per_frame + g.ct_first[pc];
            uint32_t pos = 0;
            for (int tb = 0; tb < nt; tb += 8) {
                uint32_t v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = tb + k < nt ? nd[tb + k] : 0u;
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (tb + k < nt) { nd[tb + k] = pos; pos += v[k]; }
            }
            const uint32_t run = T.lines[g.line_first + T.pc_lines[g.pc_line_first[pc]]].run;
            B.run_cnt[(size_t)f * L.runs_per_frame + g.run_first + run] = pos;
            ndec = pos;
            if (pos + 8u > g.dec_cap[pc]) {
                const unsigned long long ns = g.pc_samples[pc];
                atomicMax(&B.status[0], ((unsigned long long)(pos + 8u) * 256ull + ns - 1) / ns + 1ull);
            }
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) ndec += __shfl_xor_sync(0xFFFFFFFFu, ndec, d);
    if ((threadIdx.x & 31) == 0 && ndec) atomicAdd(&B.status[3], ndec);
}

// ------------------------------------------------------------------------------------------------ k_replay_ctx
// Lane <-> state slot permutation.  put_symbol_inline visits the slots of a symbol in the order
//   0 | 1..e+1 | 22+e-1 .. 22 | 11+e                                  (ffv1enc.c:202-229, e <= 9)
// With slots 22..31 held in REVERSE order by lanes 11..20 and slots 11..21 by lanes 21..31, that order is increasing
// in lane index for every e, so a lane's decision sits at popc(visit & lanes_below) inside the symbol.
__device__ __forceinline__ int slot_of_lane(int lane) { return lane <= 10 ? lane : (lane <= 20 ? 42 - lane : lane - 10); }
__device__ __forceinline__ int lane_of_slot(int slot) { return slot <= 10 ? slot : (slot >= 22 ? 42 - slot : slot + 10); }

__device__ __forceinline__ void symbol_masks(int d, uint32_t &visit, uint32_t &bits)
{
    if (d == 0) { visit = 1u; bits = 1u; return; }                         // "is zero" flag = 1
    const uint32_t a = (uint32_t)abs(d);
    const int e = min(31 - __clz(a), 9);                                   // e > 9 symbols take the sequential path
    const uint32_t ones_e = (1u << e) - 1u;
    const uint32_t mant = e ? (__brev(a & ones_e) >> (32 - e)) : 0u;       // lane 21-e+t <- bit e-1-t of |d|
    visit = 1u | (((2u << e) - 1u) << 1) | (ones_e << (21 - e)) | (1u << (21 + e));
    bits = (ones_e << 1) | (mant << (21 - e)) | ((d < 0 ? 1u : 0u) << (21 + e));
}

// One warp per context list, lane = state slot: the 32-byte state of the context lives in one register per lane, a
// symbol costs a handful of instructions (its visit / bit masks are computed 32 symbols at a time and broadcast), and
// the only serial dependency is state -> table lookup -> state of the slots a symbol touches.  The warps of a CTA
// (one CTA per chain) take the chain's lists from a shared counter, longest first.
__device__ __forceinline__ uint32_t lut_ld(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

__device__ __forceinline__ void visit_step(uint16_t *addr, uint32_t val, uint32_t &st, uint32_t saddr, uint32_t pred)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t@p st.global.u16 [%1], %2;\n\t@p ld.shared.u8 %0, [%3];\n\t}"
                 : "+r"(st) : "l"(addr), "h"((unsigned short)val), "r"(saddr), "r"(pred) : "memory");
}

// GSTATE: the chain's model stays in global memory (B.state_seg; the large context model's 7563 x 32 bytes do not fit an
// SM's shared memory).  A list only reads its row when it is taken up and writes it back when the frame's part is done.
template <bool HIGH_E, bool GSTATE>
__global__ void __launch_bounds__(kCtxThreads, 1) k_replay_ctx(const EncDeviceTables T, const EncBatch B)
{
    extern __shared__ __align__(16) unsigned char s_state_raw[];     // [ctx_count][32] the chain's model
    __shared__ uint8_t s_lut[512];
    __shared__ uint4 s_blk_all[kCtxThreads];
    __shared__ int s_next;
    const Layout &L = T.layout;
    uint8_t *s_state = GSTATE ? B.state_seg + (size_t)blockIdx.x * ((size_t)L.ctx_count * 32) : s_state_raw;
    const int tid = threadIdx.x, lane = tid & 31;
    uint4 *s_blk = s_blk_all + (tid & ~31);
    const uint32_t lut_base = (uint32_t)__cvta_generic_to_shared(s_lut);
    for (int i = tid; i < 512; i += kCtxThreads) s_lut[i] = T.trans_lut[i];
    if (B.status[0]) return;
    const int chain = chain_of_block(L, T.slices);          // longest chains first
    const int pc = chain % L.npc, s = (chain / L.npc) % L.nslices, seg = chain / (L.npc * L.nslices);
    const SliceGeom &g = T.slices[s];
    const int f0 = B.seg_first[seg], f1 = B.seg_first[seg + 1];
    const int nctx = L.ctx_count;
    const bool key = B.frame_key[f0] != 0;
    const bool hand_over = f1 == B.nframes;
    const size_t coff = ((size_t)s * L.npc + pc) * ((size_t)nctx * 32);
    // ---- the chain's model: the initial states on keyframes (128, or a two-pass encode's table; ffv1.c:177-202), else what the previous batch left
    {
        const uint4 *cin4 = reinterpret_cast<const uint4 *>(B.carry_in + coff);
        uint4 *st4 = reinterpret_cast<uint4 *>(s_state);
        for (int i = tid; i < nctx * 2; i += kCtxThreads)
            st4[i] = key ? (T.init_state ? reinterpret_cast<const uint4 *>(T.init_state)[i] : make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u)) : cin4[i];
    }

    const uint2 *chain_list = B.lists + (size_t)f0 * L.samples_per_frame + (size_t)(f1 - f0) * g.list_off[pc];
    const uint32_t *lstart = B.list_start + (size_t)chain * nctx;
    const uint32_t *lcount = B.list_count + (size_t)chain * nctx;
    const uint16_t *order = B.list_order + (size_t)chain * nctx;
    uint16_t *dec_pc = B.dec + g.dec_off[pc];
    const int slot = slot_of_lane(lane);
    const uint32_t lanebit = 1u << lane, lt_mask = lanebit - 1u;
    const int t0 = g.ct_first[pc];
    // list bookkeeping of the current frame in shared memory (behind the model, when that is there too): where the
    // frame's part of every list starts, how long it is, and the processing order -- taking the next list costs no
    // global load (with 7563 mostly short lists the dependent loads of a take were most of the kernel's time)
    uint32_t *s_at = reinterpret_cast<uint32_t *>(s_state_raw + (GSTATE ? 0 : (size_t)nctx * 32));
    uint32_t *s_n = s_at + nctx;
    uint16_t *s_order = reinterpret_cast<uint16_t *>(s_n + nctx);
    __shared__ int s_nlists;
    if (tid == 0) s_nlists = 0;
    __syncthreads();
    for (int i = tid; i < nctx; i += kCtxThreads) {
        const uint16_t c = order[i];
        s_order[i] = c;
        if (lcount[c]) atomicMax(&s_nlists, i + 1);             // the order is by list length: the non-empty lists come first
    }

    // Frame after frame: all warps of the CTA work on the same (chain, frame), so the decision region they scatter
    // into (one frame of one slice-plane-context, < 1 MB) is completed while it is still in L2 -- with every list
    // walked through all frames of the GOP at once, each 2..20-byte write was a read-modify-write in DRAM.
    for (int f = f0; f < f1; f++) {
      __syncthreads();                                              // model loaded / previous frame finished
      if (tid == 0) s_next = 0;
      {
        // symbols of context c that precede frame f (f+1) inside the chain = scanned histogram of the chain's first tile
        const uint32_t *before = B.ctx_hist + ((size_t)f * L.ctiles_per_frame + t0) * nctx;
        const uint32_t *before_next = B.ctx_hist + ((size_t)(f + 1) * L.ctiles_per_frame + t0) * nctx;
        const bool more = f + 1 < f1;
        for (int i = tid; i < nctx; i += kCtxThreads) {
            const uint32_t b0 = before[i], b1 = more ? before_next[i] : lcount[i];
            s_at[i] = lstart[i] + b0; s_n[i] = b1 - b0;
        }
      }
      __syncthreads();
      const int nlists = s_nlists;
      for (;;) {
        int oi = 0;
        if (lane == 0) oi = atomicAdd(&s_next, 1);
        oi = __shfl_sync(0xFFFFFFFFu, oi, 0);
        if (oi >= nlists) break;
        const int c = s_order[oi];
        const uint32_t n = s_n[c];
        if (n == 0u) continue;
        uint32_t st = s_state[c * 32 + slot];
        const uint2 *lp = chain_list + s_at[c];
        uint2 nx = lane < n ? lp[lane] : make_uint2(0u, 0u);
        for (uint32_t j0 = 0; j0 < n; j0 += 32u) {
            const uint32_t m = min(32u, n - j0);
            const uint2 en = nx;
            if (j0 + 32u + lane < n) nx = lp[j0 + 32u + lane];
            const int d = (int)(int16_t)(en.y & 0xFFFFu);
            uint32_t vis, bts;
            symbol_masks(d, vis, bts);
            if (HIGH_E && (uint32_t)abs(d) >= 1024u) vis = 0u;  // marks a symbol for the sequential path
            // decision positions relative to the block's first symbol (a list never jumps more than one frame inside a block)
            const unsigned long long off = (unsigned long long)(en.y >> 16) * L.dec_per_frame + en.x;
            const unsigned long long off0 = __shfl_sync(0xFFFFFFFFu, off, 0);
            const uint32_t rel = (uint32_t)(off - off0);
            const bool wide = __any_sync(0xFFFFFFFFu, lane < m && ((off - off0) >> 32) != 0ull);   // rare context, huge GOP
            uint16_t *o0 = dec_pc + off0;
            asm volatile("" : "+l"(o0));                                    // keep the block base in one register pair
            // the block's (visit, bits, position) triples go through shared memory: one broadcast 16-byte load per symbol
            __syncwarp();
            s_blk[lane] = make_uint4(vis, bts, rel, 0u);
            __syncwarp();
            if (!wide && !(HIGH_E && __any_sync(0xFFFFFFFFu, lane < m && vis == 0u))) {
#pragma unroll 4
                for (uint32_t k = 0; k < m; k++) {
                    const uint4 q = s_blk[k];
                    const uint32_t idx = q.z + __popc(q.x & lt_mask);
                    const uint32_t bit = (q.y & lanebit) ? 0x100u : 0u;
                    // predicated (not branched): emit the decision and step the state only in the lanes the symbol visits
                    visit_step(reinterpret_cast<uint16_t *>(reinterpret_cast<char *>(o0) + (size_t)idx * 2u), st | bit, st,
                               lut_base + bit + st, q.x & lanebit);
                }
            } else {
                for (uint32_t k = 0; k < m; k++) {
                    const uint4 q = s_blk[k];
                    uint16_t *ok = wide ? dec_pc + __shfl_sync(0xFFFFFFFFu, off, k) : o0 + q.z;
                    if (HIGH_E && q.x == 0u) {
                        // large magnitudes (ffv1enc.c:217-228): slots 1+9 and 22+9 repeat, walk the decisions one by one
                        const int dk = __shfl_sync(0xFFFFFFFFu, d, k);
                        const uint32_t a = (uint32_t)abs(dk);
                        const int ex = 31 - __clz(a);
                        const int nd = 2 * ex + 3;
                        for (int qi = 0; qi < nd; qi++) {
                            int sl; uint32_t bit;
                            if (qi == 0) { sl = 0; bit = 0u; }
                            else if (qi <= ex) { sl = 1 + min(qi - 1, 9); bit = 0x100u; }
                            else if (qi == ex + 1) { sl = 1 + 9; bit = 0u; }
                            else if (qi <= 2 * ex + 1) { const int i = ex - 1 - (qi - ex - 2); sl = 22 + min(i, 9); bit = ((a >> i) & 1u) << 8; }
                            else { sl = 11 + 10; bit = dk < 0 ? 0x100u : 0u; }
                            if (slot == sl) {
                                ok[qi] = (uint16_t)(st | bit);
                                st = lut_ld(lut_base + bit + st);
                            }
                        }
                        continue;
                    }
                    if (q.x & lanebit) {
                        const uint32_t bit = (q.y & lanebit) ? 0x100u : 0u;
                        ok[__popc(q.x & lt_mask)] = (uint16_t)(st | bit);
                        st = lut_ld(lut_base + bit + st);
                    }
                }
            }
        }
        s_state[c * 32 + slot] = (uint8_t)st;
      }
    }
    // ---- hand the model to the next batch when this segment runs to the end of the batch
    __syncthreads();
    if (hand_over) {
        const uint4 *st4 = reinterpret_cast<const uint4 *>(s_state);
        uint4 *cout4 = reinterpret_cast<uint4 *>(B.carry_out + coff);
        for (int i = tid; i < nctx * 2; i += kCtxThreads) cout4[i] = st4[i];
    }
}

// ------------------------------------------------------------------------------------------------ k_replay_grp
// Same replay, but a warp works on TWO context lists at a time, 16 lanes each.  A symbol with exponent e <= EMAX only
// touches the zero flag, e+1 exponent slots, e mantissa slots and one sign slot (ffv1enc.c:202-229), i.e. at most
// 3*EMAX+3 of the 32 state slots; with EMAX = 4 (|residual| < 32) they fit 15 lanes:
//     lane 0            slot 0                     "is zero"
//     lanes 1..5        slots 1..5                 exponent, unary
//     lanes 6..9        slots 25,24,23,22          mantissa bits 3..0 (the coding order walks them downwards)
//     lanes 10..14      slots 11..15               sign, one slot per exponent
// so the visiting order of a symbol is again increasing in lane index.  Larger residuals (rare on 8-bit content) are
// replayed by one lane of the group against the shared-memory copy of the state row.  Halving the lanes per symbol
// halves the instructions issued per symbol, which is what bounds this kernel; two CTAs of 512 threads share an SM so
// that one chain's end-of-frame tail (its longest list) overlaps the other chain's work.
template <int EMAX>
__device__ __forceinline__ void symbol_masks_grp(int d, uint32_t &visit, uint32_t &bits, bool &slow)
{
    slow = false;
    if (d == 0) { visit = 1u; bits = 1u; return; }
    const uint32_t a = (uint32_t)abs(d);
    const int e = 31 - __clz(a);
    if (e > EMAX) { visit = 0u; bits = 0u; slow = true; return; }
    const uint32_t ones_e = (1u << e) - 1u;
    const uint32_t mant = e ? (__brev(a & ones_e) >> (32 - e)) : 0u;       // bit t <- bit e-1-t of |d|
    visit = 1u | (((2u << e) - 1u) << 1) | (ones_e << (2 * EMAX + 2 - e)) | (1u << (2 * EMAX + 2 + e));
    bits = (ones_e << 1) | (mant << (2 * EMAX + 2 - e)) | ((d < 0 ? 1u : 0u) << (2 * EMAX + 2 + e));
}

// put_symbol_inline (ffv1enc.c:185-231) decision by decision against a state row in shared memory (any magnitude)
__device__ __noinline__ void replay_symbol_serial(uint8_t *row, const uint8_t *lut, uint16_t *o, int d)
{
    if (d == 0) { const uint32_t s = row[0]; o[0] = (uint16_t)(s | 0x100u); row[0] = lut[0x100u + s]; return; }
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

// TILED: the lists are kept tile by tile (k_tile_sort): a window is one tile, the window's part of a list is the
// context's run inside the tile's block, and the entries' positions count from the tile's first decision.  Three
// 256-thread CTAs per SM (80 registers, no spills) beat four (64 registers): 141 vs 146 ms per 2048 frames; five (48
// registers) 176 ms, six 128-thread CTAs 184 ms, two of 384 threads 150 ms -- the time follows the windows in flight.
template <int EMAX, int THREADS, bool TILED>
__global__ void __launch_bounds__(THREADS, (TILED ? 768 : 1024) / THREADS) k_replay_grp(const EncDeviceTables T, const EncBatch B, const int window_rt)
{
    const int window = TILED ? 1 : window_rt;
    constexpr int G = 16;
    static_assert(3 * EMAX + 3 <= G, "roles must fit a group");
    extern __shared__ __align__(16) unsigned char s_state_raw[];     // [ctx_count][32] the chain's model
    __shared__ uint8_t s_lut[512];
    __shared__ uint4 s_blk_all[THREADS];
    __shared__ int s_next;
    __shared__ uint2 s_mask[64];                                     // (visit, bits) of the residuals -32..31 (visit 0: serial path)
    uint8_t *s_state = s_state_raw;
    const Layout &L = T.layout;
    const int tid = threadIdx.x, lane = tid & 31, g = lane & (G - 1);
    const uint32_t gmask = 0xFFFFu << (lane & 16);                   // the lanes of my group
    uint4 *s_blk = s_blk_all + (tid & ~(G - 1));                     // my group's block of G entries
    const uint32_t lut_base = (uint32_t)__cvta_generic_to_shared(s_lut);
    for (int i = tid; i < 512; i += THREADS) s_lut[i] = T.trans_lut[i];
    if (tid < 64) {
        uint32_t v = 0u, bt = 0u; bool sl = false;
        symbol_masks_grp<EMAX>(tid - 32, v, bt, sl);
        s_mask[tid] = make_uint2(sl ? 0u : v, bt);
    }
    if (B.status[0]) return;
    const int chain = chain_of_block(L, T.slices);          // longest chains first
    const int pc = chain % L.npc, s = (chain / L.npc) % L.nslices, seg = chain / (L.npc * L.nslices);
    const SliceGeom &sg = T.slices[s];
    const int f0 = B.seg_first[seg], f1 = B.seg_first[seg + 1];
    const int nctx = L.ctx_count;
    // list bookkeeping of the chain / of the current window in shared memory: taking the next list costs no global load
    uint32_t *s_lstart = reinterpret_cast<uint32_t *>(s_state_raw + (size_t)nctx * 32);     // [nctx] start of every list
    uint32_t *s_b0 = s_lstart + nctx, *s_b1 = s_b0 + nctx;                                // [nctx] the window's part of every list
    uint16_t *s_order = reinterpret_cast<uint16_t *>(s_b1 + nctx);                          // [nctx] contexts, longest list first
    __shared__ int s_nlists;
    __shared__ int s_nl[2];                                          // TILED: lists queued in the current / the next window
    const bool key = B.frame_key[f0] != 0;
    const bool hand_over = f1 == B.nframes;
    const size_t coff = ((size_t)s * L.npc + pc) * ((size_t)nctx * 32);
    {
        const uint4 *cin4 = reinterpret_cast<const uint4 *>(B.carry_in + coff);
        uint4 *st4 = reinterpret_cast<uint4 *>(s_state);
        for (int i = tid; i < nctx * 2; i += THREADS)
            st4[i] = key ? (T.init_state ? reinterpret_cast<const uint4 *>(T.init_state)[i] : make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u)) : cin4[i];
    }
    const uint32_t *chain_list = reinterpret_cast<const uint32_t *>(B.lists) + (size_t)f0 * L.samples_per_frame + (size_t)(f1 - f0) * sg.list_off[pc];
    const uint32_t *lstart = B.list_start + (size_t)chain * nctx;
    const uint32_t *lcount = B.list_count + (size_t)chain * nctx;
    const uint16_t *order = B.list_order + (size_t)chain * nctx;
    uint16_t *dec_pc = B.dec + sg.dec_off[pc];
    const int t0 = sg.ct_first[pc];
    const int nt = sg.ct_count[pc];
    if (tid == 0) { s_nlists = 0; s_nl[0] = 0; s_nl[1] = 0; }
    __syncthreads();
    if (TILED) {
        // processing order: by how often the contexts occur in the chain's first frame (what matters is that the long
        // lists of a window start first; contexts that frame does not use come last)
        for (int i = tid; i < nctx; i += THREADS) {
            uint32_t n = 0;
            for (int tt = 0; tt < nt; tt++) {                        // the tiles of the chain's first frame
                const uint16_t *tab = B.tile_tab + ((size_t)f0 * L.ctiles_per_frame + t0 + tt) * B.tile_tab_pitch;
                n += (uint32_t)tab[i + 1] - (uint32_t)tab[i];
            }
            s_b0[i] = n;
        }
        __syncthreads();
        for (int c = tid; c < nctx; c += THREADS) {
            const uint32_t mine = s_b0[c];
            int rank = 0;
            for (int o = 0; o < nctx; o++) {
                const uint32_t v = s_b0[o];
                rank += (v > mine) || (v == mine && o < c);
            }
            s_order[rank] = (uint16_t)c;
        }
    } else
    for (int i = tid; i < nctx; i += THREADS) {
        s_lstart[i] = lstart[i];
        const uint16_t c = order[i];
        s_order[i] = c;
        if (lcount[c]) atomicMax(&s_nlists, i + 1);           // the order is by list length: the non-empty lists come first
    }
    // role of the lane inside its group
    const int slot = g == 0 ? 0 : (g <= EMAX + 1 ? g : (g <= 2 * EMAX + 1 ? 22 + (2 * EMAX + 1 - g) : (g <= 3 * EMAX + 2 ? 11 + g - (2 * EMAX + 2) : -1)));
    const bool has_slot = slot >= 0;
    const uint32_t lanebit = 1u << g, lt_mask = lanebit - 1u;
    const int rot = (g - 8) & 31;                                    // rotr(bits, rot) puts my bit at bit 8

    // Window after window (`window` context tiles of a frame): all warps of the CTA work on the same stretch of the
    // (frame, slice, plane context) decision region, which is then completed while it is still in L2 (with whole frames
    // per round, 296 resident CTAs x 0.8 MB of half-written sectors thrashed the 126 MB L2: 3x DRAM traffic).  Measured
    // optimum on B200: 256-thread CTAs (4 per SM) x 3 tiles; the time follows (resident CTAs x window), i.e. the L2 footprint.
    int wi = 0;                                                      // windows done
    for (int f = f0; f < f1; f++)
    for (int tw = 0; tw < nt; tw += window, wi++) {
        uint16_t *dec_f = dec_pc + (size_t)f * L.dec_per_frame;
        const uint32_t *blk = chain_list;                            // TILED: the tile's block of entries
        __syncthreads();                                             // model loaded / previous window finished
        if (tid == 0) s_next = 0;
        const bool last_win = tw + window >= nt;
        if (TILED) {
            // the window's part of every list, indexed by the list's place in the processing order; the queue ends behind
            // the last context that occurs in the tile (the tail of the order is contexts the content hardly ever uses)
            const size_t ti = (size_t)f * L.ctiles_per_frame + t0 + tw;
            const uint16_t *tab = B.tile_tab + ti * B.tile_tab_pitch;
            if (tid == 0) s_nl[(wi + 1) & 1] = 0;
            int last = 0;
            for (int i = tid; i < nctx; i += THREADS) {
                const int c = s_order[i];
                const uint32_t a = tab[c], b = tab[c + 1];
                s_b0[i] = a; s_b1[i] = b;
                if (b > a) last = i + 1;
            }
            last = __reduce_max_sync(0xFFFFFFFFu, last);
            if (lane == 0 && last) atomicMax(&s_nl[wi & 1], last);
            // the next window's table and entries are on their way into L2 while this one is replayed
            {
                int nf = f, ntw = tw + 1;
                if (ntw >= nt) { ntw = 0; nf = f + 1; }
                if (nf < f1) {
                    const size_t nti = (size_t)nf * L.ctiles_per_frame + t0 + ntw;
                    const char *ntab = reinterpret_cast<const char *>(B.tile_tab + nti * B.tile_tab_pitch);
                    const CtxTile nct = T.ctiles[t0 + ntw];
                    const char *nblk = reinterpret_cast<const char *>(chain_list + (size_t)(nf - f0) * sg.pc_samples[pc] + nct.sample_first);
                    if (tid * 128 < B.tile_tab_pitch * 2) asm volatile("prefetch.global.L2 [%0];" :: "l"(ntab + tid * 128));
                    for (uint32_t o = tid * 128u; o < nct.nsamples * 4u; o += THREADS * 128u) asm volatile("prefetch.global.L2 [%0];" :: "l"(nblk + o));
                    if (tid == THREADS - 1) asm volatile("prefetch.global.L2 [%0];" :: "l"(B.tile_nd + nti));
                }
            }
            blk = chain_list + (size_t)(f - f0) * sg.pc_samples[pc] + T.ctiles[t0 + tw].sample_first;
            dec_f += B.tile_nd[ti];
        } else {
            const uint32_t *bf = B.ctx_hist + ((size_t)f * L.ctiles_per_frame + t0 + tw) * nctx;
            const uint32_t *bn = last_win ? B.ctx_hist + ((size_t)(f + 1) * L.ctiles_per_frame + t0) * nctx
                                          : B.ctx_hist + ((size_t)f * L.ctiles_per_frame + t0 + tw + window) * nctx;
            const bool use_bn = !last_win || f + 1 < f1;
            for (int i = tid; i < nctx; i += THREADS) { s_b0[i] = bf[i]; s_b1[i] = use_bn ? bn[i] : lcount[i]; }
        }
        __syncthreads();
        const int nlists = TILED ? s_nl[wi & 1] : s_nlists;

        uint32_t n_left = 0u, st = 0u;
        uint32_t nx = 0u;
        int c = -1;
        bool exhausted = false;                                      // no lists left in this frame for my group
        const uint32_t *lp = chain_list;
        for (;;) {
            // ---- groups that finished their list take the next one (longest first)
            if (n_left == 0u && !exhausted) {
                if (c >= 0 && has_slot) s_state[c * 32 + slot] = (uint8_t)st;
                c = -1;
                for (;;) {
                    int oi = 0;
                    if (g == 0) oi = atomicAdd(&s_next, 1);
                    oi = __shfl_sync(gmask, oi, 0, G);
                    if (oi >= nlists) { exhausted = true; break; }
                    const int cc = s_order[oi];
                    const uint32_t b0 = s_b0[TILED ? oi : cc], b1 = s_b1[TILED ? oi : cc];
                    if (b1 == b0) continue;
                    c = cc; n_left = b1 - b0;
                    lp = (TILED ? blk : chain_list + s_lstart[cc]) + b0;
                    st = has_slot ? s_state[cc * 32 + slot] : 0u;
                    nx = (uint32_t)g < n_left ? lp[g] : 0u;
                    break;
                }
            }
            if (!__any_sync(0xFFFFFFFFu, n_left != 0u)) break;
            // ---- a block of up to G symbols per group: masks and positions, one symbol per lane
            const uint32_t m = min((uint32_t)G, n_left);
            const uint32_t ent = (uint32_t)g < m ? nx : 0u;
            const uint2 en = make_uint2(ent & ((1u << kGrpPosBits) - 1u), (uint32_t)((int)ent >> kGrpPosBits) & 0xFFFFu);   // position, residual
            // the next block of the list is fetched while this one is replayed (the longest list of a frame is the
            // critical path of the whole CTA: its loads must not be exposed)
            if ((uint32_t)(G + g) < n_left) nx = lp[G + g];
            const int d = (int)(int16_t)(en.y & 0xFFFFu);
            uint32_t vis = 0u, bts = 0u;
            bool slow = false;
            if ((uint32_t)g < m) {                                   // masks from the table; anything else takes the serial path
                const uint32_t di = (uint32_t)(d + 32);
                const uint2 mk = s_mask[di & 63u];
                slow = di >= 64u || mk.x == 0u;
                if (!slow) { vis = mk.x; bts = mk.y; }
            }
            // every entry of the window lies in frame f: positions relative to the block's first symbol, 32-bit arithmetic
            const uint32_t off0 = __shfl_sync(0xFFFFFFFFu, en.x, lane & 16);
            uint16_t *o0 = dec_f + off0;
            __syncwarp();
            s_blk[g] = make_uint4(vis, bts, en.x - off0, (slow ? 0x80000000u : 0u) | (en.y & 0xFFFFu));
            __syncwarp();
            const uint32_t mm = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, 16));
            if (!__any_sync(0xFFFFFFFFu, slow)) {
#pragma unroll 4
                for (uint32_t k = 0; k < mm; k++) {
                    const uint4 q = s_blk[k];
                    const uint32_t idx = q.z + __popc(q.x & lt_mask);
                    const uint32_t val = (__funnelshift_r(q.y, q.y, rot) & 0x100u) | st;
                    visit_step(reinterpret_cast<uint16_t *>(reinterpret_cast<char *>(o0) + (size_t)idx * 2u), val, st,
                               lut_base + val, q.x & lanebit);
                }
            } else {
                for (uint32_t k = 0; k < mm; k++) {
                    const uint4 q = s_blk[k];
                    if (q.w >> 31) {                                 // uniform inside a group
                        if (has_slot) s_state[c * 32 + slot] = (uint8_t)st;
                        __syncwarp(gmask);
                        if (g == 0) replay_symbol_serial(s_state + c * 32, s_lut, o0 + q.z, (int)(int16_t)(q.w & 0xFFFFu));
                        __syncwarp(gmask);
                        if (has_slot) st = s_state[c * 32 + slot];
                    } else if (q.x & lanebit) {
                        const uint32_t val = (__funnelshift_r(q.y, q.y, rot) & 0x100u) | st;
                        o0[q.z + __popc(q.x & lt_mask)] = (uint16_t)val;
                        st = lut_ld(lut_base + val);
                    }
                }
            }
            lp += m; n_left -= m;
        }
        if (c >= 0 && has_slot) s_state[c * 32 + slot] = (uint8_t)st;
    }
    __syncthreads();
    if (hand_over) {
        const uint4 *st4 = reinterpret_cast<const uint4 *>(s_state);
        uint4 *cout4 = reinterpret_cast<uint4 *>(B.carry_out + coff);
        for (int i = tid; i < nctx * 2; i += THREADS) cout4[i] = st4[i];
    }
}

static int replay_grp_smem(const Layout &L) { return L.ctx_count * (32 + 12 + 2) + 16; }
static int tile_sort_smem(const Layout &L, int pitch) { return kTileSortRows * L.ctx_count * 4 + pitch * 2 + kTiledMaxSamples * 4; }

static bool big_model(const Layout &L) { return L.ctx_count > kMaxListCtx; }

bool ctx_replay_supported(const Layout &L)
{
    return !L.golomb && L.ctx_count <= kMaxBigListCtx;
}

bool ctx_replay_needs_global_state(const Layout &L) { return big_model(L); }

bool ctx_lists_configurable(const Layout &L) { return L.ctx_count <= kMaxListCtx; }

int ctx_scatter_smem_bytes(const Layout &L) { return (kScatterThreads / 32) * L.ctx_count * 4; }
int ctx_scatter_sm_smem_bytes(const Layout &L, int fmt = 0) { return 9 * L.ctx_count * 4 + 3 * ((L.ctx_count + 7) & ~7) * 2 + kScatterSmMaxSamples * (fmt == 1 ? 4 : 8); }

cudaError_t configure_ctx_replay(const Layout &L)
{
    cudaError_t e;
    if (big_model(L)) {
        e = cudaFuncSetAttribute(k_ctx_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, L.ctx_count * 8);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_replay_ctx<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.ctx_count * 10);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_replay_ctx<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.ctx_count * 10);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(k_ctx_scatter<0, 13>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx_scatter_smem_bytes(L));
    }
    if (L.tiled_lists) {
        e = cudaFuncSetAttribute(k_tile_sort<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_sort_smem(L, (L.ctx_count + 8) & ~7));
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(k_tile_sort<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_sort_smem(L, (L.ctx_count + 8) & ~7));
    }
    e = cudaFuncSetAttribute(k_ctx_scatter_sm<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx_scatter_sm_smem_bytes(L));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_ctx_scatter_sm<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx_scatter_sm_smem_bytes(L, 1));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_ctx_scatter_sm<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx_scatter_sm_smem_bytes(L));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_ctx_scatter<0, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx_scatter_smem_bytes(L));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_ctx_scatter<1, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx_scatter_smem_bytes(L));
}

// Golomb-Rice mode: per-context lists of the samples that get a VLC code ({record index, residual | frame << 16})
bool golomb_lists_supported(const Layout &L, int max_tile_samples)
{
    return L.golomb && L.ctx_count <= kMaxListCtx && (L.tiled_lists || max_tile_samples <= kScatterSmMaxSamples) &&
           (unsigned long long)L.rec_per_frame * 2ull <= L.dec_per_frame;
}

void launch_golomb_lists(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s)
{
    const Layout &L = t.layout;
    const int nchains = b.nseg * L.nslices * L.npc;
    dim3 tiles(L.ctiles_per_frame, b.nframes);
    if (L.tiled_lists) {            // tile-sorted lists: k_gr_replay<TILED> walks them tile after tile
        k_tile_sort<true><<<tiles, kTileSortThreads, tile_sort_smem(L, b.tile_tab_pitch), s>>>(t, b);
        return;
    }
    k_ctx_hist<true><<<tiles, kHistThreads, L.ctx_count * 4, s>>>(t, b);
    k_ctx_scan<<<nchains, kScanThreads, L.ctx_count * 8, s>>>(t, b);
    k_ctx_scatter_sm<true, 0><<<tiles, kScatterSmThreads, ctx_scatter_sm_smem_bytes(L), s>>>(t, b);
}

// tile-sorted lists: sort -> decision layout -> replay (a window per tile)
static void launch_tiled_replay(const EncDeviceTables &t, const EncBatch &b, int nchains, cudaStream_t s)
{
    const Layout &L = t.layout;
    dim3 tiles(L.ctiles_per_frame, b.nframes);
    const int n = b.nframes * L.nslices * L.npc;
    k_tile_sort<false><<<tiles, kTileSortThreads, tile_sort_smem(L, b.tile_tab_pitch), s>>>(t, b);
    k_tile_layout<<<(n + 127) / 128, 128, 0, s>>>(t, b);
    int threads = 256;
    if (const char *v = getenv("FFV1B200_REPLAY_THREADS")) threads = atoi(v);
    if (threads == 128)      k_replay_grp<4, 128, true><<<nchains, 128, replay_grp_smem(L), s>>>(t, b, 1);
    else if (threads == 384) k_replay_grp<4, 384, true><<<nchains, 384, replay_grp_smem(L), s>>>(t, b, 1);
    else if (threads == 512) k_replay_grp<4, 512, true><<<nchains, 512, replay_grp_smem(L), s>>>(t, b, 1);
    else                     k_replay_grp<4, 256, true><<<nchains, 256, replay_grp_smem(L), s>>>(t, b, 1);
}

void launch_ctx_replay(const EncDeviceTables &t, const EncBatch &b, int max_tile_samples, uint32_t max_dec_cap, cudaStream_t s)
{
    const Layout &L = t.layout;
    const int nchains = b.nseg * L.nslices * L.npc;
    dim3 tiles(L.ctiles_per_frame, b.nframes);
    const int n = b.nframes * L.nslices * L.npc;
    if (L.tiled_lists) { launch_tiled_replay(t, b, nchains, s); return; }
    k_ctx_hist<false><<<tiles, kHistThreads, L.ctx_count * 4, s>>>(t, b);
    k_ctx_scan<<<nchains, kScanThreads, L.ctx_count * 8, s>>>(t, b);
    k_dec_layout<<<(n + 127) / 128, 128, 0, s>>>(t, b);
    dim3 stiles((L.ctiles_per_frame + kScatterThreads / 32 - 1) / (kScatterThreads / 32), b.nframes);
    if (big_model(L)) {
        // large context model: lists straight from the records (a tile's per-line histograms would not fit shared memory),
        // one CTA per chain replays them against the model in global memory
        k_ctx_scatter<0, 13><<<stiles, kScatterThreads, ctx_scatter_smem_bytes(L), s>>>(t, b);
        if (L.coded_bits <= 10) k_replay_ctx<false, true><<<nchains, kCtxThreads, L.ctx_count * 10, s>>>(t, b);
        else                    k_replay_ctx<true, true><<<nchains, kCtxThreads, L.ctx_count * 10, s>>>(t, b);
        return;
    }
    int grp = 2;
    if (const char *v = getenv("FFV1B200_REPLAY_GROUPS")) grp = atoi(v);
    // 8-bit content (residuals folded to <= 9 bits): two lists per warp (k_replay_grp) over 4-byte list entries
    // (position | residual << 22): a (slice, plane context) decision region must stay below 2^22 entries
    const bool grp_ok = (grp == 1 || grp == 2) && L.coded_bits <= 10 && L.dec_per_frame < 0x7FFFFFFFu && max_dec_cap < (1u << kGrpPosBits);
    // tiles that fit the shared-memory sort (slices up to 352 samples wide) take the staged scatter
    int staged = 1;
    if (const char *v = getenv("FFV1B200_SCATTER")) staged = strcmp(v, "direct") ? 1 : 0;
    if (staged && max_tile_samples <= kScatterSmMaxSamples) {
        if (grp_ok) k_ctx_scatter_sm<false, 1><<<tiles, kScatterSmThreads, ctx_scatter_sm_smem_bytes(L, 1), s>>>(t, b);
        else        k_ctx_scatter_sm<false, 0><<<tiles, kScatterSmThreads, ctx_scatter_sm_smem_bytes(L), s>>>(t, b);
    } else {
        if (grp_ok) k_ctx_scatter<1, 10><<<stiles, kScatterThreads, ctx_scatter_smem_bytes(L), s>>>(t, b);
        else        k_ctx_scatter<0, 10><<<stiles, kScatterThreads, ctx_scatter_smem_bytes(L), s>>>(t, b);
    }
    int window = 3;
    if (const char *v = getenv("FFV1B200_REPLAY_WINDOW")) { window = atoi(v); if (window < 1) window = 1 << 20; }
    if (grp == 1 && grp_ok) k_replay_grp<4, 1024, false><<<nchains, 1024, replay_grp_smem(L), s>>>(t, b, window);
    else if (grp == 2 && grp_ok) {
        int threads = 256;
        if (const char *v = getenv("FFV1B200_REPLAY_THREADS")) threads = atoi(v);
        if (threads == 128)      k_replay_grp<4, 128, false><<<nchains, 128, replay_grp_smem(L), s>>>(t, b, window);
        else if (threads == 256) k_replay_grp<4, 256, false><<<nchains, 256, replay_grp_smem(L), s>>>(t, b, window);
        else                     k_replay_grp<4, 512, false><<<nchains, 512, replay_grp_smem(L), s>>>(t, b, window);
    }
    else if (L.coded_bits <= 10) k_replay_ctx<false, false><<<nchains, kCtxThreads, L.ctx_count * 42, s>>>(t, b);
    else                    k_replay_ctx<true, false><<<nchains, kCtxThreads, L.ctx_count * 42, s>>>(t, b);
}

} // namespace ffv1
