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
//   k_replay_ctx   one thread per (chain, context) list, longest lists first: put_symbol_inline's binarisation
//                  (ffv1enc.c:185-231) on a 32-byte state held in shared memory; writes p | bit<<8 at the recorded
//                  positions of the decision stream that k_rangecode consumes
#include "ffv1_enc_kernels.cuh"

namespace ffv1 {

constexpr int kHistThreads = 256;
constexpr int kScatterThreads = 32 * kCtxTileLines;
constexpr int kCtxThreads = 128;
constexpr int kMaxListCtx = 1024;

__device__ __forceinline__ uint32_t cr_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t n = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

__device__ __forceinline__ uint32_t decisions_of(int d)
{
    return d ? (uint32_t)(2 * (31 - __clz((uint32_t)abs(d))) + 3) : 1u;       // put_symbol_inline: 1 or 2e+3
}

// ------------------------------------------------------------------------------------------------ k_ctx_hist
__global__ void __launch_bounds__(kHistThreads) k_ctx_hist(const EncDeviceTables T, const EncBatch B)
{
    __shared__ uint32_t s_hist[kMaxListCtx];
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
__global__ void __launch_bounds__(256) k_ctx_scan(const EncDeviceTables T, const EncBatch B)
{
    __shared__ uint32_t s_total[kMaxListCtx];
    __shared__ uint32_t s_start[kMaxListCtx];
    __shared__ uint32_t s_warp[8];
    const Layout &L = T.layout;
    const int chain = blockIdx.x;
    const int pc = chain % L.npc, s = (chain / L.npc) % L.nslices, seg = chain / (L.npc * L.nslices);
    const SliceGeom &g = T.slices[s];
    const int f0 = B.seg_first[seg], f1 = B.seg_first[seg + 1];
    const int nctx = L.ctx_count, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t0 = g.ct_first[pc], nt = g.ct_count[pc];
    for (int c = tid; c < nctx; c += 256) {
        uint32_t run = 0;
        for (int f = f0; f < f1; f++) {
            uint32_t *h = B.ctx_hist + ((size_t)f * L.ctiles_per_frame + t0) * nctx + c;
            for (int t = 0; t < nt; t++) {
                const uint32_t v = h[(size_t)t * nctx];
                h[(size_t)t * nctx] = run;
                run += v;
            }
        }
        s_total[c] = run;
    }
    __syncthreads();
    // exclusive scan over the contexts: 4 consecutive contexts per thread
    {
        uint32_t v[4], sum = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { const int c = tid * 4 + k; v[k] = c < nctx ? s_total[c] : 0u; sum += v[k]; }
        const uint32_t incl = cr_incl_scan(sum, lane);
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t base = incl - sum;
        for (int w = 0; w < warp; w++) base += s_warp[w];
#pragma unroll
        for (int k = 0; k < 4; k++) { const int c = tid * 4 + k; if (c < nctx) s_start[c] = base; base += v[k]; }
    }
    __syncthreads();
    for (int c = tid; c < nctx; c += 256) {
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
    if (idx >= B.nframes * L.nslices * L.npc) return;
    const int pc = idx % L.npc, s = (idx / L.npc) % L.nslices, f = idx / (L.npc * L.nslices);
    const SliceGeom &g = T.slices[s];
    const int nl = g.pc_nlines[pc];
    if (!nl) return;
    const int32_t *my_lines = T.pc_lines + g.pc_line_first[pc];
    uint32_t *line_pos = B.line_pos + (size_t)f * L.lines_per_frame + g.line_first;
    uint32_t *run_cnt = B.run_cnt + (size_t)f * L.runs_per_frame + g.run_first;
    uint32_t pos = 0, run_start = 0, cur_run = T.lines[g.line_first + my_lines[0]].run;
    unsigned long long ndec = 0;
    for (int i = 0; i < nl; i++) {
        const int line = my_lines[i];
        const uint32_t run = T.lines[g.line_first + line].run;
        if (run != cur_run) {                                   // a run of this plane context ended: next one starts 16 B aligned
            run_cnt[cur_run] = pos - run_start;
            pos = (pos + 7u) & ~7u;
            run_start = pos; cur_run = run;
        }
        const uint32_t nd = line_pos[line];
        line_pos[line] = pos;
        pos += nd;
        ndec += nd;
    }
    run_cnt[cur_run] = pos - run_start;
    atomicAdd(&B.status[3], ndec);
    if (pos + 8u > g.dec_cap[pc]) {
        const unsigned long long ns = g.pc_samples[pc];
        atomicMax(&B.status[0], ((unsigned long long)(pos + 8u) * 256ull + ns - 1) / ns + 1ull);
    }
}

// ------------------------------------------------------------------------------------------------ k_ctx_scatter
__global__ void __launch_bounds__(kScatterThreads) k_ctx_scatter(const EncDeviceTables T, const EncBatch B)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *s_off = reinterpret_cast<uint32_t *>(smem_raw);           // [kCtxTileLines][nctx]
    const Layout &L = T.layout;
    if (B.status[0]) return;
    const CtxTile ct = T.ctiles[blockIdx.x];
    const int f = blockIdx.y;
    const SliceGeom &g = T.slices[ct.slice];
    const int nctx = L.ctx_count, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int i = tid; i < kCtxTileLines * nctx; i += kScatterThreads) s_off[i] = 0;
    __syncthreads();
    const uint32_t *rec_slice = B.rec + (size_t)f * L.rec_per_frame + g.rec_first;
    const int32_t *my_lines = T.pc_lines + g.pc_line_first[ct.pc] + ct.first;
    int line = 0;
    LineDesc ld;
    ld.w = 0; ld.rec_off = 0;
    if (warp < ct.nlines) { line = my_lines[warp]; ld = T.lines[g.line_first + line]; }
    const uint32_t *recp = rec_slice + ld.rec_off;
    uint32_t *my_off = s_off + warp * nctx;
    // ---- phase 1: context histogram of every line (one warp per line, no atomics: the row belongs to the warp)
    for (int x0 = 0; x0 < ld.w; x0 += 32) {
        const bool act = x0 + lane < ld.w;
        const uint32_t ctx = act ? recp[x0 + lane] >> 16 : 0u;
        const uint32_t grp = __match_any_sync(0xFFFFFFFFu, act ? ctx : 0x10000u + lane);
        if (act && (grp & lt_mask) == 0u) my_off[ctx] += __popc(grp);
        __syncwarp();
    }
    __syncthreads();
    // ---- phase 2: line histograms -> list positions (tile base from k_ctx_scan + lines above inside the tile)
    {
        const int seg = B.frame_seg[f];
        const int f0 = B.seg_first[seg], seglen = B.seg_first[seg + 1] - f0;
        const int chain = (seg * L.nslices + ct.slice) * L.npc + ct.pc;
        const uint32_t *tile_base = B.ctx_hist + ((size_t)f * L.ctiles_per_frame + blockIdx.x) * nctx;
        const uint32_t *lstart = B.list_start + (size_t)chain * nctx;
        for (int c = tid; c < nctx; c += kScatterThreads) {
            uint32_t base = tile_base[c] + lstart[c];
            for (int l = 0; l < ct.nlines; l++) {
                const uint32_t v = s_off[l * nctx + c];
                s_off[l * nctx + c] = base;
                base += v;
            }
        }
        __syncthreads();
        // ---- phase 3: stable scatter, coding order preserved inside every context
        if (warp < ct.nlines) {
            uint2 *list = B.lists + (size_t)f0 * L.samples_per_frame + (size_t)seglen * g.list_off[ct.pc];
            uint32_t pos = B.line_pos[(size_t)f * L.lines_per_frame + g.line_first + line];
            for (int x0 = 0; x0 < ld.w; x0 += 32) {
                const bool act = x0 + lane < ld.w;
                const uint32_t r = act ? recp[x0 + lane] : 0u;
                const uint32_t ctx = r >> 16;
                const uint32_t nd = act ? decisions_of((int)(int16_t)(r & 0xFFFFu)) : 0u;
                const uint32_t incl = cr_incl_scan(nd, lane);
                const uint32_t grp = __match_any_sync(0xFFFFFFFFu, act ? ctx : 0x10000u + lane);
                const uint32_t rank = __popc(grp & lt_mask);
                if (act) list[my_off[ctx] + rank] = make_uint2(pos + incl - nd, (r & 0xFFFFu) | ((uint32_t)f << 16));
                __syncwarp();
                if (act && rank == 0u) my_off[ctx] += __popc(grp);
                __syncwarp();
                pos += __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ k_replay_ctx
constexpr int kCtxStateStride = kCtxThreads + 4;

// Every thread walks per-context lists (longest first, taken from a shared counter); the 32 lanes of a warp advance in
// lock step, one symbol of each lane's current list per iteration, so the loads of all decisions of 32 symbols are in
// flight together: slots visited by one symbol are distinct for e <= 9, hence phase 1 loads every probability, phase 2
// looks every successor up, phase 3 stores states and emits the decisions.
template <int MAXE>
__global__ void __launch_bounds__(kCtxThreads) k_replay_ctx(const EncDeviceTables T, const EncBatch B)
{
    __shared__ uint8_t s_state[32 * kCtxStateStride];
    __shared__ uint8_t s_lut[512];
    __shared__ int s_next;
    const Layout &L = T.layout;
    const int tid = threadIdx.x;
    for (int i = tid; i < 512; i += kCtxThreads) s_lut[i] = T.trans_lut[i];
    if (tid == 0) s_next = 0;
    if (B.status[0]) return;
    const int chain = blockIdx.x;
    const int pc = chain % L.npc, s = (chain / L.npc) % L.nslices, seg = chain / (L.npc * L.nslices);
    const SliceGeom &g = T.slices[s];
    const int f0 = B.seg_first[seg], f1 = B.seg_first[seg + 1];
    const int nctx = L.ctx_count;
    const bool key = B.frame_key[f0] != 0;
    const bool hand_over = f1 == B.nframes;
    const size_t coff = ((size_t)s * L.npc + pc) * ((size_t)nctx * 32);
    const uint4 *cin = reinterpret_cast<const uint4 *>(B.carry_in + coff);
    uint4 *cout = reinterpret_cast<uint4 *>(B.carry_out + coff);
    // contexts without symbols in this batch still hand their state to the next one
    if (hand_over)
        for (int i = tid; i < nctx * 2; i += kCtxThreads)
            cout[i] = key ? make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u) : cin[i];
    __syncthreads();

    const uint2 *chain_list = B.lists + (size_t)f0 * L.samples_per_frame + (size_t)(f1 - f0) * g.list_off[pc];
    const uint32_t *lstart = B.list_start + (size_t)chain * nctx;
    const uint32_t *lcount = B.list_count + (size_t)chain * nctx;
    const uint16_t *order = B.list_order + (size_t)chain * nctx;
    uint16_t *dec_pc = B.dec + g.dec_off[pc];
    uint8_t *S = s_state + tid;
    constexpr int ST = kCtxStateStride;

    uint32_t n = 0, j = 0;
    int c = -1;
    bool done = false;
    const uint2 *lp = chain_list;
    uint2 nx = make_uint2(0u, 0u);
    for (;;) {
        if (!done && j == n) {
            if (c >= 0 && hand_over) {                          // the finished context's state goes to the next batch
                uint32_t w[8];
#pragma unroll
                for (int k = 0; k < 8; k++)
                    w[k] = (uint32_t)S[(4 * k) * ST] | ((uint32_t)S[(4 * k + 1) * ST] << 8) |
                           ((uint32_t)S[(4 * k + 2) * ST] << 16) | ((uint32_t)S[(4 * k + 3) * ST] << 24);
                cout[c * 2] = make_uint4(w[0], w[1], w[2], w[3]);
                cout[c * 2 + 1] = make_uint4(w[4], w[5], w[6], w[7]);
            }
            const int oi = atomicAdd(&s_next, 1);
            c = oi < nctx ? (int)order[oi] : -1;
            n = c >= 0 ? lcount[c] : 0u;
            if (n == 0u) { done = true; c = -1; }               // lists are ordered longest first
            else {
                if (key) {
#pragma unroll
                    for (int k = 0; k < 32; k++) S[k * ST] = 128;
                } else {
                    const uint4 a = cin[c * 2], b = cin[c * 2 + 1];
                    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int k = 0; k < 32; k++) S[k * ST] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
                }
                lp = chain_list + lstart[c];
                j = 0;
                nx = lp[0];
            }
        }
        if (__all_sync(0xFFFFFFFFu, done)) break;
        const bool act = !done;
        const uint2 en = nx;
        if (act && j + 1u < n) nx = lp[j + 1u];
        j += act ? 1u : 0u;

        const int d = act ? (int)(int16_t)(en.y & 0xFFFFu) : 0;
        const uint32_t a = (uint32_t)abs(d);
        const int ex = 31 - __clz(a | 1u);
        const bool slow = act && d != 0 && ex > MAXE;           // only deeper-than-10-bit content can get here
        const bool nz = act && d != 0 && !slow;
        const int el = nz ? ex : -1;
        const int emax = __reduce_max_sync(0xFFFFFFFFu, el);
        uint16_t *o = dec_pc + (size_t)(en.y >> 16) * L.dec_per_frame + en.x;
        const uint32_t sign = d < 0 ? 0x100u : 0u;

        uint32_t p0 = 0, ps = 0, pu[MAXE + 1], pm[MAXE];
        if (act) p0 = S[0];
        if (emax >= 0) {
#pragma unroll
            for (int i = 0; i <= MAXE; i++) {
                if (i > emax) break;
                pu[i] = (i <= el) ? S[(1 + i) * ST] : 0u;
            }
#pragma unroll
            for (int i = 0; i < MAXE; i++) {
                if (i >= emax) break;
                pm[i] = (i < el) ? S[(22 + i) * ST] : 0u;
            }
            if (nz) ps = S[(11 + ex) * ST];
        }
        uint32_t n0 = 0, ns = 0, nu[MAXE + 1], nm[MAXE];
        const uint32_t bit0 = d == 0 ? 0x100u : 0u;             // "is zero" flag
        if (act) n0 = s_lut[bit0 + p0];
        if (emax >= 0) {
#pragma unroll
            for (int i = 0; i <= MAXE; i++) {
                if (i > emax) break;
                nu[i] = (i <= el) ? s_lut[(i < el ? 0x100u : 0u) + pu[i]] : 0u;
            }
#pragma unroll
            for (int i = 0; i < MAXE; i++) {
                if (i >= emax) break;
                nm[i] = (i < el) ? s_lut[(((a >> i) & 1u) << 8) + pm[i]] : 0u;
            }
            if (nz) ns = s_lut[sign + ps];
        }
        if (act) { S[0] = (uint8_t)n0; o[0] = (uint16_t)(p0 | bit0); }
        if (emax >= 0) {
#pragma unroll
            for (int i = 0; i <= MAXE; i++) {
                if (i > emax) break;
                if (i <= el) { S[(1 + i) * ST] = (uint8_t)nu[i]; o[1 + i] = (uint16_t)(pu[i] | (i < el ? 0x100u : 0u)); }
            }
#pragma unroll
            for (int i = 0; i < MAXE; i++) {
                if (i >= emax) break;
                if (i < el) { S[(22 + i) * ST] = (uint8_t)nm[i]; o[2 * ex + 1 - i] = (uint16_t)(pm[i] | (((a >> i) & 1u) << 8)); }
            }
            if (nz) { S[(11 + ex) * ST] = (uint8_t)ns; o[2 * ex + 2] = (uint16_t)(ps | sign); }
        }
        if (slow) {
            // large magnitudes (ffv1enc.c:217-228): slots 1+9 and 22+9 repeat, the decisions are walked one by one
            int k = 1;
            for (int i = 0; i <= ex; i++, k++) {
                uint8_t *q = S + (1 + min(i, 9)) * ST;
                const uint32_t p = *q, bit = i < ex ? 0x100u : 0u;
                o[k] = (uint16_t)(p | bit);
                *q = s_lut[bit + p];
            }
            for (int i = ex - 1; i >= 0; i--, k++) {
                uint8_t *q = S + (22 + min(i, 9)) * ST;
                const uint32_t p = *q, bit = ((a >> i) & 1u) << 8;
                o[k] = (uint16_t)(p | bit);
                *q = s_lut[bit + p];
            }
            uint8_t *q = S + (11 + min(ex, 10)) * ST;
            const uint32_t p = *q;
            o[k] = (uint16_t)(p | sign);
            *q = s_lut[sign + p];
        }
    }
}

bool ctx_replay_supported(const Layout &L)
{
    return !L.golomb && L.ctx_count <= kMaxListCtx;
}

int ctx_scatter_smem_bytes(const Layout &L) { return kCtxTileLines * L.ctx_count * 4; }

cudaError_t configure_ctx_replay(const Layout &L)
{
    return cudaFuncSetAttribute(k_ctx_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx_scatter_smem_bytes(L));
}

void launch_ctx_replay(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s)
{
    const Layout &L = t.layout;
    const int nchains = b.nseg * L.nslices * L.npc;
    dim3 tiles(L.ctiles_per_frame, b.nframes);
    k_ctx_hist<<<tiles, kHistThreads, 0, s>>>(t, b);
    k_ctx_scan<<<nchains, 256, 0, s>>>(t, b);
    const int n = b.nframes * L.nslices * L.npc;
    k_dec_layout<<<(n + 127) / 128, 128, 0, s>>>(t, b);
    k_ctx_scatter<<<tiles, kScatterThreads, ctx_scatter_smem_bytes(L), s>>>(t, b);
    if (L.coded_bits <= 8) k_replay_ctx<7><<<nchains, kCtxThreads, 0, s>>>(t, b);
    else                   k_replay_ctx<9><<<nchains, kCtxThreads, 0, s>>>(t, b);
}

} // namespace ffv1
