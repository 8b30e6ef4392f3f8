// ffv1_enc_kernels.cu -- hand-written sm_100a kernels of the FFV1 encode hot path.
//
//   k_pixel      per-pixel pass: sample fetch (+RCT), slice-local neighbourhood, get_context, median predictor,
//                residual, sign flip, fold  ->  (context<<16 | diff) records in coding order + decisions per line.
//                Reference: encode_plane / encode_rgb_frame / encode_line up to the symbol call
//                (ffv1enc.c:373-473, 271-321; ffv1.h:148-190).  HBM bound: 1-2 B read + 4 B written per sample.
//   k_scan_*     exclusive scans that place every line's decisions in one compact decision stream.
//   k_replay     adaptive-state replay: put_symbol_inline's binarisation (ffv1enc.c:185-231) applied to the
//                per-context 32-byte states, one warp per (GOP segment, slice, plane context), lane = state slot.
//                Emits the (probability, bit) pair of every binary decision; carries state across frames of a GOP
//                (the reference's "P-frames": state is only reset on keyframes, ffv1enc.c:1171-1172).
//   k_rangecode  put_rac / renorm_encoder / ff_rac_terminate (rangecoder.h:52-102, rangecoder.c:104-116):
//                one sequential interval coder per (frame, slice), thousands in flight.
//   k_pack_*     packet assembly (ffv1enc.c:1326-1354): slice sizes -> offsets, copy, 24-bit length, CRC-32
//                computed in parallel chunks and combined in GF(2) (libavutil/crc.c semantics).
#include "ffv1_enc_kernels.cuh"
#include <cstdio>

namespace ffv1 {

// =================================================================================================
// k_pixel
// =================================================================================================
struct PixelSmem {
    // dynamic shared memory carve-up (all int16 / uint32):
    //   quant[ctx_inputs][256] | cnt[planes_in_tile][rows] | S[planes_in_tile][rows+2][kPixelRowElems]
};

__device__ __forceinline__ int mid3(int a, int b, int c)
{
    // median of three (mathops.h:95-119)
    return max(min(a, b), min(max(a, b), c));
}

template <int SRC>
__device__ __forceinline__ int raw_planar(const Layout &L, const uint8_t *const *pl, const int32_t *ls,
                                          const SliceGeom &g, int p, int x, int y)
{
    const PlaneInfo &pi = L.plane[p];
    const uint8_t *src = pl[pi.src_plane] + (size_t)(g.py0[p] + y) * ls[pi.src_plane];
    if (SRC == SRC_PLANAR8)
        return src[(g.px0[p] + x) * pi.pstep + pi.poff];
    // 16-bit container: LSB-aligned 9/10-bit values are used as they are, 16-bit ones wrap into the int16
    // line buffer exactly like the reference's int16_t sample_buffer (ffv1enc.c:396-403, ffv1.h:111)
    unsigned v = *reinterpret_cast<const uint16_t *>(src + 2 * (g.px0[p] + x));
    return (int)(int16_t)(v >> L.sample_shift);
}

template <int SRC>
__device__ __forceinline__ void raw_rgb(const Layout &L, const uint8_t *const *pl, const int32_t *ls,
                                        const SliceGeom &g, int x, int y, int out[4])
{
    // ffv1enc.c:431-458: b,g,r[,a] from packed BGRA or from data[0],data[1],data[2] (the reference's naming; for
    // GBRP that makes plane 1 the RCT base channel), then b-=g; r-=g; g+=(b+r)>>2; b+=off; r+=off
    int b, gg, r, a = 0;
    if (SRC == SRC_RGB32) {
        unsigned v = *reinterpret_cast<const uint32_t *>(pl[0] + (size_t)(g.y0 + y) * ls[0] + 4 * (g.x0 + x));
        b = v & 0xFF; gg = (v >> 8) & 0xFF; r = (v >> 16) & 0xFF; a = v >> 24;
    } else {
        b  = *reinterpret_cast<const uint16_t *>(pl[0] + (size_t)(g.y0 + y) * ls[0] + 2 * (g.x0 + x));
        gg = *reinterpret_cast<const uint16_t *>(pl[1] + (size_t)(g.y0 + y) * ls[1] + 2 * (g.x0 + x));
        r  = *reinterpret_cast<const uint16_t *>(pl[2] + (size_t)(g.y0 + y) * ls[2] + 2 * (g.x0 + x));
    }
    b -= gg; r -= gg;
    gg += (b + r) >> 2;
    b += L.rct_offset; r += L.rct_offset;
    out[0] = (int16_t)gg; out[1] = (int16_t)b; out[2] = (int16_t)r; out[3] = (int16_t)a;
}

template <int SRC, int NIN>
__global__ void __launch_bounds__(kPixelThreads)
k_pixel(const EncDeviceTables T, const EncBatch B)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Layout &L = T.layout;
    const TileDesc td = T.tiles[blockIdx.x];
    const int f = blockIdx.y;
    const SliceGeom &g = T.slices[td.slice];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kPixelThreads / 32;
    constexpr bool RGB = (SRC == SRC_RGB32 || SRC == SRC_GBRP16);

    int16_t *s_quant = reinterpret_cast<int16_t *>(smem_raw);
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_quant + NIN * 256);
    const int npl = td.nplanes, nrows = td.nrows;
    int16_t *S = reinterpret_cast<int16_t *>(s_cnt + 4 * kTileRows);
    const int plane_stride = (kTileRows + 2) * kPixelRowElems;

    for (int i = tid; i < NIN * 256; i += kPixelThreads) s_quant[i] = T.quant[i];
    for (int i = tid; i < 4 * kTileRows; i += kPixelThreads) s_cnt[i] = 0;

    const uint8_t *pl[4];
    int32_t ls[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { pl[i] = B.planes[f * 4 + i]; ls[i] = B.linesize[i]; }

    const int w = g.pw[td.plane];
    const int bits = L.coded_bits;
    uint32_t *rec_slice = B.rec + (size_t)f * L.rec_per_frame + g.rec_first;
    const LineDesc *lines = T.lines + g.line_first;

    for (int cx0 = 0; cx0 < w; cx0 += kPixelChunk) {
        const int cw = min(kPixelChunk, w - cx0);
        __syncthreads();
        // ---- stage rows y0-2 .. y0+nrows-1, columns cx0-2 .. cx0+cw, applying the slice-local edge rules
        //      (ffv1enc.c:381-388): outside rows are 0, S[y][-1] = S[y-1][0], S[y][-2] = 0,
        //      S[y][w] = S[y][w-1].
        for (int rr = warp; rr < nrows + 2; rr += NW) {
            const int y = td.y0 + rr - 2;
            for (int cc = lane; cc < cw + 3; cc += 32) {
                int x = cx0 + cc - 2;
                int yy = y;
                bool zero = (y < 0) || (x == -2);
                if (x == -1) { x = 0; yy = y - 1; zero = zero || (yy < 0); }
                if (x >= w) x = w - 1;
                const int col = cc - 2 + kPixelPadL;
                if (!RGB) {
                    int v = zero ? 0 : raw_planar<SRC>(L, pl, ls, g, td.plane, x, yy);
                    S[rr * kPixelRowElems + col] = (int16_t)v;
                } else {
                    int v[4] = {0, 0, 0, 0};
                    if (!zero) raw_rgb<SRC>(L, pl, ls, g, x, yy, v);
#pragma unroll
                    for (int p = 0; p < 4; p++)
                        if (p < npl) S[p * plane_stride + rr * kPixelRowElems + col] = (int16_t)v[p];
                }
            }
        }
        __syncthreads();
        // ---- context / prediction / residual, 4 consecutive samples per thread
        const int groups = (cw + 3) >> 2;
        for (int r = warp; r < nrows; r += NW) {
            for (int pp = 0; pp < npl; pp++) {
                const int16_t *cur = S + pp * plane_stride + (r + 2) * kPixelRowElems + kPixelPadL;
                const int16_t *top = cur - kPixelRowElems;
                const int16_t *top2 = top - kPixelRowElems;
                const LineDesc ld = lines[td.line_first + r * td.line_step + pp];
                uint32_t *rec_line = rec_slice + ld.rec_off + cx0;
                uint32_t cnt = 0;
                for (int gi = lane; gi < groups; gi += 32) {
                    const int x = gi << 2;
                    // cur[x-2..x+3], top[x-1..x+4], top2[x..x+3]
                    int c_[6], t_[6], tt_[4];
                    {
                        const uint32_t a0 = *reinterpret_cast<const uint32_t *>(cur + x - 2);
                        const uint2 a1 = *reinterpret_cast<const uint2 *>(cur + x);
                        c_[0] = (int16_t)(a0 & 0xFFFF); c_[1] = (int16_t)(a0 >> 16);
                        c_[2] = (int16_t)(a1.x & 0xFFFF); c_[3] = (int16_t)(a1.x >> 16);
                        c_[4] = (int16_t)(a1.y & 0xFFFF); c_[5] = (int16_t)(a1.y >> 16);
                        const uint32_t b0 = *reinterpret_cast<const uint32_t *>(top + x - 2);
                        const uint2 b1 = *reinterpret_cast<const uint2 *>(top + x);
                        const uint32_t b2 = *reinterpret_cast<const uint32_t *>(top + x + 4);
                        t_[0] = (int16_t)(b0 >> 16);
                        t_[1] = (int16_t)(b1.x & 0xFFFF); t_[2] = (int16_t)(b1.x >> 16);
                        t_[3] = (int16_t)(b1.y & 0xFFFF); t_[4] = (int16_t)(b1.y >> 16);
                        t_[5] = (int16_t)(b2 & 0xFFFF);
                        if (NIN == 5) {
                            const uint2 d1 = *reinterpret_cast<const uint2 *>(top2 + x);
                            tt_[0] = (int16_t)(d1.x & 0xFFFF); tt_[1] = (int16_t)(d1.x >> 16);
                            tt_[2] = (int16_t)(d1.y & 0xFFFF); tt_[3] = (int16_t)(d1.y >> 16);
                        }
                    }
                    uint32_t out[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int X = c_[i + 2], Lf = c_[i + 1], LL = c_[i];
                        const int LT = t_[i], Tp = t_[i + 1], RT = t_[i + 2];
                        int ctx = s_quant[(Lf - LT) & 0xFF] + s_quant[256 + ((LT - Tp) & 0xFF)] + s_quant[512 + ((Tp - RT) & 0xFF)];
                        if (NIN == 5)
                            ctx += s_quant[768 + ((LL - Lf) & 0xFF)] + s_quant[1024 + ((tt_[i] - Tp) & 0xFF)];
                        int diff = X - mid3(Lf, Lf + Tp - LT, Tp);
                        if (ctx < 0) { ctx = -ctx; diff = -diff; }
                        diff = (diff << (32 - bits)) >> (32 - bits);        // fold(): sign-extend the low `bits` bits
                        const int ad = abs(diff);
                        if (x + i < cw) cnt += diff ? (uint32_t)(2 * (31 - __clz(ad)) + 3) : 1u;
                        out[i] = ((uint32_t)ctx << 16) | ((uint32_t)diff & 0xFFFFu);
                    }
                    if (x + 3 < cw) {
                        *reinterpret_cast<uint4 *>(rec_line + x) = make_uint4(out[0], out[1], out[2], out[3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            if (x + i < cw) rec_line[x + i] = out[i];
                    }
                }
                cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
                if (lane == 0) s_cnt[pp * kTileRows + r] += cnt;     // each (plane,row) is owned by one warp
            }
        }
    }
    __syncthreads();
    // decisions per line (range-coder mode; ignored by the golomb path)
    uint32_t *cnt_slice = B.line_cnt + (size_t)f * L.lines_per_frame + g.line_first;
    for (int i = tid; i < npl * nrows; i += kPixelThreads) {
        const int pp = i / nrows, r = i - pp * nrows;
        cnt_slice[td.line_first + r * td.line_step + pp] = s_cnt[pp * kTileRows + r];
    }
}

int pixel_smem_bytes(const Layout &L)
{
    const int planes = L.rgb ? L.nplanes : 1;
    return L.ctx_inputs * 256 * 2 + 4 * kTileRows * 4 + planes * (kTileRows + 2) * kPixelRowElems * 2;
}

template <int SRC, int NIN>
static void launch_pixel_t(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s)
{
    dim3 grid(t.layout.tiles_per_frame, b.nframes);
    k_pixel<SRC, NIN><<<grid, kPixelThreads, pixel_smem_bytes(t.layout), s>>>(t, b);
}

void launch_pixel(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s)
{
    const bool five = t.layout.ctx_inputs == 5;
    switch (t.layout.src_kind) {
    case SRC_PLANAR8:  five ? launch_pixel_t<SRC_PLANAR8, 5>(t, b, s)  : launch_pixel_t<SRC_PLANAR8, 3>(t, b, s);  break;
    case SRC_PLANAR16: five ? launch_pixel_t<SRC_PLANAR16, 5>(t, b, s) : launch_pixel_t<SRC_PLANAR16, 3>(t, b, s); break;
    case SRC_RGB32:    five ? launch_pixel_t<SRC_RGB32, 5>(t, b, s)    : launch_pixel_t<SRC_RGB32, 3>(t, b, s);    break;
    default:           five ? launch_pixel_t<SRC_GBRP16, 5>(t, b, s)   : launch_pixel_t<SRC_GBRP16, 3>(t, b, s);   break;
    }
}

// =================================================================================================
// k_scan_lines / k_scan_slices
// =================================================================================================
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t n = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

// one warp per (frame, slice): exclusive scan of the slice's per-line decision counts
__global__ void __launch_bounds__(256) k_scan_lines(const EncDeviceTables T, const EncBatch B)
{
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int n = B.nframes * T.layout.nslices;
    if (gw >= n) return;
    const int f = gw / T.layout.nslices, s = gw - f * T.layout.nslices;
    const SliceGeom &g = T.slices[s];
    const uint32_t *cnt = B.line_cnt + (size_t)f * T.layout.lines_per_frame + g.line_first;
    uint32_t *off = B.line_off + (size_t)f * T.layout.lines_per_frame + g.line_first;
    uint32_t run = 0;
    for (int i0 = 0; i0 < g.nlines; i0 += 32) {
        const int i = i0 + lane;
        const uint32_t v = i < g.nlines ? cnt[i] : 0;
        const uint32_t inc = warp_incl_scan(v, lane);
        if (i < g.nlines) off[i] = run + inc - v;
        run += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
    if (lane == 0) B.slice_ndec[gw] = run;
}

// single CTA: exclusive scan over all (frame, slice) of the 8-entry-aligned decision counts
__global__ void __launch_bounds__(1024) k_scan_slices(const EncDeviceTables T, const EncBatch B)
{
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_run;
    const int n = B.nframes * T.layout.nslices;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + tid;
        unsigned long long v = i < n ? (((unsigned long long)B.slice_ndec[i] + 7ull) & ~7ull) : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = s_warp[lane], winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, winc, d);
                if (lane >= d) winc += o;
            }
            s_warp[lane] = winc - w;       // exclusive prefix of the warp totals
        }
        __syncthreads();
        const unsigned long long base = s_run + s_warp[warp];
        if (i < n) B.slice_base[i] = base + inc - v;
        __syncthreads();
        if (tid == 1023) s_run = base + inc;
        __syncthreads();
    }
    if (tid == 0) {
        B.status[3] = s_run;                                   // entries used
        if (s_run > B.dec_capacity) B.status[0] = s_run;       // overflow: the host grows dec[] and re-runs
    }
}

void launch_scan(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s)
{
    const int n = b.nframes * t.layout.nslices;
    k_scan_lines<<<(n * 32 + 255) / 256, 256, 0, s>>>(t, b);
    k_scan_slices<<<1, 1024, 0, s>>>(t, b);
}

// =================================================================================================
// k_replay: adaptive state replay, one warp per (GOP segment, slice, plane context)
// =================================================================================================
// Lane <-> state slot permutation.  put_symbol_inline visits the slots in the order
//   0 | 1..e+1 | 22+e-1 .. 22 | 11+e                                  (ffv1enc.c:202-229, e <= 9)
// With slots 22..31 stored in REVERSE order in lanes 11..20 and slots 11..21 in lanes 21..31, that order is
// increasing in lane index for every e, so the position of a lane's decision inside the symbol is simply
// popc(visited_mask & lanes_below).
__device__ __forceinline__ int lane_of_slot(int slot)
{
    return slot <= 10 ? slot : (slot >= 22 ? 42 - slot : slot + 10);
}

struct SymMasks { uint32_t visit, bits; int nd; };

// masks in lane space for a symbol with e <= 9
__device__ __forceinline__ SymMasks symbol_masks(int d)
{
    SymMasks m;
    if (d == 0) { m.visit = 1u; m.bits = 1u; m.nd = 1; return m; }
    const uint32_t a = (uint32_t)abs(d);
    const int e = 31 - __clz(a);
    const uint32_t ones_e = (1u << e) - 1u;
    const uint32_t mant = e ? (__brev(a & ones_e) >> (32 - e)) : 0u;    // bit k <- bit e-1-k of a
    m.visit = 1u | (((2u << e) - 1u) << 1) | (ones_e << (21 - e)) | (1u << (21 + e));
    m.bits = (ones_e << 1) | (mant << (21 - e)) | ((d < 0 ? 1u : 0u) << (21 + e));
    m.nd = 2 * e + 3;
    return m;
}

template <bool SMEM_STATE, bool HIGH_E>
__global__ void __launch_bounds__(kReplayWarps * 32)
k_replay(const EncDeviceTables T, const EncBatch B)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Layout &L = T.layout;
    uint8_t *s_lut = smem_raw;                               // [0..255] zero_state, [256..511] one_state
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_lut[i] = T.trans_lut[i];
    __syncthreads();
    if (B.status[0]) return;                                 // decision stream does not fit: host will retry

    const int nchains = B.nseg * L.nslices * L.npc;
    const int chain = blockIdx.x * kReplayWarps + warp;
    if (chain >= nchains) return;
    const int pc = chain % L.npc;
    const int s = (chain / L.npc) % L.nslices;
    const int seg = chain / (L.npc * L.nslices);
    const SliceGeom &g = T.slices[s];
    const int f0 = B.seg_first[seg], f1 = B.seg_first[seg + 1];
    const size_t state_bytes = (size_t)L.ctx_count * 32;

    uint8_t *st;
    if (SMEM_STATE) st = smem_raw + 512 + (size_t)warp * state_bytes;
    else            st = B.state_seg + ((size_t)(seg * L.nslices + s) * L.npc + pc) * state_bytes;

    // ---- initial state: 128 on keyframes (ffv1.c:177-202), else carried over from the previous batch
    {
        const size_t coff = ((size_t)s * L.npc + pc) * state_bytes;
        const bool key = B.frame_key[f0];
        uint32_t *st4 = reinterpret_cast<uint32_t *>(st);
        const uint32_t *in4 = reinterpret_cast<const uint32_t *>(B.carry_in + coff);
        for (size_t i = lane; i < state_bytes / 4; i += 32) st4[i] = key ? 0x80808080u : in4[i];
        __syncwarp();
    }

    const uint32_t lt_mask = (1u << lane) - 1u;
    const int32_t *my_lines = T.pc_lines + g.pc_line_first[pc];
    const int nl = g.pc_nlines[pc];

    for (int f = f0; f < f1; f++) {
        const uint32_t *rec_slice = B.rec + (size_t)f * L.rec_per_frame + g.rec_first;
        const uint32_t *loff = B.line_off + (size_t)f * L.lines_per_frame + g.line_first;
        uint16_t *dec_slice = B.dec + B.slice_base[f * L.nslices + s];
        for (int li = 0; li < nl; li++) {
            const int line = my_lines[li];
            const LineDesc ld = T.lines[g.line_first + line];
            const uint32_t *recp = rec_slice + ld.rec_off;
            uint16_t *out = dec_slice + loff[line];
            const int w = ld.w;
            for (int x0 = 0; x0 < w; x0 += 32) {
                const int n = min(32, w - x0);
                const uint32_t r = (lane < n) ? recp[x0 + lane] : 0u;
                const int d_own = (int)(int16_t)(r & 0xFFFFu);
                SymMasks m = symbol_masks(d_own);
                int e_own = 0;
                if (HIGH_E) {
                    e_own = d_own ? 31 - __clz((uint32_t)abs(d_own)) : 0;
                    if (e_own > 9) m.nd = 2 * e_own + 3;
                }
                uint32_t incl = warp_incl_scan(lane < n ? (uint32_t)m.nd : 0u, lane);
                const uint32_t pos_own = incl - (lane < n ? (uint32_t)m.nd : 0u);
                const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                for (int j = 0; j < n; j++) {
                    const uint32_t rj = __shfl_sync(0xFFFFFFFFu, r, j);
                    const uint32_t vis = __shfl_sync(0xFFFFFFFFu, m.visit, j);
                    const uint32_t bts = __shfl_sync(0xFFFFFFFFu, m.bits, j);
                    const uint32_t pj = __shfl_sync(0xFFFFFFFFu, pos_own, j);
                    uint8_t *row = st + (size_t)(rj >> 16) * 32;
                    if (HIGH_E) {
                        const int ej = __shfl_sync(0xFFFFFFFFu, e_own, j);
                        if (ej > 9) {
                            // rare large-magnitude branch (ffv1enc.c:217-228): slots 1+9 and 22+9 are visited
                            // repeatedly, so the decisions are walked one by one (warp-uniform loop).
                            const int dj = (int)(int16_t)(rj & 0xFFFFu);
                            const uint32_t a = (uint32_t)abs(dj);
                            const int ndj = 2 * ej + 3;
                            for (int k = 0; k < ndj; k++) {
                                int slot, bit;
                                if (k == 0) { slot = 0; bit = 0; }
                                else if (k <= ej) { slot = 1 + min(k - 1, 9); bit = 1; }
                                else if (k == ej + 1) { slot = 1 + 9; bit = 0; }
                                else if (k <= 2 * ej + 1) { const int i = ej - 1 - (k - ej - 2); slot = 22 + min(i, 9); bit = (a >> i) & 1; }
                                else { slot = 11 + 10; bit = dj < 0; }
                                if (lane == lane_of_slot(slot)) {
                                    const uint8_t p = row[lane];
                                    out[pj + k] = (uint16_t)(p | (bit << 8));
                                    row[lane] = s_lut[(bit << 8) + p];
                                }
                                __syncwarp();
                            }
                            continue;
                        }
                    }
                    const uint8_t p = row[lane];
                    const uint32_t bit = (bts >> lane) & 1u;
                    if ((vis >> lane) & 1u) {
                        out[pj + __popc(vis & lt_mask)] = (uint16_t)(p | (bit << 8));
                        row[lane] = s_lut[(bit << 8) + p];
                    }
                }
                out += total;
            }
        }
    }
    // ---- hand the state to the next batch when this segment runs to the end of the batch
    if (f1 == B.nframes) {
        __syncwarp();
        const size_t coff = ((size_t)s * L.npc + pc) * state_bytes;
        const uint32_t *st4 = reinterpret_cast<const uint32_t *>(st);
        uint32_t *out4 = reinterpret_cast<uint32_t *>(B.carry_out + coff);
        for (size_t i = lane; i < state_bytes / 4; i += 32) out4[i] = st4[i];
    }
}

int replay_smem_bytes(const Layout &L)
{
    return 512 + kReplayWarps * L.ctx_count * 32;
}

void launch_replay(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s)
{
    const int nchains = b.nseg * t.layout.nslices * t.layout.npc;
    const int grid = (nchains + kReplayWarps - 1) / kReplayWarps;
    const bool high = t.layout.coded_bits > 10;
    if (t.state_in_smem) {
        const int smem = replay_smem_bytes(t.layout);
        if (high) k_replay<true, true><<<grid, kReplayWarps * 32, smem, s>>>(t, b);
        else      k_replay<true, false><<<grid, kReplayWarps * 32, smem, s>>>(t, b);
    } else {
        if (high) k_replay<false, true><<<grid, kReplayWarps * 32, 512, s>>>(t, b);
        else      k_replay<false, false><<<grid, kReplayWarps * 32, 512, s>>>(t, b);
    }
}

// =================================================================================================
// k_rangecode: one interval coder per (frame, slice)
// =================================================================================================
struct Rac {
    uint32_t low, range;
    int out_byte;        // -1: none pending
    uint32_t out_count;  // pending 0xFF bytes
    uint8_t *buf;
    uint32_t pos, cap;
};

__device__ __forceinline__ void rac_emit(Rac &c, int b)
{
    if (c.pos < c.cap) c.buf[c.pos] = (uint8_t)b;
    c.pos++;
}

__device__ __forceinline__ void rac_shift(Rac &c)
{
    // one iteration of renorm_encoder's loop (rangecoder.h:52-75)
    if (c.out_byte < 0) {
        c.out_byte = c.low >> 8;
    } else if (c.low <= 0xFF00u) {
        rac_emit(c, c.out_byte);
        for (; c.out_count; c.out_count--) rac_emit(c, 0xFF);
        c.out_byte = c.low >> 8;
    } else if (c.low >= 0x10000u) {
        rac_emit(c, c.out_byte + 1);
        for (; c.out_count; c.out_count--) rac_emit(c, 0x00);
        c.out_byte = (c.low >> 8) & 0xFF;
    } else {
        c.out_count++;
    }
    c.low = (c.low & 0xFFu) << 8;
    c.range <<= 8;
}

__device__ __forceinline__ void rac_code(Rac &c, uint32_t entry)
{
    // put_rac (rangecoder.h:85-102); the state update already happened in k_replay
    const uint32_t p = entry & 0xFFu;
    const uint32_t r1 = (c.range * p) >> 8;
    const uint32_t r0 = c.range - r1;
    if (entry & 0x100u) { c.low += r0; c.range = r1; } else c.range = r0;
    if (c.range < 0x100u) rac_shift(c);       // p >= 1 and range >= 0x100 before: at most one shift
}

__global__ void __launch_bounds__(128) k_rangecode(const EncDeviceTables T, const EncBatch B)
{
    const Layout &L = T.layout;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B.nframes * L.nslices) return;
    if (B.status[0]) return;
    const int f = idx / L.nslices, s = idx - f * L.nslices;
    const SliceGeom &g = T.slices[s];
    const int key = B.frame_key[f] ? 1 : 0;

    Rac c;
    c.low = 0; c.range = 0xFF00u; c.out_byte = -1; c.out_count = 0;       // ff_init_range_encoder
    c.buf = B.scratch + (size_t)f * L.scratch_per_frame + g.scratch_off;
    c.pos = 0; c.cap = g.scratch_cap;

    const uint16_t *pre = T.prefix + (size_t)(s * 2 + key) * kMaxPrefix;
    const int npre = T.prefix_len[s * 2 + key];
    for (int i = 0; i < npre; i++) rac_code(c, pre[i]);

    const uint32_t n = B.slice_ndec[idx];
    const uint4 *src = reinterpret_cast<const uint4 *>(B.dec + B.slice_base[idx]);
    const uint32_t nvec = n >> 3;
    uint4 v = nvec ? src[0] : make_uint4(0, 0, 0, 0);
    for (uint32_t i = 0; i < nvec; i++) {
        const uint4 nx = (i + 1 < nvec || (n & 7u)) ? src[i + 1] : make_uint4(0, 0, 0, 0);
        rac_code(c, v.x & 0xFFFFu); rac_code(c, v.x >> 16);
        rac_code(c, v.y & 0xFFFFu); rac_code(c, v.y >> 16);
        rac_code(c, v.z & 0xFFFFu); rac_code(c, v.z >> 16);
        rac_code(c, v.w & 0xFFFFu); rac_code(c, v.w >> 16);
        v = nx;
    }
    if (n & 7u) {
        if (!nvec) v = src[0];
        const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
        for (uint32_t k = 0; k < (n & 7u); k++) rac_code(c, (wds[k >> 1] >> ((k & 1u) * 16)) & 0xFFFFu);
    }
    rac_code(c, 129u);                                        // put_rac(state 129, 0): ffv1enc.c:1331-1333
    // ff_rac_terminate (rangecoder.c:104-116)
    c.range = 0xFFu; c.low += 0xFFu;
    while (c.range < 0x100u) rac_shift(c);
    c.range = 0xFFu;
    while (c.range < 0x100u) rac_shift(c);

    B.slice_bytes[idx] = c.pos;
    if (c.pos > c.cap) atomicMax(&B.status[1], (unsigned long long)c.pos);
}

void launch_rangecode(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s)
{
    const int n = b.nframes * t.layout.nslices;
    k_rangecode<<<(n + 127) / 128, 128, 0, s>>>(t, b);
}

// =================================================================================================
// k_pack_layout / k_pack_slices: packet assembly + CRC
// =================================================================================================
__global__ void __launch_bounds__(1024) k_pack_layout(const EncDeviceTables T, const EncBatch B)
{
    // packet size of every frame, then packet offsets (serial prefix; nframes is small)
    const int ns = T.layout.nslices;
    const int trailer = (T.version > 2 ? 3 : 0) + (T.ec ? 5 : 0);
    for (int f = threadIdx.x; f < B.nframes; f += blockDim.x) {
        uint32_t sz = 0;
        for (int s = 0; s < ns; s++) {
            sz += B.slice_bytes[f * ns + s] + trailer;
            if (T.version <= 2 && s > 0) sz += 3;
        }
        B.pkt_size[f] = sz;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int f = 0; f < B.nframes; f++) { B.pkt_off[f] = run; run += B.pkt_size[f]; }
        B.pkt_off[B.nframes] = run;
        if (run > B.out_capacity) B.status[2] = run;
    }
}

__device__ __forceinline__ uint32_t gf_mulmod(uint32_t a, uint32_t b)
{
    // a(x)*b(x) mod P(x) over GF(2), P = x^32 + 0x04C11DB7, bit k = coefficient of x^k
    uint32_t r = 0;
#pragma unroll 8
    for (int i = 31; i >= 0; i--) {
        r = (r << 1) ^ ((r & 0x80000000u) ? 0x04C11DB7u : 0u);
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}

constexpr int kPackThreads = 128;

// One CTA per (frame, slice): copy the coder output to its final place, append the 24-bit length and the
// error-check trailer.  CRC: every thread runs a table CRC over one contiguous chunk; partial CRCs are
// merged pairwise, crc(A|B) = crc(A)*x^(8|B|) + crc(B)  (init 0, no final xor => plain polynomial remainder).
__global__ void __launch_bounds__(kPackThreads) k_pack_slices(const EncDeviceTables T, const EncBatch B)
{
    __shared__ uint32_t s_tab[256];
    __shared__ uint32_t s_part[kPackThreads];
    const Layout &L = T.layout;
    const int ns = L.nslices;
    const int f = blockIdx.x / ns, s = blockIdx.x - f * ns;
    const int tid = threadIdx.x;
    if (B.status[0] | B.status[1] | B.status[2]) return;

    for (int n = tid; n < 256; n += kPackThreads) {
        uint32_t c = (uint32_t)n << 24;
#pragma unroll
        for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : (c << 1);
        s_tab[n] = c;
    }
    const bool has_len = (T.version > 2) || s > 0;
    const int trailer = (has_len ? 3 : 0) + (T.ec ? 5 : 0);
    unsigned long long off = B.pkt_off[f];
    for (int k = 0; k < s; k++) off += B.slice_bytes[f * ns + k] + ((T.version > 2 || k > 0) ? 3 : 0) + (T.ec ? 5 : 0);
    const uint32_t nb = B.slice_bytes[f * ns + s];
    const uint8_t *src = B.scratch + (size_t)f * L.scratch_per_frame + T.slices[s].scratch_off;
    uint8_t *dst = B.out + off;
    __syncthreads();

    // message covered by the CRC: payload | len24 | 0x00
    const uint32_t mlen = nb + (has_len ? 3 : 0) + (T.ec ? 1 : 0);
    const uint32_t chunk = (mlen + kPackThreads - 1) / kPackThreads;
    // chunks are aligned to the END of the message; a short/empty first chunk is equivalent to leading zero
    // bytes, which do not change a CRC with zero initial value.
    const long long beg = (long long)mlen - (long long)(kPackThreads - tid) * chunk;
    uint32_t crc = 0;
    for (long long i = beg < 0 ? 0 : beg; i < beg + (long long)chunk; i++) {
        uint32_t byte;
        if (i < nb) byte = src[i];
        else {
            const uint32_t k = (uint32_t)i - nb;
            byte = (has_len && k < 3) ? ((nb >> (16 - 8 * k)) & 0xFFu) : 0u;
        }
        dst[i] = (uint8_t)byte;
        crc = (crc << 8) ^ s_tab[(crc >> 24) ^ byte];
    }
    if (!T.ec) return;
    // x^(8*chunk) mod P by square-and-multiply
    uint32_t mult = 1u, base = 0x100u;
    for (uint32_t e = chunk; e; e >>= 1) {
        if (e & 1u) mult = gf_mulmod(mult, base);
        base = gf_mulmod(base, base);
    }
    s_part[tid] = crc;
    __syncthreads();
    for (int stride = 1; stride < kPackThreads; stride <<= 1) {
        uint32_t v = 0;
        const bool active = (tid % (2 * stride)) == 0;
        if (active) v = gf_mulmod(s_part[tid], mult) ^ s_part[tid + stride];
        __syncthreads();
        if (active) s_part[tid] = v;
        mult = gf_mulmod(mult, mult);
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t c = s_part[0];
        dst[mlen + 0] = (uint8_t)(c >> 24); dst[mlen + 1] = (uint8_t)(c >> 16);
        dst[mlen + 2] = (uint8_t)(c >> 8);  dst[mlen + 3] = (uint8_t)c;
    }
    (void)trailer;
}

void launch_pack(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s)
{
    k_pack_layout<<<1, 1024, 0, s>>>(t, b);
    k_pack_slices<<<b.nframes * t.layout.nslices, kPackThreads, 0, s>>>(t, b);
}

cudaError_t configure_kernels(const Layout &L)
{
    cudaError_t e;
    const int psm = pixel_smem_bytes(L);
#define SET_PIXEL(K, N) do { e = cudaFuncSetAttribute(k_pixel<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, psm); if (e != cudaSuccess) return e; } while (0)
    if (psm > 48 * 1024) {
        SET_PIXEL(SRC_PLANAR8, 3); SET_PIXEL(SRC_PLANAR8, 5); SET_PIXEL(SRC_PLANAR16, 3); SET_PIXEL(SRC_PLANAR16, 5);
        SET_PIXEL(SRC_RGB32, 3); SET_PIXEL(SRC_RGB32, 5); SET_PIXEL(SRC_GBRP16, 3); SET_PIXEL(SRC_GBRP16, 5);
    }
#undef SET_PIXEL
    const int rsm = replay_smem_bytes(L);
    if (rsm <= 227 * 1024) {
        e = cudaFuncSetAttribute(k_replay<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, rsm); if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_replay<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, rsm); if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

} // namespace ffv1
