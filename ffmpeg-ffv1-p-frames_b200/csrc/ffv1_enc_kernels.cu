// ffv1_enc_kernels.cu -- hand-written sm_100a kernels of the FFV1 encode hot path.
//
//   k_pixel      per-pixel pass: sample fetch (+RCT), slice-local neighbourhood, get_context, median predictor,
//                residual, sign flip, fold  ->  (context<<16 | diff) records in coding order.
//                Reference: encode_plane / encode_rgb_frame / encode_line up to the symbol call
//                (ffv1enc.c:373-473, 271-321; ffv1.h:148-190).  HBM bound: 1-2 B read + 4 B written per sample.
//   k_replay     adaptive-state replay: put_symbol_inline's binarisation (ffv1enc.c:185-231) applied to the
//                per-context 32-byte states, one warp per (GOP segment, slice, plane context), lane = sample.
//                Emits the (probability, bit) pair of every binary decision; carries state across frames of a GOP
//                (the reference's "P-frames": state is only reset on keyframes, ffv1enc.c:1171-1172).
//   k_rangecode  put_rac / renorm_encoder / ff_rac_terminate (rangecoder.h:52-102, rangecoder.c:104-116):
//                one sequential interval coder per (frame, slice), thousands in flight.
//   k_pack_*     packet assembly (ffv1enc.c:1326-1354): slice sizes -> offsets, copy, 24-bit length, CRC-32
//                computed in parallel chunks and combined in GF(2) (libavutil/crc.c semantics).
#include "ffv1_enc_kernels.cuh"
#include <cstdio>
#include <cstdlib>

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

// {ry, by} of the 15 candidates of choose_rct_params (ffv1enc.c:1066-1084), packed ry | by << 4
__constant__ uint8_t c_rct_coef[15] = {0x00, 0x11, 0x22, 0x20, 0x02, 0x04, 0x40, 0x30, 0x03, 0x13, 0x31, 0x21, 0x12, 0x10, 0x01};

template <int SRC>
__device__ __forceinline__ void load_bgr(const uint8_t *const *pl, const int32_t *ls, const SliceGeom &g, int x, int y,
                                         int &b, int &gg, int &r, int &a)
{
    // ffv1enc.c:431-442 / 1103-1112: packed BGRA, or data[0], data[1], data[2] under the reference's names b, g, r
    a = 0;
    if (SRC == SRC_RGB32) {
        unsigned v = *reinterpret_cast<const uint32_t *>(pl[0] + (size_t)(g.y0 + y) * ls[0] + 4 * (g.x0 + x));
        b = v & 0xFF; gg = (v >> 8) & 0xFF; r = (v >> 16) & 0xFF; a = v >> 24;
    } else {
        b  = *reinterpret_cast<const uint16_t *>(pl[0] + (size_t)(g.y0 + y) * ls[0] + 2 * (g.x0 + x));
        gg = *reinterpret_cast<const uint16_t *>(pl[1] + (size_t)(g.y0 + y) * ls[1] + 2 * (g.x0 + x));
        r  = *reinterpret_cast<const uint16_t *>(pl[2] + (size_t)(g.y0 + y) * ls[2] + 2 * (g.x0 + x));
    }
}

template <int SRC>
__device__ __forceinline__ void raw_rgb(const Layout &L, const uint8_t *const *pl, const int32_t *ls,
                                        const SliceGeom &g, int x, int y, int out[4], int by, int ry)
{
    // ffv1enc.c:431-458: b,g,r[,a] from packed BGRA or from data[0],data[1],data[2] (the reference's naming; for
    // GBRP that makes plane 1 the RCT base channel), then b-=g; r-=g; g+=(b+r)>>2; b+=off; r+=off
    int b, gg, r, a;
    load_bgr<SRC>(pl, ls, g, x, y, b, gg, r, a);
    b -= gg; r -= gg;
    gg += (b * by + r * ry) >> 2;                       // version <= 3: by = ry = 1 (ffv1enc.c:447-451, 1165-1167)
    b += L.rct_offset; r += L.rct_offset;
    out[0] = (int16_t)gg; out[1] = (int16_t)b; out[2] = (int16_t)r; out[3] = (int16_t)a;
}

// choose_rct_params (ffv1enc.c:1064-1144): for every slice of an RGB frame, the coefficient pair that minimises the sum of
// |second-order difference of the luma-like channel| over the slice.  The reference walks the slice row by row keeping
// the previous row's horizontal differences; spelled per pixel that is the 2x2 stencil d(x,y) = h(x,y) - h(x,y-1) with
// h(x,y) = v(x,y) - v(x-1,y), for x >= 1 and y >= 1.  One CTA per (slice, frame); the 15 sums are 32-bit and wrap like the
// reference's int accumulators; the first minimum wins.
constexpr int kRctThreads = 256;

template <int SRC>
__global__ void __launch_bounds__(kRctThreads) k_rct_search(const EncDeviceTables T, const EncBatch B, uint8_t *rct_idx)
{
    __shared__ uint32_t s_part[kRctThreads / 32][15];
    const Layout &L = T.layout;
    const int s = blockIdx.x, f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SliceGeom &g = T.slices[s];
    const uint8_t *pl[4];
    int32_t ls[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { pl[i] = B.planes[f * 4 + i]; ls[i] = B.linesize[i]; }
    uint32_t stat[15];
#pragma unroll
    for (int i = 0; i < 15; i++) stat[i] = 0u;
    const int w1 = g.w - 1, n = w1 * (g.h - 1);
    for (int i = tid; i < n; i += kRctThreads) {
        const int y = 1 + i / w1, x = 1 + i - (y - 1) * w1;
        int b[4], gg[4], r[4], a;
        load_bgr<SRC>(pl, ls, g, x, y, b[0], gg[0], r[0], a);
        load_bgr<SRC>(pl, ls, g, x - 1, y, b[1], gg[1], r[1], a);
        load_bgr<SRC>(pl, ls, g, x, y - 1, b[2], gg[2], r[2], a);
        load_bgr<SRC>(pl, ls, g, x - 1, y - 1, b[3], gg[3], r[3], a);
        // the previous row's differences pass through the reference's int16_t sample buffer
        const int bg = (gg[0] - gg[1]) - (int)(int16_t)(gg[2] - gg[3]);
        int bb = (b[0] - b[1]) - (int)(int16_t)(b[2] - b[3]);
        int br = (r[0] - r[1]) - (int)(int16_t)(r[2] - r[3]);
        br -= bg; bb -= bg;
#pragma unroll
        for (int k = 0; k < 15; k++) {
            const int ry = c_rct_coef[k] & 15, by = c_rct_coef[k] >> 4;
            stat[k] += (uint32_t)abs(bg + ((br * ry + bb * by) >> 2));
        }
    }
#pragma unroll
    for (int k = 0; k < 15; k++) {
        uint32_t v = stat[k];
#pragma unroll
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
        if (lane == 0) s_part[warp][k] = v;
    }
    __syncthreads();
    if (tid == 0) {
        int best = 0, best_v = 0;
        for (int k = 0; k < 15; k++) {
            uint32_t v = 0;
            for (int wv = 0; wv < kRctThreads / 32; wv++) v += s_part[wv][k];
            if (k == 0 || (int)v < best_v) { best = k; best_v = (int)v; }
        }
        rct_idx[f * L.nslices + s] = (uint8_t)best;
    }
}

void launch_rct_search(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s)
{
    dim3 grid(t.layout.nslices, b.nframes);
    uint8_t *out = const_cast<uint8_t *>(b.rct_idx);
    if (t.layout.src_kind == SRC_RGB32) k_rct_search<SRC_RGB32><<<grid, kRctThreads, 0, s>>>(t, b, out);
    else                                k_rct_search<SRC_GBRP16><<<grid, kRctThreads, 0, s>>>(t, b, out);
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
    const int npl = td.nplanes, nrows = td.nrows;
    int16_t *S = s_quant + NIN * 256;
    const int plane_stride = (kTileRows + 2) * kPixelRowElems;

    for (int i = tid; i < NIN * 256; i += kPixelThreads) s_quant[i] = T.quant[i];

    const uint8_t *pl[4];
    int32_t ls[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { pl[i] = B.planes[f * 4 + i]; ls[i] = B.linesize[i]; }

    const int w = g.pw[td.plane];
    const int bits = L.coded_bits;
    int rct_by = 1, rct_ry = 1;
    if (RGB && B.rct_idx) { const int cf = c_rct_coef[B.rct_idx[f * L.nslices + td.slice]]; rct_ry = cf & 15; rct_by = cf >> 4; }
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
                    if (!zero) raw_rgb<SRC>(L, pl, ls, g, x, yy, v, rct_by, rct_ry);
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
            }
        }
    }
}

int pixel_smem_bytes(const Layout &L)
{
    const int planes = L.rgb ? L.nplanes : 1;
    return L.ctx_inputs * 256 * 2 + planes * (kTileRows + 2) * kPixelRowElems * 2;
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
// k_replay: adaptive state replay, one warp per (GOP segment, slice, plane context) chain
// =================================================================================================
// The 32 lanes take 32 consecutive samples of a line.  Samples that fall into DIFFERENT contexts touch disjoint
// 32-byte state rows, so they are replayed in parallel; samples sharing a context are ordered by their position
// (match.any groups -> rank -> one round per rank).  Inside one symbol (ffv1enc.c:202-229) all decisions hit
// distinct state slots as long as e <= 9, so their probability loads are issued together.
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t n = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

constexpr int kStateStride = 36;      // shared-memory bytes per context row: 9 words, so consecutive contexts start in different banks
constexpr int kOnePowBytes = 33 * 256;

template <bool SMEM> struct StateIO;
template <> struct StateIO<true> {
    static __device__ __forceinline__ uint32_t ld(const uint8_t *p) { return *p; }
    static __device__ __forceinline__ void st(uint8_t *p, uint32_t v) { *p = (uint8_t)v; }
};
template <> struct StateIO<false> {   // state rows in global memory, cached by the SM's L1: only this warp touches them
    static __device__ __forceinline__ uint32_t ld(const uint8_t *p) { return *p; }
    static __device__ __forceinline__ void st(uint8_t *p, uint32_t v) { *p = (uint8_t)v; }
};

// large magnitudes (e > 9, ffv1enc.c:217-228): slots 1+9 and 22+9 repeat, so the decisions are walked one by one
template <bool SMEM>
__device__ __noinline__ void replay_symbol_slow(uint8_t *row, const uint8_t *lut, uint16_t *o, int d, bool wr)
{
    typedef StateIO<SMEM> IO;
    const uint32_t a = (uint32_t)abs(d);
    const int e = 31 - __clz(a);
    const uint32_t sign = d < 0 ? 0x100u : 0u;
    int k = 1;
    for (int i = 0; i < e; i++, k++) {
        uint8_t *q = row + 1 + min(i, 9);
        const uint32_t p = IO::ld(q);
        if (wr) o[k] = (uint16_t)(p | 0x100u);
        IO::st(q, lut[256 + p]);
    }
    {
        uint8_t *q = row + 1 + 9;
        const uint32_t p = IO::ld(q);
        if (wr) o[k] = (uint16_t)p;
        IO::st(q, lut[p]);
        k++;
    }
    for (int i = e - 1; i >= 0; i--, k++) {
        uint8_t *q = row + 22 + min(i, 9);
        const uint32_t p = IO::ld(q), bit = ((a >> i) & 1u) << 8;
        if (wr) o[k] = (uint16_t)(p | bit);
        IO::st(q, lut[bit + p]);
    }
    uint8_t *q = row + 11 + 10;
    const uint32_t p = IO::ld(q);
    if (wr) o[k] = (uint16_t)(p | sign);
    IO::st(q, lut[sign + p]);
}

// One round: every lane with `mine` set replays put_symbol_inline(c, state, d, is_signed=1) (ffv1enc.c:185-231) on
// its own context row (rows of the participating lanes are distinct).  Emits p | bit<<8 per decision and updates
// the row.  Loops are bounded by the largest exponent in the round (warp-uniform), loads are issued before the
// dependent table lookups and stores.
template <bool SMEM, int MAXE, bool HIGH_E>
__device__ __forceinline__ void replay_round(bool mine, uint8_t *row, const uint8_t *lut, uint16_t *o, int d, bool wr)
{
    typedef StateIO<SMEM> IO;
    const uint32_t a = (uint32_t)abs(d);
    const int e = 31 - __clz(a | 1u);
    const bool slow = HIGH_E && mine && e > 9;
    const bool nz = mine && d != 0 && !slow;
    const int el = nz ? e : -1;
    const int emax = __reduce_max_sync(0xFFFFFFFFu, el);
    const uint32_t sign = d < 0 ? 0x100u : 0u;
    const bool wm = wr && mine, wz = wr && nz;

    uint32_t p0 = 0, ps = 0, pu[MAXE + 1], pm[MAXE];
    if (mine) p0 = IO::ld(row);
    if (emax >= 0) {
#pragma unroll
        for (int i = 0; i <= MAXE; i++) {
            if (i > emax) break;
            pu[i] = (i <= el) ? IO::ld(row + 1 + i) : 0u;
        }
#pragma unroll
        for (int i = 0; i < MAXE; i++) {
            if (i >= emax) break;
            pm[i] = (i < el) ? IO::ld(row + 22 + i) : 0u;
        }
        if (nz) ps = IO::ld(row + 11 + e);
    }
    {
        const uint32_t bit = d == 0 ? 0x100u : 0u;                      // "is zero" flag
        if (wm) o[0] = (uint16_t)(p0 | bit);
        if (mine) IO::st(row, lut[bit + p0]);
    }
    if (emax >= 0) {
#pragma unroll
        for (int i = 0; i <= MAXE; i++) {
            if (i > emax) break;
            const uint32_t bit = i < el ? 0x100u : 0u;                  // unary exponent, then its terminator
            if (i <= el) {
                if (wz) o[1 + i] = (uint16_t)(pu[i] | bit);
                IO::st(row + 1 + i, lut[bit + pu[i]]);
            }
        }
#pragma unroll
        for (int i = 0; i < MAXE; i++) {
            if (i >= emax) break;
            const uint32_t bit = ((a >> i) & 1u) << 8;                  // mantissa, most significant bit first
            if (i < el) {
                if (wz) o[2 * e + 1 - i] = (uint16_t)(pm[i] | bit);
                IO::st(row + 22 + i, lut[bit + pm[i]]);
            }
        }
        if (nz) {
            if (wz) o[2 * e + 2] = (uint16_t)(ps | sign);
            IO::st(row + 11 + e, lut[sign + ps]);
        }
    }
    if (HIGH_E && __any_sync(0xFFFFFFFFu, slow)) {
        if (slow) replay_symbol_slow<SMEM>(row, lut, o, d, wr);
    }
}

template <bool SMEM, int MAXE, bool HIGH_E>
__global__ void __launch_bounds__(kReplayWarps * 32)
k_replay(const EncDeviceTables T, const EncBatch B)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Layout &L = T.layout;
    uint8_t *s_lut = smem_raw;                               // [0..255] zero_state, [256..511] one_state
    uint8_t *s_pow = smem_raw + 512;                         // [k][p]: one_state applied k times, k = 0..32
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 512 / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(s_lut)[i] = reinterpret_cast<const uint32_t *>(T.trans_lut)[i];
    for (int i = threadIdx.x; i < kOnePowBytes / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(s_pow)[i] = reinterpret_cast<const uint32_t *>(T.one_pow)[i];
    __syncthreads();

    const int nchains = B.nseg * L.nslices * L.npc;
    const int chain = blockIdx.x * kReplayWarps + warp;
    if (chain >= nchains) return;
    const int pc = chain % L.npc;
    const int s = (chain / L.npc) % L.nslices;
    const int seg = chain / (L.npc * L.nslices);
    const SliceGeom &g = T.slices[s];
    const int nl = g.pc_nlines[pc];
    if (nl == 0) return;
    const int f0 = B.seg_first[seg], f1 = B.seg_first[seg + 1];
    const int nctx = L.ctx_count;
    constexpr int RS = SMEM ? kStateStride : 32;
    const int state_smem = (nctx * kStateStride + 15) & ~15;

    uint8_t *st;
    if (SMEM) st = smem_raw + 512 + kOnePowBytes + (size_t)warp * state_smem;
    else      st = B.state_seg + ((size_t)(seg * L.nslices + s) * L.npc + pc) * ((size_t)nctx * 32);

    // ---- initial state: 128 on keyframes (ffv1.c:177-202), else carried over from the previous batch
    const size_t coff = ((size_t)s * L.npc + pc) * ((size_t)nctx * 32);
    {
        const bool key = B.frame_key[f0];
        const uint32_t *in4 = reinterpret_cast<const uint32_t *>(B.carry_in + coff);
        uint32_t *st4 = reinterpret_cast<uint32_t *>(st);
        for (int i = lane; i < nctx * 8; i += 32)
            st4[(i >> 3) * (RS / 4) + (i & 7)] = key ? (T.init_state ? reinterpret_cast<const uint32_t *>(T.init_state)[i] : 0x80808080u) : in4[i];
        __syncwarp();
    }

    const uint32_t lt_mask = (1u << lane) - 1u;
    const int32_t *my_lines = T.pc_lines + g.pc_line_first[pc];
    const LineDesc *slice_lines = T.lines + g.line_first;
    const uint32_t cap = g.dec_cap[pc];
    uint32_t need = 0;
    unsigned long long ndec = 0;
    bool overflow = false;

    for (int f = f0; f < f1; f++) {
        const uint32_t *rec_slice = B.rec + (size_t)f * L.rec_per_frame + g.rec_first;
        uint16_t *out = B.dec + (size_t)f * L.dec_per_frame + g.dec_off[pc];
        uint32_t *run_cnt = B.run_cnt + (size_t)f * L.runs_per_frame + g.run_first;
        uint32_t pos = 0, run_start = 0;
        LineDesc ld = slice_lines[my_lines[0]];
        uint32_t cur_run = ld.run;
        uint32_t r_next = lane < ld.w ? rec_slice[ld.rec_off + lane] : 0u;
        for (int li = 0; li < nl; li++) {
            LineDesc ld_next = ld;
            if (li + 1 < nl) ld_next = slice_lines[my_lines[li + 1]];
            if (ld.run != cur_run) {                                     // a run of this plane context ended
                if (lane == 0) run_cnt[cur_run] = pos - run_start;
                pos = (pos + 7u) & ~7u;
                run_start = pos; cur_run = ld.run;
            }
            const uint32_t *recp = rec_slice + ld.rec_off;
            const int w = ld.w;
            for (int x0 = 0; x0 < w; x0 += 32) {
                const int n = min(32, w - x0);
                const uint32_t r = r_next;
                if (x0 + 32 < w)     r_next = (x0 + 32 + lane < w) ? recp[x0 + 32 + lane] : 0u;
                else if (li + 1 < nl) r_next = lane < ld_next.w ? rec_slice[ld_next.rec_off + lane] : 0u;
                const bool act = lane < n;
                const int ctx = (int)(r >> 16);
                const int d = (int)(int16_t)(r & 0xFFFFu);
                uint32_t nd = 0;
                if (act) nd = d ? (uint32_t)(2 * (31 - __clz((uint32_t)abs(d))) + 3) : 1u;
                const uint32_t incl = warp_incl_scan(nd, lane);
                const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                const bool wr = pos + total <= cap;                      // warp-uniform
                overflow |= !wr;
                uint16_t *o = out + pos + incl - nd;
                uint8_t *row = st + (size_t)ctx * RS;
                const int ctx0 = __shfl_sync(0xFFFFFFFFu, ctx, 0);
                if (__all_sync(0xFFFFFFFFu, !act || (d == 0 && ctx == ctx0))) {
                    // a run of zero residuals in one context: lane j sees the state after j "is zero" decisions
                    const uint32_t p0 = StateIO<SMEM>::ld(st + (size_t)ctx0 * RS);
                    if (act && wr) o[0] = (uint16_t)(s_pow[lane * 256 + p0] | 0x100u);
                    __syncwarp();
                    if (lane == 0) StateIO<SMEM>::st(st + (size_t)ctx0 * RS, s_pow[n * 256 + p0]);
                    __syncwarp();
                } else {
                    const uint32_t grp = __match_any_sync(0xFFFFFFFFu, act ? (uint32_t)ctx : 0x10000u + lane);
                    const int rank = __popc(grp & lt_mask);
                    const int last = __reduce_max_sync(0xFFFFFFFFu, act ? rank : 0);
                    for (int round = 0; round <= last; round++) {
                        replay_round<SMEM, MAXE, HIGH_E>(act && rank == round, row, s_lut, o, d, wr);
                        __syncwarp();
                    }
                }
                pos += total;
                ndec += total;
            }
            ld = ld_next;
        }
        if (lane == 0) run_cnt[cur_run] = pos - run_start;
        need = max(need, pos + 8u);
    }
    if (lane == 0) {
        atomicAdd(&B.status[3], ndec);
        if (overflow) {
            const unsigned long long ns = g.pc_samples[pc];
            atomicMax(&B.status[0], ((unsigned long long)need * 256ull + ns - 1) / ns + 1ull);
        }
    }
    // ---- hand the state to the next batch when this segment runs to the end of the batch
    if (f1 == B.nframes) {
        __syncwarp();
        const uint32_t *st4 = reinterpret_cast<const uint32_t *>(st);
        uint32_t *out4 = reinterpret_cast<uint32_t *>(B.carry_out + coff);
        for (int i = lane; i < nctx * 8; i += 32) out4[i] = st4[(i >> 3) * (RS / 4) + (i & 7)];
    }
}

int replay_smem_bytes(const Layout &L)
{
    return 512 + kOnePowBytes + kReplayWarps * ((L.ctx_count * kStateStride + 15) & ~15);
}

template <bool SMEM>
static void launch_replay_t(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s, int grid, int smem)
{
    const int bits = t.layout.coded_bits;
    if (bits <= 8)       k_replay<SMEM, 7, false><<<grid, kReplayWarps * 32, smem, s>>>(t, b);
    else if (bits <= 10) k_replay<SMEM, 9, false><<<grid, kReplayWarps * 32, smem, s>>>(t, b);
    else                 k_replay<SMEM, 9, true><<<grid, kReplayWarps * 32, smem, s>>>(t, b);
}

void launch_replay(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s)
{
    const int nchains = b.nseg * t.layout.nslices * t.layout.npc;
    const int grid = (nchains + kReplayWarps - 1) / kReplayWarps;
    if (t.state_in_smem) launch_replay_t<true>(t, b, s, grid, replay_smem_bytes(t.layout));
    else                 launch_replay_t<false>(t, b, s, grid, 512 + kOnePowBytes);
}

// =================================================================================================
// k_rangecode: one interval coder per (frame, slice); put_rac / renorm_encoder / ff_rac_terminate
// (rangecoder.h:52-102, rangecoder.c:104-116).  The adaptive part already happened in the state replay, so a step is
// range1 = range*p >> 8, the interval update and the renormalisation.
// =================================================================================================
// Every lane of a warp runs its own coder and the warp NEVER diverges inside the decision loop:
//  * Carry-free output.  renorm_encoder delays bytes (outstanding_byte / outstanding_count) because a later carry may
//    still increment them.  Here a renormalisation emits the 9-bit value low >> 8 (byte + carry bit 16 of low) as a
//    16-bit entry and moves on; the carries are added up afterwards, in parallel, by the packet-assembly kernel
//    (byte k = (v[k] + carry out of v[k+1..]) & 0xFF, a carry-lookahead over generate = v >> 8 and propagate =
//    (v & 0xFF) == 0xFF; see pack_resolved_be).  The bytes are the reference's: its delayed writes perform exactly this
//    addition, and the entry of the last renormalisation (what would stay in outstanding_byte for ever) is dropped
//    like ff_rac_terminate drops it.  No "0xFF byte waits for its carry" path is left: a decision is 12 straight-line
//    instructions, split between the FMA pipe (mul.hi, subtract, add, shifts as mad) and the ALU pipe (select,
//    compares, permutes) -- each pipe takes a warp instruction every other cycle, so the mix is what sets the speed.
//  * The decision stream of a lane is produced by a small generator (prefix | runs of its slice | closing decision) that
//    runs a chunk of 14 16-byte vectors ahead of the coder (bulk copies, see k_rangecode), across run boundaries.
//  * A vector holds 8 decisions; entries behind the end of a run are replaced by 0x0000, which is an exact no-op for
//    the coder (p = 0, bit = 0: range1 = 0, nothing moves, no renormalisation; real states are 1..255).
//  * Output entries go through a per-lane ring in shared memory laid out [row][lane] (a lane always hits its own bank:
//    no conflicts whatever the lanes' positions) and leave as 16-byte stores of 8 entries.
constexpr int kRingRows = 32;             // entries per lane
constexpr uint32_t kRowUnit = 1u << 27;   // ring positions are kept as row << 27: they wrap around the 32 rows by themselves

struct Rac {
    uint32_t low, range;
    uint32_t posx;       // (entries produced mod 32) << 27
    uint32_t flx;        // (entries moved from the ring to the scratch region mod 32) << 27; whole groups of 8
    uint32_t nfl;        // entries moved to the scratch region
    uint32_t ring;       // shared-memory address of the lane's column of the ring
    uint32_t cap;        // entries the slice's scratch region holds
    uint16_t *out;       // the slice's scratch region (256-byte aligned), 16-bit entries
};

// put_rac (rangecoder.h:85-102) for one decision (p24 = probability state << 24, one = coded bit, any non-zero value)
// and its renormalisation.  p >= 1 and range >= 0x100 before: at most one shift.  What is stored is low >> 8 (9
// significant bits) from a register of its own: `low` is rewritten right behind the store.  rowmul = bytes between two rows of the ring * 32: the address of
// row (posx >> 27) is ring + ((posx * rowmul) >> 32), one mad.hi on the FMA pipe.
template <uint32_t ROWMUL>
__device__ __forceinline__ void rac_code(Rac &c, uint32_t p24, uint32_t one, uint32_t rowmul_rt)
{
    const uint32_t rowmul = ROWMUL ? ROWMUL : rowmul_rt;
    asm volatile("{\n\t.reg .pred one, sh;\n\t.reg .u32 r1, r0, a, b;\n\t"
        "mul.hi.u32 r1, %1, %3;\n\t"                        // (range * p) >> 8
        "mad.lo.u32 r0, r1, 0xFFFFFFFF, %1;\n\t"            // range - range1
        "setp.ne.u32 one, %4, 0;\n\t"
        "selp.u32 %1, r1, r0, one;\n\t"
        "@one mad.lo.u32 %0, r0, 1, %0;\n\t"                // low += range - range1
        "setp.lt.u32 sh, %1, 0x100;\n\t"
        "mad.hi.u32 a, %2, %6, %5;\n\t"                     // ring + row * row pitch
        "shr.u32 b, %0, 8;\n\t"
        "@sh st.shared.u32 [a], b;\n\t"
        "@sh add.u32 %2, %2, 0x8000000;\n\t"                // next row (mod 32)
        "@sh mad.lo.u32 %1, %1, 256, 0;\n\t"                // range <<= 8
        "@sh prmt.b32 %0, %0, 0, 0x4404;\n\t}"              // low = (low & 0xFF) << 8
        : "+r"(c.low), "+r"(c.range), "+r"(c.posx)
        : "r"(p24), "r"(one), "r"(c.ring), "r"(rowmul) : "memory");
}

template <uint32_t ROWMUL>
__device__ __forceinline__ void rac_code_word(Rac &c, uint32_t w, uint32_t rowmul_rt)
{
    rac_code<ROWMUL>(c, w << 24, w & 0x100u, rowmul_rt);
    rac_code<ROWMUL>(c, __byte_perm(w, 0u, 0x2444), w & 0x1000000u, rowmul_rt);
}

// 8 finished entries of the lane: ring -> scratch region, predicated (the lanes of a warp fill their rings at different
// speeds); full = the lane has 8 or more entries waiting
template <uint32_t ROWMUL>
__device__ __forceinline__ void rac_flush8(Rac &c, uint32_t full, uint32_t rowmul_rt)
{
    const uint32_t rowmul = ROWMUL ? ROWMUL : rowmul_rt;
    const uint32_t pitch = rowmul >> 5;
    const uint32_t at = c.ring + __umulhi(c.flx, rowmul);   // 8 entries never wrap: flx counts whole groups of 8, the ring holds 32
    uint32_t e[8];
#pragma unroll
    for (int k = 0; k < 8; k++) e[k] = 0u;
    if (full) {
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e[k]) : "r"(at + (uint32_t)k * pitch) : "memory");
    }
    if (full && c.nfl + 8u <= c.cap) {
        uint4 v;
        v.x = __byte_perm(e[0], e[1], 0x5410); v.y = __byte_perm(e[2], e[3], 0x5410);     // 16 bits of every entry
        v.z = __byte_perm(e[4], e[5], 0x5410); v.w = __byte_perm(e[6], e[7], 0x5410);
        *reinterpret_cast<uint4 *>(c.out + c.nfl) = v;
    }
    if (full) { c.flx += 8u * kRowUnit; c.nfl += 8u; }
}

// decision source of one coder: -1 = the slice's prefix (keyframe bit, slice header), 0..nruns-1 = sample runs,
// nruns = the closing put_rac(state 129, 0) of ffv1enc.c:1331-1333
struct RacGen {
    const uint4 *ptr;
    uint32_t rem;          // vectors left in the current source
    uint32_t last_valid;   // decisions in its last vector (1..8)
    int r;                 // current source
    uint32_t cur0, cur1, cur2;
};

constexpr int kRangeThreads = 32;
// One 224-byte chunk in flight per lane (it has the ~1800 decisions' time of the chunk being coded to arrive): the per-lane
// issue of the bulk copies is a fifth of the kernel's instructions at 128-byte chunks, and 14 vectors is what still lets
// eleven warps share an SM (19.5 KB each).  8 vectors x 3 slots -> 14 x 2: 34.4 -> 30.7 ms per 2048 frames (44.7 -> 34.7 on
// the boxes of the pool where the kernel runs slow).  The descriptor keeps the vector count in 4 bits: at most 15.
constexpr int kRangeChunk = 14;           // 16-byte decision vectors per chunk (224 bytes per lane)
constexpr int kRangeDepth = 2;            // chunk slots per lane: kRangeDepth-1 chunks in flight
static_assert(kRangeChunk <= 15 && (kRangeChunk & 1) == 0, "4-bit vector count; chunk + 16 bytes must be an odd number of 16-byte units");
constexpr int kRangeLanePitch = kRangeChunk * 16 + 16;    // bytes between the chunks of neighbouring lanes inside a slot: an odd
                                                          // number of 16-byte units, so the 128-bit reads of a warp spread over all banks

__device__ __forceinline__ void rc_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void rc_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rc_mbar_wait(uint32_t bar, uint32_t phase)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(bar), "r"(phase) : "memory");
}

// The decision streams come in through the TMA engine, one bulk copy (cp.async.bulk, SASS UBLKCP) per lane and chunk:
// the 32 coders of a warp read 32 different streams that lie tens of megabytes apart, i.e. every access of the warp
// touches 32 different pages.  Issued as ordinary (or cp.async) loads, these accesses sit in the SM's in-order
// load/store path while their addresses are translated -- with batches beyond the reach of the TLBs (>= 1024 frames)
// every coder step then waited behind them (the kernel took twice as long per decision as with 512-frame batches).  A
// bulk copy is a descriptor handed to the copy engine: translation and transfer happen off the warp's path, one request
// per chunk of 224 bytes instead of fourteen loads.
template <int LANES>        // coders per warp known at compile time (32: the common case, immediate strides) or 0 = `lanes_rt`
__global__ void __launch_bounds__(kRangeThreads) k_rangecode(const EncDeviceTables T, const EncBatch B, const int lanes_rt)
{
    const int lanes = LANES ? LANES : lanes_rt;
    const Layout &L = T.layout;
    const int idx = blockIdx.x * lanes + threadIdx.x;
    const bool live = (int)threadIdx.x < lanes && idx < B.nframes * L.nslices;
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, live);
    if (!live) return;
    if (B.status[0]) return;
    // a warp holds the same slice of consecutive frames: their streams have similar lengths
    const int s = idx / B.nframes, f = idx - s * B.nframes;
    const SliceGeom &g = T.slices[s];
    const int key = B.frame_key[f] ? 1 : 0;
    const int nruns = g.nruns;

    // sized by the lanes in use, so that half-filled warps do not halve the warps an SM holds:
    // [ring: 32 rows x lanes x 4][chunk slots: kRangeDepth x lanes x kRangeLanePitch][mbarriers: kRangeDepth x 8]
    extern __shared__ __align__(16) uint8_t s_range_dyn[];
    constexpr uint32_t kRowMul = (uint32_t)LANES * 4u * 32u;
    const uint32_t stride = (uint32_t)lanes * 4u;                                        // bytes between the rows of the ring
    const uint32_t rowmul = stride * 32u;
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(s_range_dyn);
    const uint32_t slotstride = (uint32_t)lanes * kRangeLanePitch;
    const uint32_t vec_base = smem0 + (uint32_t)kRingRows * stride + threadIdx.x * kRangeLanePitch;
    const uint32_t bar0 = smem0 + (uint32_t)kRingRows * stride + (uint32_t)kRangeDepth * slotstride;
    Rac c;
    c.low = 0; c.range = 0xFF00u;                                                        // ff_init_range_encoder
    c.out = reinterpret_cast<uint16_t *>(B.scratch) + (size_t)f * L.scratch_per_frame + g.scratch_off;
    c.posx = 0; c.flx = 0; c.nfl = 0; c.cap = g.scratch_cap;
    c.ring = smem0 + threadIdx.x * 4u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kRangeDepth; i++) rc_mbar_init(bar0 + 8u * i, (uint32_t)__popc(mask));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp(mask);

    const uint16_t *dec_frame = B.dec + (size_t)f * L.dec_per_frame;
    const uint32_t *run_cnt = B.run_cnt + (size_t)f * L.runs_per_frame + g.run_first;
    const uint8_t *run_pc = T.run_pc + g.run_first;

    RacGen gen;
    gen.ptr = nullptr; gen.rem = 0; gen.last_valid = 0; gen.r = -2; gen.cur0 = gen.cur1 = gen.cur2 = 0u;

    // Next chunk of the lane's stream: up to kRangeChunk consecutive 16-byte vectors of the current source, brought into
    // chunk slot `slot` by one bulk copy that signals the slot's mbarrier; every lane of the warp arrives on that barrier
    // with the bytes it expects (none once its stream has ended).  Returns nvec | decisions_in_last_vector << 4 (0 = ended).
    auto fetch = [&](uint32_t slot) -> uint32_t {
        uint32_t desc = 0u;
        if (gen.rem == 0u) {
            while (gen.r < nruns) {                           // next non-empty source
                const int r = ++gen.r;
                const uint16_t *src;
                uint32_t n;
                if (r < 0) {
                    const int var = B.rct_idx ? B.rct_idx[f * L.nslices + s] : 0;
                    src = T.prefix + ((size_t)(s * 2 + key) * T.nvar + var) * kMaxPrefix;
                    n = (uint32_t)T.prefix_len[(s * 2 + key) * T.nvar + var];
                } else if (r == nruns) {
                    src = T.prefix + (size_t)L.nslices * 2 * T.nvar * kMaxPrefix;      // one extra row holding the entry "129"
                    n = 1u;
                } else {
                    const int pc = run_pc[r];
                    n = run_cnt[r];
                    const uint32_t at = pc == 0 ? gen.cur0 : (pc == 1 ? gen.cur1 : gen.cur2);
                    src = dec_frame + g.dec_off[pc] + at;
                    const uint32_t nx = (at + n + 7u) & ~7u;
                    if (pc == 0) gen.cur0 = nx; else if (pc == 1) gen.cur1 = nx; else gen.cur2 = nx;
                }
                if (n) {
                    gen.ptr = reinterpret_cast<const uint4 *>(src);
                    gen.rem = (n + 7u) >> 3;
                    gen.last_valid = n - 8u * (gen.rem - 1u);
                    break;
                }
            }
        }
        const uint32_t bar = bar0 + 8u * slot;
        const uint32_t nv = min(gen.rem, (uint32_t)kRangeChunk);
        rc_mbar_expect_tx(bar, nv * 16u);
        if (nv) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(vec_base + slot * slotstride), "l"(gen.ptr), "r"(nv * 16u), "r"(bar) : "memory");
            gen.ptr += nv;
            gen.rem -= nv;
            desc = nv | (gen.rem ? 8u : gen.last_valid) << 4;
        }
        return desc;
    };

    // kRangeDepth - 1 chunks in flight per lane; dq holds the descriptors, 8 bits each, oldest in the low bits
    uint32_t dq = 0u;
#pragma unroll
    for (int i = 0; i < kRangeDepth - 1; i++) dq |= fetch((uint32_t)i) << (8 * i);
    uint32_t slot = 0u, phases = 0u;                          // bit i of phases: parity the next wait on slot i looks for
#pragma unroll 1
    for (;;) {
        const uint32_t nv = dq & 0xFu, last_valid = (dq >> 4) & 0xFu;
        if (!__any_sync(mask, nv != 0u)) break;               // the warp leaves together: every lane arrives on every barrier
        // the slot refilled now is the one the whole warp has finished reading in the previous round
        uint32_t nslot = slot + (uint32_t)(kRangeDepth - 1);
        if (nslot >= (uint32_t)kRangeDepth) nslot -= (uint32_t)kRangeDepth;
        dq = (dq >> 8) | fetch(nslot) << (8 * (kRangeDepth - 2));
        rc_mbar_wait(bar0 + 8u * slot, (phases >> slot) & 1u);
        phases ^= 1u << slot;
        uint32_t at = vec_base + slot * slotstride;
        slot = slot + 1u == (uint32_t)kRangeDepth ? 0u : slot + 1u;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (nv) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(at) : "memory");
#pragma unroll 1
        for (uint32_t j = nv; j; j--) {
            uint4 cur = v;
            at += 16u;
            // the next vector of the chunk is read while this one is coded (the chunk has landed as a whole)
            if (j > 1u) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(at) : "memory");
            if (j == 1u && last_valid < 8u) {                     // end of a run: what follows in the vector is not ours
                const uint32_t k = last_valid;
                if (k <= 6u) cur.w = 0u; else if (k == 7u) cur.w &= 0xFFFFu;
                if (k <= 4u) cur.z = 0u; else if (k == 5u) cur.z &= 0xFFFFu;
                if (k <= 2u) cur.y = 0u; else if (k == 3u) cur.y &= 0xFFFFu;
                if (k == 1u) cur.x &= 0xFFFFu;
            }
            rac_code_word<kRowMul>(c, cur.x, rowmul);
            rac_code_word<kRowMul>(c, cur.y, rowmul);
            rac_code_word<kRowMul>(c, cur.z, rowmul);
            rac_code_word<kRowMul>(c, cur.w, rowmul);
            // a vector adds at most 8 entries: a lane with 16 or more waiting makes room now (rare: the scheduled flush
            // below keeps a lane under 8 + what 8 vectors produce, ~14 on camera content; bursts of improbable bits do more)
            if (__any_sync(__activemask(), c.posx - c.flx >= 16u * kRowUnit)) {
                rac_flush8<kRowMul>(c, c.posx - c.flx >= 16u * kRowUnit, rowmul);
                rac_flush8<kRowMul>(c, c.posx - c.flx >= 16u * kRowUnit, rowmul);
            }
        }
        __syncwarp(mask);
        // scheduled flush, once per chunk: every lane moves its whole groups of 8 entries
        while (__any_sync(mask, c.posx - c.flx >= 8u * kRowUnit))
            rac_flush8<kRowMul>(c, c.posx - c.flx >= 8u * kRowUnit, rowmul);
    }
    // ff_rac_terminate (rangecoder.c:104-116): range = 0xFF, low += 0xFF, renormalise; range = 0xFF, renormalise.  A
    // decision with p = 0, bit = 0 changes nothing, so forcing the range below 0x100 in front of one is exactly a
    // forced renormalisation.
    c.low += 0xFFu;
#pragma unroll 1
    for (int t = 0; t < 2; t++) {
        c.range = 0xFFu;
        rac_code<kRowMul>(c, 0u, 0u, rowmul);
    }
    // the entry of the last renormalisation is what the reference keeps in outstanding_byte and never writes
    const uint32_t nent = c.nfl + ((c.posx - c.flx) >> 27);
    // the rest of the ring (the last group may run past the last entry: the region is padded and nobody reads that)
    while (c.nfl < nent) rac_flush8<kRowMul>(c, 1u, rowmul);

    B.slice_bytes[f * L.nslices + s] = nent - 1u;
    if (((nent + 7u) & ~7u) > c.cap) atomicMax(&B.status[1], (unsigned long long)nent + 8ull);
}

static int rangecode_smem(int lanes) { return lanes * (kRingRows * 4 + kRangeDepth * kRangeLanePitch) + kRangeDepth * 8; }

void launch_rangecode(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s)
{
    const int n = b.nframes * t.layout.nslices;
    // FFV1B200_RANGE_LANES forces the coders per warp (read at every launch: the parity tests walk through all variants)
    int forced = 0;
    if (const char *v = getenv("FFV1B200_RANGE_LANES")) { forced = atoi(v); if (forced < 1 || forced > 32 || (forced & (forced - 1))) forced = 0; }
    // A coder is one dependent chain (~25 cycles per decision) and a warp costs the same issue slots whatever the number
    // of its lanes that carry a coder: full warps, unless that would leave SM sub-partitions without any warp at all
    int lanes = forced;
    if (!lanes) {
        lanes = 32;
        while (lanes > 4 && n / lanes < 148 * 4) lanes >>= 1;
    }
    static bool attr = false;
    if (!attr) {                                    // ~10 resident warps of 20 KB need the large shared-memory carveout
        cudaFuncSetAttribute(k_rangecode<32>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_rangecode<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr = true;
    }
    const int smem = rangecode_smem(lanes);
    if (getenv("FFV1B200_DEBUG")) fprintf(stderr, "k_rangecode: %d coders, %d lanes\n", n, lanes);
    if (lanes == 32) k_rangecode<32><<<(n + lanes - 1) / lanes, kRangeThreads, smem, s>>>(t, b, lanes);
    else k_rangecode<0><<<(n + lanes - 1) / lanes, kRangeThreads, smem, s>>>(t, b, lanes);
}

// =================================================================================================
// k_golomb: Golomb-Rice / run-mode coder (ffv1enc.c:240-269, 327-367; golomb.h:508-563; put_bits.h), one warp per
// (GOP segment, slice) chain: the VlcState of a context carries over from frame to frame inside a GOP.
// First version: lane 0 walks the records sequentially (the other lanes only move state and prefix bytes).
// =================================================================================================
__constant__ uint8_t c_enc_log2_run[41] = {
    0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7,
    8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24,
};

struct BitW {                 // MSB-first bit writer (put_bits.h:152-195)
    uint8_t *buf;
    uint32_t pos, cap;        // bytes written / capacity
    unsigned long long acc;
    int nbits;
};

__device__ __forceinline__ void bw_put(BitW &w, int n, uint32_t v)
{
    w.acc = (w.acc << n) | v;
    w.nbits += n;
    while (w.nbits >= 8) {
        w.nbits -= 8;
        if (w.pos < w.cap) w.buf[w.pos] = (uint8_t)(w.acc >> w.nbits);
        w.pos++;
    }
}

struct __align__(8) VlcEnc { int16_t drift; uint16_t error_sum; int8_t bias; uint8_t count; uint16_t pad; };

__device__ __forceinline__ void vlc_put(BitW &w, VlcEnc *sp, int v, int bits)
{
    // put_vlc_symbol (ffv1enc.c:240-269) + set_sr_golomb (golomb.h:554-563) + update_vlc_state (ffv1.h:192-224)
    VlcEnc s = *sp;
    v = ((v - s.bias) << (32 - bits)) >> (32 - bits);                     // fold
    int k = 0;
    for (int i = s.count; i < s.error_sum; i += i) k++;
    const int code = v ^ ((2 * s.drift + s.count) >> 31);
    int m = -2 * code - 1;
    m ^= m >> 31;
    const int e = m >> k;
    if (e < 12) bw_put(w, e + k + 1, (1u << k) + ((uint32_t)m & ((1u << k) - 1u)));
    else        bw_put(w, 12 + bits, (uint32_t)(m - 11));
    int drift = s.drift, count = s.count, esum = s.error_sum, bias = s.bias;
    esum += abs(v);
    drift += v;
    if (count == 128) { count >>= 1; drift >>= 1; esum >>= 1; }
    count++;
    if (drift <= -count) {
        if (bias > -128) bias--;
        drift += count;
        if (drift <= -count) drift = -count + 1;
    } else if (drift > 0) {
        if (bias < 127) bias++;
        drift -= count;
        if (drift > 0) drift = 0;
    }
    s.drift = (int16_t)drift; s.count = (uint8_t)count; s.error_sum = (uint16_t)esum; s.bias = (int8_t)bias;
    *sp = s;
}

__global__ void __launch_bounds__(32) k_golomb(const EncDeviceTables T, const EncBatch B, const int sm_state)
{
    const Layout &L = T.layout;
    const int lane = threadIdx.x;
    const int chain = blockIdx.x;
    const int seg = chain / L.nslices, s = chain - seg * L.nslices;
    const SliceGeom &g = T.slices[s];
    const int f0 = B.seg_first[seg], f1 = B.seg_first[seg + 1];
    extern __shared__ __align__(8) unsigned char s_vlc[];                 // the chain's VlcStates when they fit (small context model)
    const size_t pc_bytes = (size_t)L.ctx_count * 32;                     // footprint of a plane context in the carry / state areas
    uint8_t *st = sm_state ? s_vlc : B.state_seg + (size_t)chain * L.npc * pc_bytes;
    const size_t st_pc = sm_state ? (size_t)L.ctx_count : pc_bytes / 8;   // uint2 entries between the plane contexts of `st`
    const size_t coff = (size_t)s * L.npc * pc_bytes;
    {
        const bool key = B.frame_key[f0];
        const uint2 *in8 = reinterpret_cast<const uint2 *>(B.carry_in + coff);
        uint2 *st8 = reinterpret_cast<uint2 *>(st);
        // VlcState {drift 0, error_sum 4, bias 0, count 1} (ffv1.c:194-199)
        for (int pc = 0; pc < L.npc; pc++)
            for (int i = lane; i < L.ctx_count; i += 32)
                st8[pc * st_pc + i] = key ? make_uint2(0x00040000u, 0x00000100u) : in8[pc * (pc_bytes / 8) + i];
        __syncwarp();
    }
    const int bits = L.coded_bits;
    for (int f = f0; f < f1; f++) {
        const int key = B.frame_key[f] ? 1 : 0;
        uint8_t *out = B.scratch + (size_t)f * L.scratch_per_frame + g.scratch_off + kScratchLead;
        const int hv = (s * 2 + key) * T.nvar + (B.rct_idx ? B.rct_idx[f * L.nslices + s] : 0);       // header variant
        const int npre = T.gprefix_len[hv];
        const uint8_t *pre = T.gprefix + (size_t)hv * kMaxGolombPrefix;
        for (int i = lane; i < npre; i += 32) out[i] = pre[i];
        // every lane walks the lines (32 records per coalesced load, handed to lane 0 through shuffles); lane 0 codes
        {
            BitW w;
            w.buf = out; w.pos = (uint32_t)npre; w.cap = g.scratch_cap - kScratchLead; w.acc = 0; w.nbits = 0;
            const uint32_t *rec_slice = B.rec + (size_t)f * L.rec_per_frame + g.rec_first;
            int run_index = 0;
            for (int li = 0; li < g.nlines; li++) {
                const LineDesc ld = T.lines[g.line_first + li];
                if (!L.rgb && ld.y == 0) run_index = 0;                   // encode_plane starts a new plane (ffv1enc.c:379)
                VlcEnc *vs = reinterpret_cast<VlcEnc *>(st + (size_t)ld.pc * (sm_state ? (size_t)L.ctx_count * 8 : pc_bytes));
                const uint32_t *recp = rec_slice + ld.rec_off;
                int run_count = 0, run_mode = 0;
                uint32_t nx = lane < ld.w ? recp[lane] : 0u;
                for (int x0 = 0; x0 < ld.w; x0 += 32) {
                    const uint32_t rr = nx;
                    if (x0 + 32 + lane < ld.w) nx = recp[x0 + 32 + lane];  // next group in flight while this one is coded
                    const int n = min(32, (int)ld.w - x0);
                    for (int kk = 0; kk < n; kk++) {
                        const uint32_t r = __shfl_sync(0xFFFFFFFFu, rr, kk);
                        if (lane != 0) continue;
                        const int ctx = (int)(r >> 16);
                        int diff = (int)(int16_t)(r & 0xFFFFu);
                        if (ctx == 0) run_mode = 1;
                        if (run_mode) {
                            if (diff) {
                                while (run_count >= 1 << c_enc_log2_run[run_index]) {
                                    run_count -= 1 << c_enc_log2_run[run_index];
                                    run_index++;
                                    bw_put(w, 1, 1u);
                                }
                                bw_put(w, 1 + c_enc_log2_run[run_index], (uint32_t)run_count);
                                if (run_index) run_index--;
                                run_count = 0; run_mode = 0;
                                if (diff > 0) diff--;
                            } else
                                run_count++;
                        }
                        if (!run_mode) vlc_put(w, vs + ctx, diff, bits);
                    }
                }
                if (lane == 0 && run_mode) {                               // end-of-line run flush (ffv1enc.c:358-367)
                    while (run_count >= 1 << c_enc_log2_run[run_index]) {
                        run_count -= 1 << c_enc_log2_run[run_index];
                        run_index++;
                        bw_put(w, 1, 1u);
                    }
                    if (run_count) bw_put(w, 1, 1u);
                }
            }
            if (lane == 0) {
                if (w.nbits) bw_put(w, 8 - w.nbits, 0u);                  // flush_put_bits: zero padding to a byte
                B.slice_bytes[f * L.nslices + s] = w.pos;
                if (w.pos > w.cap) atomicMax(&B.status[1], (unsigned long long)w.pos);
            }
        }
        __syncwarp();
    }
    if (f1 == B.nframes) {
        const uint2 *st8 = reinterpret_cast<const uint2 *>(st);
        uint2 *out8 = reinterpret_cast<uint2 *>(B.carry_out + coff);
        for (int pc = 0; pc < L.npc; pc++)
            for (int i = lane; i < L.ctx_count; i += 32) out8[pc * (pc_bytes / 8) + i] = st8[pc * st_pc + i];
    }
}

// =================================================================================================
// Golomb-Rice mode, decomposed (small context model): the adaptive part -- which Rice parameter k and which sign
// flip a sample gets, put_vlc_symbol / update_vlc_state (ffv1enc.c:240-269, ffv1.h:192-224) -- depends only on the
// samples coded in the same context, so it is replayed per context list like the range coder's states
// (ffv1_ctx_replay.cu builds the lists: hist / scan / scatter in Golomb mode); what is left for the sequential walk of
// a slice is the run mode and the bit writer, and that walk no longer chains the frames of a GOP.
//   k_gr_replay   one lane per (chain, context) list: VlcState in registers -> code word (value | length << 26) of
//                 every coded sample, stored at the sample's record index in the code array (the decision buffer)
//   k_gr_pack     one warp per (frame, slice): run mode (ffv1enc.c:327-367) + MSB-first bit writer over the code words
// =================================================================================================
constexpr int kGrThreads = 128;           // lanes of a chain's CTA: one (or a few) context lists per lane for the whole chain
constexpr int kGrSlots = 8;               // context lists per lane (ctx_count <= 1024)

// One CTA per (GOP segment, slice, plane context) chain; lane <-> context list(s), VlcState in registers for the whole
// chain.  The CTA walks the chain window by window (`window` context tiles of a frame, bounds from k_ctx_scan's tile
// bases staged in shared memory) so that the code words it scatters into the code array land in a stretch that is
// completed while it is still in L2 -- with every lane running through its whole list at its own pace each 4-byte
// store was a read-modify-write of a DRAM sector.
// put_vlc_symbol + set_sr_golomb (ffv1enc.c:240-269, golomb.h:554-563) and update_vlc_state (ffv1.h:192-224) for one
// sample, branch-free: the lanes of a warp run different lists, and the longest list of the chain is the critical path
// of the CTA.  Writes the code word value | length << 26.
__device__ __forceinline__ void gr_code_one(uint32_t *dst, int v, int bits, int &drift, int &esum, int &bias, int &count)
{
    const int cn = count, es = esum;
    v = ((v - bias) << (32 - bits)) >> (32 - bits);      // fold
    // k = smallest k with (count << k) >= error_sum: floor-log2 difference, plus one if that falls short
    const int t = max(__clz(cn) - __clz(es), 0);
    const int k = es > cn ? t + (((uint32_t)cn << t) < (uint32_t)es ? 1 : 0) : 0;
    const int cd = v ^ ((2 * drift + cn) >> 31);
    int m = -2 * cd - 1;
    m ^= m >> 31;
    const int e = m >> k;
    const bool esc = e >= 12;
    const uint32_t len = esc ? (uint32_t)(12 + bits) : (uint32_t)(e + k + 1);
    const uint32_t val = esc ? (uint32_t)(m - 11) : (1u << k) + ((uint32_t)m & ((1u << k) - 1u));
    *dst = val | (len << 26);
    int d2 = drift + v, e2 = es + abs(v), c2 = cn;
    const int half = cn == 128 ? 1 : 0;
    c2 >>= half; d2 >>= half; e2 >>= half;
    c2++;
    const bool neg = d2 <= -c2, pos = d2 > 0;
    int b2 = bias + (pos ? 1 : 0) - (neg ? 1 : 0);
    b2 = max(-128, min(127, b2));
    d2 += neg ? c2 : (pos ? -c2 : 0);
    if (neg) d2 = max(d2, -c2 + 1);
    if (pos) d2 = min(d2, 0);
    drift = d2; esum = e2; count = c2; bias = b2;
}

// TILED (Layout::tiled_lists, ffv1_ctx_replay.cu:k_tile_sort<GOLOMB>): the lists are kept tile by tile -- a window is one
// tile, a list's part of it is the context's run inside the tile's block (4-byte entries: record index inside the
// tile | residual << 22), and the lane <-> context assignment is by how often the contexts occur in the chain's first frame.
template <bool TILED>
__global__ void __launch_bounds__(kGrThreads, 8) k_gr_replay(const EncDeviceTables T, const EncBatch B, const int window_rt)
{
    const int window = TILED ? 1 : window_rt;
    extern __shared__ __align__(16) uint32_t s_gr[];                        // [2][nctx] the window's part of every list
    const Layout &L = T.layout;
    const int tid = threadIdx.x;
    const int chain = chain_of_block(L, T.slices);          // longest chains first
    const int pc = chain % L.npc, s = (chain / L.npc) % L.nslices, seg = chain / (L.npc * L.nslices);
    const SliceGeom &g = T.slices[s];
    const int f0 = B.seg_first[seg], f1 = B.seg_first[seg + 1];
    const int nctx = L.ctx_count;
    uint32_t *s_b0 = s_gr, *s_b1 = s_gr + nctx;
    const uint32_t *lcount = B.list_count + (size_t)chain * nctx;
    const uint2 *chain_list = B.lists + (size_t)f0 * L.samples_per_frame + (size_t)(f1 - f0) * g.list_off[pc];
    const size_t pc_bytes = (size_t)nctx * 32;
    const size_t sbase = ((size_t)s * L.npc + pc) * (pc_bytes / 8);
    const bool key = B.frame_key[f0] != 0;
    const int bits = L.coded_bits;
    const int t0 = g.ct_first[pc], nt = g.ct_count[pc];
    uint16_t *s_ord = reinterpret_cast<uint16_t *>(s_b1);                       // TILED: contexts by decreasing frequency (set-up only)
    if (TILED) {
        for (int i = tid; i < nctx; i += kGrThreads) {
            uint32_t n = 0;
            for (int tt = 0; tt < nt; tt++) {
                const uint16_t *tab = B.tile_tab + ((size_t)f0 * L.ctiles_per_frame + t0 + tt) * B.tile_tab_pitch;
                n += (uint32_t)tab[i + 1] - (uint32_t)tab[i];
            }
            s_b0[i] = n;
        }
        __syncthreads();
        for (int c = tid; c < nctx; c += kGrThreads) {
            const uint32_t mine = s_b0[c];
            int rank = 0;
            for (int o = 0; o < nctx; o++) {
                const uint32_t v = s_b0[o];
                rank += (v > mine) || (v == mine && o < c);
            }
            s_ord[rank] = (uint16_t)c;
        }
        __syncthreads();
    }
    // my lists: contexts order[tid], order[tid + 128], ... (the order is by decreasing list length)
    int ctx[kGrSlots], drift[kGrSlots], esum[kGrSlots], bias[kGrSlots], count[kGrSlots];
    uint32_t lstart[kGrSlots];
#pragma unroll
    for (int j = 0; j < kGrSlots; j++) {
        const int oi = tid + j * kGrThreads;
        ctx[j] = -1; drift[j] = 0; esum[j] = 4; bias[j] = 0; count[j] = 1; lstart[j] = 0u;   // {0,4,0,1} on keyframes (ffv1.c:194-199)
        if (oi < nctx) {
            const int c = TILED ? (int)s_ord[oi] : (int)B.list_order[(size_t)chain * nctx + oi];
            ctx[j] = c;
            if (!TILED) lstart[j] = B.list_start[(size_t)chain * nctx + c];
            if (!key) {
                const uint2 v = reinterpret_cast<const uint2 *>(B.carry_in)[sbase + c];
                drift[j] = (int)(int16_t)(v.x & 0xFFFFu); esum[j] = (int)(v.x >> 16);
                bias[j] = (int)(int8_t)(v.y & 0xFFu); count[j] = (int)((v.y >> 8) & 0xFFu);
            }
        }
    }
    const size_t code_frame = L.dec_per_frame / 2;                          // 32-bit words per frame in the code array
    const uint32_t *chain_list32 = reinterpret_cast<const uint32_t *>(B.lists) + (size_t)f0 * L.samples_per_frame + (size_t)(f1 - f0) * g.list_off[pc];
    __syncthreads();                                                        // (TILED: s_ord is read by everybody before s_b1 is reused)
    for (int f = f0; f < f1; f++) {
        uint32_t *code = reinterpret_cast<uint32_t *>(B.dec) + (size_t)f * code_frame + g.rec_first;
        for (int tw = 0; tw < nt; tw += window) {
            __syncthreads();                                                // everybody is done with the previous window's bounds
            const bool last_win = tw + window >= nt;
            const uint32_t *blk = chain_list32;
            uint32_t *code_w = code;
            if (TILED) {
                const CtxTile ct = T.ctiles[t0 + tw];
                const uint16_t *tab = B.tile_tab + ((size_t)f * L.ctiles_per_frame + t0 + tw) * B.tile_tab_pitch;
                for (int i = tid; i < nctx; i += kGrThreads) { s_b0[i] = tab[i]; s_b1[i] = tab[i + 1]; }
                blk = chain_list32 + (size_t)(f - f0) * g.pc_samples[pc] + ct.sample_first;
                // the tile's record indices count from its first line
                code_w = code + T.lines[g.line_first + T.pc_lines[g.pc_line_first[pc] + ct.first]].rec_off;
            } else {
            const uint32_t *bf = B.ctx_hist + ((size_t)f * L.ctiles_per_frame + t0 + tw) * nctx;
            const uint32_t *bn = last_win ? B.ctx_hist + ((size_t)(f + 1) * L.ctiles_per_frame + t0) * nctx
                                          : B.ctx_hist + ((size_t)f * L.ctiles_per_frame + t0 + tw + window) * nctx;
            const bool use_bn = !last_win || f + 1 < f1;
            for (int i = tid; i < nctx; i += kGrThreads) { s_b0[i] = bf[i]; s_b1[i] = use_bn ? bn[i] : lcount[i]; }
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < kGrSlots; j++) {
                if (ctx[j] < 0) continue;
                const uint32_t b0 = s_b0[ctx[j]], b1 = s_b1[ctx[j]];
                if (TILED) {
                    // same walk over 4-byte entries of the tile's block
                    uint32_t nxt[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) nxt[q] = b0 + q < b1 ? blk[b0 + q] : 0u;
                    for (uint32_t i = b0; i < b1; i += 4u) {
                        uint32_t cur[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) { cur[q] = nxt[q]; if (i + 4u + q < b1) nxt[q] = blk[i + 4u + q]; }
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            if (i + q >= b1) break;
                            gr_code_one(code_w + (cur[q] & 0x3FFFFFu), (int)cur[q] >> 22, bits, drift[j], esum[j], bias[j], count[j]);
                        }
                    }
                    continue;
                }
                const uint2 *lp = chain_list + lstart[j];
                // four entries per round, the next round's loads in flight while this one is coded: with ~2000 lists
                // streaming per SM the lines do not survive in L1, so every load is an L2 (or DRAM) round trip
                uint2 nxt[4];
#pragma unroll
                for (int q = 0; q < 4; q++) nxt[q] = b0 + q < b1 ? lp[b0 + q] : make_uint2(0u, 0u);
                for (uint32_t i = b0; i < b1; i += 4u) {
                    uint2 cur[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) { cur[q] = nxt[q]; if (i + 4u + q < b1) nxt[q] = lp[i + 4u + q]; }
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if (i + q >= b1) break;
                        const uint2 en = cur[q];
                        gr_code_one(code + en.x, (int)(int16_t)(en.y & 0xFFFFu), bits, drift[j], esum[j], bias[j], count[j]);
                    }
                }
            }
        }
    }
    if (f1 == B.nframes) {
#pragma unroll
        for (int j = 0; j < kGrSlots; j++)
            if (ctx[j] >= 0)
                reinterpret_cast<uint2 *>(B.carry_out)[sbase + ctx[j]] = make_uint2(((uint32_t)drift[j] & 0xFFFFu) | ((uint32_t)esum[j] << 16),
                                                                                    ((uint32_t)bias[j] & 0xFFu) | ((uint32_t)count[j] << 8));
    }
}

constexpr int kGrPackWarps = 4;            // (frame, slice) units per CTA
constexpr int kGrPackWin = 96;             // 32-bit words of the per-warp bit window (a group appends at most ~75)

// put_bits (put_bits.h:152-195) into the warp's shared-memory window: `len` bits of `val` at bit position p counted
// from the start of the slice's scratch payload; wbase = index of the window's first word.  MSB-first.
__device__ __forceinline__ void gr_window_or(uint32_t *win, uint32_t wbase, uint32_t p, uint32_t len, uint32_t val, bool atomic)
{
    const uint32_t sh = p & 31u, wi = (p >> 5) - wbase;
    const unsigned long long v64 = (unsigned long long)val << (64u - len - sh);     // len + sh <= 57
    const uint32_t hi = (uint32_t)(v64 >> 32), lo = (uint32_t)v64;
    if (atomic) { atomicOr(&win[wi], hi); if (lo) atomicOr(&win[wi + 1], lo); }
    else        { win[wi] |= hi; win[wi + 1] |= lo; }
}

// One warp per (frame, slice).  Groups of 32 samples without a run-mode sample (the normal case on natural content) are
// packed by all lanes at once: exclusive scan of the code lengths, every lane ORs its code word into the window, the
// completed words go out as big-endian 32-bit stores.  Groups that touch run mode (ffv1enc.c:327-367) are walked by
// lane 0 through the same window.
__global__ void __launch_bounds__(32 * kGrPackWarps) k_gr_pack(const EncDeviceTables T, const EncBatch B)
{
    __shared__ uint32_t s_win[kGrPackWarps][kGrPackWin];
    const Layout &L = T.layout;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int unit = blockIdx.x * kGrPackWarps + warp;
    if (unit >= B.nframes * L.nslices) return;
    const int f = unit / L.nslices, s = unit - f * L.nslices;
    const SliceGeom &g = T.slices[s];
    const int key = B.frame_key[f] ? 1 : 0;
    uint8_t *out = B.scratch + (size_t)f * L.scratch_per_frame + g.scratch_off + kScratchLead;     // 4-byte aligned
    uint32_t *out32 = reinterpret_cast<uint32_t *>(out);
    const uint32_t cap = g.scratch_cap - kScratchLead;
    const int hv = (s * 2 + key) * T.nvar + (B.rct_idx ? B.rct_idx[f * L.nslices + s] : 0);           // header variant
    const int npre = T.gprefix_len[hv];
    const uint8_t *pre = T.gprefix + (size_t)hv * kMaxGolombPrefix;
    for (int i = lane; i < (npre & ~3); i += 32) if ((uint32_t)i < cap) out[i] = pre[i];             // whole words of the prefix
    uint32_t *win = s_win[warp];
    for (int i = lane; i < kGrPackWin; i += 32) win[i] = 0u;
    __syncwarp();
    if (lane < (npre & 3)) atomicOr(&win[0], (uint32_t)pre[(npre & ~3) + lane] << (24 - 8 * lane)); // its last, partial word
    __syncwarp();
    uint32_t P = (uint32_t)npre * 8u;                                      // bits written (warp-uniform)
    uint32_t wbase = P >> 5;
    const uint32_t *rec_slice = B.rec + (size_t)f * L.rec_per_frame + g.rec_first;
    const uint32_t *code_slice = reinterpret_cast<const uint32_t *>(B.dec) + (size_t)f * (L.dec_per_frame / 2) + g.rec_first;
    int run_index = 0;                                                     // warp-uniform (lane 0 changes it on the serial paths)
    for (int li = 0; li < g.nlines; li++) {
        const LineDesc ld = T.lines[g.line_first + li];
        if (!L.rgb && ld.y == 0) run_index = 0;                           // encode_plane starts a new plane (ffv1enc.c:379)
        const uint32_t *recp = rec_slice + ld.rec_off, *codep = code_slice + ld.rec_off;
        int run_count = 0;                                                 // lane 0; run mode at a group start = carry
        uint32_t carry = 0u, run_open = 0u;
        uint32_t nr = lane < ld.w ? recp[lane] : 0xFFFF0001u, nc = lane < ld.w ? codep[lane] : 0u;
        for (int x0 = 0; x0 < ld.w; x0 += 32) {
            const bool act = x0 + lane < ld.w;
            const uint32_t rr = nr, cc = nc;
            nr = 0xFFFF0001u;
            if (x0 + 32 + lane < ld.w) { nr = recp[x0 + 32 + lane]; nc = codep[x0 + 32 + lane]; }
            const uint32_t zero = __ballot_sync(0xFFFFFFFFu, act && (rr & 0xFFFFu) == 0u);
            const uint32_t ctx0 = __ballot_sync(0xFFFFFFFFu, act && (rr >> 16) == 0u);
            const uint32_t cin = carry;
            // same carry chain as in the list builder (gr_run_members, ffv1_ctx_replay.cu)
            const unsigned long long sum = (unsigned long long)zero + (zero & ctx0) + cin;
            const uint32_t mem = ctx0 | ((uint32_t)sum ^ zero ^ (zero & ctx0));
            carry = (uint32_t)(sum >> 32);
            run_open = ((mem & zero) >> (min(32, (int)ld.w - x0) - 1)) & 1u;   // the line's last sample so far is inside a run
            uint32_t total;
            if (((mem & zero) | cin) == 0u) {
                // ---- no sample is absorbed into a run: all lanes at once.  A run-mode sample of such a group ends a run of
                //      length 0 at once: 1 + log2_run[run_index] zero bits in front of its code, then run_index-- (ffv1enc.c:
                //      338-346), so its run_index is the group's minus the run-mode samples before it
                const int ri = max(run_index - (int)__popc(mem & ((1u << lane) - 1u)), 0);
                const uint32_t inrun = (mem >> lane) & 1u;
                const uint32_t len = act ? (cc >> 26) + (inrun ? 1u + c_enc_log2_run[ri] : 0u) : 0u;
                run_index = max(run_index - (int)__popc(mem), 0);
                uint32_t incl = len;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += n; }
                total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                if (len) gr_window_or(win, wbase, P + incl - len, len, cc & 0x3FFFFFFu, true);
            } else {
                // ---- lane 0 walks the group: run mode + codes
                uint32_t p = P;
                const int n = min(32, (int)ld.w - x0);
                for (int kk = 0; kk < n; kk++) {
                    const uint32_t r = __shfl_sync(0xFFFFFFFFu, rr, kk), cw = __shfl_sync(0xFFFFFFFFu, cc, kk);
                    if (lane != 0) continue;
                    const bool nz = (r & 0xFFFFu) != 0u;
                    int run_mode = (int)((mem >> kk) & 1u);              // M(kk): this sample is coded in run mode
                    if (run_mode) {
                        if (nz) {
                            while (run_count >= 1 << c_enc_log2_run[run_index]) {
                                run_count -= 1 << c_enc_log2_run[run_index];
                                run_index++;
                                gr_window_or(win, wbase, p, 1u, 1u, false); p += 1u;
                            }
                            const uint32_t ln = 1u + c_enc_log2_run[run_index];
                            gr_window_or(win, wbase, p, ln, (uint32_t)run_count, false); p += ln;
                            if (run_index) run_index--;
                            run_count = 0; run_mode = 0;
                        } else
                            run_count++;
                    }
                    if (!run_mode) { const uint32_t ln = cw >> 26; gr_window_or(win, wbase, p, ln, cw & 0x3FFFFFFu, false); p += ln; }
                }
                total = __shfl_sync(0xFFFFFFFFu, p, 0) - P;
                run_index = __shfl_sync(0xFFFFFFFFu, run_index, 0);
            }
            P += total;
            __syncwarp();
            // ---- completed words -> scratch (big-endian), the partial one moves to the front of the window
            const uint32_t nfull = (P >> 5) - wbase;
            uint32_t keep = 0u;
            if (nfull) {
                for (uint32_t i = lane; i < nfull; i += 32)
                    if ((wbase + i) * 4u + 4u <= cap) out32[wbase + i] = __byte_perm(win[i], 0u, 0x0123);
                keep = win[nfull];
                __syncwarp();
                for (uint32_t i = lane; i <= nfull + 1u && i < (uint32_t)kGrPackWin; i += 32) win[i] = 0u;
                __syncwarp();
                if (lane == 0) win[0] = keep;
                wbase += nfull;
                __syncwarp();
            }
        }
        // end-of-line run flush (ffv1enc.c:358-367): lane 0, a handful of bits
        if (run_open) {
            uint32_t p = P;
            if (lane == 0) {
                while (run_count >= 1 << c_enc_log2_run[run_index]) {
                    run_count -= 1 << c_enc_log2_run[run_index];
                    run_index++;
                    gr_window_or(win, wbase, p, 1u, 1u, false); p += 1u;
                }
                if (run_count) { gr_window_or(win, wbase, p, 1u, 1u, false); p += 1u; }
            }
            P = __shfl_sync(0xFFFFFFFFu, p, 0);
            run_index = __shfl_sync(0xFFFFFFFFu, run_index, 0);
            __syncwarp();
            const uint32_t nfull = (P >> 5) - wbase;
            if (nfull) {
                for (uint32_t i = lane; i < nfull; i += 32)
                    if ((wbase + i) * 4u + 4u <= cap) out32[wbase + i] = __byte_perm(win[i], 0u, 0x0123);
                const uint32_t keep = win[nfull];
                __syncwarp();
                for (uint32_t i = lane; i <= nfull + 1u && i < (uint32_t)kGrPackWin; i += 32) win[i] = 0u;
                __syncwarp();
                if (lane == 0) win[0] = keep;
                wbase += nfull;
                __syncwarp();
            }
        }
    }
    // flush_put_bits: zero padding to a byte; the last, partial word goes out byte by byte
    const uint32_t nbytes = (P + 7u) >> 3;
    {
        const uint32_t first = wbase * 4u;
        const uint32_t wv = win[0];
        if (lane < 4 && first + lane < nbytes && first + lane < cap) out[first + lane] = (uint8_t)(wv >> (24 - 8 * lane));
    }
    if (lane == 0) {
        B.slice_bytes[f * L.nslices + s] = nbytes;
        if (nbytes > cap) atomicMax(&B.status[1], (unsigned long long)nbytes);
    }
}

void launch_golomb_coder(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s)
{
    const Layout &L = t.layout;
    int window = 3;
    if (const char *v = getenv("FFV1B200_GOLOMB_WINDOW")) { window = atoi(v); if (window < 1) window = 1 << 20; }
    if (L.tiled_lists) k_gr_replay<true><<<b.nseg * L.nslices * L.npc, kGrThreads, 2 * L.ctx_count * sizeof(uint32_t), s>>>(t, b, 1);
    else k_gr_replay<false><<<b.nseg * L.nslices * L.npc, kGrThreads, 2 * L.ctx_count * sizeof(uint32_t), s>>>(t, b, window);
    k_gr_pack<<<(b.nframes * L.nslices + kGrPackWarps - 1) / kGrPackWarps, 32 * kGrPackWarps, 0, s>>>(t, b);
}

void launch_golomb(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s)
{
    // VlcStates of the chain (8 bytes per context and plane context) in shared memory when they fit
    const size_t bytes = (size_t)t.layout.npc * t.layout.ctx_count * 8;
    const int sm_state = bytes <= 40 * 1024;
    k_golomb<<<b.nseg * t.layout.nslices, 32, sm_state ? bytes : 0, s>>>(t, b, sm_state);
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

constexpr int kPackThreads = 256;

// ---- carry resolution of k_rangecode's output (16-bit entries v[k] = byte | carry << 8, see k_rangecode) ----------------
// What is added to entry k-1: the carry out of entries k, k+1, ... -- an entry of exactly 0x0FF passes a carry on, an
// entry with bit 8 set generates one, anything else absorbs it (renorm_encoder's outstanding_count / outstanding_byte
// bookkeeping, rangecoder.h:52-75, read backwards).  n = number of entries (the dropped last one included).
__device__ __forceinline__ uint32_t pack_carry_from(const uint16_t *v, uint32_t k, uint32_t n)
{
    for (; k < n; k++) {
        const uint32_t e = v[k];
        if (e != 0xFFu) return e >> 8;
    }
    return 0u;
}

// payload bytes k..k+3 (k + 3 < n), first byte in the most significant position
__device__ __forceinline__ uint32_t pack_resolved_be(const uint16_t *v, uint32_t k, uint32_t n)
{
    const uint32_t e0 = v[k], e1 = v[k + 1], e2 = v[k + 2], e3 = v[k + 3];
    return (e0 << 24) + (e1 << 16) + (e2 << 8) + e3 + pack_carry_from(v, k + 4u, n);
}

__device__ __forceinline__ uint32_t pack_resolved_byte(const uint16_t *v, uint32_t k, uint32_t n)
{
    return (v[k] + pack_carry_from(v, k + 1u, n)) & 0xFFu;
}

// One CTA per (frame, slice): copy the coder output to its final place, append the 24-bit length and the
// error-check trailer.
//   CRC (libavutil AV_CRC_32_IEEE: MSB-first 0x04C11DB7, init 0, no final xor = plain polynomial remainder
//   M(x) * x^32 mod P), computed in the SAME pass as the copy and over the same destination-aligned words: thread t
//   takes the words t, t + 256, t + 512, ... (the accesses of a warp are contiguous) and keeps
//       c = sum_j W[t + 256 j] * x^(32 * 256 * (J - 1 - j)) mod P      (Horner: c = c * x^8192 mod P  xor  W)
//   where the multiplication by the constant x^8192 is four table lookups on the bytes of c -- the cost of an ordinary
//   slicing-by-4 step.  By linearity crc(message) = XOR over threads of c * x^(8 * (bytes behind the thread's last word)
//   + 32) mod P (square-and-multiply over precomputed x^(8*2^j)), plus the few head / tail / trailer bytes, which one
//   thread each folds byte by byte.  (Before: a second pass in which every thread walked a contiguous chunk -- the 32
//   lanes of a load 400 bytes apart -- took twice the copy's time.)
//   Copy: the packet position has arbitrary byte alignment, so destination-aligned 32-bit words are built from two
//   source words with a funnel shift (Golomb-Rice) or from the 16-bit entries; the stores of a warp are contiguous.
__device__ uint32_t g_pack_tab0[256];          // (b * x^32) mod P: the byte-wise table
__device__ uint32_t g_pack_mul[4][256];        // (b * x^(8k + 32 * kPackThreads)) mod P
__device__ uint32_t g_pack_pow[32];            // x^(8 * 2^j) mod P

__global__ void __launch_bounds__(1024) k_pack_init()
{
    __shared__ uint32_t s_pow[32];
    const int tid = threadIdx.x;
    if (tid == 0) {
        uint32_t p = 0x100u;                    // x^8
        for (int j = 0; j < 32; j++) { s_pow[j] = p; g_pack_pow[j] = p; p = gf_mulmod(p, p); }
    }
    __syncthreads();
    if (tid < 256) {
        uint32_t c = (uint32_t)tid << 24;
#pragma unroll
        for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : (c << 1);
        g_pack_tab0[tid] = c;
    }
    // x^(32 * 256) = x^(8 * 2^10)
    static_assert(kPackThreads == 256, "the stride multiplier below is x^(8 * 2^10)");
    const int k = tid >> 8, bb = tid & 255;
    g_pack_mul[k][bb] = gf_mulmod((uint32_t)bb << (8 * k), s_pow[10]);
}

template <bool C16>       // C16: the coder output is k_rangecode's 16-bit entries (range-coder modes); else plain bytes (Golomb-Rice)
__global__ void __launch_bounds__(kPackThreads) k_pack_slices(const EncDeviceTables T, const EncBatch B)
{
    __shared__ uint32_t s_tab0[256];            // (b * x^32) mod P
    __shared__ uint32_t s_mul[4][256];          // (b * x^(8k + 8192)) mod P
    __shared__ uint32_t s_pow[32];              // x^(8 * 2^j) mod P
    __shared__ uint32_t s_red[kPackThreads / 32];
    const Layout &L = T.layout;
    const int ns = L.nslices;
    const int f = blockIdx.x / ns, s = blockIdx.x - f * ns;
    const int tid = threadIdx.x, lane = tid & 31;
    if (B.status[0] | B.status[1] | B.status[2]) return;

    s_tab0[tid] = g_pack_tab0[tid];
#pragma unroll
    for (int k = 0; k < 4; k++) s_mul[k][tid] = g_pack_mul[k][tid];
    if (tid < 32) s_pow[tid] = g_pack_pow[tid];
    const bool has_len = (T.version > 2) || s > 0;
    unsigned long long off = B.pkt_off[f];
    for (int k = 0; k < s; k++) off += B.slice_bytes[f * ns + k] + ((T.version > 2 || k > 0) ? 3 : 0) + (T.ec ? 5 : 0);
    const uint32_t nb = B.slice_bytes[f * ns + s];
    const uint8_t *src = B.scratch + (size_t)f * L.scratch_per_frame + T.slices[s].scratch_off + kScratchLead;   // 4-byte aligned
    uint8_t *dst = B.out + off;
    const uint16_t *v16 = reinterpret_cast<const uint16_t *>(B.scratch) + (size_t)f * L.scratch_per_frame + T.slices[s].scratch_off;
    const uint32_t nent = nb + 1u;                                        // entries of the slice (C16)
    __syncthreads();

    // ---- head bytes up to the first aligned destination word, aligned words, tail bytes
    const uint32_t head = min(nb, (uint32_t)((4u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u));
    const uint32_t nwords = (nb - head) >> 2;
    const uint32_t done = head + nwords * 4u;
    const uint32_t tail_bytes = (has_len ? 3u : 0u) + 1u;                 // len24 | 0x00 (the CRC covers them)
    uint32_t *dst32 = reinterpret_cast<uint32_t *>(dst + head);
    const uint32_t *src32 = reinterpret_cast<const uint32_t *>(src);
    const uint32_t sh = head * 8u;                                        // source byte offset inside a word (Golomb-Rice)
    uint32_t c = 0u;                                                      // bit k = coefficient of x^k
    uint32_t last_w = 0u;
    bool any = false;
    for (uint32_t w = tid; w < nwords; w += kPackThreads) {
        uint32_t le;                                                      // the word as it lies in memory
        if (C16) le = __byte_perm(pack_resolved_be(v16, head + 4u * w, nent), 0u, 0x0123);
        else {
            const uint32_t lo = src32[w];
            const uint32_t hi = head ? src32[w + 1] : 0u;                 // within the scratch region (cap is padded)
            le = head ? __funnelshift_r(lo, hi, sh) : lo;
        }
        dst32[w] = le;
        c = s_mul[0][c & 0xFFu] ^ s_mul[1][(c >> 8) & 0xFFu] ^ s_mul[2][(c >> 16) & 0xFFu] ^ s_mul[3][c >> 24];
        c ^= __byte_perm(le, 0u, 0x0123);                                 // first message byte = highest degree
        last_w = w; any = true;
    }
    if ((uint32_t)tid < head) dst[tid] = C16 ? (uint8_t)pack_resolved_byte(v16, (uint32_t)tid, nent) : src[tid];
    if ((uint32_t)tid < nb - done) dst[done + tid] = C16 ? (uint8_t)pack_resolved_byte(v16, done + (uint32_t)tid, nent) : src[done + tid];
    if (tid == 0) {
        // trailer: 24-bit big-endian payload length, then 0x00 (the CRC itself is written below)
        uint32_t q = nb;
        if (has_len) { dst[q] = (uint8_t)(nb >> 16); dst[q + 1] = (uint8_t)(nb >> 8); dst[q + 2] = (uint8_t)nb; q += 3; }
        if (T.ec) dst[q] = 0;
    }
    if (!T.ec) return;

    // ---- CRC over payload | len24 | 0x00: scale the partial remainders to the end of the message and XOR them
    uint32_t crc = 0u, after = 0u;               // crc = (this thread's bytes) * x^32 mod P, `after` message bytes behind them
    if (any) {
        // c is a plain remainder (no x^32 yet): four more bytes of shift
        crc = c;
        after = (nwords - 1u - last_w) * 4u + (nb - done) + tail_bytes + 4u;
    }
    if (crc) {
        uint32_t mult = 1u;
        for (int j = 0; after; j++, after >>= 1) if (after & 1u) mult = gf_mulmod(mult, s_pow[j]);
        crc = gf_mulmod(crc, mult);
    }
    // the head bytes and the tail + trailer bytes: one thread each, byte by byte
    uint32_t extra = 0u;
    if (tid == 0) {
        uint32_t h = 0u;
        for (uint32_t i = 0; i < head; i++) h = (h << 8) ^ s_tab0[(h >> 24) ^ (C16 ? pack_resolved_byte(v16, i, nent) : (uint32_t)src[i])];
        if (h) {
            uint32_t aft = nb - head + tail_bytes, mult = 1u;
            for (int j = 0; aft; j++, aft >>= 1) if (aft & 1u) mult = gf_mulmod(mult, s_pow[j]);
            extra = gf_mulmod(h, mult);
        }
    }
    if (tid == 1) {
        uint32_t t = 0u;
        for (uint32_t i = done; i < nb; i++) t = (t << 8) ^ s_tab0[(t >> 24) ^ (C16 ? pack_resolved_byte(v16, i, nent) : (uint32_t)src[i])];
        if (has_len) {
            t = (t << 8) ^ s_tab0[(t >> 24) ^ ((nb >> 16) & 0xFFu)];
            t = (t << 8) ^ s_tab0[(t >> 24) ^ ((nb >> 8) & 0xFFu)];
            t = (t << 8) ^ s_tab0[(t >> 24) ^ (nb & 0xFFu)];
        }
        t = (t << 8) ^ s_tab0[t >> 24];                                  // the 0x00 byte
        extra = t;
    }
    crc ^= extra;
#pragma unroll
    for (int d = 16; d; d >>= 1) crc ^= __shfl_xor_sync(0xFFFFFFFFu, crc, d);
    if (lane == 0) s_red[tid >> 5] = crc;
    __syncthreads();
    if (tid == 0) {
        uint32_t r = 0;
        for (int i = 0; i < kPackThreads / 32; i++) r ^= s_red[i];
        const uint32_t mlen = nb + tail_bytes;
        dst[mlen + 0] = (uint8_t)(r >> 24); dst[mlen + 1] = (uint8_t)(r >> 16);
        dst[mlen + 2] = (uint8_t)(r >> 8);  dst[mlen + 3] = (uint8_t)r;
    }
}

void launch_pack(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s)
{
    k_pack_layout<<<1, 1024, 0, s>>>(t, b);
    k_pack_init<<<1, 1024, 0, s>>>();            // constant tables (per device; a microsecond)
    if (t.layout.golomb) k_pack_slices<false><<<b.nframes * t.layout.nslices, kPackThreads, 0, s>>>(t, b);
    else                 k_pack_slices<true><<<b.nframes * t.layout.nslices, kPackThreads, 0, s>>>(t, b);
}

// =================================================================================================
// k_pass1_stats: the statistics a first pass collects (AV_CODEC_FLAG_PASS1; put_symbol_inline's rc_stat / rc_stat2
// arguments, ffv1enc.c:193-200, 320-322), from what the encode left in device memory:
//   rc_stat[state][bit]            <- the decision entries (state before the update | bit << 8) of the sample runs
//   rc_stat2[context][slot][bit]   <- the records: which slots a residual visits and with which bits does not depend on
//                                     the adaptive state (ffv1enc.c:202-229)
// One CTA per (slice, frame).  PRIV: the per-context counters of the CTA are gathered in shared memory first (small
// context model: 666 x 64 counters); otherwise every count is a global atomic.  Not on the speed path.
// =================================================================================================
constexpr int kStatsThreads = 512;

template <bool PRIV>
__global__ void __launch_bounds__(kStatsThreads) k_pass1_stats(const EncDeviceTables T, const EncBatch B, unsigned long long *rc_stat,
                                                               unsigned long long *rc_stat2)
{
    extern __shared__ uint32_t s_stat2[];                    // PRIV: [ctx_count * 64]
    __shared__ uint32_t s_stat[512];
    const Layout &L = T.layout;
    const int s = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const SliceGeom &g = T.slices[s];
    if (B.status[0] | B.status[1] | B.status[2]) return;     // a scratch area overflowed: the batch is coded again
    for (int i = tid; i < 512; i += kStatsThreads) s_stat[i] = 0u;
    if (PRIV) for (int i = tid; i < L.ctx_count * 64; i += kStatsThreads) s_stat2[i] = 0u;
    __syncthreads();
    // ---- records -> per-context slot/bit counts
    const uint32_t *rec_slice = B.rec + (size_t)f * L.rec_per_frame + g.rec_first;
    for (int li = 0; li < g.nlines; li++) {
        const LineDesc ld = T.lines[g.line_first + li];
        const uint32_t *recp = rec_slice + ld.rec_off;
        for (int x = tid; x < ld.w; x += kStatsThreads) {
            const uint32_t r = recp[x];
            const uint32_t base = (r >> 16) * 64u;
            const int d = (int)(int16_t)(r & 0xFFFFu);
            auto count = [&](int slot, int bit) {
                if (PRIV) atomicAdd(&s_stat2[base + slot * 2 + bit], 1u);
                else atomicAdd(&rc_stat2[base + slot * 2 + bit], 1ull);
            };
            if (d == 0) { count(0, 1); continue; }
            const uint32_t a = (uint32_t)abs(d);
            const int e = 31 - __clz(a);
            count(0, 0);
            for (int i = 0; i < e; i++) count(1 + min(i, 9), 1);
            count(1 + min(e, 9), 0);
            for (int i = e - 1; i >= 0; i--) count(22 + min(i, 9), (a >> i) & 1u);
            count(11 + min(e, 10), d < 0);
        }
    }
    // ---- decisions of the slice's sample runs -> counts per (state, bit)
    {
        const uint16_t *dec_frame = B.dec + (size_t)f * L.dec_per_frame;
        const uint32_t *run_cnt = B.run_cnt + (size_t)f * L.runs_per_frame + g.run_first;
        const uint8_t *run_pc = T.run_pc + g.run_first;
        uint32_t cur[3] = {0u, 0u, 0u};
        for (int r = 0; r < g.nruns; r++) {
            const int pc = run_pc[r];
            const uint32_t n = run_cnt[r];
            const uint16_t *src = dec_frame + g.dec_off[pc] + cur[pc];
            cur[pc] = (cur[pc] + n + 7u) & ~7u;
            for (uint32_t i = tid; i < n; i += kStatsThreads) atomicAdd(&s_stat[src[i] & 0x1FFu], 1u);
        }
    }
    __syncthreads();
    for (int i = tid; i < 512; i += kStatsThreads)
        if (s_stat[i]) atomicAdd(&rc_stat[(i & 0xFF) * 2 + (i >> 8)], (unsigned long long)s_stat[i]);
    if (PRIV)
        for (int i = tid; i < L.ctx_count * 64; i += kStatsThreads)
            if (s_stat2[i]) atomicAdd(&rc_stat2[i], (unsigned long long)s_stat2[i]);
}

cudaError_t launch_pass1_stats(const EncDeviceTables &t, const EncBatch &b, unsigned long long *rc_stat, unsigned long long *rc_stat2,
                               cudaStream_t s)
{
    const Layout &L = t.layout;
    dim3 grid(L.nslices, b.nframes);
    const int smem = L.ctx_count * 64 * 4;
    if (smem <= 200 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_pass1_stats<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k_pass1_stats<true><<<grid, kStatsThreads, smem, s>>>(t, b, rc_stat, rc_stat2);
    } else
        k_pass1_stats<false><<<grid, kStatsThreads, 0, s>>>(t, b, rc_stat, rc_stat2);
    return cudaGetLastError();
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
        e = cudaFuncSetAttribute(k_replay<true, 7, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, rsm); if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_replay<true, 9, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, rsm); if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_replay<true, 9, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, rsm); if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

} // namespace ffv1
