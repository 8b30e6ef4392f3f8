// ffv1_pixel_fast.cu -- the per-pixel pass for planar formats, tuned for the HBM roofline (sm_100a).
//
// Same function as k_pixel (ffv1_enc_kernels.cu): sample fetch, slice-local neighbourhood, get_context (ffv1.h:170-190),
// median predictor (ffv1.h:161-168), residual, sign flip, fold (ffv1.h:148-159) -> (context<<16 | diff) records.
// Differences that make it fast:
//   * persistent CTAs (2 per SM) walk the (frame, slice-plane, 16-row) tiles; the rows of the NEXT tile are fetched
//     by the TMA unit (cp.async.bulk global->shared, one bulk copy per row, completion on an mbarrier) while the
//     current tile is computed: no thread ever issues a global load;
//   * the quantisation tables live in shared memory replicated once per lane (entry e of lane l at e*256 + l*4), so the
//     data-dependent lookups of a warp never collide in a bank; entries are indexed by (difference*256) & 0xFF00,
//     which is what a byte extraction with PRMT (value << 8) yields for free;
//   * each thread handles 8 consecutive samples held in registers and writes one 32-byte (256-bit) record vector.
// Requirements checked on the host (else the generic kernel runs): 8- or 16-bit planar source, every source plane
// 16-byte aligned with a 16-byte multiple linesize, every slice-plane starting at a multiple of 4 samples and a
// multiple of 8 samples wide.
#include "ffv1_enc_kernels.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <algorithm>
#include <cstring>
#include <cstdlib>

namespace ffv1 {

constexpr int kFastThreads  = 1024;                      // one CTA per SM: two groups of 512 threads share the tables
constexpr int kFastGroup    = 512;                       // threads working on one item
constexpr int kFastBufs     = 3;                         // staging buffers per group (TMA runs two items ahead)
constexpr int kFastChunkBytes = 512;                     // bytes of a row per work item (512 / 256 samples for 8- / 16-bit sources)
constexpr int kFastRows     = kTileRows + 2;             // two rows above the tile are needed (T, and LT of x=0 / TT)
constexpr int kFastTabAB    = 256 * 256;                 // [e][A: 32 lanes x (Q1,Q2) | B: 32 lanes x (Q0,Q3)]
constexpr int kFastTabC     = 256 * 128;                 // [e][32 lanes x (Q4,-)]   (large context model only)

template <int BYTES> struct FastGeom {
    static constexpr int kChunk = kFastChunkBytes / BYTES;
    static constexpr int kRowBytes = 16 + 16 + kFastChunkBytes + 32;        // left block | misalignment | chunk | next sample + padding (multiple of 64)
    static constexpr int kBufBytes = (kFastRows + 2) * kRowBytes;           // +2: a slice's first tile lands two rows down
};

// One tiled tensor map per source plane: {linesize / 4 (32-bit elements), rows, frames}; the box is one staging buffer
// worth of rows, so a work item is fetched by a single TMA request (SASS UTMALDG) instead of one bulk copy per row.
struct FastMaps { CUtensorMap m[4]; int32_t row_bytes[4]; };    // row_bytes: box width = staged row pitch for that plane

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_row(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// prmt.b32 with the full 4-bit selectors (bit 3 of a nibble replicates the sign of the selected byte; __byte_perm masks it)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// 16-bit table read through a 32-bit shared-space address (+ compile-time offset): one LDS.S16, no address arithmetic
template <int OFF>
__device__ __forceinline__ int lds_s16(uint32_t saddr)
{
    short v;
    asm("ld.shared.s16 %0, [%1+%2];" : "=h"(v) : "r"(saddr), "n"(OFF));
    return (int)v;
}

struct FastItem {
    int f, slice, plane, y0, nrows, cx0, cw, w;
    int line_first, line_step;
    int m;                 // misalignment (bytes) of the chunk's first sample against a 16-byte boundary
    bool last_chunk;
};

template <int BYTES>
__device__ __forceinline__ bool fast_decode(const EncDeviceTables &T, long item, int items_per_frame, int maxchunks, FastItem &it)
{
    it.f = (int)(item / items_per_frame);
    const int r = (int)(item - (long)it.f * items_per_frame);
    const int ti = r / maxchunks, ch = r - ti * maxchunks;
    const TileDesc td = T.tiles[ti];
    const SliceGeom &g = T.slices[td.slice];
    it.slice = td.slice; it.plane = td.plane; it.y0 = td.y0; it.nrows = td.nrows;
    it.line_first = td.line_first; it.line_step = td.line_step;
    it.w = g.pw[td.plane];
    it.cx0 = ch * FastGeom<BYTES>::kChunk;
    it.cw = min(FastGeom<BYTES>::kChunk, it.w - it.cx0);
    it.last_chunk = it.cx0 + it.cw >= it.w;
    it.m = ((g.px0[td.plane] + it.cx0) * BYTES) & 15;
    return it.cw > 0;
}

// warp 0: start the bulk copies of one work item into `buf`
template <int BYTES>
__device__ __forceinline__ void fast_issue(const EncDeviceTables &T, const EncBatch &B, const FastItem &it, bool valid,
                                           unsigned char *buf, uint64_t *bar, int lane)
{
    uint32_t bytes = 0, nvalid = 0;
    const unsigned char *src = nullptr;
    unsigned char *dst = nullptr;
    if (valid) {
        const SliceGeom &g = T.slices[it.slice];
        const PlaneInfo &pi = T.layout.plane[it.plane];
        const int gx = (g.px0[it.plane] + it.cx0) * BYTES;            // byte column of the chunk's first sample
        const int a0 = (gx & ~15) - 16;                                 // smem offset 0 <-> this byte column
        const int a = it.cx0 > 0 ? a0 : a0 + 16;                        // first chunk: nothing left of the slice is needed
        const int e = gx + (it.cw + (it.last_chunk ? 0 : 1)) * BYTES;
        const int b = (e + 15) & ~15;
        bytes = (uint32_t)(b - a);
        const int ytop = it.y0 - 2;
        nvalid = (uint32_t)(it.nrows + 2 - (ytop < 0 ? -ytop : 0));
        src = B.planes[it.f * 4 + pi.src_plane] + a;
        dst = buf + (a - a0);
    }
    // rows above the slice are zero (ffv1enc.c:376): written here, while the buffer is idle; the release of the
    // mbarrier arrival below publishes them to the consumers
    if (valid && it.y0 < 2)
        for (int rr = 0; rr < 2 - it.y0; rr++)
            for (int i = lane * 16; i < FastGeom<BYTES>::kRowBytes; i += 32 * 16)
                *reinterpret_cast<uint4 *>(buf + rr * FastGeom<BYTES>::kRowBytes + i) = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, bytes * nvalid);
    }
    __syncwarp();
    if (valid) {
        const int ls = B.linesize[T.layout.plane[it.plane].src_plane];
        const int row0 = T.slices[it.slice].py0[it.plane] + it.y0 - 2;            // plane row staged at buffer row 0
        for (int rr = lane; rr < it.nrows + 2; rr += 32)
            if (it.y0 - 2 + rr >= 0) tma_row(dst + rr * FastGeom<BYTES>::kRowBytes, src + (ptrdiff_t)(row0 + rr) * ls, bytes, bar);
    }
}

// warp 0 of a group: one tensor-map TMA request for the whole item
template <int BYTES>
__device__ __forceinline__ void fast_issue_tensor(const EncDeviceTables &T, const FastMaps &maps, const FastItem &it, bool valid,
                                                  unsigned char *buf, uint64_t *bar, int lane)
{
    const int rowb = valid ? maps.row_bytes[T.layout.plane[it.plane].src_plane] : 0;
    const bool top = valid && it.y0 == 0;          // first tile of a slice: the two rows above it are zero, not the neighbour slice
    if (top)
        for (int i = lane * 16; i < 2 * rowb; i += 32 * 16)
            *reinterpret_cast<uint4 *>(buf + i) = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (!valid) { mbar_expect_tx(bar, 0u); return; }
        const SliceGeom &g = T.slices[it.slice];
        const int src_plane = T.layout.plane[it.plane].src_plane;
        const int gx = (g.px0[it.plane] + it.cx0) * BYTES;            // byte column of the chunk's first sample
        const int c0 = ((gx & ~15) - 16) >> 2;                          // may be negative: out-of-bounds elements read as zero
        const int c1 = g.py0[it.plane] + it.y0 - (top ? 0 : 2);
        mbar_expect_tx(bar, (uint32_t)(kFastRows * rowb));
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     :: "r"(smem_u32(buf + (top ? 2 * rowb : 0))), "l"(reinterpret_cast<uint64_t>(&maps.m[src_plane])),
                        "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(it.f) : "memory");
    }
}

template <int BYTES, int NIN, bool TENSOR>
__global__ void __launch_bounds__(kFastThreads, 1)
k_pixel_fast(const EncDeviceTables T, const EncBatch B, const int maxchunks, const __grid_constant__ FastMaps maps)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    typedef FastGeom<BYTES> G;
    const Layout &L = T.layout;
    const int tid = threadIdx.x & (kFastGroup - 1), lane = tid & 31, warp = tid >> 5;
    const int group = threadIdx.x / kFastGroup;
    constexpr int kGroups = kFastThreads / kFastGroup;
    // Shared-memory map.  The (Q1,Q2 | Q0,Q3) table sits at a shared-space address that is a multiple of 64 KB, so that
    // "(difference & 0xFF00) | lane offset | table base" is ONE LOP3 (no add): [group 0 buffers][table AB][table C]
    // [group 1 buffers][mbarriers]
    const uint32_t s0 = smem_u32(smem);
    const uint32_t tab_s = (s0 + 0xFFFFu) & ~0xFFFFu;
    if (tab_s - s0 < (uint32_t)(kFastBufs * G::kBufBytes)) __trap();         // the launch reserves 64 KB in front (see pixel_fast_smem_bytes)
    unsigned char *tabAB = smem + (tab_s - s0);
    unsigned char *tabC = tabAB + kFastTabAB;
    unsigned char *after = tabC + (NIN == 5 ? kFastTabC : 0);
    unsigned char *bufs = group == 0 ? smem : after;
    uint64_t *bars = reinterpret_cast<uint64_t *>(after + kFastBufs * G::kBufBytes) + group * 2 * kFastBufs;   // "full" barriers
    uint64_t *empty = bars + kFastBufs;                                                                        // "empty" barriers
    const uint32_t laneA = tab_s | ((uint32_t)(threadIdx.x & 31) * 4u);       // bits 8..15 are free for the table index
    const uint32_t laneC = (tab_s + kFastTabAB) | ((uint32_t)(threadIdx.x & 31) * 4u);

    // ---- lane-replicated quantisation tables
    for (int i = threadIdx.x; i < 256 * 32; i += kFastThreads) {
        const int e = i >> 5, l = i & 31;
        const uint32_t q0 = (uint16_t)T.quant[e], q1 = (uint16_t)T.quant[256 + e], q2 = (uint16_t)T.quant[512 + e];
        const uint32_t q3 = NIN == 5 ? (uint16_t)T.quant[768 + e] : 0u;
        *reinterpret_cast<uint32_t *>(tabAB + e * 256 + l * 4) = q1 | (q2 << 16);
        *reinterpret_cast<uint32_t *>(tabAB + e * 256 + 128 + l * 4) = q0 | (q3 << 16);
        if (NIN == 5) *reinterpret_cast<uint32_t *>(tabC + e * 128 + l * 4) = (uint16_t)T.quant[1024 + e];
    }
    if (tid == 0) {
        for (int i = 0; i < kFastBufs; i++) { mbar_init(&bars[i], 1); mbar_init(&empty[i], kFastGroup / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int items_per_frame = L.tiles_per_frame * maxchunks;
    const long total = (long)items_per_frame * B.nframes;
    const int bits = L.coded_bits;
    const long step = (long)gridDim.x * kGroups;

    FastItem cur, pre;
    long item = (long)blockIdx.x * kGroups + group;
    // ---- prologue: the TMA unit starts on the first kFastBufs-1 items of this group
    if (warp == 0)
        for (int j = 0; j < kFastBufs - 1; j++) {
            const long it = item + j * step;
            if (it < total) {
                const bool v = fast_decode<BYTES>(T, it, items_per_frame, maxchunks, pre);
                if (TENSOR) fast_issue_tensor<BYTES>(T, maps, pre, v, bufs + j * G::kBufBytes, &bars[j], lane);
                else        fast_issue<BYTES>(T, B, pre, v, bufs + j * G::kBufBytes, &bars[j], lane);
            }
        }
    // Producer / consumer ring without a group-wide barrier: a warp that has finished its units of an item moves on to
    // the next item (already staged) at once; the issuing warp re-uses a buffer only after all 16 warps of the group
    // have arrived on its "empty" barrier.
    int bi = 0, pi = kFastBufs - 1;           // buffer of the current item / of the item fetched now
    uint32_t phase = 0, ephase = 0;           // parity of full[bi] / of empty[pi]
    for (int k = 0; item < total; item += step, k++) {
        unsigned char *buf = bufs + bi * G::kBufBytes;
        const long pitem = item + (kFastBufs - 1) * step;
        if (warp == 0 && pitem < total) {
            if (k > 0) mbar_wait(&empty[pi], ephase);           // item k-1 (same buffer) has been consumed by every warp
            const bool v = fast_decode<BYTES>(T, pitem, items_per_frame, maxchunks, pre);
            if (TENSOR) fast_issue_tensor<BYTES>(T, maps, pre, v, bufs + pi * G::kBufBytes, &bars[pi], lane);
            else        fast_issue<BYTES>(T, B, pre, v, bufs + pi * G::kBufBytes, &bars[pi], lane);
        }
        const bool cur_valid = fast_decode<BYTES>(T, item, items_per_frame, maxchunks, cur);
        mbar_wait(&bars[bi], phase);
        if (cur_valid) {
            const SliceGeom &g = T.slices[cur.slice];
            const int o0 = 16 + cur.m;
            uint32_t *rec_slice = B.rec + (size_t)cur.f * L.rec_per_frame + g.rec_first;
            // lines of one plane follow each other in the record area, each padded to 32 records
            uint32_t *rec_tile = rec_slice + T.lines[g.line_first + cur.line_first].rec_off + cur.cx0;
            const uint32_t rec_stride = (uint32_t)(cur.w + 31) & ~31u;
            constexpr int U = 8;                                                // samples per thread and iteration
            const int rowb = TENSOR ? maps.row_bytes[L.plane[cur.plane].src_plane] : G::kRowBytes;   // pitch of the staged rows
            const int upr = (cur.cw + U - 1) / U;                               // units per row
            const uint32_t magic = (1048576u + (uint32_t)upr - 1u) / (uint32_t)upr;   // exact u / upr for u < 2^12
            const int nunits = upr * cur.nrows;
            for (int u = tid; u < nunits; u += kFastGroup) {
                const int r = (int)__umulhi((uint32_t)u << 12, magic);      // u * magic >> 20
                const int ux = u - r * upr;
                const unsigned char *crow = buf + (r + 2) * rowb + o0 + ux * U * BYTES;
                const unsigned char *trow = crow - rowb;
                // values are kept multiplied by 256: (a - b) & 0xFF00 is then directly the byte offset of table entry
                // (a - b) & 255 (ffv1.h:181-189 masks the differences with 0xFF even for deeper samples)
                int X[U], Tt[U + 2], Lx, LLx = 0, TT[U];
                const bool first = cur.cx0 == 0 && ux == 0;
                if (BYTES == 1) {
                    const uint32_t c0 = *reinterpret_cast<const uint32_t *>(crow), c1 = *reinterpret_cast<const uint32_t *>(crow + 4);
                    const uint32_t t0 = *reinterpret_cast<const uint32_t *>(trow), t1 = *reinterpret_cast<const uint32_t *>(trow + 4);
                    const uint32_t tn = *reinterpret_cast<const uint32_t *>(trow + 8);
                    X[0] = __byte_perm(c0, 0, 0x4404); X[1] = __byte_perm(c0, 0, 0x4414);
                    X[2] = __byte_perm(c0, 0, 0x4424); X[3] = __byte_perm(c0, 0, 0x4434);
                    X[4] = __byte_perm(c1, 0, 0x4404); X[5] = __byte_perm(c1, 0, 0x4414);
                    X[6] = __byte_perm(c1, 0, 0x4424); X[7] = __byte_perm(c1, 0, 0x4434);
                    Tt[1] = __byte_perm(t0, 0, 0x4404); Tt[2] = __byte_perm(t0, 0, 0x4414);
                    Tt[3] = __byte_perm(t0, 0, 0x4424); Tt[4] = __byte_perm(t0, 0, 0x4434);
                    Tt[5] = __byte_perm(t1, 0, 0x4404); Tt[6] = __byte_perm(t1, 0, 0x4414);
                    Tt[7] = __byte_perm(t1, 0, 0x4424); Tt[8] = __byte_perm(t1, 0, 0x4434);
                    Tt[9] = __byte_perm(tn, 0, 0x4404);
                    if (first) {
                        // slice-local left edge (ffv1enc.c:381-388): L = T, LT = the sample two rows up, LL = 0
                        Lx = Tt[1];
                        Tt[0] = __byte_perm(*reinterpret_cast<const uint32_t *>(trow - rowb), 0, 0x4404);
                    } else {
                        const uint32_t cp = *reinterpret_cast<const uint32_t *>(crow - 4);
                        Lx = __byte_perm(cp, 0, 0x4434);
                        Tt[0] = __byte_perm(*reinterpret_cast<const uint32_t *>(trow - 4), 0, 0x4434);
                        if (NIN == 5) LLx = __byte_perm(cp, 0, 0x4424);
                    }
                    if (NIN == 5) {
                        const uint32_t u0 = *reinterpret_cast<const uint32_t *>(trow - rowb);
                        const uint32_t u1 = *reinterpret_cast<const uint32_t *>(trow - rowb + 4);
                        TT[0] = __byte_perm(u0, 0, 0x4404); TT[1] = __byte_perm(u0, 0, 0x4414);
                        TT[2] = __byte_perm(u0, 0, 0x4424); TT[3] = __byte_perm(u0, 0, 0x4434);
                        TT[4] = __byte_perm(u1, 0, 0x4404); TT[5] = __byte_perm(u1, 0, 0x4414);
                        TT[6] = __byte_perm(u1, 0, 0x4424); TT[7] = __byte_perm(u1, 0, 0x4434);
                    }
                } else {
                    // 16-bit containers: LSB-aligned 9..15-bit values as they are; 16-bit values wrap into int16 like the
                    // reference's int16_t sample_buffer (ffv1enc.c:396-403); MSB-aligned input is shifted down first
                    const int sh = L.sample_shift;
#define S16(v) (((int)(int16_t)(((v) & 0xFFFFu) >> sh)) << 8)
                    const uint2 ca = *reinterpret_cast<const uint2 *>(crow), cb = *reinterpret_cast<const uint2 *>(crow + 8);
                    const uint2 ta = *reinterpret_cast<const uint2 *>(trow), tb = *reinterpret_cast<const uint2 *>(trow + 8);
                    const uint32_t tn = *reinterpret_cast<const uint32_t *>(trow + 16);
                    X[0] = S16(ca.x); X[1] = S16(ca.x >> 16); X[2] = S16(ca.y); X[3] = S16(ca.y >> 16);
                    X[4] = S16(cb.x); X[5] = S16(cb.x >> 16); X[6] = S16(cb.y); X[7] = S16(cb.y >> 16);
                    Tt[1] = S16(ta.x); Tt[2] = S16(ta.x >> 16); Tt[3] = S16(ta.y); Tt[4] = S16(ta.y >> 16);
                    Tt[5] = S16(tb.x); Tt[6] = S16(tb.x >> 16); Tt[7] = S16(tb.y); Tt[8] = S16(tb.y >> 16);
                    Tt[9] = S16(tn);
                    if (first) {
                        Lx = Tt[1];
                        Tt[0] = S16(*reinterpret_cast<const uint32_t *>(trow - rowb));
                    } else {
                        const uint32_t cp = *reinterpret_cast<const uint32_t *>(crow - 4);
                        Lx = S16(cp >> 16);
                        Tt[0] = S16(*reinterpret_cast<const uint32_t *>(trow - 4) >> 16);
                        if (NIN == 5) LLx = S16(cp);
                    }
                    if (NIN == 5) {
                        const uint2 ua = *reinterpret_cast<const uint2 *>(trow - rowb), ub = *reinterpret_cast<const uint2 *>(trow - rowb + 8);
                        TT[0] = S16(ua.x); TT[1] = S16(ua.x >> 16); TT[2] = S16(ua.y); TT[3] = S16(ua.y >> 16);
                        TT[4] = S16(ub.x); TT[5] = S16(ub.x >> 16); TT[6] = S16(ub.y); TT[7] = S16(ub.y >> 16);
                    }
#undef S16
                }
                // slice-local right edge: RT of the last sample = its T
                // (pixel_fast_geometry_ok guarantees that rows end with a full unit)
                if (cur.last_chunk && ux == upr - 1) Tt[U + 1] = Tt[U];
                uint32_t out[U];
                // Q1 term of sample i uses the difference LT-T = Tt[i]-Tt[i+1]; the same difference is the T-RT term
                // (Q2) of sample i-1, so one address serves both tables (A holds Q1 low, Q2 high)
                uint32_t aA = ((uint32_t)(Tt[0] - Tt[1]) & 0xFF00u) | laneA;
                int q1 = lds_s16<0>(aA);
                int Lv = Lx, LLv = LLx;
#pragma unroll
                for (int i = 0; i < U; i++) {
                    const int LT = Tt[i], Tp = Tt[i + 1], RT = Tt[i + 2];
                    aA = ((uint32_t)(Tp - RT) & 0xFF00u) | laneA;
                    const int q2 = lds_s16<2>(aA);
                    const uint32_t aB = ((uint32_t)(Lv - LT) & 0xFF00u) | laneA;
                    int ctx = lds_s16<128>(aB) + q1 + q2;
                    q1 = lds_s16<0>(aA);
                    if (NIN == 5) {
                        const uint32_t a3 = ((uint32_t)(LLv - Lv) & 0xFF00u) | laneA;
                        const uint32_t a4 = (((uint32_t)(TT[i] - Tp) >> 1) & 0x7F80u) | laneC;
                        ctx += lds_s16<130>(a3) + lds_s16<0>(a4);
                    }
                    const int pred = max(min(Lv, Tp), min(max(Lv, Tp), Lv + Tp - LT));     // median (mathops.h:95-119)
                    int diff = X[i] - pred;                                                // residual * 256
                    const int sg = (ctx >> 31) | 1;                                        // context < 0: negate both
                    ctx *= sg; diff *= sg;
                    if (BYTES == 1) {
                        // fold() to int8 and pack: byte0 = diff, byte1 = its sign, bytes 2..3 = context
                        out[i] = prmt((uint32_t)diff, (uint32_t)ctx, 0x5491u);
                    } else {
                        const int d = (diff << (24 - bits)) >> (32 - bits);                // fold() to `bits` bits
                        out[i] = __byte_perm((uint32_t)d, (uint32_t)ctx, 0x5410);
                    }
                    LLv = Lv; Lv = X[i];
                }
                // one 256-bit store per thread (a full 32-byte sector; record lines are padded to 32 records, so the
                // last, possibly partial, unit may store all eight)
                char *dst = reinterpret_cast<char *>(rec_tile) + (size_t)((uint32_t)r * rec_stride + (uint32_t)ux * U) * 4u;
                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                             :: "l"(dst), "r"(out[0]), "r"(out[1]), "r"(out[2]), "r"(out[3]), "r"(out[4]), "r"(out[5]), "r"(out[6]), "r"(out[7])
                             : "memory");
            }
        }
        // this warp is done with the buffer
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[bi]);
        if (k > 0 && pi == kFastBufs - 1) ephase ^= 1u;        // empty[] parity flips once per trip round the ring
        pi = bi;
        if (++bi == kFastBufs) { bi = 0; phase ^= 1u; }
    }
}

int pixel_fast_smem_bytes(const Layout &L)
{
    const int bytes = L.src_kind == SRC_PLANAR16 ? 2 : 1;
    const int buf = bytes == 2 ? FastGeom<2>::kBufBytes : FastGeom<1>::kBufBytes;
    // 64 KB in front of the 64 KB-aligned table (holds group 0's buffers) + tables + group 1's buffers + mbarriers
    return 65536 + kFastTabAB + (L.ctx_inputs == 5 ? kFastTabC : 0) + kFastBufs * buf + 256;
}

// static part of the eligibility test (geometry); pointer / linesize alignment is checked per call by the host
bool pixel_fast_geometry_ok(const Layout &L, const SliceGeom *slices, int nslices)
{
    if (L.rgb || (L.src_kind != SRC_PLANAR8 && L.src_kind != SRC_PLANAR16)) return false;
    for (int p = 0; p < L.nplanes; p++)
        if (L.plane[p].pstep != (L.src_kind == SRC_PLANAR16 ? 2 : 1)) return false;          // ya8 interleaves two planes
    for (int s = 0; s < nslices; s++)
        for (int p = 0; p < L.nplanes; p++)
            if ((slices[s].px0[p] & 3) || (slices[s].pw[p] & 7)) return false;      // 4-byte aligned starts, rows of whole 8-sample units
    return true;
}

template <int BYTES, int NIN>
static cudaError_t set_attr(int sm)
{
    cudaError_t e = cudaFuncSetAttribute(k_pixel_fast<BYTES, NIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_pixel_fast<BYTES, NIN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
}

cudaError_t configure_pixel_fast(const Layout &L)
{
    const int sm = pixel_fast_smem_bytes(L);
    cudaError_t e;
    if ((e = set_attr<1, 3>(sm)) != cudaSuccess) return e;
    if ((e = set_attr<1, 5>(sm)) != cudaSuccess) return e;
    if ((e = set_attr<2, 3>(sm)) != cudaSuccess) return e;
    return set_attr<2, 5>(sm);
}

static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder()
{
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
        cudaGetLastError();
    }
    return fn;
}

template <int BYTES, int NIN>
static void launch_t(const EncDeviceTables &t, const EncBatch &b, int maxchunks, int grid, int sm, cudaStream_t s,
                     const uint8_t *const *frame0_planes, long long frame_stride, const SliceGeom *slices, int nslices)
{
    FastMaps maps;
    memset(&maps, 0, sizeof(maps));
    bool tensor = frame_stride >= 0 && (frame_stride & 15) == 0 && getenv("FFV1B200_PIXEL_ROWCOPY") == nullptr;
    PFN_cuTensorMapEncodeTiled_v12000 enc = tensor ? tensor_map_encoder() : nullptr;
    tensor = tensor && enc != nullptr;
    const Layout &L = t.layout;
    for (int p = 0; p < L.nplanes && tensor; p++) {
        const int sp = L.plane[p].src_plane;
        const int rows = (L.height + (1 << L.plane[p].vshift) - 1) >> L.plane[p].vshift;
        const cuuint64_t dims[3] = {(cuuint64_t)(b.linesize[sp] / 4), (cuuint64_t)rows, (cuuint64_t)b.nframes};
        const cuuint64_t strides[2] = {(cuuint64_t)b.linesize[sp], (cuuint64_t)(b.nframes > 1 ? frame_stride : b.linesize[sp] * (long long)rows)};
        // box width: left block + misalignment + the widest chunk of this plane + next sample, as a multiple of 64 bytes
        // (so that two rows are a multiple of the 128-byte alignment TMA wants for its shared-memory destination)
        int wmax = 0;
        for (int si = 0; si < nslices; si++) wmax = std::max(wmax, slices[si].pw[p]);
        const int rowb = std::min(FastGeom<BYTES>::kRowBytes, (16 + 15 + std::min(wmax, FastGeom<BYTES>::kChunk) * BYTES + BYTES + 63) & ~63);
        maps.row_bytes[sp] = std::max(maps.row_bytes[sp], rowb);
        const cuuint32_t box[3] = {(cuuint32_t)(maps.row_bytes[sp] / 4), (cuuint32_t)kFastRows, 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        if (strides[1] & 15) { tensor = false; break; }
        CUresult r = enc(&maps.m[sp], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t *>(frame0_planes[sp]), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) tensor = false;
    }
    if (tensor) k_pixel_fast<BYTES, NIN, true><<<grid, kFastThreads, sm, s>>>(t, b, maxchunks, maps);
    else        k_pixel_fast<BYTES, NIN, false><<<grid, kFastThreads, sm, s>>>(t, b, maxchunks, maps);
}

// frame0_planes: host copy of the first frame's plane pointers; frame_stride: bytes between the same plane of consecutive
// frames when that distance is constant over the batch (tensor-map TMA), else -1 (one bulk copy per row)
void launch_pixel_fast(const EncDeviceTables &t, const EncBatch &b, int max_plane_width, int num_sms, cudaStream_t s,
                       const uint8_t *const *frame0_planes, long long frame_stride, const SliceGeom *slices, int nslices)
{
    const Layout &L = t.layout;
    const int chunk = kFastChunkBytes / (L.src_kind == SRC_PLANAR16 ? 2 : 1);
    const int maxchunks = (max_plane_width + chunk - 1) / chunk;
    const long total = (long)L.tiles_per_frame * maxchunks * b.nframes;
    const int grid = (int)std::min<long>((total + 1) / 2, (long)num_sms);
    const int sm = pixel_fast_smem_bytes(L);
    const bool five = L.ctx_inputs == 5;
    if (L.src_kind == SRC_PLANAR8) {
        if (five) launch_t<1, 5>(t, b, maxchunks, grid, sm, s, frame0_planes, frame_stride, slices, nslices);
        else      launch_t<1, 3>(t, b, maxchunks, grid, sm, s, frame0_planes, frame_stride, slices, nslices);
    } else {
        if (five) launch_t<2, 5>(t, b, maxchunks, grid, sm, s, frame0_planes, frame_stride, slices, nslices);
        else      launch_t<2, 3>(t, b, maxchunks, grid, sm, s, frame0_planes, frame_stride, slices, nslices);
    }
}

} // namespace ffv1
