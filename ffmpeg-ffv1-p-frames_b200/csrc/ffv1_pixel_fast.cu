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
//   * each thread handles 4 consecutive samples held in registers and writes one 16-byte record vector.
// Requirements checked on the host (else the generic kernel runs): 8- or 16-bit planar source, every source plane
// 16-byte aligned with a 16-byte multiple linesize, every slice-plane starting at a multiple of 4 samples.
#include "ffv1_enc_kernels.cuh"
#include <algorithm>

namespace ffv1 {

constexpr int kFastThreads  = 256;
constexpr int kFastChunk    = 512;                       // samples of a row per work item
constexpr int kFastRows     = kTileRows + 2;             // two rows above the tile are needed (T, and LT of x=0 / TT)
constexpr int kFastTabAB    = 256 * 256;                 // [e][A: 32 lanes x (Q1,Q2) | B: 32 lanes x (Q0,Q3)]
constexpr int kFastTabC     = 256 * 128;                 // [e][32 lanes x (Q4,-)]   (large context model only)

template <int BYTES> struct FastGeom {
    static constexpr int kRowBytes = 16 + 16 + kFastChunk * BYTES + 16;     // left block | misalignment | chunk | one more sample + padding
    static constexpr int kBufBytes = kFastRows * kRowBytes;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
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
    it.cx0 = ch * kFastChunk;
    it.cw = min(kFastChunk, it.w - it.cx0);
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
        const int y = ytop + lane;
        if (lane < it.nrows + 2 && y >= 0) {
            src = B.planes[it.f * 4 + pi.src_plane] + (size_t)(g.py0[it.plane] + y) * B.linesize[pi.src_plane] + a;
            dst = buf + lane * FastGeom<BYTES>::kRowBytes + (a - a0);
        }
    }
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, bytes * nvalid);
    }
    __syncwarp();
    if (src) tma_row(dst, src, bytes, bar);
}

template <int BYTES, int NIN>
__global__ void __launch_bounds__(kFastThreads, 2)
k_pixel_fast(const EncDeviceTables T, const EncBatch B, const int maxchunks)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    typedef FastGeom<BYTES> G;
    const Layout &L = T.layout;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char *tabAB = smem;
    unsigned char *tabC = smem + kFastTabAB;
    unsigned char *bufs = smem + kFastTabAB + (NIN == 5 ? kFastTabC : 0);
    uint64_t *bars = reinterpret_cast<uint64_t *>(bufs + 2 * G::kBufBytes);

    // ---- lane-replicated quantisation tables
    for (int i = tid; i < 256 * 32; i += kFastThreads) {
        const int e = i >> 5, l = i & 31;
        const uint32_t q0 = (uint16_t)T.quant[e], q1 = (uint16_t)T.quant[256 + e], q2 = (uint16_t)T.quant[512 + e];
        const uint32_t q3 = NIN == 5 ? (uint16_t)T.quant[768 + e] : 0u;
        *reinterpret_cast<uint32_t *>(tabAB + e * 256 + l * 4) = q1 | (q2 << 16);
        *reinterpret_cast<uint32_t *>(tabAB + e * 256 + 128 + l * 4) = q0 | (q3 << 16);
        if (NIN == 5) *reinterpret_cast<uint32_t *>(tabC + e * 128 + l * 4) = (uint16_t)T.quant[1024 + e];
    }
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int items_per_frame = L.tiles_per_frame * maxchunks;
    const long total = (long)items_per_frame * B.nframes;
    const int bits = L.coded_bits;
    const uint32_t lane4 = (uint32_t)lane * 4u;

    FastItem cur, nxt;
    bool cur_valid = false, nxt_valid = false;
    long item = blockIdx.x;
    if (item < total) {
        cur_valid = fast_decode<BYTES>(T, item, items_per_frame, maxchunks, cur);
        if (warp == 0) fast_issue<BYTES>(T, B, cur, cur_valid, bufs, &bars[0], lane);
    }
    for (int k = 0; item < total; item += gridDim.x, k++) {
        unsigned char *buf = bufs + (k & 1) * G::kBufBytes;
        const long nitem = item + gridDim.x;
        if (nitem < total) {
            nxt_valid = fast_decode<BYTES>(T, nitem, items_per_frame, maxchunks, nxt);
            if (warp == 0) fast_issue<BYTES>(T, B, nxt, nxt_valid, bufs + ((k + 1) & 1) * G::kBufBytes, &bars[(k + 1) & 1], lane);
        }
        mbar_wait(&bars[k & 1], (uint32_t)(k >> 1) & 1u);
        if (cur_valid) {
            const int o0 = 16 + cur.m;                     // smem offset of the chunk's first sample inside a staged row
            // ---- slice-local edge rules (ffv1enc.c:381-388), written into the staged rows
            if (tid < cur.nrows + 2) {
                unsigned char *row = buf + tid * G::kRowBytes;
                const int y = cur.y0 - 2 + tid;
                if (y < 0) {
                    for (int i = 0; i < G::kRowBytes; i += 16) *reinterpret_cast<uint4 *>(row + i) = make_uint4(0, 0, 0, 0);
                } else {
                    if (cur.cx0 == 0) {
                        // sample[-1] = the sample above x=0 (0 on the first row); sample[-2] = 0
                        if (BYTES == 1) {
                            row[o0 - 1] = (tid > 0 && y > 0) ? (row - G::kRowBytes)[o0] : 0;
                            row[o0 - 2] = 0;
                        } else {
                            reinterpret_cast<uint16_t *>(row + o0)[-1] = (tid > 0 && y > 0) ? *reinterpret_cast<uint16_t *>(row - G::kRowBytes + o0) : 0;
                            reinterpret_cast<uint16_t *>(row + o0)[-2] = 0;
                        }
                    }
                    if (cur.last_chunk) {
                        if (BYTES == 1) row[o0 + cur.cw] = row[o0 + cur.cw - 1];
                        else reinterpret_cast<uint16_t *>(row + o0)[cur.cw] = reinterpret_cast<uint16_t *>(row + o0)[cur.cw - 1];
                    }
                }
            }
        }
        __syncthreads();
        if (cur_valid) {
            const SliceGeom &g = T.slices[cur.slice];
            const int o0 = 16 + cur.m;
            uint32_t *rec_slice = B.rec + (size_t)cur.f * L.rec_per_frame + g.rec_first;
            // lines of one plane follow each other in the record area, each padded to 32 records
            uint32_t *rec_tile = rec_slice + T.lines[g.line_first + cur.line_first].rec_off + cur.cx0;
            const uint32_t rec_stride = (uint32_t)(cur.w + 31) & ~31u;
            const int upr = (cur.cw + 3) >> 2;                                  // 4-sample units per row
            const uint32_t magic = (1048576u + (uint32_t)upr - 1u) / (uint32_t)upr;   // exact u / upr for u < 2^11
            const int nunits = upr * cur.nrows;
            for (int u = tid; u < nunits; u += kFastThreads) {
                const int r = (int)(((uint32_t)u * magic) >> 20);
                const int ux = u - r * upr;
                const unsigned char *crow = buf + (r + 2) * G::kRowBytes + o0 + ux * 4 * BYTES;
                const unsigned char *trow = crow - G::kRowBytes;
                // values are kept multiplied by 256: (a - b) & 0xFF00 is then directly the byte offset of table entry
                // (a - b) & 255 (ffv1.h:181-189 masks the differences with 0xFF even for deeper samples)
                int X[4], Tt[6], Lx, LLx = 0, TT[4] = {0, 0, 0, 0};
                if (BYTES == 1) {
                    const uint32_t c4 = *reinterpret_cast<const uint32_t *>(crow);
                    const uint32_t cp = *reinterpret_cast<const uint32_t *>(crow - 4);
                    const uint32_t t4 = *reinterpret_cast<const uint32_t *>(trow);
                    const uint32_t tp = *reinterpret_cast<const uint32_t *>(trow - 4);
                    const uint32_t tn = *reinterpret_cast<const uint32_t *>(trow + 4);
                    X[0] = __byte_perm(c4, 0, 0x4404); X[1] = __byte_perm(c4, 0, 0x4414);
                    X[2] = __byte_perm(c4, 0, 0x4424); X[3] = __byte_perm(c4, 0, 0x4434);
                    Lx = __byte_perm(cp, 0, 0x4434);
                    Tt[0] = __byte_perm(tp, 0, 0x4434);
                    Tt[1] = __byte_perm(t4, 0, 0x4404); Tt[2] = __byte_perm(t4, 0, 0x4414);
                    Tt[3] = __byte_perm(t4, 0, 0x4424); Tt[4] = __byte_perm(t4, 0, 0x4434);
                    Tt[5] = __byte_perm(tn, 0, 0x4404);
                    if (NIN == 5) {
                        LLx = __byte_perm(cp, 0, 0x4424);
                        const uint32_t u4 = *reinterpret_cast<const uint32_t *>(trow - G::kRowBytes);
                        TT[0] = __byte_perm(u4, 0, 0x4404); TT[1] = __byte_perm(u4, 0, 0x4414);
                        TT[2] = __byte_perm(u4, 0, 0x4424); TT[3] = __byte_perm(u4, 0, 0x4434);
                    }
                } else {
                    // 16-bit containers: LSB-aligned 9..15-bit values as they are; 16-bit values wrap into int16 like the
                    // reference's int16_t sample_buffer (ffv1enc.c:396-403); MSB-aligned input is shifted down first
                    const int sh = L.sample_shift;
                    const uint2 c4 = *reinterpret_cast<const uint2 *>(crow);
                    const uint32_t cp = *reinterpret_cast<const uint32_t *>(crow - 4);
                    const uint2 t4 = *reinterpret_cast<const uint2 *>(trow);
                    const uint32_t tp = *reinterpret_cast<const uint32_t *>(trow - 4);
                    const uint32_t tn = *reinterpret_cast<const uint32_t *>(trow + 8);
#define S16(v) (((int)(int16_t)(((v) & 0xFFFFu) >> sh)) << 8)
                    X[0] = S16(c4.x); X[1] = S16(c4.x >> 16); X[2] = S16(c4.y); X[3] = S16(c4.y >> 16);
                    Lx = S16(cp >> 16);
                    Tt[0] = S16(tp >> 16);
                    Tt[1] = S16(t4.x); Tt[2] = S16(t4.x >> 16); Tt[3] = S16(t4.y); Tt[4] = S16(t4.y >> 16);
                    Tt[5] = S16(tn);
                    if (NIN == 5) {
                        LLx = S16(cp);
                        const uint2 u4 = *reinterpret_cast<const uint2 *>(trow - G::kRowBytes);
                        TT[0] = S16(u4.x); TT[1] = S16(u4.x >> 16); TT[2] = S16(u4.y); TT[3] = S16(u4.y >> 16);
                    }
#undef S16
                }
                uint32_t out[4];
                // Q1 term of sample i uses the difference LT-T = Tt[i]-Tt[i+1]; the same difference is the T-RT term
                // (Q2) of sample i-1, so one address serves both tables (A holds Q1 low, Q2 high)
                uint32_t aA = ((uint32_t)(Tt[0] - Tt[1]) & 0xFF00u) | lane4;
                int q1 = *reinterpret_cast<const int16_t *>(tabAB + aA);
                int Lv = Lx, LLv = LLx;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int LT = Tt[i], Tp = Tt[i + 1], RT = Tt[i + 2];
                    aA = ((uint32_t)(Tp - RT) & 0xFF00u) | lane4;
                    const int q2 = *reinterpret_cast<const int16_t *>(tabAB + aA + 2);
                    const uint32_t aB = ((uint32_t)(Lv - LT) & 0xFF00u) | lane4;
                    int ctx = *reinterpret_cast<const int16_t *>(tabAB + aB + 128) + q1 + q2;
                    q1 = *reinterpret_cast<const int16_t *>(tabAB + aA);
                    if (NIN == 5) {
                        const uint32_t a3 = ((uint32_t)(LLv - Lv) & 0xFF00u) | lane4;
                        const uint32_t a4 = (((uint32_t)(TT[i] - Tp) >> 1) & 0x7F80u) | lane4;
                        ctx += *reinterpret_cast<const int16_t *>(tabAB + a3 + 130) + *reinterpret_cast<const int16_t *>(tabC + a4);
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
                uint32_t *dst = rec_tile + ((uint32_t)r * rec_stride + (uint32_t)ux * 4u);
                if (ux * 4 + 3 < cur.cw) {
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(out[0], out[1], out[2], out[3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        if (ux * 4 + i < cur.cw) dst[i] = out[i];
                }
            }
        }
        __syncthreads();
        cur = nxt; cur_valid = nxt_valid;
    }
}

int pixel_fast_smem_bytes(const Layout &L)
{
    const int bytes = L.src_kind == SRC_PLANAR16 ? 2 : 1;
    const int buf = bytes == 2 ? FastGeom<2>::kBufBytes : FastGeom<1>::kBufBytes;
    return kFastTabAB + (L.ctx_inputs == 5 ? kFastTabC : 0) + 2 * buf + 16;
}

// static part of the eligibility test (geometry); pointer / linesize alignment is checked per call by the host
bool pixel_fast_geometry_ok(const Layout &L, const SliceGeom *slices, int nslices)
{
    if (L.rgb || (L.src_kind != SRC_PLANAR8 && L.src_kind != SRC_PLANAR16)) return false;
    for (int p = 0; p < L.nplanes; p++)
        if (L.plane[p].pstep != (L.src_kind == SRC_PLANAR16 ? 2 : 1)) return false;          // ya8 interleaves two planes
    for (int s = 0; s < nslices; s++)
        for (int p = 0; p < L.nplanes; p++)
            if (slices[s].px0[p] & 3) return false;
    return true;
}

cudaError_t configure_pixel_fast(const Layout &L)
{
    const int sm = pixel_fast_smem_bytes(L);
    cudaError_t e;
    e = cudaFuncSetAttribute(k_pixel_fast<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm); if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_pixel_fast<1, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm); if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_pixel_fast<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm); if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_pixel_fast<2, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm); if (e != cudaSuccess) return e;
    return cudaSuccess;
}

void launch_pixel_fast(const EncDeviceTables &t, const EncBatch &b, int max_plane_width, int num_sms, cudaStream_t s)
{
    const Layout &L = t.layout;
    const int maxchunks = (max_plane_width + kFastChunk - 1) / kFastChunk;
    const long total = (long)L.tiles_per_frame * maxchunks * b.nframes;
    const int grid = (int)std::min<long>(total, 2L * num_sms);
    const int sm = pixel_fast_smem_bytes(L);
    const bool five = L.ctx_inputs == 5;
    if (L.src_kind == SRC_PLANAR8) {
        if (five) k_pixel_fast<1, 5><<<grid, kFastThreads, sm, s>>>(t, b, maxchunks);
        else      k_pixel_fast<1, 3><<<grid, kFastThreads, sm, s>>>(t, b, maxchunks);
    } else {
        if (five) k_pixel_fast<2, 5><<<grid, kFastThreads, sm, s>>>(t, b, maxchunks);
        else      k_pixel_fast<2, 3><<<grid, kFastThreads, sm, s>>>(t, b, maxchunks);
    }
}

} // namespace ffv1
