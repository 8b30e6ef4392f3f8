// ffv1_pixel_fast.cu -- the per-pixel pass for planar formats, tuned for the HBM roofline (sm_100a).
//
// Same function as k_pixel (ffv1_enc_kernels.cu): sample fetch, slice-local neighbourhood, get_context (ffv1.h:170-190),
// median predictor (ffv1.h:161-168), residual, sign flip, fold (ffv1.h:148-159) -> (context<<16 | diff) records.
// Differences that make it fast:
//   * one persistent CTA per SM = two groups of 16 warps.  A group walks its own sequence of work items (32 rows x
//     <= 512 bytes of one slice-plane, described by a host-built table: no division or dependent table walk per item);
//     warp 15 of the group is a PRODUCER that only issues TMA requests (one tensor-map request per item, up to five
//     items ahead), the other 15 warps are CONSUMERS.  Full/empty mbarriers per staging buffer, no group-wide barrier:
//     no thread ever issues a global load, and a consumer warp that is done with an item moves on at once;
//   * the 8-sample units of an item are dealt to the consumer threads round-robin ACROSS items, so the partial last
//     round of every item lands on different warps and all warps get the same share;
//   * the quantisation tables live in shared memory replicated once per lane (entry e of lane l at e*256 + l*4), so the
//     data-dependent lookups of a warp never collide in a bank; the table sits at a 64 KB aligned shared address, so
//     "difference byte -> bits 8..15 of the lane's table base" is a single PRMT;
//   * 8-bit sources with the small context model run on 16x2 SIMD integer instructions (two samples per register);
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

constexpr int kFastThreads  = 1024;                      // one CTA per SM: two groups of 16 warps share the tables
constexpr int kFastGroup    = 512;                       // threads of a group: 15 consumer warps + 1 producer warp (TMA issue only)
constexpr int kFastComp     = kFastGroup - 32;           // consumer threads of a group
constexpr int kFastMaxBufs  = 6;                         // staging buffers per group (the producer runs up to 5 items ahead)
constexpr int kFastChunkBytes = 512;                     // bytes of a row per work item (512 / 256 samples for 8- / 16-bit sources)
constexpr int kFastRows     = kTileRows + 2;             // two rows above the tile are needed (T, and LT of x=0 / TT)
constexpr int kFastBufRows  = kFastRows + 2;             // a slice's first tile lands two rows down (zero rows above it)
constexpr int kFastTabAB    = 256 * 256;                 // [e][A: 32 lanes x (Q1,Q2) | B: 32 lanes x (Q0,Q3)]
constexpr int kFastTabC     = 256 * 128;                 // [e][32 lanes x (Q4,-)]   (large context model only)
constexpr int kFastMaxSmem  = 227 * 1024;

// One tiled tensor map per source plane: {linesize / 4 (32-bit elements), rows, frames}; the box is one staging buffer
// worth of rows, so a work item is fetched by a single TMA request (SASS UTMALDG) instead of one bulk copy per row.
struct FastMaps { CUtensorMap m[4]; };

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

// ---- generic unit: 8 samples of one row, any depth, 3- or 5-input context (one sample per register)
template <int BYTES, int NIN>
__device__ __forceinline__ void fast_unit_generic(const unsigned char *crow, const int rowb, const bool first, const bool last,
                                                  const uint32_t laneA, const uint32_t laneC, const int sample_shift, const int bits,
                                                  uint32_t (&out)[8])
{
    constexpr int U = 8;
    const unsigned char *trow = crow - rowb;
    // values are kept multiplied by 256: (a - b) & 0xFF00 is then directly the byte offset of table entry
    // (a - b) & 255 (ffv1.h:181-189 masks the differences with 0xFF even for deeper samples)
    int X[U], Tt[U + 2], Lx, LLx = 0, TT[U];
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
        const int sh = sample_shift;
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
    if (last) Tt[U + 1] = Tt[U];
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
}

// ---- 8-bit sources, 3-input context: two samples per 32-bit register (16x2 SIMD: VIMNMX3.U16x2, IADD3 with a per-half
// bias so that no borrow crosses the halves).  With j = i-1, x = current row, t = row above (slice-local edge rules
// already substituted into the raw words):
//     context(i)    = Q0[V(j)] + Q1[H(j)] + Q2[H(j+1)],   V(j) = x[j]-t[j],  H(j) = t[j]-t[j+1]        (mod 256)
//     x[i]-pred(i)  = x[i] + t[j] - min3(x[j], t[j+1], t[j]) - max3(x[j], t[j+1], t[j])
// because median(L, T, L+T-LT) = L+T-median(L, T, LT) and median(a,b,c) = a+b+c-min3-max3 (ffv1.h:161-190).
__device__ __forceinline__ uint32_t ld_s32(const unsigned char *p) { return *reinterpret_cast<const uint32_t *>(p); }

__device__ __forceinline__ void fast_unit8_packed(const unsigned char *crow, const int rowb, const bool first, const bool last,
                                                  const uint32_t laneA, uint32_t (&out)[8])
{
    const unsigned char *trow = crow - rowb;
    uint32_t cp = ld_s32(crow - 4);
    const uint32_t c0 = ld_s32(crow), c1 = ld_s32(crow + 4);
    uint32_t tp = ld_s32(trow - 4);
    const uint32_t t0 = ld_s32(trow), t1 = ld_s32(trow + 4);
    uint32_t tn = ld_s32(trow + 8);
    // slice-local edges (ffv1enc.c:381-388): left of x=0: L = T, LT = the sample two rows up; RT of the last = T
    if (first) { cp = t0 << 24; tp = ld_s32(trow - rowb) << 24; }
    if (last) tn = t1 >> 24;
    // P*[k+1] = (s[2k], s[2k+1]), P*[0] ends with s[-1];  O*[k] = (s[2k-1], s[2k])
    uint32_t PX[5], PT[6], OX[4], OT[5];
    PX[0] = prmt(cp, 0u, 0x4342u);
    PX[1] = prmt(c0, 0u, 0x4140u); PX[2] = prmt(c0, 0u, 0x4342u); PX[3] = prmt(c1, 0u, 0x4140u); PX[4] = prmt(c1, 0u, 0x4342u);
    PT[0] = prmt(tp, 0u, 0x4342u);
    PT[1] = prmt(t0, 0u, 0x4140u); PT[2] = prmt(t0, 0u, 0x4342u); PT[3] = prmt(t1, 0u, 0x4140u); PT[4] = prmt(t1, 0u, 0x4342u);
    PT[5] = tn & 0xFFu;
#pragma unroll
    for (int k = 0; k < 4; k++) OX[k] = __funnelshift_r(PX[k], PX[k + 1], 16);
#pragma unroll
    for (int k = 0; k < 5; k++) OT[k] = __funnelshift_r(PT[k], PT[k + 1], 16);
    uint32_t Vw[4], Hw[5], MG[4];
#pragma unroll
    for (int k = 0; k < 5; k++) Hw[k] = OT[k] + 0x01000100u - PT[k + 1];          // bytes 0 / 2: H(j) for j = 2k-1, 2k
#pragma unroll
    for (int k = 0; k < 4; k++) {
        Vw[k] = OX[k] + 0x01000100u - OT[k];                                      // bytes 0 / 2: V(j) for j = 2k-1, 2k
        const uint32_t mn = __vimin3_u16x2(OX[k], PT[k + 1], OT[k]), mx = __vimax3_u16x2(OX[k], PT[k + 1], OT[k]);
        const uint32_t D = PX[k + 1] + OT[k] + 0x01000100u - mn - mx;              // halves: residual + 256 of samples 2k, 2k+1
        const uint32_t ND = 0x02000200u - D;                                       // halves: 256 - residual
        MG[k] = prmt(D, ND, 0x6240u);                                              // bytes: d(2k), -d(2k), d(2k+1), -d(2k+1)
    }
    // table addresses: the difference byte goes straight into bits 8..15 of the lane's table base (one PRMT)
    uint32_t aH = prmt(Hw[0], laneA, 0x7604u);
    int q1 = lds_s16<0>(aH);
#pragma unroll
    for (int n = 0; n < 8; n++) {
        aH = prmt(Hw[(n + 1) >> 1], laneA, ((n + 1) & 1) ? 0x7624u : 0x7604u);
        const uint32_t aV = prmt(Vw[n >> 1], laneA, (n & 1) ? 0x7624u : 0x7604u);
        const int ctx = lds_s16<128>(aV) + q1 + lds_s16<2>(aH);
        if (n < 7) q1 = lds_s16<0>(aH);
        // context < 0: negate context and residual (ffv1enc.c:311-316); fold() to int8 and sign-extend to 16 bits in the PRMT
        // the selector is built on the FMA pipe (IMAD.HI + IMAD): the ALU pipe is what bounds this kernel
        const int sgn = __mulhi(ctx, 1);                                           // -1 / 0
        const uint32_t sel = (uint32_t)(sgn * -0x11 + ((n & 1) ? 0x54A2 : 0x5480));
        out[n] = prmt(MG[n >> 1], (uint32_t)abs(ctx), sel);
    }
}


struct FastParams {
    const FastItemDesc *items;     // [items_per_frame]
    int32_t items_per_frame;
    int32_t nbuf;                  // staging buffers per group
    int32_t buf_bytes;             // bytes per staging buffer (multiple of 128)
};

// producer warp: one tensor-map TMA request for the whole item
__device__ __forceinline__ void fast_issue_tensor(const FastMaps &maps, const FastItemDesc &d, int f, unsigned char *buf, uint64_t *bar, int lane)
{
    const bool top = d.flags & 1;                  // first tile of a slice: the two rows above it are zero, not the neighbour slice
    const int rowb = d.rowb;
    if (top)
        for (int i = lane * 16; i < 2 * rowb; i += 32 * 16)
            *reinterpret_cast<uint4 *>(buf + i) = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, (uint32_t)(kFastRows * rowb));
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     :: "r"(smem_u32(buf + (top ? 2 * rowb : 0))), "l"(reinterpret_cast<uint64_t>(&maps.m[d.src_plane])),
                        "r"(smem_u32(bar)), "r"(d.c0), "r"(d.c1), "r"(f) : "memory");
    }
}

// producer warp, frames that are not equally spaced in memory: one bulk copy per row
template <int BYTES>
__device__ __forceinline__ void fast_issue_rows(const EncBatch &B, const FastItemDesc &d, int f, unsigned char *buf, uint64_t *bar, int lane)
{
    const bool top = d.flags & 1;
    const int rowb = d.rowb;
    const int a0 = d.c0 * 4;                                          // byte column staged at buffer column 0
    const int a = (d.flags & 2) ? a0 + 16 : a0;                       // first chunk: nothing left of the slice is needed
    const int e = a0 + d.o0 + (d.upr * 8 + ((d.flags & 4) ? 0 : 1)) * BYTES;
    const uint32_t bytes = (uint32_t)(((e + 15) & ~15) - a);
    const int skip = top ? 2 : 0;                                     // buffer rows that are not fetched
    if (top)
        for (int i = lane * 16; i < 2 * rowb; i += 32 * 16)
            *reinterpret_cast<uint4 *>(buf + i) = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, bytes * (uint32_t)(d.nrows + 2 - skip));
    }
    __syncwarp();
    const int ls = B.linesize[d.src_plane];
    const unsigned char *src = B.planes[f * 4 + d.src_plane] + a;
    for (int rr = skip + lane; rr < d.nrows + 2; rr += 32)            // d.c1 = plane row staged at buffer row `skip`
        tma_row(buf + rr * rowb + (a - a0), src + (ptrdiff_t)(d.c1 + rr - skip) * ls, bytes, bar);
}

template <int BYTES, int NIN, bool TENSOR>
__global__ void __launch_bounds__(kFastThreads, 1)
k_pixel_fast(const EncDeviceTables T, const EncBatch B, const FastParams P, const __grid_constant__ FastMaps maps)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const Layout &L = T.layout;
    const int tid = threadIdx.x & (kFastGroup - 1), lane = tid & 31, warp = tid >> 5;
    const int group = threadIdx.x / kFastGroup;
    constexpr int kGroups = kFastThreads / kFastGroup;
    // Shared-memory map: [front buffers][table AB, 64 KB aligned][table C][back buffers][mbarriers].  The table sits at a
    // shared-space address that is a multiple of 64 KB, so that "difference byte | lane offset | table base" is ONE PRMT.
    const uint32_t s0 = smem_u32(smem);
    uint32_t dyn_bytes;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
    const uint32_t tab_s = (s0 + 0xFFFFu) & ~0xFFFFu;
    unsigned char *tabAB = smem + (tab_s - s0);
    unsigned char *tabC = tabAB + kFastTabAB;
    unsigned char *back = tabC + (NIN == 5 ? kFastTabC : 0);
    const int nbuf = P.nbuf, buf_bytes = P.buf_bytes;
    const int n_front = min((int)((tab_s - s0) / (uint32_t)buf_bytes), kGroups * nbuf);
    uint64_t *bars = reinterpret_cast<uint64_t *>(back + (kGroups * nbuf - n_front) * buf_bytes) + group * 2 * kFastMaxBufs;   // "full"
    uint64_t *empty = bars + kFastMaxBufs;                                                                                       // "empty"
    if ((uint32_t)(back - smem) + (uint32_t)((kGroups * nbuf - n_front) * buf_bytes) + kGroups * 2 * kFastMaxBufs * 8u > dyn_bytes) __trap();
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
        for (int i = 0; i < nbuf; i++) { mbar_init(&bars[i], 1); mbar_init(&empty[i], kFastComp / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- the group's item sequence: gi, gi + step, ...  as (frame, item of the frame), advanced without divisions
    const int ipf = P.items_per_frame;
    const int step = (int)gridDim.x * kGroups;
    const int step_f = step / ipf, step_r = step - step_f * ipf;
    const int gi = (int)blockIdx.x * kGroups + group;
    int f = gi / ipf, r = gi - f * ipf;
    const int nframes = B.nframes;
    const int bits = L.coded_bits;
    int slot = 0;
    uint32_t phase = 0;
    auto buf_of = [&](int sl) -> unsigned char * {
        const int b = group * nbuf + sl;
        return b < n_front ? smem + b * buf_bytes : back + (b - n_front) * buf_bytes;
    };

    if (warp == kFastComp / 32) {
        // ================= producer warp: keeps up to nbuf items of the group in flight
        bool wrapped = false;
        for (; f < nframes; ) {
            if (wrapped) mbar_wait(&empty[slot], phase);            // every consumer warp is done with the item that used this buffer
            const uint4 *dp = reinterpret_cast<const uint4 *>(P.items + r);
            const uint4 d0 = __ldg(dp), d1 = __ldg(dp + 1);
            FastItemDesc d;
            *reinterpret_cast<uint4 *>(&d) = d0; *(reinterpret_cast<uint4 *>(&d) + 1) = d1;
            if (TENSOR) fast_issue_tensor(maps, d, f, buf_of(slot), &bars[slot], lane);
            else        fast_issue_rows<BYTES>(B, d, f, buf_of(slot), &bars[slot], lane);
            if (++slot == nbuf) { slot = 0; if (wrapped) phase ^= 1u; wrapped = true; }
            f += step_f; r += step_r;
            if (r >= ipf) { r -= ipf; f++; }
        }
        return;
    }

    // ================= consumer warps
    int rot = 0;                              // consumer thread that takes unit 0 of the current item
    for (; f < nframes; ) {
        const uint4 *dp = reinterpret_cast<const uint4 *>(P.items + r);
        const uint4 d0 = __ldg(dp), d1 = __ldg(dp + 1);
        FastItemDesc d;
        *reinterpret_cast<uint4 *>(&d) = d0; *(reinterpret_cast<uint4 *>(&d) + 1) = d1;
        const unsigned char *buf = buf_of(slot);
        uint32_t *rec_tile = B.rec + (size_t)f * L.rec_per_frame + d.rec_off;
        const int rowb = d.rowb, upr = d.upr, nunits = d.nunits;
        const uint32_t magic = d.magic, rec_stride = d.rec_stride;
        const unsigned char *base = buf + 2 * rowb + d.o0;
        const bool first_chunk = d.flags & 2, last_chunk = d.flags & 4;
        int u = tid - rot;
        if (u < 0) u += kFastComp;
        mbar_wait(&bars[slot], phase);
        // the units of an item start at the thread after the one that took the last unit of the previous item
        for (; u < nunits; u += kFastComp) {
            const int rr = (int)__umulhi((uint32_t)u << 12, magic);      // u / upr  (exact for u < 2^12)
            const int ux = u - rr * upr;
            const unsigned char *crow = base + rr * rowb + ux * (8 * BYTES);
            const bool first = first_chunk && ux == 0, last = last_chunk && ux == upr - 1;
            uint32_t out[8];
            if constexpr (BYTES == 1 && NIN == 3) fast_unit8_packed(crow, rowb, first, last, laneA, out);
            else fast_unit_generic<BYTES, NIN>(crow, rowb, first, last, laneA, laneC, L.sample_shift, bits, out);
            // one 256-bit store per thread (a full 32-byte sector; record lines are padded to 32 records)
            char *dst = reinterpret_cast<char *>(rec_tile) + (size_t)((uint32_t)rr * rec_stride + (uint32_t)ux * 8u) * 4u;
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         :: "l"(dst), "r"(out[0]), "r"(out[1]), "r"(out[2]), "r"(out[3]), "r"(out[4]), "r"(out[5]), "r"(out[6]), "r"(out[7])
                         : "memory");
        }
        rot += d.nunits_mod;
        if (rot >= kFastComp) rot -= kFastComp;
        // this warp is done with the buffer
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
        if (++slot == nbuf) { slot = 0; phase ^= 1u; }
        f += step_f; r += step_r;
        if (r >= ipf) { r -= ipf; f++; }
    }
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

// Work-item table of one frame (the same for every frame) and the shared-memory plan.
void build_pixel_fast_plan(const Tables &tab, FastPlan &plan)
{
    const Layout &L = tab.layout;
    const int BYTES = L.src_kind == SRC_PLANAR16 ? 2 : 1;
    const int chunk = kFastChunkBytes / BYTES;
    plan.items.clear();
    for (int p = 0; p < 4; p++) plan.row_bytes[p] = 0;
    // staged row pitch per source plane: left block + misalignment + the widest chunk of the plane + next sample, as a
    // multiple of 64 bytes (two rows are then a multiple of the 128-byte alignment TMA wants for its destination)
    for (int p = 0; p < L.nplanes; p++) {
        int wmax = 0;
        for (const SliceGeom &g : tab.slices) wmax = std::max(wmax, g.pw[p]);
        const int rowb = (16 + 15 + std::min(wmax, chunk) * BYTES + BYTES + 63) & ~63;
        const int sp = L.plane[p].src_plane;
        plan.row_bytes[sp] = std::max(plan.row_bytes[sp], rowb);
    }
    int max_rowb = 64;
    for (int p = 0; p < 4; p++) max_rowb = std::max(max_rowb, plan.row_bytes[p]);
    for (const TileDesc &td : tab.tiles) {
        const SliceGeom &g = tab.slices[td.slice];
        const int p = td.plane, w = g.pw[p], sp = L.plane[p].src_plane;
        for (int cx0 = 0; cx0 < w; cx0 += chunk) {
            const int cw = std::min(chunk, w - cx0);
            const int gx = (g.px0[p] + cx0) * BYTES;
            const bool top = td.y0 == 0;
            FastItemDesc d;
            memset(&d, 0, sizeof(d));
            d.c0 = ((gx & ~15) - 16) >> 2;
            d.c1 = g.py0[p] + td.y0 - (top ? 0 : 2);
            d.rec_off = g.rec_first + tab.lines[g.line_first + td.line_first].rec_off + (uint32_t)cx0;
            d.rec_stride = (uint16_t)((w + 31) & ~31);
            d.upr = (uint16_t)((cw + 7) / 8);
            d.magic = (1048576u + d.upr - 1u) / d.upr;
            d.nunits = (uint16_t)(d.upr * td.nrows);
            d.nunits_mod = (uint16_t)(d.nunits % kFastComp);
            d.o0 = (uint16_t)(16 + (gx & 15));
            d.rowb = (uint16_t)plan.row_bytes[sp];
            d.src_plane = (uint8_t)sp;
            d.flags = (uint8_t)((top ? 1 : 0) | (cx0 == 0 ? 2 : 0) | (cx0 + cw >= w ? 4 : 0));
            d.nrows = (uint8_t)td.nrows;
            plan.items.push_back(d);
        }
    }
    plan.items_per_frame = (int32_t)plan.items.size();
    plan.buf_bytes = kFastBufRows * max_rowb;                        // multiple of 128 (36 rows x a multiple of 64)
    // 1 KB of the 228 KB is reserved by the system in front of the dynamic region, so the 64 KB aligned table starts
    // 63 KB into it (the kernel checks the real addresses and traps if the plan does not fit)
    const int tables = kFastTabAB + (L.ctx_inputs == 5 ? kFastTabC : 0);
    const int front = 65536 - 1024, back = kFastMaxSmem - front - tables - 2 * 2 * kFastMaxBufs * 8;
    int nbuf = kFastMaxBufs;
    while (nbuf > 2 && 2 * nbuf - std::min(front / plan.buf_bytes, 2 * nbuf) > back / plan.buf_bytes) nbuf--;
    plan.nbuf = nbuf;
    plan.smem_bytes = kFastMaxSmem;
}

template <int BYTES, int NIN>
static cudaError_t set_attr(int sm)
{
    cudaError_t e = cudaFuncSetAttribute(k_pixel_fast<BYTES, NIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_pixel_fast<BYTES, NIN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
}

cudaError_t configure_pixel_fast(const FastPlan &plan)
{
    const int sm = plan.smem_bytes;
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
static void launch_t(const EncDeviceTables &t, const EncBatch &b, const FastPlan &plan, const FastParams &P, int grid, cudaStream_t s,
                     const uint8_t *const *frame0_planes, long long frame_stride)
{
    FastMaps maps;
    memset(&maps, 0, sizeof(maps));
    bool tensor = frame_stride >= 0 && (frame_stride & 15) == 0 && getenv("FFV1B200_PIXEL_ROWCOPY") == nullptr;
    PFN_cuTensorMapEncodeTiled_v12000 enc = tensor ? tensor_map_encoder() : nullptr;
    tensor = tensor && enc != nullptr;
    const Layout &L = t.layout;
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    if (const char *v = getenv("FFV1B200_TMA_L2")) { const int q = atoi(v); promo = q == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (q == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : (q == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B)); }
    for (int p = 0; p < L.nplanes && tensor; p++) {
        const int sp = L.plane[p].src_plane;
        const int rows = (L.height + (1 << L.plane[p].vshift) - 1) >> L.plane[p].vshift;
        const cuuint64_t dims[3] = {(cuuint64_t)(b.linesize[sp] / 4), (cuuint64_t)rows, (cuuint64_t)b.nframes};
        const cuuint64_t strides[2] = {(cuuint64_t)b.linesize[sp], (cuuint64_t)(b.nframes > 1 ? frame_stride : b.linesize[sp] * (long long)rows)};
        const cuuint32_t box[3] = {(cuuint32_t)(plan.row_bytes[sp] / 4), (cuuint32_t)kFastRows, 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        if (strides[1] & 15) { tensor = false; break; }
        CUresult r = enc(&maps.m[sp], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t *>(frame0_planes[sp]), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) tensor = false;
    }
    if (tensor) k_pixel_fast<BYTES, NIN, true><<<grid, kFastThreads, plan.smem_bytes, s>>>(t, b, P, maps);
    else        k_pixel_fast<BYTES, NIN, false><<<grid, kFastThreads, plan.smem_bytes, s>>>(t, b, P, maps);
}

// frame0_planes: host copy of the first frame's plane pointers; frame_stride: bytes between the same plane of consecutive
// frames when that distance is constant over the batch (tensor-map TMA), else -1 (one bulk copy per row)
void launch_pixel_fast(const EncDeviceTables &t, const EncBatch &b, const FastPlan &plan, const FastItemDesc *d_items, int num_sms,
                       cudaStream_t s, const uint8_t *const *frame0_planes, long long frame_stride)
{
    const Layout &L = t.layout;
    FastParams P;
    P.items = d_items; P.items_per_frame = plan.items_per_frame; P.nbuf = plan.nbuf; P.buf_bytes = plan.buf_bytes;
    const long total = (long)plan.items_per_frame * b.nframes;
    const int grid = (int)std::min<long>((total + 1) / 2, (long)num_sms);
    const bool five = L.ctx_inputs == 5;
    if (L.src_kind == SRC_PLANAR8) {
        if (five) launch_t<1, 5>(t, b, plan, P, grid, s, frame0_planes, frame_stride);
        else      launch_t<1, 3>(t, b, plan, P, grid, s, frame0_planes, frame_stride);
    } else {
        if (five) launch_t<2, 5>(t, b, plan, P, grid, s, frame0_planes, frame_stride);
        else      launch_t<2, 3>(t, b, plan, P, grid, s, frame0_planes, frame_stride);
    }
}

} // namespace ffv1
