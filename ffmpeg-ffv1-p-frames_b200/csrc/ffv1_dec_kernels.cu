// ffv1_dec_kernels.cu -- hand-written sm_100a kernels of the FFV1 decode path.
//
//   k_dec_crc   per-slice CRC-32 check (ffv1dec.c:963-976): chunk CRCs in parallel, combined in GF(2).
//   k_decode    decode_slice / decode_slice_header / decode_plane / decode_rgb_frame / decode_line
//               (ffv1dec.c:361-474, 282-359, 183-280, 100-181) with get_symbol_inline / get_rac (ffv1dec.c:42-63,
//               rangecoder.h:104-145) or the Golomb-Rice reader (ffv1dec.c:70-98, golomb.h:270-300, 367-372).
//               Decoding is strictly serial inside a slice (every context needs the reconstructed left neighbour), and
//               a non-keyframe needs the model state the previous frame left behind, so the unit of parallelism is the
//               chain (GOP segment, slice): one warp per chain.  Lane 0 runs the entropy decoder and reconstruction
//               into an int16 line ring (the reference's sample_buffer); after every line the whole warp converts
//               the line (inverse RCT for RGB) and writes it to the output frame with coalesced stores.
//   k_conceal   damaged slices are replaced by the co-located pixels of the previous frame (ffv1dec.c:998-1021).
#include <cstdio>
#include <cstdlib>
#include "ffv1_dec_kernels.cuh"

namespace ffv1 {

__constant__ uint8_t c_log2_run[41] = {
    0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7,
    8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24,
};

// ------------------------------------------------------------------------------------------------ CRC
__device__ __forceinline__ uint32_t gf_mulmod32(uint32_t a, uint32_t b)
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

constexpr int kCrcThreads = 128;

__global__ void __launch_bounds__(kCrcThreads) k_dec_crc(const DecDeviceTables T, const DecBatch B)
{
    __shared__ uint32_t s_tab[256];
    __shared__ uint32_t s_part[kCrcThreads];
    const int f = blockIdx.x / T.max_slices, s = blockIdx.x - f * T.max_slices;
    const int tid = threadIdx.x;
    if (s >= B.slice_count[f]) return;
    for (int n = tid; n < 256; n += kCrcThreads) {
        uint32_t c = (uint32_t)n << 24;
#pragma unroll
        for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : (c << 1);
        s_tab[n] = c;
    }
    __syncthreads();
    const uint8_t *src = B.pkt + B.pkt_off[f] + B.slice_start[f * T.max_slices + s];
    const uint32_t len = B.slice_size[f * T.max_slices + s];
    const uint32_t chunk = (len + kCrcThreads - 1) / kCrcThreads;
    // chunks are aligned to the END of the slice: leading zero bytes do not change a CRC whose initial value is 0
    const long long beg = (long long)len - (long long)(kCrcThreads - tid) * chunk;
    uint32_t crc = 0;
    for (long long i = beg < 0 ? 0 : beg; i < beg + (long long)chunk; i++) crc = (crc << 8) ^ s_tab[(crc >> 24) ^ src[i]];
    uint32_t mult = 1u, base = 0x100u;                        // x^(8*chunk) mod P
    for (uint32_t e = chunk; e; e >>= 1) {
        if (e & 1u) mult = gf_mulmod32(mult, base);
        base = gf_mulmod32(base, base);
    }
    s_part[tid] = crc;
    __syncthreads();
    for (int stride = 1; stride < kCrcThreads; stride <<= 1) {
        uint32_t v = 0;
        const bool active = (tid % (2 * stride)) == 0;
        if (active) v = gf_mulmod32(s_part[tid], mult) ^ s_part[tid + stride];
        __syncthreads();
        if (active) s_part[tid] = v;
        mult = gf_mulmod32(mult, mult);
        __syncthreads();
    }
    if (tid == 0 && s_part[0] != 0) B.damaged[f * T.max_slices + s] |= 1u;
}

void launch_dec_crc(const DecDeviceTables &t, const DecBatch &b, cudaStream_t s)
{
    if (!t.ec) return;
    k_dec_crc<<<b.nframes * t.max_slices, kCrcThreads, 0, s>>>(t, b);
}

// ------------------------------------------------------------------------------------------------ entropy readers
struct RDec {                // rangecoder.h:35-45, decoder side
    uint32_t low, range;
    const uint8_t *ptr, *end, *start;
};

__device__ __forceinline__ void rd_init(RDec &c, const uint8_t *buf, uint32_t size)
{
    // ff_init_range_decoder (rangecoder.c:53-61)
    c.start = buf; c.end = buf + size;
    c.range = 0xFF00u;
    c.low = size >= 2 ? ((uint32_t)buf[0] << 8 | buf[1]) : (size == 1 ? (uint32_t)buf[0] << 8 : 0u);
    c.ptr = buf + 2;
}

__device__ __forceinline__ int rd_get(RDec &c, uint8_t *state, const uint16_t *lut)
{
    // get_rac + refill (rangecoder.h:104-145).  lut[s] = zero_state[s] | one_state[s] << 8: both successors of the
    // state are fetched while the interval is split, the decoded bit only selects one (the dependent chain
    // state -> bit -> table -> state is what a slice decoder's speed hangs on)
    const uint32_t s = *state;
    const uint32_t zz = lut[s];
    const uint32_t range1 = (c.range * s) >> 8;
    const uint32_t r0 = c.range - range1;
    const int bit = c.low >= r0;
    c.low -= bit ? r0 : 0u;
    c.range = bit ? range1 : r0;
    *state = (uint8_t)(bit ? zz >> 8 : zz);
    if (c.range < 0x100u) {
        c.range <<= 8;
        c.low <<= 8;
        if (c.ptr < c.end) c.low += *c.ptr;
        c.ptr++;
    }
    return bit;
}

__device__ int rd_symbol(RDec &c, uint8_t *state, const uint16_t *lut, bool is_signed, int &err)
{
    // get_symbol_inline (ffv1dec.c:42-63)
    if (rd_get(c, state, lut)) return 0;
    int e = 0;
    while (rd_get(c, state + 1 + min(e, 9), lut)) {
        if (++e > 31) { err = 1; return 0; }
    }
    int a = 1;
    for (int i = e - 1; i >= 0; i--) a += a + rd_get(c, state + 22 + min(i, 9), lut);
    if (is_signed && rd_get(c, state + 11 + min(e, 10), lut)) return -a;
    return a;
}

struct BitR {                // MSB-first reader (get_bits.h), zero bits past the end
    const uint8_t *buf;
    uint32_t nbytes;
    uint32_t pos;            // bit position
    uint32_t wbyte;          // first byte held in `win`
    unsigned long long win;  // bytes wbyte .. wbyte+7, big endian
};

__device__ __forceinline__ void br_load(BitR &r, uint32_t byte)
{
    unsigned long long w = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        w <<= 8;
        if (byte + i < r.nbytes) w |= r.buf[byte + i];
    }
    r.win = w; r.wbyte = byte;
}

__device__ __forceinline__ uint32_t br_peek32(BitR &r)
{
    const uint32_t byte = r.pos >> 3;
    if (byte - r.wbyte > 3u) br_load(r, byte);           // also true when byte < wbyte (never happens: pos only grows)
    const uint32_t sh = r.pos - r.wbyte * 8u;            // 0..31
    return (uint32_t)((r.win << sh) >> 32);
}

__device__ __forceinline__ uint32_t br_get(BitR &r, int n)
{
    if (!n) return 0u;
    const uint32_t v = br_peek32(r) >> (32 - n);
    r.pos += n;
    return v;
}

struct __align__(8) VlcDev { int16_t drift; uint16_t error_sum; int8_t bias; uint8_t count; uint16_t pad; };

__device__ __forceinline__ int fold_bits(int d, int bits) { return (d << (32 - bits)) >> (32 - bits); }

__device__ int br_vlc(BitR &r, VlcDev *sp, int bits)
{
    // get_vlc_symbol (ffv1dec.c:70-98) + get_sr_golomb (golomb.h:270-300, 367-372) + update_vlc_state (ffv1.h:192-224)
    VlcDev s = *sp;
    int k = 0;
    for (int i = s.count; i < s.error_sum; i += i) k++;
    const uint32_t buf = br_peek32(r);
    const int lz = buf ? __clz(buf) : 32;
    uint32_t u;
    if (lz < 12) {
        r.pos += lz + 1;
        u = ((uint32_t)lz << k) + br_get(r, k);
    } else {
        r.pos += 12;
        u = br_get(r, bits) + 11u;
    }
    int v = (int)(u >> 1) ^ -(int)(u & 1u);
    v ^= (2 * s.drift + s.count) >> 31;
    const int ret = fold_bits(v + s.bias, bits);
    // update
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
    return ret;
}

__device__ __forceinline__ int median3(int a, int b, int c) { return max(min(a, b), min(max(a, b), c)); }

struct SliceRd {
    RDec rc;
    BitR br;
    int golomb;
    int run_index;
    int err;
};

// one line of one plane into the ring (ffv1dec.c:100-181).  cur/top/top2 point at x = 0 of the ring rows.
__device__ void dec_line(SliceRd &sr, uint8_t *model, const int16_t *q, const uint16_t *lut, int16_t *cur, const int16_t *top,
                         const int16_t *top2, int w, int bits)
{
    const bool five = q[3 * 256 + 127] != 0;
    const int mask = (1 << bits) - 1;
    int run_count = 0, run_mode = 0;
    int L = cur[-1], LL = cur[-2], LT = top[-1], Tp = top[0];
    for (int x = 0; x < w; x++) {
        const int RT = top[x + 1];
        int ctx = q[(L - LT) & 0xFF] + q[256 + ((LT - Tp) & 0xFF)] + q[512 + ((Tp - RT) & 0xFF)];
        if (five) ctx += q[768 + ((LL - L) & 0xFF)] + q[1024 + ((top2[x] - Tp) & 0xFF)];
        bool sign = false;
        if (ctx < 0) { ctx = -ctx; sign = true; }
        int diff;
        if (!sr.golomb) {
            diff = rd_symbol(sr.rc, model + (size_t)ctx * 32, lut, true, sr.err);
        } else {
            VlcDev *vs = reinterpret_cast<VlcDev *>(model) + ctx;
            if (ctx == 0 && run_mode == 0) run_mode = 1;
            if (run_mode) {
                if (run_count == 0 && run_mode == 1) {
                    if (br_get(sr.br, 1)) {
                        run_count = 1 << c_log2_run[sr.run_index];
                        if (x + run_count <= w) sr.run_index++;
                    } else {
                        const int lr = c_log2_run[sr.run_index];
                        run_count = lr ? (int)br_get(sr.br, lr) : 0;
                        if (sr.run_index) sr.run_index--;
                        run_mode = 2;
                    }
                }
                run_count--;
                if (run_count < 0) {
                    run_mode = 0; run_count = 0;
                    diff = br_vlc(sr.br, vs, bits);
                    if (diff >= 0) diff++;
                } else
                    diff = 0;
            } else
                diff = br_vlc(sr.br, vs, bits);
        }
        if (sign) diff = -diff;
        const int v = (int)(int16_t)((median3(L, Tp, L + Tp - LT) + diff) & mask);      // ffv1dec.c:178, int16 line buffer
        cur[x] = (int16_t)v;
        LL = L; L = v; LT = Tp; Tp = RT;
    }
}

// ---- range-coder sample loop with the coder's working set in registers (same arithmetic as rd_get / rd_symbol).
// A chain's speed is the number of instructions its one serial lane issues per sample (a lone warp pays ~4.5 cycles per
// issued instruction, profiles/r01_k_decode_probe.txt), so the reader is built to issue few:
//  * the 32 state bytes of the sample's context come in with two 128-bit loads; a symbol touches every byte at most once
//    while its exponent is < 9 (bytes 1+e, 22+i, 11+e are distinct), so the decisions read them from registers and the
//    successor states leave with byte stores that nothing waits for (larger exponents continue on the memory copy);
//  * the successor pair lut[s] is fetched before the decision, the decoded bit only selects;
//  * the bitstream comes through a 32-bit register window refilled from an aligned word loaded four renormalisations
//    ahead;
//  * the part of the next sample's context that does not depend on the sample being decoded (Q1, Q2, Q4) is looked up
//    one sample ahead.
struct FastRc {
    uint32_t low, range;
    uint32_t win, cnt;       // `cnt` unread bytes in the top of `win`
    uint32_t nxt;            // the aligned word at p (big endian, bytes past the end zero)
    const uint8_t *p, *end;
};

__device__ __forceinline__ uint32_t be32_masked(const uint8_t *p, const uint8_t *end)
{
    if (p >= end) return 0u;
    uint32_t w = __byte_perm(*reinterpret_cast<const uint32_t *>(p), 0u, 0x0123);
    const long left = end - p;
    if (left < 4) w &= 0xFFFFFFFFu << (8 * (4 - (int)left));
    return w;
}

__device__ __forceinline__ void fr_open(FastRc &c, const RDec &rc)
{
    const uint32_t m = (uint32_t)(reinterpret_cast<uintptr_t>(rc.ptr) & 3u);
    const uint8_t *p0 = rc.ptr - m;                      // packets start 16-byte aligned in the staging buffer
    c.low = rc.low; c.range = rc.range; c.end = rc.end;
    c.win = be32_masked(p0, rc.end) << (8 * m);
    c.cnt = 4 - m;
    c.p = p0 + 4;
    c.nxt = be32_masked(c.p, rc.end);
}

__device__ __forceinline__ void fr_close(const FastRc &c, RDec &rc)
{
    rc.low = c.low; rc.range = c.range; rc.ptr = c.p - c.cnt;
}

// the table reads are spelled as shared-memory loads from a 32-bit address (the tables never change while the kernel
// runs, so the loads are pure)
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
    uint16_t v;
    asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int lds_s16(uint32_t addr)
{
    short v;
    asm("ld.shared.s16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ uint32_t fr_bit(FastRc &c, uint32_t s)
{
    const uint32_t range1 = (c.range * s) >> 8;
    const uint32_t r0 = c.range - range1;
    const bool bit = c.low >= r0;
    if (bit) c.low -= r0;
    c.range = bit ? range1 : r0;
    if (c.range < 0x100u) {
        c.range <<= 8;
        c.low = (c.low << 8) | (c.win >> 24);
        c.win <<= 8;
        if (--c.cnt == 0) {
            c.win = c.nxt; c.cnt = 4; c.p += 4;
            c.nxt = be32_masked(c.p, c.end);
        }
    }
    return bit;
}

__device__ __forceinline__ uint32_t fr_mem(FastRc &c, uint8_t *state, const uint16_t *lut)
{
    const uint32_t s = *state;
    const uint32_t zz = lut[s];
    const uint32_t bit = fr_bit(c, s);
    *state = (uint8_t)(bit ? zz >> 8 : zz);
    return bit;
}

#define FR_DECIDE(sv, rowptr)                                                            \
    do {                                                                                 \
        s = (sv);                                                                        \
        zz = lds_u16(lut_sa + 2u * s);                                                   \
        bit = fr_bit(c, s);                                                             \
        *(rowptr) = (uint8_t)(bit ? zz >> 8 : zz);                                       \
    } while (0)

__device__ __forceinline__ int fr_symbol(FastRc &c, uint8_t *row, const uint16_t *lut, uint32_t lut_sa, int &err)
{
    const uint4 ra = *reinterpret_cast<const uint4 *>(row), rb = *reinterpret_cast<const uint4 *>(row + 16);
    const uint32_t w[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
    uint32_t s, zz, bit;
    FR_DECIDE(__byte_perm(w[0], 0u, 0x4440), row);
    if (bit) return 0;
    int e = 0;
    bool open = true;
#pragma unroll
    for (int k = 0; k < 9; k++) {
        FR_DECIDE(__byte_perm(w[(1 + k) >> 2], 0u, 0x4440 + ((1 + k) & 3)), row + 1 + k);
        if (!bit) { open = false; break; }
        e = k + 1;
    }
    if (open) {
        while (fr_mem(c, row + 1 + min(e, 9), lut)) {
            if (++e > 31) { err = 1; return 0; }
        }
        int a = 1;
        for (int i = e - 1; i >= 0; i--) a += a + (int)fr_mem(c, row + 22 + min(i, 9), lut);
        return fr_mem(c, row + 11 + min(e, 10), lut) ? -a : a;
    }
    const uint32_t sgw = e == 0 ? w[2] >> 24 : (e <= 4 ? w[3] : w[4]) >> (8 * ((e - 1) & 3));
    unsigned long long mq = ((unsigned long long)__funnelshift_r(w[6], w[7], 16) << 32) | __funnelshift_r(w[5], w[6], 16);
    mq <<= (8 * (8 - e)) & 63;
    int a = 1;
    for (int i = e - 1; i >= 0; i--) {
        const uint32_t sv = (uint32_t)(mq >> 56);
        mq <<= 8;
        FR_DECIDE(sv, row + 22 + i);
        a += a + (int)bit;
    }
    FR_DECIDE(sgw & 0xFFu, row + 11 + e);
    return bit ? -a : a;
}

// decode_line for the range coder (ffv1dec.c:100-181 with ac != AC_GOLOMB_RICE)
template <bool FIVE>
__device__ __forceinline__ void dec_line_rc(SliceRd &sr, uint8_t *model, const int16_t *q, const uint16_t *lut, int16_t *cur,
                                            const int16_t *top, const int16_t *top2, int w, int bits)
{
    FastRc c;
    fr_open(c, sr.rc);
    const int mask = (1 << bits) - 1;
    int err = 0;
    uint32_t lut_sa = (uint32_t)__cvta_generic_to_shared(lut);
    asm volatile("" : "+r"(lut_sa));
    int L = cur[-1], LL = cur[-2], LT = top[-1], Tp = top[0], RT = top[1];
    int qn = q[256 + ((LT - Tp) & 0xFF)] + q[512 + ((Tp - RT) & 0xFF)];
    if (FIVE) qn += q[1024 + ((top2[0] - Tp) & 0xFF)];
    for (int x = 0; x < w; x++) {
        int ctx = q[(L - LT) & 0xFF] + qn;
        if (FIVE) ctx += q[768 + ((LL - L) & 0xFF)];
        const int RT2 = top[x + 2];                       // inside the row's padding at the right edge, not used there
        qn = q[256 + ((Tp - RT) & 0xFF)] + q[512 + ((RT - RT2) & 0xFF)];
        if (FIVE) qn += q[1024 + ((top2[x + 1] - RT) & 0xFF)];
        const int pred = median3(L, Tp, L + Tp - LT);
        const bool neg = ctx < 0;
        int diff = fr_symbol(c, model + (size_t)abs(ctx) * 32, lut, lut_sa, err);
        if (neg) diff = -diff;
        const int v = (int)(int16_t)((pred + diff) & mask);                             // ffv1dec.c:178, int16 line buffer
        cur[x] = (int16_t)v;
        LL = L; L = v; LT = Tp; Tp = RT; RT = RT2;
    }
    sr.err |= err;
    fr_close(c, sr.rc);
}

// Version 4, slice_coding_mode 1 ("PCM", ffv1dec.c:111-122): every sample is `bits` raw bits, each decoded with a
// fresh state 128 (an even split of the interval); no prediction, no adaptive state.
__device__ void dec_line_pcm(SliceRd &sr, const uint16_t *lut, int16_t *cur, int w, int bits)
{
    for (int x = 0; x < w; x++) {
        int v = 0;
        for (int i = 0; i < bits; i++) {
            uint8_t st = 128;
            v += v + rd_get(sr.rc, &st, lut);
        }
        cur[x] = (int16_t)v;
    }
}

// ---- the common case (planar, range coder, three-table context, model and lines in shared memory): what one lane has to
// do per sample is cut down to what depends on the sample decoded just before.  The warp prepares every line in parallel:
// per x a record {T = top[x], q12 = Q1[LT-T] + Q2[T-RT]} from the finished line above; the serial lane reads one record,
// adds Q0[L-LT], decodes, and stores the sample into the record; the warp then writes the line out and turns `cur` into
// the next line's `top`.
// FIVE: the five-table context model (ffv1.h:176-186) adds Q3[LL - L], which the serial lane looks up, and Q4[TT - T],
// which the warp folds into q12 from the line two rows up (kept in `top2`).
struct __align__(8) LineRec { int16_t top, q12, cur, top2; };

template <bool FIVE>
__device__ __forceinline__ void dec_line_rec(SliceRd &sr, uint8_t *model, const int16_t *q, const uint16_t *lut, LineRec *rec,
                                             int w, int bits, int topm1)
{
    FastRc c;
    fr_open(c, sr.rc);
    const int mask = (1 << bits) - 1;
    int err = 0;
    uint32_t lut_sa = (uint32_t)__cvta_generic_to_shared(lut), q_sa = (uint32_t)__cvta_generic_to_shared(q);
    asm volatile("" : "+r"(lut_sa), "+r"(q_sa));         // keep them in registers (otherwise re-derived from %cluster_ctarank per sample)
    int L = rec[0].top, LT = topm1;                      // sample[1][-1] = sample[0][0] (ffv1dec.c:199)
    int LL = 0;                                          // sample[1][-2] stays 0
    LineRec *r = rec, *const rend = rec + w;
    do {
        const int Tp = r->top, q12 = r->q12;
        const int d = L - LT;
        int ctx = lds_s16(q_sa + 2u * (uint32_t)(d & 0xFF)) + q12;
        if (FIVE) { ctx += lds_s16(q_sa + 2u * (768u + (uint32_t)((LL - L) & 0xFF))); LL = L; }
        const int pred = median3(L, Tp, d + Tp);
        int diff = fr_symbol(c, model + (size_t)abs(ctx) * 32, lut, lut_sa, err);
        if (ctx < 0) diff = -diff;
        L = (int)(int16_t)((pred + diff) & mask);
        r->cur = (int16_t)L;
        LT = Tp;
        r++;
    } while (r != rend);
    sr.err |= err;
    fr_close(c, sr.rc);
}

constexpr int kDecSmemQuant = 2 * 5 * 256;

constexpr int kDecWarps = 2;               // chains per CTA (they share the quantisation / transition tables)

// The rectangle the slice grid gives slice `si` (ffv1.c:124-143), cleared in the planes psel picks (-1: all, 0: the first,
// 1: all but the first).  Used when the pictures are written straight into a buffer nobody has cleared (DecBatch::zero_fill)
// for slices that are missing, refused, or announce another rectangle: what the staged path's memset leaves there.
__device__ void dec_zero_slice(const DecDeviceTables &T, uint8_t *cur, int si, int psel, int lane)
{
    const int gx = si % T.num_h_slices, gy = si / T.num_h_slices;
    const int x0 = T.width * gx / T.num_h_slices, x1 = T.width * (gx + 1) / T.num_h_slices;
    const int y0 = T.height * gy / T.num_v_slices, y1 = T.height * (gy + 1) / T.num_v_slices;
    const int nsrc = T.colorspace ? (T.rgb32 ? 1 : 3) : (T.ya8 ? 1 : 1 + 2 * T.chroma_planes + T.transparency);
    for (int p = 0; p < nsrc; p++) {
        if ((psel == 0 && p != 0) || (psel == 1 && p == 0)) continue;
        const bool chroma = !T.colorspace && !T.ya8 && T.chroma_planes && (p == 1 || p == 2);
        const int hs = chroma ? T.hshift : 0, vs = chroma ? T.vshift : 0;
        const int src_plane = (!T.colorspace && !T.ya8 && !T.chroma_planes && p == 1) ? 3 : p;
        const int bpp = T.rgb32 ? 4 : (T.ya8 ? 2 : (T.bits > 8 ? 2 : 1));
        const int bx0 = (x0 >> hs) * bpp, bx1 = (-((-x1) >> hs)) * bpp;
        const int ry0 = y0 >> vs, ry1 = -((-y1) >> vs);
        const int rb = bx1 - bx0;
        for (int i = lane; i < rb * (ry1 - ry0); i += 32)
            cur[(size_t)T.plane_off[src_plane] + (size_t)(ry0 + i / rb) * T.plane_pitch[src_plane] + bx0 + i % rb] = 0;
    }
}

// PIPE (planar YUV): the two warps of a CTA work on ONE chain.  Inside a slice the planes follow each other
// in one coder stream, but across the frames of a GOP the only thing a plane needs from the frame before is the model of
// its own plane context (ffv1dec.c:419-420) -- luma of frame f+1 can be decoded while chroma of frame f still is.  Warp 0
// decodes the luma plane of every frame (its model stays in shared memory) and hands the coder (low, range, position) and
// the slice header to warp 1 through a small mailbox ring; warp 1 decodes the remaining planes (models in global memory:
// it has half the samples and time to spare) and does the end-of-slice check.  A chain then takes the time of its luma
// planes: 2/3 of the frames' samples for 4:2:0.  Used when the batch's chains fit the SMs at one chain per CTA.
struct DecMail { uint32_t low, range, pos; int32_t err, sx, sy, sw, sh, qti, v4, bad; uint32_t br_off, br_nbytes, br_pos; };
constexpr int kDecMailSlots = 4;
constexpr int kBadSeen = 3;             // `bad` value in warp 1: warp 0 refused the slice and has flagged it already

// MINB = resident CTAs per SM the register allocation aims at (8: no spills; 12: for batches with more chains than 8 CTAs hold)
template <int MINB, bool PIPE>
__global__ void __launch_bounds__(32 * kDecWarps, MINB) k_decode(const DecDeviceTables T, const DecBatch B)
{
    extern __shared__ __align__(16) unsigned char s_dyn[];      // per warp: [model of the current plane context][line ring]
    __shared__ int16_t s_quant[kDecSmemQuant];
    __shared__ uint16_t s_lut[256];                              // zero_state | one_state << 8
    __shared__ DecMail s_mail[kDecMailSlots];
    __shared__ int s_pub, s_ack;                                 // PIPE: slices handed over by warp 0 / taken by warp 1
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kDecSmemQuant; i += 32 * kDecWarps) s_quant[i] = T.quant[i];
    for (int i = threadIdx.x; i < 256; i += 32 * kDecWarps) s_lut[i] = (uint16_t)(T.lut[i] | (T.lut[256 + i] << 8));
    if (threadIdx.x == 0) { s_pub = 0; s_ack = 0; }
    __syncthreads();

    const int chain = PIPE ? (int)blockIdx.x : (int)blockIdx.x * kDecWarps + warp;
    if (chain >= B.nseg * T.max_slices) return;
    const bool luma_warp = !PIPE || warp == 0, rest_warp = !PIPE || warp == 1;
    const int ring_bytes = T.smem_ring_w * kDecSmemRingBytes;
    const int per_warp = T.smem_model + ring_bytes;
    // PIPE: [luma model][ring of warp 0][ring of warp 1]
    uint8_t *s_model = PIPE ? s_dyn : s_dyn + (size_t)warp * per_warp;
    int16_t *s_ring = reinterpret_cast<int16_t *>(PIPE ? s_dyn + T.smem_model + (size_t)warp * ring_bytes : s_model + T.smem_model);
    int nhand = 0;                          // PIPE: slices handed over / taken so far by this warp
    const int seg = chain / T.max_slices, si = chain - seg * T.max_slices;
    const int f0 = B.seg_first[seg], f1 = B.seg_first[seg + 1];
    uint8_t *models = B.state + ((size_t)B.seg_set[seg] * T.max_slices + si) * 3 * T.state_stride;
    int16_t *ring = B.ring + (size_t)chain * 4 * 3 * T.ring_w + kDecRingPad;
    const int golomb = T.ac == 0;
    const int bits = T.coded_bits;
    const int nplanes = T.colorspace ? 3 + T.transparency : (T.ya8 ? 2 : 1 + 2 * T.chroma_planes + T.transparency);

    int sm_pc = -1, sm_nb = 0;             // plane context whose model is in shared memory, its size in 16-byte units
#ifdef FFV1_DEC_PROBE
    long long probe_serial = 0, probe_samples = 0;
    const long long probe_t0 = clock64();
#endif
    for (int f = f0; f < f1; f++) {
        const int zsel = PIPE ? warp : -1;                       // the planes this warp writes
        if (si >= B.slice_count[f]) {
            if (B.zero_fill) dec_zero_slice(T, B.out + (size_t)f * T.frame_bytes, si, zsel, lane);
            continue;
        }
        const int fs = f * T.max_slices + si;
        const uint8_t *sbeg = B.pkt + B.pkt_off[f] + B.slice_start[fs];
        const uint32_t ssize = B.slice_size[fs];
        const bool key = B.frame_key[f] != 0;

        // ---- lane 0: coder init, keyframe bit, slice header (ffv1dec.c:282-359)
        SliceRd sr;
        int sx = 0, sy = 0, sw = 0, sh = 0, qti[3] = {0, 0, 0};
        int bad = 0;
        int v4 = 1 | 1 << 4;                // version 4 header fields: rct_by | rct_ry << 4 | coding mode << 8 | reset contexts << 16
        if (lane == 0) {
            sr.golomb = golomb; sr.run_index = 0; sr.err = 0;
            rd_init(sr.rc, sbeg, ssize);
            if (PIPE && warp == 1) {
                // the coder as warp 0 left it behind the luma plane, and the slice header it read
                volatile int *pub = &s_pub;
                int spins = 0;
                while (*pub <= nhand && ++spins < (1 << 25)) __nanosleep(1000);   // (a luma plane takes milliseconds)
                __threadfence_block();
                if (*pub <= nhand) bad = 1;                      // (never seen: warp 0 hands over every slice it starts)
                else {
                    const DecMail m = s_mail[nhand & (kDecMailSlots - 1)];
                    __threadfence_block();
                    *(volatile int *)&s_ack = nhand + 1;
                    sr.rc.low = m.low; sr.rc.range = m.range; sr.rc.ptr = sbeg + m.pos; sr.err = m.err;
                    if (golomb) {                                // the bit reader behind the luma plane
                        sr.br.buf = sbeg + m.br_off; sr.br.nbytes = m.br_nbytes; sr.br.pos = m.br_pos;
                        br_load(sr.br, m.br_pos >> 3);
                    }
                    sx = m.sx; sy = m.sy; sw = m.sw; sh = m.sh; v4 = m.v4; bad = m.bad ? kBadSeen : 0;
                    qti[0] = m.qti & 3; qti[1] = (m.qti >> 2) & 3; qti[2] = (m.qti >> 4) & 3;
                }
            } else
            if (T.version < 2) {
                // versions 0/1: one slice = the frame; the host has read the keyframe bit and the in-band header
                // (ffv1dec.c:646-696) and hands over the coder state behind them
                const uint32_t *is = B.init_state + (size_t)f * 3;
                sr.rc.low = is[0]; sr.rc.range = is[1]; sr.rc.ptr = sbeg + is[2];
                sx = 0; sy = 0; sw = T.width; sh = T.height;
                if (sw + 2 * kDecRingPad > T.ring_w) bad = 1;
                if (golomb) {                                                           // ffv1dec.c:427-434 (no state-129 bit before version 3.2)
                    const uint32_t acb = (uint32_t)(sr.rc.ptr - sr.rc.start) - 1u;
                    sr.br.buf = sr.rc.start + acb;
                    sr.br.nbytes = ssize - acb;
                    sr.br.pos = 0;
                    br_load(sr.br, 0);
                }
            } else {
            if (si == 0) { uint8_t ks = 128; rd_get(sr.rc, &ks, s_lut); }              // keyframe bit (ffv1dec.c:924)
            uint8_t st[32];
            for (int i = 0; i < 32; i++) st[i] = 128;
            const uint32_t ux = (uint32_t)rd_symbol(sr.rc, st, s_lut, false, sr.err) * (uint32_t)T.width;
            const uint32_t uy = (uint32_t)rd_symbol(sr.rc, st, s_lut, false, sr.err) * (uint32_t)T.height;
            const uint32_t uw = ((uint32_t)rd_symbol(sr.rc, st, s_lut, false, sr.err) + 1u) * (uint32_t)T.width + ux;
            const uint32_t uh = ((uint32_t)rd_symbol(sr.rc, st, s_lut, false, sr.err) + 1u) * (uint32_t)T.height + uy;
            const uint32_t x0 = ux / T.num_h_slices, y0 = uy / T.num_v_slices;
            const uint32_t w0 = uw / T.num_h_slices - x0, h0 = uh / T.num_v_slices - y0;
            if (w0 > (uint32_t)T.width || h0 > (uint32_t)T.height || (unsigned long long)x0 + w0 > (uint32_t)T.width ||
                (unsigned long long)y0 + h0 > (uint32_t)T.height || w0 == 0 || h0 == 0 || sr.err) bad = 1;
            sx = (int)x0; sy = (int)y0; sw = (int)w0; sh = (int)h0;
            for (int k = 0; k < T.plane_count && !bad; k++) {
                const int idx = rd_symbol(sr.rc, st, s_lut, false, sr.err);
                if ((unsigned)idx >= 2u) bad = 1; else if (k < 3) qti[k] = idx;
            }
            if (!bad) {
                rd_symbol(sr.rc, st, s_lut, false, sr.err);      // picture structure
                rd_symbol(sr.rc, st, s_lut, false, sr.err);      // sample aspect ratio num
                rd_symbol(sr.rc, st, s_lut, false, sr.err);      // sample aspect ratio den
                if (T.version > 3) {                             // ffv1dec.c:345-356
                    const int reset = rd_get(sr.rc, st, s_lut);
                    const int mode = rd_symbol(sr.rc, st, s_lut, false, sr.err);
                    int by = 1, ry = 1;
                    if (mode != 1) {
                        by = rd_symbol(sr.rc, st, s_lut, false, sr.err);
                        ry = rd_symbol(sr.rc, st, s_lut, false, sr.err);
                        if ((unsigned long long)(unsigned)by + (unsigned)ry > 4ull) bad = 1;   // "slice_rct_y_coef out of range"
                    }
                    if ((unsigned)mode > 255u || (mode == 1 && golomb)) bad = 1;               // PCM slices are range coded (ffv1enc.c:1209)
                    v4 = (by & 15) | (ry & 15) << 4 | (mode & 255) << 8 | reset << 16;
                }
            }
            if (sw + 2 * kDecRingPad > T.ring_w) bad = 1;
            if (!bad && golomb) {                                                       // ffv1dec.c:427-434
                if (T.micro_version > 1 || T.version > 3) { uint8_t s129 = 129; rd_get(sr.rc, &s129, s_lut); }
                const uint32_t acb = (uint32_t)(sr.rc.ptr - sr.rc.start) - 1u;
                sr.br.buf = sr.rc.start + acb;
                sr.br.nbytes = ssize - acb;
                sr.br.pos = 0;
                br_load(sr.br, 0);
            }
            }
        }
        bad = __shfl_sync(0xFFFFFFFFu, bad, 0);
        if (PIPE) nhand++;
        if (bad) {
            if (B.zero_fill) dec_zero_slice(T, B.out + (size_t)f * T.frame_bytes, si, zsel, lane);
            if (lane == 0) {
                if (bad != kBadSeen) B.damaged[fs] |= 2u;
                if (PIPE && warp == 0) {
                    volatile int *ack = &s_ack;
                    while (nhand - 1 - *ack >= kDecMailSlots) __nanosleep(200);
                    DecMail m{}; m.bad = 1;
                    s_mail[(nhand - 1) & (kDecMailSlots - 1)] = m;
                    __threadfence_block();
                    *(volatile int *)&s_pub = nhand;
                }
            }
            continue;
        }
        sx = __shfl_sync(0xFFFFFFFFu, sx, 0); sy = __shfl_sync(0xFFFFFFFFu, sy, 0);
        sw = __shfl_sync(0xFFFFFFFFu, sw, 0); sh = __shfl_sync(0xFFFFFFFFu, sh, 0);
        qti[0] = __shfl_sync(0xFFFFFFFFu, qti[0], 0); qti[1] = __shfl_sync(0xFFFFFFFFu, qti[1], 0);
        qti[2] = __shfl_sync(0xFFFFFFFFu, qti[2], 0);
        v4 = __shfl_sync(0xFFFFFFFFu, v4, 0);
        if (B.zero_fill) {
            // a slice header may announce any rectangle: what it leaves of the grid's rectangle is cleared first
            const int gx = si % T.num_h_slices, gy = si / T.num_h_slices;
            const int nx0 = T.width * gx / T.num_h_slices, nx1 = T.width * (gx + 1) / T.num_h_slices;
            const int ny0 = T.height * gy / T.num_v_slices, ny1 = T.height * (gy + 1) / T.num_v_slices;
            if (sx != nx0 || sy != ny0 || sw != nx1 - nx0 || sh != ny1 - ny0) {
                dec_zero_slice(T, B.out + (size_t)f * T.frame_bytes, si, zsel, lane);
                __syncwarp();
            }
        }
        const int rct_by = v4 & 15, rct_ry = (v4 >> 4) & 15, coding_mode = (v4 >> 8) & 255;
        const bool pcm = coding_mode == 1;

        // ---- keyframe (or a version-4 slice that asks for it, ffv1dec.c:419-420): reset the models (ffv1.c:177-202)
        if (key || (v4 >> 16)) {
            for (int pc = 0; pc < 3; pc++) {
                if (PIPE && (pc == 0) != (warp == 0)) continue;  // every warp resets the models it decodes with
                uint8_t *m = models + (size_t)pc * T.state_stride;
                const int set = qti[pc < T.plane_count ? pc : 0];
                const int nctx = T.ctx_count[set];
                if (!golomb) {
                    uint32_t *m4 = reinterpret_cast<uint32_t *>(m);
                    const uint32_t *i4 = reinterpret_cast<const uint32_t *>(T.model_init[set]);      // ffv1.c:188-190
                    for (int i = lane; i < nctx * 8; i += 32) m4[i] = i4 ? i4[i] : 0x80808080u;
                } else {
                    uint2 *m8 = reinterpret_cast<uint2 *>(m);
                    // VlcState {drift 0, error_sum 4, bias 0, count 1}
                    for (int i = lane; i < nctx; i += 32) m8[i] = make_uint2(0x00040000u, 0x00000100u);
                }
            }
        }
        __syncwarp();

        uint8_t *frame = B.out + (size_t)f * T.frame_bytes;
        if (!T.colorspace) {
            // ---- decode_plane per plane (ffv1dec.c:183-224, 436-455)
            for (int p = luma_warp ? 0 : 1; p < (rest_warp ? nplanes : 1); p++) {
                int src, hs = 0, vs = 0, pc, pstep = T.bits > 8 ? 2 : 1, poff = 0;
                if (T.ya8) { src = 0; pc = p; pstep = 2; poff = p; }
                else if (p == 0) { src = 0; pc = 0; }
                else if (T.chroma_planes && p <= 2) { src = p; hs = T.hshift; vs = T.vshift; pc = 1; }
                else { src = 3; pc = 2; }
                const int w = -((-sw) >> hs), h = -((-sh) >> vs);
                const int px0 = sx >> hs, py0 = sy >> vs;
                const int16_t *q = s_quant + qti[pc < T.plane_count ? pc : 0] * 5 * 256;
                uint8_t *model = models + (size_t)pc * T.state_stride;
                // the plane context's model moves to shared memory while its planes are decoded (U and V share one)
                // (the slice header picks the quantisation table set per plane: only a set whose model fits is moved)
                const int model_nb = (T.ctx_count[qti[pc < T.plane_count ? pc : 0]] * (golomb ? 8 : 32) + 15) >> 4;
                const bool model_sm = T.smem_model && model_nb * 16 <= T.smem_model && (!PIPE || warp == 0);
                if (model_sm) {
                    if (pc != sm_pc) {
                        const int nb = model_nb;
                        if (sm_pc >= 0) {
                            uint4 *dst = reinterpret_cast<uint4 *>(models + (size_t)sm_pc * T.state_stride);
                            for (int i = lane; i < sm_nb; i += 32) dst[i] = reinterpret_cast<const uint4 *>(s_model)[i];
                        }
                        __syncwarp();
                        const uint4 *src = reinterpret_cast<const uint4 *>(model);
                        for (int i = lane; i < nb; i += 32) reinterpret_cast<uint4 *>(s_model)[i] = src[i];
                        sm_pc = pc; sm_nb = nb;
                    }
                    model = s_model;
                }
                const bool ring_sm = T.smem_ring_w && w + 2 * kDecRingPad <= T.smem_ring_w;
                const int rw = ring_sm ? T.smem_ring_w : T.ring_w;
                if (!golomb && ring_sm && !pcm) {
                    const bool five = q[3 * 256 + 127] != 0;
                    LineRec *rec = reinterpret_cast<LineRec *>(s_ring);
                    for (int x = lane; x < w; x += 32) rec[x] = LineRec{0, 0, 0, 0};
                    __syncwarp();
                    int topm1 = 0;                                                       // top[-1] of the line being decoded
                    for (int y = 0; y < h; y++) {
                        const int cm1 = rec[0].top;
                        for (int x = lane; x < w; x += 32) {
                            const int Tp = rec[x].top, LTv = x ? rec[x - 1].top : topm1, RTv = x + 1 < w ? rec[x + 1].top : Tp;
                            int q12 = q[256 + ((LTv - Tp) & 0xFF)] + q[512 + ((Tp - RTv) & 0xFF)];
                            if (five) q12 += q[1024 + ((rec[x].top2 - Tp) & 0xFF)];
                            rec[x].q12 = (int16_t)q12;
                        }
                        __syncwarp();
#ifdef FFV1_DEC_PROBE
                        const long long pt0 = clock64();
#endif
                        if (lane == 0) {
                            if (five) dec_line_rec<true>(sr, model_sm ? s_model : model, q, s_lut, rec, w, bits, topm1);
                            else if (model_sm) dec_line_rec<false>(sr, s_model, q, s_lut, rec, w, bits, topm1);
                            else dec_line_rec<false>(sr, model, q, s_lut, rec, w, bits, topm1);
                        }
#ifdef FFV1_DEC_PROBE
                        probe_serial += clock64() - pt0; probe_samples += w;
#endif
                        topm1 = cm1;                                                     // the next line's top[-1] is this line's cur[-1]
                        __syncwarp();
                        uint8_t *dst = frame + T.plane_off[src] + (size_t)(py0 + y) * T.plane_pitch[src];
                        if (T.bits <= 8) {
                            for (int x = lane; x < w; x += 32) {
                                const int16_t v = rec[x].cur;
                                dst[(px0 + x) * pstep + poff] = (uint8_t)v;
                                rec[x].top2 = rec[x].top;
                                rec[x].top = v;
                            }
                        } else {
                            const int shl = T.packed_at_lsb ? 0 : 16 - T.bits;          // ffv1dec.c:211-219
                            uint16_t *d16 = reinterpret_cast<uint16_t *>(dst) + px0;
                            for (int x = lane; x < w; x += 32) {
                                const int16_t v = rec[x].cur;
                                d16[x] = (uint16_t)((uint16_t)v << shl);
                                rec[x].top2 = rec[x].top;
                                rec[x].top = v;
                            }
                        }
                        __syncwarp();
                    }
                    continue;
                }
                int16_t *rows = ring_sm ? s_ring + kDecRingPad : ring + (PIPE && warp ? 3 * T.ring_w : 0);   // plane slot 0 (1): three rows
                for (int i = lane - kDecRingPad; i < 3 * rw - kDecRingPad; i += 32) rows[i] = 0;
                __syncwarp();
                if (lane == 0) sr.run_index = 0;
                for (int y = 0; y < h; y++) {
                    int16_t *cur = rows + (y % 3) * rw;
                    int16_t *top = rows + ((y + 2) % 3) * rw;
                    int16_t *top2 = rows + ((y + 1) % 3) * rw;
                    if (lane == 0) {
                        cur[-1] = top[0];                 // ffv1dec.c:199-200
                        top[w] = top[w - 1];
                        if (pcm)
                            dec_line_pcm(sr, s_lut, cur, w, bits);
                        else if (golomb)
                            dec_line(sr, model, q, s_lut, cur, top, top2, w, bits);
                        else if (model_sm && ring_sm) {
                            // the common case, spelled out on the shared-memory arrays so that the loads are LDS
                            int16_t *r0 = s_ring + kDecRingPad;
                            int16_t *c1 = r0 + (y % 3) * T.smem_ring_w;
                            const int16_t *t1 = r0 + ((y + 2) % 3) * T.smem_ring_w, *t2 = r0 + ((y + 1) % 3) * T.smem_ring_w;
                            if (q[3 * 256 + 127]) dec_line_rc<true>(sr, s_model, q, s_lut, c1, t1, t2, w, bits);
                            else dec_line_rc<false>(sr, s_model, q, s_lut, c1, t1, t2, w, bits);
                        } else if (q[3 * 256 + 127])
                            dec_line_rc<true>(sr, model, q, s_lut, cur, top, top2, w, bits);
                        else
                            dec_line_rc<false>(sr, model, q, s_lut, cur, top, top2, w, bits);
                    }
                    __syncwarp();
                    uint8_t *dst = frame + T.plane_off[src] + (size_t)(py0 + y) * T.plane_pitch[src];
                    if (T.bits <= 8) {
                        for (int x = lane; x < w; x += 32) dst[(px0 + x) * pstep + poff] = (uint8_t)cur[x];
                    } else {
                        const int shl = T.packed_at_lsb ? 0 : 16 - T.bits;              // ffv1dec.c:211-219
                        uint16_t *d16 = reinterpret_cast<uint16_t *>(dst) + px0;
                        for (int x = lane; x < w; x += 32) d16[x] = (uint16_t)((uint16_t)cur[x] << shl);
                    }
                    __syncwarp();
                }
            }
            // the model goes back to global memory at the end of the slice (the keyframe reset and a different slice
            // geometry in the next frame work on the global copy)
            if (sm_pc >= 0) {
                uint4 *dst = reinterpret_cast<uint4 *>(models + (size_t)sm_pc * T.state_stride);
                for (int i = lane; i < sm_nb; i += 32) dst[i] = reinterpret_cast<const uint4 *>(s_model)[i];
                sm_pc = -1;
                __syncwarp();
            }
            if (PIPE && warp == 0) {
                // the luma plane is done: coder and header go to warp 1, this warp moves on to the next frame
                if (lane == 0) {
                    volatile int *ack = &s_ack;
                    while (nhand - 1 - *ack >= kDecMailSlots) __nanosleep(200);
                    DecMail m;
                    m.low = sr.rc.low; m.range = sr.rc.range; m.pos = (uint32_t)(sr.rc.ptr - sbeg); m.err = sr.err;
                    m.sx = sx; m.sy = sy; m.sw = sw; m.sh = sh; m.v4 = v4; m.bad = 0;
                    m.qti = qti[0] | qti[1] << 2 | qti[2] << 4;
                    m.br_off = golomb ? (uint32_t)(sr.br.buf - sbeg) : 0u; m.br_nbytes = sr.br.nbytes; m.br_pos = sr.br.pos;
                    s_mail[(nhand - 1) & (kDecMailSlots - 1)] = m;
                    __threadfence_block();
                    *(volatile int *)&s_pub = nhand;
                }
                __syncwarp();
                continue;
            }
        } else {
            // ---- decode_rgb_frame (ffv1dec.c:226-280): planes interleaved per row, shared run_index
            for (int i = lane - kDecRingPad; i < 4 * 3 * T.ring_w - kDecRingPad; i += 32) ring[i] = 0;
            __syncwarp();
            if (lane == 0) sr.run_index = 0;
            const int offset = 1 << (T.bits <= 8 ? 8 : T.bits);
            for (int y = 0; y < sh; y++) {
                if (lane == 0) {
                    for (int p = 0; p < nplanes; p++) {
                        int16_t *rows = ring + p * 3 * T.ring_w;
                        int16_t *cur = rows + (y % 3) * T.ring_w;
                        int16_t *top = rows + ((y + 2) % 3) * T.ring_w;
                        int16_t *top2 = rows + ((y + 1) % 3) * T.ring_w;
                        const int pc = (p + 1) / 2;
                        cur[-1] = top[0];
                        top[sw] = top[sw - 1];
                        uint8_t *model = models + (size_t)pc * T.state_stride;
                        const int16_t *q = s_quant + qti[pc < T.plane_count ? pc : 0] * 5 * 256;
                        if (pcm) dec_line_pcm(sr, s_lut, cur, sw, bits - 1);                    // ffv1dec.c:255: no extra bit without the RCT
                        else if (golomb) dec_line(sr, model, q, s_lut, cur, top, top2, sw, bits);
                        else if (q[3 * 256 + 127]) dec_line_rc<true>(sr, model, q, s_lut, cur, top, top2, sw, bits);
                        else dec_line_rc<false>(sr, model, q, s_lut, cur, top, top2, sw, bits);
                    }
                }
                __syncwarp();
                const int16_t *gr = ring + (y % 3) * T.ring_w, *br = gr + 3 * T.ring_w, *rr = gr + 6 * T.ring_w, *ar = gr + 9 * T.ring_w;
                for (int x = lane; x < sw; x += 32) {
                    int g = gr[x], b = br[x], r = rr[x];
                    const int a = nplanes == 4 ? ar[x] : 0;
                    if (!pcm) {                                                               // ffv1dec.c:263-269
                        b -= offset; r -= offset;
                        g -= (b * rct_by + r * rct_ry) >> 2;
                        b += g; r += g;
                    }
                    if (T.rgb32) {
                        uint32_t *d = reinterpret_cast<uint32_t *>(frame + T.plane_off[0] + (size_t)(sy + y) * T.plane_pitch[0]) + sx + x;
                        *d = (uint32_t)(b & 0xFF) | ((uint32_t)(g & 0xFF) << 8) | ((uint32_t)(r & 0xFF) << 16) | ((uint32_t)(a & 0xFF) << 24);
                    } else {
                        // planar RGB keeps the reference's naming quirk: data[0] <- b, data[1] <- g, data[2] <- r (ffv1dec.c:274-276)
                        reinterpret_cast<uint16_t *>(frame + T.plane_off[0] + (size_t)(sy + y) * T.plane_pitch[0])[sx + x] = (uint16_t)b;
                        reinterpret_cast<uint16_t *>(frame + T.plane_off[1] + (size_t)(sy + y) * T.plane_pitch[1])[sx + x] = (uint16_t)g;
                        reinterpret_cast<uint16_t *>(frame + T.plane_off[2] + (size_t)(sy + y) * T.plane_pitch[2])[sx + x] = (uint16_t)r;
                    }
                }
                __syncwarp();
            }
        }
        // ---- end-of-slice check (ffv1dec.c:459-467)
        if (lane == 0) {
            uint32_t flag = sr.err ? 2u : 0u;
            if (!golomb && T.version > 2) {
                uint8_t s129 = 129;
                rd_get(sr.rc, &s129, s_lut);
                const long long v = (long long)(sr.rc.end - sr.rc.ptr) - 2 - 5 * T.ec;
                if (v) flag |= 2u;
            }
            if (flag) B.damaged[fs] |= flag;
        }
        __syncwarp();
    }
#ifdef FFV1_DEC_PROBE
    if (lane == 0 && (chain == 0 || chain == 101))
        printf("chain %d: %lld samples, serial %lld cycles (%.1f per sample), total %lld cycles\n", chain, probe_samples, probe_serial,
               (double)probe_serial / (double)max(probe_samples, 1ll), clock64() - probe_t0);
#endif
}

void launch_decode(const DecDeviceTables &t_in, const DecBatch &b, cudaStream_t s)
{
    DecDeviceTables t = t_in;
    {
        const char *e = getenv("FFV1B200_DEC_SMEM");
        if (e && atoi(e) == 0) t.smem_model = 0;
    }
    const int chains = b.nseg * t.max_slices;
    const bool debug = getenv("FFV1B200_DEBUG") != nullptr;
    static int nsm = -1;
    if (nsm < 0) {
        cudaFuncSetAttribute(k_decode<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDecWarps * (24 * 1024 + 11 * 1024));
        cudaFuncSetAttribute(k_decode<12, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDecWarps * (24 * 1024 + 11 * 1024));
        cudaFuncSetAttribute(k_decode<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 1024 + kDecWarps * 11 * 1024);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    }
    // Planar YUV, few enough chains for one CTA each: two warps per chain, luma one frame ahead of
    // chroma (k_decode<.., PIPE>).  FFV1B200_DEC_PIPE=0/1 overrides the choice (1: whenever the stream allows it).
    if (t.smem_ring_w && !t.colorspace && t.chroma_planes && !t.ya8 && b.nframes >= 2 * b.nseg) {      // (intra-only: nothing to overlap)
        // Measured and dropped: the second warp's models in shared memory as well (no faster -- the luma warp sets the pace --
        // and fewer CTAs per SM); for the small context model, every model in global memory with the 80-register build for
        // batches of up to 1776 chains (768 frames: 1044 against 1062 frames/s with one warp per chain, 1024 frames: 1062
        // against 1186).  The large context model (no model fits shared memory: t.smem_model == 0) comes with few, long
        // chains (4 slices per frame): there the second warp is pure gain.
        const int smem = t.smem_model + kDecWarps * t.smem_ring_w * kDecSmemRingBytes;
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_decode<8, true>, 32 * kDecWarps, smem);
        const char *ev = getenv("FFV1B200_DEC_PIPE");
        if (ev ? atoi(ev) != 0 : chains <= nb * nsm) {
            if (debug) fprintf(stderr, "k_decode<8, pipe>: grid %d, %d B dynamic smem, %d CTAs per SM\n", chains, smem, nb);
            k_decode<8, true><<<chains, 32 * kDecWarps, smem, s>>>(t, b);
            return;
        }
    }
    // Models in shared memory cap the chains an SM holds (a 21 KB model per chain); a batch with more chains than that
    // runs faster with the models in global memory and twice the warps per scheduler to hide its latency.
    if (t.smem_model) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_decode<8, false>, 32 * kDecWarps, kDecWarps * (t.smem_model + t.smem_ring_w * kDecSmemRingBytes));
        if (chains > nb * nsm * kDecWarps) t.smem_model = 0;
    }
    const int smem = kDecWarps * (t.smem_model + t.smem_ring_w * kDecSmemRingBytes);
    const int grid = (chains + kDecWarps - 1) / kDecWarps;
    // more chains than the spill-free register allocation keeps resident, but few enough for the 80-register one to
    // hold in a single wave: take that one (2048 frames of 1080p: 1421 instead of 1239 frames/s)
    int nb8 = 0, nb12 = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb8, k_decode<8, false>, 32 * kDecWarps, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb12, k_decode<12, false>, 32 * kDecWarps, smem);
    const char *ev = getenv("FFV1B200_DEC_MINB");
    const bool dense = ev ? atoi(ev) > 8 : (grid > nb8 * nsm && grid <= nb12 * nsm);
    if (debug) fprintf(stderr, "k_decode<%d>: grid %d, %d B dynamic smem\n", dense ? 12 : 8, grid, smem);
    if (dense) k_decode<12, false><<<grid, 32 * kDecWarps, smem, s>>>(t, b);
    else k_decode<8, false><<<grid, 32 * kDecWarps, smem, s>>>(t, b);
}

// ------------------------------------------------------------------------------------------------ concealment
__global__ void __launch_bounds__(256) k_conceal(const DecDeviceTables T, const DecBatch B, const int f)
{
    const int si = blockIdx.x;
    if (si >= B.slice_count[f] || !B.damaged[f * T.max_slices + si]) return;
    const uint8_t *prev = f > 0 ? B.out + (size_t)(f - 1) * T.frame_bytes : B.prev_frame;
    if (!prev) return;
    uint8_t *cur = B.out + (size_t)f * T.frame_bytes;
    // the reference copies the rectangle of the slice grid position (ffv1dec.c:1006-1018)
    const int gx = si % T.num_h_slices, gy = si / T.num_h_slices;
    const int x0 = T.width * gx / T.num_h_slices, x1 = T.width * (gx + 1) / T.num_h_slices;
    const int y0 = T.height * gy / T.num_v_slices, y1 = T.height * (gy + 1) / T.num_v_slices;
    const int nsrc = T.colorspace ? (T.rgb32 ? 1 : 3) : (T.ya8 ? 1 : 1 + 2 * T.chroma_planes + T.transparency);
    for (int p = 0; p < nsrc; p++) {
        const bool chroma = !T.colorspace && !T.ya8 && T.chroma_planes && (p == 1 || p == 2);
        const int hs = chroma ? T.hshift : 0, vs = chroma ? T.vshift : 0;
        const int src_plane = (!T.colorspace && !T.ya8 && !T.chroma_planes && p == 1) ? 3 : p;
        const int bpp = T.rgb32 ? 4 : (T.ya8 ? 2 : (T.bits > 8 ? 2 : 1));
        const int bx0 = (x0 >> hs) * bpp, bx1 = (-((-x1) >> hs)) * bpp;
        const int ry0 = y0 >> vs, ry1 = -((-y1) >> vs);
        const int rb = bx1 - bx0;
        for (int i = threadIdx.x; i < rb * (ry1 - ry0); i += blockDim.x) {
            const int y = ry0 + i / rb, x = bx0 + i % rb;
            const size_t o = (size_t)T.plane_off[src_plane] + (size_t)y * T.plane_pitch[src_plane] + x;
            cur[o] = prev[o];
        }
    }
}

void launch_conceal(const DecDeviceTables &t, const DecBatch &b, int frame, cudaStream_t s)
{
    k_conceal<<<t.max_slices, 256, 0, s>>>(t, b, frame);
}

} // namespace ffv1
