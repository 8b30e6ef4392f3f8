/*
 * ffv1_oracle.c -- TEST INFRASTRUCTURE ONLY (see ffv1_oracle.h).
 *
 * A from-scratch, scalar restatement of the reference FFV1 (versions 0, 1 and 3) bitstream
 * algorithm.  It is written for clarity, not speed: samples of a slice plane are first gathered
 * into a w*h int16 array and every neighbour is fetched with the slice-local edge rules spelled
 * out explicitly (the reference gets the same values implicitly from its ring buffer).
 * Every function cites the reference lines (relative to /root/reference/libavcodec) it follows.
 */
#include "ffv1_oracle.h"
#include <stdlib.h>
#include <string.h>

#define E_INVAL   (-22)
#define E_NOSYS   (-38)
#define E_NOMEM   (-12)
#define E_INVALIDDATA (-1094995529)   /* AVERROR_INVALIDDATA = -MKTAG('I','N','D','A') */

/* ------------------------------------------------------------------------------------------
 * constant tables of the format
 * ---------------------------------------------------------------------------------------- */

/* The four context quantisation curves (ffv1enc.c:44-118) are step functions on |d|, mirrored for
 * negative d (q[256-i] = -q[i], q[128] = -q[127]); they are stored here as run lengths per level. */
static void build_quant_curve(int8_t q[256], const int *runs, int nruns)
{
    int i = 0, level, k;
    for (level = 0; level < nruns; level++)
        for (k = 0; k < runs[level]; k++)
            q[i++] = (int8_t)level;
    for (i = 1; i < 128; i++)
        q[256 - i] = (int8_t)-q[i];
    q[128] = (int8_t)-q[127];
}
static const int runs_quant11[]      = { 1, 1, 3, 7, 23, 93 };
static const int runs_quant5[]       = { 1, 3, 124 };
static const int runs_quant9_10bit[] = { 5, 8, 14, 29, 72 };
static const int runs_quant5_10bit[] = { 11, 39, 78 };

/* custom state-transition table used by coder=1/2 ("ver2_state", ffv1enc.c:120-137) */
static const uint8_t custom_one_state[256] = {
  0, 10, 10, 10, 10, 16, 16, 16, 28, 16, 16, 29, 42, 49, 20, 49, 59, 25, 26, 26, 27, 31, 33, 33, 33, 34, 34,
  37, 67, 38, 39, 39, 40, 40, 41, 79, 43, 44, 45, 45, 48, 48, 64, 50, 51, 52, 88, 52, 53, 74, 55, 57, 58, 58,
  74, 60, 101, 61, 62, 84, 66, 66, 68, 69, 87, 82, 71, 97, 73, 73, 82, 75, 111, 77, 94, 78, 87, 81, 83, 97, 85,
  83, 94, 86, 99, 89, 90, 99, 111, 92, 93, 134, 95, 98, 105, 98, 105, 110, 102, 108, 102, 118, 103, 106, 106,
  113, 109, 112, 114, 112, 116, 125, 115, 116, 117, 117, 126, 119, 125, 121, 121, 123, 145, 124, 126, 131, 127,
  129, 165, 130, 132, 138, 133, 135, 145, 136, 137, 139, 146, 141, 143, 142, 144, 148, 147, 155, 151, 149, 151,
  150, 152, 157, 153, 154, 156, 168, 158, 162, 161, 160, 172, 163, 169, 164, 166, 184, 167, 170, 177, 174, 171,
  173, 182, 176, 180, 178, 175, 189, 179, 181, 186, 183, 192, 185, 200, 187, 191, 188, 190, 197, 193, 196, 197,
  194, 195, 196, 198, 202, 199, 201, 210, 203, 207, 204, 205, 206, 208, 214, 209, 211, 221, 212, 213, 215, 224,
  216, 217, 218, 219, 220, 222, 228, 223, 225, 226, 224, 227, 229, 240, 230, 231, 232, 233, 234, 235, 236, 238,
  239, 237, 242, 241, 243, 242, 244, 245, 246, 247, 248, 249, 250, 251, 252, 252, 253, 254, 255,
};

/* run-length exponent table for golomb run mode (bitstream.c:40-47) */
static const uint8_t log2_run[41] = {
    0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7,
    8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24,
};

/* rangecoder.c:63-101 (ff_build_rac_states) specialised to how FFV1 calls it:
 * factor = (int)(0.05 * 2^32), max_p = 248.  Pure 64-bit integer arithmetic. */
void ffv1o_default_state_tables(uint8_t zero_state[256], uint8_t one_state[256])
{
    const int64_t one = (int64_t)1 << 32;
    const int64_t factor = (int64_t)(int)(0.05 * (double)one);
    const int max_p = 248;
    int64_t p = one / 2;
    int prev = 0, i;

    memset(zero_state, 0, 256);
    memset(one_state, 0, 256);
    for (i = 0; i < 128; i++) {
        int p8 = (int)((256 * p + one / 2) >> 32);
        if (p8 <= prev)
            p8 = prev + 1;
        if (prev && prev < 256 && p8 <= max_p)
            one_state[prev] = (uint8_t)p8;
        p += ((one - p) * factor + one / 2) >> 32;
        prev = p8;
    }
    for (i = 256 - max_p; i <= max_p; i++) {
        int p8;
        if (one_state[i])
            continue;
        p  = (i * one + 128) >> 8;
        p += ((one - p) * factor + one / 2) >> 32;
        p8 = (int)((256 * p + one / 2) >> 32);
        if (p8 <= i)
            p8 = i + 1;
        if (p8 > max_p)
            p8 = max_p;
        one_state[i] = (uint8_t)p8;
    }
    for (i = 1; i < 255; i++)
        zero_state[i] = (uint8_t)(256 - one_state[256 - i]);
}

/* libavutil/crc.c:357-380 with the AV_CRC_32_IEEE table (crc.c:303): plain MSB-first CRC-32,
 * polynomial 0x04C11DB7, no reflection, no final xor.  (libavutil keeps the register byte-swapped;
 * the bytes it stores with AV_WL32 equal the big-endian form of this value.) */
uint32_t ffv1o_crc32(uint32_t crc, const uint8_t *buf, size_t len)
{
    static uint32_t table[256];
    static int ready;
    size_t i;
    if (!ready) {
        int n, k;
        for (n = 0; n < 256; n++) {
            uint32_t c = (uint32_t)n << 24;
            for (k = 0; k < 8; k++)
                c = (c << 1) ^ ((c & 0x80000000u) ? 0x04C11DB7u : 0);
            table[n] = c;
        }
        ready = 1;
    }
    for (i = 0; i < len; i++)
        crc = (crc << 8) ^ table[(crc >> 24) ^ buf[i]];
    return crc;
}

/* ------------------------------------------------------------------------------------------
 * range coder, encoder side (rangecoder.h:52-102, rangecoder.c:42-51,104-116)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int low, range, outstanding_count, outstanding_byte;
    uint8_t *buf;
    long pos, cap;
    int overflow;
    uint8_t zero_state[256], one_state[256];
    uint64_t decisions;
} RcEnc;

static void rce_init(RcEnc *c, uint8_t *buf, long cap)
{
    c->low = 0;
    c->range = 0xFF00;
    c->outstanding_count = 0;
    c->outstanding_byte = -1;
    c->buf = buf; c->pos = 0; c->cap = cap; c->overflow = 0;
}
static void rce_byte(RcEnc *c, int b)
{
    if (c->pos < c->cap) c->buf[c->pos] = (uint8_t)b; else c->overflow = 1;
    c->pos++;
}
static void rce_renorm(RcEnc *c)
{
    while (c->range < 0x100) {
        if (c->outstanding_byte < 0) {
            c->outstanding_byte = c->low >> 8;
        } else if (c->low <= 0xFF00) {
            rce_byte(c, c->outstanding_byte);
            for (; c->outstanding_count; c->outstanding_count--) rce_byte(c, 0xFF);
            c->outstanding_byte = c->low >> 8;
        } else if (c->low >= 0x10000) {
            rce_byte(c, c->outstanding_byte + 1);
            for (; c->outstanding_count; c->outstanding_count--) rce_byte(c, 0x00);
            c->outstanding_byte = (c->low >> 8) & 0xFF;
        } else {
            c->outstanding_count++;
        }
        c->low = (c->low & 0xFF) << 8;
        c->range <<= 8;
    }
}
static void rce_put(RcEnc *c, uint8_t *state, int bit)
{
    int range1 = (c->range * (*state)) >> 8;
    if (!bit) {
        c->range -= range1;
        *state = c->zero_state[*state];
    } else {
        c->low += c->range - range1;
        c->range = range1;
        *state = c->one_state[*state];
    }
    c->decisions++;
    rce_renorm(c);
}
static long rce_terminate(RcEnc *c)
{
    c->range = 0xFF;
    c->low += 0xFF;
    rce_renorm(c);
    c->range = 0xFF;
    rce_renorm(c);
    return c->pos;
}
static int ilog2(unsigned v) { int n = 0; while (v >>= 1) n++; return n; }

/* ffv1enc.c:185-231: zero flag, unary exponent, mantissa, sign on a 32-byte state */
static void rce_symbol(RcEnc *c, uint8_t *state, int v, int is_signed)
{
    if (!v) { rce_put(c, state + 0, 1); return; }
    {
        const int a = v < 0 ? -v : v;
        const int e = ilog2((unsigned)a);
        int i;
        rce_put(c, state + 0, 0);
        for (i = 0; i < e; i++)
            rce_put(c, state + 1 + (i < 9 ? i : 9), 1);
        rce_put(c, state + 1 + (e < 9 ? e : 9), 0);
        for (i = e - 1; i >= 0; i--)
            rce_put(c, state + 22 + (i < 9 ? i : 9), (a >> i) & 1);
        if (is_signed)
            rce_put(c, state + 11 + (e < 10 ? e : 10), v < 0);
    }
}

/* ------------------------------------------------------------------------------------------
 * range coder, decoder side (rangecoder.h:104-145, rangecoder.c:53-61)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int low, range;
    const uint8_t *start, *ptr, *end;
    uint8_t zero_state[256], one_state[256];
} RcDec;

static void rcd_init(RcDec *c, const uint8_t *buf, long size)
{
    c->start = buf; c->end = buf + size;
    c->range = 0xFF00;
    c->low = (size >= 2) ? (buf[0] << 8 | buf[1]) : (size == 1 ? buf[0] << 8 : 0);
    c->ptr = buf + 2;
}
static int rcd_get(RcDec *c, uint8_t *state)
{
    int range1 = (c->range * (*state)) >> 8;
    int bit;
    c->range -= range1;
    if (c->low < c->range) {
        *state = c->zero_state[*state];
        bit = 0;
    } else {
        c->low -= c->range;
        *state = c->one_state[*state];
        c->range = range1;
        bit = 1;
    }
    if (c->range < 0x100) {
        c->range <<= 8;
        c->low <<= 8;
        if (c->ptr < c->end)
            c->low += c->ptr[0];
        c->ptr++;
    }
    return bit;
}
/* ffv1dec.c:42-63 */
static int rcd_symbol(RcDec *c, uint8_t *state, int is_signed, int *err)
{
    int e = 0, a = 1, i;
    if (rcd_get(c, state + 0))
        return 0;
    while (rcd_get(c, state + 1 + (e < 9 ? e : 9))) {
        e++;
        if (e > 31) { if (err) *err = 1; return 0; }
    }
    for (i = e - 1; i >= 0; i--)
        a += a + rcd_get(c, state + 22 + (i < 9 ? i : 9));
    if (is_signed && rcd_get(c, state + 11 + (e < 10 ? e : 10)))
        return -a;
    return a;
}

/* ------------------------------------------------------------------------------------------
 * MSB-first bit writer / reader for the Golomb-Rice path (put_bits.h, get_bits.h)
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint8_t *buf; long cap; uint64_t bitpos; int overflow; } BitW;
static void bw_put(BitW *w, int n, unsigned v)
{
    int i;
    for (i = n - 1; i >= 0; i--) {
        long byte = (long)(w->bitpos >> 3);
        if (byte < w->cap) {
            if (!(w->bitpos & 7)) w->buf[byte] = 0;
            w->buf[byte] |= (uint8_t)(((v >> i) & 1) << (7 - (w->bitpos & 7)));
        } else w->overflow = 1;
        w->bitpos++;
    }
}
typedef struct { const uint8_t *buf; uint64_t nbits, bitpos; } BitR;
static unsigned br_peek32(const BitR *r)
{
    unsigned v = 0; int i;
    uint64_t byte = r->bitpos >> 3;
    uint64_t nbytes = (r->nbits + 7) >> 3;
    uint64_t acc = 0;
    for (i = 0; i < 5; i++) {
        acc <<= 8;
        if (byte + i < nbytes) acc |= r->buf[byte + i];
    }
    v = (unsigned)((acc >> (8 - (r->bitpos & 7))) & 0xFFFFFFFFu);
    return v;
}
static unsigned br_get(BitR *r, int n)
{
    unsigned v;
    if (!n) return 0;
    v = br_peek32(r) >> (32 - n);
    r->bitpos += n;
    return v;
}

/* ------------------------------------------------------------------------------------------
 * parameters / option resolution
 * ---------------------------------------------------------------------------------------- */
typedef struct { const char *name; int colorspace, bits, hs, vs, ncomp, layout, lsb; } PixFmt;
static const PixFmt pixfmts[] = {
    /* name            cs bits hs vs nc layout               packed_at_lsb          ffv1enc.c:721-815 */
    { "yuv420p",       0,  8, 1, 1, 3, FFV1O_LAYOUT_PLANAR, 0 }, { "yuva420p",     0,  8, 1, 1, 4, FFV1O_LAYOUT_PLANAR, 0 },
    { "yuva422p",      0,  8, 1, 0, 4, FFV1O_LAYOUT_PLANAR, 0 }, { "yuv444p",      0,  8, 0, 0, 3, FFV1O_LAYOUT_PLANAR, 0 },
    { "yuva444p",      0,  8, 0, 0, 4, FFV1O_LAYOUT_PLANAR, 0 }, { "yuv440p",      0,  8, 0, 1, 3, FFV1O_LAYOUT_PLANAR, 0 },
    { "yuv422p",       0,  8, 1, 0, 3, FFV1O_LAYOUT_PLANAR, 0 }, { "yuv411p",      0,  8, 2, 0, 3, FFV1O_LAYOUT_PLANAR, 0 },
    { "yuv410p",       0,  8, 2, 2, 3, FFV1O_LAYOUT_PLANAR, 0 },
    { "bgr0",          1,  8, 0, 0, 3, FFV1O_LAYOUT_RGB32,  0 }, { "bgra",         1,  8, 0, 0, 4, FFV1O_LAYOUT_RGB32,  0 },
    { "yuv420p16le",   0, 16, 1, 1, 3, FFV1O_LAYOUT_PLANAR, 0 }, { "yuv422p16le",  0, 16, 1, 0, 3, FFV1O_LAYOUT_PLANAR, 0 },
    { "yuv444p16le",   0, 16, 0, 0, 3, FFV1O_LAYOUT_PLANAR, 0 },
    { "yuv444p9le",    0,  9, 0, 0, 3, FFV1O_LAYOUT_PLANAR, 1 }, { "yuv422p9le",   0,  9, 1, 0, 3, FFV1O_LAYOUT_PLANAR, 1 },
    { "yuv420p9le",    0,  9, 1, 1, 3, FFV1O_LAYOUT_PLANAR, 1 },
    { "yuv420p10le",   0, 10, 1, 1, 3, FFV1O_LAYOUT_PLANAR, 1 }, { "yuv422p10le",  0, 10, 1, 0, 3, FFV1O_LAYOUT_PLANAR, 1 },
    { "yuv444p10le",   0, 10, 0, 0, 3, FFV1O_LAYOUT_PLANAR, 1 },
    { "yuva444p16le",  0, 16, 0, 0, 4, FFV1O_LAYOUT_PLANAR, 0 }, { "yuva422p16le", 0, 16, 1, 0, 4, FFV1O_LAYOUT_PLANAR, 0 },
    { "yuva420p16le",  0, 16, 1, 1, 4, FFV1O_LAYOUT_PLANAR, 0 },
    { "yuva444p10le",  0, 10, 0, 0, 4, FFV1O_LAYOUT_PLANAR, 1 }, { "yuva422p10le", 0, 10, 1, 0, 4, FFV1O_LAYOUT_PLANAR, 1 },
    { "yuva420p10le",  0, 10, 1, 1, 4, FFV1O_LAYOUT_PLANAR, 1 },
    { "yuva444p9le",   0,  9, 0, 0, 4, FFV1O_LAYOUT_PLANAR, 1 }, { "yuva422p9le",  0,  9, 1, 0, 4, FFV1O_LAYOUT_PLANAR, 1 },
    { "yuva420p9le",   0,  9, 1, 1, 4, FFV1O_LAYOUT_PLANAR, 1 },
    { "gray16le",      0, 16, 0, 0, 1, FFV1O_LAYOUT_PLANAR, 0 }, { "gray",         0,  8, 0, 0, 1, FFV1O_LAYOUT_PLANAR, 0 },
    { "gbrp9le",       1,  9, 0, 0, 3, FFV1O_LAYOUT_GBRP,   0 }, { "gbrp10le",     1, 10, 0, 0, 3, FFV1O_LAYOUT_GBRP,   0 },
    { "gbrp12le",      1, 12, 0, 0, 3, FFV1O_LAYOUT_GBRP,   0 }, { "gbrp14le",     1, 14, 0, 0, 3, FFV1O_LAYOUT_GBRP,   0 },
    { "ya8",           0,  8, 0, 0, 2, FFV1O_LAYOUT_YA8,    0 },
};

static void fill_tables(ffv1o_params *p)
{
    /* ffv1enc.c:846-871 */
    int8_t qa[256], qb[256];
    int i;
    if (p->bits <= 8) {
        build_quant_curve(qa, runs_quant11, 6);
        build_quant_curve(qb, runs_quant5, 3);
    } else {
        build_quant_curve(qa, runs_quant9_10bit, 5);
        build_quant_curve(qb, runs_quant5_10bit, 3);
    }
    memset(p->quant_tables, 0, sizeof(p->quant_tables));
    for (i = 0; i < 256; i++) {
        p->quant_tables[0][0][i] = qa[i];
        p->quant_tables[0][1][i] = 11 * qa[i];
        p->quant_tables[0][2][i] = 11 * 11 * qa[i];
        p->quant_tables[1][0][i] = qa[i];
        p->quant_tables[1][1][i] = 11 * qa[i];
        p->quant_tables[1][2][i] = 11 * 11 * qb[i];
        p->quant_tables[1][3][i] = 5 * 11 * 11 * qb[i];
        p->quant_tables[1][4][i] = 5 * 5 * 11 * 11 * qb[i];
    }
    p->context_count[0] = (11 * 11 * 11 + 1) / 2;
    p->context_count[1] = (11 * 11 * 5 * 5 * 5 + 1) / 2;
}

int ffv1o_resolve(ffv1o_params *p, int width, int height, const char *pix_fmt, int gop_size,
                  int level, int coder, int context, int slices, int slicecrc)
{
    return ffv1o_resolve_ex(p, width, height, pix_fmt, gop_size, level, coder, context, slices, slicecrc, 0);
}

int ffv1o_resolve_ex(ffv1o_params *p, int width, int height, const char *pix_fmt, int gop_size,
                     int level, int coder, int context, int slices, int slicecrc, int strict_experimental)
{
    const PixFmt *pf = NULL;
    size_t i;
    int version = 0;
    memset(p, 0, sizeof(*p));
    if (width <= 0 || height <= 0) return E_INVALIDDATA;                   /* ffv1.c:46-47 */
    for (i = 0; i < sizeof(pixfmts) / sizeof(pixfmts[0]); i++)
        if (!strcmp(pixfmts[i].name, pix_fmt)) pf = &pixfmts[i];
    if (!pf) return E_NOSYS;                                               /* ffv1enc.c:816-818 */
    if (context < 0 || context > 1 || coder < -2 || coder > 2 || slicecrc < -1 || slicecrc > 1)
        return E_INVAL;                                                    /* AVOption ranges, ffv1enc.c:1383-1399 */

    /* ffv1enc.c:678-697 */
    if (slices > 1) version = 2;
    if (slices == 0 && level < 0 && width * height > 720 * 576) version = 2;
    if (level <= 0 && version == 2) version = 3;
    if (level >= 0 && level <= 4) {
        if (level < version) return E_INVAL;
        version = level;
    }
    p->ec = slicecrc < 0 ? (version >= 3) : slicecrc;                      /* 699-701 */
    if ((version == 2 || version > 3) && !strict_experimental) return E_INVALIDDATA;   /* 703-706: experimental only */
    if (version == 2) return E_NOSYS;                                      /* the abandoned version-2 bitstream is not restated */

    /* 715-718 */
    if (coder == 1) coder = FFV1O_AC_RANGE_CUSTOM;
    else if (coder == -2) coder = FFV1O_AC_RANGE_DEFAULT;
    else if (coder == -1) coder = FFV1O_AC_GOLOMB;

    p->width = width; p->height = height;
    p->colorspace = pf->colorspace;
    p->bits = pf->bits;
    p->layout = pf->layout;
    p->packed_at_lsb = pf->lsb;
    p->chroma_h_shift = pf->hs;
    p->chroma_v_shift = pf->vs;
    if (pf->colorspace == 0) {
        p->chroma_planes = pf->ncomp < 3 ? 0 : 1;                          /* 769 */
        p->transparency = (pf->ncomp == 4 || pf->ncomp == 2);              /* 771 */
        if (pf->bits > 8) {
            if (coder == FFV1O_AC_GOLOMB) coder = FFV1O_AC_RANGE_CUSTOM;   /* 755-759 */
            if (version < 1) version = 1;                                  /* 760 */
        }
    } else {
        p->chroma_planes = 1;
        p->transparency = (pf->layout == FFV1O_LAYOUT_RGB32 && pf->ncomp == 4);
        if (pf->layout == FFV1O_LAYOUT_GBRP) {
            if (version < 1) version = 1;                                  /* 809 */
            if (coder == FFV1O_AC_GOLOMB) coder = FFV1O_AC_RANGE_CUSTOM;   /* 810-814 */
        }
    }
    p->ac = coder;
    p->version = version;
    p->micro_version = version == 3 ? 4 : (version == 4 ? 2 : 0);          /* 565-569 */
    p->context_model = context;
    p->gop_size = gop_size;
    p->intra = gop_size < 2;                                               /* 610 */

    /* 836-844 */
    if (p->ac == FFV1O_AC_RANGE_CUSTOM) {
        memcpy(p->state_transition, custom_one_state, 256);
        p->state_transition[0] = 0;
    } else {
        uint8_t z[256], o[256];
        ffv1o_default_state_tables(z, o);
        memcpy(p->state_transition, o, 256);
        p->state_transition[0] = 0;
    }
    fill_tables(p);

    p->plane_count = p->transparency ? 3 : 2;                              /* 720, 890-891 */
    if (!p->chroma_planes && version > 3) p->plane_count--;                /* 892-893 */

    p->num_h_slices = p->num_v_slices = 1;
    if (version > 1) {                                                     /* 988-1000 */
        int nv, nh, ok = 0;
        for (nv = (width > 352 || height > 288 || !slices) ? 2 : 1; nv < 9 && !ok; nv++)
            for (nh = nv; nh < 2 * nv; nh++)
                if ((slices == nh * nv && slices <= 64) || !slices) {
                    p->num_h_slices = nh; p->num_v_slices = nv; ok = 1; break;
                }
        if (!ok) return E_NOSYS;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * extradata (ffv1enc.c:545-619 writer, ffv1dec.c:476-636 reader)
 * ---------------------------------------------------------------------------------------- */
static void put_quant_table(RcEnc *c, const int16_t *q)
{
    /* ffv1enc.c:475-488: lengths of the constant runs of q[0..127], each minus one */
    uint8_t st[32];
    int last = 0, i;
    memset(st, 128, sizeof(st));
    for (i = 1; i < 128; i++)
        if (q[i] != q[i - 1]) {
            rce_symbol(c, st, i - last - 1, 0);
            last = i;
        }
    rce_symbol(c, st, i - last - 1, 0);
}

int ffv1o_write_extradata(const ffv1o_params *p, uint8_t *dst, int cap)
{
    RcEnc c;
    uint8_t st[32];
    uint8_t *tmp;
    int i, t, n;
    uint32_t crc;
    if (p->version < 2) return 0;
    tmp = malloc(65536);
    if (!tmp) return E_NOMEM;
    memset(&c, 0, sizeof(c));
    rce_init(&c, tmp, 65536 - 4);
    ffv1o_default_state_tables(c.zero_state, c.one_state);
    memset(st, 128, sizeof(st));

    rce_symbol(&c, st, p->version, 0);
    if (p->version > 2)
        rce_symbol(&c, st, p->micro_version, 0);
    rce_symbol(&c, st, p->ac, 0);
    if (p->ac == FFV1O_AC_RANGE_CUSTOM)
        for (i = 1; i < 256; i++)
            rce_symbol(&c, st, p->state_transition[i] - c.one_state[i], 1);
    rce_symbol(&c, st, p->colorspace, 0);
    rce_symbol(&c, st, p->bits, 0);
    rce_put(&c, st, p->chroma_planes);
    rce_symbol(&c, st, p->chroma_h_shift, 0);
    rce_symbol(&c, st, p->chroma_v_shift, 0);
    rce_put(&c, st, p->transparency);
    rce_symbol(&c, st, p->num_h_slices - 1, 0);
    rce_symbol(&c, st, p->num_v_slices - 1, 0);
    rce_symbol(&c, st, 2, 0);                            /* quant_table_count */
    for (t = 0; t < 2; t++)
        for (i = 0; i < 5; i++)
            put_quant_table(&c, p->quant_tables[t][i]);
    for (t = 0; t < 2; t++)
        rce_put(&c, st, 0);                              /* initial states all 128 (no 2-pass) */
    if (p->version > 2) {
        rce_symbol(&c, st, p->ec, 0);
        rce_symbol(&c, st, p->intra, 0);
    }
    n = (int)rce_terminate(&c);
    crc = ffv1o_crc32(0, tmp, n);
    tmp[n] = crc >> 24; tmp[n + 1] = crc >> 16; tmp[n + 2] = crc >> 8; tmp[n + 3] = crc;
    n += 4;
    if (n > cap || c.overflow) { free(tmp); return E_INVAL; }
    memcpy(dst, tmp, n);
    free(tmp);
    return n;
}

static int get_quant_table(RcDec *c, int16_t *q, int scale)
{
    /* ffv1dec.c:476-500 */
    uint8_t st[32];
    int v, i = 0, err = 0;
    memset(st, 128, sizeof(st));
    for (v = 0; i < 128; v++) {
        unsigned len = (unsigned)rcd_symbol(c, st, 0, &err) + 1;
        if (err || len > (unsigned)(128 - i) || !len) return E_INVALIDDATA;
        while (len--) q[i++] = (int16_t)(scale * v);
    }
    for (i = 1; i < 128; i++) q[256 - i] = (int16_t)-q[i];
    q[128] = (int16_t)-q[127];
    return 2 * v - 1;
}
static int get_quant_tables(RcDec *c, int16_t q[5][256])
{
    int i, count = 1;
    for (i = 0; i < 5; i++) {
        int r = get_quant_table(c, q[i], count);
        if (r < 0) return r;
        count *= r;
        if ((unsigned)count > 32768U) return E_INVALIDDATA;
    }
    return (count + 1) / 2;
}

static int derive_layout(ffv1o_params *p)
{
    /* the pix_fmt selection of ffv1dec.c:698-786 reduced to what the oracle needs */
    if (p->colorspace == 0) {
        if (p->transparency && !p->chroma_planes) {
            if (p->bits > 8) return E_NOSYS;
            p->layout = FFV1O_LAYOUT_YA8;
        } else
            p->layout = FFV1O_LAYOUT_PLANAR;
        p->packed_at_lsb = (p->bits == 9 || p->bits == 10);
        if (p->bits > 8 && p->bits != 9 && p->bits != 10 && p->bits != 16) return E_NOSYS;
    } else if (p->colorspace == 1) {
        if (p->chroma_h_shift || p->chroma_v_shift) return E_NOSYS;
        if (p->bits <= 8) p->layout = FFV1O_LAYOUT_RGB32;
        else if ((p->bits == 9 || p->bits == 10 || p->bits == 12 || p->bits == 14) && !p->transparency)
            p->layout = FFV1O_LAYOUT_GBRP;
        else return E_NOSYS;
    } else return E_NOSYS;
    return 0;
}

int ffv1o_parse_extradata(ffv1o_params *p, int width, int height, const uint8_t *d, int n)
{
    RcDec c;
    uint8_t st[32];
    int i, err = 0, qtc;
    memset(p, 0, sizeof(*p));
    p->width = width; p->height = height;
    memset(st, 128, sizeof(st));
    rcd_init(&c, d, n);
    ffv1o_default_state_tables(c.zero_state, c.one_state);

    p->version = rcd_symbol(&c, st, 0, &err);
    if (p->version < 2) return E_INVALIDDATA;
    if (p->version > 2) {
        c.end -= 4;
        p->micro_version = rcd_symbol(&c, st, 0, &err);
        if (p->micro_version < 0) return E_INVALIDDATA;
    }
    p->ac = rcd_symbol(&c, st, 0, &err);
    if (p->ac == FFV1O_AC_RANGE_CUSTOM)
        for (i = 1; i < 256; i++)
            p->state_transition[i] = (uint8_t)(rcd_symbol(&c, st, 1, &err) + c.one_state[i]);
    else
        for (i = 1; i < 256; i++) p->state_transition[i] = c.one_state[i];
    p->colorspace = rcd_symbol(&c, st, 0, &err);
    p->bits = rcd_symbol(&c, st, 0, &err);
    p->chroma_planes = rcd_get(&c, st);
    p->chroma_h_shift = rcd_symbol(&c, st, 0, &err);
    p->chroma_v_shift = rcd_symbol(&c, st, 0, &err);
    p->transparency = rcd_get(&c, st);
    p->plane_count = 1 + (p->chroma_planes || p->version < 4) + p->transparency;
    p->num_h_slices = 1 + rcd_symbol(&c, st, 0, &err);
    p->num_v_slices = 1 + rcd_symbol(&c, st, 0, &err);
    if ((unsigned)p->chroma_h_shift > 4U || (unsigned)p->chroma_v_shift > 4U) return E_INVALIDDATA;
    if (p->num_h_slices > width || p->num_h_slices <= 0 || p->num_v_slices > height || p->num_v_slices <= 0)
        return E_INVALIDDATA;
    qtc = rcd_symbol(&c, st, 0, &err);
    if (qtc != 2) return E_NOSYS;        /* the reference encoder always writes 2 (ffv1enc.c:847) */
    for (i = 0; i < qtc; i++) {
        p->context_count[i] = get_quant_tables(&c, p->quant_tables[i]);
        if (p->context_count[i] < 0) return E_INVALIDDATA;
    }
    for (i = 0; i < qtc; i++)
        if (rcd_get(&c, st)) return E_NOSYS;   /* 2-pass initial states: out of scope */
    if (p->version > 2) {
        p->ec = rcd_symbol(&c, st, 0, &err);
        if (p->micro_version > 2) p->intra = rcd_symbol(&c, st, 0, &err);
    }
    if (err) return E_INVALIDDATA;
    if (p->version > 2) {
        if (n < 4 || ffv1o_crc32(0, d, n)) return E_INVALIDDATA;           /* ffv1dec.c:610-617 */
    }
    if (!p->bits) p->bits = 8;
    p->context_model = 0;   /* per-slice quant_table_index decides */
    return derive_layout(p);
}

/* ------------------------------------------------------------------------------------------
 * per-slice model state
 * ---------------------------------------------------------------------------------------- */
typedef struct { int16_t drift; uint16_t error_sum; int8_t bias; uint8_t count; } Vlc;   /* ffv1.h:61-66 */

typedef struct {
    uint8_t *state[3];      /* [ctx*32 + slot]; plane contexts 0 = Y/G, 1 = chroma (shared) or YA8 alpha, 2 = alpha */
    Vlc     *vlc[3];
    int      alloc_ctx[3];
    int      x0, y0, w, h;  /* luma geometry */
    int      damaged;
    int      coding_mode, rct_by, rct_ry;   /* version 4 slice header fields */
} SliceModel;

static void slice_geometry(const ffv1o_params *p, int i, int *x0, int *y0, int *w, int *h)
{
    /* ffv1.c:124-143 */
    int sx = i % p->num_h_slices, sy = i / p->num_h_slices;
    int xs = p->width * sx / p->num_h_slices, xe = p->width * (sx + 1) / p->num_h_slices;
    int ys = p->height * sy / p->num_v_slices, ye = p->height * (sy + 1) / p->num_v_slices;
    *x0 = xs; *y0 = ys; *w = xe - xs; *h = ye - ys;
}

static int model_alloc(SliceModel *m, int pc, int nctx, int golomb)
{
    if (m->alloc_ctx[pc] >= nctx && (golomb ? (m->vlc[pc] != NULL) : (m->state[pc] != NULL))) return 0;
    free(m->state[pc]); free(m->vlc[pc]);
    m->state[pc] = NULL; m->vlc[pc] = NULL;
    if (golomb) m->vlc[pc] = malloc(sizeof(Vlc) * nctx);
    else        m->state[pc] = malloc(32 * (size_t)nctx);
    m->alloc_ctx[pc] = nctx;
    return (golomb ? (m->vlc[pc] != NULL) : (m->state[pc] != NULL)) ? 0 : E_NOMEM;
}
static void model_reset(SliceModel *m, int pc, int nctx, int golomb)
{
    /* ffv1.c:177-202 */
    int j;
    if (golomb)
        for (j = 0; j < nctx; j++) { m->vlc[pc][j].drift = 0; m->vlc[pc][j].error_sum = 4; m->vlc[pc][j].bias = 0; m->vlc[pc][j].count = 1; }
    else
        memset(m->state[pc], 128, 32 * (size_t)nctx);
}
static void model_free(SliceModel *m)
{
    int i;
    for (i = 0; i < 3; i++) { free(m->state[i]); free(m->vlc[i]); }
}

/* ------------------------------------------------------------------------------------------
 * sample gathering and neighbourhood (ffv1enc.c:373-473, ffv1.h:148-190)
 * ---------------------------------------------------------------------------------------- */
typedef struct { int16_t *s; int w, h; } Plane16;

static int sample_at(const Plane16 *P, int x, int y) { return P->s[(size_t)y * P->w + x]; }
static int nb_T (const Plane16 *P, int x, int y) { return y > 0 ? sample_at(P, x, y - 1) : 0; }
static int nb_L (const Plane16 *P, int x, int y) { return x > 0 ? sample_at(P, x - 1, y) : nb_T(P, 0, y); }
static int nb_LT(const Plane16 *P, int x, int y) { return x > 0 ? nb_T(P, x - 1, y) : (y >= 2 ? sample_at(P, 0, y - 2) : 0); }
static int nb_RT(const Plane16 *P, int x, int y) { return x < P->w - 1 ? nb_T(P, x + 1, y) : nb_T(P, P->w - 1, y); }
static int nb_TT(const Plane16 *P, int x, int y) { return y >= 2 ? sample_at(P, x, y - 2) : 0; }
static int nb_LL(const Plane16 *P, int x, int y) { return x >= 2 ? sample_at(P, x - 2, y) : (x == 1 ? nb_T(P, 0, y) : 0); }

static int median3(int a, int b, int c)
{
    /* mathops.h:95-119 */
    if (a > b) { int t = a; a = b; b = t; }
    if (b > c) b = c;
    if (a > b) b = a;
    return b;
}
static int predict_at(const Plane16 *P, int x, int y)
{
    int L = nb_L(P, x, y), T = nb_T(P, x, y), LT = nb_LT(P, x, y);
    return median3(L, L + T - LT, T);
}
static int context_at(const int16_t q[5][256], const Plane16 *P, int x, int y)
{
    int L = nb_L(P, x, y), T = nb_T(P, x, y), LT = nb_LT(P, x, y), RT = nb_RT(P, x, y);
    int c = q[0][(L - LT) & 0xFF] + q[1][(LT - T) & 0xFF] + q[2][(T - RT) & 0xFF];
    if (q[3][127])
        c += q[3][(nb_LL(P, x, y) - L) & 0xFF] + q[4][(nb_TT(P, x, y) - T) & 0xFF];
    return c;
}
static int fold_diff(int d, int bits)
{
    /* ffv1.h:148-159 */
    if (bits == 8) return (int8_t)d;
    d += 1 << (bits - 1);
    d &= (1 << bits) - 1;
    d -= 1 << (bits - 1);
    return d;
}

/* number of sample planes coded per slice and the bit width they are coded with */
static int coded_bits(const ffv1o_params *p)
{
    if (p->colorspace == 1) return (p->bits <= 8 ? 8 : p->bits) + 1;     /* ffv1enc.c:465-468 */
    return p->bits <= 8 ? 8 : p->bits;
}

/* Gather the int16 sample arrays of slice (x0,y0,w,h).  Returns the number of arrays in out[]:
 *   YUV planar: Y, [U, V], [A]      YA8: Y, A      RGB: G', B', R', [A] after the RCT.
 * pc[] receives the plane-context index each array is coded with, cw/ch its size. */
static int gather_slice_ex(const ffv1o_params *p, const uint8_t *const planes[4], const int strides[4],
                           int x0, int y0, int w, int h, Plane16 out[4], int pc[4], int by, int ry, int pcm);
static int gather_slice(const ffv1o_params *p, const uint8_t *const planes[4], const int strides[4],
                        int x0, int y0, int w, int h, Plane16 out[4], int pc[4])
{
    return gather_slice_ex(p, planes, strides, x0, y0, w, h, out, pc, 1, 1, 0);
}

/* by, ry: RCT coefficients of the slice (1, 1 before version 4); pcm: slice_coding_mode 1 keeps b, g, r as they are
 * (ffv1enc.c:447-453) */
static int gather_slice_ex(const ffv1o_params *p, const uint8_t *const planes[4], const int strides[4],
                           int x0, int y0, int w, int h, Plane16 out[4], int pc[4], int by, int ry, int pcm)
{
    int n = 0, x, y, k;
    if (p->colorspace == 0) {
        int nplanes_src[4], hs[4], vs[4], pstep[4], poff[4], srcidx[4];
        if (p->layout == FFV1O_LAYOUT_YA8) {
            n = 2;
            srcidx[0] = 0; poff[0] = 0; pstep[0] = 2; hs[0] = vs[0] = 0; pc[0] = 0;
            srcidx[1] = 0; poff[1] = 1; pstep[1] = 2; hs[1] = vs[1] = 0; pc[1] = 1;   /* ffv1enc.c:1199-1201 */
        } else {
            srcidx[n] = 0; poff[n] = 0; pstep[n] = 1; hs[n] = vs[n] = 0; pc[n] = 0; n++;
            if (p->chroma_planes) {
                for (k = 1; k <= 2; k++) { srcidx[n] = k; poff[n] = 0; pstep[n] = 1; hs[n] = p->chroma_h_shift; vs[n] = p->chroma_v_shift; pc[n] = 1; n++; }
            }
            if (p->transparency) { srcidx[n] = 3; poff[n] = 0; pstep[n] = 1; hs[n] = vs[n] = 0; pc[n] = 2; n++; }
        }
        (void)nplanes_src;
        for (k = 0; k < n; k++) {
            /* ffv1enc.c:1186-1189: chroma origin x>>shift, size ceil(w / 2^shift) */
            int cw = hs[k] ? -((-w) >> hs[k]) : w, chh = vs[k] ? -((-h) >> vs[k]) : h;
            int cx = x0 >> hs[k], cy = y0 >> vs[k];
            const uint8_t *src = planes[srcidx[k]];
            int st = strides[srcidx[k]];
            out[k].w = cw; out[k].h = chh;
            out[k].s = malloc(sizeof(int16_t) * (size_t)cw * chh + 2);
            if (!out[k].s) return E_NOMEM;
            for (y = 0; y < chh; y++)
                for (x = 0; x < cw; x++) {
                    int v;
                    if (p->bits <= 8)
                        v = src[(size_t)(cy + y) * st + (size_t)(cx + x) * pstep[k] + poff[k]];
                    else {
                        const uint8_t *q = src + (size_t)(cy + y) * st + 2 * (size_t)(cx + x);
                        v = q[0] | (q[1] << 8);
                        if (!p->packed_at_lsb) v >>= 16 - p->bits;                  /* ffv1enc.c:396-403 */
                        v = (int16_t)v;                                             /* int16 ring buffer: SURVEY A.7 */
                    }
                    out[k].s[(size_t)y * cw + x] = (int16_t)v;
                }
        }
    } else {
        /* ffv1enc.c:413-458 */
        int offset = 1 << (p->bits <= 8 ? 8 : p->bits);
        n = 3 + (p->transparency ? 1 : 0);
        for (k = 0; k < n; k++) {
            out[k].w = w; out[k].h = h; pc[k] = (k + 1) / 2;
            out[k].s = malloc(sizeof(int16_t) * (size_t)w * h + 2);
            if (!out[k].s) return E_NOMEM;
        }
        for (y = 0; y < h; y++)
            for (x = 0; x < w; x++) {
                int b, g, r, a = 0;
                if (p->layout == FFV1O_LAYOUT_RGB32) {
                    const uint8_t *q = planes[0] + (size_t)(y0 + y) * strides[0] + 4 * (size_t)(x0 + x);
                    b = q[0]; g = q[1]; r = q[2]; a = q[3];
                } else {
                    /* the reference reads "b,g,r" from data[0],data[1],data[2], i.e. from the G,B,R planes
                     * of GBRP -- keep that naming */
                    const uint8_t *q0 = planes[0] + (size_t)(y0 + y) * strides[0] + 2 * (size_t)(x0 + x);
                    const uint8_t *q1 = planes[1] + (size_t)(y0 + y) * strides[1] + 2 * (size_t)(x0 + x);
                    const uint8_t *q2 = planes[2] + (size_t)(y0 + y) * strides[2] + 2 * (size_t)(x0 + x);
                    b = q0[0] | (q0[1] << 8); g = q1[0] | (q1[1] << 8); r = q2[0] | (q2[1] << 8);
                }
                if (!pcm) {
                    b -= g; r -= g;
                    g += (b * by + r * ry) >> 2;
                    b += offset; r += offset;
                }
                out[0].s[(size_t)y * w + x] = (int16_t)g;
                out[1].s[(size_t)y * w + x] = (int16_t)b;
                out[2].s[(size_t)y * w + x] = (int16_t)r;
                if (n == 4) out[3].s[(size_t)y * w + x] = (int16_t)a;
            }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * line coder, encoder (ffv1enc.c:271-371, 240-269; golomb.h:508-563; ffv1.h:192-224)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    RcEnc rc;
    BitW  bw;
    int   golomb;
    int   run_index;
    int   pcm;           /* version 4: slice_coding_mode 1 */
    uint64_t symbols;
} SliceCoder;

static void vlc_update(Vlc *s, int v)
{
    int drift = s->drift, count = s->count;
    s->error_sum += v < 0 ? -v : v;
    drift += v;
    if (count == 128) { count >>= 1; drift >>= 1; s->error_sum >>= 1; }
    count++;
    if (drift <= -count) {
        if (s->bias > -128) s->bias--;
        drift += count;
        if (drift <= -count) drift = -count + 1;
    } else if (drift > 0) {
        if (s->bias < 127) s->bias++;
        drift -= count;
        if (drift > 0) drift = 0;
    }
    s->drift = (int16_t)drift;
    s->count = (uint8_t)count;
}
static int vlc_k(const Vlc *s)
{
    int i = s->count, k = 0;
    while (i < s->error_sum) { k++; i += i; }
    return k;
}
static void put_vlc(BitW *w, Vlc *s, int v, int bits)
{
    int k, code, m, e;
    v = fold_diff(v - s->bias, bits);
    k = vlc_k(s);
    code = v ^ ((2 * s->drift + s->count) >> 31);
    /* set_sr_golomb(code, k, 12, bits) */
    m = -2 * code - 1;
    m ^= (m >> 31);
    e = m >> k;
    if (e < 12) bw_put(w, e + k + 1, (1u << k) + ((unsigned)m & ((1u << k) - 1)));
    else        bw_put(w, 12 + bits, (unsigned)(m - 12 + 1));
    vlc_update(s, v);
}

static void encode_line(const ffv1o_params *p, SliceCoder *sc, SliceModel *m, int pcidx, const int16_t q[5][256],
                        const Plane16 *P, int y, int bits)
{
    int x, run_count = 0, run_mode = 0;
    if (sc->pcm) {                                     /* ffv1enc.c:294-304: `bits` raw bits, each on a fresh state 128 */
        for (x = 0; x < P->w; x++) {
            int i, v = sample_at(P, x, y);
            for (i = bits - 1; i >= 0; i--) { uint8_t st = 128; rce_put(&sc->rc, &st, (v >> i) & 1); }
        }
        return;
    }
    for (x = 0; x < P->w; x++) {
        int ctx = context_at(q, P, x, y);
        int diff = sample_at(P, x, y) - predict_at(P, x, y);
        if (ctx < 0) { ctx = -ctx; diff = -diff; }
        diff = fold_diff(diff, bits);
        sc->symbols++;
        if (!sc->golomb) {
            rce_symbol(&sc->rc, m->state[pcidx] + 32 * (size_t)ctx, diff, 1);
        } else {
            if (ctx == 0) run_mode = 1;
            if (run_mode) {
                if (diff) {
                    while (run_count >= 1 << log2_run[sc->run_index]) {
                        run_count -= 1 << log2_run[sc->run_index];
                        sc->run_index++;
                        bw_put(&sc->bw, 1, 1);
                    }
                    bw_put(&sc->bw, 1 + log2_run[sc->run_index], run_count);
                    if (sc->run_index) sc->run_index--;
                    run_count = 0; run_mode = 0;
                    if (diff > 0) diff--;
                } else
                    run_count++;
            }
            if (!run_mode)
                put_vlc(&sc->bw, &m->vlc[pcidx][ctx], diff, bits);
        }
    }
    if (run_mode) {
        while (run_count >= 1 << log2_run[sc->run_index]) {
            run_count -= 1 << log2_run[sc->run_index];
            sc->run_index++;
            bw_put(&sc->bw, 1, 1);
        }
        if (run_count) bw_put(&sc->bw, 1, 1);
    }
    (void)p;
}

/* ------------------------------------------------------------------------------------------
 * encoder
 * ---------------------------------------------------------------------------------------- */
struct ffv1o_encoder {
    ffv1o_params p;
    int picture_number;
    int slice_count;
    SliceModel *sm;
    uint64_t decisions;
};

ffv1o_encoder *ffv1o_encoder_new(const ffv1o_params *p)
{
    ffv1o_encoder *e = calloc(1, sizeof(*e));
    int i;
    if (!e) return NULL;
    e->p = *p;
    e->slice_count = p->num_h_slices * p->num_v_slices;
    e->sm = calloc(e->slice_count, sizeof(SliceModel));
    for (i = 0; i < e->slice_count; i++)
        slice_geometry(p, i, &e->sm[i].x0, &e->sm[i].y0, &e->sm[i].w, &e->sm[i].h);
    return e;
}
void ffv1o_encoder_set_force_pcm(ffv1o_encoder *e, int on) { e->p.force_pcm = on; }

void ffv1o_encoder_free(ffv1o_encoder *e)
{
    int i;
    if (!e) return;
    for (i = 0; i < e->slice_count; i++) model_free(&e->sm[i]);
    free(e->sm); free(e);
}
uint64_t ffv1o_encoder_decisions(const ffv1o_encoder *e) { return e->decisions; }

static void load_transition(const ffv1o_params *p, uint8_t zero_state[256], uint8_t one_state[256])
{
    /* default table first (ffv1enc.c:1288), custom overrides entries 1..255 (1309-1315, ffv1.c:95-100) */
    int j;
    ffv1o_default_state_tables(zero_state, one_state);
    if (p->ac == FFV1O_AC_RANGE_CUSTOM)
        for (j = 1; j < 256; j++) {
            one_state[j] = p->state_transition[j];
            zero_state[256 - j] = (uint8_t)(256 - one_state[j]);
        }
}

static void write_v01_header(const ffv1o_params *p, RcEnc *c)
{
    /* ffv1enc.c:498-524 (version < 2 branch) */
    uint8_t st[32];
    int i;
    memset(st, 128, sizeof(st));
    rce_symbol(c, st, p->version, 0);
    rce_symbol(c, st, p->ac, 0);
    if (p->ac == FFV1O_AC_RANGE_CUSTOM)
        for (i = 1; i < 256; i++)
            rce_symbol(c, st, p->state_transition[i] - c->one_state[i], 1);
    rce_symbol(c, st, p->colorspace, 0);
    if (p->version > 0) rce_symbol(c, st, p->bits, 0);
    rce_put(c, st, p->chroma_planes);
    rce_symbol(c, st, p->chroma_h_shift, 0);
    rce_symbol(c, st, p->chroma_v_shift, 0);
    rce_put(c, st, p->transparency);
    for (i = 0; i < 5; i++) put_quant_table(c, p->quant_tables[p->context_model][i]);
}

/* choose_rct_params (ffv1enc.c:1064-1144): the candidate whose luma-like channel has the smallest sum of absolute
 * second-order differences over the slice; {ry, by} per candidate, first minimum wins, int accumulators */
static const int rct_candidates[15][2] = {
    {0, 0}, {1, 1}, {2, 2}, {0, 2}, {2, 0}, {4, 0}, {0, 4}, {0, 3}, {3, 0}, {3, 1}, {1, 3}, {1, 2}, {2, 1}, {0, 1}, {1, 0},
};

static void rct_search(const ffv1o_params *p, const uint8_t *const planes[4], const int strides[4],
                       int x0, int y0, int w, int h, int *by, int *ry)
{
    unsigned stat[15] = { 0 };
    int16_t *prev = calloc((size_t)w * 3 + 3, sizeof(int16_t));
    int x, y, i, best = 0;
    for (y = 0; y < h; y++) {
        int last[3] = { 0, 0, 0 };
        for (x = 0; x < w; x++) {
            int v[3], d[3];                            /* g, b, r under the reference's plane naming */
            if (p->layout == FFV1O_LAYOUT_RGB32) {
                const uint8_t *q = planes[0] + (size_t)(y0 + y) * strides[0] + 4 * (size_t)(x0 + x);
                v[1] = q[0]; v[0] = q[1]; v[2] = q[2];
            } else {
                const uint8_t *q0 = planes[0] + (size_t)(y0 + y) * strides[0] + 2 * (size_t)(x0 + x);
                const uint8_t *q1 = planes[1] + (size_t)(y0 + y) * strides[1] + 2 * (size_t)(x0 + x);
                const uint8_t *q2 = planes[2] + (size_t)(y0 + y) * strides[2] + 2 * (size_t)(x0 + x);
                v[1] = q0[0] | (q0[1] << 8); v[0] = q1[0] | (q1[1] << 8); v[2] = q2[0] | (q2[1] << 8);
            }
            for (i = 0; i < 3; i++) d[i] = v[i] - last[i];
            if (x && y) {
                int bg = d[0] - prev[x * 3], bb = d[1] - prev[x * 3 + 1], br = d[2] - prev[x * 3 + 2];
                br -= bg; bb -= bg;
                for (i = 0; i < 15; i++) {
                    int t = bg + ((br * rct_candidates[i][0] + bb * rct_candidates[i][1]) >> 2);
                    stat[i] += (unsigned)(t < 0 ? -t : t);
                }
            }
            for (i = 0; i < 3; i++) { prev[x * 3 + i] = (int16_t)d[i]; last[i] = v[i]; }
        }
    }
    for (i = 1; i < 15; i++)
        if ((int)stat[i] < (int)stat[best]) best = i;
    free(prev);
    *by = rct_candidates[best][1];
    *ry = rct_candidates[best][0];
}

static void write_slice_header(const ffv1o_params *p, RcEnc *c, const SliceModel *m,
                               int sar_num, int sar_den, int picture_structure)
{
    /* ffv1enc.c:1031-1051 */
    uint8_t st[32];
    int j;
    memset(st, 128, sizeof(st));
    rce_symbol(c, st, (m->x0 + 1) * p->num_h_slices / p->width, 0);
    rce_symbol(c, st, (m->y0 + 1) * p->num_v_slices / p->height, 0);
    rce_symbol(c, st, (m->w + 1) * p->num_h_slices / p->width - 1, 0);
    rce_symbol(c, st, (m->h + 1) * p->num_v_slices / p->height - 1, 0);
    for (j = 0; j < p->plane_count; j++) rce_symbol(c, st, p->context_model, 0);
    rce_symbol(c, st, picture_structure, 0);
    rce_symbol(c, st, sar_num, 0);
    rce_symbol(c, st, sar_den, 0);
    if (p->version > 3) {                              /* ffv1enc.c:1052-1061 */
        rce_put(c, st, m->coding_mode == 1);           /* "reset contexts", on state[0] */
        rce_symbol(c, st, m->coding_mode, 0);
        if (m->coding_mode != 1) {
            rce_symbol(c, st, m->rct_by, 0);
            rce_symbol(c, st, m->rct_ry, 0);
        }
    }
}

long ffv1o_encode_frame(ffv1o_encoder *e, const uint8_t *const planes[4], const int strides[4],
                        int sar_num, int sar_den, int picture_structure,
                        uint8_t *dst, long cap, int *key_frame)
{
    const ffv1o_params *p = &e->p;
    const int key = (p->gop_size == 0 || e->picture_number % p->gop_size == 0);   /* ffv1enc.c:1299 */
    const int golomb = p->ac == FFV1O_AC_GOLOMB;
    const int bits = coded_bits(p);
    const int nctx = p->context_count[p->context_model];
    long out = 0;
    int si;

    for (si = 0; si < e->slice_count; si++) {
        SliceModel *m = &e->sm[si];
        SliceCoder sc;
        Plane16 pl[4];
        int pc[4], npl, k, y, r;
        long scap = 65536 + (long)m->w * m->h * 4 * 5, bytes, ac_bytes = 0;
        uint8_t *sbuf = malloc(scap);
        if (!sbuf) return E_NOMEM;
        memset(&sc, 0, sizeof(sc));
        sc.golomb = golomb;
        rce_init(&sc.rc, sbuf, scap);
        ffv1o_default_state_tables(sc.rc.zero_state, sc.rc.one_state);
        if (si == 0) {
            uint8_t keystate = 128;                                                /* 1299-1307 */
            rce_put(&sc.rc, &keystate, key);
            if (key && p->version < 2) write_v01_header(p, &sc.rc);
        }
        load_transition(p, sc.rc.zero_state, sc.rc.one_state);

        for (k = 0; k < 3; k++) {
            int used = (k == 0) || (k == 1 && (p->chroma_planes || p->layout == FFV1O_LAYOUT_YA8)) ||
                       (k == 2 && p->transparency && p->layout != FFV1O_LAYOUT_YA8);
            if (!used) continue;
            if ((r = model_alloc(m, k, nctx, golomb)) < 0) { free(sbuf); return r; }
            if (key) model_reset(m, k, nctx, golomb);                              /* 1171-1172 */
        }
        /* version 4 (1162-1168): RGB slices search their RCT coefficients.  (The reference runs the same search over
         * planar YUV / gray frames, reading their first plane as packed RGB and beyond its rows -- values no decoder
         * uses and no other encoder can reproduce; the neutral pair 1, 1 is coded for those.) */
        m->coding_mode = (p->version > 3 && p->force_pcm && !golomb) ? 1 : 0;
        m->rct_by = m->rct_ry = 1;
        if (p->version > 3 && p->colorspace == 1) rct_search(p, planes, strides, m->x0, m->y0, m->w, m->h, &m->rct_by, &m->rct_ry);
        sc.pcm = m->coding_mode == 1;
        if (sc.pcm)                                                                /* 1054-1055: a PCM slice clears its state */
            for (k = 0; k < 3; k++)
                if (m->alloc_ctx[k]) model_reset(m, k, nctx, golomb);
        if (p->version > 2) write_slice_header(p, &sc.rc, m, sar_num, sar_den, picture_structure);
        if (golomb) {                                                              /* 1176-1183 */
            if (p->version > 2) { uint8_t s129 = 129; rce_put(&sc.rc, &s129, 0); }
            ac_bytes = (p->version > 2 || (m->x0 == 0 && m->y0 == 0)) ? rce_terminate(&sc.rc) : 0;
            sc.bw.buf = sbuf + ac_bytes; sc.bw.cap = scap - ac_bytes; sc.bw.bitpos = 0;
        }

        npl = gather_slice_ex(p, planes, strides, m->x0, m->y0, m->w, m->h, pl, pc, m->rct_by, m->rct_ry, sc.pcm);
        if (npl < 0) { free(sbuf); return npl; }
        if (p->colorspace == 0) {
            for (k = 0; k < npl; k++) {                                            /* 1185-1201 */
                sc.run_index = 0;                                                  /* 379 */
                for (y = 0; y < pl[k].h; y++)
                    encode_line(p, &sc, m, pc[k], p->quant_tables[p->context_model], &pl[k], y, bits);
            }
        } else {
            sc.run_index = 0;                                                      /* 423 */
            for (y = 0; y < m->h; y++)
                for (k = 0; k < npl; k++)                                          /* 459-469: PCM slices code the raw depth */
                    encode_line(p, &sc, m, pc[k], p->quant_tables[p->context_model], &pl[k], y, bits - sc.pcm);
        }
        for (k = 0; k < npl; k++) free(pl[k].s);

        if (!golomb) {                                                             /* 1331-1334 */
            uint8_t s129 = 129;
            rce_put(&sc.rc, &s129, 0);
            bytes = rce_terminate(&sc.rc);
        } else {                                                                   /* 1336-1337 */
            bytes = ac_bytes + (long)((sc.bw.bitpos + 7) / 8);
        }
        e->decisions += golomb ? sc.symbols : sc.rc.decisions;
        if (sc.rc.overflow || sc.bw.overflow || out + bytes + 8 > cap) { free(sbuf); return E_INVAL; }
        memcpy(dst + out, sbuf, bytes);
        free(sbuf);
        if (si > 0 || p->version > 2) {                                            /* 1339-1345 */
            dst[out + bytes] = (uint8_t)(bytes >> 16); dst[out + bytes + 1] = (uint8_t)(bytes >> 8); dst[out + bytes + 2] = (uint8_t)bytes;
            bytes += 3;
        }
        if (p->ec) {                                                               /* 1346-1352 */
            uint32_t crc;
            dst[out + bytes++] = 0;
            crc = ffv1o_crc32(0, dst + out, bytes);
            dst[out + bytes] = crc >> 24; dst[out + bytes + 1] = crc >> 16; dst[out + bytes + 2] = crc >> 8; dst[out + bytes + 3] = crc;
            bytes += 4;
        }
        out += bytes;
    }
    e->picture_number++;
    if (key_frame) *key_frame = key;
    return out;
}

long ffv1o_slice_records(const ffv1o_params *p, const uint8_t *const planes[4], const int strides[4],
                         int slice_index, uint32_t *rec, long cap)
{
    SliceModel m;
    Plane16 pl[4];
    int pc[4], npl, k, x, y;
    const int bits = coded_bits(p);
    long n = 0;
    memset(&m, 0, sizeof(m));
    slice_geometry(p, slice_index, &m.x0, &m.y0, &m.w, &m.h);
    npl = gather_slice(p, planes, strides, m.x0, m.y0, m.w, m.h, pl, pc);
    if (npl < 0) return npl;
#define EMIT_LINE(P, yy) \
    for (x = 0; x < (P)->w; x++) { \
        int ctx = context_at(p->quant_tables[p->context_model], (P), x, (yy)); \
        int diff = sample_at((P), x, (yy)) - predict_at((P), x, (yy)); \
        if (ctx < 0) { ctx = -ctx; diff = -diff; } \
        diff = fold_diff(diff, bits); \
        if (n < cap) rec[n] = ((uint32_t)ctx << 16) | ((uint32_t)diff & 0xFFFFu); \
        n++; \
    }
    if (p->colorspace == 0) {
        for (k = 0; k < npl; k++)
            for (y = 0; y < pl[k].h; y++) EMIT_LINE(&pl[k], y)
    } else {
        for (y = 0; y < m.h; y++)
            for (k = 0; k < npl; k++) EMIT_LINE(&pl[k], y)
    }
#undef EMIT_LINE
    for (k = 0; k < npl; k++) free(pl[k].s);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * decoder (ffv1dec.c:70-474, 638-1035)
 * ---------------------------------------------------------------------------------------- */
struct ffv1o_decoder {
    ffv1o_params p;
    int have_params;
    int key_frame_ok;
    int slice_count;
    int max_slices;
    SliceModel *sm;
};

ffv1o_decoder *ffv1o_decoder_new(int width, int height, const uint8_t *extradata, int extradata_size)
{
    ffv1o_decoder *d = calloc(1, sizeof(*d));
    if (!d) return NULL;
    d->p.width = width; d->p.height = height;
    d->p.num_h_slices = d->p.num_v_slices = 1;
    if (extradata_size > 0) {
        if (ffv1o_parse_extradata(&d->p, width, height, extradata, extradata_size) < 0) { free(d); return NULL; }
        d->have_params = 1;
    }
    d->max_slices = d->p.num_h_slices * d->p.num_v_slices;
    d->sm = calloc(d->max_slices, sizeof(SliceModel));
    return d;
}
const ffv1o_params *ffv1o_decoder_params(const ffv1o_decoder *d) { return &d->p; }
void ffv1o_decoder_free(ffv1o_decoder *d)
{
    int i;
    if (!d) return;
    for (i = 0; i < d->max_slices; i++) model_free(&d->sm[i]);
    free(d->sm); free(d);
}

static int get_vlc(BitR *r, Vlc *s, int bits)
{
    /* ffv1dec.c:70-98 with get_sr_golomb/get_ur_golomb (golomb.h:270-300, 367-372) */
    int k = vlc_k(s), v, ret;
    unsigned buf = br_peek32(r), u;
    int lz = 0;
    while (lz < 32 && !(buf & (0x80000000u >> lz))) lz++;
    if (lz < 12) {
        r->bitpos += lz + 1;
        u = ((unsigned)lz << k) + br_get(r, k);
    } else {
        r->bitpos += 12;
        u = br_get(r, bits) + 11;
    }
    v = (int)(u >> 1) ^ -(int)(u & 1);
    v ^= ((2 * s->drift + s->count) >> 31);
    ret = fold_diff(v + s->bias, bits);
    vlc_update(s, v);
    return ret;
}

typedef struct {
    RcDec rc;
    BitR  br;
    int   golomb;
    int   run_index;
    int   pcm;           /* version 4: slice_coding_mode 1 */
} SliceReader;

static void decode_line(SliceReader *sr, SliceModel *m, int pcidx, const int16_t q[5][256],
                        Plane16 *P, int y, int bits, int *err)
{
    /* ffv1dec.c:100-181 */
    int x, run_count = 0, run_mode = 0;
    if (sr->pcm) {                                     /* 111-122 */
        for (x = 0; x < P->w; x++) {
            int i, v = 0;
            for (i = 0; i < bits; i++) { uint8_t st = 128; v += v + rcd_get(&sr->rc, &st); }
            P->s[(size_t)y * P->w + x] = (int16_t)v;
        }
        return;
    }
    for (x = 0; x < P->w; x++) {
        int ctx = context_at(q, P, x, y), sign = 0, diff;
        if (ctx < 0) { ctx = -ctx; sign = 1; }
        if (!sr->golomb) {
            diff = rcd_symbol(&sr->rc, m->state[pcidx] + 32 * (size_t)ctx, 1, err);
        } else {
            if (ctx == 0 && run_mode == 0) run_mode = 1;
            if (run_mode) {
                if (run_count == 0 && run_mode == 1) {
                    if (br_get(&sr->br, 1)) {
                        run_count = 1 << log2_run[sr->run_index];
                        if (x + run_count <= P->w) sr->run_index++;
                    } else {
                        run_count = log2_run[sr->run_index] ? (int)br_get(&sr->br, log2_run[sr->run_index]) : 0;
                        if (sr->run_index) sr->run_index--;
                        run_mode = 2;
                    }
                }
                run_count--;
                if (run_count < 0) {
                    run_mode = 0; run_count = 0;
                    diff = get_vlc(&sr->br, &m->vlc[pcidx][ctx], bits);
                    if (diff >= 0) diff++;
                } else
                    diff = 0;
            } else
                diff = get_vlc(&sr->br, &m->vlc[pcidx][ctx], bits);
        }
        if (sign) diff = -diff;
        /* (predict + diff) mod 2^bits, stored into the int16 line buffer (ffv1dec.c:178) */
        P->s[(size_t)y * P->w + x] = (int16_t)((predict_at(P, x, y) + diff) & ((1 << bits) - 1));
    }
}

static void scatter_slice(const ffv1o_params *p, uint8_t *const planes[4], const int strides[4],
                          int x0, int y0, int w, int h, Plane16 pl[4], int npl, int by, int ry, int pcm)
{
    int x, y, k;
    if (p->colorspace == 0) {
        int idx = 0;
        for (k = 0; k < npl; k++) {
            int srcidx, pstep = 1, poff = 0, hs = 0, vs = 0;
            if (p->layout == FFV1O_LAYOUT_YA8) { srcidx = 0; pstep = 2; poff = k; }
            else if (k == 0) srcidx = 0;
            else if (p->chroma_planes && k <= 2) { srcidx = k; hs = p->chroma_h_shift; vs = p->chroma_v_shift; }
            else srcidx = 3;
            (void)idx;
            for (y = 0; y < pl[k].h; y++)
                for (x = 0; x < pl[k].w; x++) {
                    int v = pl[k].s[(size_t)y * pl[k].w + x];
                    uint8_t *dstp = planes[srcidx] + (size_t)((y0 >> vs) + y) * strides[srcidx];
                    if (p->bits <= 8) dstp[(size_t)((x0 >> hs) + x) * pstep + poff] = (uint8_t)v;
                    else {
                        unsigned u = (uint16_t)v;
                        if (!p->packed_at_lsb) u = (uint16_t)(u << (16 - p->bits));   /* ffv1dec.c:211-219 */
                        dstp[2 * (size_t)((x0 >> hs) + x)] = (uint8_t)u;
                        dstp[2 * (size_t)((x0 >> hs) + x) + 1] = (uint8_t)(u >> 8);
                    }
                }
        }
    } else {
        int offset = 1 << (p->bits <= 8 ? 8 : p->bits);                                 /* ffv1dec.c:226-280 */
        for (y = 0; y < h; y++)
            for (x = 0; x < w; x++) {
                int g = pl[0].s[(size_t)y * w + x], b = pl[1].s[(size_t)y * w + x], r = pl[2].s[(size_t)y * w + x];
                int a = npl == 4 ? pl[3].s[(size_t)y * w + x] : 0;
                if (!pcm) {                                                              /* ffv1dec.c:263-269 */
                    b -= offset; r -= offset;
                    g -= (b * by + r * ry) >> 2;
                    b += g; r += g;
                }
                if (p->layout == FFV1O_LAYOUT_RGB32) {
                    uint8_t *q = planes[0] + (size_t)(y0 + y) * strides[0] + 4 * (size_t)(x0 + x);
                    unsigned v = (unsigned)b + ((unsigned)g << 8) + ((unsigned)r << 16) + ((unsigned)a << 24);
                    q[0] = v; q[1] = v >> 8; q[2] = v >> 16; q[3] = v >> 24;
                } else {
                    uint8_t *q0 = planes[0] + (size_t)(y0 + y) * strides[0] + 2 * (size_t)(x0 + x);
                    uint8_t *q1 = planes[1] + (size_t)(y0 + y) * strides[1] + 2 * (size_t)(x0 + x);
                    uint8_t *q2 = planes[2] + (size_t)(y0 + y) * strides[2] + 2 * (size_t)(x0 + x);
                    q0[0] = b; q0[1] = b >> 8; q1[0] = g; q1[1] = g >> 8; q2[0] = r; q2[1] = r >> 8;
                }
            }
    }
}

static int read_v01_header(ffv1o_decoder *d, RcDec *c)
{
    /* ffv1dec.c:646-696 + 788-795 */
    ffv1o_params *p = &d->p;
    uint8_t st[32];
    int i, err = 0, v, cc;
    memset(st, 128, sizeof(st));
    v = rcd_symbol(c, st, 0, &err);
    if (v >= 2 || v < 0) return E_INVALIDDATA;
    p->version = v;
    p->ac = rcd_symbol(c, st, 0, &err);
    if (p->ac == FFV1O_AC_RANGE_CUSTOM)
        for (i = 1; i < 256; i++) p->state_transition[i] = (uint8_t)(rcd_symbol(c, st, 1, &err) + c->one_state[i]);
    else
        for (i = 1; i < 256; i++) p->state_transition[i] = c->one_state[i];
    p->colorspace = rcd_symbol(c, st, 0, &err);
    p->bits = p->version > 0 ? rcd_symbol(c, st, 0, &err) : 8;
    if (!p->bits) p->bits = 8;
    p->chroma_planes = rcd_get(c, st);
    p->chroma_h_shift = rcd_symbol(c, st, 0, &err);
    p->chroma_v_shift = rcd_symbol(c, st, 0, &err);
    p->transparency = rcd_get(c, st);
    p->plane_count = 2 + p->transparency;
    if ((unsigned)p->chroma_h_shift > 4U || (unsigned)p->chroma_v_shift > 4U || err) return E_INVALIDDATA;
    if ((i = derive_layout(p)) < 0) return i;
    cc = get_quant_tables(c, p->quant_tables[0]);
    if (cc < 0) return cc;
    p->context_count[0] = cc;
    p->context_model = 0;
    d->have_params = 1;
    return 0;
}

int ffv1o_decode_frame(ffv1o_decoder *d, const uint8_t *pkt, long size,
                       uint8_t *const planes[4], const int strides[4], int *key_frame,
                       uint64_t *damaged_mask)
{
    ffv1o_params *p = &d->p;
    RcDec c0;
    uint8_t keystate = 128;
    int key, si, trailer, err = 0;
    long ends[FFV1O_MAX_SLICES], starts[FFV1O_MAX_SLICES];
    const uint8_t *bp;

    if (damaged_mask) *damaged_mask = 0;
    memset(&c0, 0, sizeof(c0));
    rcd_init(&c0, pkt, size);
    ffv1o_default_state_tables(c0.zero_state, c0.one_state);
    key = rcd_get(&c0, &keystate);                                                   /* ffv1dec.c:924 */
    if (key) {
        d->key_frame_ok = 0;
        if (!d->have_params || p->version < 2) {
            int r = read_v01_header(d, &c0);
            if (r < 0) return r;
        }
        if (p->version >= 3) {                                                       /* 804-813 */
            const uint8_t *q = pkt + size;
            int n;
            trailer = 3 + 5 * !!p->ec;
            for (n = 0; n < FFV1O_MAX_SLICES && 3 < q - pkt; n++) {
                long sz = (q[-trailer] << 16) | (q[-trailer + 1] << 8) | q[-trailer + 2];
                if (sz + trailer > q - pkt) break;
                q -= sz + trailer;
            }
            d->slice_count = n;
        } else
            d->slice_count = d->max_slices;
        if (d->slice_count <= 0 || d->slice_count > d->max_slices) return E_INVALIDDATA;
        d->key_frame_ok = 1;
    } else if (!d->key_frame_ok)
        return E_INVALIDDATA;                                                        /* 930-935 */
    if (key_frame) *key_frame = key;

    /* 948-989: walk the footers back to front */
    trailer = 3 + 5 * !!p->ec;
    bp = pkt + size;
    for (si = d->slice_count - 1; si >= 0; si--) {
        long v;
        if (si || p->version > 2) v = ((bp[-trailer] << 16) | (bp[-trailer + 1] << 8) | bp[-trailer + 2]) + trailer;
        else v = bp - pkt;
        if (bp - pkt < v) return E_INVALIDDATA;
        bp -= v;
        starts[si] = bp - pkt; ends[si] = starts[si] + v;
        d->sm[si].damaged = 0;
        if (p->ec && ffv1o_crc32(0, bp, v)) d->sm[si].damaged = 1;
    }

    for (si = 0; si < d->slice_count; si++) {
        SliceModel *m = &d->sm[si];
        SliceReader sr;
        Plane16 pl[4];
        int pc[4], npl = 0, k, y, bits, nctx, qti[3] = { 0, 0, 0 }, golomb, reset_contexts = 0;
        const uint8_t *sbeg = pkt + starts[si];
        memset(&sr, 0, sizeof(sr));
        if (si == 0) { sr.rc = c0; sr.rc.end = pkt + ends[0]; }                      /* 984-987 */
        else { rcd_init(&sr.rc, sbeg, ends[si] - starts[si]); }
        load_transition(p, sr.rc.zero_state, sr.rc.one_state);
        if (p->ac != FFV1O_AC_RANGE_CUSTOM) ffv1o_default_state_tables(sr.rc.zero_state, sr.rc.one_state);

        if (p->version > 2) {                                                        /* 282-359 */
            uint8_t st[32];
            unsigned sx, sy, sw, sh, ps;
            memset(st, 128, sizeof(st));
            sx = (unsigned)rcd_symbol(&sr.rc, st, 0, &err) * p->width;
            sy = (unsigned)rcd_symbol(&sr.rc, st, 0, &err) * p->height;
            sw = ((unsigned)rcd_symbol(&sr.rc, st, 0, &err) + 1) * p->width + sx;
            sh = ((unsigned)rcd_symbol(&sr.rc, st, 0, &err) + 1) * p->height + sy;
            sx /= p->num_h_slices; sy /= p->num_v_slices;
            sw = sw / p->num_h_slices - sx; sh = sh / p->num_v_slices - sy;
            if (sw > (unsigned)p->width || sh > (unsigned)p->height ||
                (uint64_t)sx + sw > (unsigned)p->width || (uint64_t)sy + sh > (unsigned)p->height || err) {
                m->damaged = 1; continue;
            }
            m->x0 = sx; m->y0 = sy; m->w = sw; m->h = sh;
            for (k = 0; k < p->plane_count; k++) {
                int idx = rcd_symbol(&sr.rc, st, 0, &err);
                if ((unsigned)idx >= 2U) { m->damaged = 1; break; }
                qti[k] = idx;
            }
            if (m->damaged && k < p->plane_count) continue;
            ps = rcd_symbol(&sr.rc, st, 0, &err); (void)ps;
            rcd_symbol(&sr.rc, st, 0, &err);     /* SAR num */
            rcd_symbol(&sr.rc, st, 0, &err);     /* SAR den */
            m->coding_mode = 0; m->rct_by = m->rct_ry = 1; reset_contexts = 0;
            if (p->version > 3) {                /* ffv1dec.c:345-356 */
                reset_contexts = rcd_get(&sr.rc, st);
                m->coding_mode = rcd_symbol(&sr.rc, st, 0, &err);
                if (m->coding_mode != 1) {
                    m->rct_by = rcd_symbol(&sr.rc, st, 0, &err);
                    m->rct_ry = rcd_symbol(&sr.rc, st, 0, &err);
                    if ((uint64_t)(unsigned)m->rct_by + (uint64_t)(unsigned)m->rct_ry > 4) { m->damaged = 1; continue; }
                }
            }
        } else {
            m->coding_mode = 0; m->rct_by = m->rct_ry = 1; reset_contexts = 0;
            slice_geometry(p, si, &m->x0, &m->y0, &m->w, &m->h);
        }
        golomb = sr.golomb = (p->ac == FFV1O_AC_GOLOMB);
        for (k = 0; k < 3; k++) {
            int used = (k == 0) || (k == 1 && (p->chroma_planes || p->layout == FFV1O_LAYOUT_YA8)) ||
                       (k == 2 && p->transparency && p->layout != FFV1O_LAYOUT_YA8);
            int r;
            if (!used) continue;
            nctx = p->context_count[qti[k < p->plane_count ? k : 0]];
            if ((r = model_alloc(m, k, nctx, golomb)) < 0) return r;
            if (key || reset_contexts) model_reset(m, k, nctx, golomb);              /* 419-420 */
        }
        if (golomb) {                                                                /* 427-434 */
            long acb;
            if ((p->version == 3 && p->micro_version > 1) || p->version > 3) { uint8_t s129 = 129; rcd_get(&sr.rc, &s129); }
            acb = (p->version > 2 || (!m->x0 && !m->y0)) ? (sr.rc.ptr - sr.rc.start - 1) : 0;
            sr.br.buf = sr.rc.start + acb;
            sr.br.nbits = (uint64_t)(sr.rc.end - sr.rc.start - acb) * 8;
            sr.br.bitpos = 0;
        }
        bits = coded_bits(p);
        /* allocate zeroed sample arrays with the encoder's geometry */
        {
            Plane16 tmp[4]; int tpc[4];
            /* geometry only: reuse gather_slice's layout rules through a tiny re-derivation */
            int n = 0;
            if (p->colorspace == 0) {
                if (p->layout == FFV1O_LAYOUT_YA8) { tmp[0].w = tmp[1].w = m->w; tmp[0].h = tmp[1].h = m->h; tpc[0] = 0; tpc[1] = 1; n = 2; }
                else {
                    tmp[n].w = m->w; tmp[n].h = m->h; tpc[n] = 0; n++;
                    if (p->chroma_planes) {
                        int cw = -((-m->w) >> p->chroma_h_shift), chh = -((-m->h) >> p->chroma_v_shift), j;
                        for (j = 0; j < 2; j++) { tmp[n].w = cw; tmp[n].h = chh; tpc[n] = 1; n++; }
                    }
                    if (p->transparency) { tmp[n].w = m->w; tmp[n].h = m->h; tpc[n] = 2; n++; }
                }
            } else {
                n = 3 + (p->transparency ? 1 : 0);
                for (k = 0; k < n; k++) { tmp[k].w = m->w; tmp[k].h = m->h; tpc[k] = (k + 1) / 2; }
            }
            npl = n;
            for (k = 0; k < n; k++) {
                pl[k] = tmp[k]; pc[k] = tpc[k];
                pl[k].s = calloc((size_t)pl[k].w * pl[k].h + 1, sizeof(int16_t));
                if (!pl[k].s) return E_NOMEM;
            }
        }
        sr.pcm = m->coding_mode == 1;
        if (p->colorspace == 0) {
            for (k = 0; k < npl; k++) {
                sr.run_index = 0;
                for (y = 0; y < pl[k].h; y++)
                    decode_line(&sr, m, pc[k], p->quant_tables[qti[pc[k] < p->plane_count ? pc[k] : 0]], &pl[k], y, bits, &err);
            }
        } else {
            sr.run_index = 0;
            for (y = 0; y < m->h; y++)
                for (k = 0; k < npl; k++)
                    decode_line(&sr, m, pc[k], p->quant_tables[qti[pc[k] < p->plane_count ? pc[k] : 0]], &pl[k], y, bits - sr.pcm, &err);
        }
        if (!golomb && p->version > 2) {                                             /* 459-467 */
            uint8_t s129 = 129;
            long v;
            rcd_get(&sr.rc, &s129);
            v = sr.rc.end - sr.rc.ptr - 2 - 5 * p->ec;
            if (v) m->damaged = 1;
        }
        scatter_slice(p, planes, strides, m->x0, m->y0, m->w, m->h, pl, npl, m->rct_by, m->rct_ry, sr.pcm);
        for (k = 0; k < npl; k++) free(pl[k].s);
    }
    if (damaged_mask)
        for (si = 0; si < d->slice_count && si < 64; si++)
            if (d->sm[si].damaged) *damaged_mask |= (uint64_t)1 << si;
    return err ? E_INVALIDDATA : 0;
}
