// ffv1_model.cpp -- see ffv1_model.h.  Host-only; no CUDA here.
#include "ffv1_model.h"
#include "../../include/ffv1_b200.h"
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cinttypes>

namespace ffv1 {

// ------------------------------------------------------------------------------------------------
// constant tables of the bitstream format
// ------------------------------------------------------------------------------------------------

// Context quantisers (ffv1enc.c:44-118) are odd step functions; {first |d| of level 1, 2, ...}.
struct QuantCurve { int nlevels; int start[6]; };
static const QuantCurve kQ11   = {5, {1, 2, 5, 12, 35}};       // quant11      : 11 levels, 8-bit
static const QuantCurve kQ5    = {2, {1, 4}};                  // quant5       :  5 levels, 8-bit
static const QuantCurve kQ9hi  = {4, {5, 13, 27, 56}};         // quant9_10bit :  9 levels, >8-bit
static const QuantCurve kQ5hi  = {2, {11, 50}};                // quant5_10bit :  5 levels, >8-bit

static void expand_curve(const QuantCurve &q, int scale, int16_t out[256])
{
    for (int d = 0; d < 128; d++) {
        int lvl = 0;
        while (lvl < q.nlevels && d >= q.start[lvl]) lvl++;
        out[d] = (int16_t)(lvl * scale);
    }
    for (int d = 1; d < 128; d++) out[256 - d] = (int16_t)-out[d];
    out[128] = (int16_t)-out[127];
}

// "ver2_state": the state-transition table selected by coder=1/2 (ffv1enc.c:120-137), stored as the
// delta to the identity (one_state[i] - i) to keep the listing compact.
static const int8_t kCustomDelta[256] = {
    0,  9,  8,  7,  6, 11, 10,  9, 20,  7,  6, 18, 30, 36,  6, 34, 43,  8,  8,  7,  7, 10, 11, 10,  9,  9,  8, 10, 39,  9,  9,  8,
    8,  7,  7, 44,  7,  7,  7,  6,  8,  7, 22,  7,  7,  7, 42,  5,  5, 25,  5,  6,  6,  5, 20,  5, 45,  4,  4, 25,  6,  5,  6,  6,
   23, 17,  5, 30,  5,  4, 12,  4, 39,  4, 20,  3, 11,  4,  5, 18,  5,  2, 12,  3, 15,  4,  4, 12, 23,  3,  3, 43,  3,  5, 11,  3,
    9, 13,  4,  9,  2, 17,  1,  3,  2,  8,  3,  5,  6,  3,  6, 14,  3,  3,  3,  2, 10,  2,  7,  2,  1,  2, 23,  1,  2,  6,  1,  2,
   37,  1,  2,  7,  1,  2, 11,  1,  1,  2,  8,  2,  3,  1,  2,  5,  3, 10,  5,  2,  3,  1,  2,  6,  1,  1,  2, 13,  2,  5,  3,  1,
   12,  2,  7,  1,  2, 19,  1,  3,  9,  5,  1,  2, 10,  3,  6,  3, -1, 12,  1,  2,  6,  2, 10,  2, 16,  2,  5,  1,  2,  8,  3,  5,
    5,  1,  1,  1,  2,  5,  1,  2, 10,  2,  5,  1,  1,  1,  2,  7,  1,  2, 11,  1,  1,  2, 10,  1,  1,  1,  1,  1,  2,  7,  1,  2,
    2, -1,  1,  2, 12,  1,  1,  1,  1,  1,  1,  1,  2,  2, -1,  3,  1,  2,  0,  1,  1,  1,  1,  1,  1,  1,  1,  1,  0,  0,  0,  0,
};

void default_state_tables(uint8_t zero_state[256], uint8_t one_state[256])
{
    // rangecoder.c:63-101 with FFV1's arguments factor=(int)(0.05*2^32), max_p=256-8 (ffv1enc.c:562,1288)
    const int64_t kOne = int64_t(1) << 32;
    const int64_t kFactor = (int64_t)(int)(0.05 * (double)kOne);
    const int kMaxP = 248;
    memset(zero_state, 0, 256);
    memset(one_state, 0, 256);
    int64_t prob = kOne / 2;
    int last = 0;
    for (int i = 0; i < 128; i++) {
        int p8 = (int)((256 * prob + kOne / 2) >> 32);
        p8 = std::max(p8, last + 1);
        if (last && last < 256 && p8 <= kMaxP) one_state[last] = (uint8_t)p8;
        prob += ((kOne - prob) * kFactor + kOne / 2) >> 32;
        last = p8;
    }
    for (int i = 256 - kMaxP; i <= kMaxP; i++) {
        if (one_state[i]) continue;
        prob = (i * kOne + 128) >> 8;
        prob += ((kOne - prob) * kFactor + kOne / 2) >> 32;
        int p8 = (int)((256 * prob + kOne / 2) >> 32);
        p8 = std::min(std::max(p8, i + 1), kMaxP);
        one_state[i] = (uint8_t)p8;
    }
    for (int i = 1; i < 255; i++) zero_state[i] = (uint8_t)(256 - one_state[256 - i]);
}

void coder_state_tables(const Config &c, uint8_t zero_state[256], uint8_t one_state[256])
{
    // ffv1enc.c:1288 then 1309-1315 (slice 0) / ffv1.c:95-100 (other slices)
    default_state_tables(zero_state, one_state);
    if (c.ac == AC_RANGE_CUSTOM)
        for (int j = 1; j < 256; j++) {
            one_state[j] = c.state_transition[j];
            zero_state[256 - j] = (uint8_t)(256 - one_state[j]);
        }
}

uint32_t crc32_ieee(uint32_t crc, const uint8_t *buf, size_t len)
{
    // libavutil AV_CRC_32_IEEE (crc.c:303,357-380): MSB-first 0x04C11DB7, init 0, no final xor
    static uint32_t tab[256];
    static bool init = false;
    if (!init) {
        for (uint32_t n = 0; n < 256; n++) {
            uint32_t c = n << 24;
            for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : (c << 1);
            tab[n] = c;
        }
        init = true;
    }
    while (len--) crc = (crc << 8) ^ tab[(crc >> 24) ^ *buf++];
    return crc;
}

// ------------------------------------------------------------------------------------------------
// host binary coder: either produces range-coder bytes (extradata, golomb-mode slice prefixes) or just
// records the (probability, bit) decisions for the device coder to consume (range-mode slice prefixes).
// rangecoder.h:52-102, rangecoder.c:42-51,104-116.
// ------------------------------------------------------------------------------------------------
class BinCoder {
public:
    uint8_t zero_state[256], one_state[256];
    std::vector<uint8_t> bytes;
    std::vector<uint16_t> decisions;
    bool record_only = false;

    BinCoder() { default_state_tables(zero_state, one_state); }

    void put(uint8_t *state, int bit)
    {
        decisions.push_back((uint16_t)(*state | (bit ? 0x100 : 0)));
        if (!record_only) {
            int r1 = (range_ * *state) >> 8;
            if (bit) { low_ += range_ - r1; range_ = r1; } else range_ -= r1;
            renorm();
        }
        *state = bit ? one_state[*state] : zero_state[*state];
    }
    void put_symbol(uint8_t *st, int v, bool is_signed)
    {
        // ffv1enc.c:185-231
        if (v == 0) { put(st, 1); return; }
        const int a = std::abs(v);
        int e = 0;
        while ((a >> (e + 1)) != 0) e++;
        put(st, 0);
        for (int i = 0; i < e; i++) put(st + 1 + std::min(i, 9), 1);
        put(st + 1 + std::min(e, 9), 0);
        for (int i = e - 1; i >= 0; i--) put(st + 22 + std::min(i, 9), (a >> i) & 1);
        if (is_signed) put(st + 11 + std::min(e, 10), v < 0);
    }
    void terminate()
    {
        range_ = 0xFF; low_ += 0xFF; renorm();
        range_ = 0xFF; renorm();
    }
private:
    int low_ = 0, range_ = 0xFF00, out_count_ = 0, out_byte_ = -1;
    void renorm()
    {
        while (range_ < 0x100) {
            if (out_byte_ < 0) out_byte_ = low_ >> 8;
            else if (low_ <= 0xFF00) { bytes.push_back((uint8_t)out_byte_); flush(0xFF); out_byte_ = low_ >> 8; }
            else if (low_ >= 0x10000) { bytes.push_back((uint8_t)(out_byte_ + 1)); flush(0x00); out_byte_ = (low_ >> 8) & 0xFF; }
            else out_count_++;
            low_ = (low_ & 0xFF) << 8;
            range_ <<= 8;
        }
    }
    void flush(uint8_t v) { for (; out_count_; out_count_--) bytes.push_back(v); }
};

// range decoder for the extradata record (rangecoder.h:104-145)
class BinDecoder {
public:
    uint8_t zero_state[256], one_state[256];
    BinDecoder(const uint8_t *d, int n) : p_(d + 2), end_(d + n)
    {
        default_state_tables(zero_state, one_state);
        low_ = n >= 2 ? (d[0] << 8 | d[1]) : 0;
    }
    void shrink_end(int n) { end_ -= n; }
    int get(uint8_t *state)
    {
        int r1 = (range_ * *state) >> 8, bit;
        range_ -= r1;
        if (low_ < range_) { *state = zero_state[*state]; bit = 0; }
        else { low_ -= range_; *state = one_state[*state]; range_ = r1; bit = 1; }
        if (range_ < 0x100) {
            range_ <<= 8; low_ <<= 8;
            if (p_ < end_) low_ += *p_;
            p_++;
        }
        return bit;
    }
    int get_symbol(uint8_t *st, bool is_signed)
    {
        // ffv1dec.c:42-63
        if (get(st)) return 0;
        int e = 0;
        while (get(st + 1 + std::min(e, 9))) { if (++e > 31) { bad = true; return 0; } }
        int a = 1;
        for (int i = e - 1; i >= 0; i--) a += a + get(st + 22 + std::min(i, 9));
        if (is_signed && get(st + 11 + std::min(e, 10))) return -a;
        return a;
    }
    bool bad = false;
    int low() const { return low_; }
    int range() const { return range_; }
    const uint8_t *pos() const { return p_; }
private:
    int low_ = 0, range_ = 0xFF00;
    const uint8_t *p_, *end_;
};

// ------------------------------------------------------------------------------------------------
// pixel formats (ffv1enc.c:720-820 / 1425-1439)
// ------------------------------------------------------------------------------------------------
static bool parse_pix_fmt(const std::string &name, Config &c)
{
    c.pix_fmt = name;
    c.colorspace = 0; c.bits = 8; c.bytes_per_sample = 1; c.packed_at_lsb = 0; c.ya8 = 0;
    c.chroma_planes = 0; c.transparency = 0; c.chroma_h_shift = c.chroma_v_shift = 0;
    for (int i = 0; i < 4; i++) { c.src_hshift[i] = c.src_vshift[i] = 0; c.pixel_bytes[i] = 1; }

    if (name == "bgr0" || name == "bgra") {           // AV_PIX_FMT_0RGB32 / RGB32 on little endian
        c.colorspace = 1; c.chroma_planes = 1; c.transparency = (name == "bgra");
        c.nb_src_planes = 1; c.pixel_bytes[0] = 4;
        return true;
    }
    if (name == "ya8") { c.transparency = 1; c.ya8 = 1; c.nb_src_planes = 1; c.pixel_bytes[0] = 2; return true; }
    if (name == "gray") { c.nb_src_planes = 1; return true; }
    if (name == "gray16le") { c.bits = 16; c.bytes_per_sample = 2; c.nb_src_planes = 1; c.pixel_bytes[0] = 2; return true; }
    if (name.compare(0, 4, "gbrp") == 0) {
        int bits = atoi(name.c_str() + 4);
        if ((bits != 9 && bits != 10 && bits != 12 && bits != 14) || name != "gbrp" + std::to_string(bits) + "le") return false;
        c.colorspace = 1; c.chroma_planes = 1; c.bits = bits; c.bytes_per_sample = 2; c.nb_src_planes = 3;
        for (int i = 0; i < 3; i++) c.pixel_bytes[i] = 2;
        return true;
    }
    // yuv[a]4xyp[9|10|16le]
    size_t pos = 0;
    bool alpha = false;
    if (name.compare(0, 4, "yuva") == 0) { alpha = true; pos = 4; }
    else if (name.compare(0, 3, "yuv") == 0) pos = 3;
    else return false;
    if (name.size() < pos + 4 || name[pos + 3] != 'p') return false;
    std::string sub = name.substr(pos, 3), depth = name.substr(pos + 4);
    int hs, vs;
    if (sub == "420") { hs = 1; vs = 1; } else if (sub == "422") { hs = 1; vs = 0; }
    else if (sub == "444") { hs = 0; vs = 0; } else if (sub == "440") { hs = 0; vs = 1; }
    else if (sub == "411") { hs = 2; vs = 0; } else if (sub == "410") { hs = 2; vs = 2; }
    else return false;
    int bits = 8;
    if (!depth.empty()) {
        if (depth == "9le") bits = 9; else if (depth == "10le") bits = 10; else if (depth == "16le") bits = 16; else return false;
        if (sub == "440" || sub == "411" || sub == "410") return false;       // not in ff_ffv1_encoder.pix_fmts
    }
    if (alpha && (sub == "440" || sub == "411" || sub == "410")) return false;
    c.bits = bits; c.bytes_per_sample = bits > 8 ? 2 : 1;
    c.packed_at_lsb = (bits == 9 || bits == 10);
    c.chroma_planes = 1; c.transparency = alpha;
    c.chroma_h_shift = hs; c.chroma_v_shift = vs;
    c.nb_src_planes = alpha ? 4 : 3;
    c.src_hshift[1] = c.src_hshift[2] = hs; c.src_vshift[1] = c.src_vshift[2] = vs;
    for (int i = 0; i < 4; i++) c.pixel_bytes[i] = c.bytes_per_sample;
    return true;
}

void Config::plane_dims(int i, int *rows, int *row_bytes) const
{
    int cw = -((-width) >> src_hshift[i]), ch = -((-height) >> src_vshift[i]);
    *rows = ch; *row_bytes = cw * pixel_bytes[i];
}
int64_t Config::frame_bytes() const
{
    int64_t n = 0;
    for (int i = 0; i < nb_src_planes; i++) { int r, b; plane_dims(i, &r, &b); n += (int64_t)r * b; }
    return n;
}

static void fill_quant_tables(Config &c)
{
    // ffv1enc.c:846-871
    const QuantCurve &a = c.bits <= 8 ? kQ11 : kQ9hi, &b = c.bits <= 8 ? kQ5 : kQ5hi;
    memset(c.quant_tables, 0, sizeof(c.quant_tables));
    expand_curve(a, 1, c.quant_tables[0][0]);
    expand_curve(a, 11, c.quant_tables[0][1]);
    expand_curve(a, 11 * 11, c.quant_tables[0][2]);
    expand_curve(a, 1, c.quant_tables[1][0]);
    expand_curve(a, 11, c.quant_tables[1][1]);
    expand_curve(b, 11 * 11, c.quant_tables[1][2]);
    expand_curve(b, 5 * 11 * 11, c.quant_tables[1][3]);
    expand_curve(b, 5 * 5 * 11 * 11, c.quant_tables[1][4]);
    c.context_count[0] = (11 * 11 * 11 + 1) / 2;
    c.context_count[1] = (11 * 11 * 5 * 5 * 5 + 1) / 2;
}

int resolve_encoder(const EncOptions &o, Config &c, std::string &err)
{
    c = Config();
    if (o.width <= 0 || o.height <= 0) { err = "invalid picture size"; return FFV1B200_ERR_INVALIDDATA; }  // ffv1.c:46-47
    if (o.context < 0 || o.context > 1) { err = "Invalid context model, valid values are 0 and 1"; return FFV1B200_ERR_EINVAL; }
    if (o.coder < -2 || o.coder > 2 || o.slicecrc < -1 || o.slicecrc > 1) { err = "option out of range"; return FFV1B200_ERR_EINVAL; }
    c.width = o.width; c.height = o.height;
    if (!parse_pix_fmt(o.pix_fmt, c)) { err = "format not supported"; return FFV1B200_ERR_ENOSYS; }        // ffv1enc.c:816-818

    // a caller-set bits_per_raw_sample replaces the depth of 16-bit containers (ffv1enc.c:728-748, 796-805); 8-bit planar
    // formats ignore it, and packed RGB would trip the reference's av_assert0(bits_per_raw_sample >= 8)
    if (o.bits_per_raw_sample) {
        if (c.bits > 8) {
            if (o.bits_per_raw_sample <= 8 || o.bits_per_raw_sample > 16) { err = "bits_per_raw_sample invalid"; return FFV1B200_ERR_INVALIDDATA; }
            c.bits = o.bits_per_raw_sample;
        } else if (c.colorspace == 1) { err = "bits_per_raw_sample cannot be set for packed RGB"; return FFV1B200_ERR_EINVAL; }
    }

    // version selection, ffv1enc.c:676-706
    int version = 0;
    if (o.slices > 1 || (o.pass_flags & (kPass1 | kPass2))) version = 2;
    if (o.slices == 0 && o.level < 0 && o.width * o.height > 720 * 576) version = 2;
    if (o.level <= 0 && version == 2) version = 3;
    if (o.level >= 0 && o.level <= 4) {
        if (o.level < version) { err = "Version " + std::to_string(version) + " needed for requested features but " + std::to_string(o.level) + " requested"; return FFV1B200_ERR_EINVAL; }
        version = o.level;
    }
    c.ec = o.slicecrc < 0 ? (version >= 3) : o.slicecrc;
    // ffv1enc.c:703-706: versions 2 and 4 are experimental.  Version 4 (micro version 2: per-slice RCT coefficients, slice
    // coding mode, context reset flag) is provided behind -strict experimental; version 2 is not.
    if ((version == 2 || version > 3) && !o.strict_experimental) {
        err = "Version 2 needed for requested features but version 2 is experimental and not enabled";
        return FFV1B200_ERR_INVALIDDATA;
    }
    if (version == 2) { err = "FFV1 version 2 (an abandoned experimental bitstream) is not provided"; return FFV1B200_ERR_ENOSYS; }

    // coder remaps, ffv1enc.c:708-718, 755-759, 810-814
    int ac = o.coder;
    if (ac == 1) ac = AC_RANGE_CUSTOM;
    else if (ac == -2) ac = AC_RANGE_DEFAULT;
    else if (ac == -1) ac = AC_GOLOMB;
    const bool high_bits = c.bits > 8;
    if (high_bits) {
        if (ac == AC_GOLOMB) ac = AC_RANGE_CUSTOM;       // "bits_per_raw_sample > 8, forcing range coder"
        version = std::max(version, 1);
    }
    c.ac = ac;
    c.version = version;
    c.micro_version = version == 3 ? 4 : (version == 4 ? 2 : 0);              // ffv1enc.c:565-569
    c.context_model = o.context;
    c.gop_size = o.gop_size;
    c.intra = o.gop_size < 2;

    if (c.ac == AC_RANGE_CUSTOM) {
        for (int i = 1; i < 256; i++) c.state_transition[i] = (uint8_t)(i + kCustomDelta[i]);
    } else {
        uint8_t z[256], one[256];
        default_state_tables(z, one);
        for (int i = 1; i < 256; i++) c.state_transition[i] = one[i];
    }
    fill_quant_tables(c);
    c.plane_count = c.transparency ? 3 : 2;               // ffv1enc.c:720, 890-893
    if (!c.chroma_planes && version > 3) c.plane_count--;

    c.num_h_slices = c.num_v_slices = 1;
    if (version > 1) {                                    // ffv1enc.c:988-1000
        bool found = false;
        for (int nv = (o.width > 352 || o.height > 288 || !o.slices) ? 2 : 1; nv < 9 && !found; nv++)
            for (int nh = nv; nh < 2 * nv; nh++)
                if ((o.slices == nh * nv && o.slices <= 64) || !o.slices) { c.num_h_slices = nh; c.num_v_slices = nv; found = true; break; }
        if (!found) {
            err = "Unsupported number " + std::to_string(o.slices) + " of slices requested, please specify a supported number with -slices (ex:4,6,9,12,16, ...)";
            return FFV1B200_ERR_ENOSYS;
        }
    }
    c.pass_flags = o.pass_flags;
    if (!o.stats_in.empty()) {                            // ffv1enc.c:906-986 (runs before the slice grid is chosen; independent of it)
        if (c.version < 2) { err = "two-pass statistics need FFV1 version 2 or later"; return FFV1B200_ERR_EINVAL; }
        int r = apply_stats(c, o.stats_in, err);
        if (r < 0) return r;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// two-pass coding: statistics text, state-transition table sorting, initial states
// ------------------------------------------------------------------------------------------------
std::string format_stats(const Config &c, const PassStats &st)
{
    // the text encode_frame leaves in stats_out when it is flushed (ffv1enc.c:1261-1275)
    std::string out;
    char tmp[64];
    for (int j = 0; j < 256; j++) {
        snprintf(tmp, sizeof(tmp), "%" PRIu64 " %" PRIu64 " ", st.rc_stat[j][0], st.rc_stat[j][1]);
        out += tmp;
    }
    // (the reference writes a newline here and overwrites it with the next number: its pointer is not advanced, ffv1enc.c:1266)
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < c.context_count[i]; j++)
            for (int m = 0; m < 32; m++) {
                const uint64_t *v = &st.rc_stat2[i][((size_t)j * 32 + m) * 2];
                snprintf(tmp, sizeof(tmp), "%" PRIu64 " %" PRIu64 " ", v[0], v[1]);
                out += tmp;
            }
    snprintf(tmp, sizeof(tmp), "%d\n", st.gob_count);
    out += tmp;
    return out;
}

namespace {
// cost in bits of coding the decisions counted in `n` (n[0] zeros, n[1] ones) with probability state `state`
inline double stat_term0(const uint64_t n[2], int state) { return n[0] * -log2((256 - state) / 256.0); }
inline double stat_term1(const uint64_t n[2], int state) { return n[1] * -log2(state / 256.0); }

// sort_stt (ffv1enc.c:621-667): neighbouring entries of the transition table are exchanged while that shortens the
// first pass's decisions.  The sums keep the reference's left-to-right order of the eight terms (its cost macros expand
// without parentheses), and the counters are exchanged through an int like its FFSWAP(int, ...).
void sort_transition_table(uint64_t rc_stat[256][2], uint8_t stt[256])
{
    auto cost8 = [&](int a, int an, int b, int bn) {
        // COST2(a, an) + COST2(b, bn)
        double s = stat_term0(rc_stat[a], an);
        s = s + stat_term1(rc_stat[a], an);
        s = s + stat_term0(rc_stat[256 - a], 256 - an);
        s = s + stat_term1(rc_stat[256 - a], 256 - an);
        s = s + stat_term0(rc_stat[b], bn);
        s = s + stat_term1(rc_stat[b], bn);
        s = s + stat_term0(rc_stat[256 - b], 256 - bn);
        s = s + stat_term1(rc_stat[256 - b], 256 - bn);
        return s;
    };
    auto swap_counts = [&](uint64_t &x, uint64_t &y) { const int t = (int)y; y = x; x = (uint64_t)(int64_t)t; };
    auto swap_u8 = [&](uint8_t &x, uint8_t &y) { const uint8_t t = y; y = x; x = t; };
    bool changed;
    do {
        changed = false;
        for (int i = 12; i < 244; i++)
            for (int i2 = i + 1; i2 < 245 && i2 < i + 4; i2++) {
                const double keep = cost8(i, i, i2, i2), swapped = cost8(i, i2, i2, i);
                if (!(keep - swapped > keep * (1e-14)) || i == 128 || i2 == 128) continue;
                swap_u8(stt[i], stt[i2]);
                swap_counts(rc_stat[i][0], rc_stat[i2][0]);
                swap_counts(rc_stat[i][1], rc_stat[i2][1]);
                const bool mirror = i != 256 - i2;
                if (mirror) {
                    swap_u8(stt[256 - i], stt[256 - i2]);
                    swap_counts(rc_stat[256 - i][0], rc_stat[256 - i2][0]);
                    swap_counts(rc_stat[256 - i][1], rc_stat[256 - i2][1]);
                }
                for (int j = 1; j < 256; j++) {
                    if (stt[j] == i) stt[j] = (uint8_t)i2;
                    else if (stt[j] == i2) stt[j] = (uint8_t)i;
                    if (mirror) {
                        if (stt[256 - j] == 256 - i) stt[256 - j] = (uint8_t)(256 - i2);
                        else if (stt[256 - j] == 256 - i2) stt[256 - j] = (uint8_t)(256 - i);
                    }
                }
                changed = true;
            }
    } while (changed);
}

// find_best_state (ffv1enc.c:139-183): best[i][k] = the start state from which k further decisions of a source with
// P(one) = i/256 cost least, found by following the occupancy of the states through the transition table.
void best_start_states(std::vector<uint8_t> &best, const uint8_t one_state[256])
{
    best.assign(256 * 256, 0);
    double l2[256];
    l2[0] = 0.0;
    for (int i = 1; i < 256; i++) l2[i] = log2(i / 256.0);
    std::vector<double> shortest(256);
    for (int i = 0; i < 256; i++) {
        const double p = i / 256.0;
        std::fill(shortest.begin(), shortest.end(), (double)(1 << 30));
        for (int j = std::max(i - 10, 1); j < std::min(i + 11, 256); j++) {
            if (!one_state[j]) continue;
            double occ[256] = {0}, next[256];
            double len = 0;
            occ[j] = 1.0;
            for (int k = 0; k < 256; k++) {
                for (int m = 1; m < 256; m++)
                    if (occ[m]) len -= occ[m] * (p * l2[m] + (1 - p) * l2[256 - m]);
                if (len < shortest[k]) { shortest[k] = len; best[(size_t)i * 256 + k] = (uint8_t)j; }
                std::fill(next, next + 256, 0.0);
                for (int m = 1; m < 256; m++)
                    if (occ[m]) {
                        next[one_state[m]] += occ[m] * p;
                        next[256 - one_state[256 - m]] += occ[m] * (1 - p);
                    }
                memcpy(occ, next, sizeof(occ));
            }
        }
    }
}

inline int clip_int(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline int clip_u8(int v) { return (v & ~0xFF) ? (~v >> 31) & 0xFF : v; }
} // namespace

int apply_stats(Config &c, const std::string &stats_in, std::string &err)
{
    // parse (ffv1enc.c:906-948): blocks of 256x2 + contexts x 32 x 2 counters + gob_count; the last block counts
    uint64_t rc_stat[256][2];
    std::vector<uint64_t> rc_stat2[2];
    for (int i = 0; i < 2; i++) rc_stat2[i].assign((size_t)c.context_count[i] * 64, 0);
    int gob_count = 0;
    const char *p = stats_in.c_str();
    char *next;
    for (;;) {
        for (int j = 0; j < 256; j++)
            for (int i = 0; i < 2; i++) {
                rc_stat[j][i] = (uint64_t)strtol(p, &next, 0);
                if (next == p) { err = "2Pass file invalid at " + std::to_string(j) + " " + std::to_string(i); return FFV1B200_ERR_INVALIDDATA; }
                p = next;
            }
        for (int i = 0; i < 2; i++)
            for (size_t n = 0; n < rc_stat2[i].size(); n++) {
                rc_stat2[i][n] = (uint64_t)strtol(p, &next, 0);
                if (next == p) { err = "2Pass file invalid (context statistics)"; return FFV1B200_ERR_INVALIDDATA; }
                p = next;
            }
        gob_count = (int)strtol(p, &next, 0);
        if (next == p || gob_count <= 0) { err = "2Pass file invalid"; return FFV1B200_ERR_INVALIDDATA; }
        p = next;
        while (*p == '\n' || *p == ' ') p++;
        if (!*p) break;
    }
    if (c.ac == AC_RANGE_CUSTOM) sort_transition_table(rc_stat, c.state_transition);
    std::vector<uint8_t> best;
    best_start_states(best, c.state_transition);
    // initial states (ffv1enc.c:954-984): contexts with few observations share the estimate of the ones before them
    for (int i = 0; i < 2; i++) {
        const int nctx = c.context_count[i];
        std::vector<uint8_t> &init = c.initial_states[i];
        init.assign((size_t)nctx * 32, 128);
        for (int k = 0; k < 32; k++) {
            double a = 0, b = 0;
            int jp = 0;
            for (int j = 0; j < nctx; j++) {
                const uint64_t n0 = rc_stat2[i][((size_t)j * 32 + k) * 2], n1 = rc_stat2[i][((size_t)j * 32 + k) * 2 + 1];
                double pr = 128;
                if ((n0 + n1 > 200 && j) || a + b > 200) {
                    if (a + b) pr = 256.0 * b / (a + b);
                    init[(size_t)jp * 32 + k] = best[(size_t)clip_int((int)round(pr), 1, 255) * 256 + clip_u8((int)((a + b) / gob_count))];
                    for (jp++; jp < j; jp++) init[(size_t)jp * 32 + k] = init[(size_t)(jp - 1) * 32 + k];
                    a = b = 0;
                }
                a += n0;
                b += n1;
                if (a + b) pr = 256.0 * b / (a + b);
                init[(size_t)j * 32 + k] = best[(size_t)clip_int((int)round(pr), 1, 255) * 256 + clip_u8((int)((a + b) / gob_count))];
            }
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// extradata
// ------------------------------------------------------------------------------------------------
static void put_quant_table(BinCoder &bc, const int16_t *q)
{
    // ffv1enc.c:475-488
    uint8_t st[kStateSlots];
    memset(st, 128, sizeof(st));
    int last = 0;
    for (int i = 1; i < 128; i++)
        if (q[i] != q[i - 1]) { bc.put_symbol(st, i - last - 1, false); last = i; }
    bc.put_symbol(st, 128 - last - 1, false);
}

std::vector<uint8_t> write_extradata(const Config &c)
{
    // ffv1enc.c:545-619 (version >= 2 only; versions 0/1 carry the header in band)
    if (c.version < 2) return {};
    BinCoder bc;
    uint8_t st[kStateSlots];
    memset(st, 128, sizeof(st));
    bc.put_symbol(st, c.version, false);
    if (c.version > 2) bc.put_symbol(st, c.micro_version, false);
    bc.put_symbol(st, c.ac, false);
    if (c.ac == AC_RANGE_CUSTOM)
        for (int i = 1; i < 256; i++) bc.put_symbol(st, c.state_transition[i] - bc.one_state[i], true);
    bc.put_symbol(st, c.colorspace, false);
    bc.put_symbol(st, c.bits, false);
    bc.put(st, c.chroma_planes);
    bc.put_symbol(st, c.chroma_h_shift, false);
    bc.put_symbol(st, c.chroma_v_shift, false);
    bc.put(st, c.transparency);
    bc.put_symbol(st, c.num_h_slices - 1, false);
    bc.put_symbol(st, c.num_v_slices - 1, false);
    bc.put_symbol(st, 2, false);
    for (int t = 0; t < 2; t++)
        for (int i = 0; i < 5; i++) put_quant_table(bc, c.quant_tables[t][i]);
    uint8_t st2[32][kStateSlots];                          // ffv1enc.c:591-606: one coder state row per slot, shared by both sets
    memset(st2, 128, sizeof(st2));
    for (int t = 0; t < 2; t++) {
        const std::vector<uint8_t> &init = c.initial_states[t];
        bool coded = false;
        for (uint8_t v : init) if (v != 128) { coded = true; break; }
        bc.put(st, coded);
        if (!coded) continue;
        for (int j = 0; j < c.context_count[t]; j++)
            for (int k = 0; k < 32; k++) {
                const int pred = j ? init[(size_t)(j - 1) * 32 + k] : 128;
                bc.put_symbol(st2[k], (int8_t)(init[(size_t)j * 32 + k] - pred), true);
            }
    }
    if (c.version > 2) {
        bc.put_symbol(st, c.ec, false);
        bc.put_symbol(st, c.intra, false);
    }
    bc.terminate();
    std::vector<uint8_t> out = bc.bytes;
    uint32_t crc = crc32_ieee(0, out.data(), out.size());
    for (int s = 24; s >= 0; s -= 8) out.push_back((uint8_t)(crc >> s));
    return out;
}

static int read_quant_table(BinDecoder &bd, int16_t *q, int scale)
{
    // ffv1dec.c:476-500
    uint8_t st[kStateSlots];
    memset(st, 128, sizeof(st));
    int i = 0, v = 0;
    for (; i < 128; v++) {
        unsigned len = (unsigned)bd.get_symbol(st, false) + 1;
        if (bd.bad || len > (unsigned)(128 - i) || !len) return -1;
        while (len--) q[i++] = (int16_t)(scale * v);
    }
    for (i = 1; i < 128; i++) q[256 - i] = (int16_t)-q[i];
    q[128] = (int16_t)-q[127];
    return 2 * v - 1;
}

static bool select_pix_fmt(Config &c)
{
    // ffv1dec.c:698-786
    std::string name;
    const int sub = 16 * c.chroma_h_shift + c.chroma_v_shift;
    auto yuv = [&](bool alpha) -> std::string {
        const char *s = nullptr;
        switch (sub) { case 0x00: s = "444"; break; case 0x01: s = "440"; break; case 0x10: s = "422"; break;
                       case 0x11: s = "420"; break; case 0x20: s = "411"; break; case 0x22: s = "410"; break; }
        if (!s) return "";
        if (alpha && (sub == 0x01 || sub == 0x20 || sub == 0x22)) return "";
        if (c.bits > 8 && (sub == 0x01 || sub == 0x20 || sub == 0x22)) return "";
        std::string n = std::string(alpha ? "yuva" : "yuv") + s + "p";
        if (c.bits > 8) n += std::to_string(c.bits) + "le";
        return n;
    };
    if (c.colorspace == 0) {
        if (!c.transparency && !c.chroma_planes) name = c.bits <= 8 ? "gray" : "gray16le";
        else if (c.transparency && !c.chroma_planes) { if (c.bits <= 8) name = "ya8"; }
        else if (c.bits <= 8 || c.bits == 9 || c.bits == 10 || c.bits == 16) name = yuv(c.transparency);
    } else if (c.colorspace == 1) {
        if (c.chroma_h_shift || c.chroma_v_shift) return false;
        if (c.bits <= 8) name = c.transparency ? "bgra" : "bgr0";
        else if (!c.transparency && (c.bits == 9 || c.bits == 10 || c.bits == 12 || c.bits == 14)) name = "gbrp" + std::to_string(c.bits) + "le";
    }
    if (name.empty()) return false;
    const int bits = c.bits, cp = c.chroma_planes, tr = c.transparency, hs = c.chroma_h_shift, vs = c.chroma_v_shift, cs = c.colorspace;
    if (!parse_pix_fmt(name, c)) return false;
    // parse_pix_fmt re-derives the same fields; keep the header's values authoritative
    c.bits = bits; c.chroma_planes = cp; c.transparency = tr; c.chroma_h_shift = hs; c.chroma_v_shift = vs; c.colorspace = cs;
    return true;
}

int parse_extradata(const uint8_t *d, int n, int width, int height, Config &c, std::string &err)
{
    // ffv1dec.c:521-636
    c = Config();
    c.width = width; c.height = height;
    if (width <= 0 || height <= 0) { err = "invalid picture size"; return FFV1B200_ERR_INVALIDDATA; }
    if (n < 2) { err = "extradata too small"; return FFV1B200_ERR_INVALIDDATA; }
    BinDecoder bd(d, n);
    uint8_t st[kStateSlots];
    memset(st, 128, sizeof(st));
    c.version = bd.get_symbol(st, false);
    if (c.version < 2) { err = "Invalid version in global header"; return FFV1B200_ERR_INVALIDDATA; }
    if (c.version > 2) {
        bd.shrink_end(4);
        c.micro_version = bd.get_symbol(st, false);
        if (c.micro_version < 0) return FFV1B200_ERR_INVALIDDATA;
    }
    c.ac = bd.get_symbol(st, false);
    for (int i = 1; i < 256; i++)
        c.state_transition[i] = (uint8_t)((c.ac == AC_RANGE_CUSTOM ? bd.get_symbol(st, true) : 0) + bd.one_state[i]);
    c.colorspace = bd.get_symbol(st, false);
    c.bits = bd.get_symbol(st, false);
    c.chroma_planes = bd.get(st);
    c.chroma_h_shift = bd.get_symbol(st, false);
    c.chroma_v_shift = bd.get_symbol(st, false);
    c.transparency = bd.get(st);
    c.plane_count = 1 + (c.chroma_planes || c.version < 4) + c.transparency;
    c.num_h_slices = 1 + bd.get_symbol(st, false);
    c.num_v_slices = 1 + bd.get_symbol(st, false);
    if ((unsigned)c.chroma_h_shift > 4U || (unsigned)c.chroma_v_shift > 4U) { err = "chroma shift parameters are invalid"; return FFV1B200_ERR_INVALIDDATA; }
    if (c.num_h_slices > width || c.num_h_slices <= 0 || c.num_v_slices > height || c.num_v_slices <= 0) { err = "slice count invalid"; return FFV1B200_ERR_INVALIDDATA; }
    int qtc = bd.get_symbol(st, false);
    if (qtc <= 0 || qtc > 8) { err = "quant table count is invalid"; return FFV1B200_ERR_INVALIDDATA; }
    if (qtc != 2) { err = "streams with quant_table_count != 2 are not produced by the reference encoder; unsupported"; return FFV1B200_ERR_ENOSYS; }
    for (int t = 0; t < qtc; t++) {
        int count = 1;
        for (int i = 0; i < 5; i++) {
            int r = read_quant_table(bd, c.quant_tables[t][i], count);
            if (r < 0) { err = "read_quant_table error"; return FFV1B200_ERR_INVALIDDATA; }
            count *= r;
            if ((unsigned)count > 32768U) { err = "read_quant_table error"; return FFV1B200_ERR_INVALIDDATA; }
        }
        c.context_count[t] = (count + 1) / 2;
    }
    uint8_t st2[32][kStateSlots];                          // ffv1dec.c:591-600: initial states of a two-pass encode
    memset(st2, 128, sizeof(st2));
    for (int t = 0; t < qtc; t++) {
        c.initial_states[t].clear();
        if (!bd.get(st)) continue;
        std::vector<uint8_t> &init = c.initial_states[t];
        init.assign((size_t)c.context_count[t] * 32, 128);
        for (int j = 0; j < c.context_count[t]; j++)
            for (int k = 0; k < 32; k++) {
                const int pred = j ? init[(size_t)(j - 1) * 32 + k] : 128;
                init[(size_t)j * 32 + k] = (uint8_t)((pred + bd.get_symbol(st2[k], true)) & 0xFF);
            }
    }
    if (c.version > 2) {
        c.ec = bd.get_symbol(st, false);
        if (c.micro_version > 2) c.intra = bd.get_symbol(st, false);
    }
    if (bd.bad) { err = "damaged global header"; return FFV1B200_ERR_INVALIDDATA; }
    if (c.version > 2 && (n < 4 || crc32_ieee(0, d, n) != 0)) { err = "CRC mismatch in global header"; return FFV1B200_ERR_INVALIDDATA; }
    if (c.version != 3 && c.version != 4) { err = "only FFV1 version 3 and 4 extradata is supported"; return FFV1B200_ERR_ENOSYS; }
    if (!c.bits) c.bits = 8;
    if (!select_pix_fmt(c)) { err = "format not supported"; return FFV1B200_ERR_ENOSYS; }
    return 0;
}

int parse_frame_prefix_v01(const uint8_t *pkt, int size, int width, int height, Config &c, bool have_config,
                           PrefixState &ps, std::string &err)
{
    // decode_frame + read_header for version 0/1 streams (ffv1dec.c:920-937, 646-696, 788-795): keyframe bit, then on
    // keyframes the whole parameter set, coded with the default state table by the slice coder itself
    if (size < 2) { err = "packet too small"; return FFV1B200_ERR_INVALIDDATA; }
    BinDecoder bd(pkt, size);
    uint8_t keystate = 128;
    ps.key = bd.get(&keystate) != 0;
    if (ps.key) {
        Config n;
        n.width = width; n.height = height;
        uint8_t st[kStateSlots];
        memset(st, 128, sizeof(st));
        n.version = bd.get_symbol(st, false);
        if (n.version < 0 || n.version >= 2) { err = "Invalid version in frame header"; return FFV1B200_ERR_INVALIDDATA; }
        n.ac = bd.get_symbol(st, false);
        for (int i = 1; i < 256; i++)
            n.state_transition[i] = (uint8_t)((n.ac == AC_RANGE_CUSTOM ? bd.get_symbol(st, true) : 0) + bd.one_state[i]);
        n.colorspace = bd.get_symbol(st, false);
        n.bits = n.version > 0 ? bd.get_symbol(st, false) : 8;
        if (!n.bits) n.bits = 8;
        n.chroma_planes = bd.get(st);
        n.chroma_h_shift = bd.get_symbol(st, false);
        n.chroma_v_shift = bd.get_symbol(st, false);
        n.transparency = bd.get(st);
        n.plane_count = 2 + n.transparency;
        n.num_h_slices = n.num_v_slices = 1;
        if ((unsigned)n.chroma_h_shift > 4U || (unsigned)n.chroma_v_shift > 4U || bd.bad) { err = "damaged frame header"; return FFV1B200_ERR_INVALIDDATA; }
        int count = 1;
        for (int i = 0; i < 5; i++) {
            int r = read_quant_table(bd, n.quant_tables[0][i], count);
            if (r < 0) { err = "read_quant_table error"; return FFV1B200_ERR_INVALIDDATA; }
            count *= r;
            if ((unsigned)count > 32768U) { err = "read_quant_table error"; return FFV1B200_ERR_INVALIDDATA; }
        }
        memcpy(n.quant_tables[1], n.quant_tables[0], sizeof(n.quant_tables[0]));
        n.context_count[0] = n.context_count[1] = (count + 1) / 2;
        n.context_model = 0;
        if (!select_pix_fmt(n)) { err = "format not supported"; return FFV1B200_ERR_ENOSYS; }
        if (have_config) {
            // the device tables are built once: a stream that changes its parameters mid-way is refused
            if (n.version != c.version || n.ac != c.ac || n.pix_fmt != c.pix_fmt || n.context_count[0] != c.context_count[0] ||
                memcmp(n.quant_tables[0], c.quant_tables[0], sizeof(n.quant_tables[0])) || memcmp(n.state_transition, c.state_transition, 256)) {
                err = "stream parameters changed at a keyframe"; return FFV1B200_ERR_ENOSYS;
            }
        } else
            c = n;
    } else if (!have_config) {
        err = "Cannot decode non-keyframe without valid keyframe"; return FFV1B200_ERR_INVALIDDATA;
    }
    ps.low = (uint32_t)bd.low(); ps.range = (uint32_t)bd.range(); ps.pos = (uint32_t)(bd.pos() - pkt);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// slice geometry and per-slice prefix
// ------------------------------------------------------------------------------------------------
void slice_rect(const Config &c, int i, int *x0, int *y0, int *w, int *h)
{
    const int sx = i % c.num_h_slices, sy = i / c.num_h_slices;                      // ffv1.c:124-131
    const int xs = c.width * sx / c.num_h_slices, xe = c.width * (sx + 1) / c.num_h_slices;
    const int ys = c.height * sy / c.num_v_slices, ye = c.height * (sy + 1) / c.num_v_slices;
    *x0 = xs; *y0 = ys; *w = xe - xs; *h = ye - ys;
}

static void put_v01_header(BinCoder &bc, const Config &c)
{
    // ffv1enc.c:498-524: versions 0/1 repeat the global parameters in every keyframe (slice 0's coder,
    // default transition table still active)
    uint8_t st[kStateSlots];
    memset(st, 128, sizeof(st));
    bc.put_symbol(st, c.version, false);
    bc.put_symbol(st, c.ac, false);
    if (c.ac == AC_RANGE_CUSTOM)
        for (int i = 1; i < 256; i++) bc.put_symbol(st, c.state_transition[i] - bc.one_state[i], true);
    bc.put_symbol(st, c.colorspace, false);
    if (c.version > 0) bc.put_symbol(st, c.bits, false);
    bc.put(st, c.chroma_planes);
    bc.put_symbol(st, c.chroma_h_shift, false);
    bc.put_symbol(st, c.chroma_v_shift, false);
    bc.put(st, c.transparency);
    for (int i = 0; i < 5; i++) put_quant_table(bc, c.quant_tables[c.context_model][i]);
}

const int kRctCoef[kRctVariants][2] = {                   // {ry, by}: the candidates of choose_rct_params (ffv1enc.c:1066-1084)
    {0, 0}, {1, 1}, {2, 2}, {0, 2}, {2, 0}, {4, 0}, {0, 4}, {0, 3}, {3, 0}, {3, 1}, {1, 3}, {1, 2}, {2, 1}, {0, 1}, {1, 0},
};

int rct_variants(const Config &c) { return c.version > 3 && c.colorspace == 1 ? kRctVariants : 1; }

static void put_prefix(BinCoder &bc, const Config &c, int si, bool key, int sar_num, int sar_den, int ps, int variant)
{
    if (si == 0) {
        uint8_t keystate = 128;                                                       // ffv1enc.c:1299-1307
        bc.put(&keystate, key ? 1 : 0);
        if (key && c.version < 2) put_v01_header(bc, c);
    }
    coder_state_tables(c, bc.zero_state, bc.one_state);                               // ffv1enc.c:1309-1315
    if (c.version > 2) {                                                              // ffv1enc.c:1031-1051
        int x0, y0, w, h;
        slice_rect(c, si, &x0, &y0, &w, &h);
        uint8_t st[kStateSlots];
        memset(st, 128, sizeof(st));
        bc.put_symbol(st, (x0 + 1) * c.num_h_slices / c.width, false);
        bc.put_symbol(st, (y0 + 1) * c.num_v_slices / c.height, false);
        bc.put_symbol(st, (w + 1) * c.num_h_slices / c.width - 1, false);
        bc.put_symbol(st, (h + 1) * c.num_v_slices / c.height - 1, false);
        for (int j = 0; j < c.plane_count; j++) bc.put_symbol(st, c.context_model, false);
        bc.put_symbol(st, ps, false);
        bc.put_symbol(st, sar_num, false);
        bc.put_symbol(st, sar_den, false);
        if (c.version > 3) {                                                          // ffv1enc.c:1052-1061
            // slice_coding_mode is always 0 here (the PCM fallback needs > 12 bytes per pixel); planar
            // YUV / gray content has no RCT and codes the neutral coefficients 1, 1
            const int by = c.colorspace == 1 ? kRctCoef[variant][1] : 1, ry = c.colorspace == 1 ? kRctCoef[variant][0] : 1;
            bc.put(st, 0);                                                            // "reset contexts" flag, on state[0]
            bc.put_symbol(st, 0, false);
            bc.put_symbol(st, by, false);
            bc.put_symbol(st, ry, false);
        }
    }
}

std::vector<uint16_t> slice_prefix_decisions(const Config &c, int si, bool key, int sar_num, int sar_den, int ps, int variant)
{
    BinCoder bc;
    bc.record_only = true;
    put_prefix(bc, c, si, key, sar_num, sar_den, ps, variant);
    return bc.decisions;
}

std::vector<uint8_t> slice_prefix_bytes(const Config &c, int si, bool key, int sar_num, int sar_den, int ps, int variant)
{
    // Golomb-Rice mode (ffv1enc.c:1176-1183): the range-coded part of a slice is closed before the bit stream starts
    BinCoder bc;
    put_prefix(bc, c, si, key, sar_num, sar_den, ps, variant);
    if (c.version > 2) { uint8_t s129 = 129; bc.put(&s129, 0); }
    if (c.version > 2 || si == 0) { bc.terminate(); return bc.bytes; }
    return std::vector<uint8_t>();
}

// ------------------------------------------------------------------------------------------------
// kernel tables
// ------------------------------------------------------------------------------------------------
static inline int ceil_rshift(int v, int s) { return -((-v) >> s); }

void build_tables(const Config &c, Tables &t)
{
    Layout &L = t.layout;
    memset(&L, 0, sizeof(L));
    L.width = c.width; L.height = c.height;
    L.raw_bits = c.bits;
    L.rgb = c.colorspace == 1;
    L.golomb = c.ac == AC_GOLOMB;
    L.ctx_inputs = c.context_model ? 5 : 3;
    L.ctx_count = c.context_count[c.context_model];
    L.sample_shift = (c.bits > 8 && !c.packed_at_lsb && !L.rgb) ? 16 - c.bits : 0;
    if (L.rgb) {
        L.src_kind = c.bits <= 8 ? SRC_RGB32 : SRC_GBRP16;
        L.coded_bits = (c.bits <= 8 ? 8 : c.bits) + 1;                    // ffv1enc.c:465-468
        L.rct_offset = 1 << (c.bits <= 8 ? 8 : c.bits);
        L.nplanes = 3 + (c.transparency ? 1 : 0);
        for (int p = 0; p < L.nplanes; p++) L.plane[p] = PlaneInfo{p, 0, 0, (p + 1) / 2, c.pixel_bytes[0], 0};   // ffv1enc.c:459-469
    } else {
        L.src_kind = c.bits <= 8 ? SRC_PLANAR8 : SRC_PLANAR16;
        L.coded_bits = c.bits <= 8 ? 8 : c.bits;
        int n = 0;
        if (c.ya8) {                                                      // ffv1enc.c:1199-1201
            L.plane[n++] = PlaneInfo{0, 0, 0, 0, 2, 0};
            L.plane[n++] = PlaneInfo{0, 0, 0, 1, 2, 1};
        } else {                                                          // ffv1enc.c:1185-1198
            L.plane[n++] = PlaneInfo{0, 0, 0, 0, c.bytes_per_sample, 0};
            if (c.chroma_planes) {
                L.plane[n++] = PlaneInfo{1, c.chroma_h_shift, c.chroma_v_shift, 1, c.bytes_per_sample, 0};
                L.plane[n++] = PlaneInfo{2, c.chroma_h_shift, c.chroma_v_shift, 1, c.bytes_per_sample, 0};
            }
            if (c.transparency) L.plane[n++] = PlaneInfo{3, 0, 0, 2, c.bytes_per_sample, 0};
        }
        L.nplanes = n;
    }
    L.npc = 0;
    for (int p = 0; p < L.nplanes; p++) L.npc = std::max(L.npc, L.plane[p].pc + 1);
    L.nslices = c.slice_count();

    t.slices.assign(L.nslices, SliceGeom());
    t.lines.clear(); t.pc_lines.clear(); t.tiles.clear(); t.run_pc.clear(); t.ctiles.clear();
    uint32_t list_cursor = 0;
    uint32_t rec_cursor = 0, scratch_cursor = 0;
    for (int si = 0; si < L.nslices; si++) {
        SliceGeom &g = t.slices[si];
        memset(&g, 0, sizeof(g));
        slice_rect(c, si, &g.x0, &g.y0, &g.w, &g.h);
        for (int p = 0; p < L.nplanes; p++) {
            const PlaneInfo &pi = L.plane[p];
            g.pw[p] = ceil_rshift(g.w, pi.hshift);                        // ffv1enc.c:1186-1189
            g.ph[p] = ceil_rshift(g.h, pi.vshift);
            g.px0[p] = g.x0 >> pi.hshift;
            g.py0[p] = g.y0 >> pi.vshift;
        }
        g.line_first = (int32_t)t.lines.size();
        g.run_first = (int32_t)t.run_pc.size();
        g.nruns = 0;
        g.rec_first = rec_cursor;
        uint32_t off = 0, nsamp = 0;
        auto add_line = [&](int p, int y) {
            LineDesc ld;
            ld.rec_off = off; ld.w = (uint16_t)g.pw[p]; ld.pc = (uint8_t)L.plane[p].pc; ld.plane = (uint8_t)p; ld.y = (uint32_t)y;
            if (g.nruns == 0 || t.run_pc.back() != ld.pc) { t.run_pc.push_back(ld.pc); g.nruns++; }
            ld.run = (uint32_t)(g.nruns - 1);
            t.lines.push_back(ld);
            off += (uint32_t)((g.pw[p] + 31) & ~31);
            nsamp += (uint32_t)g.pw[p];
        };
        if (!L.rgb) {
            for (int p = 0; p < L.nplanes; p++) {
                for (int y0 = 0; y0 < g.ph[p]; y0 += kTileRows) {
                    TileDesc td;
                    td.slice = (uint16_t)si; td.plane = (uint8_t)p; td.nplanes = 1; td.y0 = (uint16_t)y0;
                    td.nrows = (uint16_t)std::min(kTileRows, g.ph[p] - y0);
                    td.line_first = (int32_t)t.lines.size() - g.line_first + y0;
                    td.line_step = 1;
                    t.tiles.push_back(td);
                }
                for (int y = 0; y < g.ph[p]; y++) add_line(p, y);
            }
        } else {
            const int rows = std::max(1, kTileRows / 2);
            for (int y0 = 0; y0 < g.h; y0 += rows) {
                TileDesc td;
                td.slice = (uint16_t)si; td.plane = 0; td.nplanes = (uint8_t)L.nplanes; td.y0 = (uint16_t)y0;
                td.nrows = (uint16_t)std::min(rows, g.h - y0);
                td.line_first = y0 * L.nplanes;
                td.line_step = L.nplanes;
                t.tiles.push_back(td);
            }
            for (int y = 0; y < g.h; y++)
                for (int p = 0; p < L.nplanes; p++) add_line(p, y);
        }
        g.nlines = (int32_t)t.lines.size() - g.line_first;
        g.rec_count = off;
        g.nsamples = nsamp;
        rec_cursor += off;
        // per-plane-context line lists, coding order
        for (int pc = 0; pc < 3; pc++) {
            g.pc_line_first[pc] = (int32_t)t.pc_lines.size();
            uint32_t ns = 0;
            for (int li = 0; li < g.nlines; li++)
                if (t.lines[g.line_first + li].pc == pc) { t.pc_lines.push_back(li); ns += t.lines[g.line_first + li].w; }
            g.pc_nlines[pc] = (int32_t)t.pc_lines.size() - g.pc_line_first[pc];
            g.pc_samples[pc] = ns;
            g.list_off[pc] = list_cursor;
            list_cursor += ns;
        }
        // coder output scratch: raw size of the slice + 12.5 % + slack; grown on demand (overflow is detected)
        const uint64_t raw = (uint64_t)nsamp * (L.coded_bits > 8 ? 2 : 1);
        g.scratch_off = scratch_cursor;
        g.scratch_cap = (uint32_t)(((raw + raw / 8 + 4096) + 255) & ~255ull);
        scratch_cursor += g.scratch_cap;
    }
    // ---- context tiles: the units of the per-context list builders (ffv1_ctx_replay.cu).  16 consecutive lines of a
    // (slice, plane context); 64 for the large context model (a tile carries a histogram of all 7563 contexts); 32 for the
    // tile-sorted lists of 8-bit planar range-coded content (see Layout::tiled_lists)
    {
        bool tiled = L.ctx_count <= 1024 && L.coded_bits <= 10;
        if (const char *v = getenv("FFV1B200_TILED")) { if (atoi(v) == 0) tiled = false; }     // A/B switch: the chain-wide lists
        for (const SliceGeom &g : t.slices)
            for (int pc = 0; pc < 3 && tiled; pc++) {
                uint32_t widest = 0;
                for (int i = 0; i < g.pc_nlines[pc]; i++) {
                    const LineDesc &ld = t.lines[g.line_first + t.pc_lines[g.pc_line_first[pc] + i]];
                    if (ld.run != t.lines[g.line_first + t.pc_lines[g.pc_line_first[pc]]].run) tiled = false;   // one run per plane context
                    widest = std::max<uint32_t>(widest, ld.w);
                }
                if (widest * (kTiledLines / 4) > (uint32_t)kTiledMaxSamples) tiled = false;      // slices up to 1408 samples wide
            }
        L.tiled_lists = tiled ? 1 : 0;
        for (size_t si = 0; si < t.slices.size(); si++) {
            SliceGeom &g = t.slices[si];
            for (int pc = 0; pc < 3; pc++) {
                int tile_lines = L.ctx_count > 1024 ? 4 * kCtxTileLines : (tiled ? kTiledLines : kCtxTileLines);
                if (tiled) {
                    // a tile is the replay's window, and the sort's per-tile set-up is the same whatever the tile holds
                    uint32_t widest = 0;
                    for (int i = 0; i < g.pc_nlines[pc]; i++)
                        widest = std::max<uint32_t>(widest, t.lines[g.line_first + t.pc_lines[g.pc_line_first[pc] + i]].w);
                    // as many lines as the sort's shared-memory image holds: 96 (planes <= 176 wide), 48 (<= 352), 24, 12
                    tile_lines = 2 * kTiledLines;
                    while (tile_lines > kTiledLines / 4 && widest * tile_lines > (uint32_t)kTiledMaxSamples) tile_lines /= 2;
                }
                g.ct_first[pc] = (int32_t)t.ctiles.size();
                uint32_t before = 0;
                for (int l0 = 0; l0 < g.pc_nlines[pc]; l0 += tile_lines) {
                    CtxTile ct;
                    ct.slice = (uint16_t)si; ct.pc = (uint8_t)pc; ct.first = (uint32_t)l0;
                    ct.nlines = (uint8_t)std::min(tile_lines, g.pc_nlines[pc] - l0);
                    ct.sample_first = before;
                    uint32_t ns = 0;
                    for (int i = 0; i < ct.nlines; i++) ns += t.lines[g.line_first + t.pc_lines[g.pc_line_first[pc] + l0 + i]].w;
                    ct.nsamples = ns;
                    before += ns;
                    t.ctiles.push_back(ct);
                }
                g.ct_count[pc] = (int32_t)t.ctiles.size() - g.ct_first[pc];
            }
        }
    }
    L.lines_per_frame = (int32_t)t.lines.size();
    L.tiles_per_frame = (int32_t)t.tiles.size();
    L.rec_per_frame = rec_cursor;
    L.scratch_per_frame = scratch_cursor;
    L.runs_per_frame = (int32_t)t.run_pc.size();
    L.ctiles_per_frame = (int32_t)t.ctiles.size();
    L.samples_per_frame = list_cursor;
    layout_decisions(t, 5.0);
}

void layout_decisions(Tables &t, double entries_per_sample)
{
    // one region per (slice, plane context); every run starts on an 8-entry (16 byte) boundary inside its region
    uint64_t cur = 0;
    for (auto &g : t.slices) {
        int runs_of_pc[3] = {0, 0, 0};
        for (int r = 0; r < g.nruns; r++) runs_of_pc[t.run_pc[g.run_first + r]]++;
        for (int pc = 0; pc < 3; pc++) {
            uint64_t cap = g.pc_samples[pc] ? (uint64_t)((double)g.pc_samples[pc] * entries_per_sample) + 8ull * runs_of_pc[pc] + 64 : 0;
            cap = (cap + 7) & ~7ull;
            g.dec_off[pc] = (uint32_t)cur;
            g.dec_cap[pc] = (uint32_t)cap;
            cur += cap;
        }
    }
    t.layout.dec_per_frame = (uint32_t)cur;
}

} // namespace ffv1
