// ffv1_model.h -- host-side model of the FFV1 format knobs that are not kernels:
// option resolution, quantisation / state-transition tables, slice grid and per-slice coding-order
// tables, the extradata record and the per-slice header decision lists.
// Mirrors (behaviour, not code) libavcodec/ffv1enc.c:669-1029 (encode_init), 545-619 (write_extradata),
// 1031-1062 (encode_slice_header), libavcodec/ffv1.c:117-160 and libavcodec/ffv1dec.c:476-636.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ffv1 {

enum Coder { AC_GOLOMB = 0, AC_RANGE_DEFAULT = 1, AC_RANGE_CUSTOM = 2 };

// how K_pixel fetches raw samples
enum SrcKind { SRC_PLANAR8 = 0, SRC_PLANAR16 = 1, SRC_RGB32 = 2, SRC_GBRP16 = 3 };

constexpr int kMaxSlices = 256;
constexpr int kStateSlots = 32;      // CONTEXT_SIZE, ffv1.h:50

// ---- structs shared with the device (plain data) -----------------------------------------------

// one coded sample plane of a slice ("Y", "U", "V", "A" or post-RCT "G'", "B'", "R'", "A")
struct PlaneInfo {
    int32_t src_plane;   // AVFrame.data[] index the samples come from (RGB kinds: unused)
    int32_t hshift, vshift;
    int32_t pc;          // plane-context index (ffv1.h PlaneContext): Y/G=0, chroma & B/R=1 (shared), alpha=2 (YA8: 1)
    int32_t pstep;       // bytes between horizontally adjacent samples in the source
    int32_t poff;        // byte offset of the sample inside a pixel (YA8 alpha = 1)
};

struct SliceGeom {
    int32_t x0, y0, w, h;            // luma geometry (ffv1.c:124-143)
    int32_t pw[4], ph[4];            // size of each coded sample plane
    int32_t px0[4], py0[4];          // origin of each coded sample plane inside its source plane (samples)
    int32_t line_first;              // index of this slice's first line in the per-frame line table
    int32_t nlines;
    uint32_t rec_first;              // first record of this slice inside a frame's record area
    uint32_t rec_count;              // records incl. line padding
    uint32_t nsamples;               // real samples
    uint32_t scratch_off;            // byte offset of this slice's coder output inside a frame's scratch area
    uint32_t scratch_cap;
    int32_t pc_line_first[3];        // per plane-context: start in pc_lines[] (indices into the line table)
    int32_t pc_nlines[3];
    uint32_t pc_samples[3];
    int32_t run_first;               // first entry of this slice in the per-frame run table
    int32_t nruns;                   // runs = maximal groups of consecutive lines (coding order) of one plane context
    uint32_t dec_off[3];             // per plane-context: first entry of its decision region inside a frame's dec area
    uint32_t dec_cap[3];             // capacity of that region in entries (multiple of 8)
    int32_t ct_first[3];             // per plane-context: first context tile (see CtxTile) in the per-frame table
    int32_t ct_count[3];
    uint32_t list_off[3];            // samples of all (slice, plane context) pairs that precede this one in a frame
};

// one line of samples in coding order
struct LineDesc {
    uint32_t rec_off;    // record offset inside the slice's record region (multiple of 32)
    uint16_t w;
    uint8_t  pc;
    uint8_t  plane;
    uint32_t y;
    uint32_t run;        // slice-relative index of the run this line belongs to
};

// unit of work of the per-pixel kernel: up to kTileRows rows of one plane (all planes for RGB) of one slice
struct TileDesc {
    uint16_t slice;
    uint8_t  plane;      // first coded plane handled
    uint8_t  nplanes;    // 1 (YUV) or 3/4 (RGB: planes are produced together by the RCT)
    uint16_t y0;
    uint16_t nrows;
    int32_t  line_first; // line-table index (slice relative) of (y0, plane)
    int32_t  line_step;  // distance between consecutive rows of the same plane in the line table
};

// unit of work of the per-context list builder: up to kCtxTileLines consecutive lines of one (slice, plane context)
struct CtxTile {
    uint16_t slice;
    uint8_t  pc;
    uint8_t  nlines;
    uint32_t first;        // index of its first line in the (slice, plane context) line list
    uint32_t sample_first; // samples of the (slice, plane context) in front of the tile
    uint32_t nsamples;
};
constexpr int kCtxTileLines = 16;
constexpr int kTiledLines = 48;            // tile-sorted lists: lines per tile, and the largest tile the sort holds
constexpr int kTiledMaxSamples = 16896;    // (48 lines of <= 352 samples)

struct Layout {
    int32_t width, height;
    int32_t src_kind;        // SrcKind
    int32_t raw_bits;        // bits_per_raw_sample
    int32_t coded_bits;      // bits the residual is folded to (raw_bits, +1 for RGB; 8-bit RGB = 9)
    int32_t sample_shift;    // right shift applied to 16-bit containers (16 - bits when not packed at lsb)
    int32_t nplanes;         // coded sample planes per slice
    int32_t ctx_inputs;      // 3 or 5 (context model)
    int32_t ctx_count;       // 666 / 7563
    int32_t npc;             // plane contexts in use (1..3)
    int32_t nslices;
    int32_t lines_per_frame;
    int32_t tiles_per_frame;
    int32_t rgb;             // colorspace == 1
    int32_t golomb;
    uint32_t rec_per_frame;  // records per frame incl. padding
    uint32_t scratch_per_frame;
    int32_t runs_per_frame;
    int32_t ctiles_per_frame;
    uint32_t samples_per_frame;  // coded samples (no padding)
    uint32_t dec_per_frame;  // decision entries per frame (all regions; set by layout_decisions)
    int32_t tiled_lists;     // the per-context lists are kept tile by tile (k_tile_sort / k_replay_grp<TILED>): 8..10-bit content,
                             // small context model, one run per plane context, slices <= 1408 samples wide
    int32_t rct_offset;      // 1 << bits for RGB
    PlaneInfo plane[4];
};

// ---- resolved configuration ---------------------------------------------------------------------
struct Config {
    int width = 0, height = 0;
    std::string pix_fmt;
    int version = 0, micro_version = 0;
    int ac = 0;
    int colorspace = 0, bits = 8;
    int chroma_planes = 0, chroma_h_shift = 0, chroma_v_shift = 0, transparency = 0;
    int packed_at_lsb = 0;
    int ya8 = 0;
    int context_model = 0;
    int ec = 0, intra = 0, gop_size = 0;
    int num_h_slices = 1, num_v_slices = 1;
    int plane_count = 2;                       // quant_table_index entries per slice header
    uint8_t state_transition[256] = {0};
    int16_t quant_tables[2][5][256];
    int context_count[2] = {0, 0};
    int bytes_per_sample = 1;
    // two-pass coding (ffv1enc.c:906-986): per quantisation-table set, [context][32] state bytes a keyframe starts from;
    // empty = every state 128.  Written into the extradata (ffv1enc.c:591-606), read back by ffv1dec.c:591-606.
    std::vector<uint8_t> initial_states[2];
    int pass_flags = 0;                        // kPass1 / kPass2 (AV_CODEC_FLAG_PASS1 / _PASS2)
    int nb_src_planes = 0;                     // AVFrame planes of the pix_fmt
    int src_hshift[4] = {0,0,0,0}, src_vshift[4] = {0,0,0,0};
    int pixel_bytes[4] = {1,1,1,1};            // bytes per pixel in each source plane

    int slice_count() const { return num_h_slices * num_v_slices; }
    int64_t frame_bytes() const;               // tightly packed (av_image_get_buffer_size(align=1))
    void plane_dims(int i, int *rows, int *row_bytes) const;
};

constexpr int kPass1 = 1, kPass2 = 2;
struct EncOptions {
    int width, height; std::string pix_fmt; int gop_size, level, slices, coder, context, slicecrc;
    int pass_flags = 0;          // kPass1: collect statistics; kPass2 (or any stats_in): code with the tables derived from them
    std::string stats_in;        // AVCodecContext.stats_in: the text a first pass left in stats_out
    int strict_experimental = 0; // AVCodecContext.strict_std_compliance <= FF_COMPLIANCE_EXPERIMENTAL
    int bits_per_raw_sample = 0; // AVCodecContext.bits_per_raw_sample: 0 = the format's depth; else the depth coded for 16-bit
                                 // containers (ffv1enc.c:728-748, 796-805)
};

// encode_init: returns 0 or a negative AVERROR-style code with a message in err
int resolve_encoder(const EncOptions &o, Config &c, std::string &err);
// read_extra_header + pix_fmt selection of read_header (version >= 2 streams)
int parse_extradata(const uint8_t *data, int size, int width, int height, Config &c, std::string &err);

// Version 0/1 streams carry their parameters in every keyframe, coded by slice 0's own range coder.  The host reads that
// prefix (keyframe bit [+ header]) and hands the coder state behind it to the decode kernel.
struct PrefixState { bool key; uint32_t low, range, pos; };
int parse_frame_prefix_v01(const uint8_t *pkt, int size, int width, int height, Config &c, bool have_config,
                           PrefixState &ps, std::string &err);

std::vector<uint8_t> write_extradata(const Config &c);

// Two-pass statistics (ffv1enc.c:193-200, 1236-1276): rc_stat[state][bit] over all sample decisions,
// rc_stat2[table set][context][slot][bit]; gob_count = keyframes coded.  format_stats writes AVCodecContext.stats_out's text.
struct PassStats {
    uint64_t rc_stat[256][2];
    std::vector<uint64_t> rc_stat2[2];         // [context_count[i] * 32 * 2]
    int gob_count = 0;
};
std::string format_stats(const Config &c, const PassStats &st);
// stats_in -> sorted state-transition table (custom-table coder) and initial states; 0 or a negative error
int apply_stats(Config &c, const std::string &stats_in, std::string &err);

void default_state_tables(uint8_t zero_state[256], uint8_t one_state[256]);
// transition tables a slice coder runs with: default table, custom entries 1..255 overriding for AC_RANGE_CUSTOM
void coder_state_tables(const Config &c, uint8_t zero_state[256], uint8_t one_state[256]);

// (probability state, bit) pairs, packed p | bit<<8, of everything a slice codes BEFORE its first sample:
// [keyframe bit on slice 0] + slice header symbols (version 3) [+ the state-129 bit in golomb mode].
// Version 4: the header of an RGB slice ends with the RCT coefficients chosen for it (ffv1enc.c:1052-1061), one of
// kRctVariants pairs; `variant` selects the pair (ignored otherwise).  kRctCoef[v] = {ry, by}.
constexpr int kRctVariants = 15;
extern const int kRctCoef[kRctVariants][2];
int rct_variants(const Config &c);           // header variants per (slice, keyframe flag): 15 for version-4 RGB, else 1
std::vector<uint16_t> slice_prefix_decisions(const Config &c, int slice_index, bool key_frame,
                                             int sar_num, int sar_den, int picture_structure, int variant = 0);

// Golomb-Rice mode: the BYTES a slice starts with (range-coded key bit / header, the state-129 bit of version 3,
// ff_rac_terminate); the MSB-first bit stream of the samples follows them (ffv1enc.c:1176-1183).
std::vector<uint8_t> slice_prefix_bytes(const Config &c, int slice_index, bool key_frame,
                                        int sar_num, int sar_den, int picture_structure, int variant = 0);

void slice_rect(const Config &c, int i, int *x0, int *y0, int *w, int *h);

// Builds every geometry table the kernels use.
struct Tables {
    Layout layout;
    std::vector<SliceGeom> slices;
    std::vector<LineDesc>  lines;       // all slices, coding order inside a slice
    std::vector<int32_t>   pc_lines;    // per slice, per plane context: slice-relative line indices
    std::vector<TileDesc>  tiles;
    std::vector<uint8_t>   run_pc;      // per frame: plane context of every run, slices back to back
    std::vector<CtxTile>   ctiles;
};
constexpr int kTileRows = 32;
void build_tables(const Config &c, Tables &t);
// Sizes the per-(slice, plane context) decision regions for `entries_per_sample` binary decisions per coded sample.
void layout_decisions(Tables &t, double entries_per_sample);

uint32_t crc32_ieee(uint32_t crc, const uint8_t *buf, size_t len);

} // namespace ffv1
