/*
 * ffv1_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded restatement of the reference's FFV1 encode/decode hot path
 * (theacetoace/FFMPEG-FFV1-P-FRAMES = FFmpeg 3.0.git: libavcodec/ffv1.h, ffv1.c, ffv1enc.c,
 * ffv1dec.c, rangecoder.[ch], golomb.h, put_bits.h, libavutil/crc.c).  It is the checker the
 * CUDA path is compared against; it is never linked into, imported by, or called from the
 * product (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it).
 *
 * Parity of this oracle is PINNED: tests/test_oracle_vs_ref.py compares it packet-by-packet and
 * extradata-by-extradata with oracle/_ref/libffv1ref.so (the unmodified reference sources compiled
 * by oracle/Makefile) and tests/test_oracle_golden.py against committed vectors produced by that
 * reference build (tests/golden/).
 */
#ifndef FFV1_ORACLE_H
#define FFV1_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#define FFV1O_MAX_SLICES   256
#define FFV1O_MAX_CONTEXTS 7563   /* (11*11*5*5*5+1)/2, ffv1enc.c:871 */

/* coder ids after resolution (ffv1.h:56-59) */
#define FFV1O_AC_GOLOMB 0
#define FFV1O_AC_RANGE_DEFAULT 1
#define FFV1O_AC_RANGE_CUSTOM 2

/* pixel layouts understood by the oracle */
enum ffv1o_layout {
    FFV1O_LAYOUT_PLANAR = 0,   /* 1..4 separate planes, 1 or 2 bytes per sample */
    FFV1O_LAYOUT_YA8    = 1,   /* gray+alpha interleaved bytes, pixel stride 2 */
    FFV1O_LAYOUT_RGB32  = 2,   /* packed 32-bit little-endian B,G,R,(A|X) */
    FFV1O_LAYOUT_GBRP   = 3,   /* planar G,B,R 16-bit containers, 9..14 significant bits */
};

typedef struct ffv1o_params {
    int width, height;
    int version, micro_version;
    int ac;                 /* resolved coder: 0 golomb, 1 range default table, 2 range custom table */
    int colorspace;         /* 0 = YCbCr/gray, 1 = RGB via RCT */
    int bits;               /* bits_per_raw_sample (8..16) */
    int chroma_planes, chroma_h_shift, chroma_v_shift, transparency;
    int layout;             /* enum ffv1o_layout */
    int packed_at_lsb;      /* 9..14-bit YUV keep the raw LSB-aligned value */
    int context_model;      /* 0: 3-input 666 contexts, 1: 5-input 7563 contexts */
    int ec;                 /* per-slice CRC */
    int intra;              /* gop_size < 2 */
    int gop_size;
    int num_h_slices, num_v_slices;
    int plane_count;        /* number of quant_table_index entries in a slice header */
    uint8_t state_transition[256];
    int16_t quant_tables[2][5][256];
    int context_count[2];
    /* version 4 (micro version 2, needs strict_experimental): per-slice RCT coefficients, slice coding mode, reset flag.
     * force_pcm is a TEST KNOB: every slice is coded in slice_coding_mode 1 ("PCM", ffv1enc.c:294-304, 1207-1217), which
     * the reference only does when a slice outgrows its buffer -- used to give the decoders PCM streams. */
    int force_pcm;
} ffv1o_params;

/* Option resolution = encode_init (ffv1enc.c:669-1029).  Returns 0, or a negative code:
 * -22 EINVAL, -38 ENOSYS (unsupported format / slice count), -1094995529 INVALIDDATA. */
int ffv1o_resolve(ffv1o_params *p, int width, int height, const char *pix_fmt, int gop_size,
                  int level, int coder, int context, int slices, int slicecrc);

/* same, with AVCodecContext.strict_std_compliance <= FF_COMPLIANCE_EXPERIMENTAL: unlocks level 4 (ffv1enc.c:703-706) */
int ffv1o_resolve_ex(ffv1o_params *p, int width, int height, const char *pix_fmt, int gop_size,
                     int level, int coder, int context, int slices, int slicecrc, int strict_experimental);

int ffv1o_write_extradata(const ffv1o_params *p, uint8_t *dst, int cap);                 /* ffv1enc.c:545-619 */
int ffv1o_parse_extradata(ffv1o_params *p, int width, int height, const uint8_t *d, int n); /* ffv1dec.c:521-636 */

typedef struct ffv1o_encoder ffv1o_encoder;
ffv1o_encoder *ffv1o_encoder_new(const ffv1o_params *p);
/* picture_structure: 3 progressive, 1 TFF, 2 BFF (ffv1enc.c:1044-1047). Returns packet size or <0. */
long ffv1o_encode_frame(ffv1o_encoder *e, const uint8_t *const planes[4], const int strides[4],
                        int sar_num, int sar_den, int picture_structure,
                        uint8_t *dst, long cap, int *key_frame);
/* total binary range-coder decisions (or golomb symbols) coded so far -- used for reporting */
uint64_t ffv1o_encoder_decisions(const ffv1o_encoder *e);
/* test knob (see ffv1o_params.force_pcm), switchable between frames */
void ffv1o_encoder_set_force_pcm(ffv1o_encoder *e, int on);
void ffv1o_encoder_free(ffv1o_encoder *e);

typedef struct ffv1o_decoder ffv1o_decoder;
/* For version>=2 streams pass the extradata; for version 0/1 pass NULL/0 plus the pix layout hints
 * the container would supply (bits8 only). */
ffv1o_decoder *ffv1o_decoder_new(int width, int height, const uint8_t *extradata, int extradata_size);
const ffv1o_params *ffv1o_decoder_params(const ffv1o_decoder *d);
/* Decodes into caller planes (same layouts as the encoder). Returns 0 ok, <0 error.
 * damaged_mask (may be NULL) receives one bit per slice whose CRC or end check failed (first 64). */
int ffv1o_decode_frame(ffv1o_decoder *d, const uint8_t *pkt, long size,
                       uint8_t *const planes[4], const int strides[4], int *key_frame,
                       uint64_t *damaged_mask);
void ffv1o_decoder_free(ffv1o_decoder *d);

/* ---- intermediate products, for kernel-level parity tests ---- */

/* Per-sample (context, folded diff) records of one slice in coding order (what encode_line feeds the
 * entropy coder, ffv1enc.c:311-321).  rec[i] = (context << 16) | (diff & 0xFFFF).  Returns count. */
long ffv1o_slice_records(const ffv1o_params *p, const uint8_t *const planes[4], const int strides[4],
                         int slice_index, uint32_t *rec, long cap);

/* libavutil AV_CRC_32_IEEE over a buffer (crc.c:357-380): MSB-first poly 0x04C11DB7, init 0. */
uint32_t ffv1o_crc32(uint32_t crc, const uint8_t *buf, size_t len);

/* default range-coder state table (rangecoder.c:63-101 with factor 0.05*2^32, max_p 248) */
void ffv1o_default_state_tables(uint8_t zero_state[256], uint8_t one_state[256]);

#endif
