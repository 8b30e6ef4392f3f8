// ffv1_enc_kernels.cuh -- launch interfaces of the encoder kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>
#include "ffv1_model.h"

namespace ffv1 {

constexpr int kPixelThreads   = 256;
constexpr int kPixelChunk     = 512;          // samples of one row handled per smem pass
constexpr int kPixelPadL      = 8;            // int16 elements left of x=0 in a staged row (keeps 16 B alignment)
constexpr int kPixelRowElems  = kPixelChunk + 16;
constexpr int kReplayWarps    = 2;            // chains (warps) per CTA of the state-replay kernel
 constexpr int kMaxGolombPrefix = 4096;        // bytes (a version 0/1 keyframe carries the whole header)
constexpr int kScratchLead    = 4;            // bytes of a slice's scratch region before its payload (see struct Rac)
constexpr int kMaxPrefix      = 8192;         // decisions before the first sample (v0/v1 keyframes carry a whole header)

// per-batch device buffers and scalars handed to the kernels
struct EncBatch {
    int32_t nframes;
    int32_t nseg;                       // GOP segments inside the batch
    const uint8_t *const *planes;       // [nframes*4] device pointers
    int32_t linesize[4];
    uint32_t *rec;                      // [nframes][rec_per_frame]
    uint32_t *run_cnt;                  // [nframes][runs_per_frame]   decisions of every run (written by k_replay)
    uint16_t *dec;                      // [nframes][dec_per_frame]    decision entries p | bit<<8, one region per
                                        //                             (slice, plane context), runs 8-entry aligned
    const int32_t *seg_first;           // [nseg+1] first frame of each segment
    const uint8_t *frame_key;           // [nframes]
    uint8_t *scratch;                   // [nframes][scratch_per_frame]
    uint32_t *slice_bytes;              // [nframes][nslices] coder output bytes (payload)
    uint32_t *pkt_size;                 // [nframes]
    uint64_t *pkt_off;                  // [nframes+1]
    uint8_t *out;                       // packets, back to back
    uint64_t out_capacity;
    // per-context replay (ffv1_ctx_replay.cu)
    const int32_t *frame_seg;           // [nframes] GOP segment of every frame
    uint32_t *line_pos;                 // [nframes][lines_per_frame] decisions per line, then first decision of the line
    uint32_t *ctx_hist;                 // [nframes][ctiles_per_frame][ctx_count]
    uint32_t *list_start;               // [chains][ctx_count] first entry of a context's list inside the chain's area
    uint32_t *list_count;               // [chains][ctx_count]
    uint16_t *list_order;               // [chains][ctx_count] contexts, longest list first
    uint2 *lists;                       // [nframes * samples_per_frame] {decision position, residual | frame << 16}
    // tile-sorted lists (Layout::tiled_lists; the buffers alias ctx_hist / line_pos)
    uint16_t *tile_tab;                 // [nframes][ctiles_per_frame][tile_tab_pitch] start of every context's run inside the tile's block
    uint32_t *tile_nd;                  // [nframes][ctiles_per_frame] decisions of the tile, then its first decision in the region
    int32_t tile_tab_pitch;             // entries per table: ctx_count + 1 rounded up to 8
    const uint8_t *rct_idx;             // version-4 RGB: [nframes][nslices] index of the slice's RCT coefficient pair (else null)
    // adaptive state
    uint8_t *state_seg;                 // global-state mode: [nseg][nslices][npc][ctx_count*32]
    const uint8_t *carry_in;            // [nslices][npc][ctx_count*32]
    uint8_t *carry_out;
    // [0] decision-region overflow: needed entries per sample * 256, [1] scratch overflow (bytes needed),
    // [2] out overflow (bytes needed), [3] total binary decisions of the batch
    unsigned long long *status;
};

struct EncDeviceTables {
    Layout layout;                      // by value (kernel parameter)
    const SliceGeom *slices;
    const LineDesc *lines;
    const int32_t *pc_lines;
    const TileDesc *tiles;
    const CtxTile *ctiles;
    const int16_t *quant;               // [5][256]
    const uint8_t *trans_lut;           // [512]: zero_state, one_state of the slice coders
    const uint8_t *init_state;          // [ctx_count][32] states a keyframe starts from (two-pass encode), or null = all 128
    const uint8_t *one_pow;             // [33][256]: one_state applied k times (runs of zero residuals in one context)
    const uint8_t *run_pc;              // [runs_per_frame] plane context of every run
    const uint8_t *gprefix;             // Golomb-Rice mode: [nslices][2][kMaxGolombPrefix] bytes every slice starts with
    const int32_t *gprefix_len;         // [nslices][2]
    const uint16_t *prefix;             // [nslices][2][kMaxPrefix]
    const int32_t *prefix_len;          // [nslices][2]
    int32_t ec;
    int32_t version;
    int32_t state_in_smem;
    int32_t nvar;                       // slice-header variants per (slice, keyframe flag): 15 for version-4 RGB, else 1
};

// One CTA per (GOP segment, slice, plane context) chain: CTAs are handed out in blockIdx order, so the chains of the plane
// context with the most samples (luma: twice the chroma chains' work for 4:2:0; B+R for planar RGB) come first and what
// runs while the grid drains is short chains.  Returns the chain index (seg * nslices + slice) * npc + pc of this CTA.
#ifdef __CUDACC__
__device__ __forceinline__ int chain_of_block(const Layout &L, const SliceGeom *slices)
{
    const int per = (int)gridDim.x / L.npc, grp = (int)blockIdx.x / per, rest = (int)blockIdx.x - grp * per;
    const SliceGeom &g0 = slices[0];
    uint32_t n0 = g0.pc_samples[0], n1 = L.npc > 1 ? g0.pc_samples[1] : 0u, n2 = L.npc > 2 ? g0.pc_samples[2] : 0u;
    int p0 = 0, p1 = 1, p2 = 2;
    if (n1 > n0) { const int t = p0; p0 = p1; p1 = t; const uint32_t u = n0; n0 = n1; n1 = u; }
    if (n2 > n1) {
        { const int t = p1; p1 = p2; p2 = t; const uint32_t u = n1; n1 = n2; n2 = u; }
        if (n1 > n0) { const int t = p0; p0 = p1; p1 = t; }
    }
    return rest * L.npc + (grp == 0 ? p0 : (grp == 1 ? p1 : p2));
}
#endif

int  pixel_smem_bytes(const Layout &L);
void launch_pixel(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s);
// version 4, RGB: per-slice choice among the 15 RCT coefficient pairs (choose_rct_params, ffv1enc.c:1064-1144) -> b.rct_idx
void launch_rct_search(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s);
int  replay_smem_bytes(const Layout &L);
void launch_replay(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s);
void launch_rangecode(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s);
void launch_golomb(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s);
// Golomb-Rice mode decomposed by context (small context model): lists (ffv1_ctx_replay.cu), then VlcState replay + bit writer
bool golomb_lists_supported(const Layout &L, int max_tile_samples);
bool ctx_lists_configurable(const Layout &L);
void launch_golomb_lists(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s);
void launch_golomb_coder(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s);
void launch_pack(const EncDeviceTables &t, const EncBatch &b, cudaStream_t s);
cudaError_t configure_kernels(const Layout &L);
// first-pass statistics of a coded batch (range-coder modes): rc_stat[256][2], rc_stat2[ctx_count][32][2], accumulated
cudaError_t launch_pass1_stats(const EncDeviceTables &t, const EncBatch &b, unsigned long long *rc_stat, unsigned long long *rc_stat2,
                               cudaStream_t s);
// context-decomposed state replay (ffv1_ctx_replay.cu)
bool ctx_replay_supported(const Layout &L);
bool ctx_replay_needs_global_state(const Layout &L);      // large context model: B.state_seg holds one model per chain
cudaError_t configure_ctx_replay(const Layout &L);
void launch_ctx_replay(const EncDeviceTables &t, const EncBatch &b, int max_tile_samples, uint32_t max_dec_cap, cudaStream_t s);
// tuned per-pixel pass for planar sources (ffv1_pixel_fast.cu)
// one work item of a frame: up to 32 rows x <= 512 bytes of one slice-plane (everything the kernel needs, precomputed)
struct alignas(16) FastItemDesc {
    int32_t  c0;          // tensor-map x coordinate (32-bit elements) staged at buffer column 0 (may be negative)
    int32_t  c1;          // first source row that is fetched
    uint32_t rec_off;     // first record of the item inside a frame's record area
    uint32_t magic;       // ceil(2^20 / upr): unit index -> row without a division
    uint16_t rec_stride;  // records between consecutive rows
    uint16_t upr;         // 8-sample units per row
    uint16_t nunits;      // upr * rows
    uint16_t nunits_mod;  // nunits modulo the consumer threads of a group
    uint16_t o0;          // byte offset of the item's first sample inside a staged row
    uint16_t rowb;        // staged row pitch (bytes)
    uint8_t  src_plane;
    uint8_t  flags;       // 1 = first tile of the slice-plane (zero rows above), 2 = leftmost chunk, 4 = rightmost chunk
    uint8_t  nrows;
    uint8_t  pad;
};
static_assert(sizeof(FastItemDesc) == 32, "two 16-byte loads per item");
struct FastPlan {
    std::vector<FastItemDesc> items;
    int32_t items_per_frame = 0, row_bytes[4] = {0, 0, 0, 0}, buf_bytes = 0, nbuf = 0, smem_bytes = 0;
};
bool pixel_fast_geometry_ok(const Layout &L, const SliceGeom *slices, int nslices);
void build_pixel_fast_plan(const Tables &tab, FastPlan &plan);
cudaError_t configure_pixel_fast(const FastPlan &plan);
void launch_pixel_fast(const EncDeviceTables &t, const EncBatch &b, const FastPlan &plan, const FastItemDesc *d_items, int num_sms,
                       cudaStream_t s, const uint8_t *const *frame0_planes, long long frame_stride);

} // namespace ffv1
