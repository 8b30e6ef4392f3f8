// ffv1_dec_kernels.cuh -- launch interfaces of the decoder kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ffv1 {

constexpr int kDecSmemRingBytes = 8;     // shared-memory bytes per ring sample: three int16 rows, or one 8-byte line record
constexpr int kDecRingPad = 8;          // int16 elements kept left of x = 0 in a ring row (positions -1, -2 are used)

struct DecDeviceTables {
    int32_t width, height, version, micro_version, ac, colorspace, bits, coded_bits;
    int32_t chroma_planes, hshift, vshift, transparency, packed_at_lsb, ya8, ec, rgb32;
    int32_t num_h_slices, num_v_slices, max_slices, plane_count;
    int32_t ctx_count[2];
    int32_t plane_off[4];               // byte offset of every output plane inside a tightly packed frame
    int32_t plane_pitch[4];             // bytes per row
    int64_t frame_bytes;
    const int16_t *quant;               // [2][5][256]
    const uint8_t *lut;                 // [512] zero_state | one_state the slice coders run with
    const uint8_t *model_init[2];       // per table set: [ctx_count][32] states a keyframe starts from (two-pass streams), or null = 128
    int64_t state_stride;               // bytes of one (slice, plane context) model
    int32_t ring_w;                     // int16 elements per ring row
    int32_t smem_model;                 // bytes of shared memory per chain for the current plane context's model (0 = keep it in global memory)
    int32_t smem_ring_w;                // int16 elements per shared-memory ring row (0 = ring in global memory)
};

struct DecBatch {
    int32_t nframes, nseg;
    const int32_t *seg_first;           // [nseg+1] first frame of every segment (a segment = frames sharing model state)
    const int32_t *seg_set;             // [nseg]   model-state set used by the segment
    const uint8_t *frame_key;           // [nframes]
    const uint8_t *pkt;                 // packets, back to back (each starts 16-byte aligned)
    const uint64_t *pkt_off;            // [nframes]
    const uint32_t *slice_start;        // [nframes][max_slices] offset inside the packet
    const uint32_t *slice_size;         // [nframes][max_slices] payload + trailer bytes
    const int32_t *slice_count;         // [nframes]
    uint8_t *out;                       // [nframes][frame_bytes]
    const uint8_t *prev_frame;          // last frame of the previous batch (concealment source for frame 0) or null
    uint8_t *state;                     // [sets][max_slices][3][state_stride]
    int16_t *ring;                      // [nseg*max_slices][4][3][ring_w]
    const uint32_t *init_state;         // version 0/1: [nframes][3] range-coder low, range, bytes consumed behind the in-band header
    uint32_t *damaged;                  // [nframes][max_slices]: bit0 CRC mismatch, bit1 header/end-of-slice check failed
    int32_t zero_fill;                  // `out` has not been cleared (pictures written straight into the caller's pinned buffer):
                                        // the kernel clears the rectangle of every slice it does not decode as announced by the grid
};

void launch_dec_crc(const DecDeviceTables &t, const DecBatch &b, cudaStream_t s);
void launch_decode(const DecDeviceTables &t, const DecBatch &b, cudaStream_t s);
void launch_conceal(const DecDeviceTables &t, const DecBatch &b, int frame, cudaStream_t s);

} // namespace ffv1
