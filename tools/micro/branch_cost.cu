// Micro-benchmark: what a lone warp (one active lane, like the decoder's serial lane) pays for a taken forward
// branch around a rare block, compared with dependent ALU instructions and predicated instructions.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o branch_cost branch_cost.cu ; run: ./branch_cost
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(unsigned *out, long long *cyc, unsigned seed, unsigned never, int iters)
{
    if (threadIdx.x != 0) return;
    unsigned x = seed, y = seed * 3u, acc = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        // 8 dependent ALU operations
        x = x * 5u + 1u; x ^= x >> 7; x += y; x ^= x << 3; x = x * 3u + 7u; x ^= x >> 5; x += 0x9E37u; x ^= x << 2;
        if (MODE == 1) {
            // rare block behind a branch (never taken at run time, but the compiler cannot know)
            if ((x & never) == 0x12345u) {
                y = y * 7u + x; acc ^= y >> 3; acc += out[x & 15u]; y ^= acc << 1; acc += 3u; out[(y & 15u) + 16] = acc;
            }
        } else if (MODE == 2) {
            // the same test, result consumed by 5 predicated (select) operations
            const bool p = (x & never) == 0x12345u;
            y = p ? y * 7u + x : y; acc = p ? acc ^ (y >> 3) : acc; acc = p ? acc + 3u : acc; y = p ? y ^ (acc << 1) : y; acc = p ? acc + y : acc;
        } else if (MODE == 3) {
            // two rare blocks (renormalisation + nested refill shape)
            if ((x & never) == 0x12345u) {
                y = y * 7u + x; acc ^= y >> 3;
                if ((y & never) == 0x54321u) { acc += out[x & 15u]; y ^= acc << 1; out[(y & 15u) + 16] = acc; }
            }
            if ((x & never) == 0x22222u) { acc += 1; }
        }
    }
    const long long t1 = clock64();
    out[0] = x + y + acc;
    cyc[MODE] = t1 - t0;
}

int main()
{
    unsigned *out; long long *cyc;
    cudaMalloc(&out, 4096); cudaMemset(out, 0, 4096);
    cudaMallocManaged(&cyc, 64);
    const int iters = 1 << 20;
    for (int rep = 0; rep < 2; rep++) {
        k<0><<<1, 32>>>(out, cyc, 12345u, 0xFFFFFFFFu, iters);
        k<1><<<1, 32>>>(out, cyc, 12345u, 0xFFFFFFFFu, iters);
        k<2><<<1, 32>>>(out, cyc, 12345u, 0xFFFFFFFFu, iters);
        k<3><<<1, 32>>>(out, cyc, 12345u, 0xFFFFFFFFu, iters);
        cudaDeviceSynchronize();
    }
    printf("cycles per iteration: base (8 dependent ALU + loop) %.1f | + rare block behind a branch %.1f | + 5 predicated ops %.1f | + nested rare blocks %.1f\n",
           (double)cyc[0] / iters, (double)cyc[1] / iters, (double)cyc[2] / iters, (double)cyc[3] / iters);
    return 0;
}
