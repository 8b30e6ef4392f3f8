// ffv1_internal.h -- small host-side helpers shared by the encoder and decoder translation units.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <memory>
#include <cstddef>

namespace ffv1 {

void set_last_error(const std::string &msg);

// RAII device buffer of n elements
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t alloc(size_t count)
    {
        release();
        if (!count) count = 1;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count; else p = nullptr;
        return e;
    }
    cudaError_t upload(const T *src, size_t count, cudaStream_t s)
    {
        if (n < count || !p) { cudaError_t e = alloc(count); if (e != cudaSuccess) return e; }
        if (!count) return cudaSuccess;
        // source may be pageable: the copy is staged by the runtime before the call returns
        return cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
};

// RAII pinned host buffer
template <typename T>
struct PinnedBuf {
    T *p = nullptr;
    size_t n = 0;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    cudaError_t alloc(size_t count)
    {
        if (p) cudaFreeHost(p);
        p = nullptr; n = 0;
        if (!count) count = 1;
        cudaError_t e = cudaMallocHost((void **)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count; else p = nullptr;
        return e;
    }
};

} // namespace ffv1
