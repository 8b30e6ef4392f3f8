// ffv1_api.cu -- library-level entry points: version, error text, device probing.
#include "../../include/ffv1_b200.h"
#include "ffv1_internal.h"
#include <cuda_runtime.h>
#include <string>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sched.h>
#include <unistd.h>
#include <sys/syscall.h>
#include <sys/mman.h>
#include <map>
#include <mutex>

namespace ffv1 {
static thread_local std::string g_last_error;
void set_last_error(const std::string &msg) { g_last_error = msg; }
__global__ void k_probe(int *out) { if (threadIdx.x == 0) *out = 100; }
}

extern "C" {

const char *ffv1b200_version(void) { return "ffv1_b200 0.1 (sm_100a)"; }

const char *ffv1b200_last_error(void) { return ffv1::g_last_error.c_str(); }

const char *ffv1b200_strerror(int err)
{
    switch (err) {
    case 0: return "success";
    case FFV1B200_ERR_EINVAL: return "Invalid argument";
    case FFV1B200_ERR_ENOMEM: return "Cannot allocate memory";
    case FFV1B200_ERR_ENOSYS: return "Function not implemented";
    case FFV1B200_ERR_INVALIDDATA: return "Invalid data found when processing input";
    case FFV1B200_ERR_EXTERNAL: return "Generic error in an external library (CUDA)";
    case FFV1B200_ERR_BUFFER_TOO_SMALL: return "Buffer too small";
    default: return "unknown error";
    }
}

int ffv1b200_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        ffv1::set_last_error(std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                             " (this library has no CPU fallback)");
        cudaGetLastError();
        return FFV1B200_ERR_EXTERNAL;
    }
    // the kernel image is sm_100a only: make sure it can actually be loaded on device 0
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, ffv1::k_probe);
    if (e != cudaSuccess) {
        ffv1::set_last_error(std::string("sm_100a kernel image not loadable on this device: ") + cudaGetErrorString(e));
        cudaGetLastError();
        return FFV1B200_ERR_EXTERNAL;
    }
    return n;
}

static int numa_node_of_device(int device, char *bdf_out)
{
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof(bdf), device) != cudaSuccess) { cudaGetLastError(); return -2; }
    for (char *q = bdf; *q; q++) if (*q >= 'A' && *q <= 'F') *q += 'a' - 'A';
    if (bdf_out) memcpy(bdf_out, bdf, sizeof(bdf));
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bdf);
    int node = -1;
    if (FILE *fh = fopen(path, "r")) { if (fscanf(fh, "%d", &node) != 1) node = -1; fclose(fh); }
    return node;
}

// Pinned host memory for frame / packet staging, placed on the device's NUMA node when the topology is visible (the
// pages are bound before they are first touched, then registered with CUDA), without touching the caller's affinity.
static std::mutex g_host_mu;
static std::map<void *, size_t> g_host_allocs;

void *ffv1b200_host_alloc(size_t bytes, int device)
{
    if (!bytes) bytes = 1;
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    bytes = (bytes + page - 1) / page * page;
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) { ffv1::set_last_error("mmap failed"); return nullptr; }
#ifdef SYS_mbind
    const int node = numa_node_of_device(device, nullptr);
    if (node >= 0 && node < 1024) {
        unsigned long mask[16] = {0};
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        syscall(SYS_mbind, p, bytes, 1 /* MPOL_PREFERRED */, mask, 1024ul + 1ul, 0u);      // best effort
    }
#endif
    int cur = -1;
    cudaGetDevice(&cur);
    if (device >= 0 && cur != device) cudaSetDevice(device);
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (device >= 0 && cur >= 0 && cur != device) cudaSetDevice(cur);
    if (e != cudaSuccess) {
        cudaGetLastError();
        munmap(p, bytes);
        ffv1::set_last_error(std::string("cudaHostRegister: ") + cudaGetErrorString(e));
        return nullptr;
    }
    std::lock_guard<std::mutex> g(g_host_mu);
    g_host_allocs[p] = bytes;
    return p;
}

void ffv1b200_host_free(void *p)
{
    if (!p) return;
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> g(g_host_mu);
        auto it = g_host_allocs.find(p);
        if (it == g_host_allocs.end()) return;
        bytes = it->second;
        g_host_allocs.erase(it);
    }
    cudaHostUnregister(p);
    cudaGetLastError();
    munmap(p, bytes);
}

// Host side of the copy path: the staging buffers of a GPU should live on the NUMA node its PCIe root port hangs off, and
// the thread that fills / submits them should run there.  Reads the node from sysfs, pins the calling thread to the
// node's CPUs and makes the node the preferred one for the thread's future allocations (pinned buffers included: they
// are placed when they are allocated).  Returns the node, or FFV1B200_ERR_ENOSYS when the topology is not visible
// (single-node machines, containers without sysfs): nothing is changed then.
int ffv1b200_bind_thread_to_device(int device)
{
    char bdf[32] = {0};
    char path[128];
    const int node = numa_node_of_device(device, bdf);
    if (node == -2) { ffv1::set_last_error("no such CUDA device"); return FFV1B200_ERR_EINVAL; }
    if (node < 0) { ffv1::set_last_error(std::string("NUMA node of ") + bdf + " not visible"); return FFV1B200_ERR_ENOSYS; }
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE *fh = fopen(path, "r");
    if (!fh) { ffv1::set_last_error(std::string(path) + " not readable"); return FFV1B200_ERR_ENOSYS; }
    char list[4096] = {0};
    const bool got = fgets(list, sizeof(list), fh) != nullptr;
    fclose(fh);
    cpu_set_t allowed, want;
    CPU_ZERO(&want);
    if (!got || sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return FFV1B200_ERR_ENOSYS;
    int ncpu = 0;
    for (char *q = list; *q && *q != '\n';) {                       // "0-31,64-95"
        char *end;
        long a = strtol(q, &end, 10), b = a;
        if (end == q) break;
        if (*end == '-') { q = end + 1; b = strtol(q, &end, 10); }
        for (long c = a; c <= b && c < CPU_SETSIZE; c++) if (CPU_ISSET(c, &allowed)) { CPU_SET(c, &want); ncpu++; }
        q = *end == ',' ? end + 1 : end;
    }
    if (!ncpu) { ffv1::set_last_error("no allowed CPU on the device's NUMA node"); return FFV1B200_ERR_ENOSYS; }
    if (sched_setaffinity(0, sizeof(want), &want) != 0) return FFV1B200_ERR_ENOSYS;
#ifdef SYS_set_mempolicy
    if (node < 1024) {
        unsigned long mask[16] = {0};
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, 1024ul + 1ul);     // best effort (may be filtered in containers)
    }
#endif
    return node;
}

} // extern "C"
