// common.cuh -- error plumbing and grow-only device scratch buffers shared by the .cu files.
#pragma once
#include <atomic>
#include <cstdint>
#include <string>

#include <cuda_runtime.h>

#include "../../include/svfm.h"

namespace svfm {

extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launches;

#define SVFM_CUDA(expr)                                                                          \
    do {                                                                                         \
        cudaError_t svfm_e_ = (expr);                                                            \
        if (svfm_e_ != cudaSuccess) {                                                            \
            ::svfm::g_last_error = std::string(#expr) + ": " + cudaGetErrorString(svfm_e_);      \
            return svfm_e_ == cudaErrorMemoryAllocation ? SVFM_ERR_NOMEM : SVFM_ERR_CUDA;        \
        }                                                                                        \
    } while (0)

// Grow-only device allocation.  Growing frees the old block with cudaFree, which waits for the device,
// so work already enqueued on the old block finishes first.
struct DeviceBuffer {
    void* ptr = nullptr;
    uint64_t cap = 0;
    int reserve(uint64_t bytes) {
        if (bytes <= cap) return SVFM_OK;
        if (ptr) { cudaFree(ptr); ptr = nullptr; cap = 0; }
        uint64_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            e = cudaMalloc(&ptr, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            ptr = nullptr;
            g_last_error = std::string("cudaMalloc(scratch): ") + cudaGetErrorString(e);
            return e == cudaErrorMemoryAllocation ? SVFM_ERR_NOMEM : SVFM_ERR_CUDA;
        }
        cap = want;
        return SVFM_OK;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
    ~DeviceBuffer() { release(); }
    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    DeviceBuffer(DeviceBuffer&& o) noexcept : ptr(o.ptr), cap(o.cap) { o.ptr = nullptr; o.cap = 0; }
    DeviceBuffer& operator=(DeviceBuffer&& o) noexcept {
        if (this != &o) { release(); ptr = o.ptr; cap = o.cap; o.ptr = nullptr; o.cap = 0; }
        return *this;
    }
};

}  // namespace svfm
