// lvkb200_stream: one video stream == one lvk::StabilizationFilter instance.
#pragma once

#include <memory>
#include <vector>

#include "common.hpp"

namespace lvkb200
{

// Grow-only device / pinned-host buffers.
struct DeviceBuffer
{
    void* ptr = nullptr;
    size_t capacity = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= capacity) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        capacity = 0;
        cudaError_t e = cudaMalloc(&ptr, bytes);
        if (e == cudaSuccess) capacity = bytes;
        return e;
    }
    void release()
    {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        capacity = 0;
    }
    template <typename T>
    T* as() const { return static_cast<T*>(ptr); }
};

struct PinnedBuffer
{
    void* ptr = nullptr;
    size_t capacity = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= capacity) return cudaSuccess;
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        capacity = 0;
        // mapped: kernels may read/write it directly over PCIe (zero-copy), bypassing the copy engines
        cudaError_t e = cudaHostAlloc(&ptr, bytes, cudaHostAllocMapped);
        if (e == cudaSuccess) capacity = bytes;
        return e;
    }
    // Device-side address of this host allocation (zero-copy access).
    template <typename T>
    T* device_view() const
    {
        void* d = nullptr;
        if (!ptr || cudaHostGetDevicePointer(&d, ptr, 0) != cudaSuccess) return nullptr;
        return static_cast<T*>(d);
    }
    void release()
    {
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        capacity = 0;
    }
    template <typename T>
    T* as() const { return static_cast<T*>(ptr); }
};

}  // namespace lvkb200
