/*
 * ek_misc.cu -- trivial helper kernels: fill / reverse (reference: src/cuda/common.cu:56-102)
 * and an L2-flush writer used by the benchmark between timed iterations.
 */
#include <cuda_runtime.h>
#include <stdint.h>

namespace {
template <typename T> __global__ void fill_kernel(T *out, T value, size_t n) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
        out[i] = value;
}
template <typename T> __global__ void reverse_kernel(T *out, const T *in, size_t n) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
        out[i] = in[n - 1 - i];
}
__global__ void flush_kernel(uint4 *buf, size_t n16) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t) gridDim.x * blockDim.x)
        buf[i] = make_uint4((uint32_t) i, 0u, 0u, 0u);
}
inline unsigned grid_for(size_t n) { size_t g = (n + 255) / 256; return (unsigned) (g > 148 * 8 ? 148 * 8 : (g ? g : 1)); }
}

void ek_launch_fill(void *ptr, size_t elem_size, uint64_t value, size_t n, cudaStream_t stream) {
    switch (elem_size) {
        case 1: fill_kernel<uint8_t><<<grid_for(n), 256, 0, stream>>>((uint8_t *) ptr, (uint8_t) value, n); break;
        case 2: fill_kernel<uint16_t><<<grid_for(n), 256, 0, stream>>>((uint16_t *) ptr, (uint16_t) value, n); break;
        case 4: fill_kernel<uint32_t><<<grid_for(n), 256, 0, stream>>>((uint32_t *) ptr, (uint32_t) value, n); break;
        default: fill_kernel<uint64_t><<<grid_for(n), 256, 0, stream>>>((uint64_t *) ptr, value, n); break;
    }
}
void ek_launch_reverse(void *out, const void *in, size_t elem_size, size_t n, cudaStream_t stream) {
    switch (elem_size) {
        case 1: reverse_kernel<uint8_t><<<grid_for(n), 256, 0, stream>>>((uint8_t *) out, (const uint8_t *) in, n); break;
        case 2: reverse_kernel<uint16_t><<<grid_for(n), 256, 0, stream>>>((uint16_t *) out, (const uint16_t *) in, n); break;
        case 4: reverse_kernel<uint32_t><<<grid_for(n), 256, 0, stream>>>((uint32_t *) out, (const uint32_t *) in, n); break;
        default: reverse_kernel<uint64_t><<<grid_for(n), 256, 0, stream>>>((uint64_t *) out, (const uint64_t *) in, n); break;
    }
}
void ek_launch_flush(void *buf, size_t bytes, cudaStream_t stream) {
    flush_kernel<<<148 * 8, 256, 0, stream>>>((uint4 *) buf, bytes / 16);
}
