// Probe: how fast can a B200 WRITE decoded frames? (tools/probes, not product.) Compares a linear fill with the 4x4-block
// pattern of reconstruct_k (a thread writes four 16-byte row pieces, 7680 bytes apart at 1080p), per-frame loop inside the
// kernel or one thread per (frame, block).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/write_pattern tools/probes/write_pattern.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void linear_k(uint4* out, size_t n16, uint32_t v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) out[i] = make_uint4(v, v, v, v);
}
// grid (cdiv(B,128)), loops over frames
template <bool CS>
__global__ void __launch_bounds__(128) block_loop_k(uint32_t* out, uint32_t frames, uint32_t W, uint32_t H, uint32_t v) {
    const uint32_t bw = W >> 2, B = bw * (H >> 2);
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const size_t pix0 = (size_t)(b / bw) * 4 * W + (b % bw) * 4, P = (size_t)W * H;
    for (uint32_t k = 0; k < frames; k++) {
        uint32_t* d = out + k * P + pix0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (CS) __stcs(reinterpret_cast<uint4*>(d + (size_t)j * W), make_uint4(v, v + k, v, v));
            else *reinterpret_cast<uint4*>(d + (size_t)j * W) = make_uint4(v, v + k, v, v);
        }
    }
}
// grid (cdiv(B,128), frames)
__global__ void __launch_bounds__(128) block_par_k(uint32_t* out, uint32_t W, uint32_t H, uint32_t v) {
    const uint32_t bw = W >> 2, B = bw * (H >> 2);
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const size_t pix0 = (size_t)(b / bw) * 4 * W + (b % bw) * 4, P = (size_t)W * H;
    uint32_t* d = out + blockIdx.y * P + pix0;
#pragma unroll
    for (int j = 0; j < 4; j++) *reinterpret_cast<uint4*>(d + (size_t)j * W) = make_uint4(v, v, v, v);
}
// a warp writes whole rows: thread per 16 bytes, linear inside the frame (what a row-major painter would do)
__global__ void __launch_bounds__(256) row_par_k(uint32_t* out, uint32_t W, uint32_t H, uint32_t v) {
    const size_t P = (size_t)W * H;
    const size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4;
    if (i >= P) return;
    *reinterpret_cast<uint4*>(out + blockIdx.y * P + i) = make_uint4(v, v, v, v);
}

int main() {
    const uint32_t W = 1920, H = 1080, F = 1497;
    const size_t P = (size_t)W * H, bytes = P * 4 * F;
    uint32_t* out;
    CK(cudaMalloc(&out, bytes));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const uint32_t B = (W / 4) * (H / 4);
    float ms;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(a); cudaMemsetAsync(out, 1, bytes); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms, a, b);
        printf("cudaMemset            %7.2f ms  %6.0f GB/s\n", ms, bytes / ms / 1e6);
        cudaEventRecord(a); linear_k<<<148 * 16, 256>>>(reinterpret_cast<uint4*>(out), bytes / 16, 2); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms, a, b);
        printf("linear uint4 stores   %7.2f ms  %6.0f GB/s\n", ms, bytes / ms / 1e6);
        cudaEventRecord(a); block_loop_k<false><<<(B + 127) / 128, 128>>>(out, F, W, H, 3); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms, a, b);
        printf("block pattern, frame loop in kernel        %7.2f ms  %6.0f GB/s\n", ms, bytes / ms / 1e6);
        cudaEventRecord(a); block_loop_k<true><<<(B + 127) / 128, 128>>>(out, F, W, H, 3); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms, a, b);
        printf("block pattern, frame loop, st.cs           %7.2f ms  %6.0f GB/s\n", ms, bytes / ms / 1e6);
        cudaEventRecord(a); block_par_k<<<dim3((B + 127) / 128, F), 128>>>(out, W, H, 4); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms, a, b);
        printf("block pattern, thread per (frame, block)   %7.2f ms  %6.0f GB/s\n", ms, bytes / ms / 1e6);
        cudaEventRecord(a); row_par_k<<<dim3((unsigned)((P / 4 + 255) / 256), F), 256>>>(out, W, H, 5); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms, a, b);
        printf("row-major, thread per (frame, 16 bytes)    %7.2f ms  %6.0f GB/s\n", ms, bytes / ms / 1e6);
    }
    CK(cudaGetLastError());
    return 0;
}
