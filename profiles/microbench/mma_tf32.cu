// Microbenchmark: legacy mma.sync TF32 (m16n8k8) and BF16 (m16n8k16) throughput on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_tf32 mma_tf32.cu && ./mma_tf32
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    unsigned a0 = threadIdx.x * 2654435761u, a1 = a0 ^ 0x3f800000u, a2 = a0 + 7, a3 = a1 + 9, b0 = a0 >> 3, b1 = a1 >> 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4000;
    for (int mode = 0; mode < 2; ++mode)
        for (int occ = 1; occ <= 8; occ *= 2)
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<148 * occ, 256>>>(out, iters);
                else k<1><<<148 * occ, 256>>>(out, iters);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                double mac = 148.0 * occ * 8 * (double)iters * 8 * (mode == 0 ? 16 * 8 * 8 : 16 * 8 * 16);
                if (rep) printf("%s CTAs/SM=%d: %.3f ms  %.1f TFLOP/s  (%.2f MMA/clk/SM @1.9GHz)\n", mode ? "bf16 m16n8k16" : "tf32 m16n8k8 ", occ, ms,
                                2 * mac / ms / 1e9, 148.0 * occ * 8 * (double)iters * 8 / (ms * 1e-3) / 148 / 1.9e9);
            }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
