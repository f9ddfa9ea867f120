// Read-bandwidth probe for the BatchNorm-statistics access pattern (per-channel sum and sum of squares over an
// NCHW tensor [B=8, C=256, hw=9216] fp32, 75.5 MB): which load mechanism gets a reduction closest to the HBM peak?
//   A: one CTA per channel, LDG.128 grid-stride (the r1c bn_stats kernel)
//   B: one CTA per (channel, batch-half), 8 independent LDG.128 per thread in flight
//   C: one CTA per channel, cp.async.bulk (TMA 1-D) into an mbarrier ring, all threads reduce from shared memory
//   D: as C with two CTAs per channel
// Rotates over 8 tensors (604 MB) so that nothing is served from the 126 MB L2.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o readbw readbw.cu && ./readbw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int B = 8, C = 256, HW = 9216;

__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ void block_out(float s, float q, float* out, int idx, bool atomic) {
    __shared__ float rs[32], rq[32];
    s = warp_sum(s), q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) rs[threadIdx.x >> 5] = s, rq[threadIdx.x >> 5] = q;
    __syncthreads();
    if (threadIdx.x == 0) {
        float S = 0, Q = 0;
        for (int i = 0; i < (int)blockDim.x / 32; ++i) S += rs[i], Q += rq[i];
        if (atomic) atomicAdd(out + 2 * idx, S), atomicAdd(out + 2 * idx + 1, Q);
        else out[2 * idx] = S, out[2 * idx + 1] = Q;
    }
}

__global__ void __launch_bounds__(512) varA(const float* x, float* out) {
    const int c = blockIdx.x, nv = HW / 4, total = B * nv;
    float s = 0, q = 0;
#pragma unroll 4
    for (int i = threadIdx.x; i < total; i += 512) {
        int b = i / nv, j = i - b * nv;
        float4 v = *reinterpret_cast<const float4*>(x + ((size_t)b * C + c) * HW + 4 * j);
        s += v.x + v.y + v.z + v.w;
        q += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    block_out(s, q, out, c, false);
}

template <int SPLIT>
__global__ void __launch_bounds__(256) varB(const float* x, float* out) {
    const int c = blockIdx.x / SPLIT, part = blockIdx.x % SPLIT;
    constexpr int BP = B / SPLIT, NV = HW / 4;  // 2304 float4 per plane = 9 x 256
    float s = 0, q = 0;
    for (int b = part * BP; b < (part + 1) * BP; ++b) {
        const float4* p = reinterpret_cast<const float4*>(x + ((size_t)b * C + c) * HW);
        float4 v[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) v[i] = p[threadIdx.x + 256 * i];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            s += v[i].x + v[i].y + v[i].z + v[i].w;
            q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
        }
    }
    block_out(s, q, out, c, true);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// CHUNK floats per stage; a plane is HW/CHUNK chunks
template <int SPLIT, int STAGES, int CHUNK, int THREADS>
__global__ void __launch_bounds__(THREADS) varC(const float* x, float* out) {
    extern __shared__ __align__(128) float ring[];  // [STAGES][CHUNK]
    __shared__ uint64_t full[STAGES];
    const int c = blockIdx.x / SPLIT, part = blockIdx.x % SPLIT;
    constexpr int BP = B / SPLIT, CPP = HW / CHUNK, NCH = BP * CPP;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto src = [&](int ch) {
        int b = part * BP + ch / CPP, k = ch % CPP;
        return x + ((size_t)b * C + c) * HW + (size_t)k * CHUNK;
    };
    if (threadIdx.x == 0)
        for (int i = 0; i < STAGES && i < NCH; ++i) {
            mbar_expect_tx(&full[i], CHUNK * 4);
            bulk_load(ring + i * CHUNK, src(i), CHUNK * 4, &full[i]);
        }
    float s = 0, q = 0;
    for (int ch = 0; ch < NCH; ++ch) {
        const int st = ch % STAGES;
        mbar_wait(&full[st], (ch / STAGES) & 1);
        const float4* p = reinterpret_cast<const float4*>(ring + st * CHUNK);
#pragma unroll
        for (int i = threadIdx.x; i < CHUNK / 4; i += THREADS) {
            float4 v = p[i];
            s += v.x + v.y + v.z + v.w;
            q += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        __syncthreads();  // stage consumed by everyone
        if (threadIdx.x == 0 && ch + STAGES < NCH) {
            mbar_expect_tx(&full[st], CHUNK * 4);
            bulk_load(ring + st * CHUNK, src(ch + STAGES), CHUNK * 4, &full[st]);
        }
    }
    block_out(s, q, out, c, SPLIT > 1);
}

template <typename F>
float time_it(F launch, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) launch(i);
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) launch(i);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
    return ms / reps;
}

int main() {
    const size_t n = (size_t)B * C * HW;
    float *x, *out;
    cudaMalloc(&x, 8 * n * 4);
    cudaMalloc(&out, 2 * C * 4 * 4);
    cudaMemset(x, 0, 8 * n * 4);
    cudaMemset(out, 0, 2 * C * 4 * 4);
    const double mb = n * 4 / 1e6;
    auto rep = [&](const char* name, float ms) { printf("%-44s %7.2f us  %7.1f GB/s\n", name, ms * 1e3, mb / ms); };
    rep("A  CTA/channel LDG.128 512thr", time_it([&](int i) { varA<<<C, 512>>>(x + (i % 8) * n, out); }, 40));
    rep("B2 CTA/(channel,half) 9xLDG.128 256thr", time_it([&](int i) { varB<2><<<C * 2, 256>>>(x + (i % 8) * n, out); }, 40));
    rep("B4 CTA/(channel,quarter) 9xLDG.128 256thr", time_it([&](int i) { varB<4><<<C * 4, 256>>>(x + (i % 8) * n, out); }, 40));
    rep("B8 CTA/plane 9xLDG.128 256thr", time_it([&](int i) { varB<8><<<C * 8, 256>>>(x + (i % 8) * n, out); }, 40));
#define RUNC(SPLIT, ST, CH, TH, NAME)                                                                     \
    {                                                                                                     \
        auto k = varC<SPLIT, ST, CH, TH>;                                                                 \
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * CH * 4);                \
        rep(NAME, time_it([&](int i) { k<<<C * SPLIT, TH, ST * CH * 4>>>(x + (i % 8) * n, out); }, 40)); \
    }
    RUNC(1, 4, 2304, 256, "C  bulk 4x9KB ring, CTA/channel 256thr");
    RUNC(1, 8, 2304, 256, "C  bulk 8x9KB ring, CTA/channel 256thr");
    RUNC(1, 4, 4608, 256, "C  bulk 4x18KB ring, CTA/channel 256thr");
    RUNC(2, 4, 2304, 256, "D  bulk 4x9KB ring, 2 CTAs/channel 256thr");
    RUNC(2, 6, 2304, 128, "D  bulk 6x9KB ring, 2 CTAs/channel 128thr");
    RUNC(4, 4, 2304, 128, "D  bulk 4x9KB ring, 4 CTAs/channel 128thr");
    RUNC(2, 4, 4608, 256, "D  bulk 4x18KB ring, 2 CTAs/channel 256thr");
    cudaDeviceSynchronize();
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
