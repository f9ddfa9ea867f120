// Folding the memory into the 1x1 convolution that follows the read (include/pinmem_b200.h, "score planes"):
//   W' [Co, C + PM_PLANES] = [ W1 | W2 . M^T | 0 ]   from  W = [W1 | W2] [Co, 2C]  and  M [K, C]
// and the gradient of that map w.r.t. W. Two tiny kernels (one CTA per output channel) instead of a dozen
// slice / pad / matmul / cat launches and their autograd mirrors.
#include "pm_common.cuh"

namespace pm {

__global__ void __launch_bounds__(256) fold_weight_fwd_kernel(const float* __restrict__ W, const float* __restrict__ M,
                                                              float* __restrict__ Wp, int C, int K) {
    extern __shared__ float w2[];  // [C] second half of this output channel's row
    const int co = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float* wr = W + (size_t)co * 2 * C;
    float* out = Wp + (size_t)co * (C + PM_PLANES);
    for (int c = tid; c < C; c += 256) {
        out[c] = wr[c];
        w2[c] = wr[C + c];
    }
    __syncthreads();
    for (int k = wid; k < PM_PLANES; k += 8) {
        float acc = 0.f;
        if (k < K)
            for (int c = lane; c < C; c += 32) acc = fmaf(w2[c], __ldg(M + (size_t)k * C + c), acc);
        acc = warp_sum(acc);
        if (lane == 0) out[C + k] = acc;
    }
}

__global__ void __launch_bounds__(256) fold_weight_bwd_kernel(const float* __restrict__ dWp, const float* __restrict__ M,
                                                              float* __restrict__ dW, int C, int K) {
    __shared__ float dg[PM_PLANES];
    const int co = blockIdx.x, tid = threadIdx.x;
    const float* in = dWp + (size_t)co * (C + PM_PLANES);
    float* out = dW + (size_t)co * 2 * C;
    if (tid < PM_PLANES) dg[tid] = (tid < K) ? in[C + tid] : 0.f;
    __syncthreads();
    for (int c = tid; c < C; c += 256) {
        out[c] = in[c];
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(dg[k], __ldg(M + (size_t)k * C + c), acc);
        out[C + c] = acc;
    }
}

}  // namespace pm

extern "C" int pm_fold_weight_fwd(const float* W, const float* M, float* Wp, int Co, int C, int K, void* stream) {
    if (!W || !M || !Wp) return PM_ERR_NULL;
    if (Co <= 0 || C <= 0 || C > 4096) return PM_ERR_SHAPE;
    if (K < 1 || K > 31) return PM_ERR_SLOTS;
    pm::fold_weight_fwd_kernel<<<Co, 256, sizeof(float) * C, (cudaStream_t)stream>>>(W, M, Wp, C, K);
    PM_CHECK_LAUNCH();
    return 0;
}

extern "C" int pm_fold_weight_bwd(const float* dWp, const float* M, float* dW, int Co, int C, int K, void* stream) {
    if (!dWp || !M || !dW) return PM_ERR_NULL;
    if (Co <= 0 || C <= 0 || C > 4096) return PM_ERR_SHAPE;
    if (K < 1 || K > 31) return PM_ERR_SLOTS;
    pm::fold_weight_bwd_kernel<<<Co, 256, 0, (cudaStream_t)stream>>>(dWp, M, dW, C, K);
    PM_CHECK_LAUNCH();
    return 0;
}
