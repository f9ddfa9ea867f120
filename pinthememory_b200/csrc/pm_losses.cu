// The two write losses as stand-alone entry points: `Memory_sup.diversityloss(mem)` (memory.py:264-272) and
// `Memory_sup.classification_loss(mem)` (memory.py:259-262) of the reference's module interface, forward and
// backward. Inside `write()` both are fused into pm_update_fwd / pm_update_bwd; these exist so that the methods of
// the reference class stay callable on an arbitrary [K,C] memory (rows need not be unit length).
// One CTA; K <= 31 slots, any C. Everything is [K,C]-sized (19 x 256 floats): latency, not throughput.
#include "pm_common.cuh"

namespace pm {

constexpr int ML_THREADS = 256;
constexpr int ML_KMAX = 32;

// out[0] = div = (sum_{i != j} max(<m_i, m_j>, 0)) / (K (K-1));  out[1] = cls = mean_i CE(W m_i + b, i)
// gram [K*K] and prob [K*K] (softmax rows) are kept for the backward.
__global__ void __launch_bounds__(ML_THREADS) memory_losses_fwd_kernel(const float* __restrict__ mem, const float* __restrict__ W,
                                                                       const float* __restrict__ bias, int K, int C,
                                                                       float* __restrict__ out, float* __restrict__ gram,
                                                                       float* __restrict__ prob) {
    __shared__ float g[ML_KMAX * ML_KMAX], z[ML_KMAX * ML_KMAX];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = ML_THREADS / 32;
    for (int p = warp; p < K * K; p += nw) {  // one warp per (i, j): both dot products
        const int i = p / K, j = p - i * K;
        float a = 0.f, b = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float mi = mem[i * C + c];
            a = fmaf(mi, mem[j * C + c], a);
            if (W != nullptr) b = fmaf(mi, W[j * C + c], b);
        }
        a = warp_sum(a), b = warp_sum(b);
        if (lane == 0) {
            g[p] = a;
            z[p] = W != nullptr ? b + bias[j] : 0.f;
        }
    }
    __syncthreads();
    if (warp == 0) {
        float div = 0.f;
        for (int p = lane; p < K * K; p += 32) {
            const int i = p / K, j = p - i * K;
            if (i != j) div += fmaxf(g[p], 0.f);
            if (gram != nullptr) gram[p] = g[p];
        }
        div = warp_sum(div);
        float cls = 0.f;
        if (W != nullptr && lane < K) {  // lane = row i
            float mx = -INFINITY;
            for (int k = 0; k < K; ++k) mx = fmaxf(mx, z[lane * K + k]);
            float s = 0.f;
            for (int k = 0; k < K; ++k) s += expf(z[lane * K + k] - mx);
            const float lse = mx + logf(s);
            cls = lse - z[lane * K + lane];
            if (prob != nullptr)
                for (int k = 0; k < K; ++k) prob[lane * K + k] = expf(z[lane * K + k] - lse);
        }
        cls = warp_sum(cls);
        if (lane == 0) {
            out[0] = div / (float)(K * (K - 1));
            out[1] = W != nullptr ? cls / (float)K : 0.f;
        }
    }
}

// dmem [K,C] = g_div * d(div)/d(mem) + g_cls * d(cls)/d(mem);  dW [K,C], db [K] (only when W is given)
__global__ void __launch_bounds__(ML_THREADS) memory_losses_bwd_kernel(const float* __restrict__ mem, const float* __restrict__ W,
                                                                       const float* __restrict__ gram, const float* __restrict__ prob,
                                                                       const float* __restrict__ g_div, const float* __restrict__ g_cls,
                                                                       int K, int C, float* __restrict__ dmem,
                                                                       float* __restrict__ dW, float* __restrict__ db) {
    const float gd = g_div != nullptr ? *g_div / (float)(K * (K - 1)) : 0.f;
    const float gc = (g_cls != nullptr && W != nullptr) ? *g_cls / (float)K : 0.f;
    for (int e = threadIdx.x; e < K * C; e += ML_THREADS) {
        const int i = e / C, c = e - i * C;
        float acc = 0.f, accw = 0.f;
        for (int j = 0; j < K; ++j) {
            // d/dm_i sum_{a != b} relu(<m_a, m_b>): the pair (i,j) appears twice
            if (j != i && gram[i * K + j] > 0.f) acc = fmaf(2.f * gd, mem[j * C + c], acc);
            if (W != nullptr) {
                const float dz_ij = gc * (prob[i * K + j] - (i == j ? 1.f : 0.f));  // d cls / d z[i][j]
                acc = fmaf(dz_ij, W[j * C + c], acc);
                const float dz_ji = gc * (prob[j * K + i] - (i == j ? 1.f : 0.f));  // row j, class i -> dW[i]
                accw = fmaf(dz_ji, mem[j * C + c], accw);
            }
        }
        dmem[e] = acc;
        if (W != nullptr && dW != nullptr) dW[e] = accw;
    }
    if (W != nullptr && db != nullptr && threadIdx.x < K) {
        float s = 0.f;
        for (int j = 0; j < K; ++j) s += gc * (prob[j * K + threadIdx.x] - (j == (int)threadIdx.x ? 1.f : 0.f));
        db[threadIdx.x] = s;
    }
}

}  // namespace pm

extern "C" int pm_memory_losses_fwd(const float* mem, const float* W_cls, const float* b_cls, int K, int C, float* out,
                                    float* gram, float* prob, void* stream) {
    if (!mem || !out || ((W_cls == nullptr) != (b_cls == nullptr))) return PM_ERR_NULL;
    if (K < 2 || K > 31) return PM_ERR_SLOTS;
    if (C <= 0) return PM_ERR_SHAPE;
    pm::memory_losses_fwd_kernel<<<1, pm::ML_THREADS, 0, (cudaStream_t)stream>>>(mem, W_cls, b_cls, K, C, out, gram, prob);
    PM_CHECK_LAUNCH();
    return 0;
}

extern "C" int pm_memory_losses_bwd(const float* mem, const float* W_cls, const float* gram, const float* prob,
                                    const float* g_div, const float* g_cls, int K, int C, float* dmem, float* dW_cls,
                                    float* db_cls, void* stream) {
    if (!mem || !gram || !dmem || (W_cls != nullptr && !prob)) return PM_ERR_NULL;
    if (K < 2 || K > 31) return PM_ERR_SLOTS;
    if (C <= 0) return PM_ERR_SHAPE;
    pm::memory_losses_bwd_kernel<<<1, pm::ML_THREADS, 0, (cudaStream_t)stream>>>(mem, W_cls, gram, prob, g_div, g_cls, K, C, dmem,
                                                                                  dW_cls, db_cls);
    PM_CHECK_LAUNCH();
    return 0;
}
