// Score-side kernels: the dim-0 (over all pixels) softmax and the feature-cohesion read loss.
#include "pm_common.cuh"
#include "pm_internal.h"

namespace pm {

// ------------------------------------------------------------------ score_query: softmax over pixels
// memory.py:183/186. Two small launches: per-CTA online (max, sum) per column, then normalise.
// Threads per CTA = K * floor(256/K) so a thread's column is fixed while its flat index strides rows.

constexpr int CS_MAXG = PM_COLPART_ROWS;

__global__ void __launch_bounds__(256) colsoftmax_stats_kernel(const float* __restrict__ s, const float* __restrict__ g,
                                                               float* __restrict__ partial, int N, int K, int KP,
                                                               int rows_per_cta) {
    __shared__ float sm_m[256], sm_l[256];
    const int t = threadIdx.x, k = t % K, r0 = t / K, rstep = blockDim.x / K;
    const int start = blockIdx.x * rows_per_cta, end = min(N, start + rows_per_cta);
    float m = -INFINITY, l = 0.f;
    for (int r = start + r0; r < end; r += rstep) {
        float z = __ldg(s + (size_t)r * KP + k);
        if (g != nullptr) z += __ldg(g + (size_t)r * K + k);
        if (z > m) {
            l = l * expf(m - z) + 1.f;
            m = z;
        } else {
            l += expf(z - m);
        }
    }
    sm_m[t] = m;
    sm_l[t] = l;
    __syncthreads();
    if (t < K) {
        float M = -INFINITY;
        for (int i = 0; i < rstep; ++i) M = fmaxf(M, sm_m[i * K + t]);
        float L = 0.f;
        for (int i = 0; i < rstep; ++i) {
            float mi = sm_m[i * K + t];
            if (mi > -INFINITY) L += sm_l[i * K + t] * expf(mi - M);
        }
        partial[(size_t)blockIdx.x * 64 + t] = M;
        partial[(size_t)blockIdx.x * 64 + 32 + t] = L;
    }
}

// Combine the PM_COLPART_ROWS per-CTA (max, sum) partials of every column ONCE (one CTA) into the final column maximum and
// 1/sum, stored in row PM_COLPART_ROWS of the partials buffer ([0..32) max, [32..64) 1/sum). The apply kernel that follows
// then runs on as many small CTAs as there are rows to normalise (the round-1 version re-combined the 296 x 64 partials
// in every one of its <= 296 CTAs: 75 KB of L2 reads and a dependent latency chain per CTA, 18 us for 11 MB of work).
__global__ void __launch_bounds__(256) colsoftmax_combine_kernel(float* __restrict__ partial, int K) {  // all PM_COLPART_ROWS rows
    __shared__ float sm_m[256], sm_l[256];
    const int t = threadIdx.x, k = t % K, r0 = t / K, rstep = blockDim.x / K;
    constexpr int MAXJ = (CS_MAXG + 7) / 8;  // rstep >= 8 for K <= 31
    float pm_[MAXJ], pl_[MAXJ];
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {  // all loads issued before the first use (independent L2 hits)
        const int i = r0 + j * rstep;
        const bool ok = i < CS_MAXG;
        pm_[j] = ok ? __ldcg(partial + (size_t)i * 64 + k) : -INFINITY;
        pl_[j] = ok ? __ldcg(partial + (size_t)i * 64 + 32 + k) : 0.f;
    }
    float M = -INFINITY, L = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) M = fmaxf(M, pm_[j]);
#pragma unroll
    for (int j = 0; j < MAXJ; ++j)
        if (pm_[j] > -INFINITY) L += pl_[j] * expf(pm_[j] - M);
    sm_m[t] = M;
    sm_l[t] = L;
    __syncthreads();
    if (t < K) {
        float Mx = -INFINITY;
        for (int i = 0; i < rstep; ++i) Mx = fmaxf(Mx, sm_m[i * K + t]);
        float Ls = 0.f;
        for (int i = 0; i < rstep; ++i) {
            const float mi = sm_m[i * K + t];
            if (mi > -INFINITY) Ls += sm_l[i * K + t] * expf(mi - Mx);
        }
        partial[(size_t)CS_MAXG * 64 + t] = Mx;
        partial[(size_t)CS_MAXG * 64 + 32 + t] = 1.f / Ls;
    }
}

__global__ void __launch_bounds__(256) colsoftmax_apply_kernel(const float* __restrict__ s, const float* __restrict__ g,
                                                               const float* __restrict__ partial,
                                                               float* __restrict__ out, int N, int K, int KP,
                                                               int rows_per_cta) {
    const int t = threadIdx.x, k = t % K, r0 = t / K, rstep = blockDim.x / K;
    const float M = __ldg(partial + (size_t)CS_MAXG * 64 + k), il = __ldg(partial + (size_t)CS_MAXG * 64 + 32 + k);
    const int start = blockIdx.x * rows_per_cta, end = min(N, start + rows_per_cta);
    for (int r = start + r0; r < end; r += rstep) {
        float z = __ldg(s + (size_t)r * KP + k);
        if (g != nullptr) z += __ldg(g + (size_t)r * K + k);
        out[(size_t)r * K + k] = expf(z - M) * il;
    }
}

// -------------------------------------------------------- score_memory for the external get_score path
// Row softmax of s (+ gumbel_m) -> dense [N,K]. (Inside Memory_sup.forward this is fused in read_fwd.)
template <int KP>
__global__ void __launch_bounds__(256) rowsoftmax_kernel(const float* __restrict__ s, const float* __restrict__ g,
                                                         float* __restrict__ out, int N, int K) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float z[KP], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        z[k] = (k < K) ? __ldg(s + (size_t)n * KP + k) + (g ? __ldg(g + (size_t)n * K + k) : 0.f) : -INFINITY;
        mx = fmaxf(mx, z[k]);
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        z[k] = (k < K) ? expf(z[k] - mx) : 0.f;
        sum += z[k];
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int k = 0; k < KP; ++k)
        if (k < K) out[(size_t)n * K + k] = z[k] * inv;
}

}  // namespace pm

extern "C" int pm_score_stride(int K) { return K <= 19 ? 20 : 32; }
extern "C" int pm_colsoftmax_workspace_floats(int K) { return (pm::CS_MAXG + 2) * 64; }  // + the combined row + the ticket row

namespace pm {
int colsoftmax_stats(const float* s, const float* gumbel_q, float* partial, int N, int K, cudaStream_t st) {
    const int KP = pm_score_stride(K), tpb = K * (256 / K);
    const int rows_per_cta = (N + CS_MAXG - 1) / CS_MAXG;  // CTAs past the end write (-inf, 0)
    colsoftmax_stats_kernel<<<CS_MAXG, tpb, 0, st>>>(s, gumbel_q, partial, N, K, KP, rows_per_cta);
    colsoftmax_combine_kernel<<<1, tpb, 0, st>>>(partial, K);  // -> row PM_COLPART_ROWS: column maximum, 1/sum
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}
}  // namespace pm

extern "C" int pm_colsoftmax_apply(const float* s, const float* gumbel_q, const float* col_partials, float* score_q,
                                   int N, int K, void* stream) {
    if (!s || !score_q || !col_partials) return PM_ERR_NULL;
    if (K < 1 || K > 31) return PM_ERR_SLOTS;
    if (N <= 0) return PM_ERR_SHAPE;
    const int KP = pm_score_stride(K), tpb = K * (256 / K), rstep = tpb / K;
    // the combined row (column maximum, 1/sum) was left by pm_read_fwd[_planes] / the statistics pass
    const int rows_per_cta = rstep * 5;   // 5 rows per thread: ~1100 small CTAs at cfg 2
    const int G = (N + rows_per_cta - 1) / rows_per_cta;
    pm::colsoftmax_apply_kernel<<<G, tpb, 0, (cudaStream_t)stream>>>(s, gumbel_q, col_partials, score_q, N, K, KP,
                                                                       rows_per_cta);
    PM_CHECK_LAUNCH();
    return 0;
}

extern "C" int pm_colsoftmax(const float* s, const float* gumbel_q, float* score_q, float* workspace, int N, int K,
                             void* stream) {
    if (!s || !score_q || !workspace) return PM_ERR_NULL;
    if (K < 1 || K > 31) return PM_ERR_SLOTS;
    if (N <= 0) return PM_ERR_SHAPE;
    if (int e = pm::colsoftmax_stats(s, gumbel_q, workspace, N, K, (cudaStream_t)stream)) return e;
    return pm_colsoftmax_apply(s, gumbel_q, workspace, score_q, N, K, stream);
}

extern "C" int pm_rowsoftmax(const float* s, const float* gumbel_m, float* score_m, int N, int K, void* stream) {
    if (!s || !score_m) return PM_ERR_NULL;
    if (K < 1 || K > 31) return PM_ERR_SLOTS;
    if (N <= 0) return PM_ERR_SHAPE;
    const int grid = (N + 255) / 256;
    if (K <= 19)
        pm::rowsoftmax_kernel<20><<<grid, 256, 0, (cudaStream_t)stream>>>(s, gumbel_m, score_m, N, K);
    else
        pm::rowsoftmax_kernel<32><<<grid, 256, 0, (cudaStream_t)stream>>>(s, gumbel_m, score_m, N, K);
    PM_CHECK_LAUNCH();
    return 0;
}
