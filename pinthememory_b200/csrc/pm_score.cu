// Score-side kernels: the dim-0 (over all pixels) softmax and the feature-cohesion read loss.
#include "pm_common.cuh"

namespace pm {

// ------------------------------------------------------------------ score_query: softmax over pixels
// memory.py:183/186. Two small launches: per-CTA online (max, sum) per column, then normalise.
// Threads per CTA = K * floor(256/K) so a thread's column is fixed while its flat index strides rows.

constexpr int CS_MAXG = 296;

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

__global__ void __launch_bounds__(256) colsoftmax_apply_kernel(const float* __restrict__ s, const float* __restrict__ g,
                                                               const float* __restrict__ partial,
                                                               float* __restrict__ out, int N, int K, int KP,
                                                               int rows_per_cta, int G) {
    __shared__ float col_m[32], col_il[32];
    const int t = threadIdx.x;
    if (t < 32) {  // one warp: lanes stride the G partials of each column
        for (int k = 0; k < K; ++k) {
            float M = -INFINITY;
            for (int i = t; i < G; i += 32) M = fmaxf(M, partial[(size_t)i * 64 + k]);
            M = warp_max(M);
            float L = 0.f;
            for (int i = t; i < G; i += 32) {
                float mi = partial[(size_t)i * 64 + k];
                if (mi > -INFINITY) L += partial[(size_t)i * 64 + 32 + k] * expf(mi - M);
            }
            L = warp_sum(L);
            if (t == 0) {
                col_m[k] = M;
                col_il[k] = 1.f / L;
            }
        }
    }
    __syncthreads();
    const int k = t % K, r0 = t / K, rstep = blockDim.x / K;
    const int start = blockIdx.x * rows_per_cta, end = min(N, start + rows_per_cta);
    const float M = col_m[k], il = col_il[k];
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

// ------------------------------------------------------------------------------- read loss (forward)
// CE(bilinear_up(s/T), labels) without the [B,K,Hm,Wm] logits. A "cell" is the set of label pixels whose
// four bilinear taps are the feature pixels (cy,cx),(cy,cx+1),(cy+1,cx),(cy+1,cx+1). One thread walks
// the label pixels of one cell (optionally 1/RS of its rows): inside a cell the K logits are bilinear
// in (lambda_y, lambda_x), so per label row it forms A_k + lambda_x*B_k and each pixel costs K FMAs +
// K ex2. The gradient w.r.t. the four taps is accumulated in registers (softmax part) and in
// thread-private shared-memory columns (one-hot part and label histogram: dynamic class index, no
// atomics), then merged into the CTA's tap tile in conflict-free phases and flushed to ds_rl with
// 16-byte vector REDs (only tile borders are shared between CTAs).

constexpr int RL_TX = 32;          // cells per CTA along x (= lanes)
constexpr int RL_LDX = RL_TX + 1;  // taps per tile row

__device__ __forceinline__ int bil_index(float scale, int dst, int n_in) {
    int i0 = (int)(scale * (float)dst);
    return i0 > n_in - 1 ? n_in - 1 : i0;
}
// smallest dst in [0, n_out] whose source index is >= c
__device__ __forceinline__ int first_ge(int c, float scale, int n_out, int n_in) {
    if (c <= 0) return 0;
    if (c > n_in - 1 || scale <= 0.f) return n_out;
    int y = (int)ceilf((float)c / scale);
    y = max(0, min(y, n_out));
    while (y > 0 && bil_index(scale, y - 1, n_in) >= c) --y;
    while (y < n_out && bil_index(scale, y, n_in) < c) ++y;
    return y;
}

template <int KP>
__global__ void __launch_bounds__(256) readloss_kernel(const float* __restrict__ s, const long long* __restrict__ labels,
                                                       float inv_T, float temperature, int h, int w, int Hm, int Wm,
                                                       int K, float sy, float sx, int RS, int TYC, int tiles_x,
                                                       int tiles_y, float* __restrict__ ds_rl,
                                                       unsigned long long* __restrict__ ws, float* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    const int nthr = blockDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile_elems = (TYC + 1) * RL_LDX * KP;
    float* s_tile = smem;                     // [(TYC+1)][33][KP]
    float* ds_tile = s_tile + tile_elems;     // same shape
    float* priv = ds_tile + tile_elems;       // [5][KP][nthr]: 4 one-hot tap weights + int counts
    int* hist_sm = reinterpret_cast<int*>(priv + 5 * KP * nthr);  // [KP]
    float* red = reinterpret_cast<float*>(hist_sm + KP);          // [8]

    int bid = blockIdx.x;
    const int tx_i = bid % tiles_x;
    bid /= tiles_x;
    const int ty_i = bid % tiles_y, b = bid / tiles_y;
    const int fy0 = ty_i * TYC, fx0 = tx_i * RL_TX;

    for (int i = tid; i < tile_elems; i += nthr) {
        int e = i / KP, k = i - e * KP;
        int ty = e / RL_LDX, tx = e - ty * RL_LDX;
        int fy = min(fy0 + ty, h - 1), fx = min(fx0 + tx, w - 1);
        s_tile[i] = __ldg(s + ((size_t)(b * h + fy) * w + fx) * KP + k);
        ds_tile[i] = 0.f;
    }
    for (int i = tid; i < 5 * KP * nthr; i += nthr) priv[i] = 0.f;
    if (tid < KP) hist_sm[tid] = 0;
    __syncthreads();

    const int cyl = wid / RS, split = wid - cyl * RS;
    const int cy = fy0 + cyl, cx = fx0 + lane;
    const bool active = (cy < h) && (cx < w) && (cyl < TYC);
    const int ly0 = cyl, ly1 = min(cy + 1, h - 1) - fy0;
    const int lx0 = lane, lx1 = min(cx + 1, w - 1) - fx0;
    float G00[KP], G01[KP], G10[KP], G11[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) G00[k] = G01[k] = G10[k] = G11[k] = 0.f;
    float loss = 0.f;
    float* pv = priv + tid;  // element (row, k) at pv[(row*KP + k)*nthr]
    int* pcnt = reinterpret_cast<int*>(priv + 4 * KP * nthr) + tid;

    if (active) {
        const int Ya = first_ge(cy, sy, Hm, h), Yb = first_ge(cy + 1, sy, Hm, h);
        const int Xa = first_ge(cx, sx, Wm, w), Xb = first_ge(cx + 1, sx, Wm, w);
        const float c2 = inv_T * 1.4426950408889634f;
        const float* t00 = s_tile + (ly0 * RL_LDX + lx0) * KP;
        const float* t01 = s_tile + (ly0 * RL_LDX + lx1) * KP;
        const float* t10 = s_tile + (ly1 * RL_LDX + lx0) * KP;
        const float* t11 = s_tile + (ly1 * RL_LDX + lx1) * KP;
        const long long* lab_b = labels + (size_t)b * Hm * Wm;
        for (int Y = Ya + split; Y < Yb; Y += RS) {
            float lamy = fminf(fmaxf(sy * (float)Y - (float)cy, 0.f), 1.f);
            float A[KP], Bc[KP];
#pragma unroll
            for (int q = 0; q < KP / 4; ++q) {
                float4 v00 = reinterpret_cast<const float4*>(t00)[q], v01 = reinterpret_cast<const float4*>(t01)[q];
                float4 v10 = reinterpret_cast<const float4*>(t10)[q], v11 = reinterpret_cast<const float4*>(t11)[q];
                float l0, l1;
                l0 = fmaf(lamy, v10.x - v00.x, v00.x), l1 = fmaf(lamy, v11.x - v01.x, v01.x);
                A[4 * q + 0] = l0 * c2, Bc[4 * q + 0] = (l1 - l0) * c2;
                l0 = fmaf(lamy, v10.y - v00.y, v00.y), l1 = fmaf(lamy, v11.y - v01.y, v01.y);
                A[4 * q + 1] = l0 * c2, Bc[4 * q + 1] = (l1 - l0) * c2;
                l0 = fmaf(lamy, v10.z - v00.z, v00.z), l1 = fmaf(lamy, v11.z - v01.z, v01.z);
                A[4 * q + 2] = l0 * c2, Bc[4 * q + 2] = (l1 - l0) * c2;
                l0 = fmaf(lamy, v10.w - v00.w, v00.w), l1 = fmaf(lamy, v11.w - v01.w, v01.w);
                A[4 * q + 3] = l0 * c2, Bc[4 * q + 3] = (l1 - l0) * c2;
            }
#pragma unroll
            for (int k = 0; k < KP; ++k)
                if (k >= K) A[k] = -INFINITY, Bc[k] = 0.f;
            const long long* lrow = lab_b + (size_t)Y * Wm;
            for (int X = Xa; X < Xb; ++X) {
                const long long lv = __ldg(lrow + X);
                const int cls = map_label(lv, K);
                pcnt[cls * nthr] += 1;
                if (cls == K) {
                    if (lv != PM_IGNORE_LABEL) atomicAdd(ws + PM_WS_BAD, 1ULL);
                    continue;
                }
                const float lamx = fminf(fmaxf(sx * (float)X - (float)cx, 0.f), 1.f);
                float e[KP], mx = -INFINITY;
#pragma unroll
                for (int k = 0; k < KP; ++k) {
                    e[k] = fmaf(lamx, Bc[k], A[k]);
                    mx = fmaxf(mx, e[k]);
                }
                float sum = 0.f;
#pragma unroll
                for (int k = 0; k < KP; ++k) {
                    e[k] = exp2f(e[k] - mx);
                    sum += e[k];
                }
                // logit of the labelled class, recomputed from the tile (dynamic index)
                float a = t00[cls], bq = t01[cls], c = t10[cls], d = t11[cls];
                float l0 = fmaf(lamy, c - a, a), l1 = fmaf(lamy, d - bq, bq);
                float zy = fmaf(lamx, (l1 - l0) * c2, l0 * c2);
                loss += (mx - zy) * 0.6931471805599453f + logf(sum);
                const float inv = 1.f / sum;
                const float hy = 1.f - lamy, hx = 1.f - lamx;
                const float w00 = hy * hx, w01 = hy * lamx, w10 = lamy * hx, w11 = lamy * lamx;
                const float i00 = w00 * inv, i01 = w01 * inv, i10 = w10 * inv, i11 = w11 * inv;
#pragma unroll
                for (int k = 0; k < KP; ++k) {
                    G00[k] = fmaf(e[k], i00, G00[k]);
                    G01[k] = fmaf(e[k], i01, G01[k]);
                    G10[k] = fmaf(e[k], i10, G10[k]);
                    G11[k] = fmaf(e[k], i11, G11[k]);
                }
                pv[(0 * KP + cls) * nthr] += w00;
                pv[(1 * KP + cls) * nthr] += w01;
                pv[(2 * KP + cls) * nthr] += w10;
                pv[(3 * KP + cls) * nthr] += w11;
            }
        }
        // subtract the one-hot part; fold degenerate taps (last row / column clamp onto themselves)
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            G00[k] -= pv[(0 * KP + k) * nthr];
            G01[k] -= pv[(1 * KP + k) * nthr];
            G10[k] -= pv[(2 * KP + k) * nthr];
            G11[k] -= pv[(3 * KP + k) * nthr];
        }
        if (ly1 == ly0) {
#pragma unroll
            for (int k = 0; k < KP; ++k) G00[k] += G10[k], G01[k] += G11[k], G10[k] = 0.f, G11[k] = 0.f;
        }
        if (lx1 == lx0) {
#pragma unroll
            for (int k = 0; k < KP; ++k) G00[k] += G01[k], G10[k] += G11[k], G01[k] = 0.f, G11[k] = 0.f;
        }
    }

    // merge into the tap tile: within one (tap, split) phase every active thread owns a distinct tap
    for (int sp = 0; sp < RS; ++sp) {
#pragma unroll
        for (int tap = 0; tap < 4; ++tap) {
            const bool fold = (tap >= 2 && ly1 == ly0) || ((tap & 1) && lx1 == lx0);
            if (active && split == sp && !fold) {
                const int ly = (tap >= 2) ? ly1 : ly0, lx = (tap & 1) ? lx1 : lx0;
                float4* dst = reinterpret_cast<float4*>(ds_tile + (ly * RL_LDX + lx) * KP);
                const float* G = tap == 0 ? G00 : tap == 1 ? G01 : tap == 2 ? G10 : G11;
#pragma unroll
                for (int q = 0; q < KP / 4; ++q) {
                    float4 v = dst[q];
                    v.x += G[4 * q], v.y += G[4 * q + 1], v.z += G[4 * q + 2], v.w += G[4 * q + 3];
                    dst[q] = v;
                }
            }
            __syncthreads();
        }
    }
    // flush taps that exist (clamped duplicates were folded and stay zero)
    for (int i = tid; i < tile_elems / 4; i += nthr) {
        int e = i / (KP / 4), q = i - e * (KP / 4);
        int ty = e / RL_LDX, tx = e - ty * RL_LDX;
        int fy = fy0 + ty, fx = fx0 + tx;
        if (fy < h && fx < w) {
            float4 v = reinterpret_cast<const float4*>(ds_tile)[i];
            if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
                atomicAdd(reinterpret_cast<float4*>(ds_rl + ((size_t)(b * h + fy) * w + fx) * KP) + q, v);
        }
    }
    // label histogram + loss sum
    for (int k = 0; k <= K; ++k) {
        int c = __reduce_add_sync(0xffffffffu, pcnt[k * nthr]);
        if (lane == 0 && c != 0) atomicAdd(hist_sm + k, c);
    }
    loss = warp_sum(loss);
    if (lane == 0) red[wid] = loss;
    __syncthreads();
    if (tid == 0) {
        float tot = 0.f;
        for (int i = 0; i < (nthr >> 5); ++i) tot += red[i];
        atomicAdd(reinterpret_cast<double*>(ws + PM_WS_LOSS_SUM), (double)tot);
    }
    if (tid <= K && hist_sm[tid] != 0) atomicAdd(ws + PM_WS_HIST + tid, (unsigned long long)hist_sm[tid]);
    // last CTA: readloss = loss_sum / V ; scale = 1 / (V*T)
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned long long ticket = atomicAdd(ws + PM_WS_COUNTER, 1ULL);
        if (ticket == (unsigned long long)gridDim.x - 1) {
            __threadfence();
            unsigned long long V = 0;
            for (int k = 0; k < K; ++k) V += atomicAdd(ws + PM_WS_HIST + k, 0ULL);
            double sum = __longlong_as_double((long long)atomicAdd(ws + PM_WS_LOSS_SUM, 0ULL));
            out[0] = (float)(sum / (double)V);  // V == 0 -> 0/0 = NaN like torch
            out[1] = (float)(1.0 / ((double)V * (double)temperature));
        }
    }
}

}  // namespace pm

extern "C" int pm_score_stride(int K) { return K <= 19 ? 20 : 32; }
extern "C" int pm_colsoftmax_workspace_floats(int K) { return pm::CS_MAXG * 64; }

extern "C" int pm_colsoftmax(const float* s, const float* gumbel_q, float* score_q, float* workspace, int N, int K,
                             void* stream) {
    if (!s || !score_q || !workspace) return PM_ERR_NULL;
    if (K < 1 || K > 31) return PM_ERR_SLOTS;
    if (N <= 0) return PM_ERR_SHAPE;
    const int KP = pm_score_stride(K);
    const int tpb = K * (256 / K);
    const int rstep = tpb / K;
    int G = (N + rstep * 8 - 1) / (rstep * 8);  // >= 8 rows per thread
    if (G > pm::CS_MAXG) G = pm::CS_MAXG;
    if (G < 1) G = 1;
    const int rows_per_cta = (N + G - 1) / G;
    G = (N + rows_per_cta - 1) / rows_per_cta;
    cudaStream_t st = (cudaStream_t)stream;
    pm::colsoftmax_stats_kernel<<<G, tpb, 0, st>>>(s, gumbel_q, workspace, N, K, KP, rows_per_cta);
    PM_CHECK_LAUNCH();
    pm::colsoftmax_apply_kernel<<<G, tpb, 0, st>>>(s, gumbel_q, workspace, score_q, N, K, KP, rows_per_cta, G);
    PM_CHECK_LAUNCH();
    return 0;
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

extern "C" int pm_readloss_fwd(const float* s, const int64_t* labels, float temperature, int B, int h, int w, int Hm,
                               int Wm, int K, float* ds_rl, void* ws, float* out, void* stream) {
    if (!s || !labels || !ds_rl || !ws || !out) return PM_ERR_NULL;
    if (K < 1 || K > 31) return PM_ERR_SLOTS;
    if (B <= 0 || h <= 0 || w <= 0 || Hm <= 0 || Wm <= 0) return PM_ERR_SHAPE;
    if (((uintptr_t)s & 15) || ((uintptr_t)ds_rl & 15) || ((uintptr_t)ws & 7)) return PM_ERR_ALIGN;
    const int KP = pm_score_stride(K);
    // PyTorch's align_corners scale: (in-1)/(out-1) in fp32, 0 when out == 1
    const float sy = Hm > 1 ? (float)(h - 1) / (float)(Hm - 1) : 0.f;
    const float sx = Wm > 1 ? (float)(w - 1) / (float)(Wm - 1) : 0.f;
    // row-split factor: enough threads to fill the chip when there are few cells
    const long long cells = (long long)B * h * w;
    int rows_per_cell = h > 1 ? (Hm + h - 2) / (h - 1) : Hm;
    int RS = 1;
    while (RS < 8 && cells * RS < 148LL * 256 && RS * 2 <= rows_per_cell) RS *= 2;
    const int TYC = RS >= 4 ? 1 : 4 / RS;
    const int nthr = 32 * TYC * RS;
    const int tiles_x = (w + pm::RL_TX - 1) / pm::RL_TX, tiles_y = (h + TYC - 1) / TYC;
    const size_t smem = sizeof(float) * ((size_t)2 * (TYC + 1) * pm::RL_LDX * KP + (size_t)5 * KP * nthr + KP + 8);
    const long long grid = (long long)B * tiles_x * tiles_y;
    if (grid > 0x7fffffffLL) return PM_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (KP == 20) {
        e = cudaFuncSetAttribute(pm::readloss_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        pm::readloss_kernel<20><<<(int)grid, nthr, smem, st>>>(s, (const long long*)labels, 1.f / temperature,
                                                                temperature, h, w, Hm, Wm, K, sy, sx, RS, TYC, tiles_x,
                                                                tiles_y, ds_rl, (unsigned long long*)ws, out);
    } else {
        e = cudaFuncSetAttribute(pm::readloss_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        pm::readloss_kernel<32><<<(int)grid, nthr, smem, st>>>(s, (const long long*)labels, 1.f / temperature,
                                                                temperature, h, w, Hm, Wm, K, sy, sx, RS, TYC, tiles_x,
                                                                tiles_y, ds_rl, (unsigned long long*)ws, out);
    }
    PM_CHECK_LAUNCH();
    return 0;
}
