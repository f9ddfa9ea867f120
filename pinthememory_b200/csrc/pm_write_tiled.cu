// Memory write, pipelined variants (fast path when hw is a multiple of the 16-byte chunk).
#include <cstdlib>

#include "pm_common.cuh"
#include "pm_tma.cuh"
#include "pm_internal.h"

namespace pm {

// ----------------------------------------------------------------------------- class sums (forward)
// Same algorithm as write_reduce_kernel (thread = channel, warp-uniform class index, run-length register
// accumulator, private column of a CTA-resident [K+1][C+4] tile, one vector RED per touched class row),
// but the [C][32] feature tile arrives through a 2-stage ring of 16-byte async copies and is stored with
// XOR-swizzled chunks, so a thread reads 4 (fp32) / 8 (bf16) pixels of its channel per conflict-free
// LDS.128; the label taps of the NEXT tile are fetched while the current one is reduced.

template <typename T, int C, int KP>
__global__ void __launch_bounds__(C) write_reduce_tiled_kernel(const T* __restrict__ f, const void* __restrict__ labels, int lab_u8,
                                                                float* __restrict__ SD, int h, int w, int Hm, int Wm,
                                                                int K, float sy, float sx, int tiles_per_img,
                                                                int ntiles) {
    constexpr int NSTAGE = 2, EPC = 16 / (int)sizeof(T), CPR = 32 / EPC, NW = C / 32, CS = C + 4;
    extern __shared__ __align__(16) unsigned char smraw[];
    float* S_tile = reinterpret_cast<float*>(smraw);               // [KP][CS]
    float* pn = S_tile + KP * CS;                                   // [NW][32]
    float* invr = pn + NW * 32;                                     // [32]
    float4* ent = reinterpret_cast<float4*>(invr + 32);             // [2][32][4] (class bits, w/|f|, w, -)
    unsigned* multi = reinterpret_cast<unsigned*>(ent + 2 * 32 * 4);  // [2] + pad
    T* ft = reinterpret_cast<T*>(multi + 4);                        // [NSTAGE][C][32], swizzled chunks

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, hw = h * w;
    for (int i = tid; i < KP * CS; i += C) S_tile[i] = 0.f;
    int cur = -1;
    unsigned seen = 0u;
    float acc = 0.f, accD = 0.f;

    auto tile_coords = [&](int t, int& b, int& px0) {
        b = t / tiles_per_img;
        px0 = (t - b * tiles_per_img) * 32;
    };
    auto store_entries = [&](int buf, const LabelTaps& t, int n) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            ent[(buf * 32 + lane) * 4 + j] = make_float4(__int_as_float(t.cls[j]), 0.f, t.w[j], 0.f);
        const unsigned m = __ballot_sync(0xffffffffu, n > 1);
        if (lane == 0) multi[buf] = m;
    };
    auto taps_for = [&](int t) {
        LabelTaps r;
        int b, px0;
        tile_coords(t, b, px0);
        const int px = px0 + lane;
        if (t < ntiles && px < hw) {
            const int fy = px / w, fx = px - fy * w;
            r = label_taps(label_image(labels, (size_t)b * Hm * Wm, lab_u8), lab_u8, Hm, Wm, fy, fx, sy, sx, K);
        } else {
            r.cls[0] = r.cls[1] = r.cls[2] = r.cls[3] = K;
            r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0.f;
        }
        return r;
    };

    int tile = blockIdx.x;
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
        const int t = tile + s * gridDim.x;
        if (t < ntiles) {
            int b, px0;
            tile_coords(t, b, px0);
            tile_load_async_swz<T, C, C>(ft + s * C * 32, f + (size_t)b * C * hw, hw, px0);
        }
        cp_async_commit();
    }
    if (wid == 0) {
        LabelTaps t0 = taps_for(tile);
        const int n = compact_taps(t0);
        store_entries(0, t0, n);
    }

    int stage = 0, ebuf = 0;
    for (; tile < ntiles; tile += gridDim.x) {
        // raw taps of the next tile: loads are issued now, consumed at the end of this iteration
        LabelTaps tn;
        if (wid == 0) tn = taps_for(tile + gridDim.x);
        cp_async_wait<NSTAGE - 1>();
        __syncthreads();
        const T* xt = ft + stage * C * 32;
        int b, px0;
        tile_coords(tile, b, px0);
        const int nvalid = min(32, hw - px0);
        {  // |f|^2 per pixel: lanes = pixels, each warp sums C/NW of the C rows
            float n2 = 0.f;
            if (NW % 8 == 0) {
                // rows wid, wid+NW, ... share one swizzle phase: the lane's offset inside a row is constant
                const int pc = ((lane / EPC) ^ tile_swz<T>(wid)) * EPC + (lane % EPC);
                const T* col = xt + wid * 32 + pc;
#pragma unroll 8
                for (int i = 0; i < C / NW; ++i) {
                    const float v = to_float(col[i * NW * 32]);
                    n2 = fmaf(v, v, n2);
                }
            } else {
#pragma unroll 8
                for (int c = wid; c < C; c += NW) {
                    const int pc = ((lane / EPC) ^ tile_swz<T>(c)) * EPC + (lane % EPC);
                    const float v = to_float(xt[c * 32 + pc]);
                    n2 = fmaf(v, v, n2);
                }
            }
            pn[wid * 32 + lane] = n2;
        }
        __syncthreads();
        if (tid < 32) {
            float sacc = 0.f;
#pragma unroll
            for (int i = 0; i < NW; ++i) sacc += pn[i * 32 + tid];
            const float ir = 1.f / fmaxf(sqrtf(sacc), PM_NORM_EPS);
            float4* e4 = ent + (ebuf * 32 + tid) * 4;  // fold 1/|f| into the tap weights of this pixel
#pragma unroll
            for (int j = 0; j < 4; ++j) e4[j].y = e4[j].z * ir;
        }
        __syncthreads();
        {
            const unsigned mm = multi[ebuf];
            const float4* eb = ent + ebuf * 32 * 4;
            const int swz = tile_swz<T>(tid);
            // entries are warp-uniform; a zero-weight entry (pixel past the end of the image) adds zeros
            auto add_entry = [&](const float4 en, float val) {
                const int cls = __float_as_int(en.x);
                if (cls != cur) {
                    if (cur >= 0) {
                        S_tile[cur * CS + tid] += acc;
                        if (tid == 0) S_tile[cur * CS + C] += accD;
                    }
                    acc = 0.f;
                    accD = 0.f;
                    cur = cls;
                    seen |= 1u << cls;
                }
                acc = fmaf(en.y, val, acc);
                accD += en.z;
            };
#pragma unroll 1
            for (int q = 0; q < CPR; ++q) {
                float v[EPC];
                if (sizeof(T) == 4) {
                    const float4 r = reinterpret_cast<const float4*>(xt + tid * 32)[q ^ swz];
                    v[0] = r.x, v[1] = r.y, v[2] = r.z, v[3] = r.w;
                } else {
                    const uint4 r = reinterpret_cast<const uint4*>(xt + tid * 32)[q ^ swz];
                    const unsigned rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v[(2 * i) % EPC] = __uint_as_float(rr[i] << 16);
                        v[(2 * i + 1) % EPC] = __uint_as_float(rr[i] & 0xffff0000u);
                    }
                }
#pragma unroll
                for (int j = 0; j < EPC; ++j) {
                    const int px = q * EPC + j;
                    add_entry(eb[px * 4], v[j]);
                    if ((mm >> px) & 1u) {  // warp-uniform: this pixel straddles classes
#pragma unroll 1
                        for (int e = 1; e < 4; ++e) {
                            const float4 en = eb[px * 4 + e];
                            if (en.z != 0.f) add_entry(en, v[j]);
                        }
                    }
                }
            }
        }
        __syncthreads();  // stage and entry buffer consumed
        const int next = tile + NSTAGE * gridDim.x;
        if (next < ntiles) {
            int nb, npx0;
            tile_coords(next, nb, npx0);
            tile_load_async_swz<T, C, C>(ft + stage * C * 32, f + (size_t)nb * C * hw, hw, npx0);
        }
        cp_async_commit();
        if (wid == 0) {
            const int n = compact_taps(tn);
            store_entries(ebuf ^ 1, tn, n);
        }
        stage = (stage + 1 == NSTAGE) ? 0 : stage + 1;
        ebuf ^= 1;
    }
    cp_async_wait<0>();
    if (cur >= 0) {
        S_tile[cur * CS + tid] += acc;
        if (tid == 0) S_tile[cur * CS + C] += accD;
    }
    __syncthreads();
    for (int k = 0; k <= K; ++k) {
        if (!((seen >> k) & 1u)) continue;
        for (int i = tid; i < CS / 4; i += C)
            atomicAdd(reinterpret_cast<float4*>(SD + (size_t)k * CS) + i,
                      reinterpret_cast<const float4*>(S_tile + k * CS)[i]);
    }
}

// ---- tensor-core variant (fp32): S[class][channel] += Omega^T[class][px] . V[px][channel] -----------------------
// The class sums ARE a skinny GEMM (the reference computes them with torch.matmul, memory.py:227): per tile of
// 32 pixels, M = classes (one or two 16-row tiles), N = channels (each warp 32 = four 8-column tiles), K = pixels.
// mma.sync m16n8k8 TF32 with 3xTF32 compensation (see pm_read_tiled.cu); accumulators stay in registers across
// all tiles of the persistent CTA. The soft label weights of a tile are expanded to a dense [32 px][40] table
// in shared memory (row stride 40: conflict-free A fragments), B fragments come straight from the
// chunk-swizzled f tile (conflict-free) and are scaled by 1/|f| on the fly. Cost no longer depends on how
// fragmented the label map is (the run-length kernel above degenerates on noisy labels).

// TMA = true: the tiles arrive as SWIZZLE_128B tensor-map boxes (the same chunk ^ (row & 7) layout) on an mbarrier ring
template <int C, int KP, bool TMA>
__global__ void __launch_bounds__(C) write_reduce_mma_kernel(const __grid_constant__ CUtensorMap tm_f,
                                                              const float* __restrict__ f, const void* __restrict__ labels, int lab_u8,
                                                              float* __restrict__ SD, int h, int w, int Hm, int Wm,
                                                              int K, float sy, float sx, int tiles_per_img,
                                                              int ntiles) {
    constexpr int NSTAGE = 2, NW = C / 32, CS = C + 4, OMLD = 40;
    extern __shared__ __align__(16) unsigned char smraw_[];
    __shared__ __align__(8) uint64_t full[NSTAGE];
    unsigned char* smraw = smem_align(smraw_, 1024);  // the swizzled TMA boxes need 1 KB alignment
    float* ft = reinterpret_cast<float*>(smraw);       // [NSTAGE][C][32] swizzled; reused as S_tile [32][CS] at the end
    constexpr size_t RING = (size_t)NSTAGE * C * 32 > (size_t)32 * CS ? (size_t)NSTAGE * C * 32 : (size_t)32 * CS;
    float* om = ft + RING;                             // [32 px][OMLD] dense label weights of the current tile
    float* pn = om + 32 * OMLD;                        // [NW][32]
    float* invr = pn + NW * 32;                        // [32]
    unsigned* cmask = reinterpret_cast<unsigned*>(invr + 32);  // [4]

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, hw = h * w;
    const int g = lane >> 2, t = lane & 3;
    float acc[2][4][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f;
    float Dacc = 0.f;  // lane = class, partial over this warp's pixels
    unsigned seen = 0u;

    auto tile_coords = [&](int tl, int& b, int& px0) {
        b = tl / tiles_per_img;
        px0 = (tl - b * tiles_per_img) * 32;
    };
    auto taps_for = [&](int tl) {
        LabelTaps r;
        int b, px0;
        tile_coords(tl, b, px0);
        const int px = px0 + lane;
        if (tl < ntiles && px < hw) {
            const int fy = px / w, fx = px - fy * w;
            r = label_taps(label_image(labels, (size_t)b * Hm * Wm, lab_u8), lab_u8, Hm, Wm, fy, fx, sy, sx, K);
        } else {
            r.cls[0] = r.cls[1] = r.cls[2] = r.cls[3] = K;
            r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0.f;
        }
        return r;
    };

    int tile = blockIdx.x;
    auto issue = [&](int tl, int s) {  // TMA: one thread
        int b, px0;
        tile_coords(tl, b, px0);
        mbar_expect_tx(&full[s], (uint32_t)(C * 32 * sizeof(float)));
        tma_load_2d(ft + s * C * 32, &tm_f, px0, b * C, &full[s]);
    };
    if constexpr (TMA) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
            for (int s = 0; s < NSTAGE; ++s)
                if (tile + s * (int)gridDim.x < ntiles) issue(tile + s * gridDim.x, s);
        }
    } else {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) {
            const int tl = tile + s * gridDim.x;
            if (tl < ntiles) {
                int b, px0;
                tile_coords(tl, b, px0);
                tile_load_async_swz<float, C, C>(ft + s * C * 32, f + (size_t)b * C * hw, hw, px0);
            }
            cp_async_commit();
        }
    }
    LabelTaps tcur;
    if (wid == 0) tcur = taps_for(tile);

    int stage = 0;
    unsigned phase = 0;
    for (; tile < ntiles; tile += gridDim.x) {
        LabelTaps tnext;
        if (wid == 0) tnext = taps_for(tile + gridDim.x);  // loads in flight across this tile
        if constexpr (!TMA) cp_async_wait<NSTAGE - 1>();
        __syncthreads();
        if constexpr (TMA) mbar_wait(&full[stage], phase);
        const float* xt = ft + stage * C * 32;
        for (int i = tid; i < 32 * OMLD / 4; i += C) reinterpret_cast<float4*>(om)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        {  // |f|^2 per pixel: lanes = pixels, each warp sums 32 of the C rows (rows wid, wid+NW, .. share a swizzle
           // phase when NW % 8 == 0)
            float n2 = 0.f;
            if (NW % 8 == 0) {
                const float* col = xt + wid * 32 + (((lane >> 2) ^ (wid & 7)) << 2) + (lane & 3);
#pragma unroll 8
                for (int i = 0; i < C / NW; ++i) {
                    const float v = col[i * NW * 32];
                    n2 = fmaf(v, v, n2);
                }
            } else {
                for (int c = wid; c < C; c += NW) {
                    const float v = xt[c * 32 + (((lane >> 2) ^ (c & 7)) << 2) + (lane & 3)];
                    n2 = fmaf(v, v, n2);
                }
            }
            pn[wid * 32 + lane] = n2;
        }
        __syncthreads();
        if (wid == 0) {
            float sacc = 0.f;
#pragma unroll
            for (int i = 0; i < NW; ++i) sacc += pn[i * 32 + lane];
            invr[lane] = 1.f / fmaxf(sqrtf(sacc), PM_NORM_EPS);
            unsigned m = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (tcur.w[j] != 0.f) {  // merged taps: distinct classes per pixel
                    om[lane * OMLD + tcur.cls[j]] = tcur.w[j];
                    m |= 1u << tcur.cls[j];
                }
            m = __reduce_or_sync(0xffffffffu, m);
            if (lane == 0) cmask[0] = m;
        }
        __syncthreads();
        const unsigned cm = cmask[0];
        seen |= cm;
        const bool hi_tile = (cm >> 16) != 0u;  // classes 16.. present in this tile (warp-uniform)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const float ir0 = invr[8 * ks + t], ir1 = invr[8 * ks + t + 4];
            unsigned ah[2][4], al[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                if (m == 0 || hi_tile) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float v = om[(8 * ks + t + 4 * (e >> 1)) * OMLD + 16 * m + g + 8 * (e & 1)];
                        ah[m][e] = tf32_hi(v);
                        al[m][e] = tf32_lo(v);
                    }
                }
            }
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const float* row = xt + (wid * 32 + 8 * n + g) * 32 + t;
                const float b0 = row[((2 * ks) ^ g) << 2] * ir0, b1 = row[((2 * ks + 1) ^ g) << 2] * ir1;
                const unsigned b0h = tf32_hi(b0), b1h = tf32_hi(b1), b0l = tf32_lo(b0), b1l = tf32_lo(b1);
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    if (m == 0 || hi_tile) {
                        mma_tf32(acc[m][n], al[m][0], al[m][1], al[m][2], al[m][3], b0h, b1h);
                        mma_tf32(acc[m][n], ah[m][0], ah[m][1], ah[m][2], ah[m][3], b0l, b1l);
                        mma_tf32(acc[m][n], ah[m][0], ah[m][1], ah[m][2], ah[m][3], b0h, b1h);
                    }
                }
            }
        }
        {  // soft counts: lane = class, every warp sums its own 32/NW pixels (combined at the end)
#pragma unroll
            for (int px = wid; px < 32; px += NW) Dacc += om[px * OMLD + lane];
        }
        __syncthreads();  // stage and label table consumed
        const int next = tile + NSTAGE * gridDim.x;
        if constexpr (TMA) {
            if (tid == 0 && next < ntiles) issue(next, stage);
        } else {
            if (next < ntiles) {
                int nb, npx0;
                tile_coords(next, nb, npx0);
                tile_load_async_swz<float, C, C>(ft + stage * C * 32, f + (size_t)nb * C * hw, hw, npx0);
            }
            cp_async_commit();
        }
        if (wid == 0) tcur = tnext;
        if (++stage == NSTAGE) stage = 0, phase ^= 1u;
    }
    if constexpr (!TMA) cp_async_wait<0>();
    __syncthreads();
    // accumulators -> shared [32][CS] (over the tile ring), then one vector RED per touched class row
    float* S_tile = ft;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const int ch = wid * 32 + 8 * n + 2 * t;
            *reinterpret_cast<float2*>(S_tile + (16 * m + g) * CS + ch) = make_float2(acc[m][n][0], acc[m][n][1]);
            *reinterpret_cast<float2*>(S_tile + (16 * m + g + 8) * CS + ch) = make_float2(acc[m][n][2], acc[m][n][3]);
        }
    pn[wid * 32 + lane] = Dacc;  // pn is free now
    __syncthreads();
    if (wid == 0) {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < NW; ++i) d += pn[i * 32 + lane];
        S_tile[lane * CS + C] = d;
        S_tile[lane * CS + C + 1] = S_tile[lane * CS + C + 2] = S_tile[lane * CS + C + 3] = 0.f;
    }
    __syncthreads();
    for (int k = 0; k <= K; ++k) {
        if (!((seen >> k) & 1u)) continue;
        for (int i = tid; i < CS / 4; i += C)
            atomicAdd(reinterpret_cast<float4*>(SD + (size_t)k * CS) + i,
                      reinterpret_cast<const float4*>(S_tile + k * CS)[i]);
    }
}

// ---- tensor-core variant for bf16 maps: mma.sync m16n8k16, fp32 accumulate ---------------------------------------
// Same GEMM. The f tile ([C][32 px] bf16, TMA SWIZZLE_64B) IS the B operand: a 32-bit word of a channel row holds the
// pixel pair (2t, 2t+1) an m16n8k16 B fragment wants, and with the 64-byte swizzle the eight rows of a fragment load
// fall into distinct bank groups. 1/|f| scales the A operand (the label weights of the pixel) instead of the features,
// so the bf16 feature values enter the product unrounded; A is split hi + lo (two MMAs).
template <int C, int KP, bool TMA>
__global__ void __launch_bounds__(C) write_reduce_mma16_kernel(const __grid_constant__ CUtensorMap tm_f,
                                                                const __nv_bfloat16* __restrict__ f,
                                                                const void* __restrict__ labels, int lab_u8,
                                                                float* __restrict__ SD, int h, int w, int Hm, int Wm, int K,
                                                                float sy, float sx, int tiles_per_img, int ntiles) {
    using T = __nv_bfloat16;
    constexpr int NSTAGE = 2, NW = C / 32, CS = C + 4, OMLD = 40;
    extern __shared__ __align__(16) unsigned char smraw_[];
    __shared__ __align__(8) uint64_t full[NSTAGE];
    unsigned char* smraw = smem_align(smraw_, 1024);
    T* ft = reinterpret_cast<T*>(smraw);  // [NSTAGE][C][32] swizzled; the region is reused as S_tile [32][CS] fp32 at the end
    constexpr size_t RING_B = sizeof(T) * NSTAGE * C * 32 > sizeof(float) * 32 * CS ? sizeof(T) * NSTAGE * C * 32 : sizeof(float) * 32 * CS;
    float* om = reinterpret_cast<float*>(smraw + RING_B);  // [32 px][OMLD] dense label weights of the current tile
    float* pn = om + 32 * OMLD;                            // [NW][32]
    float* invr = pn + NW * 32;                            // [32]
    unsigned* cmask = reinterpret_cast<unsigned*>(invr + 32);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, hw = h * w;
    const int g = lane >> 2, t = lane & 3;
    float acc[2][4][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f;
    float Dacc = 0.f;
    unsigned seen = 0u;

    auto tile_coords = [&](int tl, int& b, int& px0) {
        b = tl / tiles_per_img;
        px0 = (tl - b * tiles_per_img) * 32;
    };
    auto taps_for = [&](int tl) {
        LabelTaps r;
        int b, px0;
        tile_coords(tl, b, px0);
        const int px = px0 + lane;
        if (tl < ntiles && px < hw) {
            const int fy = px / w, fx = px - fy * w;
            r = label_taps(label_image(labels, (size_t)b * Hm * Wm, lab_u8), lab_u8, Hm, Wm, fy, fx, sy, sx, K);
        } else {
            r.cls[0] = r.cls[1] = r.cls[2] = r.cls[3] = K;
            r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0.f;
        }
        return r;
    };
    auto load_tile = [&](int tl, int s) {  // cp.async ring: all threads, the SWIZZLE_64B pattern by hand
        int b, px0;
        tile_coords(tl, b, px0);
        const T* base = f + (size_t)b * C * hw;
        const int ch = tid % 4, px = px0 + ch * 8;
        const bool valid = px < hw;
        for (int row = tid / 4; row < C; row += C / 4)
            cp_async16(ft + (size_t)s * C * 32 + row * 32 + ((ch ^ ((row >> 1) & 3)) << 3), base + (size_t)row * hw + (valid ? px : 0),
                       valid);
    };
    int tile = blockIdx.x;
    auto issue = [&](int tl, int s) {  // TMA: one thread
        int b, px0;
        tile_coords(tl, b, px0);
        mbar_expect_tx(&full[s], (uint32_t)(C * 32 * sizeof(T)));
        tma_load_2d(ft + (size_t)s * C * 32, &tm_f, px0, b * C, &full[s]);
    };
    if constexpr (TMA) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
            for (int s = 0; s < NSTAGE; ++s)
                if (tile + s * (int)gridDim.x < ntiles) issue(tile + s * gridDim.x, s);
        }
    } else {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) {
            if (tile + s * (int)gridDim.x < ntiles) load_tile(tile + s * gridDim.x, s);
            cp_async_commit();
        }
    }
    LabelTaps tcur;
    if (wid == 0) tcur = taps_for(tile);

    int stage = 0;
    unsigned phase = 0;
    for (; tile < ntiles; tile += gridDim.x) {
        LabelTaps tnext;
        if (wid == 0) tnext = taps_for(tile + gridDim.x);
        if constexpr (!TMA) cp_async_wait<NSTAGE - 1>();
        __syncthreads();
        if constexpr (TMA) mbar_wait(&full[stage], phase);
        const T* xt = ft + (size_t)stage * C * 32;
        const unsigned* xw = reinterpret_cast<const unsigned*>(xt);
        for (int i = tid; i < 32 * OMLD / 4; i += C) reinterpret_cast<float4*>(om)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        {  // |f|^2 per pixel: lane l sums pixel pair (2(l&15), +1) over the rows wid*32 + 2i + (l>>4); halves combined below
            const int half = lane >> 4, pp = (lane & 15) * 2;
            float na = 0.f, nb = 0.f;
#pragma unroll 8
            for (int i = 0; i < 16; ++i) {
                const int row = wid * 32 + 2 * i + half;
                const unsigned v = xw[(row * 32 + ((((pp >> 3) ^ ((row >> 1) & 3)) << 3) | (pp & 7))) >> 1];
                const float lo = __uint_as_float(v << 16), hi = __uint_as_float(v & 0xffff0000u);
                na = fmaf(lo, lo, na), nb = fmaf(hi, hi, nb);
            }
            na += __shfl_xor_sync(0xffffffffu, na, 16);
            nb += __shfl_xor_sync(0xffffffffu, nb, 16);
            if (half == 0) pn[wid * 32 + pp] = na, pn[wid * 32 + pp + 1] = nb;
        }
        __syncthreads();
        if (wid == 0) {
            float sacc = 0.f;
#pragma unroll
            for (int i = 0; i < NW; ++i) sacc += pn[i * 32 + lane];
            invr[lane] = 1.f / fmaxf(sqrtf(sacc), PM_NORM_EPS);
            unsigned m = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (tcur.w[j] != 0.f) {
                    om[lane * OMLD + tcur.cls[j]] = tcur.w[j];
                    m |= 1u << tcur.cls[j];
                }
            m = __reduce_or_sync(0xffffffffu, m);
            if (lane == 0) cmask[0] = m;
        }
        __syncthreads();
        const unsigned cm = cmask[0];
        seen |= cm;
        const bool hi_tile = (cm >> 16) != 0u;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {  // 16 pixels per step
            // A[class][px]: a0 = (class g, px 2t,2t+1), a1 = (class g+8, ..), a2 = (class g, px 2t+8,..), a3 = (class g+8, ..)
            unsigned ah[2][4], al[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                if (m == 0 || hi_tile) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int px = 16 * ks + 2 * t + 8 * (e >> 1), cls = 16 * m + g + 8 * (e & 1);
                        const float v0 = om[px * OMLD + cls] * invr[px], v1 = om[(px + 1) * OMLD + cls] * invr[px + 1];
                        const float h0 = __bfloat162float(__float2bfloat16_rn(v0)), h1 = __bfloat162float(__float2bfloat16_rn(v1));
                        __nv_bfloat162 hh = __floats2bfloat162_rn(h0, h1), ll = __floats2bfloat162_rn(v0 - h0, v1 - h1);
                        ah[m][e] = *reinterpret_cast<unsigned*>(&hh);
                        al[m][e] = *reinterpret_cast<unsigned*>(&ll);
                    }
                }
            }
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const int row = wid * 32 + 8 * n + g, sw = (row >> 1) & 3;
                const unsigned b0 = xw[row * 16 + (((2 * ks) ^ sw) << 2) + t], b1 = xw[row * 16 + (((2 * ks + 1) ^ sw) << 2) + t];
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    if (m == 0 || hi_tile) {
                        asm volatile(
                            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                            : "+f"(acc[m][n][0]), "+f"(acc[m][n][1]), "+f"(acc[m][n][2]), "+f"(acc[m][n][3])
                            : "r"(al[m][0]), "r"(al[m][1]), "r"(al[m][2]), "r"(al[m][3]), "r"(b0), "r"(b1));
                        asm volatile(
                            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                            : "+f"(acc[m][n][0]), "+f"(acc[m][n][1]), "+f"(acc[m][n][2]), "+f"(acc[m][n][3])
                            : "r"(ah[m][0]), "r"(ah[m][1]), "r"(ah[m][2]), "r"(ah[m][3]), "r"(b0), "r"(b1));
                    }
                }
            }
        }
        {
#pragma unroll
            for (int px = wid; px < 32; px += NW) Dacc += om[px * OMLD + lane];
        }
        __syncthreads();
        const int next = tile + NSTAGE * gridDim.x;
        if constexpr (TMA) {
            if (tid == 0 && next < ntiles) issue(next, stage);
        } else {
            if (next < ntiles) load_tile(next, stage);
            cp_async_commit();
        }
        if (wid == 0) tcur = tnext;
        if (++stage == NSTAGE) stage = 0, phase ^= 1u;
    }
    if constexpr (!TMA) cp_async_wait<0>();
    __syncthreads();
    float* S_tile = reinterpret_cast<float*>(smraw);
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const int ch = wid * 32 + 8 * n + 2 * t;
            *reinterpret_cast<float2*>(S_tile + (16 * m + g) * CS + ch) = make_float2(acc[m][n][0], acc[m][n][1]);
            *reinterpret_cast<float2*>(S_tile + (16 * m + g + 8) * CS + ch) = make_float2(acc[m][n][2], acc[m][n][3]);
        }
    pn[wid * 32 + lane] = Dacc;
    __syncthreads();
    if (wid == 0) {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < NW; ++i) d += pn[i * 32 + lane];
        S_tile[lane * CS + C] = d;
        S_tile[lane * CS + C + 1] = S_tile[lane * CS + C + 2] = S_tile[lane * CS + C + 3] = 0.f;
    }
    __syncthreads();
    for (int k = 0; k <= K; ++k) {
        if (!((seen >> k) & 1u)) continue;
        for (int i = tid; i < CS / 4; i += C)
            atomicAdd(reinterpret_cast<float4*>(SD + (size_t)k * CS) + i, reinterpret_cast<const float4*>(S_tile + k * CS)[i]);
    }
}

template <int C, int KP>
int launch_write_reduce_mma16(const void* f, const void* labels, int lab_u8, float* SD, int B, int h, int w, int Hm, int Wm,
                              int K, cudaStream_t st) {
    using T = __nv_bfloat16;
    const size_t ring = sizeof(T) * (size_t)2 * C * 32, stile = sizeof(float) * (size_t)32 * (C + 4);
    const size_t smem = sizeof(float) * (32 * 40 + (C / 32) * 32 + 32 + 4) + (ring > stile ? ring : stile) + 1024;
    const int hw = h * w, tiles = (hw + 31) / 32, ntiles = B * tiles;
    const float sy = h > 1 ? (float)(Hm - 1) / (float)(h - 1) : 0.f;
    const float sx = w > 1 ? (float)(Wm - 1) / (float)(w - 1) : 0.f;
    CUtensorMap tm_f;
    const bool tma = tma_enabled() && make_map_2d<T>(&tm_f, f, (size_t)B * C, hw, C, 32, 3);
    auto kern = tma ? write_reduce_mma16_kernel<C, KP, true> : write_reduce_mma16_kernel<C, KP, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int grid = 148 * per_sm;
    if (grid > ntiles) grid = ntiles;
    kern<<<grid, C, smem, st>>>(tm_f, (const T*)f, labels, lab_u8, SD, h, w, Hm, Wm, K, sy, sx, tiles, ntiles);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

template <int C, int KP>
int launch_write_reduce_mma(const void* f, const void* labels, int lab_u8, float* SD, int B, int h, int w, int Hm, int Wm, int K,
                            cudaStream_t st) {
    const size_t ring = sizeof(float) * (size_t)2 * C * 32, stile = sizeof(float) * (size_t)32 * (C + 4);
    const size_t smem = sizeof(float) * (32 * 40 + (C / 32) * 32 + 32 + 4) + (ring > stile ? ring : stile) + 1024;
    const int hw = h * w, tiles = (hw + 31) / 32, ntiles = B * tiles;
    const float sy = h > 1 ? (float)(Hm - 1) / (float)(h - 1) : 0.f;
    const float sx = w > 1 ? (float)(Wm - 1) / (float)(w - 1) : 0.f;
    CUtensorMap tm_f;
    const bool tma = tma_enabled() && make_map_2d<float>(&tm_f, f, (size_t)B * C, hw, C, 32, true);
    auto kern = tma ? write_reduce_mma_kernel<C, KP, true> : write_reduce_mma_kernel<C, KP, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int grid = 148 * per_sm;
    if (grid > ntiles) grid = ntiles;
    kern<<<grid, C, smem, st>>>(tm_f, (const float*)f, labels, lab_u8, SD, h, w, Hm, Wm, K, sy, sx, tiles, ntiles);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

template <typename T, int C, int KP>
int launch_write_reduce_tiled(const void* f, const void* labels, int lab_u8, float* SD, int B, int h, int w, int Hm, int Wm,
                              int K, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)KP * (C + 4) + (C / 32) * 32 + 32) + sizeof(float4) * 2 * 32 * 4 + 16 +
                        sizeof(T) * (size_t)2 * C * 32;
    auto kern = write_reduce_tiled_kernel<T, C, KP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int hw = h * w, tiles = (hw + 31) / 32, ntiles = B * tiles;
    int per_sm = (int)(220 * 1024 / (smem + 1024));
    if (per_sm > 2048 / C) per_sm = 2048 / C;
    if (per_sm < 1) per_sm = 1;
    int grid = 148 * per_sm;
    if (grid > ntiles) grid = ntiles;
    const float sy = h > 1 ? (float)(Hm - 1) / (float)(h - 1) : 0.f;
    const float sx = w > 1 ? (float)(Wm - 1) / (float)(w - 1) : 0.f;
    kern<<<grid, C, smem, st>>>((const T*)f, labels, lab_u8, SD, h, w, Hm, Wm, K, sy, sx, tiles, ntiles);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

// ---------------------------------------------------------------------------- write backward (to f)
// dv[n] = sum_taps w * dS[class]; df = (dv - v (v.dv)) / |f|, v = f/|f|. Persistent CTAs, 2 per SM (measured:
// one 16-warp CTA with a 5-stage ring is 45% slower -- the per-tile barriers need a second CTA to hide
// behind, bytes in flight are not the limit); the [C][32] tiles of f arrive through a 2-stage async-copy
// ring, lanes = pixels, 8 warps split the channels,
// dS sits in shared memory with row stride C+1 (lanes with different classes hit different banks). The
// label taps of the next tile are fetched by warp 0 while the current tile is processed.

constexpr int WBT_THREADS = 256, WBT_WARPS = 8, WBT_STAGES = 2;

// TMA = true: the f tiles arrive as 2-D tensor-map boxes completing on one mbarrier per stage (one thread issues a
// stage) instead of 8 LDGSTS per thread; `tm_f` is unused otherwise.
template <typename T, int C, int KP, bool TMA>
__global__ void __launch_bounds__(WBT_THREADS, 2)
    write_bwd_tiled_kernel(const __grid_constant__ CUtensorMap tm_f, const float* __restrict__ dS,
                           const T* __restrict__ f, const void* __restrict__ labels, int lab_u8,
                           T* __restrict__ df, int h, int w, int Hm, int Wm, int K, float sy, float sx,
                           int tiles_per_img, int ntiles) {
    constexpr int NSTAGE = WBT_STAGES, CW = C / WBT_WARPS, LDS_ = C + 1;
    extern __shared__ __align__(16) unsigned char smraw[];
    float* dSs = reinterpret_cast<float*>(smraw);      // [KP][C+1]; rows >= K are zero (ignore class)
    float* pn = dSs + KP * LDS_;                        // [8][32]
    float* pd = pn + WBT_WARPS * 32;                    // [8][32]
    float2* ent = reinterpret_cast<float2*>(pd + WBT_WARPS * 32 + ((KP * LDS_) & 1));  // [2][32][4]
    T* ft = reinterpret_cast<T*>(smem_align(reinterpret_cast<unsigned char*>(ent + 2 * 32 * 4), 128));  // [NSTAGE][C][32]
    __shared__ __align__(8) uint64_t full[NSTAGE];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, hw = h * w;
    auto tile_coords = [&](int t, int& b, int& px0) {
        b = t / tiles_per_img;
        px0 = (t - b * tiles_per_img) * 32;
    };
    auto taps_for = [&](int t) {
        LabelTaps r;
        int b, px0;
        tile_coords(t, b, px0);
        const int px = px0 + lane;
        if (t < ntiles && px < hw) {
            const int fy = px / w, fx = px - fy * w;
            r = label_taps(label_image(labels, (size_t)b * Hm * Wm, lab_u8), lab_u8, Hm, Wm, fy, fx, sy, sx, K);
        } else {
            r.cls[0] = r.cls[1] = r.cls[2] = r.cls[3] = K;
            r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0.f;
        }
        return r;
    };
    auto store_entries = [&](int buf, const LabelTaps& t) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ent[(buf * 32 + lane) * 4 + j] = make_float2(__int_as_float(t.cls[j]), t.w[j]);
    };

    int tile = blockIdx.x;
    auto issue = [&](int t, int s) {  // TMA: one thread
        int b, px0;
        tile_coords(t, b, px0);
        mbar_expect_tx(&full[s], (uint32_t)(C * 32 * sizeof(T)));
        tma_load_2d(ft + s * C * 32, &tm_f, px0, b * C, &full[s]);
    };
    if constexpr (TMA) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
            for (int s = 0; s < NSTAGE; ++s)
                if (tile + s * (int)gridDim.x < ntiles) issue(tile + s * gridDim.x, s);
        }
    } else {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) {
            const int t = tile + s * gridDim.x;
            if (t < ntiles) {
                int b, px0;
                tile_coords(t, b, px0);
                tile_load_async<T, C, WBT_THREADS>(ft + s * C * 32, f + (size_t)b * C * hw, hw, px0);
            }
            cp_async_commit();
        }
    }
    for (int i = tid; i < KP * LDS_; i += WBT_THREADS) {
        const int k = i / LDS_, c = i - k * LDS_;
        dSs[i] = (k < K && c < C) ? __ldg(dS + (size_t)k * C + c) : 0.f;
    }
    if (wid == 0) {
        LabelTaps t0 = taps_for(tile);
        compact_taps(t0);
        store_entries(0, t0);
    }

    int stage = 0, ebuf = 0;
    unsigned phase = 0;
    for (; tile < ntiles; tile += gridDim.x) {
        LabelTaps tn;
        if (wid == 0) tn = taps_for(tile + gridDim.x);  // loads in flight across this tile
        if constexpr (!TMA) cp_async_wait<NSTAGE - 1>();
        __syncthreads();  // dS / tap entries of this tile visible (and, with cp.async, the tile itself)
        if constexpr (TMA) mbar_wait(&full[stage], phase);
        const T* xt = ft + stage * C * 32;
        int b, px0;
        tile_coords(tile, b, px0);
        const int nvalid = min(32, hw - px0);
        // this pixel's taps
        int cls[4];
        float wt[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 e = ent[(ebuf * 32 + lane) * 4 + j];
            cls[j] = __float_as_int(e.x);
            wt[j] = e.y;
        }
        const bool multi = wt[1] != 0.f;
        const T* fcol = xt + wid * CW * 32 + lane;
        const float* r0 = dSs + cls[0] * LDS_ + wid * CW;
        float dv[CW], n2 = 0.f, dotf = 0.f;
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            const float fv = to_float(fcol[j * 32]);
            n2 = fmaf(fv, fv, n2);
            dv[j] = wt[0] * r0[j];
        }
        if (multi) {  // pixel straddles classes (rare for real label maps)
            const float* r1 = dSs + cls[1] * LDS_ + wid * CW;
            const float* r2 = dSs + cls[2] * LDS_ + wid * CW;
            const float* r3 = dSs + cls[3] * LDS_ + wid * CW;
#pragma unroll
            for (int j = 0; j < CW; ++j) dv[j] = fmaf(wt[1], r1[j], fmaf(wt[2], r2[j], fmaf(wt[3], r3[j], dv[j])));
        }
#pragma unroll
        for (int j = 0; j < CW; ++j) dotf = fmaf(to_float(fcol[j * 32]), dv[j], dotf);
        pn[wid * 32 + lane] = n2;
        pd[wid * 32 + lane] = dotf;
        __syncthreads();
        float nn = 0.f, dd = 0.f;
#pragma unroll
        for (int i = 0; i < WBT_WARPS; ++i) {
            nn += pn[i * 32 + lane];
            dd += pd[i * 32 + lane];
        }
        const float nrm = sqrtf(nn), ir = 1.f / fmaxf(nrm, PM_NORM_EPS);
        // v.dv = ir * (f.dv); df = (dv - v (v.dv)) * ir = dv*ir - f * (ir^3 * f.dv)
        const float coef = (nrm <= PM_NORM_EPS) ? 0.f : ir * ir * ir * dd;
        if (lane < nvalid) {
            T* dfp = df + ((size_t)b * C + wid * CW) * hw + px0 + lane;
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                stf(dfp, fmaf(-coef, to_float(fcol[j * 32]), dv[j] * ir));
                dfp += hw;
            }
        }
        __syncthreads();  // stage, entries and pn/pd consumed
        const int next = tile + NSTAGE * gridDim.x;
        if constexpr (TMA) {
            if (tid == 0 && next < ntiles) issue(next, stage);
        } else {
            if (next < ntiles) {
                int nb, npx0;
                tile_coords(next, nb, npx0);
                tile_load_async<T, C, WBT_THREADS>(ft + stage * C * 32, f + (size_t)nb * C * hw, hw, npx0);
            }
            cp_async_commit();
        }
        if (wid == 0) {
            compact_taps(tn);
            store_entries(ebuf ^ 1, tn);
        }
        if (++stage == NSTAGE) stage = 0, phase ^= 1u;
        ebuf ^= 1;
    }
    if constexpr (!TMA) cp_async_wait<0>();
}

template <typename T, int C, int KP>
int launch_write_bwd_tiled(const float* dS, const void* f, const void* labels, int lab_u8, void* df, int B, int h, int w, int Hm,
                           int Wm, int K, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)KP * (C + 1) + 2 * WBT_WARPS * 32 + 1) + sizeof(float2) * 2 * 32 * 4 +
                        sizeof(T) * (size_t)WBT_STAGES * C * 32 + 128;
    const int hw = h * w, tiles = (hw + 31) / 32, ntiles = B * tiles;
    int grid = 2 * 148;
    if (grid > ntiles) grid = ntiles;
    const float sy = h > 1 ? (float)(Hm - 1) / (float)(h - 1) : 0.f;
    const float sx = w > 1 ? (float)(Wm - 1) / (float)(w - 1) : 0.f;
    CUtensorMap tm_f;
    cudaError_t e;
    if (tma_enabled() && make_map_2d<T>(&tm_f, f, (size_t)B * C, hw, C, 32, false)) {
        auto kern = write_bwd_tiled_kernel<T, C, KP, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<grid, WBT_THREADS, smem, st>>>(tm_f, dS, (const T*)f, labels, lab_u8, (T*)df, h, w, Hm, Wm, K, sy,
                                              sx, tiles, ntiles);
    } else {
        auto kern = write_bwd_tiled_kernel<T, C, KP, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<grid, WBT_THREADS, smem, st>>>(tm_f, dS, (const T*)f, labels, lab_u8, (T*)df, h, w, Hm, Wm, K, sy,
                                              sx, tiles, ntiles);
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

#define PM_WT_SWITCH_C(T, KP, FN, ...)                \
    switch (C) {                                      \
        case 32: return FN<T, 32, KP>(__VA_ARGS__);   \
        case 64: return FN<T, 64, KP>(__VA_ARGS__);   \
        case 128: return FN<T, 128, KP>(__VA_ARGS__); \
        case 256: return FN<T, 256, KP>(__VA_ARGS__); \
        default: return PM_ERR_CHANNELS;              \
    }

int write_reduce_tiled(const void* f, const void* labels, int lab_u8, float* SD, int B, int C, int h, int w, int Hm, int Wm,
                       int K, int dtype, cudaStream_t st) {
    if (dtype == PM_F32) {
        switch (C) {  // tensor-core variant for every fp32 case (K + 1 <= 32 classes fit two 16-row MMA tiles)
            case 32: return launch_write_reduce_mma<32, 32>(f, labels, lab_u8, SD, B, h, w, Hm, Wm, K, st);
            case 64: return launch_write_reduce_mma<64, 32>(f, labels, lab_u8, SD, B, h, w, Hm, Wm, K, st);
            case 128: return launch_write_reduce_mma<128, 32>(f, labels, lab_u8, SD, B, h, w, Hm, Wm, K, st);
            case 256: return launch_write_reduce_mma<256, 32>(f, labels, lab_u8, SD, B, h, w, Hm, Wm, K, st);
            default: return PM_ERR_CHANNELS;
        }
    } else {
        static const bool mma_off = getenv("PM_BF16_MMA_OFF") != nullptr;  // A/B switch: the run-length FFMA2 kernel
        if (!mma_off) {
            switch (C) {  // tensor-core variant (whole 32-channel warps, >= 64 threads for the tile loads)
                case 64: return launch_write_reduce_mma16<64, 32>(f, labels, lab_u8, SD, B, h, w, Hm, Wm, K, st);
                case 128: return launch_write_reduce_mma16<128, 32>(f, labels, lab_u8, SD, B, h, w, Hm, Wm, K, st);
                case 256: return launch_write_reduce_mma16<256, 32>(f, labels, lab_u8, SD, B, h, w, Hm, Wm, K, st);
                default: break;
            }
        }
        if (K <= 19) { PM_WT_SWITCH_C(__nv_bfloat16, 20, launch_write_reduce_tiled, f, labels, lab_u8, SD, B, h, w, Hm, Wm, K, st) }
        else { PM_WT_SWITCH_C(__nv_bfloat16, 32, launch_write_reduce_tiled, f, labels, lab_u8, SD, B, h, w, Hm, Wm, K, st) }
    }
}

int write_bwd_tiled(const float* dS, const void* f, const void* labels, int lab_u8, void* df, int B, int C, int h, int w, int Hm,
                    int Wm, int K, int dtype, cudaStream_t st) {
    if (dtype == PM_F32) {
        if (K <= 19) { PM_WT_SWITCH_C(float, 20, launch_write_bwd_tiled, dS, f, labels, lab_u8, df, B, h, w, Hm, Wm, K, st) }
        else { PM_WT_SWITCH_C(float, 32, launch_write_bwd_tiled, dS, f, labels, lab_u8, df, B, h, w, Hm, Wm, K, st) }
    } else {
        if (K <= 19) { PM_WT_SWITCH_C(__nv_bfloat16, 20, launch_write_bwd_tiled, dS, f, labels, lab_u8, df, B, h, w, Hm, Wm, K, st) }
        else { PM_WT_SWITCH_C(__nv_bfloat16, 32, launch_write_bwd_tiled, dS, f, labels, lab_u8, df, B, h, w, Hm, Wm, K, st) }
    }
}

}  // namespace pm
