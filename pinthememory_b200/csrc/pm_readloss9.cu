// Feature-cohesion read loss, third-generation kernel: one thread per LABEL ROW of a cell.
//
// Same mathematics and the same cell decomposition as pm_readloss8.cu (a "cell" = the label pixels whose four bilinear taps
// are the feature pixels (cy,cx)..(cy+1,cx+1); inside a cell row exp2(z_k) is a geometric sequence in x). What changed is
// who does what. The second kernel gave a thread a whole cell (or half of one): 4 x 20 tap accumulators + 4 x 20 row
// state = 248 registers -> 8 warps per SM, and ~3 200 instructions of per-thread prologue / merge / one-hot epilogue
// against ~3 900 in the label loop (ncu: 32.6 M warp-instructions, issue slots 35 % busy). Here
//   * a warp = 32 neighbouring cells of ONE label row Y; the 8 warps of a CTA take the label rows of a band of CR cell rows
//     round-robin. A thread only carries the row state (E, ratio, the two x-weighted row sums: 80 registers) and finishes
//     by parking its 2 x 20 row sums in shared memory (10 x STS.128);
//   * the y-weighting into the four taps happens afterwards, once per CTA, by 8 threads per cell reading those records:
//     1 LDS.128 + 8 FMA per label row and thread, no barriers between cell rows (a thread owns its (cell, slot quad) in
//     the left / right tap tiles for the whole band);
//   * the one-hot term (a function of the labels alone) is accumulated in a fixed-point uint32 tile with native shared
//     integer atomics -- fp32 shared atomics are a CAS loop on sm_100 -- run-length compressed along the row. It is
//     deterministic, replaces the 40 KB of thread-private fp32 columns, and gives sum_px z_label = <tile, s> for the loss;
//   * tap tiles leave with one float4 atomicAdd per 4 slots as before.
// Work per label pixel drops from ~220 to ~125 thread-instructions and 16 warps per SM are resident.
#include "pm_common.cuh"
#include <cstdlib>

namespace pm {
namespace rw {

constexpr int TX_ = 32;
#ifndef PM_RL_MINB
#define PM_RL_MINB 2
#endif
#ifndef PM_RL_THREADS
#define PM_RL_THREADS 256
#endif
constexpr int THREADS = PM_RL_THREADS, WARPS = THREADS / 32, TX = 32, KP = 20, NH = KP / 2;
constexpr int TPC = THREADS / TX;  // threads per cell in the y-weighting step
constexpr int REC = 44;      // floats per (tile row, cell) in the tap tiles: 2 x 20 sums, padded so that 128-bit accesses are conflict-free
constexpr int SLOTF = TX_ * 40 + (TX_ / 4) * 4;   // floats per row slot of the records: cell c at c*40 + (c/4)*4 (conflict-free too)
constexpr int OSTR = 21;     // words per feature pixel in the one-hot tile (odd: neighbouring cells hit distinct banks)
constexpr int TCOLS = TX + 1;
constexpr int CHUNK = 3;     // label pixels per unrolled group
constexpr int SEG = 9;       // label pixels whose labels are loaded together (cells are 8-9 / 16-17 pixels wide)
constexpr int MAX_CR = 4;

__device__ __forceinline__ float ex2a(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2a(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcpa(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ int bil_index(float scale, int dst, int n_in) {
    const int i0 = (int)(scale * (float)dst);
    return i0 > n_in - 1 ? n_in - 1 : i0;
}
// smallest dst in [0, n_out] whose bilinear source index is >= c
__device__ __forceinline__ int first_ge(int c, float scale, int n_out, int n_in) {
    if (c <= 0) return 0;
    if (c > n_in - 1 || scale <= 0.f) return n_out;
    int y = (int)ceilf((float)c / scale);
    y = max(0, min(y, n_out));
    while (y > 0 && bil_index(scale, y - 1, n_in) >= c) --y;
    while (y < n_out && bil_index(scale, y, n_in) < c) ++y;
    return y;
}

// One label row of one cell: softmax sums along x (R0 = sum p*(1-lambda_x), R1 = sum p*lambda_x per slot) and the one-hot
// term into the fixed-point tile. A = z at lambda_x = 0, Bc = dz/dlambda_x (log2 units, stabilised by the tile maximum).
template <bool EXACT>
__device__ __forceinline__ void row_pixels(const float2 (&A)[NH], const float2 (&Bc)[NH], float sx, float cxf, int Xa,
                                           int ncols, int nc_max, const unsigned char* __restrict__ lrow, int K, float hy,
                                           float lamy, unsigned* __restrict__ o_top, int o_rowstride, float fscale,
                                           float2 (&R0)[NH], float2 (&R1)[NH], float& lossacc) {
    // labels of the first segment go out before the exp2 set-up below, so their latency is covered
    int labs[SEG];
#pragma unroll
    for (int i = 0; i < SEG; ++i) labs[i] = (i < ncols) ? (int)__ldg(lrow + i) : 255;
    float2 E[NH], Rt[NH];
    const float lam_first = fminf(fmaxf(sx * (float)Xa - cxf, 0.f), 1.f);
    const float2 lf2 = make_float2(lam_first, lam_first), sx2 = make_float2(sx, sx);
#pragma unroll
    for (int i = 0; i < NH; ++i) {
        if (!EXACT) {
            const float2 e0 = __ffma2_rn(lf2, Bc[i], A[i]), rr = __fmul2_rn(sx2, Bc[i]);
            E[i] = make_float2(ex2a(e0.x), ex2a(e0.y));
            Rt[i] = make_float2(ex2a(rr.x), ex2a(rr.y));
        } else {
            E[i] = A[i];
            Rt[i] = Bc[i];
        }
        R0[i] = R1[i] = make_float2(0.f, 0.f);
    }
    int cur = -1;
    float ox0 = 0.f, ox1 = 0.f;
    auto flush = [&]() {
        if (cur >= 0 && cur < K) {
            unsigned* p = o_top + cur;
            const float a0 = ox0 * fscale, a1 = ox1 * fscale;
            atomicAdd(p, __float2uint_rn(hy * a0));
            atomicAdd(p + OSTR, __float2uint_rn(hy * a1));
            atomicAdd(p + o_rowstride, __float2uint_rn(lamy * a0));
            atomicAdd(p + o_rowstride + OSTR, __float2uint_rn(lamy * a1));
        }
    };
    for (int S0 = 0; S0 < nc_max; S0 += SEG) {   // warp-uniform; pixels past a lane's row are predicated off (label 255)
        if (S0 > 0) {
#pragma unroll
            for (int i = 0; i < SEG; ++i) labs[i] = (S0 + i < ncols) ? (int)__ldg(lrow + S0 + i) : 255;
        }
#pragma unroll
        for (int c = 0; c < SEG / CHUNK; ++c) {
            const int X0 = S0 + c * CHUNK;
            if (X0 >= nc_max) break;
            float lamx[CHUNK], hx[CHUNK];
#pragma unroll
            for (int i = 0; i < CHUNK; ++i) {
                lamx[i] = fminf(fmaxf(sx * (float)(Xa + X0 + i) - cxf, 0.f), 1.f);
                hx[i] = 1.f - lamx[i];
            }
#pragma unroll
            for (int i = 0; i < CHUNK; ++i) {
                const bool valid = labs[c * CHUNK + i] < K;  // 0..K-1 real class; K = ignore; 255 = outside this lane's row
                float pshift = 0.f;       // the tile is stored stabilised: only the exact path adds a per-pixel maximum
                float2 e[NH];
                if (EXACT) {
                    float m = -INFINITY;
#pragma unroll
                    for (int q = 0; q < NH; ++q) {
                        e[q] = __ffma2_rn(make_float2(lamx[i], lamx[i]), Rt[q], E[q]);
                        m = fmaxf(m, fmaxf(e[q].x, e[q].y));
                    }
                    pshift = m;
#pragma unroll
                    for (int q = 0; q < NH; ++q) e[q] = make_float2(ex2a(e[q].x - m), ex2a(e[q].y - m));
                }
                const float2* ev = EXACT ? e : E;
                float2 s01 = __fadd2_rn(ev[0], ev[1]), s23 = __fadd2_rn(ev[2], ev[3]), s45 = __fadd2_rn(ev[4], ev[5]);
                float2 s67 = __fadd2_rn(ev[6], ev[7]), s89 = __fadd2_rn(ev[8], ev[9]);
                s01 = __fadd2_rn(s01, s23), s45 = __fadd2_rn(s45, s67);
                s01 = __fadd2_rn(__fadd2_rn(s01, s45), s89);
                const float sum = s01.x + s01.y;
                // straight-line on purpose (a 0/1 factor instead of `if (valid)`): with a predicated tree the compiler
                // fences every pixel into its own region and the three pixels of a group no longer overlap
                const float vf = valid ? 1.f : 0.f;
                lossacc = fmaf(vf, pshift + lg2a(sum), lossacc);
                const float inv = rcpa(sum) * vf;
                const float2 ihx = make_float2(inv * hx[i], inv * hx[i]), ilx = make_float2(inv * lamx[i], inv * lamx[i]);
#pragma unroll
                for (int q = 0; q < NH; ++q) {
                    R0[q] = __ffma2_rn(ev[q], ihx, R0[q]);
                    R1[q] = __ffma2_rn(ev[q], ilx, R1[q]);
                }
                if (!EXACT) {
#pragma unroll
                    for (int q = 0; q < NH; ++q) E[q] = __fmul2_rn(E[q], Rt[q]);
                }
            }
            // one-hot term, run-length compressed: flush when the class changes
#pragma unroll
            for (int i = 0; i < CHUNK; ++i) {
                const int cls = labs[c * CHUNK + i];
                const bool ok = cls != 255;
                if (ok && cls != cur) {
                    flush();
                    ox0 = ox1 = 0.f;
                    cur = cls;
                }
                ox0 += ok ? hx[i] : 0.f;
                ox1 += ok ? lamx[i] : 0.f;
            }
        }
    }
    flush();
}

__global__ void __launch_bounds__(THREADS, PM_RL_MINB)
    readloss_rows_kernel(const float* __restrict__ s, const unsigned char* __restrict__ lab8, float inv_T, float temperature, int h,
                         int w, int Hm, int Wm, int K, float sy, float sx, int CR, int slots, int tiles_x, int bands, float fscale,
                         float* __restrict__ ds_rl, unsigned long long* __restrict__ ws, float* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int TR = CR + 1;                               // feature rows of the tile
    float* s_tile = smem;                                // [TR][33][KP] similarities (clamped at the map's edge)
    float* lr_tile = s_tile + TR * TCOLS * KP;           // [TR][32][REC] tap sums: floats 0..19 left tap, 20..39 right tap
    unsigned* o_tile = reinterpret_cast<unsigned*>(lr_tile + TR * TX * REC);  // [TR][33][OSTR] fixed-point one-hot weights
    float* rec = reinterpret_cast<float*>(o_tile + TR * TCOLS * OSTR + 3 - (TR * TCOLS * OSTR + 3) % 4);  // [slots][SLOTF]
    float* lamy_s = rec + (size_t)slots * SLOTF;      // [slots]
    int* geo = reinterpret_cast<int*>(lamy_s + slots);   // [0..33] Xa of the tile's cells (+1), [40..40+CR] Ya of the band's cell rows
    float* red = reinterpret_cast<float*>(geo + 48);     // [16]

    int bid = blockIdx.x;
    const int tx_i = bid % tiles_x;
    bid /= tiles_x;
    const int band = bid % bands, b = bid / bands;
    const int cy0 = band * CR, fx0 = tx_i * TX;

    // ---- prologue: geometry, zeroed accumulators, and the tile: z' = s * c2 - max(tile) in log2 units, padded slots at
    //      -1e20 (their exp2 is 0 and their slope 0, so the row code needs no slot masks). One stabiliser per CTA is enough
    //      when the tile's spread is small (always, for similarities of normalised vectors); otherwise the exact path runs.
    const float c2 = inv_T * 1.4426950408889634f;
    float tmax = -INFINITY, tmin = INFINITY;
    for (int i = tid; i < TR * TCOLS * (KP / 4); i += THREADS) {
        const int e = i / (KP / 4), q = i - e * (KP / 4);
        const int ty = e / TCOLS, tx = e - ty * TCOLS;
        const int fy = min(cy0 + ty, h - 1), fx = min(fx0 + tx, w - 1);
        float4 v = __ldg(reinterpret_cast<const float4*>(s + ((size_t)(b * h + fy) * w + fx) * KP) + q);
        v.x *= c2, v.y *= c2, v.z *= c2, v.w *= c2;
        if (4 * q + 0 < K) tmax = fmaxf(tmax, v.x), tmin = fminf(tmin, v.x); else v.x = -1e20f;
        if (4 * q + 1 < K) tmax = fmaxf(tmax, v.y), tmin = fminf(tmin, v.y); else v.y = -1e20f;
        if (4 * q + 2 < K) tmax = fmaxf(tmax, v.z), tmin = fminf(tmin, v.z); else v.z = -1e20f;
        if (4 * q + 3 < K) tmax = fmaxf(tmax, v.w), tmin = fminf(tmin, v.w); else v.w = -1e20f;
        if (!(v.x == v.x && v.y == v.y && v.z == v.z && v.w == v.w)) tmin = -INFINITY;  // NaN: force the exact path
        reinterpret_cast<float4*>(s_tile)[i] = v;
    }
    for (int i = tid; i < TR * TX * REC / 4; i += THREADS) reinterpret_cast<float4*>(lr_tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < (TR * TCOLS * OSTR + 3) / 4; i += THREADS) reinterpret_cast<uint4*>(o_tile)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid <= TX) geo[tid] = first_ge(fx0 + tid, sx, Wm, w);
    else if (tid >= 64 && tid <= 64 + CR) geo[40 + tid - 64] = first_ge(min(cy0 + tid - 64, h), sy, Hm, h);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
    }
    if (lane == 0) red[wid] = tmax, red[WARPS + wid] = tmin;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < WARPS; ++i) tmax = fmaxf(tmax, red[i]), tmin = fminf(tmin, red[WARPS + i]);
    const float shift = tmax;
    const bool exact = !(tmax - tmin < 60.f);  // also inf / NaN tiles
    for (int i = tid; i < TR * TCOLS * (KP / 4); i += THREADS) {   // each thread re-visits the elements it wrote
        float4 v = reinterpret_cast<float4*>(s_tile)[i];
        const int q = i % (KP / 4);
        if (4 * q + 0 < K) v.x -= shift;
        if (4 * q + 1 < K) v.y -= shift;
        if (4 * q + 2 < K) v.z -= shift;
        if (4 * q + 3 < K) v.w -= shift;
        reinterpret_cast<float4*>(s_tile)[i] = v;
    }
    __syncthreads();

    const int cx = fx0 + lane;
    const int Xa = geo[lane], Xb = geo[lane + 1];
    const int ncols = Xb - Xa;  // 0 for cells beyond the map
    const int nc_max = __reduce_max_sync(0xffffffffu, ncols);
    const int Y0 = geo[40], Y1 = geo[40 + CR];  // label rows of the band
    const float cxf = (float)cx;
    const unsigned char* lab_b = lab8 + (size_t)b * Hm * Wm;
    float lossacc = 0.f;  // log2 units

    for (int base = Y0; base < Y1; base += slots) {   // one pass unless the band has more label rows than record slots
        const int pend = min(Y1, base + slots);
        for (int Y = base + wid; Y < pend; Y += WARPS) {
            const int cy = bil_index(sy, Y, h), cyl = cy - cy0;
            const float lamy = fminf(fmaxf(sy * (float)Y - (float)cy, 0.f), 1.f), hy = 1.f - lamy;
            // stabilised logits of the row at the cell's left (A) and right edge, then the slope Bc = right - left
            const float4* p00 = reinterpret_cast<const float4*>(s_tile + (cyl * TCOLS + lane) * KP);
            const float4 *p01 = p00 + KP / 4, *p10 = p00 + TCOLS * (KP / 4), *p11 = p10 + KP / 4;
            const float2 ly2 = make_float2(lamy, lamy), neg1 = make_float2(-1.f, -1.f);
            float2 A[NH], Bc[NH];
#pragma unroll
            for (int q = 0; q < KP / 4; ++q) {
                const float4 a = p00[q], bq = p01[q], c = p10[q], d = p11[q];
                const float2 a0 = make_float2(a.x, a.y), a1 = make_float2(a.z, a.w), b0 = make_float2(bq.x, bq.y), b1 = make_float2(bq.z, bq.w);
                const float2 l0 = __ffma2_rn(ly2, __ffma2_rn(a0, neg1, make_float2(c.x, c.y)), a0);
                const float2 l1 = __ffma2_rn(ly2, __ffma2_rn(a1, neg1, make_float2(c.z, c.w)), a1);
                const float2 r0 = __ffma2_rn(ly2, __ffma2_rn(b0, neg1, make_float2(d.x, d.y)), b0);
                const float2 r1 = __ffma2_rn(ly2, __ffma2_rn(b1, neg1, make_float2(d.z, d.w)), b1);
                A[2 * q] = l0, A[2 * q + 1] = l1;
                Bc[2 * q] = __ffma2_rn(l0, neg1, r0), Bc[2 * q + 1] = __ffma2_rn(l1, neg1, r1);
            }
            float2 R0[NH], R1[NH];
            unsigned* o_top = o_tile + (cyl * TCOLS + lane) * OSTR;
            const unsigned char* lrow = lab_b + (size_t)Y * Wm + Xa;
            if (!exact) row_pixels<false>(A, Bc, sx, cxf, Xa, ncols, nc_max, lrow, K, hy, lamy, o_top, TCOLS * OSTR, fscale, R0, R1, lossacc);
            else row_pixels<true>(A, Bc, sx, cxf, Xa, ncols, nc_max, lrow, K, hy, lamy, o_top, TCOLS * OSTR, fscale, R0, R1, lossacc);
            float4* dst = reinterpret_cast<float4*>(rec + (size_t)(Y - base) * SLOTF + lane * 40 + (lane >> 2) * 4);
#pragma unroll
            for (int q = 0; q < NH / 2; ++q) {
                dst[q] = make_float4(R0[2 * q].x, R0[2 * q].y, R0[2 * q + 1].x, R0[2 * q + 1].y);
                dst[NH / 2 + q] = make_float4(R1[2 * q].x, R1[2 * q].y, R1[2 * q + 1].x, R1[2 * q + 1].y);
            }
            if (lane == 0) lamy_s[Y - base] = lamy;
        }
        __syncthreads();
        // ---- y-weighting of the row records into the tap tiles: thread = (cell, quad of the 10 float4 of a record)
        {
            const int cell = tid / TPC, j = tid % TPC;
            for (int cyl = 0; cyl < CR; ++cyl) {
                const int ra = max(geo[40 + cyl], base), rb = min(geo[40 + cyl + 1], pend);
                if (ra >= rb) continue;
                float4* top = reinterpret_cast<float4*>(lr_tile + ((size_t)cyl * TX + cell) * REC);
                float4* bot = top + TX * REC / 4;
#pragma unroll
                for (int idx = j; idx < KP / 2; idx += TPC) {
                    float2 ta = make_float2(0.f, 0.f), tb = ta, ba = ta, bb = ta;
                    const float4* r = reinterpret_cast<const float4*>(rec + (size_t)(ra - base) * SLOTF + cell * 40 + (cell >> 2) * 4) + idx;
                    for (int Y = ra; Y < rb; ++Y, r += SLOTF / 4) {
                        const float ly = lamy_s[Y - base];
                        const float2 ly2 = make_float2(ly, ly), hy2 = make_float2(1.f - ly, 1.f - ly);
                        const float4 v = *r;
                        ta = __ffma2_rn(hy2, make_float2(v.x, v.y), ta), tb = __ffma2_rn(hy2, make_float2(v.z, v.w), tb);
                        ba = __ffma2_rn(ly2, make_float2(v.x, v.y), ba), bb = __ffma2_rn(ly2, make_float2(v.z, v.w), bb);
                    }
                    float4 a = top[idx];
                    a.x += ta.x, a.y += ta.y, a.z += tb.x, a.w += tb.y;
                    top[idx] = a;
                    a = bot[idx];
                    a.x += ba.x, a.y += ba.y, a.z += bb.x, a.w += bb.y;
                    bot[idx] = a;
                }
            }
        }
        __syncthreads();
    }

    // ---- flush: tap (row, col) = left tap of cell col + right tap of cell col-1, minus the one-hot weights; the loss gets
    //      -c2 * <one-hot weights, s>. Clamped rows / columns fold onto the edge pixel through the atomics.
    {
        const float inv_fs = 1.f / fscale;
        float zy = 0.f;
        for (int i = tid; i < TR * TCOLS * (KP / 4); i += THREADS) {
            const int e = i / (KP / 4), q = i - e * (KP / 4);
            const int ty = e / TCOLS, tx = e - ty * TCOLS;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tx < TX) v = reinterpret_cast<const float4*>(lr_tile + ((size_t)ty * TX + tx) * REC)[q];
            if (tx > 0) {
                const float4 r = reinterpret_cast<const float4*>(lr_tile + ((size_t)ty * TX + tx - 1) * REC)[KP / 4 + q];
                v.x += r.x, v.y += r.y, v.z += r.z, v.w += r.w;
            }
            const unsigned* o = o_tile + e * OSTR + 4 * q;
            const float o0 = (float)o[0] * inv_fs, o1 = (float)o[1] * inv_fs, o2 = (float)o[2] * inv_fs, o3 = (float)o[3] * inv_fs;
            const float4 sv = reinterpret_cast<const float4*>(s_tile)[i];
            zy = fmaf(o0, sv.x, fmaf(o1, sv.y, fmaf(o2, sv.z, fmaf(o3, sv.w, zy))));
            v.x -= o0, v.y -= o1, v.z -= o2, v.w -= o3;
            if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
                const int fy = min(cy0 + ty, h - 1), fx = min(fx0 + tx, w - 1);
                atomicAdd(reinterpret_cast<float4*>(ds_rl + ((size_t)(b * h + fy) * w + fx) * KP) + q, v);
            }
        }
        lossacc -= zy;  // the tile holds z' = s*c2 - shift: sum_px z'_label, in the same log2 units as lg2(sum)
    }
    lossacc = warp_sum(lossacc);
    __syncthreads();  // `red` was read by everyone in the prologue
    if (lane == 0) red[wid] = lossacc;
    __syncthreads();
    if (tid == 0) {
        float tot = 0.f;
        for (int i = 0; i < WARPS; ++i) tot += red[i];
        atomicAdd(reinterpret_cast<double*>(ws + PM_WS_LOSS_SUM), (double)tot * 0.6931471805599453);
        // last CTA: readloss = loss_sum / V ; scale = 1 / (V*T)   (V from the histogram pm_labels_pack left in ws)
        __threadfence();
        const unsigned long long ticket = atomicAdd(ws + PM_WS_COUNTER, 1ULL);
        if (ticket == (unsigned long long)gridDim.x - 1) {
            __threadfence();
            unsigned long long V = 0;
            for (int k = 0; k < K; ++k) V += atomicAdd(ws + PM_WS_HIST + k, 0ULL);
            const double sum = __longlong_as_double((long long)atomicAdd(ws + PM_WS_LOSS_SUM, 0ULL));
            out[0] = (float)(sum / (double)V);  // V == 0 -> 0/0 = NaN like torch
            out[1] = (float)(1.0 / ((double)V * (double)temperature));
        }
    }
}

}  // namespace rw
}  // namespace pm

// Returns 0 when launched, a positive CUDA error, or -1 when the shape is outside what the fixed-point one-hot tile covers
// (the caller then runs the second-generation kernel).
int pm_readloss_rows_launch(const float* s, const uint8_t* lab8, float temperature, int B, int h, int w, int Hm, int Wm, int K,
                            float* ds_rl, void* ws, float* out, cudaStream_t st) {
    using namespace pm::rw;
    const float sy = Hm > 1 ? (float)(h - 1) / (float)(Hm - 1) : 0.f;
    const float sx = Wm > 1 ? (float)(w - 1) / (float)(Wm - 1) : 0.f;
    // label rows / columns per cell (upper bounds)
    const int rmax = h > 1 ? (Hm - 1) / (h - 1) + 2 : Hm, cmax = w > 1 ? (Wm - 1) / (w - 1) + 2 : Wm;
    const double bound = 4.0 * (double)rmax * (double)cmax;   // > sum of the weights one tile element can receive
    int bits = 0;
    while (bits < 30 && (double)(1ull << (bits + 1)) * bound <= 4294967296.0) ++bits;
    if (bits < 16) return -1;
    // narrow cells (label map < 6x the feature map: the head's main loss at stride 4): a row of 4-5 pixels does not repay
    // the per-row exp2 set-up -- the one-thread-per-cell kernel is faster there (158 vs 240 us at 192x192 -> 768x768)
    if (w > 1 && (Wm - 1) / (w - 1) < 6) return -1;
    const float fscale = (float)(1u << bits);
    const int rtyp = h > 1 ? ((Hm - 1) / (h - 1) > 1 ? (Hm - 1) / (h - 1) : 1) : Hm;  // typical label rows per cell
    int CR = 1;  // cell rows per CTA: about 8 label rows (one per warp) when cells are short
    while (CR < MAX_CR && (CR + 1) * rtyp <= WARPS && CR < h) ++CR;
    if (const char* e = getenv("PM_RL_CR")) CR = atoi(e) < 1 ? 1 : (atoi(e) > MAX_CR ? MAX_CR : atoi(e));  // tuning switch
    int slots = CR * rmax;
    if (slots > 16) slots = 16;
    const int TR = CR + 1;
    const int tiles_x = (w + TX - 1) / TX, bands = (h + CR - 1) / CR;
    const int o_words = TR * TCOLS * OSTR;
    const size_t smem = sizeof(float) * ((size_t)TR * TCOLS * KP + (size_t)TR * TX * REC + (o_words + 3 - (o_words + 3) % 4) +
                                         (size_t)slots * SLOTF + slots + 48 + 16);
    const long long grid = (long long)B * tiles_x * bands;
    if (grid > 0x7fffffffLL) return -1;
    cudaError_t e = cudaFuncSetAttribute(readloss_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    readloss_rows_kernel<<<(int)grid, THREADS, smem, st>>>(s, lab8, 1.f / temperature, temperature, h, w, Hm, Wm, K, sy, sx, CR, slots,
                                                           tiles_x, bands, fscale, ds_rl, (unsigned long long*)ws, out);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}
