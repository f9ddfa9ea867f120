// One pass over the label map for the whole step: int64 (or uint8) class ids -> a compact uint8 class map that the
// read loss, the class-sum reduction and its backward all read (4.7 MB instead of 3 x 37.7 MB at 8 x 768 x 768), plus the
// bit-exact integer label histogram and the count of out-of-range values.
//
// Reference: memory.py:220 `tempmask[tempmask == 255] = memory_size` before `one_hot(K+1)`, and the
// `CrossEntropyLoss(ignore_index=255)` of memory.py:117,176. Class ids are 0..K-1, 255 = ignore -> K; anything else
// (negative, >= K and != 255) would make torch's one_hot / CE raise a device assert: here it is mapped to the ignore
// slot AND counted in ws[PM_WS_BAD] so the host side can raise (Memory_sup.check_labels).
#include "pm_common.cuh"

namespace pm {

template <typename TIN>
__global__ void __launch_bounds__(256) labels_pack_kernel(const TIN* __restrict__ labels, unsigned char* __restrict__ lab8,
                                                          long long n, int K, unsigned long long* __restrict__ ws) {
    __shared__ int hist[33];
    __shared__ int bad_sm;
    if (threadIdx.x < 33) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) bad_sm = 0;
    __syncthreads();
    int bad = 0;
    // 8 labels per thread per iteration: one 8-byte store of the packed classes
    const long long n8 = n / 8;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < n8; base += (long long)gridDim.x * blockDim.x) {
        const long long i = base + threadIdx.x;  // the loop itself is CTA-uniform: every lane reaches the match below
        const bool live = i < n8;
        unsigned long long packed = 0ull;
        long long v[8];
        if (live) {
            if constexpr (sizeof(TIN) == 8) {
                const longlong2* p = reinterpret_cast<const longlong2*>(labels + 8 * i);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const longlong2 t = __ldg(p + j);
                    v[2 * j] = t.x, v[2 * j + 1] = t.y;
                }
            } else {
                const unsigned long long t = __ldg(reinterpret_cast<const unsigned long long*>(labels + 8 * i));
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = (long long)((t >> (8 * j)) & 0xffull);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = PM_IGNORE_LABEL;
        }
        int cls[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool ok = v[j] >= 0 && v[j] < K;
            cls[j] = ok ? (int)v[j] : K;
            bad += (!ok && v[j] != PM_IGNORE_LABEL) ? 1 : 0;
            packed |= (unsigned long long)cls[j] << (8 * j);
        }
        // warp-aggregated histogram: lanes with the same class elect one adder (bin 32 swallows the idle lanes). Label
        // maps are piecewise constant, so usually every lane's 8 pixels are one class: one match instead of eight.
        const bool uniform = packed == (unsigned long long)cls[0] * 0x0101010101010101ull;
        if (__all_sync(0xffffffffu, uniform)) {
            const int bin = live ? cls[0] : 32;
            const unsigned peers = __match_any_sync(0xffffffffu, bin);
            if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[bin], 8 * __popc(peers));
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int bin = live ? cls[j] : 32;
                const unsigned peers = __match_any_sync(0xffffffffu, bin);
                if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[bin], __popc(peers));
            }
        }
        if (live) *reinterpret_cast<unsigned long long*>(lab8 + 8 * i) = packed;
    }
    // tail (n not a multiple of 8): first CTA, one label per thread
    if (blockIdx.x == 0) {
        for (long long i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) {
            const long long v = (long long)labels[i];
            const bool ok = v >= 0 && v < K;
            const int c = ok ? (int)v : K;
            bad += (!ok && v != PM_IGNORE_LABEL) ? 1 : 0;
            lab8[i] = (unsigned char)c;
            atomicAdd(&hist[c], 1);
        }
    }
    if (bad) atomicAdd(&bad_sm, bad);
    __syncthreads();
    if (threadIdx.x <= K && hist[threadIdx.x] != 0)
        atomicAdd(ws + PM_WS_HIST + threadIdx.x, (unsigned long long)hist[threadIdx.x]);
    if (threadIdx.x == 0 && bad_sm != 0) atomicAdd(ws + PM_WS_BAD, (unsigned long long)bad_sm);
}

}  // namespace pm

extern "C" int pm_labels_pack(const void* labels, int labels_are_u8, long long n, int K, uint8_t* lab8, void* ws,
                              void* stream) {
    if (!labels || !lab8 || !ws) return PM_ERR_NULL;
    if (K < 1 || K > 31) return PM_ERR_SLOTS;
    if (n <= 0) return PM_ERR_SHAPE;
    if (((uintptr_t)labels & 15) || ((uintptr_t)lab8 & 7) || ((uintptr_t)ws & 7)) return PM_ERR_ALIGN;
    long long blocks = (n / 8 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (labels_are_u8)
        pm::labels_pack_kernel<unsigned char><<<(int)blocks, 256, 0, st>>>((const unsigned char*)labels, lab8, n, K,
                                                                         (unsigned long long*)ws);
    else
        pm::labels_pack_kernel<long long><<<(int)blocks, 256, 0, st>>>((const long long*)labels, lab8, n, K,
                                                                     (unsigned long long*)ws);
    PM_CHECK_LAUNCH();
    return 0;
}
