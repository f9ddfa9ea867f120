// Internal cross-file declarations (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pm {

// true when the pipelined (16-byte async copy) kernels can be used for these feature pointers
bool tiled_ok(const void* p0, const void* p1, const void* p2, int hw, int dtype);

int read_fwd_tiled(const void* x, const float* M, const float* gum_m, const float* gum_q, void* u, float* s, float* p,
                   float* colpart, int B, int C, int hw, int K, int dtype, int planes, cudaStream_t st);

// needs a [N][stride] ds buffer (never NULL); dx_add: NULL, or a second gradient of x to sum into dx
int read_bwd_tiled(const void* du, const void* x, const float* M, const float* p, const float* ds_rl,
                   const float* g_loss, const float* rl_out, void* dx, const void* dx_add, float* ds, int B, int C, int hw,
                   int K, int dtype, int planes, cudaStream_t st);

int write_reduce_tiled(const void* f, const void* labels, int lab_u8, float* SD, int B, int C, int h, int w, int Hm, int Wm,
                       int K, int dtype, cudaStream_t st);

int write_bwd_tiled(const float* dS, const void* f, const void* labels, int lab_u8, void* df, int B, int C, int h, int w, int Hm,
                    int Wm, int K, int dtype, cudaStream_t st);

// fills `partial` ([PM_COLPART_ROWS][64]: max at [k], sum at [32+k]) from s (+ gumbel_q) and combines them (row PM_COLPART_ROWS)
int colsoftmax_stats(const float* s, const float* gumbel_q, float* partial, int N, int K, cudaStream_t st);

}  // namespace pm
