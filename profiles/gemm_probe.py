"""Bring-up / timing probe of the tcgen05 1x1-convolution GEMMs (csrc/pm_gemm.cu) against fp64 torch matmuls.
Run on a B200:  python profiles/gemm_probe.py [--time]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pinthememory_b200 import capi

torch.manual_seed(0)
dev = "cuda"


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def check_fwd(B, K, M, h, w, dtype, transpose=False, stats=True, accumulate=False, ints=False):
    hw = h * w
    if ints:
        x = torch.randint(-4, 5, (B, K, h, w), device=dev).to(dtype)
        W = torch.randint(-4, 5, ((K, M) if transpose else (M, K)), device=dev).float()
    else:
        x = torch.randn(B, K, h, w, device=dev).to(dtype)
        W = torch.randn((K, M) if transpose else (M, K), device=dev) / K ** 0.5
    hi, lo = capi.conv1x1_prep(W, transpose, dtype)
    Wm = (W.t() if transpose else W)
    if dtype == torch.bfloat16:
        Wm = Wm.bfloat16()
    ref = torch.einsum("mk,bkp->bmp", Wm.double(), x.double().view(B, K, hw)).view(B, M, h, w)
    st = torch.zeros(2 * M, dtype=torch.float64, device=dev) if stats else None
    y0 = None
    if accumulate:
        y0 = torch.randn(B, M, h, w, device=dev).to(dtype)
        ref = ref + y0.double()
        y0 = y0.clone()
    y = capi.conv1x1_fwd(x, hi, lo, M, y=y0, stats=st, accumulate=accumulate)
    torch.cuda.synchronize()
    e = rel(y, ref)
    msg = f"fwd B={B} K={K} M={M} hw={h}x{w} {str(dtype)[6:]} T={int(transpose)} acc={int(accumulate)} ints={int(ints)}: relL2={e[0]:.2e} maxrel={e[1]:.2e}"
    if stats and not accumulate:
        yy = y.double()
        s_ref = torch.cat([yy.sum((0, 2, 3)), (yy * yy).sum((0, 2, 3))])
        es = rel(st, s_ref)
        msg += f" stats relL2={es[0]:.2e}"
    print(msg, flush=True)
    return e[0]


def check_wgrad(B, M, N, h, w, dtype, accumulate=False):
    hw = h * w
    dy = torch.randn(B, M, h, w, device=dev).to(dtype)
    x = torch.randn(B, N, h, w, device=dev).to(dtype)
    ref = torch.einsum("bmp,bnp->mn", dy.double().view(B, M, hw), x.double().view(B, N, hw))
    dW0 = None
    if accumulate:
        dW0 = torch.randn(M, N, device=dev)
        ref = ref + dW0.double()
    dW = capi.conv1x1_wgrad(dy, x, dW=dW0, accumulate=accumulate)
    torch.cuda.synchronize()
    e = rel(dW, ref)
    print(f"wgrad B={B} M={M} N={N} hw={h}x{w} {str(dtype)[6:]} acc={int(accumulate)}: relL2={e[0]:.2e} maxrel={e[1]:.2e}", flush=True)
    return e[0]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def main():
    f32, bf = torch.float32, torch.bfloat16
    print(torch.cuda.get_device_name(0), flush=True)
    check_fwd(1, 32, 128, 8, 16, f32, ints=True)
    check_fwd(1, 32, 128, 8, 16, f32)
    check_fwd(1, 64, 256, 16, 16, f32)
    check_fwd(2, 288, 256, 96, 96, f32)
    check_fwd(2, 256, 288, 96, 96, f32, transpose=True)
    check_fwd(2, 512, 256, 48, 48, f32)
    check_fwd(2, 256, 512, 48, 48, f32, transpose=True)
    check_fwd(3, 256, 256, 10, 10, f32)
    check_fwd(2, 256, 256, 6, 6, f32)
    check_fwd(2, 288, 256, 48, 48, f32, accumulate=True)
    check_fwd(1, 64, 128, 8, 16, bf, ints=True)
    check_fwd(2, 288, 256, 96, 96, bf)
    check_fwd(2, 256, 288, 96, 96, bf, transpose=True)
    check_fwd(2, 256, 256, 10, 8, bf)
    check_fwd(2, 288, 256, 48, 48, bf, accumulate=True)
    check_wgrad(1, 128, 32, 8, 16, f32)
    check_wgrad(2, 256, 288, 96, 96, f32)
    check_wgrad(2, 256, 256, 48, 48, f32)
    check_wgrad(2, 256, 512, 48, 48, f32)
    check_wgrad(3, 256, 256, 10, 10, f32, accumulate=True)
    check_wgrad(2, 256, 288, 96, 96, bf)
    check_wgrad(2, 256, 512, 48, 48, bf)
    check_wgrad(2, 256, 256, 10, 8, bf)
    if "--time" in sys.argv:
        for dtype in (f32, bf):
            B, h, w = 8, 96, 96
            x = torch.randn(B, 288, h, w, device=dev).to(dtype)
            g = torch.randn(B, 256, h, w, device=dev).to(dtype)
            W = torch.randn(256, 288, device=dev) / 17
            hi, lo = capi.conv1x1_prep(W, False, dtype)
            hit, lot = capi.conv1x1_prep(W, True, dtype)
            y = torch.empty(B, 256, h, w, device=dev, dtype=dtype)
            dx = torch.empty(B, 288, h, w, device=dev, dtype=dtype)
            st = torch.zeros(512, dtype=torch.float64, device=dev)
            dW = torch.empty(256, 288, device=dev)
            print(str(dtype), "fwd  us", timeit(lambda: capi.conv1x1_fwd(x, hi, lo, 256, y=y, stats=st)), flush=True)
            print(str(dtype), "dgrad us", timeit(lambda: capi.conv1x1_fwd(g, hit, lot, 288, y=dx)), flush=True)
            print(str(dtype), "wgrad us", timeit(lambda: capi.conv1x1_wgrad(g, x, dW=dW)), flush=True)
            Wc = W.to(dtype).view(256, 288, 1, 1)
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            print(str(dtype), "cudnn fwd (tf32 off) us", timeit(lambda: torch.nn.functional.conv2d(x, Wc)), flush=True)
            torch.backends.cudnn.allow_tf32 = True
            print(str(dtype), "cudnn fwd (tf32 on) us", timeit(lambda: torch.nn.functional.conv2d(x, Wc)), flush=True)


if __name__ == "__main__":
    main()
