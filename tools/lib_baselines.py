"""Library comparators on the same GPU (context for profiles/, not part of the product path):
  * cuSOLVER FP64 potrf (torch.linalg.cholesky) and cuBLAS DGEMM at the configurations' matrix sizes;
  * batched cuSOLVER potrf + potri-equivalent (cholesky_inverse) at the C4 shape;
  * cuBLASLt INT8 GEMM (torch._int_mm) -- the ceiling an Ozaki-style FP64 emulation on the low-precision tensor cores would
    work against.
usage: python tools/lib_baselines.py"""
import json
import torch


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def spd(n, batch=None):
    g = torch.Generator(device="cuda").manual_seed(0)
    shape = (n, n) if batch is None else (batch, n, n)
    a = torch.randn(shape, dtype=torch.float64, device="cuda", generator=g)
    return a @ a.transpose(-1, -2) / n + torch.eye(n, dtype=torch.float64, device="cuda")


out = {}
for n in (5000, 16384):
    A = spd(n)
    ms = timed(lambda: torch.linalg.cholesky(A), reps=3, warm=1)
    out[f"cusolver_potrf_n{n}"] = {"ms": ms, "tflops": n ** 3 / 3 / ms / 1e9}
    B = torch.randn((n, n), dtype=torch.float64, device="cuda")
    ms = timed(lambda: torch.mm(A, B), reps=3, warm=1)
    out[f"cublas_dgemm_n{n}"] = {"ms": ms, "tflops": 2 * n ** 3 / ms / 1e9}
    del A, B
n, batch = 600, 2000
A = spd(n, batch)
ms = timed(lambda: torch.linalg.cholesky(A), reps=3, warm=1)
out["cusolver_potrf_batched_n600_b2000"] = {"ms": ms, "tflops": batch * n ** 3 / 3 / ms / 1e9}
L = torch.linalg.cholesky(A)
ms = timed(lambda: torch.cholesky_inverse(L), reps=3, warm=1)
out["cholesky_inverse_batched_n600_b2000"] = {"ms": ms, "tflops": batch * 2 * n ** 3 / 3 / ms / 1e9}
del A, L
for n in (640, 2048, 4096, 8192, 16384):
    a = torch.randint(-127, 127, (n, n), dtype=torch.int8, device="cuda")
    b = torch.randint(-127, 127, (n, n), dtype=torch.int8, device="cuda")
    try:
        ms = timed(lambda: torch._int_mm(a, b), reps=5, warm=2)
        out[f"cublaslt_int8_gemm_n{n}"] = {"ms": ms, "tops": 2 * n ** 3 / ms / 1e9}
    except Exception as e:  # noqa: BLE001
        out[f"cublaslt_int8_gemm_n{n}"] = {"error": str(e)[:200]}
print(json.dumps(out, indent=1))
