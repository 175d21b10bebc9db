"""Library FP64 baselines on the GPU box: cuBLAS DGEMM (roofline denominator) and
cuSOLVER/MAGMA batched potrf via torch (sanity baseline, not the product). SURVEY.md §8(d)."""
import json
import time

import torch


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    out = {}
    dev = torch.device("cuda:0")
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        ms = ev_time(lambda: torch.matmul(a, b))
        out[f"dgemm_{n}_tflops_burst"] = 2 * n**3 / ms * 1e-9
        if n == 8192:
            torch.cuda.synchronize()
            t0 = time.time()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            k = 0
            while k < 60:
                torch.matmul(a, b)
                k += 1
            e1.record()
            e1.synchronize()
            out["dgemm_8192_tflops_sustained"] = 2 * n**3 * k / e0.elapsed_time(e1) * 1e-9
            out["dgemm_sustained_secs"] = time.time() - t0
        del a, b
    # batched K=128 update shape: C[64,2048,2048] -= A[64,2048,128] @ A^T
    a = torch.randn(64, 2048, 128, dtype=torch.float64, device=dev)
    c = torch.randn(64, 2048, 2048, dtype=torch.float64, device=dev)
    ms = ev_time(lambda: torch.baddbmm(c, a, a.transpose(1, 2), alpha=-1.0))
    out["bgemm_64x2048x2048x128_tflops"] = 64 * 2 * 2048 * 2048 * 128 / ms * 1e-9
    del a, c
    for n, p in ((512, 64), (2048, 64)):
        x = torch.randn(p, n, n, dtype=torch.float64, device=dev)
        k = x @ x.transpose(1, 2) + n * torch.eye(n, dtype=torch.float64, device=dev)
        del x
        ms = ev_time(lambda: torch.linalg.cholesky(k), reps=3, warm=1)
        out[f"torch_cholesky_n{n}_p{p}_ms"] = ms
        out[f"torch_cholesky_n{n}_p{p}_tflops"] = p * n**3 / 3 / ms * 1e-9
        ms1 = ev_time(lambda: torch.linalg.cholesky(k[0]), reps=3, warm=1)
        out[f"torch_cholesky_n{n}_single_ms"] = ms1
        del k
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
