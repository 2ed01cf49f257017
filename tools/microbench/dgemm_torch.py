"""cuBLAS fp64 GEMM throughput via torch (library number, for the roofline denominator only)."""
import torch, time
dev = "cuda:0"
def run(m, n, k, reps=10):
    a = torch.randn(m, k, device=dev, dtype=torch.float64)
    b = torch.randn(k, n, device=dev, dtype=torch.float64)
    torch.matmul(a, b); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"dgemm {m}x{n}x{k}: {2.0*m*n*k/best/1e9:.2f} TFLOP/s ({best:.3f} ms)")
run(8192, 8192, 8192, 5)
run(4096, 4096, 4096)
run(64, 20000, 1000)
run(1000, 20000, 64)
run(32, 2500, 1000)
