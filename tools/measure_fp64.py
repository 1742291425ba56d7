"""Measure the FP64 / TF32 GEMM yardsticks of this GPU the way MEASURED_PEAKS.json was measured
(torch.matmul, CUDA events, best of N) and this library's own DMMA GEMM beside them."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from informative_path_planning_b200 import _lib as L

def best_ms(fn, n=6):
    best = None
    for i in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i: best = a.elapsed_time(b) if best is None else min(best, a.elapsed_time(b))
    return best

out = {}
n = 8192
A = torch.randn(n, n, dtype=torch.float64, device='cuda'); B = torch.randn(n, n, dtype=torch.float64, device='cuda')
C = torch.empty(n, n, dtype=torch.float64, device='cuda')
out['cublas_dgemm_tflops'] = 2.0 * n**3 / best_ms(lambda: torch.matmul(A, B.T, out=C)) / 1e9
out['ipp_dgemm_nt_tflops'] = 2.0 * n**3 / best_ms(lambda: L.check(L.lib.ipp_dgemm_nt(n, n, n, 1.0, L.ptr(A), n, L.ptr(B), n, 0.0, L.ptr(C), n, None, 0, L.stream_ptr()))) / 1e9
ref = A @ B.T
out['ipp_dgemm_max_abs_err'] = float((C - ref).abs().max())
del A, B, C, ref
Af = torch.randn(n, n, device='cuda'); Bf = torch.randn(n, n, device='cuda')
torch.backends.cuda.matmul.allow_tf32 = True
out['cublas_tf32_tflops'] = 2.0 * n**3 / best_ms(lambda: torch.matmul(Af, Bf)) / 1e9
torch.backends.cuda.matmul.allow_tf32 = False
out['cublas_fp32_tflops'] = 2.0 * n**3 / best_ms(lambda: torch.matmul(Af, Bf)) / 1e9
print(json.dumps(out))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/fp64_peaks.json', 'w'), indent=1)
