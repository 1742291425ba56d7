"""Pure-write HBM bandwidth on this GPU (the roofline of kernels that only write: rbf_queries_kernel's Kxt chunk, the feature
pass of the grid stage): cudaMemset via torch.zero_ and a torch fill, 1 GiB, best of 10, CUDA events; the copy figure beside it."""
import torch
n = 1 << 27                                   # 1 GiB of float64
a = torch.empty(n, dtype=torch.float64, device='cuda')
b = torch.empty(n, dtype=torch.float64, device='cuda')
def best(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
t = best(lambda: a.zero_())
print('zero_ (memset) 1 GiB: %.3f ms = %.0f GB/s written' % (t, 8 * n / t / 1e6))
t = best(lambda: a.fill_(1.5))
print('fill_ kernel  1 GiB: %.3f ms = %.0f GB/s written' % (t, 8 * n / t / 1e6))
t = best(lambda: b.copy_(a))
print('copy_ 1 GiB -> 1 GiB: %.3f ms = %.0f GB/s read + written' % (t, 16 * n / t / 1e6))
t = best(lambda: torch.sum(a))
print('sum (pure read) 1 GiB: %.3f ms = %.0f GB/s read' % (t, 8 * n / t / 1e6))
