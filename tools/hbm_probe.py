import torch, time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device='cuda'); b = torch.empty_like(a)
def t(f, reps=10):
    for _ in range(3): f()
    best = 1e9
    for _ in range(reps):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: b.copy_(a)); print('copy  2 GiB read + 2 GiB write: %.3f ms -> %.0f GB/s (read+write)' % (ms, 2 * n * 2 / ms / 1e6))
ms = t(lambda: b.zero_()); print('memset 2 GiB: %.3f ms -> %.0f GB/s written' % (ms, n * 2 / ms / 1e6))
ms = t(lambda: a.sum()); print('reduce 2 GiB: %.3f ms -> %.0f GB/s read' % (ms, n * 2 / ms / 1e6))
