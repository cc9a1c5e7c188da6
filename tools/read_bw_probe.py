"""Reference points for the HBM roofline on this box: read-only streaming (torch.sum) and copy (b.copy_(a))."""
import torch, json
x = torch.randn(1 << 30, device="cuda")  # 4 GiB fp32
y = torch.empty_like(x)
def t(f, n=10):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms_sum = t(lambda: x.sum())
ms_cp = t(lambda: y.copy_(x))
h = x.view(torch.float16)
ms_sum16 = t(lambda: h.sum(dtype=torch.float32))
print(json.dumps({"read_only_sum_GBps": x.numel() * 4 / ms_sum / 1e6, "copy_rw_GBps": 2 * x.numel() * 4 / ms_cp / 1e6,
                  "read_only_sum_f16view_GBps": x.numel() * 4 / ms_sum16 / 1e6}))
