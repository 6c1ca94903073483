"""Read-only HBM stream rate on this GPU (library reductions over a buffer far larger than L2), beside the copy rate of
MEASURED_PEAKS.json (read + write bytes).  Context for the roofline of the GLM stream kernel, which only reads."""
import json, torch
out = {}
n = 1 << 28                      # 2 GiB of float64
x = torch.ones(n, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
def best(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return min(t)
ms = best(lambda: x.sum())
out["sum_f64_2GiB"] = {"ms": ms, "read_GBps": 8.0 * n / (ms * 1e-3) / 1e9}
ms = best(lambda: torch.dot(x, x))
out["dot_f64_2GiB_same_buffer"] = {"ms": ms, "read_GBps": 8.0 * n / (ms * 1e-3) / 1e9}
xf = x.view(torch.float32)
ms = best(lambda: xf.sum())
out["sum_f32_2GiB"] = {"ms": ms, "read_GBps": 8.0 * n / (ms * 1e-3) / 1e9}
ms = best(lambda: y.copy_(x))
out["copy_f64_2GiB"] = {"ms": ms, "read_plus_write_GBps": 16.0 * n / (ms * 1e-3) / 1e9}
print(json.dumps(out))
