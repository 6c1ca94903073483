"""Library INT8 tensor-core GEMM rate on this GPU (torch._int_mm -> cuBLASLt), as the yardstick for an FP64-accurate
sliced-integer (Ozaki-scheme) GLM contraction: DESIGN.md section 8.  Not part of the product path."""
import json, sys, torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
a = torch.randint(-127, 127, (n, n), dtype=torch.int8, device="cuda")
b = torch.randint(-127, 127, (n, n), dtype=torch.int8, device="cuda")
for _ in range(3):
    torch._int_mm(a, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    torch._int_mm(a, b)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(json.dumps({"op": "torch._int_mm int8 x int8 -> int32", "n": n, "ms": ms, "tera_ops_per_s": 2.0 * n ** 3 / (ms * 1e-3) / 1e12}))
