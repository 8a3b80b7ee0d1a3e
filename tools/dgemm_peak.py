# cuBLAS DGEMM peak on this box (FP64 roofline denominator). Scratch tool; result copied to profiles/.
import torch, json, time
torch.backends.cuda.matmul.allow_tf32 = False
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(3): c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(8):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}_tflops_burst"] = 2 * n**3 / best * 1e-9
    # sustained 3 s
    t0 = time.time(); cnt = 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(5): c = a @ b
        cnt += 5
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    res[f"dgemm_{n}_tflops_sustained"] = 2 * n**3 * cnt / e0.elapsed_time(e1) * 1e-9
print(json.dumps(res))
