import torch, time
x = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for _ in range(2): y.copy_(x, non_blocking=True); torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print("pinned H2D: %.1f GB/s" % (1.073741824 / dt))
import subprocess; print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv,noheader", shell=True, capture_output=True, text=True).stdout.strip())
