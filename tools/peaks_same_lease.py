"""The roofline denominators measured in ONE lease, so that the int8 tensor-pipe peak of the product's roofline can be tied to the
driver's MEASURED_PEAKS.json method (torch bf16 8192^3): tcgen05.mma kind::i8 probe (tools/qf_probe.cu), torch bf16 matmul burst and
sustained, cuBLAS DGEMM, and a device copy (HBM).   usage: python tools/peaks_same_lease.py out.json"""
import json, os, re, subprocess, sys, time
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
res = {"gpu": torch.cuda.get_device_name(0)}
exe = os.path.join(HERE, "qf_probe")
if not os.path.exists(exe):
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", exe, os.path.join(HERE, "qf_probe.cu")], check=True)
out = subprocess.run([exe], capture_output=True, text=True).stdout
res["qf_probe_stdout"] = out.strip().splitlines()
tops = [float(m.group(1)) for m in re.finditer(r"([0-9.]+) TOPS int8", out)]
res["umma_i8_tops"] = max(tops) if tops else None

def gemm(dtype, n, secs):
    a = torch.randn(n, n, device="cuda").to(dtype); b = torch.randn(n, n, device="cuda").to(dtype)
    for _ in range(3): a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.time(); cnt = 0
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(10): a @ b
        cnt += 10
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return 2 * n ** 3 / best * 1e-9, 2 * n ** 3 * cnt / e0.elapsed_time(e1) * 1e-9

res["bf16_tflops_burst"], res["bf16_tflops_sustained"] = gemm(torch.bfloat16, 8192, 4.0)
res["fp64_tflops_burst"], res["fp64_tflops_sustained"] = gemm(torch.float64, 8192, 3.0)
a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
best = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
res["hbm_copy_gbs"] = 2 * a.numel() * 2 / best * 1e-6
res["i8_over_2x_bf16_burst"] = res["umma_i8_tops"] / (2 * res["bf16_tflops_burst"]) if res["umma_i8_tops"] else None
print(json.dumps(res, indent=1))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
