import torch, time
for mb in (8, 64, 256, 1024):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    t0 = time.perf_counter()
    for _ in range(10): h2.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt2 = (time.perf_counter() - t0) / 10
    print(mb, "MB  H2D %.1f GB/s  D2H %.1f GB/s" % (n / dt / 1e9, n / dt2 / 1e9))
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max", "--format=csv"], capture_output=True, text=True).stdout)
