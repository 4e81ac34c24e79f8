import time, pynvml as nv, torch
nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
def t(f, n=20):
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3
print("idle  : clock %.3f ms  reasons %.3f ms" % (t(lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), t(lambda: nv.nvmlDeviceGetCurrentClocksEventReasons(h))))
hbuf = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); d = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
for _ in range(200): d.copy_(hbuf, non_blocking=True)
print("copying: clock %.3f ms  reasons %.3f ms" % (t(lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), t(lambda: nv.nvmlDeviceGetCurrentClocksEventReasons(h))))
torch.cuda.synchronize()
