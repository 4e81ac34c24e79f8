#!/usr/bin/env python
"""Where does the fp32 e2e path lose PCIe bandwidth?  Pure pinned H2D copies of one episode's bytes vs the pipeline."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tap-net_b200"))
import torch
dev = torch.device("cuda:0")
def bw(nbytes, reps, stream=None, label=""):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print("%-28s %8.2f MB x %4d : %6.2f GB/s (events)  %6.2f GB/s (wall)" % (label, nbytes / 1e6, reps, nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9, nbytes * reps / wall / 1e9))
bw(64 << 20, 8, label="64 MiB")
bw(11141120, 50, label="one C2 episode (11.1 MB)")
bw(11141120, 50, label="again")
bw(1 << 20, 200, label="1 MiB")
# with a background thread polling NVML every 2 ms (what bench.py's clock sampler used to be)
import threading
import bench
s = bench.ClockSampler(0)
stop = []
def poll():
    while not stop:
        s.sample(); time.sleep(0.002)
th = threading.Thread(target=poll, daemon=True); th.start()
bw(11141120, 50, label="11.1 MB + NVML thread")
bw(64 << 20, 8, label="64 MiB + NVML thread")
stop.append(1); th.join()
print(s.result())
