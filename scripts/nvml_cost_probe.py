#!/usr/bin/env python
"""How long do the NVML queries of the clock sampler take, idle and while kernels run?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = torch.device("cuda:0")
x = torch.randn(64, 1024, 1024, device=dev)
def cost(fn, n=100):
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    ts.sort(); return f"median {ts[n//2]*1e6:7.1f} us  p90 {ts[int(n*.9)]*1e6:7.1f} us  max {ts[-1]*1e6:7.1f} us"
fns = {"clock": lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
       "reasons": lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h),
       "power": lambda: pynvml.nvmlDeviceGetPowerUsage(h)}
for name, fn in fns.items(): print("idle ", name, cost(fn))
for _ in range(200): y = x * 1.0001
for name, fn in fns.items():
    for _ in range(300): y = x * 1.0001
    print("busy ", name, cost(fn))
torch.cuda.synchronize()
# launch-rate impact: time to enqueue 2000 tiny kernels with / without a polling thread
import threading
z = torch.zeros(8, device=dev)
def enqueue(n=3000):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): z.add_(1.0)
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e6
print("launch us, no poller:", round(enqueue(), 2))
for period, what in ((0.005, ("clock", "reasons")), (0.005, ("clock",)), (0.025, ("clock", "reasons"))):
    on = True
    def loop():
        while on:
            for w in what: fns[w]()
            time.sleep(period)
    th = threading.Thread(target=loop, daemon=True); th.start()
    print(f"launch us, poller {what} every {period*1e3:.0f} ms:", round(enqueue(), 2))
    on = False; th.join()
