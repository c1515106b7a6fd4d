"""Which NVML poll perturbs the estimator loop?  Per-step wall times of the C2
workload under different samplers (development probe, not a bench)."""
import sys, time, threading
sys.path.insert(0, '.')
import numpy as np, torch, pynvml
from triumvirate_b200 import core

n, ng, nb, L = 10**7, 512, 20, 1000.
pos = np.random.default_rng(42).uniform(0., L, size=(3, n))
d = torch.from_numpy(pos).to('cuda:0'); torch.cuda.synchronize()
kw = dict(boxsize=L, ngrid=ng, assignment='pcs', degrees=(0, 0, 0), form='full',
          bin_range=(0.005, 0.205), num_bins=nb, norm_factor=1.)
def step():
    return core.threept_box_arrays('bispec', n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), True, **kw)
for _ in range(4): step()
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
def run(label, fn, period):
    stop = threading.Event(); cnt = [0]; tcall = []
    def loop():
        while not stop.is_set():
            t = time.perf_counter(); fn(); tcall.append(time.perf_counter() - t); cnt[0] += 1
            stop.wait(period)
    th = None
    if fn is not None:
        th = threading.Thread(target=loop, daemon=True); th.start()
    ts = []
    for _ in range(100):
        t = time.perf_counter(); step(); ts.append((time.perf_counter() - t) * 1e3)
    if th: stop.set(); th.join()
    print(f"{label:40s} mean {np.mean(ts):6.2f} med {np.median(ts):6.2f} max {np.max(ts):6.2f} ms; polls {cnt[0]}, "
          f"poll ms mean {1e3*np.mean(tcall) if tcall else 0:.2f}")
clk = lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
rsn = lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
both = lambda: (clk(), rsn())

alt_state = [0]
def alt():
    alt_state[0] ^= 1
    return clk() if alt_state[0] else rsn()
run("no sampler", None, 0)
run("clock only / 20 ms", clk, 0.02)
run("reasons only / 20 ms", rsn, 0.02)
run("alternating / 20 ms", alt, 0.02)
run("both / 20 ms", both, 0.02)
run("both / 50 ms", both, 0.05)
run("clock only / 50 ms", clk, 0.05)
run("no sampler (again)", None, 0)
# main-thread polling between steps
ts = []
for i in range(100):
    t = time.perf_counter(); step()
    if i % 5 == 0: clk()
    rsn()
    ts.append((time.perf_counter() - t) * 1e3)
print(f"main-thread polls: mean {np.mean(ts):.2f} med {np.median(ts):.2f} max {np.max(ts):.2f}")
