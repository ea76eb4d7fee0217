"""Per-(entry point, shape) table of one bench workload: ms per call, achieved GB/s or TFLOP/s of algorithmic work.
    python scratch/per_call.py [workload] [steps]"""
import sys, collections, torch
sys.path.insert(0, '.')
import bench
from vivit_b200 import kernels
name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
st = bench.Stepper(bench.WORKLOADS[name], torch.float32, torch.device('cuda:0'))
for _ in range(3): st.step_device()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
kernels.timing_start()
for _ in range(steps):
    flush.fill_(1); st.step_device()
recs = kernels.timing_stop()
agg = collections.OrderedDict()
for n, ms, shp, nl, oshp in recs:
    a = agg.setdefault((n, tuple(map(tuple, shp))), [0.0, 0, oshp])
    a[0] += ms; a[1] += 1
rows = []
for (n, shp), (ms, c, oshp) in agg.items():
    kind, amount = bench.algorithmic_work(n, [list(s) for s in shp], 4, oshp) or (None, 0)
    per = ms / c
    ach = amount / (per * 1e-3) / (1e9 if kind == 'hbm' else 1e12) if kind else 0.0
    rows.append((ms / steps, n, shp, c / steps, per, kind, ach))
for tot, n, shp, c, per, kind, ach in sorted(rows, reverse=True)[:45]:
    unit = 'GB/s' if kind == 'hbm' else 'TFLOP/s' if kind else ''
    print(f"{tot:7.3f} ms/step {n:28s} x{c:4.1f} {per*1e3:8.1f} us/call {ach:8.1f} {unit:8s} {list(shp)}")
