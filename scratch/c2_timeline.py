"""GPU timeline of one c2 step (torch.profiler / CUPTI): busy time, idle gaps and where they are."""
import sys, json, torch
sys.path.insert(0, '.')
import bench
from torch.profiler import profile, ProfilerActivity
w = bench.WORKLOADS['c2']
st = bench.Stepper(w, torch.float32, torch.device('cuda:0'))
for _ in range(4): st.step_device()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    st.step_device(); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
busy = 0.0; cur_end = t0; gaps = []
for e in ev:
    s, en = e.time_range.start, e.time_range.end
    if s > cur_end:
        gaps.append((s - cur_end, cur_end - t0, e.name[:60]))
    if en > cur_end:
        busy += en - max(s, cur_end); cur_end = en
print(f"span {(t1-t0)/1e3:.2f} ms, busy {busy/1e3:.2f} ms, idle {(t1-t0-busy)/1e3:.2f} ms, kernels {len(ev)}")
gaps.sort(reverse=True)
print("largest gaps (us, at ms, next kernel):")
for g, at, name in gaps[:25]: print(f"  {g:8.1f} us at {at/1e3:7.2f} ms before {name}")
# idle by phase: before / during / after the solver
names = [(e.time_range.start - t0, e.name) for e in ev]
first_solver = next((t for t, n in names if 'onesided' in n or 'chol' in n), None)
last_solver = max((e.time_range.end - t0 for e in ev if 'onesided' in e.name or 'jacobi' in e.name or 'sweep_end' in e.name), default=None)
print('solver window', first_solver, last_solver)
tot = {}
for e in ev:
    k = e.name.split('<')[0].split('(')[0][:50]
    tot[k] = tot.get(k, 0) + (e.time_range.end - e.time_range.start)
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:25]: print(f"  {v/1e3:7.3f} ms {k}")
