"""Per-step times of the end-to-end leg of bench.py (c2): is a slow mean one outlier step or all of them?"""
import sys, time, torch
sys.path.insert(0, '.')
import bench
torch.backends.cudnn.allow_tf32 = False
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'c2']
st = bench.Stepper(w, torch.float32, torch.device('cuda:0'))
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(3): st.step_device()
torch.cuda.synchronize()
for leg, fn in (('device', st.step_device), ('e2e', st.step_e2e), ('device', st.step_device), ('e2e', st.step_e2e)):
    times, walls = [], []
    for i in range(8):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        walls.append((time.time() - t0) * 1e3)
        times.append(e0.elapsed_time(e1))
    print(leg, 'event ms', [round(t, 1) for t in times], 'wall ms', [round(t, 1) for t in walls], flush=True)
