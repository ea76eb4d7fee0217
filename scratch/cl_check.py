import os, sys, time, torch
sys.path.insert(0, '.')
import vivit_b200.kernels as k
torch.manual_seed(0)
def run(R, rank, dt, cls, decay=-3):
    B = torch.randn(R, rank, dtype=torch.float64, device='cuda') * torch.logspace(0, decay, rank, dtype=torch.float64, device='cuda')
    G = (B @ B.t()).to(dt)
    want = torch.linalg.eigvalsh(G.double())
    for cl in cls:
        if cl is None:
            os.environ.pop('VVT_SYEVJ_CL', None)
        else:
            os.environ['VVT_SYEVJ_CL'] = str(cl)
        k.syevj(G, True); torch.cuda.synchronize()
        t0 = time.time(); ev, U = k.syevj(G, True); torch.cuda.synchronize(); ms = (time.time() - t0) * 1e3
        err = (ev.double() - want).abs().max().item() / want.abs().max().item()
        print(f"R={R} {dt} CL={cl}: {ms:.1f} ms {k.last_syevj_info} evalerr={err:.2e}", flush=True)
    os.environ.pop('VVT_SYEVJ_CL', None)
f32, f64 = torch.float32, torch.float64
if len(sys.argv) > 1:
    run(1280, 1152, f32, [None]); run(640, 576, f32, [None]); run(320, 288, f32, [None]); run(1280, 1152, f64, [None])
    sys.exit(0)
run(5120, 4608, f32, [None])
run(3840, 3456, f32, [None])
run(2560, 2304, f32, [None])
run(1920, 1728, f32, [None, 3, 5])
run(1600, 1440, f32, [None, 3, 4])
run(2560, 2304, f64, [None, 4, 6])
run(1920, 1728, f64, [None, 3, 5, 6])
run(1280, 1152, f64, [None])
run(320, 288, f32, [None, 2, 3, 4])
run(640, 576, f32, [None, 4])
