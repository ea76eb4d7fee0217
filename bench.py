#!/usr/bin/env python
"""Benchmark of the low-rank GGN hot path (BASELINE.json: "ms per GGN eigh step + Gram TFLOP/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--dtype f32]
    python bench.py --impl reference ...      # the reference's algorithm on the host cores

A *step* is one pass of the hot path over one batch of synthetic input.  The default
workload is BASELINE ``configs[1]``: cifar10_3c3d, N=128, C=10, exact GGN --
``EighComputation`` (top-10 eigenpairs) followed by ``DirectionalDerivativesComputation``
(top-10 directions): two forward/backward passes with extension + hook through
``get_result``, as in the reference.

Prints ONE JSON line (rank 0).  ``value`` is device-timed with inputs resident in HBM;
``e2e`` is the same step through the public API with host buffers (pinned H2D of the batch,
D2H of every result) inside the timed region.  At N>1 the parameter dimension is sharded
over the ranks (``process_group=``) and the partial Grams are summed with one NCCL
all-reduce: the total work is fixed, so ``scaling`` is "strong".
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch import nn  # noqa: E402

# torch's own convolutions (the model's forward/backward, not part of the hot path) default to TF32;
# the metric is quoted for fp32 arithmetic, so they are pinned to fp32 for both arms
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

METRIC = "ms_per_ggn_curvature_step"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


# --------------------------------------------------------------------------
# workloads (BASELINE.json configs; model shapes from DeepOBS)
# --------------------------------------------------------------------------


def mlp_c1():
    return nn.Sequential(nn.Linear(784, 64), nn.ReLU(), nn.Linear(64, 32), nn.ReLU(), nn.Linear(32, 10))


def cnn_3c3d():
    """cifar10_3c3d: tf 'same' 3x3/2 pooling == ceil-mode pooling on post-ReLU maps."""
    return nn.Sequential(
        nn.Conv2d(3, 64, 5), nn.ReLU(), nn.MaxPool2d(3, 2, ceil_mode=True),
        nn.Conv2d(64, 96, 3), nn.ReLU(), nn.MaxPool2d(3, 2, ceil_mode=True),
        nn.Conv2d(96, 128, 3, padding=1), nn.ReLU(), nn.MaxPool2d(3, 2, ceil_mode=True),
        nn.Flatten(), nn.Linear(1152, 512), nn.ReLU(), nn.Linear(512, 256), nn.ReLU(),
        nn.Linear(256, 10),
    )


def allcnnc():
    """cifar100_allcnnc without dropout: 9 convolutions + global average pooling."""
    def c(i, o, k, s=1, p=0):
        return [nn.Conv2d(i, o, k, stride=s, padding=p), nn.ReLU()]
    layers = (
        c(3, 96, 3, p=1) + c(96, 96, 3, p=1) + c(96, 96, 3, s=2, p=1)
        + c(96, 192, 3, p=1) + c(192, 192, 3, p=1) + c(192, 192, 3, s=2, p=1)
        + c(192, 192, 3) + c(192, 192, 1) + c(192, 100, 1)
    )
    return nn.Sequential(*layers, nn.AvgPool2d(6), nn.Flatten())


def mlp_c4():
    return nn.Sequential(
        nn.Linear(784, 4096), nn.ReLU(), nn.Linear(4096, 4096), nn.ReLU(),
        nn.Linear(4096, 4096), nn.ReLU(), nn.Linear(4096, 10),
    )


def mlp_c5():
    """config 5: a deep MLP whose full-network Gram has R = C N = 10240."""
    layers, width = [nn.Linear(784, 2048), nn.ReLU()], 2048
    for _ in range(3):
        layers += [nn.Linear(width, width), nn.ReLU()]
    return nn.Sequential(*layers, nn.Linear(width, 10))


def top_k(k):
    return lambda ev: list(range(max(0, ev.numel() - k), ev.numel()))


def const_damping(evals, evecs, gammas, lambdas):
    return torch.ones_like(evals)


WORKLOADS = {
    "c1": dict(name="mlp_784-64-32-10 N=32 C=10 EigvalshComputation exact, one group",
               model=mlp_c1, n=32, in_shape=(784,), classes=10, calls=("eigvalsh",), grouping="one"),
    "c2": dict(name="cifar10_3c3d N=128 C=10 EighComputation top-10 + DirectionalDerivativesComputation, exact GGN, one group",
               model=cnn_3c3d, n=128, in_shape=(3, 32, 32), classes=10, calls=("eigh", "dirderiv"), grouping="one"),
    "c3": dict(name="cifar100_allcnnc N=128 C=100 mc_samples=1 subsampling_ggn=[0..31] DirectionalDampedNewtonComputation",
               model=allcnnc, n=128, in_shape=(3, 32, 32), classes=100, calls=("newton",), grouping="one",
               mc=1, sub_ggn=list(range(32))),
    "c4": dict(name="mlp_784-4096x3-10 N=512 C=10 EighComputation top-10, per-layer block-diagonal groups (one batched solve)",
               model=mlp_c4, n=512, in_shape=(784,), classes=10, calls=("eigh",), grouping="layer"),
    "c5": dict(name="deep_mlp_784-2048x4-10 N=1024 C=10 EighComputation top-10, full-network Gram R=10240 "
                    "(parameter-sharded over the ranks, one all-reduce)",
               model=mlp_c5, n=1024, in_shape=(784,), classes=10, calls=("eigh",), grouping="one"),
}


def make_groups(model, grouping):
    extra = {"criterion": top_k(10), "damping": const_damping}
    if grouping == "one":
        return [{"params": list(model.parameters()), **extra}]
    return [{"params": list(m.parameters()), **extra} for m in model if list(m.parameters())]


def make_problem(w, dtype):
    torch.manual_seed(0)
    model = w["model"]().to(dtype)
    x = torch.rand(w["n"], *w["in_shape"]).to(dtype)
    y = torch.randint(0, w["classes"], (w["n"],))
    return model, x, y


# --------------------------------------------------------------------------
# the step through the public API
# --------------------------------------------------------------------------


class Stepper:
    def __init__(self, w, dtype, device, process_group=None):
        from vivit_b200 import extend

        self.w, self.device, self.pg = w, device, process_group
        self.calls = tuple(w["calls"])
        self.batch_solves = True
        model, x, y = make_problem(w, dtype)
        self.model = extend(model.to(device))
        self.loss_fn = extend(nn.CrossEntropyLoss())
        self.groups = make_groups(self.model, w["grouping"])
        self.parallelism = "single GPU"
        if process_group is not None:
            import torch.distributed as dist

            world = dist.get_world_size(process_group)
            rank = dist.get_rank(process_group)
            self.calls = tuple(w["calls"])
            if w["grouping"] != "layer" and len(w["calls"]) == 2:
                # the step is two independent computations (two backward passes in the reference):
                # they run side by side on the two halves of the ranks; a half of more than one rank shards the
                # parameter dimension of its pass.  (The eigensolver of a pass is latency-bound at R = 1280 and
                # does not shard: without this split every rank would repeat both solves.)
                half = world // 2
                halves = [dist.new_group(list(range(half))), dist.new_group(list(range(half, world)))]
                mine = 0 if rank < half else 1
                self.calls = (w["calls"][mine],)
                size = half if mine == 0 else world - half
                self.pg = halves[mine] if size > 1 else None
                self.parallelism = (
                    f"{w['calls'][0]} on ranks 0..{half - 1} next to {w['calls'][1]} on ranks {half}..{world - 1}; "
                    f"inside a half the Gram is parameter-sharded (one NCCL all-reduce per group), results stay on "
                    f"the half that computed them")
            elif w["grouping"] == "layer":
                # block-diagonal groups are independent: whole groups per rank, no collective
                # (more ranks than groups: a team of ranks per group, Gram parameter-sharded and eigensolver rounds
                # distributed inside the team)
                from vivit_b200.dist import team_groups

                n_groups = len(self.groups)
                self.groups, self.pg, teams = team_groups(self.groups, process_group)
                if world <= n_groups:
                    self.parallelism = f"{n_groups} block-diagonal groups assigned to {world} ranks, no collective"
                else:
                    self.parallelism = (
                        f"{n_groups} block-diagonal groups on teams of {[len(t) for t in teams]} ranks: inside a team "
                        f"the Gram is parameter-sharded and the eigensolver rounds are distributed (vvt_syevj_dist); "
                        f"no collective between teams")
            else:
                self.parallelism = (
                    f"parameter-sharded Gram over {world} ranks, one all-reduce per group; the block pairs of every "
                    f"round of the two-level eigensolver (fp32, R >= 4096) distributed over the same ranks, blocks "
                    f"changing owner through peer memory (vvt_syevj_dist)")
        self.x_host, self.y_host = x.pin_memory(), y.pin_memory()
        self.x, self.y = x.to(device), y.to(device)
        self.host_out = None
        self.copy_stream = None
        self.h2d_bytes = x.numel() * x.element_size() + y.numel() * y.element_size()
        self.d2h_bytes = 0
        self.mc_ids = None
        if w.get("mc"):
            g = torch.Generator().manual_seed(1)
            self.mc_ids = torch.randint(0, w["classes"], (w["mc"], len(w["sub_ggn"])), generator=g).to(device)

    def _backward(self, comp, x, y):
        from vivit_b200 import backpack

        hook = comp.get_extension_hook(self.groups)
        exts = comp.get_extensions()
        with backpack(*exts, extension_hook=hook):
            self.loss_fn(self.model(x), y).backward()
        for p in self.model.parameters():
            p.grad = None

    def _pass(self, comp, x, y):
        self._backward(comp, x, y)
        return [comp.get_result(g) for g in self.groups]

    def run(self, x, y, on_results=None):
        """One step on device-resident ``x, y``; returns the flat list of result tensors.  ``on_results`` is
        called with the tensors of every computation as soon as that computation has produced them."""
        import vivit_b200 as vv

        w, out = self.w, []
        # sharded runs return the WHOLE eigenvectors / steps (all-gathered), the same deliverable as on one GPU
        kw = {"process_group": self.pg} if self.pg is not None else {}
        gkw = {**kw, "gather": True} if self.pg is not None else {}
        # several computations of one step (c2: eigenpairs, then directional derivatives -- two backward passes, as
        # in the reference) share a SolveQueue: their Gram matrices are decomposed by ONE batched solver call when
        # the first result is read.  --sequential-solves restores the reference's order (solve inside each hook).
        queue = vv.SolveQueue() if self.batch_solves and len(self.calls) > 1 else None
        qkw = {"solve_queue": queue} if queue is not None else {}
        comps = []
        for call in self.calls:
            if call == "eigvalsh":
                comp = vv.EigvalshComputation(**kw)
            elif call == "eigh":
                comp = vv.EighComputation(**gkw, **qkw)
            elif call == "dirderiv":
                comp = vv.DirectionalDerivativesComputation(**kw, **qkw)
            elif call == "newton":
                comp = vv.DirectionalDampedNewtonComputation(
                    subsampling_ggn=w.get("sub_ggn"), mc_samples_ggn=w.get("mc", 0), **gkw)
                comp._mc_state = self.mc_ids  # class ids pre-sampled on the host
            self._backward(comp, x, y)
            comps.append((call, comp))
            if queue is None:
                out += self._collect(call, comp, on_results)
        if queue is not None:
            for call, comp in comps:
                out += self._collect(call, comp, on_results)
        return out

    def _collect(self, call, comp, on_results):
        got = []
        for res in (comp.get_result(g) for g in self.groups):
            if call == "eigvalsh":
                got.append(res)
            elif call == "eigh":
                got += [res[0], *res[1]]
            else:  # (gammas, lambdas) / Newton steps per parameter
                got += list(res)
        if on_results is not None:
            on_results(got)
        return got

    def step_device(self):
        return self.run(self.x, self.y)

    def step_e2e(self):
        """The step through the public API with HOST buffers: pinned H2D of the batch, D2H of every result.  The
        results of a computation start their way to the host (side stream, pinned buffers) as soon as that
        computation is done, i.e. while the next one runs; the step ends when every copy has landed."""
        x = self.x_host.to(self.device, non_blocking=True)
        y = self.y_host.to(self.device, non_blocking=True)
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(device=self.device)
        compute, host, first = torch.cuda.current_stream(), [], self.host_out is None

        def ship(tensors):
            ready = torch.cuda.Event()
            ready.record(compute)
            self.copy_stream.wait_event(ready)
            with torch.cuda.stream(self.copy_stream):
                for t in tensors:
                    h = torch.empty(t.shape, dtype=t.dtype).pin_memory() if first else self.host_out[len(host)]
                    h.copy_(t, non_blocking=True)  # `t` stays referenced by run()'s result list until the copies have landed
                    host.append(h)

        self.run(x, y, on_results=ship)
        self.copy_stream.synchronize()
        compute.synchronize()
        if first:
            self.host_out = host
            self.d2h_bytes = sum(h.numel() * h.element_size() for h in host)
        return host


def pcie_bandwidth(device):
    """Pinned-memory copy bandwidth of this box (GB/s, 64 MiB, best of 3): the end-to-end leg moves 36 MB of
    eigenvectors per step at c2, and boxes of the pool differ by more than an order of magnitude here."""
    n = 64 << 20
    dev, host = torch.empty(n, dtype=torch.uint8, device=device), torch.empty(n, dtype=torch.uint8).pin_memory()
    out = {}
    for name, (dst, src) in {"d2h_gbs": (host, dev), "h2d_gbs": (dev, host)}.items():
        best = 0.0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        out[name] = round(best, 1)
    return out


# --------------------------------------------------------------------------
# roofline bookkeeping
# --------------------------------------------------------------------------


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                p = json.load(f)
            return {
                "hbm_gbs": float(p.get("hbm_gbs", FALLBACK_PEAKS["hbm_gbs"])),
                "bf16_tflops": float(p.get("bf16_tflops", FALLBACK_PEAKS["bf16_tflops"])),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 0.0))),
            }, "measured"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), "fallback"


def algorithmic_work(name, shapes, es, out_shapes=()):
    """(kind, amount): algorithmic flops ("tensor") or bytes ("hbm") of one call of a kernel
    entry point, from the shapes of its tensor arguments and results.
    Symmetric-aware minimum for the Gram kernels, one read of each operand + one write of each
    result for the bandwidth-bound ones."""
    prod = lambda s: int(torch.Size(s).numel())  # noqa: E731
    if name == "gram_dense_accum":
        (R, _), (_, D) = shapes[0], shapes[1]
        return "tensor", R * (R + 1) * D
    if name == "gram_cross_accum":
        (R, ng), (_, D) = shapes[0], shapes[1]
        return "tensor", 2 * R * ng * D
    if name == "gram_linear_accum":
        (R, _), (C, N, out), (_, n_in) = shapes[0], shapes[1], shapes[2]
        return "tensor", R * (R + 1) * out + N * (N + 1) * n_in + R * (R + 1) // 2
    if name == "gram_cross_linear_accum":
        (R, ng), (C, N, out), (_, n_in) = shapes[0], shapes[1], shapes[2]
        return "tensor", 2 * R * ng * out + 2 * N * ng * n_in + R * ng
    if name == "sqrt_backprop_linear":
        S, W = shapes[0], shapes[1]
        return "tensor", 2 * (prod(S) // W[0]) * W[0] * W[1]
    if name == "sqrt_backprop_conv2d":
        S, W = shapes[0], shapes[1]  # [V,N,Co,Ho,Wo], [Co,Ci,kh,kw]
        return "tensor", 2 * prod(S) * W[1] * W[2] * W[3]
    if name == "v_emit_conv2d":
        S, X = shapes[0], shapes[1]  # [V,N,Co,Ho,Wo], [N,Ci,H,W] -> factor [V,N,Co,Ci,kh,kw] (written once)
        return "hbm", es * (prod(S) + prod(X) + sum(prod(o) for o in out_shapes))
    if name == "v_emit_bias":
        return "hbm", es * (prod(shapes[0]) + sum(prod(o) for o in out_shapes))
    if name in ("backtransform_linear", "v_apply_linear"):
        # E[k,o,i] = sum_n (sum_c U[k,c,n] S[c,n,o]) Z[n,i]: two products, K (C N out + N out in) MACs
        U, S, Z = shapes[0], shapes[1], shapes[2]
        C, N, n_out = S
        K = max(1, prod(U) // (C * N))
        return "tensor", 2 * K * N * n_out * (C + Z[1])
    if name == "center_rows":
        return "hbm", es * 2 * prod(shapes[0])
    if name in ("backtransform_dense", "v_apply_dense"):
        U, V = shapes[0], shapes[1]
        K = prod(U) // V[0]
        return "hbm", es * (prod(V) + prod(U) + K * V[1])
    if name in ("sqrt_backprop_elementwise", "sqrt_backprop_maxpool2d", "sqrt_backprop_avgpool2d"):
        # one read of the incoming factor, one write of the outgoing one (pools: the larger, un-pooled map)
        out_elems = sum(prod(o) for o in out_shapes) or prod(shapes[0])
        return "hbm", es * (prod(shapes[0]) + out_elems)
    if name == "scale_rows_rsqrt":
        return "hbm", es * 2 * prod(shapes[0])
    return None, 0


def lookup_traffic(name, shapes):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the roofline kernel from the
    committed ``ncu --set full`` capture, if one exists for this entry point and shape (profiles/ncu_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            table = json.load(f)
    except Exception:
        return None
    key = name + ":" + "x".join(",".join(str(d) for d in s_) for s_ in shapes)
    entry = table.get(key)
    return entry["dram_bytes_per_launch"] if entry else None


def eigensolver_report(stepper, batch=1):
    """The per-group Gram of the workload solved by ``vvt_syevj`` and, beside it, by cuSOLVER through
    ``torch.linalg.eigh`` on the same GPU: ms (CUDA events, 3 repetitions), sweeps, residuals."""
    import vivit_b200 as vv
    from vivit_b200 import kernels

    grabbed = []
    orig, orig_b = kernels.syevj, kernels.syevj_batched

    def spy(G, vectors=True, **kw):
        grabbed.append(G.clone())
        return orig(G, vectors, **kw)

    def spy_b(G, vectors=True, **kw):
        grabbed.extend(g.clone() for g in G)
        return orig_b(G, vectors, **kw)

    kernels.syevj, kernels.syevj_batched = spy, spy_b
    try:
        call = stepper.w["calls"][0]
        comp = {"eigvalsh": vv.EigvalshComputation, "eigh": vv.EighComputation,
                "dirderiv": vv.DirectionalDerivativesComputation}.get(call)
        if comp is None:
            return None
        stepper._pass(comp(), stepper.x, stepper.y)
    finally:
        kernels.syevj, kernels.syevj_batched = orig, orig_b
    if not grabbed:
        return None
    G = max(grabbed, key=lambda t: t.shape[0])
    R = G.shape[0]

    def ms_of(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, out

    ours_ms, (ev, U, sinfo) = ms_of(lambda: orig(G, True, return_info=True))
    sweeps = sinfo["sweeps"]
    lib_ms, (wv, _) = ms_of(lambda: torch.linalg.eigh(G))
    lib_vals_ms, _ = ms_of(lambda: torch.linalg.eigvalsh(G))
    batched = None
    if batch > 1:  # what the shared SolveQueue of the step asks for: `batch` such matrices in one call
        GG = torch.stack([G] * batch).contiguous()
        b_ms, _ = ms_of(lambda: orig_b(GG, True, return_info=True))
        batched = {"batch": batch, "ms": round(b_ms, 3)}
    Gd, Ud, evd = G.double(), U.double(), ev.double()
    want = torch.linalg.eigvalsh(Gd)
    scale = want.abs().max().item() or 1.0
    return {
        "R": R, "ms": round(ours_ms, 3), "sweeps": sweeps,
        "eigenvalue_error_rel_max": float((evd - want).abs().max().item() / scale),
        "residual_fro": float(((Gd @ Ud - Ud * evd[None]).norm() / Gd.norm()).item()),
        "orthogonality_max_abs": float((Ud.t() @ Ud - torch.eye(R, dtype=torch.float64, device=G.device)).abs().max().item()),
        "cusolver_eigh_ms": round(lib_ms, 3), "cusolver_eigvalsh_ms": round(lib_vals_ms, 3), "batched": batched,
        "note": "vvt_syevj: Cholesky-preconditioned one-sided block Jacobi + one refinement step, all on the GPU",
    }


def summarize_kernels(records, steps, es, peaks):
    agg = {}
    for name, ms, shapes, launches, out_shapes in records:
        a = agg.setdefault(name, {"ms": 0.0, "calls": 0, "launches": 0, "tensor": 0, "hbm": 0})
        a["ms"] += ms
        a["calls"] += 1
        a["launches"] += launches
        kind, amount = algorithmic_work(name, shapes, es, out_shapes)
        if kind:
            a[kind] += amount
    total = sum(a["ms"] for a in agg.values()) or 1.0
    rows = []
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        row = {"kernel": name, "ms_per_step": round(a["ms"] / steps, 4), "share": round(a["ms"] / total, 4),
               "calls_per_step": a["calls"] / steps, "launches_per_step": a["launches"] / steps}
        sec = a["ms"] / 1e3
        if a["tensor"] and sec > 0:
            row.update(bound="tensor", achieved=round(a["tensor"] / sec / 1e12, 3), unit="TFLOP/s")
        elif a["hbm"] and sec > 0:
            row.update(bound="hbm", achieved=round(a["hbm"] / sec / 1e9, 1), unit="GB/s")
        rows.append(row)
    return rows


# --------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------

_CLOCK_FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")


class ClockSampler:
    def __init__(self, index):
        self.proc, self.path = None, f"/tmp/vvt_clocks_{os.getpid()}.csv"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={_CLOCK_FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for n, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------
# the reference's algorithm on the host cores (oracle; the one place bench.py runs oracle/)
# --------------------------------------------------------------------------


def reference_step(w, model, x, y):
    from oracle import reference_path as ref

    loss = nn.CrossEntropyLoss()
    groups = make_groups(model, w["grouping"])
    for call in w["calls"]:
        if call == "eigvalsh":
            ref.eigvalsh(model, loss, x, y, groups)
        elif call == "eigh":
            ref.eigh(model, loss, x, y, groups)
        elif call == "dirderiv":
            ref.directional_derivatives(model, loss, x, y, groups)
        elif call == "newton":
            ref.directional_damped_newton(model, loss, x, y, groups, None, w.get("sub_ggn"),
                                          mc_samples_ggn=w.get("mc", 0))


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm runs on rank 0 alone and may use every core."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def time_reference(w, dtype, steps, warmup, budget_s):
    """Median ms per step of the oracle on the host cores; stops early once ``budget_s`` of CPU
    time is spent (each step is the full workload, the sample is the number of steps)."""
    use_all_host_threads()
    model, x, y = make_problem(w, dtype)
    times, t_all = [], time.time()
    for i in range(warmup + steps):
        t0 = time.time()
        reference_step(w, model, x, y)
        dt = (time.time() - t0) * 1e3
        if i >= warmup:
            times.append(dt)
        elif time.time() - t_all > budget_s / 2:
            times.append(dt)  # budget nearly gone in warm-up already: keep what we have
            break
        if time.time() - t_all > budget_s and times:
            break
    return statistics.median(times), len(times)


# --------------------------------------------------------------------------
# main
# --------------------------------------------------------------------------


_REAL_STDOUT = None


def _guard_stdout():
    """The contract is ONE JSON line on stdout: libraries that print banners there (NCCL's version line at
    NCCL_DEBUG=VERSION/WARN) are sent to stderr by pointing fd 1 at fd 2 for the duration of the run."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0)
    ap.add_argument("--sequential-solves", action="store_true",
                    help="decompose every Gram matrix inside the hook that assembled it (the reference's order) "
                         "instead of one batched solve per step for the computations of the step")
    ap.add_argument("--ncu-step", action="store_true",
                    help="warm up, run ONE step between cudaProfilerStart/Stop and exit "
                         "(for `ncu --profile-from-start off`; prints no bench line)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    w = WORKLOADS[args.workload]
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    es = 4 if args.dtype == "f32" else 8
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": w["name"], "batch": w["n"], "l2": "flushed (256 MiB write) between timed steps"}

    if args.impl == "reference":
        if rank != 0:
            return
        cores = use_all_host_threads()
        ms, n = time_reference(w, dtype, args.steps, args.warmup, args.cpu_budget_s * 1.5)
        line = {
            "impl": "reference", "metric": METRIC, "value": round(ms, 3), "unit": "ms", "n_gpus": args.gpus,
            "steps": n, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": round(ms, 3), "unit": "ms", "cores": cores, "kind": "port",
                             "sample": f"{n} full step(s) of the workload (oracle/reference_path.py, torch CPU)"},
            "e2e": {"value": round(ms, 3), "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        _emit(line)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; vivit_b200 has no CPU path (use --impl reference)")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    pg = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=device)
        pg = dist.group.WORLD

    from vivit_b200 import kernels

    stepper = Stepper(w, dtype, device, pg)
    stepper.batch_solves = not args.sequential_solves
    # (how the work is laid out is reported beside `config`, not inside it: `config` names the workload and is the
    # same object in the reference arm's line)
    schedule = {"parallelism": stepper.parallelism}
    shared_queue = stepper.batch_solves and len(stepper.calls) > 1
    if len(stepper.calls) > 1:
        schedule["solves"] = (
            "the computations of the step (one backward pass each, as in the reference) share a SolveQueue: their "
            "Gram matrices are decomposed by one batched vvt_syevj_batched call" if shared_queue else
            "every Gram matrix decomposed inside the hook that assembled it (the reference's order)")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, with_kernels=False):
        # one collection up front, so a generation-2 pass over leftovers of the warm-up does not land in the timed
        # region.  (The product itself leaves nothing for the cyclic collector: the hook/closure and factor-partner
        # reference cycles that used to keep gigabytes of factor tensors alive across steps -- one 60 ... 200 ms step
        # among 42 ms ones, scratch/e2e_steps.py -- are gone, tests/test_host_cpu.py::test_step_leaves_no_garbage.)
        import gc

        evs = []
        gc.collect()
        barrier()
        if with_kernels:
            kernels.timing_start()
        l0 = kernels.launch_count()
        t0 = time.time()
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        wall = (time.time() - t0) * 1e3 / steps
        launches = kernels.launch_count() - l0
        recs = kernels.timing_stop() if with_kernels else []
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, wall, launches, recs

    for _ in range(args.warmup):
        stepper.step_device()
    if args.ncu_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        stepper.step_device()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    sampler = ClockSampler(local) if rank == 0 else None
    ms, wall, launches, recs = timed(stepper.step_device, args.steps, with_kernels=True)
    clocks = sampler.stop() if sampler else None
    for _ in range(3):  # the first call also allocates the pinned result buffers
        stepper.step_e2e()
    ms_e2e, _, _, _ = timed(stepper.step_e2e, args.steps)
    sequential = None
    if shared_queue:  # the same step in the reference's order, for the record
        stepper.batch_solves = False
        for _ in range(3):
            stepper.step_device()
        seq_ms, _, seq_launches, _ = timed(stepper.step_device, args.steps)
        sequential = {"ms_per_step": round(seq_ms, 4), "gpu_launches": seq_launches,
                      "note": "--sequential-solves: each computation decomposes its Gram matrix inside its own hook"}
        stepper.batch_solves = True

    if rank != 0:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()
        return

    peaks, peak_src = load_peaks()
    rows = summarize_kernels(recs, args.steps, es, peaks)
    # roofline of the dominant kernel that has a roofline (the eigensolver is reported in ms only)
    # Roofline of the dominant roofline-scored kernel: the per-(entry point, shape) group with the largest
    # device time among the tensor- / HBM-bound kernels.  The eigensolver (largest share of the step) is
    # latency-bound and is reported separately below in ms / sweeps / residuals next to cuSOLVER.
    groups = {}
    floor_tensor = floor_hbm = 0.0  # algorithmic work of one step, all roofline-scored kernels
    for name, t_ms, shapes, n_launch, out_shapes in recs:
        kind, amount = algorithmic_work(name, shapes, es, out_shapes)
        if not kind:
            continue
        if kind == "tensor":
            floor_tensor += amount / args.steps
        else:
            floor_hbm += amount / args.steps
        g = groups.setdefault((name, tuple(shapes)), {"ms": 0.0, "calls": 0, "amount": amount, "kind": kind,
                                                      "launches": 0})
        g["ms"] += t_ms
        g["calls"] += 1
        g["launches"] += n_launch
    roof = None
    if groups:
        step_ms = sum(r[1] for r in recs) or 1.0
        (name, shapes), g = max(groups.items(), key=lambda kv: kv[1]["ms"])
        per_call_ms = g["ms"] / g["calls"]
        traffic = lookup_traffic(name, shapes)
        if g["kind"] == "tensor":
            peak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
            achieved = g["amount"] / (per_call_ms * 1e-3) / 1e12
            roof = {"kernel": name, "shapes": [list(s_) for s_ in shapes], "bound": "tensor",
                    "achieved": round(achieved, 2), "peak": peak, "unit": "TFLOP/s",
                    "frac": round(achieved / peak, 4), "traffic": traffic,
                    "peak_source": f"{peak_src} dense bf16 (sustained)",
                    "ms_per_call": round(per_call_ms, 4), "calls_per_step": g["calls"] / args.steps,
                    "launches_per_call": g["launches"] / g["calls"], "share_of_step": round(g["ms"] / step_ms, 4)}
            if args.dtype == "f32":  # TF32 = bf16 / 2 and three TF32 MMAs per fp32 product
                roof["peak_3xtf32"] = round(peak / 6.0, 1)
                roof["frac_of_3xtf32_peak"] = round(achieved / (peak / 6.0), 4)
        else:
            peak = peaks["hbm_gbs"]
            achieved = g["amount"] / (per_call_ms * 1e-3) / 1e9
            roof = {"kernel": name, "shapes": [list(s_) for s_ in shapes], "bound": "hbm",
                    "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                    "traffic": traffic, "peak_source": f"{peak_src} HBM copy", "ms_per_call": round(per_call_ms, 4),
                    "calls_per_step": g["calls"] / args.steps, "share_of_step": round(g["ms"] / step_ms, 4)}

    eig = eigensolver_report(stepper, len(stepper.calls) if shared_queue else 1) if world == 1 else None
    if roof is not None:
        # the whole step against its own floor: algorithmic flops / tensor peak + algorithmic bytes / HBM peak of
        # every roofline-scored kernel (the eigensolver has no such figure and counts as zero work: it can only
        # lower the fraction), divided by the measured step
        tensor_peak = (peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]) / (6.0 if args.dtype == "f32" else 48.0)
        floor_ms = floor_tensor / (tensor_peak * 1e12) * 1e3 + floor_hbm / (peaks["hbm_gbs"] * 1e9) * 1e3
        roof["step_floor_ms"] = round(floor_ms, 4)
        roof["frac_step"] = round(floor_ms / ms, 4)
        roof["step_algorithmic"] = {"tensor_flops": int(floor_tensor), "hbm_bytes": int(floor_hbm),
                                    "tensor_peak_tflops": round(tensor_peak, 1), "hbm_peak_gbs": peaks["hbm_gbs"]}
        top = max(rows, key=lambda r: r["ms_per_step"]) if rows else None
        if top is not None:
            roof["dominant_kernel_of_step"] = {"kernel": top["kernel"], "ms_per_step": top["ms_per_step"],
                                               "share_of_kernel_time": top["share"]}
            if top["kernel"].startswith("syevj") and eig:
                roof["dominant_kernel_of_step"].update(
                    note="latency-bound eigensolver: no tensor/HBM roofline; reported against cuSOLVER on the same GPU",
                    ms_per_solve=eig["ms"], cusolver_ms_per_solve=eig["cusolver_eigh_ms"], R=eig["R"], sweeps=eig["sweeps"],
                    batched=eig["batched"])

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        use_all_host_threads()
        cms, n = time_reference(w, dtype, 3, 1, args.cpu_budget_s)  # median of <= 3 steps after one warm-up, budget-bounded
        cpu = {"value": round(cms, 3), "unit": "ms", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{n} full step(s) of the workload (oracle/reference_path.py, torch CPU)"}

    line = {
        "metric": METRIC, "value": round(ms, 4), "unit": "ms", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 4), "wall_ms_per_step": round(wall, 4),
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic", "config": config, "schedule": schedule,
        "e2e": {"value": round(ms_e2e, 4), "unit": "ms", "h2d_bytes_per_step": stepper.h2d_bytes,
                "d2h_bytes_per_step": stepper.d2h_bytes, "pinned_copy_bandwidth": pcie_bandwidth(device)},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "eigensolver": eig, "cpu_baseline": cpu,
        "sequential_solves": sequential,
        # the part of the step that shards over the ranks (factor emit + Gram / cross-term assembly, rank 0) next
        # to the part that every rank repeats (the eigensolver): SURVEY 8e
        "gram_assembly_ms_per_step": round(sum(r["ms_per_step"] for r in rows if r["kernel"].startswith(("gram_", "v_emit_"))), 4),
        "eigensolver_ms_per_step": round(sum(r["ms_per_step"] for r in rows if r["kernel"].startswith("syevj")), 4),
        "kernels": rows,
    }
    _emit(line)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
