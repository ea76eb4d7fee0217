"""Per-entry-point checks of the C ABI on a real GPU.

Every kernel of ``libvivit_b200.so`` is called through ``vivit_b200.kernels``
(ctypes -> C ABI) and compared with the plain-torch test double evaluated in
float64 on the same inputs.  Tolerances: fp32 kernels use the 3xTF32 split and
must reach fp32-grade accuracy (``north_star``: rtol 1e-4; here 2e-5 of the
result scale), fp64 kernels 1e-10.
"""

import math
import os

import pytest
import torch

import tests._torch_kernels as ref

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.float64]


@pytest.fixture(scope="module")
def k():
    import vivit_b200.kernels as kernels

    return kernels


def dev():
    return torch.device("cuda:0")


def rnd(*shape, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g, dtype=torch.float64) * scale).to(dtype).to(dev())


def close(got, want, dtype, what=""):
    want = want.to(torch.float64)
    got = got.to(torch.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    scale = max(want.abs().max().item(), 1e-30)
    tol = 2e-5 if dtype == torch.float32 else 1e-10
    err = (got - want).abs().max().item() / scale
    assert err <= tol, f"{what}: rel-to-scale error {err:.3e} > {tol:.1e}"


def f64(*ts):
    return [None if t is None else (t.double() if t.is_floating_point() else t) for t in ts]


# --------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("N,C,sub", [(7, 5, None), (32, 10, [3, 0, 9]), (5, 100, None), (4, 1, None)])
def test_loss_factor_ce(k, dtype, N, C, sub):
    logits = rnd(N, C, dtype=dtype, scale=3.0)
    subt = None if sub is None else torch.tensor(sub, device=dev())
    for mean in (True, False):
        close(k.loss_sqrt_hessian_ce(logits, subt, mean), ref.loss_sqrt_hessian_ce(logits.double(), subt, mean), dtype, "ce")
    n_sub = N if sub is None else len(sub)
    ids = torch.randint(0, C, (3, n_sub), generator=torch.Generator().manual_seed(1)).to(dev())
    close(k.loss_sqrt_hessian_ce_mc(logits, subt, ids, True), ref.loss_sqrt_hessian_ce_mc(logits.double(), subt, ids, True), dtype, "ce_mc")


@pytest.mark.parametrize("dtype", DTYPES)
def test_loss_factor_mse_and_scale(k, dtype):
    like = torch.zeros(1, dtype=dtype, device=dev())
    close(k.loss_sqrt_hessian_mse(6, 5, 0.37, like), ref.loss_sqrt_hessian_mse(6, 5, 0.37, like.double()), dtype)
    t = rnd(1000, 7, dtype=dtype)
    want = t.double() * 1.7
    close(k.scale_(t, 1.7), want, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("numel", [1, 1000 * 7, 3 * 1024 * 1024 + 5])
def test_axpy(k, dtype, numel):
    y, x = rnd(numel, dtype=dtype), rnd(numel, dtype=dtype, seed=1)
    want = y.double() - 0.25 * x.double()
    got = k.axpy_(y, x, -0.25)
    assert got is y
    close(got, want, dtype)
    empty = torch.empty(0, dtype=dtype, device=dev())
    assert k.axpy_(empty, empty.clone()).numel() == 0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,n_out,n_in", [(30, 10, 32), (320, 64, 784), (1280, 256, 10), (17, 3, 5), (1, 1, 1)])
def test_backprop_linear(k, dtype, rows, n_out, n_in):
    S, W = rnd(rows, n_out, dtype=dtype), rnd(n_out, n_in, dtype=dtype, seed=1)
    close(k.sqrt_backprop_linear(S, W), ref.sqrt_backprop_linear(*f64(S, W)), dtype)
    S3 = S.reshape(1, rows, n_out)
    assert k.sqrt_backprop_linear(S3, W).shape == (1, rows, n_in)


CONV_CASES = [
    # V, N, ci, h, w, co, k, stride, pad, dil
    (2, 3, 3, 6, 6, 2, 3, 1, 1, 1),
    (3, 2, 2, 11, 11, 3, 3, 1, 0, 1),
    (2, 2, 4, 8, 8, 4, 3, 2, 1, 1),
    (2, 2, 3, 9, 7, 5, 5, 1, 0, 1),
    (1, 3, 4, 5, 5, 6, 1, 1, 0, 1),
    (2, 2, 2, 9, 9, 3, 3, 2, 0, 1),
    (2, 2, 2, 9, 9, 3, 3, 1, 2, 2),
    (10, 4, 16, 12, 12, 24, 3, 1, 1, 1),
    (3, 5, 7, 11, 13, 130, 3, 2, 1, 1),   # odd sizes, stride 2, c_out > one tile, spatial pitch not a multiple of 4
    (10, 8, 64, 14, 14, 96, 3, 1, 0, 1),  # cifar10_3c3d conv2 geometry at a small batch
    (1, 6, 3, 8, 8, 5, 1, 1, 0, 1),       # 1x1 kernel, a single "class" (BatchGrad emit)
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_backprop_and_emit(k, dtype, case):
    V, N, ci, h, w, co, ks, st, pd, dl = case
    ho = (h + 2 * pd - dl * (ks - 1) - 1) // st + 1
    wo = (w + 2 * pd - dl * (ks - 1) - 1) // st + 1
    S = rnd(V, N, co, ho, wo, dtype=dtype)
    W = rnd(co, ci, ks, ks, dtype=dtype, seed=1)
    X = rnd(N, ci, h, w, dtype=dtype, seed=2)
    args = ((st, st), (pd, pd), (dl, dl))
    close(k.sqrt_backprop_conv2d(S, W, (h, w), *args), ref.sqrt_backprop_conv2d(*f64(S, W), (h, w), *args), dtype, "dgrad")
    close(k.v_emit_conv2d(S, X, (ks, ks), *args), ref.v_emit_conv2d(*f64(S, X), (ks, ks), *args), dtype, "emit")
    close(k.v_emit_bias(S), ref.v_emit_bias(S.double()), dtype, "bias")


def test_conv2d_without_workspace_uses_the_generic_kernel(k):
    """The C ABI accepts a NULL workspace for the Conv2d entry points (generic implicit-GEMM kernel);
    results agree with the tensor-core path that kernels.py selects."""
    import ctypes

    from vivit_b200 import _lib

    lib = _lib.load()
    V, N, ci, h, w, co, ks = 4, 3, 6, 9, 9, 10, 3
    S = rnd(V, N, co, 7, 7, dtype=torch.float32)
    W = rnd(co, ci, ks, ks, dtype=torch.float32, seed=1)
    X = rnd(N, ci, h, w, dtype=torch.float32, seed=2)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    out = torch.empty(V, N, ci, h, w, device=dev())
    st = lib.vvt_sqrt_backprop_conv2d(P(out), P(S), P(W), V * N, co, 7, 7, ci, h, w, ks, ks, 1, 1, 0, 0, 1, 1,
                                      None, 0, 0, stream)
    assert st == 0
    Vt = torch.empty(V, N, co, ci, ks, ks, device=dev())
    st = lib.vvt_v_emit_conv2d(P(Vt), P(S), P(X), V, N, co, 7, 7, ci, h, w, ks, ks, 1, 1, 0, 0, 1, 1, None, 0, 0, stream)
    assert st == 0
    args = ((1, 1), (0, 0), (1, 1))
    close(out, ref.sqrt_backprop_conv2d(*f64(S, W), (h, w), *args), torch.float32, "dgrad (generic)")
    close(Vt, ref.v_emit_conv2d(*f64(S, X), (ks, ks), *args), torch.float32, "emit (generic)")
    close(out, k.sqrt_backprop_conv2d(S, W, (h, w), *args).double(), torch.float32, "dgrad paths agree")
    assert lib.vvt_conv2d_workspace_bytes(0, V, N, co, 7, 7, ci, ks, ks, 0) > 0
    assert lib.vvt_conv2d_workspace_bytes(1, V, N, co, 7, 7, ci, ks, ks, 1) == 0  # fp64: generic kernel


@pytest.mark.parametrize("dtype", DTYPES)
def test_elementwise_and_pools(k, dtype):
    S = rnd(3, 4, 5, 7, 7, dtype=dtype)
    r = rnd(4, 5, 7, 7, dtype=dtype, seed=3)
    for act in range(9):
        close(k.sqrt_backprop_elementwise(S, r, act, 2.0), ref.sqrt_backprop_elementwise(S.double(), r.double(), act, 2.0), dtype, f"act{act}")
    x = rnd(4, 5, 9, 9, dtype=dtype, seed=4)
    for kern, st, pd, ceil, dil in [(3, 2, 0, False, 1), (3, 2, 0, True, 1), (2, 2, 0, False, 1), (3, 1, 1, False, 1),
                                    (3, 2, 1, True, 1), (3, 3, 0, False, 1), (2, 1, 0, False, 2), (4, 3, 1, False, 1)]:
        y, idx = torch.nn.functional.max_pool2d(x, kern, st, pd, dil, ceil, return_indices=True)
        # the native arg-max kernel chooses what torch chooses -- also among tied values (a ReLU in front of the pool)
        assert torch.equal(k.maxpool2d_argmax(x, (kern, kern), (st, st), (pd, pd), (dil, dil), ceil), idx), (kern, st, pd, ceil, dil)
        xr = torch.relu(x - 0.3)
        assert torch.equal(k.maxpool2d_argmax(xr, (kern, kern), (st, st), (pd, pd), (dil, dil), ceil),
                           torch.nn.functional.max_pool2d(xr, kern, st, pd, dil, ceil, return_indices=True)[1]), "ties"
        Sp = rnd(3, *y.shape, dtype=dtype, seed=5)
        a = ((9, 9), (kern, kern), (st, st), (pd, pd), (dil, dil))
        close(k.sqrt_backprop_maxpool2d(Sp, idx, *a), ref.sqrt_backprop_maxpool2d(Sp.double(), idx, *a), dtype, "maxpool")
    xb = rnd(2, 3, 30, 28, dtype=dtype, seed=8)  # a map larger than one block of threads
    y, idx = torch.nn.functional.max_pool2d(xb, 3, 2, 0, 1, True, return_indices=True)
    assert torch.equal(k.maxpool2d_argmax(xb, (3, 3), (2, 2), (0, 0), (1, 1), True), idx)
    xn = xb.clone()
    xn[0, 1, 4, 5] = float("nan")  # NaN wins its windows, as in torch
    assert torch.equal(k.maxpool2d_argmax(xn, (3, 3), (2, 2), (0, 0), (1, 1), True),
                       torch.nn.functional.max_pool2d(xn, 3, 2, 0, 1, True, return_indices=True)[1])
    x_rect = rnd(2, 3, 11, 6, dtype=dtype, seed=10)  # rectangular window, stride, padding and dilation
    assert torch.equal(k.maxpool2d_argmax(x_rect, (3, 2), (2, 1), (1, 0), (2, 1), False),
                       torch.nn.functional.max_pool2d(x_rect, (3, 2), (2, 1), (1, 0), (2, 1), False, return_indices=True)[1])
    Sp = rnd(2, *y.shape, dtype=dtype, seed=9)
    a = ((30, 28), (3, 3), (2, 2), (0, 0), (1, 1))
    close(k.sqrt_backprop_maxpool2d(Sp, idx, *a), ref.sqrt_backprop_maxpool2d(Sp.double(), idx, *a), dtype, "maxpool big")
    for kern, st, pd in [(3, 3, 0), (2, 2, 0), (3, 1, 1), (9, 9, 0)]:
        y = torch.nn.functional.avg_pool2d(x, kern, st, pd)
        Sp = rnd(2, *y.shape, dtype=dtype, seed=6)
        a = ((9, 9), (kern, kern), (st, st), (pd, pd))
        close(k.sqrt_backprop_avgpool2d(Sp, *a), ref.sqrt_backprop_avgpool2d(Sp.double(), *a), dtype, "avgpool")
    Sl, Z = rnd(3, 4, 6, dtype=dtype), rnd(4, 5, dtype=dtype, seed=7)
    close(k.v_emit_linear(Sl, Z), ref.v_emit_linear(Sl.double(), Z.double()), dtype, "emit_linear")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
@pytest.mark.parametrize("M,N,K", [(64, 64, 64), (130, 70, 33), (1, 5, 7), (257, 129, 1000), (40, 300, 5)])
def test_gemm(k, dtype, ta, tb, M, N, K):
    A = rnd(*((K, M) if ta else (M, K)), dtype=dtype)
    B = rnd(*((K, N) if tb else (N, K)), dtype=dtype, seed=1)
    close(k.gemm(A, B, ta, tb), ref.gemm(A.double(), B.double(), ta, tb), dtype)
    out = rnd(M, N, dtype=dtype, seed=2)
    want = ref.gemm(A.double(), B.double(), ta, tb, out=out.double().clone(), alpha=0.5, beta=2.0)
    close(k.gemm(A, B, ta, tb, out=out, alpha=0.5, beta=2.0), want, dtype, "alpha/beta")
    Ab = rnd(3, *A.shape, dtype=dtype, seed=3)
    Bb = rnd(3, *B.shape, dtype=dtype, seed=4)
    close(k.gemm(Ab, Bb, ta, tb), ref.gemm(Ab.double(), Bb.double(), ta, tb), dtype, "batched")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("R,D", [(320, 640), (1280, 5000), (50, 7), (129, 100000), (32, 1)])
def test_gram_dense_and_cross(k, dtype, R, D):
    V = rnd(R, D, dtype=dtype)
    G0 = rnd(R, R, dtype=dtype, seed=1)
    G0 = G0 + G0.t()
    close(k.gram_dense_accum(G0.clone(), V), ref.gram_dense_accum(G0.double(), V.double()), dtype, "dense")
    g = rnd(37, D, dtype=dtype, seed=2)
    X0 = rnd(R, 37, dtype=dtype, seed=3)
    close(k.gram_cross_accum(X0.clone(), V, g), ref.gram_cross_accum(X0.double(), V.double(), g.double()), dtype, "cross")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C,N,n_out,n_in", [(10, 32, 64, 784), (10, 32, 10, 32), (5, 3, 5, 6), (3, 50, 300, 17), (10, 128, 512, 1152)])
@pytest.mark.parametrize("bias", [False, True])
def test_gram_linear(k, dtype, C, N, n_out, n_in, bias):
    S, Z = rnd(C, N, n_out, dtype=dtype), rnd(N, n_in, dtype=dtype, seed=1)
    G0 = torch.zeros(C * N, C * N, dtype=dtype, device=dev())
    close(k.gram_linear_accum(G0.clone(), S, Z, bias), ref.gram_linear_accum(G0.double(), S.double(), Z.double(), bias), dtype, "gram_linear")
    Dl, Zg = rnd(9, n_out, dtype=dtype, seed=2), rnd(9, n_in, dtype=dtype, seed=3)
    X0 = rnd(C * N, 9, dtype=dtype, seed=4)
    close(
        k.gram_cross_linear_accum(X0.clone(), S, Z, Dl, Zg, bias),
        ref.gram_cross_linear_accum(X0.double(), *f64(S, Z, Dl, Zg), bias),
        dtype,
        "cross_linear",
    )


def _psd(R, rank, dtype, seed=0):
    B = rnd(R, rank, dtype=torch.float64, seed=seed)
    return (B @ B.t() / rank).to(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("R,rank", [(1, 1), (5, 5), (32, 3), (33, 33), (100, 40), (320, 288), (640, 64), (1280, 1152)])
def test_syevj(k, dtype, R, rank):
    G = _psd(R, rank, dtype, seed=R)
    evals, evecs, info = k.syevj(G, True, return_info=True)
    assert info["converged"]
    want = torch.linalg.eigvalsh(G.double())
    tol = 2e-5 if dtype == torch.float32 else 1e-11
    assert (evals.double() - want).abs().max() <= tol * want.abs().max(), (evals.double() - want).abs().max()
    assert (evals[1:] >= evals[:-1]).all()
    U = evecs.double()
    eye = torch.eye(R, dtype=torch.float64, device=dev())
    orth = (U.t() @ U - eye).abs().max().item()
    resid = (G.double() @ U - U * evals.double()[None]).norm().item() / max(G.double().norm().item(), 1e-30)
    lim = 5e-5 if dtype == torch.float32 else 1e-11
    assert orth <= lim and resid <= lim, (orth, resid, info)
    ev_only, none = k.syevj(G, False)
    assert none is None
    assert (ev_only.double() - want).abs().max() <= tol * want.abs().max()


def _wide_round(L, rnd_idx):
    """One round of the two-level solver through the library's test hook."""
    import ctypes

    from vivit_b200 import _lib

    fn = _lib.load().vvt_dbg_wide_round
    Np = L.shape[0]
    pairs = Np // 128
    H = torch.zeros(pairs, 128, 128, device=L.device)
    Qt = torch.zeros(pairs, 128, 128, device=L.device)
    flag = torch.zeros(pairs, dtype=torch.int32, device=L.device)
    st = fn(L.data_ptr(), H.data_ptr(), Qt.data_ptr(), flag.data_ptr(), Np, rnd_idx,
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(st, "vvt_dbg_wide_round")
    torch.cuda.synchronize()
    return H, Qt, flag


def _rr_pair(n, rnd_idx, idx):
    m = n - 1
    if idx == 0:
        return rnd_idx, m
    return (rnd_idx + idx) % m, (rnd_idx - idx + m) % m


@pytest.mark.parametrize("Np,rnd_idx", [(256, -1), (256, 0), (512, 2), (1280, -1), (1280, 7)])
def test_wide_round_kernels(k, Np, rnd_idx):
    """The three kernels of a wide round, one by one: pair Grams on tcgen05 with MN-major operands (against
    float64 ``P^T P``), the rotation kernel (``Q`` orthogonal, rotated pairs closer to diagonal), the tcgen05
    apply (``P <- P Q`` in place, untouched columns bit-identical)."""
    g = torch.Generator(device="cpu").manual_seed(Np + rnd_idx)
    L0 = torch.tril(torch.randn(Np, Np, generator=g, dtype=torch.float64)) / math.sqrt(Np)
    L0 = (L0 * torch.logspace(0, -2, Np, dtype=torch.float64)[None, :]).float().to(dev())
    L0 = L0 / L0.norm()
    L = L0.clone()
    H, Qt, flag = _wide_round(L, rnd_idx)
    nbw, pairs = Np // 64, Np // 128
    for pr in range(pairs):
        wa, wb = (2 * pr, 2 * pr + 1) if rnd_idx < 0 else _rr_pair(nbw, rnd_idx, pr)
        cols = torch.cat([torch.arange(wa * 64, wa * 64 + 64), torch.arange(wb * 64, wb * 64 + 64)]).to(dev())
        P = L0[:, cols].double()
        want = P.t() @ P
        scale = want.abs().max().item()
        # (the tensor core adds into its accumulator with truncation: a bias of about -3e-6 on sums of squares)
        # (VVT_WIDE_CROSS_GRAM=1: cross rounds form only the columns of b, H_ab and H_bb; H_aa comes from the
        # solver's diagonal-block cache)
        part = slice(64, None) if rnd_idx >= 0 and os.environ.get("VVT_WIDE_CROSS_GRAM") else slice(None)
        assert (H[pr][:, part].double() - want[:, part]).abs().max().item() <= 1e-5 * scale, ("gram", pr)
        Pn = L[:, cols].double()
        if flag[pr].item() == 0:
            assert torch.equal(L[:, cols], L0[:, cols])
            continue
        Q = Qt[pr].double().t()
        eye = torch.eye(128, dtype=torch.float64, device=dev())
        assert (Q.t() @ Q - eye).abs().max().item() <= 2e-5, ("orthogonal", pr)
        assert (Pn - P @ Q).abs().max().item() <= 1e-5 * P.abs().max().item(), ("apply", pr)
        # every Jacobi rotation lowers the off-diagonal norm of the pair Gram by 2 h_pq^2
        Hn = Pn.t() @ Pn
        off = lambda M: (M - torch.diag(M.diagonal())).norm().item()  # noqa: E731
        assert off(Hn) < off(want), (off(Hn), off(want))


@pytest.mark.parametrize("R,rank", [(2048, 1800), (2560, 2304), (5120, 4608)])
def test_syevj_two_level(k, R, rank, monkeypatch):
    """Large fp32 problems run the two-level (wide, tcgen05) rounds; forced here from 2048 columns on."""
    monkeypatch.setenv("VVT_SYEVJ_WIDE_MIN", "2048")
    G = _psd(R, rank, torch.float32, seed=R)
    evals, evecs, info = k.syevj(G, True, return_info=True)
    assert info["converged"], info
    want = torch.linalg.eigvalsh(G.double())
    assert (evals.double() - want).abs().max() <= 2e-5 * want.abs().max()
    assert (evals[1:] >= evals[:-1]).all()
    U = evecs.double()
    eye = torch.eye(R, dtype=torch.float64, device=dev())
    orth = (U.t() @ U - eye).abs().max().item()
    resid = (G.double() @ U - U * evals.double()[None]).norm().item() / G.double().norm().item()
    assert orth <= 1e-4 and resid <= 5e-5, (orth, resid, info)


@pytest.mark.parametrize("R", [1279, 2047, 4607])
def test_syevj_sizes_that_are_not_a_multiple_of_four(k, R):
    """Row pitches that are not a multiple of 16 bytes keep the GEMMs of the refinement step off the TMA-fed tcgen05
    kernel (they run on ``mma.sync``, whose tensor-core accumulation is not promoted to round-to-nearest adds): the
    eigenvalue error is 3-7x that of the neighbouring aligned size (1.1e-5 against 6e-6 of the largest eigenvalue at
    R = 4607 / 4608, with contractions cut into slabs of 2048) and stays inside the north star's 1e-4.  R = 4607 also runs the two-level rounds on a padded
    panel (4608 columns)."""
    G = _psd(R, int(0.8 * R), torch.float32, seed=R)
    evals, evecs, info = k.syevj(G, True, return_info=True)
    assert info["converged"], info
    want = torch.linalg.eigvalsh(G.double())
    assert (evals.double() - want).abs().max() <= 1e-4 * want.abs().max()
    U = evecs.double()
    eye = torch.eye(R, dtype=torch.float64, device=dev())
    orth = (U.t() @ U - eye).abs().max().item()
    resid = (G.double() @ U - U * evals.double()[None]).norm().item() / G.double().norm().item()
    assert orth <= 1e-3 and resid <= 1e-4, (orth, resid, info)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,R", [(1, 33), (3, 64), (4, 320), (7, 100), (2, 1280)])
def test_syevj_batched(k, dtype, B, R):
    """Batched solve (the per-group Grams of block-diagonal groups): every problem matches its own single
    solve and float64 eigvalsh within tolerance; problems of very different scale converge independently."""
    Gs = torch.stack([_psd(R, max(1, R - 7 * b - 3), dtype, seed=R + b) * (10.0 ** b) for b in range(B)])
    evals, evecs, infos = k.syevj_batched(Gs, True, return_info=True)
    assert all(i["converged"] for i in infos), infos
    for b in range(B):
        want = torch.linalg.eigvalsh(Gs[b].double())
        tol = 2e-5 if dtype == torch.float32 else 1e-11
        assert (evals[b].double() - want).abs().max() <= tol * want.abs().max()
        U = evecs[b].double()
        resid = (Gs[b].double() @ U - U * evals[b].double()[None]).norm().item() / Gs[b].double().norm().item()
        lim = 5e-5 if dtype == torch.float32 else 1e-11
        assert resid <= lim, (b, resid)
        single_vals, _ = k.syevj(Gs[b].contiguous(), True)
        assert (single_vals.double() - evals[b].double()).abs().max() <= tol * want.abs().max()
    ev_only, none = k.syevj_batched(Gs, False)
    assert none is None
    for b in range(B):
        scale = evals[b].abs().max().item()
        assert (ev_only[b].double() - evals[b].double()).abs().max().item() <= (2e-5 if dtype == torch.float32 else 1e-11) * scale


def test_syevj_batched_two_level(k, monkeypatch):
    monkeypatch.setenv("VVT_SYEVJ_WIDE_MIN", "2048")
    Gs = torch.stack([_psd(2048, 1500 + 200 * b, torch.float32, seed=b) for b in range(2)])
    evals, evecs = k.syevj_batched(Gs, True)
    for b in range(2):
        want = torch.linalg.eigvalsh(Gs[b].double())
        assert (evals[b].double() - want).abs().max() <= 2e-5 * want.abs().max()
        U = evecs[b].double()
        assert (Gs[b].double() @ U - U * evals[b].double()[None]).norm() / Gs[b].double().norm() <= 5e-5


def test_symeig_indefinite_matrix(k):
    """``vivit/utils/eig.py`` ``symeig`` takes any symmetric matrix; the one-sided kernel is for PSD input, so
    the wrapper shifts by a Gershgorin bound."""
    from vivit_b200.utils.eig import symeig, symeig_psd

    g = torch.Generator(device="cpu").manual_seed(5)
    A = torch.randn(60, 60, generator=g, dtype=torch.float64)
    A = ((A + A.t()) / 2).to(dev())  # indefinite
    for dtype, tol in ((torch.float64, 1e-10), (torch.float32, 2e-5)):
        M = A.to(dtype)
        evals, evecs = symeig_psd(M, eigenvectors=True)
        want = torch.linalg.eigvalsh(M.double())
        assert (evals.double() - want).abs().max() <= tol * want.abs().max()
        U = evecs.double()
        assert (M.double() @ U - U * evals.double()[None]).abs().max() <= 10 * tol * want.abs().max()
        nz_evals, _ = symeig(M, eigenvectors=False)
        assert nz_evals.numel() == 60 and (nz_evals < 0).any()


def test_mixed_dtype_operands_are_rejected(k):
    S, W = rnd(4, 6, dtype=torch.float32), rnd(6, 3, dtype=torch.float64)
    with pytest.raises(TypeError):
        k.sqrt_backprop_linear(S, W)
    with pytest.raises(TypeError):
        k.gram_dense_accum(torch.zeros(4, 4, device=dev(), dtype=torch.float64), S)


def test_newton_coeff_many_directions(k):
    """More directions than one 48 KB shared-memory tile of coefficients (K = 7000 > 6144)."""
    R, K = 7000, 7000
    U = rnd(R, K, dtype=torch.float64, scale=1e-2)
    gam, lam = rnd(3, K, dtype=torch.float64), rnd(2, K, dtype=torch.float64).abs() + 1.0
    dl, ev = torch.ones(K, dtype=torch.float64, device=dev()), rnd(K, dtype=torch.float64).abs() + 0.5
    got = k.newton_coeff(U, gam, lam, dl, ev, 1.5)
    close(got, ref.newton_coeff(U, gam, lam, dl, ev, 1.5), torch.float64, "newton_coeff K=7000")


def test_syevj_reads_upper_triangle_and_keeps_input(k):
    G = _psd(40, 40, torch.float64)
    Gu = torch.triu(G) + torch.tril(torch.full_like(G, 123.0), -1)  # garbage below the diagonal
    before = Gu.clone()
    evals, _ = k.syevj(Gu, True)
    assert torch.equal(Gu, before)
    assert torch.allclose(evals, torch.linalg.eigvalsh(G), rtol=1e-10, atol=1e-12)


def test_syevj_lapack_killer(k):
    """The reference's only binary fixture: LAPACK syevd does not converge on it
    (test/utils/test_stable_symeig.py:25-45); Jacobi must, without a shift."""
    import os

    path = os.path.join(os.path.dirname(__file__), "golden", "symeig_killer.pt")
    G = torch.load(path).to(dev())
    evals, evecs = k.syevj(G, True)  # raises if the sweep limit is hit
    top = evals[-1].double().item()
    Gd = G.double()
    sym = torch.triu(Gd) + torch.triu(Gd, 1).t()
    assert abs(top - 6.28e7) / 6.28e7 < 1e-2
    v = evecs[:, -1].double()
    assert ((sym @ v) - top * v).norm() / abs(top) < 1e-5
    assert (evecs.double().t() @ evecs.double() - torch.eye(128, device=dev(), dtype=torch.float64)).abs().max() < 1e-4


@pytest.mark.parametrize("dtype", DTYPES)
def test_filter_nonzero(k, dtype):
    ev = torch.tensor([0.0, 5e-8, -5e-8, 2e-7, 1e-3, -1.0], dtype=dtype, device=dev())
    assert torch.equal(k.filter_nonzero(ev), ref.filter_nonzero(ev))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("K,R,D", [(1, 50, 33), (10, 1280, 4803), (3, 320, 100000), (20, 64, 300), (40, 32, 1000)])
def test_backtransform_dense(k, dtype, K, R, D):
    U, V = rnd(K, R, dtype=dtype), rnd(R, D, dtype=dtype, seed=1)
    n2 = torch.zeros(K, dtype=torch.float64, device=dev())
    n2r = torch.zeros(K, dtype=torch.float64, device=dev())
    E = k.backtransform_dense(U, V, n2)
    Er = ref.backtransform_dense(U.double(), V.double(), n2r)
    close(E, Er, dtype, "E")
    close(n2, n2r, dtype, "norm2")
    close(k.scale_rows_rsqrt(E, n2), ref.scale_rows_rsqrt(Er, n2r), dtype, "normalised")
    close(k.v_apply_dense(U[0].contiguous(), V), ref.v_apply_dense(U[0].double(), V.double()), dtype, "v_apply")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("K,C,N,n_out,n_in", [(1, 3, 4, 5, 6), (10, 10, 32, 64, 784), (4, 10, 128, 10, 256), (7, 2, 9, 130, 3)])
def test_structured_linear_products(k, dtype, K, C, N, n_out, n_in):
    U, S, Z = rnd(K, C * N, dtype=dtype), rnd(C, N, n_out, dtype=dtype, seed=1), rnd(N, n_in, dtype=dtype, seed=2)
    n2 = torch.zeros(K, dtype=torch.float64, device=dev())
    n2r = torch.zeros(K, dtype=torch.float64, device=dev())
    close(k.backtransform_linear(U, S, Z, n2), ref.backtransform_linear(*f64(U, S, Z), n2r), dtype, "E")
    close(n2, n2r, dtype, "norm2")
    close(k.v_apply_linear(U[0].contiguous(), S, Z), ref.v_apply_linear(*f64(U[0], S, Z)), dtype, "v_apply_linear")
    M = rnd(K, n_out, n_in, dtype=dtype, seed=3)
    close(k.vt_mat_prod_linear(S, Z, M), ref.vt_mat_prod_linear(*f64(S, Z, M)), dtype, "vt_mat_prod")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C,N_ggn,n_g,K,N", [(10, 32, 32, 10, 32), (5, 2, 3, 1, 4), (1, 32, 128, 7, 128), (10, 128, 128, 10, 128)])
def test_dirderiv_and_newton(k, dtype, C, N_ggn, n_g, K, N):
    R = C * N_ggn
    G = _psd(R, max(R // 2, K), dtype, seed=3)
    X = rnd(R, n_g, dtype=dtype, seed=4)
    ev, evec = torch.linalg.eigh(G.double() * (N / N_ggn))
    evals = ev[-K:].to(dtype).contiguous()
    U = evec[:, -K:].to(dtype).contiguous()
    gam, lam = k.dirderiv_epilogue(G, X, U, evals, C, N_ggn, N)
    gam_r, lam_r = ref.dirderiv_epilogue(*f64(G, X, U, evals), C, N_ggn, N)
    close(gam, gam_r, dtype, "gammas")
    close(lam, lam_r, dtype, "lambdas")
    deltas = torch.full((K,), 0.5, dtype=dtype, device=dev())
    corr = math.sqrt(N / N_ggn)
    close(k.newton_coeff(U, gam, lam, deltas, evals, corr), ref.newton_coeff(*f64(U, gam_r, lam_r, deltas, evals), corr), dtype, "newton v")


def test_cpu_tensor_is_refused(k):
    from vivit_b200._lib import KernelLibraryError

    with pytest.raises(KernelLibraryError):
        k.scale_(torch.ones(3), 2.0)


def test_launch_counter_moves(k):
    before = k.launch_count()
    k.scale_(torch.ones(3, device=dev()), 2.0)
    assert k.launch_count() == before + 1


# --------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(3, 6, 7), (128, 4096), (5, 10), (1, 33), (33, 1), (9, 130), (257, 3, 3, 3)])
def test_center_rows(k, dtype, shape):
    """``vvt_center_rows``: vector path (row length a multiple of 16 bytes), scalar path, more rows than
    warps, single row / single column, empty; out of place and in place."""
    g = rnd(*shape, dtype=dtype) + 0.5
    want = ref.center_rows(g.double())
    keep = g.clone()
    close(k.center_rows(g), want, dtype, "center_rows")
    assert torch.equal(g, keep)  # out of place leaves the input alone
    out = k.center_rows(g, inplace=True)
    assert out.data_ptr() == g.data_ptr()
    close(g, want, dtype, "center_rows (in place)")
    assert k.center_rows(torch.zeros(2, 0, dtype=dtype, device=dev())).shape == (2, 0)
    # an odd base address (a view starting one element in) must take the scalar path
    if len(shape) == 2 and shape[1] >= 8:
        flat = rnd(shape[0] * shape[1] + 1, dtype=dtype)
        view = flat[1:].reshape(shape)
        close(k.center_rows(view), ref.center_rows(view.double()), dtype, "center_rows (unaligned)")
