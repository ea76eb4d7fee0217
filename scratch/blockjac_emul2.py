"""Accuracy per sweep of one-sided Jacobi (scalar rotations, parallel round-robin order) on
W=G versus W=chol(G+eps I); fp32; real c2 Gram.  Vectorised: one round = 640 disjoint column pairs."""
import sys, numpy as np, time
G = np.load('scratch/G_c2.npy').astype(np.float64); G = (G + G.T) / 2
R = G.shape[0]
wref, Uref = np.linalg.eigh(G)
lmax = wref[-1]

def rr_round(n, rnd):
    m = n - 1
    idx = np.arange(1, n // 2)
    a = np.r_[rnd, (rnd + idx) % m]; b = np.r_[m, (rnd - idx + m) % m]
    return np.minimum(a, b), np.maximum(a, b)

def run(W0, is_chol, shift, tol, dtype=np.float32, max_sweeps=24, label='', sort=False):
    scale = np.linalg.norm(W0)
    W = (W0 / scale).astype(dtype)
    if sort:
        order = np.argsort(-np.sum(W.astype(np.float64)**2, axis=0)); W = W[:, order]
    eps = np.finfo(dtype).eps
    abs2 = dtype(eps * eps)
    for sweep in range(1, max_sweeps + 1):
        nrot = 0
        for rnd in range(R - 1):
            p, q = rr_round(R, rnd)
            Wp, Wq = W[:, p], W[:, q]
            hpp = np.einsum('ij,ij->j', Wp, Wp); hqq = np.einsum('ij,ij->j', Wq, Wq); hpq = np.einsum('ij,ij->j', Wp, Wq)
            need = (hpq * hpq > dtype(tol * tol) * np.abs(hpp * hqq)) & (np.abs(hpq) > abs2)
            nrot += int(need.sum())
            with np.errstate(all='ignore'):
                z = (hqq - hpp) / (2 * hpq)
                t = np.sign(z) / (np.abs(z) + np.sqrt(1 + z * z))
                t = np.where(z == 0, 1.0, t)
            c = 1 / np.sqrt(1 + t * t); s = t * c
            c = np.where(need, c, 1).astype(dtype); s = np.where(need, s, 0).astype(dtype)
            W[:, p] = c * Wp - s * Wq
            W[:, q] = s * Wp + c * Wq
        n2 = np.sum(W.astype(np.float64) ** 2, axis=0)
        lam = (n2 * scale**2 - shift) if is_chol else np.sqrt(n2) * scale
        o = np.argsort(lam)
        err = np.abs(lam[o] - wref).max() / lmax
        # top-10 eigenvectors
        Ut = W[:, o[-10:]].astype(np.float64); Ut /= np.linalg.norm(Ut, axis=0)
        sub = np.linalg.norm(Ut - Uref[:, -10:] @ (Uref[:, -10:].T @ Ut), axis=0).max()
        Hn = (W.astype(np.float64).T @ W.astype(np.float64)) / np.sqrt(np.outer(n2, n2) + 1e-300)
        live = n2 > 1e-9 * n2.max()
        Hl = np.abs(Hn[np.ix_(live, live)]); np.fill_diagonal(Hl, 0)
        print(f'{label} sweep {sweep}: rotations {nrot} ({nrot/(R*(R-1)/2):.3f})  eval err {err:.2e}  top10 subspace err {sub:.2e}  maxcos(live {live.sum()}) {Hl.max():.2e}', flush=True)
        if nrot == 0: break

which = sys.argv[1]
if which == 'A':
    run(G, False, 0.0, 1e-5, label='A W=G')
elif which == 'As':
    run(G, False, 0.0, 1e-5, label='As W=G sorted', sort=True)
elif which == 'L':
    shift = 1e-6 * lmax
    L = np.linalg.cholesky(G + shift * np.eye(R))
    run(L, True, shift, 1e-5, label='L W=chol')
elif which == 'L3':
    shift = 1e-6 * lmax
    L = np.linalg.cholesky(G + shift * np.eye(R))
    run(L, True, shift, 1e-3, label='L3 W=chol tol 1e-3')
elif which.startswith('S'):
    shift = float(which[1:]) * np.linalg.norm(G)
    print('shift/lmax', shift/lmax)
    L = np.linalg.cholesky(G + shift * np.eye(R))
    run(L, True, shift, 1e-5, label=which)
elif which == 'Lt':
    shift = 1e-6 * lmax
    L = np.linalg.cholesky(G + shift * np.eye(R))
    run(L.T.copy(), True, shift, 1e-5, label='Lt W=chol^T')
