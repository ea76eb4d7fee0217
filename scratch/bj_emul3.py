"""Vectorised numpy emulation of the block one-sided Jacobi round kernel (scalar rotations inside the
panel Gram, batched over the block pairs of a round): sweeps to convergence for several variants."""
import sys, numpy as np, time
G = np.load('scratch/G_c2.npy').astype(np.float64); G = (G + G.T) / 2
R = G.shape[0]; nrm = np.linalg.norm(G)

def rr(n, rnd):
    m = n - 1
    idx = np.arange(1, n // 2)
    return np.r_[rnd, (rnd + idx) % m], np.r_[m, (rnd - idx + m) % m]

def inner_pairs(OB, mode):
    """list of rounds; each round = (p[], q[]) disjoint index pairs in the 2*OB panel"""
    out = []
    if mode == 'cross':
        for r in range(OB):
            a = np.arange(OB); out.append((a, OB + (a + r) % OB))
    elif mode == 'intra':
        for r in range(OB - 1):
            x, y = rr(OB, r)
            p = np.minimum(x, y); q = np.maximum(x, y)
            out.append((np.r_[p, OB + p], np.r_[q, OB + q]))
    elif mode == 'full':
        for r in range(2 * OB - 1):
            x, y = rr(2 * OB, r)
            out.append((np.minimum(x, y), np.maximum(x, y)))
    return out

def rotate_batch(H, rounds, tol, abs2):
    """H [B, n, n] -> Q [B, n, n]; applies the rounds of scalar rotations (two-sided on H)"""
    B, n, _ = H.shape
    H = H.copy(); Q = np.tile(np.eye(n), (B, 1, 1))
    ar = np.arange(B)[:, None]
    for p, q in rounds:
        hpp = H[:, p, p]; hqq = H[:, q, q]; hpq = H[:, p, q]
        need = (hpq * hpq > tol * tol * np.abs(hpp * hqq)) & (np.abs(hpq) > abs2)
        with np.errstate(all='ignore'):
            z = (hqq - hpp) / (2 * hpq)
            t = np.where(z >= 0, 1.0, -1.0) / (np.abs(z) + np.sqrt(1 + z * z))
        c = 1 / np.sqrt(1 + t * t); s = t * c
        c = np.where(need, c, 1.0); s = np.where(need, s, 0.0)
        Rm = np.tile(np.eye(n), (B, 1, 1))
        Rm[ar, p, p] = c; Rm[ar, q, q] = c; Rm[ar, p, q] = s; Rm[ar, q, p] = -s
        H = np.einsum('bji,bjk,bkl->bil', Rm, H, Rm); Q = Q @ Rm
    return Q

def maxcos(W):
    H = W.T @ W; d = np.sqrt(np.maximum(np.diag(H), 1e-300))
    C = np.abs(H) / np.outer(d, d); np.fill_diagonal(C, 0); return C.max()

def run(W, OB, variant, label, tol=1e-5, max_sweeps=24, inner_sweeps=1):
    W = W / np.linalg.norm(W); abs2 = (1.2e-7) ** 2
    nb = W.shape[1] // OB
    cross, intra, full = inner_pairs(OB, 'cross'), inner_pairs(OB, 'intra'), inner_pairs(OB, 'full')
    t0 = time.time()
    for sweep in range(1, max_sweeps + 1):
        nrot = 0
        rounds = range(-1, nb - 1) if variant == 'current' else range(0, nb - 1)
        for rnd in rounds:
            if rnd < 0: a = 2 * np.arange(nb // 2); b = a + 1
            else: a, b = rr(nb, rnd)
            cols = np.concatenate([a[:, None] * OB + np.arange(OB)[None], b[:, None] * OB + np.arange(OB)[None]], axis=1)
            P = W[:, cols]
            H = np.einsum('rpi,rpj->pij', P, P)
            if variant == 'current':
                Q = rotate_batch(H, intra if rnd < 0 else cross, tol, abs2)
            else:
                Q = rotate_batch(H, full * inner_sweeps, tol, abs2)
            nrot += int((np.abs(Q - np.eye(2 * OB)).reshape(len(a), -1).max(1) > 0).sum())
            W[:, cols] = np.einsum('rpi,pij->rpj', P, Q)
        print(f'{label} OB={OB} {variant} sweep {sweep}: pairs rotated {nrot}  maxcos {maxcos(W):.2e}  t={time.time()-t0:.0f}s', flush=True)
        if nrot <= (nb // 2 * nb) // 256: break
    return sweep

which, variant = sys.argv[1], sys.argv[2]
OB = int(sys.argv[3]) if len(sys.argv) > 3 else 16
isw = int(sys.argv[4]) if len(sys.argv) > 4 else 1
import os
SHIFT = float(os.environ.get("SHIFT", "1e-3"))
A = G + SHIFT * nrm * np.eye(R)
if which == 'L': W = np.linalg.cholesky(A)
elif which == 'Ls':
    o = np.argsort(-np.diag(A)); W = np.linalg.cholesky(A[np.ix_(o, o)])
elif which == 'Lp':
    import scipy.linalg.lapack as lp
    c, piv, rank, info = lp.dpstrf(A, lower=1); W = np.tril(c)
elif which == 'L2':
    L = np.linalg.cholesky(A); W = np.linalg.cholesky(L.T @ L)
elif which == 'L3':
    L = np.linalg.cholesky(A); L = np.linalg.cholesky(L.T @ L); W = np.linalg.cholesky(L.T @ L)
elif which == 'Lp2':
    import scipy.linalg.lapack as lp
    c, piv, rank, info = lp.dpstrf(A, lower=1); L = np.tril(c); W = np.linalg.cholesky(L.T @ L)
run(W, OB, variant, which, inner_sweeps=isw)
