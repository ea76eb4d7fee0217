"""Block one-sided Jacobi on the c2 Gram, idealised inner solve (exact eigh of the 2*OB panel Gram):
sweeps to convergence for different preconditioners / column orders.  fp64 emulation."""
import sys, numpy as np, scipy.linalg as sla
G = np.load('scratch/G_c2.npy').astype(np.float64); G = (G + G.T) / 2
R = G.shape[0]
wref = np.linalg.eigvalsh(G)
nrm = np.linalg.norm(G)

def rr(n, rnd):
    m = n - 1
    idx = np.arange(1, n // 2)
    a = np.r_[rnd, (rnd + idx) % m]; b = np.r_[m, (rnd - idx + m) % m]
    return a, b

def maxcos(W):
    H = W.T @ W
    d = np.sqrt(np.maximum(np.diag(H), 1e-300))
    C = np.abs(H) / np.outer(d, d)
    np.fill_diagonal(C, 0)
    return C.max()

def run(W, OB, label, tol=1e-5, sort_inner=True, max_sweeps=20, inner='exact'):
    W = W / np.linalg.norm(W)
    n = W.shape[1]
    nb = n // OB
    for sweep in range(1, max_sweeps + 1):
        nrot = 0
        for rnd in range(nb - 1):
            a, b = rr(nb, rnd)
            cols = np.concatenate([a[:, None] * OB + np.arange(OB)[None], b[:, None] * OB + np.arange(OB)[None]], axis=1)  # [pairs, 2OB]
            P = W[:, cols]  # [rows, pairs, 2OB]
            H = np.einsum('rpi,rpj->pij', P, P)
            d = np.sqrt(np.maximum(np.einsum('pii->pi', H), 1e-300))
            C = np.abs(H) / (d[:, :, None] * d[:, None, :])
            C[:, np.arange(2 * OB), np.arange(2 * OB)] = 0
            need = C.reshape(len(a), -1).max(1) > tol
            nrot += int(need.sum())
            if inner == 'exact':
                w, Q = np.linalg.eigh(H)
                if sort_inner:
                    Q = Q[:, :, ::-1]
                # make Q close to identity?  (not needed for one-sided)
            Q[~need] = np.eye(2 * OB)
            W[:, cols] = np.einsum('rpi,pij->rpj', P, Q)
        mc = maxcos(W)
        lam = np.sort(np.sum(W * W, axis=0))
        print(f'{label} OB={OB} sweep {sweep}: pairs rotated {nrot}/{(nb-1)*nb//2}  maxcos {mc:.2e}', flush=True)
        if nrot == 0 or mc < tol: break
    return sweep

which = sys.argv[1]; OB = int(sys.argv[2]) if len(sys.argv) > 2 else 16
shift = 1e-3 * nrm
A = G + shift * np.eye(R)
if which == 'G':
    run(G.copy(), OB, 'W=G')
elif which == 'L':
    run(np.linalg.cholesky(A), OB, 'W=chol')
elif which == 'Ls':   # sort diagonal descending first
    o = np.argsort(-np.diag(A)); L = np.linalg.cholesky(A[np.ix_(o, o)]); run(L, OB, 'W=chol(sorted diag)')
elif which == 'Lp':   # fully pivoted Cholesky
    import scipy.linalg.lapack as lp
    c, piv, rank, info = lp.dpstrf(A, lower=1)
    L = np.tril(c); print('rank', rank, info)
    run(L, OB, 'W=pivoted chol')
elif which == 'L2':   # two LR steps
    L = np.linalg.cholesky(A); L2 = np.linalg.cholesky(L.T @ L); run(L2, OB, 'W=chol(L^T L)')
elif which == 'Lt':
    L = np.linalg.cholesky(A); run(L.T.copy(), OB, 'W=chol^T')
elif which == 'Lpt':
    import scipy.linalg.lapack as lp
    c, piv, rank, info = lp.dpstrf(A, lower=1)
    L = np.tril(c); run(L.T.copy(), OB, 'W=pivoted chol^T')
