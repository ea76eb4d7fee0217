"""numpy emulation of one-sided block Jacobi variants on the real c2 Gram: sweeps to convergence."""
import sys, numpy as np, time
f = np.float32
G = np.load('scratch/G_c2.npy').astype(np.float64)
G = (G + G.T) / 2
R = G.shape[0]
wref = np.linalg.eigvalsh(G)

def rr_pair(n, rnd, idx):
    m = n - 1
    if idx == 0: return rnd, m
    return (rnd + idx) % m, (rnd - idx + m) % m

def offcos(W, floor):
    H = W.T.astype(np.float64) @ W.astype(np.float64)
    d = np.sqrt(np.maximum(np.diag(H), 0))
    live = d > floor
    C = np.abs(H) / np.maximum(np.outer(d, d), 1e-300)
    C[~live, :] = 0; C[:, ~live] = 0
    np.fill_diagonal(C, 0)
    return C.max()

def inner_single_pass(H, intra, tol, abs2):
    """emulate the current kernel: one pass of rotations over cross (or intra) pairs"""
    n = H.shape[0]; b = n // 2
    H = H.copy(); Q = np.eye(n, dtype=H.dtype)
    rounds = b - 1 if intra else b
    for r in range(rounds):
        pairs = []
        for a in range(b):
            if not intra:
                p, q = a, b + ((a + r) % b)
            else:
                half, idx = a // (b // 2), a % (b // 2)
                x, y = rr_pair(b, r, idx)
                p, q = half * b + min(x, y), half * b + max(x, y)
            pairs.append((p, q))
        Rm = np.eye(n, dtype=H.dtype)
        for p, q in pairs:
            hpp, hqq, hpq = H[p, p], H[q, q], H[p, q]
            if hpq * hpq > tol * tol * abs(hpp * hqq) and abs(hpq) > abs2:
                z = (hqq - hpp) / (2 * hpq)
                t = np.sign(z) / (abs(z) + np.sqrt(1 + z * z)) if z != 0 else 1.0
                c = 1 / np.sqrt(1 + t * t); s = t * c
                Rm[p, p] = c; Rm[q, q] = c; Rm[p, q] = s; Rm[q, p] = -s
        H = Rm.T @ H @ Rm; Q = Q @ Rm
    return Q

def inner_full(H):
    w, Q = np.linalg.eigh(H.astype(np.float64))
    Q = Q[:, ::-1]  # descending
    # make it close to a permutation-free rotation: greedy column assignment to maximise the diagonal
    n = H.shape[0]
    return Q.astype(H.dtype)

def run(W0, OB, mode, dtype, tol, max_sweeps=30, sort_cols=False, accumulate=False, label=''):
    W = W0.astype(dtype).copy()
    Np = W.shape[1]
    nb = Np // OB
    scale = np.linalg.norm(W0)
    abs2 = (np.finfo(dtype).eps * 1.0) ** 2
    W = W / dtype(scale)
    t0 = time.time()
    for sweep in range(1, max_sweeps + 1):
        rotated = 0
        for rnd in range(-1, nb - 1):
            for pair in range(nb // 2):
                if rnd < 0: ba, bb = 2 * pair, 2 * pair + 1
                else: ba, bb = rr_pair(nb, rnd, pair)
                cols = np.r_[ba * OB:(ba + 1) * OB, bb * OB:(bb + 1) * OB]
                P = W[:, cols]
                H = P.T @ P
                d = np.sqrt(np.abs(np.diag(H)))
                if mode == 'single':
                    Q = inner_single_pass(H, rnd < 0, tol, abs2)
                    if not np.allclose(Q, np.eye(2 * OB)): rotated += 1
                else:
                    C = np.abs(H) / np.maximum(np.outer(d, d), 1e-300)
                    C[np.abs(H) <= abs2] = 0
                    np.fill_diagonal(C, 0)
                    if C.max() <= tol: continue
                    rotated += 1
                    Q = inner_full(H)
                W[:, cols] = P @ Q
        mc = offcos(W, np.finfo(dtype).eps)
        lam = np.sort(np.sum(W.astype(np.float64) ** 2, axis=0)) ** (0.5 if mode_is_G[0] else 1.0) * (scale if mode_is_G[0] else scale ** 2)
        print(f'{label} sweep {sweep}: rotated {rotated}/{nb//2*nb} maxcos {mc:.2e}  t={time.time()-t0:.0f}s', flush=True)
        if rotated == 0: break
    return sweep

mode_is_G = [True]
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
dt = np.float32
if which in ('A', 'all'):
    run(G, 16, 'single', dt, 1e-5, label='A: W=G OB16 single-pass')
if which in ('B', 'all'):
    run(G, 16, 'full', dt, 1e-5, label='B: W=G OB16 full-inner')
if which in ('B32', 'all'):
    run(G, 32, 'full', dt, 1e-5, label='B32: W=G OB32 full-inner')
if which in ('B64', 'all'):
    run(G, 64, 'full', dt, 1e-5, label='B64: W=G OB64 full-inner')
if which in ('C', 'all'):
    mode_is_G[0] = False
    eps = 1e-6 * np.trace(G) / R * 0 + 4e-7 * wref[-1]
    L = np.linalg.cholesky(G + eps * np.eye(R))
    run(L, 32, 'full', dt, 1e-5, label='C: W=chol(G+eps) OB32 full-inner')
    # sorted diagonal
if which in ('Ct', 'all'):
    mode_is_G[0] = False
    eps = 4e-7 * wref[-1]
    L = np.linalg.cholesky(G + eps * np.eye(R))
    run(L.T.copy(), 32, 'full', dt, 1e-5, label='Ct: W=chol(G+eps)^T OB32 full-inner')
