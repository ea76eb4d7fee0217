"""Numpy emulation of a two-level block one-sided Jacobi: column blocks of width b, every round solves the
2b x 2b pair Gram COMPLETELY (np.linalg.eigh, eigenvector matrix permuted/sign-fixed to be closest to I)
and applies it.  Question for round 2: outer sweeps to convergence vs the current scalar-rotation kernel."""
import sys, time, numpy as np

def rr(n, rnd):
    m = n - 1
    idx = np.arange(1, n // 2)
    return np.r_[rnd, (rnd + idx) % m], np.r_[m, (rnd - idx + m) % m]

def closest_to_identity(Q):
    # greedy column permutation so that |Q_ii| is large, then sign fix (keeps the update small: convergence theory)
    n = Q.shape[0]
    A = np.abs(Q).copy(); perm = -np.ones(n, int)
    for _ in range(n):
        i, j = np.unravel_index(np.argmax(A), A.shape)
        perm[i] = j; A[i, :] = -1; A[:, j] = -1
    Q = Q[:, perm]
    return Q * np.sign(np.diag(Q))[None, :]

def maxcos(W):
    H = W.T @ W; d = np.sqrt(np.maximum(np.diag(H), 1e-300))
    C = np.abs(H) / np.outer(d, d); np.fill_diagonal(C, 0); return C.max()

def run(W, b, tol=1e-5, max_sweeps=20, dtype=np.float64):
    W = (W / np.linalg.norm(W)).astype(dtype)
    nb = W.shape[1] // b
    for sweep in range(1, max_sweeps + 1):
        t0 = time.time(); nrot = 0
        for rnd in range(-1, nb - 1):
            if rnd < 0: a = 2 * np.arange(nb // 2); bb = a + 1
            else: a, bb = rr(nb, rnd)
            for pa, pb in zip(a, bb):
                cols = np.r_[pa * b + np.arange(b), pb * b + np.arange(b)]
                P = W[:, cols]
                H = (P.T @ P).astype(np.float64)
                d = np.sqrt(np.maximum(np.diag(H), 1e-300))
                C = np.abs(H) / np.outer(d, d); np.fill_diagonal(C, 0)
                if C.max() <= tol: continue
                nrot += 1
                _, Q = np.linalg.eigh(H)
                Q = closest_to_identity(Q)
                W[:, cols] = (P @ Q.astype(dtype))
        mc = maxcos(W.astype(np.float64))
        print(f'b={b} sweep {sweep}: pairs rotated {nrot} of {nb//2*nb}  maxcos {mc:.2e}  ({time.time()-t0:.0f}s)', flush=True)
        if mc < tol or nrot == 0: break
    return sweep

which = sys.argv[1]; b = int(sys.argv[2])
if which == 'c2':
    G = np.load('scratch/G_c2.npy').astype(np.float64); G = (G + G.T) / 2
else:
    R = int(which); rng = np.random.default_rng(0); rank = int(0.9 * R)
    B = rng.standard_normal((R, rank)) * np.logspace(0, -3, rank)
    G = B @ B.T
R = G.shape[0]
A = G + 1e-3 * np.linalg.norm(G) * np.eye(R)
L = np.linalg.cholesky(A)
run(L, b, dtype=np.float32 if len(sys.argv) > 3 else np.float64)


# ---- variant: inner solve = k cyclic sweeps of scalar two-sided Jacobi on the pair Gram (not a full eigh) ----
def inner_sweeps_Q(H, k, tol=1e-5):
    n = H.shape[0]; H = H.copy(); Q = np.eye(n)
    for _ in range(k):
        for r in range(n - 1):
            x, y = rr(n, r)
            p = np.minimum(x, y); q = np.maximum(x, y)
            hpp, hqq, hpq = H[p, p], H[q, q], H[p, q]
            need = (hpq * hpq > tol * tol * np.abs(hpp * hqq))
            with np.errstate(all='ignore'):
                z = (hqq - hpp) / (2 * hpq)
                t = np.where(z >= 0, 1.0, -1.0) / (np.abs(z) + np.sqrt(1 + z * z))
            c = 1 / np.sqrt(1 + t * t); s = t * c
            c = np.where(need, c, 1.0); s = np.where(need, s, 0.0)
            Rm = np.eye(n); Rm[p, p] = c; Rm[q, q] = c; Rm[p, q] = s; Rm[q, p] = -s
            H = Rm.T @ H @ Rm; Q = Q @ Rm
    return Q

def run_inner(W, b, k, tol=1e-5, max_sweeps=24):
    W = (W / np.linalg.norm(W)).astype(np.float32)
    nb = W.shape[1] // b
    for sweep in range(1, max_sweeps + 1):
        t0 = time.time(); nrot = 0
        for rnd in range(-1, nb - 1):
            if rnd < 0: a = 2 * np.arange(nb // 2); bb = a + 1
            else: a, bb = rr(nb, rnd)
            for pa, pb in zip(a, bb):
                cols = np.r_[pa * b + np.arange(b), pb * b + np.arange(b)]
                P = W[:, cols]
                H = (P.T @ P).astype(np.float64)
                d = np.sqrt(np.maximum(np.diag(H), 1e-300))
                C = np.abs(H) / np.outer(d, d); np.fill_diagonal(C, 0)
                if C.max() <= tol: continue
                nrot += 1
                W[:, cols] = P @ inner_sweeps_Q(H, k, tol).astype(np.float32)
        mc = maxcos(W.astype(np.float64))
        print(f'b={b} inner={k} sweep {sweep}: pairs rotated {nrot} of {nb//2*nb}  maxcos {mc:.2e}  ({time.time()-t0:.0f}s)', flush=True)
        if mc < tol or nrot == 0: break

if len(sys.argv) > 4:
    run_inner(L, b, int(sys.argv[4]))
