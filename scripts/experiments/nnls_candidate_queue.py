"""CPU experiment: Lawson-Hanson with a queue of the top-q dual candidates per full dual pass."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from scipy.optimize import nnls
from oracle import ref_port as rp
from pyneapple_b200 import synth

base = synth.CONFIGS["C3"]
b, y, _ = synth.sample_voxels(base, 300, z=0)
bins = rp.nnls_bins(0.0008, 0.5, 250)
A = np.concatenate([rp.nnls_basis(b, bins), rp.regularization_matrix(250, 2, 0.02)], axis=0)
G = A.T @ A
n = 250

def lh(h, q, tol, spread=0):
    P = []; x = np.zeros(n)
    passes = adds = removes = skipped = 0
    while True:
        w = h - G @ x; passes += 1
        w[P] = -np.inf
        order = np.argsort(-w)
        cand = []
        for j in order:
            if w[j] <= tol or len(cand) >= q: break
            if spread and any(abs(j - c) < spread for c in cand): continue
            cand.append(j)
        if not cand: break
        for ci, j in enumerate(cand):
            if ci > 0:
                if j in P: continue
                wj = h[j] - G[j] @ x
                if wj <= tol: skipped += 1; continue
            P.append(j); adds += 1
            while True:
                s = np.zeros(n)
                s[P] = np.linalg.solve(G[np.ix_(P, P)], h[P])
                if (s[P] > 0).all(): x = s; break
                neg = [i for i in P if s[i] <= 0]
                alpha = min(x[i] / (x[i] - s[i]) for i in neg)
                x = x + alpha * (s - x)
                drop = [i for i in P if x[i] <= 1e-15 * max(1.0, abs(x).max()) and s[i] <= 0]
                if not drop: drop = [min(neg, key=lambda i: x[i] / (x[i] - s[i]))]
                for i in drop: P.remove(i); x[i] = 0.0; removes += 1
                if not P: break
            if adds > 2000: return x, passes, adds, removes, skipped
    return x, passes, adds, removes, skipped

for q, spread in ((1, 0), (2, 0), (4, 0), (8, 0), (4, 3), (4, 8), (8, 8)):
    tot = np.zeros(4); md = 0.0
    for v in range(y.shape[0]):
        yy = np.concatenate([y[v], np.zeros(n)])
        h = A.T @ yy
        if q == 1 and v < 300 or True:
            pass
        x, *c = lh(h, q, 1e-12 * abs(h).max(), spread)
        tot += c
        if v < 40:
            xr, _ = nnls(A, yy, maxiter=250)
            md = max(md, np.abs(x - xr).max())
    tot /= y.shape[0]
    print(f"q={q} spread={spread}: dual passes {tot[0]:5.1f}  adds {tot[1]:5.1f}  removes {tot[2]:5.1f}  stale-skipped {tot[3]:5.1f}  max|dx| vs scipy (40 vox) {md:.2e}", flush=True)
