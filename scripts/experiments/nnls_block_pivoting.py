"""CPU experiment: block principal pivoting vs Lawson-Hanson on the C3 system (iteration counts, agreement)."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from scipy.optimize import nnls
from oracle import ref_port as rp
from pyneapple_b200 import synth

base = synth.CONFIGS["C3"]
b, y, _ = synth.sample_voxels(base, 400, z=0)
bins = rp.nnls_bins(0.0008, 0.5, 250)
A = np.concatenate([rp.nnls_basis(b, bins), rp.regularization_matrix(250, 2, 0.02)], axis=0)
G = A.T @ A
n = 250
print("cond(G) = %.3e" % np.linalg.cond(G))

def bpp(h, tol, max_it=100):
    F = np.zeros(n, bool)
    x = np.zeros(n); w = -h.copy()          # w = G x - h  (dual, must be >= 0 off the support)
    p_max = 3; p = p_max; best = n + 1
    for it in range(1, max_it + 1):
        infF = F & (x < 0); infG = (~F) & (w < -tol)
        nv = infF.sum() + infG.sum()
        if nv == 0:
            return x, it - 1, F
        if nv < best:
            best = nv; p = p_max; chF, chG = infF, infG
        elif p > 0:
            p -= 1; chF, chG = infF, infG
        else:  # backup rule: the largest infeasible index only
            idx = np.flatnonzero(infF | infG).max()
            chF = np.zeros(n, bool); chG = np.zeros(n, bool)
            (chF if F[idx] else chG)[idx] = True
        F = (F & ~chF) | chG
        x = np.zeros(n)
        if F.any():
            GF = G[np.ix_(F, F)]
            try:
                L = np.linalg.cholesky(GF)
                x[F] = np.linalg.solve(L.T, np.linalg.solve(L, h[F]))
            except np.linalg.LinAlgError:
                return None, -it, F
        w = G @ x - h
    return None, -max_it, F

its, ks, lh_its, bad = [], [], [], 0
maxdiff = 0.0
for v in range(y.shape[0]):
    yy = np.concatenate([y[v], np.zeros(n)])
    xr, rn = nnls(A, yy, maxiter=250)
    h = A.T @ yy
    tol = 10 * max(A.shape) * np.finfo(float).eps * np.abs(h).max()  # rough
    x, it, F = bpp(h, 0.0)
    if x is None:
        bad += 1; print("voxel", v, "failed", it); continue
    its.append(it); ks.append(int((xr > 0).sum()))
    d = np.abs(x - xr).max(); maxdiff = max(maxdiff, d)
    if d > 1e-6: print("voxel", v, "diff", d, "k_ref", ks[-1], "k_bpp", int(F.sum()), "it", it)
its = np.array(its); ks = np.array(ks)
print("BPP iterations: mean %.2f  median %d  p95 %d  max %d ; failures %d" % (its.mean(), np.median(its), np.percentile(its, 95), its.max(), bad))
print("support size: mean %.1f max %d ; max |x - x_scipy| = %.3e" % (ks.mean(), ks.max(), maxdiff))
