"""Numpy prototype of the Gram/Cholesky Lawson-Hanson used by the CUDA NNLS kernel (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ref_port, c_oracle


def nnls_gram(Bm, rtr, y, maxiter, refine=True):
    m, n = Bm.shape
    h = Bm.T @ y
    w = h.copy()
    x = np.zeros(n)
    P = []
    L = np.zeros((n, n))
    u = np.zeros(n)
    inP = np.zeros(n, bool)
    it = 0
    mode = 1
    G = lambda i, j: Bm[:, i] @ Bm[:, j] + rtr[i, j]

    def backsolve(k):
        z = np.zeros(k)
        for i in range(k - 1, -1, -1):
            z[i] = (u[i] - L[i + 1:k, i] @ z[i + 1:k]) / L[i, i]
        return z

    def remove(q):
        nonlocal P
        k = len(P)
        idx = P[q]
        x[idx] = 0.0
        inP[idx] = False
        Lk = np.delete(L[:k, :k], q, axis=0)  # (k-1, k)
        for i in range(q, k - 1):
            a, b = Lk[i, i], Lk[i, i + 1]
            if abs(a) > abs(b):
                xr = b / a; yr = np.sqrt(1 + xr * xr); c = np.copysign(1 / yr, a); s = c * xr; sig = abs(a) * yr
            elif b != 0:
                xr = a / b; yr = np.sqrt(1 + xr * xr); s = np.copysign(1 / yr, b); c = s * xr; sig = abs(b) * yr
            else:
                sig, c, s = 0.0, 0.0, 1.0
            ci, cj = Lk[:, i].copy(), Lk[:, i + 1].copy()
            Lk[:, i] = c * ci + s * cj
            Lk[:, i + 1] = -s * ci + c * cj
            Lk[i, i] = sig; Lk[i, i + 1] = 0.0
            ui, uj = u[i], u[i + 1]
            u[i] = c * ui + s * uj; u[i + 1] = -s * ui + c * uj
        L[:k - 1, :k - 1] = Lk[:, :k - 1]
        P = P[:q] + P[q + 1:]

    while True:
        k = len(P)
        if k >= n:
            break
        accepted = False
        while True:
            cand = np.where(~inP, w, -np.inf)
            j = int(np.argmax(cand))
            if not (cand[j] > 0):
                j = -1
                break
            g = np.array([G(p, j) for p in P])
            l = np.zeros(k)
            for c in range(k):
                l[c] = (g[c] - L[c, :c] @ l[:c]) / L[c, c]
            unorm = np.sqrt(l @ l)
            piv2 = G(j, j) - l @ l
            a = np.sqrt(max(piv2, 0.0))
            if (unorm + a * 0.01) - unorm > 0:
                t = (h[j] - l @ u[:k]) / a
                if t / a > 0:
                    accepted = True
                    break
            w[j] = 0.0
        if not accepted:
            break
        L[k, :k] = l; L[k, k] = a; u[k] = t
        P.append(j); inP[j] = True; w[j] = 0.0
        z = backsolve(len(P))
        fail = False
        while True:
            it += 1
            if it >= maxiter:
                mode = 3; fail = True
                break
            alpha, jj = 2.0, -1
            for ip, idx in enumerate(P):
                if z[ip] <= 0:
                    tt = -x[idx] / (z[ip] - x[idx])
                    if alpha > tt:
                        alpha, jj = tt, ip
            if jj < 0:
                break
            for ip, idx in enumerate(P):
                x[idx] += alpha * (z[ip] - x[idx])
            remove(jj)
            while True:
                bad = [ip for ip, idx in enumerate(P) if x[idx] <= 0]
                if not bad:
                    break
                remove(bad[0])
            z = backsolve(len(P))
        if fail:
            break
        for ip, idx in enumerate(P):
            x[idx] = z[ip]
        r = y - Bm[:, P] @ x[P]
        w = Bm.T @ r - rtr @ x
        if refine and len(P):
            k = len(P)
            # one step of iterative refinement with the true residual (corrected semi-normal equations)
            rhs = w[P]
            t1 = np.zeros(k)
            for c in range(k):
                t1[c] = (rhs[c] - L[c, :c] @ t1[:c]) / L[c, c]
            dz = np.zeros(k)
            for i in range(k - 1, -1, -1):
                dz[i] = (t1[i] - L[i + 1:k, i] @ dz[i + 1:k]) / L[i, i]
            xn = x[P] + dz
            if (xn > 0).all():
                x[P] = xn
        w[P] = 0.0
    if mode == 3:
        return np.zeros(n), np.linalg.norm(y), 3, it
    r = y - Bm @ x
    return x, np.sqrt(r @ r + x @ (rtr @ x)), 1, it


if __name__ == "__main__":
    sys.path.insert(0, "tests")
    from _util import load
    for name in ["nnls_c3_reg2", "nnls_c3_reg1", "nnls_c3_reg3", "nnls_c3_reg0", "nnls_small_reg1", "nnls_c3_degenerate", "nnls_c3_maxiter20"]:
        g = load(name)
        nb = int(g["n_bins"]); bins = ref_port.nnls_bins(g["d_range"][0], g["d_range"][1], nb)
        Bm = ref_port.nnls_basis(g["b"], bins); R = ref_port.regularization_matrix(nb, int(g["reg_order"]), float(g["mu"]))
        rtr = R.T @ R
        A = np.concatenate([Bm, R]); Bx = np.concatenate([g["y"], np.zeros((g["y"].shape[0], nb))], 1)
        ref = c_oracle.nnls(A, Bx, int(g["max_iter"]))
        nv = min(64, g["y"].shape[0])
        for refine in (False, True):
            errs, rerr, dit, st = [], [], [], []
            for v in range(nv):
                x, rn, mode, it = nnls_gram(Bm, rtr, g["y"][v], int(g["max_iter"]), refine)
                errs.append(np.abs(x - g["coefficients"][v]).max()); rerr.append(abs(rn - g["residual"][v])); dit.append(it - ref["iters"][v]); st.append((mode == 1) == bool(g["success"][v]))
            print(f"{name:20s} refine={refine} maxabs={max(errs):.2e} med={np.median(errs):.1e} rnorm={max(rerr):.1e} iter_diff={np.abs(dit).max()} success_same={all(st)}")
