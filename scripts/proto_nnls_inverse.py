"""Numpy prototype: Lawson-Hanson on an explicitly updated inverse of the active Gram block (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ref_port, c_oracle


def nnls_inv(Bm, rtr, y, maxiter, refine_every=1):
    m, n = Bm.shape
    h = Bm.T @ y
    w = h.copy()
    x = np.zeros(n)
    P = []            # active bins (slot order)
    H = np.zeros((0, 0))
    z = np.zeros(0)
    inP = np.zeros(n, bool)
    it = 0
    mode = 1
    G = Bm.T @ Bm + rtr
    nout = 0
    while True:
        k = len(P)
        if k >= n:
            break
        accepted = False
        while True:
            cand = np.where(~inP, w, -np.inf)
            j = int(np.argmax(cand))
            if not (cand[j] > 0):
                break
            g = G[P, j]
            v = H @ g
            unorm2 = g @ v
            s = G[j, j] - unorm2
            a = np.sqrt(max(s, 0.0)); unorm = np.sqrt(max(unorm2, 0.0))
            if (unorm + a * 0.01) - unorm > 0:
                zeta = (h[j] - g @ z) / s
                if zeta > 0:
                    accepted = True
                    break
            w[j] = 0.0
        if not accepted:
            break
        # bordering
        Hn = np.zeros((k + 1, k + 1))
        Hn[:k, :k] = H + np.outer(v, v) / s
        Hn[:k, k] = -v / s; Hn[k, :k] = -v / s; Hn[k, k] = 1.0 / s
        H = Hn
        z = np.append(z - v * zeta, zeta)
        P.append(j); inP[j] = True; w[j] = 0.0
        fail = False
        while True:
            it += 1
            if it >= maxiter:
                mode = 3; fail = True
                break
            xp = x[P]
            neg = z <= 0
            if not neg.any():
                break
            t = np.where(neg, -xp / np.where(neg, z - xp, 1.0), np.inf)
            jj = int(np.argmin(t)); alpha = t[jj]
            x[P] = xp + alpha * (z - xp)

            def remove(q):
                nonlocal H, z, P
                hq = H[:, q].copy(); zq = z[q]; d = H[q, q]
                H = H - np.outer(hq, hq) / d
                z = z - hq * (zq / d)
                keep = [i for i in range(len(P)) if i != q]
                x[P[q]] = 0.0; inP[P[q]] = False
                H = H[np.ix_(keep, keep)]; z = z[keep]; P = [P[i] for i in keep]
            remove(jj)
            while True:
                bad = [ip for ip, idx in enumerate(P) if x[idx] <= 0]
                if not bad:
                    break
                remove(bad[0])
        if fail:
            break
        x[P] = z
        r = y - Bm[:, P] @ z
        w = Bm.T @ r - rtr @ x
        nout += 1
        if refine_every and nout % refine_every == 0 and len(P):
            dz = H @ w[P]
            zn = z + dz
            if (zn > 0).all():
                z = zn; x[P] = z
        w[P] = 0.0
    if mode == 3:
        return np.zeros(n), np.linalg.norm(y), 3, it
    # final polish: two refinement steps with freshly computed residuals
    for _ in range(2):
        if len(P):
            r = y - Bm[:, P] @ x[P]
            wp = (Bm.T @ r - rtr @ x)[P]
            zn = x[P] + H @ wp
            if (zn > 0).all():
                x[P] = zn
    r = y - Bm @ x
    return x, np.sqrt(r @ r + x @ (rtr @ x)), 1, it


if __name__ == "__main__":
    sys.path.insert(0, "tests")
    from _util import load
    for name in ["nnls_c3_reg2", "nnls_c3_reg1", "nnls_c3_reg3", "nnls_c3_reg0", "nnls_small_reg1", "nnls_c3_degenerate"]:
        g = load(name)
        nb = int(g["n_bins"]); bins = ref_port.nnls_bins(g["d_range"][0], g["d_range"][1], nb)
        Bm = ref_port.nnls_basis(g["b"], bins); R = ref_port.regularization_matrix(nb, int(g["reg_order"]), float(g["mu"]))
        rtr = R.T @ R
        A = np.concatenate([Bm, R]); Bx = np.concatenate([g["y"], np.zeros((g["y"].shape[0], nb))], 1)
        ref = c_oracle.nnls(A, Bx, int(g["max_iter"]))
        nv = min(64, g["y"].shape[0])
        for re in (1, 0):
            errs, dit, st = [], [], []
            for v in range(nv):
                x, rn, mode, it = nnls_inv(Bm, rtr, g["y"][v], int(g["max_iter"]), re)
                errs.append(np.abs(x - g["coefficients"][v]).max()); dit.append(it - ref["iters"][v]); st.append((mode == 1) == bool(g["success"][v]))
            print(f"{name:20s} refine_every={re} maxabs={max(errs):.2e} med={np.median(errs):.1e} iter_diff max={np.abs(dit).max()} n_diff={np.count_nonzero(dit)} success_same={all(st)}")

def study():
    sys.path.insert(0, "tests")
    from _util import load
    g = load("nnls_c3_reg2")
    nb = 250; bins = ref_port.nnls_bins(0.0008, 0.5, nb)
    Bm = ref_port.nnls_basis(g["b"], bins)
    for order in (2, 1):
        for mu in (0.02, 5e-3, 1e-3, 2e-4, 1e-5):
            R = ref_port.regularization_matrix(nb, order, mu); rtr = R.T @ R
            A = np.concatenate([Bm, R]); Bx = np.concatenate([g["y"][:24], np.zeros((24, nb))], 1)
            ref = c_oracle.nnls(A, Bx, 750)
            errs, dit = [], []
            for v in range(24):
                x, rn, mode, it = nnls_inv(Bm, rtr, g["y"][v], 750, 0)
                errs.append(np.abs(x - ref["x"][v]).max()); dit.append(it - ref["iters"][v])
            print(f"order {order} mu {mu:g}: maxabs {max(errs):.2e} iter_diff {np.abs(dit).max()} active max {(ref['x']>0).sum(1).max()}")
if len(sys.argv) > 1 and sys.argv[1] == "study":
    study()
