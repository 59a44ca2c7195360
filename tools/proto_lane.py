"""Numpy prototype of the lane-per-solve scan solver (ibs_scan_solver.cu): validates the division-free
recurrence, the scaled S accumulation, the fixed matching row, the node count, the multigrid start and the
evaluation counts against the oracle.  Not part of the product path.

    python tools/proto_lane.py [fixture] [nth0]
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import ballooning_oracle as bo  # noqa: E402


def poly_arrays(D, i, j, stride=1):
    """G0,G1,G2 (g), C0,C1 (h^2 c), R (h^2 f = g R) of fixture line (i, j) on the grid with the given stride."""
    B = D["geo_bmag"][i, j][::stride]
    gp = np.abs(D["geo_gradpar_theta_pest"][i, j][::stride])
    cv, cv0 = D["geo_cvdrift"][i, j][::stride], D["geo_cvdrift0"][i, j][::stride]
    g0, g1, g2 = D["geo_gds2"][i, j][::stride], D["geo_gds21"][i, j][::stride], D["geo_gds22"][i, j][::stride]
    th = D["theta"][::stride]
    h = th[1] - th[0]
    dP = D["dPdrho"][i, j]
    gpB = gp * B
    mdP = -h * h * dP / gpB
    return dict(G0=gp / B * g0, G1=2 * gp / B * g1, G2=gp / B * g2, C0=mdP * cv, C1=mdP * cv0, R=h * h / gpB ** 2,
                h=h, N=len(th))


def coefs(P, th0):
    """per-lane coefficient rows: g (L,N), C, F"""
    t = th0[:, None]
    g = P["G0"][None] + t * (P["G1"][None] + t * P["G2"][None])
    C = P["C0"][None] + t * P["C1"][None]
    F = g * P["R"][None]
    return g, C, F


def rescale(X, W, S, E):
    m = np.maximum(np.abs(X), np.abs(W))
    e = np.where(m > 0, np.floor(np.log2(np.where(m > 0, m, 1.0))), 0.0)
    s = np.exp2(-e)
    return X * s, W * s, S * s * s, E + e


def evaluate(g, C, F, lam, k, want_z=False):
    """One evaluation at shift lam (L,) with matching row k (scalar, 1 <= k <= M).  Returns r', S, nodes (and z)."""
    L, N = g.shape
    M = N - 2
    tp = 2.0 * (C - lam[:, None] * F)                    # t' = 2 (C - lam F)
    a = g[:, 1:] + g[:, :-1]                             # a[j] = 2 gh between j and j+1, j = 0..N-2
    # forward: rows 1..k
    X = np.zeros(L); W = np.ones(L); S = np.zeros(L); E = np.zeros(L)
    nodes = np.zeros(L, dtype=int)
    sgn = np.zeros(L, dtype=bool)
    if want_z:
        zf = np.zeros((L, N)); zfE = np.zeros((L, N))
    for j in range(1, k + 1):
        aj = a[:, j - 1]
        Xn = aj * X + W
        Wn = aj * W - tp[:, j] * Xn
        S = aj * aj * S + F[:, j] * Xn * Xn
        X, W = Xn, Wn
        s2 = X < 0
        nodes += (s2 != sgn)
        sgn = s2
        if want_z:
            zf[:, j] = X; zfE[:, j] = E
        if j % 16 == 0:
            X, W, S, E = rescale(X, W, S, E)
            # the stored z are in the scale valid at the time; fix below with the exponent record
    Xf, Wf, Sf, Ef = X, W, S, E
    # backward: start at row M, steps down to row k
    X = np.ones(L); W = -a[:, M]; S = F[:, M] * 1.0; E = np.zeros(L)
    sgn = np.zeros(L, dtype=bool)
    if want_z:
        zb = np.zeros((L, N)); zbE = np.zeros((L, N))
        zb[:, M] = X
    for j in range(M, k, -1):                            # from row j to row j-1
        tmp = W + tp[:, j] * X
        aj = a[:, j - 1]
        Xn = aj * X - tmp
        Wn = aj * tmp
        S = aj * aj * S + (F[:, j - 1] * Xn * Xn if j - 1 > k else 0.0)
        X, W = Xn, Wn
        s2 = X < 0
        nodes += (s2 != sgn)
        sgn = s2
        if want_z:
            zb[:, j - 1] = X; zbE[:, j - 1] = E
        if (M - j) % 16 == 15:
            X, W, S, E = rescale(X, W, S, E)
    Xb, Wb, Sb = X, W, S
    r = Wb / Xb - Wf / Xf
    Stot = Sf / (Xf * Xf) + Sb / (Xb * Xb)
    if not want_z:
        return r, Stot, nodes
    return r, Stot, nodes, (zf, zfE, Xf, Ef, zb, zbE, Xb, E)


def unscaled_z(g, C, F, lam, k):
    """Matched vector z (z_k = 1) by the plain (divided) recurrence in log-safe form -- for the output pass."""
    L, N = g.shape
    M = N - 2
    t = C - lam[:, None] * F
    gh = 0.5 * (g[:, 1:] + g[:, :-1])
    z = np.zeros((L, N))
    # forward with running renormalisation
    x = np.zeros(L); w = np.ones(L)
    xs = np.zeros((L, N)); es = np.zeros((L, N)); e = np.zeros(L)
    for j in range(1, k + 1):
        x = x + w / gh[:, j - 1]
        w = w - t[:, j] * x
        xs[:, j] = x; es[:, j] = e
        if j % 16 == 0:
            m = np.floor(np.log2(np.maximum(np.abs(x), np.abs(w)))); x = x * np.exp2(-m); w = w * np.exp2(-m); e = e + m
    xk = xs[:, k]; ek = es[:, k]
    z[:, 1:k + 1] = xs[:, 1:k + 1] / xk[:, None] * np.exp2(es[:, 1:k + 1] - ek[:, None])
    x = 1.0 / gh[:, M]; w = -np.ones(L); e = np.zeros(L)
    xs[:, M] = x; es[:, M] = e
    for j in range(M, k, -1):
        w = w + t[:, j] * x
        x = x - w / gh[:, j - 1]
        xs[:, j - 1] = x; es[:, j - 1] = e
        if (M - j) % 16 == 15:
            m = np.floor(np.log2(np.maximum(np.abs(x), np.abs(w)))); x = x * np.exp2(-m); w = w * np.exp2(-m); e = e + m
    xk = xs[:, k]; ek = es[:, k]
    z[:, k + 1:M + 1] = xs[:, k + 1:M + 1] / xk[:, None] * np.exp2(es[:, k + 1:M + 1] - ek[:, None])
    return z


def iterate(g, C, F, lam0, lo, hi, U, k, warm, maxit=64, tol_rel=2.0 ** -49, stop_rel=None):
    """The bracketed Rayleigh-quotient iteration of ibs_solver.cu (per lane), vectorised evaluation.
    Returns lam (evaluated shift of the last evaluation), rho, its (number of evaluations per lane)."""
    L = g.shape[0]
    lam = lam0.copy(); rho = lam0.copy()
    lo = lo.copy(); hi = hi.copy()
    its = np.zeros(L, dtype=int)
    active = np.ones(L, dtype=bool)
    st = [dict(b1=0.0, N1=0.0, b2=0.0, N2=0.0, dprev=1e300, nabove=0, collapsed=False) for _ in range(L)]
    tol = tol_rel * np.maximum(np.abs(U), 1e-3)
    tol_stag = 1e-10 * np.maximum(np.abs(U), 1e-3)
    stop = tol if stop_rel is None else stop_rel * np.maximum(np.abs(U), 1e-3)
    nev = 0
    while active.any() and nev < maxit:
        r, S, nodes = evaluate(g, C, F, lam, k)
        nev += 1
        for l in np.nonzero(active)[0]:
            s = st[l]
            its[l] += 1
            rh = lam[l] + r[l] / (2.0 * S[l])
            rho[l] = rh
            pos = nodes[l] == 0
            above = pos and not (r[l] > 0)
            inbasin = pos and (r[l] > 0)
            if above:
                hi[l] = min(hi[l], lam[l])
                s["b1"], s["N1"] = s["b2"], s["N2"]
                s["b2"], s["N2"] = lam[l], lam[l] - rh
                s["nabove"] += 1
                if rh == rh: lo[l] = max(lo[l], min(rh, hi[l]))
            else:
                lo[l] = max(lo[l], lam[l])
                if inbasin and rh == rh: lo[l] = max(lo[l], min(rh, hi[l]))
            done = False
            if pos:
                dl = abs(rh - lam[l])
                if dl <= stop[l] or (dl < tol_stag[l] and dl >= 0.25 * s["dprev"]): done = True
                s["dprev"] = dl
            if not done and s["collapsed"]: done = True
            if not done:
                if hi[l] - lo[l] <= tol[l]:
                    nxt = 0.5 * (lo[l] + hi[l]); s["collapsed"] = True
                elif inbasin and rh > lam[l] and rh <= hi[l]:
                    nxt = rh
                elif above:
                    pw = 0.5
                    if s["nabove"] >= 2 and s["N1"] - s["N2"] > 0: pw = (s["b1"] - s["b2"]) / (s["N1"] - s["N2"])
                    pw = min(1.0, max(0.4, pw))
                    if pw > 0.8: pw = 1.0
                    if warm and s["nabove"] == 1: pw = 1.0
                    if s["N2"] < 0.05 * (U[l] - s["b2"]): pw = 1.0
                    nxt = s["b2"] - pw * s["N2"]
                    if not (nxt >= lo[l] and nxt < hi[l]): nxt = 0.5 * (lo[l] + hi[l])
                else:
                    nxt = 0.5 * (lo[l] + hi[l])
                if nxt == lam[l]: done = True
                else: lam[l] = nxt
            if done: active[l] = False
    return lam, rho, its, nev


def bounds(g, C, F):
    a = g[:, 1:] + g[:, :-1]
    U = np.max(C[:, 1:-1] / F[:, 1:-1], axis=1)
    U = U + 1e-6 * np.abs(U) + 1e-300
    numer = np.min(C[:, 1:-1], axis=1) - 4.0 * np.max(0.5 * a, axis=1)
    Lb = 1.000001 * np.where(numer < 0, numer / np.min(F[:, 1:-1], axis=1), numer / np.max(F[:, 1:-1], axis=1)) - 1e-300
    return Lb, U


def main():
    fx = sys.argv[1] if len(sys.argv) > 1 else "synthetic_d3d"
    nth0 = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    D = np.load(f"tests/golden/{fx}.npz")
    i, j = 1, 0
    th0 = np.linspace(0, np.pi / 2, nth0)
    P = poly_arrays(D, i, j)
    g, C, F = coefs(P, th0)
    N = P["N"]; M = N - 2
    k = (M + 1) // 2
    # reference eigenvalues (LAPACK, lambda_max of the same pencil)
    lam_ref = []
    for t0 in th0:
        cv = D["geo_cvdrift"][i, j] + t0 * D["geo_cvdrift0"][i, j]
        gd = D["geo_gds2"][i, j] + 2 * t0 * D["geo_gds21"][i, j] + t0 ** 2 * D["geo_gds22"][i, j]
        info = {}
        gam, X, dX, *_ = bo.gamma_ball_full(D["dPdrho"][i, j], D["theta"], D["geo_bmag"][i, j], D["geo_gradpar_theta_pest"][i, j],
                                            cv, gd, method="lambda_max", info=info)
        lam_ref.append((info["lambda_matrix"], gam, X, info["gap"]))
    lm = np.array([x[0] for x in lam_ref])
    print("lambda_matrix ref", lm[:4], "gap", [x[3] for x in lam_ref][:4])

    # --- check count semantics: nodes + (r>0) == #eigenvalues above lam
    Lb, U = bounds(g, C, F)
    for frac in (0.0, 0.3):
        lamq = lm - frac * np.abs(lm) - 1e-6
        r, S, nodes = evaluate(g, C, F, lamq, k)
        print("count at lam_max - small:", (nodes + (r > 0))[:6])
    r, S, nodes = evaluate(g, C, F, lm + 1e-9 * np.abs(lm), k)
    print("count just above:", (nodes + (r > 0))[:6], " rho-lam_ref:", (lm + 1e-9 * np.abs(lm) + r / (2 * S) - lm)[:4])

    # --- cold start on the fine grid
    lam, rho, its, nev = iterate(g, C, F, U.copy(), Lb, U, U, k, warm=False)
    print(f"fine cold: evals/lane mean {its.mean():.2f} max {its.max()} warp evals {nev}; err {np.max(np.abs(rho - lm) / np.abs(lm)):.2e}")

    # --- multigrid: strides 8, 4, 2 -> 1
    cost = 0.0
    lam_lv = {}
    prev = None
    for stride in (8, 4, 2):
        Pc = poly_arrays(D, i, j, stride)
        gc_, Cc, Fc = coefs(Pc, th0)
        Lbc, Uc = bounds(gc_, Cc, Fc)
        kc = (Pc["N"] - 2 + 1) // 2
        if prev is None:
            l0 = Uc.copy(); warm = False
        else:
            l0 = np.clip(prev, Lbc + 1e-300, Uc); warm = True
        lamc, rhoc, itc, nevc = iterate(gc_, Cc, Fc, l0, Lbc, Uc, Uc, kc, warm, stop_rel=1e-7)
        lam_lv[stride] = rhoc
        cost += nevc / stride
        if stride == 8:
            prev = rhoc
        else:
            prev = rhoc - (lam_lv[stride * 2] - rhoc) / 4.0      # Richardson estimate of the next finer level
        print(f"stride {stride}: warp evals {nevc} (mean {itc.mean():.2f}), est. err of next-level start {np.max(np.abs(prev - lm) / np.abs(lm)):.2e}")
    l0 = np.clip(prev, Lb, U)
    lam, rho, its, nev = iterate(g, C, F, l0, Lb, U, U, k, warm=True)
    cost += nev
    print(f"fine warm: warp evals {nev} (mean {its.mean():.2f}); total fine-equivalents {cost:.2f}; err {np.max(np.abs(rho - lm) / np.abs(lm)):.2e}")

    # --- output pass: z at lam, normalise, Simpson RQ
    z = unscaled_z(g, C, F, lam, k)
    Xn = z / np.max(np.abs(z), axis=1)[:, None]
    h = P["h"]
    err_g = 0.0; err_X = 0.0
    for l in range(len(th0)):
        X = Xn[l]
        gam, X2, dX = bo.postprocess(X[1:-1], h, g[l], C[l] / h ** 2, F[l] / h ** 2)
        err_g = max(err_g, abs(gam - lam_ref[l][1]) / abs(lam_ref[l][1]))
        Xr = lam_ref[l][2]
        Xr = Xr * np.sign(Xr[np.argmax(np.abs(Xr))])
        err_X = max(err_X, np.max(np.abs(X2 - Xr)))
    print(f"gam rel err {err_g:.2e}  X abs err {err_X:.2e}")


if __name__ == "__main__":
    main()
