"""Numerical prototype (numpy) of the chunk-parallel shooting eigen-solver used by the CUDA
kernel.  Development aid only -- emulates T threads, each owning a contiguous chunk of rows.

Problem: (K - lam F) x = 0 on interior rows j=1..N-2, Dirichlet ends, with
  (K x)_j = gh_j (x_{j+1}-x_j) - gh_{j-1} (x_j - x_{j-1}) + C_j x_j,   C = h^2 c,  F = h^2 f.
State (x_j, w_j), w_j = gh_j (x_{j+1}-x_j):   x_j = x_{j-1} + w_{j-1}/gh_{j-1};  w_j = w_{j-1} - (C_j - lam F_j) x_j.
"""
import numpy as np


def setup(g, c, f, h):
    N = len(g)
    gh = g[:-1] + 0.5 * (g[1:] - g[:-1])          # half points j+1/2, j=0..N-2
    ig = 1.0 / gh
    C = h * h * c
    F = h * h * f
    return dict(N=N, M=N - 2, gh=gh, ig=ig, C=C, F=F, h=h, U=np.max(c[1:-1] / f[1:-1]),
                Lb=np.min((C[1:-1] - 2 * (gh[1:] + gh[:-1])) / F[1:-1]))


def seq_eval(P, lam):
    N, ig, C, F = P['N'], P['ig'], P['C'], P['F']
    x = np.zeros(N); w = 1.0
    for j in range(1, N - 1):
        x[j] = x[j - 1] + w * ig[j - 1]
        w = w - (C[j] - lam * F[j]) * x[j]
    x[N - 1] = x[N - 2] + w * ig[N - 2]
    s = np.sign(x[1:]); s[s == 0] = 1
    return x, int(np.sum(s[1:] != s[:-1]))


def _norm2(x, w):
    m = np.maximum(np.abs(x), np.abs(w))
    e = np.where(m > 0, np.floor(np.log2(np.where(m > 0, m, 1.0))), 0).astype(int)
    return np.ldexp(x, -e), np.ldexp(w, -e), e


def chunk_eval(P, lam, T=32, want_z=False):
    N, M, ig, C, F = P['N'], P['M'], P['ig'], P['C'], P['F']
    Lc = -(-M // T)
    j0 = 1 + np.arange(T) * Lc
    j1 = np.minimum(j0 + Lc, M + 1)
    j0 = np.minimum(j0, M + 1)
    t = C - lam * F
    a = np.ones(T); b = np.zeros(T); cc = np.zeros(T); d = np.ones(T)
    for i in range(Lc):
        j = j0 + i
        act = j < j1
        jj = np.where(act, j, 1)
        igm = ig[jj - 1]; tj = t[jj]
        na = a + cc * igm; nb = b + d * igm
        ncc = cc - tj * na; nd = d - tj * nb
        a = np.where(act, na, a); b = np.where(act, nb, b); cc = np.where(act, ncc, cc); d = np.where(act, nd, d)
    fx = np.zeros(T + 1); fw = np.zeros(T + 1); fe = np.zeros(T + 1, dtype=int)
    fx[0], fw[0] = 0.0, 1.0
    for tau in range(T):
        x, w = a[tau] * fx[tau] + b[tau] * fw[tau], cc[tau] * fx[tau] + d[tau] * fw[tau]
        fx[tau + 1], fw[tau + 1], e = _norm2(x, w); fe[tau + 1] = fe[tau] + e
    bx = np.zeros(T + 1); bw = np.zeros(T + 1); be = np.zeros(T + 1, dtype=int)
    bx[T], bw[T] = ig[N - 2], -1.0
    for tau in range(T - 1, -1, -1):
        x, w = d[tau] * bx[tau + 1] - b[tau] * bw[tau + 1], -cc[tau] * bx[tau + 1] + a[tau] * bw[tau + 1]
        bx[tau], bw[tau], e = _norm2(x, w); be[tau] = be[tau + 1] + e
    valid = (j1 - j0) > 0
    cand = np.arange(1, T + 1)[valid]
    with np.errstate(divide='ignore'):
        score = np.log2(np.abs(fx[cand] * bx[cand])) + fe[cand] + be[cand]
    kb = cand[np.argmax(score)]
    r = bw[kb] / bx[kb] - fw[kb] / fx[kb]
    # phase B (forward chains for chunks tau < kb) / phase C (backward chains for tau >= kb)
    tau = np.arange(T)
    fwd = tau < kb
    x = np.where(fwd, fx[:-1], bx[1:]); w = np.where(fwd, fw[:-1], bw[1:])
    scl = np.where(fwd, np.ldexp(1.0, np.clip(fe[:-1] - fe[kb], -1000, 1000)) / fx[kb],
                   np.ldexp(1.0, np.clip(be[1:] - be[kb], -1000, 1000)) / bx[kb])
    acc = np.zeros(T); nodes = np.zeros(T, dtype=int)
    z = np.zeros(N) if want_z else None
    prev = x.copy()        # fwd: x_{j0-1}; bwd: x_{j1-1} (first processed)
    for i in range(Lc):
        jf = j0 + i; jb = j1 - 1 - i
        act = jf < j1
        j = np.where(act, np.where(fwd, jf, jb), 1)
        # forward step: x_j = x + w*ig[j-1]; w = w - t_j x_j ; accumulate at x_j
        xf = x + w * ig[j - 1]; wf = w - t[j] * xf
        # backward: current (x,w) is (x_j,w_j): accumulate at x_j, then step to j-1
        wb = w + t[j] * x; xb = x - wb * ig[j - 1]
        xa = np.where(fwd, xf, x)                       # the x_j this step accounts for
        acc = np.where(act, acc + F[j] * xa * xa, acc)
        if want_z:
            z[j[act]] = (xa * scl)[act]
        # nodes: fwd compares x_j with x_{j-1} (skip j==1); bwd compares x_j with x_{j+1} (skip first)
        chg_f = act & fwd & (j > 1) & ((xf < 0) != (prev < 0))
        chg_b = act & (~fwd) & (i > 0) & ((x < 0) != (prev < 0))
        nodes += chg_f.astype(int) + chg_b.astype(int)
        prev = np.where(act, xa, prev)
        x = np.where(act, np.where(fwd, xf, xb), x); w = np.where(act, np.where(fwd, wf, wb), w)
    # backward chunks: after loop x = x_{j0-1}; compare with x_{j0} (prev)
    chg_end = (~fwd) & valid & ((x < 0) != (prev < 0))
    nodes += chg_end.astype(int)
    S = np.sum(acc * scl * scl)
    count = int(nodes.sum()) + (1 if r > 0 else 0)
    return dict(rho=lam + r / S, r=r, kb=kb, S=S, count=count, z=z, nodes=int(nodes.sum()))
