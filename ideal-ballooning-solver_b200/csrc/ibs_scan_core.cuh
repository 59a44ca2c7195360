// Lane-per-solve scan solver: the per-lane arithmetic (host/device code).
//
// Scan-shaped batches (ball_scan.py:248-274: every field line is solved for a whole row of theta0) have a structure
// the team-per-solve kernel (ibs_solver.cu) cannot use: all theta0 of a line share the same six theta0-independent
// coefficient rows (see poly_prep).  Here ONE LANE owns one solve and a warp owns 32 (or 64) consecutive theta0 of
// one line, so that every coefficient load is warp-uniform (one broadcast LDS.128 serves 32 solves), there is no
// per-solve set-up, no transfer-matrix scan (a lane runs its whole chain: half the FP64 work per row of the team
// kernel) and no idle lane.  The recurrences are the same as in ibs_solver.cu,
//     x_j = x_{j-1} + w_{j-1} / gh_{j-1},    w_j = w_{j-1} - (C_j - lam F_j) x_j          (utils.py:1564-1592),
// run forward from the left Dirichlet end and backward from the right one to a matching row k (twisted
// factorisation); inside the iteration they are used in a DIVISION-FREE form (state scaled by the running product of
// the gh): 13 DFMA per row including the coefficient polynomials and the sum F z^2 of the Rayleigh-quotient step.
// Without a chain of predecessors to warm-start from, the start value comes from the same pencil on coarser grids
// (every 8th, 4th, 2nd point; Richardson-extrapolated), which costs ~4 fine-grid evaluations instead of ~12.
//
// This header holds everything a lane computes, templated on a context that supplies the coefficient records and the
// warp votes: the CUDA context (ibs_scan_solver.cu) streams the records through shared memory with TMA bulk copies;
// tools/scan_core_host.cpp instantiates the same code for one lane on the CPU (test harness only).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define IBS_HD __host__ __device__ __forceinline__
#else
#define IBS_HD inline
#endif

namespace ibs {
namespace scan {

constexpr int TR = 32;            // records per tile (one pipeline stage holds one forward and one backward tile)
constexpr int REC = 6;            // doubles per record: G0 G1 G2 (g = G0 + th0 G1 + th0^2 G2), C0 C1 (2 h^2 c), R (2 h^2 f = g R)
constexpr int MAXLEV = 3;         // coarse levels (strides 2, 4, 8)
constexpr int MIN_COARSE_N = 65;  // a level is used only if it has at least this many points
constexpr int MAXIT_LEVEL = 64;   // evaluations per level before giving up
constexpr int K_MARGIN = 8;       // matching row kept this far from both ends
constexpr double LOWQ = 1.0e3;    // max|z| / |z_k| above which the matching row is moved (see solve_item)

constexpr int FLAG_NOT_CONVERGED = 1, FLAG_BAD_INPUT = 2, FLAG_SIGMA_NOT_MAX = 4;   // = IBS_FLAG_* of include/ibs_b200.h

// ---- bit helpers ---------------------------------------------------------------------------------
IBS_HD int hi_word(double v) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(v);
#else
    uint64_t u; std::memcpy(&u, &v, 8); return (int)(u >> 32);
#endif
}
IBS_HD double pow2(int e) {       // 2^e; 0 below the normal range, 2^1023 above
    if (e < -1022) return 0.0;
    if (e > 1023) e = 1023;
#if defined(__CUDA_ARCH__)
    return __hiloint2double((e + 1023) << 20, 0);
#else
    uint64_t u = (uint64_t)(e + 1023) << 52; double d; std::memcpy(&d, &u, 8); return d;
#endif
}
IBS_HD int imin(int a, int b) { return a < b ? a : b; }
IBS_HD int imax(int a, int b) { return a > b ? a : b; }
IBS_HD unsigned sign_bit(double v) { return (unsigned)hi_word(v) >> 31; }
// exponent of the larger of |a|, |b| (0 for zero / subnormal / inf / nan), clamped to +-1000
IBS_HD int exp_max2(double a, double b) {
    const int ha = hi_word(a) & 0x7fffffff, hb = hi_word(b) & 0x7fffffff;
    const int hm = imax(ha, hb);
    const int e = (hm >> 20) - 1023;
    return (hm >= 0x00100000 && hm < 0x7ff00000) ? imax(-1000, imin(1000, e)) : 0;
}
IBS_HD bool not_pos_normal(double v) { return (unsigned)(hi_word(v) - 0x00100000) >= 0x7fe00000u; }
IBS_HD bool not_finite(double v) { return (unsigned)(hi_word(v) & 0x7fffffff) >= 0x7ff00000u; }

// reciprocal of a positive normal double, ~1 ulp (MUFU.RCP64H seed + Newton; see ibs_solver.cu fast_rcp)
IBS_HD double rcp_fast(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    e = fma(e, e, e);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
#else
    return 1.0 / x;
#endif
}

// ---- coefficient records ------------------------------------------------------------------------------
struct Rec { double G0, G1, G2, C0, C1, R; };

IBS_HD Rec load_rec(const double* p) {
    Rec r;
#if defined(__CUDA_ARCH__)
    const double2* q = reinterpret_cast<const double2*>(p);      // 48-byte records, 16-byte aligned: three LDS.128
    const double2 a = q[0], b = q[1], c = q[2];
    r.G0 = a.x; r.G1 = a.y; r.G2 = b.x; r.C0 = b.y; r.C1 = c.x; r.R = c.y;
#else
    r.G0 = p[0]; r.G1 = p[1]; r.G2 = p[2]; r.C0 = p[3]; r.C1 = p[4]; r.R = p[5];
#endif
    return r;
}
IBS_HD void coef(const Rec& r, double th0, double& g, double& C, double& F) {
    g = fma(th0, fma(th0, r.G2, r.G1), r.G0);
    C = fma(th0, r.C1, r.C0);
    F = g * r.R;
}

// number of points of level `lev` (stride 2^lev) and its first record within a line's block
IBS_HD int level_n(int N, int lev) { return ((N - 1) >> lev) + 1; }
IBS_HD int level_offset(int N, int lev) { int o = 0; for (int l = 0; l < lev; ++l) o += level_n(N, l); return o; }
IBS_HD int num_levels(int N) {            // coarse levels usable for N points
    int n = 0;
    while (n < MAXLEV && ((N - 1) % (2 << n)) == 0 && level_n(N, n + 1) >= MIN_COARSE_N) ++n;
    return n;
}
IBS_HD int clamp_k(int k, int Nl) { return imax(K_MARGIN, imin(Nl - 1 - K_MARGIN, k)); }

// ---- division-free chain steps (iteration) ------------------------------------------------------------
// state (X, W) = P (x, w') with P the running product of a = 2 gh;  S = P^2 sum F x^2
IBS_HD void fwd_step(const Rec& rc, double th0, double lam, double& X, double& W, double& S, double& gp) {
    double g, C, F;
    coef(rc, th0, g, C, F);
    const double t = fma(-lam, F, C);
    const double a = g + gp;
    gp = g;
    const double Xn = fma(a, X, W);
    W = fma(-t, Xn, a * W);
    S = fma(F * Xn, Xn, (a * a) * S);
    X = Xn;
}
template <bool ADDS>
IBS_HD void bwd_step(const Rec& rc, double th0, double lam, double& X, double& W, double& S, double& gp, double& tcur) {
    double g, C, F;
    coef(rc, th0, g, C, F);                    // the row being stepped TO
    const double a = g + gp;
    gp = g;
    const double tmp = fma(tcur, X, W);
    const double Xn = fma(a, X, -tmp);
    W = a * tmp;
    const double a2S = (a * a) * S;
    S = ADDS ? fma(F * Xn, Xn, a2S) : a2S;
    X = Xn;
    tcur = fma(-lam, F, C);
}
IBS_HD void rescale3(double& X, double& W, double& S) {
    const double s = pow2(-exp_max2(X, W));
    X *= s; W *= s; S *= s * s;
}
IBS_HD int sign_changes32(unsigned m, unsigned enter_sign) {     // bit (31 - i) of m = sign after step i
#if defined(__CUDA_ARCH__)
    return __popc((m ^ (m >> 1)) & 0x7fffffffu) + (int)(((m >> 31) & 1u) ^ enter_sign);
#else
    return __builtin_popcount((m ^ (m >> 1)) & 0x7fffffffu) + (int)(((m >> 31) & 1u) ^ enter_sign);
#endif
}

// One evaluation at the shifts lam[]: twisted residual r' (= 2 r), S' = sum 2F z^2 (z_k = 1), node count.
// rho = lam + r' / S' is the Rayleigh quotient of z; #eigenvalues above lam = nodes + (r' > 0).
template <int SPL, class Ctx>
IBS_HD void eval_pass(Ctx& ctx, int lev, int Nl, int k, const double (&th0)[SPL], const double (&lam)[SPL],
                      double (&r)[SPL], double (&S)[SPL], int (&nodes)[SPL]) {
    const int qf_end = k, qb_end = Nl - 1 - k;           // last record of each direction (row k)
    const int qmin = imin(qf_end, qb_end), qmax = imax(qf_end, qb_end);
    const int nst = qmax / TR + 1;
    ctx.begin_pass(lev, Nl, k, nst);
    double Xf[SPL], Wf[SPL], Sf[SPL], gf[SPL], Xb[SPL], Wb[SPL], Sb[SPL], gb[SPL], tb[SPL];
#pragma unroll
    for (int q = 0; q < SPL; ++q) nodes[q] = 0;
    for (int s = 0; s < nst; ++s) {
        ctx.wait(s);
        int i0 = 0;
        if (s == 0) {
            const Rec f0 = load_rec(ctx.frec(0)), f1 = load_rec(ctx.frec(1));
            const Rec b0 = load_rec(ctx.brec(0)), b1 = load_rec(ctx.brec(1));
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                double g, C, F;
                coef(f0, th0[q], g, C, F);
                gf[q] = g; Xf[q] = 0.0; Wf[q] = 1.0; Sf[q] = 0.0;
                fwd_step(f1, th0[q], lam[q], Xf[q], Wf[q], Sf[q], gf[q]);           // row 1: X_1 = 1 > 0
                coef(b0, th0[q], g, C, F);
                const double gN = g;
                coef(b1, th0[q], g, C, F);                                          // row M = Nl - 2
                gb[q] = g; Xb[q] = 1.0; Wb[q] = -(g + gN); Sb[q] = F; tb[q] = fma(-lam[q], F, C);
            }
            i0 = 2;
        }
        if (s > 0 && TR * s + TR - 1 < qmin) {
            // fast path: a whole tile of both directions, no per-step tests; node counts from sign histories
            unsigned mf[SPL], mb[SPL], ef[SPL], eb[SPL];
#pragma unroll
            for (int q = 0; q < SPL; ++q) { mf[q] = 0; mb[q] = 0; ef[q] = sign_bit(Xf[q]); eb[q] = sign_bit(Xb[q]); }
#pragma unroll 1
            for (int hh = 0; hh < TR; hh += 16) {
#pragma unroll 2
                for (int ii = 0; ii < 16; ++ii) {
                    const Rec rf = load_rec(ctx.frec(hh + ii)), rb = load_rec(ctx.brec(hh + ii));
#pragma unroll
                    for (int q = 0; q < SPL; ++q) {
                        fwd_step(rf, th0[q], lam[q], Xf[q], Wf[q], Sf[q], gf[q]);
                        mf[q] = (mf[q] << 1) | sign_bit(Xf[q]);
                        bwd_step<true>(rb, th0[q], lam[q], Xb[q], Wb[q], Sb[q], gb[q], tb[q]);
                        mb[q] = (mb[q] << 1) | sign_bit(Xb[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < SPL; ++q) { rescale3(Xf[q], Wf[q], Sf[q]); rescale3(Xb[q], Wb[q], Sb[q]); }
            }
#pragma unroll
            for (int q = 0; q < SPL; ++q) nodes[q] += sign_changes32(mf[q], ef[q]) + sign_changes32(mb[q], eb[q]);
        } else {
#pragma unroll 1
            for (int i = i0; i < TR; ++i) {
                const int qq = TR * s + i;
                if (qq <= qf_end) {
                    const Rec rf = load_rec(ctx.frec(i));
#pragma unroll
                    for (int q = 0; q < SPL; ++q) {
                        const unsigned s0 = sign_bit(Xf[q]);
                        fwd_step(rf, th0[q], lam[q], Xf[q], Wf[q], Sf[q], gf[q]);
                        nodes[q] += (int)(s0 ^ sign_bit(Xf[q]));
                    }
                }
                if (qq <= qb_end) {
                    const Rec rb = load_rec(ctx.brec(i));
#pragma unroll
                    for (int q = 0; q < SPL; ++q) {
                        const unsigned s0 = sign_bit(Xb[q]);
                        if (qq < qb_end) bwd_step<true>(rb, th0[q], lam[q], Xb[q], Wb[q], Sb[q], gb[q], tb[q]);
                        else bwd_step<false>(rb, th0[q], lam[q], Xb[q], Wb[q], Sb[q], gb[q], tb[q]);     // row k: counted by the forward sum
                        nodes[q] += (int)(s0 ^ sign_bit(Xb[q]));
                    }
                }
                if ((i & 15) == 15) {
#pragma unroll
                    for (int q = 0; q < SPL; ++q) { rescale3(Xf[q], Wf[q], Sf[q]); rescale3(Xb[q], Wb[q], Sb[q]); }
                }
            }
        }
        ctx.release(s);
    }
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        const double inv = 1.0 / (Xf[q] * Xb[q]);
        r[q] = fma(Wb[q], Xf[q], -(Wf[q] * Xb[q])) * inv;
        S[q] = fma(Sf[q] * Xb[q], Xb[q], Sb[q] * Xf[q] * Xf[q]) * inv * inv;
    }
}

// ---- output passes (plain recurrence with reciprocals: the un-scaled eigenfunction is needed) ------------
// One direction of one solve.  The window holds the four previous rows' values; the sums are split by row parity
// (Simpson weights 4/3 and 2/3) and carry the running power-of-two scale 2^-E of x.
struct Sweep {
    double x, w; int E;
    double W1, W2, W3, W4;        // x of the previous rows: W1 = the row computed last, ... (running scale; O2: normalised X)
    double gp, gpp;               // g of the last row and of the one before
    double tcur;                  // backward: t' of the current row
    double a0[2], a1[2], aD[2];   // sum t' X^2, sum F' X^2, sum g D^2 by row parity
    double aEnd;                  // g D^2 of the Dirichlet end point
    double vmax; int jmax;        // largest |x| so far (running scale) and its row
    bool bad;
};

struct SolveOut {                 // what the first output pass returns per solve
    double gam, zmax; int jmax; bool bad;
    double xkf, xkb; int Ekf, Ekb;
    double Dm1, D0, Dp1;          // seam stencils (z scale): rows k-1, k, k+1
};

constexpr double C23 = 2.0 / 3.0, C12 = 1.0 / 12.0;

template <bool WRITE>
IBS_HD void sweep_rescale(Sweep& sw, double& fsc, double cnorm, int& ex) {
    const int e = exp_max2(sw.x, sw.w);
    const double s = pow2(-e);
    sw.x *= s; sw.w *= s; sw.E += e;
    if (!WRITE) {
        const double s2 = s * s;
        sw.W1 *= s; sw.W2 *= s; sw.W3 *= s; sw.W4 *= s;
        sw.a0[0] *= s2; sw.a0[1] *= s2; sw.a1[0] *= s2; sw.a1[1] *= s2; sw.aD[0] *= s2; sw.aD[1] *= s2; sw.aEnd *= s2;
        sw.vmax *= s;
    } else {
        ex += e;
        fsc = cnorm * pow2(ex);
    }
}

// normalised output value: the largest element must come out as exactly 1 (utils.py:1605 divides by the maximum)
IBS_HD double norm_clamp(double v) { return (fabs(v) > 1.0 - 1e-15) ? copysign(1.0, v) : v; }

// One step of an output pass for one direction.  DIR = +1 forward (new row = qq), -1 backward (new row = Nl-1-qq).
// `last` (backward only): the step that reaches row k (its X^2 terms and its X belong to the forward sweep).
template <int DIR, bool WRITE>
IBS_HD void out_step(Sweep& sw, const Rec& rc, double th0, double lam, int qq, int Nl, bool last, double fsc, double ih,
                     double* Xrow, double* dXrow) {
    double g, C, F;
    coef(rc, th0, g, C, F);
    const double a = g + sw.gp;
    const double tnew = fma(-lam, F, C);
    const int row = (DIR > 0) ? qq : Nl - 1 - qq;
    if (!WRITE) sw.bad |= not_pos_normal(a) | ((row >= 1 && row <= Nl - 2) && (not_pos_normal(F) | not_finite(C)));
    const double ia = rcp_fast(a);
    double xn;
    if (DIR > 0) { xn = fma(sw.w, ia, sw.x); sw.w = fma(-tnew, xn, sw.w); }
    else         { sw.w = fma(sw.tcur, sw.x, sw.w); xn = fma(-sw.w, ia, sw.x); sw.tcur = tnew; }
    sw.x = xn;
    const int par = qq & 1;                    // = parity of the row (N odd)
    double v = xn;                             // value entering the window
    if (WRITE) {
        v = norm_clamp(xn * fsc);
        if (Xrow && !last) Xrow[row] = v;
    } else {
        if (!last) {
            const double x2 = xn * xn;
            sw.a0[par] = fma(tnew, x2, sw.a0[par]);
            sw.a1[par] = fma(F, x2, sw.a1[par]);
            const double ax = fabs(xn);
            if (ax > sw.vmax) { sw.vmax = ax; sw.jmax = row; }
        }
    }
    // stencil of the row two behind (h-free: D = h dX)
    if (qq >= 4) {
        const double D = fma(C23, sw.W1 - sw.W3, -(C12 * (v - sw.W4)));     // forward orientation; backward: -D
        if (WRITE) { if (dXrow) dXrow[(DIR > 0) ? row - 2 : row + 2] = ((DIR > 0) ? D : -D) * ih; }
        else sw.aD[par] = fma(sw.gpp, D * D, sw.aD[par]);
    } else if (qq == 2) {
        // Dirichlet end point: one-sided formula (utils.py:1610, 1613), weight 1/3
        const double D = fma(2.0, sw.W1, -0.5 * v);                     // forward: dX_0 h;  backward: -dX_{N-1} h
        if (WRITE) { if (dXrow) dXrow[(DIR > 0) ? 0 : Nl - 1] = ((DIR > 0) ? D : -D) * ih; }
        else sw.aEnd = sw.gpp * D * D;
    } else if (qq == 3) {
        // rows 1 and N-2: second-order formula (utils.py:1611-1612)
        const double D = 0.5 * sw.W1;                                   // forward: dX_1 h = X_2 / 2;  backward: -dX_{N-2} h
        if (WRITE) { if (dXrow) dXrow[(DIR > 0) ? 1 : Nl - 2] = ((DIR > 0) ? D : -D) * ih; }
        else sw.aD[par] = fma(sw.gpp, D * D, sw.aD[par]);
    }
    sw.W4 = sw.W3; sw.W3 = sw.W2; sw.W2 = sw.W1; sw.W1 = v;
    sw.gpp = sw.gp; sw.gp = g;
}

IBS_HD void sweep_zero(Sweep& sw) {
    sw.E = 0; sw.W1 = sw.W2 = sw.W3 = sw.W4 = 0.0; sw.tcur = 0.0;
    sw.a0[0] = sw.a0[1] = sw.a1[0] = sw.a1[1] = sw.aD[0] = sw.aD[1] = 0.0; sw.aEnd = 0.0;
    sw.vmax = 0.0; sw.jmax = 0; sw.bad = false;
}

// Output pass at the shifts lam[] with matching row k.
//   WRITE = false: Simpson Rayleigh quotient gam (utils.py:1605-1621), max|z| and its row, validity -> out[]
//   WRITE = true : X = z / max z (and dX) written for the solves with wr[q] set, using out[] of the first pass
template <int SPL, bool WRITE, class Ctx>
IBS_HD void out_pass(Ctx& ctx, int lev, int Nl, int k, const double (&th0)[SPL], const double (&lam)[SPL], double h,
                     SolveOut (&out)[SPL], const bool (&wr)[SPL], double* const (&Xrow_in)[SPL], double* const (&dXrow_in)[SPL]) {
    const int qf_end = k, qb_end = Nl - 1 - k;
    const int qmin = imin(qf_end, qb_end), qmax = imax(qf_end, qb_end);
    const int nst = qmax / TR + 1;
    const double ih = 1.0 / h;
    ctx.begin_pass(lev, Nl, k, nst);
    Sweep F_[SPL], B_[SPL];
    double fscf[SPL], fscb[SPL], cnf[SPL], cnb[SPL];
    int exf[SPL], exb[SPL];
    double* Xrow[SPL]; double* dXrow[SPL];
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        Xrow[q] = (WRITE && wr[q]) ? Xrow_in[q] : nullptr;
        dXrow[q] = (WRITE && wr[q]) ? dXrow_in[q] : nullptr;
        fscf[q] = fscb[q] = cnf[q] = cnb[q] = 0.0; exf[q] = exb[q] = 0;
        if (WRITE) {
            const bool ok = !out[q].bad && out[q].zmax > 0.0 && out[q].zmax < 1e300;
            cnf[q] = ok ? 1.0 / (out[q].xkf * out[q].zmax) : 0.0;
            cnb[q] = ok ? 1.0 / (out[q].xkb * out[q].zmax) : 0.0;
            exf[q] = -out[q].Ekf; exb[q] = -out[q].Ekb;
            fscf[q] = cnf[q] * pow2(exf[q]); fscb[q] = cnb[q] * pow2(exb[q]);
        }
    }
    for (int s = 0; s < nst; ++s) {
        ctx.wait(s);
        int i0 = 0;
        if (s == 0) {
            const Rec f0 = load_rec(ctx.frec(0)), f1 = load_rec(ctx.frec(1));
            const Rec b0 = load_rec(ctx.brec(0)), b1 = load_rec(ctx.brec(1));
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                double g, C, Fv;
                Sweep& f = F_[q]; Sweep& b = B_[q];
                sweep_zero(f); sweep_zero(b);
                // forward: row 0 (x = 0, w' = 1), then the ordinary step to row 1
                coef(f0, th0[q], g, C, Fv);
                f.x = 0.0; f.w = 1.0; f.gp = g; f.gpp = g;
                if (WRITE) { if (Xrow[q]) { Xrow[q][0] = 0.0; Xrow[q][Nl - 1] = 0.0; } }
                out_step<+1, WRITE>(f, f1, th0[q], lam[q], 1, Nl, false, fscf[q], ih, Xrow[q], dXrow[q]);
                // backward: row N-1 (x = 0), row M = N-2 with x = 1, w' = -a_M
                coef(b0, th0[q], g, C, Fv);
                const double gN = g;
                coef(b1, th0[q], g, C, Fv);
                const double a = g + gN;
                if (!WRITE) b.bad |= not_pos_normal(a) | not_pos_normal(Fv) | not_finite(C);
                b.x = 1.0; b.w = -a; b.gp = g; b.gpp = gN; b.tcur = fma(-lam[q], Fv, C);
                double v = 1.0;
                if (WRITE) { v = norm_clamp(fscb[q]); if (Xrow[q]) Xrow[q][Nl - 2] = v; }
                else { b.a0[1] = b.tcur; b.a1[1] = Fv; b.vmax = 1.0; b.jmax = Nl - 2; }
                b.W1 = v;
            }
            i0 = 2;
        }
        if (s > 0 && TR * s + TR - 1 < qmin) {
#pragma unroll 1
            for (int hh = 0; hh < TR; hh += 16) {
#pragma unroll 2
                for (int ii = 0; ii < 16; ++ii) {
                    const int qq = TR * s + hh + ii;
                    const Rec rf = load_rec(ctx.frec(hh + ii)), rb = load_rec(ctx.brec(hh + ii));
#pragma unroll
                    for (int q = 0; q < SPL; ++q) {
                        out_step<+1, WRITE>(F_[q], rf, th0[q], lam[q], qq, Nl, false, fscf[q], ih, Xrow[q], dXrow[q]);
                        out_step<-1, WRITE>(B_[q], rb, th0[q], lam[q], qq, Nl, false, fscb[q], ih, Xrow[q], dXrow[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < SPL; ++q) {
                    sweep_rescale<WRITE>(F_[q], fscf[q], cnf[q], exf[q]);
                    sweep_rescale<WRITE>(B_[q], fscb[q], cnb[q], exb[q]);
                }
            }
        } else {
#pragma unroll 1
            for (int i = i0; i < TR; ++i) {
                const int qq = TR * s + i;
                if (qq <= qf_end) {
                    const Rec rf = load_rec(ctx.frec(i));
#pragma unroll
                    for (int q = 0; q < SPL; ++q)
                        out_step<+1, WRITE>(F_[q], rf, th0[q], lam[q], qq, Nl, false, fscf[q], ih, Xrow[q], dXrow[q]);
                }
                if (qq <= qb_end) {
                    const Rec rb = load_rec(ctx.brec(i));
#pragma unroll
                    for (int q = 0; q < SPL; ++q)
                        out_step<-1, WRITE>(B_[q], rb, th0[q], lam[q], qq, Nl, qq == qb_end, fscb[q], ih, Xrow[q], dXrow[q]);
                }
                if ((i & 15) == 15) {
#pragma unroll
                    for (int q = 0; q < SPL; ++q) {
                        if (qq < qf_end) sweep_rescale<WRITE>(F_[q], fscf[q], cnf[q], exf[q]);      // a finished sweep keeps its final scale
                        if (qq < qb_end) sweep_rescale<WRITE>(B_[q], fscb[q], cnb[q], exb[q]);
                    }
                }
            }
        }
        ctx.release(s);
    }
    // ---- seam (rows k-1, k, k+1) and totals
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        const Sweep& f = F_[q]; const Sweep& b = B_[q];
        if (!WRITE) {
            SolveOut& o = out[q];
            o.xkf = f.x; o.xkb = b.x; o.Ekf = f.E; o.Ekb = b.E;
            o.bad = f.bad | b.bad;
            const double zf = 1.0 / f.x, zb = 1.0 / b.x;
            // window values in the z scale: forward W1..W4 = rows k..k-3, backward W1..W4 = rows k..k+3
            const double Zm3 = f.W4 * zf, Zm2 = f.W3 * zf, Zm1 = f.W2 * zf, Zp1 = b.W2 * zb, Zp2 = b.W3 * zb, Zp3 = b.W4 * zb;
            o.Dm1 = fma(C23, 1.0 - Zm2, -(C12 * (Zp1 - Zm3)));
            o.D0 = fma(C23, Zp1 - Zm1, -(C12 * (Zp2 - Zm2)));
            o.Dp1 = fma(C23, Zp2 - 1.0, -(C12 * (Zp3 - Zm1)));
            const double w43 = 4.0 / 3.0, w23 = 2.0 / 3.0, w13 = 1.0 / 3.0;
            const double wk = (k & 1) ? w43 : w23, wk1 = (k & 1) ? w23 : w43;       // rows k and k +- 1
            const double zf2 = zf * zf, zb2 = zb * zb;
            const double sD = zf2 * (w43 * f.aD[1] + w23 * f.aD[0] + w13 * f.aEnd) + zb2 * (w43 * b.aD[1] + w23 * b.aD[0] + w13 * b.aEnd) +
                              wk1 * (f.gpp * o.Dm1 * o.Dm1 + b.gpp * o.Dp1 * o.Dp1) + wk * (f.gp * o.D0 * o.D0);
            const double sX0 = zf2 * (w43 * f.a0[1] + w23 * f.a0[0]) + zb2 * (w43 * b.a0[1] + w23 * b.a0[0]);
            const double sX1 = zf2 * (w43 * f.a1[1] + w23 * f.a1[0]) + zb2 * (w43 * b.a1[1] + w23 * b.a1[0]);
            o.gam = lam[q] + (sX0 - 2.0 * sD) / sX1;
            const double mf = f.vmax * fabs(zf), mb = b.vmax * fabs(zb);
            o.zmax = fmax(mf, mb);
            o.jmax = (mb > mf) ? b.jmax : f.jmax;
            if (!(o.zmax == o.zmax)) o.zmax = 1e308;          // NaN: treat as unusable
        } else if (dXrow[q]) {
            // the seam stencils in the normalised scale (the windows hold normalised X here)
            const double Xm3 = f.W4, Xm2 = f.W3, Xm1 = f.W2, X0 = f.W1, Xp1 = b.W2, Xp2 = b.W3, Xp3 = b.W4;
            dXrow[q][k - 1] = fma(C23, X0 - Xm2, -(C12 * (Xp1 - Xm3))) * ih;
            dXrow[q][k] = fma(C23, Xp1 - Xm1, -(C12 * (Xp2 - Xm2))) * ih;
            dXrow[q][k + 1] = fma(C23, Xp2 - X0, -(C12 * (Xp3 - Xm1))) * ih;
        }
    }
}

// ---- bracketed Rayleigh-quotient iteration (per solve; the logic of ibs_solver.cu) -----------------------
struct Iter {
    double lam, rho, lo, hi, b1, N1, b2, N2, dprev, dprev2;
    int nabove, it;
    bool collapsed, done, conv, warm;
};

IBS_HD void iter_init(Iter& s, double l0, double Lb, double U, bool frozen) {
    s.lo = Lb; s.hi = U; s.b1 = s.N1 = s.b2 = s.N2 = 0.0; s.dprev = 1e300; s.dprev2 = 1e300; s.nabove = 0; s.it = 0;
    s.collapsed = false; s.done = frozen; s.conv = frozen;
    s.warm = (l0 > Lb && l0 < U);
    s.lam = s.warm ? l0 : U;
    s.rho = s.lam;
}

IBS_HD void iter_update(Iter& s, double r, double S, int nodes, double U, double tol, double tol_stag, double stop, bool predictive) {
    if (s.done) return;
    ++s.it;
    const double lam = s.lam;
    const double rho = lam + r / S;
    s.rho = rho;
    const bool pos = nodes == 0;
    const bool above = pos && !(r > 0.0);
    const bool inbasin = pos && (r > 0.0);
    if (above) {
        s.hi = fmin(s.hi, lam);
        s.b1 = s.b2; s.N1 = s.N2; s.b2 = lam; s.N2 = lam - rho; ++s.nabove;
        if (rho == rho) s.lo = fmax(s.lo, fmin(rho, s.hi));
    } else {
        s.lo = fmax(s.lo, lam);
        if (inbasin && rho == rho) s.lo = fmax(s.lo, fmin(rho, s.hi));
    }
    bool done = false;
    if (pos) {
        const double dl = fabs(rho - lam);
        if (dl <= stop || (dl < tol_stag && dl >= 0.25 * s.dprev)) { s.conv = true; done = true; }
        // quadratic convergence: the error of rho is ~ dl^3 / dprev^2 once two consecutive corrections contract
        if (predictive && !done && dl < 1e-3 * s.dprev && s.dprev < 1e-2 * fmax(fabs(U), 1e-3)) {
            const double q = dl / s.dprev;
            if (dl * q * q <= 0.01 * tol) { s.conv = true; done = true; }
        }
        s.dprev2 = s.dprev; s.dprev = dl;
    }
    if (!done && s.collapsed) { s.conv = true; done = true; }
    if (!done) {
        double nxt;
        if (s.hi - s.lo <= tol) {
            nxt = 0.5 * (s.lo + s.hi);
            s.collapsed = true;
        } else if (inbasin && rho > lam && rho <= s.hi) {
            nxt = rho;
        } else if (above) {
            double pw = 0.5;
            if (s.nabove >= 2 && s.N1 - s.N2 > 0.0) pw = (s.b1 - s.b2) / (s.N1 - s.N2);
            pw = fmin(1.0, fmax(0.4, pw));
            if (pw > 0.8) pw = 1.0;
            if (s.warm && s.nabove == 1) pw = 1.0;
            if (s.N2 < 0.05 * (U - s.b2)) pw = 1.0;
            nxt = s.b2 - pw * s.N2;
            if (!(nxt >= s.lo && nxt < s.hi)) nxt = 0.5 * (s.lo + s.hi);
        } else {
            nxt = 0.5 * (s.lo + s.hi);
        }
        if (nxt == lam || s.it >= MAXIT_LEVEL) { s.conv = (nxt == lam); done = true; }
        else s.lam = nxt;
    }
    s.done = done;
}

struct ItemProblem {
    int N, nlev;
    double h, U, Lb;
    bool want_X, want_dX;
};
struct ItemResult { double gam, rho; int info; };

// Everything for the SPL solves of one lane.  act[q] = false: the slot duplicates a valid solve and writes nothing.
template <int SPL, class Ctx>
IBS_HD void solve_item(Ctx& ctx, const ItemProblem& P, const double (&th0)[SPL], const bool (&act)[SPL], const double (&sigma)[SPL],
                       const bool has_sigma, double* const (&Xrow)[SPL], double* const (&dXrow)[SPL], ItemResult (&res)[SPL]) {
    const double scale = fmax(fabs(P.U), 1e-3);
    const double tol = 1.7763568394002505e-15 * scale, tol_stag = 1e-10 * scale;
    Iter it[SPL];
    double lam_eval[SPL], r[SPL], S[SPL], rho1[SPL], rho2[SPL];
    int nodes[SPL], nev[SPL], flags[SPL];
    const double qnan = NAN;
#pragma unroll
    for (int q = 0; q < SPL; ++q) { rho1[q] = qnan; rho2[q] = qnan; nev[q] = 0; flags[q] = 0; }
    int k = 0;
    SolveOut out[SPL];
    bool nowr[SPL];
#pragma unroll
    for (int q = 0; q < SPL; ++q) nowr[q] = false;
    double* nullrow[SPL];
#pragma unroll
    for (int q = 0; q < SPL; ++q) nullrow[q] = nullptr;

    for (int lev = P.nlev; lev >= 0; --lev) {
        const int Nl = level_n(P.N, lev);
        k = (lev == P.nlev) ? clamp_k((Nl - 1) / 2, Nl) : clamp_k(2 * k, Nl);
        const double stop = (lev > 0) ? 1e-7 * scale : tol;
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            double l0 = rho1[q];
            if (rho2[q] == rho2[q]) l0 = rho1[q] - 0.25 * (rho2[q] - rho1[q]);      // Richardson: the error is ~ h^2
            iter_init(it[q], l0, P.Lb, P.U, false);
        }
        for (;;) {
#pragma unroll
            for (int q = 0; q < SPL; ++q) lam_eval[q] = it[q].lam;
            eval_pass<SPL>(ctx, lev, Nl, k, th0, lam_eval, r, S, nodes);
            bool alldone = true;
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                if (!it[q].done) ++nev[q];
                iter_update(it[q], r[q], S[q], nodes[q], P.U, tol, tol_stag, stop, lev == 0);
                alldone &= it[q].done;
            }
            if (ctx.all(alldone)) break;
        }
#pragma unroll
        for (int q = 0; q < SPL; ++q) { rho2[q] = rho1[q]; rho1[q] = (it[q].conv && it[q].rho == it[q].rho) ? it[q].rho : qnan; }
        if (lev == P.nlev && lev > 0) {
            // matching row from the coarsest eigenfunctions: the middle of the range of the lanes' peaks
            double sh[SPL];
#pragma unroll
            for (int q = 0; q < SPL; ++q) sh[q] = (rho1[q] == rho1[q]) ? rho1[q] : it[q].lam;
            out_pass<SPL, false>(ctx, lev, Nl, k, th0, sh, P.h, out, nowr, nullrow, nullrow);
            int jlo = 1 << 30, jhi = -1;
#pragma unroll
            for (int q = 0; q < SPL; ++q) { jlo = imin(jlo, out[q].jmax); jhi = imax(jhi, out[q].jmax); }
            jlo = ctx.min_i(jlo); jhi = ctx.max_i(jhi);
            k = clamp_k((jlo + jhi) / 2, Nl);
        }
    }
    // ---- fine level done: output passes; the matching row is moved if some solve has |z_k| << max|z|
    const int N = P.N;
    bool fin[SPL];
    double shift[SPL];
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        fin[q] = false;
        shift[q] = (it[q].conv && it[q].rho == it[q].rho) ? it[q].rho : it[q].lam;
        if (!it[q].conv) flags[q] |= FLAG_NOT_CONVERGED;
        res[q].gam = qnan; res[q].rho = qnan;
    }
    for (int round = 0; round < 4; ++round) {
        out_pass<SPL, false>(ctx, 0, N, k, th0, shift, P.h, out, nowr, nullrow, nullrow);
        bool newly[SPL], lowq_any = false, wr_any = false;
        int jsel = -1;
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            const bool lowq = !fin[q] && !out[q].bad && out[q].zmax > LOWQ && round < 3;
            newly[q] = !fin[q] && !lowq;
            if (newly[q]) {
                fin[q] = true;
                res[q].gam = out[q].bad ? qnan : out[q].gam;
                res[q].rho = out[q].bad ? qnan : shift[q];
                if (out[q].bad) flags[q] = FLAG_BAD_INPUT;
            }
            if (lowq && jsel < 0) jsel = out[q].jmax;
            lowq_any |= lowq;
            wr_any |= newly[q] && act[q];
        }
        if ((P.want_X || P.want_dX) && ctx.any(wr_any)) {
            bool wr[SPL];
#pragma unroll
            for (int q = 0; q < SPL; ++q) wr[q] = newly[q] && act[q];
            out_pass<SPL, true>(ctx, 0, N, k, th0, shift, P.h, out, wr, Xrow, dXrow);
        }
        if (!ctx.any(lowq_any)) break;
        // move the matching row to the peak of the first low-quality solve and re-converge those solves there
        k = clamp_k(ctx.first_i(jsel), N);
#pragma unroll
        for (int q = 0; q < SPL; ++q) iter_init(it[q], shift[q], P.Lb, P.U, fin[q]);
        for (;;) {
#pragma unroll
            for (int q = 0; q < SPL; ++q) lam_eval[q] = it[q].lam;
            eval_pass<SPL>(ctx, 0, N, k, th0, lam_eval, r, S, nodes);
            bool alldone = true;
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                if (!it[q].done) ++nev[q];
                iter_update(it[q], r[q], S[q], nodes[q], P.U, tol, tol_stag, tol, true);
                alldone &= it[q].done;
            }
            if (ctx.all(alldone)) break;
        }
#pragma unroll
        for (int q = 0; q < SPL; ++q)
            if (!fin[q]) {
                shift[q] = (it[q].conv && it[q].rho == it[q].rho) ? it[q].rho : it[q].lam;
                if (!it[q].conv) flags[q] |= FLAG_NOT_CONVERGED;
            }
    }
    // ---- nearest-sigma check (utils.py:1597 returns the eigenvalue nearest sigma; the engine returns lambda_max)
    if (has_sigma) {
        bool need[SPL], any_need = false;
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            need[q] = (flags[q] & FLAG_BAD_INPUT) == 0 && sigma[q] < shift[q];
            lam_eval[q] = need[q] ? 2.0 * sigma[q] - shift[q] : shift[q];
            any_need |= need[q];
        }
        if (ctx.any(any_need)) {
            eval_pass<SPL>(ctx, 0, N, k, th0, lam_eval, r, S, nodes);
#pragma unroll
            for (int q = 0; q < SPL; ++q)
                if (need[q] && nodes[q] + (r[q] > 0.0 ? 1 : 0) > 1) flags[q] |= FLAG_SIGMA_NOT_MAX;
        }
    }
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        const int itc = (flags[q] & FLAG_BAD_INPUT) ? 0 : ((flags[q] & FLAG_NOT_CONVERGED) ? 64 : imin(nev[q], 0xffff));
        res[q].info = itc | (flags[q] << 16);
    }
}

// ---- coefficient preparation of one point (poly_prep): records of every level that contains the point ------
// base8: the eight base arrays of the line at this point in IBS_BASE_* order (bmag, gradpar, cvdrift, cvdrift0, gds2,
// gds21, gds22, gbdrift).  Returns the un-scaled record of level 0 (sigma and the level factors are applied by the caller).
IBS_HD Rec raw_record(double B, double gradpar, double cv, double cv0, double gds2, double gds21, double gds22, double dP, double h2) {
    const double gp = fabs(gradpar);
    const double gpB = gp * B, gpoB = gp / B, mdP = -2.0 * h2 * dP / gpB;
    Rec r;
    r.G0 = gpoB * gds2; r.G1 = 2.0 * gpoB * gds21; r.G2 = gpoB * gds22;
    r.C0 = mdP * cv; r.C1 = mdP * cv0;
    r.R = 2.0 * h2 / (gpB * gpB);
    return r;
}
// range of g, C over theta0 in [t0, t1] for one record (g is a convex parabola where the input is valid)
IBS_HD void record_ranges(const Rec& r, double t0, double t1, double& gmin, double& gmax, double& Cmin, double& Cmax) {
    const double ga = fma(t0, fma(t0, r.G2, r.G1), r.G0), gb = fma(t1, fma(t1, r.G2, r.G1), r.G0);
    gmin = fmin(ga, gb); gmax = fmax(ga, gb);
    if (r.G2 > 0.0) {
        const double tv = -0.5 * r.G1 / r.G2;
        if (tv > t0 && tv < t1) gmin = fmin(gmin, fma(tv, fma(tv, r.G2, r.G1), r.G0));
    }
    const double Ca = fma(t0, r.C1, r.C0), Cb = fma(t1, r.C1, r.C0);
    Cmin = fmin(Ca, Cb); Cmax = fmax(Ca, Cb);
}

}  // namespace scan
}  // namespace ibs
