// Lane-per-CHAIN scan solver: the per-lane arithmetic (host/device code).
//
// Second form of the lane-per-solve kernel of ibs_scan_core.cuh (same pencil, same recurrences, same multigrid start,
// same state machine; see there).  In that kernel one lane runs BOTH chains of a solve's twisted factorisation (forward
// from the left Dirichlet end, backward from the right one), which costs ~240 registers -- 8 warps per SM, two per
// scheduler -- and although its streaming loops keep the FP64 pipe ~87 % busy, a warp in any other phase (pass
// prologue, iteration bookkeeping, the Simpson pass's integer work) leaves its scheduler to ONE other warp: ncu shows
// the pipe 48 % active over the whole kernel.  Here a lane runs ONE chain:
//     lane = 16 h + i,   h = 0: forward chain of solve i,   h = 1: backward chain of solve i   (i = 0..15),
// a warp owns 16 consecutive theta0 of one line.  Half the state per lane -> twice the resident warps, and both halves of
// the warp execute the SAME instruction stream: the backward chain is the forward recurrence on the mirrored line
// (row N-1-q instead of q), so the only per-lane difference is the direction in which the records are read.
//   * Both chains are stepped to the matching row k and joined there (ibs_scan_core.cuh stops the backward chain one
//     update short of k; here row k is counted by both and taken out again in the join, which needs t_k and F_k: one
//     record read per pass).
//   * Every step is a pipelined step.  The chains end at multiples of 16 steps (the matching row is snapped to a multiple
//     of 16: any row near the eigenfunction's peak is as good), the pipelined chunks of 16 steps end exactly there, and a
//     half-warp whose chain is the shorter one simply sits out the remaining chunks -- there is no one-step-at-a-time
//     path for unequal chains (7 x the cost of a pipelined step in the two-chain kernel).
//   * Tiles of the record ring are shifted by three records (stage 0: steps 0..34, stage s: 32 s + 3 .. 32 s + 34) so
//     that a chunk that ENDS at a multiple of 16 never prefetches (two steps ahead) across a stage boundary in its middle.
// The pass functions below are per-lane and contain no cross-lane operation; the join of the two chains is a separate
// function of (mine, partner's).  The CUDA context gets the partner's values with shuffles (lane ^ 16); the CPU harness
// (tools/scan2_core_host.cpp, test infrastructure only) runs the two lanes one after the other.
#pragma once
#include "ibs_scan_core.cuh"

namespace ibs {
namespace scan2 {
using namespace scan;

constexpr int NH = 16;            // solves per warp
constexpr int T0 = TR + 3;        // records of stage 0: steps 0 .. 34
constexpr int NPRE = 5;           // steps 0 .. 4 are done one at a time (start values, the end-point stencils); chunks start at step 5
constexpr int KQ = 16;            // chain ends (and chunk ends) are multiples of KQ
#ifndef IBS_SCAN2_RCP_EXTRA
#define IBS_SCAN2_RCP_EXTRA 0       // one more Newton step on the reciprocal of the output passes: not needed (see o_single); 2.93 -> 2.84 ms without
#endif
#ifndef IBS_SCAN2_EVAL4
#define IBS_SCAN2_EVAL4 0          // iteration pass in blocks of four steps: measured 2.97 vs 2.92 ms (two-step blocks stay)
#endif
#ifndef IBS_SCAN2_EXTRAP
#define IBS_SCAN2_EXTRAP 1
#endif
constexpr int WREC = MAXLEV + 2;  // warm-start record: the eigenvalue of every level (index = level) + how far the neighbour was
constexpr double WARM_FAR = 0.05; // relative distance of the coarsest-level eigenvalues beyond which a neighbour is no help

IBS_HD int stage_of(int q) { return q < T0 ? 0 : (q - 3) / TR; }
IBS_HD int stage_first(int s) { return s == 0 ? 0 : TR * s + 3; }
IBS_HD int stage_len(int s, int Nl) { const int f = stage_first(s), l = (s == 0) ? T0 : TR; return imin(l, Nl - f); }
// matching row: nearest multiple of 16 with both chains at least 16 steps long (Nl - 1 is a multiple of 16)
IBS_HD int snap_k(int k, int Nl) { const int r = ((k + KQ / 2) / KQ) * KQ; return imax(KQ, imin(Nl - 1 - KQ, r)); }
IBS_HD bool scan2_size_ok(int N) {        // every level's N - 1 must be a multiple of 16 and leave room for two chains
    return (N & 1) == 1 && N >= 65 && ((N - 1) % KQ) == 0 && ((level_n(N, num_levels(N)) - 1) % KQ) == 0;
}

// ---- iteration pass: one chain of one solve --------------------------------------------------------------------------
struct EvalEnd { double X, W, S; int nodes; };

// chains of the current step (coefficients c) + coefficients of the next one from the record n; one chain per lane
IBS_HD void chain_step(const Rec& n, double th0, double lam, Co& c, double& gp, double& X, double& W, double& S) {
    const double p1 = fma(th0, n.G2, n.G1);
    const double Xn = fma(c.a, X, W);
    const double Cn = fma(th0, n.C1, n.C0);
    const double aW = c.a * W;
    const double a2 = c.a * c.a;
    const double gn = fma(th0, p1, n.G0);
    W = fma(-c.t, Xn, aW);
    const double FX = c.F * Xn;
    const double a2S = a2 * S;
    const double Fn = gn * n.R;
    S = fma(FX, Xn, a2S);
    X = Xn;
    c.a = gn + gp; gp = gn; c.F = Fn;
    c.t = fma(-lam, Fn, Cn);
}

template <class Ctx>
IBS_HD EvalEnd eval_lane(Ctx& ctx, int lev, int Nl, int q_end, int q_max, double th0, double lam) {
    ctx.begin_pass(lev, Nl, q_max);
    ctx.wait(0);
    const int ds = ctx.dstep();
    const double* pr = ctx.ptr(0, 0);
    double X = 0.0, W = 1.0, S = 0.0, gp;
    int nodes = 0;
    {
        double g, C, F;
        coef(load_rec(pr), th0, g, C, F);                // the Dirichlet end: x = 0, w' = 1
        gp = g;
        fwd_step(load_rec(pr + ds), th0, lam, X, W, S, gp);        // step 1: X = 1 > 0
        for (int q = 2; q < NPRE; ++q) {
            const unsigned s0 = sign_bit(X);
            fwd_step(load_rec(pr + q * ds), th0, lam, X, W, S, gp);
            nodes += (int)(s0 ^ sign_bit(X));
        }
    }
    Co cf;
    coef_next(load_rec(pr + NPRE * ds), th0, lam, gp, cf);
    Rec rf = load_rec(pr + (NPRE + 1) * ds);
    pr += (NPRE + 2) * ds;
    const int nchunk = q_max / KQ, cmine = q_end / KQ;      // chunk c ends at step 16 c + 16; my chain ends after chunk cmine - 1
#pragma unroll 1
    for (int c = 0; c < nchunk; ++c) {
        const bool newstage = c >= 2 && !(c & 1);
        if (newstage) { ctx.wait(c >> 1); pr = ctx.ptr(c >> 1, KQ * c + 3); }
        if (c < cmine) {                                      // (a half-warp whose chain has ended sits the chunk out)
            unsigned hist = 0;
            const unsigned enter = sign_bit(X);
            const int nblk = (c == 0) ? (KQ - NPRE + 1) / 2 : KQ / 2;        // 6 or 8 pairs of steps
#if IBS_SCAN2_EVAL4
#pragma unroll 1
            for (int b = 0; b < nblk / 2; ++b) {              // blocks of four steps (records two steps ahead, as before)
                const Rec n1 = load_rec(pr);
                chain_step(rf, th0, lam, cf, gp, X, W, S);
                hist = (hist << 1) | sign_bit(X);
                const Rec n2 = load_rec(pr + ds);
                chain_step(n1, th0, lam, cf, gp, X, W, S);
                hist = (hist << 1) | sign_bit(X);
                const Rec n3 = load_rec(pr + 2 * ds);
                chain_step(n2, th0, lam, cf, gp, X, W, S);
                hist = (hist << 1) | sign_bit(X);
                rf = load_rec(pr + 3 * ds);
                chain_step(n3, th0, lam, cf, gp, X, W, S);
                hist = (hist << 1) | sign_bit(X);
                pr += 4 * ds;
            }
#else
#pragma unroll 1
            for (int b = 0; b < nblk; ++b) {
                const Rec nf = load_rec(pr);
                chain_step(rf, th0, lam, cf, gp, X, W, S);
                hist = (hist << 1) | sign_bit(X);
                rf = load_rec(pr + ds);
                chain_step(nf, th0, lam, cf, gp, X, W, S);
                hist = (hist << 1) | sign_bit(X);
                pr += 2 * ds;
            }
#endif
            rescale3(X, W, S);
            nodes += sign_changes_n(hist, 2 * nblk, enter);
        }
        if (newstage) ctx.release((c >> 1) - 1);              // its last records were consumed by this chunk's first block
    }
    EvalEnd e; e.X = X; e.W = W; e.S = S; e.nodes = nodes;
    return e;
}

// join of the two chains at row k: twisted residual r' and S' = sum 2F z^2 (z_k = 1), node count
// (t_k, F_k) of the matching row from its record
IBS_HD void row_tF(const Rec& rk, double th0, double lam, double& tk, double& Fk) {
    double g, C;
    coef(rk, th0, g, C, Fk);
    tk = fma(-lam, Fk, C);
}
IBS_HD void eval_join(const EvalEnd& f, const EvalEnd& b, double tk, double F, double th0, double lam, double& r, double& S, int& nodes) {
    const double ixf = 1.0 / f.X, ixb = 1.0 / b.X;
    r = -(f.W * ixf + b.W * ixb + tk);
    S = fma(f.S * ixf, ixf, b.S * ixb * ixb) - F;            // row k was counted by both chains
    nodes = f.nodes + b.nodes;
}

// ---- output passes: one sweep of one solve ---------------------------------------------------------------------------
// general step (first rows: their stencils differ), forward arithmetic on the lane's own (possibly mirrored) row order
// CHECK (first pass only): test every row's coefficients for validity.  The prep kernel certifies whole lines (all rows, the
// whole theta0 range); their passes run without the tests (15 of 104 instructions per two steps).
template <bool WRITE, bool CHECK>
IBS_HD void o_step(Sweep& sw, const Rec& rc, double th0, double lam, int qq, int row, double* Xw) {
    double g, C, F;
    coef(rc, th0, g, C, F);
    const double a = g + sw.gp;
    const double tnew = fma(-lam, F, C);
    if (!WRITE && CHECK) sw.bad |= not_pos_normal(a) | not_pos_normal(F) | not_finite(C);
    const double ia = rcp_fast(a);
    const double xn = fma(sw.w, ia, sw.x);
    sw.w = fma(-tnew, xn, sw.w);
    sw.x = xn;
    if (WRITE) {
        if (Xw) Xw[row] = norm_value(xn, sw.fsc);
        sw.gp = g;
        return;
    }
    const int par = qq & 1;
    const double x2 = xn * xn;
    if (par) { sw.a0o = fma(tnew, x2, sw.a0o); sw.a1o = fma(F, x2, sw.a1o); }
    else     { sw.a0e = fma(tnew, x2, sw.a0e); sw.a1e = fma(F, x2, sw.a1e); }
    if (fabs(xn) > fabs(sw.vmax)) { sw.vmax = xn; sw.jmax = row; }      // (vmax keeps the SIGNED value: no instruction for |x|)
    if (qq >= 4) {
        const double D = fma(C23, sw.W1 - sw.W3, -(C12 * (xn - sw.W4)));
        if (par) sw.aDo = fma(sw.gpp, D * D, sw.aDo); else sw.aDe = fma(sw.gpp, D * D, sw.aDe);
    } else if (qq == 2) {
        const double D = fma(2.0, sw.W1, -0.5 * xn);        // Dirichlet end point: one-sided formula (utils.py:1610, 1613), weight 1/3
        sw.aEnd = sw.gpp * D * D;
    } else if (qq == 3) {
        const double D = 0.5 * sw.W1;                       // rows 1 and N-2: second-order formula (utils.py:1611-1612)
        sw.aDo = fma(sw.gpp, D * D, sw.aDo);
    }
    sw.W4 = sw.W3; sw.W3 = sw.W2; sw.W2 = sw.W1; sw.W1 = xn;
    sw.gpp = sw.gp; sw.gp = g;
}

// pipelined step: chain + sums of the current row (coefficients c) and, interleaved, the coefficients of the next row
template <int PAR, bool WRITE, bool CHECK>
IBS_HD void o_single(const Rec& n, double th0, double lam, OCo& c, Sweep& s, int row, double* Xw) {
    const double pA = fma(th0, n.G2, n.G1);
    const double CA = fma(th0, n.C1, n.C0);
    const double gA = fma(th0, pA, n.G0);
    const double aA = gA + c.g;
#if defined(__CUDA_ARCH__)
    double rA;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rA) : "d"(aA));
#else
    double rA = 1.0 / aA;
#endif
    const double FA = gA * n.R;
    if (!WRITE && CHECK) s.bad |= not_pos_normal(aA) | not_pos_normal(FA) | not_finite(CA);
    // current row
    const double xn = fma(s.w, c.ia, s.x);
    s.w = fma(-c.t, xn, s.w);
    s.x = xn;
    double eA = fma(-aA, rA, 1.0);
    const double tA = fma(-lam, FA, CA);
    if (WRITE) {
        if (Xw) Xw[row] = norm_value(xn, s.fsc);
    } else {
        const double d1 = s.W1 - s.W3;
        const double x2 = xn * xn;
        const double d2 = xn - s.W4;
        if (PAR) { s.a0o = fma(c.t, x2, s.a0o); s.a1o = fma(c.F, x2, s.a1o); }
        else     { s.a0e = fma(c.t, x2, s.a0e); s.a1e = fma(c.F, x2, s.a1e); }
        const double D = fma(C23, d1, -(C12 * d2));
        if (fabs(xn) > fabs(s.vmax)) { s.vmax = xn; s.jmax = row; }
        const double D2 = D * D;
        if (PAR) s.aDo = fma(s.gpp, D2, s.aDo); else s.aDe = fma(s.gpp, D2, s.aDe);
        s.W4 = s.W3; s.W3 = s.W2; s.W2 = s.W1; s.W1 = xn;
        s.gpp = s.gp;
    }
    s.gp = c.g;
    eA = fma(eA, eA, eA);
    rA = fma(rA, eA, rA);                                 // seed 2^-23 -> (2^-23)^3: below the rounding of a double
#if defined(__CUDA_ARCH__) && IBS_SCAN2_RCP_EXTRA
    eA = fma(-aA, rA, 1.0);
    rA = fma(rA, eA, rA);
#endif
    c.ia = rA; c.t = tA; c.F = FA; c.g = gA;
}

// what the seam needs from one sweep (first pass) -- the Sweep itself is returned
template <bool WRITE, bool CHECK, class Ctx>
IBS_HD Sweep out_lane(Ctx& ctx, int lev, int Nl, int q_end, int q_max, bool mirrored, double th0, double lam, double cn, int ex0, double* Xw) {
    ctx.begin_pass(lev, Nl, q_max);
    ctx.wait(0);
    const int ds = ctx.dstep();
    const double* pr = ctx.ptr(0, 0);
    Sweep sw;
    sweep_zero(sw);
    if (WRITE) {
        sw.cn = cn; sw.ex = ex0; sw.fsc = cn * pow2(ex0);
        if (Xw) Xw[mirrored ? Nl - 1 : 0] = 0.0;
    }
    {
        double g, C, F;
        coef(load_rec(pr), th0, g, C, F);
        sw.x = 0.0; sw.w = 1.0; sw.gp = g; sw.gpp = g;
        for (int q = 1; q < NPRE; ++q)
            o_step<WRITE, CHECK>(sw, load_rec(pr + q * ds), th0, lam, q, mirrored ? Nl - 1 - q : q, Xw);
    }
    OCo cf;
    {
        bool bad = false;
        out_coef<WRITE>(load_rec(pr + NPRE * ds), th0, lam, sw.gp, cf, bad);
        if (CHECK) sw.bad |= bad;
    }
    Rec rf = load_rec(pr + (NPRE + 1) * ds);
    pr += (NPRE + 2) * ds;
    int q = NPRE;                                             // step of the next chain update
    const int dr = mirrored ? -1 : 1;
    int row = mirrored ? Nl - 1 - NPRE : NPRE;
    const int nchunk = q_max / KQ, cmine = q_end / KQ;
#pragma unroll 1
    for (int c = 0; c < nchunk; ++c) {
        const bool newstage = c >= 2 && !(c & 1);
        if (newstage) { ctx.wait(c >> 1); pr = ctx.ptr(c >> 1, KQ * c + 3); }
        if (c < cmine) {
            const int nblk = (c == 0) ? (KQ - NPRE + 1) / 2 : KQ / 2;        // 6 or 8 blocks of two steps
            if (WRITE) {
#pragma unroll 1
                for (int b = 0; b < nblk; ++b) {              // steps q (odd), q + 1 (even)
                    const Rec nf = load_rec(pr);
                    o_single<1, WRITE, CHECK>(rf, th0, lam, cf, sw, row, Xw);
                    rf = load_rec(pr + ds);
                    o_single<0, WRITE, CHECK>(nf, th0, lam, cf, sw, row + dr, Xw);
                    pr += 2 * ds; row += 2 * dr;
                }
            } else {
                // first pass: blocks of FOUR steps -- the five-row window W1..W4 of the derivative stencil rotates once per
                // block, so no register moves are left (21 of 104 instructions per two steps in the two-step form)
#pragma unroll 1
                for (int b = 0; b < nblk / 2; ++b) {
                    const Rec n1 = load_rec(pr);
                    o_single<1, WRITE, CHECK>(rf, th0, lam, cf, sw, row, Xw);
                    const Rec n2 = load_rec(pr + ds);
                    o_single<0, WRITE, CHECK>(n1, th0, lam, cf, sw, row + dr, Xw);
                    const Rec n3 = load_rec(pr + 2 * ds);
                    o_single<1, WRITE, CHECK>(n2, th0, lam, cf, sw, row + 2 * dr, Xw);
                    rf = load_rec(pr + 3 * ds);
                    o_single<0, WRITE, CHECK>(n3, th0, lam, cf, sw, row + 3 * dr, Xw);
                    pr += 4 * ds; row += 4 * dr;
                }
            }
            q += 2 * nblk;
            sweep_rescale<WRITE>(sw);
        }
        if (newstage) ctx.release((c >> 1) - 1);
    }
    (void)q;
    return sw;
}

// seam (rows k-1, k, k+1) and totals of the first output pass: Simpson Rayleigh quotient, max |z| and its row, validity,
// and the two sweeps' value / scale at row k for the writing pass.  f: forward sweep, b: backward (mirrored) sweep; both
// have counted row k in their X^2 sums (taken out here with t_k, F_k of the record rk).
IBS_HD void out_join(const Sweep& f, const Sweep& b, double tk, double Fk, double th0, double lam, int k, SolveOut& o) {
    o.xkf = f.x; o.xkb = b.x; o.Ekf = f.E; o.Ekb = b.E;
    o.bad = f.bad | b.bad;
    const double zf = 1.0 / f.x, zb = 1.0 / b.x;
    const double Zm3 = f.W4 * zf, Zm2 = f.W3 * zf, Zm1 = f.W2 * zf, Zp1 = b.W2 * zb, Zp2 = b.W3 * zb, Zp3 = b.W4 * zb;
    const double Dm1 = fma(C23, 1.0 - Zm2, -(C12 * (Zp1 - Zm3)));
    const double D0 = fma(C23, Zp1 - Zm1, -(C12 * (Zp2 - Zm2)));
    const double Dp1 = fma(C23, Zp2 - 1.0, -(C12 * (Zp3 - Zm1)));
    const double w43 = 4.0 / 3.0, w23 = 2.0 / 3.0, w13 = 1.0 / 3.0;
    const double wk = (k & 1) ? w43 : w23, wk1 = (k & 1) ? w23 : w43;
    const double zf2 = zf * zf, zb2 = zb * zb;
    const double sD = zf2 * (w43 * f.aDo + w23 * f.aDe + w13 * f.aEnd) + zb2 * (w43 * b.aDo + w23 * b.aDe + w13 * b.aEnd) +
                      wk1 * (f.gpp * Dm1 * Dm1 + b.gpp * Dp1 * Dp1) + wk * (f.gp * D0 * D0);
    const double sX0 = zf2 * (w43 * f.a0o + w23 * f.a0e) + zb2 * (w43 * b.a0o + w23 * b.a0e) - wk * tk;
    const double sX1 = zf2 * (w43 * f.a1o + w23 * f.a1e) + zb2 * (w43 * b.a1o + w23 * b.a1e) - wk * Fk;
    o.gam = lam + (sX0 - 2.0 * sD) / sX1;
    // the same pass seen as an iteration pass (eval_join): twisted residual at row k and S = sum F z^2 with z_k = 1
    {
        const double r = -(f.w * zf + b.w * zb + tk);
        const double S = fma(zf2, f.a1o + f.a1e, zb2 * (b.a1o + b.a1e)) - Fk;
        o.dlt = r / S;
    }
    const double mf = fabs(f.vmax) * fabs(zf), mb = fabs(b.vmax) * fabs(zb);
    o.zmax = fmax(mf, mb);
    o.jmax = (mb > mf) ? b.jmax : f.jmax;
    if (!(o.zmax == o.zmax)) o.zmax = 1e308;
}

// ---- the state machine of one solve (both lanes of a pair run it identically) -----------------------------------------
// Ctx supplies:  eval(lev, Nl, k, th0, lam, r, S, nodes);  out1(lev, Nl, k, th0, lam, SolveOut&, check);  out2(lev, Nl, k, th0, lam,
// SolveOut, Xw);  all / any / min_i / max_i / first_i;  fixup(wr, X, dX, N, bad, h, want_dX);
// warm_in(token, lev) / warm_out(token, lev, value): the per-level eigenvalues of the NEIGHBOUR -- the same theta0 on the previous
// field line of the batch (the adjacent surface / alpha of a scan grid), solved one round earlier: token w_in, -1 = none -- and
// of this solve for its own successor (token w_out, -1 = nobody needs them).
//
// Warm start.  lam_max is smooth along the scan grid, and so is the error of every start value below: with the neighbour's level
// eigenvalues n_l known, the coarsest level starts at n_nlev instead of the Gershgorin bound (~4 passes instead of ~15),
// and every finer level starts at  E(own coarser levels) + [ n_l - E(neighbour's coarser levels) ]  -- the extrapolation E
// corrected by the error it made for the neighbour.  Whether the neighbour is close enough for that (adjacent surfaces of a
// tokamak scan: ~1 %; adjacent alphas of a stellarator scan: ~30 %, useless) is measured on the coarsest level, |lam - n| /
// |lam| <= WARM_FAR, and handed on with the record: a solve whose predecessor found its own neighbour far starts cold.
template <class Ctx>
IBS_HD void solve_item2(Ctx& ctx, const ItemProblem& P, double th0, bool act, double sigma, bool has_sigma, double* Xrow, double* dXrow,
                        ItemResult& res, ColdState<1>& cs, int w_in = -1, int w_out = -1) {
    const double scale = fmax(fabs(P.U), 1e-3);
    const double tol = 1.7763568394002505e-15 * scale, tol_stag = 1e-10 * scale;
    const double tol_lvl = 1e-9 * scale;                  // coarse levels: the error left in the level's eigenvalue
    const double qnan = NAN;
    const int N = P.N;
    Iter& it = cs.it[0];
    SolveOut& out = cs.out[0];
    double& rho1 = cs.rho1[0]; double& rho2 = cs.rho2[0]; double& rho3 = cs.rho3[0]; double& rbest = cs.rbest[0];
    double& Cq = cs.Cq[0];
    int& nev = cs.nev[0]; int& flags = cs.flags[0];
    bool& fin = cs.fin[0]; bool& wr = cs.wr[0]; bool& need = cs.need[0];
    double sh, r = 0.0, S = 1.0;
    int nodes = 0;
    const bool want_out = P.want_X || P.want_dX;
    rho1 = qnan; rho2 = qnan; rho3 = qnan; Cq = 0.0; nev = 0; flags = 0; fin = false; wr = false; need = false; rbest = qnan;
    cs.wi[0] = w_in; cs.wo[0] = w_out;
    {
        const double n4 = ctx.warm_in(w_in, P.nlev);
        iter_init(it, (ctx.warm_in(w_in, MAXLEV + 1) > WARM_FAR) ? qnan : n4, P.Lb, P.U, false);
    }
    sh = it.lam;
    res.gam = qnan; res.rho = qnan;
    int lev = P.nlev, Nl = level_n(N, lev), k = snap_k((Nl - 1) / 2, Nl), round = 0;
    int phase = PH_ITER;
    bool lowq_any = false, fix_any = false;
    bool prov = false;              // fine level: the first output pass doubles as the confirming iteration pass
    int jsel = -1;
    for (;;) {
        const int kind = (phase == PH_ITER || phase == PH_SIGMA) ? 1 : (phase == PH_PEAK || phase == PH_O1) ? 2 : 3;
        if (kind == 1) ctx.eval(lev, Nl, k, th0, sh, r, S, nodes);
        else if (kind == 2) ctx.out1(lev, Nl, k, th0, sh, out, !P.safe);
        else ctx.out2(lev, Nl, k, th0, sh, out, (wr && !out.bad) ? Xrow : nullptr);
        if (phase == PH_ITER || phase == PH_SIGMA) {
            if (phase == PH_SIGMA) {
                if (need && nodes + (r > 0.0 ? 1 : 0) > 1) flags |= FLAG_SIGMA_NOT_MAX;
                break;
            }
            const bool fine = lev == 0;
            if (!it.done) {
                nev += (1 << MAXLEV) >> lev;
                const double dp0 = it.dprev;
                iter_update(it, r, S, nodes, P.U, fine ? tol : tol_lvl, tol_stag, fine ? tol : 1e-7 * scale, true);
                // Each pass restarts from e_k, so the error obeys e' = C e^2 with C a property of the spectrum and of the
                // matching row (about the same on every level): two consecutive in-basin corrections measure it ...
                // (e0 = d0 + e1, e1 = C e0^2 ~ d1: C = d1 / (d0 + d1)^2)
                if (nodes == 0 && dp0 < 1e-2 * scale && it.dprev < 0.7 * dp0 && it.dprev > 0.0) {
                    const double e0 = dp0 + it.dprev;
                    Cq = it.dprev / (e0 * e0);
                }
                // ... and on the next levels ONE pass is enough when the error it leaves, C dl^2, is predicted (x 10) below
                // what the extrapolation to the finer level can use
                if (!it.done && !fine && it.rq && Cq > 0.0 && 10.0 * Cq * it.dprev * it.dprev <= tol_lvl) { it.done = true; it.conv = true; }
            }
#if IBS_SCAN2_EXTRAP
            // ... and what the step just taken leaves, C dl^2 (the Rayleigh quotient is below lam_max from either side), is
            // added to it: one order of convergence more per pass in the slow phase (C dl ~ 0.1 - 0.5) of a cold start.  An
            // overshoot lands just above lam_max, where the next pass is a full step again.
            if (!it.done && it.rq && Cq > 0.0) {
                const double x = Cq * it.dprev;
                if (x < 0.25) it.lam = fmin(it.lam + x * it.dprev / (1.0 - 2.0 * x), 0.5 * (it.lam + it.hi));
            }
#endif
            sh = it.lam;
            // Fine level: after an in-basin step whose successor is predicted tiny (q = C dl), go straight to the first output
            // pass at lam + dl -- it IS an iteration pass (out_join returns its correction) and confirms or rejects the step
            const bool ready = it.done || (fine && it.rq && Cq > 0.0 && 10.0 * Cq * it.dprev <= 1e-5);
            if (!ctx.all(ready)) continue;
            if (lev > 0) {
                rho3 = rho2; rho2 = rho1; rho1 = (it.conv && it.rho == it.rho) ? it.rho : qnan;
                ctx.warm_out(cs.wo[0], lev, rho1);
                if (lev == P.nlev) {
                    // how far was the neighbour?  (far or unknown: no corrections from it on the finer levels either)
                    const double n4 = ctx.warm_in(cs.wi[0], lev);
                    const double far = fabs(rho1 - n4) / fmax(fabs(rho1), 0.05 * scale);
                    if (!(far <= WARM_FAR)) cs.wi[0] = -1;
                    ctx.warm_out(cs.wo[0], MAXLEV + 1, (far == far) ? far : 0.0);
                    sh = (rho1 == rho1) ? rho1 : it.lam;
                    phase = PH_PEAK;
                    continue;
                }
            } else {
                if (!fin) {
                    prov = !it.done;
                    rbest = (!it.done || (it.conv && it.rho == it.rho)) ? it.rho : it.lam;       // (provisional: it.lam = it.rho)
                    sh = rbest;
                    if (it.done && !it.conv) flags |= FLAG_NOT_CONVERGED;
                }
                phase = PH_O1;
                continue;
            }
        } else if (phase == PH_PEAK) {
            const int jlo = ctx.min_i(out.jmax), jhi = ctx.max_i(out.jmax);
            const int kp = (jlo + jhi) / 2, km = (Nl - 1) / 2;
            k = snap_k((kp > km ? kp - km : km - kp) * PEAK_SNAP <= Nl ? km : kp, Nl);
        } else if (phase == PH_O1) {
            bool retry = false;
            if (prov && !fin) {
                // the confirming pass: d1 = |correction at lam + d0|.  Accept at the rounding level, or when d1 / d0 = q says
                // both that the eigenvalue after this correction is exact (d1 q^2 <= tol, the iteration's own predictive
                // stop) and that the eigenvector of THIS pass is clean (its contamination is ~ C d1 = q^2 <= 1e-10)
                const double d1 = fabs(out.dlt), d0 = it.dprev, q = d1 / d0;
                const bool acc = out.bad || d1 <= tol || (d1 < tol_stag && d1 >= 0.25 * d0) || (q <= 1e-5 && d1 * q * q <= tol);
                if (acc) {
                    if (!out.bad && out.dlt == out.dlt) rbest = sh + out.dlt;
                    it.done = true; it.conv = true; it.rho = rbest;
                } else {
                    retry = true;                             // go on iterating from here (an in-basin step when it is one)
                    it.lo = fmax(it.lo, sh);
                    if (out.dlt > 0.0 && sh + out.dlt <= it.hi) { it.lam = sh + out.dlt; it.dprev = d1; }
                    Cq = 0.0;
                }
                prov = false;
            }
            if (ctx.any(retry)) { sh = it.lam; phase = PH_ITER; continue; }
            const bool lowq = !fin && !out.bad && out.zmax > LOWQ && round < 3;
            const bool newly = !fin && !lowq;
            if (newly) {
                fin = true;
                res.gam = out.bad ? qnan : out.gam;
                res.rho = out.bad ? qnan : rbest;
                if (out.bad) flags = FLAG_BAD_INPUT;
                ctx.warm_out(cs.wo[0], 0, res.rho);
            }
            jsel = lowq ? out.jmax : -1;
            lowq_any = lowq;
            wr = newly && act;
            fix_any = wr && (out.bad || P.want_dX);
            ++round;
            if (want_out && ctx.any(wr)) { phase = PH_O2; continue; }
        } else {       // PH_O2
            if (ctx.any(fix_any)) ctx.fixup(wr && (out.bad || P.want_dX), Xrow, dXrow, N, out.bad, P.h, P.want_dX);
        }
        if (lev > 0) {
            --lev;
            Nl = level_n(N, lev);
            k = snap_k(2 * k, Nl);
            // start of the finer level: the eigenvalue is lam* + a h^2 + b h^4 + ...; two known levels remove a, three a and b
            double l0 = rho1;
            int order = 1;
            if (rho2 == rho2) {
                const double d12 = rho1 - rho2, d23 = rho2 - rho3;
                l0 = rho1 + 0.25 * d12;
                order = 2;
                if (rho3 == rho3 && fabs(d23) > 2.0 * fabs(d12) && fabs(d23) < 8.0 * fabs(d12)) { l0 = rho1 + 0.3125 * d12 - 0.015625 * d23; order = 3; }
            }
            {
                // the same extrapolation on the neighbour's levels, and the error it made there
                const int wi = cs.wi[0];
                const double n0 = ctx.warm_in(wi, lev), n1 = ctx.warm_in(wi, lev + 1);
                double e = n1;
                if (order >= 2) {
                    const double n2 = ctx.warm_in(wi, lev + 2), d12 = n1 - n2;
                    e = n1 + 0.25 * d12;
                    if (order == 3) e = n1 + 0.3125 * d12 - 0.015625 * (n2 - ctx.warm_in(wi, lev + 3));
                }
                const double corr = n0 - e;
                if (corr == corr && l0 == l0) l0 += corr;            // (NaN: no neighbour, or one of its levels did not converge)
            }
            iter_init(it, l0, P.Lb, P.U, false);
            sh = it.lam;
            phase = PH_ITER;
            continue;
        }
        if (ctx.any(lowq_any)) {
            k = snap_k(ctx.first_i(jsel), N);
            lowq_any = false;
            Cq = 0.0;                                         // (a property of the matching row too)
            iter_init(it, sh, P.Lb, P.U, fin);
            if (!fin) sh = it.lam;
            phase = PH_ITER;
            continue;
        }
        if (!has_sigma) break;
        need = (flags & FLAG_BAD_INPUT) == 0 && sigma < rbest;
        if (need) sh = 2.0 * sigma - rbest;
        if (!ctx.any(need)) break;
        phase = PH_SIGMA;
    }
    const int itc = (flags & FLAG_BAD_INPUT) ? 0 : ((flags & FLAG_NOT_CONVERGED) ? 64 : imin((nev + (1 << MAXLEV) - 1) >> MAXLEV, 63));
    res.info = itc | (flags << 16);
}

}  // namespace scan2
}  // namespace ibs
