// Batched bound-constrained refinement of the coarse maxima: the per-problem logic (host/device code).
//
// Replaces the per-surface scipy.optimize.minimize(obj_w_grad, jac=True, bounds=..., options={ftol, gtol, maxiter}) of
// /root/reference/ball_scan.py:305-314 (L-BFGS-B on the two variables (alpha, theta0)) by a projected quasi-Newton method
// that advances ALL surfaces in lock step: one batched evaluation of obj_w_grad per round (K1 for the three field lines of
// every surface + K2/K3 + K4) followed by one launch of the step kernel, which consumes the values / gradients, updates
// every problem's state and writes the next trial points in the layout the evaluation takes -- no host arithmetic, no
// host round trip per optimiser step, no thread per surface.
//
// Method per problem (minimise F = -lambda over the box lo <= x <= hi, n = 2):
//   * direction d = -H g on the free variables (a variable sitting on a bound whose gradient points outward is frozen),
//     H = BFGS approximation of the inverse Hessian (2 x 2, updated when s.y > 0, reset when the active set changes);
//   * projected line search: Armijo (c1 = 1e-4) with backtracking, and EXPANSION (step x 4, the last good point kept)
//     while the slope along d is still steeper than 0.9 of the initial one -- without it a region of negative curvature
//     (s.y <= 0: no BFGS update) would be crossed in steps of |g| ~ 1e-3; first trial step of a fresh H: 1 / |d|, a unit
//     step in (alpha, theta0), like L-BFGS-B;
//   * stopping tests of L-BFGS-B with the reference's options: projected-gradient inf-norm <= gtol (2e-8),
//     (F_k - F_k+1) / max(|F_k|, |F_k+1|, 1) <= ftol (5e-11), iterations >= maxiter (30).
// The iterates differ from scipy's (another quasi-Newton variant); parity is asserted on the objective / gradient
// evaluations and on the optimum reached (tests/test_ball_scan_gpu.py).
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define IBS_RHD __host__ __device__ __forceinline__
#else
#define IBS_RHD inline
#endif

namespace ibs {
namespace refine {

constexpr int ST_FIRST = 0, ST_SEARCH = 1, ST_DONE = 2;
constexpr int WHY_NONE = 0, WHY_PGTOL = 1, WHY_FTOL = 2, WHY_MAXITER = 3, WHY_STEP = 4, WHY_FAILED = 5;

// One problem's state: 30 doubles (kept as a plain array on the device: state[problem][NSTATE])
struct State {
    double x[2], f, g[2];        // accepted point, F and grad F there
    double H[3];                 // inverse-Hessian approximation (h00, h01, h11)
    double d[2], t;              // search direction and current step
    double xt[2];                // trial point (evaluated in the next round)
    double slope;                // g . d at the start of the line search
    double lo[2], hi[2];
    double status, why, nit, nfev, fresh;     // integers kept as doubles (one plain array per problem)
    double fbest;                // best F seen (monotone by construction; diagnostic)
    double xg[2], fg, gg[2], nexp;   // line search, expansion phase: last good trial point (value, gradient) and the number of expansions
};
constexpr int NSTATE = (int)(sizeof(State) / sizeof(double));
static_assert(NSTATE == 30, "layout documented in include/ibs_b200.h (status = 18, why = 19, nit = 20, nfev = 21)");

struct Options { double ftol, gtol; int maxiter; };

IBS_RHD double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// projected-gradient inf-norm: | x - P(x - g) |
IBS_RHD double pg_norm(const State& s) {
    double m = 0.0;
    for (int i = 0; i < 2; ++i) {
        const double p = s.x[i] - clampd(s.x[i] - s.g[i], s.lo[i], s.hi[i]);
        m = fmax(m, fabs(p));
    }
    return m;
}

// new search direction at the accepted point (free variables only) and the first trial point of its line search
IBS_RHD void new_direction(State& s) {
    bool fixed[2];
    for (int i = 0; i < 2; ++i)
        fixed[i] = (s.x[i] <= s.lo[i] && s.g[i] > 0.0) || (s.x[i] >= s.hi[i] && s.g[i] < 0.0);
    double d0 = -(s.H[0] * s.g[0] + s.H[1] * s.g[1]);
    double d1 = -(s.H[1] * s.g[0] + s.H[2] * s.g[1]);
    if (fixed[0]) { d0 = 0.0; d1 = -s.H[2] * s.g[1]; }
    if (fixed[1]) { d1 = 0.0; d0 = fixed[0] ? 0.0 : -s.H[0] * s.g[0]; }
    double slope = s.g[0] * d0 + s.g[1] * d1;
    if (!(slope < 0.0)) {                       // not a descent direction (indefinite H): steepest descent on the free variables
        s.H[0] = 1.0; s.H[1] = 0.0; s.H[2] = 1.0; s.fresh = 1.0;
        d0 = fixed[0] ? 0.0 : -s.g[0];
        d1 = fixed[1] ? 0.0 : -s.g[1];
        slope = s.g[0] * d0 + s.g[1] * d1;
    }
    s.d[0] = d0; s.d[1] = d1; s.slope = slope;
    const double dn = sqrt(d0 * d0 + d1 * d1);
    s.t = (s.fresh != 0.0 && dn > 0.0) ? 1.0 / dn : 1.0;                  // L-BFGS-B's first step: 1 / |d| (a unit step in x)
    s.nexp = 0.0;
    s.xt[0] = clampd(s.x[0] + s.t * d0, s.lo[0], s.hi[0]);
    s.xt[1] = clampd(s.x[1] + s.t * d1, s.lo[1], s.hi[1]);
}

IBS_RHD void init(State& s, double a0, double t0, double alo, double ahi, double tlo, double thi) {
    s.lo[0] = alo; s.hi[0] = ahi; s.lo[1] = tlo; s.hi[1] = thi;
    s.x[0] = s.xt[0] = clampd(a0, alo, ahi);
    s.x[1] = s.xt[1] = clampd(t0, tlo, thi);
    s.f = 0.0; s.g[0] = s.g[1] = 0.0;
    s.H[0] = 1.0; s.H[1] = 0.0; s.H[2] = 1.0;
    s.d[0] = s.d[1] = 0.0; s.t = 1.0; s.slope = 0.0;
    s.status = ST_FIRST; s.why = WHY_NONE; s.nit = 0.0; s.nfev = 0.0; s.fresh = 1.0; s.fbest = 0.0;
    s.xg[0] = s.xg[1] = 0.0; s.fg = 0.0; s.gg[0] = s.gg[1] = 0.0; s.nexp = 0.0;
}

IBS_RHD void finish(State& s, int why) { s.status = ST_DONE; s.why = why; s.xt[0] = s.x[0]; s.xt[1] = s.x[1]; }

// Consume the evaluation (ft, gt) of the trial point s.xt; `failed` = the eigen-solve flagged the point.
IBS_RHD void consume(State& s, double ft, double gt0, double gt1, bool failed, const Options& o) {   // (ft, gt: by value, may be replaced)
    if ((int)s.status == ST_DONE) return;
    s.nfev += 1.0;
    const bool finite = (ft == ft) && (gt0 == gt0) && (gt1 == gt1) && fabs(ft) < 1e300;
    if ((int)s.status == ST_FIRST) {
        if (failed || !finite) { finish(s, WHY_FAILED); return; }
        s.f = ft; s.g[0] = gt0; s.g[1] = gt1; s.fbest = ft;
        if (pg_norm(s) <= o.gtol) { finish(s, WHY_PGTOL); return; }
        s.status = ST_SEARCH;
        new_direction(s);
        return;
    }
    // ---- line search along the projected path
    const bool usable = !failed && finite;
    double s0 = s.xt[0] - s.x[0], s1 = s.xt[1] - s.x[1];
    const double gs = s.g[0] * s0 + s.g[1] * s1;                        // decrease predicted by the projected step
    const bool moved = (s0 != 0.0) || (s1 != 0.0);
    bool accept = usable && moved && ft <= s.f + 1e-4 * gs;
    if (accept && !(s.nexp > 0.0 && ft > s.fg)) {
        // Armijo holds (and an expanded trial did not overshoot).  Still descending steeply along d?  Then try further out.
        const double dd = gt0 * s.d[0] + gt1 * s.d[1];
        const double n0 = clampd(s.x[0] + 4.0 * s.t * s.d[0], s.lo[0], s.hi[0]), n1 = clampd(s.x[1] + 4.0 * s.t * s.d[1], s.lo[1], s.hi[1]);
        const bool can_move = (n0 != s.xt[0]) || (n1 != s.xt[1]);
        if (dd < 0.9 * s.slope && can_move && s.nexp < 8.0) {
            s.xg[0] = s.xt[0]; s.xg[1] = s.xt[1]; s.fg = ft; s.gg[0] = gt0; s.gg[1] = gt1;
            s.nexp += 1.0;
            s.t *= 4.0;
            s.xt[0] = n0; s.xt[1] = n1;
            return;
        }
    } else if (s.nexp > 0.0) {
        // the expanded trial overshot (or failed): take the last good point
        s.xt[0] = s.xg[0]; s.xt[1] = s.xg[1]; ft = s.fg; gt0 = s.gg[0]; gt1 = s.gg[1];
        s0 = s.xt[0] - s.x[0]; s1 = s.xt[1] - s.x[1];
        accept = true;
    }
    if (accept) {
        // accept: BFGS update of H with (s, y) when the curvature condition holds
        const double y0 = gt0 - s.g[0], y1 = gt1 - s.g[1];
        const double sy = s0 * y0 + s1 * y1;
        const double fold = s.f;
        if (sy > 1e-12 * sqrt((s0 * s0 + s1 * s1) * (y0 * y0 + y1 * y1))) {
            if (s.fresh != 0.0) {                                        // scale the initial H (Nocedal & Wright 6.20)
                const double sc = sy / (y0 * y0 + y1 * y1);
                s.H[0] = sc; s.H[1] = 0.0; s.H[2] = sc; s.fresh = 0.0;
            }
            const double rho = 1.0 / sy;
            const double Hy0 = s.H[0] * y0 + s.H[1] * y1, Hy1 = s.H[1] * y0 + s.H[2] * y1;
            const double yHy = y0 * Hy0 + y1 * Hy1;
            const double c = (1.0 + rho * yHy) * rho;
            s.H[0] += c * s0 * s0 - rho * (Hy0 * s0 + s0 * Hy0);
            s.H[1] += c * s0 * s1 - rho * (Hy0 * s1 + s0 * Hy1);
            s.H[2] += c * s1 * s1 - rho * (Hy1 * s1 + s1 * Hy1);
        }
        s.x[0] = s.xt[0]; s.x[1] = s.xt[1]; s.f = ft; s.g[0] = gt0; s.g[1] = gt1;
        s.fbest = fmin(s.fbest, ft);
        s.nit += 1.0;
        if (pg_norm(s) <= o.gtol) { finish(s, WHY_PGTOL); return; }
        if ((fold - ft) <= o.ftol * fmax(fmax(fabs(fold), fabs(ft)), 1.0)) { finish(s, WHY_FTOL); return; }
        if ((int)s.nit >= o.maxiter) { finish(s, WHY_MAXITER); return; }
        new_direction(s);
        return;
    }
    // reject: shrink the step (safeguarded quadratic interpolation when the trial value is usable)
    double shrink = 0.5;
    if (!failed && finite && moved && s.slope < 0.0) {
        const double q = -s.slope * s.t / (2.0 * (ft - s.f - s.slope * s.t) / s.t);          // minimiser of the interpolating parabola
        if (q == q) shrink = clampd(q / s.t, 0.1, 0.5);
    }
    s.t *= shrink;
    const double n0 = clampd(s.x[0] + s.t * s.d[0], s.lo[0], s.hi[0]), n1 = clampd(s.x[1] + s.t * s.d[1], s.lo[1], s.hi[1]);
    if ((n0 == s.x[0] && n1 == s.x[1]) || s.t < 1e-14) { finish(s, WHY_STEP); return; }     // no representable progress left
    s.xt[0] = n0; s.xt[1] = n1;
}

}  // namespace refine
}  // namespace ibs
