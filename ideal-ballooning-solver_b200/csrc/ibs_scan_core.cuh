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
// (every 16th, 8th, 4th, 2nd point; Richardson-extrapolated): ~5 fine-grid-equivalent evaluations instead of ~12.
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
// The streaming passes: optionally compiled as real functions (IBS_SCAN_NOINLINE), so that the caller's cold state is
// saved once at the call instead of competing for registers with the pass's loops.
#if defined(__CUDACC__) && defined(IBS_SCAN_NOINLINE)
#define IBS_PASS __host__ __device__ __noinline__
#else
#define IBS_PASS IBS_HD
#endif

namespace ibs {
namespace scan {

#ifndef IBS_SCAN_BLK
#define IBS_SCAN_BLK 4
#endif
#ifndef IBS_SCAN_BLK_OUT
#define IBS_SCAN_BLK_OUT 2
#endif
// Steps per iteration of the streaming loops.  The loops must stay small: two warps share a scheduler's L0 instruction cache
// and ncu shows `no_instruction` stalls at EVERY 128-byte line of a loop that does not fit beside its neighbour's (25 % of the
// samples of a 5.3 KB output-pass loop).  Measured on the bench batch (ms per 303 104 solves, prep included; round 1: 4.72):
// iteration / output = 2 / 2: 4.24, 4 / 4: 4.32, 4 / 2: 4.23, 2 / 4: 4.28; two inlined copies of the 4-step loops: 5.6.
constexpr int SCAN_BLK = IBS_SCAN_BLK;           // iteration pass: 4 steps = 2.3 KB (2 steps: 1.2 KB)
constexpr int SCAN_BLK_OUT = IBS_SCAN_BLK_OUT;   // output passes: 2 steps = 3.0 / 1.9 KB (4 steps: 5.3 / 3.7 KB)
constexpr int TR = 32;            // records per tile (one pipeline stage holds one forward and one backward tile)
constexpr int REC = 6;            // doubles per record: G0 G1 G2 (g = G0 + th0 G1 + th0^2 G2), C0 C1 (2 h^2 c), R (2 h^2 f = g R)
constexpr int MAXLEV = 4;         // coarse levels (strides 2, 4, 8, 16)
constexpr int MIN_COARSE_N = 65;  // a level is used only if it has at least this many points
constexpr int MAXIT_LEVEL = 64;   // evaluations per level before giving up
constexpr int K_MARGIN = 8;       // matching row kept this far from both ends
constexpr double LOWQ = 1.0e3;    // max|z| / |z_k| above which the matching row is moved (see solve_item)
// (Tried and removed: letting the output pass stand in for the last fine-level evaluation.  It saves ~1 of ~6 fine-grid
// passes per lane, but its vector is then taken at a shift that is only ~1e-11 converged and the Simpson Rayleigh
// quotient -- which weighs the far end of the spectrum heavily -- moves by up to 6e-11 relative; and at warp level the
// acceptance test has to hold for all 32 lanes, so there was no measurable speed-up.)
#ifndef IBS_PEAK_SNAP
#define IBS_PEAK_SNAP 8
#endif
// A coarsest-level peak within Nl / PEAK_SNAP rows of the middle keeps the middle matching row.  With unequal chains the rows
// beyond the shorter one run on the one-step-at-a-time path (~7x the cost of a pipelined step), so the row is moved only
// when it has to be: measured on the bench's D3D-like lines, 16 -> 8 takes the general steps from 15 % of all steps to
// 1 per pass with the same number of passes (NCSX-like lines: 28 % -> 1 per pass, 24.0 -> 28.4 passes per solve).
constexpr int PEAK_SNAP = IBS_PEAK_SNAP;

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

IBS_HD int sign_changes_n(unsigned m, int n, unsigned enter_sign) {   // n <= 32 steps recorded; bit (n-1-i) = sign after step i
    const unsigned pm = (n >= 32) ? 0x7fffffffu : ((1u << (n - 1)) - 1u);
    const unsigned v = (m ^ (m >> 1)) & pm;
#if defined(__CUDA_ARCH__)
    return __popc(v) + (int)(((m >> (n - 1)) & 1u) ^ enter_sign);
#else
    return __builtin_popcount(v) + (int)(((m >> (n - 1)) & 1u) ^ enter_sign);
#endif
}

// ---- software-pipelined form of the chain steps ------------------------------------------------------------------
// ptxas keeps the source order of a loop body almost unchanged and a warp issues in order, so the fast paths are
// written as a pipeline by hand: the records of step i+2 are loaded, the coefficients of step i+1 are formed and the
// chains of step i are advanced in ONE interleaved instruction stream (forward and backward statements alternate),
// so that neither the shared-memory latency nor the 8-cycle DFMA latency is exposed.
struct Co { double a, t, F; };        // of the row being stepped to: a = g + g_prev (= 2 gh of the half cell), t', F'

IBS_HD void coef_next(const Rec& r, double th0, double lam, double& gp, Co& c) {
    const double g = fma(th0, fma(th0, r.G2, r.G1), r.G0);
    const double C = fma(th0, r.C1, r.C0);
    c.F = g * r.R;
    c.a = g + gp;
    gp = g;
    c.t = fma(-lam, c.F, C);
}
IBS_HD void fwd_chain(const Co& c, double& X, double& W, double& S) {
    const double Xn = fma(c.a, X, W);
    W = fma(-c.t, Xn, c.a * W);
    S = fma(c.F * Xn, Xn, (c.a * c.a) * S);
    X = Xn;
}
IBS_HD void bwd_chain(const Co& c, double& X, double& W, double& S, double& tcur) {
    const double tmp = fma(tcur, X, W);
    const double Xn = fma(c.a, X, -tmp);
    W = c.a * tmp;
    S = fma(c.F * Xn, Xn, (c.a * c.a) * S);
    X = Xn;
    tcur = c.t;
}
// chains of the current joint step (coefficients cf, cb) + coefficients of the next one from the records nf, nb
IBS_HD void joint_step(const Rec& nf, const Rec& nb, double th0, double lam, Co& cf, Co& cb, double& gf, double& gb,
                       double& Xf, double& Wf, double& Sf, double& Xb, double& Wb, double& Sb, double& tb) {
    const double p1 = fma(th0, nf.G2, nf.G1);
    const double XnF = fma(cf.a, Xf, Wf);
    const double p2 = fma(th0, nb.G2, nb.G1);
    const double tmp = fma(tb, Xb, Wb);
    const double CF = fma(th0, nf.C1, nf.C0);
    const double aWF = cf.a * Wf;
    const double CB = fma(th0, nb.C1, nb.C0);
    const double a2F = cf.a * cf.a;
    const double gF = fma(th0, p1, nf.G0);
    const double XnB = fma(cb.a, Xb, -tmp);
    const double gB = fma(th0, p2, nb.G0);
    Wf = fma(-cf.t, XnF, aWF);
    const double a2B = cb.a * cb.a;
    Wb = cb.a * tmp;
    const double FXF = cf.F * XnF;
    const double a2SF = a2F * Sf;
    const double FFn = gF * nf.R;
    const double FXB = cb.F * XnB;
    const double FBn = gB * nb.R;
    const double a2SB = a2B * Sb;
    Sf = fma(FXF, XnF, a2SF);
    Sb = fma(FXB, XnB, a2SB);
    tb = cb.t;
    Xf = XnF; Xb = XnB;
    cf.a = gF + gf; gf = gF; cf.F = FFn;
    cb.a = gB + gb; gb = gB; cb.F = FBn;
    cf.t = fma(-lam, FFn, CF);
    cb.t = fma(-lam, FBn, CB);
}

// One evaluation at the shifts lam[]: twisted residual r' (= 2 r), S' = sum 2F z^2 (z_k = 1), node count.
// rho = lam + r' / S' is the Rayleigh quotient of z; #eigenvalues above lam = nodes + (r' > 0).
template <int SPL, class Ctx>
IBS_PASS void eval_pass(Ctx& ctx, int lev, int Nl, int k, const double (&th0)[SPL], const double (&lam)[SPL],
                      double (&r)[SPL], double (&S)[SPL], int (&nodes)[SPL]) {
    const int qf_end = k, qb_end = Nl - 1 - k;           // last record of each direction (row k)
    const int qmin = imin(qf_end, qb_end), qmax = imax(qf_end, qb_end);
    const int nst = qmax / TR + 1;
    const int nfs = qmin / TR;                            // stages whose 32 steps are interior in both directions
    ctx.begin_pass(lev, Nl, k, nst);
    double Xf[SPL], Wf[SPL], Sf[SPL], gf[SPL], Xb[SPL], Wb[SPL], Sb[SPL], gb[SPL], tb[SPL];
#pragma unroll
    for (int q = 0; q < SPL; ++q) nodes[q] = 0;
    // ---- stage 0: records 0 and 1 of both directions start the chains
    ctx.wait(0);
    {
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
    }
    int s_next = 0;          // first stage the general loop below has to handle (stage 0 from step 2 if there is no fast region)
    int i_first = 2;
    if (nfs > 0) {
        // ---- fast region: steps 2 .. 32 nfs - 1, ONE software pipeline across the stage boundaries (records two steps
        // ahead through running pointers, coefficients one step ahead).  Control flow is kept out of the hot code: its
        // only branch is the back edge of a counted loop over blocks of SCAN_BLK steps (a taken branch costs a warp
        // ~15 idle cycles, and with two warps per scheduler nothing hides them; round 1's loop took three per two steps).
        // The steps 32 m - 2 .. 32 m + 29 read records of stage m only: wait / pointer reset before them, rescaling
        // every 16 steps, release of the previous stage after the first 16 (its last records were consumed by the first
        // block), sign histories of exactly 32 steps counted after the second 16.
        const int gend = TR * nfs;
        Co cf[SPL], cb[SPL];
        unsigned mf[SPL], mb[SPL], ef[SPL], eb[SPL];
        const double* pf = ctx.frec(2);
        const double* pb = ctx.brec(2);
        {
            const Rec f = load_rec(pf), b = load_rec(pb);
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                coef_next(f, th0[q], lam[q], gf[q], cf[q]);
                coef_next(b, th0[q], lam[q], gb[q], cb[q]);
                // the first trip records 28 steps only: the history is pre-filled with the entering sign
                ef[q] = sign_bit(Xf[q]); eb[q] = sign_bit(Xb[q]);
                mf[q] = 0u - ef[q]; mb[q] = 0u - eb[q];
            }
        }
        pf += REC; pb -= REC;
        Rec rf = load_rec(pf), rb = load_rec(pb);         // records of step 3
        pf += REC; pb -= REC;                              // -> records of step 4
        auto step1 = [&](const Rec& af, const Rec& ab) {   // chains of one step + coefficients of the next one from (af, ab)
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                joint_step(af, ab, th0[q], lam[q], cf[q], cb[q], gf[q], gb[q], Xf[q], Wf[q], Sf[q], Xb[q], Wb[q], Sb[q], tb[q]);
                mf[q] = (mf[q] << 1) | sign_bit(Xf[q]);
                mb[q] = (mb[q] << 1) | sign_bit(Xb[q]);
            }
        };
        auto rescale_all = [&]() {
#pragma unroll
            for (int q = 0; q < SPL; ++q) { rescale3(Xf[q], Wf[q], Sf[q]); rescale3(Xb[q], Wb[q], Sb[q]); }
        };
        // One copy of the block loop (instruction cache: two warps share a scheduler's L0), run once per HALF stage:
        // half hf covers the steps 16 hf - 2 .. 16 hf + 13 (hf = 0: 2 .. 13); the few tests per half are off the hot path.
#pragma unroll 1
        for (int hf = 0; hf < 2 * nfs; ++hf) {
            if ((hf & 1) == 0 && hf > 0) { ctx.wait(hf >> 1); pf = ctx.frec(0); pb = ctx.brec(0); }
            const int nblk = (hf > 0) ? 16 / SCAN_BLK : 12 / SCAN_BLK;
#pragma unroll 1
            for (int b = 0; b < nblk; ++b) {               // SCAN_BLK steps; loads the records of the steps g+2 .. g+SCAN_BLK+1
                Rec nf = load_rec(pf), nb = load_rec(pb);
                step1(rf, rb);
                rf = load_rec(pf + REC); rb = load_rec(pb - REC);
                step1(nf, nb);
                if (SCAN_BLK == 4) {
                    nf = load_rec(pf + 2 * REC); nb = load_rec(pb - 2 * REC);
                    step1(rf, rb);
                    rf = load_rec(pf + 3 * REC); rb = load_rec(pb - 3 * REC);
                    step1(nf, nb);
                }
                pf += SCAN_BLK * REC; pb -= SCAN_BLK * REC;
            }
            rescale_all();
            if (hf & 1) {
#pragma unroll
                for (int q = 0; q < SPL; ++q) {
                    nodes[q] += sign_changes32(mf[q], ef[q]) + sign_changes32(mb[q], eb[q]);
                    ef[q] = sign_bit(Xf[q]); eb[q] = sign_bit(Xb[q]);
                    mf[q] = 0; mb[q] = 0;
                }
            } else if (hf > 0) {
                ctx.release((hf >> 1) - 1);                // its last records were consumed by the first block of this half
            }
        }
        // steps gend - 2 (its successor's coefficients from the record already loaded) and gend - 1 (chains only)
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            joint_step(rf, rb, th0[q], lam[q], cf[q], cb[q], gf[q], gb[q], Xf[q], Wf[q], Sf[q], Xb[q], Wb[q], Sb[q], tb[q]);
            mf[q] = (mf[q] << 1) | sign_bit(Xf[q]);
            mb[q] = (mb[q] << 1) | sign_bit(Xb[q]);
            fwd_chain(cf[q], Xf[q], Wf[q], Sf[q]);
            bwd_chain(cb[q], Xb[q], Wb[q], Sb[q], tb[q]);
            mf[q] = (mf[q] << 1) | sign_bit(Xf[q]);
            mb[q] = (mb[q] << 1) | sign_bit(Xb[q]);
            nodes[q] += sign_changes_n(mf[q], 2, ef[q]) + sign_changes_n(mb[q], 2, eb[q]);
        }
        rescale_all();
        ctx.release(nfs - 1);
        s_next = nfs;
        i_first = 0;
    }
    // ---- general steps (the ends of the chains, unequal chains): one step at a time with per-step tests
    for (int s = s_next; s < nst; ++s) {
        if (s > 0 || nfs > 0) ctx.wait(s);
#pragma unroll 1
        for (int i = (s == s_next) ? i_first : 0; i < TR && TR * s + i <= qmax; ++i) {
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
        ctx.release(s);
    }
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        const double inv = 1.0 / (Xf[q] * Xb[q]);
        r[q] = fma(Wb[q], Xf[q], -(Wf[q] * Xb[q])) * inv;
        S[q] = fma(Sf[q] * Xb[q], Xb[q], Sb[q] * Xf[q] * Xf[q]) * inv * inv;
    }
}

// ---- output pass (plain recurrence with reciprocals: the un-scaled eigenfunction is needed) ---------------
// One direction of one solve.  The window holds the four previous rows' values; the sums are split by row parity
// (Simpson weights 4/3 and 2/3) and carry the running power-of-two scale 2^-E of x.
struct Sweep {
    double x, w; int E;
    double W1, W2, W3, W4;        // x of the previous rows: W1 = the row computed last, ... (running scale)
    double gp, gpp;               // g of the last row and of the one before
    double tcur;                  // backward: t' of the current row
    double a0e, a0o, a1e, a1o, aDe, aDo;   // sum t' X^2, sum F' X^2, sum g D^2 of the even / odd rows (no dynamic indexing)
    double aEnd;                  // g D^2 of the Dirichlet end point
    double vmax; int jmax;        // largest |x| so far (running scale) and its row
    bool bad;
    double fsc, cn; int ex;       // writing pass: X = x * fsc, fsc = cn 2^ex (ex follows the rescalings)
};

struct SolveOut {                 // what the output pass returns per solve
    double gam, zmax; int jmax; bool bad;
    double dlt;                        // lane-per-chain kernel: the Rayleigh-quotient correction r / S of this pass's shift (out_join)
    double xkf, xkb; int Ekf, Ekb;     // the two sweeps at the matching row: value and scale exponent
};

constexpr double C23 = 2.0 / 3.0, C12 = 1.0 / 12.0;
constexpr int EBLK = 16;          // rows between two rescalings

template <bool WRITE>
IBS_HD void sweep_rescale(Sweep& sw) {
    const int e = exp_max2(sw.x, sw.w);
    const double s = pow2(-e);
    sw.x *= s; sw.w *= s; sw.E += e;
    if (WRITE) {
        sw.ex += e;
        sw.fsc = sw.cn * pow2(sw.ex);
    } else {
        const double s2 = s * s;
        sw.W1 *= s; sw.W2 *= s; sw.W3 *= s; sw.W4 *= s;
        sw.a0e *= s2; sw.a0o *= s2; sw.a1e *= s2; sw.a1o *= s2; sw.aDe *= s2; sw.aDo *= s2; sw.aEnd *= s2;
        sw.vmax *= s;
    }
}

// Normalised output value.  The largest element must come out as exactly 1 (utils.py:1605 divides by the maximum):
// the scale factor is inflated by a few ulp (NORM_INFLATE) and the product clamped from above.
constexpr double NORM_INFLATE = 1.0 + 1.7763568394002505e-15;      // 1 + 2^-49
IBS_HD double norm_value(double x, double fsc) { return fmin(x * fsc, 1.0); }

// One step of the output pass for one direction, general form (first rows, last rows, unequal chains).
// DIR = +1 forward (new row = qq), -1 backward (new row = Nl-1-qq).  `last` (backward only): the step that reaches
// row k (its X^2 terms and its X belong to the forward sweep).  WRITE: the normalised X goes to Xw (if not null) and
// nothing is summed.
template <int DIR, bool WRITE>
IBS_HD void out_step(Sweep& sw, const Rec& rc, double th0, double lam, int qq, int Nl, bool last, double* Xw) {
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
    if (WRITE) {
        if (!last && Xw) Xw[row] = norm_value(xn, sw.fsc);
        sw.gp = g;
        return;
    }
    const int par = qq & 1;                    // = parity of the row (N odd)
    if (!last) {
        const double x2 = xn * xn;
        if (par) { sw.a0o = fma(tnew, x2, sw.a0o); sw.a1o = fma(F, x2, sw.a1o); }
        else     { sw.a0e = fma(tnew, x2, sw.a0e); sw.a1e = fma(F, x2, sw.a1e); }
        const double ax = fabs(xn);
        if (ax > sw.vmax) { sw.vmax = ax; sw.jmax = row; }
    }
    // stencil of the row two behind (h-free: D = h dX; only D^2 is needed here)
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

IBS_HD void sweep_zero(Sweep& sw) {
    sw.E = 0; sw.W1 = sw.W2 = sw.W3 = sw.W4 = 0.0; sw.tcur = 0.0;
    sw.a0e = sw.a0o = sw.a1e = sw.a1o = sw.aDe = sw.aDo = 0.0; sw.aEnd = 0.0;
    sw.vmax = 0.0; sw.jmax = 0; sw.bad = false;
}

// ---- pipelined form of the output-pass step (interior rows: generic stencil, never the last backward step) -----
struct OCo { double ia, t, F, g; };   // of the row being stepped to: 1 / (g + g_prev), t', F', g

// chain + tail of one direction for the row whose coefficients are c (PAR = parity of the row, compile time)
template <int DIR, int PAR, bool WRITE>
IBS_HD void out_chain_tail(Sweep& sw, const OCo& c, int row, double* Xw) {
    double xn;
    if (DIR > 0) { xn = fma(sw.w, c.ia, sw.x); sw.w = fma(-c.t, xn, sw.w); }
    else         { sw.w = fma(sw.tcur, sw.x, sw.w); xn = fma(-sw.w, c.ia, sw.x); sw.tcur = c.t; }
    sw.x = xn;
    if (WRITE) {
        if (Xw) Xw[row] = norm_value(xn, sw.fsc);
        sw.gp = c.g;
        return;
    }
    const double d1 = sw.W1 - sw.W3;
    const double x2 = xn * xn;
    const double d2 = xn - sw.W4;
    const double ax = fabs(xn);
    if (PAR) { sw.a0o = fma(c.t, x2, sw.a0o); sw.a1o = fma(c.F, x2, sw.a1o); }
    else     { sw.a0e = fma(c.t, x2, sw.a0e); sw.a1e = fma(c.F, x2, sw.a1e); }
    const double D = fma(C23, d1, -(C12 * d2));
    if (ax > sw.vmax) { sw.vmax = ax; sw.jmax = row; }
    const double D2 = D * D;
    if (PAR) sw.aDo = fma(sw.gpp, D2, sw.aDo); else sw.aDe = fma(sw.gpp, D2, sw.aDe);
    sw.W4 = sw.W3; sw.W3 = sw.W2; sw.W2 = sw.W1; sw.W1 = xn;
    sw.gpp = sw.gp; sw.gp = c.g;
}

// coefficients (with the reciprocal) of the next row of one direction, not interleaved (prologue of a tile)
template <bool WRITE>
IBS_HD void out_coef(const Rec& r, double th0, double lam, double gprev, OCo& c, bool& bad) {
    double C;
    coef(r, th0, c.g, C, c.F);
    const double a = c.g + gprev;
    if (!WRITE) bad |= not_pos_normal(a) | not_pos_normal(c.F) | not_finite(C);
    c.ia = rcp_fast(a);
    c.t = fma(-lam, c.F, C);
}

// One joint step: chains + tails of the current rows (coefficients cf, cb) and, interleaved with them, the
// coefficients of the next rows from the records nf, nb (the two reciprocal chains are the long dependences).
template <int PAR, bool WRITE>
IBS_HD void out_joint(const Rec& nf, const Rec& nb, double th0, double lam, OCo& cf, OCo& cb, Sweep& f, Sweep& b, int qq, int Nl,
                      double* Xw) {
    // next coefficients, part 1: g, a and the reciprocal seeds
    const double pA = fma(th0, nf.G2, nf.G1);
    const double pB = fma(th0, nb.G2, nb.G1);
    const double CA = fma(th0, nf.C1, nf.C0);
    const double CB = fma(th0, nb.C1, nb.C0);
    const double gA = fma(th0, pA, nf.G0);
    const double gB = fma(th0, pB, nb.G0);
    const double aA = gA + cf.g;
    const double aB = gB + cb.g;
#if defined(__CUDA_ARCH__)
    double rA, rB;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rA) : "d"(aA));
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rB) : "d"(aB));
#else
    double rA = 1.0 / aA, rB = 1.0 / aB;
#endif
    const double FA = gA * nf.R;
    const double FB = gB * nb.R;
    if (!WRITE) {
        f.bad |= not_pos_normal(aA) | not_pos_normal(FA) | not_finite(CA);
        b.bad |= not_pos_normal(aB) | not_pos_normal(FB) | not_finite(CB);
    }
    // current forward row
    out_chain_tail<+1, PAR, WRITE>(f, cf, qq, Xw);
    // next coefficients, part 2: first Newton step of the reciprocals (cubic)
    double eA = fma(-aA, rA, 1.0);
    double eB = fma(-aB, rB, 1.0);
    const double tA = fma(-lam, FA, CA);
    const double tB = fma(-lam, FB, CB);
    eA = fma(eA, eA, eA);
    eB = fma(eB, eB, eB);
    rA = fma(rA, eA, rA);
    rB = fma(rB, eB, rB);
    // current backward row
    out_chain_tail<-1, PAR, WRITE>(b, cb, Nl - 1 - qq, Xw);
    // next coefficients, part 3: second Newton step
#if defined(__CUDA_ARCH__)
    eA = fma(-aA, rA, 1.0);
    eB = fma(-aB, rB, 1.0);
    rA = fma(rA, eA, rA);
    rB = fma(rB, eB, rB);
#endif
    cf.ia = rA; cf.t = tA; cf.F = FA; cf.g = gA;
    cb.ia = rB; cb.t = tB; cb.F = FB; cb.g = gB;
}

// Output pass at the shifts lam[] with matching row k.
//   WRITE = false: Simpson Rayleigh quotient gam (utils.py:1605-1621), max|z| and its row, validity -> out[]
//   WRITE = true : the chains are recomputed (bit-identically) and X = z / max|z| is written to Xw[q] (where not null),
//                  using the scales out[] of the first pass
template <int SPL, bool WRITE, class Ctx>
IBS_PASS void out_pass(Ctx& ctx, int lev, int Nl, int k, const double (&th0)[SPL], const double (&lam)[SPL],
                     SolveOut (&out)[SPL], double* const (&Xw)[SPL]) {
    const int qf_end = k, qb_end = Nl - 1 - k;
    const int qmin = imin(qf_end, qb_end), qmax = imax(qf_end, qb_end);
    const int nst = qmax / TR + 1;
    const int nfs = qmin / TR;                            // stages whose 32 steps are interior rows in both directions
    ctx.begin_pass(lev, Nl, k, nst);
    Sweep F_[SPL], B_[SPL];
    auto rescale_all = [&](bool do_f, bool do_b) {
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            if (do_f) sweep_rescale<WRITE>(F_[q]);
            if (do_b) sweep_rescale<WRITE>(B_[q]);
        }
    };
    // ---- stage 0: rows 0, 1 (and N-1, N-2) start the sweeps
    ctx.wait(0);
    {
        const Rec f0 = load_rec(ctx.frec(0)), f1 = load_rec(ctx.frec(1));
        const Rec b0 = load_rec(ctx.brec(0)), b1 = load_rec(ctx.brec(1));
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            double g, C, Fv;
            Sweep& f = F_[q]; Sweep& b = B_[q];
            sweep_zero(f); sweep_zero(b);
            if (WRITE) {
                const bool ok = !out[q].bad && out[q].zmax > 0.0 && out[q].zmax < 1e300;
                f.cn = ok ? NORM_INFLATE / (out[q].xkf * out[q].zmax) : 0.0;
                b.cn = ok ? NORM_INFLATE / (out[q].xkb * out[q].zmax) : 0.0;
                f.ex = -out[q].Ekf; b.ex = -out[q].Ekb;
                f.fsc = f.cn * pow2(f.ex); b.fsc = b.cn * pow2(b.ex);
                if (Xw[q]) { Xw[q][0] = 0.0; Xw[q][Nl - 1] = 0.0; }
            }
            // forward: row 0 (x = 0, w' = 1), then the ordinary step to row 1
            coef(f0, th0[q], g, C, Fv);
            f.x = 0.0; f.w = 1.0; f.gp = g; f.gpp = g;
            out_step<+1, WRITE>(f, f1, th0[q], lam[q], 1, Nl, false, Xw[q]);
            // backward: row N-1 (x = 0), row M = N-2 with x = 1, w' = -a_M
            coef(b0, th0[q], g, C, Fv);
            const double gN = g;
            coef(b1, th0[q], g, C, Fv);
            const double a = g + gN;
            b.x = 1.0; b.w = -a; b.gp = g; b.gpp = gN; b.tcur = fma(-lam[q], Fv, C);
            if (WRITE) { if (Xw[q]) Xw[q][Nl - 2] = norm_value(1.0, b.fsc); }
            else {
                b.bad |= not_pos_normal(a) | not_pos_normal(Fv) | not_finite(C);
                b.a0o = b.tcur; b.a1o = Fv; b.vmax = 1.0; b.jmax = Nl - 2;
                b.W1 = 1.0;
            }
        }
    }
    int s_next = 0, i_first = 2;
    if (nfs > 0) {
        // rows 2 and 3 (and N-3, N-4) have their own stencils, rows 4 and 5 bring the pipeline to a step = 2 (mod 4):
        // general steps
        for (int i = 2; i < 6; ++i) {
            const Rec rf = load_rec(ctx.frec(i)), rb = load_rec(ctx.brec(i));
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                out_step<+1, WRITE>(F_[q], rf, th0[q], lam[q], i, Nl, false, Xw[q]);
                out_step<-1, WRITE>(B_[q], rb, th0[q], lam[q], i, Nl, false, Xw[q]);
            }
        }
        // ---- fast region: steps 6 .. 32 nfs - 1 as ONE software pipeline across the stage boundaries (coefficients one
        // step ahead, records two steps ahead), in blocks of four steps with no branch but the back edge (see eval_pass;
        // four steps also let the 4-row stencil window rotate through register names instead of being moved).  A trip of
        // the stage loop covers the steps 32 m - 2 .. 32 m + 29, whose records all lie in stage m.
        const int gend = TR * nfs;
        OCo cf[SPL], cb[SPL];
        const double* pf = ctx.frec(6);
        const double* pb = ctx.brec(6);
        {
            const Rec f0 = load_rec(pf), b0 = load_rec(pb);
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                out_coef<WRITE>(f0, th0[q], lam[q], F_[q].gp, cf[q], F_[q].bad);
                out_coef<WRITE>(b0, th0[q], lam[q], B_[q].gp, cb[q], B_[q].bad);
            }
        }
        pf += REC; pb -= REC;
        Rec rf = load_rec(pf), rb = load_rec(pb);         // records of step 7
        pf += REC; pb -= REC;                              // -> records of step 8
        int g = 6;
#pragma unroll 1
        for (int hf = 0; hf < 2 * nfs; ++hf) {             // half hf: steps 16 hf - 2 .. 16 hf + 13 (hf = 0: 6 .. 13); one copy of the block loop
            if ((hf & 1) == 0 && hf > 0) { ctx.wait(hf >> 1); pf = ctx.frec(0); pb = ctx.brec(0); }
            const int nblk = (hf > 0) ? 16 / SCAN_BLK_OUT : 8 / SCAN_BLK_OUT;
#pragma unroll 1
            for (int b = 0; b < nblk; ++b) {
                Rec nf = load_rec(pf), nb = load_rec(pb);
#pragma unroll
                for (int q = 0; q < SPL; ++q) out_joint<0, WRITE>(rf, rb, th0[q], lam[q], cf[q], cb[q], F_[q], B_[q], g, Nl, Xw[q]);
                rf = load_rec(pf + REC); rb = load_rec(pb - REC);
#pragma unroll
                for (int q = 0; q < SPL; ++q) out_joint<1, WRITE>(nf, nb, th0[q], lam[q], cf[q], cb[q], F_[q], B_[q], g + 1, Nl, Xw[q]);
                if (SCAN_BLK_OUT == 4) {
                    nf = load_rec(pf + 2 * REC); nb = load_rec(pb - 2 * REC);
#pragma unroll
                    for (int q = 0; q < SPL; ++q) out_joint<0, WRITE>(rf, rb, th0[q], lam[q], cf[q], cb[q], F_[q], B_[q], g + 2, Nl, Xw[q]);
                    rf = load_rec(pf + 3 * REC); rb = load_rec(pb - 3 * REC);
#pragma unroll
                    for (int q = 0; q < SPL; ++q) out_joint<1, WRITE>(nf, nb, th0[q], lam[q], cf[q], cb[q], F_[q], B_[q], g + 3, Nl, Xw[q]);
                }
                pf += SCAN_BLK_OUT * REC; pb -= SCAN_BLK_OUT * REC;
                g += SCAN_BLK_OUT;
            }
            rescale_all(true, true);
            if ((hf & 1) == 0 && hf > 0) ctx.release((hf >> 1) - 1);
        }
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            // step gend - 2 (its successor's coefficients from the record already loaded), then step gend - 1 on its own
            out_joint<0, WRITE>(rf, rb, th0[q], lam[q], cf[q], cb[q], F_[q], B_[q], gend - 2, Nl, Xw[q]);
            out_chain_tail<+1, 1, WRITE>(F_[q], cf[q], gend - 1, Xw[q]);
            out_chain_tail<-1, 1, WRITE>(B_[q], cb[q], Nl - 1 - (gend - 1), Xw[q]);
        }
        rescale_all(true, true);
        ctx.release(nfs - 1);
        s_next = nfs;
        i_first = 0;
    }
    // ---- general steps (the first rows when there is no fast region, the ends of the sweeps, unequal sweeps)
    for (int s = s_next; s < nst; ++s) {
        if (s > 0) ctx.wait(s);
#pragma unroll 1
        for (int i = (s == s_next) ? i_first : 0; i < TR && TR * s + i <= qmax; ++i) {
            const int qq = TR * s + i;
            if (qq <= qf_end) {
                const Rec rf = load_rec(ctx.frec(i));
#pragma unroll
                for (int q = 0; q < SPL; ++q) out_step<+1, WRITE>(F_[q], rf, th0[q], lam[q], qq, Nl, false, Xw[q]);
            }
            if (qq <= qb_end) {
                const Rec rb = load_rec(ctx.brec(i));
#pragma unroll
                for (int q = 0; q < SPL; ++q) out_step<-1, WRITE>(B_[q], rb, th0[q], lam[q], qq, Nl, qq == qb_end, Xw[q]);
            }
            if ((i & (EBLK - 1)) == EBLK - 1) rescale_all(qq < qf_end, qq < qb_end);       // a finished sweep keeps its final scale
        }
        ctx.release(s);
    }
    if (WRITE) return;
    // ---- seam (rows k-1, k, k+1) and totals
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        const Sweep& f = F_[q]; const Sweep& b = B_[q];
        SolveOut& o = out[q];
        o.xkf = f.x; o.xkb = b.x; o.Ekf = f.E; o.Ekb = b.E;
        o.bad = f.bad | b.bad;
        const double zf = 1.0 / f.x, zb = 1.0 / b.x;
        // window values in the z scale: forward W1..W4 = rows k..k-3, backward W1..W4 = rows k..k+3
        const double Zm3 = f.W4 * zf, Zm2 = f.W3 * zf, Zm1 = f.W2 * zf, Zp1 = b.W2 * zb, Zp2 = b.W3 * zb, Zp3 = b.W4 * zb;
        const double Dm1 = fma(C23, 1.0 - Zm2, -(C12 * (Zp1 - Zm3)));
        const double D0 = fma(C23, Zp1 - Zm1, -(C12 * (Zp2 - Zm2)));
        const double Dp1 = fma(C23, Zp2 - 1.0, -(C12 * (Zp3 - Zm1)));
        const double w43 = 4.0 / 3.0, w23 = 2.0 / 3.0, w13 = 1.0 / 3.0;
        const double wk = (k & 1) ? w43 : w23, wk1 = (k & 1) ? w23 : w43;       // rows k and k +- 1
        const double zf2 = zf * zf, zb2 = zb * zb;
        const double sD = zf2 * (w43 * f.aDo + w23 * f.aDe + w13 * f.aEnd) + zb2 * (w43 * b.aDo + w23 * b.aDe + w13 * b.aEnd) +
                          wk1 * (f.gpp * Dm1 * Dm1 + b.gpp * Dp1 * Dp1) + wk * (f.gp * D0 * D0);
        const double sX0 = zf2 * (w43 * f.a0o + w23 * f.a0e) + zb2 * (w43 * b.a0o + w23 * b.a0e);
        const double sX1 = zf2 * (w43 * f.a1o + w23 * f.a1e) + zb2 * (w43 * b.a1o + w23 * b.a1e);
        o.gam = lam[q] + (sX0 - 2.0 * sD) / sX1;
        const double mf = f.vmax * fabs(zf), mb = b.vmax * fabs(zb);
        o.zmax = fmax(mf, mb);
        o.jmax = (mb > mf) ? b.jmax : f.jmax;
        if (!(o.zmax == o.zmax)) o.zmax = 1e308;          // NaN: treat as unusable
    }
}

// Finish the eigenfunction output of ONE solve with the rows dealt out over `nl` cooperating lanes (coalesced):
// zero_X: X = 0 (invalid input);  dX (if not null) from the normalised X by the reference's stencils (utils.py:1610-1614).
// ld/sync come from the context (the rows were written by other lanes).
template <class Ctx>
IBS_HD void fixup_solve(Ctx& ctx, double* X, double* dX, int N, bool zero_X, double h, int lane, int nl) {
    if (zero_X)
        for (int j = lane; j < N; j += nl) X[j] = 0.0;
    if (dX) {
        ctx.sync_mem();
        const double c23h = 2.0 / (3.0 * h), i12h = 1.0 / (12.0 * h), i2h = 1.0 / (2.0 * h), ih = 1.0 / h;
        constexpr int UD = 4;                  // rows per lane whose loads are in flight together
        for (int j0 = lane; j0 < N; j0 += UD * nl) {
            double xm2[UD], xm1[UD], xp1[UD], xp2[UD];
#pragma unroll
            for (int u = 0; u < UD; ++u) {
                const int j = j0 + u * nl;
                const bool in = j < N;
                xm2[u] = (in && j >= 2) ? ctx.ld(X + j - 2) : 0.0;
                xm1[u] = (in && j >= 1) ? ctx.ld(X + j - 1) : 0.0;
                xp1[u] = (in && j + 1 < N) ? ctx.ld(X + j + 1) : 0.0;
                xp2[u] = (in && j + 2 < N) ? ctx.ld(X + j + 2) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < UD; ++u) {
                const int j = j0 + u * nl;
                if (j < N) {
                    double d;
                    if (j == 0) d = (2.0 * xp1[u] - 0.5 * xp2[u]) * ih;                    // X_0 = 0 (utils.py:1610)
                    else if (j == 1) d = xp1[u] * i2h;                                       // (X_2 - X_0) / 2h
                    else if (j == N - 2) d = -xm1[u] * i2h;                                  // (X_{N-1} - X_{N-3}) / 2h
                    else if (j == N - 1) d = (0.5 * xm2[u] - 2.0 * xm1[u]) * ih;             // utils.py:1613
                    else d = c23h * (xp1[u] - xm1[u]) - (xp2[u] - xm2[u]) * i12h;
                    dX[j] = d;
                }
            }
        }
    }
}

// ---- bracketed Rayleigh-quotient iteration (per solve; the logic of ibs_solver.cu) -----------------------
struct Iter {
    double lam, rho, lo, hi, b1, N1, b2, N2, dprev;
    int nabove, it;
    bool collapsed, done, conv, warm;
    bool rq;        // the last update took the full Rayleigh-quotient step (lam = rho; in the basin, or from just above); dprev = its size
};

IBS_HD void iter_init(Iter& s, double l0, double Lb, double U, bool frozen) {
    s.lo = Lb; s.hi = U; s.b1 = s.N1 = s.b2 = s.N2 = 0.0; s.dprev = 1e300; s.nabove = 0; s.it = 0;
    s.collapsed = false; s.done = frozen; s.conv = frozen; s.rq = false;
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
    s.rq = false;
    if (pos) {
        const double dl = fabs(rho - lam);
        if (dl <= stop || (dl < tol_stag && dl >= 0.25 * s.dprev)) { s.conv = true; done = true; }
        // quadratic convergence: the error of rho is ~ dl^3 / dprev^2 once two consecutive corrections contract
        if (predictive && !done && dl < 1e-3 * s.dprev && s.dprev < 1e-2 * fmax(fabs(U), 1e-3)) {
            const double q = dl / s.dprev;
            if (dl * q * q <= 1.0 * tol) { s.conv = true; done = true; }
        }
        s.dprev = dl;
    }
    if (!done && s.collapsed) { s.conv = true; done = true; }
    if (!done) {
        double nxt;
        if (s.hi - s.lo <= tol) {
            nxt = 0.5 * (s.lo + s.hi);
            s.collapsed = true;
        } else if (inbasin && rho > lam && rho <= s.hi) {
            nxt = rho;
            s.rq = true;
        } else if (above) {
            double pw = 0.5;
            if (s.nabove >= 2 && s.N1 - s.N2 > 0.0) pw = (s.b1 - s.b2) / (s.N1 - s.N2);
            pw = fmin(1.0, fmax(0.4, pw));
            if (pw > 0.8) pw = 1.0;
            if (s.warm && s.nabove == 1) pw = 1.0;
            if (s.N2 < 0.05 * (U - s.b2)) pw = 1.0;
            nxt = s.b2 - pw * s.N2;
            if (!(nxt >= s.lo && nxt < s.hi)) nxt = 0.5 * (s.lo + s.hi);
            else if (pw == 1.0) s.rq = true;               // the full Rayleigh-quotient step, taken from above
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
    bool safe = false;      // lane-per-chain kernel: the prep kernel has certified every row of the line for the whole theta0 range
};
struct ItemResult { double gam, rho; int info; };

// The state of a lane's solves that no streaming loop touches (iteration bookkeeping, results of the output pass).
// It lives in memory the context provides (CUDA: shared memory, one block per thread with an odd stride in doubles, so
// that the lanes of a warp hit different banks) instead of competing for registers with the passes' loops.
template <int SPL>
struct ColdState {
    Iter it[SPL];
    SolveOut out[SPL];
    double rho1[SPL], rho2[SPL], rbest[SPL];
    double rho3[SPL], Cq[SPL];          // lane-per-chain kernel: third level for the extrapolation, quadratic-convergence constant
    int wi[SPL], wo[SPL];               // lane-per-chain kernel: tokens of the warm-start records read / written (-1: none)
    int nev[SPL], flags[SPL];
    bool fin[SPL], wr[SPL], need[SPL];
};
template <int SPL> struct ColdStride { static constexpr int value = (int)((sizeof(ColdState<SPL>) + 7) / 8) | 1; };     // doubles, odd

// Everything for the SPL solves of one lane.  act[q] = false: the slot duplicates a valid solve and writes nothing.
// Written as a state machine with ONE call site per kind of pass (iteration / output pass), so that the kernel holds
// a single copy of each streaming loop (instruction cache).
//   ITER   bracketed Rayleigh-quotient iteration on level `lev`, coarsest first; start = Richardson estimate of the
//          two coarser levels' eigenvalues
//   PEAK   (coarsest level only) first output pass there: the lanes' eigenfunction peaks give the matching row
//   O1     fine level: Simpson Rayleigh quotient, max|z|, validity;  solves with |z_k| << max|z| are re-converged
//          with the matching row moved to their peak (at most three times) before they are accepted
//          then X (stored raw by O1) is normalised and dX formed by the context's fix-up (coalesced over the rows)
//   SIGMA  one count at 2 sigma - lambda (utils.py:1597 returns the eigenvalue nearest sigma; the engine lambda_max)
enum { PH_ITER = 0, PH_PEAK = 1, PH_O1 = 2, PH_O2 = 3, PH_SIGMA = 4 };
// MODE_FULL: everything.  The two-kernel form splits it where the register needs differ (the iteration needs about half
// of what the output passes need): MODE_ITER = the level iterations only (matching row = middle row; res[].rho / .info
// carry the converged shift and the counters out), MODE_OUT = output passes, fallback and sigma check, started from the
// res[].rho / .info of a MODE_ITER run.
enum { MODE_FULL = 0, MODE_ITER = 1, MODE_OUT = 2 };

// The output passes need about twice the registers of the iteration pass, so with two solves per lane they are run one
// solve at a time (the iteration keeps both in flight: twice the instruction-level parallelism, half the record loads).
template <int SPL, bool WRITE, class Ctx>
IBS_HD void out_pass_each(Ctx& ctx, int lev, int Nl, int k, const double (&th0)[SPL], const double (&lam)[SPL],
                          SolveOut (&out)[SPL], double* const (&Xw)[SPL]) {
    if (SPL == 1) {
        out_pass<SPL, WRITE>(ctx, lev, Nl, k, th0, lam, out, Xw);
    } else {
#pragma unroll 1
        for (int q = 0; q < SPL; ++q) {
            const double th1[1] = {q ? th0[SPL - 1] : th0[0]};
            const double lam1[1] = {q ? lam[SPL - 1] : lam[0]};
            double* const Xw1[1] = {q ? Xw[SPL - 1] : Xw[0]};
            SolveOut o1[1];
            if (WRITE) o1[0] = out[q];
            out_pass<1, WRITE>(ctx, lev, Nl, k, th1, lam1, o1, Xw1);
            if (!WRITE) out[q] = o1[0];
        }
    }
}

template <int SPL, int MODE, class Ctx>
IBS_HD void solve_item(Ctx& ctx, const ItemProblem& P, const double (&th0)[SPL], const bool (&act)[SPL], const double (&sigma)[SPL],
                       const bool has_sigma, double* const (&Xrow)[SPL], double* const (&dXrow)[SPL], ItemResult (&res)[SPL],
                       ColdState<SPL>& cs) {
    // Xrow[q]: the solve's X row (needed, as scratch, also when only dX is wanted)
    const double scale = fmax(fabs(P.U), 1e-3);
    const double tol = 1.7763568394002505e-15 * scale, tol_stag = 1e-10 * scale;
    const double qnan = NAN;
    const int N = P.N;
    Iter (&it)[SPL] = cs.it;
    SolveOut (&out)[SPL] = cs.out;
    double (&rho1)[SPL] = cs.rho1; double (&rho2)[SPL] = cs.rho2;
    double (&rbest)[SPL] = cs.rbest;                    // best eigenvalue estimate of the matrix pencil
    int (&nev)[SPL] = cs.nev; int (&flags)[SPL] = cs.flags;
    bool (&fin)[SPL] = cs.fin; bool (&wr)[SPL] = cs.wr; bool (&need)[SPL] = cs.need;
    double sh[SPL], r[SPL], S[SPL];                     // sh: the shifts of the next pass
    int nodes[SPL];
    double* Xraw[SPL];
    const bool want_out = P.want_X || P.want_dX;
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        Xraw[q] = nullptr;
        rho1[q] = qnan; rho2[q] = qnan; nev[q] = 0; flags[q] = 0; fin[q] = false; wr[q] = false; need[q] = false; rbest[q] = qnan;
        iter_init(it[q], qnan, P.Lb, P.U, false);
        sh[q] = it[q].lam;
        if (MODE == MODE_OUT) {          // continue a MODE_ITER run
            sh[q] = res[q].rho; rbest[q] = res[q].rho;
            flags[q] = res[q].info >> 16; nev[q] = (res[q].info & 0xffff) << MAXLEV;
        }
        res[q].gam = qnan; res[q].rho = qnan;
    }
    int lev = (MODE == MODE_OUT) ? 0 : P.nlev, Nl = level_n(N, lev), k = clamp_k((Nl - 1) / 2, Nl), round = 0;
    int phase = (MODE == MODE_OUT) ? PH_O1 : PH_ITER;
    bool lowq_any = false;
    int jsel = -1;
    bool fix_any = false;
    for (;;) {
        // ---- the streaming pass of this phase: ONE call site per kind of pass
        const int kind = (phase == PH_ITER || phase == PH_SIGMA) ? 1 : (phase == PH_PEAK || phase == PH_O1) ? 2 : 3;
        if (kind == 1 || MODE == MODE_ITER) eval_pass<SPL>(ctx, lev, Nl, k, th0, sh, r, S, nodes);
        else if (kind == 2) out_pass_each<SPL, false>(ctx, lev, Nl, k, th0, sh, out, Xraw);
        else out_pass_each<SPL, true>(ctx, lev, Nl, k, th0, sh, out, Xraw);
        // ---- what the phase does with it
        if (phase == PH_ITER || phase == PH_SIGMA) {
            if (phase == PH_SIGMA) {
#pragma unroll
                for (int q = 0; q < SPL; ++q)
                    if (need[q] && nodes[q] + (r[q] > 0.0 ? 1 : 0) > 1) flags[q] |= FLAG_SIGMA_NOT_MAX;
                break;
            }
            bool alldone = true;
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                if (!it[q].done) nev[q] += (1 << MAXLEV) >> lev;     // in units of the coarsest level's share of a fine-grid evaluation
                iter_update(it[q], r[q], S[q], nodes[q], P.U, tol, tol_stag, (lev > 0) ? 1e-7 * scale : tol, lev == 0);
                alldone &= it[q].done;
                sh[q] = it[q].lam;
            }
            if (!ctx.all(alldone)) continue;
            // ---- this level has converged
            if (lev > 0) {
#pragma unroll
                for (int q = 0; q < SPL; ++q) { rho2[q] = rho1[q]; rho1[q] = (it[q].conv && it[q].rho == it[q].rho) ? it[q].rho : qnan; }
                if (MODE != MODE_ITER && lev == P.nlev) {
#pragma unroll
                    for (int q = 0; q < SPL; ++q) sh[q] = (rho1[q] == rho1[q]) ? rho1[q] : it[q].lam;
                    phase = PH_PEAK;
                    continue;
                }
            } else {
#pragma unroll
                for (int q = 0; q < SPL; ++q)
                    if (!fin[q]) {
                        rbest[q] = (it[q].conv && it[q].rho == it[q].rho) ? it[q].rho : it[q].lam;
                        sh[q] = rbest[q];
                        if (!it[q].conv) flags[q] |= FLAG_NOT_CONVERGED;
                    }
                if (MODE == MODE_ITER) {
#pragma unroll
                    for (int q = 0; q < SPL; ++q) res[q].rho = rbest[q];
                    break;
                }
                phase = PH_O1;
                continue;
            }
        } else if (phase == PH_PEAK) {
            // matching row from the coarsest eigenfunctions: the middle of the range of the lanes' peaks -- unless that is
            // close to the middle row, which keeps the two chains equally long (everything on the fast path)
            int jlo = 1 << 30, jhi = -1;
#pragma unroll
            for (int q = 0; q < SPL; ++q) { jlo = imin(jlo, out[q].jmax); jhi = imax(jhi, out[q].jmax); }
            jlo = ctx.min_i(jlo); jhi = ctx.max_i(jhi);
            const int kp = (jlo + jhi) / 2, km = (Nl - 1) / 2;
            k = clamp_k((kp > km ? kp - km : km - kp) * PEAK_SNAP <= Nl ? km : kp, Nl);
        } else if (phase == PH_O1) {
            bool wr_any = false;
            fix_any = false;
            lowq_any = false; jsel = -1;
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                const bool lowq = !fin[q] && !out[q].bad && out[q].zmax > LOWQ && round < 3;
                const bool newly = !fin[q] && !lowq;
                if (newly) {
                    fin[q] = true;
                    res[q].gam = out[q].bad ? qnan : out[q].gam;
                    res[q].rho = out[q].bad ? qnan : rbest[q];
                    if (out[q].bad) flags[q] = FLAG_BAD_INPUT;
                }
                if (lowq && jsel < 0) jsel = out[q].jmax;
                lowq_any |= lowq;
                wr[q] = newly && act[q];
                wr_any |= wr[q];
                fix_any |= wr[q] && (out[q].bad || P.want_dX);
            }
            ++round;
            if (want_out && ctx.any(wr_any)) {
                // second pass: X of the solves accepted in this round (invalid ones are zero-filled by the fix-up)
#pragma unroll
                for (int q = 0; q < SPL; ++q) Xraw[q] = (wr[q] && !out[q].bad) ? Xrow[q] : nullptr;
                phase = PH_O2;
                continue;
            }
        } else {       // PH_O2
#pragma unroll
            for (int q = 0; q < SPL; ++q) Xraw[q] = nullptr;
            if (ctx.any(fix_any)) ctx.template fixup<SPL>(wr, Xrow, dXrow, N, out, P.h, P.want_dX);
        }
        // ---- what follows a finished level (lev > 0), the peak finder, or the output passes of a round
        if (lev > 0) {
            --lev;
            Nl = level_n(N, lev);
            k = clamp_k(2 * k, Nl);
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                double l0 = rho1[q];
                if (rho2[q] == rho2[q]) l0 = rho1[q] - 0.25 * (rho2[q] - rho1[q]);      // Richardson: the error is ~ h^2
                iter_init(it[q], l0, P.Lb, P.U, false);
                sh[q] = it[q].lam;
            }
            phase = PH_ITER;
            continue;
        }
        if (ctx.any(lowq_any)) {
            // move the matching row to the peak of the first low-quality solve and re-converge the open solves there
            k = clamp_k(ctx.first_i(jsel), N);
            lowq_any = false;
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                iter_init(it[q], sh[q], P.Lb, P.U, fin[q]);
                if (!fin[q]) sh[q] = it[q].lam;
            }
            phase = PH_ITER;
            continue;
        }
        if (!has_sigma) break;
        bool any_need = false;
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            need[q] = (flags[q] & FLAG_BAD_INPUT) == 0 && sigma[q] < rbest[q];
            if (need[q]) sh[q] = 2.0 * sigma[q] - rbest[q];
            any_need |= need[q];
        }
        if (!ctx.any(any_need)) break;
        phase = PH_SIGMA;
    }
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        // iterations reported = fine-grid-equivalent evaluations of the iteration (coarse levels count by their size)
        const int itc = (flags[q] & FLAG_BAD_INPUT) ? 0 : ((flags[q] & FLAG_NOT_CONVERGED) ? 64 : imin((nev[q] + (1 << MAXLEV) - 1) >> MAXLEV, 63));
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
