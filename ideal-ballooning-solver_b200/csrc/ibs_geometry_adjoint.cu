// Reverse mode of the field-line geometry: d(lambda) / d(Fourier table coefficient)  (SURVEY.md section 8 row f3).
//
// The reference obtains the Jacobian of the ballooning penalty by finite differences over ndofs + 1 perturbed equilibria,
// each with a full scan (/root/reference/sims_runner_NCSX.py:245-262).  With the eigen-pair of the base equilibrium in hand
// the Hellmann-Feynman formula the reference itself uses for its (alpha, theta0) gradients (utils.py:1676-1680) gives the
// first-order change of lambda for ANY change of the coefficient arrays,
//     d lam = sum_j ( s^g_j dg_j + s^c_j dc_j + s^f_j df_j ),        s^g, s^c, s^f from ibs_adjoint_sensitivities,
// and g, c, f are explicit functions of the 19 mode sums of each point (utils.py:432-720) and of dPdrho (a mean over the line,
// ball_scan.py:262).  This file chains that back to the tables:
//   stage A (one thread per point):  theta_vmec (Newton), the 19 mode sums S_r and their theta_vmec-derivatives; then
//       v_r = d Phi_j / d S_r   with   Phi_j = s^g g - dP s^c chat + s^f f + (Q / 2N) (cvdrift - gbdrift) bmag^2
//       (chat = c / (-dP), Q = sum_j s^c_j chat_j: the dPdrho term folded back onto the points), by FORWARD-mode automatic
//       differentiation of the pointwise algebra (a value + one derivative, 19 sweeps: no hand-derived adjoint to get wrong),
//       and the weight u = -(d Phi / d theta_vmec) / (1 + d lambda / d theta_vmec) of the implicit theta_vmec(l_mn) dependence;
//   stage B (one thread per (surface, mode)):  the transposed mode sum over the points of the line,
//       d lam / d rmnc[mn] = sum_j ( v_R cos a - m v_Rt sin a + n v_Rp sin a ),  a = m theta_vmec - n phi,   etc.
// Gradients are with respect to the per-surface tables tab_mn (6 rows) and tab_nyq (7 rows) at fixed scalars (iota, shat,
// p', iota'): the same perturbations scan.hellmann_feynman_gamma takes.  Not a hot kernel: ns field lines (each surface's
// arg-max line), direct mode sums.
#include "ibs_common.cuh"

namespace ibs {

// ---- forward-mode scalar: value + one directional derivative -----------------------------------------------------------
struct D1 {
    double v, d;
    __device__ __forceinline__ D1() : v(0.0), d(0.0) {}
    __device__ __forceinline__ D1(double v_) : v(v_), d(0.0) {}
    __device__ __forceinline__ D1(double v_, double d_) : v(v_), d(d_) {}
};
__device__ __forceinline__ D1 operator+(D1 a, D1 b) { return {a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ D1 operator-(D1 a, D1 b) { return {a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ D1 operator-(D1 a) { return {-a.v, -a.d}; }
__device__ __forceinline__ D1 operator*(D1 a, D1 b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ D1 operator/(D1 a, D1 b) { const double q = a.v / b.v; return {q, (a.d - q * b.d) / b.v}; }
__device__ __forceinline__ D1 dabs(D1 a) { return a.v < 0.0 ? D1{-a.v, -a.d} : a; }

constexpr int NSUM = 19;
enum { S_R, S_Rs, S_Rt, S_Rp, S_Zs, S_Zt, S_Zp, S_Ls, S_Lt, S_Lp, S_sqrtg, S_B, S_Bs, S_Bt, S_Bp, S_Bsupp, S_Bsubs, S_Bsubt, S_Bsubp };

struct PointConst {
    double sp, cp, dphi;                  // sin / cos(phi), phi - phi_center
    double iota, d_iota, dpds, shat, s_val, psi_e, L_ref;
    double th0, dP, sg, sc, sf, qw;       // theta0, dPdrho, the three sensitivities of this point, Q / (2 N)
};

// Phi_j as a function of the 19 mode sums (the pointwise algebra of geo_point_epilogue, ibs_geometry.cu, on T = D1)
__device__ D1 phi_point(const D1 (&S)[NSUM], const PointConst& k) {
    const D1 R = S[S_R], R_s = S[S_Rs], R_t = S[S_Rt], R_p = S[S_Rp], Z_s = S[S_Zs], Z_t = S[S_Zt], Z_p = S[S_Zp];
    const D1 L_s = S[S_Ls], L_t = S[S_Lt], L_p = S[S_Lp];
    const D1 sqrtg = S[S_sqrtg], B = S[S_B], B_s = S[S_Bs], B_t = S[S_Bt], B_p = S[S_Bp], Bsup_p = S[S_Bsupp];
    const D1 Bsub_s = S[S_Bsubs], Bsub_t = S[S_Bsubt], Bsub_p = S[S_Bsubp];
    const D1 sp = k.sp, cp = k.cp;
    const D1 X_t = R_t * cp, X_p = R_p * cp - R * sp, X_s = R_s * cp;
    const D1 Y_t = R_t * sp, Y_p = R_p * sp + R * cp, Y_s = R_s * sp;
    const D1 isg = D1(1.0) / sqrtg;
    const D1 gsx = (Y_t * Z_p - Z_t * Y_p) * isg, gsy = (Z_t * X_p - X_t * Z_p) * isg, gsz = (X_t * Y_p - Y_t * X_p) * isg;
    const D1 gtx = (Y_p * Z_s - Z_p * Y_s) * isg, gty = (Z_p * X_s - X_p * Z_s) * isg, gtz = (X_p * Y_s - Y_p * X_s) * isg;
    const D1 gpx = (Y_s * Z_t - Z_s * Y_t) * isg, gpy = (Z_s * X_t - X_s * Z_t) * isg, gpz = (X_s * Y_t - Y_s * X_t) * isg;
    const D1 psi_e = k.psi_e;
    const D1 a_s = L_s - D1(k.dphi * k.d_iota);
    const D1 a_t = D1(1.0) + L_t, a_p = L_p - D1(k.iota);
    const D1 gax = a_s * gsx + (a_t * gtx + a_p * gpx), gay = a_s * gsy + (a_t * gty + a_p * gpy), gaz = a_s * gsz + (a_t * gtz + a_p * gpz);
    const D1 gqx = gsx * psi_e, gqy = gsy * psi_e, gqz = gsz * psi_e;
    const D1 lpi = a_p;
    const D1 BxgB_ga = (Bsub_s * B_t * lpi + Bsub_t * B_p * a_s + Bsub_p * B_s * a_t - Bsub_p * B_t * a_s - Bsub_t * B_s * lpi -
                        Bsub_s * B_p * a_t) * isg;
    const D1 ga_ga = gax * gax + gay * gay + gaz * gaz, ga_gq = gax * gqx + gay * gqy + gaz * gqz, gq_gq = gqx * gqx + gqy * gqy + gqz * gqz;
    const D1 BxgB_gq = (Bsub_t * B_p - Bsub_p * B_t) * isg * psi_e;
    const double L_ref = k.L_ref, B_ref = 2.0 * fabs(k.psi_e) / (L_ref * L_ref);
    const double sgn = (k.psi_e > 0.0) ? 1.0 : ((k.psi_e < 0.0) ? -1.0 : 0.0);
    const double sqrt_s = sqrt(k.s_val), mu0 = 4.0 * 3.141592653589793 * 1.0e-7;
    const D1 B3 = B * B * B;
    const D1 bmag = B / D1(B_ref);
    const D1 gradpar = D1(L_ref * k.iota) * Bsup_p / B;
    const D1 gds2 = ga_ga * D1(L_ref * L_ref * k.s_val);
    const D1 gds21 = ga_gq * D1(k.shat / B_ref);
    const D1 gds22 = gq_gq * D1(k.shat * k.shat / (L_ref * L_ref * B_ref * B_ref * k.s_val));
    const D1 gbdrift = D1(-2.0 * B_ref * L_ref * L_ref * sqrt_s * sgn) * BxgB_ga / B3;
    const D1 gbdrift0 = D1(-2.0 * k.shat * sgn / sqrt_s) * BxgB_gq / B3;
    const D1 cvdrift = gbdrift - D1(2.0 * B_ref * L_ref * L_ref * sqrt_s * mu0 * k.dpds * sgn / k.psi_e) / (B * B);
    // theta0 shift and the coefficients (ball_scan.py:267-268, utils.py:1560-1562)
    const D1 cv_f = cvdrift + D1(k.th0) * gbdrift0;
    const D1 gd_f = gds2 + D1(2.0 * k.th0) * gds21 + D1(k.th0 * k.th0) * gds22;
    const D1 gp = dabs(gradpar);
    const D1 g = gp * gd_f / bmag;
    const D1 chat = cv_f / (gp * bmag);                        // c = -dPdrho chat
    const D1 f = gd_f / (bmag * bmag * bmag * gp);
    return D1(k.sg) * g - D1(k.dP * k.sc) * chat + D1(k.sf) * f + D1(k.qw) * (cvdrift - gbdrift) * bmag * bmag;
}

struct GeoAdjParams {
    const double* tab_mn; const double* tab_nyq; const double* scal;
    const double* xm; const double* xn; const double* xm_nyq; const double* xn_nyq;
    int ns, mnmax, mnmax_nyq, nl;
    const double* alpha; const double* theta; const double* theta0; const double* dPdrho;
    const double* sg; const double* sc; const double* sf; const double* Q;
    double phi_center, psi_e, L_ref;
    double* work;          // [ns][NSUM + 3][nl]: v_r (19), u, theta_vmec, phi
    double* grad_mn; double* grad_nyq;
};

__global__ void __launch_bounds__(128)
geo_adjoint_point_kernel(const GeoAdjParams p) {
    const long long pt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= (long long)p.ns * p.nl) return;
    const int js = (int)(pt / p.nl), jl = (int)(pt - (long long)js * p.nl);
    const double* sc = p.scal + (size_t)js * IBS_NSCAL;
    PointConst k;
    k.s_val = sc[0]; k.iota = sc[1]; k.d_iota = sc[2]; k.dpds = sc[3]; k.shat = sc[4];
    k.psi_e = p.psi_e; k.L_ref = p.L_ref;
    const double theta_p = p.theta[jl], phi = p.phi_center + (theta_p - p.alpha[js]) / k.iota;
    sincos(phi, &k.sp, &k.cp);
    k.dphi = phi - p.phi_center;
    const size_t o = (size_t)js * p.nl + jl;
    k.th0 = p.theta0[js]; k.dP = p.dPdrho[js]; k.sg = p.sg[o]; k.sc = p.sc[o]; k.sf = p.sf[o]; k.qw = p.Q[js] / (2.0 * p.nl);
    const int mn = p.mnmax, mq = p.mnmax_nyq;
    const double* tm = p.tab_mn + (size_t)js * IBS_TAB_MN_ROWS * mn;
    const double* tq = p.tab_nyq + (size_t)js * IBS_TAB_NYQ_ROWS * mq;
    // theta_vmec (utils.py:391-416), Newton with the analytic derivative
    double th = theta_p;
    for (int it = 0; it < 50; ++it) {
        double f = th - theta_p, d = 1.0;
        for (int q = 0; q < mn; ++q) {
            double sa, ca;
            sincos(p.xm[q] * th - p.xn[q] * phi, &sa, &ca);
            f = fma(tm[2 * mn + q], sa, f);
            d = fma(tm[2 * mn + q] * p.xm[q], ca, d);
        }
        const double dth = f / d;
        th -= dth;
        if (fabs(dth) <= 4.5e-16 * fmax(1.0, fabs(th))) break;
    }
    // the 19 sums and their theta_vmec-derivatives
    double S[NSUM], T[NSUM];
    for (int r = 0; r < NSUM; ++r) { S[r] = 0.0; T[r] = 0.0; }
    for (int q = 0; q < mn; ++q) {
        const double m = p.xm[q], n = p.xn[q];
        double sa, ca;
        sincos(m * th - n * phi, &sa, &ca);
        const double r = tm[q], z = tm[mn + q], l = tm[2 * mn + q], rs = tm[3 * mn + q], zs = tm[4 * mn + q], ls = tm[5 * mn + q];
        S[S_R] += r * ca;        T[S_R] -= r * m * sa;
        S[S_Rs] += rs * ca;      T[S_Rs] -= rs * m * sa;
        S[S_Rt] -= r * m * sa;   T[S_Rt] -= r * m * m * ca;
        S[S_Rp] += r * n * sa;   T[S_Rp] += r * n * m * ca;
        S[S_Zs] += zs * sa;      T[S_Zs] += zs * m * ca;
        S[S_Zt] += z * m * ca;   T[S_Zt] -= z * m * m * sa;
        S[S_Zp] -= z * n * ca;   T[S_Zp] += z * n * m * sa;
        S[S_Ls] += ls * sa;      T[S_Ls] += ls * m * ca;
        S[S_Lt] += l * m * ca;   T[S_Lt] -= l * m * m * sa;
        S[S_Lp] -= l * n * ca;   T[S_Lp] += l * n * m * sa;
    }
    for (int q = 0; q < mq; ++q) {
        const double m = p.xm_nyq[q], n = p.xn_nyq[q];
        double sa, ca;
        sincos(m * th - n * phi, &sa, &ca);
        const double g = tq[q], b = tq[mq + q], bs = tq[2 * mq + q], bv = tq[3 * mq + q], s_ = tq[4 * mq + q], u_ = tq[5 * mq + q], v_ = tq[6 * mq + q];
        S[S_sqrtg] += g * ca;    T[S_sqrtg] -= g * m * sa;
        S[S_B] += b * ca;        T[S_B] -= b * m * sa;
        S[S_Bs] += bs * ca;      T[S_Bs] -= bs * m * sa;
        S[S_Bt] -= b * m * sa;   T[S_Bt] -= b * m * m * ca;
        S[S_Bp] += b * n * sa;   T[S_Bp] += b * n * m * ca;
        S[S_Bsupp] += bv * ca;   T[S_Bsupp] -= bv * m * sa;
        S[S_Bsubs] += s_ * sa;   T[S_Bsubs] += s_ * m * ca;
        S[S_Bsubt] += u_ * ca;   T[S_Bsubt] -= u_ * m * sa;
        S[S_Bsubp] += v_ * ca;   T[S_Bsubp] -= v_ * m * sa;
    }
    // v_r = d Phi / d S_r by 19 forward sweeps; dPhi/dtheta_vmec = sum_r v_r T_r
    double* w = p.work + (size_t)js * (NSUM + 3) * p.nl + jl;
    double dth = 0.0;
    for (int r = 0; r < NSUM; ++r) {
        D1 Sd[NSUM];
        for (int q = 0; q < NSUM; ++q) Sd[q] = D1(S[q], q == r ? 1.0 : 0.0);
        const double v = phi_point(Sd, k).d;
        w[(size_t)r * p.nl] = v;
        dth = fma(v, T[r], dth);
    }
    w[(size_t)NSUM * p.nl] = -dth / (1.0 + S[S_Lt]);          // weight of d(lmns) sin a (implicit theta_vmec)
    w[(size_t)(NSUM + 1) * p.nl] = th;
    w[(size_t)(NSUM + 2) * p.nl] = phi;
}

// stage B: one thread per (surface, mode); modes 0 .. mnmax-1 of the (R, Z, lambda) set, then the Nyquist set
__global__ void __launch_bounds__(128)
geo_adjoint_mode_kernel(const GeoAdjParams p) {
    const int js = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int mn = p.mnmax, mq = p.mnmax_nyq;
    if (q >= mn + mq) return;
    const bool nyq = q >= mn;
    const int kq = nyq ? q - mn : q;
    const double m = nyq ? p.xm_nyq[kq] : p.xm[kq], n = nyq ? p.xn_nyq[kq] : p.xn[kq];
    const double* w = p.work + (size_t)js * (NSUM + 3) * p.nl;
    const double* th = w + (size_t)(NSUM + 1) * p.nl;
    const double* ph = w + (size_t)(NSUM + 2) * p.nl;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0;
    for (int j = 0; j < p.nl; ++j) {
        double sa, ca;
        sincos(m * th[j] - n * ph[j], &sa, &ca);
#define V(r) w[(size_t)(r) * p.nl + j]
        if (!nyq) {
            a0 += V(S_R) * ca - m * V(S_Rt) * sa + n * V(S_Rp) * sa;                 // rmnc
            a1 += m * V(S_Zt) * ca - n * V(S_Zp) * ca;                                // zmns
            a2 += m * V(S_Lt) * ca - n * V(S_Lp) * ca + V(NSUM) * sa;                 // lmns (explicit + through theta_vmec)
            a3 += V(S_Rs) * ca;                                                       // d_rmnc_d_s
            a4 += V(S_Zs) * sa;                                                       // d_zmns_d_s
            a5 += V(S_Ls) * sa;                                                       // d_lmns_d_s
        } else {
            a0 += V(S_sqrtg) * ca;                                                    // gmnc
            a1 += V(S_B) * ca - m * V(S_Bt) * sa + n * V(S_Bp) * sa;                  // bmnc
            a2 += V(S_Bs) * ca;                                                       // d_bmnc_d_s
            a3 += V(S_Bsupp) * ca;                                                    // bsupvmnc
            a4 += V(S_Bsubs) * sa;                                                    // bsubsmns
            a5 += V(S_Bsubt) * ca;                                                    // bsubumnc
            a6 += V(S_Bsubp) * ca;                                                    // bsubvmnc
        }
#undef V
    }
    if (!nyq) {
        double* g = p.grad_mn + (size_t)js * IBS_TAB_MN_ROWS * mn + kq;
        g[0] = a0; g[mn] = a1; g[2 * mn] = a2; g[3 * mn] = a3; g[4 * mn] = a4; g[5 * mn] = a5;
    } else {
        double* g = p.grad_nyq + (size_t)js * IBS_TAB_NYQ_ROWS * mq + kq;
        g[0] = a0; g[mq] = a1; g[2 * mq] = a2; g[3 * mq] = a3; g[4 * mq] = a4; g[5 * mq] = a5; g[6 * mq] = a6;
    }
}

int geometry_adjoint_dispatch(const double* tab_mn, const double* tab_nyq, const double* scal, const double* xm, const double* xn,
                              const double* xm_nyq, const double* xn_nyq, int ns, int mnmax, int mnmax_nyq, double phiedge,
                              double aminor_p, const double* alpha, const double* theta, int nl, double phi_center,
                              const double* theta0, const double* dPdrho, const double* sg, const double* sc, const double* sf,
                              const double* Q, double* grad_mn, double* grad_nyq, cudaStream_t st) {
    if (ns == 0) return IBS_OK;
    GeoAdjParams p;
    p.tab_mn = tab_mn; p.tab_nyq = tab_nyq; p.scal = scal; p.xm = xm; p.xn = xn; p.xm_nyq = xm_nyq; p.xn_nyq = xn_nyq;
    p.ns = ns; p.mnmax = mnmax; p.mnmax_nyq = mnmax_nyq; p.nl = nl; p.alpha = alpha; p.theta = theta; p.theta0 = theta0; p.dPdrho = dPdrho;
    p.sg = sg; p.sc = sc; p.sf = sf; p.Q = Q; p.phi_center = phi_center; p.psi_e = -phiedge / (2.0 * 3.141592653589793); p.L_ref = aminor_p;
    p.grad_mn = grad_mn; p.grad_nyq = grad_nyq;
    keep_pool_cached();
    IBS_CUDA_CHECK(cudaMallocAsync((void**)&p.work, (size_t)ns * (NSUM + 3) * nl * sizeof(double), st));
    const long long npt = (long long)ns * nl;
    geo_adjoint_point_kernel<<<(unsigned)((npt + 127) / 128), 128, 0, st>>>(p);
    int rc = (cudaGetLastError() == cudaSuccess) ? IBS_OK : IBS_ERR_CUDA;
    if (rc == IBS_OK) {
        geo_adjoint_mode_kernel<<<dim3((mnmax + mnmax_nyq + 127) / 128, ns), 128, 0, st>>>(p);
        if (cudaGetLastError() != cudaSuccess) rc = IBS_ERR_CUDA;
    }
    if (rc != IBS_OK) set_error("geometry adjoint kernel launch failed");
    cudaFreeAsync(p.work, st);
    return rc;
}

}  // namespace ibs
