// Full-output field-line geometry: every per-point array of the Struct that vmec_fieldlines returns
// (/root/reference/utils.py:161-864, list of fields at :723-864), plus what vmec_fieldlines_axisym adds (:872-1542).
//
// This is the API-completeness companion of K1 (ibs_geometry.cu), not a hot kernel: the ballooning path reads eight of
// these arrays and K1 produces exactly those with separable recurrences; gyrokinetic-geometry consumers of the reference
// want the other ~70 too.  One thread = one point; the mode sums are evaluated directly (one
// sincos per mode and point), so any mode list is accepted (no dense (m, n) packing, no limit on |n| / nfp or mpol).
//   mode 0: `grid` is theta_pest; phi = phi_center + (theta_pest - alpha) / iota          (utils.py:371-373), root solve for theta_vmec
//   mode 1: `grid` is phi;        theta_pest = alpha + iota (phi - phi_center)             (utils.py:364-369), root solve for theta_vmec
//   mode 2: `grid` is theta_vmec (axisymmetric routine: uniform theta_vmec, no root solve, utils.py:972-978); the mode sums use
//           theta_vmec + theta_shift (pi when the poloidal angle has to be flipped, :993-1010) and phi = 0; then
//           theta_pest = theta_vmec + lambda and phi = phi_center + (theta_pest - alpha) / iota (:1043-1046)
#include "ibs_common.cuh"

namespace ibs {

// Field order of out[line][NF][nl]; mirrored by FULL_FIELDS in reference_api.py (checked by tests/test_abi.py)
enum FullField {
    FF_phi, FF_theta_pest, FF_theta_vmec, FF_lambda,
    FF_R, FF_d_R_d_s, FF_d_R_d_theta_vmec, FF_d_R_d_phi, FF_Z, FF_d_Z_d_s, FF_d_Z_d_theta_vmec, FF_d_Z_d_phi,
    FF_d_lambda_d_s, FF_d_lambda_d_theta_vmec, FF_d_lambda_d_phi,
    FF_sqrt_g_vmec, FF_modB, FF_d_B_d_s, FF_d_B_d_theta_vmec, FF_d_B_d_phi, FF_B_sup_theta_vmec, FF_B_sup_phi,
    FF_B_sub_s, FF_B_sub_theta_vmec, FF_B_sub_phi, FF_B_sup_theta_pest, FF_sqrt_g_vmec_alt, FF_sinphi, FF_cosphi,
    FF_d_X_d_theta_vmec, FF_d_X_d_phi, FF_d_X_d_s, FF_d_Y_d_theta_vmec, FF_d_Y_d_phi, FF_d_Y_d_s,
    FF_grad_s_X, FF_grad_s_Y, FF_grad_s_Z, FF_grad_theta_vmec_X, FF_grad_theta_vmec_Y, FF_grad_theta_vmec_Z,
    FF_grad_phi_X, FF_grad_phi_Y, FF_grad_phi_Z, FF_grad_psi_X, FF_grad_psi_Y, FF_grad_psi_Z,
    FF_grad_alpha_X, FF_grad_alpha_Y, FF_grad_alpha_Z, FF_grad_B_X, FF_grad_B_Y, FF_grad_B_Z, FF_B_X, FF_B_Y, FF_B_Z,
    FF_B_cross_grad_s_dot_grad_alpha, FF_B_cross_grad_s_dot_grad_alpha_alternate,
    FF_B_cross_grad_B_dot_grad_alpha, FF_B_cross_grad_B_dot_grad_alpha_alternate,
    FF_B_cross_grad_B_dot_grad_psi, FF_B_cross_kappa_dot_grad_psi, FF_B_cross_kappa_dot_grad_alpha,
    FF_grad_alpha_dot_grad_alpha, FF_grad_alpha_dot_grad_psi, FF_grad_psi_dot_grad_psi,
    FF_bmag, FF_gradpar_theta_pest, FF_gradpar_phi, FF_gds2, FF_gds21, FF_gds22, FF_gbdrift, FF_gbdrift0, FF_cvdrift, FF_cvdrift0,
    FF_B_p,
    FF_COUNT
};

struct FullParams {
    const double* tab_mn; const double* tab_nyq; const double* bsupu; const double* scal;
    const double* xm; const double* xn; const double* xm_nyq; const double* xn_nyq;
    int ns, mnmax, mnmax_nyq;
    const double* alpha; int nalpha; const double* grid; int nl; int mode;
    double theta_shift, phi_center, psi_e, L_ref;
    int zero_xn_nyq;
    double* out; int* info;
};

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 cross(const V3& a, const V3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 scale(const V3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 add3(const V3& a, const V3& b, const V3& c) { return {a.x + (b.x + c.x), a.y + (b.y + c.y), a.z + (b.z + c.z)}; }
// a . (b x c) written out like the reference's "_alternate" triple products (utils.py:593-601)
__device__ __forceinline__ double triple(const V3& a, const V3& b, const V3& c) {
    return a.x * b.y * c.z + a.y * b.z * c.x + a.z * b.x * c.y - a.z * b.y * c.x - a.x * b.z * c.y - a.y * b.x * c.z;
}

__global__ void __launch_bounds__(128)
geometry_full_kernel(const FullParams p) {
    const long long pt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_surface = (long long)p.nalpha * p.nl;
    if (pt >= per_surface * p.ns) return;
    const int js = (int)(pt / per_surface);
    const int rem = (int)(pt - (long long)js * per_surface);
    const int ja = rem / p.nl, jl = rem - ja * p.nl;
    const double* sc = p.scal + (size_t)js * IBS_NSCAL;
    const double s_val = sc[0], iota = sc[1], d_iota = sc[2], dpds = sc[3], shat = sc[4];
    const double al = p.alpha[ja];
    const double* tm = p.tab_mn + (size_t)js * IBS_TAB_MN_ROWS * p.mnmax;
    const double* tq = p.tab_nyq + (size_t)js * IBS_TAB_NYQ_ROWS * p.mnmax_nyq;
    const double* bu = p.bsupu + (size_t)js * p.mnmax_nyq;
    const int mn = p.mnmax, mq = p.mnmax_nyq;

    double theta_p, phi, th, ang_phi;             // ang_phi: the phi used inside the mode sums
    int nit = 0; bool ok = true;
    if (p.mode == 2) {
        th = p.grid[jl];
        ang_phi = 0.0;                            // utils.py:961-982: phi is still zero when the angles are formed
        theta_p = 0.0; phi = 0.0;                 // set after lambda is known
    } else {
        if (p.mode == 0) { theta_p = p.grid[jl]; phi = p.phi_center + (theta_p - al) / iota; }
        else { phi = p.grid[jl]; theta_p = al + iota * (phi - p.phi_center); }
        ang_phi = phi;
        // theta_vmec + sum l sin(m theta_vmec - n phi) = theta_pest (utils.py:391-416): Newton with the analytic derivative
        th = theta_p;
        ok = false;
        for (nit = 1; nit <= 50; ++nit) {
            double f = th - theta_p, d = 1.0;
            for (int k = 0; k < mn; ++k) {
                double sa, ca;
                sincos(p.xm[k] * th - p.xn[k] * phi, &sa, &ca);
                f = fma(tm[2 * mn + k], sa, f);
                d = fma(tm[2 * mn + k] * p.xm[k], ca, d);
            }
            const double dth = f / d;
            th -= dth;
            if (fabs(dth) <= 4.5e-16 * fmax(1.0, fabs(th))) { ok = true; break; }
        }
    }
    // ---- mode sums (utils.py:420-468)
    const double tha = th + ((p.mode == 2) ? p.theta_shift : 0.0);
    double R = 0, R_s = 0, R_t = 0, R_p = 0, Z = 0, Z_s = 0, Z_t = 0, Z_p = 0, lam = 0, L_s = 0, L_t = 0, L_p = 0;
    for (int k = 0; k < mn; ++k) {
        const double m = p.xm[k], n = p.xn[k];
        double sa, ca;
        sincos(m * tha - n * ang_phi, &sa, &ca);
        const double r = tm[k], z = tm[mn + k], l = tm[2 * mn + k];
        R = fma(r, ca, R); R_s = fma(tm[3 * mn + k], ca, R_s); R_t = fma(-r * m, sa, R_t); R_p = fma(r * n, sa, R_p);
        Z = fma(z, sa, Z); Z_s = fma(tm[4 * mn + k], sa, Z_s); Z_t = fma(z * m, ca, Z_t); Z_p = fma(-z * n, ca, Z_p);
        lam = fma(l, sa, lam); L_s = fma(tm[5 * mn + k], sa, L_s); L_t = fma(l * m, ca, L_t); L_p = fma(-l * n, ca, L_p);
    }
    if (p.mode == 2) {
        theta_p = th + lam;                                   // utils.py:1043
        phi = p.phi_center + (theta_p - al) / iota;           // utils.py:1046
    }
    double sqrtg = 0, B = 0, B_s = 0, B_t = 0, B_p = 0, Bsup_t = 0, Bsup_p = 0, Bsub_s = 0, Bsub_t = 0, Bsub_p = 0;
    for (int k = 0; k < mq; ++k) {
        const double m = p.xm_nyq[k], n = p.zero_xn_nyq ? 0.0 : p.xn_nyq[k];
        double sa, ca;
        sincos(m * tha - n * ang_phi, &sa, &ca);
        const double b = tq[mq + k];
        sqrtg = fma(tq[k], ca, sqrtg); B = fma(b, ca, B); B_s = fma(tq[2 * mq + k], ca, B_s);
        B_t = fma(-b * m, sa, B_t); B_p = fma(b * n, sa, B_p);
        Bsup_t = fma(bu[k], ca, Bsup_t); Bsup_p = fma(tq[3 * mq + k], ca, Bsup_p);
        Bsub_s = fma(tq[4 * mq + k], sa, Bsub_s); Bsub_t = fma(tq[5 * mq + k], ca, Bsub_t); Bsub_p = fma(tq[6 * mq + k], ca, Bsub_p);
    }
    // ---- pointwise algebra (utils.py:470-720)
    const double psi_e = p.psi_e;
    double sp, cp;
    sincos(phi, &sp, &cp);
    const V3 e_t = {R_t * cp, R_t * sp, Z_t};                       // d(X, Y, Z)/d theta_vmec
    const V3 e_p = {R_p * cp - R * sp, R_p * sp + R * cp, Z_p};     // d/d phi
    const V3 e_s = {R_s * cp, R_s * sp, Z_s};                       // d/d s
    const double isg = 1.0 / sqrtg;
    const V3 grad_s = scale(cross(e_t, e_p), isg), grad_t = scale(cross(e_p, e_s), isg), grad_p = scale(cross(e_s, e_t), isg);
    const V3 grad_psi = scale(grad_s, psi_e);
    const double a_s = L_s - (phi - p.phi_center) * d_iota, a_t = 1.0 + L_t, a_p = -iota + L_p;
    const V3 grad_a = add3(scale(grad_s, a_s), scale(grad_t, a_t), scale(grad_p, a_p));
    const V3 grad_B = add3(scale(grad_s, B_s), scale(grad_t, B_t), scale(grad_p, B_p));
    const double bf = psi_e * isg;
    const V3 Bv = {bf * (a_t * e_p.x + (iota - L_p) * e_t.x), bf * (a_t * e_p.y + (iota - L_p) * e_t.y), bf * (a_t * e_p.z + (iota - L_p) * e_t.z)};
    const double lpi = L_p - iota;
    const double Bxgs_ga = (Bsub_p * a_t - Bsub_t * lpi) * isg;
    const double BxgB_ga = (Bsub_s * B_t * lpi + Bsub_t * B_p * a_s + Bsub_p * B_s * a_t - Bsub_p * B_t * a_s - Bsub_t * B_s * lpi -
                            Bsub_s * B_p * a_t) * isg;
    const double BxgB_gq = (Bsub_t * B_p - Bsub_p * B_t) * isg * psi_e;
    const double mu0 = 4.0 * 3.141592653589793 * 1.0e-7;
    const double L_ref = p.L_ref, B_ref = 2.0 * fabs(psi_e) / (L_ref * L_ref);
    const double sgn = (psi_e > 0.0) ? 1.0 : ((psi_e < 0.0) ? -1.0 : 0.0);
    const double sqrt_s = sqrt(s_val), B3 = B * B * B;
    const double ga_ga = dot(grad_a, grad_a), ga_gq = dot(grad_a, grad_psi), gq_gq = dot(grad_psi, grad_psi);
    const double Bsup_tp = iota * Bsup_p;
    const double gbdrift = -1.0 * 2.0 * B_ref * L_ref * L_ref * sqrt_s * BxgB_ga / B3 * sgn;
    const double gbdrift0 = -1.0 * BxgB_gq * 2.0 * shat / (B3 * sqrt_s) * sgn;
    const double cvdrift = gbdrift - 2.0 * B_ref * L_ref * L_ref * sqrt_s * mu0 * dpds * sgn / (psi_e * B * B);

    double* o = p.out + ((size_t)js * p.nalpha + ja) * FF_COUNT * p.nl + jl;
#define PUT(f, v) o[(size_t)(FF_##f) * p.nl] = (v)
    PUT(phi, phi); PUT(theta_pest, theta_p); PUT(theta_vmec, th); PUT(lambda, lam);
    PUT(R, R); PUT(d_R_d_s, R_s); PUT(d_R_d_theta_vmec, R_t); PUT(d_R_d_phi, R_p);
    PUT(Z, Z); PUT(d_Z_d_s, Z_s); PUT(d_Z_d_theta_vmec, Z_t); PUT(d_Z_d_phi, Z_p);
    PUT(d_lambda_d_s, L_s); PUT(d_lambda_d_theta_vmec, L_t); PUT(d_lambda_d_phi, L_p);
    PUT(sqrt_g_vmec, sqrtg); PUT(modB, B); PUT(d_B_d_s, B_s); PUT(d_B_d_theta_vmec, B_t); PUT(d_B_d_phi, B_p);
    PUT(B_sup_theta_vmec, Bsup_t); PUT(B_sup_phi, Bsup_p); PUT(B_sub_s, Bsub_s); PUT(B_sub_theta_vmec, Bsub_t); PUT(B_sub_phi, Bsub_p);
    PUT(B_sup_theta_pest, Bsup_tp); PUT(sqrt_g_vmec_alt, R * (Z_s * R_t - R_s * Z_t)); PUT(sinphi, sp); PUT(cosphi, cp);
    PUT(d_X_d_theta_vmec, e_t.x); PUT(d_X_d_phi, e_p.x); PUT(d_X_d_s, e_s.x);
    PUT(d_Y_d_theta_vmec, e_t.y); PUT(d_Y_d_phi, e_p.y); PUT(d_Y_d_s, e_s.y);
    PUT(grad_s_X, grad_s.x); PUT(grad_s_Y, grad_s.y); PUT(grad_s_Z, grad_s.z);
    PUT(grad_theta_vmec_X, grad_t.x); PUT(grad_theta_vmec_Y, grad_t.y); PUT(grad_theta_vmec_Z, grad_t.z);
    PUT(grad_phi_X, grad_p.x); PUT(grad_phi_Y, grad_p.y); PUT(grad_phi_Z, grad_p.z);
    PUT(grad_psi_X, grad_psi.x); PUT(grad_psi_Y, grad_psi.y); PUT(grad_psi_Z, grad_psi.z);
    PUT(grad_alpha_X, grad_a.x); PUT(grad_alpha_Y, grad_a.y); PUT(grad_alpha_Z, grad_a.z);
    PUT(grad_B_X, grad_B.x); PUT(grad_B_Y, grad_B.y); PUT(grad_B_Z, grad_B.z);
    PUT(B_X, Bv.x); PUT(B_Y, Bv.y); PUT(B_Z, Bv.z);
    PUT(B_cross_grad_s_dot_grad_alpha, Bxgs_ga); PUT(B_cross_grad_s_dot_grad_alpha_alternate, triple(Bv, grad_s, grad_a));
    PUT(B_cross_grad_B_dot_grad_alpha, BxgB_ga); PUT(B_cross_grad_B_dot_grad_alpha_alternate, triple(Bv, grad_B, grad_a));
    PUT(B_cross_grad_B_dot_grad_psi, BxgB_gq); PUT(B_cross_kappa_dot_grad_psi, BxgB_gq / B);
    PUT(B_cross_kappa_dot_grad_alpha, BxgB_ga / B + mu0 * dpds / psi_e);
    PUT(grad_alpha_dot_grad_alpha, ga_ga); PUT(grad_alpha_dot_grad_psi, ga_gq); PUT(grad_psi_dot_grad_psi, gq_gq);
    PUT(bmag, B / B_ref); PUT(gradpar_theta_pest, L_ref * Bsup_tp / B); PUT(gradpar_phi, L_ref * Bsup_p / B);
    PUT(gds2, ga_ga * L_ref * L_ref * s_val); PUT(gds21, ga_gq * shat / B_ref);
    PUT(gds22, gq_gq * shat * shat / (L_ref * L_ref * B_ref * B_ref * s_val));
    PUT(gbdrift, gbdrift); PUT(gbdrift0, gbdrift0); PUT(cvdrift, cvdrift); PUT(cvdrift0, gbdrift0);
    PUT(B_p, sqrt(Bsub_t * fabs(psi_e) * iota));                // utils.py:1282-1286 (axisymmetric routine only)
#undef PUT
    if (p.info) atomicMax(p.info + ((size_t)js * p.nalpha + ja), nit | (ok ? 0 : (1 << 16)));
}

int geometry_full_nfields() { return FF_COUNT; }

int geometry_full_dispatch(const double* tab_mn, const double* tab_nyq, const double* bsupumnc, const double* scal,
                           const double* xm, const double* xn, const double* xm_nyq, const double* xn_nyq,
                           int ns, int mnmax, int mnmax_nyq, double phiedge, double aminor_p, const double* alpha, int nalpha,
                           const double* grid, int nl, int mode, double theta_shift, int zero_xn_nyq, double phi_center,
                           double* out, int* info, cudaStream_t st) {
    FullParams p;
    p.tab_mn = tab_mn; p.tab_nyq = tab_nyq; p.bsupu = bsupumnc; p.scal = scal;
    p.xm = xm; p.xn = xn; p.xm_nyq = xm_nyq; p.xn_nyq = xn_nyq;
    p.ns = ns; p.mnmax = mnmax; p.mnmax_nyq = mnmax_nyq; p.alpha = alpha; p.nalpha = nalpha; p.grid = grid; p.nl = nl; p.mode = mode;
    p.theta_shift = theta_shift; p.phi_center = phi_center; p.psi_e = -phiedge / (2.0 * 3.141592653589793); p.L_ref = aminor_p;
    p.zero_xn_nyq = zero_xn_nyq; p.out = out; p.info = info;
    const long long npt = (long long)ns * nalpha * nl;
    if (npt == 0) return IBS_OK;
    if (info) IBS_CUDA_CHECK(cudaMemsetAsync(info, 0, (size_t)ns * nalpha * sizeof(int), st));
    geometry_full_kernel<<<(unsigned)((npt + 127) / 128), 128, 0, st>>>(p);
    IBS_CUDA_CHECK(cudaGetLastError());
    return IBS_OK;
}

}  // namespace ibs
