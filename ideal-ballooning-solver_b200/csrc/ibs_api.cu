// C ABI of the library (see include/ibs_b200.h).  Thin argument checking + dispatch; no torch types.
#include <cstring>
#include <string>
#include <vector>

#include "ibs_common.cuh"

namespace ibs {

// ---- defined in the kernel translation units -------------------------------------------------------
int solve_dispatch(const SolveParams& p, bool base, bool count_only, cudaStream_t stream);
int geometry_dispatch(const double* tab_mn, const double* tab_nyq, const double* scal,
                      const double* xm, const double* xn, const double* xm_nyq, const double* xn_nyq,
                      int ns, int mnmax, int mnmax_nyq, double phiedge, double aminor_p,
                      const double* alpha, int nalpha, int alpha_per_surface, const double* theta, int nl,
                      double phi_center, double* base_out, double* dPdrho_out, double* theta_vmec_out,
                      int* info_out, cudaStream_t st);
int launch_adjoint(const double* lam, const double* X, const double* dX, const double* f, const double* g_p,
                   const double* c_p, const double* f_p, int nsolve, int nparam, int N, double* grad, cudaStream_t st);
int launch_sensitivity(const double* lam, const double* X, const double* dX, const double* f, int nsolve, int N,
                       double* dg, double* dc, double* df, cudaStream_t st);
int launch_obj_grad(const double* base3, const double* dPdrho3, const double* theta0, const double* lam,
                    const double* X, const double* dX, int npoint, int N, double del_alpha, double* val,
                    double* grad, cudaStream_t st);
int launch_centre_lines(int* line, int n, cudaStream_t st);
int launch_argmax(const double* gamma, int ns, int ngrid, double* val, int* idx, double* sigma0, cudaStream_t st);
int launch_gather_best(const double* X, const int* idx, int ns, int ngrid, int N, double* out, cudaStream_t st);
int launch_best_setup(const int* idx, const double* theta0, int ns, int ngrid, int nth0, int* line_out, double* th0_out,
                      cudaStream_t st);

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return IBS_ERR_CUDA;
}
int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// Warm-start chain used by the scan entry points: consecutive theta0 of one field line (the reference chains
// its ARPACK start vector through the same loop, ball_scan.py:265-274).  The run length divides nth0.
static int scan_chain_len(int nth0) {
    int k = nth0 < 16 ? nth0 : 16;
    while (k > 1 && nth0 % k) --k;
    return k;
}

void keep_pool_cached() {
    static bool done[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ULL;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done[dev] = true;
}

static SolveParams blank_params() { SolveParams p; std::memset(&p, 0, sizeof(p)); return p; }

}  // namespace ibs

using namespace ibs;

extern "C" {

int ibs_version(void) { return 100; }
const char* ibs_last_error(void) { return g_err.c_str(); }

int ibs_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    IBS_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    IBS_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return IBS_OK;
}

int ibs_geometry_batch(const double* tab_mn, const double* tab_nyq, const double* scal,
                       const double* xm, const double* xn, const double* xm_nyq, const double* xn_nyq,
                       int ns, int mnmax, int mnmax_nyq, double phiedge, double aminor_p,
                       const double* alpha, int nalpha, int alpha_per_surface,
                       const double* theta, int nl, double phi_center,
                       double* base_out, double* dPdrho_out, double* theta_vmec_out, int* info_out, void* stream) {
    IBS_REQUIRE(tab_mn && tab_nyq && scal && xm && xn && xm_nyq && xn_nyq && alpha && theta && base_out, "null pointer");
    IBS_REQUIRE(ns >= 0 && nalpha >= 0 && nl >= 1 && mnmax >= 1 && mnmax_nyq >= 1, "bad sizes");
    IBS_REQUIRE(aminor_p > 0.0 && phiedge != 0.0, "Aminor_p must be > 0 and phiedge != 0");
    if (ns == 0 || nalpha == 0) return IBS_OK;
    return geometry_dispatch(tab_mn, tab_nyq, scal, xm, xn, xm_nyq, xn_nyq, ns, mnmax, mnmax_nyq, phiedge, aminor_p, alpha,
                             nalpha, alpha_per_surface, theta, nl, phi_center, base_out, dPdrho_out, theta_vmec_out,
                             info_out, (cudaStream_t)stream);
}

int ibs_solve_gcf_batch(const double* g, const double* c, const double* f, int nsolve, int N, double h,
                        const double* lam0, const double* sigma, int chain_len, double* lam_out, double* lam_matrix_out,
                        double* X_out, double* dX_out, int* info_out, void* stream) {
    IBS_REQUIRE(nsolve >= 0 && N >= 5, "need nsolve >= 0 and N >= 5");
    if (nsolve == 0) return IBS_OK;                 // empty batches are legal (and have null pointers)
    IBS_REQUIRE(g && c && f && lam_out, "null pointer");
    IBS_REQUIRE(h > 0.0, "h must be positive");
    SolveParams p = blank_params();
    p.g = g; p.c = c; p.f = f; p.nsolve = nsolve; p.N = N; p.h = h; p.lam0 = lam0; p.sigma = sigma; p.chain_len = chain_len;
    p.lam_out = lam_out; p.lam_matrix_out = lam_matrix_out; p.X_out = X_out; p.dX_out = dX_out; p.info_out = info_out;
    return solve_dispatch(p, false, false, (cudaStream_t)stream);
}

int ibs_solve_base_batch(const double* base, const double* dPdrho, const double* theta0, const int* line_of_solve,
                         int nth0, int nsolve, int N, double h, const double* lam0, const double* sigma, int chain_len,
                         double* lam_out, double* lam_matrix_out, double* X_out, double* dX_out, double* g_out,
                         double* c_out, double* f_out, int* info_out, void* stream) {
    IBS_REQUIRE(nsolve >= 0 && N >= 5, "need nsolve >= 0 and N >= 5");
    if (nsolve == 0) return IBS_OK;
    IBS_REQUIRE(base && dPdrho && theta0 && lam_out, "null pointer");
    IBS_REQUIRE(line_of_solve || nth0 >= 1, "need line_of_solve or nth0 >= 1");
    keep_pool_cached();
    IBS_REQUIRE(h > 0.0, "h must be positive");
    SolveParams p = blank_params();
    p.base = base; p.dPdrho = dPdrho; p.theta0 = theta0; p.line_of_solve = line_of_solve; p.nth0 = nth0 > 0 ? nth0 : 1;
    p.nsolve = nsolve; p.N = N; p.h = h; p.lam0 = lam0; p.sigma = sigma; p.chain_len = chain_len;
    p.lam_out = lam_out; p.lam_matrix_out = lam_matrix_out; p.X_out = X_out; p.dX_out = dX_out;
    p.g_out = g_out; p.c_out = c_out; p.f_out = f_out; p.info_out = info_out;
    return solve_dispatch(p, true, false, (cudaStream_t)stream);
}

int ibs_count_above_batch(const double* g, const double* c, const double* f, int nsolve, int N, double h,
                          const double* lam, int* count_out, void* stream) {
    IBS_REQUIRE(nsolve >= 0 && N >= 3 && h > 0.0, "bad sizes");
    if (nsolve == 0) return IBS_OK;
    IBS_REQUIRE(g && c && f && lam && count_out, "null pointer");
    SolveParams p = blank_params();
    p.g = g; p.c = c; p.f = f; p.nsolve = nsolve; p.N = N; p.h = h; p.lam_query = lam; p.count_out = count_out;
    return solve_dispatch(p, false, true, (cudaStream_t)stream);
}

int ibs_adjoint_batch(const double* lam, const double* X, const double* dX, const double* f, const double* g_p,
                      const double* c_p, const double* f_p, int nsolve, int nparam, int N, double* grad_out,
                      void* stream) {
    IBS_REQUIRE(nsolve >= 0 && nparam >= 0 && N >= 3, "bad sizes");
    if (nsolve == 0 || nparam == 0) return IBS_OK;
    IBS_REQUIRE(lam && X && dX && f && g_p && c_p && f_p && grad_out, "null pointer");
    return launch_adjoint(lam, X, dX, f, g_p, c_p, f_p, nsolve, nparam, N, grad_out, (cudaStream_t)stream);
}

int ibs_adjoint_sensitivities(const double* lam, const double* X, const double* dX, const double* f, int nsolve,
                              int N, double* dlam_dg, double* dlam_dc, double* dlam_df, void* stream) {
    IBS_REQUIRE(nsolve >= 0 && N >= 3, "bad sizes");
    if (nsolve == 0) return IBS_OK;
    IBS_REQUIRE(lam && X && dX && f, "null pointer");
    return launch_sensitivity(lam, X, dX, f, nsolve, N, dlam_dg, dlam_dc, dlam_df, (cudaStream_t)stream);
}

int ibs_obj_w_grad_batch(const double* base3, const double* dPdrho3, const double* theta0, int npoint, int N,
                         double h, double del_alpha, const double* lam0, double* val_out, double* grad_out,
                         double* X_out, double* dX_out, int* info_out, void* stream) {
    IBS_REQUIRE(npoint >= 0 && N >= 5 && h > 0.0 && del_alpha != 0.0, "bad sizes");
    if (npoint == 0) return IBS_OK;
    IBS_REQUIRE(base3 && dPdrho3 && theta0 && val_out && grad_out, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    // stream-ordered scratch: centre-line indices, lambda, and X / dX when the caller does not want them
    int* line = nullptr; double* lam = nullptr; double* Xw = nullptr; double* dXw = nullptr;
    IBS_CUDA_CHECK(cudaMallocAsync((void**)&line, (size_t)npoint * sizeof(int), st));
    IBS_CUDA_CHECK(cudaMallocAsync((void**)&lam, (size_t)npoint * sizeof(double), st));
    if (!X_out) IBS_CUDA_CHECK(cudaMallocAsync((void**)&Xw, (size_t)npoint * N * sizeof(double), st));
    if (!dX_out) IBS_CUDA_CHECK(cudaMallocAsync((void**)&dXw, (size_t)npoint * N * sizeof(double), st));
    double* X = X_out ? X_out : Xw;
    double* dX = dX_out ? dX_out : dXw;
    int rc = launch_centre_lines(line, npoint, st);
    if (rc == IBS_OK) {
        SolveParams p = blank_params();
        p.base = base3; p.dPdrho = dPdrho3; p.theta0 = theta0; p.line_of_solve = line; p.nth0 = 1;
        p.nsolve = npoint; p.N = N; p.h = h; p.lam0 = lam0;
        p.lam_out = lam; p.X_out = X; p.dX_out = dX; p.info_out = info_out;
        rc = solve_dispatch(p, true, false, st);
    }
    if (rc == IBS_OK) rc = launch_obj_grad(base3, dPdrho3, theta0, lam, X, dX, npoint, N, del_alpha, val_out, grad_out, st);
    cudaFreeAsync(line, st); cudaFreeAsync(lam, st);
    if (Xw) cudaFreeAsync(Xw, st);
    if (dXw) cudaFreeAsync(dXw, st);
    return rc;
}

int ibs_scan_argmax(const double* gamma, int ns, int ngrid, double* val_out, int* idx_out, double* sigma0_out,
                    void* stream) {
    IBS_REQUIRE(gamma && val_out && idx_out, "null pointer");
    IBS_REQUIRE(ns >= 0 && ngrid >= 1, "bad sizes");
    return launch_argmax(gamma, ns, ngrid, val_out, idx_out, sigma0_out, (cudaStream_t)stream);
}

int ibs_scan_host(const double* tab_mn, const double* tab_nyq, const double* scal,
                  const double* xm, const double* xn, const double* xm_nyq, const double* xn_nyq,
                  int ns, int mnmax, int mnmax_nyq, double phiedge, double aminor_p,
                  const double* alpha, int nalpha, const double* theta0, int nth0,
                  const double* theta, int nl, double h,
                  double* gamma_out, double* val_out, int* idx_out, double* sigma0_out, double* xbest_out, int* nbad_out) {
    IBS_REQUIRE(tab_mn && tab_nyq && scal && alpha && theta0 && theta && gamma_out, "null pointer");
    IBS_REQUIRE(ns >= 1 && nalpha >= 1 && nth0 >= 1 && nl >= 3, "bad sizes");
    keep_pool_cached();
    cudaStream_t st = nullptr;
    IBS_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const size_t nlines = (size_t)ns * nalpha, nsolve = nlines * nth0;
    const size_t b_tab_mn = (size_t)ns * 6 * mnmax * 8, b_tab_nyq = (size_t)ns * 7 * mnmax_nyq * 8, b_scal = (size_t)ns * IBS_NSCAL * 8;
    // one device arena: inputs | base | dPdrho | theta0 per solve | gamma | val | sigma0 | idx | info
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_tmn = take(b_tab_mn), o_tnq = take(b_tab_nyq), o_sc = take(b_scal), o_al = take((size_t)nalpha * 8),
                 o_th = take((size_t)nl * 8), o_t0 = take(nsolve * 8), o_base = take(nlines * IBS_NBASE * nl * 8),
                 o_dp = take(nlines * 8), o_gam = take(nsolve * 8), o_val = take((size_t)ns * 8), o_sig = take((size_t)ns * 8),
                 o_idx = take((size_t)ns * 4), o_info = take(nsolve * 4),
                 o_xb = take(xbest_out ? (size_t)ns * nl * 8 : 0), o_bl = take((size_t)ns * 4), o_bt = take((size_t)ns * 8),
                 o_bg = take((size_t)ns * 8);
    char* d = nullptr;
    int rc = IBS_OK;
    std::vector<double> t0_rep(nsolve);
    for (size_t i = 0; i < nsolve; ++i) t0_rep[i] = theta0[i % nth0];
    std::vector<int> info(nsolve);
#define IBS_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { rc = cuda_fail(_e, #expr); goto done; } } while (0)
    IBS_TRY(cudaMallocAsync((void**)&d, off, st));
    IBS_TRY(cudaMemcpyAsync(d + o_tmn, tab_mn, b_tab_mn, cudaMemcpyHostToDevice, st));
    IBS_TRY(cudaMemcpyAsync(d + o_tnq, tab_nyq, b_tab_nyq, cudaMemcpyHostToDevice, st));
    IBS_TRY(cudaMemcpyAsync(d + o_sc, scal, b_scal, cudaMemcpyHostToDevice, st));
    IBS_TRY(cudaMemcpyAsync(d + o_al, alpha, (size_t)nalpha * 8, cudaMemcpyHostToDevice, st));
    IBS_TRY(cudaMemcpyAsync(d + o_th, theta, (size_t)nl * 8, cudaMemcpyHostToDevice, st));
    IBS_TRY(cudaMemcpyAsync(d + o_t0, t0_rep.data(), nsolve * 8, cudaMemcpyHostToDevice, st));
    rc = geometry_dispatch((double*)(d + o_tmn), (double*)(d + o_tnq), (double*)(d + o_sc), xm, xn, xm_nyq, xn_nyq, ns, mnmax,
                           mnmax_nyq, phiedge, aminor_p, (double*)(d + o_al), nalpha, 0, (double*)(d + o_th), nl, 0.0,
                           (double*)(d + o_base), (double*)(d + o_dp), nullptr, nullptr, st);
    if (rc != IBS_OK) goto done;
    {
        SolveParams p = blank_params();
        p.base = (double*)(d + o_base); p.dPdrho = (double*)(d + o_dp); p.theta0 = (double*)(d + o_t0); p.nth0 = nth0;
        p.nsolve = (int)nsolve; p.N = nl; p.h = h; p.lam_out = (double*)(d + o_gam); p.info_out = (int*)(d + o_info);
        p.chain_len = scan_chain_len(nth0);
        rc = solve_dispatch(p, true, false, st);
        if (rc != IBS_OK) goto done;
    }
    rc = launch_argmax((double*)(d + o_gam), ns, nalpha * nth0, (double*)(d + o_val), (int*)(d + o_idx), (double*)(d + o_sig), st);
    if (rc != IBS_OK) goto done;
    if (xbest_out) {
        // eigenfunction of each surface's arg-max only: re-solve those ns problems with the eigenvector written out
        // (instead of writing nsolve eigenvectors and gathering ns of them)
        rc = launch_best_setup((int*)(d + o_idx), (double*)(d + o_t0), ns, nalpha * nth0, nth0, (int*)(d + o_bl), (double*)(d + o_bt), st);
        if (rc != IBS_OK) goto done;
        SolveParams pb = blank_params();
        pb.base = (double*)(d + o_base); pb.dPdrho = (double*)(d + o_dp); pb.theta0 = (double*)(d + o_bt);
        pb.line_of_solve = (int*)(d + o_bl); pb.nth0 = 1; pb.nsolve = ns; pb.N = nl; pb.h = h;
        pb.lam_out = (double*)(d + o_bg); pb.X_out = (double*)(d + o_xb);
        rc = solve_dispatch(pb, true, false, st);
        if (rc != IBS_OK) goto done;
        IBS_TRY(cudaMemcpyAsync(xbest_out, d + o_xb, (size_t)ns * nl * 8, cudaMemcpyDeviceToHost, st));
    }
    IBS_TRY(cudaMemcpyAsync(gamma_out, d + o_gam, nsolve * 8, cudaMemcpyDeviceToHost, st));
    if (val_out) IBS_TRY(cudaMemcpyAsync(val_out, d + o_val, (size_t)ns * 8, cudaMemcpyDeviceToHost, st));
    if (idx_out) IBS_TRY(cudaMemcpyAsync(idx_out, d + o_idx, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
    if (sigma0_out) IBS_TRY(cudaMemcpyAsync(sigma0_out, d + o_sig, (size_t)ns * 8, cudaMemcpyDeviceToHost, st));
    IBS_TRY(cudaMemcpyAsync(info.data(), d + o_info, nsolve * 4, cudaMemcpyDeviceToHost, st));
    IBS_TRY(cudaStreamSynchronize(st));
    if (nbad_out) {
        int nb = 0;
        for (size_t i = 0; i < nsolve; ++i) nb += ((info[i] >> 16) & (IBS_FLAG_NOT_CONVERGED | IBS_FLAG_BAD_INPUT)) ? 1 : 0;
        *nbad_out = nb;
    }
done:
#undef IBS_TRY
    if (d) cudaFreeAsync(d, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}

}  // extern "C"
