// CPU harness for the per-problem logic of the batched refine (csrc/ibs_refine_core.cuh).  TEST INFRASTRUCTURE ONLY
// (tests/test_refine_core_host.py): drives refine::init / refine::consume exactly as the CUDA step kernel does, with the
// objective evaluated by a Python callback.  Not part of the library.
//   g++ -O2 -std=c++17 -shared -fPIC -o refine_core_host.so tools/refine_core_host.cpp
#include "../ideal-ballooning-solver_b200/csrc/ibs_refine_core.cuh"

using namespace ibs::refine;
extern "C" int refine_host_nstate() { return NSTATE; }
extern "C" void refine_host_init(double* state, double a0, double t0, double alo, double ahi, double tlo, double thi) {
    init(*reinterpret_cast<State*>(state), a0, t0, alo, ahi, tlo, thi);
}
extern "C" void refine_host_consume(double* state, double f, double g0, double g1, int failed, double ftol, double gtol, int maxiter) {
    Options o; o.ftol = ftol; o.gtol = gtol; o.maxiter = maxiter;
    consume(*reinterpret_cast<State*>(state), f, g0, g1, failed != 0, o);
}
