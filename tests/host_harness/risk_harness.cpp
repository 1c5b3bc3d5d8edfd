// Host build of pilco_b200/csrc/risk_math.cuh (the scalar formulas the device kernels execute) so that the
// CPU test suite can check them against the oracle without a GPU.  Test infrastructure only.
#include "../../pilco_b200/csrc/risk_math.cuh"

extern "C" double risk_box_eval_host(int Ds, const double* prm, const double* m, const double* s, double* dm, double* dv) {
    return risk_box_eval(Ds, prm, m, s, dm, dv);
}
