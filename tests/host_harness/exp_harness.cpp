// Host build of pilco_b200/csrc/exp_table.cuh -- the table exp the tile kernels execute -- so the CPU suite can
// check its accuracy and clamping against libm without a GPU.  Test infrastructure only.
#include "../../pilco_b200/csrc/exp_table.cuh"

extern "C" {
// same values as the table compiled into the library (scripts/gen_exp_table.py); EXP_STRIDE interleaved copies
void exp_harness_table(double* tab) {
    for (int j = 0; j < EXP_TAB; ++j)
        for (int c = 0; c < EXP_STRIDE; ++c) tab[j * EXP_STRIDE + c] = (double)exp2l((long double)j / (long double)EXP_TAB);
}
double exp_harness_scale(void) { return EXP_SC; }
int exp_harness_entries(void) { return EXP_TAB; }
int exp_harness_stride(void) { return EXP_STRIDE; }
// exp_scaled on PRE-SCALED arguments xs (= x * EXP_SC): returns 2^(xs / EXP_TAB)
void exp_harness_scaled(int n, const double* xs, const double* tab, double* out) {
    for (int i = 0; i < n; ++i) out[i] = exp_scaled(xs[i], EXP_LANE_TAB(tab, i));       // element i plays lane i
}
// exp((c + A)/EXP_SC) the way a tile row computes it: exp_row_split(A) once, exp_shifted(c, am) per element, row factor
void exp_harness_shifted(int n, const double* c, double A, const double* tab, double* out, double* rowfac_out) {
    double am, rowfac;
    exp_row_split(A, am, rowfac);
    *rowfac_out = rowfac;
    for (int i = 0; i < n; ++i) out[i] = exp_shifted(c[i], am, EXP_LANE_TAB(tab, i)) * rowfac;
}
}
