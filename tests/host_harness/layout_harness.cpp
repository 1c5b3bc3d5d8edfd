// Host build of the workspace-layout functions (pilco_b200/csrc/{mm_kernels,mm_backward,rollout}.cuh) so the CPU suite
// can check their invariants (ordering, alignment, sizes) against the sizes the shared library reports.
#include "../../pilco_b200/csrc/mm_backward.cuh"
#include "../../pilco_b200/csrc/rollout.cuh"

extern "C" {
// fills out[0..10): zeta, betap, Bq, Tpart, Wm, Wc, Qab, Ufrag, Arow, per_r (doubles); returns np
int layout_mm(int n, int D, int E, int ordered, unsigned long long* out) {
    const MMWs L = mm_ws_layout(n, D, E, ordered != 0);
    const size_t v[10] = {L.zeta, L.betap, L.Bq, L.Tpart, L.Wm, L.Wc, L.Qab, L.Ufrag, L.Arow, L.per_r};
    for (int i = 0; i < 10; ++i) out[i] = v[i];
    return L.np;
}
unsigned long long layout_bwd_per_r(int n, int D, int E, int need_param) { return mm_bws_layout(n, D, E, need_param).per_r; }
int layout_pair(int q, int* a, int* b) { int aa, bb; pair_decode(q, aa, bb); *a = aa; *b = bb; return pair_index(aa, bb); }
int layout_misc(int D, int* ldz, int* ks) { *ldz = ldz_of(D); *ks = ksteps_of(D); return pad64(D); }
}
