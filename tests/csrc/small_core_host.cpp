// Host build of krypy_b200/csrc/kry_small_core.h for the CPU test tier (tests/test_small_core_cpu.py):
// the very code givens_z_kernel / tri_solve_z_kernel run in their single-thread sections.
#include "../../krypy_b200/csrc/kry_small_core.h"

extern "C" {
void host_zrotg(double fr, double fi, double gr, double gi, double* out3) {
    kryc_zrotg(fr, fi, gr, gi, out3, out3 + 1, out3 + 2);
}
void host_drotg(double a, double b, double* out2) { kryc_drotg(a, b, out2, out2 + 1); }
double host_givens_step(int k, double* r, const double* rot, double* rot_new, double* y2) {
    return kryc_givens_step(k, r, rot, rot_new, y2);
}
void host_tri_solve(int k, const double* R, long long ldr, double* x) { kryc_tri_solve(k, R, ldr, x); }
}
