// matvec_lut.cu -- table-driven bed mat-vecs (generation 1).  Placeholder wiring: until the table
// kernels land both entry points run the generation-0 kernels.
#include "gvb_internal.cuh"

int gvb_ax_lut(gvb_ctx* c, const double* v, double* out) { return gvb_ax_simple(c, v, out); }
int gvb_atx_lut(gvb_ctx* c, const double* u, double* out) { return gvb_atx_simple(c, u, out); }
