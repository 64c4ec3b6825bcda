// Host driver for hephaestus_b200/csrc/interp_body.h (TEST INFRASTRUCTURE): runs the per-thread program of interp_kernel
// for every output dof in a plain loop, so the index arithmetic of the discrete gradient / curl kernels is checked
// against the oracle without a GPU (tests/test_interp.py).
#include <cstring>

#include "interp_body.h"

extern "C" int emu_interp_apply(int kind, int p, int64_t n_elem, int nd_in, int nd_out, const uint32_t *in_map, const uint32_t *out_map,
                                const double *D, const double *x, double *y, int geom, const double *Mloc) {
  hcurl::InterpArgs a;
  memset(&a, 0, sizeof(a));
  a.kind = kind; a.p = p; a.n_elem = n_elem; a.nd_in = nd_in; a.nd_out = nd_out; a.in_map = in_map; a.out_map = out_map; a.x = x; a.y = y;
  a.geom = geom;
  if (geom) memcpy(a.Mloc, Mloc, sizeof(double) * 24);
  else memcpy(a.D, D, sizeof(double) * (size_t)p * (p + 1));
  for (int64_t gid = 0; gid < n_elem * nd_out; gid++) hcurl::interp_thread(a, gid);
  return 0;
}
