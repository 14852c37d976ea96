#include "fx_oracle.h"
int fxo_inverse(int64_t npts, const double *o, const double *theta, double *D_at_knots,
                double *sorptivity_pchip, double *sorptivity_trapz) {
  (void)npts; (void)o; (void)theta; (void)D_at_knots; (void)sorptivity_pchip; (void)sorptivity_trapz;
  return -1;
}
