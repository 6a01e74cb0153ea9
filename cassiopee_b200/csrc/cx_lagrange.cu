// placeholder, replaced below
#include "cx_internal.h"
using namespace cx;
int cx_lagrange_pipeline(cx_ctx* ctx, int order, int n, const double* d_xr, const double* d_yr,
                         const double* d_zr, int nz, const ZoneView* d_zones, const ZoneView* h_zones,
                         int nature, int penalty, int* noblk, int* type, int* dind, double* cf, int ncfmax,
                         unsigned long long* d_stats)
{ cx_set_error("order %d not built yet", order); return -4; }
