// tests/hostemu/hostemu.cpp — TEST INFRASTRUCTURE ONLY.
// Compiles the device header cassiopee_b200/csrc/cx_core.cuh with g++ (CX_D
// empty) so the per-thread search logic and the level-synchronous ADT build
// can be checked against the oracle on the GPU-less builder box.  Nothing in
// the product loads this library; the product has no CPU path.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "../../cassiopee_b200/csrc/cx_core.cuh"
using namespace cx;

struct EmuZone {
  int topo = 1; double h[3] = {0., 0., 0.};
  int ni, nj, nk, npts, ncells, depth;
  std::vector<double> x, y, z, c;
  std::vector<AdtNode> nodes;
  double bbox[6];
  ZoneView view() const {
    ZoneView V; V.x = x.data(); V.y = y.data(); V.z = z.data(); V.cellN = c.data(); V.nodes = nodes.data();
    V.ni = ni; V.nj = nj; V.nk = nk; V.ncells = ncells; V.depth = depth;
    for (int i = 0; i < 6; i++) V.bbox[i] = bbox[i];
    V.topo = topo;
    for (int i = 0; i < 3; i++) { V.h[i] = h[i]; V.hinv[i] = topo == 0 ? 1.0 / h[i] : 0.; V.hs2[i] = 0.5 * h[i]; quant_params(bbox, i, V.qorg[i], V.qscl[i]); }
    return V;
  }
};

extern "C" {

void* emu_zone_create(int ni, int nj, int nk, const double* x, const double* y, const double* z, const double* cellN,
                      const double* bbox)
{
  EmuZone* Z = new EmuZone;
  Z->ni = ni; Z->nj = nj; Z->nk = nk; Z->npts = ni * nj * nk; Z->ncells = (ni - 1) * (nj - 1) * (nk > 1 ? nk - 1 : 1);
  Z->x.assign(x, x + Z->npts); Z->y.assign(y, y + Z->npts); Z->z.assign(z, z + Z->npts);
  if (cellN) Z->c.assign(cellN, cellN + Z->npts); else Z->c.assign(Z->npts, 1.0);
  for (int i = 0; i < 6; i++) Z->bbox[i] = bbox[i];
  int n = Z->ncells;
  Z->nodes.resize(n);
  std::vector<double> keys(6 * (size_t)n), lo(6 * (size_t)n), hi(6 * (size_t)n);
  std::vector<int> cur(n, 0), dep(n, 0); std::vector<unsigned char> br(n, 0);
  BuildView B; B.keys = keys.data(); B.lo = lo.data(); B.hi = hi.data(); B.cur = cur.data(); B.br = br.data(); B.depth = dep.data();
  B.nodes = Z->nodes.data(); B.ncells = n;
  std::vector<int> child(2 * (size_t)n, CX_NULL); B.child = child.data();
  for (int d = 0; d < 6; d++) { B.zb.lo[d] = bbox[d % 3]; B.zb.hi[d] = bbox[3 + d % 3]; }
  for (int c = 0; c < n; c++) adt_cell_keys(ni, nj, nk, x, y, z, c, n, B.keys, B.nodes, bbox, bbox[2], bbox[5]);
  freeze_node(B.nodes[0], 0, B.zb.lo[0], B.zb.hi[0]);
  std::vector<int> act, next;
  for (int c = 1; c < n; c++) act.push_back(c);
  // shuffle to prove order independence of the level-synchronous build
  for (size_t i = act.size(); i > 1; i--) { size_t j = (size_t)rand() % i; std::swap(act[i - 1], act[j]); }
  int L = 0, depth = 0;
  while (!act.empty())
  {
    for (int c : act)
    {
      int p, right; adt_claim(B, L, c, p, right);
      int* slot = &B.child[2 * (size_t)p + (right ? 1 : 0)];
      if (c < *slot) *slot = c;   // atomicMin
    }
    next.clear();
    for (int c : act) { if (!adt_resolve(B, L, c)) next.push_back(c); else depth = L + 1; }
    act.swap(next);
    L++;
  }
  Z->depth = depth;
  for (int c = 0; c < n; c++) adt_set_children(B, c);
  for (int lv = depth - 1; lv >= 0; lv--)
    for (int c = 0; c < n; c++) if (dep[c] == lv) adt_subtree_box(B.nodes, c);
  return Z;
}

// regular Cartesian donor: same set-up arithmetic as cx_zone_create_cart (cx_api.cu)
void* emu_zone_create_cart(int ni, int nj, int nk, const double* x, const double* y, const double* z, const double* cellN)
{
  EmuZone* Z = new EmuZone;
  Z->ni = ni; Z->nj = nj; Z->nk = nk; Z->npts = ni * nj * nk; Z->ncells = (ni - 1) * (nj - 1) * (nk - 1); Z->depth = 0;
  Z->x.assign(x, x + Z->npts); Z->y.assign(y, y + Z->npts); Z->z.assign(z, z + Z->npts);
  if (cellN) Z->c.assign(cellN, cellN + Z->npts); else Z->c.assign(Z->npts, 1.0);
  Z->topo = 0;
  Z->h[0] = x[1] - x[0]; Z->h[1] = y[ni] - y[0]; Z->h[2] = z[(size_t)ni * nj] - z[0];
  Z->bbox[0] = x[0]; Z->bbox[1] = y[0]; Z->bbox[2] = z[0];
  Z->bbox[3] = Z->bbox[0] + (ni - 1) * Z->h[0];
  Z->bbox[4] = Z->bbox[1] + (nj - 1) * Z->h[1];
  Z->bbox[5] = Z->bbox[2] + (nk - 1) * Z->h[2];
  return Z;
}

// InterpCart order 3 of one zone (the device does this in k_lag_search for topo 0)
void emu_cart_o3(void* h, int n, const double* xr, const double* yr, const double* zr, int* found, int* dind, double* cf)
{
  ZoneView V = ((EmuZone*)h)->view();
  for (int r = 0; r < n; r++)
  {
    int ic = 0, jc = 0, kc = 0;
    found[r] = cart_search_o3(V, xr[r], yr[r], zr[r], ic, jc, kc, cf + (size_t)r * 9);
    dind[r] = (ic - 1) + (jc - 1) * V.ni + (kc - 1) * V.ni * V.nj;
  }
}

void emu_cart_o4(void* h, int n, const double* xr, const double* yr, const double* zr, int* found, int* dind, double* cf)
{
  ZoneView V = ((EmuZone*)h)->view();
  for (int r = 0; r < n; r++)
  {
    int ic = 0, jc = 0, kc = 0;
    found[r] = cart_search_o4(V, xr[r], yr[r], zr[r], ic, jc, kc, cf + (size_t)r * 12);
    dind[r] = (ic - 1) + (jc - 1) * V.ni + (kc - 1) * V.ni * V.nj;
  }
}

void emu_zone_free(void* h) { delete (EmuZone*)h; }
int emu_zone_depth(void* h) { return ((EmuZone*)h)->depth; }

int emu_candidates(void* h, double x, double y, double z, double alphaTol, int* out, int cap)
{
  ZoneView V = ((EmuZone*)h)->view();
  if (zone_reject(V, x, y, z, alphaTol)) return 0;
  AdtQuery q; query_init(q, V, x, y, z, alphaTol);
  int c, nid, nv = 0, n = 0;
  while ((c = query_next(q, V.nodes, &nid, &nv)) >= 0) { if (n < cap) out[n] = c; n++; }
  return n;
}

static int use_scratch = 1;
void emu_use_scratch(int on) { use_scratch = on; }

void emu_set_interp_data(int n, const double* xr, const double* yr, const double* zr, int nz, void** zones,
                         int nature, int penalty, int extrap, int* noblk, int* type, int* dind, int* isext,
                         double* cf, long long* stats)
{
  std::vector<ZoneView> V(nz);
  for (int i = 0; i < nz; i++) V[i] = ((EmuZone*)zones[i])->view();
  SearchStats st; st.nvisit = st.ncand = st.ntet = 0;
  for (int r = 0; r < n; r++)
  {
    RcvOut o;
    if (use_scratch)
    {
      // the version the kernels run (search state in a Scratch area; here plain arrays, stride 1)
      int stk[CX_STACK]; double vtx[24];
      Scratch S; S.stk = stk; S.vtx = vtx; S.ss = 1; S.cap = CX_STACK;
      interp_o2_receptor_s(xr[r], yr[r], zr[r], nz, V.data(), nature, penalty, o, S, &st);
    }
    else interp_o2_receptor(xr[r], yr[r], zr[r], nz, V.data(), nature, penalty, o, &st);
    isext[r] = 0;
    if (o.noblk == 0 && extrap == 1)
    {
      extrap_receptor(xr[r], yr[r], zr[r], nz, V.data(), nature, penalty, o, &st);
      if (o.noblk > 0) isext[r] = 1;
    }
    noblk[r] = o.noblk; type[r] = o.noblk > 0 ? o.type : 0; dind[r] = o.noblk > 0 ? o.dind : -1;
    for (int k = 0; k < 8; k++) cf[(size_t)r * 8 + k] = o.noblk > 0 ? o.cf[k] : 0.;
  }
  if (stats) { stats[0] = st.nvisit; stats[1] = st.ncand; }
}

double emu_cell_volume(void* h, int ind) { ZoneView V = ((EmuZone*)h)->view(); return cell_volume(V, ind); }

}  // extern "C"
