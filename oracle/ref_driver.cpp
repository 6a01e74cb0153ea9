/* ORACLE — TEST INFRASTRUCTURE ONLY. Never imported, linked or called by the
 * product path (cassiopee_b200/). Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it, as the checker.
 *
 * Thin extern "C" driver around the reference's OWN compiled C++:
 * oracle/Makefile compiles KCore/KCore/Interp/*.cpp, Metric/compVolOfStructCell.cpp,
 * CompGeom/compGeom.cpp and Def/DefCplusPlusConst.cpp unchanged from
 * /root/reference into oracle/_ref/libcassref.so together with this file.
 * Nothing of the algorithm lives here: this file only
 *   (1) loops over receptors exactly like
 *       Connector/Connector/setInterpData.cpp:978-1114 (getInterpolationCell,
 *       then getExtrapolationCell when allowed) and returns the per-receptor
 *       result instead of per-thread chunk buffers (the chunk concatenation at
 *       :1183-1252 yields ascending receptor order per donor zone, which the
 *       Python side re-creates with a stable selection);
 *   (2) runs the reference transfer loops by #including the reference's own
 *       fragment headers commonInterpTransfers_{indirect,direct}.h /
 *       commonInterpTransfers.h / commonCellNTransfersStrict.h with the local
 *       variable names those fragments expect
 *       (Connector/Connector/setInterpTransfers.cpp:285-467,
 *        setInterpTransfersD.cpp:106-303);
 *   (3) exposes the ADT candidate list for debugging the device ADT replica.
 * It bypasses the CPython/numpy marshalling of KCore (not buildable here
 * without scons+gfortran).
 */
#include <vector>
#include <list>
#include <algorithm>
#include <cstring>
#include <omp.h>
#include "Interp/Interp.h"
#include "Interp/InterpCart.h"
#include "Metric/metric.h"
#include "CompGeom/compGeom.h"
#include "parallel.h"

using namespace std;
using namespace K_FLD;

// Connector/Connector/getInterpolatedPoints.cpp:370 (C++ linkage: declared outside the extern "C" block)
namespace K_CONNECTOR {
void searchMaskInterpolatedCellsStructOpt(E_Int imc, E_Int jmc, E_Int kmc, E_Int depth, E_Int dir,
                                          E_Float* cellN, E_Float* cellN_tmp);
}


/* SIZECF: Connector/Connector/connector.h:27-37 (pulled by macro re-statement
 * because connector.h drags the whole Connector module in). */
#define SIZECF(type, meshtype, sizecf)                     \
  {                                                        \
    if (type == 1) sizecf = 1;                             \
    else if (type == 2 && meshtype == 1) sizecf = 8;       \
    else if (type == 3 && meshtype == 1) sizecf = 9;       \
    else if (type == 44 && meshtype == 1) sizecf = 12;     \
    else if (type == 4 && meshtype == 1) sizecf = 8;       \
    else if (type == 4 && meshtype == 2) sizecf = 4;       \
    else if (type == 5 && meshtype == 1) sizecf = 15;      \
    else if (type == 22 && meshtype == 1) sizecf = 4;      \
    else sizecf = -1;                                      \
  }

struct RefZone
{
  E_Int ni, nj, nk;
  FldArrayF* f;  // (npts, 4): x, y, z, cellN
  K_INTERP::InterpAdt* adt;
  K_INTERP::InterpCart* cart;   // interpDataType 0 (setInterpData.cpp:790-804)
  K_INTERP::InterpData* data() { return adt ? (K_INTERP::InterpData*)adt : (K_INTERP::InterpData*)cart; }
};

extern "C" {

int ref_num_threads() { return omp_get_max_threads(); }
void ref_set_num_threads(int n) { omp_set_num_threads(n); }

/* Donor zone = coordinates + cellN + ADT (setInterpData.cpp:772-787). */
void* ref_zone_create(int ni, int nj, int nk, const double* x, const double* y,
                      const double* z, const double* cellN, int buildAdt)
{
  RefZone* zn = new RefZone;
  zn->ni = ni; zn->nj = nj; zn->nk = nk;
  E_Int npts = ni * nj * nk;
  zn->f = new FldArrayF(npts, 4);
  memcpy(zn->f->begin(1), x, sizeof(double) * npts);
  memcpy(zn->f->begin(2), y, sizeof(double) * npts);
  memcpy(zn->f->begin(3), z, sizeof(double) * npts);
  if (cellN) memcpy(zn->f->begin(4), cellN, sizeof(double) * npts);
  else for (E_Int i = 0; i < npts; i++) (*zn->f)(i, 4) = 1.;
  zn->adt = NULL; zn->cart = NULL;
  if (buildAdt == 2)
  {
    /* CART branch of setInterpData.cpp:790-804, verbatim arithmetic */
    E_Float* xt = zn->f->begin(1); E_Float* yt = zn->f->begin(2); E_Float* zt = zn->f->begin(3);
    E_Float x0 = xt[0]; E_Float y0 = yt[0]; E_Float z0 = zt[0];
    E_Float hi = xt[1]-xt[0];
    E_Float hj = yt[ni]-yt[0];
    E_Float hk = zt[ni*nj]-zt[0];
    zn->cart = new K_INTERP::InterpCart(ni,nj,nk,hi,hj,hk,x0,y0,z0);
  }
  else if (buildAdt)
  {
    E_Int built = 0;
    zn->adt = new K_INTERP::InterpAdt(npts, zn->f->begin(1), zn->f->begin(2),
                                      zn->f->begin(3), &zn->ni, &zn->nj,
                                      &zn->nk, built);
    if (built != 1) { delete zn->adt; zn->adt = NULL; }
  }
  return zn;
}

int ref_zone_has_adt(void* h) { return ((RefZone*)h)->data() != NULL; }

void ref_zone_free(void* h)
{
  RefZone* zn = (RefZone*)h;
  if (zn->adt) delete zn->adt;
  if (zn->cart) delete zn->cart;
  delete zn->f;
  delete zn;
}

void ref_zone_bbox(void* h, double* bb)
{
  K_INTERP::InterpData* a = ((RefZone*)h)->data();
  bb[0] = a->_xmin; bb[1] = a->_ymin; bb[2] = a->_zmin;
  bb[3] = a->_xmax; bb[4] = a->_ymax; bb[5] = a->_zmax;
}

/* InterpAdt::getListOfCandidateCells (InterpAdt.cpp:658-929). */
int ref_candidates(void* h, double x, double y, double z, double alphaTol,
                   int* out, int cap)
{
  list<E_Int> l;
  ((RefZone*)h)->adt->getListOfCandidateCells(x, y, z, l, alphaTol);
  int n = 0;
  for (list<E_Int>::iterator it = l.begin(); it != l.end(); ++it, ++n)
    if (n < cap) out[n] = *it;
  return n;
}

/* Per-receptor loop of K_CONNECTOR::setInterpData (setInterpData.cpp:978-1114).
 * Outputs, one entry per receptor:
 *   noblk[n]   0 = orphan, else 1-based donor zone
 *   type[n]    InterpolantsType (1,2,3,5,22,44)
 *   dind[n]    donor index
 *   isext[n]   1 when produced by getExtrapolationCell
 *   cf[n*16+k] coefficients (ncfmax<=15 used)
 *   vol[n]     selection volume
 */
int ref_set_interp_data(int nrcv, const double* xr, const double* yr,
                        const double* zr, int nz, void** zones, int order,
                        int nature, int penalty, int extrap, int* noblk_o,
                        int* type_o, int* dind_o, int* isext_o, double* cf_o,
                        double* vol_o)
{
  K_INTERP::InterpData::InterpolationType interpType;
  E_Int nindi = 1, ncfmax;
  switch (order)  // setInterpData.cpp:624-652
  {
    case 2: interpType = K_INTERP::InterpData::O2CF; ncfmax = 8; break;
    case 3: interpType = K_INTERP::InterpData::O3ABC; ncfmax = 9; break;
    case 4: interpType = K_INTERP::InterpData::O4ABC; ncfmax = 12; break;
    case 5: interpType = K_INTERP::InterpData::O5ABC; ncfmax = 15; break;
    default: interpType = K_INTERP::InterpData::O2CF; ncfmax = 8; break;
  }
  vector<K_INTERP::InterpData*> interpDatas;
  vector<FldArrayF*> fields;
  vector<void*> a2, a3, a4, a5;
  vector<E_Int> posxs, posys, poszs, poscs;
  for (int i = 0; i < nz; i++)
  {
    RefZone* zn = (RefZone*)zones[i];
    if (!zn->data()) return -1;
    interpDatas.push_back(zn->data());
    fields.push_back(zn->f);
    a2.push_back(&zn->ni); a3.push_back(&zn->nj); a4.push_back(&zn->nk);
    a5.push_back(NULL);
    posxs.push_back(1); posys.push_back(2); poszs.push_back(3); poscs.push_back(4);
  }
  E_Int nbI = nrcv;
#pragma omp parallel
  {
    FldArrayI indi(nindi); FldArrayF cf(ncfmax);
    FldArrayI tmpIndi(nindi); FldArrayF tmpCf(ncfmax);
    E_Float x, y, z, vol;
    E_Int noblk = 0;
    short ok;
    E_Int type, isExtrapolated;
#pragma omp for
    for (E_Int ind = 0; ind < nbI; ind++)
    {
      x = xr[ind]; y = yr[ind]; z = zr[ind];
      type = 0;
      ok = K_INTERP::getInterpolationCell(
        x, y, z, interpDatas, fields, a2, a3, a4, a5, posxs, posys, poszs,
        poscs, vol, indi, cf, tmpIndi, tmpCf, type, noblk, interpType, nature,
        penalty);
      isExtrapolated = 0;
      if (ok != 1 && extrap == 1)
      {
        ok = K_INTERP::getExtrapolationCell(
          x, y, z, interpDatas, fields, a2, a3, a4, a5, posxs, posys, poszs,
          poscs, vol, indi, cf, type, noblk, interpType, nature, penalty);
        if (noblk > 0) isExtrapolated = 1;
      }
      noblk_o[ind] = noblk;
      isext_o[ind] = isExtrapolated;
      vol_o[ind] = vol;
      for (int k = 0; k < 16; k++) cf_o[(size_t)ind * 16 + k] = 0.;
      if (noblk > 0)
      {
        type_o[ind] = type;
        dind_o[ind] = indi[0];
        E_Int ncf = cf.getSize() < ncfmax ? cf.getSize() : ncfmax;
        for (E_Int k = 0; k < ncf; k++) cf_o[(size_t)ind * 16 + k] = cf[k];
      }
      else { type_o[ind] = 0; dind_o[ind] = -1; }
    }
  }
  return 0;
}

/* Double wall: the receptor loop of K_CONNECTOR::setInterpDataDW (Connector/Connector/setInterpData.cpp:1606-1625) on raw
 * pointers: receptor ind seen from donor zone noz is (xt[noz][ind], yt[noz][ind], zt[noz][ind]); the reference's own
 * K_INTERP::getInterpolationCellDW / getExtrapolationCellDW (Interp.cpp:176-255, :371-) do the work; extrapolation is
 * always tried.  Same per-receptor tables as ref_set_interp_data (the packing, with its type-1 slot rule, is
 * oracle.py:set_interp_data_dw). */
int ref_set_interp_data_dw(int nrcv, const double* const* xt_, const double* const* yt_, const double* const* zt_,
                           int nz, void** zones, int order, int nature, int penalty, int* noblk_o, int* type_o,
                           int* dind_o, int* isext_o, double* cf_o)
{
  K_INTERP::InterpData::InterpolationType interpType;
  E_Int nindi = 1, ncfmax;
  switch (order)  // setInterpData.cpp:1366-1391
  {
    case 2: interpType = K_INTERP::InterpData::O2CF; ncfmax = 8; break;
    case 3: interpType = K_INTERP::InterpData::O3ABC; ncfmax = 9; break;
    case 5: interpType = K_INTERP::InterpData::O5ABC; ncfmax = 15; break;
    default: interpType = K_INTERP::InterpData::O2CF; ncfmax = 8; break;
  }
  vector<K_INTERP::InterpData*> interpDatas;
  vector<FldArrayF*> fields;
  vector<void*> a2, a3, a4, a5;
  vector<E_Int> posxs, posys, poszs, poscs;
  for (int i = 0; i < nz; i++)
  {
    RefZone* zn = (RefZone*)zones[i];
    if (!zn->data()) return -1;
    interpDatas.push_back(zn->data());
    fields.push_back(zn->f);
    a2.push_back(&zn->ni); a3.push_back(&zn->nj); a4.push_back(&zn->nk);
    a5.push_back(NULL);
    posxs.push_back(1); posys.push_back(2); poszs.push_back(3); poscs.push_back(4);
  }
  FldArrayI indi(nindi); FldArrayF cf(ncfmax);
  vector<E_Float> xt(nz), yt(nz), zt(nz);
  E_Float vol; E_Int type, noblk = 0;
  for (E_Int ind = 0; ind < nrcv; ind++)
  {
    for (E_Int noz = 0; noz < nz; noz++) { xt[noz] = xt_[noz][ind]; yt[noz] = yt_[noz][ind]; zt[noz] = zt_[noz][ind]; }
    type = 0;
    short ok = K_INTERP::getInterpolationCellDW(&xt[0], &yt[0], &zt[0], interpDatas, fields, a2, a3, a4, a5, posxs,
                                                posys, poszs, poscs, vol, indi, cf, type, noblk, interpType, nature,
                                                penalty);
    E_Int isExtrapolated = 0;
    if (ok != 1)
    {
      ok = K_INTERP::getExtrapolationCellDW(&xt[0], &yt[0], &zt[0], interpDatas, fields, a2, a3, a4, a5, posxs, posys,
                                            poszs, poscs, vol, indi, cf, type, noblk, interpType, nature, penalty);
      if (noblk > 0) isExtrapolated = 1;
    }
    noblk_o[ind] = noblk;
    isext_o[ind] = isExtrapolated;
    for (int k = 0; k < 16; k++) cf_o[(size_t)ind * 16 + k] = 0.;
    if (noblk > 0)
    {
      type_o[ind] = type;
      dind_o[ind] = indi[0];
      E_Int ncf = cf.getSize() < ncfmax ? cf.getSize() : ncfmax;
      for (E_Int k = 0; k < ncf; k++) cf_o[(size_t)ind * 16 + k] = cf[k];
    }
    else { type_o[ind] = 0; dind_o[ind] = -1; }
  }
  return 0;
}

/* _setInterpTransfers general path (setInterpTransfers.cpp:436-449) on raw
 * pointers: fieldR[v][rcvPts[n]] = sum cf*fieldD[v][stencil]. */
void ref_transfer(int nvars_, double** dnrFields, double** rcvFields, int imd_,
                  int jmd_, int nrcv, const int* rcvPts_, const int* donorPts_,
                  const int* types_, const double* coefs, int ncoefs)
{
  E_Int imd = imd_, jmd = jmd_, imdjmd, nvars = nvars_, meshtype = 1;
  E_Int nbRcvPts = nrcv;
  E_Int* rcvPts = (E_Int*)rcvPts_;
  E_Int* donorPts = (E_Int*)donorPts_;
  E_Int* types = (E_Int*)types_;
  E_Int* ptrcnd = NULL; E_Int cnNfldD = 0;
  FldArrayF donorCoefs(ncoefs > 0 ? ncoefs : 1);
  if (ncoefs > 0) memcpy(donorCoefs.begin(), coefs, sizeof(double) * ncoefs);
  FldArrayF* donorCoefsF = &donorCoefs;
  vector<E_Float*> vectOfRcvFields(nvars), vectOfDnrFields(nvars);
  for (E_Int eq = 0; eq < nvars; eq++)
  {
    vectOfRcvFields[eq] = rcvFields[eq];
    vectOfDnrFields[eq] = dnrFields[eq];
  }
  (void)ptrcnd; (void)cnNfldD;
# include "commonInterpTransfers_indirect.h"
}

/* _setInterpTransfersD (setInterpTransfersD.cpp:106-303): dense (nvars,nrcv)
 * output, indR = noind. */
void ref_transferD(int nvars_, double** dnrFields, double* out, int imd_,
                   int jmd_, int nrcv, const int* donorPts_, const int* types_,
                   const double* coefs, int ncoefs)
{
  E_Int imd = imd_, jmd = jmd_, imdjmd, nvars = nvars_, meshtype = 1;
  E_Int nbRcvPts = nrcv;
  E_Int* donorPts = (E_Int*)donorPts_;
  E_Int* types = (E_Int*)types_;
  E_Int* ptrcnd = NULL; E_Int cnNfldD = 0;
  FldArrayF donorCoefs(ncoefs > 0 ? ncoefs : 1);
  if (ncoefs > 0) memcpy(donorCoefs.begin(), coefs, sizeof(double) * ncoefs);
  FldArrayF* donorCoefsF = &donorCoefs;
  vector<E_Float*> vectOfRcvFields(nvars), vectOfDnrFields(nvars);
  for (E_Int eq = 0; eq < nvars; eq++)
  {
    vectOfRcvFields[eq] = out + (size_t)eq * nrcv;
    vectOfDnrFields[eq] = dnrFields[eq];
  }
  (void)ptrcnd; (void)cnNfldD;
# include "commonInterpTransfers_direct.h"
}

/* Strict cellN transfer (setInterpTransfers.cpp:452-467). */
void ref_transfer_celln(double* cellND, double* cellNR, int imd_, int jmd_,
                        int nrcv, const int* rcvPts_, const int* donorPts_,
                        const int* types_, const double* coefs)
{
  E_Int imd = imd_, jmd = jmd_, imdjmd = imd * jmd, meshtype = 1;
  E_Int nbRcvPts = nrcv;
  E_Int* rcvPts = (E_Int*)rcvPts_;
  E_Int* donorPts = (E_Int*)donorPts_;
  E_Int* types = (E_Int*)types_;
  E_Int* ptrcnd = NULL; E_Int cnNfldD = 0;
  E_Int indR, type, nocf;
  E_Int indD0, indD, ncfLoc;
  E_Int noi = 0;
  E_Int sizecoefs = 0;
  E_Float* ptrCoefs = (E_Float*)coefs;
  (void)ptrcnd; (void)cnNfldD; (void)ncfLoc;
  for (E_Int noind = 0; noind < nbRcvPts; noind++)
  {
    indR = rcvPts[noind];
# include "commonCellNTransfersStrict.h"
  }
}

/* Periodicity by rotation after a transfer (setInterpTransfers.cpp:281-285,
 * :468-472): the reference's own fragment includeTransfers.h on raw pointers.
 * posv* / posm*: positions of VelocityX/Y/Z and MomentumX/Y/Z (-1: absent). */
void ref_rotate(int nvars_, double** rcvFields, int nrcv, const int* rcvPts_,
                double AngleX, double AngleY, double AngleZ, int posvx_, int posvy_,
                int posvz_, int posmx_, int posmy_, int posmz_)
{
  E_Int nvars = nvars_, nbRcvPts = nrcv;
  E_Int* rcvPts = (E_Int*)rcvPts_;
  E_Int posvx = posvx_, posvy = posvy_, posvz = posvz_, posmx = posmx_, posmy = posmy_, posmz = posmz_;
  vector<E_Float*> vectOfRcvFields(nvars);
  for (E_Int eq = 0; eq < nvars; eq++) vectOfRcvFields[eq] = rcvFields[eq];
  E_Int dirR = 0; E_Float theta = 0.;
  if (K_FUNC::E_abs(AngleX) > 0.) {dirR = 1; theta=AngleX;}
  else if (K_FUNC::E_abs(AngleY) > 0.){dirR=2; theta=AngleY;}
  else if (K_FUNC::E_abs(AngleZ) > 0.) {dirR=3; theta=AngleZ;}
  if (dirR != 0)
  {
# include "includeTransfers.h"
  }
}

/* ___setInterpTransfers (setInterpTransfers.cpp:761-1410), the transfer the Fast solver calls on the flattened
 * donor tree (param_int / param_real of compactTransfers.py:miseAPlatDonorTree__): the reader loop restated for
 * chimera (ID) racs of structured, cell-centred zones with 5 or 6 equations, steady racs, implicit / global
 * explicit time stepping (exploc == 0); the per-type arithmetic is the reference's own fragments
 * commonInterpTransfers_reorder_5eq.h / _6eq.h, #included.  Sequential (one "thread": pt_deb = ideb, pt_fin = ifin;
 * the reference only splits each type range over its OpenMP threads).
 *   roR[nd], roD[nd]  first variable of zone nd's compact solution (receptor tree centres / donor tree nodes)
 *   zparam[nd]        zone nd's .Solver#ownData/Parameter_int (NIJK, NDIMDX, IFLOW, LEVEL are read)
 *   threadmax_sdm     __NUMTHREADS__ of the build that wrote dtloc (stride of its task records)
 * Returns the number of racs transferred, or -1 for a rac outside that scope. */
#define REF_NIJK 0
#define REF_IFLOW 27
#define REF_NDIMDX 41
#define REF_LEVEL 53
int ref_transfer_compact(int nidomR, double** roR, double** roD, int** zparam, const int* iptdtloc,
                         const int* ipt_param_int, const double* ipt_param_real, int vartype, int type_transfert,
                         int no_transfert, int nstep, int threadmax_sdm)
{
  E_Int TypeTransfert = type_transfert;
  E_Int pass_deb, pass_fin;
  if     (TypeTransfert==0) { pass_deb =1; pass_fin =2; }//ID
  else if(TypeTransfert==1) { pass_deb =0; pass_fin =1; }//IBCD
  else                      { pass_deb =0; pass_fin =2; }//ALL
  E_Int NoTransfert = no_transfert;
  E_Int nvars;
  if (vartype <= 3 && vartype >= 1) nvars = 5; else nvars = 6;

  E_Int nbcomIBC = ipt_param_int[2];
  E_Int nbcomID  = ipt_param_int[3+nbcomIBC];
  E_Int shift_graph = nbcomIBC + nbcomID + 3;
  E_Int ech            = ipt_param_int[ NoTransfert +shift_graph];
  E_Int nrac           = ipt_param_int[ ech +1 ];
  E_Int nrac_inst      = ipt_param_int[ ech +2 ];
  E_Int timelevel      = ipt_param_int[ ech +3 ];
  E_Int nrac_steady    = nrac - nrac_inst;
  if (nrac_inst > 0) return -1;

  std::vector<E_Int> impli_local(nidomR, 0);
  E_Int nssiter  = iptdtloc[0];
  E_Int shift_omp= iptdtloc[11];
  const E_Int* ipt_omp = iptdtloc + shift_omp;
  E_Int nbtask   = ipt_omp[nstep-1];
  E_Int ptiter   = ipt_omp[nssiter+ nstep-1];
  for (E_Int ntask = 0; ntask < nbtask; ntask++)
  {
    E_Int pttask = ptiter + ntask*(6+threadmax_sdm*7);
    E_Int nd = ipt_omp[ pttask ];
    impli_local[nd] = 1;
  }
  E_Int maxlevel      =  iptdtloc[ 9];
  E_Int it_cycl_lbm   =  iptdtloc[10];
  E_Int max_it        = pow(2, maxlevel-1);
  E_Int level_next_it =  maxlevel;
  if (it_cycl_lbm != max_it -1 ) { level_next_it = iptdtloc[13 +it_cycl_lbm +1];}
  for (E_Int nd = 0; nd < nidomR; nd++)
  {
    if (zparam[nd][REF_IFLOW] == 4)
    {
      E_Int level = zparam[nd][REF_LEVEL];
      if (level <=level_next_it +1 ){impli_local[nd]=1;}
      else                          {impli_local[nd]=0;}
    }
  }

  std::vector<E_Float*> RcvFields(nvars), DnrFields(nvars);
  E_Float** vectOfRcvFields = RcvFields.data();
  E_Float** vectOfDnrFields = DnrFields.data();
  E_Int done = 0;
  E_Int indR, type;
  E_Int indD0, indD, ncfLoc, indCoef, noi, sizecoefs, imd, jmd, imdjmd;
  E_Int meshtype = 1; E_Int cnNfldD = 0; E_Int* ptrcnd = NULL;
  (void)ncfLoc; (void)indD; (void)indD0; (void)cnNfldD; (void)ptrcnd;
  for (E_Int ipass_typ=pass_deb; ipass_typ < pass_fin; ipass_typ++)
  {
    for (E_Int irac=0; irac< nrac_steady; irac++)
    {
      E_Int shift_rac =  ech + 4 + timelevel*2 + irac;
      E_Int NoD0 =  ipt_param_int[ shift_rac + nrac*5 ];
      if (impli_local[NoD0] != 1) continue;                  // autorisation_transferts (:1033-1038)
      E_Int ibcType =  ipt_param_int[shift_rac+nrac*3];
      E_Int ibc = 1;
      if (ibcType < 0) ibc = 0;
      if(1-ibc != ipass_typ)  continue;
      if (ibc == 1) return -1;
      E_Int NoD      =  ipt_param_int[ shift_rac + nrac*5  ];
      E_Int loc      =  ipt_param_int[ shift_rac + nrac*9  ];
      E_Int NoR      =  ipt_param_int[ shift_rac + nrac*11 ];
      E_Int nvars_loc=  ipt_param_int[ shift_rac + nrac*13 ];
      E_Int rotation =  ipt_param_int[ shift_rac + nrac*14 ];
      if (loc == 0 || rotation != 0 || !(nvars_loc == 5 || nvars_loc == 6) || nvars_loc > nvars) return -1;
      for (E_Int eq = 0; eq < nvars_loc; eq++)
      {
        vectOfRcvFields[eq] = roR[ NoR] + eq*zparam[ NoR][ REF_NDIMDX ];
        vectOfDnrFields[eq] = roD[ NoD] + eq*zparam[ NoD][ REF_NDIMDX ];
      }
      imd= zparam[ NoD ][ REF_NIJK ]; jmd= zparam[ NoD ][ REF_NIJK+1];
      imdjmd = imd*jmd;

      E_Int pos;
      pos  = ipt_param_int[ shift_rac + nrac*7 ]; const E_Int* ntype    = ipt_param_int  + pos;
      pos  = pos +1 + ntype[0]                  ; const E_Int* types    = ipt_param_int  + pos;
      pos  = ipt_param_int[ shift_rac + nrac*6 ]; const E_Int* donorPts = ipt_param_int  + pos;
      pos  = ipt_param_int[ shift_rac + nrac*12]; const E_Int* rcvPts   = ipt_param_int  + pos;
      pos  = ipt_param_int[ shift_rac + nrac*8 ]; const E_Float* ptrCoefs = ipt_param_real + pos;

      E_Int ideb      = 0;
      E_Int ifin      = 0;
      E_Int shiftCoef = 0;
      E_Int shiftDonor  = 0;
      for (E_Int ndtyp = 0; ndtyp < ntype[0]; ndtyp++)
      {
        type      = types[ifin];
        SIZECF(type, meshtype, sizecoefs);
        ifin =  ifin + ntype[ 1 + ndtyp];
        E_Int pt_deb = ideb, pt_fin = ifin;
        noi       = shiftDonor;
        indCoef   = (pt_deb-ideb)*sizecoefs +  shiftCoef;
        E_Int shiftv  =0;
        if (nvars_loc==5)
        {
#           include "commonInterpTransfers_reorder_5eq.h"
        }
        else
        {
#           include "commonInterpTransfers_reorder_6eq.h"
        }
        ideb       =  ideb + ntype[ 1 + ndtyp];
        shiftCoef  = shiftCoef  +  ntype[1+ndtyp]*sizecoefs;
        shiftDonor = shiftDonor +  ntype[1+ndtyp];
      }
      done++;
    }
  }
  return done;
}

/* setHoleInterpolatedPoints on one structured cellN array: the reference's OWN routine
 * (K_CONNECTOR::searchMaskInterpolatedCellsStructOpt, Connector/Connector/getInterpolatedPoints.cpp:370-, compiled
 * unchanged), called the way _getOversetHolesInterpNodes / CellCenters call it (:1124-1144): on a copy, then copied
 * back.  dir 0..3. */
void ref_hole_fringe(int im, int jm, int km, int depth, int dir, double* cellN)
{
  size_t n = (size_t)im * jm * km;
  std::vector<double> tmp(cellN, cellN + n);
  K_CONNECTOR::searchMaskInterpolatedCellsStructOpt(im, jm, km, depth, dir, cellN, tmp.data());
  memcpy(cellN, tmp.data(), sizeof(double) * n);
}

/* connector.applyBCOverlapStruct (Connector/Connector/applyBCOverlaps.cpp:27-125) without its Python argument
 * parsing: the window arithmetic (:69-96) and the loop (:98-112) restated statement by statement.  Returns -1 for
 * the windows the reference refuses (:61-66, :73-78). */
int ref_apply_bc_overlap(int im, int jm, int km, int imin, int jmin, int kmin, int imax, int jmax, int kmax,
                         int depth, int loc, int cellNInterpValue, double* cellNt)
{
  E_Int shift = 0;
  if (loc == 0) shift = 1; // loc='nodes'
  E_Float interpolatedValue = cellNInterpValue;
  if (loc == 0) // nodes
  {
    if (imin < 1 || imax > im || jmin < 1 || jmax > jm || kmin < 1 || kmax > km) return -1;
  }
  else // centers
  {
    if (imin < 1 || imax > im+1 || jmin < 1 || jmax > jm+1 || kmin < 1 || kmax > km+1) return -1;
  }
  if (imin == imax)
  {
    if (imin == 1) {imin = 1; imax = K_FUNC::E_min(depth,im);}
    else { imin =  K_FUNC::E_max(imin-depth+shift,1); imax = imax-1+shift;}
  }
  else
  {imax = imax-1+shift;}
  if (jmin == jmax )
  {
    if (jmin == 1) {jmin = 1; jmax = K_FUNC::E_min(depth, jm);}
    else {jmin = K_FUNC::E_max(jmin-depth+shift,1); jmax = jmax-1+shift;}
  }
  else
  {jmax = jmax-1+shift;}
  if (kmin == kmax)
  {
    if (kmin == 1) { kmin = 1 ; kmax = K_FUNC::E_min(depth, km);}
    else {kmin = K_FUNC::E_max(kmin-depth+shift,1); kmax = kmax-1+shift;}
  }
  else {kmax = kmax-1+shift;}
  E_Int imjm = im*jm;
  for (E_Int k = kmin; k <= kmax; k++)
    for (E_Int j = jmin; j <= jmax; j++)
      for (E_Int i = imin; i <= imax; i++)
      {
        E_Int ind = (i-1) + (j-1)* im + (k-1)*imjm;
        if (cellNt[ind] != 0.) cellNt[ind] = interpolatedValue;
      }
  return 0;
}

/* K_METRIC::compVolOfStructCell3D (selection volume, for unit checks). */
double ref_cell_volume(void* h, int indnode)
{
  RefZone* zn = (RefZone*)h;
  E_Float vol;
  K_METRIC::compVolOfStructCell3D(zn->ni, zn->nj, zn->nk, -1, indnode,
                                  zn->f->begin(1), zn->f->begin(2),
                                  zn->f->begin(3), vol);
  return vol;
}

}  // extern "C"
