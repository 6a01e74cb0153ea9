/* cassiopee_b200.h — C ABI of libcassiopee_b200.so (sm_100a CUDA).
 *
 * Drop-in boundary for the chimera interpolation hot path of ONERA Cassiopee's
 * Connector module.  Each entry point names the reference interface it
 * replaces (paths relative to /root/reference/Cassiopee/).  Plain pointers and
 * sizes only; E_Int = int32, E_Float = double, as in the reference's default
 * build (KCore/KCore/Def/DefTypes.h:60-86).  Every function returns 0 on
 * success, a negative status on error (message: cx_last_error()); nothing
 * throws across the ABI.  There is no CPU fallback: without a CUDA device the
 * context cannot be created and every other call needs a context.
 *
 * Host-pointer entry points copy host<->device inside the call (what a
 * reference user sees); *_dev entry points work on device-resident arrays
 * obtained from cx_dev_alloc (what a GPU-resident solver uses every
 * sub-iteration).
 */
#ifndef CASSIOPEE_B200_H
#define CASSIOPEE_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cx_ctx cx_ctx;        /* one CUDA device + stream */
typedef struct cx_zone cx_zone;      /* donor zone resident on the device + its ADT replica */
typedef struct cx_result cx_result;  /* output of one setInterpData call */
typedef struct cx_plan cx_plan;      /* interpData of one (donor zone, receptor zone) pair, device resident */
typedef struct cx_receptors cx_receptors;   /* receptor points of one zone (cellN == 2), device resident */

const char* cx_last_error(void);
int cx_version(void);
int cx_device_count(int* n);

/* ---- context ----------------------------------------------------------- */
int cx_ctx_create(int device, cx_ctx** out);
int cx_ctx_destroy(cx_ctx* ctx);
int cx_ctx_sync(cx_ctx* ctx);
long long cx_ctx_launches(cx_ctx* ctx);              /* kernels launched by this library so far */
int cx_ctx_info(cx_ctx* ctx, int* sm_count, size_t* free_bytes, size_t* total_bytes);
/* CUDA-event timing on the context stream (which = 0 start, 1 stop) */
int cx_event_record(cx_ctx* ctx, int which);
int cx_event_elapsed_ms(cx_ctx* ctx, float* ms);     /* synchronises on the stop event */
int cx_flush_l2(cx_ctx* ctx);                        /* writes a 256 MiB scratch buffer (> 126 MB L2) */

/* ---- raw device / pinned memory (resident SoA fields) -------------------- */
int cx_dev_alloc(cx_ctx* ctx, size_t bytes, void** dptr);
int cx_dev_free(cx_ctx* ctx, void* dptr);
int cx_dev_upload(cx_ctx* ctx, void* dptr, const void* host, size_t bytes);
int cx_dev_download(cx_ctx* ctx, void* host, const void* dptr, size_t bytes);
int cx_dev_memset(cx_ctx* ctx, void* dptr, int value, size_t bytes);
int cx_dev_upload_async(cx_ctx* ctx, void* dptr, const void* host, size_t bytes);   /* ordered by cx_ctx_sync */
int cx_dev_download_async(cx_ctx* ctx, void* host, const void* dptr, size_t bytes);
int cx_host_register(void* hptr, size_t bytes);   /* page-lock a caller-owned range in place; 1 = already registered */
int cx_host_unregister(void* hptr);
int cx_pinned_alloc(size_t bytes, void** hptr);
int cx_pinned_free(void* hptr);

/* ---- donor zone = device ADT hook ---------------------------------------
 * Replaces: new K_INTERP::InterpAdt(...) per donor zone in
 * Connector/Connector/setInterpData.cpp:772-787 (KCore/KCore/Interp/InterpAdt.cpp:83-108,
 * :284-433 buildStructAdt, :506-646 insert) and the hook of
 * Converter/Converter/hook.cpp:578-632 (C.createHook(z,'adt')).
 * x,y,z,cellN: host arrays of ni*nj*nk doubles, i fastest; cellN may be NULL
 * (all 1, as OversetData.py:1193 _addCellN__ does).  nk == 1: 2-D zone in a z = constant plane (quads extruded by 1 in z,
 * InterpAdt.cpp:391-431; results are type 22). */
int cx_zone_create(cx_ctx* ctx, int ni, int nj, int nk, const double* x, const double* y, const double* z,
                   const double* cellN, cx_zone** out);
/* Regular Cartesian donor, interpDataType=0: replaces new K_INTERP::InterpCart(ni,nj,nk,hi,hj,hk,x0,y0,z0)
 * (setInterpData.cpp:790-804; KCore/KCore/Interp/InterpCart.cpp:42-66).  No tree: cells are located in closed
 * form (searchInterpolationCellCartO2/O3, InterpCart.cpp:71-318; searchExtrapolationCellCart :478-571).
 * Like the reference, order 5 finds nothing in such a donor (Interp2.cpp:464-467). */
int cx_zone_create_cart(cx_ctx* ctx, int ni, int nj, int nk, const double* x, const double* y, const double* z,
                        const double* cellN, cx_zone** out);
int cx_zone_set_celln(cx_zone* zone, const double* cellN);
int cx_zone_info(cx_zone* zone, double bbox[6], int* ncells, int* depth, float* build_ms);
/* test/debug: candidate cells of one point in the reference's list order
 * (InterpAdt::getListOfCandidateCells, InterpAdt.cpp:658-929), device traversal */
/* Hooks of a batch of zones: cx_zone_create_noadt uploads coordinates + cellN only, cx_zones_build_adt then builds
 * the ADTs of all of them concurrently on the device (a 10^6-cell tree is ~100 short dependent launches: alone it
 * leaves the GPU idle; eight at a time fill it).  The result is the tree cx_zone_create builds. */
int cx_zone_create_noadt(cx_ctx* ctx, int ni, int nj, int nk, const double* x, const double* y, const double* z,
                         const double* cellN, cx_zone** out);
/* cx_zone_adt_begin starts the build of one such zone on a side stream and returns at once (it then overlaps the upload
 * of the next zones); cx_zones_build_adt finishes the builds in flight and builds the rest. */
int cx_zone_adt_begin(cx_zone* zone);
int cx_zones_build_adt(cx_ctx* ctx, int nzones, cx_zone* const* zones);
int cx_zone_candidates(cx_zone* zone, double x, double y, double z, double alphaTol, int* out, int cap, int* n);
/* test/debug: copy the ADT replica to the host, 128-byte nodes (see cx_core.cuh AdtNode) */
int cx_zone_export_adt(cx_zone* zone, void* nodes_out, size_t bytes);
int cx_zone_destroy(cx_zone* zone);

/* ---- cellN preparers either side of the path ------------
 * cx_apply_bc_overlap: connector.applyBCOverlapStruct(array, imin,jmin,kmin, imax,jmax,kmax, depth, loc, val, cellNName)
 * (Connector/Connector/applyBCOverlaps.cpp:27-125): the `depth` layers next to the 1-based window get `val` unless
 * blanked (cellN == 0); loc 0 = nodes, 1 = centres; cellN: (im,jm,km) doubles, i fastest, in place.
 * cx_hole_fringe: connector._getOversetHolesInterpNodes / _getOversetHolesInterpCellCenters(array, depth, dir, cellNName)
 * (getInterpolatedPoints.cpp:370-650, searchMaskInterpolatedCellsStructOpt): cellN == 1 -> 2 where a point of the
 * stencil (dir 0: cross, dir 1: star; dir 2 / 3 return -4) has cellN == 0; depth < 0 grows the fringe into the
 * blanked region (Connector/Connector/Connector.py:176-182).  cx_apply_bc_overlaps applies all the windows of a zone
 * ([nwin][6], host pointer) with one upload and one download.  The _dev forms work on device arrays
 * (cx_dev_alloc), asynchronously on the context's stream; cx_hole_fringe_dev writes d_out (d_in is read only). */
int cx_apply_bc_overlaps(cx_ctx* ctx, int im, int jm, int km, int nwin, const int* win /* [nwin][6]: imin,jmin,kmin,imax,jmax,kmax */,
                         int depth, int loc, double val, double* cellN);
int cx_apply_bc_overlap_dev(cx_ctx* ctx, int im, int jm, int km, int imin, int jmin, int kmin, int imax, int jmax,
                            int kmax, int depth, int loc, double val, double* d_cellN);
int cx_hole_fringe(cx_ctx* ctx, int im, int jm, int km, int depth, int dir, double* cellN);
int cx_hole_fringe_dev(cx_ctx* ctx, int im, int jm, int km, int depth, int dir, const double* d_in, double* d_out);

/* ---- setInterpData ------------------------------------------------------
 * Replaces: connector.setInterpData(interpPts, donorArrays, order, nature,
 * penalty, extrap, interpDataTypeList, hooks) — Connector/Connector/setInterpData.cpp:585-1322
 * (and below it KCore/KCore/Interp/Interp.cpp:81-165,:282-360, Interp2.cpp:84-665,
 * InterpAdt.cpp:658-1400, InterpData.cpp:56-947, Metric/compVolOfStructCell.cpp:91-218).
 * xr,yr,zr: receptor coordinates (host, nrcv doubles each), zones: ordered
 * donor list (order matters: commonTypesForExtrapAndInterp.h:16).
 * order 2|3|4|5 (setInterpData.cpp:624-652: types 2, 3, 44, 5), nature 0|1, penalty 0|1, extrap 0|1 as in the reference. */
int cx_set_interp_data(cx_ctx* ctx, int nrcv, const double* xr, const double* yr, const double* zr, int nz,
                       cx_zone* const* zones, int order, int nature, int penalty, int extrap, cx_result** out);
/* same with receptor coordinates already on the device */
int cx_set_interp_data_dev(cx_ctx* ctx, int nrcv, const double* d_xr, const double* d_yr, const double* d_zr, int nz,
                           cx_zone* const* zones, int order, int nature, int penalty, int extrap, cx_result** out);
/* Receptor extraction on the device.  Replaces connector.getInterpolatedPoints(array)
 * (Connector/Connector/getInterpolatedPoints.cpp:1675-1732: points with |cellN-2| <= 1e-15, source order,
 * original index 'indcell') and, for loc=1 (centres), K_LOC::node2centerStruct restricted to the selected cells
 * (KCore/KCore/Loc/node2Center.cpp:118-146, same summation order).  x,y,z: host node coordinates (ni*nj*nk);
 * cellN: host, at nodes (loc 0) or at centres (loc 1: (ni-1)*(nj-1)*(nk-1) values). */
int cx_receptors_create(cx_ctx* ctx, int ni, int nj, int nk, const double* x, const double* y, const double* z,
                        const double* cellN, int loc, cx_receptors** out);
/* the same for a zone that is also resident as a donor hook (a tree interpolated from itself, loc = nodes): the hook's
 * node coordinates and node cellN are the receptor zone's own, nothing is uploaded */
int cx_receptors_from_zone(cx_ctx* ctx, cx_zone* zone, cx_receptors** out);
int cx_receptors_count(cx_receptors* rcv, int* n);
/* copies to the host: indcell[n] (index in the zone at loc, ascending) and/or the coordinates; any may be NULL */
int cx_receptors_fetch(cx_receptors* rcv, int* indcell, double* x, double* y, double* z);
int cx_receptors_destroy(cx_receptors* rcv);
/* setInterpData on extracted receptors (receptor positions in the result = positions in indcell) */
int cx_set_interp_data_rcv(cx_ctx* ctx, cx_receptors* rcv, int nz, cx_zone* const* zones, int order, int nature,
                           int penalty, int extrap, cx_result** out);
/* two-phase fetch (sizes, then copy into caller-allocated numpy): the 7-list
 * returned by the reference at setInterpData.cpp:1272 —
 * rcv[n], donor[n], type[n] (int32), coefs (ragged f64), extrap[ne] per donor
 * zone, ascending receptor position; orphan[no] global. */
/* Double wall: connector.setInterpDataDW(receiverArrays, donorArrays, order, nature, penalty, hook)
 * (Connector/Connector/setInterpData.cpp:1346-1890, K_INTERP::getInterpolationCellDW Interp.cpp:176-255): xr[noz], yr[noz],
 * zr[noz] (host, nrcv values each) are the receptor points as donor zone noz must see them (projected onto its walls by
 * the caller); extrapolation is always on; orders 2, 3 and 5 like the reference (4 returns -4).  The result is read like
 * cx_set_interp_data's. */
int cx_set_interp_data_dw(cx_ctx* ctx, int nrcv, const double* const* xr, const double* const* yr,
                          const double* const* zr, int nzones, cx_zone* const* zones, int order, int nature, int penalty,
                          cx_result** out);
/* cx_set_interp_data_rcv / _dev return as soon as the search is queued; the first accessor of the result waits for it.
 * cx_ctx_search_slot(ctx, k), k = 1..8: the searches issued next run on side stream k (those of several receptor
 * zones then overlap on the device); 0: back on the main stream. */
int cx_ctx_search_slot(cx_ctx* ctx, int slot);
int cx_result_sizes(cx_result* res, int* cnt, int* ncoef, int* nextrap, int* norphan);
int cx_result_fetch_zone(cx_result* res, int zone, int* rcv, int* donor, int* type, double* coefs, int* extrap);
int cx_result_fetch_orphans(cx_result* res, int* orphan);
/* every list of the call in one go: rcv / donor / type hold sum(cnt) entries (the zones one behind the other, in call
 * order), coefs sum(ncoef), extrap sum(nextrap), orphan norphan; page-locked destinations make the copies overlap */
int cx_result_fetch_all(cx_result* res, int* rcv, int* donor, int* type, double* coefs, int* extrap, int* orphan);
/* per-receptor table (tests): noblk (0 orphan / 1-based zone), type, donor, isext, cf[ncfmax][nrcv] */
int cx_result_fetch_raw(cx_result* res, int* noblk, int* type, int* dind, int* isext, double* cf);
/* out[0..7]: search kernel ms, extrapolation ms, nodes visited, candidates tested, ncfmax, ... */
int cx_result_stats(cx_result* res, double* out);
int cx_result_destroy(cx_result* res);

/* ---- setInterpTransfers -------------------------------------------------
 * A plan is the device-resident interpData of one ID_* ZoneSubRegion:
 * donor = PointList, rcv = PointListDonor, type = InterpolantsType,
 * coefs = InterpolantsDonor (ragged, SIZECF connector.h:27-37).
 * imd,jmd,kmd: donor zone node dims; nptsR: size of the receptor fields. */
int cx_plan_create(cx_ctx* ctx, int imd, int jmd, int kmd, long long nptsR, int n, const int* donor, const int* rcv,
                   const int* type, const double* coefs, long long ncoefs, cx_plan** out);
/* Replaces connector._setInterpTransfers (Connector/Connector/setInterpTransfers.cpp:259-508,
 * commonInterpTransfers_indirect.h, commonInterpTransfers.h): in place,
 * fieldR[v][rcv[e]] = sum_k cf_k * fieldD[v][stencil_k].  Host arrays. */
int cx_transfer(cx_plan* plan, int nvars, const double* const* fieldD, double* const* fieldR);
/* Replaces connector._setInterpTransfersD (setInterpTransfersD.cpp:106-303):
 * dense out[v*n + e]. Host arrays. */
int cx_transferD(cx_plan* plan, int nvars, const double* const* fieldD, double* out);
/* cx_transfer followed by the periodicity-by-rotation correction of connector._setInterpTransfers
 * (setInterpTransfers.cpp:281-285 dirR/theta from AngleX/Y/Z, :468-472, includeTransfers.h:1-60): the ntrip
 * vector triplets trip[3*t..3*t+2] (indices into the variable list: VelocityX/Y/Z, MomentumX/Y/Z) are rotated
 * at the plan's receptors by theta about axis dirR (1: x, 2: y, 3: z); ctheta/stheta = cos/sin(theta) computed by
 * the caller's libm like the reference.  dirR = 0: plain cx_transfer.  Host arrays. */
int cx_transfer_rot(cx_plan* plan, int nvars, const double* const* fieldD, double* const* fieldR, int dirR,
                    double ctheta, double stheta, int ntrip, const int* trip);
/* A batch of host-pointer transfers = one connector loop over the subregions of a tree (the body of
 * OversetData._setInterpTransfers, Connector/Connector/OversetData.py:1962-2099, which hands the same donor-zone
 * arrays to connector._setInterpTransfers once per ID_* subregion): between cx_batch_begin and cx_batch_end a donor
 * field (identified by its host pointer) is uploaded once and reused by the following cx_transfer / cx_transferD /
 * cx_transfer_rot calls; a host array a transfer writes into is dropped from the cache, so the sequential semantics of
 * the reference loop are kept.  budget_bytes bounds the device memory held (0: default 8 GiB); cx_batch_flush drops
 * the cached fields (between donor zones); cx_batch_end closes the batch and reports the uploads done. */
int cx_batch_begin(cx_ctx* ctx, long long budget_bytes);
int cx_batch_flush(cx_ctx* ctx);
int cx_batch_end(cx_ctx* ctx, long long* uploads);
/* the rotation alone, in place on device-resident receptor fields, at the plan's receptors */
int cx_rotate_dev(cx_plan* plan, int dirR, double ctheta, double stheta, double* d_fx, double* d_fy, double* d_fz);
/* strict cellN rule (commonCellNTransfersStrict.h; setInterpTransfers.cpp:452-467) */
int cx_transfer_celln(cx_plan* plan, const double* cellND, double* cellNR);
/* device-resident variants (pointers from cx_dev_alloc); asynchronous on the
 * context stream */
int cx_transfer_dev(cx_plan* plan, int nvars, const double* const* d_fieldD, double* const* d_fieldR);
int cx_transferD_dev(cx_plan* plan, int nvars, const double* const* d_fieldD, double* d_out, long long ld);
int cx_transfer_celln_dev(cx_plan* plan, const double* d_cellND, double* d_cellNR);
/* receptor-side scatter of a dense (nvars, ld) buffer: fieldR[v][rcv[e]] = in[v*ld+e]
 * (what Connector/Mpi.py:396-440 does with C._setPartialFields on arrival) */
int cx_scatter_dev(cx_ctx* ctx, int nvars, const double* d_in, long long ld, int n, const int* d_rcv,
                   double* const* d_fieldR);
/* Sparse download of a receptor zone (the host tree only needs the values a transfer wrote, what
 * C._setPartialFields(zone, [array], [indices]) stores in Connector/Mpi.py:396-440): device side
 * out[v*ld+e] = field[v][idx[e]], host side dst[idx[e]] = src[e] (distinct indices). */
int cx_gather_dev(cx_ctx* ctx, int nvars, double* d_out, long long ld, int n, const int* d_idx,
                  double* const* d_field);
int cx_host_scatter(double* dst, const int* idx, const double* src, long long n);
/* alg_fixed: sum over entries of 8*ncf+12 bytes; distinct: number of distinct
 * donor nodes referenced (U of SURVEY.md 8(d)); B_alg = alg_fixed + 8*N*(n + U) */
/* receptor indices in the plan's own (regrouped, donor-sorted) entry order = the order of a dense block written
 * by a plan set with mode 2 */
int cx_plan_sorted_rcv(cx_plan* plan, int* out);
int cx_plan_info(cx_plan* plan, int* n, int* ngroups, long long* alg_fixed, long long* distinct);
/* donor points per variable a host-pointer transfer (cx_transfer / cx_transferD) of this plan uploads: for a donor zone
 * of at least 2^20 points of which the plan reads less than 1/4, only the points it reads (gathered on the host into a
 * page-locked block, stored by index on the device); else the whole zone.  CX_SPARSE_H2D=0 disables it. */
int cx_plan_upload_points(cx_plan* plan, long long* npts);
int cx_plan_destroy(cx_plan* plan);

/* ---- batched transfers: every plan of a rank in one launch ------------------
 * A solver calls setInterpTransfers every sub-iteration over ALL ID_* subregions
 * (OversetData.py:2036-2098: Python loop over donor zones x subregions, one C
 * call each).  A plan set fuses that loop into one kernel launch: modes[p]=0
 * writes plan p in place into its receptor zone's fields, modes[p]=1 writes its
 * dense (nvars, ld[p]) block into a send buffer (the _setInterpTransfersD shape
 * the distributed path of Connector/Mpi.py:396-440 needs).  Field pointer
 * lists are device pointers (cx_dev_alloc), nvars per plan. */
typedef struct cx_planset cx_planset;
typedef struct cx_scatterset cx_scatterset;
int cx_planset_create(cx_ctx* ctx, int nplans, cx_plan* const* plans, int nvars, const int* modes,
                      const double* const* const* d_fieldD, double* const* const* d_fieldR, double* const* d_dense,
                      const long long* ld, cx_planset** out);
int cx_planset_run(cx_planset* set);
int cx_planset_destroy(cx_planset* set);
/* receptor-side: scatter nsets received dense blocks into the receptor fields in one launch */
int cx_scatterset_create(cx_ctx* ctx, int nsets, int nvars, const int* n, const int* const* d_rcv,
                         const double* const* d_in, const long long* ld, double* const* const* d_fieldR,
                         cx_scatterset** out);
int cx_scatterset_run(cx_scatterset* set);
int cx_scatterset_destroy(cx_scatterset* set);

/* one solver iteration of a rank: remote plans (dense blocks into the per-peer send
 * buffers) -> grouped NCCL send/recv + scatter on a second stream, overlapped with the
 * local in-place plans; niter iterations are enqueued per call (asynchronous). */
typedef struct cx_step cx_step;
int cx_step_create(cx_ctx* ctx, cx_planset* ps_remote, cx_planset* ps_local, cx_scatterset* ss, int npeers,
                   const int* peers, const double* const* d_send, const long long* sendcounts,
                   double* const* d_recv, const long long* recvcounts, cx_step** out);
/* Peer-memory transport of the same step (cx_p2p.cu): the remote plans store their dense blocks straight into the
 * receptor ranks' receive buffers (exported with CUDA IPC, double-buffered by iteration parity: ps_remote0/1,
 * ss0/1), and a flag per rank pair replaces the message: peer_flags[p] = address in peer p's memory of the
 * flag this rank raises, d_my_flags[p] = the flag peer p raises here.  No NCCL call inside the iteration. */
int cx_step_create_p2p(cx_ctx* ctx, cx_planset* ps_remote0, cx_planset* ps_remote1, cx_planset* ps_local,
                       cx_scatterset* ss0, cx_scatterset* ss1, int npeers, int* const* peer_flags, int* d_my_flags,
                       cx_step** out);
int cx_step_run(cx_step* step, int niter, int overlap);
/* the same restricted to the variables [v0, v0 + nv) of the step's sets (nv < 0: all from v0).  With the peer-memory
 * transport (or a single rank) a sub-range is a complete iteration; the NCCL transport only takes the full range
 * (cx_step_splittable tells). */
int cx_step_run_vars(cx_step* step, int niter, int overlap, int v0, int nv);
int cx_step_splittable(cx_step* step);
int cx_step_destroy(cx_step* step);

/* Host trees in, host trees out: one iteration with its copies, what Connector/Connector/Mpi.py:396-440
 * (__setInterpTransfers on host arrays, then C._setPartialFields of the received values) does per call.  An item is
 * the nvars fields of one zone in one direction (dir 0: host -> device before the plans run, 1: device -> host after):
 * whole arrays of n points (h_idx == NULL) or only the n listed points (packed through a page-locked block; the donor
 * points a step reads come from cx_plan_mark_donors).  cx_step_run_host pipelines per variable over three streams
 * (upload of v+1 | plans + exchange of v | download of v-1) and returns when the host arrays hold the result. */
typedef struct cx_hostio cx_hostio;
int cx_hostio_create(cx_ctx* ctx, int nvars, cx_hostio** out);
int cx_hostio_add(cx_hostio* io, int dir, long long n, const int* h_idx, double* const* h_field, double* const* d_field);
/* a whole-array item cut to the box [i0,i1] x [j0,j1] x [k0,k1] (inclusive) of (imd, jmd, kmd) arrays: one strided copy per variable */
int cx_hostio_add_box(cx_hostio* io, int dir, int imd, int jmd, int kmd, const int* box6, double* const* h_field,
                      double* const* d_field);
int cx_hostio_bytes(cx_hostio* io, long long* h2d_bytes, long long* d2h_bytes);
int cx_step_run_host(cx_step* step, cx_hostio* io, int pipelined);
int cx_hostio_destroy(cx_hostio* io);
/* mark[node] = 1 for every donor node the plan reads; d_mark: nptsD bytes on the device, zeroed by the caller */
int cx_plan_mark_donors(cx_plan* plan, unsigned char* d_mark);
/* dst[e] = src[idx[e]] on the host (OpenMP team): the host side of a sparse upload */
int cx_host_gather(double* dst, const int* idx, const double* src, long long n);

/* ---- multi-GPU (one process per GPU; NCCL over NVLink) -------------------
 * Replaces the mpi4py pickled isend/recv of Converter/Converter/Mpi4py.py:149-168
 * used by Connector/Mpi.py:396-440: one packed device buffer per peer rank. */
/* CUDA IPC: 72-byte descriptor (cudaIpcMemHandle_t of the containing allocation + offset) of a device pointer,
 * and its mapping in another process of the same node */
int cx_ipc_export(void* dptr, void* handle72);
int cx_ipc_open(cx_ctx* ctx, const void* handle72, void** dptr);
int cx_ipc_close(const void* handle72);   /* unmap; the exporter may free the memory only after this */
int cx_ipc_close_all(void);
int cx_comm_unique_id(char* id128);
int cx_comm_init(cx_ctx* ctx, int rank, int nranks, const char* id128);
int cx_comm_destroy(cx_ctx* ctx);
int cx_comm_barrier(cx_ctx* ctx);
/* grouped ncclSend/ncclRecv: counts in doubles; device buffers */
int cx_exchange_dev(cx_ctx* ctx, int npeers, const int* peers, const double* const* d_send,
                    const long long* sendcounts, double* const* d_recv, const long long* recvcounts);
/* same for raw bytes (one-off shipping of donor coordinates and index lists) */
int cx_exchange_bytes_dev(cx_ctx* ctx, int npeers, const int* peers, const void* const* d_send,
                          const long long* sendbytes, void* const* d_recv, const long long* recvbytes);
/* Device-to-device replication of donor zones, the device form of Cmpi._addXZones(aD, graph, variables=['cellN'])
 * (Converter/Converter/Mpi.py; called by Connector/Connector/Mpi.py:941 before the local setInterpData calls):
 * the resident (x, y, z, cellN) block and the ADT replica of each listed zone travel over NCCL; the receiver gets
 * ready-to-query zone handles (no host round trip, no tree rebuild).  cx_zone_meta exports what the receiver
 * needs beside the bulk data (meta_i[6]: ni, nj, nk, depth, levels, topo; meta_d[9]: bbox[6], h[3]); the caller
 * exchanges it.  Between one pair of ranks both sides must list the zones in the same order. */
int cx_zone_meta(cx_zone* zone, int* meta_i, double* meta_d);
int cx_zone_exchange(cx_ctx* ctx, int nsend, cx_zone* const* send_zones, const int* send_peers, int nrecv,
                     const int* recv_peers, const int* meta_i, const double* meta_d, cx_zone** out);

#ifdef __cplusplus
}
#endif
#endif
