/* eqlb_b200.h - C ABI of the B200-native flux-equilibration hot path.
 *
 * Drop-in boundary: these entry points are what the successor of the reference's
 * pybind11 module (`python/dolfinx_eqlb/wrappers.cpp`, module `dolfinx_eqlb.cpp`)
 * binds.  The pybind layer extracts plain arrays from the DOLFINx objects it is
 * handed and calls into this library; nothing above it changes.
 *
 *   reference entry point (wrappers.cpp)                    replaced by
 *   ------------------------------------------------------  -----------------------
 *   reconstruct_fluxes_semiexplt             :97-115        eqlb_se_run
 *   reconstruct_fluxes_semiexplt_with_kornconst :117-137    eqlb_se_run (korn != NULL)
 *   reconstruct_fluxes_minimisation          :85-95         eqlb_ev_run
 *   local_solver_cholesky / _lu / _cg        :54-79         eqlb_local_project
 *   BoundaryData (the part the hot path reads) :235-256     eqlb_set_bcs
 *
 * All functions return 0 on success; on failure a negative code, with the text
 * of the `std::runtime_error` the reference would have thrown available from
 * eqlb_last_error().  Buffers are caller-owned; `memspace` says whether data
 * pointers are host (EQLB_HOST) or device (EQLB_DEVICE) memory.  There is no
 * CPU fallback: every compute entry point runs CUDA kernels and fails with
 * EQLB_ERR_CUDA if no device is usable.
 */
#ifndef EQLB_B200_H
#define EQLB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EQLB_HOST 0
#define EQLB_DEVICE 1
/* host pointers; sigma is taken as zero on entry and not read (what the reference does in
 * practice: `FluxEqlbSE.py:79-81` creates zero flux functions and equilibrates into them
 * once), which saves the host->device copy of sigma */
#define EQLB_HOST_ZEROED 2
/* G and f are host pointers, sigma is a device pointer (accumulated in place, result stays
 * on the device, e.g. for the halo sum of a distributed run); returns when the inputs have
 * been consumed, the kernels are ordered on the handle's stream */
#define EQLB_HOST_IN 3

#define EQLB_OK 0
#define EQLB_ERR_INPUT (-1)   /* the reference's std::runtime_error on bad input */
#define EQLB_ERR_CUDA (-2)    /* CUDA runtime / no device */
#define EQLB_ERR_STATE (-3)   /* call order (e.g. run before set_bcs) */

/* flags for eqlb_create */
#define EQLB_FLAG_STRESS 1u   /* first gdim fluxes are rows of a stress tensor (weak symmetry) */
#define EQLB_FLAG_ATOMIC 2u   /* accumulate with fp64 atomics instead of colour-ordered launches */
#define EQLB_FLAG_GENERIC 4u  /* always use the generic patch kernel (no degree-2 streaming kernel) */
#define EQLB_FLAG_K2_THREAD 8u /* reserved (former per-thread degree-2 kernel, removed); ignored */
/* host-pointer calls stream the mesh through the GPU in spatial stages: host->device copy of
 * stage s+1, patch kernels of stage s and device->host copy of finished DOF ranges overlap on
 * three streams (PCIe both directions busy); costs a few extra launches on device-pointer calls */
#define EQLB_FLAG_HOST_PIPELINE 16u
/* distributed runs (node_owned given): patches that touch a cell shared with another rank are
 * ordered first, so that eqlb_set_part can launch them separately and the halo exchange of
 * their DOFs overlaps the interior patches (ignored together with EQLB_FLAG_HOST_PIPELINE) */
#define EQLB_FLAG_INTERFACE_FIRST 32u

/* wire values, `base/Patch.hpp:20-33` */
enum eqlb_patch_type { EQLB_PATCH_INTERNAL = 0, EQLB_PATCH_ESSNT_DUAL = 1,
                       EQLB_PATCH_ESSNT_PRIMAL = 2, EQLB_PATCH_MIXED = 3 };
enum eqlb_facet_type { EQLB_FCT_INTERNAL = 0, EQLB_FCT_ESSNT_PRIMAL = 1,
                       EQLB_FCT_ESSNT_DUAL = 2 };

/* Mesh arrays: exactly the DOLFINx accessors the reference reads
 * (`se/Patch.cpp:23-28`, `se/reconstruction.hpp:83-93`,
 *  `se/solve_patch_semiexplt.hpp:223-224,454-463`, `ev/reconstruction.hpp:79-107`).
 * Always HOST pointers (read once at eqlb_create and copied to the device). */
typedef struct eqlb_mesh {
  int32_t nnode, ncell, nfct;
  const double*   x;              /* [nnode*3]  geometry().x()                       */
  const int32_t*  cell_node;      /* [ncell*3]  geometry().dofmap() == connectivity(2,0) */
  const int32_t*  cell_fct;       /* [ncell*3]  connectivity(2,1)                    */
  const int32_t*  fct_node;       /* [nfct*2]   connectivity(1,0)                    */
  const int32_t*  fct_cell_off;   /* [nfct+1]   connectivity(1,2) offsets            */
  const int32_t*  fct_cell;       /*            connectivity(1,2) array              */
  const int32_t*  node_cell_off;  /* [nnode+1]  connectivity(0,2)                    */
  const int32_t*  node_cell;
  const int32_t*  node_fct_off;   /* [nnode+1]  connectivity(0,1)                    */
  const int32_t*  node_fct;
  const uint8_t*  fct_perms;      /* [ncell*3]  topology().get_facet_permutations()  */
  const uint32_t* cell_perm_info; /* [ncell]    topology().get_cell_permutation_info() */
  const int32_t*  dg_dofmap;      /* [ncell*ndg] dofmap of the DG_p space of G and f; NULL = the DOLFINx layout
                                     cell*ndg + i (any other map: generic patch kernel, no projection calls) */
  const uint8_t*  node_owned;     /* [nnode] 1 = equilibrate the patch of this node (index_map(0) owned
                                     nodes, `se/reconstruction.hpp:90`); NULL = all nodes          */
} eqlb_mesh;

/* Reference-element tables (dolfinx_eqlb_b200/tables.py; Basix-derived in a
 * DOLFINx deployment).  HOST pointers. Shapes in tables.py::Tables. */
typedef struct eqlb_tables {
  int32_t k, p, nrt, ndg, ndg_fct, nq, nqf, ndiv, nadd, npk;
  /* quadrature-style tables (what se::KernelData holds) */
  const double *qpts, *qwts, *fpts_s, *fwts, *M, *rt_q, *rt_f, *dg_q, *dg_f, *hat_q, *hat_f, *trafo;
  const int32_t *fct_closure, *div_lm;
  /* reference-matrix tables (exact integrals on the reference cell) */
  const double *rt_mass, *fct_mom, *cell_mom_f, *cell_mom_g, *bc_mat, *rt_p1;
  const double *dg_mono, *hat_dg_rt, *mono_int; /* EV load vector / divergence data */
  /* change of basis hierarchic RT -> Basix "RT" (Legendre variant): [k][k] facet part, [(k*k-k)][nrt]
   * interior part (tables.py::basix_rt_legendre_maps); may be NULL (eqlb_ev_to_basix_rt then fails) */
  const double *rt_basix_fct, *rt_basix_int;
  /* primal space P_k (npk dofs per cell, Basix order) for eqlb_set_primal_space / eqlb_*_run_primal /
   * eqlb_estimate_*: gradients at the DG_p nodes [ndg][npk][2], L2 projection P_k -> DG_p [ndg][npk], basis /
   * gradient / second derivatives at the cell quadrature points [nq][npk](x2, x3), divergence of the RT basis at
   * the quadrature points [nq][nrt]; may be NULL (those entry points then fail) */
  const double *pk_grad_dg, *pk_to_dg, *pk_q, *pk_gq, *pk_hq, *rt_div_q;
} eqlb_tables;

typedef struct eqlb_handle eqlb_handle;

/* Build the device-resident problem: copies mesh + tables to the GPU.
 * Replaces the one-time part of `se::reconstruction` (`se/reconstruction.hpp:62-153`)
 * and `ev::reconstruction` (`ev/reconstruction.hpp:64-127`).
 * Errors like the reference: 1-cell patches (`se/Patch.cpp:353-359`), bad degrees
 * (`se/reconstruction.hpp:358-388`). */
int eqlb_create(const eqlb_mesh* mesh, const eqlb_tables* tables, int nrhs,
                uint32_t flags, eqlb_handle** out);
void eqlb_destroy(eqlb_handle* h);

/* CUDA stream all later launches go to (a cudaStream_t; NULL = default stream). */
int eqlb_set_stream(eqlb_handle* h, void* cuda_stream);

/* Boundary data consumed by the hot path = the state of `base::BoundaryData`
 * after its constructor (`base/BoundaryData.cpp:279-633`):
 *   facet_type  [nrhs*nfct] int8  (eqlb_facet_type)
 *   bflux[i]    [ncell*nrt] f64   boundary function of rhs i (DRT layout), may be
 *                                 NULL when rhs i has no flux BC
 *   local_fct_id[nfct]      int8  cell-local id of each flux-BC facet (validated against the mesh;
 *                                 may be NULL: the hot path derives it from the patch maps)
 *   node_on_stress_bnd [nnode] int8 (stress only, else NULL)
 * Runs the device patch builder (integer maps, patch types, colouring).
 * HOST pointers. */
int eqlb_set_bcs(eqlb_handle* h, const int8_t* facet_type, const double* const* bflux,
                 const int8_t* local_fct_id, const int8_t* node_on_stress_bnd);

/* Device-side construction of the same boundary data for tractions given as polynomials on the
 * boundary facets: replaces the host loop of the `base::BoundaryData` constructor
 * (`base/BoundaryData.cpp:279-633`, interpolation branch `:580-597`, `KernelDataBC::interpolate_flux
 * :149-161,229-250`) and the FluxBC objects of `wrappers.cpp:144-232` for this case.
 *   eqlb_fluxbc: on facet facets[i] the prescribed OUTWARD normal flux is
 *                g(s) = sum_j coeffs[i*ncoef + j] s^j, s = facet parameter of the adjacent cell
 *                ((1-s,s), (0,s), (s,0) for local facets 0, 1, 2)
 *   prime_facets[r][nprime[r]]  facets with essential BCs of the primal problem of rhs r
 *   bcs[r][nbc[r]]              flux BCs of rhs r
 * One thread per boundary facet computes facet types, cell-local facet ids, the boundary DOFs
 * (hierarchic facet moments, DRT layout) and - for stress problems - the node markers; then the
 * patch builder runs as in eqlb_set_bcs.  HOST pointers. */
typedef struct eqlb_fluxbc {
  int32_t nfct;
  const int32_t* facets;
  int32_t ncoef;
  const double* coeffs;
} eqlb_fluxbc;
int eqlb_set_bcs_poly(eqlb_handle* h, const int32_t* nprime, const int32_t* const* prime_facets,
                      const int32_t* nbc, const eqlb_fluxbc* const* bcs);
/* The boundary data resident on the device (after eqlb_set_bcs or eqlb_set_bcs_poly), for parity
 * tests: facet_type [nrhs*nfct], bflux[r] [ncell*nrt], local_fct_id [nfct], node_on_stress_bnd
 * [nnode] (stress handles).  Any pointer may be NULL.  HOST buffers. */
int eqlb_get_boundary_data(eqlb_handle* h, int8_t* facet_type, double* const* bflux, int8_t* local_fct_id,
                           int8_t* node_on_stress_bnd);

/* Semi-explicit equilibration: sigma[i] += equilibrated corrector of rhs i.
 *   G[i]     [ncell*ndg*2] projected flux (blocked bs=2), f[i] [ncell*ndg] projected RHS
 *   sigma[i] [ncell*nrt]   DRT_k vector, ACCUMULATED like the reference
 *                          (`se/solve_patch_semiexplt.hpp:1159`)
 *   korn     [ncell] or NULL: += 3 * squared Korn constants (`se/reconstruction.hpp:248-260`) */
int eqlb_se_run(eqlb_handle* h, const double* const* G, const double* const* f,
                double* const* sigma, double* korn, int memspace);

/* Constrained-minimisation (Ern-Vohralik) equilibration on the mixed RT_k x DG_(k-1)
 * patch spaces; sigma[i] [nfct*k + ncell*(k*k-k)] conforming hierarchic-RT vector,
 * ACCUMULATED (`ev/solve_patch.hpp:216-227`). */
int eqlb_ev_run(eqlb_handle* h, const double* const* G, const double* const* f,
                double* const* sigma, int memspace);

/* EV output in the space the reference's FluxEqlbEV returns (`python/dolfinx_eqlb/eqlb/FluxEqlbEV.py:95`:
 * Basix "RT", Legendre variant): converts conforming hierarchic-RT vectors (layout of eqlb_ev_run) into the
 * DOF vectors of the SAME functions w.r.t. the Basix functionals, same layout [facet dofs nfct*k (facets in
 * their global low->high vertex orientation)][cell dofs ncell*(k*k-k)]; a DOLFINx binding scatters them
 * through `V_flux.dofmap`.  in == out allowed only for different buffers (out of place).
 * The Basix functionals are restated from Basix' definition (third party, not in the reference tree):
 * see tables.py::basix_rt_legendre_maps for what is and is not pinned. */
int eqlb_ev_to_basix_rt(eqlb_handle* h, int nfun, const double* const* sigma_hier, double* const* sigma_basix, int memspace);

/* ---- fused input stage and error estimator ----
 * eqlb_set_primal_space: the continuous P_k space of the primal solution (k = flux degree): cell dofmap
 *   [ncell*npk] (Basix DOF order: vertices, edge interiors, cell interior; HOST pointer, copied) and its size.
 * eqlb_project_primal: what `lsolver/projection.py:17-77` computes before the equilibration, on the device:
 *   G[r] = Pi_{DG_p}(-grad u_h[r]) (exact: grad u_h is piecewise P_{k-1}), F[r] = Pi_{DG_p} f_h[r], both from P_k
 *   coefficient vectors; writes the [ncell*ndg*2] / [ncell*ndg] arrays eqlb_se_run / eqlb_ev_run take.
 *   Any of uh / fh (and the matching output) may be NULL.
 * eqlb_ev_run_primal / eqlb_se_run_primal: projection + equilibration in one call; host callers move the two P_k
 *   vectors (8 B per primal dof) instead of G and F (24 ndg B per cell): the host->device volume of an EV degree-2
 *   call at 1024^2 drops from 302 MB to 134 MB.
 * eqlb_estimate_poisson: cell-wise error indicators of `python/demo/poisson/demo_error_estimation.py:52-122`
 *   eta_sig2[c] = ||sigma_eqlb||^2_T (semi-explicit, discontinuous flux) or ||grad u_h + sigma_eqlb||^2_T (conforming)
 *   eta_osc2[c] = (h_T/pi)^2 ||f_h - div sigma||^2_T, sigma = sigma_eqlb - grad u_h (SE) or sigma_eqlb (EV)
 *   sigma: DRT vector (is_ev = 0) or conforming hierarchic-RT vector (is_ev = 1); only 2 * ncell doubles per
 *   function go back to the host instead of the flux vector.
 * eqlb_estimate_elasticity: `python/demo/elasticity/demo_error_estimation.py:49-135`, displacement formulation:
 *   eta[0][c] = int dsig : a(dsig), a(s) = (s - pi_1/(2+2 pi_1) tr(s) I)/2;  eta[1][c] = int (c_K (dsig_01 - dsig_10)/2)^2;
 *   eta[2][c] = (c_K h_T/pi)^2 ||f_h + div(sigma_h + dsig)||^2_T; dsig rows = DRT vectors, sigma_h rows = DG_p^2 vectors
 *   (the projected stress handed to the equilibration as -G), f_h rows = P_k vectors, korn = cell-wise c_K.
 * memspace: EQLB_HOST or EQLB_DEVICE for all data pointers of a call. */
int eqlb_set_primal_space(eqlb_handle* h, const int32_t* pk_dofmap, int64_t ndofs);
int eqlb_project_primal(eqlb_handle* h, int nfun, const double* const* uh, const double* const* fh, double* const* G,
                        double* const* F, int memspace);
int eqlb_ev_run_primal(eqlb_handle* h, const double* const* uh, const double* const* fh, double* const* sigma, int memspace);
int eqlb_se_run_primal(eqlb_handle* h, const double* const* uh, const double* const* fh, double* const* sigma, double* korn,
                       int memspace);
int eqlb_estimate_poisson(eqlb_handle* h, int nfun, const double* const* sigma, const double* const* uh,
                          const double* const* fh, double* const* eta_sig2, double* const* eta_osc2, int is_ev, int memspace);
int eqlb_estimate_elasticity(eqlb_handle* h, const double* const* dsig, const double* const* sigma_h, const double* const* fh,
                             const double* korn, double pi_1, double* const* eta, int memspace);

/* Cell-wise L2 projection into DG_p (the fixed-form fast path of
 * `base::local_solver_cholesky`, `base/local_solver.hpp:38-187`, as used by
 * `lsolver/projection.py:17-77`): out[cell*ndg+i] = dofs of the projection of the
 * function given by its values at the nq cell quadrature points.
 *   qvals [nfun][ncell*nq], out [nfun][ncell*ndg] (assigned, not accumulated). */
int eqlb_local_project(eqlb_handle* h, int nfun, const double* const* qvals,
                       double* const* out, int memspace);

/* Integer patch maps in the reference's layout, for bit-exact comparison
 * (SURVEY App. A).  All HOST output buffers, padded with -1 beyond a patch's
 * size; ncmax from eqlb_patch_dims.  Any pointer may be NULL.
 *   ncells[npatch], cells/fcts/inodes_local [npatch*(ncmax+2)],
 *   fcts_local [npatch*2*(ncmax+1)], type [npatch*nrhs],
 *   reversed [npatch*ncmax*2], reversion [npatch*nrhs], colour[npatch] */
int eqlb_patch_dims(eqlb_handle* h, int32_t* npatch, int32_t* ncmax, int32_t* ncolours);
int eqlb_get_patch_maps(eqlb_handle* h, int32_t* ncells, int32_t* cells, int32_t* fcts,
                        int8_t* inodes_local, int8_t* fcts_local, int8_t* type,
                        uint8_t* reversed, uint8_t* reversion, int32_t* colour);
/* SE 4-plane DOF map + facet DOFs of the projected flux + boundary markers:
 *   dofmap [npatch*4*(ncmax+2)*ndpc], ndpc = 2k+nadd(+3 if stress)+ndiv
 *   projflux_fct [npatch*(ncmax+1)*2*ndg_fct], bmarkers [npatch*nrhs*hzmax] */
int eqlb_get_se_dofmaps(eqlb_handle* h, int32_t* dofmap, int32_t* projflux_fct,
                        int8_t* bmarkers, int32_t* ndpc, int32_t* hzmax);

/* EV patch ordering and sub-DOF maps of the mixed RT_k x DG_(k-1) patch problem
 * (`ev/Patch.cpp:83-309` ordering, `:482-676` create_subdofmap), EV storage convention;
 * nz = nrt + ndg - k non-zero element DOFs per patch cell.  HOST buffers, -1 padded:
 *   ncells [npatch], cells/inodes_local [npatch*ncmax], fcts [npatch*(ncmax+1)],
 *   dofs_elmt/dofs_patch/dofs_global [npatch*ncmax*nz]   (_dofsnz_elmt/_patch/_global)
 *   list_patch/list_global [npatch*(ncmax*(k*k-k) + (ncmax+1)*k)] (_list_dofsnz_*_fluxhdiv)
 * global numbering: facet*k+j | nfct*k + cell*(k*k-k) + i | nflux + cell*ndg + q */
int eqlb_get_ev_dofmaps(eqlb_handle* h, int32_t* ncells, int32_t* cells, int32_t* fcts, int8_t* inodes_local,
                        int32_t* dofs_elmt, int32_t* dofs_patch, int32_t* dofs_global,
                        int32_t* list_patch, int32_t* list_global);

/* Which patches the next eqlb_se_run / eqlb_ev_run calls launch: EQLB_PART_ALL (default),
 * EQLB_PART_INTERFACE (owned patches touching a cell that holds a vertex of another rank) or
 * EQLB_PART_INTERIOR (the rest).  Needs EQLB_FLAG_INTERFACE_FIRST.  Interior patches never
 * write a DOF the halo exchange reads or updates, so
 *   run(INTERFACE); record event; run(INTERIOR) || [wait event; halo sum on a second stream]
 * gives the result of run(ALL); halo sum with the exchange hidden behind the interior kernels. */
#define EQLB_PART_ALL 0
#define EQLB_PART_INTERFACE 1
#define EQLB_PART_INTERIOR 2
int eqlb_set_part(eqlb_handle* h, int part);

/* ---- multi-GPU halo sum over NVLink peer memory (one process per GPU, no NCCL on the data path) ----
 * The reference sums shared DOFs through PETSc ghost updates after `equilibrate_fluxes`
 * (`x.scatter_reverse`); here every rank packs its shared values, publishes them with a flag
 * in the neighbour's memory and adds the neighbours' values, all in ONE kernel launch
 * (csrc/halo_p2p.cu).  Neighbours are given in ascending rank order; `idx[n]` are the local
 * DOF indices shared with neighbour n, ordered identically (by global id) on both sides.
 *   eqlb_halo_create  -> CUDA IPC handle of the own buffer + byte offsets of the send areas;
 *   the caller exchanges (handle, offsets) between the ranks (e.g. all_gather) and calls
 *   eqlb_halo_connect(h, n, neighbour's handle, neighbour's offset of the area for this rank,
 *                     this rank's position in the neighbour's neighbour list);
 *   eqlb_halo_apply(h, x, nrhs, stream): x[r][idx] += sum over neighbours, deterministic. */
typedef struct eqlb_halo eqlb_halo;
int eqlb_halo_create(int nneigh, const int64_t* counts, const int64_t* const* idx, int nrhs_max,
                     eqlb_halo** out, unsigned char* ipc_handle_out, int64_t* send_off_out);
int eqlb_halo_connect(eqlb_halo* h, int n, const unsigned char* peer_ipc_handle,
                      int64_t peer_send_off, int peer_slot);
int eqlb_halo_apply(eqlb_halo* h, double* const* x, int nrhs, void* cuda_stream);
/* synchronises the stream and reports a timed-out exchange (all device-side waits are bounded: a neighbour that
 * never arrives costs seconds, not a hung GPU) */
int eqlb_halo_status(eqlb_halo* h, void* cuda_stream);

/* Launch order of the patches (inspection / tests): order [number of owned nodes] = node ids sorted by
 * (segment = chunk * ncolours + colour, lane class, node id), grouped boundary patches first; nchunk = spatial
 * chunks (stages of the host pipeline); seg_off [nchunk*ncolours + 1] = offsets of the segments in `order`.
 * The reference has no counterpart (it loops over the nodes in index order, `se/reconstruction.hpp:150-160`). */
int eqlb_get_launch_order(eqlb_handle* h, int32_t* order, int32_t* nchunk, int32_t* seg_off);

/* Device copy of the flux vector r written by the last host-buffer call (EQLB_HOST / EQLB_HOST_ZEROED) of this
 * handle: distributed callers run the halo sum on it (eqlb_halo_apply) and fetch only the shared DOFs again,
 * instead of holding back the whole copy-out until the exchange is done (dolfinx_eqlb_b200/dist.py
 * `equilibrate_host`).  Valid until the next call on the handle.  The reference has no counterpart (its flux
 * vectors are host PETSc vectors whose ghost update is PETSc's, `se/reconstruction.hpp:150-160`). */
int eqlb_get_staged_flux(eqlb_handle* h, int r, int is_ev, double** device_ptr, int64_t* n);

/* every rank must have finished its exchanges (collective barrier + device synchronisation) before any rank
 * destroys its handle: the neighbours read this rank's buffer over NVLink */
void eqlb_halo_destroy(eqlb_halo* h);

/* Cell-wise squared L2 norm of DRT_k flux functions, eta2[i][cell] = ||sigma_i||^2_{L2(cell)}:
 * the flux part `dot(err_sig, err_sig) * v * dx` of the reference's error indicators for the
 * semi-explicit (discontinuous) flux (`python/demo/poisson/demo_error_estimation.py:96-101`),
 * i.e. the consumer right after `equilibrate_fluxes`; with EQLB_DEVICE the flux never has to
 * leave the GPU.  sigma[i] [ncell*nrt], eta2[i] [ncell] (assigned); EQLB_HOST or EQLB_DEVICE. */
int eqlb_flux_l2norm(eqlb_handle* h, int nfun, const double* const* sigma, double* const* eta2, int memspace);

/* Page-lock / unlock a caller-owned host buffer (cudaHostRegister): host-pointer calls on
 * pageable memory are staged by the driver at a fraction of the PCIe rate.  Register the
 * flux / RHS vectors once per function (e.g. the PETSc arrays), not per call. */
int eqlb_pin_host(void* ptr, size_t bytes);
int eqlb_unpin_host(void* ptr);

/* number of kernel launches issued by this handle so far (bench "gpu_launches") */
int64_t eqlb_launch_count(eqlb_handle* h);

/* Measurement aid for the FP64 roofline (not on the hot path): register-resident DFMA chains
 * on every SM of the current device for `iters` iterations, best of `reps` launches timed with
 * CUDA events; returns TFLOP/s (2 flop per FMA) in *tflops. */
int eqlb_measure_fp64_peak(int iters, int reps, double* tflops);

const char* eqlb_last_error(void);
const char* eqlb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* EQLB_B200_H */
