/* apple_b200 -- C ABI of the B200-native FEM-elasticity hot path.
 *
 * This is the drop-in boundary for liblaf/apple's Warp backend.  Every
 * entry point names the reference interface it replaces; paths are relative to
 * /root/reference/src/liblaf/apple.  Plain pointers and sizes only -- no torch, Warp or JAX types.
 *
 * Conventions
 *  - Every function returns APL_OK (0) or a negative APL_ERR_* code; the message is available from
 *    apl_last_error() (thread-local).  No exceptions cross the ABI, no call synchronises the host
 *    with the device unless its comment says so.
 *  - `dtype` selects fp32 / fp64 for every `void*` floating-point array of that call (the reference
 *    is dtype-generic through the JAX x64 flag, warp/fem/_base.py:100-104).
 *  - Nodal fields (u, p, grad, diag, prod) are device arrays of n_points rows with a leading
 *    dimension `ld` of 3 (the reference's vec3 arrays) or 4 (16-byte padded rows; the 4th column is
 *    ignored on input and receives +0 on output).  ld == 4 arrays must be 16-byte aligned, ld == 3
 *    OUTPUT arrays 8-byte aligned (vector reductions); misaligned pointers are rejected.
 *  - Operators ACCUMULATE into caller-zeroed outputs, exactly like WarpPotential
 *    (warp/model/_potential.py:19-32; zeroing by the caller, warp/model/_model.py:14,19,24,29,34).
 *  - Calls are stream-ordered on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *    stream).  A handle may be used from one host thread and one stream at a time.
 *  - Device buffers passed in are owned by the caller; handles own only their packed mesh tables.
 */
#ifndef APPLE_B200_H
#define APPLE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APL_OK 0
#define APL_ERR_INVALID (-1)  /* bad argument */
#define APL_ERR_CUDA (-2)     /* a CUDA runtime call failed (no device, OOM, launch error) */
#define APL_ERR_MESH (-3)     /* mesh violates an assumption (index range, dhdX rows not summing to 0) */
#define APL_ERR_STATE (-4)    /* handle not in a state that allows the call */

#define APL_F32 0
#define APL_F64 1

/* energies: warp/fem/_stable_neo_hookean.py, warp/fem/_arap.py, warp/fem/_stable_neo_hookean_muscle.py */
#define APL_KIND_SNH 0
#define APL_KIND_ARAP 1
#define APL_KIND_SNH_MUSCLE 2
#define APL_KIND_SNH_ARAP 3 /* an SNH and an ARAP potential on the SAME cells, summed in one pass (see below) */

/* operator bit mask; any OR of these is evaluated in ONE pass over the elements */
#define APL_OP_FUN 1        /* WarpPotentialFem.fun        warp/fem/_base.py:151-158, kernel :243-263 */
#define APL_OP_GRAD 2       /* WarpPotentialFem.grad       warp/fem/_base.py:160-167, kernel :265-291 */
#define APL_OP_HESS_DIAG 4  /* WarpPotentialFem.hess_diag  warp/fem/_base.py:169-176, kernel :293-324 */
#define APL_OP_HESS_PROD 8  /* WarpPotentialFem.hess_prod  warp/fem/_base.py:178-185, kernel :326-351 */
#define APL_OP_HESS_QUAD 16 /* WarpPotentialFem.hess_quad  warp/fem/_base.py:187-194, kernel :353-383 */
/* Opt-in supersets of the reference (BASELINE.json north star: "3x3-block Hessian diagonal", "analytic PSD
 * eigen-projection"); default behaviour and parity are unchanged when the bits are clear.  TILE assembly only.
 *  HESS_OFFD  off-diagonal entries (xy, xz, yz) of the 3x3 VERTEX BLOCKS of the assembled Hessian, accumulated into the
 *             `prod` argument of apl_fem_eval (so never together with HESS_PROD / HESS_QUAD); with HESS_DIAG -- whose
 *             entries are the block diagonals, the clamp of warp/fem/_base.py:317-320 kept -- it is the block-Jacobi
 *             preconditioner.  Combines with FUN, GRAD, HESS_DIAG.
 *  PSD        modifier: the Hessian terms (diag, offd, prod, quad) of the Stable Neo-Hookean kinds use d2Psi/dF2 with its
 *             negative eigenvalues set to zero, from the ANALYTIC eigen-system (3 twist, 3 flip, 3 scaling modes in the
 *             SVD frame of F); a superset of the clamps at warp/fem/_base.py:317-320,379-380, which stay in place.  ARAP:
 *             no-op, its clamped twist rates (warp/fem/func/_misc.py:31-43) already are the projection. */
#define APL_OP_HESS_OFFD 32
#define APL_OP_PSD 64

/* assembly strategy */
#define APL_SCATTER_TILE 0   /* TMA/mbarrier-pipelined tiles: shared-memory gather, in-tile slot reduction, one
                                vector RED per tile vertex and field (the product path) */
#define APL_SCATTER_ATOMIC 1 /* one thread per tet, direct gathers and 12 REDs per field (reference-like) */
#define APL_SCATTER_TILE_SIMPLE 2 /* same tiles without the producer warp / bulk-copy pipeline */

typedef struct apl_fem apl_fem_t;   /* one FEM potential: replaces WarpPotentialFem (warp/fem/_base.py:39) */
typedef struct apl_pncg apl_pncg_t; /* fused PNCG workspace (liblaf.peach.optim.PNCG, external to the reference) */

int apl_version(void);
const char* apl_last_error(void);
/* number of CUDA devices visible, or a negative error code (never throws; usable as a probe) */
int apl_device_count(void);

/* ---- setup: replaces WarpPotentialFem.from_region (warp/fem/_base.py:93-111) ---------------------
 * HOST arrays in the reference's own layout:
 *   cells   int32  (n_cells,4)   region.cells_global
 *   dhdX    dtype  (n_cells,4,3) region.dhdX[:,0]      (rows must sum to zero: linear tetrahedra)
 *   dV      dtype  (n_cells,)    Fraction * region.dV[:,0]
 *   mu, lambda_ dtype (n_cells,) materials (lambda_ ignored for ARAP, may be NULL)
 *   activation dtype (n_cells,6) SNH-muscle only (warp/fem/func/_misc.py:46-53 ordering), else NULL
 *   points  double (n_points,3)  optional rest positions, used only to order the tets along a Morton
 *                                curve; NULL keeps the given cell order
 * The tets are packed into tiles of <= 256 tets touching <= 256 distinct vertices, static data is
 * stored as planes of 16-byte vectors, and everything is uploaded to `device`.
 * device = -1 builds the host tables only (no CUDA call; for inspection and CPU tests). */
int apl_fem_create(int kind, int dtype, int64_t n_cells, int64_t n_points, const int32_t* cells,
                   const void* dhdX, const void* dV, const void* mu, const void* lambda_,
                   const void* activation, const double* points, int device, apl_fem_t** out);
/* Two potentials that share their cells -- what WarpModel (warp/model/_model.py:13-36) would evaluate as
 * two kernel launches per operator, each re-reading the mesh and re-gathering the vertices -- fused
 * into ONE handle whose passes return the SUM of a Stable Neo-Hookean and an ARAP potential (each with
 * its own dV = Fraction * dV and materials; clamps applied per potential exactly as in the reference). */
int apl_fem_create_snh_arap(int dtype, int64_t n_cells, int64_t n_points, const int32_t* cells,
                            const void* dhdX, const void* dV_snh, const void* mu_snh, const void* lambda_snh,
                            const void* dV_arap, const void* mu_arap, const double* points, int device,
                            apl_fem_t** out);
void apl_fem_destroy(apl_fem_t* fem);

/* ---- setup from DEVICE arrays: replaces Region.compute_grad (jax/fem/region/_region.py:84-108: dXdr, drdX,
 * dV = det / 6, dhdX = dhdr . drdX for the linear tetrahedron, jax/fem/element/_tetra.py:36-45 with the one-point
 * rule jax/fem/quadrature/_tetra.py:12-15) TOGETHER WITH WarpPotentialFem.from_region (warp/fem/_base.py:93-111:
 * dV *= Fraction, materials), for meshes that already live in HBM (the reference does this once per mesh on the
 * host in JAX; at 64 M tets that is minutes, here seconds).  DEVICE arrays:
 *   cells      int32  (n_cells,4)
 *   points     double (n_points,3)  rest positions
 *   fraction   dtype  (n_cells,)    cell_data["Fraction"] (warp/fem/utils/_material.py:19-23); NULL = 1
 *   mu, lambda_, activation         as in apl_fem_create, on the device
 *   fraction2, mu2                  the ARAP half of APL_KIND_SNH_ARAP (fraction2 NULL = 1), else NULL
 * morton != 0 orders the tets along the Morton curve of their centroids (device radix sort, same keys and the
 * same stable order as apl_fem_create with `points`), 0 keeps the given order.  The rest shape is evaluated in
 * fp64 and rounded to `dtype`.  Synchronises the device (setup-time call).  A tet with zero rest volume is an
 * error (APL_ERR_MESH), negative volumes are accepted as in the reference (region/_region.py:98-99 only warns). */
int apl_fem_create_from_mesh(int kind, int dtype, int64_t n_cells, int64_t n_points, const int32_t* cells,
                             const double* points, const void* fraction, const void* mu, const void* lambda_,
                             const void* activation, const void* fraction2, const void* mu2, int morton, int device,
                             apl_fem_t** out);

/* info[0..9] = n_cells, n_points, n_tiles, length of tile_verts, static device bytes,
 *              kind, dtype, device, length of tile_voff, number of packed tet positions (= n_cells) */
int apl_fem_info(const apl_fem_t* fem, int64_t info[10]);

/* Copies of the host tables (sizes from apl_fem_info); any pointer may be NULL.
 *   tiles      int32 (n_tiles,6): tet_start, n_tets, vert_start, n_verts, voff_start, 0
 *   order      int64 (info[9],)  : packed tet position -> caller's cell index
 *   conn       uint8 (n_cells,4) : tile-local vertex id per corner (packed order)
 *   slots      uint16(n_cells,4) : in-tile reduction slot per corner (packed order)
 *   tile_verts int32 (info[3])   : global vertex id per tile-local id, from vert_start
 *   tile_voff  uint16(info[8])   : per tile, n_verts+1 entries starting at voff_start: bits 0..11 first slot of
 *                                  the vertex's range (reduce order), bits 12..15 unused pad slots after it
 *   tile_vperm uint8 (info[3])   : per tile, the local ids in reduce order (groups of 16 by decreasing valence) */
int apl_fem_host_tables(const apl_fem_t* fem, int32_t* tiles, int64_t* order, uint8_t* conn,
                        uint16_t* slots, int32_t* tile_verts, uint16_t* tile_voff, uint8_t* tile_vperm);
/* Host-only handles (device = -1): copy of the packed static planes, [n_planes][plane_stride] 16-byte
 * vectors in packed cell order (record = Dm^-1 (9), dV, mu, lambda, activation (6) / second potential);
 * `planes` may be NULL to query the sizes. */
int apl_fem_host_planes(const apl_fem_t* fem, void* planes, int64_t* n_planes, int64_t* plane_stride);

/* Replace per-cell materials in place (HOST arrays in the caller's cell order; NULL = keep).
 * Replaces re-creating the Materials struct (warp/fem/utils/_material.py:15-31). */
int apl_fem_set_materials(apl_fem_t* fem, const void* dV, const void* mu, const void* lambda_,
                          const void* activation);

/* ---- the operators: replace the wp.launch calls at warp/fem/_base.py:151-194 ---------------------
 * ops   OR of APL_OP_*; everything requested is computed in one pass over the elements.
 * u, p  device nodal fields with leading dimension ld_in (p may be NULL unless HESS_PROD/HESS_QUAD).
 * fun, quad          device scalars (dtype[1]);      fun += sum Psi dV,  quad += sum max(p.Hp dV, 0)
 * grad, diag, prod   device nodal fields, ld_out;    accumulated (diag clamped >= 0 per entry and cell)
 * Outputs whose op bit is not set are ignored (may be NULL). */
int apl_fem_eval(apl_fem_t* fem, int ops, const void* u, const void* p, int ld_in, void* fun,
                 void* quad, void* grad, void* diag, void* prod, int ld_out, int scatter,
                 void* stream);

/* ---- mixed derivative product: what the reference's inverse problems call model.mixed_derivative_prod(state, p)
 * after the adjoint solve (exp/2026/01/28/smas/src/31-inverse-activation-stable-neo-hookean.py:472-487; the method
 * is absent from the reference's current src/, so this is the analytic derivative of the energies of this library).
 * Per cell c:  d/dq_c [ grad_u E(u) . p ]  for the cell's own material parameters q_c.
 * Outputs: DEVICE arrays in the CALLER's cell order, overwritten; NULL = not wanted.
 *   d_mu (n_cells), d_lambda (n_cells; Stable Neo-Hookean kinds), d_activation (n_cells,6; muscle kind).
 * Not available for the fused SNH+ARAP handle (evaluate its two potentials separately). */
int apl_fem_mixed_derivative_prod(apl_fem_t* fem, const void* u, const void* p, int ld_in, void* d_mu,
                                  void* d_lambda, void* d_activation, void* stream);

/* ---- multi-GPU overlap (no reference counterpart: the reference is single-GPU) -----------------------
 * apl_fem_mark_boundary: tiles that touch a vertex with vertex_flags[v] != 0 (HOST array, n_points
 * bytes; a sharded caller flags the vertices it shares with other ranks) become "boundary" tiles and
 * their headers are moved to the front of the tile list (no element data moves).  *n_boundary receives
 * their number.  NULL flags restore "no boundary tiles".
 * apl_fem_eval_part: apl_fem_eval over a part of the tiles -- APL_PART_ALL, APL_PART_BOUNDARY or
 * APL_PART_INTERIOR.  Interior tiles touch no flagged vertex, so the halo exchange of the boundary
 * results can run while the interior pass is still accumulating.  Both parts ADD to fun / quad. */
#define APL_PART_ALL 0
#define APL_PART_BOUNDARY 1
#define APL_PART_INTERIOR 2
int apl_fem_mark_boundary(apl_fem_t* fem, const uint8_t* vertex_flags, int64_t* n_boundary);
int apl_fem_eval_part(apl_fem_t* fem, int part, int ops, const void* u, const void* p, int ld_in, void* fun,
                      void* quad, void* grad, void* diag, void* prod, int ld_out, int scatter, void* stream);

/* ---- ExternalForce: replaces warp/potential/_ext_force.py:17-39 ------------------------------------
 * force dtype (k,3) and indices int32 (k,) are DEVICE arrays.  ops may contain FUN and/or GRAD:
 *   fun[0] -= sum_k force_k . u[indices_k]        grad[indices_k] -= force_k                     */
int apl_ext_force_eval(int dtype, int ops, int64_t k, const void* force, const int32_t* indices,
                       const void* u, int ld_in, void* fun, void* grad, int ld_out, void* stream);

/* ---- nodal-field helpers --------------------------------------------------------------------------
 * dst(n,ld_dst) = src(n,ld_src) for the first 3 columns; padding columns of dst are zeroed. */
int apl_field_copy(int dtype, int64_t n, const void* src, int ld_src, void* dst, int ld_dst,
                   void* stream);

/* ---- halo sums for sharded meshes (new functionality: the reference is single-GPU) ------------------
 * Between pack and unpack the caller moves `send` to the other ranks (one all-to-all of the shared rows,
 * e.g. ncclSend/ncclRecv groups or torch.distributed.all_to_all_single) into `recv`.
 *   pack:   send[i, f, 0:3] = field_f[index[i], 0:3]        (i < n rows shared with other ranks)
 *   unpack: for each of the n_shared shared vertices, field_f[shared[i]] = sum over its CSR entries
 *           src[row_ptr[i] .. row_ptr[i+1]) taken in order: -1 = this rank's own partial, j >= 0 = row j
 *           of recv.  Listing the entries in ascending rank order makes all replicas bit-identical.
 * All arrays are DEVICE arrays; up to 3 fields (f1, f2 may be NULL) of leading dimension ld. */
int apl_halo_pack(int dtype, int64_t n, const int64_t* index, int nf, const void* f0, const void* f1,
                  const void* f2, int ld, void* send, void* stream);
int apl_halo_unpack(int dtype, int64_t n_shared, const int64_t* shared, const int32_t* row_ptr,
                    const int64_t* src, int nf, void* f0, void* f1, void* f2, int ld, const void* recv,
                    void* stream);

/* ---- halo sums through PEER MEMORY over NVLink / NVSwitch (new functionality: the reference is single-GPU) --------
 * Replaces pack -> NCCL all-to-all -> unpack -> NCCL all-reduce (four host-launched operations) by two small kernels
 * with a device-side hand-shake; results are bit-identical to apl_halo_pack / apl_halo_unpack with the same plan.
 * One process per GPU of one node.  Setup: every rank creates its handle, publishes apl_xchg_ipc_handle (64 bytes)
 * to the other ranks through any host channel, and maps theirs with apl_xchg_connect (handles: world x 64 bytes in
 * rank order, the own entry is ignored).  The plan arrays are DEVICE arrays owned by the caller (they must outlive
 * the handle): row i of the send list is this rank's partial on local vertex send_index[i], stored into row
 * send_row[i] of rank send_peer[i]'s receive buffer; (shared, row_ptr, src) is the CSR of apl_halo_unpack.
 *   apl_xchg_push: stores the shared rows of up to 3 fields (leading dimension ld) and up to 16 partial scalars
 *                  (dtype, e.g. this rank's energy) into the peers' buffers and publishes the epoch.
 *   apl_xchg_pull: waits on the device until every rank's push of this epoch has landed, then sums the partials of
 *                  every shared vertex in ascending rank order into f0..f2 and the scalars, in rank order, into scal
 *                  (overwritten with the global sums: identical bits on every rank).
 * Every rank must call push and pull the same number of times with the same nf / n_scal (collective semantics);
 * push and pull of one exchange may be enqueued on different streams as long as pull is ordered after push. */
/* max_recv_rows: capacity of one receive buffer in shared rows; it MUST be the same number on every rank (use the
 * largest number of rows any rank receives): the two receive buffers of a region are addressed with it from the peers;
 * apl_xchg_connect compares the capacities published in the regions and fails with APL_ERR_INVALID on a mismatch. */
typedef struct apl_xchg apl_xchg_t;
int apl_xchg_create(int world, int rank, int device, int64_t max_recv_rows, apl_xchg_t** out);
void apl_xchg_destroy(apl_xchg_t* x);
int apl_xchg_ipc_handle(apl_xchg_t* x, void* handle64);
int apl_xchg_connect(apl_xchg_t* x, const void* handles);
int apl_xchg_set_plan(apl_xchg_t* x, int64_t n_send, const int64_t* send_index, const int32_t* send_peer,
                      const int64_t* send_row, int64_t n_shared, const int64_t* shared, const int32_t* row_ptr,
                      const int64_t* src);
int apl_xchg_push(apl_xchg_t* x, int dtype, int nf, const void* f0, const void* f1, const void* f2, int ld,
                  const void* scal, int n_scal, void* stream);
int apl_xchg_pull(apl_xchg_t* x, int dtype, int nf, void* f0, void* f1, void* f2, int ld, void* scal, int n_scal,
                  void* stream);

/* ---- fused PNCG workspace -------------------------------------------------------------------------
 * The optimizer the reference uses (liblaf.peach.optim.PNCG) is external to /root/reference; what is
 * restated here are the recurrences of the reference's own PNCG-like benchmark,
 * benches/bench_pncg_branching_backends.py:254-329,407-410,606-610,663-679, with the problem glue of
 * forward/_problem.py:24-59 folded in (fixed DOFs are masked, not gathered/scattered).
 *
 * All vectors are DEVICE arrays of n_points rows x 4 columns (16-byte rows, 4th column unused),
 * allocated and owned by the caller: x, two direction buffers p0/p1, two gradient buffers g0/g1, two
 * Hessian-diagonal buffers d0/d1.  `mask` is uint8 (n_points,4): bit 0 = free DOF, bit 1 = counted in
 * reductions (clear on ghost copies when a mesh is sharded).  `scal` is a device array of
 * APL_PNCG_NSCAL doubles holding every scalar of the iteration; the host never has to read it
 * between iterations.  Buffer roles (current / trial, current / previous) flip every iteration;
 * apl_pncg_current() tells which index is current (always scal[APL_S_K] % 2). */
#define APL_PNCG_NSCAL 96
#define APL_S_F 0            /* energy at the current iterate */
#define APL_S_F_PREV 1       /* energy before the last accepted step */
#define APL_S_GP 2           /* g . p */
#define APL_S_PHP 3          /* hess_quad(x, p) */
#define APL_S_ALPHA 4        /* step length of the last iteration */
#define APL_S_BETA 5
#define APL_S_GNORM2 6       /* |g|^2 over free DOFs at the start of the last iteration */
#define APL_S_GNORM2_FIRST 7
#define APL_S_ACCEPTED 8     /* 1 if the last line search accepted a trial point */
#define APL_S_LS_STEPS 9     /* halvings used by the last line search */
#define APL_S_K 10           /* iterations performed */
#define APL_S_N_ACCEPTED 11
#define APL_S_DIAG_MEAN 12   /* mean of the positive |hess_diag| entries */
#define APL_S_GPG 13         /* g . P g */
#define APL_S_DONE 15        /* 0 running, 1 gradient criterion met, 2 max_steps, 3 stagnation, 4 non-finite */
#define APL_S_FAILS 16       /* consecutive failed line searches */
#define APL_S_F_NEW 17
#define APL_S_J 18           /* trials evaluated so far in the current line search (device-side counter) */
#define APL_S_SUMS 20        /* 11 reduction results of APL_PHASE_REDUCE */
#define APL_S_ALPHA_J 32     /* per-trial step lengths      [16] */
#define APL_S_ACC_J 48       /* per-trial line-search state [16]: 1 accepted, 0 live, -1 gave up */
#define APL_S_FT_J 64        /* per-trial energies          [16] */

#define APL_PHASE_INIT 0      /* f, g, diag at x; resets scal and the buffer roles */
#define APL_PHASE_REDUCE 1    /* 11 masked sums over (g, g_prev, diag, p_prev) -> scal[APL_S_SUMS..] */
#define APL_PHASE_FINALIZE 2  /* termination tests, Dai-Kou beta, descent guard */
#define APL_PHASE_DIRECTION 3 /* p = -P g + beta p_prev; scal[GP]; zero the trial buffers */
#define APL_PHASE_PASS_B 4    /* scal[PHP] += hess_quad(x, p) over the registered potentials */
#define APL_PHASE_ALPHA 5     /* alpha_0 */
#define APL_PHASE_TRIAL 6     /* (j) f', g', diag' at x + alpha_j p (skipped once accepted) */
#define APL_PHASE_LS 7        /* (j) Armijo test of trial j; halve and re-zero on failure */
#define APL_PHASE_COMMIT 8    /* x += alpha p or restore g', diag'; k += 1 */

int apl_pncg_create(int dtype, int64_t n_points, int device, void* x, void* p0, void* p1, void* g0,
                    void* g1, void* d0, void* d1, const uint8_t* mask, double* scal, apl_pncg_t** out);
void apl_pncg_destroy(apl_pncg_t* ws);
/* potentials summed by the workspace's passes (WarpModel, warp/model/_model.py:9-36) */
int apl_pncg_add_fem(apl_pncg_t* ws, apl_fem_t* fem);
int apl_pncg_add_ext_force(apl_pncg_t* ws, int64_t k, const void* force, const int32_t* indices);
/* max_steps: iteration budget (forward/_forward.py:26-27); rtol_g / atol_g: stop when
 * |g| <= rtol_g |g_0| or |g| <= atol_g; max_fails: consecutive failed line searches tolerated;
 * overstep, max_step, c1, max_halvings: line search (bench :606-610, :413-456; _problem.py:29-34);
 * scatter: APL_SCATTER_*; use_graph: 0 = plain launches, 1 = replay each iteration as a CUDA graph with
 * all max_halvings + 1 flag-guarded trials, 2 = CUDA graph whose backtracking is a conditional WHILE
 * node (device-side loop: trials beyond the first cost nothing unless the Armijo test fails). */
int apl_pncg_set_params(apl_pncg_t* ws, double max_steps, double rtol_g, double atol_g, double max_fails,
                        double overstep, double max_step, double c1, int max_halvings, int scatter,
                        int use_graph);
/* OPT-IN superset of the reference's preconditioner (scalar Jacobi on the clamped diagonal, restated from
 * benches/bench_pncg_branching_backends.py:407-410): 3x3 BLOCK Jacobi.  o0 / o1 are caller-owned device arrays
 * (n_points, 4) of dtype holding the off-diagonals (xy, xz, yz) of the vertex blocks at the current / trial point; pass
 * A then evaluates FUN | GRAD | HESS_DIAG | HESS_OFFD and REDUCE / DIRECTION apply the inverse of every vertex block
 * restricted to its free components (a block that is not positive definite falls back to the scalar rule).  psd != 0:
 * passes A and B use the eigenvalue-clamped element Hessians (APL_OP_PSD).  NULL, NULL, 0 restores the reference
 * behaviour.  Call before APL_PHASE_INIT. */
int apl_pncg_set_block_jacobi(apl_pncg_t* ws, void* o0, void* o1, int psd);
/* Sharded meshes (new functionality): with an exchange attached, every phase completes its partial results over all
 * ranks on the device, right after the kernels that produced them -- the 11 sums of REDUCE, (g.p, p.Hp) after PASS_B,
 * and per trial the halo sum of g', diag' with the trial energy (one push + one pull kernel each, guarded by the same
 * skip flags as the trial) -- so apl_pncg_iterate runs the sharded iteration, CUDA graphs included, with no host round
 * trip and no host-launched collective.  `mask` bit 1 must be clear on ghost copies; every rank must enqueue the same
 * phases (the scalars are bit-identical on all ranks, so all ranks take the same decisions).  NULL detaches. */
int apl_pncg_set_exchange(apl_pncg_t* ws, apl_xchg_t* x);
int apl_pncg_current(const apl_pncg_t* ws);
int apl_pncg_flip(apl_pncg_t* ws);
/* Enqueue one phase (for callers that interleave collectives between phases: sharded meshes). */
int apl_pncg_phase(apl_pncg_t* ws, int phase, int j, void* stream);
/* Enqueue n_iters complete iterations (phases 1..8 and the role flip); never synchronises. */
int apl_pncg_iterate(apl_pncg_t* ws, int n_iters, void* stream);

/* ---- fused adjoint solve: Jacobi-preconditioned conjugate gradients on hess_prod -------------------------
 * Replaces  jax.scipy.sparse.linalg.cg(lambda p: model.hess_prod(u, p), -dLdu, tol=1e-5, atol=1e-15,
 *                                       maxiter=n_free // 10, M=lambda x: P * x),  P = 1 / model.hess_diag(u)
 * of the reference's inverse problems (exp/2025/09/24/inverse-grin/src/35-inverse-small-reg.py:223-260), where every
 * matvec is an FFI callback and the vector algebra is XLA ops; here one iteration is five launches (matvec, p.Ap,
 * {x, r update + r.Mr + r.r}, a 1-thread step, direction) replayed as a CUDA graph, scalars on the device.
 * All vectors are caller-owned DEVICE arrays (n_points, 4) of dtype: u (the state), x (solution, in: x0 unless
 * x_is_zero), b (right-hand side), r / p / Ap (work), diag (model.hess_diag(u), raw: |d| is taken and non-positive
 * entries replaced by the mean of the positive ones); mask as in apl_pncg_create (bit 0 = free DOF; fixed DOFs of x
 * stay untouched and are excluded from every product).  scal: APL_PCG_NSCAL device doubles:
 *   [0] r.Mr  [1] p.Ap  [3] r.r  [4] b.b  [5] done: 0 running, 1 converged (|r| <= max(tol |b|, atol)), 2 max_iters,
 *   3 breakdown (p.Ap <= 0: H indefinite along p, or a non-finite value)  [6] iterations  [7] beta  [8] target^2. */
#define APL_PCG_NSCAL 16
typedef struct apl_pcg apl_pcg_t;
int apl_pcg_create(int dtype, int64_t n_points, int device, const void* u, void* x, const void* b, void* r, void* p,
                   void* Ap, const void* diag, const uint8_t* mask, double* scal, apl_pcg_t** out);
void apl_pcg_destroy(apl_pcg_t* ws);
int apl_pcg_add_fem(apl_pcg_t* ws, apl_fem_t* fem);
/* psd != 0: the matvec is the PSD-projected product (APL_OP_PSD) -- the operator is then positive semi-definite by
 * construction, what the reference asserts with lx.positive_semidefinite_tag (:248) */
int apl_pcg_set_params(apl_pcg_t* ws, double tol, double atol, int64_t max_iters, int psd, int scatter, int use_graph);
int apl_pcg_init(apl_pcg_t* ws, int x_is_zero, void* stream);
/* enqueue n_iters iterations (no-ops once scal[5] != 0); never synchronises */
int apl_pcg_iterate(apl_pcg_t* ws, int n_iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* APPLE_B200_H */
