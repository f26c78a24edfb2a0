/*
 * fedmx.h -- C ABI of the B200-native FEDM hot path (libfedmx.so).
 *
 * This is the drop-in boundary below the reference's L4 (fedm/functions.py, fedm/file_io.py): everything
 * FEniCS 2019.1.0 / PETSc do underneath `nonlinear_solver.solve(problem, u_new.vector())`
 * (reference: fedm/functions.py:1047) is reached through these entry points.  Each declaration cites the
 * reference interface it replaces.  Conventions:
 *   - plain C symbols, no exceptions cross the boundary; every call returns 0 on success and a negative
 *     FEDMX_E* code otherwise, fedmx_last_error(ctx) returns the message;
 *   - pointers named h_* are HOST pointers (set-up data, copied once), pointers named d_* are caller-owned
 *     DEVICE pointers (e.g. torch tensors' data_ptr()); scalars are passed by value;
 *   - all kernels are launched on the context's stream (fedmx_set_stream); calls that return host scalars
 *     synchronise that stream, nothing else does;
 *   - no allocation happens in per-step calls: device workspaces are sized in the set-up calls.
 * Unknown layout: node-interleaved, dof = n_eq*node + comp (dolfin's mixed P1 layout); FP64 throughout,
 * int32 indices.
 */
#ifndef FEDMX_H
#define FEDMX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fedmx_ctx fedmx_ctx;

enum {
    FEDMX_OK = 0,
    FEDMX_EINVAL = -1,      /* bad argument / call order */
    FEDMX_ECUDA = -2,       /* CUDA runtime error */
    FEDMX_ENOMEM = -3,
    FEDMX_ENCCL = -4,
    FEDMX_EUNSUPPORTED = -5
};

/* model ids: which member of the FEDM form family the cell/facet kernels evaluate */
enum {
    FEDMX_MODEL_STREAMER = 1,   /* tests/integrated_tests/streamer_discharge/fedm_streamer.py:243-279 */
    FEDMX_MODEL_TOF = 2,        /* tests/integrated_tests/time_of_flight/fedm_tof.py:120-127 */
    FEDMX_MODEL_GLOW = 3,       /* tests/integrated_tests/glow_discharge/fedm_gd.py:349-385 */
    FEDMX_MODEL_POISSON = 4,    /* weak_form_Poisson_equation alone (initial potential), fedm_streamer.py:210-223 */
    FEDMX_MODEL_PROJECTION = 5, /* dolfin project() onto P1: fedm_tof.py:165, fedm_gd.py:309,432 */
    FEDMX_MODEL_PLUGIN = 100    /* row kernel generated from a UFL-lite form at run time (fedmx_plugin_load) */
};

/* SNES convergence reasons (subset of PETSc's SNESConvergedReason the reference can observe) */
enum {
    FEDMX_SNES_CONVERGED_FNORM_ABS = 2,
    FEDMX_SNES_CONVERGED_FNORM_RELATIVE = 3,
    FEDMX_SNES_CONVERGED_SNORM_RELATIVE = 4,
    FEDMX_SNES_DIVERGED_LINEAR_SOLVE = -3,
    FEDMX_SNES_DIVERGED_FNORM_NAN = -4,
    FEDMX_SNES_DIVERGED_MAX_IT = -5
};

/* ---- context ------------------------------------------------------------------------------------- */
/* Replaces the implicit dolfin/PETSc global state created by `import dolfin` (fedm/functions.py:8). */
int fedmx_ctx_create(int device, fedmx_ctx **out);
int fedmx_ctx_destroy(fedmx_ctx *ctx);
const char *fedmx_last_error(const fedmx_ctx *ctx);
const char *fedmx_version(void);
/* cudaStream_t passed as void*; NULL = the legacy default stream. */
int fedmx_set_stream(fedmx_ctx *ctx, void *cuda_stream);

/* ---- mesh, dof map, sparsity pattern (set-up, once) ------------------------------------------------ */
/* Replaces dolfin Mesh/MeshFunction/DofMap/SparsityPattern construction behind
 * `FunctionSpace(mesh, MixedElement([P1]*n_eq))` (fedm_streamer.py:141-143) and
 * `Marking_boundaries` output consumption (fedm/functions.py:86-124).
 * h_coords[n_vertices][2], h_cells[n_cells][3]; exterior facets: h_facet_vertices[n_facets][2],
 * h_facet_marker[n_facets] (0 = unmarked, 1..15).  Only the locally owned rows -- the first n_owned
 * vertices of the local numbering -- are assembled/stored (single GPU: n_owned = n_vertices); ghost
 * vertices follow, and all cells touching an owned vertex must be present (one ghost-cell layer).  Builds the node graph, the BSR sparsity pattern (every cell-dof x
 * cell-dof pair, as dolfin inserts them; blocks are stored even where the physics is zero), the
 * SELL-32 storage and the per-row cell fans used by the assembly kernels. */
int fedmx_mesh_set(fedmx_ctx *ctx, int n_eq, int64_t n_vertices, const double *h_coords,
                   int64_t n_cells, const int32_t *h_cells,
                   int64_t n_facets, const int32_t *h_facet_vertices, const int32_t *h_facet_marker,
                   int64_t n_owned);
/* Pattern queries (node-block CSR of the owned rows; expands to dolfin's scalar pattern with
 * dof = n_eq*node+comp).  Call with NULL arrays to get sizes. */
int fedmx_pattern_get(fedmx_ctx *ctx, int64_t *n_rows, int64_t *nnz_blocks, int32_t *h_rowptr, int32_t *h_colidx);
/* FEniCS 2019.1.0 numbering as an export permutation.  dolfin numbers the dofs of a serial P1 space with SCOTCH's
 * Gibbs-Poole-Stockmeyer ordering (DofMapBuilder behind `FunctionSpace(mesh, 'P', 1)`, fedm_gd.py:116-123); everything the
 * reference stores in dof order (the checkpoints of fedm/file_io.py:594-600, compared in
 * tests/integrated_tests/glow_discharge/test_glow_discharge.py:53-62) depends on it.  Mixed spaces: dof = n_eq*d + comp.
 * fedmx_gps_numbering is context-free host code: adjacency graph in CSR form (ascending neighbours) -> dof-to-vertex map. */
int fedmx_gps_numbering(int64_t n, const int32_t *h_rowptr, const int32_t *h_colidx, int32_t start, int32_t *h_dof_to_vertex);
int fedmx_dofmap_fenics(fedmx_ctx *ctx, int32_t *h_dof_to_vertex);
/* dolfin's scalar SparsityPattern in dolfin's own dof numbering (rows/columns n_eq*dof+comp, ascending columns). */
int fedmx_pattern_get_fenics(fedmx_ctx *ctx, int64_t *n_rows, int64_t *nnz, int64_t *h_rowptr, int32_t *h_colidx);
/* ---- 1-D P2 elements (BASELINE configs[0]: examples/time_of_flight_1D/fedm-tof_1d.py:87-104) -------------------------------
 * IntervalMesh + FunctionSpace(mesh, 'P', 2), one log-density unknown, Gamma = -grad(D e^u) + w e^u (fedm-tof_1d.py:114),
 * balance equation in Cartesian log form (fedm/functions.py:350-368).  dofs in coordinate order (2c, 2c+1, 2c+2 = left vertex,
 * midpoint, right vertex of cell c).  fedmx_p2line_assemble: BDF2 history into d_hist, residual into d_res and -- if d_band is
 * not NULL -- the pentadiagonal Jacobian into d_band[ndof][5] (column i-2+k); replaces assemble(F)/assemble(J) behind Problem.F/J.
 * fedmx_p2line_solve: banded LU without pivoting (the example's direct solver, linear_solver = 'mumps'); destroys band and rhs. */
int fedmx_p2line_assemble(fedmx_ctx *ctx, int64_t n_cells, double h, double dt, double dt_old, double D, double w,
                          const double *d_u, const double *d_u_old, const double *d_u_old1, const double *d_f,
                          double *d_hist, double *d_res, double *d_band);
int fedmx_p2line_solve(fedmx_ctx *ctx, int64_t ndof, double *d_band, double *d_rhs, double *d_x);
/* ILU(0) -- PETSc's default preconditioner behind `linear_solver = "gmres"` of the reference test
 * (tests/integrated_tests/streamer_discharge/fedm_streamer.py:32,302-305): ILU(0) in serial, block Jacobi + ILU(0) per rank
 * under mpirun.  One block per GPU (ghost columns dropped), multicolour ordering (fedm_b200/csrc/fedmx_ilu.cu).
 * fedmx_ilu0_factor: numeric factorisation of the currently assembled Jacobian (symbolic phase on first use);
 * fedmx_ilu0_solve: z = (LU)^-1 r.  Selected for the Krylov solver with fedmx_precond_set(ctx, 2, 0, 1), after which every
 * Jacobian assembly refreshes the factors. */
int fedmx_ilu0_factor(fedmx_ctx *ctx);
int fedmx_ilu0_solve(fedmx_ctx *ctx, const double *d_r, double *d_z);
int fedmx_ilu0_info(fedmx_ctx *ctx, int *n_colors, int64_t *n_blocks, int32_t *h_order /* elimination order, may be NULL */);
/* Copy the assembled Jacobian out as block CSR values [nnz_blocks][n_eq][n_eq] (host buffer). */
int fedmx_matrix_get_bsr(fedmx_ctx *ctx, double *h_values);

/* Dirichlet conditions on one component: node list + index into the value table passed at solve time.
 * Replaces `DirichletBC(ME.sub(k), value, subdomain)` (fedm_streamer.py:241, fedm_gd.py:345-347). */
int fedmx_dirichlet_set(fedmx_ctx *ctx, int comp, int64_t n_nodes, const int32_t *h_nodes, const int32_t *h_value_index);

/* ---- model (the "form compiler" output) ---------------------------------------------------------- */
/* Replaces UFL form construction + FFC/dijitso JIT of `F = action(F,u_new); J = derivative(F,u_new,u)`
 * (fedm_streamer.py:296-297): selects the hand-written cell/facet kernel family and its constants.
 * h_params: model-specific double array (layouts: the header comment of each model in fedm_b200/csrc/fedmx_assemble.cu and glow_model.cuh; host side: fedm_b200/streamer.py STREAMER_PARAMS, fedm_b200/glow.py). */
int fedmx_model_set(fedmx_ctx *ctx, int model_id, int n_params, const double *h_params);
/* Run-time compiled forms -- the role of FFC + dijitso for `F = action(F,u_new); J = derivative(F,u_new,u)` with arbitrary
 * user forms (fedm_streamer.py:296-297): fedm_b200/formcompiler.py lowers a UFL-lite form to CUDA source for the generic
 * row kernel (fedm_b200/csrc/row_kernel.cuh), nvcc builds a plug-in .so, fedmx_plugin_load binds it to the context and
 * fedmx_plugin_set_field tells the kernel where coefficient k lives (component `offset` of a node-interleaved device
 * vector with `stride` components).  Scalar parameters (dt, dt_old, Constants ...) go through fedmx_model_set(
 * FEDMX_MODEL_PLUGIN, n, values) before every solve.
 * Entry points the plug-in .so exports (checked by fedmx_plugin_load): fedmx_plugin_neq(), fedmx_plugin_nf(),
 * fedmx_plugin_args_size() (must equal the library's AsmArgs), fedmx_plugin_launch(args, jac, minb, stream), and optionally
 * fedmx_plugin_cell_doubles(jac): > 0 = the plug-in has the cell-wise path (every cell evaluated once into that many doubles
 * of library-owned scratch per cell, rows gather; row_kernel.cuh) and the library allocates the scratch on first use. */
int fedmx_plugin_load(fedmx_ctx *ctx, const char *so_path);
int fedmx_plugin_set_field(fedmx_ctx *ctx, int k, const double *d_ptr, int stride, int offset);

/* Per-node coefficient fields (P1 Functions: history terms, transport/rate coefficients...),
 * d_fields[n_vertices][n_fields], caller-owned. */
/* dolfin's A.ident_zeros() (fedm/functions.py:1147): all-zero matrix rows get a unit diagonal (boundary-only forms). */
int fedmx_ident_zeros_set(fedmx_ctx *ctx, int on);
int fedmx_fields_bind(fedmx_ctx *ctx, int n_fields, const double *d_fields);

/* ---- assembly (a1/a2: Problem.F / Problem.J, fedm/functions.py:188-202) --------------------------- */
/* b = F(x) with Dirichlet rows b_i = x_i - g (assemble + bc.apply(b, x)); h_bc_values = value table. */
int fedmx_assemble_residual(fedmx_ctx *ctx, const double *d_x, double dt, double dt_old, double t,
                            const double *h_bc_values, int n_bc_values, double *d_b);
/* A = dF/dx into the context's matrix storage, Dirichlet rows -> unit rows (assemble + bc.apply(A));
 * if d_b != NULL the residual is produced in the same pass. */
int fedmx_assemble_jacobian(fedmx_ctx *ctx, const double *d_x, double dt, double dt_old, double t,
                            const double *h_bc_values, int n_bc_values, double *d_b);
/* BDF2 history field hist = (-(1+w)^2 u_old + w^2 u_old1)/((1+w) dt), w = dt/dt_old
 * (the u_old/u_old1 part of fedm/functions.py:356-357), for the first n_hist components. */
int fedmx_bdf2_history(fedmx_ctx *ctx, int n_hist, const double *d_x_old, const double *d_x_old1,
                       double dt, double dt_old, double *d_fields, int n_fields, int field_offset);

/* ---- linear algebra (a15: PETSc Mat/Vec/KSP) -------------------------------------------------------- */
int fedmx_spmv(fedmx_ctx *ctx, const double *d_x, double *d_y);               /* y = A x (MatMult) */
int fedmx_norm2(fedmx_ctx *ctx, const double *d_x, int64_t n, double *out);   /* VecNorm(NORM_2) */
int fedmx_dot(fedmx_ctx *ctx, const double *d_x, const double *d_y, int64_t n, double *out);
/* GMRES(restart), left preconditioning (fedmx_precond_set), classical Gram-Schmidt, convergence on the
 * preconditioned residual: KSPGMRES + PCBJACOBI semantics (fedm_streamer.py:32 'gmres'). x starts at 0. */
int fedmx_gmres(fedmx_ctx *ctx, const double *d_b, double *d_x, int restart, double rtol, double atol,
                int max_it, int *iters, double *rnorm, int *converged);

/* ---- preconditioner (PETSc PC behind KSP; the reference selects it with nonlinear_solver.parameters["preconditioner"],
 * fedm_streamer.py:306-307) ----------------------------------------------------------------------------- */
/* kind 0: node-block Jacobi (default).  kind 1: two-stage -- algebraic-multigrid V(nu,nu)-cycle on component `comp`
 * (the potential: its block 2 pi r grad u . grad v, fedm/functions.py:401, depends on the mesh only) followed by
 * node-block Jacobi on the other components; the hierarchy is built once by the host (fedm_b200/amg.py) and
 * uploaded with fedmx_amg_add_level (finest first; host CSR arrays; P/R omitted on the coarsest level, whose dense
 * inverse goes to fedmx_amg_set_coarse).  fedmx_matrix_get_component copies entry (i,j) of every stored block
 * (block-CSR order) so that the host can extract that block from an assembled Jacobian. */
int fedmx_precond_set(fedmx_ctx *ctx, int kind, int comp, int nu);
/* z = M^-1 r with the current preconditioner (PETSc PCApply); r and z are distinct device vectors of the state layout. */
int fedmx_pc_apply(fedmx_ctx *ctx, const double *d_r, double *d_z);
int fedmx_amg_reset(fedmx_ctx *ctx);
int fedmx_amg_add_level(fedmx_ctx *ctx, int64_t n, const double *h_dinv, double omega,
                        const int32_t *a_rowptr, const int32_t *a_col, const double *a_val, int64_t n_coarse,
                        const int32_t *p_rowptr, const int32_t *p_col, const double *p_val,
                        const int32_t *r_rowptr, const int32_t *r_col, const double *r_val);
/* Finest-level options of the multigrid stage: v01 != 0 selects a V(0,1) cycle there (restrict the right-hand side, smooth the
 * prolongated correction once); with a prolongation that also holds the rows of the ghost vertices (n_rows = local vertices)
 * a partitioned V-cycle then needs no halo exchange at all, only the allreduce of the restricted right-hand side. */
int fedmx_amg_set_fine_prolongation(fedmx_ctx *ctx, int64_t n_rows, const int32_t *p_rowptr, const int32_t *p_col,
                                    const double *p_val, int v01);
int fedmx_amg_set_coarse(fedmx_ctx *ctx, int n, const double *h_inverse);
int fedmx_matrix_get_component(fedmx_ctx *ctx, int i, int j, double *h_values);

/* ---- Newton (a14: PETScSNESSolver.solve, fedm/functions.py:1047) ----------------------------------- */
typedef struct {
    double rtol, atol, stol;        /* SNES: relative_tolerance, absolute_tolerance (1e-10), solution tol (1e-16) */
    int max_it;                     /* maximum_iterations */
    int ksp_restart;                /* 30 */
    double ksp_rtol, ksp_atol;      /* 1e-5, 1e-50 (PETSc defaults); tighten for 'mumps'-grade parity runs */
    int ksp_max_it;                 /* 10000 */
} fedmx_newton_params;

typedef struct {
    int iterations;                 /* Newton iterations */
    int reason;                     /* FEDMX_SNES_* */
    int linear_iterations;          /* total Krylov iterations */
    double fnorm0, fnorm;           /* ||F(x0)||_2, final ||F||_2 */
    double fnorm_history[64];
} fedmx_newton_result;

/* Solves F(x)=0 in place (d_x: initial guess in, solution out). Never throws: the host wrapper raises
 * RuntimeError when reason < 0, which triggers the reference's retry path (fedm/functions.py:1080). */
int fedmx_newton_solve(fedmx_ctx *ctx, double *d_x, double dt, double dt_old, double t,
                       const double *h_bc_values, int n_bc_values,
                       const fedmx_newton_params *params, fedmx_newton_result *result);

/* ---- time-step control (a12/a13) ------------------------------------------------------------------- */
/* err = ||a - b + eps||_2 / ||b + eps||_2 over component comp of two node-interleaved vectors
 * (fedm/functions.py:1062-1064; eps = DOLFIN_EPS added entry-wise). */
int fedmx_error_norm(fedmx_ctx *ctx, const double *d_new, const double *d_old, int comp, double eps, double *err);
/* Same estimator on arbitrary strided views (per-field vectors: stride 1, offset 0), n entries (owned nodes). */
int fedmx_error_norm_strided(fedmx_ctx *ctx, const double *d_new, const double *d_old, int64_t stride, int64_t offset,
                             int64_t n, double eps, double *err);
/* PID / PI.3.4 / H211b controllers (fedm/functions.py:915-951); host libm, bit-exact with the reference. */
double fedmx_pid_step(double dt, const double err[3], double tol, double dt_min, double dt_max);
double fedmx_pi34_step(double dt, const double err[3], double tol, double dt_min, double dt_max);
double fedmx_h211b_step(double dt, double dt_old, const double err[3], double tol, double dt_min, double dt_max);

/* ---- coefficient interpolation (a11: np.interp on dof vectors, fedm/functions.py:627-636, 739-747) - */
/* out[i*out_stride] = np.interp(x[i*x_stride], kx, ky) / divisor for i < n (clamped, linear; the reference
 * divides transport tables by N0, fedm/functions.py:627); tables are device arrays of n_table entries. */
int fedmx_interp_table(fedmx_ctx *ctx, const double *d_x, int64_t x_stride, int64_t n,
                       const double *d_kx, const double *d_ky, int n_table, double divisor,
                       double *d_out, int64_t out_stride);

/* ---- vector plumbing (a16: FunctionAssigner.assign / Function.assign) ----------------------------- */
/* dst[i*dst_stride + dst_off] = src[i*src_stride + src_off], i < n (mixed <-> per-field split/merge). */
int fedmx_strided_copy(fedmx_ctx *ctx, const double *d_src, int64_t src_stride, int64_t src_off,
                       double *d_dst, int64_t dst_stride, int64_t dst_off, int64_t n);

/* dst[i*dst_stride+dst_off] = mul * src[i*src_stride+src_off] / div : the 'ESR' rule D = kB*Tgas*mu/e
 * (fedm/functions.py:633) and other per-node scalings, same operation order as the reference. */
int fedmx_strided_scale(fedmx_ctx *ctx, const double *d_src, int64_t src_stride, int64_t src_off,
                        double *d_dst, int64_t dst_stride, int64_t dst_off, int64_t n, double mul, double div);
/* dst[i*dst_stride+dst_off] = exp(x[i*n_eq+comp_a] - x[i*n_eq+comp_b]) : mean electron energy from ln w_e and ln n_e
 * (fedm_gd.py:452). */
int fedmx_exp_diff(fedmx_ctx *ctx, const double *d_x, int n_eq, int comp_a, int comp_b,
                   double *d_dst, int64_t dst_stride, int64_t dst_off, int64_t n);

/* ---- multi-GPU (north_star (f): mesh partition, halo exchange, allreduce over NCCL) ---------------- */
/* The host creates the id on rank 0 (fedmx_comm_unique_id), broadcasts the 128 bytes with its own
 * plumbing (torch.distributed) and every rank calls fedmx_comm_init.  Replaces MPI inside dolfin/PETSc. */
int fedmx_comm_unique_id(char id_out[128]);
int fedmx_comm_init(fedmx_ctx *ctx, int n_ranks, int rank, const char id[128]);
/* Halo plan: for every neighbour rank the owned nodes to send and the ghost nodes to receive
 * (local node indices; CSR over neighbours). */
int fedmx_halo_set(fedmx_ctx *ctx, int n_neighbors, const int32_t *h_neighbor_ranks,
                   const int32_t *h_send_ptr, const int32_t *h_send_nodes,
                   const int32_t *h_recv_ptr, const int32_t *h_recv_nodes);
/* Ghost update of a node-interleaved vector with n_comp components per node (update_ghost_values(),
 * fedm/functions.py:625; VecGhostUpdate inside MatMult). */
int fedmx_halo_exchange(fedmx_ctx *ctx, double *d_vec, int n_comp);
int fedmx_allreduce_sum(fedmx_ctx *ctx, double *d_buf, int n);

/* NVLink peer-memory transport for the collectives above (CUDA IPC symmetric heap, one per context): once opened,
 * fedmx_halo_exchange / fedmx_allreduce_sum and the solver's internal collectives store straight into the
 * neighbours' memory and synchronise with release/acquire flags instead of going through NCCL (latency: an NVLink
 * round trip instead of a collective launch).  Replaces the same MPI calls as above.
 *   create: allocates the heap (halo_capacity, big_capacity in doubles, identical on all ranks) -> 64-byte IPC handle
 *   open:   handles of all ranks, rank-major [n_ranks][64] (gathered by the host's own plumbing)
 *   halo_set: per neighbour (order of fedmx_halo_set) the node offset of this rank's segment in the NEIGHBOUR's
 *             receive list (= the neighbour's recv_ptr entry for this rank)
 *   enable: switch the peer transport on/off (off = NCCL);  error: 1 after a flag wait timed out */
int fedmx_peer_create(fedmx_ctx *ctx, int64_t halo_capacity, int64_t big_capacity, char handle_out[64]);
int fedmx_peer_open(fedmx_ctx *ctx, const char *handles);
int fedmx_peer_halo_set(fedmx_ctx *ctx, int n_neighbors, const int32_t *h_remote_offsets);
int fedmx_peer_enable(fedmx_ctx *ctx, int on);
int fedmx_peer_error(fedmx_ctx *ctx);

/* ---- instrumentation ------------------------------------------------------------------------------- */
/* Number of kernels this library launched since the last call with reset != 0. */
int64_t fedmx_launch_count(fedmx_ctx *ctx, int reset);
/* Optional per-kernel-class device timing with CUDA events on the context's stream (used by bench.py for the
 * roofline numbers).  kinds: 0 = Jacobian assembly, 1 = residual assembly, 2 = SpMV(+block-Jacobi),
 * 3 = Gram-Schmidt multi-dot + multi-axpy, 4 = two-stage preconditioner (block Jacobi + AMG V-cycle).  fedmx_profile_get synchronises, returns launches and total ms
 * since the last reset and resets the class. */
enum { FEDMX_PROF_ASM_JAC = 0, FEDMX_PROF_ASM_RES = 1, FEDMX_PROF_SPMV = 2, FEDMX_PROF_ORTHO = 3, FEDMX_PROF_PRECOND = 4,
       /* finer break-down (nested inside the classes above): multigrid finest level / separately launched coarser
        * levels / cluster-fused coarsest levels, and the collectives (halo exchanges + allreduces) */
       FEDMX_PROF_MG_FINE = 5, FEDMX_PROF_MG_MID = 6, FEDMX_PROF_MG_COARSE = 7, FEDMX_PROF_COMM = 8,
       FEDMX_PROF_KINDS = 9 };
int fedmx_profile_enable(fedmx_ctx *ctx, int on);
int fedmx_profile_get(fedmx_ctx *ctx, int kind, int64_t *count, double *total_ms);

#ifdef __cplusplus
}
#endif
#endif /* FEDMX_H */
