// Internal declarations of libfedmx (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/fedmx.h"

#define FEDMX_MAX_EQ 5
#define FEDMX_MAX_PARAMS 128
#define FEDMX_MAX_BC 16
#define FEDMX_SLICE 32
#define FEDMX_MAX_REFS 48
#define FEDMX_MAX_RANKS 8
#define FEDMX_PEER_AR_MAX 128

// Indirect coefficient reference of a plug-in model: component `offset` of a node-interleaved vector with `stride` comps
struct FieldRef {
    const double *ptr;
    int stride, offset;
};

// Device-side view of the mesh topology in SELL-32 (sliced ELLPACK, 32 rows per slice) form.
// Matrix block (row r = 32*s + l, k-th stored column) lives at block slot  slice_ptr[s] + 32*k + l ;
// its NB = n_eq*n_eq values are stored entry-major inside the (slice,k) group:
//     val[(slice_ptr[s] + 32*k) * NB + e*32 + l]
// so that a warp working on 32 consecutive rows reads/writes 256 contiguous bytes per entry.
struct SellView {
    int64_t n_rows;          // owned rows (nodes)
    int64_t n_slices;
    int64_t n_slots;         // total block slots incl. padding
    const int32_t *slice_ptr;   // [n_slices+1] block-slot offsets
    const int32_t *slice_width; // [n_slices]
    const int32_t *col;         // [n_slots] column node (padding: the row itself)
    // per-row cell fans (same slicing): entry (row, m) at inc_ptr[s] + 32*m + l
    const int32_t *inc_ptr;     // [n_slices+1]
    const int32_t *inc_width;   // [n_slices]
    const int32_t *inc_vp;      // neighbour vertex shared with the previous cell of the fan
    const int32_t *inc_vn;      // neighbour vertex shared with the next cell of the fan
    const uint32_t *inc_pack;   // see PACK_* below
    const uint32_t *inc_cid;    // cell << 4 | ka << 2 | kp: the cell of the fan entry and the positions of the row's node / of vp in it
    const uint32_t *row_pack;   // [n_slices*32]: n_inc | diag_slot<<8 | bc_value_index+1 <<16 | (row touches a marked exterior facet) <<31
    // Storage row r (slice r/32, lane r%32) holds mesh node row_node[r].  The rows are ordered in spatially compact tiles
    // of FEDMX_TILE consecutive rows (recursive coordinate bisection, fedmx_topology.cu), so that one CTA of the tiled
    // assembly kernel -- and one L1 -- sees a 2-D patch of the mesh; vectors stay in the caller's node order.
    const int32_t *row_node;    // [n_slices*32], padding rows: -1
};

#define FEDMX_TILE 128

// Cell batches of the tiled assembly kernel: tile t = storage rows [128 t, 128 t + 128).  Its vertex set is the tile's own
// nodes (local index = row - 128 t) followed by `n_halo` ring vertices; its cells are all cells incident to its rows.
struct TileInfo {
    int32_t n_cells, n_halo;
};
struct TileView {
    int64_t n_tiles = 0;
    int max_cells = 0, max_vtx = 0;
    int64_t cell_stride = 0, halo_stride = 0;    // tile t: cellv[t*cell_stride ...], halo[t*halo_stride ...] (fixed strides: the
                                                 // kernel can issue these loads without waiting for the tile descriptor)
    const TileInfo *info = nullptr;
    const int32_t *halo = nullptr;          // node ids of the ring vertices (padding: -1)
    const unsigned long long *cellv = nullptr;   // per cell: local vertex indices v0 | v1<<16 | v2<<32 (padding: 0)
    const uint32_t *inc_cell = nullptr;     // fan layout (like inc_pack): local cell | ka<<12 | kp<<14  (ka/kp: position of the row's
                                            // node / of the fan's "previous" vertex inside the cell)
};

// inc_pack bit layout
#define PACK_SLOT_P(p) ((p)&0xffu)
#define PACK_SLOT_N(p) (((p) >> 8) & 0xffu)
#define PACK_MK_N(p) (((p) >> 16) & 0xfu)   // marker of exterior facet (a,p)  [opposite n]
#define PACK_MK_P(p) (((p) >> 20) & 0xfu)   // marker of exterior facet (a,n)  [opposite p]
#define PACK_FIRST(p) (((p) >> 24) & 1u)
#define PACK_LAST(p) (((p) >> 25) & 1u)
#define PACK_CLOSED(p) (((p) >> 26) & 1u)

struct ModelParams {
    int model_id;
    int n_eq;
    int n_fields;
    int bc_comp;                 // component carrying Dirichlet conditions (-1: none)
    double p[FEDMX_MAX_PARAMS];
    double bc_values[FEDMX_MAX_BC];
    double dt, dt_old, t;
    int ident_zeros;             // dolfin's A.ident_zeros(): an all-zero matrix row gets a unit diagonal
    int pad_;
};

// Scalar SELL-32 matrix (multigrid operators): entry (row r = 32 s + l, k) at slice_ptr[s] + 32 k + l
struct ScalarSell {
    int64_t n_rows = 0, n_cols = 0, n_slices = 0;
    const int64_t *slice_ptr = nullptr;
    const int32_t *slice_width = nullptr;
    const int32_t *col = nullptr;
    const float *val = nullptr;     // multigrid operators are stored in FP32 (vectors and arithmetic stay FP64): the
                                    // preconditioner is still one fixed linear operator, so GMRES is unaffected, and
                                    // the V-cycle moves 8 instead of 12 bytes per nonzero
    // compact form of the finest operator (the only multigrid array that does not fit the L2): 4 bytes per nonzero.
    // val16 = a_ij / a_ii in FP16 (row-scaled: |.| <= O(1), no range problem near the axis where 2 pi r -> 0),
    // col16 = column - row, rdiag = a_ii in FP64.  Used when every offset fits 16 bits (one GPU; node-ordered meshes).
    const unsigned short *val16 = nullptr;
    const short *col16 = nullptr;
    const double *rdiag = nullptr;
};

struct AmgLevel {
    int64_t n = 0, n_coarse = 0;
    double omega = 1.0;
    ScalarSell A, P, R;
    double *dinv = nullptr, *b = nullptr, *x = nullptr, *x2 = nullptr, *r = nullptr;
};

struct HaloPlan {
    int n_neighbors = 0;
    std::vector<int> ranks;
    std::vector<int> send_ptr, recv_ptr;
    int32_t *d_send_nodes = nullptr, *d_recv_nodes = nullptr;
    double *d_send_buf = nullptr, *d_recv_buf = nullptr;
    int64_t n_send = 0, n_recv = 0;
};

// Symmetric peer heap (CUDA IPC) for the NVLink collectives of fedmx_peer.cu
struct PeerState {
    bool on = false, opened = false, halo_ready = false;
    char *base = nullptr;
    size_t bytes = 0;
    char *bases[FEDMX_MAX_RANKS] = {};
    size_t off_flags = 0, off_ar_slots = 0, off_halo = 0, off_big_in = 0, off_big_out = 0;
    int64_t halo_cap = 0, big_cap = 0;
    unsigned long long halo_epoch = 0, ar_epoch = 0, big_epoch = 0;
    std::vector<int32_t> remote_off;
    unsigned int *d_counter = nullptr;
    int *h_err = nullptr, *d_err = nullptr;
};

struct IluPlan;

struct fedmx_ctx {
    int device = 0;
    cudaStream_t stream = 0;
    bool own_stream = false;
    std::string err;
    int64_t launches = 0;

    // mesh / topology
    int n_eq = 0;
    int64_t n_vertices = 0;      // local vertices incl. ghosts
    int64_t n_owned = 0;         // owned vertices are numbered first
    double *d_coords = nullptr;  // [n_vertices][2]
    int32_t *d_cells = nullptr;  // [n_cells][3] local vertex indices
    int64_t n_cells = 0;
    double *d_cell_scratch = nullptr;   // per-cell element records of plug-in models that assemble cell-wise (row_kernel.cuh)
    size_t cell_scratch_doubles = 0;
    bool cell_scratch_refused = false;  // allocation failed or over the cap once: row-wise assembly from then on
    SellView sell{};
    TileView tiles{};
    std::vector<int32_t> h_row_node, h_node_row;   // storage row <-> owned node
    std::vector<void *> owned_device;   // everything cudaMalloc'ed by set-up calls
    std::vector<void *> amg_device;     // multigrid hierarchy: released by fedmx_amg_reset (re-running set-up must not leak)
    std::vector<void *> *alloc_pool = nullptr;   // where fedmx_dev_alloc records (default: owned_device)
    // host copies for pattern export
    std::vector<int32_t> h_rowptr, h_colidx;      // node-block CSR of the owned rows
    std::vector<int32_t> h_slice_ptr;
    // matrix values + block-Jacobi inverse
    double *d_val = nullptr;     // [n_slots * NB]
    double *d_dinv = nullptr;    // [n_slices*32 * NB] entry-major per slice
    // model
    ModelParams mp{};
    unsigned zero_mask = 0;      // structurally zero block entries of the model (never written, skipped by the SpMV)
    const double *d_fields = nullptr;
    // run-time compiled model (form compiler plug-in)
    void *plugin_handle = nullptr;
    int (*plugin_launch)(const void *args, int jac, int minb, void *stream) = nullptr;
    int (*plugin_cell_doubles)(int jac) = nullptr;   // optional: doubles per cell of the plug-in's cell-wise path (0: none)
    FieldRef plugin_refs[FEDMX_MAX_REFS] = {};
    double *d_nodal = nullptr;   // library-owned per-node precomputations of the model (e.g. exp(u/6))
    // Dirichlet
    std::vector<int32_t> bc_nodes, bc_value_index;
    int bc_comp = -1;
    // work vectors (sized n_eq * n_vertices)
    double *d_res = nullptr, *d_dx = nullptr, *d_w = nullptr, *d_V = nullptr;
    int krylov_restart_alloc = 0;
    double *d_red = nullptr;     // reduction scratch
    double *d_hess = nullptr;    // small device scalars for GMRES
    double *h_pinned = nullptr;  // pinned host scratch
    int64_t red_blocks = 0;
    // preconditioner: 0 = node-block Jacobi, 1 = two-stage (AMG on component amg_comp + block Jacobi), 2 = multicolour ILU(0)
    IluPlan *ilu = nullptr;
    int newton_expected_its = 0;   // Newton iterations of the previous converged solve (predictor for speculative J assembly)
    int precond_kind = 0, amg_comp = 0, amg_nu = 1;
    cudaGraphExec_t mg_graph[2] = {nullptr, nullptr};   // replicated multigrid levels, one graph per x/x2 ping-pong parity
    int mg_parity = 0, mg_graph_nodes = 0;
    bool mg_graph_off = false;       // stream capture was refused once: plain launches from then on
    bool amg_v01 = false;            // finest multigrid level runs a V(0,1) cycle (see fedmx_amg_vcycle)
    std::vector<AmgLevel> amg_levels;
    double *amg_coarse_inv = nullptr;
    double *d_tmp = nullptr;
    void *d_coarse_levels = nullptr;   // device array of fused coarse-level descriptors
    int coarse_first = -1, coarse_count = 0;
    unsigned int *d_grid_barrier = nullptr;
    int coarse_grid_blocks = 0;        // > 0: the coarse levels run in one cooperative whole-GPU kernel of this many CTAs
    // profiling (CUDA events around kernel classes)
    int prof_on = 0;                         // 0 off, 1 kernel classes, 2 + nested break-down (more event records)
    std::vector<cudaEvent_t> prof_ev[FEDMX_PROF_KINDS];     // start/stop pairs per kind
    size_t prof_used[FEDMX_PROF_KINDS] = {};
    int64_t prof_count[FEDMX_PROF_KINDS] = {};
    double prof_ms[FEDMX_PROF_KINDS] = {};
    // comm
    void *nccl_comm = nullptr;
    int n_ranks = 1, rank = 0;
    HaloPlan halo;
    PeerState peer;
};

#define FEDMX_CHECK_CUDA(ctx, expr)                                                          \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                 \
            return FEDMX_ECUDA;                                                              \
        }                                                                                    \
    } while (0)

#define FEDMX_REQUIRE(ctx, cond, msg)                                                        \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            (ctx)->err = std::string(msg);                                                   \
            return FEDMX_EINVAL;                                                             \
        }                                                                                    \
    } while (0)

template <class T>
int fedmx_dev_alloc(fedmx_ctx *ctx, T **ptr, size_t count) {
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, (count ? count : 1) * sizeof(T));
    if (e != cudaSuccess) {
        ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
        return FEDMX_ENOMEM;
    }
    (ctx->alloc_pool ? ctx->alloc_pool : &ctx->owned_device)->push_back(p);
    *ptr = (T *)p;
    return FEDMX_OK;
}

// RAII: allocations inside the scope are recorded in `pool` (e.g. the multigrid hierarchy's own list)
struct PoolScope {
    fedmx_ctx *c;
    PoolScope(fedmx_ctx *ctx, std::vector<void *> *pool) : c(ctx) { c->alloc_pool = pool; }
    ~PoolScope() { c->alloc_pool = nullptr; }
};

template <class T>
int fedmx_upload(fedmx_ctx *ctx, T **ptr, const std::vector<T> &h) {
    int rc = fedmx_dev_alloc(ctx, ptr, h.size());
    if (rc) return rc;
    if (!h.empty()) FEDMX_CHECK_CUDA(ctx, cudaMemcpy(*ptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return FEDMX_OK;
}

// RAII helper: records an event pair around a group of launches when profiling is on
struct ProfScope {
    fedmx_ctx *c;
    int kind;
    cudaEvent_t stop = nullptr;
    ProfScope(fedmx_ctx *ctx, int k);
    ~ProfScope();
};
int fedmx_profile_flush(fedmx_ctx *ctx, int kind);

// internal entry points shared between translation units
int fedmx_ensure_workspace(fedmx_ctx *ctx, int restart);
int fedmx_update_block_jacobi(fedmx_ctx *ctx);
int fedmx_spmv_precond(fedmx_ctx *ctx, const double *d_x, double *d_y, bool precond, double *d_rhs_c = nullptr, int comp = 0);
int fedmx_two_stage_after_spmv(fedmx_ctx *ctx, const double *d_v, double *d_z);
int fedmx_norm2_local(fedmx_ctx *ctx, const double *d_x, int64_t n, double *out, bool global);
int fedmx_precond_apply(fedmx_ctx *ctx, const double *d_r, double *d_z);
int fedmx_peer_halo_exchange(fedmx_ctx *ctx, double *d_vec, int n_comp);
int fedmx_peer_allreduce(fedmx_ctx *ctx, double *d_buf, int64_t n);
int fedmx_peer_destroy(fedmx_ctx *ctx);
int fedmx_comm_destroy_internal(fedmx_ctx *ctx);
int fedmx_ilu0_factor_internal(fedmx_ctx *ctx);
int fedmx_ilu0_solve_internal(fedmx_ctx *ctx, const double *d_r, double *d_z);
void fedmx_ilu0_free(fedmx_ctx *ctx);
int fedmx_bdf2_history_raw(fedmx_ctx *ctx, int64_t n, const double *d_x_old, const double *d_x_old1, double dt, double dt_old,
                           double *d_out);
double *fedmx_peer_big_buffer(fedmx_ctx *ctx, int64_t n);
