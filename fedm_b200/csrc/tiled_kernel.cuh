// Tiled (cell-batch) assembly kernel: north_star (a) "mesh and coefficient data staged per cell-batch in shared memory".
//
// One CTA = one tile of FEDMX_TILE storage rows = a compact 2-D patch of mesh nodes (fedmx_topology.cu orders the rows that
// way) plus its ring of halo vertices.  Three phases separated by __syncthreads():
//   V  stage the vertex data of the patch in shared memory (coordinates, unknowns, coefficient fields, and the model's
//      per-node precomputations, e.g. exp(u/6)): the tile's own nodes are contiguous runs => coalesced FP64 loads;
//   C  one thread per CELL evaluates everything that is constant on the cell (geometry, gradients, field magnitude, the
//      transcendental transport coefficients and their derivatives) ONCE and leaves a record in shared memory -- the generic
//      row kernel recomputes all of that in each of the three rows that visit the cell;
//   R  one thread per ROW walks its cell fan exactly like the generic kernel (same carry / single-store scheme, hence the same
//      deterministic, atomic-free scatter), but only evaluates the quadrature moments of its own test function from the staged
//      vertices and the cell record.
// Cells on the tile boundary are evaluated by every tile that owns one of their vertices (~1.25x instead of 3x).
// A model opts in by providing  CellRec  (number of doubles per record for JAC / residual-only),  stage(),  cell<JAC>()  and
// row_rec<JAC>();  facet terms fall back to the model's full row<JAC>() for the (rare) boundary rows.
#pragma once
#include "row_kernel.cuh"

namespace {

struct TiledArgs {
    AsmArgs A;
    TileView T;
    int vtx_cap, cell_cap;     // shared-memory capacities (>= max over tiles)
};

// Shared-memory view handed to the model: structure-of-arrays, index = field * cap + local id (consecutive lanes touch
// consecutive banks)
struct TileSmem {
    double *vtx;                     // [Model::NVF][vtx_cap]
    double *rec;                     // [NREC][cell_cap]
    unsigned long long *cellv;       // [cell_cap]
    int vtx_cap, cell_cap;
    __device__ __forceinline__ double &v(int f, int i) const { return vtx[f * vtx_cap + i]; }
    __device__ __forceinline__ double &r(int f, int c) const { return rec[f * cell_cap + c]; }
};

// What the fan walk needs from the SELL view, by value (taking the address of the kernel's parameter block for a non-inlined
// call would force a per-thread copy of the whole block into local memory).
struct RowCtx {
    const int32_t *slice_ptr, *inc_ptr, *inc_width;
    const uint32_t *inc_pack, *inc_cell, *row_pack;
    double *val, *res;
    const double *x;
    const double *bc_values;       // points into shared memory (copied there by the kernel)
    int bc_comp, ident_zeros;
};

// Walks the cell fan of one storage row: `eval(cw, pack, o)` produces row `a` of the element tensor of one incident cell, the
// walk accumulates the diagonal block, finishes every off-diagonal block in the "carry" registers and stores it exactly once
// (fixed order, no atomics: the same deterministic scatter as the generic row kernel), applies the Dirichlet row and writes
// the residual.
// Loaded at kernel start, long before the row phase needs it (the global-load latency hides behind phases V and C)
struct RowPrefetch {
    uint32_t rp, pack0, cw0;
    int32_t gbase, ibase;
};
// volatile: the compiler must neither sink these loads to their first use (two barriers later) nor merge them away
__device__ __forceinline__ uint32_t ldg_now(const void *p) {
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ RowPrefetch row_prefetch(const RowCtx &C, int64_t row) {
    RowPrefetch r;
    const int64_t s = row / FEDMX_SLICE;
    r.rp = ldg_now(C.row_pack + row);
    r.gbase = int32_t(ldg_now(C.slice_ptr + s));
    r.ibase = int32_t(ldg_now(C.inc_ptr + s));
    const int64_t e = int64_t(r.ibase) + (row % FEDMX_SLICE);
    r.pack0 = ldg_now(C.inc_pack + e);           // every owned row has at least one incident cell
    r.cw0 = ldg_now(C.inc_cell + e);
    return r;
}

template <int NEQ, unsigned ZMASK, bool JAC, class Eval>
__device__ __forceinline__ void fan_walk(const RowCtx &C, const RowPrefetch &pf, int64_t row, int32_t node, Eval eval) {
    constexpr int NB = NEQ * NEQ;
    const int lane = int(row % FEDMX_SLICE);
    const uint32_t rp = pf.rp;
    const int n_inc = int(rp & 0xffu);
    const int diag_slot = int((rp >> 8) & 0xffu);
    const int bc = int((rp >> 16) & 0x7fffu);
    const int bc_comp = C.bc_comp;
    const int64_t gbase = pf.gbase;
    const int64_t ibase = pf.ibase;
    AsmArgs Aw;                                  // only .val is used by store_block / add_block
    Aw.val = C.val;
    double diag[JAC ? NB : 1], carry[JAC ? NB : 1], R[NEQ];
    if (JAC) {
#pragma unroll
        for (int e = 0; e < NB; ++e) { diag[e] = 0.0; carry[e] = 0.0; }
    }
#pragma unroll
    for (int k = 0; k < NEQ; ++k) R[k] = 0.0;
    // fan entries are prefetched TWO cells ahead (one cell of arithmetic does not cover the global-load latency)
    uint32_t pack = pf.pack0, cw = pf.cw0, pack1 = 0, cw1 = 0;
    if (n_inc > 1) {
        const int64_t e = ibase + FEDMX_SLICE + lane;
        pack1 = C.inc_pack[e];
        cw1 = C.inc_cell[e];
    }
    for (int m = 0; m < n_inc; ++m) {
        uint32_t pack2 = 0, cw2 = 0;
        if (m + 2 < n_inc) {
            const int64_t e = ibase + int64_t(m + 2) * FEDMX_SLICE + lane;
            pack2 = C.inc_pack[e];
            cw2 = C.inc_cell[e];
        }
        RowOut<NEQ> o;
        eval(cw, pack, o);
#pragma unroll
        for (int k = 0; k < NEQ; ++k) R[k] += o.R[k];
        if constexpr (JAC) {
#pragma unroll
            for (int q = 0; q < NB; ++q) diag[q] += o.J[0][q];
            if (PACK_FIRST(pack)) {
#pragma unroll
                for (int q = 0; q < NB; ++q) carry[q] = 0.0;
            }
#pragma unroll
            for (int q = 0; q < NB; ++q) carry[q] += o.J[1][q];
            if (bc) zero_row<NEQ>(carry, bc_comp);
            store_block<NEQ, ZMASK>(Aw, gbase, PACK_SLOT_P(pack), lane, carry);     // block (i, vp) complete: written once
#pragma unroll
            for (int q = 0; q < NB; ++q) carry[q] = o.J[2][q];
            if (PACK_LAST(pack)) {
                if (bc) zero_row<NEQ>(carry, bc_comp);
                if (PACK_CLOSED(pack)) add_block<NEQ, ZMASK>(Aw, gbase, PACK_SLOT_N(pack), lane, carry);
                else store_block<NEQ, ZMASK>(Aw, gbase, PACK_SLOT_N(pack), lane, carry);
            }
        }
        pack = pack1;
        cw = cw1;
        pack1 = pack2;
        cw1 = cw2;
    }
    if (bc) {
        const double g = C.bc_values[bc - 1];
#pragma unroll
        for (int k = 0; k < NEQ; ++k)
            if (k == bc_comp) R[k] = C.x[int64_t(NEQ) * node + k] - g;
        if (JAC) {
            zero_row<NEQ>(diag, bc_comp);
#pragma unroll
            for (int k = 0; k < NEQ; ++k)
                if (k == bc_comp) diag[k * NEQ + k] = 1.0;
        }
    }
    if (JAC && C.ident_zeros) {
#pragma unroll
        for (int k = 0; k < NEQ; ++k) {
            bool zero = true;
#pragma unroll
            for (int l = 0; l < NEQ; ++l) zero = zero && diag[k * NEQ + l] == 0.0;
#pragma unroll
            for (int l = 0; l < NEQ; ++l)
                if (l == k && zero) diag[k * NEQ + l] = 1.0;
        }
    }
    if (JAC) store_block<NEQ, ZMASK>(Aw, gbase, diag_slot, lane, diag);
    if (C.res) {
#pragma unroll
        for (int k = 0; k < NEQ; ++k) C.res[int64_t(NEQ) * node + k] = R[k];
    }
}

// Rows that touch a marked exterior facet (flagged in row_pack bit 31): complete per-row evaluation with the model's generic
// row<JAC>() on the staged vertices.  Rare (the boundary rows of the boundary tiles); not inlined so that its register
// footprint does not burden the main path.
template <class Model, bool JAC>
__device__ __noinline__ void boundary_row(RowCtx C, TileSmem sm, typename Model::MiniParams q, int64_t row, int32_t node, int tid) {
    const ModelParams mp = Model::expand(q);
    typename Model::V a;
    Model::load_staged(sm, tid, a);
    const RowPrefetch pf = row_prefetch(C, row);
    fan_walk<Model::NEQ, Model::ZERO_MASK, JAC>(C, pf, row, node, [&](uint32_t cw, uint32_t pack, RowOut<Model::NEQ> &o) {
        const int c = int(cw & 0xfffu), ka = int((cw >> 12) & 3u), kp = int((cw >> 14) & 3u), kn = 3 - ka - kp;
        const unsigned long long cv = sm.cellv[c];
        typename Model::V p, n;
        Model::load_staged(sm, int((cv >> (16 * kp)) & 0xffffu), p);
        Model::load_staged(sm, int((cv >> (16 * kn)) & 0xffffu), n);
        Model::template row<JAC>(mp, a, p, n, PACK_MK_N(pack), PACK_MK_P(pack), o);
    });
}

template <class Model, bool JAC, int MINB>
__global__ void __launch_bounds__(FEDMX_TILE, MINB) assemble_tiled_kernel(TiledArgs P) {
    constexpr int NEQ = Model::NEQ, NVF = Model::NVF;
    constexpr int NREC = JAC ? Model::NREC_JAC : Model::NREC_RES;
    extern __shared__ double smem_raw[];
    __shared__ double s_bc[FEDMX_MAX_BC];
    const AsmArgs &A = P.A;
    TileSmem sm;
    sm.vtx_cap = P.vtx_cap;
    sm.cell_cap = P.cell_cap;
    sm.vtx = smem_raw;
    sm.rec = sm.vtx + NVF * P.vtx_cap;
    sm.cellv = reinterpret_cast<unsigned long long *>(sm.rec + NREC * P.cell_cap);
    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t row = tile * FEDMX_TILE + tid;
    const bool active = row < A.S.n_rows;
    // ---- issue every index load up front: tile descriptor, own node id, row descriptor + first fan entry, the thread's cell
    // words and its halo vertex id; they are consumed phases later
    // (the cell words and ring-vertex ids of a tile sit at fixed strides, so none of these loads waits for the descriptor)
    const TileInfo ti = P.T.info[tile];
    const int32_t node = active ? A.S.row_node[row] : -1;
    RowCtx C;
    C.slice_ptr = A.S.slice_ptr; C.inc_ptr = A.S.inc_ptr; C.inc_width = A.S.inc_width;
    C.inc_pack = A.S.inc_pack; C.inc_cell = P.T.inc_cell; C.row_pack = A.S.row_pack;
    C.val = A.val; C.res = A.res; C.x = A.x; C.bc_values = s_bc;
    C.bc_comp = A.mp.bc_comp; C.ident_zeros = A.mp.ident_zeros;
    RowPrefetch pf;
    if (active) pf = row_prefetch(C, row);
    constexpr int CPT = 3;                                             // cells per thread held in registers (more: loop below)
    unsigned long long cwords[CPT];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int c = tid + k * FEDMX_TILE;
        cwords[k] = c < P.T.cell_stride ? P.T.cellv[tile * P.T.cell_stride + c] : 0ull;          // padding words are 0
    }
    const int32_t hnode = tid < P.T.halo_stride ? P.T.halo[tile * P.T.halo_stride + tid] : -1;      // padding ids are -1
    if (tid < FEDMX_MAX_BC) s_bc[tid] = A.mp.bc_values[tid];

    // ---- phase V: own nodes (local id = tid), then the halo ring
    {   // both vertices' loads are issued before either is consumed (clamped ids instead of branches around the loads)
        const typename Model::VtxRaw own = Model::stage_load(A, active ? node : 0);
        const typename Model::VtxRaw ring = Model::stage_load(A, hnode >= 0 ? hnode : 0);
        if (active) Model::stage_store(sm, tid, own);
        if (hnode >= 0) Model::stage_store(sm, FEDMX_TILE + tid, ring);
    }
    for (int h = tid + FEDMX_TILE; h < ti.n_halo; h += FEDMX_TILE) Model::stage(A, P.T.halo[tile * P.T.halo_stride + h], sm, FEDMX_TILE + h);
    __syncthreads();
    // ---- phase C: cell records
    {   // cells tid and tid + 128 together (every tile has more than 256 cells unless it is a ragged last one): the word of an
        // absent cell is 0 = three times local vertex 0, a degenerate but harmless cell whose record is not stored
        const unsigned long long w0 = cwords[0], w1 = cwords[1];
        const auto r0 = Model::template cell_compute<JAC>(A.mp, sm, int(w0 & 0xffffu), int((w0 >> 16) & 0xffffu), int((w0 >> 32) & 0xffffu));
        const auto r1 = Model::template cell_compute<JAC>(A.mp, sm, int(w1 & 0xffffu), int((w1 >> 16) & 0xffffu), int((w1 >> 32) & 0xffffu));
        if (tid < ti.n_cells) { sm.cellv[tid] = w0; Model::template cell_store<JAC>(sm, tid, r0); }
        if (tid + FEDMX_TILE < ti.n_cells) { sm.cellv[tid + FEDMX_TILE] = w1; Model::template cell_store<JAC>(sm, tid + FEDMX_TILE, r1); }
    }
#pragma unroll
    for (int k = 2; k < CPT; ++k) {
        const int c = tid + k * FEDMX_TILE;
        if (c < ti.n_cells) {
            const unsigned long long w = cwords[k];
            sm.cellv[c] = w;
            Model::template cell<JAC>(A.mp, sm, int(w & 0xffffu), int((w >> 16) & 0xffffu), int((w >> 32) & 0xffffu), c);
        }
    }
    for (int c = tid + CPT * FEDMX_TILE; c < ti.n_cells; c += FEDMX_TILE) {
        const unsigned long long w = P.T.cellv[tile * P.T.cell_stride + c];
        sm.cellv[c] = w;
        Model::template cell<JAC>(A.mp, sm, int(w & 0xffffu), int((w >> 16) & 0xffffu), int((w >> 32) & 0xffffu), c);
    }
    __syncthreads();
    // ---- phase R: rows
    if (!active) return;
    if (pf.rp >> 31) {
        boundary_row<Model, JAC>(C, sm, Model::compact(A.mp), row, node, tid);
        return;
    }
    typename Model::RowState st;
    Model::row_begin(A.mp, sm, tid, st);
    fan_walk<NEQ, Model::ZERO_MASK, JAC>(C, pf, row, node, [&](uint32_t cw, uint32_t pack, RowOut<NEQ> &o) {
        const int c = int(cw & 0xfffu), ka = int((cw >> 12) & 3u), kp = int((cw >> 14) & 3u), kn = 3 - ka - kp;
        const unsigned long long cv = sm.cellv[c];
        Model::template row_rec<JAC>(A.mp, sm, st, c, ka, kp, kn, int((cv >> (16 * kp)) & 0xffffu), int((cv >> (16 * kn)) & 0xffffu),
                                     PACK_FIRST(pack), o);
    });
}

template <class Model, bool JAC>
inline size_t tiled_smem_bytes(int vtx_cap, int cell_cap) {
    return size_t(Model::NVF) * vtx_cap * 8 + size_t(JAC ? Model::NREC_JAC : Model::NREC_RES) * cell_cap * 8 + size_t(cell_cap) * 8;
}

// Returns false when the tile does not fit into shared memory (the caller then uses the generic row kernel).
template <class Model>
inline bool launch_tiled(const AsmArgs &A, const TileView &T, bool jac, cudaStream_t stream) {
    if (T.n_tiles == 0) return false;
    TiledArgs P;
    P.A = A;
    P.T = T;
    P.vtx_cap = (T.max_vtx + 15) & ~15;
    P.cell_cap = (T.max_cells + 15) & ~15;
    const size_t bytes = jac ? tiled_smem_bytes<Model, true>(P.vtx_cap, P.cell_cap) : tiled_smem_bytes<Model, false>(P.vtx_cap, P.cell_cap);
    if (bytes > 200 * 1024) return false;
    static bool attr_set = false;
    // FEDMX_TILED_MINB=2: the Jacobian kernel compiled for 2 instead of 3 resident CTAs per SM (255 registers, no spills) -- tuning knob
    static const bool minb2 = getenv("FEDMX_TILED_MINB") && atoi(getenv("FEDMX_TILED_MINB")) == 2;
    if (!attr_set) {
        cudaFuncSetAttribute(assemble_tiled_kernel<Model, true, Model::TILED_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(assemble_tiled_kernel<Model, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(assemble_tiled_kernel<Model, false, Model::TILED_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    if (jac && minb2) assemble_tiled_kernel<Model, true, 2><<<unsigned(T.n_tiles), FEDMX_TILE, bytes, stream>>>(P);
    else if (jac) assemble_tiled_kernel<Model, true, Model::TILED_MINB><<<unsigned(T.n_tiles), FEDMX_TILE, bytes, stream>>>(P);
    else assemble_tiled_kernel<Model, false, Model::TILED_MINB><<<unsigned(T.n_tiles), FEDMX_TILE, bytes, stream>>>(P);
    return true;
}

}  // namespace
