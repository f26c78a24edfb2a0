// Generic row-assembly kernel of libfedmx, shared by the built-in models (fedmx_assemble.cu) and by the row kernels the
// form compiler generates at run time (fedm_b200/formcompiler.py -> plug-in .so).  See fedmx_assemble.cu for the design.
#pragma once
#include <type_traits>
#include <math.h>
#include <stdlib.h>

#include "fedmx_internal.h"

namespace {

constexpr double kTwoPi = 6.283185307179586476925286766559;

// Gauss-Legendre nodes on [0,1]
__device__ __constant__ double kGL2_t[2] = {0.21132486540518711775, 0.78867513459481288225};
__device__ __constant__ double kGL3_t[3] = {0.11270166537925831148, 0.5, 0.88729833462074168852};
__device__ __constant__ double kGL3_w[3] = {0.27777777777777777778, 0.44444444444444444444, 0.27777777777777777778};

struct AsmArgs {
    SellView S;
    const double *coords;   // [n_vertices][2]
    const double *x;        // [n_vertices][n_eq]
    const double *fields;   // [n_vertices][n_fields]
    const double *nodal;    // [n_vertices][NC] library-owned per-node precomputations of the model (may be null)
    FieldRef refs[FEDMX_MAX_REFS];   // plug-in models: coefficient k of vertex v is refs[k].ptr[v*stride + offset]
    double *val;            // matrix values (may be null for residual-only)
    double *res;            // residual (may be null)
    const int32_t *cells;   // [n_cells][3] (cell-wise path of the plug-ins)
    int64_t n_cells;
    double *scratch;        // per-cell element records (null: row-wise path)
    ModelParams mp;
};

// One vertex as seen by a model: coordinates, unknowns, coefficient fields.
template <int NEQ, int NF, int NC = 0>
struct Vtx {
    int32_t id;             // local vertex index (FFC orients non-symmetric quadrature rules by ascending index)
    double r, z;
    double u[NEQ];
    double f[NF > 0 ? NF : 1];
    double c[NC > 0 ? NC : 1];
};

template <int NEQ, int NF, int NC>
__device__ __forceinline__ void load_vertex(const AsmArgs &A, int32_t v, Vtx<NEQ, NF, NC> &o) {
    const double2 c = *reinterpret_cast<const double2 *>(A.coords + 2 * int64_t(v));
    o.id = v;
    o.r = c.x;
    o.z = c.y;
#pragma unroll
    for (int k = 0; k < NEQ; ++k) o.u[k] = __ldg(A.x + int64_t(NEQ) * v + k);
    if (A.fields) {
#pragma unroll
        for (int k = 0; k < NF; ++k) o.f[k] = __ldg(A.fields + int64_t(NF) * v + k);
    } else {
#pragma unroll
        for (int k = 0; k < NF; ++k) o.f[k] = __ldg(A.refs[k].ptr + int64_t(A.refs[k].stride) * v + A.refs[k].offset);
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) o.c[k] = __ldg(A.nodal + int64_t(NC) * v + k);
}

// Row `a` of one element tensor, local vertex order (a, p, n) = (0, 1, 2).
template <int NEQ>
struct RowOut {
    double R[NEQ];
    double J[3][NEQ * NEQ];   // J[b][k*NEQ + l] = d R_a[k] / d U_b[l]
};

// =====================================================================================================
// Generic row kernel
// =====================================================================================================
// ZMASK: bit e set = entry e of every block is structurally zero for this model (e.g. d R_e / d U_i of the streamer
// system).  dolfin keeps such entries in the pattern and so does the exported pattern, but they are never written (the
// value array is zero-initialised once) and the SpMV skips them: 1/9 less matrix traffic for the streamer model.
template <int NEQ, unsigned ZMASK>
__device__ __forceinline__ void store_block(const AsmArgs &A, int64_t group_base, int slot, int lane, const double *blk) {
    constexpr int NB = NEQ * NEQ;
    const int64_t g = group_base + int64_t(slot) * FEDMX_SLICE;
    double *dst = A.val + g * NB + lane;
#pragma unroll
    for (int e = 0; e < NB; ++e)
        if (!((ZMASK >> e) & 1u)) dst[e * FEDMX_SLICE] = blk[e];
}

template <int NEQ, unsigned ZMASK>
__device__ __forceinline__ void add_block(const AsmArgs &A, int64_t group_base, int slot, int lane, const double *blk) {
    constexpr int NB = NEQ * NEQ;
    const int64_t g = group_base + int64_t(slot) * FEDMX_SLICE;
    double *dst = A.val + g * NB + lane;
#pragma unroll
    for (int e = 0; e < NB; ++e)
        if (!((ZMASK >> e) & 1u)) dst[e * FEDMX_SLICE] += blk[e];
}

// Zero block row `comp` with static register indices (comp is a run-time value: a dynamic index would push the
// accumulators into local memory).
template <int NEQ>
__device__ __forceinline__ void zero_row(double *blk, int comp) {
#pragma unroll
    for (int k = 0; k < NEQ; ++k)
#pragma unroll
        for (int l = 0; l < NEQ; ++l)
            if (k == comp) blk[k * NEQ + l] = 0.0;
}

// SACC: the two accumulator blocks (diag, carry: 2 n_eq^2 doubles per thread) live in shared memory instead of registers
// ([entry][thread]: conflict-free), which frees 4 n_eq^2 registers for a third resident CTA per SM.
template <class Model, bool JAC, int MINB, bool SACC = false>
__global__ void __launch_bounds__(128, MINB) assemble_rows_kernel(AsmArgs A) {
    constexpr int NEQ = Model::NEQ, NF = Model::NF, NC = Model::NC, NB = NEQ * NEQ;
    constexpr unsigned ZMASK = Model::ZERO_MASK;
    __shared__ double sacc[(SACC && JAC) ? 2 * NB : 1][(SACC && JAC) ? 128 : 1];
    const int tid = threadIdx.x;
    const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t s = row / FEDMX_SLICE;
    const int lane = int(row % FEDMX_SLICE);
    if (s >= A.S.n_slices) return;
    const bool active = row < A.S.n_rows;
    const uint32_t rp = A.S.row_pack[row];
    const int n_inc = active ? int(rp & 0xffu) : 0;
    const int diag_slot = int((rp >> 8) & 0xffu);
    const int bc = int((rp >> 16) & 0x7fffu);       // value index + 1, 0 = unconstrained (bit 31: row touches a marked facet)
    const int bc_comp = A.mp.bc_comp;
    const int64_t gbase = A.S.slice_ptr[s];
    const int64_t ibase = A.S.inc_ptr[s];
    const int width = A.S.inc_width[s];

    const int32_t node = active ? A.S.row_node[row] : 0;           // storage row -> mesh node (vectors are in node order)
    Vtx<NEQ, NF, NC> va;
    if (active) load_vertex<NEQ, NF, NC>(A, node, va);

    double diag[SACC ? 1 : NB], carry[SACC ? 1 : NB], R[NEQ];
    if constexpr (SACC && JAC) {
#pragma unroll
        for (int e = 0; e < 2 * NB; ++e) sacc[e][tid] = 0.0;
    } else {
#pragma unroll
        for (int e = 0; e < NB; ++e) { diag[e] = 0.0; carry[e] = 0.0; }
    }
#pragma unroll
    for (int k = 0; k < NEQ; ++k) R[k] = 0.0;

    // Walk the fan.  Consecutive cells of a fan share an edge, so the "previous" vertex of cell m+1 is the "next" vertex
    // of cell m: only ONE new vertex is gathered per cell (its indices are fetched one cell ahead).
    Vtx<NEQ, NF, NC> p, n;
    int32_t vp1 = 0, vn1 = 0;
    uint32_t pack0 = 0, pack1 = 0;
    if (n_inc > 0) {
        const int64_t e = ibase + lane;
        const int32_t vp0 = A.S.inc_vp[e], vn0 = A.S.inc_vn[e];
        pack0 = A.S.inc_pack[e];
        load_vertex<NEQ, NF, NC>(A, vp0, p);
        load_vertex<NEQ, NF, NC>(A, vn0, n);
    }
    for (int m = 0; m < width; ++m) {
        if (m >= n_inc) continue;
        const uint32_t pack = pack0;
        const bool has1 = m + 1 < n_inc;
        if (has1) {                                                      // indices of cell m+1
            const int64_t e = ibase + int64_t(m + 1) * FEDMX_SLICE + lane;
            vp1 = A.S.inc_vp[e]; vn1 = A.S.inc_vn[e]; pack1 = A.S.inc_pack[e];
        }
        RowOut<NEQ> o;
        Model::template row<JAC>(A.mp, va, p, n, PACK_MK_N(pack), PACK_MK_P(pack), o);
#pragma unroll
        for (int k = 0; k < NEQ; ++k) R[k] += o.R[k];
        if constexpr (JAC && SACC) {
            double blk[NB];
#pragma unroll
            for (int q = 0; q < NB; ++q) sacc[q][tid] += o.J[0][q];
            const bool first = PACK_FIRST(pack);
#pragma unroll
            for (int q = 0; q < NB; ++q) blk[q] = (first ? 0.0 : sacc[NB + q][tid]) + o.J[1][q];
            if (bc) zero_row<NEQ>(blk, bc_comp);
            store_block<NEQ, ZMASK>(A, gbase, PACK_SLOT_P(pack), lane, blk);   // block (i, vp) complete: write it
            if (PACK_LAST(pack)) {
#pragma unroll
                for (int q = 0; q < NB; ++q) blk[q] = o.J[2][q];
                if (bc) zero_row<NEQ>(blk, bc_comp);
                if (PACK_CLOSED(pack)) add_block<NEQ, ZMASK>(A, gbase, PACK_SLOT_N(pack), lane, blk);
                else store_block<NEQ, ZMASK>(A, gbase, PACK_SLOT_N(pack), lane, blk);
            } else {
#pragma unroll
                for (int q = 0; q < NB; ++q) sacc[NB + q][tid] = o.J[2][q];
            }
        } else if constexpr (JAC) {
#pragma unroll
            for (int q = 0; q < NB; ++q) diag[q] += o.J[0][q];
            if (PACK_FIRST(pack)) {
#pragma unroll
                for (int q = 0; q < NB; ++q) carry[q] = 0.0;
            }
#pragma unroll
            for (int q = 0; q < NB; ++q) carry[q] += o.J[1][q];
            if (bc) zero_row<NEQ>(carry, bc_comp);
            // block (i, vp) is complete (or the partial first block of a closed fan): write it
            store_block<NEQ, ZMASK>(A, gbase, PACK_SLOT_P(pack), lane, carry);
#pragma unroll
            for (int q = 0; q < NB; ++q) carry[q] = o.J[2][q];
            if (PACK_LAST(pack)) {
                if (bc) zero_row<NEQ>(carry, bc_comp);
                if (PACK_CLOSED(pack)) add_block<NEQ, ZMASK>(A, gbase, PACK_SLOT_N(pack), lane, carry);
                else store_block<NEQ, ZMASK>(A, gbase, PACK_SLOT_N(pack), lane, carry);
            }
        }
        if (has1) {
            if (PACK_FIRST(pack1)) load_vertex<NEQ, NF, NC>(A, vp1, p);  // a new fan starts (non-manifold vertex): rare
            else p = n;
            load_vertex<NEQ, NF, NC>(A, vn1, n);
        }
        pack0 = pack1;
    }
    if (!active) return;
    double dblk[NB];
    if constexpr (JAC) {
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            if constexpr (SACC) dblk[q] = sacc[q][tid];
            else dblk[q] = diag[q];
        }
    }
    if (bc) {
        const double g = A.mp.bc_values[bc - 1];
#pragma unroll
        for (int k = 0; k < NEQ; ++k)
            if (k == bc_comp) R[k] = va.u[k] - g;
        if (JAC) {
            zero_row<NEQ>(dblk, bc_comp);
#pragma unroll
            for (int k = 0; k < NEQ; ++k)
                if (k == bc_comp) dblk[k * NEQ + k] = 1.0;
        }
    }
    if (JAC && A.mp.ident_zeros) {
        // ident_zeros (fedm/functions.py:1147, Normal_vector): rows without any contribution -- forms that live on the
        // boundary only -- become identity rows.  Mass-type forms: the row is zero iff its diagonal block row is zero.
#pragma unroll
        for (int k = 0; k < NEQ; ++k) {
            bool zero = true;
#pragma unroll
            for (int l = 0; l < NEQ; ++l) zero = zero && dblk[k * NEQ + l] == 0.0;
#pragma unroll
            for (int l = 0; l < NEQ; ++l)
                if (l == k && zero) dblk[k * NEQ + l] = 1.0;
        }
    }
    if (JAC) store_block<NEQ, ZMASK>(A, gbase, diag_slot, lane, dblk);
    if (A.res) {
#pragma unroll
        for (int k = 0; k < NEQ; ++k) A.res[int64_t(NEQ) * node + k] = R[k];
    }
}


// =====================================================================================================
// Cell-wise path (models with CELLS = true: the form compiler's plug-ins)
// =====================================================================================================
// The generic row kernel evaluates a cell's integrand once per (row, cell) pair, i.e. three times per cell; for generated
// forms (forward-mode AD, several exp/pow per quadrature point) that redundancy is the whole cost.  Here the work is split:
//   1. assemble_cells_kernel: one thread per cell evaluates the element vector/tensor ONCE (Model::cell) and writes three
//      records -- one per test vertex i: R_i[NEQ], then dR_i/dU_b for b = 0..2 (NB each) -- to A.scratch.  The scratch is
//      SoA over blocks of 32 cells: element e of record (cell, i) at ((cell / 32) * 3 REC + i REC + e) * 32 + cell % 32, so
//      every store of a warp covers 256 contiguous bytes (an AoS record per thread made the Jacobian variant write-bound:
//      8-byte pieces of 32 different sectors per instruction);
//   2. gather_rows_kernel: the row threads walk their fans in the same fixed order as the row kernel and add the records
//      up with the same carry / single-store logic, so the scatter stays deterministic.  Rows that touch a marked exterior
//      facet (row_pack bit 31; O(sqrt N) of them) evaluate Model::row themselves, facet integrals included.
template <class Model, bool JAC, int MINB>
__global__ void __launch_bounds__(128, MINB) assemble_cells_kernel(AsmArgs A) {
    constexpr int NEQ = Model::NEQ, NF = Model::NF, NC = Model::NC;
    constexpr int REC = JAC ? Model::REC_JAC : Model::REC_RES;
    const int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= A.n_cells) return;
    const int32_t i0 = __ldg(A.cells + 3 * c), i1 = __ldg(A.cells + 3 * c + 1), i2 = __ldg(A.cells + 3 * c + 2);
    Vtx<NEQ, NF, NC> v0, v1, v2;
    load_vertex<NEQ, NF, NC>(A, i0, v0);
    load_vertex<NEQ, NF, NC>(A, i1, v1);
    load_vertex<NEQ, NF, NC>(A, i2, v2);
    Model::template cell<JAC>(A.mp, v0, v1, v2, A.scratch + (c >> 5) * (3 * REC * FEDMX_SLICE) + (c & 31));
}

// One row of the gather pass.  BOUNDARY: the row touches a marked exterior facet and evaluates Model::row itself.
template <class Model, bool JAC, bool BOUNDARY>
__device__ __forceinline__ void gather_row(const AsmArgs &A, int64_t row, uint32_t rp) {
    constexpr int NEQ = Model::NEQ, NF = Model::NF, NC = Model::NC, NB = NEQ * NEQ;
    constexpr int REC = JAC ? Model::REC_JAC : Model::REC_RES;
    constexpr unsigned ZMASK = Model::ZERO_MASK;
    const int64_t s = row / FEDMX_SLICE;
    const int lane = int(row % FEDMX_SLICE);
    const int n_inc = int(rp & 0xffu);
    const int diag_slot = int((rp >> 8) & 0xffu);
    const int bc = int((rp >> 16) & 0x7fffu);
    const int bc_comp = A.mp.bc_comp;
    const int64_t gbase = A.S.slice_ptr[s];
    const int64_t ibase = A.S.inc_ptr[s];
    const int32_t node = A.S.row_node[row];

    double diag[NB], carry[NB], R[NEQ];
#pragma unroll
    for (int e = 0; e < NB; ++e) { diag[e] = 0.0; carry[e] = 0.0; }
#pragma unroll
    for (int k = 0; k < NEQ; ++k) R[k] = 0.0;

    Vtx<NEQ, NF, NC> va;
    if constexpr (BOUNDARY) load_vertex<NEQ, NF, NC>(A, node, va);
    uint32_t pack = 0, cid = 0;
    if (n_inc > 0) { pack = __ldg(A.S.inc_pack + ibase + lane); cid = __ldg(A.S.inc_cid + ibase + lane); }
    for (int m = 0; m < n_inc; ++m) {
        uint32_t pack1 = 0, cid1 = 0;
        if (m + 1 < n_inc) {                                             // indices of cell m+1
            const int64_t e = ibase + int64_t(m + 1) * FEDMX_SLICE + lane;
            pack1 = __ldg(A.S.inc_pack + e); cid1 = __ldg(A.S.inc_cid + e);
        }
        double oR[NEQ], oJ0[JAC ? NB : 1], oJ1[JAC ? NB : 1], oJ2[JAC ? NB : 1];
        if constexpr (BOUNDARY) {
            const int64_t e = ibase + int64_t(m) * FEDMX_SLICE + lane;
            Vtx<NEQ, NF, NC> p, n;
            load_vertex<NEQ, NF, NC>(A, A.S.inc_vp[e], p);
            load_vertex<NEQ, NF, NC>(A, A.S.inc_vn[e], n);
            RowOut<NEQ> o;
            Model::template row<JAC>(A.mp, va, p, n, PACK_MK_N(pack), PACK_MK_P(pack), o);
#pragma unroll
            for (int k = 0; k < NEQ; ++k) oR[k] = o.R[k];
            if constexpr (JAC) {
#pragma unroll
                for (int q = 0; q < NB; ++q) { oJ0[q] = o.J[0][q]; oJ1[q] = o.J[1][q]; oJ2[q] = o.J[2][q]; }
            }
        } else {
            const int ka = int((cid >> 2) & 3u), kp = int(cid & 3u), kn = 3 - ka - kp;
            const int64_t c = int64_t(cid >> 4);
            const double *rec = A.scratch + ((c >> 5) * (3 * REC) + ka * REC) * FEDMX_SLICE + (c & 31);
#pragma unroll
            for (int k = 0; k < NEQ; ++k) oR[k] = rec[k * FEDMX_SLICE];
            if constexpr (JAC) {
                const double *j0 = rec + (NEQ + ka * NB) * FEDMX_SLICE, *j1 = rec + (NEQ + kp * NB) * FEDMX_SLICE,
                             *j2 = rec + (NEQ + kn * NB) * FEDMX_SLICE;
#pragma unroll
                for (int q = 0; q < NB; ++q) { oJ0[q] = j0[q * FEDMX_SLICE]; oJ1[q] = j1[q * FEDMX_SLICE]; oJ2[q] = j2[q * FEDMX_SLICE]; }
            }
        }
#pragma unroll
        for (int k = 0; k < NEQ; ++k) R[k] += oR[k];
        if constexpr (JAC) {
#pragma unroll
            for (int q = 0; q < NB; ++q) diag[q] += oJ0[q];
            if (PACK_FIRST(pack)) {
#pragma unroll
                for (int q = 0; q < NB; ++q) carry[q] = 0.0;
            }
#pragma unroll
            for (int q = 0; q < NB; ++q) carry[q] += oJ1[q];
            if (bc) zero_row<NEQ>(carry, bc_comp);
            store_block<NEQ, ZMASK>(A, gbase, PACK_SLOT_P(pack), lane, carry);    // block (i, vp) complete: write it
#pragma unroll
            for (int q = 0; q < NB; ++q) carry[q] = oJ2[q];
            if (PACK_LAST(pack)) {
                if (bc) zero_row<NEQ>(carry, bc_comp);
                if (PACK_CLOSED(pack)) add_block<NEQ, ZMASK>(A, gbase, PACK_SLOT_N(pack), lane, carry);
                else store_block<NEQ, ZMASK>(A, gbase, PACK_SLOT_N(pack), lane, carry);
            }
        }
        pack = pack1;
        cid = cid1;
    }
    if (bc) {
        const double g = A.mp.bc_values[bc - 1];
#pragma unroll
        for (int k = 0; k < NEQ; ++k)
            if (k == bc_comp) R[k] = __ldg(A.x + int64_t(NEQ) * node + k) - g;
        if (JAC) {
            zero_row<NEQ>(diag, bc_comp);
#pragma unroll
            for (int k = 0; k < NEQ; ++k)
                if (k == bc_comp) diag[k * NEQ + k] = 1.0;
        }
    }
    if (JAC && A.mp.ident_zeros) {                  // see assemble_rows_kernel
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
    if (JAC) store_block<NEQ, ZMASK>(A, gbase, diag_slot, lane, diag);
    if (A.res) {
#pragma unroll
        for (int k = 0; k < NEQ; ++k) A.res[int64_t(NEQ) * node + k] = R[k];
    }
}

template <class Model, bool JAC>
__device__ __noinline__ void gather_boundary_row(const AsmArgs &A, int64_t row, uint32_t rp) {
    gather_row<Model, JAC, true>(A, row, rp);
}

template <class Model, bool JAC, int MINB>
__global__ void __launch_bounds__(128, MINB) gather_rows_kernel(const __grid_constant__ AsmArgs A) {
    const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (row >= A.S.n_rows) return;
    const uint32_t rp = A.S.row_pack[row];
    if (rp >> 31) gather_boundary_row<Model, JAC>(A, row, rp);
    else gather_row<Model, JAC, false>(A, row, rp);
}

template <class Model, class = void>
struct ModelHasCells { static constexpr bool value = false; };
template <class Model>
struct ModelHasCells<Model, std::enable_if_t<Model::CELLS>> { static constexpr bool value = true; };

// Launch helper shared by the built-in dispatch and by plug-ins.  (Host emulation builds -- tests/csrc/cuda_emul, where a
// kernel is a plain function -- bring their own sequential launcher with the same dispatch.)
#ifdef FEDMX_HOST_EMULATION
#include "emul_launch.h"
#else
template <class Model>
inline void launch_rows(const AsmArgs &A, bool jac, int minb, cudaStream_t stream) {
    const int threads = 128;
    if constexpr (ModelHasCells<Model>::value) {
        if (A.scratch && A.cells && A.S.inc_cid && (!jac || Model::CELLS_JAC)) {
            const unsigned cblocks = unsigned((A.n_cells + threads - 1) / threads);
            const unsigned rblocks = unsigned((A.S.n_rows + threads - 1) / threads);
            if (jac) {
                assemble_cells_kernel<Model, true, Model::CELL_MINB><<<cblocks, threads, 0, stream>>>(A);
                gather_rows_kernel<Model, true, 4><<<rblocks, threads, 0, stream>>>(A);
            } else {
                assemble_cells_kernel<Model, false, 4><<<cblocks, threads, 0, stream>>>(A);
                gather_rows_kernel<Model, false, 4><<<rblocks, threads, 0, stream>>>(A);
            }
            return;
        }
    }
    const int64_t rows = A.S.n_slices * FEDMX_SLICE;
    const unsigned blocks = unsigned((rows + threads - 1) / threads);
    if (jac) {
        if constexpr (2 * Model::NEQ * Model::NEQ * 128 * 8 <= 48 * 1024) {      // accumulators in (static) shared memory
            if (minb == 13) { assemble_rows_kernel<Model, true, 3, true><<<blocks, threads, 0, stream>>>(A); return; }
            if (minb == 14) { assemble_rows_kernel<Model, true, 4, true><<<blocks, threads, 0, stream>>>(A); return; }
        }
        if (minb > 10) minb -= 10;
        if (minb >= 4) assemble_rows_kernel<Model, true, 4><<<blocks, threads, 0, stream>>>(A);
        else if (minb == 3) assemble_rows_kernel<Model, true, 3><<<blocks, threads, 0, stream>>>(A);
        else assemble_rows_kernel<Model, true, 2><<<blocks, threads, 0, stream>>>(A);
    } else {
        if (minb >= 4) assemble_rows_kernel<Model, false, 4><<<blocks, threads, 0, stream>>>(A);
        else if (minb == 3) assemble_rows_kernel<Model, false, 3><<<blocks, threads, 0, stream>>>(A);
        else assemble_rows_kernel<Model, false, 2><<<blocks, threads, 0, stream>>>(A);
    }
}
#endif

}  // namespace
