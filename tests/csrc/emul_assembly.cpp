// Host emulation of the plug-in assembly kernels (tests/test_kernel_emulation_cpu.py; CUDA stand-ins in tests/csrc/cuda_emul):
// libfedmx's own topology code (fedmx_topology.cu) builds the SELL / fan / inc_cid structures of a small mesh in host memory,
// and the kernels of row_kernel.cuh run as plain functions over their grids with a GENERATED model (FORM = a form_*.cu of the
// form compiler).  Checked:
//   * cell-wise path (assemble_cells_kernel + gather_rows_kernel) == row-wise path (assemble_rows_kernel) for the Jacobian and
//     the residual, on interior rows, rows with marked exterior facets, Dirichlet rows; cell counts that do not fill the last
//     block of 32 scratch cells; owned rows < vertices (a partition with ghost vertices);
//   * the shared-memory-accumulator variant of the row kernel is bit-identical to the register variant;
//   * residual-only launches give the residual of the Jacobian launches.
#include "cuda_runtime.h"

#include "fedmx_topology.cu"
#include FORM

#include <cstdio>
#include <random>

ProfScope::ProfScope(fedmx_ctx *ctx, int k) : c(ctx), kind(k) {}
ProfScope::~ProfScope() {}
int fedmx_ensure_workspace(fedmx_ctx *, int) { return FEDMX_OK; }      // Krylov work vectors: not needed here

struct Mesh {
    std::vector<double> xy;
    std::vector<int32_t> cells, fverts, fmarks;
    int64_t nv = 0, nc = 0;
};

// nx x ny "right" grid on [0,1] x [0,2], slightly perturbed inside; exterior facets marked 1 (z = 0), 2 (z = top), 3 (r = right), 0 (axis)
static Mesh make_mesh(int nx, int ny, std::mt19937 &g) {
    Mesh m;
    std::uniform_real_distribution<double> U(-0.2, 0.2);
    m.nv = int64_t(nx + 1) * (ny + 1);
    for (int j = 0; j <= ny; ++j)
        for (int i = 0; i <= nx; ++i) {
            const bool inner = i > 0 && i < nx && j > 0 && j < ny;
            m.xy.push_back((i + (inner ? U(g) : 0.0)) / nx);
            m.xy.push_back(2.0 * (j + (inner ? U(g) : 0.0)) / ny);
        }
    auto id = [&](int i, int j) { return int32_t(j * (nx + 1) + i); };
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            const int32_t a = id(i, j), b = id(i + 1, j), c = id(i, j + 1), d = id(i + 1, j + 1);
            m.cells.insert(m.cells.end(), {a, b, d});
            m.cells.insert(m.cells.end(), {a, c, d});
        }
    m.nc = int64_t(m.cells.size() / 3);
    for (int i = 0; i < nx; ++i) {
        m.fverts.insert(m.fverts.end(), {id(i, 0), id(i + 1, 0)}); m.fmarks.push_back(1);
        m.fverts.insert(m.fverts.end(), {id(i, ny), id(i + 1, ny)}); m.fmarks.push_back(2);
    }
    for (int j = 0; j < ny; ++j) {
        m.fverts.insert(m.fverts.end(), {id(nx, j), id(nx, j + 1)}); m.fmarks.push_back(3);
        m.fverts.insert(m.fverts.end(), {id(0, j), id(0, j + 1)}); m.fmarks.push_back(0);
    }
    return m;
}

static int run_case(int nx, int ny, int64_t n_owned_request, unsigned seed) {
    constexpr int NEQ = JitModel::NEQ, NF = JitModel::NF, NB = NEQ * NEQ;
    std::mt19937 g(seed);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    Mesh m = make_mesh(nx, ny, g);
    const int64_t n_owned = n_owned_request > 0 ? std::min<int64_t>(n_owned_request, m.nv) : m.nv;
    fedmx_ctx *ctx = new fedmx_ctx();
    int rc = fedmx_mesh_set(ctx, NEQ, m.nv, m.xy.data(), m.nc, m.cells.data(), int64_t(m.fmarks.size()), m.fverts.data(),
                            m.fmarks.data(), n_owned);
    if (rc) { printf("mesh_set failed: %s\n", ctx->err.c_str()); return 1; }
    // Dirichlet rows on the last component: bottom (value 0) and top (value 1) boundary nodes
    std::vector<int32_t> bnodes, bidx;
    for (int i = 0; i <= nx; ++i) { bnodes.push_back(i); bidx.push_back(0); bnodes.push_back(int32_t(ny * (nx + 1) + i)); bidx.push_back(1); }
    rc = fedmx_dirichlet_set(ctx, NEQ - 1, int64_t(bnodes.size()), bnodes.data(), bidx.data());
    if (rc) { printf("dirichlet_set failed: %s\n", ctx->err.c_str()); return 1; }

    std::vector<double> x(size_t(m.nv) * NEQ), fields(size_t(m.nv) * std::max(NF, 1));
    for (auto &v : x) v = 0.2 + U(g);
    for (auto &v : fields) v = 0.2 + U(g);
    AsmArgs A{};
    A.S = ctx->sell;
    A.coords = ctx->d_coords;
    A.x = x.data();
    for (int k = 0; k < NF; ++k) A.refs[k] = FieldRef{fields.data(), std::max(NF, 1), k};
    A.mp = ctx->mp;
    A.mp.n_eq = NEQ;
    for (int i = 0; i < FEDMX_MAX_PARAMS; ++i) A.mp.p[i] = 0.5 + U(g);
    A.mp.bc_values[0] = 0.25;
    A.mp.bc_values[1] = 1.75;
    const size_t nval = size_t(ctx->sell.n_slots) * NB, nres = size_t(m.nv) * NEQ;
    std::vector<double> res(nres), scratch;
    auto run = [&](bool jac, int minb, bool cellwise, std::vector<double> &val_out, std::vector<double> &res_out) {
        memset(ctx->d_val, 0, nval * 8);
        std::fill(res.begin(), res.end(), -77.0);
        A.val = jac ? ctx->d_val : nullptr;
        A.res = res.data();
        A.cells = cellwise ? ctx->d_cells : nullptr;
        A.n_cells = cellwise ? ctx->n_cells : 0;
        if (cellwise) {
            scratch.assign(size_t((m.nc + 31) / 32 * 32) * 3 * JitModel::REC_JAC, 1e300);     // poison: unwritten records show up
            A.scratch = scratch.data();
        } else {
            A.scratch = nullptr;
        }
        launch_rows<JitModel>(A, jac, minb, nullptr);
        val_out.assign(ctx->d_val, ctx->d_val + nval);
        res_out = res;
    };
    std::vector<double> vr, rr, vs, rs, vc, rc_, v0, r0, v1, r1;
    run(true, 2, false, vr, rr);       // row kernel, register accumulators
    run(true, 13, false, vs, rs);      // row kernel, shared-memory accumulators
    run(false, 2, false, v0, r0);      // residual only
    int bad = 0;
    if (vs != vr || rs != rr) { printf("shared-memory accumulator variant differs\n"); ++bad; }
    if (r0 != rr) { printf("residual-only launch differs from the Jacobian launch's residual\n"); ++bad; }
    for (int64_t v = n_owned; v < m.nv; ++v)
        for (int k = 0; k < NEQ; ++k)
            if (rr[v * NEQ + k] != -77.0) { printf("ghost residual written\n"); ++bad; v = m.nv; break; }
    double worst = 0.0;
    if (JitModel::CELLS) {
        run(true, 2, true, vc, rc_);
        run(false, 2, true, v1, r1);
        if (r1 != rc_) { printf("cell-wise residual-only launch differs from the cell-wise Jacobian launch\n"); ++bad; }
        // compare per row against the row's scale (entries of one block row span many orders of magnitude)
        std::vector<double> bsr_r(ctx->h_colidx.size() * NB), bsr_c(bsr_r.size());
        memcpy(ctx->d_val, vr.data(), nval * 8); fedmx_matrix_get_bsr(ctx, bsr_r.data());
        memcpy(ctx->d_val, vc.data(), nval * 8); fedmx_matrix_get_bsr(ctx, bsr_c.data());
        for (int64_t r = 0; r < n_owned; ++r)
            for (int k = 0; k < NEQ; ++k) {
                double scale = fabs(rr[r * NEQ + k]);
                for (int32_t q = ctx->h_rowptr[r]; q < ctx->h_rowptr[r + 1]; ++q)
                    for (int l = 0; l < NEQ; ++l) scale = std::max(scale, fabs(bsr_r[size_t(q) * NB + k * NEQ + l]));
                double err = fabs(rr[r * NEQ + k] - rc_[r * NEQ + k]);
                for (int32_t q = ctx->h_rowptr[r]; q < ctx->h_rowptr[r + 1]; ++q)
                    for (int l = 0; l < NEQ; ++l)
                        err = std::max(err, fabs(bsr_r[size_t(q) * NB + k * NEQ + l] - bsr_c[size_t(q) * NB + k * NEQ + l]));
                if (!(err <= 1e-11 * scale)) {
                    if (bad < 10) printf("row %lld comp %d: cell-wise differs by %g (scale %g)\n", (long long)r, k, err, scale);
                    ++bad;
                }
                if (scale > 0) worst = std::max(worst, err / scale);
            }
    }
    printf("case %dx%d owned %lld/%lld cells %lld cellwise %d: worst rel diff %.3g, bad %d\n", nx, ny, (long long)n_owned,
           (long long)m.nv, (long long)m.nc, int(JitModel::CELLS), worst, bad);
    return bad;
}

int main() {
    int bad = 0;
    bad += run_case(7, 9, 0, 1);          // 126 cells: the last scratch block is partly filled
    bad += run_case(16, 12, 0, 2);        // several slices and tiles
    bad += run_case(11, 13, 120, 3);      // owned rows < vertices: a rank's local mesh with ghost vertices
    printf("bad %d\n", bad);
    return bad != 0;
}
