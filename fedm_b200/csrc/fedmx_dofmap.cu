// FEniCS 2019.1.0 dof numbering as an EXPORT permutation (host code; no device work).
//
// dolfin's DofMapBuilder renumbers the nodes of a function space with SCOTCH's Gibbs-Poole-Stockmeyer ordering
// (un-vendored third party of the reference; call site behind `FunctionSpace(mesh, ...)`, fedm_gd.py:116-123).  Everything
// the reference writes to disk in dof order -- the checkpoints of fedm/file_io.py:594-600, compared entry by entry in
// tests/integrated_tests/glow_discharge/test_glow_discharge.py:53-62 -- therefore depends on that numbering.  libfedmx
// computes in mesh-vertex order; this unit provides the permutation and the sparsity pattern in FEniCS order:
//   fedmx_gps_numbering       dof -> vertex map of a serial P1 space from the vertex adjacency graph (context-free)
//   fedmx_dofmap_fenics       the same for the mesh of a context (graph = its node-block pattern)
//   fedmx_pattern_get_fenics  scalar CSR pattern with rows/columns numbered  n_eq * dof(node) + component
// Algorithm (pinned bit-exactly on the golden `cell_dofs` of the glow-discharge HDF5 files):
//   1. pseudo-peripheral root: BFS from the start vertex; while the depth grows, restart from the lowest-index vertex of the
//      deepest level; keep the level structure of the final root;
//   2. number by "chains": take the next queued vertex, then keep walking to the lowest-index un-numbered neighbour in the
//      SAME level; every other un-numbered neighbour met on the way is queued (once) in the order met.
#include <algorithm>
#include <deque>
#include <vector>

#include "fedmx_internal.h"

namespace {

// level structure by breadth-first search over (rowptr, colidx); returns the depth
int bfs_levels(int64_t n, const int32_t *rowptr, const int32_t *col, int32_t root, std::vector<int32_t> &lev,
               std::vector<int32_t> &queue) {
    std::fill(lev.begin(), lev.end(), -1);
    queue.clear();
    queue.push_back(root);
    lev[root] = 0;
    int depth = 0;
    for (size_t head = 0; head < queue.size(); ++head) {
        const int32_t v = queue[head];
        for (int32_t q = rowptr[v]; q < rowptr[v + 1]; ++q) {
            const int32_t w = col[q];
            if (lev[w] < 0) {
                lev[w] = lev[v] + 1;
                depth = std::max(depth, int(lev[w]));
                queue.push_back(w);
            }
        }
    }
    (void)n;
    return depth;
}

int gps_numbering(int64_t n, const int32_t *rowptr, const int32_t *col, int32_t start, int32_t *d2v) {
    if (n <= 0) return FEDMX_EINVAL;
    std::vector<int32_t> lev(n), queue;
    queue.reserve(n);
    int32_t root = start;
    int prev = -1;
    for (;;) {
        const int depth = bfs_levels(n, rowptr, col, root, lev, queue);
        if (int64_t(queue.size()) != n) return FEDMX_EUNSUPPORTED;          // disconnected graph
        if (depth <= prev) break;
        prev = depth;
        int32_t far = -1;
        for (int64_t v = 0; v < n; ++v)
            if (lev[v] == depth) { far = int32_t(v); break; }
        root = far;
    }
    std::vector<char> numbered(n, 0), queued(n, 0);
    std::deque<int32_t> todo;
    todo.push_back(root);
    queued[root] = 1;
    int64_t count = 0;
    while (!todo.empty()) {
        int32_t v = todo.front();
        todo.pop_front();
        if (numbered[v]) continue;
        while (v >= 0) {
            numbered[v] = 1;
            d2v[count++] = v;
            int32_t next = -1;
            for (int32_t q = rowptr[v]; q < rowptr[v + 1]; ++q) {
                const int32_t w = col[q];
                if (w == v || numbered[w]) continue;
                if (next < 0 && lev[w] == lev[v]) next = w;
                else if (!queued[w]) { queued[w] = 1; todo.push_back(w); }
            }
            v = next;
        }
    }
    return count == n ? FEDMX_OK : FEDMX_EUNSUPPORTED;
}

}  // namespace

// Context-free: adjacency in CSR form with ascending neighbour lists (self entries are ignored).
extern "C" int fedmx_gps_numbering(int64_t n, const int32_t *h_rowptr, const int32_t *h_colidx, int32_t start,
                                   int32_t *h_dof_to_vertex) {
    if (!h_rowptr || !h_colidx || !h_dof_to_vertex || start < 0 || start >= n) return FEDMX_EINVAL;
    return gps_numbering(n, h_rowptr, h_colidx, start, h_dof_to_vertex);
}

extern "C" int fedmx_dofmap_fenics(fedmx_ctx *ctx, int32_t *h_dof_to_vertex) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, ctx->n_vertices > 0 && h_dof_to_vertex, "fedmx_dofmap_fenics: no mesh / null output");
    FEDMX_REQUIRE(ctx, ctx->n_owned == ctx->n_vertices, "fedmx_dofmap_fenics: serial numbering needs the un-partitioned mesh");
    int rc = gps_numbering(ctx->n_vertices, ctx->h_rowptr.data(), ctx->h_colidx.data(), 0, h_dof_to_vertex);
    if (rc) ctx->err = "fedmx_dofmap_fenics: the mesh graph is not connected";
    return rc;
}

// Scalar CSR pattern in FEniCS numbering: row n_eq*d + k holds the columns n_eq*dof(j) + l of every neighbour node j of
// vertex d2v[d] (ascending).  Call with NULL arrays for the sizes.
extern "C" int fedmx_pattern_get_fenics(fedmx_ctx *ctx, int64_t *n_rows, int64_t *nnz, int64_t *h_rowptr, int32_t *h_colidx) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, ctx->n_vertices > 0 && ctx->n_eq > 0, "fedmx_pattern_get_fenics: no mesh");
    FEDMX_REQUIRE(ctx, ctx->n_owned == ctx->n_vertices, "fedmx_pattern_get_fenics: needs the un-partitioned mesh");
    const int64_t nv = ctx->n_vertices;
    const int neq = ctx->n_eq;
    if (n_rows) *n_rows = nv * neq;
    if (nnz) *nnz = int64_t(ctx->h_colidx.size()) * neq * neq;
    if (!h_rowptr || !h_colidx) return FEDMX_OK;
    std::vector<int32_t> d2v(nv), v2d(nv);
    int rc = gps_numbering(nv, ctx->h_rowptr.data(), ctx->h_colidx.data(), 0, d2v.data());
    if (rc) { ctx->err = "fedmx_pattern_get_fenics: the mesh graph is not connected"; return rc; }
    for (int64_t d = 0; d < nv; ++d) v2d[d2v[d]] = int32_t(d);
    int64_t pos = 0;
    std::vector<int32_t> cols;
    h_rowptr[0] = 0;
    for (int64_t d = 0; d < nv; ++d) {
        const int32_t v = d2v[d];
        cols.clear();
        for (int32_t q = ctx->h_rowptr[v]; q < ctx->h_rowptr[v + 1]; ++q) cols.push_back(v2d[ctx->h_colidx[q]]);
        std::sort(cols.begin(), cols.end());
        for (int k = 0; k < neq; ++k) {
            for (int32_t c : cols)
                for (int l = 0; l < neq; ++l) h_colidx[pos++] = c * neq + l;
            h_rowptr[d * neq + k + 1] = pos;
        }
    }
    return FEDMX_OK;
}
