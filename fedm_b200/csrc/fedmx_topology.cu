// Host-side set-up: node graph, BSR sparsity pattern in SELL-32 form, per-row cell fans.
//
// Replaces dolfin's DofMap/SparsityPattern construction (reference call site: FunctionSpace(mesh,
// MixedElement(...)), tests/integrated_tests/streamer_discharge/fedm_streamer.py:141-143) and the
// facet-marker consumption of `ds(i)` measures (fedm_streamer.py:127-133).  The pattern is the one dolfin
// builds: every (cell dof, cell dof) pair, i.e. a full n_eq x n_eq block for every vertex-vertex adjacency
// (plus the diagonal) -- nothing is pruned by value.
//
// For the assembly kernels every owned row gets its incident cells ordered as a *fan* (consecutive cells
// share an edge).  Walking the fan, the block (i,j) of an interior edge is complete after its two cells,
// so each matrix block is produced in registers and written exactly once, deterministically.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <stdlib.h>
#include <string>

#include "fedmx_internal.h"

namespace {

struct Inc {
    int32_t vp, vn;
    uint32_t pack;
    int32_t cell;
};

// Recursive coordinate bisection of the owned nodes into tiles of FEDMX_TILE nodes (the last one may be smaller): every split
// puts a whole number of tiles on the left and cuts the longer side, measured in GRID units (dense ranks of the coordinates,
// so that graded tensor-product meshes tile like uniform ones).  order[] is permuted in place.
struct Rcb {
    const std::vector<double> &ux, &uy;
    std::vector<int32_t> &order;
    void split(int64_t lo, int64_t hi) {
        const int64_t n = hi - lo;
        if (n <= FEDMX_TILE) return;
        const int64_t tiles = (n + FEDMX_TILE - 1) / FEDMX_TILE;
        const int64_t left = (tiles / 2) * FEDMX_TILE;
        double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
        for (int64_t i = lo; i < hi; ++i) {
            const int32_t v = order[i];
            x0 = std::min(x0, ux[v]); x1 = std::max(x1, ux[v]);
            y0 = std::min(y0, uy[v]); y1 = std::max(y1, uy[v]);
        }
        const bool cut_x = (x1 - x0) > (y1 - y0);
        auto less = [&](int32_t a, int32_t b) {
            const double ka = cut_x ? ux[a] : uy[a], kb = cut_x ? ux[b] : uy[b];
            if (ka != kb) return ka < kb;
            const double la = cut_x ? uy[a] : ux[a], lb = cut_x ? uy[b] : ux[b];
            if (la != lb) return la < lb;
            return a < b;
        };
        std::nth_element(order.begin() + lo, order.begin() + lo + left, order.begin() + hi, less);
        split(lo, lo + left);
        split(lo + left, hi);
    }
};

std::vector<double> dense_rank(const double *coords, int comp, int64_t n) {
    std::vector<double> vals(n);
    for (int64_t i = 0; i < n; ++i) vals[i] = coords[2 * i + comp];
    std::vector<double> uniq(vals);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    std::vector<double> r(n);
    for (int64_t i = 0; i < n; ++i) r[i] = double(std::lower_bound(uniq.begin(), uniq.end(), vals[i]) - uniq.begin());
    return r;
}

inline int find_slot(const std::vector<int32_t> &colidx, int32_t lo, int32_t hi, int32_t col) {
    const int32_t *b = colidx.data() + lo, *e = colidx.data() + hi;
    const int32_t *it = std::lower_bound(b, e, col);
    return (it != e && *it == col) ? int(it - b) : -1;
}

}  // namespace

extern "C" int fedmx_mesh_set(fedmx_ctx *ctx, int n_eq, int64_t n_vertices, const double *h_coords,
                              int64_t n_cells, const int32_t *h_cells, int64_t n_facets,
                              const int32_t *h_facet_vertices, const int32_t *h_facet_marker, int64_t n_owned) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, n_eq >= 1 && n_eq <= FEDMX_MAX_EQ, "fedmx_mesh_set: n_eq out of range");
    FEDMX_REQUIRE(ctx, n_vertices > 0 && n_cells > 0 && h_coords && h_cells, "fedmx_mesh_set: empty mesh");
    FEDMX_REQUIRE(ctx, n_owned > 0 && n_owned <= n_vertices, "fedmx_mesh_set: bad n_owned");
    FEDMX_REQUIRE(ctx, ctx->n_vertices == 0, "fedmx_mesh_set: mesh already set (create a new context)");
    FEDMX_REQUIRE(ctx, n_vertices < (int64_t(1) << 31) / FEDMX_MAX_EQ, "fedmx_mesh_set: too many vertices for int32 dofs");
    for (int64_t c = 0; c < 3 * n_cells; ++c)
        FEDMX_REQUIRE(ctx, h_cells[c] >= 0 && h_cells[c] < n_vertices, "fedmx_mesh_set: cell vertex out of range");

    const int64_t NV = n_vertices, NO = n_owned;
    // ---- vertex -> incident cells (only needed for owned vertices)
    std::vector<int32_t> vc_ptr(NO + 1, 0);
    for (int64_t c = 0; c < n_cells; ++c)
        for (int a = 0; a < 3; ++a) {
            int32_t v = h_cells[3 * c + a];
            if (v < NO) vc_ptr[v + 1]++;
        }
    for (int64_t v = 0; v < NO; ++v) vc_ptr[v + 1] += vc_ptr[v];
    std::vector<int32_t> vc(vc_ptr[NO]);
    {
        std::vector<int32_t> fill(vc_ptr.begin(), vc_ptr.end() - 1);
        for (int64_t c = 0; c < n_cells; ++c)
            for (int a = 0; a < 3; ++a) {
                int32_t v = h_cells[3 * c + a];
                if (v < NO) vc[fill[v]++] = int32_t(c);
            }
    }
    // ---- node-block CSR pattern of the owned rows: sorted unique {v} U neighbours
    std::vector<int32_t> &rowptr = ctx->h_rowptr, &colidx = ctx->h_colidx;
    rowptr.assign(NO + 1, 0);
    colidx.clear();
    colidx.reserve(size_t(NO) * 7);
    {
        std::vector<int32_t> tmp;
        for (int64_t v = 0; v < NO; ++v) {
            tmp.clear();
            tmp.push_back(int32_t(v));
            for (int32_t q = vc_ptr[v]; q < vc_ptr[v + 1]; ++q) {
                const int32_t *cv = h_cells + 3 * int64_t(vc[q]);
                for (int a = 0; a < 3; ++a)
                    if (cv[a] != v) tmp.push_back(cv[a]);
            }
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            FEDMX_REQUIRE(ctx, tmp.size() <= 255, "fedmx_mesh_set: vertex degree > 254 unsupported");
            colidx.insert(colidx.end(), tmp.begin(), tmp.end());
            rowptr[v + 1] = int32_t(colidx.size());
        }
    }
    // ---- exterior facet markers keyed by (min,max) vertex pair
    std::vector<std::pair<int64_t, int32_t>> fkeys;
    fkeys.reserve(n_facets);
    for (int64_t f = 0; f < n_facets; ++f) {
        int64_t a = h_facet_vertices[2 * f], b = h_facet_vertices[2 * f + 1];
        int32_t mk = h_facet_marker ? h_facet_marker[f] : 0;
        FEDMX_REQUIRE(ctx, mk >= 0 && mk <= 15, "fedmx_mesh_set: facet marker must be in 0..15");
        if (a > b) std::swap(a, b);
        fkeys.emplace_back(a * NV + b, mk);
    }
    std::sort(fkeys.begin(), fkeys.end());
    auto facet_marker = [&](int64_t a, int64_t b) -> uint32_t {
        if (a > b) std::swap(a, b);
        auto it = std::lower_bound(fkeys.begin(), fkeys.end(), std::make_pair(a * NV + b, int32_t(-1)));
        return (it != fkeys.end() && it->first == a * NV + b) ? uint32_t(it->second) : 0u;
    };

    // ---- storage order of the rows: spatially compact tiles of FEDMX_TILE nodes, ascending node index inside a tile
    const int64_t n_slices = (NO + FEDMX_SLICE - 1) / FEDMX_SLICE;
    std::vector<int32_t> &row_node = ctx->h_row_node, &node_row = ctx->h_node_row;
    row_node.assign(n_slices * FEDMX_SLICE, -1);
    node_row.assign(NO, -1);
    {
        std::vector<int32_t> order(NO);
        std::iota(order.begin(), order.end(), 0);
        static const bool tiled_rows = !(getenv("FEDMX_ROW_ORDER") && std::string(getenv("FEDMX_ROW_ORDER")) == "natural");
        if (tiled_rows && NO > FEDMX_TILE) {
            const std::vector<double> ux = dense_rank(h_coords, 0, NO), uy = dense_rank(h_coords, 1, NO);
            Rcb rcb{ux, uy, order};
            rcb.split(0, NO);
            for (int64_t t = 0; t * FEDMX_TILE < NO; ++t)
                std::sort(order.begin() + t * FEDMX_TILE, order.begin() + std::min<int64_t>(NO, (t + 1) * FEDMX_TILE));
        }
        for (int64_t r = 0; r < NO; ++r) { row_node[r] = order[r]; node_row[order[r]] = int32_t(r); }
    }
    auto rlen = [&](int64_t r) { const int32_t v = row_node[r]; return rowptr[v + 1] - rowptr[v]; };

    // ---- SELL-32 slicing of matrix slots
    std::vector<int32_t> slice_ptr(n_slices + 1, 0), slice_width(n_slices, 0);
    for (int64_t s = 0; s < n_slices; ++s) {
        int w = 0;
        for (int64_t r = s * FEDMX_SLICE; r < std::min<int64_t>(NO, (s + 1) * FEDMX_SLICE); ++r)
            w = std::max(w, rlen(r));
        slice_width[s] = w;
        int64_t next = int64_t(slice_ptr[s]) + int64_t(w) * FEDMX_SLICE;
        FEDMX_REQUIRE(ctx, next * n_eq * n_eq < (int64_t(1) << 31) * 4, "fedmx_mesh_set: matrix too large for this rank");
        FEDMX_REQUIRE(ctx, next < (int64_t(1) << 31), "fedmx_mesh_set: slot index overflow");
        slice_ptr[s + 1] = int32_t(next);
    }
    const int64_t n_slots = slice_ptr[n_slices];
    std::vector<int32_t> col(n_slots);
    for (int64_t s = 0; s < n_slices; ++s)
        for (int l = 0; l < FEDMX_SLICE; ++l) {
            int64_t r = s * FEDMX_SLICE + l;
            for (int k = 0; k < slice_width[s]; ++k) {
                int64_t slot = int64_t(slice_ptr[s]) + int64_t(k) * FEDMX_SLICE + l;
                int32_t c;
                if (r < NO && k < rlen(r)) c = colidx[rowptr[row_node[r]] + k];
                else c = int32_t(r < NO ? row_node[r] : 0);
                col[slot] = c;
            }
        }

    // ---- fans
    std::vector<std::vector<Inc>> fans(NO);
    std::vector<uint32_t> row_pack(n_slices * FEDMX_SLICE, 0);
    {
        struct Loc { int32_t cell, u, w; bool used; };
        std::vector<Loc> loc;
        for (int64_t v = 0; v < NO; ++v) {
            loc.clear();
            for (int32_t q = vc_ptr[v]; q < vc_ptr[v + 1]; ++q) {
                const int32_t *cv = h_cells + 3 * int64_t(vc[q]);
                int a = (cv[0] == v) ? 0 : (cv[1] == v ? 1 : 2);
                loc.push_back({vc[q], cv[(a + 1) % 3], cv[(a + 2) % 3], false});
            }
            const int nloc = int(loc.size());
            FEDMX_REQUIRE(ctx, nloc > 0 && nloc <= 255, "fedmx_mesh_set: vertex without cells / too many cells");
            auto count_spoke = [&](int32_t sp) {
                int n = 0;
                for (auto &L : loc) n += (L.u == sp) + (L.w == sp);
                return n;
            };
            auto next_cell = [&](int32_t sp, int self) {
                for (int i = 0; i < nloc; ++i)
                    if (i != self && !loc[i].used && (loc[i].u == sp || loc[i].w == sp)) return i;
                return -1;
            };
            const int32_t lo = rowptr[v], hi = rowptr[v + 1];
            std::vector<Inc> &fan = fans[v];
            int remaining = nloc;
            while (remaining > 0) {
                // pick a start: prefer a cell with a boundary spoke (spoke used by exactly one cell)
                int start = -1;
                int32_t p = -1;
                for (int i = 0; i < nloc && start < 0; ++i) {
                    if (loc[i].used) continue;
                    if (count_spoke(loc[i].u) == 1) { start = i; p = loc[i].u; }
                    else if (count_spoke(loc[i].w) == 1) { start = i; p = loc[i].w; }
                }
                bool closed = false;
                if (start < 0) {
                    for (int i = 0; i < nloc; ++i)
                        if (!loc[i].used) { start = i; p = loc[i].u; break; }
                    closed = true;
                }
                const size_t first_idx = fan.size();
                const int32_t first_p = p;
                int cur = start;
                while (cur >= 0) {
                    loc[cur].used = true;
                    --remaining;
                    int32_t n = (loc[cur].u == p) ? loc[cur].w : loc[cur].u;
                    int sp = find_slot(colidx, lo, hi, p), sn = find_slot(colidx, lo, hi, n);
                    FEDMX_REQUIRE(ctx, sp >= 0 && sn >= 0, "fedmx_mesh_set: internal pattern error");
                    uint32_t pack = uint32_t(sp) | (uint32_t(sn) << 8) | (facet_marker(v, p) << 16) |
                                    (facet_marker(v, n) << 20);
                    fan.push_back({p, n, pack, loc[cur].cell});
                    int nxt = next_cell(n, cur);
                    if (nxt < 0) {
                        // closed fans end where they started; otherwise this is an open end
                        if (closed && n != first_p) closed = false;
                        break;
                    }
                    p = n;
                    cur = nxt;
                }
                fan[first_idx].pack |= (1u << 24);
                fan.back().pack |= (1u << 25);
                if (closed) { fan[first_idx].pack |= (1u << 26); fan.back().pack |= (1u << 26); }
            }
            int diag = find_slot(colidx, lo, hi, int32_t(v));
            uint32_t marked = 0;
            for (const Inc &in : fan) marked |= (in.pack >> 16) & 0xffu;
            row_pack[node_row[v]] = uint32_t(nloc) | (uint32_t(diag) << 8) | (marked ? (1u << 31) : 0u);
        }
    }
    std::vector<int32_t> inc_ptr(n_slices + 1, 0), inc_width(n_slices, 0);
    for (int64_t s = 0; s < n_slices; ++s) {
        int w = 0;
        for (int64_t r = s * FEDMX_SLICE; r < std::min<int64_t>(NO, (s + 1) * FEDMX_SLICE); ++r)
            w = std::max<int>(w, int(fans[row_node[r]].size()));
        inc_width[s] = w;
        inc_ptr[s + 1] = inc_ptr[s] + w * FEDMX_SLICE;
    }
    const int64_t n_inc = inc_ptr[n_slices];
    std::vector<int32_t> inc_vp(n_inc, 0), inc_vn(n_inc, 0);
    std::vector<uint32_t> inc_pack(n_inc, 0), inc_cell(n_inc, 0), inc_cid(n_inc, 0);
    const bool cid_ok = n_cells < (int64_t(1) << 28);
    for (int64_t r = 0; r < NO; ++r) {
        int64_t s = r / FEDMX_SLICE, l = r % FEDMX_SLICE;
        const int32_t v = row_node[r];
        const std::vector<Inc> &fan = fans[v];
        for (size_t m = 0; m < fan.size(); ++m) {
            int64_t e = int64_t(inc_ptr[s]) + int64_t(m) * FEDMX_SLICE + l;
            inc_vp[e] = fan[m].vp;
            inc_vn[e] = fan[m].vn;
            inc_pack[e] = fan[m].pack;
            const int32_t *cv = h_cells + 3 * int64_t(fan[m].cell);
            const uint32_t ka = (cv[0] == v) ? 0 : (cv[1] == v ? 1 : 2);
            const uint32_t kp = (cv[0] == fan[m].vp) ? 0 : (cv[1] == fan[m].vp ? 1 : 2);
            inc_cid[e] = (uint32_t(fan[m].cell) << 4) | (ka << 2) | kp;
        }
    }

    // ---- cell batches of the tiled assembly kernel: per tile the cells incident to its rows (numbered in the order the
    // row threads first meet them: fan position major, lane minor, so that the lanes of a warp touch consecutive records),
    // the ring of vertices outside the tile, and per fan entry (local cell, position of the row's node, position of vp)
    const int64_t n_tiles = (NO + FEDMX_TILE - 1) / FEDMX_TILE;
    std::vector<TileInfo> tinfo(n_tiles);
    std::vector<std::vector<int32_t>> thalo_t(n_tiles);
    std::vector<std::vector<unsigned long long>> tcellv_t(n_tiles);
    int max_cells = 0, max_vtx = 0;
    bool tiles_ok = true;
    {
        std::vector<int32_t> cell_loc(n_cells, -1), vtx_loc(NV, -1), touched_c, touched_v;
        for (int64_t t = 0; t < n_tiles; ++t) {
            const int64_t r0 = t * FEDMX_TILE, r1 = std::min<int64_t>(NO, r0 + FEDMX_TILE);
            touched_c.clear();
            touched_v.clear();
            for (int64_t r = r0; r < r1; ++r) { vtx_loc[row_node[r]] = int32_t(r - r0); touched_v.push_back(row_node[r]); }
            int nvtx = FEDMX_TILE;                       // local ids 0..127 are the tile's rows (also for a short last tile)
            TileInfo &ti = tinfo[t];
            std::vector<int32_t> &thalo = thalo_t[t];
            std::vector<unsigned long long> &tcellv = tcellv_t[t];
            size_t maxfan = 0;
            for (int64_t r = r0; r < r1; ++r) maxfan = std::max(maxfan, fans[row_node[r]].size());
            for (size_t m = 0; m < maxfan; ++m)
                for (int64_t r = r0; r < r1; ++r) {
                    const int32_t v = row_node[r];
                    const std::vector<Inc> &fan = fans[v];
                    if (m >= fan.size()) continue;
                    const int32_t c = fan[m].cell;
                    const int32_t *cv = h_cells + 3 * int64_t(c);
                    if (cell_loc[c] < 0) {
                        cell_loc[c] = int32_t(tcellv.size());
                        touched_c.push_back(c);
                        unsigned long long w = 0;
                        for (int k = 0; k < 3; ++k) {
                            if (vtx_loc[cv[k]] < 0) {
                                vtx_loc[cv[k]] = nvtx++;
                                touched_v.push_back(cv[k]);
                                thalo.push_back(cv[k]);
                            }
                            w |= (unsigned long long)(uint32_t(vtx_loc[cv[k]]) & 0xffffu) << (16 * k);
                        }
                        tcellv.push_back(w);
                    }
                    const int ka = (cv[0] == v) ? 0 : (cv[1] == v ? 1 : 2);
                    const int kp = (cv[0] == fan[m].vp) ? 0 : (cv[1] == fan[m].vp ? 1 : 2);
                    const int64_t s = r / FEDMX_SLICE, l = r % FEDMX_SLICE;
                    inc_cell[int64_t(inc_ptr[s]) + int64_t(m) * FEDMX_SLICE + l] =
                        uint32_t(cell_loc[c]) | (uint32_t(ka) << 12) | (uint32_t(kp) << 14);
                }
            ti.n_cells = int32_t(tcellv.size());
            ti.n_halo = int32_t(thalo.size());
            max_cells = std::max(max_cells, int(ti.n_cells));
            max_vtx = std::max(max_vtx, nvtx);
            if (ti.n_cells > 4095 || nvtx > 65535) tiles_ok = false;
            for (int32_t c : touched_c) cell_loc[c] = -1;
            for (int32_t v : touched_v) vtx_loc[v] = -1;
        }
    }
    fans.clear();
    fans.shrink_to_fit();

    // ---- upload
    ctx->n_eq = n_eq;
    ctx->n_vertices = NV;
    ctx->n_owned = NO;
    ctx->h_slice_ptr = slice_ptr;
    int rc;
    {
        std::vector<double> xy(h_coords, h_coords + 2 * NV);
        if ((rc = fedmx_upload(ctx, &ctx->d_coords, xy))) return rc;
    }
    SellView &S = ctx->sell;
    S.n_rows = NO;
    S.n_slices = n_slices;
    S.n_slots = n_slots;
    int32_t *p32;
    uint32_t *pu32;
    if ((rc = fedmx_upload(ctx, &p32, slice_ptr))) return rc; S.slice_ptr = p32;
    if ((rc = fedmx_upload(ctx, &p32, slice_width))) return rc; S.slice_width = p32;
    if ((rc = fedmx_upload(ctx, &p32, col))) return rc; S.col = p32;
    if ((rc = fedmx_upload(ctx, &p32, inc_ptr))) return rc; S.inc_ptr = p32;
    if ((rc = fedmx_upload(ctx, &p32, inc_width))) return rc; S.inc_width = p32;
    if ((rc = fedmx_upload(ctx, &p32, inc_vp))) return rc; S.inc_vp = p32;
    if ((rc = fedmx_upload(ctx, &p32, inc_vn))) return rc; S.inc_vn = p32;
    if ((rc = fedmx_upload(ctx, &pu32, inc_pack))) return rc; S.inc_pack = pu32;
    S.inc_cid = nullptr;
    if (cid_ok) {                             // cell-wise path of the plug-in models (row_kernel.cuh)
        if ((rc = fedmx_upload(ctx, &pu32, inc_cid))) return rc; S.inc_cid = pu32;
        std::vector<int32_t> cv(h_cells, h_cells + 3 * int64_t(n_cells));
        if ((rc = fedmx_upload(ctx, &ctx->d_cells, cv))) return rc;
        ctx->n_cells = n_cells;
    }
    if ((rc = fedmx_upload(ctx, &pu32, row_pack))) return rc; S.row_pack = pu32;
    if ((rc = fedmx_upload(ctx, &p32, row_node))) return rc; S.row_node = p32;
    if (tiles_ok) {
        TileView &T = ctx->tiles;
        TileInfo *pti;
        unsigned long long *pcv;
        int max_halo = 0;
        for (const TileInfo &ti : tinfo) max_halo = std::max(max_halo, int(ti.n_halo));
        T.cell_stride = (max_cells + 15) & ~15;
        T.halo_stride = (max_halo + 15) & ~15;
        std::vector<int32_t> thalo(size_t(n_tiles) * T.halo_stride, -1);
        std::vector<unsigned long long> tcellv(size_t(n_tiles) * T.cell_stride, 0ull);
        for (int64_t t = 0; t < n_tiles; ++t) {
            std::copy(thalo_t[t].begin(), thalo_t[t].end(), thalo.begin() + t * T.halo_stride);
            std::copy(tcellv_t[t].begin(), tcellv_t[t].end(), tcellv.begin() + t * T.cell_stride);
        }
        if ((rc = fedmx_upload(ctx, &pti, tinfo))) return rc; T.info = pti;
        if ((rc = fedmx_upload(ctx, &p32, thalo))) return rc; T.halo = p32;
        if ((rc = fedmx_upload(ctx, &pcv, tcellv))) return rc; T.cellv = pcv;
        if ((rc = fedmx_upload(ctx, &pu32, inc_cell))) return rc; T.inc_cell = pu32;
        T.n_tiles = n_tiles;
        T.max_cells = max_cells;
        T.max_vtx = max_vtx;
    }
    const int NB = n_eq * n_eq;
    if ((rc = fedmx_dev_alloc(ctx, &ctx->d_val, size_t(n_slots) * NB))) return rc;
    if ((rc = fedmx_dev_alloc(ctx, &ctx->d_dinv, size_t(n_slices) * FEDMX_SLICE * NB))) return rc;
    FEDMX_CHECK_CUDA(ctx, cudaMemset(ctx->d_val, 0, size_t(n_slots) * NB * sizeof(double)));
    FEDMX_CHECK_CUDA(ctx, cudaMemset(ctx->d_dinv, 0, size_t(n_slices) * FEDMX_SLICE * NB * sizeof(double)));
    return fedmx_ensure_workspace(ctx, 0);
}

extern "C" int fedmx_pattern_get(fedmx_ctx *ctx, int64_t *n_rows, int64_t *nnz_blocks, int32_t *h_rowptr,
                                 int32_t *h_colidx) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, ctx->n_vertices > 0, "fedmx_pattern_get: no mesh");
    if (n_rows) *n_rows = ctx->n_owned;
    if (nnz_blocks) *nnz_blocks = int64_t(ctx->h_colidx.size());
    if (h_rowptr) std::memcpy(h_rowptr, ctx->h_rowptr.data(), ctx->h_rowptr.size() * sizeof(int32_t));
    if (h_colidx) std::memcpy(h_colidx, ctx->h_colidx.data(), ctx->h_colidx.size() * sizeof(int32_t));
    return FEDMX_OK;
}

extern "C" int fedmx_matrix_get_bsr(fedmx_ctx *ctx, double *h_values) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, ctx->d_val && h_values, "fedmx_matrix_get_bsr: no matrix");
    const int NB = ctx->n_eq * ctx->n_eq;
    std::vector<double> v(size_t(ctx->sell.n_slots) * NB);
    FEDMX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    FEDMX_CHECK_CUDA(ctx, cudaMemcpy(v.data(), ctx->d_val, v.size() * sizeof(double), cudaMemcpyDeviceToHost));
    const auto &rp = ctx->h_rowptr;
    for (int64_t r = 0; r < ctx->n_owned; ++r) {                       // r: node (the exported pattern is in node order)
        const int64_t sr = ctx->h_node_row[r];
        int64_t s = sr / FEDMX_SLICE, l = sr % FEDMX_SLICE;
        for (int32_t q = rp[r]; q < rp[r + 1]; ++q) {
            int64_t k = q - rp[r];
            const double *src = v.data() + (int64_t(ctx->h_slice_ptr[s]) + k * FEDMX_SLICE) * NB + l;
            for (int e = 0; e < NB; ++e) h_values[int64_t(q) * NB + e] = src[e * FEDMX_SLICE];
        }
    }
    return FEDMX_OK;
}

extern "C" int fedmx_dirichlet_set(fedmx_ctx *ctx, int comp, int64_t n_nodes, const int32_t *h_nodes,
                                   const int32_t *h_value_index) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, ctx->n_vertices > 0, "fedmx_dirichlet_set: set the mesh first");
    FEDMX_REQUIRE(ctx, comp >= 0 && comp < ctx->n_eq, "fedmx_dirichlet_set: bad component");
    FEDMX_REQUIRE(ctx, ctx->bc_comp < 0 || ctx->bc_comp == comp, "fedmx_dirichlet_set: one constrained component supported");
    std::vector<uint32_t> row_pack(size_t(ctx->sell.n_slices) * FEDMX_SLICE);
    FEDMX_CHECK_CUDA(ctx, cudaMemcpy(row_pack.data(), ctx->sell.row_pack, row_pack.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n_nodes; ++i) {
        int32_t v = h_nodes[i];
        int32_t vi = h_value_index[i];
        FEDMX_REQUIRE(ctx, v >= 0 && v < ctx->n_vertices, "fedmx_dirichlet_set: node out of range");
        FEDMX_REQUIRE(ctx, vi >= 0 && vi < FEDMX_MAX_BC - 1, "fedmx_dirichlet_set: value index out of range");
        if (v < ctx->n_owned) {
            const int32_t sr = ctx->h_node_row[v];
            row_pack[sr] = (row_pack[sr] & 0x8000ffffu) | (uint32_t(vi + 1) << 16);
        }
    }
    FEDMX_CHECK_CUDA(ctx, cudaMemcpy((void *)ctx->sell.row_pack, row_pack.data(), row_pack.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    ctx->bc_comp = comp;
    ctx->mp.bc_comp = comp;
    return FEDMX_OK;
}
