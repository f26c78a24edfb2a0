// Included by row_kernel.cuh (inside its anonymous namespace) instead of the <<< >>> launcher when FEDMX_HOST_EMULATION is
// defined: the same dispatch as launch_rows, every kernel run sequentially over its grid.
#pragma once
template <class K>
inline void emul_launch(K kernel, unsigned blocks, unsigned threads, const AsmArgs &A) {
    blockDim.x = threads;
    gridDim.x = blocks;
    for (unsigned b = 0; b < blocks; ++b)
        for (unsigned t = 0; t < threads; ++t) {
            blockIdx.x = b;
            threadIdx.x = t;
            kernel(A);
        }
}

template <class Model>
inline void launch_rows(const AsmArgs &A, bool jac, int minb, cudaStream_t) {
    const int threads = 128;
    if constexpr (ModelHasCells<Model>::value) {
        if (A.scratch && A.cells && A.S.inc_cid && (!jac || Model::CELLS_JAC)) {
            const unsigned cblocks = unsigned((A.n_cells + threads - 1) / threads);
            const unsigned rblocks = unsigned((A.S.n_rows + threads - 1) / threads);
            if (jac) {
                emul_launch(assemble_cells_kernel<Model, true, Model::CELL_MINB>, cblocks, threads, A);
                emul_launch(gather_rows_kernel<Model, true, 4>, rblocks, threads, A);
            } else {
                emul_launch(assemble_cells_kernel<Model, false, 4>, cblocks, threads, A);
                emul_launch(gather_rows_kernel<Model, false, 4>, rblocks, threads, A);
            }
            return;
        }
    }
    const int64_t rows = A.S.n_slices * FEDMX_SLICE;
    const unsigned blocks = unsigned((rows + threads - 1) / threads);
    if (jac) {
        if constexpr (2 * Model::NEQ * Model::NEQ * 128 * 8 <= 48 * 1024) {
            if (minb == 13) { emul_launch(assemble_rows_kernel<Model, true, 3, true>, blocks, threads, A); return; }
        }
        emul_launch(assemble_rows_kernel<Model, true, 2>, blocks, threads, A);
    } else {
        emul_launch(assemble_rows_kernel<Model, false, 2>, blocks, threads, A);
    }
}
