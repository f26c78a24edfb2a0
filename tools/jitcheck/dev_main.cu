#include FORM
#include "harness.inc"
__global__ void run_cells(const Case *cs, ModelParams mp, double *sc, int n) {
    int t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= n) return;
    if constexpr (JitModel::CELLS) JitModel::cell<true>(mp, cs[t].v[0], cs[t].v[1], cs[t].v[2], CELL_REC(sc, t));
}
__global__ void run(const Case *cs, ModelParams mp, double *out, int n) {
    constexpr int NEQ = JitModel::NEQ;
    int t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= n) return;
    RowOut<NEQ> o, o0;
    JitModel::row<true>(mp, cs[t].v[0], cs[t].v[1], cs[t].v[2], cs[t].m1, cs[t].m2, o);
    JitModel::row<false>(mp, cs[t].v[0], cs[t].v[1], cs[t].v[2], cs[t].m1, cs[t].m2, o0);
    double *w = out + size_t(t) * kPer;
    for (int k = 0; k < NEQ; ++k) { w[k] = o.R[k]; w[NEQ + k] = o0.R[k]; }
    for (int b = 0; b < 3; ++b) for (int e = 0; e < NEQ * NEQ; ++e) w[2 * NEQ + b * NEQ * NEQ + e] = o.J[b][e];
}
int main(int argc, char **argv) {
    std::vector<Case> cs; ModelParams mp; memset(&mp, 0, sizeof(mp)); make_cases(cs, mp, 256);
    Case *d; double *o; cudaMalloc(&d, cs.size() * sizeof(Case)); cudaMalloc(&o, cs.size() * kPer * 8);
    cudaMemcpy(d, cs.data(), cs.size() * sizeof(Case), cudaMemcpyHostToDevice);
    run<<<4, 64>>>(d, mp, o, 256);
    std::vector<double> h(cs.size() * kPer);
    cudaError_t e = cudaMemcpy(h.data(), o, h.size() * 8, cudaMemcpyDeviceToHost);
    printf("cuda: %s\n", cudaGetErrorString(e));
    FILE *f = fopen(argv[1], "wb"); fwrite(h.data(), 8, h.size(), f);
    if (JitModel::CELLS) {
        std::vector<double> sc((cs.size() + 31) / 32 * kCellBlock, 0.0);
        double *dsc; cudaMalloc(&dsc, sc.size() * 8); cudaMemset(dsc, 0, sc.size() * 8);
        run_cells<<<4, 64>>>(d, mp, dsc, 256);
        e = cudaMemcpy(sc.data(), dsc, sc.size() * 8, cudaMemcpyDeviceToHost);
        printf("cuda (cells): %s\n", cudaGetErrorString(e));
        fwrite(sc.data(), 8, sc.size(), f);
    }
    fclose(f);
}
