#include FORM
#include "harness.inc"
int main(int argc, char **argv) {
    constexpr int NEQ = JitModel::NEQ;
    std::vector<Case> cs; ModelParams mp; make_cases(cs, mp, 256);
    FILE *f = fopen(argv[1], "wb");
    for (auto &c : cs) {
        RowOut<NEQ> o, o0; JitModel::row<true>(mp, c.v[0], c.v[1], c.v[2], c.m1, c.m2, o); JitModel::row<false>(mp, c.v[0], c.v[1], c.v[2], c.m1, c.m2, o0);
        fwrite(o.R, 8, NEQ, f); fwrite(o0.R, 8, NEQ, f); fwrite(o.J, 8, 3 * NEQ * NEQ, f);
    }
    if (JitModel::CELLS) {
        std::vector<double> sc((cs.size() + 31) / 32 * kCellBlock, 0.0);
        for (size_t t = 0; t < cs.size(); ++t) JitModel::cell<true>(mp, cs[t].v[0], cs[t].v[1], cs[t].v[2], CELL_REC(sc.data(), t));
        fwrite(sc.data(), 8, sc.size(), f);
    }
    fclose(f);
}
