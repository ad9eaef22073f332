/* TEST INFRASTRUCTURE ONLY.  Stand-in for Torch3's EditDistance (insertions / deletions / substitutions between two word
 * index sequences with the given costs).  Only reached when a file of expected results is given to DecoderBatchTest
 * (src/DecoderBatchTest.cpp:148-200); the print format is Torch3's and is not reproduced (the oracle never passes one). */
#ifndef ORACLE_SHIM_HARNESS_EDITDISTANCE_H
#define ORACLE_SHIM_HARNESS_EDITDISTANCE_H
#include <vector>
#include "DiskXFile.h"
namespace Torch {
class EditDistance {
public:
    int ci, cd, cs, n_ins, n_del, n_sub, n_seq;
    EditDistance() : ci(1), cd(1), cs(1), n_ins(0), n_del(0), n_sub(0), n_seq(0) {}
    void setCosts(int i, int d, int s) { ci = i; cd = d; cs = s; }
    void distance(int* a, int na, int* b, int nb)
    {
        std::vector<std::vector<int> > c(na + 1, std::vector<int>(nb + 1, 0)), op(na + 1, std::vector<int>(nb + 1, 0));
        for (int i = 1; i <= na; ++i) { c[i][0] = i * ci; op[i][0] = 1; }
        for (int j = 1; j <= nb; ++j) { c[0][j] = j * cd; op[0][j] = 2; }
        for (int i = 1; i <= na; ++i)
            for (int j = 1; j <= nb; ++j) {
                const int s = c[i - 1][j - 1] + (a[i - 1] == b[j - 1] ? 0 : cs), in = c[i - 1][j] + ci, de = c[i][j - 1] + cd;
                c[i][j] = s; op[i][j] = a[i - 1] == b[j - 1] ? 0 : 3;
                if (in < c[i][j]) { c[i][j] = in; op[i][j] = 1; }
                if (de < c[i][j]) { c[i][j] = de; op[i][j] = 2; }
            }
        n_ins = n_del = n_sub = 0; n_seq = nb;
        for (int i = na, j = nb; i > 0 || j > 0;) {
            const int o = op[i][j];
            if (o == 1) { ++n_ins; --i; } else if (o == 2) { ++n_del; --j; } else { n_sub += o == 3; --i; --j; }
        }
    }
    void add(EditDistance* o) { n_ins += o->n_ins; n_del += o->n_del; n_sub += o->n_sub; n_seq += o->n_seq; }
    void print(XFile* f) { DiskXFile* d = dynamic_cast<DiskXFile*>(f); if (d) fprintf(d->file, "ins %d del %d sub %d of %d\n", n_ins, n_del, n_sub, n_seq); }
    void printRatio(XFile* f) { print(f); }
};
}
using Torch::EditDistance;
#endif
