// scan_ldg.cu — instantiations of scan_topk_kernel for the direct-load variant
// (its own translation unit so that the build compiles the variants in parallel).
#include "scan_topk.cuh"

typedef void (*ScanFn)(const ScanParams);

ScanFn b200_pick_scan_ldg(int metric, int store, int qb, int lpr) {
#define SC(M, S, Q, L) \
    if (metric == M && store == S && qb == Q && lpr == L) return scan_topk_kernel<M, S, Q, 4, B200_VARIANT_LDG, L>;
#define SC_Q(M, S) SC(M, S, 1, 32) SC(M, S, 2, 32) SC(M, S, 4, 32) SC(M, S, 8, 32) SC(M, S, 1, 16) SC(M, S, 8, 16) SC(M, S, 1, 8) SC(M, S, 8, 8)
    SC_Q(0, 0) SC_Q(0, 1) SC_Q(1, 0) SC_Q(1, 1)
#undef SC_Q
#undef SC
    return nullptr;
}
