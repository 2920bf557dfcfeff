// scan_bulk.cu — instantiations of scan_topk_kernel for the TMA-staged ring variant.  Compiled once per
// (metric, store) pair with -DSCAN_M=.. -DSCAN_S=.. (build.py) so that the build parallelises.
#include "scan_topk.cuh"

typedef void (*ScanFn)(const ScanParams);
#define SCAN_CAT2(a, b, c) a##b##_s##c
#define SCAN_CAT(a, b, c) SCAN_CAT2(a, b, c)

ScanFn SCAN_CAT(b200_pick_scan_bulk_m, SCAN_M, SCAN_S)(int qb, int lpr) {
#define SC(Q, L) \
    if (qb == Q && lpr == L) return scan_topk_kernel<SCAN_M, SCAN_S, Q, 4, B200_VARIANT_BULK, L>;
    SC(1, 32) SC(2, 32) SC(4, 32) SC(8, 32) SC(1, 16) SC(8, 16) SC(1, 8) SC(8, 8)
#undef SC
    return nullptr;
}
