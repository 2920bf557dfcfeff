/* memo_recall.c — the C ABI from plain C99: load a .memo file, run one query, print the top k.
 *
 *   gcc -std=c99 -O2 -I include examples/memo_recall.c -o memo_recall \
 *       c99_vectordb_b200/_b200flat.so -Wl,-rpath,'$ORIGIN/c99_vectordb_b200' -lm
 *   ./memo_recall db.memo 5
 *
 * What it replaces in the reference: load_index (memo_cli.py:251-261) + index.search (memo_cli.py:292).
 * The query here is the file's first row read back from the device, so the first hit is that row itself. */
#include <stdio.h>
#include <stdlib.h>

#include "b200_flat.h"

static int die(const char* what) {
    fprintf(stderr, "%s: %s\n", what, b200_last_error());
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s file.memo [k]\n", argv[0]);
        return 2;
    }
    const long k = argc > 2 ? strtol(argv[2], NULL, 10) : 5;
    if (k < 1) return 2;
    b200_memo_info info;
    if (b200_memo_probe(argv[1], &info)) return die("probe");
    printf("%s: %s index, d=%d, %s, %lld rows%s\n", argv[1], info.kind ? "id-mapped" : "flat", (int)info.d,
           info.metric == B200_METRIC_IP ? "inner product" : "L2", (long long)info.ntotal, info.from_hnsw ? " (HNSW graph skipped)" : "");
    b200_index* ix = NULL;
    if (b200_index_load(&ix, argv[1], B200_STORE_F32, 0, NULL)) return die("load");
    if (info.ntotal == 0) {
        b200_index_destroy(ix);
        return 0;
    }
    float* q = malloc(sizeof(float) * (size_t)info.d);
    float* D = malloc(sizeof(float) * (size_t)k);
    int64_t* I = malloc(sizeof(int64_t) * (size_t)k);
    if (!q || !D || !I) return 1;
    if (b200_index_get_rows(ix, 0, 1, q)) return die("get_rows");
    if (b200_index_search(ix, q, 1, k, D, I)) return die("search");
    for (long i = 0; i < k; ++i) printf("%2ld  id %lld  score %.6f\n", i + 1, (long long)I[i], D[i]);
    free(q);
    free(D);
    free(I);
    b200_index_destroy(ix);
    return 0;
}
