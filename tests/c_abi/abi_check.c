/* Plain-C client of libsasvqa_b200.so: proves the drop-in boundary needs nothing but include/sasvqa.h, pointers and
 * sizes (no torch, no C++).  Built and run by tests/test_host_cpu.py (no GPU: version, sizes, argument errors) and by
 * tests/test_gpu_parity.py with --gpu (the selection entry points against values worked out by hand / on the host).
 *   gcc -std=c99 -I include tests/c_abi/abi_check.c -L sas-vqa_b200 -lsasvqa_b200 -L/usr/local/cuda/lib64 -lcudart */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sasvqa.h"

/* the three CUDA runtime calls the GPU half needs, declared by hand so this file stays plain C99 */
extern int cudaMalloc(void** p, size_t n);
extern int cudaMemcpy(void* dst, const void* src, size_t n, int kind);
extern int cudaFree(void* p);
extern int cudaDeviceSynchronize(void);
enum { H2D = 1, D2H = 2 };

#define CHECK(cond)                                                            \
    do {                                                                       \
        if (!(cond)) {                                                         \
            fprintf(stderr, "FAILED %s:%d: %s (last error: %s)\n", __FILE__, __LINE__, #cond, sasvqa_last_error()); \
            return 1;                                                          \
        }                                                                      \
    } while (0)

static int cpu_checks(void) {
    CHECK(sasvqa_abi_version() == 1);
    CHECK(sasvqa_scorer_num_params(28996, 2) == 108311810ull);          /* bert-base-cased, 2 labels */
    CHECK(sasvqa_scorer_num_params(0, 2) == 0);
    /* argument errors are reported before any CUDA call: codes, not exceptions, and a message */
    CHECK(sasvqa_topk_strided(NULL, 1, 8, 0, 2, NULL, NULL) == SASVQA_ERR_INVALID);
    CHECK(strlen(sasvqa_last_error()) > 0);
    CHECK(sasvqa_topk_strided((const float*)16, 1, 4, 1, 5, (int32_t*)16, NULL) == SASVQA_ERR_INVALID);   /* K > candidates */
    CHECK(sasvqa_mdf_scores(NULL, 2, 8, 2, NULL, NULL, NULL) == SASVQA_ERR_INVALID);
    CHECK(sasvqa_mdf_scores(NULL, 0, 0, 2, NULL, NULL, NULL) == SASVQA_OK);                               /* empty batch */
    SasvqaEncoder* enc = NULL;
    float dummy = 0.f;
    CHECK(sasvqa_encoder_create(&dummy, 1, 0, &enc) == SASVQA_ERR_INVALID && enc == NULL);                /* wrong size */
    SasvqaScorer* sc = NULL;
    CHECK(sasvqa_scorer_create(&dummy, 1, 100, 2, 0, &sc) == SASVQA_ERR_INVALID && sc == NULL);
    sasvqa_encoder_destroy(NULL);
    sasvqa_scorer_destroy(NULL);
    return 0;
}

static int gpu_checks(void) {
    /* MIF strided top-K (gen_sample.py:87-88) on 2 score rows of 10 */
    const float scores[20] = {0.1f, 0.9f, 0.3f, 0.8f, -1.f, 0.7f, 0.95f, 0.2f, 0.0f, 0.5f,
                              5.f, 4.f, 3.f, 2.f, 1.f, 0.f, -1.f, -2.f, -3.f, 9.f};
    float* d_scores = NULL;
    int32_t* d_idx = NULL;
    int32_t idx[6];
    CHECK(cudaMalloc((void**)&d_scores, sizeof scores) == 0 && cudaMalloc((void**)&d_idx, sizeof idx) == 0);
    CHECK(cudaMemcpy(d_scores, scores, sizeof scores, H2D) == 0);
    CHECK(sasvqa_topk_strided(d_scores, 2, 10, 1, 3, d_idx, NULL) == SASVQA_OK);
    CHECK(cudaMemcpy(idx, d_idx, sizeof idx, D2H) == 0);
    CHECK(idx[0] == 6 && idx[1] == 1 && idx[2] == 3 && idx[3] == 9 && idx[4] == 0 && idx[5] == 1);
    CHECK(sasvqa_topk_strided(d_scores, 2, 10, 2, 3, d_idx, NULL) == SASVQA_OK);      /* every 2nd score, indices scaled back */
    CHECK(cudaMemcpy(idx, d_idx, sizeof idx, D2H) == 0);
    CHECK(idx[0] == 6 && idx[1] == 2 && idx[2] == 0 && idx[3] == 0 && idx[4] == 2 && idx[5] == 4);

    /* MDF greedy selection (utils.py:63-88): T = 12, W = 2, K = 3; peaks at 5 (0.9), 9 (0.8), 2 (0.7) */
    const float lcl[12] = {0.f, 0.f, 0.7f, 0.1f, 0.2f, 0.9f, 0.85f, 0.1f, 0.3f, 0.8f, 0.f, 0.f};
    float* d_lcl = NULL;
    int32_t *d_sel = NULL, *d_status = NULL, sel[3], status = -1;
    CHECK(cudaMalloc((void**)&d_lcl, sizeof lcl) == 0 && cudaMalloc((void**)&d_sel, sizeof sel) == 0 &&
          cudaMalloc((void**)&d_status, sizeof status) == 0);
    CHECK(cudaMemcpy(d_lcl, lcl, sizeof lcl, H2D) == 0);
    CHECK(sasvqa_mdf_select(d_lcl, 1, 12, 3, 2, d_sel, d_status, NULL) == SASVQA_OK);
    CHECK(cudaMemcpy(sel, d_sel, sizeof sel, D2H) == 0 && cudaMemcpy(&status, d_status, sizeof status, D2H) == 0);
    /* top = 5; open intervals [0,3) -> max 0.7 @2 and [7,12) -> max 0.8 @9 (6 is inside the +-W exclusion) */
    CHECK(status == SASVQA_STATUS_OK && sel[0] == 5 && sel[1] == 9 && sel[2] == 2);
    /* K = 5 cannot be met with spacing 2 on these intervals -> plain top-K fallback, 'Failure' */
    int32_t *d_sel5 = NULL, sel5[5];
    CHECK(cudaMalloc((void**)&d_sel5, sizeof sel5) == 0);
    CHECK(sasvqa_mdf_select(d_lcl, 1, 12, 5, 2, d_sel5, d_status, NULL) == SASVQA_OK);
    CHECK(cudaMemcpy(sel5, d_sel5, sizeof sel5, D2H) == 0 && cudaMemcpy(&status, d_status, sizeof status, D2H) == 0);
    /* greedy finds 5, 9, 2 and the 0.0 at 11, then runs dry: the reference discards them and takes topk(5) */
    CHECK(status == SASVQA_STATUS_FALLBACK);
    CHECK(sel5[0] == 5 && sel5[1] == 6 && sel5[2] == 9 && sel5[3] == 2 && sel5[4] == 8);
    CHECK(cudaDeviceSynchronize() == 0);
    CHECK(sasvqa_launch_count() >= 4);
    cudaFree(d_scores); cudaFree(d_idx); cudaFree(d_lcl); cudaFree(d_sel); cudaFree(d_status); cudaFree(d_sel5);
    return 0;
}

int main(int argc, char** argv) {
    if (cpu_checks()) return 1;
    if (argc > 1 && strcmp(argv[1], "--gpu") == 0 && gpu_checks()) return 1;
    printf("abi_check ok%s\n", argc > 1 ? " (gpu)" : "");
    return 0;
}
