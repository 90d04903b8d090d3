/* c_client.c -- the C ABI of libfos_b200.so used from plain C (no Python, no torch).
 *
 * Generates a synthetic correlated-column design in HBM, estimates the Lipschitz constant
 * (estimate_lipschitz, iterative_solvers.py:45-60), runs FISTA-Lasso with history
 * (fista, iterative_solvers.py:132-245) and prints the objective trace.
 *
 *   gcc -std=c99 -I include examples/c_client.c -o c_client -L fastoptsolver_b200 -lfos_b200 \
 *       -Wl,-rpath,$PWD/fastoptsolver_b200 -lm
 *   ./c_client [rows] [cols] [iterations]
 *
 * Without a CUDA device every compute entry fails loudly (exit code 3, message from
 * fos_last_error): there is no CPU fallback.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "fos.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int st_ = (call);                                                        \
        if (st_ != FOS_OK) {                                                     \
            fprintf(stderr, "%s failed (%d): %s\n", #call, st_, fos_last_error()); \
            return 3;                                                            \
        }                                                                        \
    } while (0)

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 20000;
    const int64_t d = argc > 2 ? atoll(argv[2]) : 1024;
    const int iters = argc > 3 ? atoi(argv[3]) : 20;
    printf("abi %d, %d CUDA device(s)\n", fos_abi_version(), fos_device_count());

    fos_design* des = NULL;
    CHECK(fos_design_create_synthetic(n, d, FOS_F64, /*seed*/ 0, /*noise*/ 0.5, 0.5, 0.7, /*row0*/ 0, /*device*/ 0, &des));

    double lam = 0.0;
    CHECK(fos_design_lambda_max(des, &lam));
    const double alpha1 = 0.1 * lam;

    /* start vector of the power iteration: any unit vector (the Python drop-in draws it from numpy) */
    double* v0 = (double*)malloc((size_t)d * sizeof(double));
    double nrm = 0.0;
    for (int64_t i = 0; i < d; ++i) {
        v0[i] = sin(0.37 * (double)(i + 1));
        nrm += v0[i] * v0[i];
    }
    for (int64_t i = 0; i < d; ++i) v0[i] /= sqrt(nrm);
    double L = 0.0;
    int pit = 0;
    float ms = 0.f;
    CHECK(fos_power_iter(des, v0, 100, 1e-6, &L, &pit, &ms));
    printf("lambda_max %.17g  L %.17g after %d power steps (%.2f ms)\n", lam, L, pit, ms);

    fos_pg_params p = {0};
    p.scheme = FOS_SCHEME_NESTEROV;
    p.alpha1 = alpha1;
    p.obj_terms = 1; /* alpha1 * |x|_1 in the recorded objective */
    p.eta = 0.5;
    p.armijo_c = 1e-2;
    p.step0 = 1.0 / L;
    p.max_iter = iters;
    p.restart_threshold = 1.0;
    p.want_history = 1;

    fos_pg_result r = {0};
    r.x = (double*)malloc((size_t)d * sizeof(double));
    r.x_hist = (double*)malloc((size_t)(iters + 1) * d * sizeof(double));
    r.obj_hist = (double*)malloc((size_t)(iters > 0 ? iters : 1) * sizeof(double));
    CHECK(fos_prox_grad(des, &p, &r));

    int nnz = 0;
    for (int64_t i = 0; i < d; ++i) nnz += r.x[i] != 0.0;
    printf("%d iterations, %d passes over A, %lld kernel launches, loop %.3f ms, nnz %d\n", r.n_iters, r.n_passes,
           (long long)r.kernel_launches, r.loop_ms, nnz);
    for (int k = 0; k < r.n_iters; ++k) printf("obj[%d] %.17g\n", k + 1, r.obj_hist[k]);

    /* the recorded objective of the last iterate == a separate objective pass on it */
    double obj = 0.0;
    CHECK(fos_objective(des, r.x, /*lasso*/ 1, alpha1, 0.0, &obj));
    printf("objective(x) %.17g\n", obj);

    free(v0);
    free(r.x);
    free(r.x_hist);
    free(r.obj_hist);
    CHECK(fos_design_destroy(des));
    return 0;
}
