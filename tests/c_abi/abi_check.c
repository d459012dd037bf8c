/* abi_check.c — a plain C11 consumer of include/pioran_b200.h, the way a foreign-language binding sees the library
 * (the Julia shim of julia/b200_solver.jl ccalls exactly these symbols).  Built by the tests with
 *     gcc -std=c11 -Wall -Werror -I include tests/c_abi/abi_check.c -ldl -o abi_check
 * (no CUDA headers, no C++).  Modes:
 *   abi_check layout LIB            -> dlopen, resolve every entry point it uses, print the struct layout as JSON
 *   abi_check call   LIB SERIES N THETA...  -> one pioran_approx_logl call on a series file (rows "t y sigma2"), prints logL
 * Test infrastructure only. */
#include <dlfcn.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pioran_b200.h"

typedef const char *(*last_error_fn)(void);
typedef int (*version_fn)(void);
typedef int (*ctx_create_fn)(int, pioran_ctx **);
typedef int (*ctx_create_multi_fn)(const int *, int, pioran_ctx **);
typedef int (*ctx_destroy_fn)(pioran_ctx *);
typedef int (*series_upload_fn)(pioran_ctx *, int64_t, const double *, const double *, const double *, int *);
typedef int (*series_free_fn)(pioran_ctx *, int);
typedef int (*approx_logl_fn)(pioran_ctx *, int, const int *, const pioran_approx_spec *, int, const double *, int, double *);

static void *must(void *lib, const char *name) {
    void *p = dlsym(lib, name);
    if (!p) { fprintf(stderr, "missing symbol %s\n", name); exit(3); }
    return p;
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: abi_check layout|call LIB ...\n"); return 2; }
    void *lib = dlopen(argv[2], RTLD_NOW | RTLD_LOCAL);
    if (!lib) { fprintf(stderr, "dlopen failed: %s\n", dlerror()); return 3; }
    last_error_fn last_error = (last_error_fn)must(lib, "pioran_last_error");
    version_fn version = (version_fn)must(lib, "pioran_version");
    ctx_create_fn ctx_create = (ctx_create_fn)must(lib, "pioran_ctx_create");
    ctx_create_multi_fn ctx_create_multi = (ctx_create_multi_fn)must(lib, "pioran_ctx_create_multi");
    ctx_destroy_fn ctx_destroy = (ctx_destroy_fn)must(lib, "pioran_ctx_destroy");
    series_upload_fn series_upload = (series_upload_fn)must(lib, "pioran_series_upload");
    series_free_fn series_free = (series_free_fn)must(lib, "pioran_series_free");
    approx_logl_fn approx_logl = (approx_logl_fn)must(lib, "pioran_approx_logl");
    (void)ctx_create_multi;

    if (!strcmp(argv[1], "layout")) {
        /* what every binding hard-codes (ctypes: pioran.jl_b200/_lib.py ApproxSpec; Julia: struct B200ApproxSpec) */
        printf("{\"version\": %d, \"sizeof_spec\": %zu, \"off_psd_model\": %zu, \"off_n_components\": %zu, \"off_basis\": %zu, "
               "\"off_is_integrated_power\": %zu, \"off_f_min\": %zu, \"off_f_max\": %zu, \"off_S_low\": %zu, \"off_S_high\": %zu, "
               "\"PIORAN_OK\": %d, \"PIORAN_EINVAL\": %d, \"PIORAN_ECUDA\": %d, \"PIORAN_ENOMEM\": %d, \"PIORAN_ESINGULAR\": %d, "
               "\"PIORAN_EUNSUPPORTED\": %d, \"sizeof_prior\": %zu, \"off_prior_kind\": %zu, \"off_prior_ref_col\": %zu, "
               "\"off_prior_p0\": %zu, \"off_prior_p1\": %zu}\n",
               version(), sizeof(pioran_approx_spec), offsetof(pioran_approx_spec, psd_model),
               offsetof(pioran_approx_spec, n_components), offsetof(pioran_approx_spec, basis),
               offsetof(pioran_approx_spec, is_integrated_power), offsetof(pioran_approx_spec, f_min),
               offsetof(pioran_approx_spec, f_max), offsetof(pioran_approx_spec, S_low), offsetof(pioran_approx_spec, S_high),
               PIORAN_OK, PIORAN_EINVAL, PIORAN_ECUDA, PIORAN_ENOMEM, PIORAN_ESINGULAR, PIORAN_EUNSUPPORTED,
               sizeof(pioran_prior_spec), offsetof(pioran_prior_spec, kind), offsetof(pioran_prior_spec, ref_col),
               offsetof(pioran_prior_spec, p0), offsetof(pioran_prior_spec, p1));
        /* argument checking needs no device */
        if (ctx_create(0, NULL) != PIORAN_EINVAL) { fprintf(stderr, "ctx_create(NULL out) must be EINVAL\n"); return 4; }
        if (!last_error()[0]) { fprintf(stderr, "no error message after a failed call\n"); return 4; }
        return 0;
    }

    if (!strcmp(argv[1], "call")) {
        /* call LIB SERIES J basis(0|1) theta0..theta5 : series rows "t y sigma2" (already transformed) */
        if (argc < 12) { fprintf(stderr, "usage: abi_check call LIB SERIES J BASIS th0 th1 th2 norm nu mu\n"); return 2; }
        FILE *f = fopen(argv[3], "r");
        if (!f) { perror("series"); return 2; }
        size_t cap = 1024, n = 0;
        double *t = malloc(cap * sizeof *t), *y = malloc(cap * sizeof *y), *s2 = malloc(cap * sizeof *s2);
        while (fscanf(f, "%lf %lf %lf", &t[n], &y[n], &s2[n]) == 3) {
            if (++n == cap) { cap *= 2; t = realloc(t, cap * sizeof *t); y = realloc(y, cap * sizeof *y); s2 = realloc(s2, cap * sizeof *s2); }
        }
        fclose(f);
        double dtmin = 1e300;
        for (size_t i = 1; i < n; i++) if (t[i] - t[i - 1] < dtmin) dtmin = t[i] - t[i - 1];
        pioran_approx_spec spec;
        memset(&spec, 0, sizeof spec);
        spec.psd_model = PIORAN_PSD_SBPL;
        spec.n_components = atoi(argv[4]);
        spec.basis = atoi(argv[5]);
        spec.is_integrated_power = 1;
        spec.f_min = 1.0 / (t[n - 1] - t[0]);           /* examples/ultranest/single_pl.jl:43-44 */
        spec.f_max = 1.0 / dtmin / 2.0;
        spec.S_low = 20.0; spec.S_high = 20.0;
        double theta[6];
        for (int k = 0; k < 6; k++) theta[k] = atof(argv[6 + k]);
        pioran_ctx *ctx = NULL;
        int sid = -1;
        double out = 0.0;
        int rc = ctx_create(0, &ctx);
        if (!rc) rc = series_upload(ctx, (int64_t)n, t, y, s2, &sid);
        if (!rc) rc = approx_logl(ctx, 1, &sid, &spec, 1, theta, 0, &out);
        if (rc) { fprintf(stderr, "error %d: %s\n", rc, last_error()); return 5; }
        printf("%.17g\n", out);
        series_free(ctx, sid);
        ctx_destroy(ctx);
        free(t); free(y); free(s2);
        return 0;
    }
    fprintf(stderr, "unknown mode %s\n", argv[1]);
    return 2;
}
