/* Plain-C consumer of libocto_b200.so — what a foreign-function binding (Julia ccall) does, without Python.
 * Builds the reference's 8-epoch fixture model (test/integration-tests.jl:8-15), evaluates two chains, prints
 * ll and gradient rows as hex floats.  Compiled and run by tests/test_gpu_parity.py::test_plain_c_consumer. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/octo_b200.h"

typedef int (*create_t)(const OctoConstants*, const OctoLayout*, const OctoObsBlock*, int32_t, int32_t, OctoCtx**);
typedef int (*grad_t)(OctoCtx*, const double*, int64_t, int64_t, double*, double*);
typedef void (*destroy_t)(OctoCtx*);
typedef void (*defc_t)(OctoConstants*);
typedef const char* (*err_t)(void);

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s libocto_b200.so\n", argv[0]); return 2; }
    void* h = dlopen(argv[1], RTLD_NOW);
    if (!h) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    create_t create = (create_t)dlsym(h, "octo_create");
    grad_t grad = (grad_t)dlsym(h, "octo_logp_grad");
    destroy_t destroy = (destroy_t)dlsym(h, "octo_destroy");
    defc_t defc = (defc_t)dlsym(h, "octo_default_constants");
    err_t lasterr = (err_t)dlsym(h, "octo_last_error");
    double ep[8] = {50000, 50120, 50240, 50360, 50480, 50600, 50720, 50840};
    double ra[8] = {-505.7637580573554, -502.570356287689, -498.2089148883798, -492.67768482682357,
                    -485.9770335870402, -478.1095526888573, -469.0801731788123, -458.89628893460525};
    double dec[8] = {-66.92982418533026, -37.47217527025044, -7.927548139010479, 21.63557115669823,
                     51.147204404903704, 80.53589069730698, 109.72870493064629, 138.65128697876773};
    double sg[8] = {10, 10, 10, 10, 10, 10, 10, 10};
    OctoObsBlock B; memset(&B, 0, sizeof B);
    B.kind = OCTO_KIND_ASTROM_RADEC; B.planet = 0; B.n_epochs = 8; B.has_cor = 0;
    B.epoch = ep; B.y1 = ra; B.y2 = dec; B.s1 = sg; B.s2 = sg; B.cor = NULL;
    B.idx_jitter = B.idx_platescale = B.idx_northangle = B.idx_offset = -1;
    OctoLayout L; memset(&L, 0, sizeof L);
    L.n_planets = 1; L.n_in = 8;                     /* columns: M, plx, a, e, i, w, W, tp */
    for (int p = 0; p < OCTO_MAX_PLANETS; ++p) L.idx_mass[p] = -1;
    L.idx_M[0] = 0; L.idx_plx[0] = 1; L.idx_a[0] = 2; L.idx_e[0] = 3; L.idx_i[0] = 4; L.idx_w[0] = 5; L.idx_W[0] = 6; L.idx_tp[0] = 7;
    OctoConstants C; defc(&C);
    OctoCtx* ctx = NULL;
    if (create(&C, &L, &B, 1, 0, &ctx)) { fprintf(stderr, "octo_create: %s\n", lasterr()); return 1; }
    enum { N = 2, LD = 3 };                          /* leading dimension larger than the batch on purpose */
    double in[LD * 8], ll[N], g[LD * 8];
    double x0[8] = {1.21, 50.01, 12.1, 0.12, 0.72, 0.65, 0.29, 41500.0};
    for (int k = 0; k < 8; ++k) { in[0 + k * LD] = x0[k]; in[1 + k * LD] = x0[k] * (k == 7 ? 1.0 : 1.01); in[2 + k * LD] = 0; }
    if (grad(ctx, in, N, LD, ll, g)) { fprintf(stderr, "octo_logp_grad: %s\n", lasterr()); return 1; }
    for (int c = 0; c < N; ++c) {
        printf("%a", ll[c]);
        for (int k = 0; k < 8; ++k) printf(" %a", g[c + k * LD]);
        printf("\n");
    }
    destroy(ctx);
    return 0;
}
