/* rhs_from_c.c - the drop-in boundary used from plain C, the way the reference's host code would use it
 * (INTEGRATION.md sections 1-3): load a mesh container, create the GPU context, hand over the forcing arrays and
 * the state, evaluate f(t, y, ydot) through the CVRhsFn-shaped entry point, print a checksum.
 *
 *   gcc -O2 -I include examples/rhs_from_c.c -L shud_up_b200 -lshud_b200 -Wl,-rpath,$PWD/shud_up_b200 -lm -o rhs_from_c
 *   ./rhs_from_c mesh.shudb200 case.bin
 * case.bin: NY doubles of y followed by 8 x Ne doubles (qEleNetPrep, qPotEvap, qPotTran, t_lai, fu_Surf, fu_Sub,
 * qElePrep, qEleE_IC), as tests/test_c_example.py writes it.  Exit code 3 = no CUDA device (there is no CPU path). */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "shud_b200.h"

static double *read_doubles(FILE *fp, size_t n) {
    double *p = (double *)malloc(sizeof(double) * (n ? n : 1));
    if (!p || fread(p, sizeof(double), n, fp) != n) { free(p); return NULL; }
    return p;
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s mesh.shudb200 case.bin\n", argv[0]); return 2; }
    shud_mesh M;
    void *block = NULL;
    if (shud_b200_mesh_load(argv[1], &M, &block) != SHUD_OK) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    const size_t Ne = (size_t)M.Ne, NY = 3 * Ne + (size_t)M.Nr + (size_t)M.Nl;
    FILE *fp = fopen(argv[2], "rb");
    if (!fp) { fprintf(stderr, "cannot read %s\n", argv[2]); return 2; }
    double *y = read_doubles(fp, NY), *arr[8];
    int ok = y != NULL;
    for (int k = 0; k < 8; k++) { arr[k] = ok ? read_doubles(fp, Ne) : NULL; ok = ok && arr[k] != NULL; }
    fclose(fp);
    if (!ok) { fprintf(stderr, "short case file\n"); return 2; }

    shud_ctx *gpu = NULL;
    int rc = shud_b200_create(&M, 0, &gpu);
    if (rc == SHUD_ERR_NO_DEVICE) { fprintf(stderr, "no CUDA device: the library has no CPU fallback\n"); return 3; }
    if (rc != SHUD_OK) { fprintf(stderr, "shud_b200_create failed: %d\n", rc); return 1; }
    shud_forcing F = {0};
    F.qEleNetPrep = arr[0]; F.qPotEvap = arr[1]; F.qPotTran = arr[2]; F.t_lai = arr[3]; F.fu_Surf = arr[4]; F.fu_Sub = arr[5];
    F.qElePrep = arr[6]; F.qEleE_IC = arr[7];
    if ((rc = shud_b200_set_forcing(gpu, &F)) != SHUD_OK) { fprintf(stderr, "set_forcing: %d\n", rc); return 1; }
    if ((rc = shud_b200_prime(gpu, y)) != SHUD_OK) { fprintf(stderr, "prime: %d\n", rc); return 1; }
    double *ydot = (double *)malloc(sizeof(double) * NY);
    rc = shud_b200_rhs(gpu, 0.0, y, ydot); /* what the CVRhsFn registered with CVodeInit calls */
    if (rc != SHUD_OK) { fprintf(stderr, "rhs: code %d (10 = NaN, 13 = data range, 1 = river BC type)\n", rc); return 1; }
    double s = 0., sa = 0.;
    for (size_t i = 0; i < NY; i++) { s += ydot[i]; sa += fabs(ydot[i]); }
    printf("Ne=%d Nr=%d Ns=%d Nl=%d NY=%zu sum(ydot)=%.17g sum|ydot|=%.17g\n", M.Ne, M.Nr, M.Ns, M.Nl, NY, s, sa);
    shud_b200_destroy(gpu);
    shud_b200_mesh_free(block);
    return 0;
}
