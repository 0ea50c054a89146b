/* abi_harness.c -- drives libmct_sm100a.so from plain C exactly as julia/MCTDevice.jl does through ccall: the same entry
 * points, the same struct layouts (mct_td_options, mct_coeffs), host buffers in and out.  Julia is not in the build image,
 * so this is the executable stand-in for the reference-side binding (tests/test_abi_harness.py builds it with gcc against
 * include/mct.h and runs it).
 *
 *   abi_harness --layout                      prints sizeof/offsetof of the two structs (compared with the Julia declaration)
 *   abi_harness in.bin out.bin                in:  int32 Nk, N; float64 rho, dt, t_max, tol; k[Nk], Sk[Nk]
 *                                             solve(MemoryEquation(0, 1, k^2/S, 0, S, 0, ModeCouplingKernel(rho,1,1,k,S)),
 *                                                   TimeDoublingSolver(N, dt, t_max, tolerance=tol))
 *                                             out: int64 nt, kernel_evals; t[nt], F[nt][Nk], K[nt][Nk]; then the evaluation
 *                                             K0 = evaluate_kernel(kernel, S, 0) [Nk] through mct_kernel_eval
 */
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mct.h"

#define CHECK(call)                                                                       \
    do {                                                                                  \
        int rc__ = (call);                                                                \
        if (rc__ != MCT_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc__, mct_last_error()); return 10 + rc__; } \
    } while (0)

int main(int argc, char **argv) {
    if (argc == 2 && strcmp(argv[1], "--layout") == 0) {
        printf("mct_td_options %zu N %zu dt %zu t_max %zu max_iterations %zu tolerance %zu verbose %zu init_with_memory %zu\n",
               sizeof(mct_td_options), offsetof(mct_td_options, N), offsetof(mct_td_options, dt), offsetof(mct_td_options, t_max),
               offsetof(mct_td_options, max_iterations), offsetof(mct_td_options, tolerance), offsetof(mct_td_options, verbose),
               offsetof(mct_td_options, init_with_memory));
        printf("mct_coeffs %zu alpha_kind %zu alpha_scalar %zu alpha %zu beta_kind %zu beta_scalar %zu beta %zu gamma_kind %zu gamma_scalar %zu gamma %zu delta %zu\n",
               sizeof(mct_coeffs), offsetof(mct_coeffs, alpha_kind), offsetof(mct_coeffs, alpha_scalar), offsetof(mct_coeffs, alpha),
               offsetof(mct_coeffs, beta_kind), offsetof(mct_coeffs, beta_scalar), offsetof(mct_coeffs, beta),
               offsetof(mct_coeffs, gamma_kind), offsetof(mct_coeffs, gamma_scalar), offsetof(mct_coeffs, gamma), offsetof(mct_coeffs, delta));
        printf("version %d\n", mct_version());
        return 0;
    }
    if (argc != 3) { fprintf(stderr, "usage: abi_harness --layout | in.bin out.bin\n"); return 2; }
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) { perror(argv[1]); return 2; }
    int32_t hdr[2]; double par[4];
    if (fread(hdr, 4, 2, fi) != 2 || fread(par, 8, 4, fi) != 4) return 3;
    const int Nk = hdr[0];
    double *k = malloc(8 * (size_t)Nk), *Sk = malloc(8 * (size_t)Nk), *gamma = malloc(8 * (size_t)Nk), *dF0 = calloc((size_t)Nk, 8);
    if (fread(k, 8, (size_t)Nk, fi) != (size_t)Nk || fread(Sk, 8, (size_t)Nk, fi) != (size_t)Nk) return 3;
    fclose(fi);
    for (int i = 0; i < Nk; i++) gamma[i] = k[i] * k[i] / Sk[i];

    mct_kernel_t kern = NULL;
    CHECK(mct_sc3d_create_from_sk(&kern, 0, Nk, par[0], 1.0, 1.0, k, Sk));
    mct_td_options opts = {hdr[1], par[1], par[2], 1000000, par[3], 0, 1};
    /* alpha = 0 (UniformScaling), beta = 1 (UniformScaling), gamma = Diagonal(k^2/S), delta = 0 */
    mct_coeffs co = {0, 0.0, NULL, 0, 1.0, NULL, 1, 0.0, gamma, NULL};
    long nt = 0, evals = 0;
    CHECK(mct_td_output_len(&opts, &nt));
    double *t = malloc(8 * (size_t)nt), *F = malloc(8 * (size_t)nt * Nk), *K = malloc(8 * (size_t)nt * Nk), *K0 = malloc(8 * (size_t)Nk);
    CHECK(mct_td_solve(kern, &co, Sk, dF0, &opts, t, F, K, &evals));
    CHECK(mct_kernel_eval(kern, Sk, 0, K0));
    /* error behaviour the binding relies on: a bad grid is an assertion (status 2), not a crash */
    {
        mct_kernel_t bad = NULL;
        double kb[4] = {0.3, 0.5, 0.7, 0.9}, Sb[4] = {1, 1, 1, 1};
        const int rc = mct_sc3d_create_from_sk(&bad, 0, 4, 0.9, 1.0, 1.0, kb, Sb);
        if (rc != MCT_BAD_ARGUMENT) { fprintf(stderr, "grid assertion: expected %d, got %d\n", MCT_BAD_ARGUMENT, rc); return 4; }
    }
    CHECK(mct_kernel_destroy(kern));
    FILE *fo = fopen(argv[2], "wb");
    if (!fo) { perror(argv[2]); return 2; }
    int64_t oh[2] = {nt, evals};
    fwrite(oh, 8, 2, fo); fwrite(t, 8, (size_t)nt, fo); fwrite(F, 8, (size_t)nt * Nk, fo); fwrite(K, 8, (size_t)nt * Nk, fo); fwrite(K0, 8, (size_t)Nk, fo);
    fclose(fo);
    printf("abi_harness: Nk=%d nt=%ld kernel_evals=%ld launches=%ld\n", Nk, nt, evals, mct_launch_count());
    return 0;
}
