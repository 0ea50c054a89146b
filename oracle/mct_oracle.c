/*
 * mct_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY, NOT A PRODUCT PATH.
 *
 * Plain-C restatement, in the reference's own loop order, of the hot path of
 * ModeCouplingTheory.jl v0.8.9 (memory-kernel evaluation + Fuchs time-doubling
 * step).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (libmct_sm100a.so)
 * never does.
 *
 * Parity status: PINNED against the reference's own regression goldens
 * (tests/test_oracle_goldens.py): test/test_MCT.jl:56, test/test_tagged.jl:59,71,
 * test/test_steady_state.jl:64,104, test/test_MCMCT.jl:192,211,212,226,
 * test/test_dDimMCT_crit_pt.jl:96-97.  The reference itself (Julia) cannot be
 * run in this image (no julia binary, no network).
 *
 * All file:line citations are relative to /root/reference/.
 * Every array that mirrors a Julia array keeps Julia's column-major layout;
 * loops are written 0-based but nest exactly as the Julia loops do.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define OK_API __attribute__((visibility("default")))

enum {
    OK_SC3D = 1,      /* ModeCouplingKernel3D            src/Kernels/ModeCouplingKernels.jl:2-17   */
    OK_TAGGED3D = 2,  /* TaggedModeCouplingKernel3D      src/Kernels/ModeCouplingKernels.jl:265-282 */
    OK_MC3D = 3,      /* MultiComponentModeCouplingKernel3D  src/Kernels/MultiComponentKernels.jl:45-60 */
    OK_MC3D_NAIVE = 4,/* NaiveMultiComponentModeCouplingKernel  test/test_MCMCT.jl:6-113 */
    OK_DDIM = 5,      /* dDimModeCouplingKernel          src/Kernels/ModeCouplingKernels.jl:19-71 */
    OK_F2 = 6,        /* SchematicF2Kernel               src/Kernels/SchematicKernels.jl:33-42 */
    OK_TABULATED = 7, /* K depends on the time index only (MSD kernels, :569-582) */
    OK_TAGGED_MC3D = 8,/* TaggedMultiComponentModeCouplingKernel3D  MultiComponentKernels.jl:227-395 */
    OK_DDIM_TAGGED = 9 /* dDimTaggedModeCouplingKernel   ModeCouplingKernels.jl:284-360 */
};

typedef struct ok_kernel {
    int type, Nk, ns;
    double *k;
    /* SC3D / TAGGED3D / TAGGED_MC3D: Nk x Nk column-major tables and scratch */
    double *V1, *V2, *V3, *A1, *A2, *A3, *T1, *T2, *T3;
    long nt;          /* number of stored trajectory points (tagged kinds) */
    double *Fc;       /* collective trajectory [nt][Nk*ncs*ncs] */
    int ncs;          /* species count of the collective trajectory (tagged MC) */
    int s;            /* tagged species (0-based) */
    /* MC3D */
    double *C;        /* [Nk][ns*ns] column-major blocks */
    double *pref;     /* [ns*ns] column-major */
    double *MA1, *MA2, *MA3; /* (ns,ns,Nk,Nk) column-major */
    double *MT1, *MT2, *MT3; /* [Nk][ns*ns] */
    /* MC3D_NAIVE */
    double *V6;       /* (Ns,Ns,Ns,Nk,Nk,Nk) column-major */
    /* DDIM */
    double *DV, *DJ, *P, prefactor;
    /* F2 */
    double nu;
    /* TABULATED */
    double *Ktab;
} ok_kernel;

static double *dalloc(size_t n) {
    double *p = (double *)calloc(n ? n : 1, sizeof(double));
    if (!p) { fprintf(stderr, "oracle: out of memory (%zu doubles)\n", n); abort(); }
    return p;
}
static double *ddup(const double *src, size_t n) {
    double *p = dalloc(n);
    memcpy(p, src, n * sizeof(double));
    return p;
}

OK_API void ok_kernel_free(ok_kernel *K) {
    if (!K) return;
    free(K->k); free(K->V1); free(K->V2); free(K->V3); free(K->A1); free(K->A2); free(K->A3);
    free(K->T1); free(K->T2); free(K->T3); free(K->Fc); free(K->C); free(K->pref);
    free(K->MA1); free(K->MA2); free(K->MA3); free(K->MT1); free(K->MT2); free(K->MT3);
    free(K->V6); free(K->DV); free(K->DJ); free(K->P); free(K->Ktab);
    free(K);
}

static ok_kernel *knew(int type, int Nk, int ns, const double *k) {
    ok_kernel *K = (ok_kernel *)calloc(1, sizeof(ok_kernel));
    K->type = type; K->Nk = Nk; K->ns = ns;
    if (k) K->k = ddup(k, Nk);
    return K;
}

static void alloc_sc_scratch(ok_kernel *K) {
    size_t n2 = (size_t)K->Nk * K->Nk;
    K->A1 = dalloc(n2); K->A2 = dalloc(n2); K->A3 = dalloc(n2);
    K->T1 = dalloc(K->Nk); K->T2 = dalloc(K->Nk); K->T3 = dalloc(K->Nk);
}

/* ModeCouplingKernel(rho,kBT,m,k_array,Sk)  -- src/Kernels/ModeCouplingKernels.jl:94-129 */
OK_API ok_kernel *ok_sc3d_new(int Nk, double rho, double kBT, double m, const double *k, const double *Sk) {
    ok_kernel *K = knew(OK_SC3D, Nk, 1, k);
    size_t n2 = (size_t)Nk * Nk;
    K->V1 = dalloc(n2); K->V2 = dalloc(n2); K->V3 = dalloc(n2);
    alloc_sc_scratch(K);
    double dk = k[1] - k[0];                                   /* :102 */
    double *Ck = dalloc(Nk);
    for (int i = 0; i < Nk; i++) Ck[i] = (Sk[i] - 1.0) / (rho * Sk[i]);   /* :105 */
    double D0 = kBT / m;                                       /* :115 */
    const double pi2x8 = 8.0 * (M_PI * M_PI);
    for (int iq = 0; iq < Nk; iq++)
        for (int ip = 0; ip < Nk; ip++) {                       /* :116-126 */
            double p = k[ip], q = k[iq], cp = Ck[ip], cq = Ck[iq];
            double s = cp + cq, dd = q * q - p * p, dc = cq - cp, dc2 = cq * cq - cp * cp;
            size_t ix = (size_t)iq + (size_t)ip * Nk;
            K->V1[ix] = p * q * (s * s) / 4.0 * D0 * rho / pi2x8 * dk * dk;
            K->V2[ix] = p * q * (dd * dd) * (dc * dc) / 4.0 * D0 * rho / pi2x8 * dk * dk;
            K->V3[ix] = p * q * dd * dc2 / 2.0 * D0 * rho / pi2x8 * dk * dk;
        }
    free(Ck);
    return K;
}

/* Same kernel object built from caller-supplied tables (what the C ABI receives). */
OK_API ok_kernel *ok_sc3d_from_tables(int Nk, const double *k, const double *V1, const double *V2, const double *V3) {
    ok_kernel *K = knew(OK_SC3D, Nk, 1, k);
    size_t n2 = (size_t)Nk * Nk;
    K->V1 = ddup(V1, n2); K->V2 = ddup(V2, n2); K->V3 = ddup(V3, n2);
    alloc_sc_scratch(K);
    return K;
}

/* TaggedModeCouplingKernel(rho,kBT,m,k_array,Sk,sol) -- ModeCouplingKernels.jl:390-426.
 * Fc_traj[it][ik] is sol.F (it = 0 is t = 0); the reference's tDict[t] float lookup (:394,:436)
 * is replaced by the integer index of t in sol.t, which the solver knows. */
OK_API ok_kernel *ok_tagged3d_new(int Nk, double rho, double kBT, double m, const double *k, const double *Sk,
                                  long nt, const double *Fc_traj) {
    ok_kernel *K = knew(OK_TAGGED3D, Nk, 1, k);
    size_t n2 = (size_t)Nk * Nk;
    K->V1 = dalloc(n2); K->V2 = dalloc(n2); K->V3 = dalloc(n2);
    alloc_sc_scratch(K);
    K->nt = nt; K->Fc = ddup(Fc_traj, (size_t)nt * Nk); K->ncs = 1;
    double dk = k[1] - k[0];
    double *Ck = dalloc(Nk);
    for (int i = 0; i < Nk; i++) Ck[i] = (Sk[i] - 1.0) / (rho * Sk[i]);   /* :402 */
    double D0 = kBT / m;
    const double pi2x4 = 4.0 * (M_PI * M_PI);
    double dk2 = dk * dk;
    for (int iq = 0; iq < Nk; iq++)
        for (int ip = 0; ip < Nk; ip++) {                       /* :413-422 */
            double p = k[ip], q = k[iq], cq = Ck[iq];
            double dd = q * q - p * p;
            size_t ix = (size_t)iq + (size_t)ip * Nk;
            K->V1[ix] = p * q * (cq * cq) / 4.0 * D0 * rho / pi2x4 * dk2;
            K->V2[ix] = p * q * (dd * dd) * (cq * cq) / 4.0 * D0 * rho / pi2x4 * dk2;
            K->V3[ix] = p * q * dd * (cq * cq) / 2.0 * D0 * rho / pi2x4 * dk2;
        }
    free(Ck);
    return K;
}

OK_API ok_kernel *ok_tagged3d_from_tables(int Nk, const double *k, const double *V1, const double *V2,
                                          const double *V3, long nt, const double *Fc_traj) {
    ok_kernel *K = knew(OK_TAGGED3D, Nk, 1, k);
    size_t n2 = (size_t)Nk * Nk;
    K->V1 = ddup(V1, n2); K->V2 = ddup(V2, n2); K->V3 = ddup(V3, n2);
    alloc_sc_scratch(K);
    K->nt = nt; K->Fc = ddup(Fc_traj, (size_t)nt * Nk); K->ncs = 1;
    return K;
}

OK_API void ok_get_tables(const ok_kernel *K, double *V1, double *V2, double *V3) {
    size_t n2 = (size_t)K->Nk * K->Nk;
    memcpy(V1, K->V1, n2 * sizeof(double));
    memcpy(V2, K->V2, n2 * sizeof(double));
    memcpy(V3, K->V3, n2 * sizeof(double));
}

/* ---- small dense helpers on ns x ns column-major blocks (stand-ins for StaticArrays SMatrix ops) ---- */
#define NSMAX 8
static void blk_mul(int ns, const double *A, const double *B, double *C) { /* C = A*B, plain left-to-right sums */
    double tmp[NSMAX * NSMAX];
    for (int j = 0; j < ns; j++)
        for (int i = 0; i < ns; i++) {
            double s = A[i] * B[j * ns];
            for (int l = 1; l < ns; l++) s += A[i + l * ns] * B[l + j * ns];
            tmp[i + j * ns] = s;
        }
    memcpy(C, tmp, sizeof(double) * ns * ns);
}
/* X = A \ B by Gaussian elimination with partial pivoting (StaticArrays `\` = closed form for n<=3, LU above;
 * not bit-pinned by the reference -- tolerance-based parity only). */
static int blk_solve(int ns, const double *A, const double *B, double *X) {
    double a[NSMAX * NSMAX], b[NSMAX * NSMAX];
    memcpy(a, A, sizeof(double) * ns * ns);
    memcpy(b, B, sizeof(double) * ns * ns);
    for (int c = 0; c < ns; c++) {
        int piv = c; double best = fabs(a[c + c * ns]);
        for (int r = c + 1; r < ns; r++) if (fabs(a[r + c * ns]) > best) { best = fabs(a[r + c * ns]); piv = r; }
        if (best == 0.0) return 1;
        if (piv != c)
            for (int j = 0; j < ns; j++) {
                double t = a[c + j * ns]; a[c + j * ns] = a[piv + j * ns]; a[piv + j * ns] = t;
                t = b[c + j * ns]; b[c + j * ns] = b[piv + j * ns]; b[piv + j * ns] = t;
            }
        for (int r = c + 1; r < ns; r++) {
            double f = a[r + c * ns] / a[c + c * ns];
            if (f == 0.0) continue;
            for (int j = c; j < ns; j++) a[r + j * ns] -= f * a[c + j * ns];
            for (int j = 0; j < ns; j++) b[r + j * ns] -= f * b[c + j * ns];
        }
    }
    for (int j = 0; j < ns; j++)
        for (int r = ns - 1; r >= 0; r--) {
            double s = b[r + j * ns];
            for (int c = r + 1; c < ns; c++) s -= a[r + c * ns] * X[c + j * ns];
            X[r + j * ns] = s / a[r + r * ns];
        }
    return 0;
}
static int blk_inv(int ns, const double *A, double *Ainv) {
    double I[NSMAX * NSMAX];
    memset(I, 0, sizeof(I));
    for (int i = 0; i < ns; i++) I[i + i * ns] = 1.0;
    return blk_solve(ns, A, I, Ainv);
}

/* direct correlation blocks C[k] = (delta./x - inv(S[k]))/rho_all -- MultiComponentKernels.jl:100-109 */
static void mc_direct_correlation(int Nk, int ns, const double *rho, const double *Sk, double *C) {
    double rho_all = 0.0; for (int a = 0; a < ns; a++) rho_all += rho[a];
    double x[NSMAX]; for (int a = 0; a < ns; a++) x[a] = rho[a] / rho_all;
    for (int i = 0; i < Nk; i++) {
        double Sinv[NSMAX * NSMAX];
        blk_inv(ns, Sk + (size_t)i * ns * ns, Sinv);
        for (int b = 0; b < ns; b++)
            for (int a = 0; a < ns; a++) {
                double d = (a == b) ? 1.0 / x[a] : 0.0 / x[a];
                C[(size_t)i * ns * ns + a + b * ns] = (d - Sinv[a + b * ns]) / rho_all;
            }
    }
}

/* MultiComponentModeCouplingKernel(rho,kBT,m,k_array,Sk) -- MultiComponentKernels.jl:84-131.
 * Sk: [Nk][ns*ns] column-major blocks.  Allocates the reference's (ns,ns,Nk,Nk) scratch. */
OK_API ok_kernel *ok_mc3d_new(int Nk, int ns, const double *rho, double kBT, const double *m, const double *k,
                              const double *Sk) {
    ok_kernel *K = knew(OK_MC3D, Nk, ns, k);
    size_t nb = (size_t)ns * ns;
    K->C = dalloc(nb * Nk); K->pref = dalloc(nb);
    K->MA1 = dalloc(nb * Nk * Nk); K->MA2 = dalloc(nb * Nk * Nk); K->MA3 = dalloc(nb * Nk * Nk);
    K->MT1 = dalloc(nb * Nk); K->MT2 = dalloc(nb * Nk); K->MT3 = dalloc(nb * Nk);
    mc_direct_correlation(Nk, ns, rho, Sk, K->C);
    double dk = k[1] - k[0];
    double rho_all = 0.0; for (int a = 0; a < ns; a++) rho_all += rho[a];
    for (int a = 0; a < ns; a++)
        for (int b = 0; b < ns; b++) {                          /* :122-127 */
            double xb = rho[b] / rho_all;
            double h = dk / 2.0 / M_PI;
            K->pref[a + b * ns] = rho_all * kBT / (8.0 * m[a] * xb) * (h * h);
        }
    return K;
}

/* NaiveMultiComponentModeCouplingKernel -- test/test_MCMCT.jl:19-62 (brute-force triangle sum, O(Ns^4 Nk^3)) */
OK_API ok_kernel *ok_mc3d_naive_new(int Nk, int ns, const double *rho, double kBT, const double *m, const double *k,
                                    const double *Sk) {
    ok_kernel *K = knew(OK_MC3D_NAIVE, Nk, ns, k);
    size_t nb = (size_t)ns * ns;
    K->C = dalloc(nb * Nk); K->pref = dalloc(nb);
    mc_direct_correlation(Nk, ns, rho, Sk, K->C);
    double dk = k[1] - k[0];
    double rho_all = 0.0; for (int a = 0; a < ns; a++) rho_all += rho[a];
    for (int a = 0; a < ns; a++)
        for (int b = 0; b < ns; b++) {                          /* test_MCMCT.jl:41-46 */
            double xb = rho[b] / rho_all;
            double h = dk / 2.0 / M_PI;
            K->pref[a + b * ns] = rho_all * kBT / (2.0 * m[a] * xb) * (h * h);
        }
    size_t n3 = (size_t)Nk * Nk * Nk, s3 = (size_t)ns * ns * ns;
    K->V6 = dalloc(s3 * n3);
    for (int ik = 0; ik < Nk; ik++)
        for (int iq = 0; iq < Nk; iq++)
            for (int ip = 0; ip < Nk; ip++) {
                int lo = abs(iq - ik) + 1, hi = (Nk < ik + iq + 1) ? Nk : ik + iq + 1; /* 1-based: |iq-ik|+1 <= ip <= min(Nk, ik+iq-1) */
                int ip1 = ip + 1;
                if (!(lo <= ip1 && ip1 <= hi)) continue;
                double q = k[iq], p = k[ip], kk = k[ik];
                for (int a = 0; a < ns; a++)
                    for (int b = 0; b < ns; b++)
                        for (int g = 0; g < ns; g++) {              /* test_MCMCT.jl:51-60 */
                            double dbg = (b == g), dag = (a == g);
                            double caq = K->C[(size_t)iq * nb + a + g * ns];
                            double cbp = K->C[(size_t)ip * nb + b + g * ns];
                            size_t ix = (size_t)a + ns * ((size_t)b + ns * ((size_t)g + ns * ((size_t)iq + Nk * ((size_t)ip + (size_t)Nk * ik))));
                            K->V6[ix] = 1.0 / (2.0 * kk) * ((kk * kk + q * q - p * p) * dbg * caq + (kk * kk + p * p - q * q) * dag * cbp);
                        }
            }
    return K;
}

/* TaggedMultiComponentModeCouplingKernel(s, ...) -- MultiComponentKernels.jl:271-333 */
OK_API ok_kernel *ok_tagged_mc3d_new(int s0, int Nk, int ncs, const double *rho, double kBT, const double *m,
                                     const double *k, const double *Sk, long nt, const double *Fc_traj) {
    ok_kernel *K = knew(OK_TAGGED_MC3D, Nk, 1, k);
    size_t n2 = (size_t)Nk * Nk, nb = (size_t)ncs * ncs;
    K->V1 = dalloc(n2); K->V2 = dalloc(n2); K->V3 = dalloc(n2);
    alloc_sc_scratch(K);
    K->s = s0; K->ncs = ncs; K->nt = nt;
    K->Fc = ddup(Fc_traj, (size_t)nt * Nk * nb);
    K->C = dalloc(nb * Nk);
    mc_direct_correlation(Nk, ncs, rho, Sk, K->C);
    double dk = k[1] - k[0];
    double rho_all = 0.0; for (int a = 0; a < ncs; a++) rho_all += rho[a];
    double twopi = 2.0 * M_PI;
    double prefactor = kBT * (dk * dk) * rho_all / (4.0 * m[s0] * (twopi * twopi));   /* :306 */
    for (int iq = 0; iq < Nk; iq++)
        for (int ip = 0; ip < Nk; ip++) {                       /* :307-318 */
            double p = k[ip], q = k[iq], dd = q * q - p * p;
            size_t ix = (size_t)iq + (size_t)ip * Nk;
            K->V1[ix] = prefactor * p * q;
            K->V2[ix] = prefactor * p * q * (dd * dd);
            K->V3[ix] = prefactor * 2.0 * p * q * dd;
        }
    return K;
}

static double surface_d_dim_unit_sphere(double d) {          /* src/HelperFunctions.jl:258-260 */
    return 2.0 * pow(M_PI, d / 2.0) / tgamma(d / 2.0);
}

/* dDimModeCouplingKernel(rho,kBT,m,k_array,Sk,d) -- ModeCouplingKernels.jl:34-71
 * tagged = 1: dDimTaggedModeCouplingKernel -- :301-341 */
OK_API ok_kernel *ok_ddim_new(int Nk, double rho, double kBT, double m, const double *k, const double *Sk, int d,
                              int tagged, long nt, const double *Fc_traj) {
    ok_kernel *K = knew(tagged ? OK_DDIM_TAGGED : OK_DDIM, Nk, 1, k);
    size_t n3 = (size_t)Nk * Nk * Nk;
    K->DV = dalloc(n3); K->DJ = dalloc(n3); K->P = dalloc(Nk);
    if (tagged) { K->nt = nt; K->Fc = ddup(Fc_traj, (size_t)nt * Nk); K->ncs = 1; }
    double dk = k[1] - k[0];
    double *Ck = dalloc(Nk);
    for (int i = 0; i < Nk; i++) Ck[i] = (1.0 - 1.0 / Sk[i]) / rho;       /* :41-46 */
    K->prefactor = (kBT / m) * rho * (dk * dk) * surface_d_dim_unit_sphere(d - 1) / pow(4.0 * M_PI, d);  /* :48 */
    if (tagged) K->prefactor = 2.0 * (kBT / m) * rho * (dk * dk) * surface_d_dim_unit_sphere(d - 1) / pow(4.0 * M_PI, d); /* :321 */
    for (int iq = 0; iq < Nk; iq++)
        for (int ik = 0; ik < Nk; ik++)
            for (int ip = 0; ip < Nk; ip++) {                   /* :53-66 */
                int lo = abs(iq - ik) + 1, hi = (Nk < ik + iq + 1) ? Nk : ik + iq + 1;
                int ip1 = ip + 1;
                if (!(lo <= ip1 && ip1 <= hi)) continue;
                double q = k[iq], p = k[ip], kk = k[ik];
                double cp = Ck[ip], ck = Ck[ik];
                double w = q * q + kk * kk - p * p;
                size_t ix = (size_t)iq + Nk * ((size_t)ik + (size_t)Nk * ip);
                K->DJ[ix] = kk * p * pow(4.0 * (q * q) * (kk * kk) - w * w, (d - 3) / 2.0) / pow(q, d);
                double v = tagged ? (w * ck) : (w * ck + (q * q - kk * kk + p * p) * cp);
                K->DV[ix] = v * v;
            }
    free(Ck);
    return K;
}

OK_API void ok_ddim_get_tables(const ok_kernel *K, double *V, double *J, double *prefactor) {
    size_t n3 = (size_t)K->Nk * K->Nk * K->Nk;
    memcpy(V, K->DV, n3 * sizeof(double));
    memcpy(J, K->DJ, n3 * sizeof(double));
    *prefactor = K->prefactor;
}

OK_API ok_kernel *ok_f2_new(double nu) {
    ok_kernel *K = knew(OK_F2, 1, 1, NULL);
    K->nu = nu;
    return K;
}

OK_API ok_kernel *ok_tabulated_new(long nt, const double *Ktab) {
    ok_kernel *K = knew(OK_TABULATED, 1, 1, NULL);
    K->nt = nt; K->Ktab = ddup(Ktab, nt);
    return K;
}

/* MSDModeCouplingKernel evaluate_kernel -- ModeCouplingKernels.jl:569-582: K(t_it) for every stored time. */
OK_API void ok_msd3d_table(int Nk, double rho, double kBT, double m, const double *k, const double *Sk, long nt,
                           const double *F, const double *Fs, double *Kout) {
    double dk = k[1] - k[0];
    double *Ck = dalloc(Nk);
    for (int i = 0; i < Nk; i++) Ck[i] = (Sk[i] - 1.0) / (rho * Sk[i]);   /* :564 */
    for (long it = 0; it < nt; it++) {
        double K = 0.0;
        for (int iq = 0; iq < Nk; iq++) {
            double k2 = k[iq] * k[iq];
            K += (k2 * k2) * (Ck[iq] * Ck[iq]) * F[it * Nk + iq] * Fs[it * Nk + iq];
        }
        K *= dk * rho * kBT / (6.0 * (M_PI * M_PI) * m);
        Kout[it] = K;
    }
    free(Ck);
}

/* MSDMultiComponentModeCouplingKernel evaluate_kernel -- MultiComponentKernels.jl:468-486 */
OK_API void ok_msd_mc3d_table(int s0, int Nk, int ns, const double *rho, double kBT, const double *m, const double *k,
                              const double *Sk, long nt, const double *F, const double *Fs, double *Kout) {
    size_t nb = (size_t)ns * ns;
    double *C = dalloc(nb * Nk);
    mc_direct_correlation(Nk, ns, rho, Sk, C);
    double dk = k[1] - k[0];
    double rho_all = 0.0; for (int a = 0; a < ns; a++) rho_all += rho[a];
    for (long it = 0; it < nt; it++) {
        double K = 0.0;
        /* Reference quirk kept on purpose: `Ns = length(Fs[1])` (:477) takes the length of a Float64, i.e. 1,
         * so the species double sum only ever visits alpha = beta = 1.  The golden test/test_MCMCT.jl:226
         * (0.6057207573375025) is produced by exactly this behaviour. */
        int nsq = 1;
        for (int a = 0; a < nsq; a++)
            for (int b = 0; b < nsq; b++)
                for (int iq = 0; iq < Nk; iq++) {
                    double k2 = k[iq] * k[iq];
                    K += (k2 * k2) * C[iq * nb + a + s0 * ns] * C[iq * nb + s0 + b * ns] *
                         F[((size_t)it * Nk + iq) * nb + a + b * ns] * Fs[(size_t)it * Nk + iq];
                }
        K *= dk * rho_all * kBT / (6.0 * (M_PI * M_PI) * m[s0]);
        Kout[it] = K;
    }
    free(C);
}

/* bengtzelius3! -- ModeCouplingKernels.jl:154-188 (sequential recursion, reference order) */
static void bengtzelius3(double *T1, double *T2, double *T3, const double *A1, const double *A2, const double *A3, int Nk) {
    double T01 = 0.0, T02 = 0.0, T03 = 0.0;
    for (int iq = 0; iq < Nk; iq++) {
        size_t ix = (size_t)iq + (size_t)iq * Nk;
        T01 += A1[ix]; T02 += A2[ix]; T03 += A3[ix];
    }
    T1[0] = T01; T2[0] = T02; T3[0] = T03;
    for (int ik = 2; ik <= Nk; ik++) {          /* 1-based ik as in the reference */
        int qmax = Nk - ik + 1;
        double t1 = T1[ik - 2], t2 = T2[ik - 2], t3 = T3[ik - 2];
        for (int iq = 1; iq <= qmax; iq++) {
            int ip = iq + ik - 1;
            size_t a = (size_t)(iq - 1) + (size_t)(ip - 1) * Nk, b = (size_t)(ip - 1) + (size_t)(iq - 1) * Nk;
            t1 += A1[a] + A1[b]; t2 += A2[a] + A2[b]; t3 += A3[a] + A3[b];
        }
        for (int iq = 1; iq <= ik - 1; iq++) {
            int ip = ik - iq;
            size_t a = (size_t)(iq - 1) + (size_t)(ip - 1) * Nk;
            t1 -= A1[a]; t2 -= A2[a]; t3 -= A3[a];
        }
        T1[ik - 1] = t1; T2[ik - 1] = t2; T3[ik - 1] = t3;
    }
}

static void sc_combine(const ok_kernel *K, double *out) {     /* ModeCouplingKernels.jl:224-227 */
    for (int ik = 0; ik < K->Nk; ik++) {
        double k = K->k[ik];
        out[ik] = k * K->T1[ik] + K->T2[ik] / (k * k * k) + K->T3[ik] / k;
    }
}

/* evaluate_kernel!(out, kernel, F, t).  t_index = index of t in the collective solution's sol.t
 * (0 = t=0); ignored by kernels that do not depend on t. Output: Nk*ns*ns doubles. */
OK_API void ok_kernel_eval(ok_kernel *K, const double *F, long t_index, double *out) {
    int Nk = K->Nk, ns = K->ns;
    switch (K->type) {
    case OK_SC3D: {
        /* fill_A! :190-208.  The reference's loop nest is under @turbo (LoopVectorization picks the order and vectorises
         * along the contiguous index iq); the element-wise products do not depend on the order, so iq is the inner loop */
        for (int ip = 0; ip < Nk; ip++)
            for (int iq = 0; iq < Nk; iq++) {
                double f4 = F[ip] * F[iq];
                size_t ix = (size_t)iq + (size_t)ip * Nk;
                K->A1[ix] = K->V1[ix] * f4; K->A2[ix] = K->V2[ix] * f4; K->A3[ix] = K->V3[ix] * f4;
            }
        bengtzelius3(K->T1, K->T2, K->T3, K->A1, K->A2, K->A3, Nk);
        sc_combine(K, out);
    } break;
    case OK_TAGGED3D: {
        /* fill_A!(kernel,F,t) :428-448 */
        const double *Fc = K->Fc + (size_t)t_index * Nk;
        for (int ip = 0; ip < Nk; ip++)
            for (int iq = 0; iq < Nk; iq++) {
                double f4 = F[ip] * Fc[iq];
                size_t ix = (size_t)iq + (size_t)ip * Nk;
                K->A1[ix] = K->V1[ix] * f4; K->A2[ix] = K->V2[ix] * f4; K->A3[ix] = K->V3[ix] * f4;
            }
        bengtzelius3(K->T1, K->T2, K->T3, K->A1, K->A2, K->A3, Nk);
        sc_combine(K, out);
    } break;
    case OK_TAGGED_MC3D: {
        /* fill_A! MultiComponentKernels.jl:335-370 */
        int nc = K->ncs; size_t nb = (size_t)nc * nc; int s0 = K->s;
        const double *Fc = K->Fc + (size_t)t_index * Nk * nb;
        for (int ip = 0; ip < Nk; ip++)
            for (int iq = 0; iq < Nk; iq++) {
                double fp = F[ip], a1 = 0.0, a2 = 0.0, a3 = 0.0;
                size_t ix = (size_t)iq + (size_t)ip * Nk;
                for (int a = 0; a < nc; a++)
                    for (int b = 0; b < nc; b++) {
                        double csa = K->C[iq * nb + s0 + a * nc], csb = K->C[iq * nb + s0 + b * nc];
                        double fq = Fc[iq * nb + a + b * nc];
                        double f4c2 = fp * fq * csa * csb;
                        a1 += K->V1[ix] * f4c2; a2 += K->V2[ix] * f4c2; a3 += K->V3[ix] * f4c2;
                    }
                K->A1[ix] = a1; K->A2[ix] = a2; K->A3[ix] = a3;
            }
        bengtzelius3(K->T1, K->T2, K->T3, K->A1, K->A2, K->A3, Nk);
        sc_combine(K, out);
    } break;
    case OK_MC3D: {
        size_t nb = (size_t)ns * ns;
        /* fill_A! MultiComponentKernels.jl:134-198 (reference loop order; O(Ns^4 Nk^2)) */
        for (int a = 0; a < ns; a++)
            for (int b = 0; b < ns; b++) {
                double pab = K->pref[a + b * ns];
                for (int iq = 0; iq < Nk; iq++) {
                    double q = K->k[iq];
                    const double *Cq = K->C + iq * nb, *Fq = F + iq * nb;
                    for (int ip = 0; ip < Nk; ip++) {
                        const double *Cp = K->C + ip * nb, *Fp = F + ip * nb;
                        double A1 = 0.0, A2 = 0.0, A3 = 0.0;
                        double p = K->k[ip];
                        double pq2 = p * p - q * q;
                        double P1 = p * q, P2 = p * q * (pq2 * pq2), P3 = 2.0 * p * q * pq2;
                        double fqab = Fq[a + b * ns], fpab = Fp[a + b * ns];
                        for (int d1 = 0; d1 < ns; d1++) {
                            double cp1a = Cp[d1 + a * ns], cq1a = Cq[d1 + a * ns];
                            double fq1b = Fq[d1 + b * ns], fp1b = Fp[d1 + b * ns];
                            for (int d2 = 0; d2 < ns; d2++) {
                                double Vpp = cp1a * Cp[d2 + b * ns], Vpq = cp1a * Cq[d2 + b * ns];
                                double Vqp = cq1a * Cp[d2 + b * ns], Vqq = cq1a * Cq[d2 + b * ns];
                                double t1 = Vpp * fqab * Fp[d1 + d2 * ns];
                                double t2 = Vpq * Fq[a + d2 * ns] * fp1b;
                                double t3 = Vqp * fq1b * Fp[a + d2 * ns];
                                double t4 = Vqq * Fq[d1 + d2 * ns] * fpab;
                                A1 += t1 + t2 + t3 + t4;
                                A2 += t1 - t2 - t3 + t4;
                                A3 += t1 - t4;
                            }
                        }
                        size_t ix = (size_t)a + ns * ((size_t)b + ns * ((size_t)iq + (size_t)Nk * ip));
                        K->MA1[ix] = A1 * P1 * pab; K->MA2[ix] = A2 * P2 * pab; K->MA3[ix] = A3 * P3 * pab;
                    }
                }
            }
        /* bengtzelius3!(T,A,Nk,Ns) :2-43, element-wise on the species block */
        for (size_t e = 0; e < nb; e++) {
            double t1 = 0.0, t2 = 0.0, t3 = 0.0;
            for (int iq = 0; iq < Nk; iq++) {
                size_t ix = e + nb * ((size_t)iq + (size_t)Nk * iq);
                t1 += K->MA1[ix]; t2 += K->MA2[ix]; t3 += K->MA3[ix];
            }
            K->MT1[e] = t1; K->MT2[e] = t2; K->MT3[e] = t3;
            for (int ik = 2; ik <= Nk; ik++) {
                for (int iq = 1; iq <= Nk - ik + 1; iq++) {
                    int ip = iq + ik - 1;
                    size_t x = e + nb * ((size_t)(iq - 1) + (size_t)Nk * (ip - 1));
                    size_t y = e + nb * ((size_t)(ip - 1) + (size_t)Nk * (iq - 1));
                    t1 += K->MA1[x]; t1 += K->MA1[y];
                    t2 += K->MA2[x]; t2 += K->MA2[y];
                    t3 += K->MA3[x]; t3 += K->MA3[y];
                }
                for (int iq = 1; iq <= ik - 1; iq++) {
                    int ip = ik - iq;
                    size_t x = e + nb * ((size_t)(iq - 1) + (size_t)Nk * (ip - 1));
                    t1 -= K->MA1[x]; t2 -= K->MA2[x]; t3 -= K->MA3[x];
                }
                K->MT1[(size_t)(ik - 1) * nb + e] = t1; K->MT2[(size_t)(ik - 1) * nb + e] = t2; K->MT3[(size_t)(ik - 1) * nb + e] = t3;
            }
        }
        for (int ik = 0; ik < Nk; ik++) {                       /* :215-218 */
            double k = K->k[ik];
            for (size_t e = 0; e < nb; e++)
                out[ik * nb + e] = k * K->MT1[ik * nb + e] + K->MT2[ik * nb + e] / (k * k * k) + K->MT3[ik * nb + e] / k;
        }
    } break;
    case OK_MC3D_NAIVE: {
        size_t nb = (size_t)ns * ns;
        /* helper! test/test_MCMCT.jl:64-88 */
        for (int a = 0; a < ns; a++)
            for (int b = 0; b < ns; b++)
                for (int ik = 0; ik < Nk; ik++) {
                    double P = 0.0;
                    for (int m2 = 0; m2 < ns; m2++)
                        for (int mu = 0; mu < ns; mu++)
                            for (int n2 = 0; n2 < ns; n2++)
                                for (int nu = 0; nu < ns; nu++)
                                    for (int iq = 0; iq < Nk; iq++)
                                        for (int ip = 0; ip < Nk; ip++) {
                                            size_t base = ns * ns * ns * ((size_t)iq + Nk * ((size_t)ip + (size_t)Nk * ik));
                                            double Va = K->V6[base + m2 + ns * (n2 + ns * a)];
                                            double Vb = K->V6[base + mu + ns * (nu + ns * b)];
                                            P += K->pref[a + b * ns] * K->k[ip] * K->k[iq] * Va * F[iq * nb + m2 + mu * ns] *
                                                 F[ip * nb + n2 + nu * ns] * Vb / K->k[ik];
                                        }
                    out[ik * nb + a + b * ns] = P;
                }
    } break;
    case OK_DDIM: case OK_DDIM_TAGGED: {
        /* ModeCouplingKernels.jl:237-256 / :343-360 */
        const double *Fa = (K->type == OK_DDIM_TAGGED) ? K->Fc + (size_t)t_index * Nk : F;
        for (int i = 0; i < Nk; i++) K->P[i] = 0.0;
        for (int iq = 0; iq < Nk; iq++)
            for (int ik = 0; ik < Nk; ik++)
                for (int ip = 0; ip < Nk; ip++) {
                    size_t ix = (size_t)iq + Nk * ((size_t)ik + (size_t)Nk * ip);
                    K->P[iq] += K->prefactor * K->DJ[ix] * K->DV[ix] * Fa[ik] * F[ip];
                }
        for (int i = 0; i < Nk; i++) out[i] = K->P[i];
    } break;
    case OK_F2: out[0] = K->nu * (F[0] * F[0]); break;          /* SchematicKernels.jl:39-42 */
    case OK_TABULATED: out[0] = K->Ktab[t_index]; break;
    default: fprintf(stderr, "oracle: unknown kernel type\n"); abort();
    }
}

/* ------------------------------------------------------------------------------------------------
 * Memory equation + solvers.  F, K, coefficients are stored as Nk blocks of ns x ns (column-major);
 * ns = 1 for the single-component kernels (K is then the `.diag` of the reference's Diagonal).
 * Multiplicative coefficients alpha/beta/gamma are either a scalar lambda (UniformScaling, kind 0) or
 * per-k blocks (Diagonal, kind 1)  -- src/MemoryEquation.jl:15-21, src/HelperFunctions.jl:73-84.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int kind;          /* 0 = lambda*I, 1 = per-k blocks */
    double lambda;
    const double *blk; /* [Nk][ns*ns] */
} ok_coef;

static int coef_iszero(const ok_coef *c, int Nk, int ns) {
    if (c->kind == 0) return c->lambda == 0.0;
    for (size_t i = 0; i < (size_t)Nk * ns * ns; i++) if (c->blk[i] != 0.0) return 0;
    return 1;
}
/* out = c * x for block k */
static void coef_mul(const ok_coef *c, int ns, int k, const double *x, double *out) {
    if (c->kind == 0) { for (int e = 0; e < ns * ns; e++) out[e] = c->lambda * x[e]; }
    else blk_mul(ns, c->blk + (size_t)k * ns * ns, x, out);
}
/* out = (-c) \ x for block k */
static void coef_negsolve(const ok_coef *c, int ns, int k, const double *x, double *out) {
    if (c->kind == 0) { for (int e = 0; e < ns * ns; e++) out[e] = x[e] / (-c->lambda); }
    else {
        double neg[NSMAX * NSMAX];
        for (int e = 0; e < ns * ns; e++) neg[e] = -c->blk[(size_t)k * ns * ns + e];
        blk_solve(ns, neg, x, out);
    }
}
/* acc += s * c  (acc is a block; UniformScaling adds on the diagonal only) */
static void coef_scaled_to_block(const ok_coef *c, int ns, int k, double s, double *out) {
    memset(out, 0, sizeof(double) * ns * ns);
    if (c->kind == 0) { for (int i = 0; i < ns; i++) out[i + i * ns] = s * c->lambda; }
    else for (int e = 0; e < ns * ns; e++) out[e] = s * c->blk[(size_t)k * ns * ns + e];
}

typedef struct {
    int Nk, ns;
    ok_coef alpha, beta, gamma;
    const double *delta;  /* [Nk][ns*ns] */
    const double *F0, *dF0;
    double *K0;
    ok_kernel *kernel;
    const long *tmap;     /* optional: kernel time index for every stored output index (NULL = identity) */
} ok_equation;

/* solve(equation, EulerSolver) restricted to what initialize_F_temp_Euler! consumes
 * -- src/EulerSolver.jl:92-112 with :26-64; called from src/TimeDoublingSolver.jl:96-109.
 * nsteps Euler steps of size dt; Fout/Kout [nsteps][dof] hold F,K at t = dt*it, it=1..nsteps. */
static void euler_bootstrap(const ok_equation *eq, double dt, int nsteps, double *Fout, double *Kout) {
    int Nk = eq->Nk, ns = eq->ns; size_t nb = (size_t)ns * ns, dof = nb * Nk;
    double *Karr = dalloc(dof * (nsteps + 1));   /* K_array: K at t_0..t_nsteps */
    double *dFarr = dalloc(dof * (nsteps + 1));  /* dF at t_0..t_nsteps (the reference keeps it reversed) */
    double *Farr = dalloc(dof * (nsteps + 1));
    double *integ = dalloc(dof), *f = dalloc(dof * (nsteps + 1));
    memcpy(Karr, eq->K0, dof * sizeof(double));
    memcpy(dFarr, eq->dF0, dof * sizeof(double));
    memcpy(Farr, eq->F0, dof * sizeof(double));
    int second_order = !coef_iszero(&eq->alpha, Nk, ns);
    for (int it = 1; it <= nsteps; it++) {
        const double *dF_old = dFarr + dof * (it - 1), *F_old = Farr + dof * (it - 1);
        /* construct_euler_integrand! :26-31: integrand[i] = K_array[i] * dF_reverse[i], i = 1..it */
        for (int i = 1; i <= it; i++)
            for (int k = 0; k < Nk; k++)
                blk_mul(ns, Karr + dof * (i - 1) + k * nb, dFarr + dof * (it - i) + k * nb, f + dof * (i - 1) + k * nb);
        /* trapezoidal_integration :33-45 */
        for (size_t e = 0; e < dof; e++) {
            if (it == 1) { integ[e] = f[e] * dt; continue; }
            double r = f[e] / 2.0;
            r += f[dof * (it - 1) + e] / 2.0;
            for (int i2 = 2; i2 <= it - 1; i2++) r += f[dof * (i2 - 1) + e];
            integ[e] = r * dt;
        }
        /* Euler_step :48-64 */
        double *Fn = Farr + dof * it, *dFn = dFarr + dof * it;
        for (int k = 0; k < Nk; k++) {
            double gF[NSMAX * NSMAX], bdF[NSMAX * NSMAX], rhs[NSMAX * NSMAX], sol[NSMAX * NSMAX];
            coef_mul(&eq->gamma, ns, k, F_old + k * nb, gF);
            if (second_order) {
                coef_mul(&eq->beta, ns, k, dF_old + k * nb, bdF);
                for (size_t e = 0; e < nb; e++) rhs[e] = bdF[e] + gF[e] + eq->delta[k * nb + e] + integ[k * nb + e];
                coef_negsolve(&eq->alpha, ns, k, rhs, sol);       /* ddF */
                for (size_t e = 0; e < nb; e++) {
                    dFn[k * nb + e] = dF_old[k * nb + e] + dt * sol[e];
                    Fn[k * nb + e] = F_old[k * nb + e] + dt * dFn[k * nb + e];
                }
            } else {
                for (size_t e = 0; e < nb; e++) rhs[e] = gF[e] + eq->delta[k * nb + e] + integ[k * nb + e];
                coef_negsolve(&eq->beta, ns, k, rhs, sol);        /* dF */
                for (size_t e = 0; e < nb; e++) {
                    dFn[k * nb + e] = sol[e];
                    Fn[k * nb + e] = F_old[k * nb + e] + dt * sol[e];
                }
            }
        }
        long tix = eq->tmap ? eq->tmap[it] : it;
        ok_kernel_eval(eq->kernel, Fn, tix, Karr + dof * it);
    }
    memcpy(Fout, Farr + dof, dof * nsteps * sizeof(double));
    memcpy(Kout, Karr + dof, dof * nsteps * sizeof(double));
    free(Karr); free(dFarr); free(Farr); free(integ); free(f);
}

OK_API long ok_td_output_len(int N, double dt, double tmax) {
    long n = 1 + 2L * N;
    while (dt < tmax * 2.0) { n += 2L * N; dt *= 2.0; }          /* src/TimeDoublingSolver.jl:523 */
    return n;
}

/* find_error -- src/HelperFunctions.jl:44-68 */
static double find_error(const double *a, const double *b, size_t n) {
    double err = 0.0;
    for (size_t i = 0; i < n; i++) { double e = fabs(a[i] - b[i]); if (e > err) err = e; }
    return err;
}

/*
 * solve(equation, TimeDoublingSolver) -- src/TimeDoublingSolver.jl:512-532 and everything it calls.
 * coefficient kinds: 0 scalar, 1 per-k blocks.  scalar_mode = 1 follows the allocating branch taken for
 * immutable F (Float64): compute_C3_immutable ordering (:213-235) and F_old reset every step (:400,:412).
 * kernel_tmap: optional [nt] table mapping this solve's output index -> kernel time index (tagged kernels);
 * NULL means identity.  Returns 0 ok, 1 not converged.
 * Outputs: t_out[nt], F_out[nt][dof], K_out[nt][dof].
 */
OK_API int ok_td_solve(ok_kernel *kernel, int Nk, int ns,
                       int akind, double alam, const double *ablk,
                       int bkind, double blam, const double *bblk,
                       int gkind, double glam, const double *gblk,
                       const double *delta, const double *F0, const double *dF0,
                       int N, double dt0, double tmax, long max_iterations, double tolerance,
                       int scalar_mode, int init_with_memory, const long *kernel_tmap,
                       double *t_out, double *F_out, double *K_out, long *kernel_evals_out, long max_evals) {
    size_t nb = (size_t)ns * ns, dof = nb * Nk;
    int n4 = 4 * N, n2 = 2 * N;
    ok_equation eq;
    eq.Nk = Nk; eq.ns = ns;
    eq.alpha = (ok_coef){akind, alam, ablk}; eq.beta = (ok_coef){bkind, blam, bblk}; eq.gamma = (ok_coef){gkind, glam, gblk};
    eq.delta = delta; eq.F0 = F0; eq.dF0 = dF0; eq.kernel = kernel; eq.tmap = kernel_tmap;
    eq.K0 = dalloc(dof);
    ok_kernel_eval(kernel, F0, kernel_tmap ? kernel_tmap[0] : 0, eq.K0);    /* src/MemoryEquation.jl:45 */

    /* allocate_temporary_arrays :57-88 (slots are 1-based in the reference: slot s lives at index s-1) */
    double *Ft = dalloc(dof * n4), *Kt = dalloc(dof * n4), *FI = dalloc(dof * n4), *KI = dalloc(dof * n4);
    double *C1 = dalloc(dof), *C2 = dalloc(dof), *C3 = dalloc(dof), *tv = dalloc(dof), *Fold = dalloc(dof), *Fold0 = dalloc(dof);
    for (int k = 0; k < Nk; k++) blk_mul(ns, eq.K0 + k * nb, F0 + k * nb, Fold + k * nb);     /* F_old = K0*F0 :64 */
    memcpy(Fold0, Fold, dof * sizeof(double));
    for (int s = 0; s < n4; s++) {
        memcpy(Ft + dof * s, Fold, dof * sizeof(double));                                        /* :82 */
        memcpy(FI + dof * s, Fold, dof * sizeof(double));                                        /* :84 */
        for (size_t e = 0; e < dof; e++) { Kt[dof * s + e] = eq.K0[e] + eq.K0[e]; KI[dof * s + e] = Kt[dof * s + e]; } /* :83,:85 */
    }
#define FT(s) (Ft + dof * ((s) - 1))
#define KT(s) (Kt + dof * ((s) - 1))
#define FINT(s) (FI + dof * ((s) - 1))
#define KINT(s) (KI + dof * ((s) - 1))

    double delta_t = dt0 / (4.0 * N);
    /* initialize_F_temp! :116-148 */
    if (init_with_memory) {
        /* initialize_F_temp_Euler! :96-109 */
        double *Fe = dalloc(dof * n2), *Ke = dalloc(dof * n2);
        euler_bootstrap(&eq, delta_t, n2, Fe, Ke);
        memcpy(Ft, Fe, dof * n2 * sizeof(double));
        free(Fe); free(Ke);
    } else {
        double *Fo = ddup(F0, dof), *dFo = ddup(dF0, dof);
        int second_order = !coef_iszero(&eq.alpha, Nk, ns);
        for (int it = 1; it <= n2; it++) {
            for (int k = 0; k < Nk; k++) {
                double gF[NSMAX * NSMAX], bdF[NSMAX * NSMAX], rhs[NSMAX * NSMAX], sol[NSMAX * NSMAX];
                coef_mul(&eq.gamma, ns, k, Fo + k * nb, gF);
                if (second_order) {
                    coef_mul(&eq.beta, ns, k, dFo + k * nb, bdF);
                    for (size_t e = 0; e < nb; e++) rhs[e] = bdF[e] + gF[e] + delta[k * nb + e];
                    coef_negsolve(&eq.alpha, ns, k, rhs, sol);
                    for (size_t e = 0; e < nb; e++) {
                        dFo[k * nb + e] = dFo[k * nb + e] + delta_t * sol[e];
                        Fo[k * nb + e] = Fo[k * nb + e] + delta_t * dFo[k * nb + e];
                    }
                } else {
                    for (size_t e = 0; e < nb; e++) rhs[e] = gF[e] + delta[k * nb + e];
                    coef_negsolve(&eq.beta, ns, k, rhs, sol);
                    for (size_t e = 0; e < nb; e++) { dFo[k * nb + e] = sol[e]; Fo[k * nb + e] = Fo[k * nb + e] + delta_t * sol[e]; }
                }
            }
            memcpy(FT(it), Fo, dof * sizeof(double));
        }
        free(Fo); free(dFo);
    }
    /* initialize_K_temp! :155-166 */
    for (int it = 1; it <= n2; it++) ok_kernel_eval(kernel, FT(it), kernel_tmap ? kernel_tmap[it] : it, KT(it));
    /* initialize_integrals! :174-199 */
    for (size_t e = 0; e < dof; e++) {
        FINT(1)[e] = (FT(1)[e] + F0[e]) / 2.0;
        KINT(1)[e] = (3.0 * KT(1)[e] - KT(2)[e]) / 2.0;
    }
    for (int it = 2; it <= n2; it++)
        for (size_t e = 0; e < dof; e++) {
            FINT(it)[e] = (FT(it)[e] + FT(it - 1)[e]) / 2.0;
            KINT(it)[e] = (KT(it)[e] + KT(it - 1)[e]) / 2.0;
        }

    /* initialize_output_arrays :507-509, allocate_results!(istart=1, iend=2N) :518 */
    long nout = 0;
    t_out[nout] = 0.0; memcpy(F_out, F0, dof * sizeof(double)); memcpy(K_out, eq.K0, dof * sizeof(double)); nout++;
    for (int it = 1; it <= n2; it++) {
        t_out[nout] = delta_t * it;
        memcpy(F_out + dof * nout, FT(it), dof * sizeof(double));
        memcpy(K_out + dof * nout, KT(it), dof * sizeof(double));
        nout++;
    }

    long kernel_evals = 0;
    int status = 0;
    double Dt = dt0;
    double blkA[NSMAX * NSMAX], blkB[NSMAX * NSMAX], blkC[NSMAX * NSMAX];
    while (Dt < tmax * 2.0 && status == 0) {                       /* :523 */
        double dlt = Dt / (4.0 * N);
        /* do_time_steps! :393-421 */
        for (int it = n2 + 1; it <= n4 && status == 0; it++) {
            int i2 = it / 2;
            if (scalar_mode) memcpy(Fold, Fold0, dof * sizeof(double));   /* immutable path: F_old = temp_arrays.F_old :400 */
            /* update_Fuchs_parameters! :300-319 */
            for (int k = 0; k < Nk; k++) {
                double *c1 = C1 + k * nb;
                coef_scaled_to_block(&eq.alpha, ns, k, 2.0 / (dlt * dlt), blkA);
                coef_scaled_to_block(&eq.beta, ns, k, 3.0 / (2.0 * dlt), blkB);
                coef_scaled_to_block(&eq.gamma, ns, k, 1.0, blkC);
                for (size_t e = 0; e < nb; e++) c1[e] = ((blkA[e] + blkB[e]) + KINT(1)[k * nb + e]) + blkC[e];   /* :310/:314 */
                for (size_t e = 0; e < nb; e++) C2[k * nb + e] = FINT(1)[k * nb + e] - F0[k * nb + e];           /* :311/:315 */
            }
            if (!scalar_mode) {
                /* compute_C3_mutable! :243-288 */
                for (size_t e = 0; e < dof; e++) tv[e] = (5.0 * FT(it - 1)[e] - 4.0 * FT(it - 2)[e] + FT(it - 3)[e]) / (dlt * dlt);
                for (int k = 0; k < Nk; k++) coef_mul(&eq.alpha, ns, k, tv + k * nb, C3 + k * nb);
                for (size_t e = 0; e < dof; e++) tv[e] = 2.0 / dlt * FT(it - 1)[e] - FT(it - 2)[e] / (2.0 * dlt);
                for (int k = 0; k < Nk; k++) { coef_mul(&eq.beta, ns, k, tv + k * nb, blkA); for (size_t e = 0; e < nb; e++) C3[k * nb + e] += blkA[e]; }
                for (int k = 0; k < Nk; k++) { blk_mul(ns, KT(it - i2) + k * nb, FT(i2) + k * nb, blkA); for (size_t e = 0; e < nb; e++) C3[k * nb + e] -= blkA[e]; }
                for (int k = 0; k < Nk; k++) { blk_mul(ns, KT(it - 1) + k * nb, FINT(1) + k * nb, blkA); for (size_t e = 0; e < nb; e++) C3[k * nb + e] += blkA[e]; }
                for (int k = 0; k < Nk; k++) { blk_mul(ns, KINT(1) + k * nb, FT(it - 1) + k * nb, blkA); for (size_t e = 0; e < nb; e++) C3[k * nb + e] += blkA[e]; }
                for (int j = 2; j <= i2; j++)
                    for (int k = 0; k < Nk; k++) {
                        for (size_t e = 0; e < nb; e++) blkB[e] = KT(it - j)[k * nb + e] - KT(it - j + 1)[k * nb + e];
                        blk_mul(ns, blkB, FINT(j) + k * nb, blkA);
                        for (size_t e = 0; e < nb; e++) C3[k * nb + e] += blkA[e];
                    }
                for (int j = 2; j <= it - i2; j++)
                    for (int k = 0; k < Nk; k++) {
                        for (size_t e = 0; e < nb; e++) blkB[e] = FT(it - j)[k * nb + e] - FT(it - j + 1)[k * nb + e];
                        blk_mul(ns, KINT(j) + k * nb, blkB, blkA);
                        for (size_t e = 0; e < nb; e++) C3[k * nb + e] += blkA[e];
                    }
                for (size_t e = 0; e < dof; e++) C3[e] -= delta[e];
            } else {
                /* compute_C3_immutable :213-235 (ns = 1 only) */
                for (size_t e = 0; e < dof; e++) {
                    double a = (eq.alpha.kind == 0) ? eq.alpha.lambda : eq.alpha.blk[e];
                    double b = (eq.beta.kind == 0) ? eq.beta.lambda : eq.beta.blk[e];
                    double c3 = a * (5.0 * FT(it - 1)[e] - 4.0 * FT(it - 2)[e] + FT(it - 3)[e]) / (dlt * dlt);
                    c3 += b * (2.0 / dlt * FT(it - 1)[e] - FT(it - 2)[e] / (2.0 * dlt));
                    c3 -= delta[e];
                    c3 += -KT(it - i2)[e] * FT(i2)[e] + KT(it - 1)[e] * FINT(1)[e] + KINT(1)[e] * FT(it - 1)[e];
                    for (int j = 2; j <= i2; j++) c3 += (KT(it - j)[e] - KT(it - j + 1)[e]) * FINT(j)[e];
                    for (int j = 2; j <= it - i2; j++) c3 += KINT(j)[e] * (FT(it - j)[e] - FT(it - j + 1)[e]);
                    C3[e] = c3;
                }
            }
            /* update_F! :326-348 with the stale K_temp[it] */
#define UPDATE_F() \
            for (int k = 0; k < Nk; k++) { \
                blk_mul(ns, KT(it) + k * nb, C2 + k * nb, blkA); \
                for (size_t e = 0; e < nb; e++) blkA[e] = -blkA[e] + C3[k * nb + e]; \
                if (ns == 1) FT(it)[k] = blkA[0] / C1[k]; \
                else blk_solve(ns, C1 + k * nb, blkA, FT(it) + k * nb); \
            }
            UPDATE_F();
            double error = INFINITY;
            long iterations = 0;
            while (error > tolerance) {                                /* :404 */
                iterations++;
                if (iterations > max_iterations) { status = 1; break; } /* :406-408 */
                ok_kernel_eval(kernel, FT(it), kernel_tmap ? kernel_tmap[nout + (it - n2 - 1)] : nout + (it - n2 - 1), KT(it));  /* update_K! :360-369 */
                UPDATE_F();
                error = find_error(FT(it), Fold, dof);                 /* :410 */
                memcpy(Fold, FT(it), dof * sizeof(double));            /* :412/:414 */
                kernel_evals++;                                        /* :416 */
                if (max_evals > 0 && kernel_evals >= max_evals) { status = 5; break; }   /* oracle-only: bounded CPU baseline */
            }
            if (status) break;
            /* update_integrals! :377-384 */
            for (size_t e = 0; e < dof; e++) {
                FINT(it)[e] = (FT(it)[e] + FT(it - 1)[e]) / 2.0;
                KINT(it)[e] = (KT(it)[e] + KT(it - 1)[e]) / 2.0;
            }
        }
        if (status) break;
        /* allocate_results! :429-438 */
        for (int it = n2 + 1; it <= n4; it++) {
            t_out[nout] = dlt * it;
            memcpy(F_out + dof * nout, FT(it), dof * sizeof(double));
            memcpy(K_out + dof * nout, KT(it), dof * sizeof(double));
            nout++;
        }
        /* new_time_mapping! :448-489 */
        for (int j = 1; j <= n2; j++)
            for (size_t e = 0; e < dof; e++) {
                FINT(j)[e] = (FINT(2 * j)[e] + FINT(2 * j - 1)[e]) / 2.0;
                KINT(j)[e] = (KINT(2 * j)[e] + KINT(2 * j - 1)[e]) / 2.0;
                FT(j)[e] = FT(2 * j)[e];
                KT(j)[e] = KT(2 * j)[e];
            }
        for (int j = n2 + 1; j <= n4; j++)
            for (size_t e = 0; e < dof; e++) { FINT(j)[e] = 0.0; KINT(j)[e] = 0.0; FT(j)[e] = 0.0; KT(j)[e] = 0.0; }
        Dt *= 2.0;
    }
    *kernel_evals_out = kernel_evals;
    free(Ft); free(Kt); free(FI); free(KI); free(C1); free(C2); free(C3); free(tv); free(Fold); free(Fold0); free(eq.K0);
    return status;
}

/* solve_steady_state_mutable -- src/SteadyStateMemoryEquation.jl:59-100 (diagonal/blocked K + gamma).
 * gamma: per-k blocks [Nk][ns*ns].  Returns 0 ok, 1 not converged. */
OK_API int ok_steady_state(ok_kernel *kernel, int Nk, int ns, const double *gamma, const double *F0, double tolerance,
                           long max_iterations, double *F_out, double *K_out, long *iters_out) {
    size_t nb = (size_t)ns * ns, dof = nb * Nk;
    double *K = dalloc(dof), *F = dalloc(dof), *Fold = ddup(F0, dof), *tv = dalloc(dof);
    ok_kernel_eval(kernel, F0, 0, K);                                  /* :66 */
    double err = tolerance * 2.0;
    long iterations = 0; int status = 0;
    while (err > tolerance) {
        iterations++;
        if (iterations > max_iterations) { status = 1; break; }
        for (int k = 0; k < Nk; k++) {
            double M[NSMAX * NSMAX];
            blk_mul(ns, K + k * nb, F0 + k * nb, tv + k * nb);         /* :81 */
            for (size_t e = 0; e < nb; e++) M[e] = K[k * nb + e] + gamma[k * nb + e];   /* :83 */
            if (ns == 1) F[k] = tv[k] / M[0];                          /* :84 */
            else blk_solve(ns, M, tv + k * nb, F + k * nb);
        }
        ok_kernel_eval(kernel, F, 0, K);                               /* :89 */
        err = find_error(F, Fold, dof);                                /* :90 */
        memcpy(Fold, F, dof * sizeof(double));
    }
    memcpy(F_out, F, dof * sizeof(double));
    memcpy(K_out, K, dof * sizeof(double));
    *iters_out = iterations;
    free(K); free(F); free(Fold); free(tv);
    return status;
}

/* find_relaxation_time -- src/RelaxationTime.jl:14-39.  mode: 0 = :log, 1 = :lin */
OK_API double ok_relaxation_time(long nt, const double *t, const double *F, double threshold, int mode) {
    for (long it = 0; it < nt; it++) {
        if (F[it] / F[0] > threshold) continue;
        if (it == 0) return 0.0;
        double y0 = F[it] / F[0], y1 = F[it - 1] / F[0], y = threshold;
        if (mode == 0) {
            double x0 = log(t[it]), x1 = log(t[it - 1]);
            return exp(((y - y0) * (x1 - x0) + x0 * (y1 - y0)) / (y1 - y0));
        }
        double x0 = t[it], x1 = t[it - 1];
        return ((y - y0) * (x1 - x0) + x0 * (y1 - y0)) / (y1 - y0);
    }
    return INFINITY;
}

/* Time the kernel evaluation alone (bench.py cpu_baseline): n evaluations, F perturbed per call so the
 * compiler cannot hoist anything.  Returns a checksum. */
OK_API double ok_kernel_eval_repeat(ok_kernel *K, const double *F, int n, double *out) {
    size_t dof = (size_t)K->Nk * K->ns * K->ns;
    double *Fw = ddup(F, dof), chk = 0.0;
    for (int i = 0; i < n; i++) {
        Fw[i % dof] *= 1.0000001;
        ok_kernel_eval(K, Fw, 0, out);
        chk += out[i % dof];
    }
    free(Fw);
    return chk;
}
