// msd.cu -- MSDModeCouplingKernel3D and the scalar memory equation it feeds.
//
// evaluate_kernel(kernel::MSDModeCouplingKernel3D, MSD, t) (src/Kernels/ModeCouplingKernels.jl:572-586) does not depend
// on the unknown: K(t_i) = dk rho kBT / (6 pi^2 m) * sum_q q^4 c_q^2 F(q,t_i) Fs(q,t_i), with F and Fs the trajectories
// of the collective and the tagged solve looked up at the same time (tDict).  Both trajectories are resident on the
// device after their solves, so the table is one pass over 2 nt Nk doubles (HBM-bound) and nothing crosses PCIe.
// The equation itself is scalar: `solve` takes the allocating branch for an immutable F0 (compute_C3_immutable,
// src/TimeDoublingSolver.jl:213-235; F_old restored from the cache every step, :400).  One thread integrates it.
#include "msd.cuh"

namespace mct {

namespace {

// One WARP per output time (8 times in flight per CTA, no block barrier): the lanes walk the wavenumbers with four
// independent partial sums (four pairs of loads in flight per lane), fixed-order combination, butterfly over the warp.
// (Round 1 used one block per time with a barrier per time: 2.6 TB/s at config-3 size, see profiles/README.md.)
__global__ void __launch_bounds__(kThreads) msd_table_kernel(const MsdParams P) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ns = P.ns, nb = ns * ns, sp = P.s;
    for (long it = (long)blockIdx.x * kWarps + w; it < P.nt; it += (long)gridDim.x * kWarps) {
        const double *Fs = P.Fs + (size_t)it * P.Nk;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        if (!P.Cmix) {
            const double *F = P.Fc + (size_t)it * P.Nk;
            for (int i0 = lane; i0 < P.Nk; i0 += 128) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int iq = i0 + 32 * j;
                    if (iq < P.Nk) {
                        const double k = P.kvec[iq], c = P.Ck[iq];
                        acc[j] += k * k * k * k * (c * c) * F[iq] * Fs[iq];       // :581
                    }
                }
            }
        } else {
            // MSDMultiComponentModeCouplingKernel3D (src/Kernels/MultiComponentKernels.jl:470-486):
            //   K += k^4 C[a,s] C[s,b] F[a,b] Fs  for a, b = 1..Ns,  with  Ns = length(Fs[1])  (:477).
            // Fs[1] is a Float64 there, so the reference's species sum only ever visits a = b = 1 and its golden
            // (test/test_MCMCT.jl:226) is that number: species_terms = 1 reproduces it, species_terms = ns is the documented sum.
            const double *F = P.Fc + (size_t)it * P.Nk * nb;
            for (int i0 = lane; i0 < P.Nk; i0 += 128) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int iq = i0 + 32 * j;
                    if (iq < P.Nk) {
                        const double k = P.kvec[iq], k4 = k * k * k * k;
                        const double *Fb = F + (size_t)iq * nb, *Cb = P.Cmix + (size_t)iq * nb;
                        for (int a = 0; a < P.species_terms; a++)
                            for (int b = 0; b < P.species_terms; b++) acc[j] += k4 * Cb[a + sp * ns] * Cb[sp + b * ns] * Fb[a + b * ns] * Fs[iq];
                    }
                }
            }
        }
        const double s = warp_sum((acc[0] + acc[1]) + (acc[2] + acc[3]));
        if (lane == 0) P.Ktab[it] = s * P.scale;                                 // :583
    }
}

// scalar time doubling with a tabulated kernel; rings in shared memory (slot s at index s-1)
__global__ void msd_scalar_td_kernel(const MsdParams P) {
    extern __shared__ double sm[];
    if (threadIdx.x != 0) return;
    const int N = P.N, n4 = 4 * N, n2 = 2 * N;
    double *Ft = sm, *Kt = sm + n4, *FI = sm + 2 * n4, *KI = sm + 3 * n4;
    double *Karr = sm + 4 * n4, *dFarr = Karr + (n2 + 1), *Farr = dFarr + (n2 + 1);
#define FT(s) Ft[(s) - 1]
#define KT(s) Kt[(s) - 1]
#define FINT(s) FI[(s) - 1]
#define KINT(s) KI[(s) - 1]
    const double a = P.alpha, b = P.beta, g = P.gamma, dlt0 = P.delta;
    const double K0 = P.Ktab[0];                                     // src/MemoryEquation.jl:45
    const double Fold0 = K0 * P.F0;                                  // :64
    for (int s = 1; s <= n4; s++) { FT(s) = Fold0; FINT(s) = Fold0; KT(s) = K0 + K0; KINT(s) = K0 + K0; }   // :82-85
    // Euler bootstrap  src/EulerSolver.jl:26-64, 92-112
    const double de = P.dt0 / (4.0 * N);
    Karr[0] = K0; dFarr[0] = P.dF0; Farr[0] = P.F0;
    for (int it = 1; it <= n2; it++) {
        double integ;
        if (it == 1) integ = Karr[0] * dFarr[0] * de;                // :34-36
        else {
            double r = Karr[0] * dFarr[it - 1] / 2.0;
            r += Karr[it - 1] * dFarr[0] / 2.0;
            for (int i = 2; i <= it - 1; i++) r += Karr[i - 1] * dFarr[it - i];
            integ = r * de;
        }
        const double Fo = Farr[it - 1], dFo = dFarr[it - 1];
        if (a != 0.0) {                                              // :55-58
            const double ddF = (b * dFo + g * Fo + dlt0 + integ) / (-a);
            dFarr[it] = dFo + de * ddF;
            Farr[it] = Fo + de * dFarr[it];
        } else {                                                     // :60-61
            const double dF = (g * Fo + dlt0 + integ) / (-b);
            dFarr[it] = dF;
            Farr[it] = Fo + de * dF;
        }
        Karr[it] = P.Ktab[it];
    }
    for (int it = 1; it <= n2; it++) { FT(it) = Farr[it]; KT(it) = Karr[it]; }       // :96-109, :155-166
    FINT(1) = (FT(1) + P.F0) / 2.0;                                  // initialize_integrals! :174-199
    KINT(1) = (3.0 * KT(1) - KT(2)) / 2.0;
    for (int it = 2; it <= n2; it++) { FINT(it) = (FT(it) + FT(it - 1)) / 2.0; KINT(it) = (KT(it) + KT(it - 1)) / 2.0; }
    long nout = 0;
    P.t_out[0] = 0.0; P.F_out[0] = P.F0; P.K_out[0] = K0; nout = 1;
    for (int it = 1; it <= n2; it++, nout++) { P.t_out[nout] = de * it; P.F_out[nout] = FT(it); P.K_out[nout] = KT(it); }
    long long evals = 0;
    int status = MCT_OK;
    double Dt = P.dt0;
    for (int blk = 0; blk < P.nblocks && status == MCT_OK; blk++, Dt *= 2.0) {
        const double dl = Dt / (4.0 * N);
        for (int it = n2 + 1; it <= n4 && status == MCT_OK; it++) {
            const int i2 = it / 2;
            double Fold = Fold0;                                     // immutable branch: F_old from the cache  :400
            const double c1 = ((2.0 / (dl * dl) * a + 3.0 / (2.0 * dl) * b) + KINT(1)) + g;     // :310
            const double c2 = FINT(1) - P.F0;                        // :311
            double c3 = a * (5.0 * FT(it - 1) - 4.0 * FT(it - 2) + FT(it - 3)) / (dl * dl);    // compute_C3_immutable :213-235
            c3 += b * (2.0 / dl * FT(it - 1) - FT(it - 2) / (2.0 * dl));
            c3 -= dlt0;
            c3 += -KT(it - i2) * FT(i2) + KT(it - 1) * FINT(1) + KINT(1) * FT(it - 1);
            for (int j = 2; j <= i2; j++) c3 += (KT(it - j) - KT(it - j + 1)) * FINT(j);
            for (int j = 2; j <= it - i2; j++) c3 += KINT(j) * (FT(it - j) - FT(it - j + 1));
            FT(it) = (-KT(it) * c2 + c3) / c1;                       // update_F! with the stale K_temp[it]  :330-331
            double err = INFINITY;
            long long iter = 0;
            while (err > P.tol) {                                    // :404
                iter++;
                if (iter > P.max_iter) { status = MCT_NOT_CONVERGED; break; }
                KT(it) = P.Ktab[nout + (it - n2 - 1)];               // update_K!: the table at this step's time  :364-365
                FT(it) = (-KT(it) * c2 + c3) / c1;
                err = fabs(FT(it) - Fold);                           // :410
                Fold = FT(it);
                evals++;                                             // :416
            }
            if (status != MCT_OK) break;
            FINT(it) = (FT(it) + FT(it - 1)) / 2.0;                  // update_integrals! :377-384
            KINT(it) = (KT(it) + KT(it - 1)) / 2.0;
        }
        if (status != MCT_OK) break;
        for (int it = n2 + 1; it <= n4; it++, nout++) { P.t_out[nout] = dl * it; P.F_out[nout] = FT(it); P.K_out[nout] = KT(it); }   // :429-438
        for (int j = 1; j <= n2; j++) {                              // new_time_mapping! :448-489
            FINT(j) = (FINT(2 * j) + FINT(2 * j - 1)) / 2.0;
            KINT(j) = (KINT(2 * j) + KINT(2 * j - 1)) / 2.0;
            FT(j) = FT(2 * j);
            KT(j) = KT(2 * j);
        }
        for (int j = n2 + 1; j <= n4; j++) { FINT(j) = 0.0; KINT(j) = 0.0; FT(j) = 0.0; KT(j) = 0.0; }
    }
    *P.status = status;
    *P.kernel_evals = evals;
#undef FT
#undef KT
#undef FINT
#undef KINT
}

__global__ void mc_species_weight_kernel(const double *F, const double *C, int Nk, int ns, int s, long nt, double *G) {
    const int nb = ns * ns;
    const size_t n = (size_t)nt * Nk;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % Nk);
        const double *Fb = F + i * nb, *Cb = C + (size_t)q * nb;     // column-major ns x ns blocks
        double g = 0.0;
        for (int a = 0; a < ns; a++)
            for (int b = 0; b < ns; b++) g += Fb[a + b * ns] * Cb[s + a * ns] * Cb[s + b * ns];      // :352-357
        G[i] = g;
    }
}

__global__ void gather_slice_kernel(const double *src, int dof, const long *times, long n_times, const int *dofs, int n_dofs,
                                    double *dst) {
    const size_t n = (size_t)n_times * n_dofs;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const long ti = (long)(i / n_dofs);
        const int dj = (int)(i - (size_t)ti * n_dofs);
        dst[i] = src[(size_t)(times ? times[ti] : ti) * dof + (dofs ? dofs[dj] : dj)];
    }
}

}  // namespace

int gather_slice_launch(const double *src, int dof, const long *d_times, long n_times, const int *d_dofs, int n_dofs,
                        double *dst, cudaStream_t stream) {
    const size_t n = (size_t)n_times * n_dofs;
    const int blocks = (int)std::min<size_t>((n + kThreads - 1) / kThreads, 148 * 8);
    gather_slice_kernel<<<std::max(blocks, 1), kThreads, 0, stream>>>(src, dof, d_times, n_times, d_dofs, n_dofs, dst);
    count_launch();
    MCT_CUDA(cudaGetLastError());
    return MCT_OK;
}

int mc_species_weight_launch(const double *F_traj, const double *Cblocks, int Nk, int ns, int s, long nt, double *G,
                             cudaStream_t stream) {
    mc_species_weight_kernel<<<148 * 8, kThreads, 0, stream>>>(F_traj, Cblocks, Nk, ns, s, nt, G);
    count_launch();
    MCT_CUDA(cudaGetLastError());
    return MCT_OK;
}

int msd3d_launch(const MsdParams &P, cudaStream_t stream) {
    const int blocks = (int)std::min<long>((P.nt + kWarps - 1) / kWarps, 148 * 8);      // one warp per output time
    msd_table_kernel<<<blocks, kThreads, 0, stream>>>(P);
    count_launch();
    MCT_CUDA(cudaGetLastError());
    const size_t smem = ((size_t)16 * P.N + 3 * (2 * P.N + 1)) * sizeof(double);
    MCT_REQUIRE(smem <= 200 * 1024, MCT_UNSUPPORTED, "N too large for the scalar device solver");
    MCT_CUDA(cudaFuncSetAttribute((const void *)msd_scalar_td_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    msd_scalar_td_kernel<<<1, 32, smem, stream>>>(P);
    count_launch();
    MCT_CUDA(cudaGetLastError());
    return MCT_OK;
}

}  // namespace mct
