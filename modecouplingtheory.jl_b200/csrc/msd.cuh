// msd.cuh -- mean-squared-displacement step of the collective -> tagged -> MSD chain.
#pragma once
#include "common.cuh"

namespace mct {

struct MsdParams {
    int Nk, N, nblocks;
    long nt;
    double dt0, tol;
    long long max_iter;
    double alpha, beta, gamma, delta, F0, dF0;
    double scale;                 // dk * rho * kBT / (6 pi^2 m)
    const double *kvec, *Ck;      // [Nk]
    const double *Fc, *Fs;        // device trajectories [nt][Nk] of the collective and the tagged solve
    double *Ktab;                 // [nt] kernel table
    double *t_out, *F_out, *K_out;   // [nt]
    int *status;
    long long *kernel_evals;
};

int msd3d_launch(const MsdParams &P, cudaStream_t stream);

// dst[i][j] = src[times[i]][dofs[j]] of a device trajectory [nt][dof]; NULL index lists mean "all"  (get_F(sol, its, iks),
// src/HelperFunctions.jl:169-210 -- SURVEY 8f-3: only the requested slice crosses PCIe)
int gather_slice_launch(const double *src, int dof, const long *d_times, long n_times, const int *d_dofs, int n_dofs,
                        double *dst, cudaStream_t stream);

}  // namespace mct
