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
    // multi-component MSD (Cmix != nullptr): Fc is the mixture trajectory [nt][Nk][ns*ns], Cmix the C(k) blocks; the species
    // sum runs over the first `species_terms` species only (1 reproduces the reference, see msd.cu)
    const double *Cmix; int ns, s, species_terms;
    const double *Fc, *Fs;        // device trajectories [nt][Nk] of the collective and the tagged solve
    double *Ktab;                 // [nt] kernel table
    double *t_out, *F_out, *K_out;   // [nt]
    int *status;
    long long *kernel_evals;
};

int msd3d_launch(const MsdParams &P, cudaStream_t stream);

// G[it][q] = sum_{a,b} C_q[s,a] C_q[s,b] F_it,q[a,b]: the collective weight of species s that both
// TaggedMultiComponentModeCouplingKernel3D (fill_A!, src/Kernels/MultiComponentKernels.jl:335-370) and
// MSDMultiComponentModeCouplingKernel3D (:470-486) contract the mixture trajectory with.
int mc_species_weight_launch(const double *F_traj, const double *Cblocks, int Nk, int ns, int s, long nt, double *G,
                             cudaStream_t stream);

// dst[i][j] = src[times[i]][dofs[j]] of a device trajectory [nt][dof]; NULL index lists mean "all"  (get_F(sol, its, iks),
// src/HelperFunctions.jl:169-210 -- SURVEY 8f-3: only the requested slice crosses PCIe)
int gather_slice_launch(const double *src, int dof, const long *d_times, long n_times, const int *d_dofs, int n_dofs,
                        double *dst, cudaStream_t stream);

}  // namespace mct
