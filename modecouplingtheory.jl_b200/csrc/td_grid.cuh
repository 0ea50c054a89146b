// td_grid.cuh -- whole-grid cooperative solver for the kernels whose degrees of freedom are Ns x Ns blocks per
// wavenumber (MultiComponentModeCouplingKernel3D) or whose evaluation is a dense triple sum (dDimModeCouplingKernel).
#pragma once
#include <vector>

#include "common.cuh"

namespace mct {

constexpr int kNsMax = 4;
constexpr int kMaxShardRanks = 8;

enum GridKernelType { GK_MC3D = 3, GK_DDIM = 5 };

struct GridParams {
    int type, Nk, ns, N, nblocks, second_order, mode;   // mode 0: time doubling, 1: steady state, 2: one evaluation
    double dt0, tol;
    long long max_iter;
    // kernel data
    const double *kvec;
    const double *C, *pref;      // MC: direct correlation blocks [Nk][nb], prefactor [nb]
    const double *W;             // dDim: prefactor * J .* V as [iq][ik][ip] (ip contiguous)
    // equation: multiplicative coefficients as per-k blocks [Nk][nb]; delta, F0, dF0 [Nk][nb]
    const double *alpha, *beta, *gamma, *delta, *F0, *dF0;
    // workspace
    double *Ft, *Kt, *FI, *KI;   // rings [4N][dof]  (slot s at index s-1)
    double *C1, *C2, *C3, *Fold, *K0, *Fcur, *Kcur;   // [dof]
    double *G;                   // MC: F, G1, G2, G3 species-major [4][nb][Nk]
    double *Inc, *Tm;            // MC: increments and prefix sums [3*nb][Nk]
    double *Lsum;                // task sums [3*nb sequences][tasks][7]  (grid_lsum_words)
    double *dFh;                 // bootstrap: dF history [2N+1][dof]
    unsigned long long *err;     // [2] rotating max-error words (bits of a non-negative double)
    double *F_out, *K_out;       // trajectory [nt][dof]  (mode 1/2: [dof])
    const double *Fin;           // mode 2: input F
    // tagged dense kernels (TaggedActiveMCTKernel): collective trajectory [nt_c][Nk]; the evaluation at output time i is
    // bilinear, sum W F[ik] Fc_i[ip].  tix0: time index of a mode-2 evaluation
    const double *Fc; long nt_c, tix0;
    // k-sharding over GPUs (one process per GPU): rank r computes the line sums of lines L with L % world == r and stores
    // them straight into every rank's exchange buffer (peer memory over NVLink, IPC-mapped); 4 rotating buffers of
    // self-validating words, each rank re-arms its own copy.  world == 1: single GPU, Lsum only.
    int rank, world;
    double *xpeer[8];            // [world] exchange buffers [4][grid_lsum_words] of every rank (own included)
    long long spin_limit_cycles;
    int *status;
    long long *kernel_evals;
    unsigned *xepoch_state;      // exchange epoch carried from one solve of a sharded equation to the next (the rotating
                                 // buffers of all ranks stay in step without a host-side re-arm)
    unsigned long long *prof;    // [12] cycles per phase seen by thread 0 (diagnostics, printed with MCT_GRID_PROF=1)
    // line-sum tasks dealt by the host (grid_build_deal): list of dealing-order slots of warp wl of species pair e at
    // deal[(e * deal_wsmax + wl) * deal_stride + j], terminated by -1; valid for a grid of deal_grid blocks of kGridThreads
    // threads (any other launch shape, or deal == nullptr, falls back to the closed-form snake deal)
    const int *deal; int deal_grid, deal_wsmax, deal_stride;
};
constexpr int kGridThreads = 256;

// Multi-component line sums are produced per tile task (td_grid.cu, phase_lines): 4-wavenumber chunks, one task per tile
// diagonal J - I >= 0 (tasks [0, chunks)) and per tile antidiagonal that holds a line s <= Nk-2 (tasks chunks + S),
// 7 lines x 3 vertices x Ns^2 words each.
__host__ __device__ inline int grid_chunks(int Nk) { return (Nk + 3) / 4; }
__host__ __device__ inline int grid_smem_stride(int Nk) { return (4 * grid_chunks(Nk) + 15) / 16 * 16; }   // doubles per staged vector
__host__ __device__ inline int grid_anti_tasks(int Nk) { return Nk >= 2 ? (Nk - 2) / 4 + 1 : 0; }
__host__ __device__ inline size_t grid_lsum_words(int Nk, int ns) {
    return (size_t)(grid_chunks(Nk) + grid_anti_tasks(Nk)) * 7 * 3 * ns * ns;
}

// Host side of the deal: longest-processing-time-first assignment of this rank's task pairs to the SM partitions of every
// species pair's blocks, then to the lighter of the partition's two warps.  Returns the table (see GridParams::deal).
void grid_build_deal(int Nk, int ns, int rank, int world, int grid, std::vector<int> &table, int &wsmax, int &stride);
size_t grid_deal_capacity(int Nk, int ns, int grid);     // ints to reserve (any rank / world)

int td_grid_launch(const GridParams &P, int n_sm, cudaStream_t stream);

// dDim: fused table W[iq][ik][ip] = prefactor * J[iq,ik,ip] * V[iq,ik,ip] from the reference's column-major arrays
int ddim_build_table(int Nk, double prefactor, const double *dV, const double *dJ, double *dW, cudaStream_t s);
// the same without the transposition, for tables stored [l, j, i] with the OUTPUT index last (ActiveMCTKernel)
int ddim_fuse_table(int Nk, double prefactor, const double *dV, const double *dJ, double *dW, cudaStream_t s);

}  // namespace mct
