// td_sc.cuh -- persistent time-doubling solver for the single-component (Ns = 1) mode-coupling kernels.
#pragma once
#include "common.cuh"
#include "layout.cuh"

namespace mct {

// One memory equation resident on the device (all arrays length Nk unless noted).
struct EqDev {
    const double *tab;      // team-layout vertex tables [3][slab_terms]
    const double *kvec;     // wavenumbers
    const double *alpha, *beta, *gamma, *delta;  // coefficients expanded to per-k values
    const double *F0, *dF0;
    double *ring;           // history rings F_temp, K_temp, F_I, K_I: [4][Nk][4N]  (src/TimeDoublingSolver.jl:3-18)
    double *F_out, *K_out;  // trajectory [nt][Nk]
    const double *Fc_traj;  // tagged kernels: collective trajectory [nt][Nk]; nullptr otherwise
    int mode;               // 0: time-doubling solve, 1: steady state (solve_steady_state)
    int *status;            // MCT_* code
    long long *kernel_evals;
};

// byte offsets of the kernel's arrays inside its dynamic shared memory (filled by td_sc_launch; kept in the
// parameter bank so that they cost no registers)
struct SmemOffsets {
    unsigned F1, F2, rho, off, wt, own, part, lane, C, X, scan, misc, eq, xd, desc, tab, ring, total;
};

struct SolveParams {
    int Nk, G, N, nblocks, n_eq, nteams, second_order;
    double dt0, tol;
    long long max_iter;
    // layout geometry shared by every equation of the launch (layout.cuh)
    int U;                      // units per warp (a unit = 32 positions x 4 rows, layout.cuh)
    int UR;                     // units per warp resident in registers (selects the kernel instantiation)
    int UT;                     // units per warp resident in tensor memory (24 of the warp's 256 columns each, <= 10)
    int US;                     // units per warp resident in shared memory; the remaining U-UR-UT-US stream from L2
    int rows_max, rb_max, ent_max, lane_cap;
    int ring_smem;              // 1: the history rings of the owned rows live in shared memory ([4][rows_max][4N+1]), 0: in global memory
    const int *udesc, *cta_meta, *ent_begin, *warp_ent, *row_owner;
    // team communication
    double *xbuf;               // [nteams][4][xbuf_words]: rotating exchange buffers of self-validating 8-byte slots
    int xbuf_words;
    int xrows;                  // offset of the rows inside an exchange buffer (the block totals come first)
    int fabric;                 // 0: exchange through L2 (any team size); 1: the team is a thread-block cluster (G <= 16), DSMEM
    long long xstride;          // doubles per team: 4*xbuf_words + 16 (arrival counter)
    int *queue;                 // next equation to solve
    double *qbox;               // [nteams][4*16]: rotating mailbox for the equation index
    int *abort_flag;
    const EqDev *eqs;
    long long spin_limit_cycles;
    SmemOffsets so;             // filled by td_sc_launch
    int dbg;                    // debug switches of profile builds (MCT_DBG)
    int snap_epoch;             // profile builds: exchange whose timeline is recorded (MCT_SNAP_EPOCH)
    long long *prof;            // [G][16] phase cycle counters of team 0 (only filled when built with -DMCT_PROFILE)
};

constexpr int kProfSlots = 32;
constexpr int kLaneStride = 33;                                    // padded column of the warp scratch: 32 lane partials
enum { OWN_K = 0, OWN_F0, OWN_AL, OWN_BE, OWN_GA, OWN_DE, OWN_K0, kOwnFields };   // per-row constants of the owner, in shared memory
__host__ __device__ inline int even_up(int n) { return (n + 1) & ~1; }

// shared memory needed by a launch with these parameters; US = 0 gives the fixed part
size_t td_sc_smem_bytes(const SolveParams &P);
int td_sc_ring_doubles(const SolveParams &P);
// largest UR worth using at this Nk (24 registers per register-resident unit; the instantiations that finish 4 or 8 rows
// per thread have fewer registers to spare)
int td_sc_max_reg_units(int Nk);
// units of a warp that fit its share of tensor memory (512 columns per lane quarter, shared by the warps w, w+4, ...)
constexpr int kTmemUnitsPerWarp = (512 / ((kWarps + 3) / 4)) / (2 * kUV);
int td_sc_launch(const SolveParams &P, cudaStream_t stream);
int td_sc_max_clusters(const SolveParams &P);

}  // namespace mct
