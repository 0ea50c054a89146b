// td_sc.cuh -- persistent time-doubling solver for the single-component (Ns = 1) mode-coupling kernels.
#pragma once
#include "common.cuh"
#include "layout.cuh"

namespace mct {

// One memory equation resident on the device (all arrays length Nk unless noted).
struct EqDev {
    const double *tab;      // team-layout vertex tables [3][total_terms]
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

struct SolveParams {
    int Nk, G, N, nblocks, n_eq, nteams, second_order;
    double dt0, tol;
    long long max_iter;
    // layout geometry shared by every equation of the launch
    long long total_terms;
    const long long *cta_off;
    const Task *tasks;
    const int *task_begin;
    const int *rows;
    const int *row_slot;
    const int *kslot;           // [Nk] exchange-slot word of every wavenumber
    int xbuf_words;             // words of one exchange buffer
    int rows_max, slots_max, tasks_max, smem_terms;
    // team communication
    double *xbuf;               // [nteams][xstride]: four rotating exchange buffers of 3*Nk self-validating slots
    long long xstride;
    int *queue;                 // next equation to solve
    int *abort_flag;
    const EqDev *eqs;
    long long spin_limit_cycles;
    long long *prof;            // [nteams][10] phase cycle counters (only filled when built with -DMCT_PROFILE)
};

size_t td_sc_smem_bytes(int Nk, int G, int slots_max, int tasks_max, int rows_max, int smem_terms);
int td_sc_max_smem_terms(int Nk, int G, int slots_max_guess, int tasks_max_guess, int rows_max, size_t smem_limit);
int td_sc_launch(const SolveParams &P, cudaStream_t stream);

}  // namespace mct
