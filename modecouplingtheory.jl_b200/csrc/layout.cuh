// layout.cuh -- "team layout" of the Bengtzelius vertex tables.
//
// The reference evaluates  T_m[ik] = T_m[ik-1] + sum_{iq<=Nk-ik+1} (A_m[iq,iq+ik-1] + A_m[iq+ik-1,iq])
//                                            - sum_{iq<ik} A_m[iq,ik-iq],     A_m[iq,ip] = V_m[iq,ip] F1[iq] F2[ip]
// (src/Kernels/ModeCouplingKernels.jl:154-208) by walking diagonals and antidiagonals of three Nk x Nk
// column-major matrices with stride Nk+-1.  Here every "row" ik owns the terms of its increment
// inc_m[ik] = T_m[ik]-T_m[ik-1]; the terms of a row are stored CONTIGUOUSLY, pre-multiplied by their sign
// (and by 2 where the symmetric form merges A[q,p]+A[p,q]), so that a warp streams them with unit stride.
// Rows are dealt to the G CTAs of a team in snake order (cost of row ik falls linearly with ik), each CTA's
// slab is what it keeps resident in shared memory for the whole solve.
#pragma once
#include "common.cuh"

namespace mct {

struct Task {      // a run of consecutive terms of one row, executed by one warp (32 bytes)
    int a0;        // index into F1 of the first term (0-based)
    int b0;        // index into F2 of the first term (0-based)
    int bdir;      // F2 index step per term: +1 (diagonals) or -1 (antidiagonals); F1 always steps +1
    int len;       // number of terms
    int off;       // offset of the first term inside the CTA slab
    int slot;      // partial-sum slot inside the CTA (shared by consecutive tasks of one row in one warp)
    int where;     // 0: slab part resident in shared memory, 1: part spilled to global (L2)
    int flush;     // 1: last task of its slot -> reduce across the warp and store the partial sum
};

struct Piece {     // a run of terms of one row, used to build the slab from the Nk x Nk tables
    int a0, b0, bdir, len;
    long long off;  // global term offset inside the layout
    double coef;    // +1, +2, -1, -2
};

struct TeamLayout {
    int Nk = 0, G = 0, sym = 0;
    int smem_terms = 0;    // per-CTA capacity (terms) of the shared-memory resident part
    int rows_max = 0;      // max rows per CTA
    int slots_max = 0;     // max partial-sum slots per CTA
    int tasks_max = 0;     // max tasks of one warp
    long long total_terms = 0;
    std::vector<long long> cta_off;  // [G+1]
    // device copies
    double *d_tab = nullptr;      // [3][total_terms]
    Task *d_tasks = nullptr;      // all tasks, grouped by (cta, warp)
    int *d_task_begin = nullptr;  // [G*kWarps + 1]
    int *d_rows = nullptr;        // [G][rows_max]   row index ik (0-based) or -1
    int *d_row_slot = nullptr;    // [G][rows_max+1] slot range of each local row
    long long *d_cta_off = nullptr;  // [G+1]
    int *d_kslot = nullptr;          // [Nk] word offset of wavenumber k's slot in an exchange buffer
    int xbuf_words = 0;              // words (doubles) of one exchange buffer
    std::vector<Piece> pieces;    // host copy (kept for building more slabs with the same geometry)
    Piece *d_pieces = nullptr;
    int n_pieces = 0;
};

// Geometry only (no table values): depends on (Nk, G, sym, smem_terms).
int layout_build_geometry(TeamLayout &L, int Nk, int G, int sym, int smem_terms);
void layout_free(TeamLayout &L);
// Fill L.d_tab (must be allocated, 3*total_terms doubles) from column-major device tables.
int layout_fill_from_tables(const TeamLayout &L, double *d_tab, const double *dV1, const double *dV2, const double *dV3,
                            cudaStream_t s);
// Fill from S(k) directly, evaluating the reference constructor's expressions in their own order
// (src/Kernels/ModeCouplingKernels.jl:105-126 collective, :402-422 tagged).
int layout_fill_from_sk(const TeamLayout &L, double *d_tab, const double *dk_array, const double *dCk, double rho, double kBT,
                        double m, int tagged, cudaStream_t s);

// Raw Nk x Nk tables from S(k) on the device (same expressions), used by the streaming evaluation kernel.
int tables_from_sk(int Nk, const double *dk_array, const double *dSk, double *dCk, double rho, double kBT, double m, int tagged,
                   double *dV1, double *dV2, double *dV3, cudaStream_t s);

}  // namespace mct
