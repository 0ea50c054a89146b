// layout.cuh -- "team layout" of the Bengtzelius vertex tables.
//
// The reference evaluates  T_m[ik] = T_m[ik-1] + sum_{iq<=Nk-ik+1} (A_m[iq,iq+ik-1] + A_m[iq+ik-1,iq])
//                                            - sum_{iq<ik} A_m[iq,ik-iq],     A_m[iq,ip] = V_m[iq,ip] F1[iq] F2[ip]
// (src/Kernels/ModeCouplingKernels.jl:154-208) by walking diagonals and antidiagonals of three Nk x Nk
// column-major matrices with stride Nk+-1.  Here every "row" ik owns the terms of its increment
// inc_m[ik] = T_m[ik]-T_m[ik-1].  A term is (a, b, coef): it contributes coef * V_m[a,b] * F1[a] * F2[b]; the
// coefficient (+1, +2, -1, -2; 2 where the symmetric form merges A[q,p]+A[p,q]) is folded into the stored table
// value.  The terms of a row are padded to a multiple of 32 ("groups": one term per lane of a warp), rows are
// dealt to the G CTAs of a team as contiguous, cost-balanced ranges, and the groups of a CTA are split evenly
// over its 16 warps: warp w of CTA c streams groups [w*T, (w+1)*T) of the CTA's group sequence.  The slab of a
// CTA (16*T*32 term slots per table) is what the CTA keeps resident in registers + shared memory for the whole
// solve; every thread owns the fixed slots  (w*T + j)*32 + lane,  j < T.
#pragma once
#include "common.cuh"

namespace mct {

// packed per-slot descriptor (one int per thread and group)
constexpr int kDescIdxBits = 11;                 // Nk <= 2048
constexpr int kDescIdxMask = (1 << kDescIdxBits) - 1;
constexpr int kDescCoefShift = 22;               // 2 bits: 0:+1 1:+2 2:-1 3:-2
constexpr int kDescValid = 1 << 24;              // slot holds a term (otherwise the table value is 0)
constexpr int kDescNewEntry = 1 << 25;           // warp-uniform: this group starts a new (warp,row) entry

struct TeamLayout {
    int Nk = 0, G = 0, sym = 0;
    int T = 0;             // groups per warp
    int rows_max = 0;      // max rows owned by one CTA
    int ent_max = 0;       // max (warp,row) entries of one CTA
    int went_max = 0;      // max entries of one warp
    long long slab_terms = 0;   // G*16*T*32: stride between the three tables
    // device copies
    int *d_desc = nullptr;       // [G][16][T][32]
    int *d_cta_meta = nullptr;   // [G][4]: row0, nrows, n_entries, 0
    int *d_ent_begin = nullptr;  // [G][rows_max+1]: entry range of each owned row
    int *d_warp_ent = nullptr;   // [G][16]: first entry of each warp
    int *d_row_owner = nullptr;  // [Nk]: owner CTA of each row
    int xbuf_words = 0;          // words (doubles) of one exchange buffer
};

// Geometry only (no table values): depends on (Nk, G, sym).
int layout_build_geometry(TeamLayout &L, int Nk, int G, int sym);
void layout_free(TeamLayout &L);
// groups per warp the geometry (Nk, G, sym) would have, without building it
int layout_groups_per_warp(int Nk, int G, int sym, int *rows_max);
// Fill d_tab (3*slab_terms doubles) from column-major device tables.
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
