// layout.cuh -- "team layout" of the Bengtzelius vertex tables.
//
// The reference evaluates  T_m[ik] = T_m[ik-1] + sum_{iq<=Nk-ik+1} (A_m[iq,iq+ik-1] + A_m[iq+ik-1,iq])
//                                            - sum_{iq<ik} A_m[iq,ik-iq],     A_m[iq,ip] = V_m[iq,ip] F1[iq] F2[ip]
// (src/Kernels/ModeCouplingKernels.jl:154-208) by walking diagonals and antidiagonals of three Nk x Nk
// column-major matrices with stride Nk+-1.  Here every "row" (0-based d = ik-1) owns the terms of its increment
// inc_m[d] = T_m[ik]-T_m[ik-1]: a run along the diagonal d, (non-symmetric tables only) a run along the lower diagonal,
// and a run along the antidiagonal.  A term contributes coef * V_m[a,b] * F1[a] * F2[b]; the coefficient (+1, +2, -1,
// -2; 2 where the symmetric form merges A[q,p]+A[p,q]) is folded into the stored table value.
//
// Register tiling: rows are grouped in ROW BLOCKS of 4 consecutive rows.  For a fixed position a along a run the four
// rows of a block need the SAME single operand S[a] and four CONSECUTIVE window operands W[w(a)+r], r = 0..3:
//     diagonal      (a, a+d0+r):      S = F1, W = F2, w(a) = d0 + a
//     antidiagonal  (a, d0-1-a+r):    S = F1, W = F2, w(a) = d0 - 1 - a
//     lower diag.   (a+d0+r, a):      S = F2, W = F1, w(a) = d0 + a          (non-symmetric tables)
// A UNIT is 32 consecutive positions a of one run of one row block: every lane loads 1 + 4 operands and does 4
// products and 12 FMAs with 12 table values it owns.  The units of a CTA (its rows form a contiguous, cost-balanced
// range) are dealt to its 8 warps in order; thread (warp w, lane l) owns the fixed table values
// tab[((c*8 + w)*U + u)*12 + r*3 + m][l], which it keeps in registers / shared memory for the whole solve.
#pragma once
#include "common.cuh"

namespace mct {

constexpr int kRB = 4;                           // rows per row block
constexpr int kUV = 3 * kRB;                     // table values per lane and unit
// packed unit descriptor (warp-uniform)
constexpr int kUA0Bits = 11;                     // a0: first position of the unit (Nk <= 2048)
constexpr int kUW0Shift = 11, kUW0Bits = 12;     // window start of lane 0, biased by kFPad (can be slightly negative)
constexpr int kUNeg = 1 << 23;                   // window index decreases with the lane (antidiagonal)
constexpr int kUSwap = 1 << 24;                  // single operand from F2, window from F1
constexpr int kUNewEntry = 1 << 25;              // first unit of a new (warp,row block) entry
constexpr int kUValid = 1 << 26;
constexpr int kURowsShift = 27;                  // 2 bits: (rows of the block that exist) - 1; the arithmetic of the others is skipped
constexpr int kFPad = 40;                        // zero padding (doubles) in front of and behind the F arrays in shared memory

struct UnitInfo {        // what the table-fill kernels need to know about a unit
    int type;            // 0 diagonal, 1 antidiagonal, 2 lower diagonal
    int d0;              // first row of the row block
    int dend;            // one past the last row of this CTA (rows >= dend are padding)
    int a0;
};

struct TeamLayout {
    int Nk = 0, G = 0, sym = 0;
    int U = 0;             // units per warp
    int rows_max = 0;      // max rows owned by one CTA
    int rb_max = 0;        // max row blocks of one CTA
    int ent_max = 0;       // max (warp,row block) entries of one CTA
    int went_max = 0;      // max entries of one warp
    long long n_units = 0; // G*8*U
    // device copies
    int *d_udesc = nullptr;      // [G][8][U]
    UnitInfo *d_uinfo = nullptr; // [G][8][U]
    int *d_cta_meta = nullptr;   // [G][4]: row0, nrows, n_entries, 0
    int *d_ent_begin = nullptr;  // [G][rb_max+1]: entry range of each row block
    int *d_warp_ent = nullptr;   // [G][9]: first entry of each warp
    int *d_row_owner = nullptr;  // [Nk]: owner CTA of each row
    int xbuf_words = 0;          // words (doubles) of one exchange buffer
    int xrows = 0;               // offset of the rows inside an exchange buffer (the block totals come first)
    long long slab_doubles() const { return n_units * kUV * 32; }
};

// Geometry only (no table values): depends on (Nk, G, sym).
int layout_build_geometry(TeamLayout &L, int Nk, int G, int sym);
void layout_free(TeamLayout &L);
// total number of units of the problem (for choosing the team size)
long long layout_total_units(int Nk, int sym);
// Fill d_tab (slab_doubles()) from column-major device tables.
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
