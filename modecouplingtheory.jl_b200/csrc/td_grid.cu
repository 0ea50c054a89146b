// td_grid.cu -- whole-grid cooperative time-doubling / steady-state solver for block-valued and dense kernels.
//
// MultiComponentModeCouplingKernel3D (src/Kernels/MultiComponentKernels.jl:84-219) and dDimModeCouplingKernel
// (src/Kernels/ModeCouplingKernels.jl:19-71, 237-256) behind the same solver loop as td_sc.cu
// (src/TimeDoublingSolver.jl:512-532), with every phase spread over the whole grid and a grid-wide barrier between
// phases.  The loop structure, the Euler bootstrap (src/EulerSolver.jl:26-64) and all per-k block algebra follow the
// reference's operation order (K*F products on the left, C1 \ rhs per wavenumber -- src/HelperFunctions.jl:11-36,
// src/TimeDoublingSolver.jl:326-337); the file is compiled with -fmad=false.
//
// Multi-component evaluation: the reference's fill_A! costs O(Ns^4 Nk^2) (:134-198).  Its four (mu', nu') sums
// factorise into Hadamard products of per-k matrices  G1 = C'FC, G2 = C'F, G3 = FC:
//     sum(term1) = F_q .* G1_p,   sum(term2) = G3_q .* G2_p,   sum(term3) = G2_q .* G3_p,   sum(term4) = G1_q .* F_p
// so an entry A_m[a,b,iq,ip] costs O(1) per (a,b) and A is never stored: a warp owns one diagonal / antidiagonal of
// the (iq, ip) plane and accumulates its 3 Ns^2 line sums; bengtzelius3! (:2-43) then is a prefix over the lines.
// Per evaluation the grid passes three cg grid barriers (2.1 us each at 148 CTAs -- profiles/gridsync_bench_r02.txt).  A
// hand-rolled arrival-counter barrier measures 1.7 us alone but changed nothing inside the solver (81.4 against 81.0 us per
// config-4 evaluation: the wait is the arrival skew of the blocks, not the barrier), so cg's is kept.
#include <cooperative_groups.h>

#include <algorithm>
#include <utility>

#include "td_grid.cuh"

namespace cg = cooperative_groups;

namespace mct {

namespace {

constexpr int kGT = kGridThreads;   // threads per block

// C = A*B, column-major NS x NS, plain left-to-right sums (StaticArrays `*`); everything stays in registers
template <int NS>
__device__ __forceinline__ void blk_mul(const double *A, const double *B, double *C) {
    double tmp[NS * NS];
#pragma unroll
    for (int j = 0; j < NS; j++)
#pragma unroll
        for (int i = 0; i < NS; i++) {
            double s = A[i] * B[j * NS];
#pragma unroll
            for (int l = 1; l < NS; l++) s += A[i + l * NS] * B[l + j * NS];
            tmp[i + j * NS] = s;
        }
#pragma unroll
    for (int e = 0; e < NS * NS; e++) C[e] = tmp[e];
}
// x = A \ b for NR right-hand sides (columns of b), Gaussian elimination with partial pivoting.  The pivot row is
// selected with predicated swaps so that every index is a compile-time constant (no local memory).
template <int NS, int NR>
__device__ __forceinline__ void blk_solve(const double *A, const double *B, double *X) {
    if (NS == 1) {
#pragma unroll
        for (int j = 0; j < NR; j++) X[j] = B[j] / A[0];
        return;
    }
    double a[NS * NS], b[NS * NR];
#pragma unroll
    for (int e = 0; e < NS * NS; e++) a[e] = A[e];
#pragma unroll
    for (int e = 0; e < NS * NR; e++) b[e] = B[e];
#pragma unroll
    for (int c = 0; c < NS; c++) {
        int piv = c; double best = fabs(a[c + c * NS]);
#pragma unroll
        for (int r = c + 1; r < NS; r++) if (fabs(a[r + c * NS]) > best) { best = fabs(a[r + c * NS]); piv = r; }
#pragma unroll
        for (int r = c + 1; r < NS; r++)
            if (piv == r) {
#pragma unroll
                for (int j = 0; j < NS; j++) { const double t = a[c + j * NS]; a[c + j * NS] = a[r + j * NS]; a[r + j * NS] = t; }
#pragma unroll
                for (int j = 0; j < NR; j++) { const double t = b[c + j * NS]; b[c + j * NS] = b[r + j * NS]; b[r + j * NS] = t; }
            }
#pragma unroll
        for (int r = c + 1; r < NS; r++) {
            const double f = a[r + c * NS] / a[c + c * NS];
            if (f != 0.0) {
#pragma unroll
                for (int j = c; j < NS; j++) a[r + j * NS] -= f * a[c + j * NS];
#pragma unroll
                for (int j = 0; j < NR; j++) b[r + j * NS] -= f * b[c + j * NS];
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NR; j++)
#pragma unroll
        for (int r = NS - 1; r >= 0; r--) {
            double s = b[r + j * NS];
#pragma unroll
            for (int c = r + 1; c < NS; c++) s -= a[r + c * NS] * X[c + j * NS];
            X[r + j * NS] = s / a[r + r * NS];
        }
}

constexpr unsigned long long kSentinel = 0xFFFFFFFFFFFFFFFFull;   // a NaN payload no arithmetic produces

__device__ __forceinline__ unsigned long long ld_relaxed_sys_bits(const double *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// L2 load that the compiler can neither sink into the branch that uses the value nor merge: the prefix phase wants
// all loads of a lane in flight before the first use (sunk loads cost one L2 round trip each, ~30 us per evaluation)
__device__ __forceinline__ double ld_cg_now(const double *p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void red_add_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// The solver context, compiled per species count so that all block algebra is unrolled into registers.
//
// Per-wavenumber work (block products, C1 \ rhs) is spread over "k groups": GW consecutive lanes of a warp own one
// wavenumber, lane ge of the group owns element (ga, gb) of its Ns x Ns block; operands held by the other lanes of
// the group come through shuffles, operands in memory are read directly (the 16 lanes of a group read one 128-byte
// line).  The summation order of every product equals blk_mul's, so both forms give the same bits.
template <int NS>
struct Ctx {
    static constexpr int NB = NS * NS;
    static constexpr int GW = NS == 1 ? 1 : NS == 2 ? 4 : 16;      // lanes per wavenumber group
    static constexpr int KPW = 32 / GW;                            // wavenumbers per warp
    const GridParams &P;
    cg::grid_group grid;
    int dof, gtid, gthreads, lane, gwarp, gwarps;
    int ge, ga, gb, gbase; bool gact;
    unsigned xepoch = 0;
    long tix = 0;                 // time index of the evaluation in progress (tagged dense kernels)
    bool failed = false;
    long long tp;
    __device__ Ctx(const GridParams &p) : P(p), grid(cg::this_grid()) {
        dof = p.Nk * NB;
        gtid = blockIdx.x * blockDim.x + threadIdx.x; gthreads = gridDim.x * blockDim.x;
        lane = threadIdx.x & 31; gwarp = gtid >> 5; gwarps = gthreads >> 5;
        const int l = lane % GW;
        gbase = lane - l; gact = l < NB; ge = gact ? l : 0; ga = ge % NS; gb = ge / NS;
        tp = clock64();
        xepoch = *p.xepoch_state;
    }
    __device__ __forceinline__ void mark(int slot) {
        if (gtid == 0) { const long long t = clock64(); P.prof[slot] += (unsigned long long)(t - tp); tp = t; }
    }
    // element idx of a block whose elements are spread over this lane's k group
    __device__ __forceinline__ double sh(double v, int idx) const { return __shfl_sync(0xffffffffu, v, gbase + idx); }
    // this lane's element of A*B: A, B blocks in memory / A in memory, B in the group's registers / A in registers, B in memory
    __device__ __forceinline__ double mm_gg(const double *A, const double *B) const {
        double s = A[ga] * B[gb * NS];
#pragma unroll
        for (int l = 1; l < NS; l++) s += A[ga + l * NS] * B[l + gb * NS];
        return s;
    }
    __device__ __forceinline__ double mm_gr(const double *A, double v) const {
        double s = A[ga] * sh(v, gb * NS);
#pragma unroll
        for (int l = 1; l < NS; l++) s += A[ga + l * NS] * sh(v, l + gb * NS);
        return s;
    }
    __device__ __forceinline__ double mm_rg(double v, const double *B) const {
        double s = sh(v, ga) * B[gb * NS];
#pragma unroll
        for (int l = 1; l < NS; l++) s += sh(v, ga + l * NS) * B[l + gb * NS];
        return s;
    }
    // this lane's element of A \ T (A in memory, T spread over the group): every lane solves the column it sits in
    __device__ __forceinline__ double solve_group(const double *A, double t) const {
        if (NS == 1) return t / A[0];
        double a[NB], b[NS], x[NS];
#pragma unroll
        for (int e = 0; e < NB; e++) a[e] = A[e];
#pragma unroll
        for (int r = 0; r < NS; r++) b[r] = sh(t, r + gb * NS);
        blk_solve<NS, 1>(a, b, x);
        double r = x[0];
#pragma unroll
        for (int i = 1; i < NS; i++) if (ga == i) r = x[i];
        return r;
    }

    // ---- (1) G1 = C'FC, G2 = C'F, G3 = FC of one wavenumber from the group's F (lane e holds F[e]), stored species-major
    //      ([e][k]) together with a copy of F so that the line walk reads unit stride along k.  Called by whole warps.
    __device__ __forceinline__ void g_block(int k, bool valid, double f) {
        const int Nk = P.Nk;
        const double *Ck = P.C + (size_t)k * NB;
        double g3 = sh(f, ga) * Ck[gb * NS];                               // F C
#pragma unroll
        for (int l = 1; l < NS; l++) g3 += sh(f, ga + l * NS) * Ck[l + gb * NS];
        double g2 = Ck[ga * NS] * sh(f, gb * NS);                          // C' F
#pragma unroll
        for (int l = 1; l < NS; l++) g2 += Ck[l + ga * NS] * sh(f, l + gb * NS);
        double g1 = Ck[ga * NS] * sh(g3, gb * NS);                         // C' (F C)
#pragma unroll
        for (int l = 1; l < NS; l++) g1 += Ck[l + ga * NS] * sh(g3, l + gb * NS);
        if (valid) {
            const size_t w = (size_t)ge * Nk + k;
            P.G[w] = f; P.G[w + dof] = g1; P.G[w + 2 * (size_t)dof] = g2; P.G[w + 3 * (size_t)dof] = g3;
        }
    }
    __device__ void phase_G(const double *F) {
        for (int k0 = gwarp * KPW; k0 < P.Nk; k0 += gwarps * KPW) {
            const int kk = k0 + lane / GW, k = min(kk, P.Nk - 1);
            g_block(k, kk < P.Nk && gact, F[(size_t)k * NB + ge]);
        }
    }

    // ---- (2) line sums of A_m[a,b,iq,ip] along the diagonals d = ip - iq and the antidiagonals s = iq + ip <= Nk-2 that
    //      bengtzelius3! (:2-43) walks.  The work is split by species pair e = (a,b) FIRST: a block stages the four operand
    //      vectors of ITS e (F, G1, G2, G3 along k) and the wavenumbers in shared memory (zero padded to a multiple of 4).
    //      The (iq, ip) plane is cut into 4 x 4 tiles; a warp task is one tile diagonal (J - I = D >= 0: lines 4D-3 .. 4D+3)
    //      or tile antidiagonal (I + J = S: lines 4S .. 4S+6), its lanes stride over the tiles of the task.  Only d >= 0
    //      is walked: for a fixed species pair the entry at (ip, iq) equals the one at (iq, ip) (t1 <-> t4, t2 <-> t3 and
    //      the sign of p^2 - q^2 swap together), so the line sums of d and -d are the same number, and an antidiagonal
    //      is twice its half with iq < ip (plus the middle entry).  A lane keeps
    //      the 2 x 4 x 5 operands of a tile and 3 x 7 line accumulators in registers: 2.5 shared-memory words and 12
    //      FP64 instructions per entry, which is what bounds this phase.
    //      The 21 partial sums of a task are written per task; phase_scan adds the (at most two) tasks of a line.
    //      Tasks are dealt to warps longest first in snake order; rank r of a k-sharded solve takes every world-th task.
    // One half (two consecutive wavenumbers) of a chunk's staged operands: F k, G1 k, G2 k, G3 k and k^2.  The staging
    // multiplies the four vectors by k so that a product of a q-side and a p-side operand carries the weight pq.
    struct Half { double2 F, g1, g2, g3, k2; };
    static __device__ __forceinline__ double el(const double2 &v, int i) { return i ? v.y : v.x; }
    static __device__ __forceinline__ void halve(Half &h) {
        h.F.x *= 0.5; h.F.y *= 0.5; h.g1.x *= 0.5; h.g1.y *= 0.5; h.g2.x *= 0.5; h.g2.y *= 0.5; h.g3.x *= 0.5; h.g3.y *= 0.5;
    }
    // entries (A0 + a, B0 + b), a, b in {0, 1}, of a tile: 12 FP64 instructions each
    template <bool DIAG, int A0, int B0>
    static __device__ __forceinline__ void quarter(const Half &x, const Half &y, double (&acc)[3][8]) {
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++) {
                const int idx = DIAG ? (B0 + b) - (A0 + a) + 3 : (A0 + a) + (B0 + b);
                // with P = pq: t1 = P F_q G1_p, t2 = P G2_p G3_q, t3 = P G2_q G3_p, t4 = P G1_q F_p; d = p^2 - q^2
                const double t4 = el(x.g1, a) * el(y.F, b), t3 = el(x.g2, a) * el(y.g3, b);
                const double v = fma(el(x.F, a), el(y.g1, b), t4), dm = fma(el(x.F, a), el(y.g1, b), -t4);
                const double u = fma(el(y.g2, b), el(x.g3, a), t3);
                const double d = el(y.k2, b) - el(x.k2, a), c = v - u;
                acc[0][idx] += v + u;                              // :186-192 (V1: pq, V2: pq(p^2-q^2)^2, V3: 2pq(p^2-q^2))
                acc[1][idx] = fma(c * d, d, acc[1][idx]);
                acc[2][idx] = fma(dm, d, acc[2][idx]);
            }
    }
    // Walk of the tiles I = lo + l, lo + l + 16, ... < hi (l = lane within its half warp) of a tile diagonal (J = I + off) or
    // antidiagonal (J = off - I); the two half warps walk two different tasks at the same time (phase_lines).
    // The tile `mid` (middle of an antidiagonal, which holds both members of its mirror pairs) enters with half weight
    // because the caller doubles the sums afterwards.  Measured and dropped (profiles/README.md, round 2): loading the
    // next tile's operands into the registers of the current one as its halves die, pinned with warp fences, and forcing
    // the first use of every load group before the next group is issued -- ptxas' own placement (all 20 loads next to
    // their first use) is 5 % faster, the two warps of an SM partition hide each other's load latency.
    template <bool DIAG>
    __device__ __forceinline__ void tile_walk(const double *sm, int gs, int lo, int hi, int off, int mid, double (&acc)[3][8]) {
        const double *sF = sm, *s1 = sm + gs, *s2 = sm + 2 * gs, *s3 = sm + 3 * gs, *sk = sm + 4 * gs;
        // the 4 operands of chunk I live in the 16-byte slots 2I and 2I+1, exchanged when bit 2 of I is set (swz): the 8
        // lanes of a quarter warp then hit 8 different bank groups instead of 4 twice
        auto ldh = [&](int chunk, int h) {
            const int o = 4 * chunk + 2 * (h ^ ((chunk >> 2) & 1));
            Half r;
            r.F = *reinterpret_cast<const double2 *>(sF + o); r.g1 = *reinterpret_cast<const double2 *>(s1 + o);
            r.g2 = *reinterpret_cast<const double2 *>(s2 + o); r.g3 = *reinterpret_cast<const double2 *>(s3 + o);
            r.k2 = *reinterpret_cast<const double2 *>(sk + o);
            return r;
        };
        int I = lo + (lane & 15);
        for (; I < hi; I += 16) {
            const int J = DIAG ? I + off : off - I;
            Half x0 = ldh(I, 0), x1 = ldh(I, 1);
            const Half y0 = ldh(J, 0), y1 = ldh(J, 1);
            if (!DIAG && __builtin_expect(I == mid, 0)) { halve(x0); halve(x1); }
            quarter<DIAG, 0, 0>(x0, y0, acc);
            quarter<DIAG, 0, 2>(x0, y1, acc);
            quarter<DIAG, 2, 0>(x1, y0, acc);
            quarter<DIAG, 2, 2>(x1, y1, acc);
        }
    }
    __device__ void phase_lines() {
        extern __shared__ double gsm[];
        const int Nk = P.Nk, world = P.world, rank = P.rank;
        const int nc = grid_chunks(Nk), nanti = grid_anti_tasks(Nk);
        const size_t xwords = grid_lsum_words(Nk, NS);
        xepoch++;
        const int e = blockIdx.x % NB;                       // species pair of this block
        const int chunk = blockIdx.x / NB, nchunk = (gridDim.x - e + NB - 1) / NB;    // blocks sharing this e
        const int gs = grid_smem_stride(Nk);
        for (int i = threadIdx.x; i < gs; i += blockDim.x) {
            const bool in = i < Nk;
            const int c = i >> 1, ph = ((c ^ ((c >> 3) & 1)) << 1) | (i & 1);      // swz: see tile_walk
            const double kq = in ? P.kvec[i] : 0.0;
#pragma unroll
            for (int j = 0; j < 4; j++) gsm[j * gs + ph] = in ? P.G[(size_t)j * dof + (size_t)e * Nk + i] * kq : 0.0;
            gsm[4 * gs + ph] = kq * kq;
        }
        __syncthreads();
        const double pab = P.pref[e];
        // the two warps of an SM partition (w and w + 4) take complementary positions of the dealing order, so that every
        // partition carries the same work (tasks are dealt longest first in snake order)
        const int wpb = blockDim.x >> 5, hw = wpb >> 1, w = threadIdx.x >> 5, wstride = nchunk * wpb;
        const int wi = w < hw ? chunk * hw + w : wstride - 1 - (chunk * hw + w - hw);
        // A warp takes a PAIR of neighbouring tasks of the same kind at a time, one per half warp (lam = 2 * pair + half): the
        // cost outside the tile loop (first operand loads, zeroing, reduction, stores) is paid once per pair, and short
        // tasks fill the lanes better.  The unit that is dealt is the pair: position t of the dealing order = 2 * pair + kind.
        const int ncp = (nc + 1) >> 1, half = lane >> 4;
        const size_t l7 = (size_t)(nc + nanti) * 7;
        const int *mine = (P.deal != nullptr && (int)gridDim.x == P.deal_grid && blockDim.x == kGridThreads)
                              ? P.deal + ((size_t)e * P.deal_wsmax + chunk * wpb + w) * P.deal_stride : nullptr;
        for (int round = 0;; round++) {
            int slot;
            if (mine) { slot = __ldg(mine + round); if (slot < 0) break; }                 // the host's deal (grid_build_deal)
            else {
                slot = round * wstride + ((round & 1) ? wstride - 1 - wi : wi);
                if (round * wstride * world >= 2 * ncp) break;
            }
            const int t = slot * world + (rank + slot) % world;   // position in the longest-first order; the rotation gives
                                                                  // every rank both kinds of task when world is even
            if (t >= 2 * ncp) continue;
            const int lam = 2 * (t >> 1) + half, kind = t & 1;    // task length nc - lam; kind 0: D = lam, 1: S = nc-lam-1
            const bool valid = lam < nc && !(kind == 1 && nc - lam - 1 >= nanti);
            double acc[3][8];
#pragma unroll
            for (int m = 0; m < 3; m++)
#pragma unroll
                for (int i = 0; i < 8; i++) acc[m][i] = 0.0;
            int task;
            if (kind == 1) {
                // the mirror argument holds along an antidiagonal too: tiles I < J count twice; the middle tile (S even)
                // holds both members of its pairs already and is walked with half weight
                const int S = nc - lam - 1; task = nc + S;
                tile_walk<false>(gsm, gs, 0, valid ? (S >> 1) + 1 : 0, S, (S & 1) ? -1 : (S >> 1), acc);
#pragma unroll
                for (int m = 0; m < 3; m++)
#pragma unroll
                    for (int i = 0; i < 7; i++) acc[m][i] += acc[m][i];
            }
            else { task = lam; tile_walk<true>(gsm, gs, 0, valid ? nc - lam : 0, lam, -1, acc); }
            // 24 values (3 x 7 used) summed over the 16 lanes of a half warp by halving: lanes exchange the half they do not
            // keep; both halves of the warp reduce their own task in the same shuffles
            double v12[12], v6[6], v4[4], v2[2];
            {
                const bool up = lane & 8;
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    const double lo = acc[i / 8][i % 8], hi = acc[(i + 12) / 8][(i + 12) % 8];
                    const double got = __shfl_xor_sync(0xffffffffu, up ? lo : hi, 8);
                    v12[i] = (up ? hi : lo) + got;
                }
            }
            {
                const bool up = lane & 4;
#pragma unroll
                for (int i = 0; i < 6; i++) { const double got = __shfl_xor_sync(0xffffffffu, up ? v12[i] : v12[i + 6], 4); v6[i] = (up ? v12[i + 6] : v12[i]) + got; }
            }
            {
                const bool up = lane & 2;
#pragma unroll
                for (int i = 0; i < 3; i++) { const double got = __shfl_xor_sync(0xffffffffu, up ? v6[i] : v6[i + 3], 2); v4[i] = (up ? v6[i + 3] : v6[i]) + got; }
                v4[3] = 0.0;
            }
            {
                const bool up = lane & 1;
#pragma unroll
                for (int i = 0; i < 2; i++) { const double got = __shfl_xor_sync(0xffffffffu, up ? v4[i] : v4[i + 2], 1); v2[i] = (up ? v4[i + 2] : v4[i]) + got; }
            }
            // this lane now holds the values number vi0, vi0 + 1 of the flattened [3][8] accumulators of its half's task
            const int sub0 = (lane & 1) ? 2 : 0;
            const int vi0 = (lane & 8 ? 12 : 0) + (lane & 4 ? 6 : 0) + (lane & 2 ? 3 : 0) + sub0;
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const int vi = vi0 + i, m = vi >> 3, idx = vi & 7;
                if (valid && sub0 + i < 3 && m < 3 && idx < 7) {
                    const double tot = v2[i] * ((m == 2) ? 2.0 * pab : pab);
                    const size_t w = (size_t)(m * NB + e) * l7 + (size_t)task * 7 + idx;          // [sequence][task][7]: see phase_scan
                    if (world == 1) P.Lsum[w] = tot;
                    else P.xpeer[rank][(size_t)(xepoch & 3u) * xwords + w] = tot;       // this rank's own buffer; the peers pull
                }
            }
        }
        if (world > 1) {
            // announce: every block adds 1 to the arrival counter of this exchange on every rank once its sums are out
            // (the sums are in this GPU's L2 -- where the peers' NVLink loads are served -- after the device-scope fence of
            // every storing thread; the announcing thread adds the system-scope fence in front of its counter updates)
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) __threadfence_system();
            if (threadIdx.x == 0)
                for (int r = 0; r < world; r++) red_add_sys((unsigned long long *)(P.xpeer[r] + 4 * xwords) + (xepoch & 3u), 1ull);
        }
    }

    // ---- (3) bengtzelius3! (:2-43) as a prefix over the line sums.  One block per (m, e) sequence: its warps take
    //      contiguous segments, form the increments (diagonal + lower diagonal - antidiagonal) straight from the task
    //      sums, scan them 32 at a time and add the totals of the segments before them.
    __device__ void phase_scan() {
        __shared__ double wtot[kGT / 32];
        const int Nk = P.Nk, world = P.world, nseq = 3 * NB;
        const int nc = grid_chunks(Nk);
        const size_t xwords = grid_lsum_words(Nk, NS), l7 = (size_t)(nc + grid_anti_tasks(Nk)) * 7;
        const double *lcur = P.Lsum;                           // sharded: filled by the landing pass below
        const long long t0 = clock64();
        // A line collects the entries of at most two tasks.  Increment of row ik (0-based): diagonal ik + diagonal -ik
        // (the same number, see phase_lines) - antidiagonal ik-1; ik = 0: the main diagonal alone.  Word addresses are
        // clamped so that the 4 loads of a row are unconditional; the values are selected afterwards.
        struct Row { size_t w[4]; bool two_d, two_a, pos; };
        auto row_of = [&](int ik, int c) {
            Row r;
            const int D0 = ik >> 2, rr = ik & 3;
            r.two_d = rr > 0 && D0 + 1 <= nc - 1;
            const size_t cb = (size_t)c * l7;                    // the sums are stored [sequence][task][7]: consecutive rows of
                                                                 // a sequence are neighbours in memory (14 instead of 32 sectors
                                                                 // per warp load)
            r.w[0] = cb + (size_t)D0 * 7 + rr + 3;
            r.w[1] = cb + (size_t)(r.two_d ? D0 + 1 : D0) * 7 + (r.two_d ? rr - 1 : rr + 3);
            const int s_ = max(ik - 1, 0), S0 = s_ >> 2, q = s_ & 3;
            r.two_a = q <= 2 && S0 >= 1;
            r.w[2] = cb + (size_t)(nc + S0) * 7 + q;
            r.w[3] = cb + (size_t)(nc + (r.two_a ? S0 - 1 : S0)) * 7 + (r.two_a ? q + 4 : q);
            r.pos = ik > 0;
            if (!r.pos) { r.w[2] = r.w[0]; r.w[3] = r.w[0]; }          // row 0 has no antidiagonal: any valid word
            return r;
        };
        auto combine = [&](const Row &r, const double (&v)[4]) -> double {
            double d = v[0]; if (r.two_d) d += v[1];
            double a = v[2]; if (r.two_a) a += v[3];
            return r.pos ? (d + d) - a : d;
        };
        auto increment = [&](int ik, int c) -> double {
            const Row r = row_of(ik, c);
            double v[4];
#pragma unroll
            for (int i = 0; i < 4; i++) v[i] = ld_cg_now(lcur + r.w[i]);
            return combine(r, v);
        };
        if (world > 1) {
            // Landing pass of a k-sharded solve.  One thread per block waits for the arrival counter of this exchange (all
            // blocks of all ranks have announced); then the grid PULLS every task's sums from the buffer of the rank that
            // computed it (NVLink loads from peer memory, 7 consecutive words per task and sequence) into the local Lsum, and re-arms this
            // rank's own buffer of two exchanges ahead.  Every word still validates itself (sentinel until stored).
            if (threadIdx.x == 0) {
                const unsigned long long *cnt = (const unsigned long long *)(P.xpeer[P.rank] + 4 * xwords) + (xepoch & 3u);
                const unsigned long long want = (unsigned long long)world * gridDim.x;
                unsigned spins = 0;
                while (ld_acquire_sys(cnt) < want)
                    if ((++spins & 255u) == 0 && clock64() - t0 > P.spin_limit_cycles) { failed = true; break; }
            }
            __syncthreads();
            const size_t slot_off = (size_t)(xepoch & 3u) * xwords;
            unsigned long long *clr = (unsigned long long *)(P.xpeer[P.rank] + (size_t)((xepoch + 2u) & 3u) * xwords);
            auto source = [&](size_t i) -> const double * {
                const int task = (int)((i % l7) / 7);              // [sequence][task][7]
                const int t = task < nc ? 2 * (task >> 1) : 2 * ((2 * nc - task - 1) >> 1) + 1;    // position of its pair in the dealing order (phase_lines)
                const int slot = t / world, owner = ((t - slot * world) - slot % world + world) % world;
                return P.xpeer[owner] + slot_off + i;
            };
            constexpr int kU = 12;                                 // remote loads in flight per thread (one NVLink round trip each)
            for (size_t i0 = gtid; i0 < xwords; i0 += (size_t)gthreads * kU) {
                double v[kU];
#pragma unroll
                for (int u = 0; u < kU; u++) {
                    const size_t i = i0 + (size_t)u * gthreads;
                    v[u] = ld_cg_now(source(i < xwords ? i : i0));
                }
#pragma unroll
                for (int u = 0; u < kU; u++) {
                    const size_t i = i0 + (size_t)u * gthreads;
                    if (i >= xwords) continue;
                    unsigned long long bits = (unsigned long long)__double_as_longlong(v[u]);
                    unsigned spins = 0;
                    while (bits == kSentinel) {
                        bits = ld_relaxed_sys_bits(source(i));
                        if ((++spins & 1023u) == 0 && clock64() - t0 > P.spin_limit_cycles) { failed = true; break; }
                    }
                    P.Lsum[i] = __longlong_as_double((long long)bits);
                    clr[i] = kSentinel;
                }
            }
            if (gtid == 0) ((unsigned long long *)(P.xpeer[P.rank] + 4 * xwords))[(xepoch + 2u) & 3u] = 0ull;
            if (failed) *P.status = MCT_CUDA_ERROR;
            grid.sync();
            failed = false;
        }
        const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        const int seg = ((Nk + nw * 32 - 1) / (nw * 32)) * 32;
        const int b0 = min(Nk, w * seg), b1 = min(Nk, b0 + seg);
        constexpr int kFast = 8;                               // segments of up to 8 x 32 rows stay in registers
        for (int c = blockIdx.x; c < nseq; c += gridDim.x) {
            double *Tm = P.Tm + (size_t)c * Nk;
            double carry = 0.0;
            if (seg <= 32 * kFast) {
                double v[kFast], raw[kFast][4];
#pragma unroll
                for (int j = 0; j < kFast; j++) {                  // rows past the segment load row Nk-1 and are dropped
                    const Row r = row_of(min(b0 + 32 * j + lane, Nk - 1), c);
#pragma unroll
                    for (int i = 0; i < 4; i++) raw[j][i] = ld_cg_now(lcur + r.w[i]);
                }
#pragma unroll
                for (int j = 0; j < kFast; j++) { const int ik = b0 + 32 * j + lane; v[j] = (ik < b1) ? combine(row_of(min(ik, Nk - 1), c), raw[j]) : 0.0; }
#pragma unroll
                for (int j = 0; j < kFast; j++) {
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, v[j], o); if (lane >= o) v[j] += u; }
                    v[j] += carry;
                    carry = __shfl_sync(0xffffffffu, v[j], 31);
                }
                if (lane == 0) wtot[w] = carry;
                __syncthreads();
                double off = 0.0;
                for (int j = 0; j < w; j++) off += wtot[j];
#pragma unroll
                for (int j = 0; j < kFast; j++) { const int ik = b0 + 32 * j + lane; if (ik < b1) Tm[ik] = v[j] + off; }
            } else {
                for (int i0 = b0; i0 < b1; i0 += 32) {
                    const int ik = i0 + lane;
                    double v = (ik < b1) ? increment(ik, c) : 0.0;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
                    v += carry;
                    if (ik < b1) Tm[ik] = v;
                    carry = __shfl_sync(0xffffffffu, v, 31);
                }
                if (lane == 0) wtot[w] = carry;
                __syncthreads();
                double off = 0.0;
                for (int j = 0; j < w; j++) off += wtot[j];
                if (w > 0) for (int ik = b0 + lane; ik < b1; ik += 32) Tm[ik] += off;     // this lane's own stores
            }
            __syncthreads();
        }
        if (failed) *P.status = MCT_CUDA_ERROR;
    }
    // K = k T1 + T2/k^3 + T3/k   (:215-218)
    __device__ __forceinline__ double K_of(int ik, int e) const {
        const double k = P.kvec[ik];
        return k * P.Tm[(size_t)e * P.Nk + ik] + P.Tm[(size_t)(NB + e) * P.Nk + ik] / (k * k * k) + P.Tm[(size_t)(2 * NB + e) * P.Nk + ik] / k;
    }

    // P[iq] = sum_{ik,ip} prefactor*J*V * F[ik]*F[ip]   (ModeCouplingKernels.jl:237-256); ends with a grid barrier
    __device__ void eval_ddim(const double *F, double *Kout) {
        const int Nk = P.Nk;
        __shared__ double red[kGT / 32];
        // tagged variant (src/Kernels/ActiveMCTKernel.jl:125-136): the second factor is the collective F at this time
        const double *F2 = P.Fc ? P.Fc + (size_t)tix * Nk : F;
        for (int iq = blockIdx.x; iq < Nk; iq += gridDim.x) {
            const double *W = P.W + (size_t)iq * Nk * Nk;
            double s = 0.0;
            for (int i = threadIdx.x; i < Nk * Nk; i += blockDim.x) {
                const int ik = i / Nk, ip = i - ik * Nk;
                s += W[i] * F[ik] * F2[ip];
            }
            s = warp_sum(s);
            if (lane == 0) red[threadIdx.x >> 5] = s;
            __syncthreads();
            if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < kGT / 32; w++) t += red[w]; Kout[iq] = t; }
            __syncthreads();
        }
        grid.sync();
    }

    // ---- K = kernel(F)  (ends with a grid barrier; Kout valid for every thread afterwards)
    __device__ void eval(const double *F, double *Kout) {
        if (P.type == GK_DDIM) { eval_ddim(F, Kout); return; }
        tp = clock64();
        phase_G(F);
        grid.sync();
        mark(0);
        phase_lines();
        grid.sync();
        mark(1);
        phase_scan();
        grid.sync();
        mark(2);
        for (size_t i = gtid; i < (size_t)dof; i += gthreads) Kout[i] = K_of((int)(i / NB), (int)(i % NB));
        grid.sync();
        mark(3);
    }

    // max |a - b| over the grid through a rotating atomic word; every thread returns the same value
    __device__ double grid_max(double local, unsigned &ecount) {
        __shared__ double red[kGT / 32];
        double m = warp_max(local);
        if (lane == 0) red[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < kGT / 32; w++) t = fmax(t, red[w]);
            atomicMax(P.err + (ecount & 1u), (unsigned long long)__double_as_longlong(t));
            if (blockIdx.x == 0) P.err[(ecount + 1u) & 1u] = 0ull;     // nobody reads or writes that word in this phase
        }
        grid.sync();
        const double r = __longlong_as_double((long long)*(volatile unsigned long long *)(P.err + (ecount & 1u)));
        ecount++;
        return r;
    }

    // this lane's element of F = C1 \ (-K*C2 + C3) at block offset o; K spread over the group  (:326-337)
    __device__ __forceinline__ double update_group(size_t o, double Kreg) const {
        double t = mm_rg(Kreg, P.C2 + o);
        t = -t + P.C3[o + ge];
        return solve_group(P.C1 + o, t);
    }

    __device__ void run() {
        const int Nk = P.Nk, N = P.N, n4 = 4 * N, n2 = 2 * N;
        unsigned ecount = 0;
#define FT(s) (P.Ft + (size_t)dof * ((s) - 1))
#define KT(s) (P.Kt + (size_t)dof * ((s) - 1))
#define FINT(s) (P.FI + (size_t)dof * ((s) - 1))
#define KINT(s) (P.KI + (size_t)dof * ((s) - 1))
        if (P.mode == 2) {                       // evaluate_kernel!(out, kernel, F, t)
            tix = P.tix0;
            eval(P.Fin, P.F_out);
            return;
        }
        eval(P.F0, P.K0);                        // K0 = evaluate_kernel(kernel, F0, 0)  src/MemoryEquation.jl:45
        if (P.mode == 1) {                       // solve_steady_state  src/SteadyStateMemoryEquation.jl:59-100
            for (int e = gtid; e < dof; e += gthreads) { P.Fold[e] = P.F0[e]; P.Kcur[e] = P.K0[e]; }
            grid.sync();
            long long iter = 0; int status = MCT_OK;
            while (true) {
                iter++;
                if (iter > P.max_iter) { status = MCT_NOT_CONVERGED; break; }
                double lerr = 0.0;
                for (int k = gtid; k < Nk; k += gthreads) {
                    const size_t o = (size_t)k * NB;
                    double t[NB], M[NB], x[NB];
                    blk_mul<NS>(P.Kcur + o, P.F0 + o, t);                                  // :81
                    for (int e = 0; e < NB; e++) M[e] = P.Kcur[o + e] + P.gamma[o + e];    // :83
                    blk_solve<NS, NS>(M, t, x);
                    for (int e = 0; e < NB; e++) { const double d = fabs(x[e] - P.Fold[o + e]); if (d > lerr) lerr = d; P.Fold[o + e] = x[e]; P.Fcur[o + e] = x[e]; }
                }
                const double err = grid_max(lerr, ecount);
                eval(P.Fcur, P.Kcur);                                                      // :89
                if (!(err > P.tol)) break;
            }
            for (int e = gtid; e < dof; e += gthreads) { P.F_out[e] = P.Fcur[e]; P.K_out[e] = P.Kcur[e]; }
            if (gtid == 0) { if (*P.status != MCT_CUDA_ERROR) *P.status = status; *P.kernel_evals = iter; }
            return;
        }
        // ---- time doubling.  F_old = K0*F0 (:64); K_temp slots keep 2 K0 until they are evaluated (:83)
        for (int k = gtid; k < Nk; k += gthreads) blk_mul<NS>(P.K0 + (size_t)k * NB, P.F0 + (size_t)k * NB, P.Fold + (size_t)k * NB);
        for (int e = gtid; e < dof; e += gthreads) {
            for (int s = n2 + 1; s <= n4; s++) KT(s)[e] = P.K0[e] + P.K0[e];
            P.dFh[e] = P.dF0[e];
            P.F_out[e] = P.F0[e]; P.K_out[e] = P.K0[e];
        }
        grid.sync();
        // bootstrap: forward Euler with the trapezoid memory integral  (src/EulerSolver.jl:26-64, 92-112)
        const double de = P.dt0 / (4.0 * N);
        for (int it = 1; it <= n2; it++) {
            for (int k = gtid; k < Nk; k += gthreads) {
                const size_t o = (size_t)k * NB;
                double integ[NB], f[NB];
                // integrand[i] = K_array[i] * dF_reverse[i] = K(t_{i-1}) * dF(t_{it-i}), i = 1..it
                if (it == 1) {
                    blk_mul<NS>(P.K0 + o, P.dFh + o, f);
                    for (int e = 0; e < NB; e++) integ[e] = f[e] * de;                     // :34-36
                } else {
                    double r[NB];
                    blk_mul<NS>(P.K0 + o, P.dFh + (size_t)dof * (it - 1) + o, f);
                    for (int e = 0; e < NB; e++) r[e] = f[e] / 2.0;
                    blk_mul<NS>(KT(it - 1) + o, P.dFh + o, f);
                    for (int e = 0; e < NB; e++) r[e] += f[e] / 2.0;
                    for (int i = 2; i <= it - 1; i++) {
                        blk_mul<NS>(KT(i - 1) + o, P.dFh + (size_t)dof * (it - i) + o, f);
                        for (int e = 0; e < NB; e++) r[e] += f[e];
                    }
                    for (int e = 0; e < NB; e++) integ[e] = r[e] * de;
                }
                const double *Fo = (it == 1) ? P.F0 + o : FT(it - 1) + o;
                const double *dFo = P.dFh + (size_t)dof * (it - 1) + o;
                double gF[NB], rhs[NB], neg[NB], sol[NB];
                blk_mul<NS>(P.gamma + o, Fo, gF);
                if (P.second_order) {                                                      // Euler_step :55-58
                    double bdF[NB];
                    blk_mul<NS>(P.beta + o, dFo, bdF);
                    for (int e = 0; e < NB; e++) { rhs[e] = bdF[e] + gF[e] + P.delta[o + e] + integ[e]; neg[e] = -P.alpha[o + e]; }
                    blk_solve<NS, NS>(neg, rhs, sol);
                    for (int e = 0; e < NB; e++) {
                        const double dFn = dFo[e] + de * sol[e];
                        P.dFh[(size_t)dof * it + o + e] = dFn;
                        FT(it)[o + e] = Fo[e] + de * dFn;
                    }
                } else {                                                                   // :60-61
                    for (int e = 0; e < NB; e++) { rhs[e] = gF[e] + P.delta[o + e] + integ[e]; neg[e] = -P.beta[o + e]; }
                    blk_solve<NS, NS>(neg, rhs, sol);
                    for (int e = 0; e < NB; e++) {
                        P.dFh[(size_t)dof * it + o + e] = sol[e];
                        FT(it)[o + e] = Fo[e] + de * sol[e];
                    }
                }
            }
            grid.sync();
            tix = it;
            eval(FT(it), KT(it));                                                          // == initialize_K_temp! :155-166
            for (int e = gtid; e < dof; e += gthreads) { P.F_out[(size_t)it * dof + e] = FT(it)[e]; P.K_out[(size_t)it * dof + e] = KT(it)[e]; }
        }
        // initialize_integrals! :174-199
        for (int e = gtid; e < dof; e += gthreads) {
            FINT(1)[e] = (FT(1)[e] + P.F0[e]) / 2.0;
            KINT(1)[e] = (3.0 * KT(1)[e] - KT(2)[e]) / 2.0;
            for (int it = 2; it <= n2; it++) {
                FINT(it)[e] = (FT(it)[e] + FT(it - 1)[e]) / 2.0;
                KINT(it)[e] = (KT(it)[e] + KT(it - 1)[e]) / 2.0;
            }
        }
        grid.sync();
        const bool mc = P.type == GK_MC3D;
        long long evals = 0;
        int status = MCT_OK;
        double Dt = P.dt0;
        for (int blk = 0; blk < P.nblocks && status == MCT_OK; blk++, Dt *= 2.0) {
            const double dl = Dt / (4.0 * N);
            for (int it = n2 + 1; it <= n4 && status == MCT_OK; it++) {
                const int i2 = it / 2;
                const long long row = 1 + (long long)n2 * (blk + 1) + (it - n2 - 1);
                tix = (long)row;
                tp = clock64();
                // update_Fuchs_parameters! :300-319 (compute_C3_mutable! :243-288), then update_F! with the stale K_temp[it]
                // and the G matrices of the first evaluation, all on the lanes of the wavenumber's group
                for (int k0 = gwarp * KPW; k0 < Nk; k0 += gwarps * KPW) {
                    const int kk = k0 + lane / GW, k = min(kk, Nk - 1);
                    const bool valid = kk < Nk && gact;
                    const size_t o = (size_t)k * NB, oe = o + ge;
                    const double c1 = ((2.0 / (dl * dl) * P.alpha[oe] + 3.0 / (2.0 * dl) * P.beta[oe]) + KINT(1)[oe]) + P.gamma[oe];   // :314
                    const double c2 = FINT(1)[oe] - P.F0[oe];                                                                            // :315
                    double tv = (5.0 * FT(it - 1)[oe] - 4.0 * FT(it - 2)[oe] + FT(it - 3)[oe]) / (dl * dl);
                    double c3 = mm_gr(P.alpha + o, tv);                                                                                  // :259-260
                    tv = 2.0 / dl * FT(it - 1)[oe] - FT(it - 2)[oe] / (2.0 * dl);
                    c3 += mm_gr(P.beta + o, tv);                                                                                         // :263-264
                    c3 -= mm_gg(KT(it - i2) + o, FT(i2) + o);                                                                            // :267
                    c3 += mm_gg(KT(it - 1) + o, FINT(1) + o);                                                                            // :268
                    c3 += mm_gg(KINT(1) + o, FT(it - 1) + o);                                                                            // :269
#pragma unroll 2
                    for (int j = 2; j <= i2; j++) {                                                                                      // :271-280
                        const double *Ka = KT(it - j) + o, *Kb = KT(it - j + 1) + o, *Fi = FINT(j) + o;
                        double t = (Ka[ga] - Kb[ga]) * Fi[gb * NS];
#pragma unroll
                        for (int l = 1; l < NS; l++) t += (Ka[ga + l * NS] - Kb[ga + l * NS]) * Fi[l + gb * NS];
                        c3 += t;
                    }
#pragma unroll 2
                    for (int j = 2; j <= it - i2; j++) {                                                                                 // :281-285
                        const double *Ki = KINT(j) + o, *Fa = FT(it - j) + o, *Fb = FT(it - j + 1) + o;
                        double t = Ki[ga] * (Fa[gb * NS] - Fb[gb * NS]);
#pragma unroll
                        for (int l = 1; l < NS; l++) t += Ki[ga + l * NS] * (Fa[l + gb * NS] - Fb[l + gb * NS]);
                        c3 += t;
                    }
                    c3 -= P.delta[oe];                                                                                                   // :286
                    const double Kst = KT(it)[oe];
                    if (valid) { P.C1[oe] = c1; P.C2[oe] = c2; P.C3[oe] = c3; }
                    __syncwarp();
                    const double x = update_group(o, Kst);
                    if (valid) FT(it)[oe] = x;
                    if (mc) g_block(k, valid, x);
                }
                grid.sync();
                mark(4);
                long long iter = 0;
                while (true) {
                    iter++;
                    if (iter > P.max_iter) { status = MCT_NOT_CONVERGED; break; }          // :406-408
                    // update_K! :360-369
                    if (mc) { phase_lines(); grid.sync(); mark(1); phase_scan(); grid.sync(); mark(2); }
                    else eval_ddim(FT(it), KT(it));
                    // K from the prefix sums, update_F! + find_error + F_old .= F, and the G matrices of the next evaluation
                    double lerr = 0.0;
                    for (int k0 = gwarp * KPW; k0 < Nk; k0 += gwarps * KPW) {
                        const int kk = k0 + lane / GW, k = min(kk, Nk - 1);
                        const bool valid = kk < Nk && gact;
                        const size_t o = (size_t)k * NB, oe = o + ge;
                        const double Kv = mc ? K_of(k, ge) : KT(it)[oe];
                        const double x = update_group(o, Kv);
                        if (valid) {
                            const double d = fabs(x - P.Fold[oe]); if (d > lerr) lerr = d;
                            P.Fold[oe] = x; FT(it)[oe] = x;
                            if (mc) KT(it)[oe] = Kv;
                        }
                        if (mc) g_block(k, valid, x);
                    }
                    const double err = grid_max(lerr, ecount);
                    mark(5);
                    evals++;                                                               // :416
                    if (!(err > P.tol)) break;
                }
                if (status != MCT_OK) break;
                for (int e = gtid; e < dof; e += gthreads) {                               // update_integrals! :377-384, allocate_results! :429-438
                    FINT(it)[e] = (FT(it)[e] + FT(it - 1)[e]) / 2.0;
                    KINT(it)[e] = (KT(it)[e] + KT(it - 1)[e]) / 2.0;
                    P.F_out[(size_t)row * dof + e] = FT(it)[e]; P.K_out[(size_t)row * dof + e] = KT(it)[e];
                }
                grid.sync();
            }
            if (status != MCT_OK) break;
            for (int e = gtid; e < dof; e += gthreads) {                                   // new_time_mapping! :448-489
                for (int j = 1; j <= n2; j++) {
                    FINT(j)[e] = (FINT(2 * j)[e] + FINT(2 * j - 1)[e]) / 2.0;
                    KINT(j)[e] = (KINT(2 * j)[e] + KINT(2 * j - 1)[e]) / 2.0;
                    FT(j)[e] = FT(2 * j)[e];
                    KT(j)[e] = KT(2 * j)[e];
                }
                for (int j = n2 + 1; j <= n4; j++) { FINT(j)[e] = 0.0; KINT(j)[e] = 0.0; FT(j)[e] = 0.0; KT(j)[e] = 0.0; }
            }
            grid.sync();
        }
        if (gtid == 0) { if (*P.status != MCT_CUDA_ERROR) *P.status = status; *P.kernel_evals = evals; }
#undef FT
#undef KT
#undef FINT
#undef KINT
    }
};

template <int NS>
__global__ void __launch_bounds__(kGT, 1) td_grid_kernel(const GridParams P) {
    Ctx<NS> c(P);
    c.run();
    c.grid.sync();                                   // every thread has read the epoch word
    if (c.gtid == 0) *P.xepoch_state = c.xepoch;
}

__global__ void ddim_table_kernel(int Nk, const double *V, const double *J, double prefactor, double *W) {
    // out W[iq][ik][ip] (ip contiguous) = prefactor * J[iq,ik,ip] * V[iq,ik,ip]   (column-major inputs: iq fastest)
    const size_t n = (size_t)Nk * Nk * Nk;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int ip = (int)(i % Nk), ik = (int)((i / Nk) % Nk), iq = (int)(i / ((size_t)Nk * Nk));
        const size_t src = (size_t)iq + (size_t)Nk * ((size_t)ik + (size_t)Nk * ip);
        W[i] = prefactor * J[src] * V[src];
    }
}

// tables that already are [output][summed][summed] with the last index contiguous (ActiveMCTKernel, src/Kernels/ActiveMCTKernel.jl:28-49)
__global__ void ddim_fuse_kernel(size_t n, const double *V, const double *J, double prefactor, double *W) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) W[i] = prefactor * J[i] * V[i];
}

}  // namespace

int ddim_build_table(int Nk, double prefactor, const double *dV, const double *dJ, double *dW, cudaStream_t s) {
    ddim_table_kernel<<<1024, 256, 0, s>>>(Nk, dV, dJ, prefactor, dW);
    count_launch();
    MCT_CUDA(cudaGetLastError());
    return MCT_OK;
}

int ddim_fuse_table(int Nk, double prefactor, const double *dV, const double *dJ, double *dW, cudaStream_t s) {
    ddim_fuse_kernel<<<1024, 256, 0, s>>>((size_t)Nk * Nk * Nk, dV, dJ, prefactor, dW);
    count_launch();
    MCT_CUDA(cudaGetLastError());
    return MCT_OK;
}

// tile rounds of the pair at position t of the dealing order (the longer of its two half-warp tasks); mirrors phase_lines
static int deal_rounds(int t, int nc, int nanti) {
    const int pair = t >> 1, kind = t & 1;
    int r = 0;
    for (int h = 0; h < 2; h++) {
        const int lam = 2 * pair + h;
        if (lam >= nc) continue;
        int L;
        if (kind) { const int S = nc - lam - 1; if (S >= nanti) continue; L = (S >> 1) + 1; }
        else L = nc - lam;
        r = std::max(r, (L + 15) / 16);
    }
    return r;
}

void grid_build_deal(int Nk, int ns, int rank, int world, int grid, std::vector<int> &table, int &wsmax, int &stride) {
    const int nb = ns * ns, wpb = kGridThreads / 32, hw = wpb / 2;
    const int nc = grid_chunks(Nk), nanti = grid_anti_tasks(Nk), ncp = (nc + 1) / 2;
    wsmax = ((grid + nb - 1) / nb) * wpb;
    const double ovh = 1.5;      // rounds; the measured time does not depend on it between 0 and 5
    // this rank's slots with their cost: tile rounds + `ovh` rounds for what a pair costs outside the tile loop
    std::vector<std::pair<double, int>> slots;
    for (int s = 0; (long long)s * world < 2 * ncp; s++) {
        const int t = s * world + (rank + s) % world;
        if (t >= 2 * ncp) continue;
        const int r = deal_rounds(t, nc, nanti);
        if (r > 0) slots.push_back({r + ovh, s});
    }
    std::stable_sort(slots.begin(), slots.end(), [](const std::pair<double, int> &a, const std::pair<double, int> &b) { return a.first > b.first; });
    std::vector<std::vector<std::vector<int>>> lists(nb);      // [e][warp] -> slots
    size_t longest = 0;
    for (int e = 0; e < nb; e++) {
        const int nchunk = grid > e ? (grid - e + nb - 1) / nb : 0, ws = nchunk * wpb;
        lists[e].assign(ws, {});
        if (ws == 0) continue;
        std::vector<double> bin(nchunk * hw, 0.0), wl(ws, 0.0);
        for (const auto &cs : slots) {
            const int b = (int)(std::min_element(bin.begin(), bin.end()) - bin.begin());
            const int chunk = b / hw, w0 = chunk * wpb + b % hw, w1 = w0 + hw;     // the two warps of that SM partition
            const int wsel = wl[w0] <= wl[w1] ? w0 : w1;
            bin[b] += cs.first; wl[wsel] += cs.first;
            lists[e][wsel].push_back(cs.second);
        }
        for (const auto &l : lists[e]) longest = std::max(longest, l.size());
    }
    stride = (int)longest + 1;
    table.assign((size_t)nb * wsmax * stride, -1);
    for (int e = 0; e < nb; e++)
        for (size_t w = 0; w < lists[e].size(); w++)
            std::copy(lists[e][w].begin(), lists[e][w].end(), table.begin() + ((size_t)e * wsmax + w) * stride);
}

size_t grid_deal_capacity(int Nk, int ns, int grid) {      // the single-GPU table is the largest; a few entries of slack per warp
    std::vector<int> t; int wsmax = 0, stride = 0;
    grid_build_deal(Nk, ns, 0, 1, grid, t, wsmax, stride);
    return t.size() + (size_t)ns * ns * wsmax * 4;
}

int td_grid_launch(const GridParams &P, int n_sm, cudaStream_t stream) {
    MCT_REQUIRE(P.ns >= 1 && P.ns <= kNsMax, MCT_UNSUPPORTED, "Ns > 4 is not supported on the device");
    const void *fn = P.ns == 1 ? (const void *)td_grid_kernel<1> : P.ns == 2 ? (const void *)td_grid_kernel<2>
                   : P.ns == 3 ? (const void *)td_grid_kernel<3> : (const void *)td_grid_kernel<4>;
    // multi-component: five zero-padded vectors of Nk doubles per block in shared memory (phase_lines)
    const size_t smem = (P.type == GK_MC3D) ? (size_t)5 * grid_smem_stride(P.Nk) * sizeof(double) : 0;
    MCT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MCT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kGT, smem));
    MCT_REQUIRE(per_sm >= 1, MCT_CUDA_ERROR, "cooperative kernel does not fit");
    GridParams Pc = P;
    void *args[] = {(void *)&Pc};
    dim3 grid(n_sm), block(kGT);
    MCT_CUDA(cudaLaunchCooperativeKernel(fn, grid, block, args, smem, stream));
    count_launch();
    return MCT_OK;
}

}  // namespace mct
