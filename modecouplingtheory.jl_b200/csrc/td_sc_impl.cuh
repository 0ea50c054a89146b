// td_sc.cu -- ONE persistent cooperative launch runs the whole Fuchs time-doubling solve
// (src/TimeDoublingSolver.jl:512-532) of one or many single-component memory equations.
//
// A "team" of G CTAs (one CTA per SM) owns one equation at a time (layout.cuh):
//   * the vertex tables stay resident on chip for the whole solve: every thread keeps the table values of its first
//     TR term slots in REGISTERS, the next TS in shared memory; only what does not fit streams from L2;
//   * rows (wavenumbers) are owned by CTAs in contiguous ranges.  Everything local in k -- the memory convolution C3
//     (:243-288), history rings, time mapping (:448-489) -- is done by the owner and never crosses CTAs;
//   * one fixed-point iteration (update_K!/update_F!/find_error, :404-417) costs ONE all-gather.  The update is
//     linear in the Bengtzelius prefix sums:  F[k] = c3/c1 - sum_m rho_m[k] T_m[k],  rho_m = (c2/c1) {k, 1/k^3, 1/k},
//     and T_m[k] = off_m(owner) + L_m[k] (offset of the owning CTA + prefix inside it).  The owner publishes
//     F_loc[k] = c3/c1 - sum_m rho_m L_m[k] and its three block totals; every CTA collects all of it, scans the G
//     totals into offsets and finishes F[k] = F_loc[k] - sum_m rho_m off_m with identical arithmetic, so all CTAs
//     hold bit-identical iterates and take the same branch with no second exchange.  K itself is only formed
//     (with the reference's expression k T1 + T2/k^3 + T3/k, :224-227) by the owner when a time step is stored.
// The all-gather needs no barrier and no fence: every 8-byte slot validates itself.  Slots hold a sentinel bit
// pattern until their owner stores the value; a reader issues all its loads at once after a short delay and
// re-issues only those that still show the sentinel.  Four buffers rotate; at exchange e a CTA re-arms its own
// slots of buffer (e+2)&3, which every CTA finished reading before this CTA could complete exchange e-1
// ("exchange protocol").  The file is compiled with -fmad=false: FMAs only where fma() is written.
#include <algorithm>

#pragma once
#include "td_sc.cuh"

namespace mct {

namespace {

constexpr unsigned long long kSentinel = 0xFFFFFFFFFFFFFFFFull;   // a NaN payload no arithmetic produces
constexpr int kTotPerLane = 5;                                     // block totals per lane of the scanning warp -> G <= 160


extern __shared__ __align__(16) unsigned char smem_raw[];

__device__ __forceinline__ void st_relaxed(double *p, double v) {
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed2(double *p, double a, double b) {
    asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st_relaxed_bits(double *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed2_bits(double *p, unsigned long long a, unsigned long long b) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_bits(const double *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ld_relaxed2_bits(const double *p, unsigned long long &a, unsigned long long &b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ double bits2d(unsigned long long v) { return __longlong_as_double((long long)v); }

// ---- thread-block cluster / distributed shared memory (teams of <= 16 CTAs exchange through DSMEM instead of L2)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_cluster(unsigned local_addr, int rank, double v) {      // store into CTA `rank` of the cluster
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    asm volatile("st.relaxed.cluster.shared::cluster.f64 [%0], %1;" ::"r"(r), "d"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_smem_relaxed_bits(const double *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.cluster.shared::cta.u64 %0, [%1];" : "=l"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- tensor memory as a per-warp scratchpad for table values (no MMA involved): warp w owns the 32-lane quarter w % 4 and
//      the column half w / 4 of the CTA's 512 columns; a unit is 24 columns (12 doubles per lane).  Measured (scripts/
//      tmem_bench.cu): 339 B/clk/SM against 127 B/clk/SM for the same values in shared memory, bit-exact round trip.
__device__ __forceinline__ void tmem_ld24(unsigned taddr, unsigned (&r)[24]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23])
                 : "r"(taddr + 16));
}
__device__ __forceinline__ void tmem_st8(unsigned taddr, const unsigned (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
constexpr int kTmemCols = 512;                    // the whole tensor memory of the SM (one CTA per SM)
constexpr int kTmemUnitCols = 2 * kUV;            // 24 columns of 32 bits = 12 doubles per lane
constexpr int kTmemColsPerWarp = kTmemCols / ((kWarps + 3) / 4);      // the warps w, w+4, ... share a lane quarter

#ifdef MCT_PROFILE
#define PROF(i) do { long long now__ = clock64(); prof[i] += now__ - tlast; tlast = now__; } while (0)
#else
#define PROF(i) do { } while (0)
#endif

#ifdef MCT_PROFILE
#define BEACON(ph) do { if ((P.dbg & 2) && tid == 0 && P.prof && !*sAbort && !*(volatile int *)P.abort_flag) ((volatile long long *)P.prof)[((size_t)blockIdx.x) * kProfSlots + 28] = (long long)epoch * 16 + (ph); } while (0)
#else
#define BEACON(ph) do { } while (0)
#endif
enum { X_USE_OFF = 1, X_TRACK = 2 };

// LEAN: the instantiation for the common large-team case -- every unit in tensor memory (UR = 0, no shared-memory or L2
// units), exchange through L2 -- without the code of the other cases: the fixed-point loop shrinks from 2.6 k to 1.7 k
// instructions (41 KB -> 27 KB), inside what the instruction cache holds (profiles/icache_bench.txt).
template <int KPT, int UR, bool LEAN = false>
struct Team {
    const SolveParams &P;
    // shared memory arrays: constant offsets (parameter bank) into the dynamic segment -- no registers, no 64-bit pointers
    unsigned oF1;                // F1 operand of the vertex sum: P.so.F1 (tagged kernels) or P.so.F2
#define sF1 ((const double *)(smem_raw + oF1))
#define sF1_home ((double *)(smem_raw + P.so.F1))
#define sF2 ((double *)(smem_raw + P.so.F2))
#define sRho ((double *)(smem_raw + P.so.rho))
#define sOff ((double *)(smem_raw + P.so.off))
#define sWT ((double *)(smem_raw + P.so.wt))
#define sOwn ((double *)(smem_raw + P.so.own))
#define sPart ((double *)(smem_raw + P.so.part))
#define sLane ((double *)(smem_raw + P.so.lane))
#define sC ((double *)(smem_raw + P.so.C))
#define sX ((double *)(smem_raw + P.so.X))
#define sScan ((double *)(smem_raw + P.so.scan))
#define sTab ((double *)(smem_raw + P.so.tab))
#define sAbort ((volatile int *)(smem_raw + P.so.misc))
#define sXD ((double *)(smem_raw + P.so.xd))
    // identity
    int cta, tid, lane, warp;
#define NkP (even_up(P.Nk))
#define GP (even_up(P.G))
#define EP (even_up(P.ent_max))
#define RP (even_up(P.rows_max))
    int row0, nrows;             // rows owned by this CTA: [row0, row0 + nrows)
    int we0, we_n;               // first (warp,row) entry of this warp and how many it has
    int eb0, eb1;                // entries of the row this thread adds up: row tid, or (3 rows_max <= 32) row prow of table pm
    int prow, pm;
    int owner[KPT];              // owner CTA of the rows this thread finishes (k = j*kThreads + tid: consecutive lanes,
                                 // consecutive rows -- conflict-free in shared memory, coalesced in the exchange buffer)
    // register-resident part of the tables: the 12 values of this lane for the first UR units of the warp
    double tab[UR > 0 ? UR : 1][kUV];
    // exchange
    double *xbase, *qbase;
    unsigned epoch, qepoch;
    double myoff1, myoff2, myoff3;   // offsets of this CTA from the last exchange
    long long slab_base;         // first slot of this CTA in the team layout
    unsigned tbase;              // tensor-memory address of this warp's first table unit
    unsigned sbase;              // shared-space address of smem_raw
#ifdef MCT_PROFILE
    long long prof[kProfSlots], tlast;
#endif

    __device__ __forceinline__ Team(const SolveParams &p) : P(p) {}
    __device__ __forceinline__ bool aborted() const { return *sAbort != 0; }
    // per-row constant of the row this thread owns (valid for tid < nrows)
    __device__ __forceinline__ double own(int field) const { return sOwn[field * RP + (tid < nrows ? tid : 0)]; }
    __device__ __forceinline__ void load_own(const EqDev &E) {
        if (tid < nrows) {
            const int k = row0 + tid;
            sOwn[OWN_K * RP + tid] = E.kvec[k]; sOwn[OWN_F0 * RP + tid] = E.F0[k];
            sOwn[OWN_AL * RP + tid] = E.alpha[k]; sOwn[OWN_BE * RP + tid] = E.beta[k];
            sOwn[OWN_GA * RP + tid] = E.gamma[k]; sOwn[OWN_DE * RP + tid] = E.delta[k];
            sOwn[OWN_K0 * RP + tid] = 0.0;
        }
    }

    // ------------------------------------------------------------------------------------------ spin helpers
    __device__ __forceinline__ bool spin_check(long long t0, unsigned &spins) {
        if ((++spins & 255u) == 0) {
            if (*(volatile int *)P.abort_flag) { *sAbort = 1; return true; }
            if (clock64() - t0 > P.spin_limit_cycles) {
#ifdef MCT_PROFILE
                if (atomicAdd(P.abort_flag + 1, 1) < 40) printf("[mct] spin timeout: cta %d tid %d epoch %u spins %u\n", cta, tid, epoch, spins);
#endif
                atomicExch(P.abort_flag, 1); *sAbort = 1; return true;
            }
        }
        return false;
    }

    // ------------------------------------------------------------------------------------------ block helpers
    __device__ __forceinline__ void warp_inscan3(double &v0, double &v1, double &v2) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double u0 = __shfl_up_sync(0xffffffffu, v0, o), u1 = __shfl_up_sync(0xffffffffu, v1, o), u2 = __shfl_up_sync(0xffffffffu, v2, o);
            if (lane >= o) { v0 += u0; v1 += u1; v2 += u2; }
        }
    }
    // inclusive block scan of three per-thread values (all threads must call; one __syncthreads inside)
    __device__ __forceinline__ void block_inscan3(double &v0, double &v1, double &v2) {
        warp_inscan3(v0, v1, v2);
        if (lane == 31) { sScan[warp] = v0; sScan[kWarps + warp] = v1; sScan[2 * kWarps + warp] = v2; }
        __syncthreads();
        double b0 = 0.0, b1 = 0.0, b2 = 0.0;
        for (int w = 0; w < warp; w++) { b0 += sScan[w]; b1 += sScan[kWarps + w]; b2 += sScan[2 * kWarps + w]; }
        v0 += b0; v1 += b1; v2 += b2;
    }

    // ------------------------------------------------------------------------------------------ kernel evaluation
    // One unit (layout.cuh): 32 positions x 4 rows.  Every lane loads its single operand and the 4 consecutive window
    // operands, forms the 4 products and feeds 12 FMAs with the 12 table values it owns.  The loop is bound by the number of
    // instructions a warp has to issue (two warps per scheduler, ~0.5 IPC), so everything that does not depend on the iterate
    // is decoded once per equation: sUA holds, per (warp, unit), the shared-memory byte addresses of lane 0's operands
    //     x = &S[a0],   y = &W[w0 - kFPad] | 1 (window index decreases with the lane) | 2 (first unit of a new entry).
    // Rows that do not exist in a trailing row block have zero table values, their FMAs add zero (no branches).  The
    // operands of unit u+1 and the addresses of unit u+2 are loaded before the arithmetic of unit u.
    // When a unit starts a new (warp, row block) entry the 12 lane partials of the finished one are parked as they are in
    // the warp scratch [entry][12][kLaneStride] (warp-uniform, rare); the lanes add the columns after the last unit.
    struct Operands { double f1, w[kRB]; };
    __device__ __forceinline__ int2 unit_addr(int u) const { return ((const int2 *)(smem_raw + P.so.desc))[warp * (P.U + 2) + u]; }
    __device__ __forceinline__ void load_operands(int2 ua, Operands &o) const {
        const unsigned l8 = 8u * (unsigned)lane;
        const unsigned aS = (unsigned)ua.x + l8;
        const unsigned aW = ((unsigned)ua.y & ~7u) + ((ua.y & 1) ? 0u - l8 : l8);
        o.f1 = *(const double *)(smem_raw + aS);
#pragma unroll
        for (int r = 0; r < kRB; r++) o.w[r] = *(const double *)(smem_raw + aW + 8 * r);
    }
    __device__ __forceinline__ void park(double (&acc)[kUV], double *&sl) const {
#pragma unroll
        for (int v = 0; v < kUV; v++) { sl[v * kLaneStride] = acc[v]; acc[v] = 0.0; }
        sl += kUV * kLaneStride;
    }
    __device__ __forceinline__ void unit(int flags, const Operands &o, const double (&t)[kUV], double (&acc)[kUV], double *&sl) const {
        if (flags & 2) park(acc, sl);
        const double f0 = o.f1 * o.w[0], f1 = o.f1 * o.w[1], f2 = o.f1 * o.w[2], f3 = o.f1 * o.w[3];
        acc[0] = fma(t[0], f0, acc[0]); acc[1] = fma(t[1], f0, acc[1]); acc[2] = fma(t[2], f0, acc[2]);
        acc[3] = fma(t[3], f1, acc[3]); acc[4] = fma(t[4], f1, acc[4]); acc[5] = fma(t[5], f1, acc[5]);
        acc[6] = fma(t[6], f2, acc[6]); acc[7] = fma(t[7], f2, acc[7]); acc[8] = fma(t[8], f2, acc[8]);
        acc[9] = fma(t[9], f3, acc[9]); acc[10] = fma(t[10], f3, acc[10]); acc[11] = fma(t[11], f3, acc[11]);
    }

    // evaluate_kernel!(K, kernel, F, t) for the whole team (src/Kernels/ModeCouplingKernels.jl:211-228, 450-467),
    // up to the prefix sums inside this CTA.  Inputs: sF1 (collective F at this time for tagged kernels, == sF2
    // otherwise) and sF2 (current iterate), both complete and visible (caller synced).  Output, in owner thread
    // tid < nrows: L_m = sum of the increments of rows row0 .. row0+tid  (so T_m[k] = off_m(cta) + L_m).
    __device__ __forceinline__ void eval_rows(const double *__restrict__ gtab, double &L1, double &L2, double &L3) {
        PROF(8);
        BEACON(7);
#ifdef MCT_PROFILE
        if (epoch + 1 == (unsigned)P.snap_epoch) { long long g__; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g__)); prof[24] = g__; }
#endif
        {
            double acc[kUV];
#pragma unroll
            for (int v = 0; v < kUV; v++) acc[v] = 0.0;
            double *const sl0 = sLane + (size_t)warp * P.lane_cap * kUV * kLaneStride;
            double *sl = sl0 + lane;
            const int UT = P.UT, US = LEAN ? 0 : P.US;             // units in tensor memory, in shared memory
            const int UG = LEAN ? 0 : P.U - UR - UT - US;          // units that fit nowhere on chip: from L2
            // software pipeline: op = operands of the current unit, fl = its flags, uan = addresses of the next unit
            Operands op, nop;
            int2 uan = unit_addr(0);
            load_operands(uan, op);
            int fl = uan.y;
            uan = unit_addr(1);
            int u = 0;
#define MCT_UNIT_STEP(TABLE)                                                  \
            {                                                                     \
                load_operands(uan, nop);                                          \
                const int fln = uan.y;                                            \
                uan = unit_addr(u + 2);                                           \
                TABLE;                                                            \
                unit(fl, op, t, acc, sl);                                         \
                op = nop; fl = fln; u++;                                          \
            }
#pragma unroll
            for (int i = 0; i < UR; i++) {
                const double (&t)[kUV] = tab[i];
                MCT_UNIT_STEP((void)0)
            }
#pragma unroll 2
            for (int i = 0; i < UT; i++) {
                double t[kUV];
                unsigned r[24];
                tmem_ld24(tbase + kTmemUnitCols * i, r);
                MCT_UNIT_STEP(tmem_wait_ld(); _Pragma("unroll") for (int v = 0; v < kUV; v++) t[v] = __hiloint2double((int)r[2 * v + 1], (int)r[2 * v]))
            }
#pragma unroll 2
            for (int i = 0; i < US; i++) {
                double t[kUV];
                const double *ts = sTab + ((size_t)(warp * US + i) * kUV) * 32 + lane;
#pragma unroll
                for (int v = 0; v < kUV; v++) t[v] = ts[v * 32];
                MCT_UNIT_STEP((void)0)
            }
            for (int i = 0; i < UG; i++) {
                const long long slot = slab_base + (long long)warp * P.U + UR + UT + US + i;
                double t[kUV];
                const double *ts = gtab + (slot * kUV) * 32 + lane;
#pragma unroll
                for (int v = 0; v < kUV; v++) t[v] = __ldg(ts + v * 32);
                MCT_UNIT_STEP((void)0)
            }
#undef MCT_UNIT_STEP
            // last entry, then the lanes add the 32 lane partials of the 12*we_n columns
            if (we_n > 0) park(acc, sl);
            __syncwarp();
            for (int c = lane; c < kUV * we_n; c += 32) {
                const double *col = sl0 + (size_t)c * kLaneStride;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                for (int i = 0; i < 32; i += 4) { s0 += col[i]; s1 += col[i + 1]; s2 += col[i + 2]; s3 += col[i + 3]; }
                sPart[we0 * kUV + c] = (s0 + s1) + (s2 + s3);
            }
        }
        PROF(0);
        BEACON(8);
        __syncthreads();
        PROF(1);
        BEACON(9);
        // increments of the owned rows, inclusive prefix over them
        if (3 * P.rows_max <= 32) {
            // warp 0, lane (m, r) = (table, row): the three prefixes run side by side in one segmented scan; the owner thread of
            // row r then fetches its three sums.  Same additions in the same order as the scan of three values per lane.
            L1 = L2 = L3 = 0.0;
            if (warp == 0) {
                const int RM = P.rows_max;
                double x = 0.0;
                const int c3 = 3 * (prow & (kRB - 1)) + pm;
                for (int e = eb0; e < eb1; e++) x += sPart[e * kUV + c3];
                for (int o = 1; o < RM; o <<= 1) { const double up = __shfl_up_sync(0xffffffffu, x, o); if (prow >= o) x += up; }
                const int sl = (lane < RM) ? lane : 0;
                L1 = __shfl_sync(0xffffffffu, x, sl); L2 = __shfl_sync(0xffffffffu, x, RM + sl); L3 = __shfl_sync(0xffffffffu, x, 2 * RM + sl);
            }
        } else {
            double i1 = 0.0, i2 = 0.0, i3 = 0.0;
            if (tid < nrows) {
                const int r3 = 3 * (tid & (kRB - 1));          // row inside its row block
                for (int e = eb0; e < eb1; e++) { i1 += sPart[e * kUV + r3]; i2 += sPart[e * kUV + r3 + 1]; i3 += sPart[e * kUV + r3 + 2]; }
            }
            if (P.rows_max <= 32) { if (warp == 0) warp_inscan3(i1, i2, i3); }
            else block_inscan3(i1, i2, i3);
            L1 = i1; L2 = i2; L3 = i3;
        }
        PROF(2);
#ifdef MCT_PROFILE
        if (epoch + 1 == (unsigned)P.snap_epoch) { long long g__; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g__)); prof[25] = g__; }
#endif
    }

    // K = k T1 + T2/k^3 + T3/k   (:224-227) for the owned row, from the prefix inside the CTA and the CTA's offset
    // (myoff_m: offsets of this CTA as of the last exchange, kept in registers because sOff is rewritten by the next one)
    __device__ __forceinline__ double K_exact(double L1, double L2, double L3) const {
        const double kval = own(OWN_K);
        const double T1 = myoff1 + L1, T2 = myoff2 + L2, T3 = myoff3 + L3;
        return kval * T1 + T2 / (kval * kval * kval) + T3 / kval;
    }

    // ------------------------------------------------------------------------------------------ exchange
    // offset of CTA c (exclusive prefix of the block totals, written by the scanning warps of the last exchange)
    __device__ __forceinline__ void cta_offsets(int c, double &o1, double &o2, double &o3) const {
        o1 = sOff[c]; o2 = sOff[GP + c]; o3 = sOff[2 * GP + c];
    }

    // All-gather.  Owner threads publish `v` for their row; the last owner thread also publishes the CTA's block
    // totals (B1,B2,B3).  Every thread collects the rows k = j*kThreads + tid, every CTA scans the totals into sOff.
    //   X_USE_OFF: out = v_k - (rho_1[k] off_1 + rho_2[k] off_2 + rho_3[k] off_3) with the offsets of k's owner
    //   X_TRACK:   compare with fold (find_error, src/HelperFunctions.jl:55-64; F_old .= F, TimeDoublingSolver.jl:414)
    //   dst_off:   byte offset (in the dynamic shared memory) of the array receiving out, or -1
    // Returns non-zero iff some |out - fold| > tol (identical in every thread of every CTA).  Ends with a barrier.
    __device__ __forceinline__ int exchange(double v, double B1, double B2, double B3, int flags, double (&fold)[KPT], int dst_off,
                                            double (&out)[KPT]) {
        PROF(3);
        unsigned long long w[KPT];
        if (!LEAN && P.G == 1) {
            if (tid < nrows) sX[tid] = v;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < KPT; j++) { const int k = j * kThreads + tid; w[j] = (k < P.Nk) ? (unsigned long long)__double_as_longlong(sX[k]) : 0ull; }
            PROF(5);
        } else if (!LEAN && P.fabric == 1) {
            // ---- cluster fabric: the team is one thread-block cluster; every CTA owns a copy of the exchange buffers in its
            //      shared memory and the publishers store straight into all of them (DSMEM).  Collecting is a local spin.
            epoch++;
            double *lcur = sXD + (size_t)(epoch & 3u) * P.xbuf_words, *lclr = sXD + (size_t)((epoch + 2u) & 3u) * P.xbuf_words;
            // re-arm the own copy of the buffer used two exchanges from now (peers write it only after they have seen this
            // CTA's data of the next exchange, which is published after the barrier that ends this one)
            for (int i = tid; i < P.xrows + P.Nk; i += kThreads) ((unsigned long long *)lclr)[i] = kSentinel;
            if (tid < nrows) {
                const unsigned a = smem_u32(lcur + P.xrows + row0 + tid);
                for (int p = 0; p < P.G; p++) st_cluster(a, p, v);
            }
            if (tid == nrows - 1) {
                const unsigned a = smem_u32(lcur + 4 * cta);
                for (int p = 0; p < P.G; p++) { st_cluster(a, p, B1); st_cluster(a + 8, p, B2); st_cluster(a + 16, p, B3); }
            }
            const long long t0 = clock64();
            unsigned spins = 0;
            PROF(4);
            // arrival: lane t < G of warp 0 spins on the totals of CTA t in the LOCAL copy; every other warp waits at named
            // barrier 1 without issuing (seven spinning warps would take the issue slots the publishing warp needs)
            unsigned long long t0w = 0, t1w = 0, t2w = 0;
            if (warp == 0 && tid < P.G) {
                const double *ts = lcur + 4 * tid;
                while ((t0w = ld_smem_relaxed_bits(ts)) == kSentinel) { if (spin_check(t0, spins)) break; }
                while ((t1w = ld_smem_relaxed_bits(ts + 1)) == kSentinel) { if (spin_check(t0, spins)) break; }
                while ((t2w = ld_smem_relaxed_bits(ts + 2)) == kSentinel) { if (spin_check(t0, spins)) break; }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                const int k = j * kThreads + tid;
                w[j] = 0;
                if (k < P.Nk) { while ((w[j] = ld_smem_relaxed_bits(lcur + P.xrows + k)) == kSentinel) { if (spin_check(t0, spins)) break; } }
            }
            PROF(5);
            if (warp == 0) {     // G <= 16: one warp scans the totals
                const double b1 = bits2d(t0w), b2 = bits2d(t1w), b3 = bits2d(t2w);
                double s1 = b1, s2 = b2, s3 = b3;
                warp_inscan3(s1, s2, s3);
                if (tid < P.G) { sOff[tid] = s1 - b1; sOff[GP + tid] = s2 - b2; sOff[2 * GP + tid] = s3 - b3; }
            }
            __syncthreads();
        } else {
            // ---- L2 fabric: publish, bump the team's arrival counter, ONE thread polls it, everybody else waits at named
            //      barrier 1 without issuing; then every thread loads its rows once (consecutive lanes, consecutive words) and
            //      warp 0 loads the block totals and scans them while the rows are in flight.  Measured alternatives
            //      (scripts/loop_bench.cu, profiles/loop_bench_r02.txt): polling the data words -- by every thread, by one
            //      warp, with or without replicated buffers -- is slower or fragile under arrival skew.
            epoch++;
            double *cur = xbase + (size_t)(epoch & 3u) * P.xbuf_words, *clr = xbase + (size_t)((epoch + 2u) & 3u) * P.xbuf_words;
#ifdef MCT_PROFILE
#define SNAP(i) do { if (epoch == (unsigned)P.snap_epoch) { long long g__; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g__)); prof[i] = g__; } } while (0)
#else
#define SNAP(i) do { } while (0)
#endif
            SNAP(19);
            BEACON(1);
            const int GT = (P.G + 15) & ~15;                 // block totals: three arrays [GT], one per vertex table
            if (tid < nrows) st_relaxed(cur + P.xrows + row0 + tid, v);
            if (tid == nrows - 1) { st_relaxed(cur + cta, B1); st_relaxed(cur + GT + cta, B2); st_relaxed(cur + 2 * GT + cta, B3); }
            if (tid < nrows) st_relaxed_bits(clr + P.xrows + row0 + tid, kSentinel);
            if (tid == nrows - 1) { st_relaxed_bits(clr + cta, kSentinel); st_relaxed_bits(clr + GT + cta, kSentinel); st_relaxed_bits(clr + 2 * GT + cta, kSentinel); }
            const long long t0 = clock64();
            unsigned spins = 0;
            PROF(4);
            BEACON(2);
            // arrival counter: only a hint, no fence orders it after the data, the words still validate themselves.  (Timed arrival
            // without any signal -- wait an adaptive delay after the own publish, load once, re-issue sentinels -- was built and
            // measured in round 2: same speed, the wait is the arrival skew of the slowest CTA, not the counter's latency.)
            if (warp == 0) {
                __syncwarp();
                if (lane == 0) {
                    unsigned long long *ctr = (unsigned long long *)(xbase + 4 * (size_t)P.xbuf_words);
                    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(ctr), "l"(1ull) : "memory");
                    const unsigned long long want = (unsigned long long)P.G * epoch;
                    // (polling harder -- four staggered loads in flight -- or later -- an initial delay, a back-off -- changes the
                    //  solve time by less than 2 %: the wait is the arrival of the slowest CTA plus the L2 latency of its stores)
                    while (ld_relaxed_bits((const double *)ctr) < want) { if (spin_check(t0, spins)) break; }
                }
                __syncwarp();
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
            SNAP(20);
            PROF(5);
            BEACON(3);
            // one shot: the rows of every thread; re-issue only what still shows the sentinel
            const double *src = cur + P.xrows + tid;
#pragma unroll
            for (int j = 0; j < KPT; j++) { w[j] = 0ull; if (j * kThreads + tid < P.Nk) w[j] = ld_relaxed_bits(src + j * kThreads); }
            if (warp < 3) {
                // warp m scans the block totals of table m while the rows are in flight: coalesced loads (lane l: CTAs l + 32 i),
                // transposed through the idle warp scratch so that lane l holds the kTotPerLane consecutive CTAs it adds
                // serially, one warp scan of the lane sums; exclusive prefix in a fixed order, identical in every CTA
                const double *tp = cur + warp * GT;
                double *stg = sLane + warp * (32 * kTotPerLane);
                unsigned long long tv[kTotPerLane];
#pragma unroll
                for (int i = 0; i < kTotPerLane; i++) { tv[i] = 0ull; if (i * 32 + lane < P.G) tv[i] = ld_relaxed_bits(tp + i * 32 + lane); }
                while (true) {
                    bool again = false;
#pragma unroll
                    for (int i = 0; i < kTotPerLane; i++)
                        if (tv[i] == kSentinel) { tv[i] = ld_relaxed_bits(tp + i * 32 + lane); again = true; }
                    const bool stop = again && spin_check(t0, spins);
                    if (!__any_sync(0xffffffffu, again) || __any_sync(0xffffffffu, stop)) break;
                }
                SNAP(24);
#pragma unroll
                for (int i = 0; i < kTotPerLane; i++) stg[i * 32 + lane] = bits2d(tv[i]);
                __syncwarp();
                double le[kTotPerLane], s = 0.0;
#pragma unroll
                for (int i = 0; i < kTotPerLane; i++) { le[i] = s; s += stg[kTotPerLane * lane + i]; }
                double e = s;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const double up = __shfl_up_sync(0xffffffffu, e, o); if (lane >= o) e += up; }
                e -= s;
                double *so = sOff + warp * GP + kTotPerLane * lane;
#pragma unroll
                for (int i = 0; i < kTotPerLane; i++) if (kTotPerLane * lane + i < P.G) so[i] = e + le[i];
                SNAP(25);
            }
            while (true) {
                bool again = false;
#pragma unroll
                for (int j = 0; j < KPT; j++)
                    if (w[j] == kSentinel) { w[j] = ld_relaxed_bits(src + j * kThreads); again = true; }
                if (!again || spin_check(t0, spins)) break;
            }
            SNAP(21);
            BEACON(4);
            __syncthreads();
            BEACON(5);
        }
        PROF(6);
        SNAP(22);
        cta_offsets(cta, myoff1, myoff2, myoff3);
        // finish the rows of this thread: branch-free (rows beyond Nk are computed on clamped addresses and discarded), all
        // shared-memory loads first, so that the KPT rows overlap instead of running one after the other
        int pred = 0;
        double o1[KPT], o2[KPT], o3[KPT], r1[KPT], r2[KPT], r3[KPT];
        if (flags & X_USE_OFF) {
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                const int kc = min(j * kThreads + tid, P.Nk - 1);
                cta_offsets(owner[j], o1[j], o2[j], o3[j]);
                r1[j] = sRho[kc]; r2[j] = sRho[NkP + kc]; r3[j] = sRho[2 * NkP + kc];
            }
        }
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            const int k = j * kThreads + tid;
            const bool valid = k < P.Nk;
            double f = bits2d(w[j]);
            if (flags & X_USE_OFF) { const double g = f - (r1[j] * o1[j] + r2[j] * o2[j] + r3[j] * o3[j]); f = valid ? g : f; }
            if (flags & X_TRACK) { pred |= (valid && fabs(f - fold[j]) > P.tol) ? 1 : 0; fold[j] = f; }
            if (dst_off >= 0 && valid) asm volatile("st.shared.f64 [%0], %1;" ::"r"(sbase + (unsigned)dst_off + 8u * (unsigned)k), "d"(f) : "memory");
            out[j] = f;
        }
        const int any = __syncthreads_or(pred);
        PROF(7);
        SNAP(23);
        BEACON(6);
        return any;
    }

    // ------------------------------------------------------------------------------------------ per-equation setup
    __device__ __forceinline__ void load_tables(const EqDev &E) {
        // operand addresses of this warp's units (see eval_rows); two dummy entries behind the last unit (prefetch)
        for (int u = lane; u < P.U + 2; u += 32) {
            int2 ua = make_int2((int)P.so.F2, (int)P.so.F2);
            if (u < P.U) {
                const int d = P.udesc[slab_base + (long long)warp * P.U + u];
                const unsigned oS = (d & kUSwap) ? P.so.F2 : oF1, oW = (d & kUSwap) ? oF1 : P.so.F2;
                ua.x = (int)(oS + 8u * (unsigned)(d & ((1 << kUA0Bits) - 1)));
                ua.y = (int)(oW + 8u * (unsigned)(((d >> kUW0Shift) & ((1 << kUW0Bits) - 1))) - 8u * (unsigned)kFPad) | ((d & kUNeg) ? 1 : 0) | ((d & kUNewEntry) ? 2 : 0);
            }
            ((int2 *)(smem_raw + P.so.desc))[warp * (P.U + 2) + u] = ua;
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < UR; u++) {
            const long long slot = slab_base + (long long)warp * P.U + u;
#pragma unroll
            for (int v = 0; v < kUV; v++) tab[u][v] = (u < P.U) ? __ldg(E.tab + (slot * kUV + v) * 32 + lane) : 0.0;
        }
        const int UT = P.UT, US = P.US;
        for (int u = 0; u < UT; u++) {
            const long long slot = slab_base + (long long)warp * P.U + UR + u;
#pragma unroll
            for (int h = 0; h < 3; h++) {
                unsigned r[8];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const double x = __ldg(E.tab + (slot * kUV + 4 * h + i) * 32 + lane);
                    r[2 * i] = (unsigned)__double2loint(x); r[2 * i + 1] = (unsigned)__double2hiint(x);
                }
                tmem_st8(tbase + kTmemUnitCols * u + 8 * h, r);
            }
        }
        if (UT > 0) tmem_wait_st();
        for (int u = 0; u < US; u++) {
            const long long slot = slab_base + (long long)warp * P.U + UR + UT + u;
            for (int v = 0; v < kUV; v++) sTab[((size_t)(warp * US + u) * kUV + v) * 32 + lane] = __ldg(E.tab + (slot * kUV + v) * 32 + lane);
        }
    }
    // F1 operand of a tagged kernel (collective F at trajectory index tix); a later barrier publishes it.
    __device__ __forceinline__ void load_F1(const EqDev &E, long long tix) {
        if (E.Fc_traj == nullptr) return;
        const double *src = E.Fc_traj + tix * P.Nk;
        for (int k = tid; k < P.Nk; k += kThreads) sF1_home[k] = __ldg(src + k);
    }

    // ------------------------------------------------------------------------------------------ steady state
    // solve_steady_state_mutable (src/SteadyStateMemoryEquation.jl:59-100) for a diagonal kernel:
    // iterate F <- (K + gamma)^-1 K F0, K <- kernel(F) until max|F - F_old| <= tolerance.
    __device__ __forceinline__ void steady_state_one(const EqDev &E) {
        oF1 = P.so.F2;
        load_tables(E);
        load_own(E);
        const int k = (tid < nrows) ? row0 + tid : -1;
        const double f0 = (k >= 0) ? E.F0[k] : 0.0, g = (k >= 0) ? E.gamma[k] : 1.0;
        double fold[KPT], out[KPT];
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            const int kk = j * kThreads + tid;
            fold[j] = (kk < P.Nk) ? E.F0[kk] : 0.0;              // Fold = copy(F0) :68
            if (kk < P.Nk) sF2[kk] = fold[j];
        }
        __syncthreads();
        double L1, L2, L3;
        eval_rows(E.tab, L1, L2, L3);                          // K = evaluate_kernel(kernel, F0, t) :66
        exchange(f0, L1, L2, L3, 0, fold, -1, out);
        double K = K_exact(L1, L2, L3), Fn = f0;
        long long iter = 0;
        int status = MCT_OK;
        while (true) {
            iter++;
            if (iter > P.max_iter) { status = MCT_NOT_CONVERGED; break; }   // :75-78
            Fn = (K * f0) / (K + g);                           // :81-84
            const int any = exchange(Fn, 0.0, 0.0, 0.0, X_TRACK, fold, (int)P.so.F2, out);   // find_error(F, Fold); Fold .= F  :90-91
            eval_rows(E.tab, L1, L2, L3);                      // :89 (the error does not depend on it)
            exchange(Fn, L1, L2, L3, 0, fold, -1, out);
            K = K_exact(L1, L2, L3);
            if (aborted()) { status = MCT_CUDA_ERROR; break; }
            if (!any) break;
        }
        if (k >= 0) { E.F_out[k] = Fn; E.K_out[k] = K; }
        if (cta == 0 && tid == 0) { *E.status = status; *E.kernel_evals = iter; }
    }

    // ------------------------------------------------------------------------------------------ time-doubling solve
    __device__ __forceinline__ void solve_one(const EqDev &E) {
        const int N = P.N, n4 = 4 * N, n2 = 2 * N;
        enum { Ft = 0, Kt = 1, FI = 2, KI = 3 };          // F_temp, K_temp, F_I, K_I  (src/TimeDoublingSolver.jl:3-18)
        // history rings of the owned rows: in shared memory when they fit (the memory convolution of a step reads ~2 x 4N slots
        // per row; from the L2-resident global rings that is a chain of ~700-cycle loads), else in global memory
        double *const rbase = P.ring_smem ? (double *)(smem_raw + P.so.ring) : E.ring;
        const int rrows = P.ring_smem ? even_up(P.rows_max) : P.Nk, rk0 = P.ring_smem ? row0 : 0, rn = P.ring_smem ? n4 + 1 : n4;
#define RING(R, k, s) rbase[((size_t)(R) * rrows + ((k) - rk0)) * rn + ((s) - 1)]   /* slot s is 1-based as in the reference */
        oF1 = (E.Fc_traj != nullptr) ? P.so.F1 : P.so.F2;        // tagged kernels: A = V * Fc[iq] * Fs[ip]
        load_tables(E);
        load_own(E);
        const int k = (tid < nrows) ? row0 + tid : -1;    // the wavenumber this thread owns (or -1)
        double fold[KPT], out[KPT], L1, L2, L3;
        // K0 = evaluate_kernel(kernel, F0, 0.0)   (src/MemoryEquation.jl:45)
        for (int kk = tid; kk < P.Nk; kk += kThreads) sF2[kk] = E.F0[kk];
        load_F1(E, 0);
        __syncthreads();
        eval_rows(E.tab, L1, L2, L3);
        exchange(0.0, L1, L2, L3, 0, fold, -1, out);
#ifdef MCT_PROFILE
        {   // isolated cost of the two building blocks (cycles of thread 0, 2000 repetitions each)
            long long c0 = clock64();
            for (int i = 0; i < 2000; i++) exchange(1.0 + i, 0.0, 0.0, 0.0, 0, fold, -1, out);
            long long c1 = clock64();
            for (int i = 0; i < 2000; i++) { double l1, l2, l3; eval_rows(E.tab, l1, l2, l3); __syncthreads(); }
            long long c2 = clock64();
            for (int i = 0; i < 2000; i++) { double l1, l2, l3; eval_rows(E.tab, l1, l2, l3); exchange(l1, l1, l2, l3, X_USE_OFF | X_TRACK, fold, -1, out); }
            long long c3 = clock64();
            prof[10] = (c1 - c0) / 2000; prof[11] = (c2 - c1) / 2000; prof[12] = (c3 - c2) / 2000;
            tlast = clock64();
        }
#endif
        {
            const double K0 = K_exact(L1, L2, L3), f0 = own(OWN_F0);
            if (k >= 0) { sOwn[OWN_K0 * RP + tid] = K0; E.F_out[k] = f0; E.K_out[k] = K0; }     // output row 0 (t = 0)
            // F_old = K0*F0 (:64) for every wavenumber
            exchange(K0 * f0, 0.0, 0.0, 0.0, 0, fold, -1, out);
        }
#pragma unroll
        for (int j = 0; j < KPT; j++) fold[j] = out[j];

        // ---- bootstrap: 2N forward-Euler steps with the trapezoid memory integral
        //      (initialize_F_temp_Euler! :96-109 -> src/EulerSolver.jl:26-64, 92-112); the dF history lives in the
        //      F_I ring until initialize_integrals!.
        const double de = P.dt0 / (4.0 * N);
        if (k >= 0) RING(FI, k, 1) = E.dF0[k];
        for (int it = 1; it <= n2; it++) {
            double Fn = 0.0;
            if (k >= 0) {
                const double *dF = &RING(FI, k, 1);           // dF[j] = dF at Euler step j
                const double K0 = own(OWN_K0), f0 = own(OWN_F0), al = own(OWN_AL), be = own(OWN_BE), ga = own(OWN_GA), de_k = own(OWN_DE);
                double integ;
                // integrand[i] = K_array[i] * dF_reverse[i], K_array[i] = K(t_{i-1}), dF_reverse[i] = dF(t_{it-i})
                if (it == 1) {
                    integ = (K0 * dF[0]) * de;                 // trapezoidal_integration it == 1 (EulerSolver.jl:34-36)
                } else {
                    double rr = (K0 * dF[it - 1]) / 2.0;
                    rr += (RING(Kt, k, it - 1) * dF[0]) / 2.0;
                    for (int i = 2; i <= it - 1; i++) rr += RING(Kt, k, i - 1) * dF[it - i];
                    integ = rr * de;
                }
                const double Fo = (it == 1) ? f0 : RING(Ft, k, it - 1);
                const double dFo = dF[it - 1];
                double dFn;
                if (P.second_order) {                          // Euler_step (EulerSolver.jl:55-58)
                    const double ddF = (be * dFo + ga * Fo + de_k + integ) / (-al);
                    dFn = dFo + de * ddF;
                    Fn = Fo + de * dFn;
                } else {                                       // :60-61
                    dFn = (ga * Fo + de_k + integ) / (-be);
                    Fn = Fo + de * dFn;
                }
                RING(Ft, k, it) = Fn;
                RING(FI, k, it + 1) = dFn;
            }
            load_F1(E, it);
            exchange(Fn, 0.0, 0.0, 0.0, 0, fold, (int)P.so.F2, out);
            eval_rows(E.tab, L1, L2, L3);
            exchange(Fn, L1, L2, L3, 0, fold, -1, out);
            if (k >= 0) {
                const double K = K_exact(L1, L2, L3);
                RING(Kt, k, it) = K;                           // == initialize_K_temp! :155-166
                E.F_out[(size_t)it * P.Nk + k] = Fn; E.K_out[(size_t)it * P.Nk + k] = K;
            }
            if (aborted()) return;
        }
        // initialize_integrals! :174-199 (overwrites the dF scratch, which is no longer needed)
        if (k >= 0) {
            const double F1v = RING(Ft, k, 1), K1v = RING(Kt, k, 1), K2v = RING(Kt, k, 2), f0 = own(OWN_F0);
            double Fprev = F1v, Kprev = K1v;
            RING(FI, k, 1) = (F1v + f0) / 2.0;
            RING(KI, k, 1) = (3.0 * K1v - K2v) / 2.0;
            for (int it = 2; it <= n2; it++) {
                const double Fv = RING(Ft, k, it), Kv = RING(Kt, k, it);
                RING(FI, k, it) = (Fv + Fprev) / 2.0;
                RING(KI, k, it) = (Kv + Kprev) / 2.0;
                Fprev = Fv; Kprev = Kv;
            }
        }
        __syncthreads();

        // ---- main loop: while solver.dt < 2 t_max (:523); block b works with dt = dt0 * 2^b
        long long evals = 0;
        int status = MCT_OK;
        double Dt = P.dt0;
        for (int blk = 0; blk < P.nblocks && status == MCT_OK; blk++, Dt *= 2.0) {
            const double dl = Dt / (4.0 * N);
            // C1 and C2 (:314-315) only change with the block: K_I[1], F_I[1] and dt are fixed inside it.  Every CTA
            // needs rho_m[k] = (c2/c1) {k, 1/k^3, 1/k} of every wavenumber.
            double c1 = 1.0, c2 = 0.0;
            if (k >= 0) {
                c1 = ((2.0 / (dl * dl) * own(OWN_AL) + 3.0 / (2.0 * dl) * own(OWN_BE)) + RING(KI, k, 1)) + own(OWN_GA);    // :314
                c2 = RING(FI, k, 1) - own(OWN_F0);                                              // :315
            }
            exchange(c2 / c1, 0.0, 0.0, 0.0, 0, fold, -1, out);
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                const int kk = j * kThreads + tid;
                if (kk < P.Nk) {
                    const double kv = E.kvec[kk];
                    sRho[kk] = out[j] * kv; sRho[NkP + kk] = out[j] / (kv * kv * kv); sRho[2 * NkP + kk] = out[j] / kv;
                }
            }
            __syncthreads();
            for (int it = n2 + 1; it <= n4 && status == MCT_OK; it++) {
                const int i2 = it / 2;
                const long long row = 1 + (long long)n2 * (blk + 1) + (it - n2 - 1);
                BEACON(10);
                // update_Fuchs_parameters! :300-319 for the owned wavenumbers: one warp per row, lanes over the history
                for (int r = warp; r < nrows; r += kWarps) {
                    const int kr = row0 + r;
                    double s1 = 0.0, s2 = 0.0;
                    for (int j = 2 + lane; j <= i2; j += 32) s1 += (RING(Kt, kr, it - j) - RING(Kt, kr, it - j + 1)) * RING(FI, kr, j);        // :271-280
                    for (int j = 2 + lane; j <= it - i2; j += 32) s2 += RING(KI, kr, j) * (RING(Ft, kr, it - j) - RING(Ft, kr, it - j + 1));  // :281-285
                    s1 = warp_sum(s1); s2 = warp_sum(s2);
                    if (lane == 0) { sC[r] = s1; sC[even_up(P.rows_max) + r] = s2; }
                }
                BEACON(11);
                __syncthreads();
                BEACON(12);
                double Fc3 = 0.0, Finit = 0.0;
                if (k >= 0) {
                    const double al = own(OWN_AL), be = own(OWN_BE), de_k = own(OWN_DE);
                    const double Fm1 = RING(Ft, k, it - 1), Fm2 = RING(Ft, k, it - 2), Fm3 = RING(Ft, k, it - 3);
                    const double KI1 = RING(KI, k, 1), FI1 = RING(FI, k, 1);
                    double v = al * ((5.0 * Fm1 - 4.0 * Fm2 + Fm3) / (dl * dl));                // :259-260
                    v += be * (2.0 / dl * Fm1 - Fm2 / (2.0 * dl));                              // :263-264
                    v -= RING(Kt, k, it - i2) * RING(Ft, k, i2);                                // :267
                    v += RING(Kt, k, it - 1) * FI1;                                             // :268
                    v += KI1 * Fm1;                                                             // :269
                    v += sC[tid]; v += sC[even_up(P.rows_max) + tid];
                    v -= de_k;                                                                  // :286
                    const double c3 = v;
                    Fc3 = c3 / c1;
                    // update_F! :326-337 with the stale K_temp[it]: 2 K0 in the first block (:83), 0 afterwards (:485)
                    const double Kst = (blk == 0) ? (own(OWN_K0) + own(OWN_K0)) : 0.0;
                    Finit = (-(Kst * c2) + c3) / c1;
                }
                load_F1(E, row);
                PROF(9);
                BEACON(13);
                exchange(Finit, 0.0, 0.0, 0.0, 0, fold, (int)P.so.F2, out);
                // fixed-point loop :404-417
                long long iter = 0;
                while (true) {
                    iter++;
                    if (iter > P.max_iter) { status = MCT_NOT_CONVERGED; break; }                // :406-408
                    eval_rows(E.tab, L1, L2, L3);                                               // update_K! :360-369
                    const int kc = (k >= 0) ? k : 0;
                    const double Floc = Fc3 - (sRho[kc] * L1 + sRho[NkP + kc] * L2 + sRho[2 * NkP + kc] * L3);   // update_F! :333-336, CTA-local part
                    const int any = exchange(Floc, L1, L2, L3, X_USE_OFF | X_TRACK, fold, (int)P.so.F2, out);   // find_error + F_old .= F
                    evals++;                                                                    // :416
                    if (aborted()) { status = MCT_CUDA_ERROR; break; }
                    if (!any) break;
                }
                if (status != MCT_OK) break;
                BEACON(14);
                // store the step + update_integrals! :377-384, allocate_results! :429-438 (owners)
                if (k >= 0) {
                    const double Fn = sF2[k], K = K_exact(L1, L2, L3);
                    RING(Ft, k, it) = Fn; RING(Kt, k, it) = K;
                    RING(FI, k, it) = (Fn + RING(Ft, k, it - 1)) / 2.0;
                    RING(KI, k, it) = (K + RING(Kt, k, it - 1)) / 2.0;
                    E.F_out[(size_t)row * P.Nk + k] = Fn; E.K_out[(size_t)row * P.Nk + k] = K;
                }
                __syncthreads();
                PROF(9);
                BEACON(15);
            }
            if (status != MCT_OK) break;
            BEACON(16);
            // new_time_mapping! :448-489 (owners, one warp per row; read everything of a chunk before writing it)
            for (int r = warp; r < nrows; r += kWarps) {
                const int kr = row0 + r;
                for (int jb = 1; jb <= n2; jb += 32) {
                    const int j = jb + lane;
                    double a = 0, b = 0, d = 0, e = 0;
                    if (j <= n2) {
                        a = (RING(FI, kr, 2 * j) + RING(FI, kr, 2 * j - 1)) / 2.0;
                        b = (RING(KI, kr, 2 * j) + RING(KI, kr, 2 * j - 1)) / 2.0;
                        d = RING(Ft, kr, 2 * j);
                        e = RING(Kt, kr, 2 * j);
                    }
                    __syncwarp();
                    if (j <= n2) { RING(FI, kr, j) = a; RING(KI, kr, j) = b; RING(Ft, kr, j) = d; RING(Kt, kr, j) = e; }
                    __syncwarp();
                }
                for (int j = n2 + 1 + lane; j <= n4; j += 32) { RING(FI, kr, j) = 0.0; RING(KI, kr, j) = 0.0; RING(Ft, kr, j) = 0.0; RING(Kt, kr, j) = 0.0; }
            }
            BEACON(17);
            __syncthreads();
            BEACON(18);
        }
        BEACON(19);
        if (cta == 0 && tid == 0) { *E.status = status; *E.kernel_evals = evals; }
#undef RING
    }

    // next equation of this team (dynamic queue); CTA 0 pops it and hands it to the others through a mailbox of
    // self-validating slots (same rotation as the exchange buffers).
    __device__ __forceinline__ int next_equation() {
        int e;
        volatile int *sm = (volatile int *)(sAbort + 1);
        if (P.G == 1) {
            if (tid == 0) sm[0] = atomicAdd(P.queue, 1);
        } else {
            qepoch++;
            double *cur = qbase + (size_t)(qepoch & 3u) * 16, *clr = qbase + (size_t)((qepoch + 2u) & 3u) * 16;
            if (tid == 0) {
                if (cta == 0) { st_relaxed(cur, (double)atomicAdd(P.queue, 1)); st_relaxed_bits(clr, kSentinel); }
                const long long t0 = clock64();
                unsigned spins = 0;
                unsigned long long v;
                while ((v = ld_relaxed_bits(cur)) == kSentinel) { if (spin_check(t0, spins)) { v = 0; break; } }
                sm[0] = (int)bits2d(v);
            }
        }
        __syncthreads();
        e = sm[0];
        __syncthreads();
        return e;
    }
};

template <int KPT, int UR, bool LEAN>
__global__ void __launch_bounds__(kThreads, 1) td_sc_kernel(const SolveParams P) {
    Team<KPT, UR, LEAN> T(P);
    T.tid = threadIdx.x; T.lane = T.tid & 31; T.warp = T.tid >> 5;
    const int team = blockIdx.x / P.G;
    T.cta = blockIdx.x % P.G;
    T.epoch = 0; T.qepoch = 0;
    T.sbase = smem_u32(smem_raw);
    T.xbase = P.xbuf + (size_t)team * P.xstride;
    T.qbase = P.qbox + (size_t)team * 64;
    T.oF1 = P.so.F1;
    T.slab_base = (long long)T.cta * kWarps * P.U;          // first unit slot of this CTA
#ifdef MCT_PROFILE
    for (int i = 0; i < kProfSlots; i++) T.prof[i] = 0;
    T.tlast = clock64();
    const long long t_begin = T.tlast;
    long long g_begin;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g_begin));
#endif
    // per-CTA slice of the layout geometry (identical for every equation of this launch)
    {
        T.row0 = P.cta_meta[4 * T.cta]; T.nrows = P.cta_meta[4 * T.cta + 1];
        T.we0 = P.warp_ent[T.cta * (kWarps + 1) + T.warp];
        T.we_n = P.warp_ent[T.cta * (kWarps + 1) + T.warp + 1] - T.we0;
        const int *eb = P.ent_begin + (size_t)T.cta * (P.rb_max + 1);
        const bool split = 3 * P.rows_max <= 32;
        T.pm = split ? min(T.tid / P.rows_max, 2) : 0;
        T.prow = split ? T.tid % P.rows_max : T.tid;
        const bool sums = T.prow < T.nrows && (!split || T.tid < 3 * P.rows_max);
        T.eb0 = sums ? eb[T.prow / kRB] : 0;
        T.eb1 = sums ? eb[T.prow / kRB + 1] : 0;
#pragma unroll
        for (int j = 0; j < KPT; j++) { const int k = j * kThreads + T.tid; T.owner[j] = (k < P.Nk) ? P.row_owner[k] : 0; }
        // zero padding around the F arrays (window operands of padding lanes)
        for (int i = T.tid; i < kFPad; i += kThreads) {
            double *f1 = (double *)(smem_raw + P.so.F1), *f2 = (double *)(smem_raw + P.so.F2);
            f1[-1 - i] = 0.0; f2[-1 - i] = 0.0; f1[P.Nk + i] = 0.0; f2[P.Nk + i] = 0.0;
        }
        for (int i = T.tid; i < 3 * even_up(P.G); i += kThreads) ((double *)(smem_raw + P.so.off))[i] = 0.0;
        if (T.tid < 24) ((double *)(smem_raw + P.so.wt))[T.tid] = 0.0;
        for (int i = T.tid; i < 3 * even_up(P.Nk); i += kThreads) ((double *)(smem_raw + P.so.rho))[i] = 0.0;
        if (T.tid == 0) *(volatile int *)(smem_raw + P.so.misc) = 0;
        if (P.fabric == 1)
            for (int i = T.tid; i < 4 * P.xbuf_words; i += kThreads) ((unsigned long long *)(smem_raw + P.so.xd))[i] = kSentinel;
    }
    T.tbase = 0;
    if (P.UT > 0) {
        // the whole tensor memory of the SM for this CTA (one CTA per SM: 64 K registers); warp w: lanes 32 (w % 4), column half w / 4
        unsigned *slot = (unsigned *)(smem_raw + P.so.misc) + 2;
        if (T.warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        T.tbase = *(volatile unsigned *)slot + ((unsigned)(32 * (T.warp & 3)) << 16) + (unsigned)((T.warp >> 2) * kTmemColsPerWarp);
    }
    __syncthreads();
    if (P.fabric == 1) cluster_sync_all();        // nobody stores into a peer before its buffers are armed
    while (true) {
        const int e = T.next_equation();
        if (e >= P.n_eq || T.aborted()) break;
        {   // the equation descriptor lives in shared memory for the solve (keeps ~30 registers free)
            const int *src = (const int *)(P.eqs + e);
            int *dst = (int *)(smem_raw + P.so.eq);
            if (T.tid < (int)(sizeof(EqDev) / 4)) dst[T.tid] = src[T.tid];
        }
        __syncthreads();
        const EqDev &E = *(const EqDev *)(smem_raw + P.so.eq);
        if (E.mode == 1) T.steady_state_one(E);
        else T.solve_one(E);
        if (T.aborted()) {
            if (T.cta == 0 && T.tid == 0) *E.status = MCT_CUDA_ERROR;
            break;
        }
        __syncthreads();
    }
    if (P.fabric == 1) cluster_sync_all();        // a CTA must not exit while peers may still store into its shared memory
    if (P.UT > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (T.warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*(volatile unsigned *)((unsigned *)(smem_raw + P.so.misc) + 2)), "n"(kTmemCols) : "memory");
    }
#ifdef MCT_PROFILE
    if (team == 0 && T.tid == 0 && P.prof) {
        long long g_end;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g_end));
        T.prof[kProfSlots - 2] = clock64() - t_begin;
        T.prof[kProfSlots - 1] = g_end - g_begin;
        for (int i = 0; i < kProfSlots; i++) if (i != 28) P.prof[T.cta * kProfSlots + i] = T.prof[i];
    }
#endif
}

#undef NkP
#undef GP
#undef EP
#undef RP


// kernel entry of one (KPT, UR) instantiation; nullptr if the combination is not built
template <int KPT>
void *pick_tr(int UR) {
    switch (UR) {
        case -1: return (void *)td_sc_kernel<KPT, 0, true>;      // lean: all units in tensor memory, L2 fabric
        case 0: return (void *)td_sc_kernel<KPT, 0, false>;
        case 2: return (void *)td_sc_kernel<KPT, 2, false>;
        case 4: return (KPT <= 2) ? (void *)td_sc_kernel<(KPT <= 2 ? KPT : 1), 4, false> : nullptr;
        default: return nullptr;
    }
}

}  // namespace
}  // namespace mct
