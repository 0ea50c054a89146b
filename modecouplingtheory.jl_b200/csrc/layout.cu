// layout.cu -- builds the team layout of the vertex tables (see layout.cuh).
#include <algorithm>
#include <cmath>

#include "layout.cuh"

namespace mct {

// Vertex expressions of the reference constructors, evaluated strictly left to right as Julia does
// (n-ary * and / are left-associative; the file is compiled with -fmad=false so nothing is contracted).
//   collective: src/Kernels/ModeCouplingKernels.jl:116-126      tagged: :413-422
__device__ __forceinline__ void vertex_collective(double p, double q, double cp, double cq, double D0, double rho,
                                                  double pi2x8, double dk, double &v1, double &v2, double &v3) {
    double s = cp + cq, dd = q * q - p * p, dc = cq - cp, dc2 = cq * cq - cp * cp;
    double pq = p * q;
    v1 = pq * (s * s) / 4.0 * D0 * rho / pi2x8 * dk * dk;
    v2 = pq * (dd * dd) * (dc * dc) / 4.0 * D0 * rho / pi2x8 * dk * dk;
    v3 = pq * dd * dc2 / 2.0 * D0 * rho / pi2x8 * dk * dk;
}
__device__ __forceinline__ void vertex_tagged(double p, double q, double cq, double D0, double rho, double pi2x4,
                                              double dk2, double &v1, double &v2, double &v3) {
    double dd = q * q - p * p, c2 = cq * cq;
    double pq = p * q;
    v1 = pq * c2 / 4.0 * D0 * rho / pi2x4 * dk2;
    v2 = pq * (dd * dd) * c2 / 4.0 * D0 * rho / pi2x4 * dk2;
    v3 = pq * dd * c2 / 2.0 * D0 * rho / pi2x4 * dk2;
}

__global__ void direct_correlation_kernel(int Nk, const double *Sk, double rho, double *Ck) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Nk) Ck[i] = (Sk[i] - 1.0) / (rho * Sk[i]);  // :105 / :402
}

__global__ void raw_tables_kernel(int Nk, const double *k, const double *Ck, double rho, double D0, double pi2c, double dk,
                                  int tagged, double *V1, double *V2, double *V3) {
    int iq = blockIdx.x * blockDim.x + threadIdx.x;  // contiguous (first) index of the column-major tables
    int ip = blockIdx.y;
    if (iq >= Nk) return;
    double v1, v2, v3;
    if (tagged) vertex_tagged(k[ip], k[iq], Ck[iq], D0, rho, pi2c, dk * dk, v1, v2, v3);
    else vertex_collective(k[ip], k[iq], Ck[ip], Ck[iq], D0, rho, pi2c, dk, v1, v2, v3);
    size_t ix = (size_t)iq + (size_t)ip * Nk;
    V1[ix] = v1; V2[ix] = v2; V3[ix] = v3;
}

// The term of (unit, lane, row r): indices (a, b) into V and its coefficient; false if the slot is padding.
__device__ __forceinline__ bool unit_term(const UnitInfo &u, int Nk, int sym, int lane, int r, int &a, int &b, double &coef) {
    const int d = u.d0 + r, pos = u.a0 + lane;
    if (d >= u.dend || d >= Nk) return false;
    if (u.type == 0) {                       // A[q, q+d] (+ A[q+d, q] merged when symmetric)
        if (pos >= Nk - d) return false;
        a = pos; b = pos + d; coef = (sym && d > 0) ? 2.0 : 1.0;
        return true;
    }
    if (u.type == 2) {                       // A[q+d, q]   (non-symmetric tables, d > 0)
        if (d == 0 || pos >= Nk - d) return false;
        a = pos + d; b = pos; coef = 1.0;
        return true;
    }
    // antidiagonal of row ik = d+1:  -sum_{iq<ik} A[iq, ik-iq]  ->  0-based (pos, d-1-pos), pos <= d-1
    const int bb = d - 1 - pos;
    if (bb < 0) return false;
    if (sym) {                               // pairs (a,b),(b,a) merged: a < b with -2, a == b with -1
        if (pos > bb) return false;
        coef = (pos < bb) ? -2.0 : -1.0;
    } else coef = -1.0;
    a = pos; b = bb;
    return true;
}

// One thread per (unit, lane): the 12 table values tab[(unit*12 + r*3 + m)*32 + lane] = coef * V_m[a,b] (0 if padding)
__global__ void fill_from_tables_kernel(int Nk, int sym, long long n_units, const UnitInfo *uinfo, const double *V1, const double *V2,
                                        const double *V3, double *tab) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_units * 32) return;
    const long long unit = i >> 5; const int lane = (int)(i & 31);
    const UnitInfo u = uinfo[unit];
    for (int r = 0; r < kRB; r++) {
        int a = 0, b = 0; double cf = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
        if (u.type >= 0 && unit_term(u, Nk, sym, lane, r, a, b, cf)) {
            const size_t ix = (size_t)a + (size_t)b * Nk;
            v1 = cf * V1[ix]; v2 = cf * V2[ix]; v3 = cf * V3[ix];
        }
        double *t = tab + (unit * kUV + r * 3) * 32 + lane;
        t[0] = v1; t[32] = v2; t[64] = v3;
    }
}

__global__ void fill_from_sk_kernel(int Nk, int sym, long long n_units, const UnitInfo *uinfo, const double *k, const double *Ck,
                                    double rho, double D0, double pi2c, double dk, int tagged, double *tab) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_units * 32) return;
    const long long unit = i >> 5; const int lane = (int)(i & 31);
    const UnitInfo u = uinfo[unit];
    for (int r = 0; r < kRB; r++) {
        int a = 0, b = 0; double cf = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
        if (u.type >= 0 && unit_term(u, Nk, sym, lane, r, a, b, cf)) {   // A[a,b] = V[a,b] F1[a] F2[b]; q = k[a], p = k[b]
            if (tagged) vertex_tagged(k[b], k[a], Ck[a], D0, rho, pi2c, dk * dk, v1, v2, v3);
            else vertex_collective(k[b], k[a], Ck[b], Ck[a], D0, rho, pi2c, dk, v1, v2, v3);
            v1 = cf * v1; v2 = cf * v2; v3 = cf * v3;
        }
        double *t = tab + (unit * kUV + r * 3) * 32 + lane;
        t[0] = v1; t[32] = v2; t[64] = v3;
    }
}

namespace {

// units of the row block starting at row d0 (rows d0 .. min(d0+4, dend)-1), appended in a fixed order
void block_units(int Nk, int sym, int d0, int dend, std::vector<UnitInfo> &out) {
    const int nd = (Nk - d0 + 31) / 32;                               // longest diagonal run of the block: row d0
    for (int i = 0; i < nd; i++) out.push_back({0, d0, dend, 32 * i});
    const int dlast = std::min(d0 + kRB, dend) - 1;
    if (!sym && dlast > 0) for (int i = 0; i < nd; i++) out.push_back({2, d0, dend, 32 * i});
    // antidiagonal: positions 0 .. d-1 (non-symmetric) or 0 .. (d-1)/2 (symmetric) for the longest row dlast
    const int na = sym ? (dlast - 1) / 2 + 1 : dlast;
    if (dlast >= 1) for (int i = 0; i < (na + 31) / 32; i++) out.push_back({1, d0, dend, 32 * i});
}

int block_unit_count(int Nk, int sym, int d0, int dend) {
    const int nd = (Nk - d0 + 31) / 32;
    const int dlast = std::min(d0 + kRB, dend) - 1;
    int n = nd;
    if (!sym && dlast > 0) n += nd;
    if (dlast >= 1) n += ((sym ? (dlast - 1) / 2 + 1 : dlast) + 31) / 32;
    return n;
}

// Contiguous partition of the rows into G ranges (every range non-empty) minimising the largest unit count.  Row blocks
// start at the first row of a range, so the cost of a range is evaluated as a whole.
long long range_units(int Nk, int sym, int r0, int r1) {
    long long n = 0;
    for (int d0 = r0; d0 < r1; d0 += kRB) n += block_unit_count(Nk, sym, d0, r1);
    return n;
}
// time a CTA spends on a range: its number of units (since round 2 every unit costs the same: rows that do not exist in a
// trailing block are no longer skipped, their table values are zero)
long long range_cost(int Nk, int sym, int r0, int r1) { return range_units(Nk, sym, r0, r1); }

void partition_rows(int Nk, int G, int sym, std::vector<int> &row0) {
    auto fits = [&](long long cap, std::vector<int> *out) {
        int c = 0, r = 0;
        if (out) out->assign(G + 1, Nk);
        while (r < Nk) {
            if (c >= G) return false;
            if (out) (*out)[c] = r;
            // largest end such that the range costs <= cap, leaving at least one row for every later CTA
            int hi = Nk - (G - 1 - c), e = r + 1;
            if (range_cost(Nk, sym, r, e) > cap) return false;
            int lo = e;
            while (lo < hi) { const int mid = (lo + hi + 1) / 2; if (range_cost(Nk, sym, r, mid) <= cap) lo = mid; else hi = mid - 1; }
            r = lo; c++;
        }
        if (out) for (int i = c; i <= G; i++) (*out)[i] = Nk;
        return c <= G;
    };
    long long lo = 1, hi = range_cost(Nk, sym, 0, Nk);
    while (lo < hi) { const long long mid = (lo + hi) / 2; if (fits(mid, nullptr)) hi = mid; else lo = mid + 1; }
    fits(lo, &row0);
    // ranges the greedy packing left empty at the end: give them one row each from their predecessor
    for (int c = G - 1; c >= 1; c--)
        if (row0[c] >= row0[c + 1]) row0[c] = row0[c + 1] - 1;
}

}  // namespace

long long layout_total_units(int Nk, int sym) { return range_units(Nk, sym, 0, Nk); }

int layout_build_geometry(TeamLayout &L, int Nk, int G, int sym) {
    MCT_REQUIRE(Nk <= (1 << kUA0Bits), MCT_UNSUPPORTED, "Nk > 2048 is not supported by the persistent solver");
    L.Nk = Nk; L.G = G; L.sym = sym;
    std::vector<int> row0;
    partition_rows(Nk, G, sym, row0);
    for (int c = 0; c < G; c++) MCT_REQUIRE(row0[c + 1] > row0[c], MCT_BAD_ARGUMENT, "team larger than the number of rows");
    long long umax = 0;
    L.rows_max = 0; L.rb_max = 0;
    for (int c = 0; c < G; c++) {
        umax = std::max(umax, range_units(Nk, sym, row0[c], row0[c + 1]));
        L.rows_max = std::max(L.rows_max, row0[c + 1] - row0[c]);
        L.rb_max = std::max(L.rb_max, (row0[c + 1] - row0[c] + kRB - 1) / kRB);
    }
    const int U = (int)((umax + kWarps - 1) / kWarps);
    L.U = U;
    L.n_units = (long long)G * kWarps * U;
    std::vector<int> udesc((size_t)L.n_units, 0), meta(4 * G, 0), ent_begin((size_t)G * (L.rb_max + 1), 0), warp_ent((size_t)G * (kWarps + 1), 0),
        owner(Nk, 0);
    std::vector<UnitInfo> uinfo((size_t)L.n_units, UnitInfo{-1, 0, 0, 0});
    L.ent_max = 1; L.went_max = 1;
    for (int c = 0; c < G; c++) {
        const int r0 = row0[c], r1 = row0[c + 1];
        for (int r = r0; r < r1; r++) owner[r] = c;
        std::vector<UnitInfo> units; std::vector<int> unit_rb;
        for (int d0 = r0, q = 0; d0 < r1; d0 += kRB, q++) {
            const size_t before = units.size();
            block_units(Nk, sym, d0, r1, units);
            for (size_t i = before; i < units.size(); i++) unit_rb.push_back(q);
        }
        const int nu = (int)units.size();
        MCT_REQUIRE(nu <= kWarps * U, MCT_CUDA_ERROR, "internal: layout overflow");
        const int nrb = (r1 - r0 + kRB - 1) / kRB;
        // deal the units to the warps in order, as evenly as possible
        int n_ent = 0, prev_rb = -1, next = 0;
        int *eb = ent_begin.data() + (size_t)c * (L.rb_max + 1);
        for (int w = 0; w < kWarps; w++) {
            warp_ent[(size_t)c * (kWarps + 1) + w] = n_ent;
            const int cnt = nu / kWarps + (w < nu % kWarps ? 1 : 0);
            int prev = -1, nw = 0;
            for (int j = 0; j < cnt; j++, next++) {
                const UnitInfo &u = units[next];
                const int q = unit_rb[next];
                const size_t slot = ((size_t)c * kWarps + w) * U + j;
                uinfo[slot] = u;
                // window start of lane 0 (see layout.cuh), biased so that it is never negative
                const int w0 = (u.type == 1) ? u.d0 - 1 - u.a0 : u.d0 + u.a0;
                int d = u.a0 | ((w0 + kFPad) << kUW0Shift) | kUValid | ((std::min(kRB, u.dend - u.d0) - 1) << kURowsShift);
                if (u.type == 1) d |= kUNeg;
                if (u.type == 2) d |= kUSwap;
                if (q != prev) {
                    if (prev >= 0) d |= kUNewEntry;
                    if (q != prev_rb) { for (int x = prev_rb + 1; x <= q; x++) eb[x] = n_ent; prev_rb = q; }
                    n_ent++; nw++; prev = q;
                }
                udesc[slot] = d;
            }
            L.went_max = std::max(L.went_max, nw);
        }
        warp_ent[(size_t)c * (kWarps + 1) + kWarps] = n_ent;
        for (int x = prev_rb + 1; x <= L.rb_max; x++) eb[x] = n_ent;
        (void)nrb;
        L.ent_max = std::max(L.ent_max, n_ent);
        meta[4 * c] = r0; meta[4 * c + 1] = r1 - r0; meta[4 * c + 2] = n_ent;
    }
    L.xrows = std::max(4 * G, 3 * (((G + 15) / 16) * 16));     // totals: [G][4] (cluster fabric) or [3][G rounded to a line] (L2 fabric)
    L.xbuf_words = L.xrows + ((Nk + 15) / 16) * 16 + 16;
    auto up = [](int **dst, const std::vector<int> &v) {
        cudaError_t e = cudaMalloc(dst, std::max<size_t>(1, v.size()) * sizeof(int));
        if (e == cudaSuccess) e = cudaMemcpy(*dst, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice);
        return e;
    };
    MCT_CUDA(up(&L.d_udesc, udesc));
    MCT_CUDA(up(&L.d_cta_meta, meta));
    MCT_CUDA(up(&L.d_ent_begin, ent_begin));
    MCT_CUDA(up(&L.d_warp_ent, warp_ent));
    MCT_CUDA(up(&L.d_row_owner, owner));
    MCT_CUDA(cudaMalloc(&L.d_uinfo, uinfo.size() * sizeof(UnitInfo)));
    MCT_CUDA(cudaMemcpy(L.d_uinfo, uinfo.data(), uinfo.size() * sizeof(UnitInfo), cudaMemcpyHostToDevice));
    return MCT_OK;
}

void layout_free(TeamLayout &L) {
    cudaFree(L.d_udesc); cudaFree(L.d_uinfo); cudaFree(L.d_cta_meta); cudaFree(L.d_ent_begin); cudaFree(L.d_warp_ent); cudaFree(L.d_row_owner);
    L = TeamLayout();
}

int layout_fill_from_tables(const TeamLayout &L, double *d_tab, const double *dV1, const double *dV2, const double *dV3,
                            cudaStream_t s) {
    const unsigned nb = (unsigned)((L.n_units * 32 + 255) / 256);
    fill_from_tables_kernel<<<nb, 256, 0, s>>>(L.Nk, L.sym, L.n_units, L.d_uinfo, dV1, dV2, dV3, d_tab);
    count_launch();
    MCT_CUDA(cudaGetLastError());
    return MCT_OK;
}

static void vertex_constants(double kBT, double m, int tagged, double &D0, double &pi2c) {
    D0 = kBT / m;
    pi2c = (tagged ? 4.0 : 8.0) * (M_PI * M_PI);
}

int layout_fill_from_sk(const TeamLayout &L, double *d_tab, const double *dk_array, const double *dCk, double rho, double kBT,
                        double m, int tagged, cudaStream_t s) {
    double D0, pi2c;
    vertex_constants(kBT, m, tagged, D0, pi2c);
    double k01[2];
    MCT_CUDA(cudaMemcpyAsync(k01, dk_array, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    MCT_CUDA(cudaStreamSynchronize(s));
    double dk = k01[1] - k01[0];
    const unsigned nb = (unsigned)((L.n_units * 32 + 255) / 256);
    fill_from_sk_kernel<<<nb, 256, 0, s>>>(L.Nk, L.sym, L.n_units, L.d_uinfo, dk_array, dCk, rho, D0, pi2c, dk, tagged, d_tab);
    count_launch();
    MCT_CUDA(cudaGetLastError());
    return MCT_OK;
}

int tables_from_sk(int Nk, const double *dk_array, const double *dSk, double *dCk, double rho, double kBT, double m, int tagged,
                   double *dV1, double *dV2, double *dV3, cudaStream_t s) {
    double D0, pi2c;
    vertex_constants(kBT, m, tagged, D0, pi2c);
    double k01[2];
    MCT_CUDA(cudaMemcpyAsync(k01, dk_array, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    MCT_CUDA(cudaStreamSynchronize(s));
    double dk = k01[1] - k01[0];
    direct_correlation_kernel<<<(Nk + 127) / 128, 128, 0, s>>>(Nk, dSk, rho, dCk);
    count_launch();
    if (dV1) {
        dim3 grid((Nk + 127) / 128, Nk);
        raw_tables_kernel<<<grid, 128, 0, s>>>(Nk, dk_array, dCk, rho, D0, pi2c, dk, tagged, dV1, dV2, dV3);
        count_launch();
    }
    MCT_CUDA(cudaGetLastError());
    return MCT_OK;
}

}  // namespace mct
