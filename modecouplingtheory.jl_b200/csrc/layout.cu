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

__device__ __forceinline__ double desc_coef(int d) {
    const int c = (d >> kDescCoefShift) & 3;
    return c == 0 ? 1.0 : c == 1 ? 2.0 : c == 2 ? -1.0 : -2.0;
}

// One thread per slab slot: tab_m[slot] = coef * V_m[a,b]  (0 where the slot is padding).
__global__ void fill_from_tables_kernel(int Nk, long long slab, const int *desc, const double *V1, const double *V2,
                                        const double *V3, double *tab) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= slab) return;
    const int d = desc[i];
    double v1 = 0.0, v2 = 0.0, v3 = 0.0;
    if (d & kDescValid) {
        const int a = d & kDescIdxMask, b = (d >> kDescIdxBits) & kDescIdxMask;
        const size_t ix = (size_t)a + (size_t)b * Nk;
        const double cf = desc_coef(d);
        v1 = cf * V1[ix]; v2 = cf * V2[ix]; v3 = cf * V3[ix];
    }
    tab[i] = v1; tab[slab + i] = v2; tab[2 * slab + i] = v3;
}

__global__ void fill_from_sk_kernel(int Nk, long long slab, const int *desc, const double *k, const double *Ck, double rho,
                                    double D0, double pi2c, double dk, int tagged, double *tab) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= slab) return;
    const int d = desc[i];
    double v1 = 0.0, v2 = 0.0, v3 = 0.0;
    if (d & kDescValid) {
        const int a = d & kDescIdxMask, b = (d >> kDescIdxBits) & kDescIdxMask;  // A[a,b] = V[a,b] F1[a] F2[b]; q = k[a], p = k[b]
        if (tagged) vertex_tagged(k[b], k[a], Ck[a], D0, rho, pi2c, dk * dk, v1, v2, v3);
        else vertex_collective(k[b], k[a], Ck[b], Ck[a], D0, rho, pi2c, dk, v1, v2, v3);
        const double cf = desc_coef(d);
        v1 = cf * v1; v2 = cf * v2; v3 = cf * v3;
    }
    tab[i] = v1; tab[slab + i] = v2; tab[2 * slab + i] = v3;
}

namespace {

struct Term { int a, b, coef; };   // coef code: 0:+1 1:+2 2:-1 3:-2

// Terms of the increment of row ik (1-based), in a fixed canonical order.
void row_terms(int Nk, int ik, int sym, std::vector<Term> &out) {
    out.clear();
    const int d = ik - 1;
    if (sym) {
        for (int i = 0; i < Nk - d; i++) out.push_back({i, i + d, d > 0 ? 1 : 0});         // A[q,q+d] + A[q+d,q]
        for (int a = 0; a < (ik - 1) / 2; a++) out.push_back({a, ik - 2 - a, 3});          // -(A[a,b] + A[b,a]), a < b
        if (ik % 2 == 0) out.push_back({ik / 2 - 1, ik / 2 - 1, 2});                       // -A[a,a]
    } else {
        for (int i = 0; i < Nk - d; i++) out.push_back({i, i + d, 0});
        if (d > 0) for (int i = 0; i < Nk - d; i++) out.push_back({i + d, i, 0});
        for (int a = 0; a <= ik - 2; a++) out.push_back({a, ik - 2 - a, 2});
    }
}

int row_nterms(int Nk, int ik, int sym) {
    const int d = ik - 1;
    if (sym) return (Nk - d) + (ik - 1) / 2 + (ik % 2 == 0 ? 1 : 0);
    return (Nk - d) * (d > 0 ? 2 : 1) + (ik - 1);
}

// Contiguous partition of the rows into G ranges minimising the largest number of groups; every range non-empty.
void partition_rows(int Nk, int G, int sym, std::vector<int> &row0) {
    std::vector<int> ng(Nk);
    long long total = 0;
    int biggest = 0;
    for (int r = 0; r < Nk; r++) { ng[r] = (row_nterms(Nk, r + 1, sym) + 31) / 32; total += ng[r]; biggest = std::max(biggest, ng[r]); }
    auto ranges_needed = [&](long long cap) {
        int n = 1; long long cur = 0;
        for (int r = 0; r < Nk; r++) { if (cur + ng[r] > cap) { n++; cur = 0; } cur += ng[r]; }
        return n;
    };
    long long lo = std::max<long long>(biggest, (total + G - 1) / G), hi = total;
    while (lo < hi) { const long long mid = (lo + hi) / 2; if (ranges_needed(mid) <= G) hi = mid; else lo = mid + 1; }
    row0.assign(G + 1, Nk);
    int c = 0; long long cur = 0;
    row0[0] = 0;
    for (int r = 0; r < Nk; r++) {
        const int rows_left = Nk - r, ctas_after = G - 1 - c;      // rows still to place (with r), CTAs after this one
        if (cur > 0 && (cur + ng[r] > lo || rows_left <= ctas_after)) { c++; row0[c] = r; cur = 0; }
        cur += ng[r];
    }
    for (int i = c + 1; i <= G; i++) row0[i] = Nk;
}

}  // namespace

int layout_groups_per_warp(int Nk, int G, int sym, int *rows_max) {
    std::vector<int> row0;
    partition_rows(Nk, G, sym, row0);
    long long gmax = 0; int rmax = 0;
    for (int c = 0; c < G; c++) {
        long long g = 0;
        for (int r = row0[c]; r < row0[c + 1]; r++) g += (row_nterms(Nk, r + 1, sym) + 31) / 32;
        gmax = std::max(gmax, g); rmax = std::max(rmax, row0[c + 1] - row0[c]);
    }
    if (rows_max) *rows_max = rmax;
    return (int)((gmax + kWarps - 1) / kWarps);
}

int layout_build_geometry(TeamLayout &L, int Nk, int G, int sym) {
    MCT_REQUIRE(Nk <= (1 << kDescIdxBits), MCT_UNSUPPORTED, "Nk > 2048 is not supported by the persistent solver");
    L.Nk = Nk; L.G = G; L.sym = sym;
    std::vector<int> row0;
    partition_rows(Nk, G, sym, row0);
    for (int c = 0; c < G; c++) MCT_REQUIRE(row0[c + 1] > row0[c], MCT_BAD_ARGUMENT, "team larger than the number of rows");
    int rows_max = 0;
    L.T = layout_groups_per_warp(Nk, G, sym, &rows_max);
    L.rows_max = rows_max;
    const int T = L.T;
    const long long cta_slots = (long long)kWarps * T * 32;
    L.slab_terms = cta_slots * G;
    std::vector<int> desc((size_t)L.slab_terms, 0), meta(4 * G, 0), ent_begin((size_t)G * (rows_max + 1), 0), warp_ent((size_t)G * (kWarps + 1), 0),
        owner(Nk, 0);
    L.ent_max = 1; L.went_max = 1;
    std::vector<Term> terms;
    for (int c = 0; c < G; c++) {
        int *dc = desc.data() + (size_t)c * cta_slots;
        const int nrows = row0[c + 1] - row0[c];
        std::vector<int> group_row;     // local row of every group of this CTA
        for (int lr = 0; lr < nrows; lr++) {
            const int r = row0[c] + lr;
            owner[r] = c;
            row_terms(Nk, r + 1, sym, terms);
            const int gs = (int)group_row.size(), ngr = ((int)terms.size() + 31) / 32;
            for (int g = 0; g < ngr; g++) group_row.push_back(lr);
            for (size_t i = 0; i < terms.size(); i++)
                dc[(size_t)gs * 32 + i] = terms[i].a | (terms[i].b << kDescIdxBits) | (terms[i].coef << kDescCoefShift) | kDescValid;
        }
        const int GC = (int)group_row.size();
        MCT_REQUIRE(GC <= kWarps * T, MCT_CUDA_ERROR, "internal: layout overflow");
        // entries: (warp,row) pairs in group order
        int n_ent = 0, prev_row = -1;
        for (int lr = 0; lr <= nrows; lr++) ent_begin[(size_t)c * (rows_max + 1) + lr] = 0;
        for (int w = 0; w < kWarps; w++) {
            warp_ent[(size_t)c * (kWarps + 1) + w] = n_ent;
            int prev = -1, nw = 0;
            for (int j = 0; j < T; j++) {
                const int g = w * T + j;
                if (g >= GC) break;
                const int lr = group_row[g];
                if (lr != prev) {
                    if (prev >= 0) for (int l = 0; l < 32; l++) dc[(size_t)g * 32 + l] |= kDescNewEntry;
                    if (lr != prev_row) {   // first entry of this row: rows before it (if any were skipped) start here too
                        for (int q = prev_row + 1; q <= lr; q++) ent_begin[(size_t)c * (rows_max + 1) + q] = n_ent;
                        prev_row = lr;
                    }
                    n_ent++; nw++; prev = lr;
                }
            }
            L.went_max = std::max(L.went_max, nw);
        }
        warp_ent[(size_t)c * (kWarps + 1) + kWarps] = n_ent;
        for (int q = prev_row + 1; q <= rows_max; q++) ent_begin[(size_t)c * (rows_max + 1) + q] = n_ent;
        L.ent_max = std::max(L.ent_max, n_ent);
        meta[4 * c] = row0[c]; meta[4 * c + 1] = nrows; meta[4 * c + 2] = n_ent;
    }
    L.xbuf_words = 4 * G + ((Nk + 15) / 16) * 16 + 16;
    auto up = [](int **dst, const std::vector<int> &v) {
        cudaError_t e = cudaMalloc(dst, std::max<size_t>(1, v.size()) * sizeof(int));
        if (e == cudaSuccess) e = cudaMemcpy(*dst, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice);
        return e;
    };
    MCT_CUDA(up(&L.d_desc, desc));
    MCT_CUDA(up(&L.d_cta_meta, meta));
    MCT_CUDA(up(&L.d_ent_begin, ent_begin));
    MCT_CUDA(up(&L.d_warp_ent, warp_ent));
    MCT_CUDA(up(&L.d_row_owner, owner));
    return MCT_OK;
}

void layout_free(TeamLayout &L) {
    cudaFree(L.d_desc); cudaFree(L.d_cta_meta); cudaFree(L.d_ent_begin); cudaFree(L.d_warp_ent); cudaFree(L.d_row_owner);
    L = TeamLayout();
}

int layout_fill_from_tables(const TeamLayout &L, double *d_tab, const double *dV1, const double *dV2, const double *dV3,
                            cudaStream_t s) {
    const unsigned nb = (unsigned)((L.slab_terms + 255) / 256);
    fill_from_tables_kernel<<<nb, 256, 0, s>>>(L.Nk, L.slab_terms, L.d_desc, dV1, dV2, dV3, d_tab);
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
    const unsigned nb = (unsigned)((L.slab_terms + 255) / 256);
    fill_from_sk_kernel<<<nb, 256, 0, s>>>(L.Nk, L.slab_terms, L.d_desc, dk_array, dCk, rho, D0, pi2c, dk, tagged, d_tab);
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
