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

__global__ void fill_from_tables_kernel(int Nk, long long total, const Piece *pieces, const double *V1, const double *V2,
                                        const double *V3, double *tab) {
    const Piece pc = pieces[blockIdx.x];
    for (int i = threadIdx.x; i < pc.len; i += blockDim.x) {
        size_t ix = (size_t)(pc.a0 + i) + (size_t)(pc.b0 + pc.bdir * i) * Nk;
        tab[pc.off + i] = pc.coef * V1[ix];
        tab[total + pc.off + i] = pc.coef * V2[ix];
        tab[2 * total + pc.off + i] = pc.coef * V3[ix];
    }
}

__global__ void fill_from_sk_kernel(int Nk, long long total, const Piece *pieces, const double *k, const double *Ck,
                                    double rho, double D0, double pi2c, double dk, int tagged, double *tab) {
    const Piece pc = pieces[blockIdx.x];
    for (int i = threadIdx.x; i < pc.len; i += blockDim.x) {
        int a = pc.a0 + i, b = pc.b0 + pc.bdir * i;  // A[a,b] = V[a,b] F1[a] F2[b]; q = k[a], p = k[b]
        double v1, v2, v3;
        if (tagged) vertex_tagged(k[b], k[a], Ck[a], D0, rho, pi2c, dk * dk, v1, v2, v3);
        else vertex_collective(k[b], k[a], Ck[b], Ck[a], D0, rho, pi2c, dk, v1, v2, v3);
        tab[pc.off + i] = pc.coef * v1;
        tab[total + pc.off + i] = pc.coef * v2;
        tab[2 * total + pc.off + i] = pc.coef * v3;
    }
}

static void row_pieces(int Nk, int ik /*1-based*/, int sym, std::vector<Piece> &out) {
    int d = ik - 1;
    if (sym) {
        out.push_back({0, d, +1, Nk - d, 0, d > 0 ? 2.0 : 1.0});
        int nhalf = (ik - 1) / 2;
        if (nhalf > 0) out.push_back({0, ik - 2, -1, nhalf, 0, -2.0});
        if (ik % 2 == 0) out.push_back({ik / 2 - 1, ik / 2 - 1, -1, 1, 0, -1.0});
    } else {
        out.push_back({0, d, +1, Nk - d, 0, 1.0});
        if (d > 0) out.push_back({d, 0, +1, Nk - d, 0, 1.0});
        if (ik >= 2) out.push_back({0, ik - 2, -1, ik - 1, 0, -1.0});
    }
}

int layout_build_geometry(TeamLayout &L, int Nk, int G, int sym, int smem_terms) {
    L.Nk = Nk; L.G = G; L.sym = sym; L.smem_terms = smem_terms;
    // All terms in row order form one sequence; CTA c owns the contiguous range [cta_off[c], cta_off[c+1]) of it, so
    // every CTA streams the same number of terms.  A row cut by a CTA boundary contributes a partial sum to the
    // earlier CTA's block total; the row itself (its T, K, F, history) belongs to the CTA holding its LAST term.
    struct RowPiece { Piece p; int row; };
    std::vector<RowPiece> all;
    std::vector<long long> row_end(Nk);
    long long total = 0;
    for (int r = 0; r < Nk; r++) {
        std::vector<Piece> ps;
        row_pieces(Nk, r + 1, sym, ps);
        for (auto &p : ps) { p.off = total; all.push_back({p, r}); total += p.len; }
        row_end[r] = total;
    }
    L.total_terms = (total + 1) / 2 * 2;
    L.pieces.clear();
    for (auto &rp : all) L.pieces.push_back(rp.p);
    L.n_pieces = (int)L.pieces.size();
    L.cta_off.assign(G + 1, 0);
    for (int c = 1; c < G; c++) L.cta_off[c] = std::min<long long>(total, (total * c / G + 1) / 2 * 2);
    L.cta_off[G] = total;

    // first pass: rows (row parts) per CTA
    std::vector<std::vector<int>> cta_rows(G);      // row index of every row part, in order
    {
        size_t pi = 0;
        for (int c = 0; c < G; c++) {
            const long long lo = L.cta_off[c], hi = L.cta_off[c + 1];
            while (pi < all.size() && all[pi].p.off + all[pi].p.len <= lo) pi++;
            for (size_t q = pi; q < all.size() && all[q].p.off < hi; q++)
                if (cta_rows[c].empty() || cta_rows[c].back() != all[q].row) cta_rows[c].push_back(all[q].row);
        }
    }
    L.rows_max = 1;
    for (int c = 0; c < G; c++) L.rows_max = std::max(L.rows_max, (int)cta_rows[c].size());

    std::vector<Task> tasks;
    std::vector<int> task_begin(G * kWarps + 1, 0);
    std::vector<int> rows(G * L.rows_max, -1), row_slot(G * (L.rows_max + 1), 0);
    L.slots_max = 1; L.tasks_max = 1;
    size_t pi = 0;
    for (int c = 0; c < G; c++) {
        const long long lo = L.cta_off[c], hi = L.cta_off[c + 1];
        const int Lc = (int)(hi - lo);
        int chunk = (((Lc + kWarps - 1) / kWarps) + 31) / 32 * 32;
        if (chunk == 0) chunk = 32;
        for (size_t lr = 0; lr < cta_rows[c].size(); lr++) {
            const int r = cta_rows[c][lr];
            // owned (its last term is here): store ik; partial row whose tail lives in a later CTA: store -(ik+2)
            rows[c * L.rows_max + lr] = (row_end[r] <= hi) ? r : -(r + 2);
        }
        int slot = -1, cur_lr = -1, last_row = -1, last_warp = -1;
        std::vector<std::vector<Task>> per_warp(kWarps);
        while (pi < all.size() && all[pi].p.off + all[pi].p.len <= lo) pi++;
        for (size_t q = pi; q < all.size() && all[q].p.off < hi; q++) {
            const Piece &pc = all[q].p;
            const int row = all[q].row;
            // part of the piece inside this CTA
            int pos = (int)std::max<long long>(0, lo - pc.off);
            const int pend = (int)std::min<long long>(pc.len, hi - pc.off);
            while (pos < pend) {
                const int t0 = (int)(pc.off + pos - lo);       // offset inside the CTA slab
                const int w = std::min(t0 / chunk, kWarps - 1);
                const int wend = (w == kWarps - 1) ? Lc : (w + 1) * chunk;
                int n = std::min(pend - pos, wend - t0);
                const int where = (t0 >= smem_terms) ? 1 : 0;
                if (!where && t0 + n > smem_terms) n = smem_terms - t0;
                if (row != last_row || w != last_warp) {        // new (warp,row) group -> new partial-sum slot
                    if (last_warp >= 0) per_warp[last_warp].back().flush = 1;
                    slot++;
                    last_warp = w;
                }
                if (row != last_row) { cur_lr++; row_slot[c * (L.rows_max + 1) + cur_lr] = slot; last_row = row; }
                Task t{pc.a0 + pos, pc.b0 + pc.bdir * pos, pc.bdir, n, t0, slot, where, 0};
                per_warp[w].push_back(t);
                pos += n;
            }
        }
        if (last_warp >= 0) per_warp[last_warp].back().flush = 1;
        slot++;   // number of slots
        for (int r = cur_lr + 1; r <= L.rows_max; r++) row_slot[c * (L.rows_max + 1) + r] = slot;
        L.slots_max = std::max(L.slots_max, slot);
        for (int w = 0; w < kWarps; w++) {
            task_begin[c * kWarps + w] = (int)tasks.size();
            L.tasks_max = std::max(L.tasks_max, (int)per_warp[w].size());
            tasks.insert(tasks.end(), per_warp[w].begin(), per_warp[w].end());
        }
    }
    task_begin[G * kWarps] = (int)tasks.size();

    // Exchange buffer geometry (see td_sc.cu): every 128-byte line has exactly ONE writer CTA.  Line c (c < G)
    // carries the three block totals of CTA c (+ one spare word); after these G lines, CTA c owns LF lines holding
    // the per-wavenumber values of the rows it owns, in row order.
    {
        const int LF = (L.rows_max + 15) / 16;
        std::vector<int> kslot(Nk, 0);
        for (int c = 0; c < G; c++) {
            int idx = 0;
            for (int lr = 0; lr < L.rows_max; lr++) {
                const int r = rows[c * L.rows_max + lr];
                if (r >= 0) kslot[r] = 16 * G + 16 * LF * c + idx++;
            }
        }
        L.xbuf_words = 16 * G + 16 * LF * G;
        MCT_CUDA(cudaMalloc(&L.d_kslot, Nk * sizeof(int)));
        MCT_CUDA(cudaMemcpy(L.d_kslot, kslot.data(), Nk * sizeof(int), cudaMemcpyHostToDevice));
    }
    MCT_CUDA(cudaMalloc(&L.d_tasks, std::max<size_t>(1, tasks.size()) * sizeof(Task)));
    MCT_CUDA(cudaMemcpy(L.d_tasks, tasks.data(), tasks.size() * sizeof(Task), cudaMemcpyHostToDevice));
    MCT_CUDA(cudaMalloc(&L.d_task_begin, task_begin.size() * sizeof(int)));
    MCT_CUDA(cudaMemcpy(L.d_task_begin, task_begin.data(), task_begin.size() * sizeof(int), cudaMemcpyHostToDevice));
    MCT_CUDA(cudaMalloc(&L.d_rows, rows.size() * sizeof(int)));
    MCT_CUDA(cudaMemcpy(L.d_rows, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice));
    MCT_CUDA(cudaMalloc(&L.d_row_slot, row_slot.size() * sizeof(int)));
    MCT_CUDA(cudaMemcpy(L.d_row_slot, row_slot.data(), row_slot.size() * sizeof(int), cudaMemcpyHostToDevice));
    MCT_CUDA(cudaMalloc(&L.d_cta_off, L.cta_off.size() * sizeof(long long)));
    MCT_CUDA(cudaMemcpy(L.d_cta_off, L.cta_off.data(), L.cta_off.size() * sizeof(long long), cudaMemcpyHostToDevice));
    MCT_CUDA(cudaMalloc(&L.d_pieces, L.pieces.size() * sizeof(Piece)));
    MCT_CUDA(cudaMemcpy(L.d_pieces, L.pieces.data(), L.pieces.size() * sizeof(Piece), cudaMemcpyHostToDevice));
    return MCT_OK;
}

void layout_free(TeamLayout &L) {
    cudaFree(L.d_tab); cudaFree(L.d_tasks); cudaFree(L.d_task_begin); cudaFree(L.d_rows); cudaFree(L.d_row_slot);
    cudaFree(L.d_cta_off); cudaFree(L.d_pieces); cudaFree(L.d_kslot);
    L = TeamLayout();
}

int layout_fill_from_tables(const TeamLayout &L, double *d_tab, const double *dV1, const double *dV2, const double *dV3,
                            cudaStream_t s) {
    fill_from_tables_kernel<<<L.n_pieces, 128, 0, s>>>(L.Nk, L.total_terms, L.d_pieces, dV1, dV2, dV3, d_tab);
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
    fill_from_sk_kernel<<<L.n_pieces, 128, 0, s>>>(L.Nk, L.total_terms, L.d_pieces, dk_array, dCk, rho, D0, pi2c, dk, tagged,
                                                   d_tab);
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
    if (dV1) {
        dim3 grid((Nk + 127) / 128, Nk);
        raw_tables_kernel<<<grid, 128, 0, s>>>(Nk, dk_array, dCk, rho, D0, pi2c, dk, tagged, dV1, dV2, dV3);
    }
    MCT_CUDA(cudaGetLastError());
    return MCT_OK;
}

}  // namespace mct
