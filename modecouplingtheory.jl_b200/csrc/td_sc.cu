// td_sc.cu -- host side of the persistent time-doubling solver: shared-memory plan, kernel selection, launch.
// The kernel itself is in td_sc_impl.cuh, instantiated per wavenumbers-per-thread in td_sc_k{1,2,4,8}.cu.
#include <algorithm>

#include "td_sc.cuh"

namespace mct {

void *td_sc_entry_k1(int UR);
void *td_sc_entry_k2(int UR);
void *td_sc_entry_k4(int UR);
void *td_sc_entry_k8(int UR);
SmemOffsets td_sc_plan_smem(int Nk, int G, int rows_max, int ent_max, int lane_cap, int U, int US, int fabric, int xbuf_words, int ring_doubles) {
    SmemOffsets p;
    size_t o = 0;
    const int NkP = even_up(Nk), NkF = even_up(Nk + 2 * kFPad);
    // F arrays: kFPad zero doubles in front and behind (window operands of padding lanes fall there); the offsets
    // point at element 0
    p.F1 = (unsigned)(o + (size_t)kFPad * 8); o += (size_t)NkF * 8;
    p.F2 = (unsigned)(o + (size_t)kFPad * 8); o += (size_t)NkF * 8;
    p.rho = (unsigned)o; o += (size_t)3 * NkP * 8;
    p.off = (unsigned)o; o += (size_t)3 * even_up(G) * 8;
    p.wt = (unsigned)o; o += (size_t)3 * 8 * 8;                      // totals of the (<= 5) warps that scan the CTA totals
    p.own = (unsigned)o; o += (size_t)kOwnFields * even_up(rows_max) * 8;
    p.part = (unsigned)o; o += (size_t)kUV * even_up(ent_max) * 8;
    p.lane = (unsigned)o; o += (size_t)kWarps * lane_cap * kUV * kLaneStride * 8;
    p.C = (unsigned)o; o += (size_t)2 * even_up(rows_max) * 8;
    p.X = (unsigned)o; o += (G == 1 ? (size_t)NkP * 8 : 16);
    p.scan = (unsigned)o; o += (size_t)3 * kWarps * 8;
    p.misc = (unsigned)o; o += 32;
    p.eq = (unsigned)o; o += (sizeof(EqDev) + 15) & ~(size_t)15;
    o = (o + 15) & ~(size_t)15;
    p.xd = (unsigned)o; o += (fabric == 1) ? (size_t)4 * xbuf_words * 8 : 0;     // cluster fabric: the exchange buffers live here
    p.desc = (unsigned)o; o += (size_t)kWarps * (U + 2) * 8;        // operand addresses of every unit (int2)
    o = (o + 15) & ~(size_t)15;
    p.tab = (unsigned)o; o += (size_t)kWarps * US * kUV * 32 * 8;
    p.ring = ring_doubles > 0 ? (unsigned)o : 0u; o += (size_t)ring_doubles * 8;
    p.total = (unsigned)o;
    return p;
}


// doubles of the shared-memory history rings: F_temp, K_temp, F_I, K_I of the owned rows, 4N slots each, rows padded by one
// word against bank conflicts
int td_sc_ring_doubles(const SolveParams &P) { return P.ring_smem ? 4 * even_up(P.rows_max) * (4 * P.N + 1) : 0; }

size_t td_sc_smem_bytes(const SolveParams &P) { return td_sc_plan_smem(P.Nk, P.G, P.rows_max, P.ent_max, P.lane_cap, P.U, P.US, P.fabric, P.xbuf_words, td_sc_ring_doubles(P)).total; }

int td_sc_max_reg_units(int Nk) { return kThreads > 256 ? 0 : ((Nk > 2 * kThreads) ? 2 : 4); }

// which instantiation: -1 = lean (every unit in tensor memory, exchange through L2), otherwise the register-unit count
static int entry_key(const SolveParams &P) {
    return (P.UR == 0 && P.US == 0 && P.UT == P.U && P.fabric == 0 && P.G > 1) ? -1 : P.UR;
}

int td_sc_launch(const SolveParams &P, cudaStream_t stream) {
    const int KPT = (P.Nk + kThreads - 1) / kThreads;
    MCT_REQUIRE(KPT <= kMaxKPT, MCT_UNSUPPORTED, "Nk > 2048 is not supported by the persistent solver");
    MCT_REQUIRE(P.G <= 160, MCT_UNSUPPORTED, "team larger than 160 CTAs");
    MCT_REQUIRE(P.rows_max <= kThreads, MCT_UNSUPPORTED, "more rows per CTA than threads: use a larger team");
    const size_t smem = td_sc_smem_bytes(P);
    void *fn = nullptr;
    const int key = entry_key(P);
    if (KPT <= 1) fn = td_sc_entry_k1(key);
    else if (KPT <= 2) fn = td_sc_entry_k2(key);
    else if (KPT <= 4) fn = td_sc_entry_k4(key);
    else fn = td_sc_entry_k8(key);
    MCT_REQUIRE(fn != nullptr, MCT_BAD_ARGUMENT, "internal: no kernel instantiation for this UR");
    MCT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SolveParams Pc = P;
    Pc.so = td_sc_plan_smem(P.Nk, P.G, P.rows_max, P.ent_max, P.lane_cap, P.U, P.US, P.fabric, P.xbuf_words, td_sc_ring_doubles(P));
    void *args[] = {(void *)&Pc};
    dim3 grid(P.nteams * P.G), block(kThreads);
    if (P.fabric == 1) {
        // one thread-block cluster per team: the hardware co-schedules the CTAs of a cluster, teams are independent
        if (P.G > 8) MCT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)P.G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        MCT_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
    } else if (P.G > 1) {
        // the exchange spins on other CTAs of the team: a cooperative launch guarantees co-residency or fails.
        MCT_CUDA(cudaLaunchCooperativeKernel(fn, grid, block, args, smem, stream));
    } else {
        MCT_CUDA(cudaLaunchKernel(fn, grid, block, args, smem, stream));
    }
    count_launch();
    return MCT_OK;
}

}  // namespace mct

namespace mct {
// how many clusters of G CTAs with this much shared memory can be resident at once (0 if the launch is impossible)
int td_sc_max_clusters(const SolveParams &P) {
    const int KPT = (P.Nk + kThreads - 1) / kThreads;
    const int key = entry_key(P);
    void *fn = KPT <= 1 ? td_sc_entry_k1(key) : KPT <= 2 ? td_sc_entry_k2(key) : KPT <= 4 ? td_sc_entry_k4(key) : td_sc_entry_k8(key);
    if (!fn) return 0;
    const size_t smem = td_sc_smem_bytes(P);
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (P.G > 8 && cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(P.G * 64); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)P.G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
}  // namespace mct
