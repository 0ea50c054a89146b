// abi.cu -- the extern "C" boundary declared in include/mct.h.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <functional>
#include <unistd.h>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "eval.cuh"
#include "msd.cuh"
#include "layout.cuh"
#include "td_grid.cuh"
#include "td_sc.cuh"

namespace mct {

static thread_local std::string g_last_error;
static std::atomic<long> g_launches{0};
void count_launch(int n) { g_launches += n; }
void set_error(const std::string &msg) { g_last_error = msg; }
int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    g_last_error = buf;
    return MCT_CUDA_ERROR;
}

enum KernelType { K_SC3D = 1, K_TAGGED3D = 2, K_MC3D = 3, K_DDIM = 5 };

// geometry of a team layout is shared by every kernel handle with the same (device, Nk, G, sym)
using GeomKey = std::tuple<int, int, int, int>;
static std::mutex g_geom_mutex;
static std::map<GeomKey, TeamLayout *> g_geoms;
static std::mutex g_stream_mutex;
static std::vector<cudaStream_t> g_stream_pool[64];
static unsigned g_stream_next[64];
constexpr int kStreamsPerDevice = 8;


// ---- device memory arena -----------------------------------------------------------------------------------------------
// cudaMalloc / cudaFree synchronise with the driver and cost milliseconds each on these boxes (measured: 12 ms per kernel
// handle created, 14 ms per handle destroyed -- 10 s of a 512-point sweep).  Every buffer of this file therefore comes from
// large chunks that are obtained once and never returned: blocks are carved off a chunk and, when released, kept in
// exact-size bins (a sweep asks for the same few sizes again and again).  Buffers exported through CUDA IPC (the k-shard
// exchange) bypass the arena: an IPC handle names a whole allocation.
namespace {
struct Arena {
    struct Chunk { char *base; size_t size, used; };
    std::mutex mu;
    std::vector<Chunk> chunks[64];
    std::map<size_t, std::vector<void *>> bins[64];
    std::map<void *, std::pair<int, size_t>> live;          // block -> (device, rounded size)
    static size_t round_up(size_t b) { return (std::max<size_t>(b, 1) + 511) & ~(size_t)511; }
    cudaError_t alloc(void **p, size_t bytes) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        dev &= 63;
        const size_t n = round_up(bytes);
        std::lock_guard<std::mutex> lock(mu);
        auto it = bins[dev].find(n);
        if (it != bins[dev].end() && !it->second.empty()) { *p = it->second.back(); it->second.pop_back(); live[*p] = {dev, n}; return cudaSuccess; }
        for (Chunk &c : chunks[dev])
            if (c.size - c.used >= n) { *p = c.base + c.used; c.used += n; live[*p] = {dev, n}; return cudaSuccess; }
        const size_t csize = std::max<size_t>(n, (size_t)512 << 20);
        void *base = nullptr;
        e = (cudaMalloc)(&base, csize);
        if (e != cudaSuccess) { cudaGetLastError(); if (csize > n) e = (cudaMalloc)(&base, n); if (e != cudaSuccess) return e; chunks[dev].push_back({(char *)base, n, n}); }
        else chunks[dev].push_back({(char *)base, csize, n});
        *p = base; live[*p] = {dev, n};
        return cudaSuccess;
    }
    cudaError_t release(void *p) {
        if (!p) return cudaSuccess;
        std::lock_guard<std::mutex> lock(mu);
        auto it = live.find(p);
        if (it == live.end()) return (cudaFree)(p);         // not ours (IPC buffers)
        bins[it->second.first][it->second.second].push_back(p);
        live.erase(it);
        return cudaSuccess;
    }
};
Arena g_arena;
inline cudaError_t raw_cuda_malloc(void **p, size_t bytes) { return cudaMalloc(p, bytes); }      // a real allocation (CUDA IPC)
template <class T> cudaError_t arena_malloc(T **p, size_t bytes) { return g_arena.alloc((void **)p, bytes); }
inline cudaError_t arena_free(void *p) { return g_arena.release(p); }
}  // namespace
#define cudaMalloc arena_malloc
#define cudaFree arena_free

}  // namespace mct

using namespace mct;

struct mct_kernel_s {
    int type = 0, device = 0, Nk = 0, Ns = 1, sym = 0;
    cudaStream_t stream = nullptr;
    double *d_k = nullptr, *d_V1 = nullptr, *d_V2 = nullptr, *d_V3 = nullptr;  // raw column-major tables (may be absent)
    // S(k)-defined kernels keep what is needed to rebuild tables in any layout
    bool from_sk = false;
    double rho = 0, kBT = 0, m = 0;
    double *d_Sk = nullptr, *d_Ck = nullptr;
    // multi-component: direct correlation blocks + prefactor; dDim: fused table
    double *d_C = nullptr, *d_pref = nullptr, *d_W = nullptr;
    // tagged
    long nt = 0;
    double *d_Fc = nullptr;
    bool owns_Fc = false;
    std::vector<double> t_c;               // output times of the collective solution (the keys of the reference's tDict)
    mct_equation_t borrowed_from = nullptr; // equation whose trajectory d_Fc points into (owns_Fc == false)
    int live_equations = 0;                // equations created on this handle and not destroyed yet
    std::map<TeamLayout *, double *> slabs;  // team-layout tables per geometry
    double *d_scratch = nullptr; size_t scratch_doubles = 0;
    double *d_io = nullptr; size_t io_doubles = 0;
    int n_sm = 0; size_t smem_optin = 0;
    // retired device buffers kept for the next request of exactly the same size (mct_td_solve in a loop: cudaMalloc and
    // above all cudaFree cost 5 ms to several 100 ms per solve otherwise -- both synchronise with the driver)
    std::vector<std::pair<void *, size_t>> pool;
};

struct mct_equation_s {
    mct_kernel_t kernel = nullptr;
    int Nk = 0, nblocks = 0, second_order = 0, mode = 0;
    long nt = 0;
    mct_td_options opts{};
    double *d_coef = nullptr;   // alpha, beta, gamma, delta, F0, dF0  [6][Nk]
    double *d_ring = nullptr, *d_F = nullptr, *d_K = nullptr;
    size_t coef_bytes = 0, ring_bytes = 0, traj_bytes = 0;   // set when the buffers come from the kernel's pool
    int *d_status = nullptr;
    long long *d_evals = nullptr;
    bool solved = false;
    // block-valued / dense kernels (td_grid.cu): one allocation holding coefficients and workspace
    int ns = 1;
    double *d_gws = nullptr;
    int *d_deal = nullptr; size_t deal_cap = 0; int deal_nsm = 0;      // inside d_gws: the host's deal of the line-sum tasks (grid_upload_deal)
    GridParams gp{};
    // k-sharding over GPUs: this rank's exchange buffer (IPC-exported) and the peers' mappings
    double *d_xshard = nullptr;
    void *peer_maps[kMaxShardRanks] = {};
    int borrowers = 0;          // tagged / active kernels chained on this equation's device-resident trajectory
    bool counted = false;       // included in kernel->live_equations
};

namespace mct {

static constexpr size_t kPoolEntries = 8;
static cudaError_t pool_alloc(mct_kernel_t h, void **p, size_t bytes) {
    for (size_t i = 0; i < h->pool.size(); i++)
        if (h->pool[i].second == bytes) { *p = h->pool[i].first; h->pool.erase(h->pool.begin() + i); return cudaSuccess; }
    return cudaMalloc(p, bytes);
}
static void pool_free(mct_kernel_t h, void *p, size_t bytes) {
    if (!p) return;
    if (h && h->pool.size() < kPoolEntries) h->pool.emplace_back(p, bytes);
    else cudaFree(p);
}

static int ensure_scratch(mct_kernel_t h, size_t doubles) {
    if (h->scratch_doubles >= doubles) return MCT_OK;
    cudaFree(h->d_scratch);
    h->d_scratch = nullptr; h->scratch_doubles = 0;
    MCT_CUDA(cudaMalloc(&h->d_scratch, doubles * sizeof(double)));
    h->scratch_doubles = doubles;
    return MCT_OK;
}
static int ensure_io(mct_kernel_t h, size_t doubles) {
    if (h->io_doubles >= doubles) return MCT_OK;
    cudaFree(h->d_io);
    h->d_io = nullptr; h->io_doubles = 0;
    MCT_CUDA(cudaMalloc(&h->d_io, doubles * sizeof(double)));
    h->io_doubles = doubles;
    return MCT_OK;
}

static bool approx(double a, double b) {  // Julia isapprox default: rtol = sqrt(eps)
    return std::fabs(a - b) <= 1.4901161193847656e-08 * std::max(std::fabs(a), std::fabs(b));
}

// @assert k_array[1] ≈ Δk/2 ; @assert all(diff(k_array) .≈ Δk)   src/Kernels/ModeCouplingKernels.jl:103-104
static int check_grid(int Nk, const double *k) {
    MCT_REQUIRE(Nk >= 2, MCT_BAD_ARGUMENT, "k_array needs at least two points");
    MCT_REQUIRE(k != nullptr, MCT_BAD_ARGUMENT, "k_array is NULL");
    const double dk = k[1] - k[0];
    MCT_REQUIRE(approx(k[0], dk / 2), MCT_BAD_ARGUMENT, "grid assertion failed: k_array[1] must equal dk/2");
    for (int i = 1; i < Nk; i++)
        MCT_REQUIRE(approx(k[i] - k[i - 1], dk), MCT_BAD_ARGUMENT, "grid assertion failed: k_array must be equidistant");
    return MCT_OK;
}

static int new_handle(mct_kernel_t *out, int type, int device, int Nk, const double *k) {
    MCT_REQUIRE(out != nullptr, MCT_BAD_ARGUMENT, "output handle pointer is NULL");
    *out = nullptr;
    int rc = check_grid(Nk, k);
    if (rc) return rc;
    int ndev = 0;
    MCT_CUDA(cudaGetDeviceCount(&ndev));
    MCT_REQUIRE(device >= 0 && device < ndev, MCT_BAD_ARGUMENT, "no such CUDA device");
    MCT_CUDA(cudaSetDevice(device));
    // device attributes, queried once per device (cudaGetDeviceProperties costs ~20 ms per call: 10 s for a 512-point sweep)
    struct DevInfo { int major = -1, n_sm = 0, smem_optin = 0; };
    static DevInfo g_dev[64];
    DevInfo &di = g_dev[device & 63];
    if (di.major < 0) {
        int major = 0;
        MCT_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
        MCT_CUDA(cudaDeviceGetAttribute(&di.n_sm, cudaDevAttrMultiProcessorCount, device));
        MCT_CUDA(cudaDeviceGetAttribute(&di.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        di.major = major;
    }
    MCT_REQUIRE(di.major == 10, MCT_CUDA_ERROR, "libmct_sm100a needs an sm_100 (B200) device; there is no CPU or other-GPU fallback");
    mct_kernel_t h = new mct_kernel_s();
    h->type = type; h->device = device; h->Nk = Nk;
    h->n_sm = di.n_sm; h->smem_optin = (size_t)di.smem_optin;
    // Kernel handles share a small set of streams per device (round robin): creating a stream costs ~9 ms on these boxes
    // (4.6 s for the 512 handles of a sweep).  Work of handles that share a stream is merely serialised.
    cudaError_t e = cudaSuccess;
    {
        std::lock_guard<std::mutex> lock(g_stream_mutex);
        auto &pool = g_stream_pool[device & 63];
        if ((int)pool.size() < kStreamsPerDevice) {
            cudaStream_t st = nullptr;
            e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
            if (e == cudaSuccess) pool.push_back(st);
        }
        if (e == cudaSuccess) h->stream = pool[g_stream_next[device & 63]++ % pool.size()];
    }
    if (e == cudaSuccess) e = cudaMalloc(&h->d_k, Nk * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->d_k, k, Nk * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "new_handle", __FILE__, __LINE__); }
    *out = h;
    return MCT_OK;
}

static int upload_tables(mct_kernel_t h, const double *V1, const double *V2, const double *V3) {
    const size_t n2 = (size_t)h->Nk * h->Nk;
    MCT_REQUIRE(V1 && V2 && V3, MCT_BAD_ARGUMENT, "vertex table pointer is NULL");
    MCT_CUDA(cudaMalloc(&h->d_V1, 3 * n2 * sizeof(double)));
    h->d_V2 = h->d_V1 + n2; h->d_V3 = h->d_V2 + n2;
    MCT_CUDA(cudaMemcpyAsync(h->d_V1, V1, n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    MCT_CUDA(cudaMemcpyAsync(h->d_V2, V2, n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    MCT_CUDA(cudaMemcpyAsync(h->d_V3, V3, n2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return MCT_OK;
}

static bool tables_symmetric(int Nk, const double *V1, const double *V2, const double *V3) {
    for (int a = 0; a < Nk; a++)
        for (int b = a + 1; b < Nk; b++) {
            size_t x = (size_t)a + (size_t)b * Nk, y = (size_t)b + (size_t)a * Nk;
            if (V1[x] != V1[y] || V2[x] != V2[y] || V3[x] != V3[y]) return false;
        }
    return true;
}

static int ensure_raw_tables(mct_kernel_t h) {
    if (h->d_V1) return MCT_OK;
    MCT_REQUIRE(h->from_sk, MCT_BAD_ARGUMENT, "kernel has no tables");
    const size_t n2 = (size_t)h->Nk * h->Nk;
    MCT_CUDA(cudaMalloc(&h->d_V1, 3 * n2 * sizeof(double)));
    h->d_V2 = h->d_V1 + n2; h->d_V3 = h->d_V2 + n2;
    return tables_from_sk(h->Nk, h->d_k, h->d_Sk, h->d_Ck, h->rho, h->kBT, h->m, h->type == K_TAGGED3D, h->d_V1, h->d_V2, h->d_V3,
                          h->stream);
}

static int setup_from_sk(mct_kernel_t h, double rho, double kBT, double m, const double *Sk) {
    MCT_REQUIRE(Sk != nullptr, MCT_BAD_ARGUMENT, "Sk is NULL");
    h->from_sk = true; h->rho = rho; h->kBT = kBT; h->m = m;
    MCT_CUDA(cudaMalloc(&h->d_Sk, 2 * (size_t)h->Nk * sizeof(double)));
    h->d_Ck = h->d_Sk + h->Nk;
    MCT_CUDA(cudaMemcpyAsync(h->d_Sk, Sk, h->Nk * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return tables_from_sk(h->Nk, h->d_k, h->d_Sk, h->d_Ck, rho, kBT, m, h->type == K_TAGGED3D, nullptr, nullptr, nullptr, h->stream);
}

static int upload_traj(mct_kernel_t h, long nt, const double *Fc) {
    MCT_REQUIRE(nt >= 1 && Fc != nullptr, MCT_BAD_ARGUMENT, "tagged kernel needs the collective trajectory");
    h->nt = nt; h->owns_Fc = true;
    MCT_CUDA(cudaMalloc(&h->d_Fc, (size_t)nt * h->Nk * sizeof(double)));
    MCT_CUDA(cudaMemcpyAsync(h->d_Fc, Fc, (size_t)nt * h->Nk * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return MCT_OK;
}

// ---- team geometry selection ------------------------------------------------------------------------------
static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

struct Residency { int UR, UT, US, lane_cap; };

// How the U units of every warp are kept: UR in registers (a kernel instantiation), US in shared memory, rest in L2.
static Residency choose_residency(const TeamLayout &L, size_t smem_optin, int fabric) {
    Residency r;
    // 24 registers per register-resident unit; the instantiation that finishes 4 rows per thread (Nk > 512) has 28
    // fewer registers to spare (measured: 2 units there, 4 otherwise, keeps spills out of the kernel)
    // Tensor memory first: a warp owns 256 of the SM's 512 columns in its lane quarter = 10 units of 24 columns, read at
    // ~2.7x the shared-memory rate and without taking registers (scripts/tmem_bench.cu).  Then registers (24 per unit; the
    // instantiations that finish 4 or 8 rows per thread have fewer to spare), then shared memory, the rest streams from L2.
    const int utmax = env_int("MCT_TMEM", 1) ? std::min(kTmemUnitsPerWarp, env_int("MCT_MAX_TMEM_UNITS", kTmemUnitsPerWarp)) : 0;
    const int urmax = std::min(td_sc_max_reg_units(L.Nk), env_int("MCT_MAX_REG_UNITS", td_sc_max_reg_units(L.Nk)));
    r.UR = 0;
    if (L.U > utmax) for (int ur : {2, 4}) if (ur <= urmax && ur <= L.U) { r.UR = ur; if (ur >= L.U - utmax) break; }
    r.UT = std::min(L.U - r.UR, utmax);
    r.lane_cap = std::max(1, L.went_max);               // one scratch column set per (warp, row block) entry
    SolveParams P{};
    P.Nk = L.Nk; P.G = L.G; P.rows_max = L.rows_max; P.ent_max = L.ent_max; P.lane_cap = r.lane_cap; P.U = L.U; P.UT = r.UT; P.US = 0;
    P.fabric = fabric; P.xbuf_words = L.xbuf_words;
    const int want = std::max(0, L.U - r.UR - r.UT);
    const size_t fixed = td_sc_smem_bytes(P);
    const size_t per_unit = (size_t)kWarps * (kUV * 32 * 8 + 4);
    const int cap = fixed + 64 < smem_optin ? (int)((smem_optin - fixed - 64) / per_unit) : 0;
    r.US = std::min(want, cap);
    return r;
}

// Returns the (cached) geometry for n_eq concurrent equations on this kernel's device.
static int choose_geometry(mct_kernel_t h, int n_eq, TeamLayout **out, int *nteams_out) {
    const int Nk = h->Nk, sym = h->sym;
    const int n_sm = h->n_sm;
    const long long units = layout_total_units(Nk, sym);
    const int g_cap = std::max(1, std::min(n_sm, Nk / 2));               // at least two rows per CTA
    const int g_min = (Nk + kThreads - 1) / kThreads;                    // one owner thread per row
    // on-chip capacity of one CTA in units per warp: tensor memory (10) + registers + what shared memory can hold
    const int u_tmem = env_int("MCT_TMEM", 1) ? std::min(kTmemUnitsPerWarp, env_int("MCT_MAX_TMEM_UNITS", kTmemUnitsPerWarp)) : 0;
    const int u_chip = u_tmem + td_sc_max_reg_units(Nk) + (int)((h->smem_optin - 72 * 1024 - (size_t)48 * Nk) / ((size_t)kWarps * kUV * 32 * 8));
    int G_fit = (int)std::min<long long>(g_cap, std::max<long long>(g_min, (units + (long long)kWarps * u_chip - 1) / ((long long)kWarps * u_chip)));
    int G;
    if (n_eq == 1) {
        // latency: as many CTAs as keep every warp busy with a couple of units per iteration
        const int target = env_int("MCT_TARGET_UNITS", 2);
        G = (int)std::min<long long>(g_cap, std::max<long long>(G_fit, (units + (long long)kWarps * target - 1) / ((long long)kWarps * target)));
        // small problems: spread to at least 32 CTAs (one unit or less per warp) on the L2 fabric.  Round 1 used a thread-block
        // cluster of <= 16 CTAs with a DSMEM exchange here; since the round-2 rebuild of the L2 exchange the larger L2 team is
        // faster (Nk=100: 3.4 us per evaluation at G=37, 3.6 at a cluster of 16, 4.0 at a cluster of 11) -- the cluster fabric
        // remains for explicitly requested team sizes (MCT_TEAM_SIZE <= 16).
        G = (int)std::min<long long>(g_cap, std::max<long long>(G, env_int("MCT_MIN_LATENCY_TEAM", 32)));
    } else {
        // throughput: the smallest team whose tables stay on chip (the exchange is a per-CTA cost)
        G = G_fit;
    }
    const int forced = env_int("MCT_TEAM_SIZE", 0);
    if (forced > 0) G = std::max(g_min, std::min(forced, g_cap));

    std::lock_guard<std::mutex> lock(g_geom_mutex);
    std::map<GeomKey, TeamLayout *>::iterator it;
    while (true) {
        GeomKey key(h->device, Nk, G, sym);
        it = g_geoms.find(key);
        if (it == g_geoms.end()) {
            TeamLayout *L = new TeamLayout();
            int rc = layout_build_geometry(*L, Nk, G, sym);
            if (rc) { delete L; return rc; }
            it = g_geoms.emplace(key, L).first;
        }
        // the warp scratch holds 12 columns per (warp, row block) entry: short rows in a small team need a larger team
        if (it->second->went_max <= 6 || G >= g_cap) break;
        G++;
    }
    MCT_REQUIRE(it->second->went_max <= 8, MCT_UNSUPPORTED, "internal: no team geometry with <= 8 row blocks per warp");
    const int nteams = std::max(1, std::min(n_eq, n_sm / G));
    *out = it->second;
    *nteams_out = nteams;
    return MCT_OK;
}

static int ensure_slab(mct_kernel_t h, TeamLayout *L, const double **tab) {
    auto it = h->slabs.find(L);
    if (it != h->slabs.end()) { *tab = it->second; return MCT_OK; }
    double *d = nullptr;
    MCT_CUDA(cudaMalloc(&d, (size_t)L->slab_doubles() * sizeof(double)));
    int rc;
    if (h->d_V1) rc = layout_fill_from_tables(*L, d, h->d_V1, h->d_V2, h->d_V3, h->stream);
    else rc = layout_fill_from_sk(*L, d, h->d_k, h->d_Ck, h->rho, h->kBT, h->m, h->type == K_TAGGED3D, h->stream);
    if (rc) { cudaFree(d); return rc; }
    h->slabs[L] = d;
    *tab = d;
    return MCT_OK;
}

static int count_blocks(const mct_td_options *o) {
    int n = 0;
    double dt = o->dt;
    while (dt < o->t_max * 2.0) { n++; dt *= 2.0; }   // src/TimeDoublingSolver.jl:523
    return n;
}

// t = dt/(4N) * it, exactly as allocate_results! forms it (src/TimeDoublingSolver.jl:431-434)
static void output_times(const mct_td_options &o, int nblocks, std::vector<double> &t) {
    const int N = o.N;
    t.clear();
    t.push_back(0.0);
    double Dt = o.dt;
    const double d0 = Dt / (4.0 * N);
    for (int it = 1; it <= 2 * N; it++) t.push_back(d0 * it);
    for (int b = 0; b < nblocks; b++, Dt *= 2.0) {
        const double dl = Dt / (4.0 * N);
        for (int it = 2 * N + 1; it <= 4 * N; it++) t.push_back(dl * it);
    }
}

// RAII pair of CUDA events (no leak on the error paths of the timed launches)
struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaError_t create() { cudaError_t e = cudaEventCreate(&a); return e == cudaSuccess ? cudaEventCreate(&b) : e; }
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
};

static int expand_coef(int kind, double lam, const double *arr, int Nk, double *dst, const char *name) {
    if (kind == 0) { for (int i = 0; i < Nk; i++) dst[i] = lam; return MCT_OK; }
    MCT_REQUIRE(kind == 1 && arr != nullptr, MCT_BAD_ARGUMENT, std::string("bad coefficient ") + name);
    memcpy(dst, arr, Nk * sizeof(double));
    return MCT_OK;
}

// One persistent launch for a batch of equations that share kernel geometry and solver options.
// Set by mct_td_solve around the solve: called once with the event that marks the end of the launch, while the kernel runs.
static thread_local std::function<void(cudaEvent_t)> *g_overlap = nullptr;

static int run_batch(int n, mct_equation_t *eqs, int mode, double tol, long long max_iter, int *status, long long *evals,
                     float *device_ms) {
    MCT_REQUIRE(n >= 1 && eqs != nullptr, MCT_BAD_ARGUMENT, "empty batch");
    mct_kernel_t h0 = eqs[0]->kernel;
    MCT_CUDA(cudaSetDevice(h0->device));
    for (int i = 0; i < n; i++) {
        mct_kernel_t h = eqs[i]->kernel;
        MCT_REQUIRE(h->device == h0->device && h->Nk == h0->Nk && h->sym == h0->sym, MCT_BAD_ARGUMENT,
                    "equations of a batch must share device, Nk and table symmetry");
        MCT_REQUIRE(h->type == K_SC3D || h->type == K_TAGGED3D, MCT_UNSUPPORTED, "kernel type not supported by this solver");
        if (mode == 0) {
            MCT_REQUIRE(eqs[i]->opts.N == eqs[0]->opts.N && eqs[i]->opts.dt == eqs[0]->opts.dt && eqs[i]->nblocks == eqs[0]->nblocks &&
                            eqs[i]->second_order == eqs[0]->second_order,
                        MCT_BAD_ARGUMENT, "equations of a batch must share the solver options");
        }
    }
    TeamLayout *L = nullptr;
    int nteams = 1;
    int rc = choose_geometry(h0, n, &L, &nteams);
    if (rc) return rc;
    cudaStream_t s = h0->stream;
    std::vector<EqDev> eh(n);
    for (int i = 0; i < n; i++) {
        mct_equation_t q = eqs[i];
        mct_kernel_t h = q->kernel;
        MCT_CUDA(cudaStreamSynchronize(h->stream));   // uploads of this handle
        const double *tab = nullptr;
        rc = ensure_slab(h, L, &tab);
        if (rc) return rc;
        MCT_CUDA(cudaStreamSynchronize(h->stream));
        const int Nk = q->Nk;
        EqDev &e = eh[i];
        e.tab = tab; e.kvec = h->d_k;
        e.alpha = q->d_coef; e.beta = q->d_coef + Nk; e.gamma = q->d_coef + 2 * Nk; e.delta = q->d_coef + 3 * Nk;
        e.F0 = q->d_coef + 4 * Nk; e.dF0 = q->d_coef + 5 * Nk;
        e.ring = q->d_ring; e.F_out = q->d_F; e.K_out = q->d_K;
        e.Fc_traj = (h->type == K_TAGGED3D) ? h->d_Fc : nullptr;
        e.mode = mode;
        e.status = q->d_status; e.kernel_evals = q->d_evals;
    }
    SolveParams P{};
    int fabric = (L->G > 1 && L->G <= 16 && env_int("MCT_CLUSTER", 1)) ? 1 : 0;
    const int nteams_wanted = nteams;
    for (;;) {
        const Residency res = choose_residency(*L, h0->smem_optin, fabric);
        P = SolveParams{};
        P.Nk = h0->Nk; P.G = L->G; P.N = eqs[0]->opts.N; P.nblocks = eqs[0]->nblocks; P.n_eq = n; P.nteams = nteams;
        P.second_order = eqs[0]->second_order;
        P.dt0 = eqs[0]->opts.dt; P.tol = tol; P.max_iter = max_iter;
        P.U = L->U; P.UR = res.UR; P.UT = res.UT; P.US = res.US; P.lane_cap = res.lane_cap;
        P.rows_max = L->rows_max; P.rb_max = L->rb_max; P.ent_max = L->ent_max;
        P.udesc = L->d_udesc; P.cta_meta = L->d_cta_meta; P.ent_begin = L->d_ent_begin; P.warp_ent = L->d_warp_ent; P.row_owner = L->d_row_owner;
        P.xbuf_words = L->xbuf_words; P.xrows = L->xrows; P.fabric = fabric;
        // history rings in shared memory when they fit beside everything else (time-doubling solves only)
        P.ring_smem = 0;
        if (mode == 0 && env_int("MCT_RING_SMEM", 1)) {
            P.ring_smem = 1;
            if (td_sc_smem_bytes(P) + 1024 > h0->smem_optin) P.ring_smem = 0;
        }
        P.xstride = 4LL * L->xbuf_words + 16;      // four rotating buffers + the arrival counter (own line)
        P.dbg = env_int("MCT_DBG", 0);
        P.snap_epoch = env_int("MCT_SNAP_EPOCH", 1000);
        P.spin_limit_cycles = (long long)env_int("MCT_SPIN_LIMIT_MS", 2000) * 2000000LL;
        if (fabric == 0) break;
        // Cluster teams: the hardware places a cluster inside one GPC, so only a few clusters of 9..16 CTAs are resident at once
        // (7 of 16 on a B200: 112 of 148 SMs).  A batch that has more equations than that runs more teams through L2 instead.
        const int maxc = td_sc_max_clusters(P);
        if (maxc >= 1 && (maxc >= nteams_wanted || maxc * L->G * 5 >= h0->n_sm * 4)) { nteams = std::max(1, std::min(nteams, maxc)); P.nteams = nteams; break; }
        MCT_REQUIRE(maxc >= 1 || L->G > 1, MCT_CUDA_ERROR, "thread-block cluster launch not possible for this team size");
        fabric = 0;
    }
    if (env_int("MCT_VERBOSE", 0))
        fprintf(stderr, "[mct] Nk=%d sym=%d n_eq=%d: teams=%d x G=%d, U=%d units/warp (%d regs, %d tmem, %d smem, %d L2), rows/CTA<=%d, entries/warp<=%d, smem=%zu B, fabric=%s\n",
                P.Nk, h0->sym, n, nteams, P.G, P.U, std::min(P.U, P.UR), P.UT, P.US, std::max(0, P.U - P.UR - P.UT - P.US), P.rows_max, L->went_max, td_sc_smem_bytes(P), P.fabric ? "cluster/DSMEM" : (P.G > 1 ? "L2" : "none"));

    // communication scratch: exchange buffers and queue mailboxes (all slots armed with the sentinel = 0xFF bytes),
    // queue, abort flag, equation table
    const size_t xbytes = (size_t)nteams * P.xstride * sizeof(double);
    const size_t qbytes = (size_t)nteams * 64 * sizeof(double);
    const size_t ebytes = (size_t)n * sizeof(EqDev);
    unsigned char *comm = nullptr;
    const size_t pbytes = (size_t)P.G * kProfSlots * sizeof(long long);
    const size_t total = xbytes + qbytes + 256 + ebytes + pbytes;
    MCT_CUDA(pool_alloc(h0, (void **)&comm, total));
    cudaError_t ce = cudaMemsetAsync(comm, 0, total, s);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(comm, 0xFF, xbytes + qbytes, s);
    for (int t = 0; t < nteams && ce == cudaSuccess; t++)      // arrival counters start at zero
        ce = cudaMemsetAsync(comm + ((size_t)t * P.xstride + 4 * (size_t)P.xbuf_words) * sizeof(double), 0, 16 * sizeof(double), s);
    P.xbuf = (double *)comm;
    P.qbox = (double *)(comm + xbytes);
    P.queue = (int *)(comm + xbytes + qbytes);
    P.abort_flag = P.queue + 32;
    EqDev *d_eqs = (EqDev *)(comm + xbytes + qbytes + 256);
    P.eqs = d_eqs;
    P.prof = (long long *)(comm + xbytes + qbytes + 256 + ebytes);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_eqs, eh.data(), ebytes, cudaMemcpyHostToDevice, s);
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ce == cudaSuccess) ce = cudaEventCreate(&ev0);
    if (ce == cudaSuccess) ce = cudaEventCreate(&ev1);
    if (ce == cudaSuccess) ce = cudaEventRecord(ev0, s);
    if (ce == cudaSuccess) {
        rc = td_sc_launch(P, s);
        if (rc == MCT_OK) {
            ce = cudaEventRecord(ev1, s);
            if (ce == cudaSuccess && g_overlap) (*g_overlap)(ev1);          // host work while the persistent kernel runs
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
            float ms = 0.f;
            if (ce == cudaSuccess) ce = cudaEventElapsedTime(&ms, ev0, ev1);
            if (device_ms) *device_ms = ms;
        }
    }
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (ce != cudaSuccess) { cudaFree(comm); return cuda_fail(ce, "run_batch", __FILE__, __LINE__); }
    if (rc != MCT_OK) { pool_free(h0, comm, total); return rc; }
#ifdef MCT_PROFILE
    {
        std::vector<long long> pr((size_t)P.G * kProfSlots);
        cudaMemcpy(pr.data(), P.prof, pbytes, cudaMemcpyDeviceToHost);
        if (getenv("MCT_DUMP_CTAS")) {
            fprintf(stderr, "per CTA (thread 0), kilo-cycles: cta rows | units+reduce  row-sums  publish  collect(wait)  loads+scan  apply  step-local\n");
            for (int c = 0; c < P.G; c++) {
                const long long *q = &pr[(size_t)c * kProfSlots];
                fprintf(stderr, "  %3d | %8lld %8lld %8lld %8lld %8lld %8lld %8lld\n", c, q[0] / 1000, q[2] / 1000, q[4] / 1000, q[5] / 1000, q[6] / 1000, q[7] / 1000, q[9] / 1000);
            }
        }
        if (getenv("MCT_DUMP_BEACON")) {
            fprintf(stderr, "beacons (cta:epoch.phase):");
            for (int c = 0; c < P.G; c++) fprintf(stderr, " %d:%lld.%lld", c, pr[(size_t)c * kProfSlots + 28] / 16, pr[(size_t)c * kProfSlots + 28] % 16);
            fprintf(stderr, "\n");
        }
        static const char *nm[kProfSlots] = {"stream+warp reduce", "barrier", "row sums+prefix", "F_loc", "publish+delay", "collect", "offset scan+barrier",
                                            "apply+or-barrier", "between", "step-local", "=exchange alone", "=eval alone", "=eval+exchange", "-", "retries: rows", "-", "-", "x:offsets", "x:or-barrier", "-", "-", "-", "-", "-", "-", "-", "-", "-", "-", "-", "TOTAL cycles", "TOTAL ns"};
        const int ctas[3] = {0, P.G / 2, P.G - 1};
        fprintf(stderr, "[mct profile] Nk=%d G=%d nteams=%d U=%d UR=%d UT=%d US=%d; thread 0 of CTA %d / %d / %d (team 0)\n", P.Nk, P.G, nteams,
                P.U, P.UR, P.UT, P.US, ctas[0], ctas[1], ctas[2]);
        for (int i = 0; i < kProfSlots; i++) {
            if (nm[i][0] == '-') continue;
            fprintf(stderr, "   %-20s", nm[i]);
            for (int c : ctas) fprintf(stderr, " %13lld (%4.1f%%)", pr[c * kProfSlots + i], 100.0 * pr[c * kProfSlots + i] / (double)pr[c * kProfSlots + kProfSlots - 2]);
            fprintf(stderr, "\n");
        }
        if (getenv("MCT_DUMP_SNAP")) {
            long long t0 = pr[19];
            for (int c = 0; c < P.G; c++) t0 = std::min(t0, pr[(size_t)c * kProfSlots + 19]);
            fprintf(stderr, "snapshot of exchange 1000 (ns after the first publish): cta publish totals_seen barrier1 rows_loaded end\n");
            for (int c = 0; c < P.G; c++) {
                fprintf(stderr, "  %3d", c);
                for (int i = 19; i <= 23; i++) fprintf(stderr, " %6lld", pr[(size_t)c * kProfSlots + i] - t0);
                fprintf(stderr, "   |");
                for (int i = 24; i < 30; i++) fprintf(stderr, " %6lld", pr[(size_t)c * kProfSlots + i] - t0);
                fprintf(stderr, "\n");
            }
        }
    }
#endif
    int worst = MCT_OK;
    for (int i = 0; i < n; i++) {
        int st = 0; long long ev = 0;
        MCT_CUDA(cudaMemcpy(&st, eqs[i]->d_status, sizeof(int), cudaMemcpyDeviceToHost));
        MCT_CUDA(cudaMemcpy(&ev, eqs[i]->d_evals, sizeof(long long), cudaMemcpyDeviceToHost));
        if (status) status[i] = st;
        if (evals) evals[i] = ev;
        eqs[i]->solved = (st == MCT_OK);
        if (st != MCT_OK && worst == MCT_OK) worst = st;
    }
    pool_free(h0, comm, total);
    if (worst == MCT_NOT_CONVERGED)
        set_error("Iteration did not converge. Either increase the number of time steps before a time doubling, or choose a "
                  "different memory kernel.");
    else if (worst == MCT_CUDA_ERROR) set_error("persistent solver aborted (team barrier timeout)");
    return worst;
}


// ---- block-valued / dense kernels: everything goes through td_grid.cu ------------------------------------------
static bool is_grid_kernel(const mct_kernel_s *h) { return h->type == K_MC3D || h->type == K_DDIM; }

// per-k blocks [Nk][nb] of a multiplicative coefficient: UniformScaling -> lambda*I, Diagonal -> the blocks given
static int expand_coef_blocks(int kind, double lam, const double *arr, int Nk, int ns, double *dst, const char *name) {
    const int nb = ns * ns;
    if (kind == 0) {
        std::fill(dst, dst + (size_t)Nk * nb, 0.0);
        for (int k = 0; k < Nk; k++) for (int a = 0; a < ns; a++) dst[(size_t)k * nb + a + a * ns] = lam;
        return MCT_OK;
    }
    MCT_REQUIRE(kind == 1 && arr != nullptr, MCT_BAD_ARGUMENT, std::string("bad coefficient ") + name);
    memcpy(dst, arr, (size_t)Nk * nb * sizeof(double));
    return MCT_OK;
}

// The host's deal of the line-sum tasks for this equation's (rank, world) on a grid of one block per SM (td_grid.cu,
// grid_build_deal); a table that does not fit the reserved space leaves the kernel on its closed-form deal.
static int grid_upload_deal(mct_equation_t q) {
    GridParams &g = q->gp;
    g.deal = nullptr; g.deal_grid = 0;
    if (g.type != GK_MC3D || !q->d_deal || env_int("MCT_GRID_DEAL", 1) == 0) return MCT_OK;
    std::vector<int> table; int wsmax = 0, stride = 0;
    grid_build_deal(g.Nk, g.ns, g.rank, g.world, q->deal_nsm, table, wsmax, stride);
    if (table.size() > q->deal_cap) return MCT_OK;
    MCT_CUDA(cudaMemcpy(q->d_deal, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice));
    g.deal = q->d_deal; g.deal_grid = q->deal_nsm; g.deal_wsmax = wsmax; g.deal_stride = stride;
    return MCT_OK;
}

// Fills q->gp and allocates its workspace.  mode 0: time doubling (needs opts), 1: steady state, 2: one evaluation.
static int grid_setup(mct_equation_t q, mct_kernel_t h, int mode, const std::vector<double> &coef_host /* [6][dof] */) {
    const int Nk = h->Nk, ns = h->Ns, nb = ns * ns;
    const size_t dof = (size_t)Nk * nb;
    const int N = (mode == 0) ? q->opts.N : 1;
    const size_t ring = (mode == 0) ? (size_t)4 * N * dof : 0;
    const size_t traj = (mode == 0) ? (size_t)q->nt * dof : dof;
    const size_t lsum = grid_lsum_words(Nk, ns);
    const size_t dfh = (mode == 0) ? (size_t)(2 * N + 1) * dof : 0;
    const size_t deal_cap = (h->type == K_MC3D) ? grid_deal_capacity(Nk, ns, h->n_sm) : 0;      // ints, behind everything else
    const size_t total = 6 * dof + 4 * ring + 7 * dof + 4 * dof + 6 * dof + lsum + dfh + 2 * traj + 40 + (deal_cap + 1) / 2;
    MCT_CUDA(cudaMalloc(&q->d_gws, total * sizeof(double)));
    cudaError_t e_init = cudaMemset(q->d_gws, 0, total * sizeof(double));
    if (e_init == cudaSuccess) e_init = cudaMemcpy(q->d_gws, coef_host.data(), 6 * dof * sizeof(double), cudaMemcpyHostToDevice);
    if (e_init != cudaSuccess) { cudaFree(q->d_gws); q->d_gws = nullptr; return cuda_fail(e_init, "grid_setup", __FILE__, __LINE__); }
    GridParams &g = q->gp;
    g = GridParams{};
    g.type = (h->type == K_MC3D) ? GK_MC3D : GK_DDIM; g.Nk = Nk; g.ns = ns; g.mode = mode;
    g.rank = 0; g.world = 1; g.spin_limit_cycles = (long long)env_int("MCT_SHARD_SPIN_LIMIT_MS", 20000) * 2000000LL;
    g.kvec = h->d_k; g.C = h->d_C; g.pref = h->d_pref; g.W = h->d_W;
    g.Fc = (h->type == K_DDIM) ? h->d_Fc : nullptr; g.nt_c = h->nt; g.tix0 = 0;
    double *p = q->d_gws;
    g.alpha = p; g.beta = p + dof; g.gamma = p + 2 * dof; g.delta = p + 3 * dof; g.F0 = p + 4 * dof; g.dF0 = p + 5 * dof; p += 6 * dof;
    g.Ft = p; g.Kt = p + ring; g.FI = p + 2 * ring; g.KI = p + 3 * ring; p += 4 * ring;
    g.C1 = p; g.C2 = p + dof; g.C3 = p + 2 * dof; g.Fold = p + 3 * dof; g.K0 = p + 4 * dof; g.Fcur = p + 5 * dof; g.Kcur = p + 6 * dof; p += 7 * dof;
    g.G = p; p += 4 * dof;
    g.Inc = p; p += 3 * dof; g.Tm = p; p += 3 * dof;
    g.Lsum = p; p += lsum;
    g.dFh = p; p += dfh;
    g.F_out = p; g.K_out = p + traj; p += 2 * traj;
    g.err = (unsigned long long *)p; p += 2;
    g.status = (int *)p; g.kernel_evals = (long long *)(p + 2); g.prof = (unsigned long long *)(p + 4); g.xepoch_state = (unsigned *)(p + 32);
    q->d_F = g.F_out; q->d_K = g.K_out;
    q->d_deal = (int *)(p + 38); q->deal_cap = deal_cap; q->deal_nsm = h->n_sm;      // p + 38 = end of the 40 doubles that start at g.err
    return grid_upload_deal(q);
}

static int grid_run(mct_equation_t q, long long *evals, int *status, float *device_ms) {
    mct_kernel_t h = q->kernel;
    MCT_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    MCT_CUDA(cudaMemsetAsync(q->gp.err, 0, 32 * sizeof(double), s));
    EventPair tm;
    MCT_CUDA(tm.create());
    MCT_CUDA(cudaEventRecord(tm.a, s));
    int rc = td_grid_launch(q->gp, h->n_sm, s);
    if (rc) return rc;
    MCT_CUDA(cudaEventRecord(tm.b, s));
    MCT_CUDA(cudaStreamSynchronize(s));
    float ms = 0.f;
    MCT_CUDA(cudaEventElapsedTime(&ms, tm.a, tm.b));
    if (device_ms) *device_ms = ms;
    int st = 0; long long ev = 0;
    MCT_CUDA(cudaMemcpy(&st, q->gp.status, sizeof(int), cudaMemcpyDeviceToHost));
    MCT_CUDA(cudaMemcpy(&ev, q->gp.kernel_evals, sizeof(long long), cudaMemcpyDeviceToHost));
    if (env_int("MCT_GRID_PROF", 0)) {
        unsigned long long pr[12];
        MCT_CUDA(cudaMemcpy(pr, q->gp.prof, sizeof(pr), cudaMemcpyDeviceToHost));
        char line[512];
        int n = snprintf(line, sizeof line, "[mct] rank %d grid phases (cycles of thread 0, %lld evals, %.3f ms):", q->gp.rank, ev, ms);
        for (int i = 0; i < 12; i++) n += snprintf(line + n, sizeof line - n, " %llu", pr[i]);
        snprintf(line + n, sizeof line - n, "\n");
        fputs(line, stderr);
    }
    if (evals) *evals = ev;
    if (status) *status = st;
    return st;
}

// A tagged / active kernel chained on the device-resident trajectory of a solved equation: it reads q->d_F in place, so the
// equation must outlive it (mct_equation_destroy refuses while borrowers exist) and it inherits the equation's time grid
// (the keys of the reference's tDict, src/Kernels/ModeCouplingKernels.jl:394).
static void borrow_trajectory(mct_kernel_t h, mct_equation_t collective) {
    h->nt = collective->nt; h->d_Fc = collective->d_F; h->owns_Fc = false;
    h->borrowed_from = collective;
    collective->borrowers++;
    output_times(collective->opts, collective->nblocks, h->t_c);
}

}  // namespace mct

// ================================================ extern "C" ================================================
extern "C" {

const char *mct_last_error(void) { return g_last_error.c_str(); }
int mct_version(void) { return 100; }
long mct_launch_count(void) { return g_launches.load(); }

int mct_device_count(int *count) {
    MCT_REQUIRE(count != nullptr, MCT_BAD_ARGUMENT, "count is NULL");
    *count = 0;
    MCT_CUDA(cudaGetDeviceCount(count));
    return MCT_OK;
}

int mct_sc3d_create(mct_kernel_t *out, int device, int Nk, const double *k_array, const double *V1, const double *V2,
                    const double *V3) {
    int rc = new_handle(out, K_SC3D, device, Nk, k_array);
    if (rc) return rc;
    rc = upload_tables(*out, V1, V2, V3);
    if (rc) { mct_kernel_destroy(*out); *out = nullptr; return rc; }
    (*out)->sym = tables_symmetric(Nk, V1, V2, V3) ? 1 : 0;
    return MCT_OK;
}

int mct_sc3d_create_from_sk(mct_kernel_t *out, int device, int Nk, double rho, double kBT, double m, const double *k_array,
                            const double *Sk) {
    int rc = new_handle(out, K_SC3D, device, Nk, k_array);
    if (rc) return rc;
    (*out)->sym = 1;   // the collective vertices are bitwise symmetric under p <-> q (:122-124)
    rc = setup_from_sk(*out, rho, kBT, m, Sk);
    if (rc) { mct_kernel_destroy(*out); *out = nullptr; }
    return rc;
}

int mct_sc3d_tagged_create(mct_kernel_t *out, int device, int Nk, const double *k_array, const double *V1, const double *V2,
                           const double *V3, long nt, const double *Fc_traj) {
    int rc = new_handle(out, K_TAGGED3D, device, Nk, k_array);
    if (rc) return rc;
    rc = upload_tables(*out, V1, V2, V3);
    if (!rc) rc = upload_traj(*out, nt, Fc_traj);
    if (rc) { mct_kernel_destroy(*out); *out = nullptr; }
    return rc;
}

int mct_sc3d_tagged_create_from_sk(mct_kernel_t *out, int device, int Nk, double rho, double kBT, double m,
                                   const double *k_array, const double *Sk, long nt, const double *Fc_traj) {
    int rc = new_handle(out, K_TAGGED3D, device, Nk, k_array);
    if (rc) return rc;
    rc = setup_from_sk(*out, rho, kBT, m, Sk);
    if (!rc) rc = upload_traj(*out, nt, Fc_traj);
    if (rc) { mct_kernel_destroy(*out); *out = nullptr; }
    return rc;
}

int mct_sc3d_tagged_create_from_equation(mct_kernel_t *out, int Nk, double rho, double kBT, double m, const double *k_array,
                                         const double *Sk, mct_equation_t collective) {
    MCT_REQUIRE(collective && collective->solved && collective->mode == 0, MCT_BAD_ARGUMENT,
                "collective equation must be solved before a tagged kernel is chained on it");
    MCT_REQUIRE(collective->Nk == Nk, MCT_BAD_ARGUMENT, "collective solution has a different Nk");
    MCT_REQUIRE(collective->ns == 1 && !collective->d_gws, MCT_BAD_ARGUMENT,
                "the collective solution must be single-component (mixtures: mct_mc3d_tagged_create_from_equation)");
    int rc = new_handle(out, K_TAGGED3D, collective->kernel->device, Nk, k_array);
    if (rc) return rc;
    rc = setup_from_sk(*out, rho, kBT, m, Sk);
    if (rc) { mct_kernel_destroy(*out); *out = nullptr; return rc; }
    borrow_trajectory(*out, collective);
    return MCT_OK;
}

int mct_mc3d_tagged_create_from_equation(mct_kernel_t *out, int s, double rho_all, double kBT, double m_s, mct_equation_t collective) {
    MCT_REQUIRE(out && collective && collective->solved && collective->mode == 0 && collective->d_gws && collective->kernel &&
                    collective->kernel->type == K_MC3D,
                MCT_BAD_ARGUMENT, "a solved multi-component equation is needed to chain a tagged kernel on it");
    mct_kernel_t hc = collective->kernel;
    const int Nk = hc->Nk, ns = hc->Ns;
    MCT_REQUIRE(s >= 0 && s < ns, MCT_BAD_ARGUMENT, "species index out of range (0-based)");
    MCT_CUDA(cudaSetDevice(hc->device));
    std::vector<double> k(Nk);
    MCT_CUDA(cudaMemcpy(k.data(), hc->d_k, Nk * sizeof(double), cudaMemcpyDeviceToHost));
    int rc = new_handle(out, K_TAGGED3D, hc->device, Nk, k.data());
    if (rc) return rc;
    mct_kernel_t h = *out;
    // vertex tables of the constructor (src/Kernels/MultiComponentKernels.jl:297-310); the c's are in the species weight
    const double dk = k[1] - k[0], pi = 3.14159265358979323846;
    const double pref = kBT * (dk * dk) * rho_all / (4 * m_s * ((2 * pi) * (2 * pi)));
    std::vector<double> V1((size_t)Nk * Nk), V2(V1.size()), V3(V1.size());
    for (int ip = 0; ip < Nk; ip++)
        for (int iq = 0; iq < Nk; iq++) {
            const double p = k[ip], q = k[iq], d2 = q * q - p * p;
            const size_t w = (size_t)iq + (size_t)Nk * ip;
            V1[w] = pref * p * q; V2[w] = pref * p * q * (d2 * d2); V3[w] = pref * 2 * p * q * d2;
        }
    rc = upload_tables(h, V1.data(), V2.data(), V3.data());
    if (!rc) {
        h->nt = collective->nt; h->owns_Fc = true;
        output_times(collective->opts, collective->nblocks, h->t_c);
        cudaError_t e = cudaMalloc(&h->d_Fc, (size_t)h->nt * Nk * sizeof(double));
        if (e != cudaSuccess) rc = cuda_fail(e, "mct_mc3d_tagged_create_from_equation", __FILE__, __LINE__);
    }
    if (!rc) rc = mc_species_weight_launch(collective->d_F, hc->d_C, Nk, ns, s, h->nt, h->d_Fc, h->stream);
    if (!rc) { cudaError_t e = cudaStreamSynchronize(h->stream); if (e != cudaSuccess) rc = cuda_fail(e, "species weight", __FILE__, __LINE__); }
    if (rc) { mct_kernel_destroy(h); *out = nullptr; }
    return rc;
}

int mct_mc3d_create(mct_kernel_t *out, int device, int Nk, int Ns, const double *k_array, const double *Cblocks, const double *prefactor) {
    MCT_REQUIRE(Ns >= 1 && Ns <= kNsMax, MCT_UNSUPPORTED, "the device supports 1 <= Ns <= 4 species");
    MCT_REQUIRE(Cblocks && prefactor, MCT_BAD_ARGUMENT, "NULL argument to mct_mc3d_create");
    int rc = new_handle(out, K_MC3D, device, Nk, k_array);
    if (rc) return rc;
    mct_kernel_t h = *out;
    h->Ns = Ns;
    const size_t nb = (size_t)Ns * Ns;
    cudaError_t e = cudaMalloc(&h->d_C, (nb * Nk + nb) * sizeof(double));
    if (e == cudaSuccess) { h->d_pref = h->d_C + nb * Nk; e = cudaMemcpyAsync(h->d_C, Cblocks, nb * Nk * sizeof(double), cudaMemcpyHostToDevice, h->stream); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->d_pref, prefactor, nb * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { mct_kernel_destroy(h); *out = nullptr; return cuda_fail(e, "mct_mc3d_create", __FILE__, __LINE__); }
    return MCT_OK;
}

int mct_ddim_create(mct_kernel_t *out, int device, int Nk, const double *k_array, double prefactor, const double *V, const double *J) {
    MCT_REQUIRE(V && J, MCT_BAD_ARGUMENT, "NULL argument to mct_ddim_create");
    int rc = new_handle(out, K_DDIM, device, Nk, k_array);
    if (rc) return rc;
    mct_kernel_t h = *out;
    const size_t n3 = (size_t)Nk * Nk * Nk;
    double *tmp = nullptr;
    cudaError_t e = cudaMalloc(&h->d_W, n3 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&tmp, 2 * n3 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(tmp, V, n3 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(tmp + n3, J, n3 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) { rc = ddim_build_table(Nk, prefactor, tmp, tmp + n3, h->d_W, h->stream); if (rc) e = cudaErrorUnknown; }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) { mct_kernel_destroy(h); *out = nullptr; return rc ? rc : cuda_fail(e, "mct_ddim_create", __FILE__, __LINE__); }
    return MCT_OK;
}

int mct_active_create(mct_kernel_t *out, int device, int Nk, const double *k_array, double prefactor, const double *J, const double *V2) {
    MCT_REQUIRE(V2 && J, MCT_BAD_ARGUMENT, "NULL argument to mct_active_create");
    int rc = new_handle(out, K_DDIM, device, Nk, k_array);
    if (rc) return rc;
    mct_kernel_t h = *out;
    const size_t n3 = (size_t)Nk * Nk * Nk;
    double *tmp = nullptr;
    cudaError_t e = cudaMalloc(&h->d_W, n3 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&tmp, 2 * n3 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(tmp, V2, n3 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(tmp + n3, J, n3 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) { rc = ddim_fuse_table(Nk, prefactor, tmp, tmp + n3, h->d_W, h->stream); if (rc) e = cudaErrorUnknown; }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) { mct_kernel_destroy(h); *out = nullptr; return rc ? rc : cuda_fail(e, "mct_active_create", __FILE__, __LINE__); }
    return MCT_OK;
}

int mct_active_tagged_create_from_equation(mct_kernel_t *out, int Nk, const double *k_array, double prefactor, const double *J,
                                           const double *V2, mct_equation_t collective) {
    MCT_REQUIRE(collective && collective->solved && collective->mode == 0 && collective->ns == 1 && collective->Nk == Nk,
                MCT_BAD_ARGUMENT, "a solved scalar-valued collective equation with the same Nk is needed");
    int rc = mct_active_create(out, collective->kernel->device, Nk, k_array, prefactor, J, V2);
    if (rc) return rc;
    borrow_trajectory(*out, collective);
    return MCT_OK;
}

int mct_kernel_set_collective_times(mct_kernel_t h, long nt, const double *t) {
    MCT_REQUIRE(h && t && nt >= 1, MCT_BAD_ARGUMENT, "NULL argument to mct_kernel_set_collective_times");
    MCT_REQUIRE(h->d_Fc != nullptr && nt == h->nt, MCT_BAD_ARGUMENT, "the kernel holds no collective trajectory of this length");
    h->t_c.assign(t, t + nt);
    return MCT_OK;
}

int mct_kernel_destroy(mct_kernel_t h) {
    if (!h) return MCT_OK;
    MCT_REQUIRE(h->live_equations == 0, MCT_BAD_ARGUMENT,
                "mct_kernel_destroy: equations created on this kernel are still alive (destroy them first)");
    if (h->borrowed_from) { h->borrowed_from->borrowers--; h->borrowed_from = nullptr; }
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_k); cudaFree(h->d_V1); cudaFree(h->d_Sk); cudaFree(h->d_C); cudaFree(h->d_W);
    if (h->owns_Fc) cudaFree(h->d_Fc);
    for (auto &kv : h->slabs) cudaFree(kv.second);
    cudaFree(h->d_scratch); cudaFree(h->d_io);
    for (auto &pe : h->pool) cudaFree(pe.first);
    // (the stream belongs to the device's shared set and stays)
    delete h;
    return MCT_OK;
}

int mct_kernel_eval_many(mct_kernel_t h, int n, const double *F, double *out, float *device_ms) {
    MCT_REQUIRE(h && F && out && n >= 1, MCT_BAD_ARGUMENT, "bad argument to mct_kernel_eval_many");
    MCT_REQUIRE(h->type == K_SC3D, MCT_UNSUPPORTED, "mct_kernel_eval_many supports the collective single-component kernel");
    MCT_CUDA(cudaSetDevice(h->device));
    int rc = ensure_raw_tables(h);
    if (rc) return rc;
    const int Nk = h->Nk;
    rc = ensure_scratch(h, sc3d_eval_scratch_doubles(Nk, n));
    if (!rc) rc = ensure_io(h, (size_t)2 * n * Nk);
    if (rc) return rc;
    double *dF = h->d_io, *dO = h->d_io + (size_t)n * Nk;
    MCT_CUDA(cudaMemcpyAsync(dF, F, (size_t)n * Nk * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    EventPair tm;
    MCT_CUDA(tm.create());
    MCT_CUDA(cudaEventRecord(tm.a, h->stream));
    rc = sc3d_eval_launch(Nk, n, h->d_k, h->d_V1, h->d_V2, h->d_V3, dF, Nk, dF, h->d_scratch, dO, h->stream);
    MCT_CUDA(cudaEventRecord(tm.b, h->stream));
    if (rc) return rc;
    MCT_CUDA(cudaMemcpyAsync(out, dO, (size_t)n * Nk * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    MCT_CUDA(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    MCT_CUDA(cudaEventElapsedTime(&ms, tm.a, tm.b));
    if (device_ms) *device_ms = ms;
    return MCT_OK;
}

int mct_kernel_eval(mct_kernel_t h, const double *F, long t_index, double *out) {
    MCT_REQUIRE(h && F && out, MCT_BAD_ARGUMENT, "bad argument to mct_kernel_eval");
    MCT_CUDA(cudaSetDevice(h->device));
    if (is_grid_kernel(h)) {
        const size_t dof = (size_t)h->Nk * h->Ns * h->Ns;
        mct_equation_s q;
        q.kernel = h; q.Nk = h->Nk; q.ns = h->Ns; q.mode = 2; q.nt = 1;
        std::vector<double> host(6 * dof, 0.0);
        int rc = grid_setup(&q, h, 2, host);
        double *dF = nullptr;
        if (!rc) { cudaError_t e = cudaMalloc(&dF, dof * sizeof(double)); if (e == cudaSuccess) e = cudaMemcpy(dF, F, dof * sizeof(double), cudaMemcpyHostToDevice); if (e != cudaSuccess) rc = cuda_fail(e, "mct_kernel_eval", __FILE__, __LINE__); }
        if (!rc && q.gp.Fc && (t_index < 0 || t_index >= h->nt)) { set_error("t is not a time of the collective solution (tDict KeyError)"); rc = MCT_BAD_ARGUMENT; }
        if (!rc) { q.gp.Fin = dF; q.gp.tix0 = t_index; rc = grid_run(&q, nullptr, nullptr, nullptr); }
        if (!rc) { cudaError_t e = cudaMemcpy(out, q.gp.F_out, dof * sizeof(double), cudaMemcpyDeviceToHost); if (e != cudaSuccess) rc = cuda_fail(e, "mct_kernel_eval", __FILE__, __LINE__); }
        cudaFree(dF); cudaFree(q.d_gws);
        return rc;
    }
    MCT_REQUIRE(h->type == K_SC3D || h->type == K_TAGGED3D, MCT_UNSUPPORTED, "kernel type not implemented on the device yet");
    int rc = ensure_raw_tables(h);
    if (rc) return rc;
    const int Nk = h->Nk;
    rc = ensure_scratch(h, sc3d_eval_scratch_doubles(Nk, 1));
    if (!rc) rc = ensure_io(h, (size_t)2 * Nk);
    if (rc) return rc;
    double *dF = h->d_io, *dO = h->d_io + Nk;
    MCT_CUDA(cudaMemcpyAsync(dF, F, Nk * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    const double *dF1 = dF;
    if (h->type == K_TAGGED3D) {
        MCT_REQUIRE(t_index >= 0 && t_index < h->nt, MCT_BAD_ARGUMENT, "t is not a time of the collective solution (tDict KeyError)");
        dF1 = h->d_Fc + (size_t)t_index * Nk;
    }
    rc = sc3d_eval_launch(Nk, 1, h->d_k, h->d_V1, h->d_V2, h->d_V3, dF1, 0, dF, h->d_scratch, dO, h->stream);
    if (rc) return rc;
    MCT_CUDA(cudaMemcpyAsync(out, dO, Nk * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    MCT_CUDA(cudaStreamSynchronize(h->stream));
    return MCT_OK;
}

int mct_td_output_len(const mct_td_options *o, long *nt) {
    MCT_REQUIRE(o && nt, MCT_BAD_ARGUMENT, "NULL argument");
    MCT_REQUIRE(o->N >= 2 && o->dt > 0 && o->t_max > 0, MCT_BAD_ARGUMENT, "bad solver options");   // same checks as mct_equation_create
    *nt = 1 + 2L * o->N + 2L * o->N * count_blocks(o);
    return MCT_OK;
}

int mct_equation_create(mct_equation_t *out, mct_kernel_t h, const mct_coeffs *c, const double *F0, const double *dF0,
                        const mct_td_options *o) {
    MCT_REQUIRE(out && h && c && F0 && dF0 && o, MCT_BAD_ARGUMENT, "NULL argument to mct_equation_create");
    *out = nullptr;
    MCT_REQUIRE(h->type == K_SC3D || h->type == K_TAGGED3D || is_grid_kernel(h), MCT_UNSUPPORTED, "kernel type not implemented on the device yet");
    MCT_REQUIRE(o->N >= 2 && o->dt > 0 && o->t_max > 0 && o->tolerance >= 0 && o->max_iterations >= 1, MCT_BAD_ARGUMENT,
                "bad solver options");
    MCT_REQUIRE(o->init_with_memory == 1, MCT_UNSUPPORTED, "init_with_memory=false is not supported on the device");
    MCT_CUDA(cudaSetDevice(h->device));
    const int Nk = h->Nk;
    mct_equation_t q = new mct_equation_s();
    q->kernel = h; q->Nk = Nk; q->opts = *o; q->nblocks = count_blocks(o);
    q->nt = 1 + 2L * o->N + 2L * o->N * q->nblocks;
    if (h->d_Fc != nullptr) {
        // the reference looks the collective solution up by time, `tDict[t]` (src/Kernels/ModeCouplingKernels.jl:394, 436):
        // every output time of this solve must be a time of the collective solution, at the same index
        bool ok = h->nt >= q->nt;
        if (ok && !h->t_c.empty()) {
            std::vector<double> t;
            output_times(q->opts, q->nblocks, t);
            for (long i = 0; ok && i < q->nt; i++) ok = (t[i] == h->t_c[i]);
        }
        if (!ok) {
            delete q;
            set_error("the time grid of this solve is not the time grid of the collective solution (tDict KeyError in the reference): "
                      "use the same N, dt and a t_max that is not larger");
            return MCT_BAD_ARGUMENT;
        }
    }
    if (is_grid_kernel(h)) {
        const int ns = h->Ns;
        const size_t dof = (size_t)Nk * ns * ns;
        q->ns = ns;
        std::vector<double> host(6 * dof, 0.0);
        int rc = expand_coef_blocks(c->alpha_kind, c->alpha_scalar, c->alpha, Nk, ns, host.data(), "alpha");
        if (!rc) rc = expand_coef_blocks(c->beta_kind, c->beta_scalar, c->beta, Nk, ns, host.data() + dof, "beta");
        if (!rc) rc = expand_coef_blocks(c->gamma_kind, c->gamma_scalar, c->gamma, Nk, ns, host.data() + 2 * dof, "gamma");
        if (rc) { delete q; return rc; }
        if (c->delta) memcpy(host.data() + 3 * dof, c->delta, dof * sizeof(double));
        memcpy(host.data() + 4 * dof, F0, dof * sizeof(double));
        memcpy(host.data() + 5 * dof, dF0, dof * sizeof(double));
        q->second_order = 0;
        for (size_t i = 0; i < dof; i++) if (host[i] != 0.0) q->second_order = 1;       // !iszero(alpha)  src/EulerSolver.jl:55
        rc = grid_setup(q, h, 0, host);
        if (rc) { mct_equation_destroy(q); return rc; }
        q->gp.N = o->N; q->gp.nblocks = q->nblocks; q->gp.second_order = q->second_order;
        q->gp.dt0 = o->dt; q->gp.tol = o->tolerance; q->gp.max_iter = o->max_iterations;
        h->live_equations++; q->counted = true;
        *out = q;
        return MCT_OK;
    }
    std::vector<double> host(6 * (size_t)Nk);
    int rc = expand_coef(c->alpha_kind, c->alpha_scalar, c->alpha, Nk, host.data(), "alpha");
    if (!rc) rc = expand_coef(c->beta_kind, c->beta_scalar, c->beta, Nk, host.data() + Nk, "beta");
    if (!rc) rc = expand_coef(c->gamma_kind, c->gamma_scalar, c->gamma, Nk, host.data() + 2 * Nk, "gamma");
    if (rc) { delete q; return rc; }
    if (c->delta) memcpy(host.data() + 3 * Nk, c->delta, Nk * sizeof(double));
    else std::fill(host.begin() + 3 * Nk, host.begin() + 4 * Nk, 0.0);
    memcpy(host.data() + 4 * Nk, F0, Nk * sizeof(double));
    memcpy(host.data() + 5 * Nk, dF0, Nk * sizeof(double));
    q->second_order = 0;
    for (int i = 0; i < Nk; i++) if (host[i] != 0.0) q->second_order = 1;   // !iszero(alpha)  src/EulerSolver.jl:55
    if (q->second_order) {
        for (int i = 0; i < Nk; i++)
            if (host[i] == 0.0) { delete q; set_error("alpha has zero and non-zero entries (singular)"); return MCT_BAD_ARGUMENT; }
    } else {
        for (int i = 0; i < Nk; i++)
            if (host[Nk + i] == 0.0) { delete q; set_error("alpha = 0 needs beta != 0"); return MCT_BAD_ARGUMENT; }
    }
    const size_t ring = (size_t)4 * Nk * 4 * o->N, traj = (size_t)q->nt * Nk;
    q->coef_bytes = host.size() * sizeof(double); q->ring_bytes = ring * sizeof(double); q->traj_bytes = 2 * traj * sizeof(double);
    cudaError_t e = pool_alloc(h, (void **)&q->d_coef, q->coef_bytes);
    if (e == cudaSuccess) e = pool_alloc(h, (void **)&q->d_ring, q->ring_bytes);
    if (e == cudaSuccess) e = pool_alloc(h, (void **)&q->d_F, q->traj_bytes);
    if (e == cudaSuccess) e = pool_alloc(h, (void **)&q->d_status, 64);
    if (e == cudaSuccess) {
        q->d_K = q->d_F + traj;
        q->d_evals = (long long *)((char *)q->d_status + 16);
        e = cudaMemcpy(q->d_coef, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaMemset(q->d_ring, 0, ring * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(q->d_status, 0, 64);
    h->live_equations++; q->counted = true;
    if (e != cudaSuccess) { mct_equation_destroy(q); return cuda_fail(e, "mct_equation_create", __FILE__, __LINE__); }
    *out = q;
    return MCT_OK;
}

int mct_equation_destroy(mct_equation_t q) {
    if (!q) return MCT_OK;
    MCT_REQUIRE(q->borrowers == 0, MCT_BAD_ARGUMENT,
                "mct_equation_destroy: a tagged kernel still reads this equation's device-resident trajectory (destroy it first)");
    if (q->counted && q->kernel) q->kernel->live_equations--;
    if (q->kernel) cudaSetDevice(q->kernel->device);
    if (q->d_gws) { cudaFree(q->d_gws); q->d_F = nullptr; }
    for (int r = 0; r < kMaxShardRanks; r++) if (q->peer_maps[r]) cudaIpcCloseMemHandle(q->peer_maps[r]);
    cudaFree(q->d_xshard);
    if (q->traj_bytes && q->kernel) {                  // single-component equation: buffers go back to the kernel's pool
        pool_free(q->kernel, q->d_coef, q->coef_bytes); pool_free(q->kernel, q->d_ring, q->ring_bytes);
        pool_free(q->kernel, q->d_F, q->traj_bytes); pool_free(q->kernel, q->d_status, 64);
    } else {
        cudaFree(q->d_coef); cudaFree(q->d_ring); cudaFree(q->d_F); cudaFree(q->d_status);
    }
    delete q;
    return MCT_OK;
}

int mct_equation_solve_batch(int n, mct_equation_t *eqs, int *status, long *kernel_evals, float *device_ms) {
    MCT_REQUIRE(n >= 1 && eqs, MCT_BAD_ARGUMENT, "empty batch");
    for (int i = 0; i < n; i++) MCT_REQUIRE(eqs[i] && eqs[i]->mode == 0, MCT_BAD_ARGUMENT, "bad equation handle");
    if (is_grid_kernel(eqs[0]->kernel)) {          // block-valued kernels: one cooperative launch per equation
        int worst = MCT_OK;
        float tot = 0.f;
        for (int i = 0; i < n; i++) {
            long e = 0; float ms = 0.f;
            int rc = mct_equation_solve(eqs[i], &e, &ms);
            if (status) status[i] = rc;
            if (kernel_evals) kernel_evals[i] = e;
            tot += ms;
            if (rc != MCT_OK && worst == MCT_OK) worst = rc;
        }
        if (device_ms) *device_ms = tot;
        return worst;
    }
    std::vector<long long> ev(n, 0);
    int rc = run_batch(n, eqs, 0, eqs[0]->opts.tolerance, eqs[0]->opts.max_iterations, status, ev.data(), device_ms);
    if (kernel_evals) for (int i = 0; i < n; i++) kernel_evals[i] = (long)ev[i];
    return rc;
}

int mct_equation_solve(mct_equation_t q, long *kernel_evals, float *device_ms) {
    MCT_REQUIRE(q && q->kernel, MCT_BAD_ARGUMENT, "NULL equation");
    if (is_grid_kernel(q->kernel)) {
        long long ev = 0; int st = 0;
        int rc = grid_run(q, &ev, &st, device_ms);
        if (kernel_evals) *kernel_evals = (long)ev;
        q->solved = (rc == MCT_OK);
        if (rc == MCT_NOT_CONVERGED)
            set_error("Iteration did not converge. Either increase the number of time steps before a time doubling, or choose a "
                      "different memory kernel.");
        return rc;
    }
    int st = 0;
    long ev = 0;
    int rc = mct_equation_solve_batch(1, &q, &st, &ev, device_ms);
    if (kernel_evals) *kernel_evals = ev;
    return rc;
}

int mct_equation_download(mct_equation_t q, double *t_out, double *F_out, double *K_out) {
    MCT_REQUIRE(q, MCT_BAD_ARGUMENT, "NULL equation");
    MCT_CUDA(cudaSetDevice(q->kernel->device));
    const size_t traj = (size_t)q->nt * q->Nk * q->ns * q->ns;
    if (F_out) MCT_CUDA(cudaMemcpy(F_out, q->d_F, traj * sizeof(double), cudaMemcpyDeviceToHost));
    if (K_out) MCT_CUDA(cudaMemcpy(K_out, q->d_K, traj * sizeof(double), cudaMemcpyDeviceToHost));
    if (t_out) {
        std::vector<double> t;
        output_times(q->opts, q->nblocks, t);
        memcpy(t_out, t.data(), t.size() * sizeof(double));
    }
    return MCT_OK;
}

int mct_equation_download_slice(mct_equation_t q, long n_times, const long *times, int n_dofs, const int *dofs,
                                double *t_out, double *F_out, double *K_out) {
    MCT_REQUIRE(q && q->mode == 0, MCT_BAD_ARGUMENT, "NULL equation");
    const int dof = q->Nk * q->ns * q->ns;
    if (!times) n_times = q->nt;
    if (!dofs) n_dofs = dof;
    MCT_REQUIRE(n_times >= 1 && n_dofs >= 1, MCT_BAD_ARGUMENT, "empty slice");
    for (long i = 0; times && i < n_times; i++) MCT_REQUIRE(times[i] >= 0 && times[i] < q->nt, MCT_BAD_ARGUMENT, "time index out of range");
    for (int j = 0; dofs && j < n_dofs; j++) MCT_REQUIRE(dofs[j] >= 0 && dofs[j] < dof, MCT_BAD_ARGUMENT, "wavenumber index out of range");
    mct_kernel_t h = q->kernel;
    MCT_CUDA(cudaSetDevice(h->device));
    const size_t n = (size_t)n_times * n_dofs;
    const size_t idx_doubles = (size_t)(times ? n_times : 0) + (size_t)(dofs ? (n_dofs + 1) / 2 : 0);
    int rc = ensure_io(h, n + idx_doubles + 2);
    if (rc) return rc;
    long *d_times = times ? (long *)(h->d_io + n) : nullptr;
    int *d_dofs = dofs ? (int *)(h->d_io + n + (times ? n_times : 0)) : nullptr;
    if (times) MCT_CUDA(cudaMemcpyAsync(d_times, times, n_times * sizeof(long), cudaMemcpyHostToDevice, h->stream));
    if (dofs) MCT_CUDA(cudaMemcpyAsync(d_dofs, dofs, n_dofs * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    for (int pass = 0; pass < 2; pass++) {
        double *out = pass ? K_out : F_out;
        if (!out) continue;
        rc = gather_slice_launch(pass ? q->d_K : q->d_F, dof, d_times, n_times, d_dofs, n_dofs, h->d_io, h->stream);
        if (rc) return rc;
        MCT_CUDA(cudaMemcpyAsync(out, h->d_io, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        MCT_CUDA(cudaStreamSynchronize(h->stream));
    }
    if (t_out) {
        std::vector<double> t;
        output_times(q->opts, q->nblocks, t);
        for (long i = 0; i < n_times; i++) t_out[i] = t[times ? times[i] : i];
    }
    return MCT_OK;
}

int mct_equation_relaxation_times(mct_equation_t q, int n_k, const int *ik, double threshold, int mode, double *tau_out) {
    MCT_REQUIRE(q && ik && tau_out && n_k >= 1, MCT_BAD_ARGUMENT, "NULL argument to mct_equation_relaxation_times");
    MCT_REQUIRE(mode == 0 || mode == 1, MCT_BAD_ARGUMENT, "the mode is not known. it must be either :lin (1), :log (0).");
    std::vector<double> F((size_t)q->nt * n_k), t;
    int rc = mct_equation_download_slice(q, 0, nullptr, n_k, ik, nullptr, F.data(), nullptr);
    if (rc) return rc;
    output_times(q->opts, q->nblocks, t);
    for (int j = 0; j < n_k; j++) {                                   // src/RelaxationTime.jl:14-39 per requested wavenumber
        double tau = INFINITY;
        const double f0 = F[j];
        for (long it = 0; it < q->nt; it++) {
            const double y0 = F[(size_t)it * n_k + j] / f0;
            if (y0 > threshold) continue;
            if (it == 0) { tau = 0.0; break; }
            const double y1 = F[(size_t)(it - 1) * n_k + j] / f0, y = threshold;
            const double x0 = mode == 0 ? std::log(t[it]) : t[it], x1 = mode == 0 ? std::log(t[it - 1]) : t[it - 1];
            const double x = ((y - y0) * (x1 - x0) + x0 * (y1 - y0)) / (y1 - y0);
            tau = mode == 0 ? std::exp(x) : x;
            break;
        }
        tau_out[j] = tau;
    }
    return MCT_OK;
}

// Download of the trajectory WHILE the persistent kernel produces it (single-component solves through mct_td_solve).
// The trajectory is pre-filled with a NaN payload no arithmetic produces; every output word is written exactly once, so a
// word that is not the sentinel is final -- no ordering between the kernel's stores is assumed.  Per time block (2N
// rows) the host polls the block's last row (8 Nk bytes), then copies the block on a second stream and accepts it when no
// sentinel is left in it.  70 MB hide behind a 330 ms solve even on a slow PCIe link; whatever is still missing when
// the kernel ends (or after an aborted solve) is copied by the plain download afterwards.
static long stream_download(mct_equation_t q, double *F_out, double *K_out, cudaEvent_t done) {
    const int Nk = q->Nk, n2 = 2 * q->opts.N;
    const unsigned long long sentinel = 0xFFFFFFFFFFFFFFFFull;
    cudaStream_t sb = nullptr;
    if (cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking) != cudaSuccess) return 0;
    auto clean = [&](const double *p, size_t n) {
        const unsigned long long *u = (const unsigned long long *)p;
        for (size_t i = 0; i < n; i++) if (u[i] == sentinel) return false;
        return true;
    };
    long row = 0;                                       // rows [0, row) are on the host
    bool alive = true;
    while (alive && row < q->nt) {
        const long rows = (row == 0) ? 1 + n2 : n2;     // t = 0 and the bootstrap first, then one time block at a time
        const size_t off = (size_t)row * Nk, n = (size_t)rows * Nk, last = off + n - Nk;
        bool got = false;
        while (!got) {
            const bool finished = cudaEventQuery(done) == cudaSuccess;
            if (cudaMemcpyAsync(F_out + last, q->d_F + last, Nk * sizeof(double), cudaMemcpyDeviceToHost, sb) != cudaSuccess ||
                cudaMemcpyAsync(K_out + last, q->d_K + last, Nk * sizeof(double), cudaMemcpyDeviceToHost, sb) != cudaSuccess ||
                cudaStreamSynchronize(sb) != cudaSuccess) { alive = false; break; }
            if (clean(F_out + last, Nk) && clean(K_out + last, Nk)) {
                if (cudaMemcpyAsync(F_out + off, q->d_F + off, n * sizeof(double), cudaMemcpyDeviceToHost, sb) != cudaSuccess ||
                    cudaMemcpyAsync(K_out + off, q->d_K + off, n * sizeof(double), cudaMemcpyDeviceToHost, sb) != cudaSuccess ||
                    cudaStreamSynchronize(sb) != cudaSuccess) { alive = false; break; }
                got = clean(F_out + off, n) && clean(K_out + off, n);
            }
            if (!got) {
                if (finished) { alive = false; break; }   // the kernel is done and the block is not there: failed solve
                usleep(300);
            }
        }
        if (got) row += rows;
    }
    cudaStreamDestroy(sb);
    return row;
}

int mct_td_solve(mct_kernel_t h, const mct_coeffs *c, const double *F0, const double *dF0, const mct_td_options *o,
                 double *t_out, double *F_out, double *K_out, long *kernel_evals) {
    mct_equation_t q = nullptr;
    const auto w0 = std::chrono::steady_clock::now();
    int rc = mct_equation_create(&q, h, c, F0, dF0, o);
    if (rc) return rc;
    const auto w1 = std::chrono::steady_clock::now();
    float dev_ms = 0.f;
    long have = 0;
    const bool streamed = !is_grid_kernel(h) && F_out && K_out && env_int("MCT_STREAM_DOWNLOAD", 1);
    if (streamed) {
        const size_t traj = (size_t)q->nt * q->Nk;
        cudaError_t e = cudaMemsetAsync(q->d_F, 0xFF, traj * sizeof(double), h->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(q->d_K, 0xFF, traj * sizeof(double), h->stream);
        if (e != cudaSuccess) { mct_equation_destroy(q); return cuda_fail(e, "mct_td_solve", __FILE__, __LINE__); }
        std::function<void(cudaEvent_t)> overlap = [&](cudaEvent_t done) { have = stream_download(q, F_out, K_out, done); };
        g_overlap = &overlap;
        rc = mct_equation_solve(q, kernel_evals, &dev_ms);
        g_overlap = nullptr;
    } else {
        rc = mct_equation_solve(q, kernel_evals, &dev_ms);
    }
    const auto w2 = std::chrono::steady_clock::now();
    const long streamed_rows = have;
    int rc2 = MCT_OK;
    if (have < q->nt) {                                  // the rest (everything, when nothing was streamed)
        const size_t off = (size_t)have * q->Nk * q->ns * q->ns, n = (size_t)(q->nt - have) * q->Nk * q->ns * q->ns;
        cudaError_t e = cudaSuccess;
        if (F_out) e = cudaMemcpy(F_out + off, q->d_F + off, n * sizeof(double), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && K_out) e = cudaMemcpy(K_out + off, q->d_K + off, n * sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc2 = cuda_fail(e, "mct_td_solve download", __FILE__, __LINE__);
    }
    if (rc2 == MCT_OK && t_out) rc2 = mct_equation_download(q, t_out, nullptr, nullptr);
    const auto w3 = std::chrono::steady_clock::now();
    const long nt_all = q->nt;
    mct_equation_destroy(q);
    if (env_int("MCT_VERBOSE", 0)) {
        const auto w4 = std::chrono::steady_clock::now();
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[mct] mct_td_solve: create %.2f ms, solve %.2f ms (kernel %.2f ms, %ld of %ld rows streamed), rest of download %.2f ms, destroy %.2f ms\n",
                ms(w0, w1), ms(w1, w2), dev_ms, streamed_rows, nt_all, ms(w2, w3), ms(w3, w4));
    }
    return rc ? rc : rc2;
}

int mct_equation_shard_prepare(mct_equation_t q, int rank, int world, unsigned char *handle_out) {
    MCT_REQUIRE(q && handle_out, MCT_BAD_ARGUMENT, "NULL argument to mct_equation_shard_prepare");
    MCT_REQUIRE(q->kernel && q->kernel->type == K_MC3D, MCT_UNSUPPORTED, "k-sharding is implemented for the multi-component kernel");
    MCT_REQUIRE(world >= 2 && world <= kMaxShardRanks && rank >= 0 && rank < world, MCT_BAD_ARGUMENT, "bad rank / world size (2..8)");
    MCT_CUDA(cudaSetDevice(q->kernel->device));
    const size_t xwords = grid_lsum_words(q->Nk, q->ns);
    // 4 rotating buffers of task sums + 4 arrival counters (td_grid.cu, phase_lines / phase_scan)
    if (!q->d_xshard) MCT_CUDA(raw_cuda_malloc((void **)&q->d_xshard, (4 * xwords + 8) * sizeof(double)));     // exported through CUDA IPC: a real allocation
    MCT_CUDA(cudaMemset(q->d_xshard, 0xFF, 4 * xwords * sizeof(double)));          // every slot armed with the sentinel
    MCT_CUDA(cudaMemset(q->d_xshard + 4 * xwords, 0, 8 * sizeof(double)));
    cudaIpcMemHandle_t hd;
    MCT_CUDA(cudaIpcGetMemHandle(&hd, q->d_xshard));
    static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle_out, &hd, 64);
    q->gp.rank = rank; q->gp.world = world;
    return grid_upload_deal(q);
}

int mct_equation_shard_connect(mct_equation_t q, const unsigned char *handles) {
    MCT_REQUIRE(q && handles && q->d_xshard && q->gp.world >= 2, MCT_BAD_ARGUMENT, "call mct_equation_shard_prepare first");
    MCT_CUDA(cudaSetDevice(q->kernel->device));
    for (int r = 0; r < q->gp.world; r++) {
        if (r == q->gp.rank) { q->gp.xpeer[r] = q->d_xshard; continue; }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, handles + 64 * r, 64);
        void *p = nullptr;
        MCT_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
        q->peer_maps[r] = p;
        q->gp.xpeer[r] = (double *)p;
    }
    return MCT_OK;
}

// Shared by the single- and multi-component MSD entry points: Fc = trajectory the tagged trajectory is contracted with
// (collective F with weights q^4 c^2, or the species weight G with weights q^4).
static int msd_solve_common(mct_equation_t tagged, const double *d_Fc, const double *Sk, const double *d_Cmix, int ns, int sp,
                            int species_terms, double rho, double kBT, double m,
                            double alpha, double beta, double gamma, double delta, double MSD0, double dMSD0,
                            double *t_out, double *msd_out, double *K_out, long *kernel_evals) {
    MCT_REQUIRE(alpha != 0.0 || beta != 0.0, MCT_BAD_ARGUMENT, "alpha and beta are both zero");
    mct_kernel_t h = tagged->kernel;
    const int Nk = tagged->Nk;
    const long nt = tagged->nt;
    MCT_CUDA(cudaSetDevice(h->device));
    std::vector<double> kh(Nk), ck(Nk, 0.0);
    MCT_CUDA(cudaMemcpy(kh.data(), h->d_k, Nk * sizeof(double), cudaMemcpyDeviceToHost));
    for (int i = 0; Sk && i < Nk; i++) ck[i] = (Sk[i] - 1.0) / (rho * Sk[i]);       // ModeCouplingKernels.jl:566
    const double dk = kh[1] - kh[0];
    const double pi = 3.14159265358979323846;
    double *d = nullptr;                                  // Ck [Nk], Ktab [nt], t/F/K out [3 nt], status, evals
    MCT_CUDA(cudaMalloc(&d, ((size_t)Nk + 4 * nt + 4) * sizeof(double)));
    cudaError_t e = cudaMemcpy(d, ck.data(), Nk * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(d + Nk + 4 * nt, 0, 4 * sizeof(double));
    if (e != cudaSuccess) { cudaFree(d); return cuda_fail(e, "mct_msd3d_solve setup", __FILE__, __LINE__); }
    MsdParams P{};
    P.Nk = Nk; P.N = tagged->opts.N; P.nblocks = tagged->nblocks; P.nt = nt;
    P.dt0 = tagged->opts.dt; P.tol = tagged->opts.tolerance; P.max_iter = tagged->opts.max_iterations;
    P.alpha = alpha; P.beta = beta; P.gamma = gamma; P.delta = delta; P.F0 = MSD0; P.dF0 = dMSD0;
    P.scale = dk * rho * kBT / (6.0 * (pi * pi) * m);                               // :583 / MultiComponentKernels.jl:486
    P.kvec = h->d_k; P.Ck = d; P.Fc = d_Fc; P.Fs = tagged->d_F;
    P.Cmix = d_Cmix; P.ns = ns; P.s = sp; P.species_terms = species_terms;
    P.Ktab = d + Nk; P.t_out = d + Nk + nt; P.F_out = d + Nk + 2 * nt; P.K_out = d + Nk + 3 * nt;
    P.status = (int *)(d + Nk + 4 * nt); P.kernel_evals = (long long *)(d + Nk + 4 * nt + 2);
    int rc = msd3d_launch(P, h->stream);
    if (rc == MCT_OK) {
        e = cudaStreamSynchronize(h->stream);
        if (e == cudaSuccess) e = cudaMemcpy(t_out, P.t_out, nt * sizeof(double), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(msd_out, P.F_out, nt * sizeof(double), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(K_out, P.K_out, nt * sizeof(double), cudaMemcpyDeviceToHost);
        int st = 0; long long ev = 0;
        if (e == cudaSuccess) e = cudaMemcpy(&st, P.status, sizeof(int), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(&ev, P.kernel_evals, sizeof(long long), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "mct_msd3d_solve", __FILE__, __LINE__);
        else {
            if (kernel_evals) *kernel_evals = (long)ev;
            rc = st;
            if (rc == MCT_NOT_CONVERGED) set_error("Iteration did not converge.");
        }
    }
    cudaFree(d);
    return rc;
}

int mct_msd3d_solve(mct_equation_t collective, mct_equation_t tagged, double rho, double kBT, double m, const double *Sk,
                    double alpha, double beta, double gamma, double delta, double MSD0, double dMSD0,
                    double *t_out, double *msd_out, double *K_out, long *kernel_evals) {
    MCT_REQUIRE(collective && tagged && Sk && t_out && msd_out && K_out, MCT_BAD_ARGUMENT, "NULL argument to mct_msd3d_solve");
    MCT_REQUIRE(collective->solved && tagged->solved, MCT_BAD_ARGUMENT, "both equations must have been solved on the device");
    MCT_REQUIRE(collective->ns == 1 && tagged->ns == 1 && !collective->d_gws && !tagged->d_gws, MCT_UNSUPPORTED,
                "mct_msd3d_solve takes single-component solutions (mixtures: mct_msd_mc3d_solve)");
    MCT_REQUIRE(collective->Nk == tagged->Nk && collective->nt == tagged->nt && collective->opts.N == tagged->opts.N &&
                    collective->opts.dt == tagged->opts.dt,
                MCT_BAD_ARGUMENT, "the two solutions must share the wavenumber grid and the time grid (tDict lookup)");
    MCT_REQUIRE(collective->kernel->device == tagged->kernel->device, MCT_BAD_ARGUMENT, "the two solutions live on different devices");
    return msd_solve_common(tagged, collective->d_F, Sk, nullptr, 1, 0, 1, rho, kBT, m, alpha, beta, gamma, delta, MSD0, dMSD0, t_out,
                            msd_out, K_out, kernel_evals);
}

int mct_msd_mc3d_solve(mct_equation_t collective, mct_equation_t tagged, int s, int full_species_sum, double rho_all, double kBT,
                       double m_s, double alpha, double beta, double gamma, double delta, double MSD0, double dMSD0,
                       double *t_out, double *msd_out, double *K_out, long *kernel_evals) {
    MCT_REQUIRE(collective && tagged && t_out && msd_out && K_out, MCT_BAD_ARGUMENT, "NULL argument to mct_msd_mc3d_solve");
    MCT_REQUIRE(collective->solved && collective->mode == 0 && collective->d_gws && collective->kernel->type == K_MC3D,
                MCT_BAD_ARGUMENT, "a solved multi-component equation is needed");
    MCT_REQUIRE(tagged->solved && tagged->ns == 1 && !tagged->d_gws, MCT_BAD_ARGUMENT, "a solved tagged-particle equation is needed");
    MCT_REQUIRE(collective->Nk == tagged->Nk && collective->nt == tagged->nt && collective->opts.N == tagged->opts.N &&
                    collective->opts.dt == tagged->opts.dt && collective->kernel->device == tagged->kernel->device,
                MCT_BAD_ARGUMENT, "the two solutions must share the wavenumber grid, the time grid (tDict lookup) and the device");
    const int ns = collective->kernel->Ns;
    MCT_REQUIRE(s >= 0 && s < ns, MCT_BAD_ARGUMENT, "species index out of range (0-based)");
    return msd_solve_common(tagged, collective->d_F, nullptr, collective->kernel->d_C, ns, s, full_species_sum ? ns : 1, rho_all, kBT,
                            m_s, alpha, beta, gamma, delta, MSD0, dMSD0, t_out, msd_out, K_out, kernel_evals);
}

int mct_steady_state(mct_kernel_t h, const double *gamma, const double *F0, double tolerance, long max_iterations,
                     double *F_out, double *K_out, long *iterations) {
    MCT_REQUIRE(h && gamma && F0 && F_out && K_out, MCT_BAD_ARGUMENT, "NULL argument to mct_steady_state");
    MCT_CUDA(cudaSetDevice(h->device));
    if (is_grid_kernel(h)) {
        const size_t dof = (size_t)h->Nk * h->Ns * h->Ns;
        mct_equation_s q;
        q.kernel = h; q.Nk = h->Nk; q.ns = h->Ns; q.mode = 1; q.nt = 1;
        std::vector<double> host(6 * dof, 0.0);
        memcpy(host.data() + 2 * dof, gamma, dof * sizeof(double));
        memcpy(host.data() + 4 * dof, F0, dof * sizeof(double));
        int rc = grid_setup(&q, h, 1, host);
        long long it = 0;
        if (!rc) { q.gp.tol = tolerance; q.gp.max_iter = max_iterations; rc = grid_run(&q, &it, nullptr, nullptr); }
        if (rc == MCT_OK || rc == MCT_NOT_CONVERGED) {
            cudaMemcpy(F_out, q.gp.F_out, dof * sizeof(double), cudaMemcpyDeviceToHost);
            cudaMemcpy(K_out, q.gp.K_out, dof * sizeof(double), cudaMemcpyDeviceToHost);
        }
        if (iterations) *iterations = (long)it;
        if (rc == MCT_NOT_CONVERGED) set_error("The recursive iteration did not converge.");
        cudaFree(q.d_gws);
        return rc;
    }
    MCT_REQUIRE(h->type == K_SC3D, MCT_UNSUPPORTED, "steady state is implemented for the collective single-component kernel");
    const int Nk = h->Nk;
    mct_equation_s q;
    q.kernel = h; q.Nk = Nk; q.mode = 1; q.nt = 1;
    q.opts.N = 2; q.opts.dt = 1.0; q.nblocks = 0;
    std::vector<double> host(6 * (size_t)Nk, 0.0);
    memcpy(host.data() + 2 * Nk, gamma, Nk * sizeof(double));
    memcpy(host.data() + 4 * Nk, F0, Nk * sizeof(double));
    MCT_CUDA(cudaMalloc(&q.d_coef, (host.size() + 2 * Nk) * sizeof(double)));
    q.d_F = q.d_coef + host.size(); q.d_K = q.d_F + Nk;
    cudaError_t e = cudaMalloc(&q.d_status, 64);
    if (e == cudaSuccess) e = cudaMemset(q.d_status, 0, 64);
    if (e == cudaSuccess) e = cudaMemcpy(q.d_coef, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice);
    q.d_evals = (long long *)((char *)q.d_status + 16);
    int rc = MCT_OK;
    long long it = 0;
    int st = 0;
    if (e != cudaSuccess) rc = cuda_fail(e, "mct_steady_state", __FILE__, __LINE__);
    if (!rc) {
        mct_equation_t qp = &q;
        rc = run_batch(1, &qp, 1, tolerance, max_iterations, &st, &it, nullptr);
    }
    if (rc == MCT_OK || rc == MCT_NOT_CONVERGED) {
        cudaMemcpy(F_out, q.d_F, Nk * sizeof(double), cudaMemcpyDeviceToHost);
        cudaMemcpy(K_out, q.d_K, Nk * sizeof(double), cudaMemcpyDeviceToHost);
    }
    if (iterations) *iterations = (long)it;
    if (rc == MCT_NOT_CONVERGED) set_error("The recursive iteration did not converge.");
    cudaFree(q.d_coef); cudaFree(q.d_status);
    return rc;
}

}  // extern "C"
