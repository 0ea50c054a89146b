/*
 * mct.h -- C ABI of libmct_sm100a.so: the B200 (sm_100a) replacement for the hot path of
 * ModeCouplingTheory.jl v0.8.9 -- the wavenumber-resolved memory-kernel evaluation and the Fuchs
 * time-doubling step that consumes it.  Plain pointers and sizes only; every entry point returns an int
 * status (MCT_OK = 0) and never calls back into the host language.  Citations (file:line) are relative to
 * the reference repository root.
 *
 * Data conventions (all Float64):
 *   - Vectors over wavenumber are length Nk.  Matrices are Julia column-major: V[iq + ip*Nk] = V[iq,ip].
 *   - Multi-component quantities are Nk blocks of Ns x Ns doubles, column-major inside a block, i.e. the
 *     memory of a Julia Vector{SMatrix{Ns,Ns,Float64}}: X[ik*Ns*Ns + a + b*Ns] = X[ik][a,b].
 *   - "dof" = Nk*Ns*Ns (Ns = 1 for the single-component kernels; K is then the `.diag` of the reference's
 *     Diagonal).
 *   - Trajectories are [nt][dof] row-major: the concatenation of the reference's Vector of per-time arrays.
 *   - Unless a function name ends in _dev, every pointer is a HOST pointer owned by the caller; the library
 *     copies in/out.  Handles own device memory and one CUDA stream; a handle must not be used from two
 *     host threads at once (the reference's kernel objects are not re-entrant either:
 *     src/Kernels/ModeCouplingKernels.jl:190-208 mutates its own scratch).
 *
 * There is no CPU fallback: without a usable sm_100 device every call returns MCT_CUDA_ERROR.
 */
#ifndef MCT_H
#define MCT_H

#if defined(__GNUC__)
#define MCT_API __attribute__((visibility("default")))
#else
#define MCT_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* status codes and the reference behaviour they map to */
#define MCT_OK 0
#define MCT_NOT_CONVERGED 1 /* DomainError("Iteration did not converge...")  src/TimeDoublingSolver.jl:406-408;
                               error("The recursive iteration did not converge...") src/SteadyStateMemoryEquation.jl:76-78 */
#define MCT_BAD_ARGUMENT 2  /* @assert on the grid, src/Kernels/ModeCouplingKernels.jl:103-104; size mismatches */
#define MCT_UNSUPPORTED 3   /* non-Float64, non-diagonal C1 (LinearSolve branch, src/TimeDoublingSolver.jl:73-77,338-344),
                               custom update_coefficients! (src/MemoryEquation.jl:44), ismutable=false */
#define MCT_CUDA_ERROR 4    /* CUDA / NCCL failure, or no sm_100 device */

typedef struct mct_kernel_s *mct_kernel_t;     /* replaces a MemoryKernel object (src/Kernels.jl:1) */
typedef struct mct_equation_s *mct_equation_t; /* a MemoryEquation resident on the device (src/MemoryEquation.jl:23-52) */

MCT_API const char *mct_last_error(void); /* message of the last failing call on this thread */
MCT_API int mct_device_count(int *count);
MCT_API int mct_version(void);
MCT_API long mct_launch_count(void); /* CUDA kernels launched by this library in this process so far (benchmark evidence) */

/* ---- memory kernels: plug-in point #2, `evaluate_kernel!(out, kernel, F, t)`  src/Kernels.jl:10-26 ---- */

/* ModeCouplingKernel3D from the tables the untouched Julia constructor produced:
 * kernel.k_array, kernel.V1, kernel.V2, kernel.V3  (src/Kernels/ModeCouplingKernels.jl:2-17, 94-129). */
MCT_API int mct_sc3d_create(mct_kernel_t *out, int device, int Nk, const double *k_array, const double *V1, const double *V2,
                    const double *V3);
/* Same kernel, tables built on the device in the constructor's own operation order (:105-126) -- avoids
 * shipping 3*Nk^2 doubles over PCIe for packing-fraction sweeps. */
MCT_API int mct_sc3d_create_from_sk(mct_kernel_t *out, int device, int Nk, double rho, double kBT, double m, const double *k_array,
                            const double *Sk);
/* TaggedModeCouplingKernel3D (:265-282, 390-467).  `Fc_traj` is sol.F of the collective solution, [nt][Nk];
 * the reference's tDict[t] float lookup (:394,:436) becomes the integer index of t in sol.t. */
MCT_API int mct_sc3d_tagged_create(mct_kernel_t *out, int device, int Nk, const double *k_array, const double *V1, const double *V2,
                           const double *V3, long nt, const double *Fc_traj);
MCT_API int mct_sc3d_tagged_create_from_sk(mct_kernel_t *out, int device, int Nk, double rho, double kBT, double m,
                                   const double *k_array, const double *Sk, long nt, const double *Fc_traj);
/* The reference finds the collective solution at the time of an evaluation with `tDict[t]`, a Dict from the collective
 * solve's output times to their index (:394, :436): a tagged solve on another time grid throws KeyError.  Pass
 * get_t(sol) (nt doubles) here after one of the two constructors above and mct_equation_create checks every output time of
 * the tagged solve against it (MCT_BAD_ARGUMENT on a mismatch); without it only the length of the trajectory is checked.
 * Kernels chained on a device-resident solution (below) inherit the grid of that solve. */
MCT_API int mct_kernel_set_collective_times(mct_kernel_t kernel, long nt, const double *t);
/* Tagged kernel chained on the device-resident trajectory of a finished collective solve (no PCIe round trip).  The
 * kernel reads that trajectory in place: mct_equation_destroy(collective) returns MCT_BAD_ARGUMENT until the kernel is
 * destroyed. */
MCT_API int mct_sc3d_tagged_create_from_equation(mct_kernel_t *out, int Nk, double rho, double kBT, double m, const double *k_array,
                                         const double *Sk, mct_equation_t collective);
/* MultiComponentModeCouplingKernel3D from kernel.C (Vector{SMatrix}) and kernel.prefactor (Ns x Ns)
 * (src/Kernels/MultiComponentKernels.jl:45-60, 84-131). */
MCT_API int mct_mc3d_create(mct_kernel_t *out, int device, int Nk, int Ns, const double *k_array, const double *C,
                    const double *prefactor);
/* dDimModeCouplingKernel from kernel.prefactor, kernel.V, kernel.J (Nk x Nk x Nk column-major [iq,ik,ip])
 * (src/Kernels/ModeCouplingKernels.jl:19-71). */
MCT_API int mct_ddim_create(mct_kernel_t *out, int device, int Nk, const double *k_array, double prefactor, const double *V,
                    const double *J);
/* ActiveMCTKernel from kernel.prefactor, kernel.J, kernel.V2 (Nk x Nk x Nk column-major [l,j,i], the OUTPUT index i last)
 * (src/Kernels/ActiveMCTKernel.jl:1-49; evaluate_kernel! :51-61 is the dDim triple sum over another table). */
MCT_API int mct_active_create(mct_kernel_t *out, int device, int Nk, const double *k_array, double prefactor, const double *J,
                      const double *V2);
/* TaggedActiveMCTKernel (src/Kernels/ActiveMCTKernel.jl:63-136) chained on the device-resident collective solution:
 * out[i] = prefactor sum_{j,l} J[l,j,i] V2[l,j,i] F_collective(t)[l] Fs[j]  (:125-136), tDict lookup = output index. */
MCT_API int mct_active_tagged_create_from_equation(mct_kernel_t *out, int Nk, const double *k_array, double prefactor, const double *J,
                                           const double *V2, mct_equation_t collective);
/* evaluate_kernel!(out, kernel, F, t): F and out are dof doubles.  t_index is used by tagged kernels only. */
MCT_API int mct_kernel_eval(mct_kernel_t kernel, const double *F, long t_index, double *out);
/* n independent evaluations K_b = kernel(F_b), F and out are [n][dof] (parity tests, roofline measurement). */
MCT_API int mct_kernel_eval_many(mct_kernel_t kernel, int n, const double *F, double *out, float *device_ms);
/* A kernel handle must outlive every equation created on it (the equation uses the kernel's stream, scratch and buffer
 * pool): while such equations exist the call changes nothing and returns MCT_BAD_ARGUMENT. */
MCT_API int mct_kernel_destroy(mct_kernel_t kernel);

/* ---- TimeDoublingSolver: plug-in point #1, `solve(equation, solver)`  src/Solvers.jl:42-46 ---- */

/* keyword arguments of TimeDoublingSolver(; N, Δt, t_max, max_iterations, tolerance, verbose, ismutable,
 * init_with_memory)  src/TimeDoublingSolver.jl:48-50 */
typedef struct mct_td_options {
    int N;
    double dt;
    double t_max;
    long max_iterations;
    double tolerance;
    int verbose;
    int init_with_memory; /* must be 1 (the default); 0 -> MCT_UNSUPPORTED */
} mct_td_options;

/* MemoryEquationCoefficients after cleaning (src/MemoryEquation.jl:15-21, src/HelperFunctions.jl:73-84):
 * kind 0 = UniformScaling: pass lambda in *_scalar;  kind 1 = Diagonal: pass `.diag` (Nk doubles for Ns = 1,
 * Nk blocks of Ns x Ns otherwise).  delta: dof doubles or NULL for zero. */
typedef struct mct_coeffs {
    int alpha_kind; double alpha_scalar; const double *alpha;
    int beta_kind;  double beta_scalar;  const double *beta;
    int gamma_kind; double gamma_scalar; const double *gamma;
    const double *delta;
} mct_coeffs;

/* nt = 1 + 2N + 2N*ceil-count of doublings while dt < 2*t_max  (src/TimeDoublingSolver.jl:518,523-528) */
MCT_API int mct_td_output_len(const mct_td_options *opts, long *nt);

/* solve(MemoryEquation(alpha,beta,gamma,delta,F0,dF0,kernel), TimeDoublingSolver(opts))
 * (src/TimeDoublingSolver.jl:512-532).  t_out[nt], F_out[nt][dof], K_out[nt][dof] are caller-allocated;
 * *kernel_evals receives solver.kernel_evals (:416,:520). */
MCT_API int mct_td_solve(mct_kernel_t kernel, const mct_coeffs *coeffs, const double *F0, const double *dF0,
                 const mct_td_options *opts, double *t_out, double *F_out, double *K_out, long *kernel_evals);

/* Device-resident form of the same call: create uploads coefficients/initial conditions and allocates the
 * history rings + trajectory; solve runs the persistent kernel only (no host<->device traffic); download copies
 * (t,F,K) back.  mct_td_solve == create + solve + download + destroy. */
MCT_API int mct_equation_create(mct_equation_t *out, mct_kernel_t kernel, const mct_coeffs *coeffs, const double *F0,
                        const double *dF0, const mct_td_options *opts);
MCT_API int mct_equation_solve(mct_equation_t eq, long *kernel_evals, float *device_ms);
MCT_API int mct_equation_download(mct_equation_t eq, double *t_out, double *F_out, double *K_out);
MCT_API int mct_equation_destroy(mct_equation_t eq);
/* Batch of independent equations with identical (Nk, Ns, options) -- e.g. a packing-fraction sweep.  All
 * equations are solved by ONE persistent launch per device; status[i]/kernel_evals[i] are per equation. */
MCT_API int mct_equation_solve_batch(int n, mct_equation_t *eqs, int *status, long *kernel_evals, float *device_ms);

/* ---- one very large multi-component equation sharded over the GPUs of a node (one process per GPU) ----
 * Every rank creates the SAME equation on its own device, then:
 *   prepare(q, rank, world, handle)   allocates this rank's exchange buffer and returns its 64-byte CUDA IPC handle;
 *   -- the host language all-gathers the world handles (any transport) --
 *   connect(q, handles[world][64])    maps the peers' buffers;
 *   mct_equation_solve(q)             on all ranks at the same time: the persistent kernel of rank r evaluates the lines
 *                                     L = r (mod world) of the vertex sum and stores their sums straight into every
 *                                     rank's buffer over NVLink (self-validating words, no host round trip, no
 *                                     collective call); everything else is replicated, so all ranks hold the solution. */
MCT_API int mct_equation_shard_prepare(mct_equation_t eq, int rank, int world, unsigned char *handle_out /* 64 bytes */);
MCT_API int mct_equation_shard_connect(mct_equation_t eq, const unsigned char *handles /* [world][64] */);

/* TaggedMultiComponentModeCouplingKernel(s, rho, kBT, m, k_array, S, sol) (src/Kernels/MultiComponentKernels.jl:268-333)
 * chained on the device-resident solution of a multi-component equation.  fill_A! (:335-370) is
 * A_m[iq,ip] = V_m[iq,ip] Fs[ip] G(iq,t) with the species weight G(q,t) = sum_ab C_q[s,a] C_q[s,b] F_ab(q,t): G is
 * contracted once on the device from the mixture trajectory and the kernel then is a tagged single-component kernel
 * (same persistent solver).  s is 0-based; rho_all = sum(rho), m_s = m[s]. */
MCT_API int mct_mc3d_tagged_create_from_equation(mct_kernel_t *out, int s, double rho_all, double kBT, double m_s,
                                         mct_equation_t collective);
/* MSDMultiComponentModeCouplingKernel(s, ...) (:400-486) + the scalar solve, like mct_msd3d_solve.  `collective` is the
 * solved mixture equation, `tagged` the solved tagged-particle equation of species s (0-based).
 * full_species_sum = 0 reproduces the reference: `Ns = length(Fs[1])` (:477) is 1 there, so its species sum only visits
 * alpha = beta = 1 and the golden test/test_MCMCT.jl:226 is that number; 1 evaluates the documented sum over all pairs. */
MCT_API int mct_msd_mc3d_solve(mct_equation_t collective, mct_equation_t tagged, int s, int full_species_sum, double rho_all,
                       double kBT, double m_s, double alpha, double beta, double gamma, double delta, double MSD0, double dMSD0,
                       double *t_out, double *msd_out, double *K_out, long *kernel_evals);

/* get_F(sol, its, iks) / get_K (src/HelperFunctions.jl:169-210) on the device-resident solution: gathers
 * F[times[i]][dofs[j]] into F_out[n_times][n_dofs] on the device and copies only that.  times == NULL: all nt output
 * times (n_times ignored); dofs == NULL: all dof entries.  t_out (n_times), F_out, K_out may each be NULL. */
MCT_API int mct_equation_download_slice(mct_equation_t eq, long n_times, const long *times, int n_dofs, const int *dofs,
                                double *t_out, double *F_out, double *K_out);
/* find_relaxation_time(t, F(k); threshold, mode) (src/RelaxationTime.jl:14-39) for n_k wavenumbers of a solved
 * single-valued equation: the columns are gathered on the device (8 nt n_k bytes cross PCIe instead of the whole
 * trajectory); the O(nt) search and interpolation run on the host as in the reference.  mode 0 = :log, 1 = :lin. */
MCT_API int mct_equation_relaxation_times(mct_equation_t eq, int n_k, const int *ik, double threshold, int mode, double *tau_out);

/* MSDModeCouplingKernel(rho, kBT, m, k_array, S, sol, taggedsol) (src/Kernels/ModeCouplingKernels.jl:552-570) and
 * solve(MemoryEquation(alpha, beta, gamma, delta, MSD0, dMSD0, kernel), TimeDoublingSolver) for it
 * (src/TimeDoublingSolver.jl:512-532, scalar/immutable branch :213-235).  `collective` and `tagged` are the two solved
 * device-resident equations (same grid, same solver options -- the reference's tDict lookup); the kernel table
 * K(t_i) (:572-586) is reduced on the device from their trajectories, nothing but nt-sized vectors crosses PCIe.
 * Sk: Nk doubles (host).  t_out, msd_out, K_out: nt doubles. */
MCT_API int mct_msd3d_solve(mct_equation_t collective, mct_equation_t tagged, double rho, double kBT, double m, const double *Sk,
                    double alpha, double beta, double gamma, double delta, double MSD0, double dMSD0,
                    double *t_out, double *msd_out, double *K_out, long *kernel_evals);

/* solve_steady_state(gamma, F0, kernel; tolerance, max_iterations)  src/SteadyStateMemoryEquation.jl:59-100.
 * gamma: dof doubles (Diagonal / blocks).  F_out, K_out: dof doubles. */
MCT_API int mct_steady_state(mct_kernel_t kernel, const double *gamma, const double *F0, double tolerance, long max_iterations,
                     double *F_out, double *K_out, long *iterations);

#ifdef __cplusplus
}
#endif
#endif /* MCT_H */
