# MCTDevice.jl -- glue a maintainer of ModeCouplingTheory.jl adds next to src/TimeDoublingSolver.jl:
#     include("MCTDevice.jl")            # in src/ModeCouplingTheory.jl after the solver includes (:22-24)
#     export DeviceTimeDoublingSolver
# It adds ONE solver subtype and ONE `solve` method (the documented extension recipe, docs/src/internals.md:7-16) and
# forwards them through `ccall` to libmct_sm100a.so (include/mct.h).  Everything else of the package is untouched:
# MemoryEquation, the ModeCouplingKernel constructors, get_t/get_F/get_K keep working on the returned solution.
# NOTE: julia is not available in the build image, so this file is not executed by the test-suite; the Python mirror
# (modecouplingtheory.jl_b200/api.py) drives exactly the same C entry points with the same argument marshalling.

const libmct = get(ENV, "LIBMCT_SM100A", "libmct_sm100a.so")

struct MctTdOptions          # struct mct_td_options
    N::Cint
    dt::Cdouble
    t_max::Cdouble
    max_iterations::Clong
    tolerance::Cdouble
    verbose::Cint
    init_with_memory::Cint
end

struct MctCoeffs             # struct mct_coeffs
    alpha_kind::Cint; alpha_scalar::Cdouble; alpha::Ptr{Cdouble}
    beta_kind::Cint;  beta_scalar::Cdouble;  beta::Ptr{Cdouble}
    gamma_kind::Cint; gamma_scalar::Cdouble; gamma::Ptr{Cdouble}
    delta::Ptr{Cdouble}
end

"Same keyword arguments as `TimeDoublingSolver` (src/TimeDoublingSolver.jl:48-50), plus the CUDA device index."
mutable struct DeviceTimeDoublingSolver <: Solver
    N::Int; Δt::Float64; t_max::Float64; max_iterations::Int; tolerance::Float64
    verbose::Bool; device::Int; kernel_evals::Int
end
DeviceTimeDoublingSolver(; N=32, Δt=1e-10, t_max=1e10, max_iterations=10^4, tolerance=1e-10, verbose=false, device=0) =
    DeviceTimeDoublingSolver(N, Δt, t_max, max_iterations, tolerance, verbose, device, 0)

function mct_check(rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:mct_last_error, libmct), Cstring, ()))
    rc == 1 && throw(DomainError("Iteration did not converge.", msg))      # TimeDoublingSolver.jl:406-408
    rc == 2 && throw(AssertionError(msg))                                   # ModeCouplingKernels.jl:103-104
    rc == 3 && throw(ArgumentError("not supported on the device: " * msg))
    error("libmct_sm100a: " * msg)                                          # no CPU fallback
end

# UniformScaling -> (0, λ, NULL); Diagonal{Float64} -> (1, 0, pointer(diag))   (MemoryEquation.jl:15-21)
coef(c::UniformScaling) = (Cint(0), Float64(c.λ), Ptr{Cdouble}(C_NULL))
coef(c::Diagonal{Float64}) = (Cint(1), 0.0, pointer(c.diag))

device_kernel(k::ModeCouplingKernel3D, dev) = begin
    h = Ref{Ptr{Cvoid}}(C_NULL)
    mct_check(ccall((:mct_sc3d_create, libmct), Cint,
        (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
        h, dev, length(k.k_array), k.k_array, k.V1, k.V2, k.V3))            # tables made by the untouched constructor (:94-129)
    h[]
end
device_kernel(k::TaggedModeCouplingKernel3D, dev) = begin
    h = Ref{Ptr{Cvoid}}(C_NULL)
    Fc = reduce(hcat, k.F)                                                  # [Nk, nt] column-major == [nt][Nk] row-major
    mct_check(ccall((:mct_sc3d_tagged_create, libmct), Cint,
        (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Clong, Ptr{Cdouble}),
        h, dev, length(k.k_array), k.k_array, k.V1, k.V2, k.V3, size(Fc, 2), Fc))
    h[]
end

function solve(eq::MemoryEquation, s::DeviceTimeDoublingSolver)
    eq.update_coefficients! === do_nothing || throw(ArgumentError("update_coefficients! callbacks are not supported on the device"))
    F₀ = eq.F₀ isa Vector{Float64} ? eq.F₀ : throw(ArgumentError("the device solver is Float64-only"))
    h = device_kernel(eq.kernel, s.device)
    opts = MctTdOptions(s.N, s.Δt, s.t_max, s.max_iterations, s.tolerance, s.verbose, 1)
    nt = Ref{Clong}(0)
    mct_check(ccall((:mct_td_output_len, libmct), Cint, (Ref{MctTdOptions}, Ref{Clong}), opts, nt))
    Nk = length(F₀)
    t = Vector{Float64}(undef, nt[]); F = Matrix{Float64}(undef, Nk, nt[]); K = Matrix{Float64}(undef, Nk, nt[])
    c = eq.coeffs
    δ = c.δ isa AbstractVector ? pointer(c.δ) : Ptr{Cdouble}(C_NULL)
    coeffs = MctCoeffs(coef(c.α)..., coef(c.β)..., coef(c.γ)..., δ)
    evals = Ref{Clong}(0)
    GC.@preserve c F₀ begin
        rc = ccall((:mct_td_solve, libmct), Cint,
            (Ptr{Cvoid}, Ref{MctCoeffs}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{MctTdOptions}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Clong}),
            h, coeffs, F₀, eq.∂ₜF₀, opts, t, F, K, evals)
    end
    ccall((:mct_kernel_destroy, libmct), Cint, (Ptr{Cvoid},), h)
    s.kernel_evals = evals[]                                                # TimeDoublingSolver.jl:416,520
    mct_check(rc)
    # same container types as the reference: Vector of per-time vectors / Diagonals (Solvers.jl:19-24)
    MemoryEquationSolution(t, [F[:, i] for i in 1:nt[]], [Diagonal(K[:, i]) for i in 1:nt[]], s)
end

# ---- the steps after the collective solve, chained on device-resident solutions (include/mct.h) -------------------
# A maintainer would keep the equation handle `q` (mct_equation_create / _solve / _download instead of mct_td_solve) in a
# `DeviceSolution <: Any` next to the MemoryEquationSolution and add these methods; nothing but handles and nt-sized
# vectors crosses PCIe.  Species / wavenumber / time indices are 0-based on the C side.

# get_F(sol, its, iks)  (src/HelperFunctions.jl:169-210)
function device_slice(q::Ptr{Cvoid}, its::Vector{Int}, iks::Vector{Int})
    t = Vector{Float64}(undef, length(its)); F = Matrix{Float64}(undef, length(iks), length(its)); K = similar(F)
    mct_check(ccall((:mct_equation_download_slice, libmct), Cint,
        (Ptr{Cvoid}, Clong, Ptr{Clong}, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
        q, length(its), Clong.(its .- 1), length(iks), Cint.(iks .- 1), t, F, K))
    t, F, K
end

# find_relaxation_time(get_t(sol), get_F(sol, :, ik); threshold, mode)  (src/RelaxationTime.jl:14-39)
function device_relaxation_times(q::Ptr{Cvoid}, iks::Vector{Int}; threshold=exp(-1), mode=:log)
    τ = Vector{Float64}(undef, length(iks))
    mct_check(ccall((:mct_equation_relaxation_times, libmct), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Cdouble, Cint, Ptr{Cdouble}),
        q, length(iks), Cint.(iks .- 1), threshold, mode == :log ? 0 : 1, τ))
    τ
end

# TaggedMultiComponentModeCouplingKernel(s, ρ, kBT, m, k_array, Sₖ, sol)  (src/Kernels/MultiComponentKernels.jl:268-333)
function device_tagged_mc_kernel(s::Int, ρ, kBT, m, q_mixture::Ptr{Cvoid})
    h = Ref{Ptr{Cvoid}}(C_NULL)
    mct_check(ccall((:mct_mc3d_tagged_create_from_equation, libmct), Cint, (Ref{Ptr{Cvoid}}, Cint, Cdouble, Cdouble, Cdouble, Ptr{Cvoid}),
        h, s - 1, sum(ρ), kBT, m[s], q_mixture))
    h[]
end

# solve(MemoryEquation(α, β, γ, δ, msd0, dmsd0, MSDModeCouplingKernel(ρ, kBT, m, k_array, Sₖ, sol, taggedsol)), solver)
# (src/Kernels/ModeCouplingKernels.jl:552-586 + the scalar branch of src/TimeDoublingSolver.jl)
function device_msd(q::Ptr{Cvoid}, q_tagged::Ptr{Cvoid}, ρ, kBT, m, Sₖ::Vector{Float64}, α, β, γ, δ, msd0, dmsd0, nt::Int)
    t = Vector{Float64}(undef, nt); msd = similar(t); K = similar(t); evals = Ref{Clong}(0)
    mct_check(ccall((:mct_msd3d_solve, libmct), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble,
         Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Clong}),
        q, q_tagged, ρ, kBT, m, Sₖ, α, β, γ, δ, msd0, dmsd0, t, msd, K, evals))
    t, msd, K, evals[]
end

# ActiveMCTKernel(ρ, k_array, wk, w0, Sk, dim)  (src/Kernels/ActiveMCTKernel.jl:28-49): J and V2 are [l, j, i], the device layout
function device_kernel(k::ActiveMCTKernel, dev::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    mct_check(ccall((:mct_active_create, libmct), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}),
        h, dev, length(k.k_array), k.k_array, k.prefactor, k.J, k.V2))
    h[]
end
