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
