# MCTDevice.jl -- glue a maintainer of ModeCouplingTheory.jl (v0.8.9) adds next to src/TimeDoublingSolver.jl:
#     include("MCTDevice.jl")            # in src/ModeCouplingTheory.jl after the solver includes (:22-24)
#     export DeviceTimeDoublingSolver, DeviceKernel, DeviceSolution, device_kernel
# It adds ONE solver subtype with ONE `solve` method (plug-in point #1, the documented extension recipe
# docs/src/internals.md:7-16) and ONE kernel subtype with `evaluate_kernel!` (plug-in point #2, src/Kernels.jl:10-26), and
# forwards them through `ccall` to libmct_sm100a.so (include/mct.h).  Everything else of the package is untouched:
# MemoryEquation, the ModeCouplingKernel / MultiComponentModeCouplingKernel / dDimModeCouplingKernel constructors,
# get_t/get_F/get_K keep working on the returned MemoryEquationSolution.
#
# julia is not available in the build image: this file is not executed by the test-suite.  What IS executed:
# tests/abi_harness.c drives the same entry points with the same struct layouts from plain C (gcc against include/mct.h),
# and the Python mirror (modecouplingtheory.jl_b200/api.py) marshals the same arguments in the same order.

using LinearAlgebra, StaticArrays

const libmct = get(ENV, "LIBMCT_SM100A", "libmct_sm100a.so")

struct MctTdOptions          # struct mct_td_options (include/mct.h:107-115); isbits, C layout
    N::Cint
    dt::Cdouble
    t_max::Cdouble
    max_iterations::Clong
    tolerance::Cdouble
    verbose::Cint
    init_with_memory::Cint
end

struct MctCoeffs             # struct mct_coeffs (include/mct.h:120-125)
    alpha_kind::Cint; alpha_scalar::Cdouble; alpha::Ptr{Cdouble}
    beta_kind::Cint;  beta_scalar::Cdouble;  beta::Ptr{Cdouble}
    gamma_kind::Cint; gamma_scalar::Cdouble; gamma::Ptr{Cdouble}
    delta::Ptr{Cdouble}
end

"""
    DeviceTimeDoublingSolver(; N, Δt, t_max, max_iterations, tolerance, verbose, ismutable, init_with_memory, device)

Same keyword arguments as `TimeDoublingSolver` (src/TimeDoublingSolver.jl:48-50) plus the CUDA device index.  The device path
is the mutable, memory-initialised one: `ismutable=false` or `init_with_memory=false` throw an `ArgumentError`.
"""
mutable struct DeviceTimeDoublingSolver <: Solver
    N::Int; Δt::Float64; t_max::Float64; kernel_evals::Int; max_iterations::Int; tolerance::Float64
    verbose::Bool; ismutable::Bool; init_with_memory::Bool; device::Int
end
function DeviceTimeDoublingSolver(; N=32, Δt=10^-10, t_max=10.0^10, max_iterations=10^4, tolerance=10^-10, verbose=false,
                                  ismutable=true, init_with_memory=true, device=0)
    ismutable || throw(ArgumentError("DeviceTimeDoublingSolver: ismutable=false (the allocating branch) is not built on the device"))
    init_with_memory || throw(ArgumentError("DeviceTimeDoublingSolver: init_with_memory=false is not built on the device"))
    DeviceTimeDoublingSolver(N, Δt, t_max, 0, max_iterations, tolerance, verbose, ismutable, init_with_memory, device)
end

function mct_check(rc::Integer)
    rc == 0 && return
    msg = unsafe_string(ccall((:mct_last_error, libmct), Cstring, ()))
    rc == 1 && throw(DomainError("Iteration did not converge.", msg))      # TimeDoublingSolver.jl:406-408
    rc == 2 && throw(AssertionError(msg))                                   # ModeCouplingKernels.jl:103-104, tDict KeyError
    rc == 3 && throw(ArgumentError("not supported on the device: " * msg))
    error("libmct_sm100a: " * msg)                                          # no CPU fallback
end

# ---------------------------------------------------------------------------------------------- kernels (plug-in point #2)
"""
    DeviceKernel(kernel; device=0)  /  device_kernel(kernel, device)

A `MemoryKernel` that holds the opaque handle of the device copy of `kernel` (built from the tables the untouched Julia
constructor produced) and forwards `evaluate_kernel!` to `mct_kernel_eval`.  `MemoryEquation(α, β, γ, δ, F₀, ∂ₜF₀, DeviceKernel(k))`
therefore evaluates K₀ on the device (src/MemoryEquation.jl:45), and `solve_steady_state` works through the generic code.
"""
mutable struct DeviceKernel{K<:MemoryKernel} <: MemoryKernel
    host::K                  # keeps the Julia kernel (and with it every array the handle was built from) alive
    handle::Ptr{Cvoid}
    device::Int
    function DeviceKernel(host::K, handle::Ptr{Cvoid}, device::Int) where {K<:MemoryKernel}
        k = new{K}(host, handle, device)
        finalizer(k) do x                                  # returns MCT_BAD_ARGUMENT (and changes nothing) while equations live on it
            x.handle == C_NULL || ccall((:mct_kernel_destroy, libmct), Cint, (Ptr{Cvoid},), x.handle)
        end
        k
    end
end
DeviceKernel(k::MemoryKernel; device=0) = DeviceKernel(k, device_kernel(k, device), device)

blocks_ptr(v::Vector{<:SMatrix{Ns,Ns,Float64}}) where {Ns} = Ptr{Cdouble}(pointer(v))   # isbits elements: contiguous [Ns,Ns,Nk]

# ModeCouplingKernel(ρ, kBT, m, k_array, Sₖ; dims=3)   src/Kernels/ModeCouplingKernels.jl:94-129
function device_kernel(k::ModeCouplingKernel3D, dev::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    mct_check(ccall((:mct_sc3d_create, libmct), Cint,
        (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
        h, dev, length(k.k_array), k.k_array, k.V1, k.V2, k.V3))
    h[]
end
# TaggedModeCouplingKernel(ρ, kBT, m, k_array, Sₖ, sol; dims=3)   :390-426; the tDict lookup (:394, :436) becomes an index check
function device_kernel(k::TaggedModeCouplingKernel3D, dev::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    Fc = reduce(hcat, k.F)                                                  # [Nk, nt] column-major == [nt][Nk] row-major
    mct_check(ccall((:mct_sc3d_tagged_create, libmct), Cint,
        (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Clong, Ptr{Cdouble}),
        h, dev, length(k.k_array), k.k_array, k.V1, k.V2, k.V3, size(Fc, 2), Fc))
    ts = sort!(collect(keys(k.tDict)))                                      # get_t(sol) of the collective solve
    mct_check(ccall((:mct_kernel_set_collective_times, libmct), Cint, (Ptr{Cvoid}, Clong, Ptr{Cdouble}), h[], length(ts), ts))
    h[]
end
# MultiComponentModeCouplingKernel(ρ, kBT, m, k_array, Sₖ; dims=3)   src/Kernels/MultiComponentKernels.jl:84-131
function device_kernel(k::MultiComponentModeCouplingKernel3D, dev::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    pref = Matrix{Float64}(k.prefactor)                                     # Ns x Ns, column-major
    GC.@preserve k pref mct_check(ccall((:mct_mc3d_create, libmct), Cint,
        (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
        h, dev, k.Nk, k.Ns, k.k_array, blocks_ptr(k.C), pref))
    h[]
end
# dDimModeCouplingKernel(ρ, kBT, m, k_array, Sₖ, d)   src/Kernels/ModeCouplingKernels.jl:34-71; V, J are [iq, ik, ip]
function device_kernel(k::dDimModeCouplingKernel, dev::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    mct_check(ccall((:mct_ddim_create, libmct), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}),
        h, dev, k.Nk, k.k_array, k.prefactor, k.V, k.J))
    h[]
end
# ActiveMCTKernel(ρ, k_array, wk, w0, Sk, dim)   src/Kernels/ActiveMCTKernel.jl:28-49; J and V2 are [l, j, i]
function device_kernel(k::ActiveMCTKernel, dev::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    mct_check(ccall((:mct_active_create, libmct), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}),
        h, dev, length(k.k_array), k.k_array, k.prefactor, k.J, k.V2))
    h[]
end
device_kernel(k::DeviceKernel, dev::Integer) = k.handle
device_kernel(k::MemoryKernel, dev::Integer) =
    throw(ArgumentError("no device implementation of $(typeof(k)); supported: ModeCouplingKernel3D, TaggedModeCouplingKernel3D, " *
                        "MultiComponentModeCouplingKernel3D, dDimModeCouplingKernel, ActiveMCTKernel (and TaggedActiveMCTKernel / the " *
                        "MSD kernels chained on a DeviceSolution)"))

dof_vector(F::Vector{Float64}) = F
dof_vector(F::Vector{<:SMatrix{Ns,Ns,Float64}}) where {Ns} = reinterpret(Float64, F)
time_index(k::DeviceKernel, t) = hasproperty(k.host, :tDict) ? Clong(k.host.tDict[t] - 1) : Clong(0)    # KeyError as in the reference

# evaluate_kernel!(out, kernel, F, t)   src/Kernels.jl:10-26 -- Diagonal{Float64} / Diagonal{SMatrix} outputs
function evaluate_kernel!(out::Diagonal, kernel::DeviceKernel, F::Vector, t)
    GC.@preserve out F mct_check(ccall((:mct_kernel_eval, libmct), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Clong, Ptr{Cdouble}),
        kernel.handle, pointer(dof_vector(F)), time_index(kernel, t), pointer(dof_vector(out.diag))))
    out
end
function evaluate_kernel(kernel::DeviceKernel, F::Vector, t)
    out = Diagonal(similar(F))
    evaluate_kernel!(out, kernel, F, t)
end

# ---------------------------------------------------------------------------------------------- solver (plug-in point #1)
# UniformScaling -> (0, λ, NULL); Diagonal -> (1, 0, pointer to .diag: Nk doubles or Nk blocks)   (MemoryEquation.jl:15-21)
coef(c::UniformScaling) = (Cint(0), Float64(c.λ), Ptr{Cdouble}(C_NULL))
coef(c::Diagonal{Float64}) = (Cint(1), 0.0, pointer(c.diag))
coef(c::Diagonal{<:SMatrix{Ns,Ns,Float64}}) where {Ns} = (Cint(1), 0.0, blocks_ptr(c.diag))
coef(c) = throw(ArgumentError("the device solver needs scalar or diagonal coefficients, got $(typeof(c)) (the dense LinearSolve branch, " *
                              "src/TimeDoublingSolver.jl:338-344, is not built on the device)"))

"The default `update_coefficients!` of MemoryEquation is the anonymous closure `(coeffs, t) -> nothing` (src/MemoryEquation.jl:44): it
cannot be recognised by identity, so the callback is PROBED -- called on a copy of the coefficients at two times; any change of the
copy means a real callback, which the device solver (time loop resident on the GPU) cannot honour."
function assert_no_coefficient_callback(eq::MemoryEquation)
    c = eq.coeffs
    probe = MemoryEquationCoefficients(deepcopy(c.α), deepcopy(c.β), deepcopy(c.γ), deepcopy(c.δ))
    for t in (0.0, 1.0)
        eq.update_coefficients!(probe, t)
    end
    (probe.α == c.α && probe.β == c.β && probe.γ == c.γ && probe.δ == c.δ) ||
        throw(ArgumentError("update_coefficients! callbacks are not supported on the device"))
end

"""
    DeviceSolution

What `solve(eq, ::DeviceTimeDoublingSolver; keep_on_device=true)` returns besides the ordinary `MemoryEquationSolution`: the handle
of the equation whose trajectory stays resident on the GPU, for the steps that follow the collective solve in every workflow of the
reference (tagged kernel, MSD, slices, relaxation times) without a PCIe round trip of the trajectory.
"""
mutable struct DeviceSolution
    handle::Ptr{Cvoid}
    kernel::DeviceKernel         # the kernel handle must outlive the equation
    nt::Int
    function DeviceSolution(handle, kernel, nt)
        s = new(handle, kernel, nt)
        finalizer(x -> x.handle == C_NULL || ccall((:mct_equation_destroy, libmct), Cint, (Ptr{Cvoid},), x.handle), s)
        s
    end
end

wrap_F(F::Matrix{Float64}, ::Vector{Float64}) = [F[:, i] for i in axes(F, 2)]
wrap_K(K::Matrix{Float64}, ::Vector{Float64}) = [Diagonal(K[:, i]) for i in axes(K, 2)]
wrap_F(F::Matrix{Float64}, ::Vector{SMatrix{Ns,Ns,Float64,L}}) where {Ns,L} =
    [collect(reinterpret(SMatrix{Ns,Ns,Float64,L}, F[:, i])) for i in axes(F, 2)]
wrap_K(K::Matrix{Float64}, F₀::Vector{SMatrix{Ns,Ns,Float64,L}}) where {Ns,L} = [Diagonal(v) for v in wrap_F(K, F₀)]

function solve(eq::MemoryEquation, s::DeviceTimeDoublingSolver; keep_on_device::Bool=false)
    F₀ = eq.F₀
    (F₀ isa Vector{Float64} || F₀ isa Vector{<:SMatrix{Ns,Ns,Float64} where Ns}) ||
        throw(ArgumentError("the device solver is Float64-only (Vector{Float64} or Vector{SMatrix{Ns,Ns,Float64}}), got $(typeof(F₀))"))
    assert_no_coefficient_callback(eq)
    dk = eq.kernel isa DeviceKernel ? eq.kernel : DeviceKernel(eq.kernel; device=s.device)
    opts = MctTdOptions(s.N, s.Δt, s.t_max, s.max_iterations, s.tolerance, s.verbose, 1)
    nt = Ref{Clong}(0)
    mct_check(ccall((:mct_td_output_len, libmct), Cint, (Ref{MctTdOptions}, Ref{Clong}), opts, nt))
    dof = length(dof_vector(F₀))
    t = Vector{Float64}(undef, nt[]); F = Matrix{Float64}(undef, dof, nt[]); K = Matrix{Float64}(undef, dof, nt[])
    c = eq.coeffs
    δ = c.δ isa AbstractVector ? pointer(dof_vector(c.δ)) : Ptr{Cdouble}(C_NULL)
    coeffs = MctCoeffs(coef(c.α)..., coef(c.β)..., coef(c.γ)..., δ)
    evals = Ref{Clong}(0)
    local dsol = nothing
    GC.@preserve c eq dk begin
        if keep_on_device                 # create + solve + download: the equation (trajectory) stays on the GPU
            q = Ref{Ptr{Cvoid}}(C_NULL); ms = Ref{Cfloat}(0)
            mct_check(ccall((:mct_equation_create, libmct), Cint,
                (Ref{Ptr{Cvoid}}, Ptr{Cvoid}, Ref{MctCoeffs}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{MctTdOptions}),
                q, dk.handle, coeffs, pointer(dof_vector(F₀)), pointer(dof_vector(eq.∂ₜF₀)), opts))
            dsol = DeviceSolution(q[], dk, nt[])
            rc = ccall((:mct_equation_solve, libmct), Cint, (Ptr{Cvoid}, Ref{Clong}, Ref{Cfloat}), q[], evals, ms)
            s.kernel_evals = evals[]                                        # TimeDoublingSolver.jl:416,520
            mct_check(rc)
            mct_check(ccall((:mct_equation_download, libmct), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), q[], t, F, K))
        else
            rc = ccall((:mct_td_solve, libmct), Cint,
                (Ptr{Cvoid}, Ref{MctCoeffs}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{MctTdOptions}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Clong}),
                dk.handle, coeffs, pointer(dof_vector(F₀)), pointer(dof_vector(eq.∂ₜF₀)), opts, t, F, K, evals)
            s.kernel_evals = evals[]
            mct_check(rc)
        end
    end
    # same container types as the reference: Vector of per-time vectors / Diagonals (Solvers.jl:19-24), so get_F(sol, :, ik),
    # sol[ik] etc. (HelperFunctions.jl:169-253) work unchanged; solver.Δt is left as the caller set it (:519-520, :529)
    sol = MemoryEquationSolution(t, wrap_F(F, F₀), wrap_K(K, F₀), s)
    keep_on_device ? (sol, dsol) : sol
end

# ---- the steps after the collective solve, chained on a DeviceSolution (include/mct.h): only handles and nt-sized vectors
# ---- cross PCIe.  Species / wavenumber / time indices are 0-based on the C side.

# get_F(sol, its, iks)  (src/HelperFunctions.jl:169-210)
function device_slice(d::DeviceSolution, its::Vector{Int}, iks::Vector{Int})
    t = Vector{Float64}(undef, length(its)); F = Matrix{Float64}(undef, length(iks), length(its)); K = similar(F)
    mct_check(ccall((:mct_equation_download_slice, libmct), Cint,
        (Ptr{Cvoid}, Clong, Ptr{Clong}, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
        d.handle, length(its), Clong.(its .- 1), length(iks), Cint.(iks .- 1), t, F, K))
    t, F, K
end

# find_relaxation_time(get_t(sol), get_F(sol, :, ik); threshold, mode)  (src/RelaxationTime.jl:14-39)
function device_relaxation_times(d::DeviceSolution, iks::Vector{Int}; threshold=exp(-1), mode=:log)
    τ = Vector{Float64}(undef, length(iks))
    mct_check(ccall((:mct_equation_relaxation_times, libmct), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Cdouble, Cint, Ptr{Cdouble}),
        d.handle, length(iks), Cint.(iks .- 1), threshold, mode == :log ? 0 : 1, τ))
    τ
end

# TaggedModeCouplingKernel(ρ, kBT, m, k_array, Sₖ, sol) on the resident collective solution (no upload of sol.F)
function device_tagged_kernel(ρ, kBT, m, k_array::Vector{Float64}, Sₖ::Vector{Float64}, d::DeviceSolution)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    mct_check(ccall((:mct_sc3d_tagged_create_from_equation, libmct), Cint,
        (Ref{Ptr{Cvoid}}, Cint, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}),
        h, length(k_array), ρ, kBT, m, k_array, Sₖ, d.handle))
    h[]
end

# TaggedActiveMCTKernel(ρ, k_array, wk, w0, Sk, Fsol, dim)  (src/Kernels/ActiveMCTKernel.jl:96-136) on the resident collective solution
function device_kernel(k::TaggedActiveMCTKernel, d::DeviceSolution)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    mct_check(ccall((:mct_active_tagged_create_from_equation, libmct), Cint,
        (Ref{Ptr{Cvoid}}, Cint, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}),
        h, length(k.k_array), k.k_array, k.prefactor, k.J, k.V2, d.handle))
    h[]
end

# TaggedMultiComponentModeCouplingKernel(s, ρ, kBT, m, k_array, Sₖ, sol)  (src/Kernels/MultiComponentKernels.jl:268-333)
function device_tagged_mc_kernel(s::Int, ρ, kBT, m, mixture::DeviceSolution)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    mct_check(ccall((:mct_mc3d_tagged_create_from_equation, libmct), Cint, (Ref{Ptr{Cvoid}}, Cint, Cdouble, Cdouble, Cdouble, Ptr{Cvoid}),
        h, s - 1, sum(ρ), kBT, m[s], mixture.handle))
    h[]
end

# solve(MemoryEquation(α, β, γ, δ, msd0, dmsd0, MSDModeCouplingKernel(ρ, kBT, m, k_array, Sₖ, sol, taggedsol)), solver)
# (src/Kernels/ModeCouplingKernels.jl:552-586 + the scalar branch of src/TimeDoublingSolver.jl)
function device_msd(col::DeviceSolution, tagged::DeviceSolution, ρ, kBT, m, Sₖ::Vector{Float64}, α, β, γ, δ, msd0, dmsd0)
    nt = col.nt
    t = Vector{Float64}(undef, nt); msd = similar(t); K = similar(t); evals = Ref{Clong}(0)
    mct_check(ccall((:mct_msd3d_solve, libmct), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble,
         Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Clong}),
        col.handle, tagged.handle, ρ, kBT, m, Sₖ, α, β, γ, δ, msd0, dmsd0, t, msd, K, evals))
    t, msd, K, evals[]
end

# solve_steady_state(γ, F₀, kernel; tolerance, max_iterations)  (src/SteadyStateMemoryEquation.jl:59-100) in one launch
function device_steady_state(γ, F₀::Vector, kernel::DeviceKernel; tolerance=10^-8, max_iterations=10^6)
    g = γ isa Diagonal ? dof_vector(γ.diag) : fill(Float64(γ), length(dof_vector(F₀)))
    F = similar(F₀); K = similar(F₀); it = Ref{Clong}(0)
    GC.@preserve g F₀ F K begin
        rc = ccall((:mct_steady_state, libmct), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Clong, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Clong}),
            kernel.handle, pointer(g), pointer(dof_vector(F₀)), tolerance, max_iterations, pointer(dof_vector(F)), pointer(dof_vector(K)), it)
    end
    rc == 1 && error("The recursive iteration did not converge. The error is larger than the tolerance after $(max_iterations) iterations.")  # :76-78
    mct_check(rc)
    F, Diagonal(K), it[]
end
