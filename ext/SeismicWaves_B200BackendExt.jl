# SeismicWaves_B200BackendExt.jl -- Julia package extension that plugs libswb200.so (this repository) into SeismicWaves.jl as
# the backend `parall = :B200`.  It is the B200 counterpart of ext/SeismicWaves_CUDABackendExt.jl (reference) and is loaded
# when the tiny trigger package `SeismicWavesB200` (exports `libswb200`, the path of the shared library) is imported:
#
#   [weakdeps]   SeismicWavesB200 = "<uuid>"          [extensions]   SeismicWaves_B200BackendExt = "SeismicWavesB200"
#
# Two levels of integration (INTEGRATION.md):
#  1. backend modules `Acoustic2D_CD_CPML_B200`, `Acoustic3D_CD_CPML_B200`, `Acoustic2D_VD_CPML_B200`, `Elastic2D_Iso_CPML_B200`
#     exposing exactly the members the reference's L3/L4 code calls on a backend module (SURVEY.md 8b): `Data.Array`, `zeros`,
#     `ones`, `forward_onestep_CPML!`, `adjoint_onestep_CPML!`, `prescale_residuals!`, `correlate_gradient*!` -- one ccall per
#     reference function, operating on `B200Array`s;
#  2. methods of `swforward_1shot!` / `swgradient_1shot!` specialised on simulations whose fields are `B200Array`s: they hand
#     the whole shot to the per-shot engine (`swb_sim_*`), which runs the time loops, checkpointing and correlation on the GPU.
#     `B200Array`s are allocated lazily (a never-written array is all zeros and owns no device memory), so the fields and the
#     LinearCheckpointer the reference's constructors create cost nothing on this path: the engine's copy is the only one.
# There is no CPU fallback: every call raises if the library reports an error (e.g. no sm_100 device).
#
# NOTE: Julia is not available in the build environment of this repository, so this file has never been executed.  What IS checked
# (tests/test_julia_boundary.py, CPU): every struct below against the C layout the library itself reports (swb_abi_layout), every
# ccall against the prototypes of include/swb200.h (symbol, return type, argument count and kinds), and the member lists of the four
# backend tuples against SURVEY.md 8b.  Its Python twin (seismicwaves.jl_b200/, same C ABI, same call sequence) is what runs on the
# GPU; keep the two in sync.
module SeismicWaves_B200BackendExt

using SeismicWaves, SeismicWavesB200
using SeismicWaves: CPMLBoundaryCondition, LocalGrid, AcousticCDCPMLWaveSimulation, AcousticVDStaggeredCPMLWaveSimulation,
                    ElasticIsoCPMLWaveSimulation, ScalarShot, MomentTensorShot, ExternalForceShot, MomentTensor2D, AbstractMisfit, AbstractField

const lib = SeismicWavesB200.libswb200

# ---------------------------------------------------------------------------------------------------------------------
# error handling
# ---------------------------------------------------------------------------------------------------------------------
function check(status::Int32)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:swb_last_error, lib), Cstring, ()))
    error("libswb200 error $status: $msg")
end

dtype_code(::Type{Float32}) = Int32(0)
dtype_code(::Type{Float64}) = Int32(1)

# ---------------------------------------------------------------------------------------------------------------------
# B200Array{T,N}: device array owned by libswb200 (plays the role of backend.Data.Array).  `ptr` stays C_NULL until the array
# is written or handed to a kernel: an untouched array reads as zeros (what backend.zeros promises) without owning memory.
# ---------------------------------------------------------------------------------------------------------------------
mutable struct B200Array{T, N} <: DenseArray{T, N}
    ptr::Ptr{Cvoid}
    dims::NTuple{N, Int}
    function B200Array{T, N}(dims::NTuple{N, Int}) where {T, N}
        a = new{T, N}(C_NULL, dims)
        finalizer(release!, a)
        return a
    end
end
function release!(a::B200Array)
    if a.ptr != C_NULL
        ccall((:swb_free, lib), Int32, (Ptr{Cvoid},), a.ptr)
        a.ptr = C_NULL
    end
    return nothing
end
ismaterialized(a::B200Array) = a.ptr != C_NULL
function materialize!(a::B200Array{T}) where {T}
    if a.ptr == C_NULL
        p = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:swb_malloc, lib), Int32, (Ref{Ptr{Cvoid}}, Csize_t), p, max(prod(a.dims), 1) * sizeof(T)))   # zero-filled
        a.ptr = p[]
    end
    return a
end
Base.size(a::B200Array) = a.dims
Base.eltype(::B200Array{T}) where {T} = T
Base.unsafe_convert(::Type{Ptr{Cvoid}}, a::B200Array) = materialize!(a).ptr
Base.getindex(::B200Array, ::Int...) = error("scalar indexing of a B200Array is not supported; copy it to the host with Array(a)")
Base.setindex!(::B200Array, v, ::Int...) = error("scalar indexing of a B200Array is not supported; build the array on the host and copy it")
function B200Array(h::Array{T, N}) where {T, N}                       # Data.Array(hostarray): H2D copy
    a = B200Array{T, N}(size(h))
    copyto!(a, h)
    return a
end
B200Array{T, N}(h::Array{T, N}) where {T, N} = B200Array(h)
function Base.copyto!(d::B200Array{T}, h::Array{T}) where {T}         # copyto!(dev, host)
    @assert length(d) == length(h)
    materialize!(d)
    check(ccall((:swb_memcpy_h2d, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), d.ptr, h, sizeof(h)))
    return d
end
function Base.copyto!(h::Array{T}, d::B200Array{T}) where {T}         # copyto!(host, dev)
    @assert length(d) == length(h)
    if !ismaterialized(d)
        return fill!(h, zero(T))
    end
    check(ccall((:swb_memcpy_d2h, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), h, d.ptr, sizeof(h)))
    return h
end
function Base.copyto!(d::B200Array{T}, s::B200Array{T}) where {T}     # copyto!(dev, dev)  (fields.jl:20,32-36)
    @assert length(d) == length(s)
    if !ismaterialized(s)
        return ismaterialized(d) ? fill!(d, zero(T)) : d              # zeros onto zeros: nothing to move
    end
    materialize!(d)
    check(ccall((:swb_memcpy_d2d, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), d.ptr, s.ptr, length(s) * sizeof(T)))
    return d
end
Base.Array(d::B200Array{T, N}) where {T, N} = copyto!(Array{T, N}(undef, size(d)), d)
Base.zero(d::B200Array{T, N}) where {T, N} = B200Array{T, N}(size(d))
Base.similar(d::B200Array{T, N}) where {T, N} = B200Array{T, N}(size(d))
Base.copy(d::B200Array{T, N}) where {T, N} = copyto!(B200Array{T, N}(size(d)), d)
Base.convert(::Type{Array}, d::B200Array) = Array(d)
Base.convert(::Type{Array{T, N}}, d::B200Array{T, N}) where {T, N} = Array(d)   # snapshotter.jl:21-25
function Base.fill!(d::B200Array{T}, v) where {T}                     # dev .= scalar (fields.jl:40-44)
    if !ismaterialized(d) && iszero(v)
        return d
    end
    materialize!(d)
    check(ccall((:swb_fill, lib), Int32, (Ptr{Cvoid}, Int32, Cdouble, Csize_t, Csize_t), d.ptr, dtype_code(T), Float64(v), 0, length(d)))
    return d
end
# a[1:n] .= v on a vector (free-surface override of the C-PML coefficients, acou_init_bc.jl:35-38)
function fillrange!(d::B200Array{T, 1}, r::UnitRange{Int}, v) where {T}
    materialize!(d)
    check(ccall((:swb_fill, lib), Int32, (Ptr{Cvoid}, Int32, Cdouble, Csize_t, Csize_t), d.ptr, dtype_code(T), Float64(v), first(r) - 1, length(r)))
    return d
end
Base.Broadcast.materialize!(dest::SubArray{T, 1, <:B200Array{T, 1}, Tuple{UnitRange{Int}}}, bc::Base.Broadcast.Broadcasted{<:Any, <:Any, typeof(identity), <:Tuple{Number}}) where {T} =
    fillrange!(parent(dest), dest.indices[1], bc.args[1])
Base.Broadcast.materialize!(dest::B200Array, bc::Base.Broadcast.Broadcasted{<:Any, <:Any, typeof(identity), <:Tuple{Number}}) = fill!(dest, bc.args[1])
# `dest .= x` reaches materialize! with the bare scalar first (Base.broadcasted(identity, x::Number) === x): catch that form too
Base.Broadcast.materialize!(dest::B200Array, x::Number) = fill!(dest, x)
Base.Broadcast.materialize!(dest::SubArray{T, 1, <:B200Array{T, 1}, Tuple{UnitRange{Int}}}, x::Number) where {T} = fillrange!(parent(dest), dest.indices[1], x)

# ---------------------------------------------------------------------------------------------------------------------
# C structs of include/swb200.h (layout must match field for field; checked by tests/test_julia_boundary.py)
# ---------------------------------------------------------------------------------------------------------------------
struct CpmlAxis                # swb_cpml_axis
    a::Ptr{Cvoid}; a_h::Ptr{Cvoid}; b::Ptr{Cvoid}; b_h::Ptr{Cvoid}
end
struct Points                  # swb_points
    n::Int64; pos::Ptr{Cvoid}; tf::Ptr{Cvoid}; nt::Int64
end
const NOPOINTS = Points(0, C_NULL, C_NULL, 0)
struct AcouCDStepArgs          # swb_acou_cd_step_args
    dtype::Int32; ndim::Int32; halo::Int32; flags::Int32
    n::NTuple{3, Int64}; spacing::NTuple{3, Float64}
    pold::Ptr{Cvoid}; pcur::Ptr{Cvoid}; pnew::Ptr{Cvoid}; fact::Ptr{Cvoid}
    psi::NTuple{3, Ptr{Cvoid}}; xi::NTuple{3, Ptr{Cvoid}}
    cpml::NTuple{3, CpmlAxis}
    src::Points; rec::Points
    it::Int64; stream::Ptr{Cvoid}
end
struct AcouVDStepArgs          # swb_acou_vd_step_args
    dtype::Int32; halo::Int32; flags::Int32; _pad::Int32
    n::NTuple{2, Int64}; spacing::NTuple{2, Float64}
    pcur::Ptr{Cvoid}; vcur::NTuple{2, Ptr{Cvoid}}
    fact_m0::Ptr{Cvoid}; fact_m1_stag::NTuple{2, Ptr{Cvoid}}
    psi::NTuple{2, Ptr{Cvoid}}; xi::NTuple{2, Ptr{Cvoid}}
    cpml::NTuple{2, CpmlAxis}
    src::Points; rec::Points
    it::Int64; stream::Ptr{Cvoid}
end
struct SincPoints              # swb_sinc_points (device CSR)
    n::Int64; off::Ptr{Cvoid}; ij::Ptr{Cvoid}; coef::Ptr{Cvoid}
end
const NOSINC = SincPoints(0, C_NULL, C_NULL, C_NULL)
struct ElaStepArgs             # swb_ela_step_args
    dtype::Int32; halo::Int32; flags::Int32; freetop::Int32
    n::NTuple{2, Int64}; spacing::NTuple{2, Float64}; dt::Float64
    uold::NTuple{2, Ptr{Cvoid}}; ucur::NTuple{2, Ptr{Cvoid}}; unew::NTuple{2, Ptr{Cvoid}}
    sigma::NTuple{3, Ptr{Cvoid}}
    lambda::Ptr{Cvoid}; mu::Ptr{Cvoid}; rho_ihalf::Ptr{Cvoid}; rho_jhalf::Ptr{Cvoid}; mu_ihalf_jhalf::Ptr{Cvoid}
    psi_dsdx::NTuple{2, Ptr{Cvoid}}; psi_dsdz::NTuple{2, Ptr{Cvoid}}; psi_dudx::NTuple{2, Ptr{Cvoid}}; psi_dudz::NTuple{2, Ptr{Cvoid}}
    cpml::NTuple{2, CpmlAxis}
    src_kind::Int32; _pad::Int32
    src_pts::NTuple{2, SincPoints}
    srctf::Ptr{Cvoid}; nt_tf::Int64
    Mxx::Ptr{Cvoid}; Mzz::Ptr{Cvoid}; Mxz::Ptr{Cvoid}
    rec_pts::NTuple{2, SincPoints}
    traces::Ptr{Cvoid}; nt_tr::Int64
    it::Int64; stream::Ptr{Cvoid}
end
struct ElaCorrelateArgs        # swb_ela_correlate_args
    dtype::Int32; flags::Int32; freetop::Int32; _pad::Int32
    n::NTuple{2, Int64}; spacing::NTuple{2, Float64}; dt::Float64
    adjucur::NTuple{2, Ptr{Cvoid}}
    u_itm2::NTuple{2, Ptr{Cvoid}}; u_itm1::NTuple{2, Ptr{Cvoid}}; u_it::NTuple{2, Ptr{Cvoid}}
    lambda::Ptr{Cvoid}; mu::Ptr{Cvoid}
    grad_rho_ihalf::Ptr{Cvoid}; grad_rho_jhalf::Ptr{Cvoid}; grad_lambda::Ptr{Cvoid}; grad_mu::Ptr{Cvoid}; grad_mu_ihalf_jhalf::Ptr{Cvoid}
    stream::Ptr{Cvoid}
end
struct SimDesc                 # swb_sim_desc
    kind::Int32; dtype::Int32; ndim::Int32; device::Int32
    n::NTuple{3, Int64}; spacing::NTuple{3, Float64}; dt::Float64; nt::Int64
    halo::Int32; freetop::Int32; gradient::Int32; check_freq::Int32
    flags::Int32; _pad::Int32
end
struct SincPointsHost          # swb_sinc_points_host
    n::Int64; off::Ptr{Int64}; ij::Ptr{Int32}; coef::Ptr{Cvoid}
end
struct L2Spec                  # swb_l2_spec
    observed::Ptr{Cvoid}; invcov_diag::Ptr{Cvoid}; mask::Ptr{Cvoid}
end
struct SlabHandle              # swb_slab_handle
    ipc::NTuple{192, UInt8}; raw::NTuple{3, UInt64}
    nz::Int64; plane_elems::Int64
    device::Int32; pid::Int32
end

pad3(t::NTuple{N, T}, z) where {N, T} = ntuple(i -> i <= N ? t[i] : z, 3)
pad2(t::NTuple{N, T}, z) where {N, T} = ntuple(i -> i <= N ? t[i] : z, 2)
vptr(a::B200Array) = materialize!(a).ptr
vptr(::Nothing) = C_NULL
cpmlaxis(c) = CpmlAxis(vptr(c.a), vptr(c.a_h), vptr(c.b), vptr(c.b_h))
const NOAXIS = CpmlAxis(C_NULL, C_NULL, C_NULL, C_NULL)

# ---------------------------------------------------------------------------------------------------------------------
# Level 1: backend modules (one ccall per reference backend function)
# ---------------------------------------------------------------------------------------------------------------------
module Data
    import ..B200Array
    const Array = B200Array
end
zeros(::Type{T}, dims::Int...) where {T} = B200Array{T, length(dims)}(dims)
ones(::Type{T}, dims::Int...) where {T} = fill!(B200Array{T, length(dims)}(dims), one(T))

function acou_cd_args(model, possrcs, srctf, posrecs, traces, it, fields::NTuple{3, String}, ψname, ξname)
    T = eltype(model.grid.fields["fact"].value)
    N = length(model.grid.size)
    g = model.grid
    ψ, ξ = g.fields[ψname].value, g.fields[ξname].value
    src = Points(size(possrcs, 1), vptr(possrcs), vptr(srctf), size(srctf, 1))
    rec = traces === nothing ? NOPOINTS : Points(size(posrecs, 1), vptr(posrecs), vptr(traces), size(traces, 1))
    return AcouCDStepArgs(dtype_code(T), N, model.cpmlparams.halo, FLAGS[],
        pad3(Int64.(g.size), Int64(1)), pad3(Float64.(g.spacing), 0.0),
        vptr(g.fields[fields[1]].value), vptr(g.fields[fields[2]].value), vptr(g.fields[fields[3]].value), vptr(g.fields["fact"].value),
        pad3(Tuple(vptr.(ψ)), C_NULL), pad3(Tuple(vptr.(ξ)), C_NULL), pad3(Tuple(cpmlaxis.(model.cpmlcoeffs)), NOAXIS),
        src, rec, it, C_NULL)
end

function rotate!(g, a, b, c)   # acoustic2D_xPU.jl:121-124
    g.fields[a] = g.fields[b]; g.fields[b] = g.fields[c]; g.fields[c] = g.fields[a]
end

# forward_onestep_CPML! / adjoint_onestep_CPML! of acoustic{2,3}D_xPU.jl
function cd_forward_onestep_CPML!(model, possrcs, srctf, posrecs, traces, it; save_trace=true)
    args = acou_cd_args(model, possrcs, srctf, save_trace ? posrecs : nothing, save_trace ? traces : nothing, it, ("pold", "pcur", "pnew"), "ψ", "ξ")
    check(ccall((:swb_acou_cd_forward_onestep, lib), Int32, (Ref{AcouCDStepArgs},), args))
    rotate!(model.grid, "pold", "pcur", "pnew")
    return nothing
end
function cd_adjoint_onestep_CPML!(model, posrecs, adjsrc, it)
    args = acou_cd_args(model, posrecs, adjsrc, nothing, nothing, it, ("adjold", "adjcur", "adjnew"), "ψ_adj", "ξ_adj")
    check(ccall((:swb_acou_cd_adjoint_onestep, lib), Int32, (Ref{AcouCDStepArgs},), args))
    rotate!(model.grid, "adjold", "adjcur", "adjnew")
    return nothing
end
function prescale_residuals!(res::B200Array{T}, posrecs, fact::B200Array{T, N}) where {T, N}
    n = collect(Int64, size(fact))
    check(ccall((:swb_prescale_residuals, lib), Int32, (Int32, Int32, Ptr{Int64}, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
        dtype_code(T), N, n, vptr(res), size(res, 1), size(res, 2), vptr(posrecs), vptr(fact), C_NULL))
end
function correlate_gradient!(grad::B200Array{T}, adjcur, p_itm2, p_itm1, p_it, dt) where {T}
    check(ccall((:swb_acou_cd_correlate_gradient, lib), Int32, (Int32, Int32, Csize_t, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
        dtype_code(T), FLAGS[], length(grad), vptr(grad), vptr(adjcur), vptr(p_itm2), vptr(p_itm1), vptr(p_it), Float64(dt), C_NULL))
end

function acou_vd_args(model, possrcs, srctf, posrecs, traces, it, pname, vname, ψname, ξname)
    g = model.grid
    T = eltype(g.fields["fact_m0"].value)
    v, m1, ψ, ξ = g.fields[vname].value, g.fields["fact_m1_stag"].value, g.fields[ψname].value, g.fields[ξname].value
    src = Points(size(possrcs, 1), vptr(possrcs), vptr(srctf), size(srctf, 1))
    rec = traces === nothing ? NOPOINTS : Points(size(posrecs, 1), vptr(posrecs), vptr(traces), size(traces, 1))
    # a 1D simulation (acoustic1D_VD_xPU.jl) is the single-row case of the same entry point: n = (nx, 1), no y arrays
    return AcouVDStepArgs(dtype_code(T), model.cpmlparams.halo, FLAGS[], 0, pad2(Tuple(Int64.(g.size)), Int64(1)), pad2(Tuple(Float64.(g.spacing)), 0.0),
        vptr(g.fields[pname].value), pad2(Tuple(vptr.(v)), C_NULL), vptr(g.fields["fact_m0"].value), pad2(Tuple(vptr.(m1)), C_NULL),
        pad2(Tuple(vptr.(ψ)), C_NULL), pad2(Tuple(vptr.(ξ)), C_NULL), pad2(Tuple(cpmlaxis.(model.cpmlcoeffs)), NOAXIS), src, rec, it, C_NULL)
end
function vd_forward_onestep_CPML!(model, possrcs, srctf, posrecs, traces, it; save_trace=true)
    args = acou_vd_args(model, possrcs, srctf, save_trace ? posrecs : nothing, save_trace ? traces : nothing, it, "pcur", "vcur", "ψ", "ξ")
    check(ccall((:swb_acou_vd_forward_onestep, lib), Int32, (Ref{AcouVDStepArgs},), args))
end
function vd_adjoint_onestep_CPML!(model, posrecs, adjsrc, it)
    args = acou_vd_args(model, posrecs, adjsrc, nothing, nothing, it, "adjpcur", "adjvcur", "ψ_adj", "ξ_adj")
    check(ccall((:swb_acou_vd_adjoint_onestep, lib), Int32, (Ref{AcouVDStepArgs},), args))
end
function correlate_gradient_m0!(grad_m0::B200Array{T}, adjp, p_it, p_itm1, dt) where {T}
    check(ccall((:swb_acou_vd_correlate_gradient_m0, lib), Int32, (Int32, Int32, Csize_t, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
        dtype_code(T), FLAGS[], length(grad_m0), vptr(grad_m0), vptr(adjp), vptr(p_it), vptr(p_itm1), Float64(dt), C_NULL))
end
function correlate_gradient_m1!(grad_m1_stag, adjv, p_it::B200Array{T, N}, spacing) where {T, N}
    n, sp = collect(Int64, pad2(size(p_it), 1)), collect(Float64, pad2(Tuple(spacing), 0.0))   # 1D: n = (nx, 1), range (2:nx-2) inside the library
    g, av = Ptr{Cvoid}[vptr(x) for x in grad_m1_stag], Ptr{Cvoid}[vptr(x) for x in adjv]
    N == 1 && (push!(g, C_NULL); push!(av, C_NULL))
    check(ccall((:swb_acou_vd_correlate_gradient_m1, lib), Int32, (Int32, Int32, Ptr{Int64}, Ptr{Cdouble}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{Cvoid}),
        dtype_code(T), FLAGS[], n, sp, g, av, vptr(p_it), C_NULL))
end

# ---- elastic P-SV (elastic2D_iso_xPU.jl:120-448, elastic/backends/shared/correlate_gradient_xPU.jl:47-83) -------------------
# The reference passes the sinc point lists as one device array pair per source / receiver (ela_forward.jl:25-32); the C entry
# points take them flattened to CSR on the device.  The lists are constant over a shot, so the CSR form is built on the first
# step and cached per list object.
mutable struct DevCSR{T}
    n::Int64
    off::B200Array{Int64, 1}
    ij::B200Array{Int32, 2}
    coef::B200Array{T, 1}
end
const CSR_CACHE = IdDict{Any, Any}()   # Vector{B200Array} (ij lists) -> DevCSR; emptied by forget_lists!() at the end of a shot
function devcsr(ijs::Vector, vals::Vector, nx::Int, ::Type{T}) where {T}
    get!(CSR_CACHE, ijs) do
        off, ij, co = csr([Array(x) for x in ijs], [Array(x) for x in vals], nx)
        DevCSR{T}(length(ijs), B200Array(off), B200Array(ij), B200Array(co))
    end
end
sincpts(c::DevCSR) = SincPoints(c.n, vptr(c.off), vptr(c.ij), vptr(c.coef))
forget_lists!() = empty!(CSR_CACHE)

function ela_args(model, names, src_kind, src_a, src_b, srctf, Ms, rec_ux, rec_uz, traces, it)
    g = model.grid
    T = eltype(g.fields["λ"].value)
    f(name) = g.fields[name].value
    two(v) = (vptr(v[1]), vptr(v[2]))
    σ = f(names.σ)
    return ElaStepArgs(dtype_code(T), model.cpmlparams.halo, FLAGS[], model.cpmlparams.freeboundtop,
        Tuple(Int64.(g.size)), Tuple(Float64.(g.spacing)), Float64(model.dt),
        two(f(names.uold)), two(f(names.ucur)), two(f(names.unew)), (vptr(σ[1]), vptr(σ[2]), vptr(σ[3])),
        vptr(f("λ")), vptr(f("μ")), vptr(f("ρ_ihalf")), vptr(f("ρ_jhalf")), vptr(f("μ_ihalf_jhalf")),
        two(f(names.ψσx)), two(f(names.ψσz)), two(f(names.ψux)), two(f(names.ψuz)),
        Tuple(cpmlaxis.(model.cpmlcoeffs)),
        Int32(src_kind), Int32(0), (src_a, src_b), vptr(srctf), size(srctf, 1), Ms[1], Ms[2], Ms[3],
        (rec_ux, rec_uz), traces === nothing ? C_NULL : vptr(traces), traces === nothing ? 0 : size(traces, 1), it, C_NULL)
end
const ELA_FWD = (σ="σ", uold="uold", ucur="ucur", unew="unew", ψσx="ψ_∂σ∂x", ψσz="ψ_∂σ∂z", ψux="ψ_∂u∂x", ψuz="ψ_∂u∂z")
const ELA_ADJ = (σ="adjσ", uold="adjuold", ucur="adjucur", unew="adjunew", ψσx="adjψ_∂σ∂x", ψσz="adjψ_∂σ∂z", ψux="adjψ_∂u∂x", ψuz="adjψ_∂u∂z")

# moment-tensor method (elastic2D_iso_xPU.jl:120-242); reduced_buf / traces_*_bk_buf are the reference's scratch buffers, unused here
function ela_forward_onestep_CPML!(model, srccoeij_xx, srccoeval_xx, srccoeij_xz, srccoeval_xz, reccoeij_ux, reccoeval_ux, reccoeij_uz, reccoeval_uz,
                                   srctf_bk, reduced_buf, traces_ux_bk_buf, traces_uz_bk_buf, traces_bk, it::Int, Mxx_bk, Mzz_bk, Mxz_bk; save_trace::Bool=true)
    T, nx = eltype(srctf_bk), model.grid.size[1]
    sa, sb = sincpts(devcsr(srccoeij_xx, srccoeval_xx, nx, T)), sincpts(devcsr(srccoeij_xz, srccoeval_xz, nx, T))
    ra, rb = save_trace ? (sincpts(devcsr(reccoeij_ux, reccoeval_ux, nx, T)), sincpts(devcsr(reccoeij_uz, reccoeval_uz, nx, T))) : (NOSINC, NOSINC)
    args = ela_args(model, ELA_FWD, 1, sa, sb, srctf_bk, (vptr(Mxx_bk), vptr(Mzz_bk), vptr(Mxz_bk)), ra, rb, save_trace ? traces_bk : nothing, it)
    check(ccall((:swb_ela_forward_onestep, lib), Int32, (Ref{ElaStepArgs},), args))
    rotate!(model.grid, "uold", "ucur", "unew")
    return nothing
end
# external-force method (elastic2D_iso_xPU.jl:244-357)
function ela_forward_onestep_CPML!(model, srccoeij_ux, srccoeval_ux, srccoeij_uz, srccoeval_uz, reccoeij_ux, reccoeval_ux, reccoeij_uz, reccoeval_uz,
                                   srctf_bk, reduced_buf, traces_ux_bk_buf, traces_uz_bk_buf, traces_bk, it::Int; save_trace::Bool=true)
    T, nx = eltype(srctf_bk), model.grid.size[1]
    sa, sb = sincpts(devcsr(srccoeij_ux, srccoeval_ux, nx, T)), sincpts(devcsr(srccoeij_uz, srccoeval_uz, nx, T))
    ra, rb = save_trace ? (sincpts(devcsr(reccoeij_ux, reccoeval_ux, nx, T)), sincpts(devcsr(reccoeij_uz, reccoeval_uz, nx, T))) : (NOSINC, NOSINC)
    args = ela_args(model, ELA_FWD, 2, sa, sb, srctf_bk, (C_NULL, C_NULL, C_NULL), ra, rb, save_trace ? traces_bk : nothing, it)
    check(ccall((:swb_ela_forward_onestep, lib), Int32, (Ref{ElaStepArgs},), args))
    rotate!(model.grid, "uold", "ucur", "unew")
    return nothing
end
# adjoint_onestep_CPML! (elastic2D_iso_xPU.jl:359-448): the residuals (nt, 2, nrec) enter as external forces through the receivers' lists
function ela_adjoint_onestep_CPML!(model, reccoeij_ux, reccoeval_ux, reccoeij_uz, reccoeval_uz, residuals_bk, it)
    T, nx = eltype(residuals_bk), model.grid.size[1]
    sa, sb = sincpts(devcsr(reccoeij_ux, reccoeval_ux, nx, T)), sincpts(devcsr(reccoeij_uz, reccoeval_uz, nx, T))
    args = ela_args(model, ELA_ADJ, 2, sa, sb, residuals_bk, (C_NULL, C_NULL, C_NULL), NOSINC, NOSINC, nothing, it)
    check(ccall((:swb_ela_adjoint_onestep, lib), Int32, (Ref{ElaStepArgs},), args))
    rotate!(model.grid, "adjuold", "adjucur", "adjunew")
    return nothing
end
# correlate_gradients! (correlate_gradient_xPU.jl:47-83), called as backend.correlate_gradients!(grid, uold_corr, ucur_corr, unew_corr, dt, freeboundtop)
function correlate_gradients!(grid, uold_corr, ucur_corr, unew_corr, dt, freeboundtop)
    f(name) = grid.fields[name].value
    T = eltype(f("λ"))
    two(v) = (vptr(v[1]), vptr(v[2]))
    args = ElaCorrelateArgs(dtype_code(T), FLAGS[], freeboundtop, 0, Tuple(Int64.(grid.size)), Tuple(Float64.(grid.spacing)), Float64(dt),
        two(f("adjucur")), two(uold_corr), two(ucur_corr), two(unew_corr), vptr(f("λ")), vptr(f("μ")),
        vptr(f("grad_ρ_ihalf")), vptr(f("grad_ρ_jhalf")), vptr(f("grad_λ")), vptr(f("grad_μ")), vptr(f("grad_μ_ihalf_jhalf")), C_NULL)
    check(ccall((:swb_ela_correlate_gradients, lib), Int32, (Ref{ElaCorrelateArgs},), args))
    return nothing
end

# the backend "modules": NamedTuples carry the same member names the reference looks up with backend.<name>
const Acoustic2D_CD_CPML_B200 = (; Data, zeros, ones, forward_onestep_CPML! = cd_forward_onestep_CPML!, adjoint_onestep_CPML! = cd_adjoint_onestep_CPML!,
                                  prescale_residuals!, correlate_gradient!)
const Acoustic3D_CD_CPML_B200 = Acoustic2D_CD_CPML_B200          # the C entry points take ndim
const Acoustic1D_CD_CPML_B200 = Acoustic2D_CD_CPML_B200
const Acoustic2D_VD_CPML_B200 = (; Data, zeros, ones, forward_onestep_CPML! = vd_forward_onestep_CPML!, adjoint_onestep_CPML! = vd_adjoint_onestep_CPML!,
                                  prescale_residuals!, correlate_gradient_m0!, correlate_gradient_m1!)
const Acoustic1D_VD_CPML_B200 = Acoustic2D_VD_CPML_B200          # n = (nx, 1)
const Elastic2D_Iso_CPML_B200 = (; Data, zeros, ones, forward_onestep_CPML! = ela_forward_onestep_CPML!, adjoint_onestep_CPML! = ela_adjoint_onestep_CPML!,
                                  correlate_gradients!)

const FT = Union{Float32, Float64}
SeismicWaves.select_backend(::CPMLBoundaryCondition, ::LocalGrid, ::Type{<:AcousticCDCPMLWaveSimulation{<:FT, 1}}, ::Type{Val{:B200}}) = Acoustic1D_CD_CPML_B200
SeismicWaves.select_backend(::CPMLBoundaryCondition, ::LocalGrid, ::Type{<:AcousticVDStaggeredCPMLWaveSimulation{<:FT, 1}}, ::Type{Val{:B200}}) = Acoustic1D_VD_CPML_B200
SeismicWaves.select_backend(::CPMLBoundaryCondition, ::LocalGrid, ::Type{<:AcousticCDCPMLWaveSimulation{<:FT, 2}}, ::Type{Val{:B200}}) = Acoustic2D_CD_CPML_B200
SeismicWaves.select_backend(::CPMLBoundaryCondition, ::LocalGrid, ::Type{<:AcousticCDCPMLWaveSimulation{<:FT, 3}}, ::Type{Val{:B200}}) = Acoustic3D_CD_CPML_B200
SeismicWaves.select_backend(::CPMLBoundaryCondition, ::LocalGrid, ::Type{<:AcousticVDStaggeredCPMLWaveSimulation{<:FT, 2}}, ::Type{Val{:B200}}) = Acoustic2D_VD_CPML_B200
SeismicWaves.select_backend(::CPMLBoundaryCondition, ::LocalGrid, ::Type{<:ElasticIsoCPMLWaveSimulation{<:FT, 2}}, ::Type{Val{:B200}}) = Elastic2D_Iso_CPML_B200

# ---------------------------------------------------------------------------------------------------------------------
# Level 2: whole shots on the per-shot engine (performance path)
# ---------------------------------------------------------------------------------------------------------------------
# The reference's simulation types are immutable structs (acou_models.jl:68,311; ela_models.jl:177): no finalizer can hang on them.
# The engine handle therefore lives in a mutable field object stored in the simulation's own field dictionary
# (`model.grid.fields["b200_engine"]`): it is created on first use, freed by its finalizer when the simulation is collected, or
# explicitly with close!(model).  B200-only knobs the reference's RunParameters has no field for -- the GPU a simulation lives on
# (default 0) and the creation flags (SWB_FLAG_FAST_F32 = 1: Float32 arithmetic; SWB_FLAG_NO_FUSION = 2; SWB_FLAG_NO_GRAPH = 4) --
# are kept there too; set them before the first shot.
mutable struct EngineField{T} <: AbstractField{T}
    h::Ptr{Cvoid}
    device::Int32
    flags::Int32
    function EngineField{T}(device::Integer, flags::Integer) where {T}
        e = new{T}(C_NULL, Int32(device), Int32(flags))
        finalizer(destroy!, e)
        return e
    end
end
function destroy!(e::EngineField)
    if e.h != C_NULL
        ccall((:swb_sim_destroy, lib), Int32, (Ptr{Cvoid},), e.h)
        e.h = C_NULL
    end
    return nothing
end
SeismicWaves.setzero!(e::EngineField) = e          # reset!(grid) walks every field (grids.jl:28-33); the engine resets itself per shot
Base.zero(e::EngineField{T}) where {T} = EngineField{T}(e.device, e.flags)

const FLAGS = Ref{Int32}(0)                       # default creation flags of new engines (and the flags of the Level-1 calls)
set_flags!(flags::Integer) = (FLAGS[] = Int32(flags))
function enginefield(model)
    T = typeof(model.dt)
    get!(() -> EngineField{T}(0, FLAGS[]), model.grid.fields, "b200_engine")::EngineField{T}
end
set_device!(model, dev::Integer) = (enginefield(model).device = Int32(dev); model)
device_of(model) = enginefield(model).device
close!(model) = destroy!(enginefield(model))

simkind(::AcousticCDCPMLWaveSimulation) = Int32(1)
simkind(::AcousticVDStaggeredCPMLWaveSimulation) = Int32(2)
simkind(::ElasticIsoCPMLWaveSimulation) = Int32(3)

function engine(model)
    e = enginefield(model)
    if e.h == C_NULL
        T = typeof(model.dt)
        N = length(model.grid.size)
        cf = model.gradparams === nothing ? 1 : model.gradparams.check_freq
        desc = SimDesc(simkind(model), dtype_code(T), N, e.device, pad3(Int64.(model.grid.size), Int64(1)), pad3(Float64.(model.grid.spacing), 0.0), Float64(model.dt), model.nt,
            model.cpmlparams.halo, model.cpmlparams.freeboundtop, model.checkpointer !== nothing, cf, e.flags, Int32(0))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:swb_sim_create, lib), Int32, (Ref{SimDesc}, Ref{Ptr{Cvoid}}), desc, h))
        e.h = h[]
    end
    return e.h
end

matfields(m::AcousticCDCPMLWaveSimulation) = (m.matprop.vp,)
matfields(m::AcousticVDStaggeredCPMLWaveSimulation) = (m.matprop.vp, m.matprop.rho)
matfields(m::ElasticIsoCPMLWaveSimulation) = (m.matprop.ρ, m.matprop.λ, m.matprop.μ)

# interpolation code of swb_sim_set_material: bit 0 = density (VD: rho; elastic: ρ), bit 1 = elastic μ; 1 = harmonic average
interpbit(::SeismicWaves.ArithmeticAverageInterpolation) = Int32(0)
interpbit(::SeismicWaves.HarmonicAverageInterpolation) = Int32(1)
interpbit(m) = error("the :B200 backend supports ArithmeticAverageInterpolation and HarmonicAverageInterpolation only (got $(typeof(m)))")
interpcode(::AcousticCDCPMLWaveSimulation) = Int32(0)
interpcode(m::AcousticVDStaggeredCPMLWaveSimulation) = interpbit(m.matprop.interp_method)
interpcode(m::ElasticIsoCPMLWaveSimulation) = interpbit(m.matprop.interp_method_ρ) + Int32(2) * interpbit(m.matprop.interp_method_μ)

# update_matprop! + precompute_fact! / precomp_elaprop! on the device: once per material update
function upload_material!(model)
    fs = matfields(model)
    ptrs = [Ptr{Cvoid}(pointer(f)) for f in fs]
    GC.@preserve fs check(ccall((:swb_sim_set_material, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}, Int32), engine(model), length(fs), ptrs, interpcode(model)))
end
# C-PML coefficient vectors of the CURRENT shot: init_shot! -> init_bdc! recomputes them per shot (they depend on the shot's dominant
# frequency, acou_init_bc.jl:20-39), so this runs after init_shot! for every shot
function upload_cpml!(model)
    for (ax, c) in enumerate(model.cpmlcoeffs)
        a, a_h, b, b_h = Array(c.a), Array(c.a_h), Array(c.b), Array(c.b_h)
        GC.@preserve a a_h b b_h check(ccall((:swb_sim_set_cpml, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), engine(model), ax - 1,
            a, a_h, b, b_h))
    end
end
upload_model!(model) = (upload_material!(model); upload_cpml!(model))

function bind!(model::Union{AcousticCDCPMLWaveSimulation, AcousticVDStaggeredCPMLWaveSimulation}, shot::ScalarShot)
    possrcs, posrecs, scal_srctf = SeismicWaves.possrcrec_scaletf(model, shot)     # acou_forward.jl:6-20,67-81 (host)
    check(ccall((:swb_sim_bind_scalar_shot, lib), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Cvoid}, Int64, Ptr{Int64}), engine(model),
        size(possrcs, 1), possrcs, scal_srctf, size(posrecs, 1), posrecs))
end

# snapshots taken by the engine (runparams.snapevery) -> model.snapshotter.snapshots[it][name] (snapshotter.jl:33-41);
# field order of swb_sim_get_snapshot: CD pcur; VD pcur, vcur[1:2]; elastic ucur[1:2], σ[1:3]
snapnames(::AcousticCDCPMLWaveSimulation) = ("pcur",)
snapnames(::AcousticVDStaggeredCPMLWaveSimulation) = ("pcur", "vcur")
snapnames(::ElasticIsoCPMLWaveSimulation) = ("ucur", "σ")
function fetch_snapshots!(model)
    model.snapshotter === nothing && return nothing
    h = engine(model)
    for (it, fields) in model.snapshotter.snapshots
        idx = 0
        for name in snapnames(model)
            v = fields[name].value
            for comp in (v isa Vector ? v : (v,))
                check(ccall((:swb_sim_get_snapshot, lib), Int32, (Ptr{Cvoid}, Int64, Int32, Ptr{Cvoid}), h, it, idx, comp))
                idx += 1
            end
        end
    end
    return nothing
end

function run_forward!(model, shot)
    upload_model!(model)
    bind!(model, shot)
    snapevery = model.runparams.snapevery === nothing ? 0 : model.runparams.snapevery
    check(ccall((:swb_sim_forward, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32), engine(model), shot.recs.seismograms, snapevery))
    fetch_snapshots!(model)
    return nothing
end

# windows + diagonal / identity inverse covariance of an L2Misfit as the per-time-sample vectors swb_sim_gradient_l2_ex takes;
# `nothing` for anything else (dense covariance, other misfits): those go through the host path
function l2_device_spec(misfit, ::Type{T}, nt::Int) where {T}
    misfit isa SeismicWaves.L2Misfit || return nothing
    ic = misfit.invcov
    w = if ic isa SeismicWaves.LinearAlgebra.Diagonal
        Vector{T}(ic.diag)
    elseif ic isa SeismicWaves.LinearAlgebra.UniformScaling
        fill(T(ic.λ), nt)
    else
        return nothing
    end
    mask = ones(T, nt)
    if length(misfit.windows) > 0
        mask .= 0
        for wnd in misfit.windows
            mask[wnd.first:wnd.second] .= 1
        end
    end
    return w, mask
end

# gradient of one shot, raw correlation included; returns after the adjoint loop (the gradient stays on the device)
function run_gradient!(model, shot, misfit::AbstractMisfit{T}) where {T}
    h = engine(model)
    spec = l2_device_spec(misfit, T, model.nt)
    if spec !== nothing        # forward -> residual, weights, window -> adjoint without leaving the device (SURVEY 8f.1)
        w, mask = spec
        obs = misfit.observed
        GC.@preserve obs w mask begin
            l2 = L2Spec(pointer(obs), pointer(w), pointer(mask))
            check(ccall((:swb_sim_gradient_l2_ex, lib), Int32, (Ptr{Cvoid}, Ref{L2Spec}, Ptr{Cvoid}, Ptr{Cdouble}), h, l2, shot.recs.seismograms, C_NULL))
        end
    else                       # any AbstractMisfit: the adjoint source is computed on the host between the two phases
        check(ccall((:swb_sim_gradient_forward, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), h, shot.recs.seismograms))
        adjsrc = .-SeismicWaves.∂χ_∂u(misfit, shot.recs)
        check(ccall((:swb_sim_gradient_adjoint, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), h, adjsrc))
    end
    return h
end

gradnames(::AcousticCDCPMLWaveSimulation) = ("vp",)
gradnames(::AcousticVDStaggeredCPMLWaveSimulation) = ("vp", "rho")
gradnames(::ElasticIsoCPMLWaveSimulation) = ("rho", "lambda", "mu")

# mute + back_interp + chain rule of THIS shot on the device, accumulated into the engine's total
function accumulate!(model, shot)
    gp = model.gradparams
    check(ccall((:swb_sim_accumulate_gradient, lib), Int32, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int32, Int64, Ptr{Cvoid}, Int32), engine(model),
        size(shot.srcs.positions, 1), shot.srcs.positions, gp.mute_radius_src, size(shot.recs.positions, 1), shot.recs.positions, gp.mute_radius_rec))
end
function download_total(model, ::Type{T}, ::Val{N}) where {T, N}
    out = Dict{String, Array{T, N}}()
    for (k, name) in enumerate(gradnames(model))
        g = Base.zeros(T, model.grid.size...)
        check(ccall((:swb_sim_get_total_gradient, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}), engine(model), k - 1, g))
        out[name] = g
    end
    return out
end
function gradient_1shot!(model, shot, misfit::AbstractMisfit{T}, ::Val{N}) where {T, N}
    upload_model!(model)
    bind!(model, shot)
    h = run_gradient!(model, shot, misfit)
    check(ccall((:swb_sim_zero_total_gradient, lib), Int32, (Ptr{Cvoid},), h))
    accumulate!(model, shot)
    return download_total(model, T, Val(N))       # the reference's loop adds it to its host total (gradient.jl:124-127)
end

# swforward_1shot! (acou_forward.jl:22-125) / swgradient_1shot! (acou_gradient.jl:4-203): the whole time loops in one call each
const AcousticB200{T, N} = Union{AcousticCDCPMLWaveSimulation{T, N, <:B200Array}, AcousticVDStaggeredCPMLWaveSimulation{T, N, <:B200Array}}
SeismicWaves.swforward_1shot!(::CPMLBoundaryCondition, model::AcousticB200{T, N}, shot::ScalarShot{T}) where {T, N} = run_forward!(model, shot)
SeismicWaves.swgradient_1shot!(::CPMLBoundaryCondition, model::AcousticB200{T, N}, shot::ScalarShot{T}, misfit::AbstractMisfit{T}) where {T, N} =
    gradient_1shot!(model, shot, misfit, Val(N))

# ---------------------------------------------------------------------------------------------------------------------
# Elastic P-SV shots through the per-shot engine (ela_forward.jl:4-159, ela_gradient.jl:4-362).  The host part of the reference
# stays as it is: `possrcrec_scaletf` (ela_models.jl:6-90) spreads the off-grid positions into Kaiser-windowed sinc point lists
# and scales the source time functions; the lists cross the ABI flattened to CSR (include/swb200.h, swb_sinc_points_host).
# ---------------------------------------------------------------------------------------------------------------------
# per-position lists (Vector of (npts, 2) index matrices, Vector of coefficient vectors) -> CSR; the points of one position in
# ascending linear index (the reference iterates a Dict; any fixed order is equivalent up to the summation order of a receiver)
function csr(ijs::Vector{<:AbstractMatrix{<:Integer}}, vals::Vector{<:AbstractVector{T}}, nx::Int) where {T}
    off = Base.zeros(Int64, length(ijs) + 1)
    for k in eachindex(ijs)
        off[k + 1] = off[k] + size(ijs[k], 1)
    end
    ij = Base.zeros(Int32, off[end], 2)
    co = Base.zeros(T, off[end])
    for k in eachindex(ijs)
        order = sortperm([(ijs[k][p, 2] - 1) * nx + ijs[k][p, 1] for p in 1:size(ijs[k], 1)])
        ij[off[k] + 1:off[k + 1], :] .= ijs[k][order, :]
        co[off[k] + 1:off[k + 1]] .= vals[k][order]
    end
    return off, ij, co
end

const ElasticShot{T} = Union{MomentTensorShot{T, 2, MomentTensor2D{T}}, ExternalForceShot{T, 2}}

function bind!(model::ElasticIsoCPMLWaveSimulation{T, 2}, shot::ElasticShot{T}) where {T}
    src_a_ij, src_a_val, src_b_ij, src_b_val, rec_ux_ij, rec_ux_val, rec_uz_ij, rec_uz_val, scal_srctf =
        SeismicWaves.possrcrec_scaletf(model, shot; sincinterp=model.sincinterp)
    nx = model.grid.size[1]
    lists = (csr(src_a_ij, src_a_val, nx), csr(src_b_ij, src_b_val, nx), csr(rec_ux_ij, rec_ux_val, nx), csr(rec_uz_ij, rec_uz_val, nx))
    GC.@preserve lists scal_srctf begin
        pts = [SincPointsHost(length(l[1]) - 1, pointer(l[1]), pointer(l[2]), pointer(l[3])) for l in lists]
        src_pts, rec_pts = pts[1:2], pts[3:4]
        if shot isa MomentTensorShot          # lists on σxx/σzz and σxz, srctf (nt, nsrc), one moment tensor per source
            M = shot.srcs.momtens
            Mxx, Mzz, Mxz = T[m.Mxx for m in M], T[m.Mzz for m in M], T[m.Mxz for m in M]
            check(ccall((:swb_sim_bind_elastic_shot, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{SincPointsHost}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{SincPointsHost}),
                engine(model), 1, src_pts, scal_srctf, Mxx, Mzz, Mxz, rec_pts))
        else                                   # lists on ux and uz, srctf (nt, 2, nsrc)
            check(ccall((:swb_sim_bind_elastic_shot, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{SincPointsHost}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{SincPointsHost}),
                engine(model), 2, src_pts, scal_srctf, C_NULL, C_NULL, C_NULL, rec_pts))
        end
    end
end

SeismicWaves.swforward_1shot!(::CPMLBoundaryCondition, model::ElasticIsoCPMLWaveSimulation{T, 2, <:B200Array}, shot::MomentTensorShot{T, 2, MomentTensor2D{T}}) where {T} =
    run_forward!(model, shot)      # seismograms (nt, 2, nrec)
SeismicWaves.swforward_1shot!(::CPMLBoundaryCondition, model::ElasticIsoCPMLWaveSimulation{T, 2, <:B200Array}, shot::ExternalForceShot{T, 2}) where {T} =
    run_forward!(model, shot)
SeismicWaves.swgradient_1shot!(::CPMLBoundaryCondition, model::ElasticIsoCPMLWaveSimulation{T, 2, <:B200Array}, shot::MomentTensorShot{T, 2, MomentTensor2D{T}},
                               misfit::AbstractMisfit{T}) where {T} = gradient_1shot!(model, shot, misfit, Val(2))
SeismicWaves.swgradient_1shot!(::CPMLBoundaryCondition, model::ElasticIsoCPMLWaveSimulation{T, 2, <:B200Array}, shot::ExternalForceShot{T, 2},
                               misfit::AbstractMisfit{T}) where {T} = gradient_1shot!(model, shot, misfit, Val(2))

# ---------------------------------------------------------------------------------------------------------------------
# Multi-GPU shot sharding (SURVEY 8e): one simulation per GPU (`set_device!(wavesim[k], k - 1)` before the first shot), one Julia task per
# simulation, contiguous shot groups from `distribsrcs` (utils.jl:28-45) exactly as the reference's `:threadpersrc` mode
# (gradient.jl:139-208), per-shot post-processing on the device, then ONE NCCL all-reduce of the per-device totals
# (`swb_sim_allreduce_total_gradient`) and the result read back from device 0.  Needs JULIA_NUM_THREADS >= number of GPUs: the
# ranks of one NCCL communicator have to enter `swb_comm_create` and the all-reduce concurrently.
# The Python twin (`seismicwaves.jl_b200/multigpu.py`, `bench.py --gpus N`) runs the same sequence with one process per GPU.
# ---------------------------------------------------------------------------------------------------------------------
function SeismicWaves.run_swgradient!(wavesim::Vector{<:Union{AcousticCDCPMLWaveSimulation{T, N, <:B200Array}, AcousticVDStaggeredCPMLWaveSimulation{T, N, <:B200Array},
                                                             ElasticIsoCPMLWaveSimulation{T, N, <:B200Array}}},
                                      matprop::SeismicWaves.MaterialProperties{T, N}, shots::Vector{<:SeismicWaves.Shot{T}},
                                      misfit::Vector{<:AbstractMisfit{T}}) where {T, N}
    nshots = length(shots)
    ndev = min(length(wavesim), nshots)           # fewer shots than GPUs: the spare simulations stay out of the communicator
    @assert ndev >= 1 "no shots"
    @assert Threads.nthreads() >= ndev "one Julia thread per GPU is needed (NCCL ranks enter collectives concurrently)"
    for w in wavesim[1:ndev]
        SeismicWaves.check_sim_consistency(w, matprop, shots)
        SeismicWaves.set_wavesim_matprop!(w, matprop)
    end
    id = Base.zeros(UInt8, 128)
    ndev > 1 && check(ccall((:swb_comm_unique_id, lib), Int32, (Ptr{UInt8},), id))
    grpshots = SeismicWaves.distribsrcs(nshots, ndev)      # ndev <= nshots: exactly ndev non-empty contiguous groups
    misfitvals = Base.zeros(T, nshots)
    compute_misfit = wavesim[1].gradparams.compute_misfit
    # engines first, one after the other: a failure here (no device, out of memory) is thrown before any rank enters a collective
    handles = [engine(w) for w in wavesim[1:ndev]]
    # phase 0: the communicator (ncclCommInitRank is itself a collective: every rank enters it, nothing else can fail before it)
    comms = fill(C_NULL, ndev)
    errors = Vector{Any}(nothing, ndev)
    if ndev > 1
        tasks = map(1:ndev) do r
            Threads.@spawn try
                comm = Ref{Ptr{Cvoid}}(C_NULL)
                check(ccall((:swb_comm_create, lib), Int32, (Ptr{UInt8}, Int32, Int32, Int32, Ref{Ptr{Cvoid}}), id, ndev, r - 1, device_of(wavesim[r]), comm))
                comms[r] = comm[]
            catch err
                errors[r] = err
            end
        end
        foreach(wait, tasks)
    end
    # phase 1 (no collective inside): every rank computes its shots; an exception is kept, not thrown, so that no rank is left
    # alone in a collective
    tasks = map(1:ndev) do r
        Threads.@spawn try
            errors[r] === nothing || return
            model, h = wavesim[r], handles[r]
            upload_material!(model)
            check(ccall((:swb_sim_zero_total_gradient, lib), Int32, (Ptr{Cvoid},), h))
            for s in grpshots[r]
                SeismicWaves.init_shot!(model, shots[s])   # check_shot + init_bdc!: the C-PML profiles of THIS shot ...
                upload_cpml!(model)                        # ... go to the engine before it runs
                bind!(model, shots[s])
                run_gradient!(model, shots[s], misfit[s])
                accumulate!(model, shots[s])               # muting is per shot: it precedes the sum
                compute_misfit && (misfitvals[s] = SeismicWaves.calcmisfit(misfit[s], shots[s].recs))
            end
        catch err
            errors[r] = err
        end
    end
    foreach(wait, tasks)
    failed = findfirst(!isnothing, errors)
    # phase 2: the all-reduce, entered by every rank or by none
    if ndev > 1 && failed === nothing && all(!=(C_NULL), comms)
        tasks = map(1:ndev) do r
            Threads.@spawn check(ccall((:swb_sim_allreduce_total_gradient, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), engine(wavesim[r]), comms[r]))
        end
        foreach(wait, tasks)
    end
    for c in comms
        c != C_NULL && ccall((:swb_comm_destroy, lib), Int32, (Ptr{Cvoid},), c)
    end
    failed === nothing || throw(errors[failed])
    totgrad = download_total(wavesim[1], T, Val(N))
    return compute_misfit ? (totgrad, sum(misfitvals)) : totgrad
end

# ---------------------------------------------------------------------------------------------------------------------
# z-slab decomposition of one 3D acoustic CD forward run (include/swb200.h section 4; no counterpart in the reference).
# `model` is the slab-local simulation (owned planes + one ghost plane per interior face), `comm` a swb_comm handle created with
# swb_comm_create; lower / upper = neighbour ranks, -1 at a true domain end.  Call before swforward_1shot!.
# ---------------------------------------------------------------------------------------------------------------------
function set_slab!(model::AcousticCDCPMLWaveSimulation{T, 3, <:B200Array}, comm::Ptr{Cvoid}, lower::Integer, upper::Integer) where {T}
    check(ccall((:swb_sim_set_slab, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32), engine(model), comm, Int32(lower), Int32(upper)))
    return model
end

# The same decomposition with the halo exchange through peer memory (no communicator): `slabs[k]` is the slab-local simulation of
# slab k, bottom to top, each on its own GPU (set_device!).  After this call the step kernels store the boundary planes straight into
# the neighbours' ghost planes over NVLink and per-step flags keep the slabs in lockstep; run every slab's swforward_1shot! from its
# own task (Threads.@spawn): a slab's call blocks until its neighbours have caught up.
function connect_slabs!(slabs::Vector{<:AcousticCDCPMLWaveSimulation{T, 3, <:B200Array}}) where {T}
    n = length(slabs)
    for (k, m) in enumerate(slabs)
        set_slab!(m, C_NULL, k > 1 ? k - 2 : -1, k < n ? k : -1)
    end
    handles = map(slabs) do m
        h = Ref{SlabHandle}()
        check(ccall((:swb_sim_slab_export, lib), Int32, (Ptr{Cvoid}, Ref{SlabHandle}), engine(m), h))
        h
    end
    for (k, m) in enumerate(slabs)
        lo = k > 1 ? Base.unsafe_convert(Ptr{SlabHandle}, handles[k - 1]) : Ptr{SlabHandle}(C_NULL)
        hi = k < n ? Base.unsafe_convert(Ptr{SlabHandle}, handles[k + 1]) : Ptr{SlabHandle}(C_NULL)
        GC.@preserve handles check(ccall((:swb_sim_slab_connect, lib), Int32, (Ptr{Cvoid}, Ptr{SlabHandle}, Ptr{SlabHandle}), engine(m), lo, hi))
    end
    return slabs
end

end # module
