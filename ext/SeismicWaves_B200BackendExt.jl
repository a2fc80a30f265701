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
# There is no CPU fallback: every call raises if the library reports an error (e.g. no sm_100 device).
#
# NOTE: Julia is not available in the build environment of this repository, so this file is exercised only through its
# Python twin (seismicwaves.jl_b200/, same C ABI, same call sequence); keep the two in sync.
module SeismicWaves_B200BackendExt

using SeismicWaves, SeismicWavesB200
using SeismicWaves: CPMLBoundaryCondition, LocalGrid, AcousticCDCPMLWaveSimulation, AcousticVDStaggeredCPMLWaveSimulation,
                    ElasticIsoCPMLWaveSimulation, ScalarShot, MomentTensorShot, ExternalForceShot, AbstractMisfit, L2Misfit

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
# B200Array{T,N}: device array owned by libswb200 (plays the role of backend.Data.Array)
# ---------------------------------------------------------------------------------------------------------------------
mutable struct B200Array{T, N} <: DenseArray{T, N}
    ptr::Ptr{Cvoid}
    dims::NTuple{N, Int}
    function B200Array{T, N}(dims::NTuple{N, Int}) where {T, N}
        p = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:swb_malloc, lib), Int32, (Ref{Ptr{Cvoid}}, Csize_t), p, max(prod(dims), 1) * sizeof(T)))   # zero-filled
        a = new{T, N}(p[], dims)
        finalizer(x -> ccall((:swb_free, lib), Int32, (Ptr{Cvoid},), x.ptr), a)
        return a
    end
end
Base.size(a::B200Array) = a.dims
Base.eltype(::B200Array{T}) where {T} = T
Base.unsafe_convert(::Type{Ptr{Cvoid}}, a::B200Array) = a.ptr
Base.getindex(::B200Array, ::Int...) = error("scalar indexing of a B200Array is not supported; copy it to the host with Array(a)")
function B200Array(h::Array{T, N}) where {T, N}                       # Data.Array(hostarray): H2D copy
    a = B200Array{T, N}(size(h))
    copyto!(a, h)
    return a
end
function Base.copyto!(d::B200Array{T}, h::Array{T}) where {T}         # copyto!(dev, host)
    @assert length(d) == length(h)
    check(ccall((:swb_memcpy_h2d, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), d.ptr, h, sizeof(h)))
    return d
end
function Base.copyto!(h::Array{T}, d::B200Array{T}) where {T}         # copyto!(host, dev)
    @assert length(d) == length(h)
    check(ccall((:swb_memcpy_d2h, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), h, d.ptr, sizeof(h)))
    return h
end
function Base.copyto!(d::B200Array{T}, s::B200Array{T}) where {T}     # copyto!(dev, dev)  (fields.jl:20,32-36)
    @assert length(d) == length(s)
    check(ccall((:swb_memcpy_d2d, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), d.ptr, s.ptr, length(s) * sizeof(T)))
    return d
end
Base.Array(d::B200Array{T, N}) where {T, N} = copyto!(Array{T, N}(undef, size(d)), d)
Base.zero(d::B200Array{T, N}) where {T, N} = B200Array{T, N}(size(d))
Base.copy(d::B200Array{T, N}) where {T, N} = copyto!(B200Array{T, N}(size(d)), d)
Base.convert(::Type{Array}, d::B200Array) = Array(d)
function Base.fill!(d::B200Array{T}, v) where {T}                     # dev .= scalar (fields.jl:40-44)
    check(ccall((:swb_fill, lib), Int32, (Ptr{Cvoid}, Int32, Cdouble, Csize_t, Csize_t), d.ptr, dtype_code(T), Float64(v), 0, length(d)))
    return d
end
# a[1:n] .= v on a vector (free-surface override of the C-PML coefficients, acou_init_bc.jl:35-38)
function fillrange!(d::B200Array{T, 1}, r::UnitRange{Int}, v) where {T}
    check(ccall((:swb_fill, lib), Int32, (Ptr{Cvoid}, Int32, Cdouble, Csize_t, Csize_t), d.ptr, dtype_code(T), Float64(v), first(r) - 1, length(r)))
    return d
end
Base.Broadcast.materialize!(dest::SubArray{T, 1, <:B200Array{T, 1}, Tuple{UnitRange{Int}}}, bc::Base.Broadcast.Broadcasted{<:Any, <:Any, typeof(identity), <:Tuple{Number}}) where {T} =
    fillrange!(parent(dest), dest.indices[1], bc.args[1])
Base.Broadcast.materialize!(dest::B200Array, bc::Base.Broadcast.Broadcasted{<:Any, <:Any, typeof(identity), <:Tuple{Number}}) = fill!(dest, bc.args[1])

# ---------------------------------------------------------------------------------------------------------------------
# C structs of include/swb200.h (layout must match field for field)
# ---------------------------------------------------------------------------------------------------------------------
struct CpmlAxis
    a::Ptr{Cvoid}; a_h::Ptr{Cvoid}; b::Ptr{Cvoid}; b_h::Ptr{Cvoid}
end
struct Points
    n::Int64; pos::Ptr{Cvoid}; tf::Ptr{Cvoid}; nt::Int64
end
const NOPOINTS = Points(0, C_NULL, C_NULL, 0)
struct AcouCDStepArgs
    dtype::Int32; ndim::Int32; halo::Int32; flags::Int32
    n::NTuple{3, Int64}; spacing::NTuple{3, Float64}
    pold::Ptr{Cvoid}; pcur::Ptr{Cvoid}; pnew::Ptr{Cvoid}; fact::Ptr{Cvoid}
    psi::NTuple{3, Ptr{Cvoid}}; xi::NTuple{3, Ptr{Cvoid}}
    cpml::NTuple{3, CpmlAxis}
    src::Points; rec::Points
    it::Int64; stream::Ptr{Cvoid}
end
struct AcouVDStepArgs
    dtype::Int32; halo::Int32; flags::Int32; _pad::Int32
    n::NTuple{2, Int64}; spacing::NTuple{2, Float64}
    pcur::Ptr{Cvoid}; vcur::NTuple{2, Ptr{Cvoid}}
    fact_m0::Ptr{Cvoid}; fact_m1_stag::NTuple{2, Ptr{Cvoid}}
    psi::NTuple{2, Ptr{Cvoid}}; xi::NTuple{2, Ptr{Cvoid}}
    cpml::NTuple{2, CpmlAxis}
    src::Points; rec::Points
    it::Int64; stream::Ptr{Cvoid}
end
struct SimDesc
    kind::Int32; dtype::Int32; ndim::Int32; device::Int32
    n::NTuple{3, Int64}; spacing::NTuple{3, Float64}; dt::Float64; nt::Int64
    halo::Int32; freetop::Int32; gradient::Int32; check_freq::Int32
    flags::Int32; _pad::Int32
end

pad3(t::NTuple{N, T}, z) where {N, T} = ntuple(i -> i <= N ? t[i] : z, 3)
vptr(a::B200Array) = a.ptr
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
    return AcouCDStepArgs(dtype_code(T), N, model.cpmlparams.halo, 0,
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
        dtype_code(T), N, n, res.ptr, size(res, 1), size(res, 2), posrecs.ptr, fact.ptr, C_NULL))
end
function correlate_gradient!(grad::B200Array{T}, adjcur, p_itm2, p_itm1, p_it, dt) where {T}
    check(ccall((:swb_acou_cd_correlate_gradient, lib), Int32, (Int32, Int32, Csize_t, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
        dtype_code(T), 0, length(grad), grad.ptr, adjcur.ptr, p_itm2.ptr, p_itm1.ptr, p_it.ptr, Float64(dt), C_NULL))
end

function acou_vd_args(model, possrcs, srctf, posrecs, traces, it, pname, vname, ψname, ξname)
    g = model.grid
    T = eltype(g.fields["fact_m0"].value)
    v, m1, ψ, ξ = g.fields[vname].value, g.fields["fact_m1_stag"].value, g.fields[ψname].value, g.fields[ξname].value
    src = Points(size(possrcs, 1), vptr(possrcs), vptr(srctf), size(srctf, 1))
    rec = traces === nothing ? NOPOINTS : Points(size(posrecs, 1), vptr(posrecs), vptr(traces), size(traces, 1))
    return AcouVDStepArgs(dtype_code(T), model.cpmlparams.halo, 0, 0, Tuple(Int64.(g.size)), Tuple(Float64.(g.spacing)),
        vptr(g.fields[pname].value), Tuple(vptr.(v)), vptr(g.fields["fact_m0"].value), Tuple(vptr.(m1)), Tuple(vptr.(ψ)), Tuple(vptr.(ξ)),
        Tuple(cpmlaxis.(model.cpmlcoeffs)), src, rec, it, C_NULL)
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
        dtype_code(T), 0, length(grad_m0), grad_m0.ptr, adjp.ptr, p_it.ptr, p_itm1.ptr, Float64(dt), C_NULL))
end
function correlate_gradient_m1!(grad_m1_stag, adjv, p_it::B200Array{T, 2}, spacing) where {T}
    n, sp = collect(Int64, size(p_it)), collect(Float64, spacing)
    g, av = [x.ptr for x in grad_m1_stag], [x.ptr for x in adjv]
    check(ccall((:swb_acou_vd_correlate_gradient_m1, lib), Int32, (Int32, Int32, Ptr{Int64}, Ptr{Cdouble}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{Cvoid}),
        dtype_code(T), 0, n, sp, g, av, p_it.ptr, C_NULL))
end

# the backend "modules": NamedTuples carry the same member names the reference looks up with backend.<name>
const Acoustic2D_CD_CPML_B200 = (; Data, zeros, ones, forward_onestep_CPML! = cd_forward_onestep_CPML!, adjoint_onestep_CPML! = cd_adjoint_onestep_CPML!,
                                  prescale_residuals!, correlate_gradient!)
const Acoustic3D_CD_CPML_B200 = Acoustic2D_CD_CPML_B200          # the C entry points take ndim
const Acoustic2D_VD_CPML_B200 = (; Data, zeros, ones, forward_onestep_CPML! = vd_forward_onestep_CPML!, adjoint_onestep_CPML! = vd_adjoint_onestep_CPML!,
                                  prescale_residuals!, correlate_gradient_m0!, correlate_gradient_m1!)
# The elastic fine-grained members (swb_ela_forward_onestep / swb_ela_adjoint_onestep / swb_ela_correlate_gradients) follow the same
# pattern with `swb_ela_step_args`; the elastic simulation is normally driven through the per-shot engine below.
const Elastic2D_Iso_CPML_B200 = (; Data, zeros, ones)

const FT = Union{Float32, Float64}
SeismicWaves.select_backend(::CPMLBoundaryCondition, ::LocalGrid, ::Type{<:AcousticCDCPMLWaveSimulation{<:FT, 2}}, ::Type{Val{:B200}}) = Acoustic2D_CD_CPML_B200
SeismicWaves.select_backend(::CPMLBoundaryCondition, ::LocalGrid, ::Type{<:AcousticCDCPMLWaveSimulation{<:FT, 3}}, ::Type{Val{:B200}}) = Acoustic3D_CD_CPML_B200
SeismicWaves.select_backend(::CPMLBoundaryCondition, ::LocalGrid, ::Type{<:AcousticVDStaggeredCPMLWaveSimulation{<:FT, 2}}, ::Type{Val{:B200}}) = Acoustic2D_VD_CPML_B200
SeismicWaves.select_backend(::CPMLBoundaryCondition, ::LocalGrid, ::Type{<:ElasticIsoCPMLWaveSimulation{<:FT, 2}}, ::Type{Val{:B200}}) = Elastic2D_Iso_CPML_B200

# ---------------------------------------------------------------------------------------------------------------------
# Level 2: whole shots on the per-shot engine (performance path)
# ---------------------------------------------------------------------------------------------------------------------
const ENGINES = IdDict{Any, Ptr{Cvoid}}()     # wavesim object -> swb_sim*
# B200-only knobs the reference's RunParameters has no field for: the GPU a simulation lives on (default 0) and the creation flags
# (SWB_FLAG_FAST_F32 = 1: Float32 arithmetic; SWB_FLAG_NO_FUSION = 2; SWB_FLAG_NO_GRAPH = 4), set before the first shot
const DEVICE_OF = IdDict{Any, Int32}()
const FLAGS = Ref{Int32}(0)
set_device!(model, dev::Integer) = (DEVICE_OF[model] = Int32(dev); model)
device_of(model) = get(DEVICE_OF, model, Int32(0))
set_flags!(flags::Integer) = (FLAGS[] = Int32(flags))

simkind(::AcousticCDCPMLWaveSimulation) = Int32(1)
simkind(::AcousticVDStaggeredCPMLWaveSimulation) = Int32(2)
simkind(::ElasticIsoCPMLWaveSimulation) = Int32(3)

function engine(model)
    get!(ENGINES, model) do
        T = typeof(model.dt)
        N = length(model.grid.size)
        cf = model.gradparams === nothing ? 1 : model.gradparams.check_freq
        desc = SimDesc(simkind(model), dtype_code(T), N, device_of(model), pad3(Int64.(model.grid.size), Int64(1)), pad3(Float64.(model.grid.spacing), 0.0), Float64(model.dt), model.nt,
            model.cpmlparams.halo, model.cpmlparams.freeboundtop, model.checkpointer !== nothing, cf, FLAGS[], Int32(0))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:swb_sim_create, lib), Int32, (Ref{SimDesc}, Ref{Ptr{Cvoid}}), desc, h))
        finalizer(_ -> ccall((:swb_sim_destroy, lib), Int32, (Ptr{Cvoid},), h[]), model)
        h[]
    end
end

matfields(m::AcousticCDCPMLWaveSimulation) = (m.matprop.vp,)
matfields(m::AcousticVDStaggeredCPMLWaveSimulation) = (m.matprop.vp, m.matprop.rho)
matfields(m::ElasticIsoCPMLWaveSimulation) = (m.matprop.ρ, m.matprop.λ, m.matprop.μ)

function upload_model!(model)      # update_matprop! + precompute_fact! / precomp_elaprop! on the device
    fs = matfields(model)
    ptrs = [Ptr{Cvoid}(pointer(f)) for f in fs]
    GC.@preserve fs check(ccall((:swb_sim_set_material, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}, Int32), engine(model), length(fs), ptrs, 0))
    for (ax, c) in enumerate(model.cpmlcoeffs)     # host-computed by init_bdc! on plain Arrays, uploaded per axis
        check(ccall((:swb_sim_set_cpml, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), engine(model), ax - 1,
            Array(c.a), Array(c.a_h), Array(c.b), Array(c.b_h)))
    end
end

function bind!(model::Union{AcousticCDCPMLWaveSimulation, AcousticVDStaggeredCPMLWaveSimulation}, shot::ScalarShot)
    possrcs, posrecs, scal_srctf = SeismicWaves.possrcrec_scaletf(model, shot)     # acou_forward.jl:6-20,67-81 (host)
    check(ccall((:swb_sim_bind_scalar_shot, lib), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Cvoid}, Int64, Ptr{Int64}), engine(model),
        size(possrcs, 1), possrcs, scal_srctf, size(posrecs, 1), posrecs))
end

# swforward_1shot! (acou_forward.jl:22-125): the whole time loop in one call
function SeismicWaves.swforward_1shot!(::CPMLBoundaryCondition, model::Union{AcousticCDCPMLWaveSimulation{T, N, <:B200Array}, AcousticVDStaggeredCPMLWaveSimulation{T, N, <:B200Array}},
                                       shot::ScalarShot{T, N}) where {T, N}
    upload_model!(model)
    bind!(model, shot)
    snapevery = model.runparams.snapevery === nothing ? 0 : model.runparams.snapevery
    check(ccall((:swb_sim_forward, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32), engine(model), shot.recs.seismograms, snapevery))
    return nothing
end

# swgradient_1shot! (acou_gradient.jl:4-203), split where the reference hands control to the pluggable misfit
function SeismicWaves.swgradient_1shot!(::CPMLBoundaryCondition, model::Union{AcousticCDCPMLWaveSimulation{T, N, <:B200Array}, AcousticVDStaggeredCPMLWaveSimulation{T, N, <:B200Array}},
                                        shot::ScalarShot{T, N}, misfit::AbstractMisfit{T}) where {T, N}
    h = engine(model)
    upload_model!(model)
    bind!(model, shot)
    check(ccall((:swb_sim_gradient_forward, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), h, shot.recs.seismograms))
    adjsrc = .-SeismicWaves.∂χ_∂u(misfit, shot.recs)                                   # any AbstractMisfit, on the host
    check(ccall((:swb_sim_gradient_adjoint, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), h, adjsrc))
    gp = model.gradparams
    check(ccall((:swb_sim_zero_total_gradient, lib), Int32, (Ptr{Cvoid},), h))
    check(ccall((:swb_sim_accumulate_gradient, lib), Int32, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int32, Int64, Ptr{Cvoid}, Int32), h,
        size(shot.srcs.positions, 1), shot.srcs.positions, gp.mute_radius_src, size(shot.recs.positions, 1), shot.recs.positions, gp.mute_radius_rec))
    names = model isa AcousticCDCPMLWaveSimulation ? ("vp",) : ("vp", "rho")
    out = Dict{String, Array{T, N}}()
    for (k, name) in enumerate(names)
        g = zeros(T, model.grid.size...)
        check(ccall((:swb_sim_get_total_gradient, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}), h, k - 1, g))
        out[name] = g
    end
    return out
end

# ---------------------------------------------------------------------------------------------------------------------
# Elastic P-SV shots through the per-shot engine (ela_forward.jl:4-159, ela_gradient.jl:4-362).  The host part of the reference
# stays as it is: `possrcrec_scaletf` (ela_models.jl:6-90) spreads the off-grid positions into Kaiser-windowed sinc point lists
# and scales the source time functions; the lists cross the ABI flattened to CSR (include/swb200.h, swb_sinc_points_host).
# ---------------------------------------------------------------------------------------------------------------------
struct SincPointsHost          # mirrors swb_sinc_points_host
    n::Int64
    off::Ptr{Int64}
    ij::Ptr{Int32}
    coef::Ptr{Cvoid}
end

# per-position lists (Vector of (npts, 2) index matrices, Vector of coefficient vectors) -> CSR; the points of one position in
# ascending linear index (the reference iterates a Dict; any fixed order is equivalent up to the summation order of a receiver)
function csr(ijs::Vector{<:AbstractMatrix{<:Integer}}, vals::Vector{<:AbstractVector{T}}, nx::Int) where {T}
    off = zeros(Int64, length(ijs) + 1)
    for k in eachindex(ijs)
        off[k + 1] = off[k] + size(ijs[k], 1)
    end
    ij = zeros(Int32, off[end], 2)
    co = zeros(T, off[end])
    for k in eachindex(ijs)
        order = sortperm([(ijs[k][p, 2] - 1) * nx + ijs[k][p, 1] for p in 1:size(ijs[k], 1)])
        ij[off[k] + 1:off[k + 1], :] .= ijs[k][order, :]
        co[off[k] + 1:off[k + 1]] .= vals[k][order]
    end
    return off, ij, co
end

function bind!(model::ElasticIsoCPMLWaveSimulation{T, 2}, shot::Union{MomentTensorShot{T, 2}, ExternalForceShot{T, 2}}) where {T}
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

function SeismicWaves.swforward_1shot!(::CPMLBoundaryCondition, model::ElasticIsoCPMLWaveSimulation{T, 2, <:B200Array},
                                       shot::Union{MomentTensorShot{T, 2}, ExternalForceShot{T, 2}}) where {T}
    upload_model!(model)
    bind!(model, shot)
    snapevery = model.runparams.snapevery === nothing ? 0 : model.runparams.snapevery
    check(ccall((:swb_sim_forward, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32), engine(model), shot.recs.seismograms, snapevery))   # (nt, 2, nrec)
    return nothing
end

function SeismicWaves.swgradient_1shot!(::CPMLBoundaryCondition, model::ElasticIsoCPMLWaveSimulation{T, 2, <:B200Array},
                                        shot::Union{MomentTensorShot{T, 2}, ExternalForceShot{T, 2}}, misfit::AbstractMisfit{T}) where {T}
    h = engine(model)
    upload_model!(model)
    bind!(model, shot)
    check(ccall((:swb_sim_gradient_forward, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), h, shot.recs.seismograms))
    adjsrc = .-SeismicWaves.∂χ_∂u(misfit, shot.recs)                                   # (nt, 2, nrec), any AbstractMisfit, on the host
    check(ccall((:swb_sim_gradient_adjoint, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), h, adjsrc))
    gp = model.gradparams
    check(ccall((:swb_sim_zero_total_gradient, lib), Int32, (Ptr{Cvoid},), h))
    # back_interp of the staggered accumulators, mutearoundmultiplepoints!, on the device (ela_gradient.jl:155-190)
    check(ccall((:swb_sim_accumulate_gradient, lib), Int32, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int32, Int64, Ptr{Cvoid}, Int32), h,
        size(shot.srcs.positions, 1), shot.srcs.positions, gp.mute_radius_src, size(shot.recs.positions, 1), shot.recs.positions, gp.mute_radius_rec))
    out = Dict{String, Array{T, 2}}()
    for (k, name) in enumerate(("rho", "lambda", "mu"))
        g = zeros(T, model.grid.size...)
        check(ccall((:swb_sim_get_total_gradient, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}), h, k - 1, g))
        out[name] = g
    end
    return out
end

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
    ndev, nshots = length(wavesim), length(shots)
    @assert Threads.nthreads() >= ndev "one Julia thread per GPU is needed (NCCL ranks enter collectives concurrently)"
    for w in wavesim
        SeismicWaves.check_sim_consistency(w, matprop, shots)
        SeismicWaves.set_wavesim_matprop!(w, matprop)
    end
    id = zeros(UInt8, 128)
    ndev > 1 && check(ccall((:swb_comm_unique_id, lib), Int32, (Ptr{UInt8},), id))
    grpshots = SeismicWaves.distribsrcs(nshots, ndev)
    misfitvals = zeros(T, nshots)
    compute_misfit = wavesim[1].gradparams.compute_misfit
    names = wavesim[1] isa AcousticCDCPMLWaveSimulation ? ("vp",) : wavesim[1] isa AcousticVDStaggeredCPMLWaveSimulation ? ("vp", "rho") : ("rho", "lambda", "mu")
    tasks = map(1:ndev) do r
        Threads.@spawn begin
            model, h = wavesim[r], engine(wavesim[r])
            comm = Ref{Ptr{Cvoid}}(C_NULL)
            ndev > 1 && check(ccall((:swb_comm_create, lib), Int32, (Ptr{UInt8}, Int32, Int32, Int32, Ref{Ptr{Cvoid}}), id, ndev, r - 1, device_of(model), comm))
            upload_model!(model)
            check(ccall((:swb_sim_zero_total_gradient, lib), Int32, (Ptr{Cvoid},), h))
            gp = model.gradparams
            for s in grpshots[r]
                SeismicWaves.init_shot!(model, shots[s])
                bind!(model, shots[s])
                check(ccall((:swb_sim_gradient_forward, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), h, shots[s].recs.seismograms))
                adjsrc = .-SeismicWaves.∂χ_∂u(misfit[s], shots[s].recs)
                check(ccall((:swb_sim_gradient_adjoint, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), h, adjsrc))
                # mute + chain rule of THIS shot, then accumulation into the device-resident total (muting is per shot: it precedes the sum)
                check(ccall((:swb_sim_accumulate_gradient, lib), Int32, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int32, Int64, Ptr{Cvoid}, Int32), h,
                    size(shots[s].srcs.positions, 1), shots[s].srcs.positions, gp.mute_radius_src, size(shots[s].recs.positions, 1), shots[s].recs.positions, gp.mute_radius_rec))
                compute_misfit && (misfitvals[s] = SeismicWaves.calcmisfit(misfit[s], shots[s].recs))
            end
            if ndev > 1
                check(ccall((:swb_sim_allreduce_total_gradient, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), h, comm[]))
                check(ccall((:swb_comm_destroy, lib), Int32, (Ptr{Cvoid},), comm[]))
            end
        end
    end
    foreach(wait, tasks)
    totgrad = Dict{String, Array{T, N}}()
    for (k, name) in enumerate(names)
        g = zeros(T, wavesim[1].grid.size...)
        check(ccall((:swb_sim_get_total_gradient, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}), engine(wavesim[1]), k - 1, g))
        totgrad[name] = g
    end
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

end # module
