# baseline/run_reference.jl -- times (and optionally dumps) the UNMODIFIED reference, SeismicWaves.jl's `parall = :threads`
# backend (src/models/acoustic/backends/Acoustic2D_CD_CPML_Threads.jl:1-15 and siblings), on a problem that bench.py /
# baseline/pin_oracle.py wrote to a directory.  Julia is not part of the build image (SURVEY.md 8c), so bench.py runs this only
# when `julia` is on PATH and SeismicWaves.jl is loadable (JULIA_PROJECT / JULIA_LOAD_PATH, or baseline/_ref/SeismicWaves.jl);
# otherwise its CPU legs time the OpenMP build of the oracle and say kind = "port".
#
#   julia --threads=auto baseline/run_reference.jl <dir> [time|dump] [warmup] [steps]
#
# <dir>/meta.txt   "key value" lines: kind (acoustic_cd | acoustic_vd), T (Float32 | Float64), ndim, n1 n2 [n3], h, dt, nt, halo,
#                  rcoef, freetop (0|1), check_freq, mute_src, mute_rec, domfreq, nsrc, nrec
# <dir>/vp.bin [rho.bin]          column-major arrays of T, grid-sized
# <dir>/srcpos.bin, recpos.bin    (nsrc, ndim), (nrec, ndim) of T, metres
# <dir>/srctf.bin                 (nt, nsrc) of T
# <dir>/observed.bin              (nt, nrec) of T
# time: prints {"impl": "reference", "seconds_per_step": ..., "threads": ...}; dump: also writes seis.bin, grad_<name>.bin, misfit.txt
using LinearAlgebra
using Logging

let ref = joinpath(@__DIR__, "_ref", "SeismicWaves.jl")
    isdir(ref) && pushfirst!(LOAD_PATH, ref)
end
using SeismicWaves

function readmeta(dir)
    meta = Dict{String, String}()
    for ln in eachline(joinpath(dir, "meta.txt"))
        s = split(strip(ln))
        length(s) >= 2 && (meta[s[1]] = join(s[2:end], " "))
    end
    return meta
end

readarr(dir, name, ::Type{T}, dims...) where {T} = reshape(reinterpret(T, read(joinpath(dir, name))), dims...) |> collect

function main()
    dir = ARGS[1]
    mode = length(ARGS) >= 2 ? ARGS[2] : "time"
    warmup = length(ARGS) >= 3 ? parse(Int, ARGS[3]) : 1
    steps = length(ARGS) >= 4 ? parse(Int, ARGS[4]) : 2
    m = readmeta(dir)
    T = m["T"] == "Float32" ? Float32 : Float64
    ndim = parse(Int, m["ndim"])
    n = ntuple(d -> parse(Int, m["n$d"]), ndim)
    h, dt = parse(T, m["h"]), parse(T, m["dt"])
    nt, halo = parse(Int, m["nt"]), parse(Int, m["halo"])
    nsrc, nrec = parse(Int, m["nsrc"]), parse(Int, m["nrec"])
    vp = readarr(dir, "vp.bin", T, n...)
    matprop = m["kind"] == "acoustic_vd" ? VpRhoAcousticVDMaterialProperties(vp, readarr(dir, "rho.bin", T, n...)) : VpAcousticCDMaterialProperties(vp)
    srcs = ScalarSources(readarr(dir, "srcpos.bin", T, nsrc, ndim), readarr(dir, "srctf.bin", T, nt, nsrc), parse(T, m["domfreq"]))
    mkshot() = ScalarShot(; srcs=srcs, recs=ScalarReceivers(readarr(dir, "recpos.bin", T, nrec, ndim), nt))
    boundcond = CPMLBoundaryConditionParameters(; halo=halo, rcoef=parse(T, m["rcoef"]), freeboundtop=m["freetop"] == "1")
    params = InputParametersAcoustic(nt, dt, n, ntuple(_ -> h, ndim), boundcond)
    runparams = RunParameters(; parall=:threads, logger=ConsoleLogger(stderr, Logging.Warn), erroronPPW=false)
    gradparams = GradParameters(; mute_radius_src=parse(Int, m["mute_src"]), mute_radius_rec=parse(Int, m["mute_rec"]), compute_misfit=true,
                                check_freq=parse(Int, m["check_freq"]))
    observed = readarr(dir, "observed.bin", T, nt, nrec)
    mkmisfit() = [SeismicWaves.L2Misfit(; observed=observed, invcov=Diagonal(ones(T, nt)))]
    wavesim = build_wavesim(params, matprop; runparams=runparams, gradparams=gradparams, gradient=true)
    local grad, mis, shots
    for _ in 1:warmup
        shots = [mkshot()]
        grad, mis = swgradient!(wavesim, matprop, shots, mkmisfit())
    end
    t0 = time_ns()
    for _ in 1:steps
        shots = [mkshot()]
        grad, mis = swgradient!(wavesim, matprop, shots, mkmisfit())
    end
    secs = (time_ns() - t0) / 1e9 / max(steps, 1)
    if mode == "dump"
        write(joinpath(dir, "seis.bin"), shots[1].recs.seismograms)
        for (k, g) in grad
            write(joinpath(dir, "grad_$(k).bin"), g)
        end
        write(joinpath(dir, "misfit.txt"), string(Float64(mis)))
    end
    println("{\"impl\": \"reference\", \"seconds_per_step\": $(secs), \"threads\": $(Threads.nthreads()), \"misfit\": $(Float64(mis))}")
end

main()
