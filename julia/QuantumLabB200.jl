# QuantumLabB200.jl -- thin Julia host side of libqlb200.so (include/qlb200.h).
#
# STATUS: never executed (no Julia in the build image); every name used below was checked by hand against the export /
# import lists of the reference modules cited next to it.
#
# This is the binding a QuantumLab.jl maintainer adds next to LibInt2Module.jl (the reference's existing FFI seam,
# src/SubModules/LibInt2Module.jl:30-40, 62-110, 193-334): a third shell-vector type, `B200Shells`, on which the
# hot-path methods dispatch.  It is deliberately thin -- every line maps 1:1 onto the ctypes harness in
# quantumlab_jl_b200/{libqlb,b200shells}.py, which IS executed by the test-suite (no Julia exists in the build image,
# so this file cannot be run there; see INTEGRATION.md).
module QuantumLabB200

using Libdl
using LinearAlgebra
using ..BaseModule        # Position, Ionization, LQuantumNumber (exported, BaseModule.jl:2)
using ..ShellModule       # AbstractShell, Shell, expandShell (ShellModule.jl:2)
using ..GeometryModule    # Geometry, computeEnergyInteratomicRepulsion, computeNumberElectrons (GeometryModule.jl:2)
using ..BasisSetModule    # BasisSet (BasisSetModule.jl:3)
using ..BasisModule       # computeBasisShells (BasisModule.jl:2)
using ..InitialGuessModule  # InitialGuess, ZeroGuess, SADGuess (InitialGuessModule.jl:2)
import ..ShellModule: nbf                                         # not exported by ShellModule (ShellModule.jl:2, :134)
import ..InitialGuessModule: computeDensityMatrixGuess            # not exported (InitialGuessModule.jl:11-16)
# the methods below EXTEND the reference's generic functions (new dispatch on B200Shells / B200RI), they do not shadow them
import ..SpecialMatricesModule: computeMatrixCoulomb, computeMatrixExchange, computeMatrixFock,
                                computeTensorElectronRepulsionIntegrals,
                                computeMatrixOverlap, computeMatrixKinetic, computeMatrixNuclearAttraction
import ..IntegralsModule: computeTensorBlockElectronRepulsionIntegrals, computeElectronRepulsionIntegral
import ..ResolutionIdentityModule: computeMatrixResolutionOfTheIdentityCoulombMetric, computeMatrixExchangeRIK,
                                   computeTensorElectronRepulsionIntegralsRICoulomb
import ..HartreeFockModule: evaluateSCF

export B200Shells, B200RI, destroy!, buildJK

const libqlb = Ref{Ptr{Nothing}}(C_NULL)

function __init__()
  # same pattern as LibInt2Module.__init__ (LibInt2Module.jl:30-40); unlike libint2 a load failure is an ERROR:
  # there is no stub / CPU fallback for this backend.
  push!(Libdl.DL_LOAD_PATH, joinpath(@__DIR__, "..", "deps", "usr", "lib"))
  libqlb[] = Libdl.dlopen("libqlb200")
  check(ccall(dlsym(libqlb[], :qlb_init), Cint, (Cint,), -1))
end

lasterror() = unsafe_string(ccall(dlsym(libqlb[], :qlb_last_error), Cstring, ()))
check(rc::Integer) = rc == 0 ? nothing : error("libqlb200 error $rc: $(lasterror())")

"Per-build statistics, layout of `struct qlb_stats` (include/qlb200.h)."
mutable struct QlbStats
  quartets_unique::Int64; quartets_kept::Int64
  quartets_kept_class::NTuple{21,Int64}; prim_quartets_class::NTuple{21,Int64}
  pairs_total::Int64; pairs_significant::Int64
  ms_total::Float64; ms_eri::Float64; ms_class::NTuple{21,Float64}
  flops_algorithmic::Float64; kernel_launches::Int32; reserved::Int32
  QlbStats() = new(0, 0, ntuple(_ -> 0, 21), ntuple(_ -> 0, 21), 0, 0, 0.0, 0.0, ntuple(_ -> 0.0, 21), 0.0, 0, 0)
end

"""
    B200Shells(shells::Vector{Shell}; tau=1e-12)
Device-resident shell list: the analogue of `Vector{LibInt2Shell}` + `LibInt2EngineCoulomb` (LibInt2Module.jl:62-68,
193-204).  Owns an opaque `qlb_basis*`; freed by `destroy!` (LibInt2Module.jl:108-110) or the finalizer.
Coefficients are passed as the Shell constructor leaves them (ShellModule.jl:25-48).
"""
mutable struct B200Shells <: AbstractVector{Shell}
  shells::Vector{Shell}
  handle::Ptr{Nothing}
  nbf::Int
  tau::Float64
  cacheP::Union{Nothing,Matrix{Float64}}
  cacheJ::Matrix{Float64}
  cacheK::Matrix{Float64}
  function B200Shells(shells::Vector{Shell}; tau::Float64=1e-12)
    centers = Float64[c for sh in shells for c in (sh.center.x, sh.center.y, sh.center.z)]
    L       = Int32[sh.lqn.exponent for sh in shells]
    nprim   = Int32[length(sh.exponents) for sh in shells]
    exps    = vcat([sh.exponents for sh in shells]...)
    coefs   = vcat([sh.coefficients for sh in shells]...)
    h = Ref{Ptr{Nothing}}(C_NULL)
    check(ccall(dlsym(libqlb[], :qlb_basis_create), Cint,
                (Cint, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Ptr{Nothing}}),
                length(shells), centers, L, nprim, exps, coefs, h))
    n = mapreduce(nbf, +, shells)
    obj = new(shells, h[], n, tau, nothing, zeros(0, 0), zeros(0, 0))
    finalizer(destroy!, obj)
    obj
  end
end
B200Shells(basSet::BasisSet, geo::Geometry; kw...) = B200Shells(computeBasisShells(basSet, geo); kw...)   # BasisModule.jl:72-83
Base.size(b::B200Shells) = size(b.shells)
Base.getindex(b::B200Shells, i::Int) = b.shells[i]

function destroy!(b::B200Shells)
  if b.handle != C_NULL
    ccall(dlsym(libqlb[], :qlb_basis_destroy), Cint, (Ptr{Nothing},), b.handle)
    b.handle = C_NULL
  end
end

"One fused, screened pass: (J, K) as SpecialMatricesModule.jl:33-100 defines them."
function buildJK(b::B200Shells, P::Matrix{Float64}; stats::Union{Nothing,QlbStats}=nothing)
  size(P) == (b.nbf, b.nbf) || error("density matrix must be $(b.nbf)x$(b.nbf)")
  J = Matrix{Float64}(undef, b.nbf, b.nbf); K = similar(J)
  st = stats === nothing ? C_NULL : pointer_from_objref(stats)
  check(ccall(dlsym(libqlb[], :qlb_build_jk), Cint,
              (Ptr{Nothing}, Ptr{Float64}, Cdouble, Ptr{Float64}, Ptr{Float64}, Ptr{Nothing}),
              b.handle, P, b.tau, J, K, st))
  J, K
end

# `coulomb(P)` followed by `exchange(P)` (HartreeFockModule.jl:123-131) costs ONE device pass: same P => cached
function jkcached(b::B200Shells, P::Matrix{Float64})
  if b.cacheP === nothing || b.cacheP != P
    b.cacheJ, b.cacheK = buildJK(b, P)
    b.cacheP = copy(P)
  end
  b.cacheJ, b.cacheK
end

computeMatrixCoulomb(b::B200Shells, density::Matrix)  = copy(jkcached(b, Matrix{Float64}(density))[1])   # :231-244
computeMatrixExchange(b::B200Shells, density::Matrix) = copy(jkcached(b, Matrix{Float64}(density))[2])   # :268-281

function onee(b::B200Shells, sym::Symbol)
  M = Matrix{Float64}(undef, b.nbf, b.nbf)
  check(ccall(dlsym(libqlb[], sym), Cint, (Ptr{Nothing}, Ptr{Float64}), b.handle, M))
  M
end
computeMatrixOverlap(b::B200Shells) = onee(b, :qlb_overlap)                                              # :102-125
computeMatrixKinetic(b::B200Shells) = onee(b, :qlb_kinetic)                                              # :127-150
function computeMatrixNuclearAttraction(b::B200Shells, geo::Geometry)                                    # :152-177
  xyz = Float64[c for a in geo.atoms for c in (a.position.x, a.position.y, a.position.z)]
  Z   = Float64[a.element.atomicNumber for a in geo.atoms]
  V = Matrix{Float64}(undef, b.nbf, b.nbf)
  check(ccall(dlsym(libqlb[], :qlb_nuclear_attraction), Cint, (Ptr{Nothing}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
              b.handle, length(Z), xyz, Z, V))
  V
end

function computeMatrixFock(b::B200Shells, geo::Geometry, density::Matrix)                                # :292-303
  H = computeMatrixKinetic(b) + computeMatrixNuclearAttraction(b, geo)
  F = similar(H)
  check(ccall(dlsym(libqlb[], :qlb_fock), Cint,
              (Ptr{Nothing}, Ptr{Float64}, Ptr{Float64}, Cdouble, Ptr{Float64}, Ptr{Nothing}),
              b.handle, H, Matrix{Float64}(density), b.tau, F, C_NULL))
  F
end

function computeTensorElectronRepulsionIntegrals(b::B200Shells)                                          # :198-207
  T = Array{Float64,4}(undef, b.nbf, b.nbf, b.nbf, b.nbf)
  check(ccall(dlsym(libqlb[], :qlb_eri_tensor), Cint, (Ptr{Nothing}, Ptr{Float64}), b.handle, T))
  T
end

"Block of one shell quartet, 1-based shell indices into `b` (SpecialMatricesModule.jl:184-186, LibInt2Module.jl:319-334)."
function computeTensorBlockElectronRepulsionIntegrals(b::B200Shells, i::Int, j::Int, k::Int, l::Int)
  blk = Array{Float64,4}(undef, nbf(b[i]), nbf(b[j]), nbf(b[k]), nbf(b[l]))
  check(ccall(dlsym(libqlb[], :qlb_eri_block), Cint, (Ptr{Nothing}, Cint, Cint, Cint, Cint, Ptr{Float64}),
              b.handle, i - 1, j - 1, k - 1, l - 1, blk))
  blk
end

"(mu nu|la si) for components (1-based) of four shells of `b`: IntegralsModule.jl:447-462 for the functions expandShell yields."
function computeElectronRepulsionIntegral(b::B200Shells, mu::Tuple{Int,Int}, nu::Tuple{Int,Int}, la::Tuple{Int,Int}, si::Tuple{Int,Int})
  blk = computeTensorBlockElectronRepulsionIntegrals(b, mu[1], nu[1], la[1], si[1])
  blk[mu[2], nu[2], la[2], si[2]]
end

# ---- resolution of the identity (ResolutionIdentityModule.jl:41-78): B stays on the device inside the handle
mutable struct B200RI
  orbital::B200Shells
  handle::Ptr{Nothing}
  naux::Int
end
function B200RI(b::B200Shells, aux::Vector{Shell})
  centers = Float64[c for sh in aux for c in (sh.center.x, sh.center.y, sh.center.z)]
  L = Int32[sh.lqn.exponent for sh in aux]; nprim = Int32[length(sh.exponents) for sh in aux]
  exps = vcat([sh.exponents for sh in aux]...); coefs = vcat([sh.coefficients for sh in aux]...)
  h = Ref{Ptr{Nothing}}(C_NULL)
  check(ccall(dlsym(libqlb[], :qlb_ri_create), Cint,
              (Ptr{Nothing}, Cint, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Ptr{Nothing}}),
              b.handle, length(aux), centers, L, nprim, exps, coefs, h))
  B200RI(b, h[], mapreduce(nbf, +, aux))
end
destroy!(r::B200RI) = (ccall(dlsym(libqlb[], :qlb_ri_destroy), Cint, (Ptr{Nothing},), r.handle); r.handle = C_NULL)
# the reference's signatures take (basis, basis_aux); these methods add the device-resident pair
function computeMatrixResolutionOfTheIdentityCoulombMetric(r::B200RI)                                    # :41-50
  B = zeros(Float64, r.orbital.nbf, r.orbital.nbf, r.naux)
  check(ccall(dlsym(libqlb[], :qlb_ri_b_tensor), Cint, (Ptr{Nothing}, Ptr{Float64}), r.handle, B)); B
end
function computeMatrixResolutionOfTheIdentityCoulombMetric(b::B200Shells, aux::Vector{Shell})
  r = B200RI(b, aux); B = computeMatrixResolutionOfTheIdentityCoulombMetric(r); destroy!(r); B
end
function computeTensorElectronRepulsionIntegralsRICoulomb(b::B200Shells, aux::Vector{Shell})            # :54-58 (N <= 64)
  r = B200RI(b, aux)
  T = Array{Float64,4}(undef, b.nbf, b.nbf, b.nbf, b.nbf)
  check(ccall(dlsym(libqlb[], :qlb_ri_eri_tensor), Cint, (Ptr{Nothing}, Ptr{Float64}), r.handle, T)); destroy!(r); T
end
function computeMatrixExchangeRIK(r::B200RI, density::Matrix)                                            # :69-78
  K = Matrix{Float64}(undef, r.orbital.nbf, r.orbital.nbf)
  check(ccall(dlsym(libqlb[], :qlb_ri_build_jk), Cint, (Ptr{Nothing}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
              r.handle, Matrix{Float64}(density), C_NULL, K, C_NULL)); K
end
function computeMatrixExchangeRIK(b::B200Shells, aux::Vector{Shell}, density::Matrix)
  r = B200RI(b, aux); K = computeMatrixExchangeRIK(r, density); destroy!(r); K
end
"RI-J and RI-K from the occupied orbital coefficients (N x nocc): the form an SCF loop uses."
function buildJKoccRI(r::B200RI, Cocc::Matrix{Float64})
  n = r.orbital.nbf
  J = Matrix{Float64}(undef, n, n); K = similar(J)
  check(ccall(dlsym(libqlb[], :qlb_ri_build_jk_occ), Cint, (Ptr{Nothing}, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
              r.handle, Cocc, size(Cocc, 2), J, K, C_NULL)); J, K
end

# ---- evaluateSCF(basis::B200Shells, geometry, guess, electronNumber = Ionization(0); ...): the convenience form
# (HartreeFockModule.jl:158-208) with the reference's argument types and keywords.  deviceSCF = true (default) keeps
# S, T, V, P, J, K, F and the eigenproblem on the device (qlb_scf_*: one iteration moves four doubles to the host);
# deviceSCF = false hands the integral-direct closures of the libint branch (:197-199, 207) to the primitive form.
# Iteration semantics are the reference's (:104-140) in both cases.
function evaluateSCF(b::B200Shells,
                     geometry::Union{Geometry,String},
                     initialGuessDensity::Union{Matrix,InitialGuess},
                     electronNumber::Union{Integer,Ionization}=Ionization(0);
                     detailedinfo::Bool=true, info::Bool=true, energyConvergenceCriterion::Float64=1e-8, deviceSCF::Bool=true)
  if isa(geometry, String)
    geometry = Geometry(geometry)
  end
  if isa(electronNumber, Ionization)
    electronNumber = Int(computeNumberElectrons(geometry, electronNumber.charge) / 2)
  end
  S = computeMatrixOverlap(b); T = computeMatrixKinetic(b); V = computeMatrixNuclearAttraction(b, geometry)
  Pinit = Matrix{Float64}(computeDensityMatrixGuess(initialGuessDensity, size(S)[1]))
  Vnn = computeEnergyInteratomicRepulsion(geometry)
  if !deviceSCF
    return evaluateSCF(Pinit, S, T, V, P -> computeMatrixCoulomb(b, P), P -> computeMatrixExchange(b, P), Vnn, electronNumber;
                       energyConvergenceCriterion=energyConvergenceCriterion, info=info, detailedinfo=detailedinfo)
  end
  h = Ref{Ptr{Nothing}}(C_NULL)
  check(ccall(dlsym(libqlb[], :qlb_scf_create), Cint, (Ptr{Nothing}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Ptr{Nothing}}),
              b.handle, S, T, V, electronNumber, h))
  scf = h[]
  e4 = zeros(Float64, 4)   # 2tr(TP), 2tr(VP), 2tr(JP), -tr(KP)  (HartreeFockModule.jl:25-35)
  build() = (check(ccall(dlsym(libqlb[], :qlb_scf_build), Cint, (Ptr{Nothing}, Cdouble, Cint, Ptr{Float64}, Ptr{Nothing}),
                         scf, b.tau, 0, e4, C_NULL)); sum(e4))
  try
    check(ccall(dlsym(libqlb[], :qlb_scf_set_density), Cint, (Ptr{Nothing}, Ptr{Float64}), scf, Pinit))
    totalEnergy = build()
    i = 0
    while true
      i += 1
      check(ccall(dlsym(libqlb[], :qlb_scf_step), Cint, (Ptr{Nothing},), scf))   # evaluateSCFStep (:73-84)
      oldEnergy = totalEnergy
      totalEnergy = build()
      if info
        println("E[$i]: $totalEnergy")
      end
      if abs(totalEnergy - oldEnergy) < energyConvergenceCriterion
        info && println("Converged!")
        break
      end
    end
    n = b.nbf
    energies = Vector{Float64}(undef, n); P = Matrix{Float64}(undef, n, n)
    check(ccall(dlsym(libqlb[], :qlb_scf_get), Cint, (Ptr{Nothing}, Cint, Ptr{Float64}), scf, 4, energies))
    check(ccall(dlsym(libqlb[], :qlb_scf_get), Cint, (Ptr{Nothing}, Cint, Ptr{Float64}), scf, 0, P))
    return energies, totalEnergy, P
  finally
    ccall(dlsym(libqlb[], :qlb_scf_destroy), Cint, (Ptr{Nothing},), scf)
  end
end

end # module
