# QuiqboxB200.jl -- reference-side binding of libqbx.so (include/qbx.h).
#
# NOT executed in the build container (no Julia there); it is the ~150 lines a Quiqbox.jl
# maintainer would add.  It overloads the two seams of SURVEY.md section 8b by multiple dispatch,
# leaves every other method of Quiqbox untouched, and falls through to Quiqbox's own Julia
# methods whenever the eligibility guard fails.  The same C entry points, in the same order,
# are exercised from Python (quiqbox.jl_b200/integrals.py) by the parity tests.
#
#   seam 1  Quiqbox.getOrbVectorIntegralCore!   src/Integration/Framework.jl:640-698
#   seam 2  Quiqbox.getGcore                    src/HartreeFock.jl:305-319
module QuiqboxB200

using Quiqbox
using Quiqbox: TwoBodyOrbIntegralInfo, OrbCorePointerVector, CoulombInteractionSampler,
               FloatingPolyGaussField, prepareOrbitalInfoCore, PrimGaussTypeOrb

const libqbx = get(ENV, "QBX_LIB", joinpath(@__DIR__, "..", "quiqbox.jl_b200", "libqbx.so"))

struct QbxError <: Exception
    code::Cint
    msg::String
end
check(rc::Cint) = rc == 0 ? nothing :
    throw(QbxError(rc, unsafe_string(ccall((:qbx_last_error, libqbx), Cstring, ()))))

function __init__()
    ndev = Ref{Cint}(0)
    check(ccall((:qbx_init, libqbx), Cint, (Cint, Ptr{Cint}), parse(Cint, get(ENV, "LOCAL_RANK", "0")), ndev))
end

"Return the library's idle device blocks to the driver (they are kept for the next basis otherwise)."
trim_pool() = check(ccall((:qbx_pool_trim, libqbx), Cint, (Ptr{Int64},), C_NULL))

# ---- opaque handle: owns the device copy of one basis set -----------------------------------
mutable struct DeviceBasis
    ptr::Ptr{Cvoid}
    nbf::Int
    function DeviceBasis(cen::Matrix{Float64}, xpn::Vector{Float64}, ang::Matrix{Int32},
                         off::Vector{Int64}, prim::Vector{Int64}, w::Vector{Float64})
        h = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve cen xpn ang off prim w check(ccall((:qbx_basis_create, libqbx), Cint,
            (Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64},
             Ptr{Ptr{Cvoid}}), length(xpn), cen, xpn, ang, length(off) - 1, off, prim, w, h))
        obj = new(h[], length(off) - 1)
        finalizer(b -> ccall((:qbx_basis_destroy, libqbx), Cint, (Ptr{Cvoid},), b.ptr), obj)
    end
end

# Flatten what the seam receives: inteInfo.source.left (primitive fields) and
# ptrVector[i].inner :: MemoryPair{OneToIndex, T} (Framework.jl:134-153).
function flatten(source, ptrVector)
    prims = map(prepareOrbitalInfoCore, source)                 # GaussianOrbitals.jl:65-78
    np = length(prims)
    cen = Matrix{Float64}(undef, 3, np); ang = Matrix{Int32}(undef, 3, np)
    xpn = Vector{Float64}(undef, np)
    for (p, o) in enumerate(prims)
        cen[:, p] .= o.cen; ang[:, p] .= o.ang; xpn[p] = o.xpn
    end
    off = Int64[0]; prim = Int64[]; w = Float64[]
    for ptr in ptrVector
        for (idx, weight) in zip(ptr.inner.left, ptr.inner.right)
            push!(prim, idx.idx - 1); push!(w, weight)          # 1-based -> 0-based
        end
        push!(off, length(prim))
    end
    DeviceBasis(cen, xpn, ang, off, prim, w)
end

eligible(source) = all(f -> f isa FloatingPolyGaussField{Float64, 3}, source)

# Resident mode.  initializeHartreeFock keeps whatever seam 1 returns in ElecHamiltonianConfig.twoBody
# (`_, eriH = computeOrbDataIntegral(style2B, eriOp, eriBasis)`, HartreeFock.jl:189-193; the result of
# getOrbVectorIntegralCore! is passed through unchanged, Framework.jl:835-840, 907-916), and that field is
# typed A4 <: AbstractArray{T,4}.  With RESIDENT[] = true seam 1 therefore returns the device handle
# (packed unique ERIs in HBM, qbx_eri_store) instead of materialising N^4 doubles on the host, and every
# later getGcore(HeeI, DJ, DK) of the SCF loop dispatches to qbx_fock_build (seam 2 below):
#     QuiqboxB200.with_device_eri() do;  runHartreeFock(nucInfo, bs);  end
const RESIDENT = Ref(false)
const SCREEN = Ref(1e-12)          # Schwarz threshold of the resident store
const SHARD = Ref((0, 1))          # (rank, nranks) of this process (one Julia process per GPU)
function with_device_eri(f; screen::Float64=1e-12, rank::Integer=0, nranks::Integer=1)
    old = (RESIDENT[], SCREEN[], SHARD[])
    RESIDENT[], SCREEN[], SHARD[] = true, screen, (Int(rank), Int(nranks))
    try
        return f()
    finally
        RESIDENT[], SCREEN[], SHARD[] = old
    end
end

# ---- seam 1: the whole N^4 tensor (elecRepulsions) ------------------------------------------
function Quiqbox.getOrbVectorIntegralCore!(
        inteInfo::TwoBodyOrbIntegralInfo{Float64, 3, Float64, <:CoulombInteractionSampler},
        ptrVector::OrbCorePointerVector{3, Float64})
    src = inteInfo.source.left
    all(==(PrimGaussTypeOrb), inteInfo.source.right) && eligible(src) ||
        return invoke(Quiqbox.getOrbVectorIntegralCore!,
                      Tuple{Quiqbox.TwoBodyOrbIntegralInfo, Quiqbox.OrbCorePointerVector}, inteInfo, ptrVector)
    RESIDENT[] && return DeviceERI(src, ptrVector; screen=SCREEN[], rank=SHARD[][1], nranks=SHARD[][2])
    b = flatten(src, ptrVector)
    n = b.nbf
    out = Array{Float64}(undef, n, n, n, n)                     # column-major, tensor[i,j,k,l] = (ij|kl)
    GC.@preserve out check(ccall((:qbx_eri_tensor, libqbx), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64),
                                 b.ptr, out, sizeof(out)))
    out
end

# ---- seam 2: device-resident ERIs that stand where HeeI::Array{T,4} stood --------------------
# ElecHamiltonianConfig.twoBody is typed A4 <: AbstractArray{T,4} (HartreeFock.jl:143-151), so
# this handle slots in without touching the SCF code.
struct DeviceERI <: AbstractArray{Float64, 4}
    basis::DeviceBasis
end
Base.size(e::DeviceERI) = ntuple(_ -> e.basis.nbf, 4)
function Base.getindex(e::DeviceERI, i::Int, j::Int, k::Int, l::Int)   # slow path, e.g. printing
    out = Ref(0.0); idx = Int64[i - 1, j - 1, k - 1, l - 1]
    GC.@preserve idx check(ccall((:qbx_eri_quartets, libqbx), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}),
                                 e.basis.ptr, 1, idx, out))
    out[]
end

# ---- multi-GPU: one Julia process per GPU.  The partial Fock matrices are summed INSIDE qbx_fock_build by the library's
# own ncclAllReduce, so that getGcore keeps its contract (the result is the full G, HartreeFock.jl:322-327).  Rank 0
# obtains the 128-byte NCCL id and hands it to the others by the host's own channel, e.g. with MPI.jl:
#     id = zeros(UInt8, 128); rank == 0 && QuiqboxB200.comm_unique_id!(id); MPI.Bcast!(id, 0, comm)
#     QuiqboxB200.comm_init(rank, nranks, id)
comm_unique_id!(id::Vector{UInt8}) = (length(id) == 128 || throw(ArgumentError("id must hold 128 bytes"));
    GC.@preserve id check(ccall((:qbx_comm_unique_id, libqbx), Cint, (Ptr{UInt8},), id)); id)
comm_init(rank::Integer, nranks::Integer, id::Vector{UInt8}) =
    GC.@preserve id check(ccall((:qbx_comm_init, libqbx), Cint, (Cint, Cint, Ptr{UInt8}), rank, nranks, id))
function comm_size()
    r = Ref{Cint}(0); n = Ref{Cint}(1)
    check(ccall((:qbx_comm_info, libqbx), Cint, (Ptr{Cint}, Ptr{Cint}), r, n))
    (Int(r[]), Int(n[]))
end

function DeviceERI(source, ptrVector; screen::Float64=1e-12, mode::Integer=0, rank::Integer=0, nranks::Integer=1)
    # a sharded store without the communicator would make getGcore return a PARTIAL G into Quiqbox's unchanged SCF loop
    nranks > 1 && comm_size() != (Int(rank), Int(nranks)) &&
        throw(ArgumentError("QuiqboxB200: nranks = $nranks needs QuiqboxB200.comm_init(rank, nranks, id) on every rank first"))
    b = flatten(source, ptrVector)
    check(ccall((:qbx_eri_store, libqbx), Cint, (Ptr{Cvoid}, Float64, Cint, Cint, Cint), b.ptr, screen, mode, rank, nranks))
    DeviceERI(b)
end

# getGcore(HeeI, DJ, DK): replaces the Threads.@threads loop over (mu, nu) by one call.  DJ and DK must be symmetric
# (they are densities); the library rejects anything else with an error instead of a mode-dependent result.
function Quiqbox.getGcore(HeeI::DeviceERI, DJ::Matrix{Float64}, DK::Matrix{Float64})
    G = similar(DJ)
    GC.@preserve DJ DK G check(ccall((:qbx_fock_build, libqbx), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), HeeI.basis.ptr, 1, DJ, DK, G))
    G
end

# UHF: both exchange densities in one pass over the stored integrals (getG, HartreeFock.jl:326-327)
function Quiqbox.getG(HeeI::DeviceERI, (Da, Db)::NTuple{2, Matrix{Float64}})
    n = size(Da, 1)
    DK = cat(Da, Db; dims=3); G = Array{Float64}(undef, n, n, 2); DJ = Da + Db
    GC.@preserve DJ DK G check(ccall((:qbx_fock_build, libqbx), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), HeeI.basis.ptr, 2, DJ, DK, G))
    (G[:, :, 1], G[:, :, 2])
end

end # module
