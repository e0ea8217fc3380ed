# StuffingB200.jl — the reference-side binding: thin `ccall` glue over libstuffing_b200.so that keeps
# Stuffing.jl's own signatures.  NOT EXECUTED in the build image (no Julia there); it is a 1:1 mirror of
# the Python host layer (stuffing.jl_b200/qtrees.py, trainer.py), which is what the tests exercise.
#
# Usage inside Stuffing.jl:   include("StuffingB200.jl"); using .StuffingB200
#   dq  = DeviceQTrees(qtrees(objs, mask=mask))      # upload + build on the GPU
#   C   = totalcollisions_spacial(dq; unique=false)  # Vector{CoItem}, 1-based, like the reference
#   batchsteps!(dq, C, Momentum(η=1/4))              # grad2d + step! + rebuild on the GPU
module StuffingB200
export DeviceQTrees, totalcollisions_spacial, totalcollisions_native, partialcollisions, locate!, collision,
       batchsteps!, fit_round!, getshift, setshift!, shift!, download!, grad2d, filttrain!, place!, ground_level,
       shard_collide!, shard_pack!, shard_step!

const LIB = get(ENV, "STUFFING_B200_LIB", joinpath(@__DIR__, "..", "stuffing.jl_b200", "libstuffing_b200.so"))
const CoItem = Pair{Tuple{Int,Int},Tuple{Int,Int,Int}}

struct StfItem            # include/stuffing_b200.h: stf_item
    i1::Int32; i2::Int32; l::Int32; a::Int32; b::Int32
end
CoItem(x::StfItem) = (Int(x.i1), Int(x.i2)) => (Int(x.l), Int(x.a), Int(x.b))
StfItem(c::CoItem) = StfItem(c.first[1], c.first[2], c.second[1], c.second[2], c.second[3])

mutable struct DeviceQTrees
    h::Ptr{Cvoid}
    n::Int
    function DeviceQTrees(qts::AbstractVector; device::Integer=0)   # qts::Vector{U8SQTree} of the reference
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:stf_create, LIB), Int32, (Int32, Ptr{Ptr{Cvoid}}), device, h)
        rc == 0 || error("stf_create failed ($rc): no CUDA device — there is no CPU fallback")
        n = length(qts)
        kernels = [qt[1].kernel for qt in qts]                       # PaddedMat.kernel :: Matrix{UInt8}, (kh+3)x(kw+3)
        GC.@preserve kernels begin
            pics  = [pointer(k, LinearIndices(k)[2, 2]) for k in kernels]   # payload starts at kernel[2,2]
            kh    = Int32[size(k, 1) - 3 for k in kernels]
            kw    = Int32[size(k, 2) - 3 for k in kernels]
            pitch = Int32[size(k, 1) for k in kernels]
            rs    = Int32[qt[1].rshift for qt in qts]
            cs    = Int32[qt[1].cshift for qt in qts]
            dflt  = UInt8[qt[1].default for qt in qts]
            canvas = n > 0 ? size(qts[1][1], 1) : 2
            rc = ccall((:stf_upload_objects, LIB), Int32,
                       (Ptr{Cvoid}, Int32, Ptr{Ptr{UInt8}}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{UInt8}, Int32, Ptr{UInt8}, Ptr{Int32}),
                       h[], n, pics, kh, kw, pitch, rs, cs, dflt, canvas, C_NULL, C_NULL)
        end
        d = new(h[], n)
        check(d, rc)
        finalizer(x -> ccall((:stf_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.h), d)
    end
end
Base.length(d::DeviceQTrees) = d.n
check(d::DeviceQTrees, rc) = rc == 0 || error(unsafe_string(ccall((:stf_last_error, LIB), Cstring, (Ptr{Cvoid},), d.h)))

# two-call pattern for result lists
function _list(f::Symbol, d::DeviceQTrees, args...; cap=1 << 14)
    buf = Vector{StfItem}(undef, cap); cnt = Ref{Int64}(0)
    rc = _call(f, d, args..., buf, cap, cnt)
    if rc == -3                                  # STF_E_CAPACITY
        resize!(buf, cnt[])
        rc = ccall((:stf_fetch_result, LIB), Int32, (Ptr{Cvoid}, Ptr{StfItem}, Int64, Ptr{Int64}), d.h, buf, cnt[], cnt)
    end
    check(d, rc)
    CoItem.(@view buf[1:cnt[]])
end
_labels(::Nothing) = (C_NULL, Int32(0))
_labels(l) = (v = Int32.(collect(l)); (v, Int32(length(v))))
function _call(f::Symbol, d, labels, unique, buf, cap, cnt)
    lb, nl = _labels(labels)
    f === :stf_totalcollisions_spacial ?
        ccall((:stf_totalcollisions_spacial, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{StfItem}, Int64, Ptr{Int64}), d.h, nl, lb, unique, buf, cap, cnt) :
    f === :stf_partialcollisions ?
        ccall((:stf_partialcollisions, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{StfItem}, Int64, Ptr{Int64}), d.h, nl, lb, unique, buf, cap, cnt) :
        ccall((:stf_totalcollisions_native, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{StfItem}, Int64, Ptr{Int64}), d.h, nl, lb, buf, cap, cnt)
end

# src/qtree_functions.jl:225-263
totalcollisions_spacial(d::DeviceQTrees, labels=nothing; unique=true, kargs...) = _list(:stf_totalcollisions_spacial, d, labels, Int32(unique))
# src/qtree_functions.jl:111-131
totalcollisions_native(d::DeviceQTrees, labels=nothing; kargs...) = _list(:stf_totalcollisions_native, d, labels, Int32(0))
# src/qtree_functions.jl:277-330 (the linked tree lives in the context; stf_linked_locate = locate!(qts, labels, tree))
partialcollisions(d::DeviceQTrees, labels; unique=true, kargs...) = _list(:stf_partialcollisions, d, labels, Int32(unique))

# src/qtree_functions.jl:43-51
function collision(d::DeviceQTrees, i1::Integer, i2::Integer)
    out = Vector{Int32}(undef, 3)
    check(d, ccall((:stf_collision, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}), d.h, i1, i2, out))
    (Int(out[1]), Int(out[2]), Int(out[3]))
end

# src/qtree_functions.jl:190-201 → Dict{Index,Vector{Int}} like HashSpacialQTree.qtree
function locate!(d::DeviceQTrees, labels=nothing)
    lb, nl = _labels(labels)
    cap = 4 * (labels === nothing ? d.n : Int(nl)) + 4
    buf = Matrix{Int32}(undef, 4, cap); cnt = Ref{Int64}(0)
    check(d, ccall((:stf_locate, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Int64, Ptr{Int64}), d.h, nl, lb, buf, cap, cnt))
    t = Dict{Tuple{Int,Int,Int},Vector{Int}}()
    for k in 1:cnt[]
        push!(get!(Vector{Int}, t, (Int(buf[1, k]), Int(buf[2, k]), Int(buf[3, k]))), Int(buf[4, k]))
    end
    t
end

# src/qtrees.jl:267-288
function _shifts!(d, labels, l, s1, s2, mode)
    lb, n = _labels(labels)
    check(d, ccall((:stf_set_shifts, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Int32}, Int32),
                   d.h, n, lb, l, Int32.(collect(s1)), Int32.(collect(s2)), mode))
end
setshift!(d::DeviceQTrees, label::Integer, l::Integer, s1::Integer, s2::Integer) = _shifts!(d, (label,), l, (s1,), (s2,), 0)
shift!(d::DeviceQTrees, label::Integer, l::Integer, s1::Integer, s2::Integer) = _shifts!(d, (label,), l, (s1,), (s2,), 1)
function getshift(d::DeviceQTrees, labels=1:d.n, l::Integer=1)
    lb, n = _labels(labels); rs = Vector{Int32}(undef, n); cs = similar(rs)
    check(d, ccall((:stf_get_shifts, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Int32}), d.h, n, lb, l, rs, cs))
    collect(zip(Int.(rs), Int.(cs)))
end
# refresh a host tree from the device before place!/reposition!/getpositions (host code stays the reference's)
function download!(d::DeviceQTrees, label::Integer, qt)
    for l in 1:length(qt)
        info = Vector{Int32}(undef, 5)
        check(d, ccall((:stf_get_level_info, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}), d.h, label, l, info))
        qt[l].rshift, qt[l].cshift = info[3], info[4]
        payload = Matrix{UInt8}(undef, info[1], info[2])
        check(d, ccall((:stf_download_level, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), d.h, label, l, payload))
        qt[l].kernel[2:end-2, 2:end-2] .= payload
    end
    qt
end

# src/fit_helper.jl:51-66
function grad2d(d::DeviceQTrees, label::Integer, l::Integer, a::Integer, b::Integer)
    out = Vector{Int32}(undef, 2)
    check(d, ccall((:stf_grad2d, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                   d.h, 1, Int32[label], Int32[l], Int32[a], Int32[b], out))
    (Int(out[1]), Int(out[2]))
end

# src/fit.jl:69-75; optimiser = the reference's Momentum/SGD structs (only η, ρ cross the boundary)
const _clock = Ref((UInt64(0), UInt64(0)))   # (seed, iteration) of the canonical-mode draws
_optkind(o) = hasproperty(o, :rho) ? (Int32(2), o.eta, o.rho) : hasproperty(o, :eta) ? (Int32(1), o.eta, 0.0) : (Int32(0), 0.0, 0.0)
function batchsteps!(d::DeviceQTrees, colist::Vector{CoItem}, optimiser=nothing)
    kind, eta, rho = optimiser === nothing ? (Int32(0), 0.0, 0.0) : _optkind(optimiser)
    check(d, ccall((:stf_set_optimiser, LIB), Int32, (Ptr{Cvoid}, Int32, Float64, Float64), d.h, kind, eta, rho))
    seed, it = _clock[]
    check(d, ccall((:stf_batchsteps, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{StfItem}, UInt64, UInt64), d.h, length(colist), StfItem.(colist), seed, it))
    _clock[] = (seed, it + 1)
    nothing
end
# device-resident totalcollisions(...; unique=false) → batchsteps!; returns length(colist)
function fit_round!(d::DeviceQTrees, labels=nothing; partial=false)
    lb, nl = _labels(labels); nh = Ref{Int64}(0); seed, it = _clock[]
    check(d, ccall((:stf_fit_round, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, UInt64, UInt64, Ptr{Int64}), d.h, partial ? 1 : 0, nl, lb, seed, it, nh))
    _clock[] = (seed, it + 1)
    Int(nh[])
end

# the same round, then getshift(t, level) of every tree, read back with the round's own synchronisation
function fit_round_positions!(d::DeviceQTrees, labels=nothing; partial=false, level=1)
    lb, nl = _labels(labels); nh = Ref{Int64}(0); seed, it = _clock[]
    n = Int(ccall((:stf_num_objects, LIB), Int32, (Ptr{Cvoid},), d.h))
    rs = Vector{Int32}(undef, n); cs = Vector{Int32}(undef, n)
    check(d, ccall((:stf_fit_round_positions, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, UInt64, UInt64, Ptr{Int64}, Int32, Ptr{Int32}, Ptr{Int32}),
                   d.h, partial ? 1 : 0, nl, lb, seed, it, nh, level, rs, cs))
    _clock[] = (seed, it + 1)
    Int(nh[]), rs, cs
end

# src/fit.jl:235-266 — one root BFS per pair of the pool (misses report -level: the nearness score), hits are stepped
function filttrain!(d::DeviceQTrees, inpool::Vector{Tuple{Int,Int}}, outpool, nearlevel2; optimiser=nothing)
    isempty(inpool) && return 0
    L = Int(ccall((:stf_num_levels, LIB), Int32, (Ptr{Cvoid},), d.h))
    items = [StfItem(i1, i2, L, 1, 1) for (i1, i2) in inpool]; out = similar(items)
    check(d, ccall((:stf_collision_bfs, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{StfItem}, Ptr{StfItem}), d.h, length(items), items, out))
    near = [o.l >= nearlevel2 for o in out]
    outpool === nothing || append!(outpool, inpool[near])
    colist = CoItem[CoItem(o) for (o, k) in zip(out, near) if k && o.l > 0]
    batchsteps!(d, colist, optimiser)
    length(colist)
end

# src/qtree_functions.jl:452-518 — the ground place!/reposition! search: deepcopy(mask) + every tree not in `inds`,
# composed on the device; findroom_uniform (the reference's own, unchanged) then runs on the downloaded ground
function ground_level(d::DeviceQTrees, l::Integer)
    info = Vector{Int32}(undef, 5)
    check(d, ccall((:stf_ground_level, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{UInt8}, Ptr{Int32}), d.h, l, C_NULL, info))
    m = Matrix{UInt8}(undef, info[1], info[2])
    check(d, ccall((:stf_ground_level, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{UInt8}, Ptr{Int32}), d.h, l, m, info))
    m, (Int(info[3]), Int(info[4]))                       # payload and (rshift, cshift): fill a PaddedMat with it
end
function place!(d::DeviceQTrees, host_ground, qts, inds; kargs...)   # host_ground: a ShiftedQTree shaped like the mask
    ex = Int32.(collect(inds))
    check(d, ccall((:stf_ground_compose, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}), d.h, 1, length(ex), ex))
    for l in 1:length(host_ground)
        m, _ = ground_level(d, l); host_ground[l].kernel[2:size(m, 1) + 1, 2:size(m, 2) + 1] .= m
    end
    ind = nothing; Q = Vector{Tuple{Int,Int,Int}}()
    for i in inds                                           # the reference's loop, qtree_functions.jl:512-517
        ind = Main.Stuffing.QTrees.place!(host_ground, qts[i], Q; kargs...)
        ind === nothing && return ind
        Main.Stuffing.QTrees.overlap!(host_ground, qts[i])
    end
    rs = Int32[qts[i][1].rshift for i in inds]; cs = Int32[qts[i][1].cshift for i in inds]
    check(d, ccall((:stf_set_shifts, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Int32}, Int32), d.h, length(ex), ex, 1, rs, cs, 0))
    ind
end

# item-parallel sharding of one scene (INTEGRATION.md "Multi-GPU from Julia"): the three calls around the all-gather
function shard_collide!(d::DeviceQTrees, rank, nranks; labels=nothing, partial=false)
    lb, nl = _labels(labels); n = Ref{Int64}(0)
    check(d, ccall((:stf_shard_collide, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Int32, Int32, Ptr{Int64}), d.h, partial ? 1 : 0, nl, lb, rank, nranks, n))
    Int(n[])
end
shard_pack!(d::DeviceQTrees, send_dev, stride) = check(d, ccall((:stf_shard_pack, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), d.h, send_dev, stride))
function shard_step!(d::DeviceQTrees, recv_dev, nranks, stride)
    nh = Ref{Int64}(0); seed, it = _clock[]
    check(d, ccall((:stf_shard_step, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, UInt64, UInt64, Ptr{Int64}), d.h, recv_dev, nranks, stride, seed, it, nh))
    _clock[] = (seed, it + 1)
    Int(nh[])
end
end # module
