# StuffingB200.jl — the reference-side binding: `ccall` glue over libstuffing_b200.so that KEEPS Stuffing.jl's own
# signatures for the hot path (SURVEY §8b).  NOT EXECUTED in the build image (Julia is absent there); it is a 1:1 mirror
# of the Python host layer (stuffing.jl_b200/qtrees.py, trainer.py), which is what the tests exercise, and
# tests/test_abi_and_host.py checks every `ccall` argument tuple in this file against the library's exports.
#
# Drop-in use inside Stuffing.jl (src/Stuffing.jl would `include` this file next to qtrees.jl):
#   qts  = qtrees(objs, mask=mask)                   # the reference's own host trees (Vector{U8SQTree})
#   dqts = DeviceQTrees(qts)                         # upload + build on the GPU; dqts isa AbstractVector{DeviceTree}
#   C    = totalcollisions(dqts)                     # same dispatcher, same keywords, Vector{CoItem}
#   collision(dqts[2], dqts[3]); grad2d(dqts[2], l, a, b); shift!(dqts[2], 1, 2, 3); getshift(dqts[2])
#   fit!(dqts, 100; trainer=trainepoch_EM2!)         # the reference's train! loop, unchanged, over the device trainers
#   sync_to_host!(dqts, qts)                         # shifts + levels back into the host trees (before place!/getpositions on them)
#
# Every function below cites the reference definition it stands in for (src/...:line).  Integers cross the boundary as
# the reference prints them (1-based labels and (l, a, b)); matrices are column-major on both sides.
module StuffingB200
import ..QTrees: AbstractStackedQTree, U8SQTree, CoItem, Index, collision, collision_dfs, totalcollisions, totalcollisions_native,
                 totalcollisions_spacial, partialcollisions, dynamiccollisions, locate!, hash_spacial_qtree, linked_spacial_qtree,
                 shift!, setshift!, setcenter!, getshift, getcenter, kernelsize, buildqtree!, inkernelbounds, EMPTY, FULL, MIX
import ..Trainer: grad2d, batchsteps!, Momentum, SGD, eta, eta!, reset!
export DeviceQTrees, DeviceTree, DeviceHashTree, DeviceLinkedTree, DeviceColliders, sync_to_host!, fit_round!, fit_rounds!, trainepoch_E_device!, filttrain!, GpuGroup

const LIB = get(ENV, "STUFFING_B200_LIB", joinpath(@__DIR__, "..", "stuffing.jl_b200", "libstuffing_b200.so"))

struct StfItem            # include/stuffing_b200.h: stf_item
    i1::Int32; i2::Int32; l::Int32; a::Int32; b::Int32
end
_coitem(x::StfItem) = (Int(x.i1), Int(x.i2)) => (Int(x.l), Int(x.a), Int(x.b))
StfItem(c::CoItem) = StfItem(c.first[1], c.first[2], c.second[1], c.second[2], c.second[3])

# ------------------------------------------------------------------------------------------------ the vector of trees
"Vector{U8SQTree} resident on one GPU (one context of the C ABI).  Indexing yields DeviceTree handles."
mutable struct DeviceQTrees <: AbstractVector{AbstractStackedQTree}
    h::Ptr{Cvoid}
    n::Int
    L::Int
    clock::Tuple{UInt64,UInt64}                   # (seed, iteration) of the canonical-mode draws, one tick per batchsteps!
    function DeviceQTrees(qts::AbstractVector; device::Integer=0, seed::Integer=0)   # qts::Vector{U8SQTree} of the reference
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:stf_create, LIB), Int32, (Int32, Ptr{Ptr{Cvoid}}), device, h)
        rc == 0 || error("stf_create failed ($rc): no CUDA device — there is no CPU fallback")
        d = new(h[], length(qts), 0, (UInt64(seed), UInt64(0)))
        finalizer(x -> ccall((:stf_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.h), d)
        upload!(d, qts)
    end
end
function upload!(d::DeviceQTrees, qts::AbstractVector)
    n = length(qts)
    kernels = [qt[1].kernel for qt in qts]                       # PaddedMat.kernel :: Matrix{UInt8}, (kh+3)x(kw+3)  qtrees.jl:109-128
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
                   d.h, n, pics, kh, kw, pitch, rs, cs, dflt, canvas, C_NULL, C_NULL)
    end
    check(d, rc)
    d.n = n
    d.L = Int(ccall((:stf_num_levels, LIB), Int32, (Ptr{Cvoid},), d.h))
    d
end
check(d::DeviceQTrees, rc) = rc == 0 || error(unsafe_string(ccall((:stf_last_error, LIB), Cstring, (Ptr{Cvoid},), d.h)))
Base.size(d::DeviceQTrees) = (d.n,)
Base.getindex(d::DeviceQTrees, i::Int) = DeviceTree(d, i)
Base.IndexStyle(::Type{DeviceQTrees}) = IndexLinear()

# ------------------------------------------------------------------------------------------------ one tree (qtrees.jl:58-105, 170-294)
"One ShiftedQTree of a DeviceQTrees: the AbstractStackedQTree interface over the C ABI."
struct DeviceTree <: AbstractStackedQTree
    owner::DeviceQTrees
    label::Int
end
Base.length(t::DeviceTree) = t.owner.L                                      # qtrees.jl:217
function levelinfo(t::DeviceTree, l::Integer)                                # kh, kw, rshift, cshift, canvas size of level l
    out = Vector{Int32}(undef, 5)
    check(t.owner, ccall((:stf_get_level_info, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}), t.owner.h, t.label, l, out))
    Int.(out)
end
Base.size(t::DeviceTree) = (s = levelinfo(t, 1)[5]; (s, s))
function Base.getindex(t::DeviceTree, l::Integer)                            # t[l] -> Matrix{UInt8} payload of level l  qtrees.jl:98
    kh, kw = levelinfo(t, l)[1:2]
    m = Matrix{UInt8}(undef, kh, kw)
    check(t.owner, ccall((:stf_download_level, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), t.owner.h, t.label, l, m))
    m
end
function Base.getindex(t::DeviceTree, l::Integer, a::Integer, b::Integer)    # t[l, a, b]: PaddedMat getindex, default outside  qtrees.jl:156-162
    out = Vector{UInt8}(undef, 1)
    check(t.owner, ccall((:stf_read_nodes, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{UInt8}),
                         t.owner.h, 1, Int32[t.label], Int32[l], Int32[a], Int32[b], out))
    out[1]
end
Base.getindex(t::DeviceTree, ind::Index) = t[ind[1], ind[2], ind[3]]
kernelsize(t::DeviceTree, l::Integer=1) = (i = levelinfo(t, l); (i[1], i[2]))       # qtrees.jl:152
getshift(t::DeviceTree, l::Integer=1) = (i = levelinfo(t, l); (i[3], i[4]))         # qtrees.jl:288
getcenter(t::DeviceTree) = getshift(t) .+ kernelsize(t) .÷ 2                        # qtrees.jl:290
inkernelbounds(t::DeviceTree, l::Integer, a::Integer, b::Integer) = (i = levelinfo(t, l); i[3] < a <= i[3] + i[1] && i[4] < b <= i[4] + i[2])  # qtrees.jl:140-149
function _shifts!(d::DeviceQTrees, labels, l, s1, s2, mode)
    lb = Int32.(collect(labels))
    check(d, ccall((:stf_set_shifts, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Int32}, Int32),
                   d.h, length(lb), lb, l, Int32.(collect(s1)), Int32.(collect(s2)), mode))
end
shift!(t::DeviceTree, l::Integer, s1::Integer, s2::Integer) = (_shifts!(t.owner, (t.label,), l, (s1,), (s2,), 1); t)        # qtrees.jl:267-275
setshift!(t::DeviceTree, l::Integer, s1::Integer, s2::Integer) = (_shifts!(t.owner, (t.label,), l, (s1,), (s2,), 0); t)     # qtrees.jl:277-285
setshift!(t::DeviceTree, l::Integer, st::Tuple{Integer,Integer}) = setshift!(t, l, st...)
setshift!(t::DeviceTree, st::Tuple{Integer,Integer}) = setshift!(t, 1, st...)
setcenter!(t::DeviceTree, c) = setshift!(t, 1, (c .- kernelsize(t) .÷ 2)...)                                                # qtrees.jl:293-294
function buildqtree!(t::DeviceTree, layer::Integer=2)                                                                       # qtrees.jl:221-237
    check(t.owner, ccall((:stf_build, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32), t.owner.h, 1, Int32[t.label], layer)); t
end
# whole vectors at once (one launch): the batched forms the host mirror uses
setshift!(d::DeviceQTrees, labels, l::Integer, s1, s2) = _shifts!(d, labels, l, s1, s2, 0)
shift!(d::DeviceQTrees, labels, l::Integer, s1, s2) = _shifts!(d, labels, l, s1, s2, 1)
function getshift(d::DeviceQTrees, labels=1:d.n, l::Integer=1)
    lb = Int32.(collect(labels)); rs = Vector{Int32}(undef, length(lb)); cs = similar(rs)
    check(d, ccall((:stf_get_shifts, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Int32}), d.h, length(lb), lb, l, rs, cs))
    collect(zip(Int.(rs), Int.(cs)))
end
"device -> host: shifts of every level and the level payloads of `labels` into the reference's own trees (before host-side place!/reposition!)"
function sync_to_host!(d::DeviceQTrees, qts::AbstractVector, labels=1:d.n)
    lb = Int32.(collect(labels)); cnt = Ref{Int64}(0)
    ccall((:stf_download_trees, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{UInt8}, Int64, Ptr{Int64}), d.h, length(lb), lb, C_NULL, 0, cnt)
    buf = Vector{UInt8}(undef, cnt[])
    check(d, ccall((:stf_download_trees, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{UInt8}, Int64, Ptr{Int64}), d.h, length(lb), lb, buf, length(buf), cnt))
    pos = 0
    for i in lb, l in 1:d.L
        m = qts[i][l]; kh, kw = size(m.kernel) .- 3
        m.kernel[2:kh+1, 2:kw+1] .= reshape(view(buf, pos+1:pos+kh*kw), kh, kw); pos += kh * kw
        m.rshift, m.cshift = getshift(DeviceTree(d, i), l)
    end
    qts
end
"getpositions / setpositions!  (Stuffing.jl:78-99): (x, y) of every object relative to the mask (label 1)"
function getpositions(d::DeviceQTrees, inds=1:d.n-1)
    s = getshift(d); m = s[1]
    [(s[i+1][2] - m[2] + 1, s[i+1][1] - m[1] + 1) for i in inds]
end
function setpositions!(d::DeviceQTrees, inds, x_y)
    m = getshift(d, (1,))[1]; inds = collect(inds)
    xy = x_y isa Tuple ? fill(x_y, length(inds)) : collect(x_y)
    setshift!(d, inds .+ 1, 1, [p[2] - 1 + m[1] for p in xy], [p[1] - 1 + m[2] for p in xy]); x_y
end

# ------------------------------------------------------------------------------------------------ pair collision (qtree_functions.jl:2-51)
function collision(Q1::DeviceTree, Q2::DeviceTree)                                        # qtree_functions.jl:43-51
    @assert Q1.owner === Q2.owner
    out = Vector{Int32}(undef, 3)
    check(Q1.owner, ccall((:stf_collision, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}), Q1.owner.h, Q1.label, Q2.label, out))
    (Int(out[1]), Int(out[2]), Int(out[3]))
end
function _pairs(f::Symbol, d::DeviceQTrees, items::Vector{StfItem})
    out = similar(items)
    rc = f === :stf_collision_dfs ?
        ccall((:stf_collision_dfs, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{StfItem}, Ptr{StfItem}), d.h, length(items), items, out) :
        ccall((:stf_collision_bfs, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{StfItem}, Ptr{StfItem}), d.h, length(items), items, out)
    check(d, rc); out
end
function collision_dfs(Q1::DeviceTree, Q2::DeviceTree, i::Index=(length(Q1), 1, 1))       # qtree_functions.jl:2-20
    o = _pairs(:stf_collision_dfs, Q1.owner, [StfItem(Q1.label, Q2.label, i...)])[1]; (Int(o.l), Int(o.a), Int(o.b))
end

# ------------------------------------------------------------------------------------------------ many-body (qtree_functions.jl:52-370)
function _list(d::DeviceQTrees, colist::Vector{CoItem}, call)                             # two-call pattern of the ABI
    cap = 1 << 14; buf = Vector{StfItem}(undef, cap); cnt = Ref{Int64}(0)
    rc = call(buf, cap, cnt)
    if rc == -3                                                                            # STF_E_CAPACITY
        resize!(buf, cnt[])
        rc = ccall((:stf_fetch_result, LIB), Int32, (Ptr{Cvoid}, Ptr{StfItem}, Int64, Ptr{Int64}), d.h, buf, cnt[], cnt)
    end
    check(d, rc)
    append!(colist, _coitem.(@view buf[1:cnt[]]))
end
_lb(::Nothing) = (C_NULL, Int32(0))
_lb(l) = (v = Int32.(collect(l)); (v, Int32(length(v))))
function totalcollisions_native(d::DeviceQTrees, labels=nothing; colist=Vector{CoItem}(), kargs...)            # qtree_functions.jl:126-131
    lb, nl = _lb(labels)
    _list(d, colist, (buf, cap, cnt) -> ccall((:stf_totalcollisions_native, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{StfItem}, Int64, Ptr{Int64}), d.h, nl, lb, buf, cap, cnt))
end
"HashSpacialQTree stand-in (qtrees.jl:399-422): the device keeps the cells as sorted records; `qtree` is filled by locate!"
struct DeviceHashTree; qtree::Dict{Index,Vector{Int}}; end
"LinkedSpacialQTree stand-in (qtrees.jl:424-552): the persistent 4-slot table lives in the context"
struct DeviceLinkedTree; owner::DeviceQTrees; end
hash_spacial_qtree(::DeviceQTrees) = DeviceHashTree(Dict{Index,Vector{Int}}())           # qtree_functions.jl:139-140
function linked_spacial_qtree(d::DeviceQTrees)                                            # qtree_functions.jl:132-138
    check(d, ccall((:stf_linked_reset, LIB), Int32, (Ptr{Cvoid},), d.h)); DeviceLinkedTree(d)
end
Base.empty!(t::DeviceHashTree) = (empty!(t.qtree); t)
function locate!(d::DeviceQTrees, labels, t::DeviceHashTree)                              # qtree_functions.jl:190-201
    lb, nl = _lb(labels)
    cap = 4 * (labels === nothing ? d.n : Int(nl)) + 4
    buf = Matrix{Int32}(undef, 4, cap); cnt = Ref{Int64}(0)
    check(d, ccall((:stf_locate, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Int64, Ptr{Int64}), d.h, nl, lb, buf, cap, cnt))
    for k in 1:cnt[]
        push!(get!(Vector{Int}, t.qtree, (Int(buf[1, k]), Int(buf[2, k]), Int(buf[3, k]))), Int(buf[4, k]))
    end
    t
end
locate!(d::DeviceQTrees, t::DeviceHashTree=hash_spacial_qtree(d)) = locate!(d, nothing, t)
function locate!(d::DeviceQTrees, labels, t::DeviceLinkedTree)
    lb, nl = _lb(labels)
    check(d, ccall((:stf_linked_locate, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}), d.h, nl, lb)); t
end
locate!(d::DeviceQTrees, t::DeviceLinkedTree) = locate!(d, nothing, t)
function totalcollisions_spacial(d::DeviceQTrees, labels=nothing; colist=Vector{CoItem}(), unique=true, kargs...)  # qtree_functions.jl:225-263
    lb, nl = _lb(labels)
    _list(d, colist, (buf, cap, cnt) -> ccall((:stf_totalcollisions_spacial, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{StfItem}, Int64, Ptr{Int64}), d.h, nl, lb, Int32(unique), buf, cap, cnt))
end
const SPACIAL_ENABLE_THRESHOLD = round(Int, 10 + 10log2(Threads.nthreads()))               # qtree_functions.jl:265
function totalcollisions(d::DeviceQTrees, labels=nothing; itemlist=nothing, pairlist=nothing, queue=nothing, spqtree=nothing, linkedspqtree=nothing,
                         treenodestack=nothing, unique=true, kargs...)                   # dispatcher + _kw shims  qtree_functions.jl:266-276,331-338
    n = labels === nothing ? d.n : length(labels)
    n > SPACIAL_ENABLE_THRESHOLD ? totalcollisions_spacial(d, labels; unique=unique, kargs...) : totalcollisions_native(d, labels; kargs...)
end
function partialcollisions(d::DeviceQTrees, t::DeviceLinkedTree=locate!(d, linked_spacial_qtree(d)), labels=1:d.n;
                           colist=Vector{CoItem}(), unique=true, kargs...)               # qtree_functions.jl:277-330
    lb, nl = _lb(labels)
    _list(d, colist, (buf, cap, cnt) -> ccall((:stf_partialcollisions, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{StfItem}, Int64, Ptr{Int64}), d.h, nl, lb, Int32(unique), buf, cap, cnt))
end
function dynamiccollisions(d::DeviceQTrees, t::DeviceLinkedTree, moved::AbstractVector{Int}; unlocated::AbstractVector{Int}, colist=Vector{CoItem}(), kargs...)  # qtree_functions.jl:340-356
    if length(moved) / d.n > 0.6
        r = totalcollisions(d; colist=colist, kargs...)
        union!(unlocated, moved)
    else
        ms = Set(moved)
        locate!(d, [i for i in unlocated if !(i in ms)], t)
        empty!(unlocated)
        r = partialcollisions(d, t, moved; colist=colist, kargs...)
    end
    empty!(moved)
    union!(moved, Iterators.flatten(first.(r)))
    r
end
"DynamicColliders (qtree_functions.jl:357-370) over device trees"
struct DeviceColliders; qtrees::DeviceQTrees; spqtree::DeviceLinkedTree; moved::Vector{Int}; unlocated::Vector{Int}; end
DeviceColliders(d::DeviceQTrees) = DeviceColliders(d, linked_spacial_qtree(d), collect(1:d.n), Int[])
Base.union!(dc::DeviceColliders, c) = union!(dc.moved, c)
dynamiccollisions(dc::DeviceColliders; kargs...) = dynamiccollisions(dc.qtrees, dc.spqtree, dc.moved; unlocated=dc.unlocated, kargs...)

# ------------------------------------------------------------------------------------------------ fit step (fit_helper.jl, fit.jl:11-75)
function grad2d(t::DeviceTree, l::Integer, a::Integer, b::Integer)                        # fit_helper.jl:51-66
    out = Vector{Int32}(undef, 2)
    check(t.owner, ccall((:stf_grad2d, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                         t.owner.h, 1, Int32[t.label], Int32[l], Int32[a], Int32[b], out))
    (Int(out[1]), Int(out[2]))
end
_optkind(o::Momentum) = (Int32(2), Float64(o.eta), Float64(o.rho))
_optkind(o::SGD) = (Int32(1), Float64(o.eta), 0.0)
_optkind(::Any) = (Int32(0), 0.0, 0.0)                                                   # (t, Δ) -> Δ ./ 6  fit.jl:25
_setopt!(d::DeviceQTrees, o) = (k = _optkind(o); check(d, ccall((:stf_set_optimiser, LIB), Int32, (Ptr{Cvoid}, Int32, Float64, Float64), d.h, k[1], k[2], k[3])))
function reset!(o::Momentum, ts::AbstractVector{DeviceTree})                              # reset!.(optimiser, ts[repositioned])  fit.jl:508
    isempty(ts) && return
    lb = Int32[t.label for t in ts]
    check(ts[1].owner, ccall((:stf_reset_velocity, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}), ts[1].owner.h, length(lb), lb))
end
function batchsteps!(d::DeviceQTrees, colist::Vector{CoItem}, optimiser=(t, Δ) -> Δ ./ 6)  # fit.jl:69-75 (canonical mode: list order)
    _setopt!(d, optimiser)
    seed, it = d.clock
    check(d, ccall((:stf_batchsteps, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{StfItem}, UInt64, UInt64), d.h, length(colist), StfItem.(colist), seed, it))
    d.clock = (seed, it + 1)
    nothing
end
"one device-resident round: totalcollisions(qtrees[, labels]; unique=false) -> batchsteps!; returns length(colist)  (fit.jl:87-96)"
function fit_round!(d::DeviceQTrees, labels=nothing; optimiser=(t, Δ) -> Δ ./ 6, partial=false)
    _setopt!(d, optimiser)
    lb, nl = _lb(labels); nh = Ref{Int64}(0); seed, it = d.clock
    check(d, ccall((:stf_fit_round, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, UInt64, UInt64, Ptr{Int64}), d.h, partial ? 1 : 0, nl, lb, seed, it, nh))
    d.clock = (seed, it + 1)
    Int(nh[])
end
"the same round + getshift of every tree at `level` in the round's one synchronisation (stf_fit_round_positions)"
function fit_round_positions!(d::DeviceQTrees, labels=nothing; optimiser=(t, Δ) -> Δ ./ 6, partial=false, level=1)
    _setopt!(d, optimiser)
    lb, nl = _lb(labels); nh = Ref{Int64}(0); seed, it = d.clock
    rs = Vector{Int32}(undef, d.n); cs = Vector{Int32}(undef, d.n)
    check(d, ccall((:stf_fit_round_positions, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, UInt64, UInt64, Ptr{Int64}, Int32, Ptr{Int32}, Ptr{Int32}),
                   d.h, partial ? 1 : 0, nl, lb, seed, it, nh, level, rs, cs))
    d.clock = (seed, it + 1)
    Int(nh[]), collect(zip(Int.(rs), Int.(cs)))
end
"a round that returns `moved := the labels of its colist` (qtree_functions.jl:354-355; the trainers' `inds`, fit.jl:90) instead of the list:
the step of a dynamic loop is ONE call and one synchronisation (stf_fit_round_moved)"
function fit_round_moved!(d::DeviceQTrees, labels=nothing; optimiser=(t, Δ) -> Δ ./ 6, partial=true)
    _setopt!(d, optimiser)
    lb, nl = _lb(labels); nh = Ref{Int64}(0); nm = Ref{Int32}(0); seed, it = d.clock
    moved = Vector{Int32}(undef, d.n)
    check(d, ccall((:stf_fit_round_moved, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, UInt64, UInt64, Ptr{Int64}, Ptr{Int32}, Int32, Ptr{Int32}),
                   d.h, partial ? 1 : 0, nl, lb, seed, it, nh, moved, length(moved), nm))
    d.clock = (seed, it + 1)
    Int(nh[]), Int.(moved[1:nm[]])
end
"`k` rounds with ONE host synchronisation (stf_fit_rounds); returns the per-round length(colist) and the positions after the last round"
function fit_rounds!(d::DeviceQTrees, k::Integer, labels=nothing; optimiser=(t, Δ) -> Δ ./ 6, partial=false, level=1)
    _setopt!(d, optimiser)
    lb, nl = _lb(labels); nh = Vector{Int64}(undef, k); seed, it = d.clock
    rs = Vector{Int32}(undef, d.n); cs = Vector{Int32}(undef, d.n)
    check(d, ccall((:stf_fit_rounds, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, UInt64, UInt64, Int32, Ptr{Int64}, Ptr{Int64}, Int32, Ptr{Int32}, Ptr{Int32}),
                   d.h, partial ? 1 : 0, nl, lb, seed, it, k, nh, C_NULL, level, rs, cs))
    d.clock = (seed, it + k)
    Int.(nh), collect(zip(Int.(rs), Int.(cs)))
end

"filttrain!(qtrees, inpool, outpool, nearlevel2; optimiser)  fit.jl:235-266 — the pool test runs on the device; `inpool === nothing`
stands for ALL pairs i < j in the reference's order (fit.jl:279), generated on the device instead of being enumerated here"
function filttrain!(d::DeviceQTrees, inpool, outpool, nearlevel2; optimiser=(t, Δ) -> Δ ./ 6, kargs...)
    pairs = inpool === nothing ? C_NULL : Int32[x for p in inpool for x in p]
    np = inpool === nothing ? 0 : length(inpool)
    near = _list(d, Vector{CoItem}(), (buf, cap, cnt) -> ccall((:stf_filter_pairs, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Int32}, Int32, Ptr{StfItem}, Int64, Ptr{Int64}),
                                                                d.h, np, pairs, Int32(ceil(nearlevel2)), buf, cap, cnt))
    outpool === nothing || append!(outpool, first.(near))
    colist = CoItem[c for c in near if c.second[1] > 0]
    batchsteps!(d, colist, optimiser)
    length(colist)
end

"trainepoch_E!(qtrees; optimiser)  fit.jl:85-99 with the round loop on the device (stf_fit_epoch_e): `inds`, the round limit and the
stop rule `ni > 8length(colist)` are evaluated there; one host synchronisation per `chunk` rounds.  Small label sets (<= the
dispatcher's threshold) finish through the host dispatcher, as in the reference."
function trainepoch_E_device!(d::DeviceQTrees; optimiser=(t, Δ) -> Δ ./ 6, chunk::Integer=8, kargs...)
    _setopt!(d, optimiser)
    nh = Vector{Int64}(undef, chunk); state = Vector{Int32}(undef, 4); inds = Vector{Int32}(undef, d.n)
    first = 0; nc = 0
    while true
        rd = Ref{Int32}(0); seed, it = d.clock
        check(d, ccall((:stf_fit_epoch_e, LIB), Int32, (Ptr{Cvoid}, UInt64, UInt64, Int32, Int32, Ptr{Int64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int32),
                       d.h, seed, it, first, chunk, nh, rd, state, inds, length(inds)))
        first == 0 && rd[] > 0 && (nc = Int(nh[1]))
        (first == 0 && rd[] == 1 && nc == 0) || (d.clock = (seed, it + rd[]))
        first += rd[]
        (state[1] != 0 || rd[] == 0) && break
    end
    if state[1] == 3                                    # the native dispatcher path takes over for a small label set
        lb = Int.(inds[1:state[3]])
        for ni in state[2]+1:state[4]
            colist = totalcollisions(d, lb; unique=false)
            batchsteps!(d, colist, optimiser)
            ni > 8length(colist) && break
        end
    end
    nc
end

# ------------------------------------------------------------------------------------------------ trainers (fit.jl:77-233) — host policy over device rounds
# The reference's trainepoch_* functions only call totalcollisions / dynamiccollisions / batchsteps! on their `qtrees`
# argument, so they run UNCHANGED on a DeviceQTrees through the methods above; train!/fit! (fit.jl:450-552) likewise
# (its `resource` Dict is passed through; `:spqtree`/`:queue`/`:itemlist`/`:pairlist` entries are accepted and ignored by
# the keyword shims, `:moved` must be a DeviceColliders for trainepoch_D!).  Two resource constructors need a device twin:
trainepoch_D_resource(d::DeviceQTrees) = Dict(:colist => Vector{CoItem}(), :moved => DeviceColliders(d), :loops => 10)

# ------------------------------------------------------------------------------------------------ several GPUs from one process
"One context per device holding the same scene; rounds are split by candidate records, the GPUs exchange their hit lists
over NVLink peer memory themselves (include/stuffing_b200.h: stf_group_*)."
# ---- the rest of the boundary: narrow phase over an explicit item list, the linked tree's own operations, level shifts of
# one PaddedMat, the optimiser's velocities, ground composition for place!, counters
"_totalcollisions_native(qtrees, coitems, colist)  qtree_functions.jl:89-110: the narrow phase over an explicit item list"
function totalcollisions_native(d::DeviceQTrees, coitems::Vector{CoItem}, colist=Vector{CoItem}())
    items = StfItem.(coitems)
    _list(d, colist, (buf, cap, cnt) -> ccall((:stf_collide_items, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{StfItem}, Ptr{StfItem}, Int64, Ptr{Int64}), d.h, length(items), items, buf, cap, cnt))
end
"the itemlist of the last many-body call (the `itemlist` keyword of totalcollisions_spacial / partialcollisions)"
function itemlist(d::DeviceQTrees)
    _list(d, Vector{CoItem}(), (buf, cap, cnt) -> ccall((:stf_fetch_items, LIB), Int32, (Ptr{Cvoid}, Ptr{StfItem}, Int64, Ptr{Int64}), d.h, buf, cap, cnt))
end
keep_itemlist!(d::DeviceQTrees, on::Bool=true) = check(d, ccall((:stf_set_keep_items, LIB), Int32, (Ptr{Cvoid}, Int32), d.h, on ? 1 : 0))
"clear!(t::LinkedSpacialQTree, label) for a label list  qtrees.jl:538-545"
function clear!(t::DeviceLinkedTree, labels)
    lb, nl = _lb(labels)
    check(t.owner, ccall((:stf_linked_clear, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}), t.owner.h, nl, lb)); t
end
"collect_tree(t::LinkedSpacialQTree)  qtrees.jl:546-552 — (index, label) of every list node"
function collect_tree(t::DeviceLinkedTree)
    cnt = Ref{Int64}(0)
    ccall((:stf_linked_collect, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int64, Ptr{Int64}), t.owner.h, C_NULL, 0, cnt)
    buf = Vector{Int32}(undef, 4 * cnt[])                                    # stf_cellrec = (l, a, b, label)
    check(t.owner, ccall((:stf_linked_collect, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int64, Ptr{Int64}), t.owner.h, buf, cnt[], cnt))
    [(Int(buf[4i+1]), Int(buf[4i+2]), Int(buf[4i+3])) => Int(buf[4i+4]) for i in 0:cnt[]-1]
end
"setrshift!/setcshift!(t[l]::PaddedMat, …)  qtrees.jl:134-135: the shift of ONE level, no rebuild (the caller rebuilds with buildqtree!)"
setlevelshift!(t::DeviceTree, l::Integer, rs::Integer, cs::Integer) =
    check(t.owner, ccall((:stf_set_level_shift, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Int32), t.owner.h, t.label, l, rs, cs))
"the Momentum velocity of a list of trees (fit_helper.jl:19-33): (v, has) — what `optimiser.velocity[t]` holds"
function velocity(d::DeviceQTrees, labels=1:d.n)
    lb = Int32.(collect(labels)); v = Vector{Float64}(undef, 2 * length(lb)); has = Vector{UInt8}(undef, length(lb))
    check(d, ccall((:stf_get_velocity, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{UInt8}), d.h, length(lb), lb, v, has))
    [(v[2i-1], v[2i]) for i in 1:length(lb)], has .!= 0
end
"overlap!(ground, trees) of place!(ground, sortedtrees, indexes)  qtree_functions.jl:452-518: ground = the mask with every tree except
`excluded` composed into it, on the device; `groundlevel(d, l)` is the level the host-side findroom reads"
function compose_ground!(d::DeviceQTrees, mask_label::Integer=1, excluded=Int[])
    ex = Int32.(collect(excluded))
    check(d, ccall((:stf_ground_compose, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}), d.h, mask_label, length(ex), ex)); d
end
function groundlevel(d::DeviceQTrees, l::Integer)
    info = Vector{Int32}(undef, 5)
    check(d, ccall((:stf_ground_level, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{UInt8}, Ptr{Int32}), d.h, l, C_NULL, info))   # sizes first
    m = Matrix{UInt8}(undef, info[1], info[2])
    check(d, ccall((:stf_ground_level, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{UInt8}, Ptr{Int32}), d.h, l, m, info))
    m, (Int(info[3]), Int(info[4])), Int(info[5])                            # kernel, (rshift, cshift), canvas size of the level
end
"counters of the last many-body call (records, items, direct, hits, overflow, moved, visited)"
function counters(d::DeviceQTrees)
    c = Vector{Int64}(undef, 8)
    check(d, ccall((:stf_counters, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}), d.h, c))
    (records=c[1], items=c[2], direct=c[3], hits=c[4], overflow=c[5], moved=c[6], visited=c[7])
end

# ---- item-parallel rounds with ONE PROCESS PER GPU (MPI.jl / NCCL.jl own the all-gather; SURVEY §8e-2): collide over this rank's
# share, pack the local hits into a fixed-stride device buffer, (caller: all-gather send -> recv), step on the union
function shard_collide!(d::DeviceQTrees, rank::Integer, nranks::Integer, labels=nothing; partial=false)
    lb, nl = _lb(labels); nloc = Ref{Int64}(0)
    check(d, ccall((:stf_shard_collide, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Int32, Int32, Ptr{Int64}), d.h, partial ? 1 : 0, nl, lb, rank, nranks, nloc))
    Int(nloc[])
end
shard_pack!(d::DeviceQTrees, send_dev::Ptr{Cvoid}, stride::Integer) =
    check(d, ccall((:stf_shard_pack, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), d.h, send_dev, stride))
function shard_step!(d::DeviceQTrees, recv_dev::Ptr{Cvoid}, nranks::Integer, stride::Integer)
    nh = Ref{Int64}(0); seed, it = d.clock
    check(d, ccall((:stf_shard_step, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, UInt64, UInt64, Ptr{Int64}), d.h, recv_dev, nranks, stride, seed, it, nh))
    d.clock = (seed, it + 1)
    Int(nh[])
end

mutable struct GpuGroup
    h::Ptr{Cvoid}
    replicas::Vector{DeviceQTrees}
end
function GpuGroup(qts::AbstractVector, devices::AbstractVector{<:Integer}; seed::Integer=0)
    h = Ref{Ptr{Cvoid}}(C_NULL); dv = Int32.(collect(devices))
    rc = ccall((:stf_group_create, LIB), Int32, (Int32, Ptr{Int32}, Ptr{Ptr{Cvoid}}), length(dv), dv, h)
    rc == 0 || error("stf_group_create failed ($rc)")
    reps = DeviceQTrees[]
    for i in 0:length(dv)-1
        d = ccall(:jl_new_struct_uninit, Any, (Any,), DeviceQTrees)::DeviceQTrees    # adopt the group's context: no finalizer, the group owns it
        d.h = ccall((:stf_group_ctx, LIB), Ptr{Cvoid}, (Ptr{Cvoid}, Int32), h[], i); d.n = 0; d.L = 0; d.clock = (UInt64(seed), UInt64(0))
        push!(reps, upload!(d, qts))
    end
    g = GpuGroup(h[], reps)
    finalizer(x -> ccall((:stf_group_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.h), g)
end
"setshift! of the listed trees on EVERY replica (the GPUs driven in parallel)  qtrees.jl:276-288"
function setshift!(g::GpuGroup, labels, l::Integer, s1, s2)
    lb = Int32.(collect(labels)); rs = Int32.(collect(s1)); cs = Int32.(collect(s2))
    rc = ccall((:stf_group_set_shifts, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Int32}, Int32), g.h, length(lb), lb, l, rs, cs, 0)
    rc == 0 || error(unsafe_string(ccall((:stf_group_last_error, LIB), Cstring, (Ptr{Cvoid},), g.h))); g
end
function reset!(o::Momentum, g::GpuGroup, labels=nothing)
    lb, nl = _lb(labels)
    rc = ccall((:stf_group_reset_velocity, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}), g.h, nl, lb)
    rc == 0 || error(unsafe_string(ccall((:stf_group_last_error, LIB), Cstring, (Ptr{Cvoid},), g.h))); o
end
Base.length(g::GpuGroup) = Int(ccall((:stf_group_size, LIB), Int32, (Ptr{Cvoid},), g.h))
"bytes pushed to peer GPUs over NVLink by the last round"
nvlink_bytes(g::GpuGroup) = Int(ccall((:stf_group_nvlink_bytes, LIB), Int64, (Ptr{Cvoid},), g.h))
function fit_round!(g::GpuGroup, labels=nothing; optimiser=(t, Δ) -> Δ ./ 6, partial=false)
    foreach(d -> _setopt!(d, optimiser), g.replicas)
    lb, nl = _lb(labels); nh = Ref{Int64}(0); seed, it = g.replicas[1].clock
    rc = ccall((:stf_group_fit_round, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, UInt64, UInt64, Ptr{Int64}), g.h, partial ? 1 : 0, nl, lb, seed, it, nh)
    rc == 0 || error(unsafe_string(ccall((:stf_group_last_error, LIB), Cstring, (Ptr{Cvoid},), g.h)))
    foreach(d -> d.clock = (seed, it + 1), g.replicas)
    Int(nh[])
end
end # module
