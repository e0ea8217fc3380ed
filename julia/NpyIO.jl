# NpyIO.jl — minimal reader/writer of the NumPy .npy format 1.0 for the scene exchange of stuffing.jl_b200/io.py
# (SURVEY §8 f-4): UInt8 / Int32 arrays, little endian; matrices are written in Fortran order, i.e. Julia's own memory
# layout, so no transposition happens on either side.  NOT EXECUTED in the build image (Julia is absent there).
#
#   save_scene("dir", qtrees)            # the reference's Vector{U8SQTree}: level-1 payloads, defaults, level-1 shifts
#   shifts = load_shifts("dir")           # n x 2 Int32 (rshift, cshift) fitted on the GPU -> setshift!(qt, (rs, cs))
module NpyIO
export npywrite, npyread, save_scene, load_shifts, load_scene_codes

const MAGIC = UInt8[0x93, 'N', 'U', 'M', 'P', 'Y', 0x01, 0x00]
_descr(::Type{UInt8}) = "|u1"
_descr(::Type{Int32}) = "<i4"
function npywrite(path::AbstractString, a::Array{T}) where {T<:Union{UInt8,Int32}}
    shape = ndims(a) == 1 ? "($(length(a)),)" : "(" * join(size(a), ", ") * ")"
    hdr = "{'descr': '$(_descr(T))', 'fortran_order': $(ndims(a) > 1 ? "True" : "False"), 'shape': $shape, }"
    pad = 64 - (length(MAGIC) + 2 + length(hdr) + 1) % 64
    hdr *= " "^(pad % 64) * "\n"
    open(path, "w") do io
        write(io, MAGIC); write(io, UInt16(length(hdr))); write(io, hdr); write(io, a)
    end
end
function npyread(path::AbstractString)
    open(path) do io
        magic = read(io, 8); @assert magic[1:6] == MAGIC[1:6]
        hlen = magic[7] == 1 ? Int(read(io, UInt16)) : Int(read(io, UInt32))
        hdr = String(read(io, hlen))
        T = occursin("u1", hdr) ? UInt8 : occursin("<i4", hdr) ? Int32 : error("unsupported dtype in $path")
        fortran = occursin("'fortran_order': True", hdr)
        dims = parse.(Int, filter(!isempty, split(strip(match(r"'shape': \(([^)]*)\)", hdr).captures[1]), r",\s*")))
        if fortran || length(dims) <= 1
            a = Array{T}(undef, dims...); read!(io, a); a
        else                                   # C order: read reversed dims, then permute
            a = Array{T}(undef, reverse(dims)...); read!(io, a); permutedims(a, length(dims):-1:1)
        end
    end
end
"the reference's trees -> the directory layout of stuffing.jl_b200/io.py:save_scene"
function save_scene(dir::AbstractString, qts)
    mkpath(dir)
    npywrite(joinpath(dir, "shifts.npy"), Int32[getfield(qt[1], f) for qt in qts, f in (:rshift, :cshift)])
    npywrite(joinpath(dir, "defaults.npy"), UInt8[qt[1].default for qt in qts])
    for (i, qt) in enumerate(qts)
        k = qt[1].kernel
        npywrite(joinpath(dir, "tree_" * lpad(i, 5, '0') * ".npy"), Matrix{UInt8}(k[2:end-2, 2:end-2]))   # payload of the PaddedMat  qtrees.jl:117-128
    end
    canvas = isempty(qts) ? 2 : size(qts[1][1], 1)
    write(joinpath(dir, "meta.json"), "{\"format\": \"stuffing-scene-npy-1\", \"n\": $(length(qts)), \"canvas\": $canvas}")
end
load_shifts(dir::AbstractString) = npyread(joinpath(dir, "shifts.npy"))
load_scene_codes(dir::AbstractString, n::Integer) = [npyread(joinpath(dir, "tree_" * lpad(i, 5, '0') * ".npy")) for i in 1:n]
end # module
