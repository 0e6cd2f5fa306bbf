# graft.jl - Julia binding of libgraft.so's packing entry points (include/graft.h).
#
# Lives beside notebooks/julia_utils/*.jl in the reference tree; included by
# notebooks/Training/transformer.jl and notebooks/Finetune/transformer.jl:
#     include("../julia_utils/graft.jl")
# Replaces the element-by-element copy loop of `concat`
# (notebooks/Training/transformer.jl:144-163) and gives the packer the user-run
# offsets the trainer recomputes per file (notebooks/Training/transformer.py:53-68).
# Julia is not installed in the build image: the same three C functions are
# exercised through ctypes in tests/test_abi.py with identical arguments.
module Graft

const libgraft = get(ENV, "LIBGRAFT", "libgraft.so")

struct GraftError <: Exception
    msg::String
end

function check(rc::Cint, what::String)
    if rc != 0
        msg = unsafe_string(ccall((:graft_last_error, libgraft), Cstring, ()))
        throw(GraftError("$what failed ($rc): $msg"))
    end
end

"""
    jagged_offsets(userid::Vector{Int32}) -> Vector{Int64}

0-based run boundaries of equal `userid`: `offsets[1] == 0`, `offsets[end] == length(userid)`.
"""
function jagged_offsets(userid::Vector{Int32})
    n = length(userid)
    out = Vector{Int64}(undef, n + 1)
    nruns = Ref{Int64}(0)
    rc = ccall((:graft_jagged_offsets_host, libgraft), Cint,
               (Ptr{Int32}, Int64, Ptr{Int64}, Ref{Int64}), userid, n, out, nruns)
    check(rc, "graft_jagged_offsets_host")
    resize!(out, nruns[] + 1)
    out
end

"""
    block_permutation(offsets, order) -> Vector{Int64}

0-based index permutation that moves whole user runs in the order `order` (0-based run ids).
"""
function block_permutation(offsets::Vector{Int64}, order::Vector{Int64})
    perm = Vector{Int64}(undef, offsets[end])
    rc = ccall((:graft_block_permutation_host, libgraft), Cint,
               (Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}), offsets, length(order), order, perm)
    check(rc, "graft_block_permutation_host")
    perm
end

"""
    concat_pad(pieces::Vector{Vector{T}}, multiple) -> Vector{T}

`concat` for one key: pieces back to back, zero padded to a multiple of `multiple`.
"""
function concat_pad(pieces::Vector{Vector{T}}, multiple::Integer) where {T}
    n = sum(length, pieces)
    total = n % multiple == 0 ? n : n + multiple - n % multiple
    dst = Vector{T}(undef, total)
    ptrs = [pointer(p) for p in pieces]
    lens = Int64[length(p) for p in pieces]
    GC.@preserve pieces begin
        rc = ccall((:graft_concat_pad_host, libgraft), Cint,
                   (Ptr{Ptr{Cvoid}}, Ptr{Int64}, Int64, Int64, Ptr{Cvoid}, Int64),
                   ptrs, lens, length(pieces), sizeof(T), dst, total)
    end
    check(rc, "graft_concat_pad_host")
    dst
end

# drop-in for `concat(d)` of notebooks/Training/transformer.jl:144-163
function concat(d::Vector, batch_size::Integer)
    r = Dict()
    ks = collect(keys(d[1]))
    Threads.@threads for k in ks
        T = eltype(d[1][k])
        r[k] = concat_pad(Vector{T}[x[k] for x in d], batch_size)
    end
    r
end

end # module
