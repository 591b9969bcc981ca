# ImmersedBoundaryB200.jl -- thin Julia binding of libibx.so (include/ibx.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no `julia` (SURVEY.md F2).  The file is kept tiny and
# mechanical so that it can be reviewed against include/ibx.h; every call below is a one-line `ccall` of an entry
# point that the Python harness (immersedboundary.jl_b200/_lib.py) exercises with the same arguments.
#
# It keeps the reference package's names (src/ImmersedBoundary.jl:34-38, :873-1409) so that user closures written
# for ImmersedBoundary.jl run unchanged on `IBXArray`s:
#
#     using ImmersedBoundaryB200
#     dom = Domain(msh; hypercube_families = ["farfield" => [(1, false), (1, true)]])
#     dom(u, ud) do part, u, ud
#         ∇u = cell_gradient(part, u, 1)
#         uL, uR = MUSCL(part, u, ∇u, 1; D = JST_sensor(part, u))
#         ud .-= green_gauss(part, uL, 1)
#     end
#
# Index conventions: the library is 0-based (`dim`, cell ids); this wrapper subtracts/adds 1 at the call.
module ImmersedBoundaryB200

export Stereolitography, merge_points, feature_regions, DistanceField, Ball, Box, Line, Mesh, Domain,
       at_owners, at_neighbors, at_faces, green_gauss, unsigned_green_gauss, cell_gradient, face_distance,
       owner_distance, neighbor_distance, face_gradient, JST_sensor, MUSCL, impose_bc!, multigrid, volume_integral,
       Fluid, FlowBC, state2primitive, primitive2state, speed_of_sound, inviscid_fluxes,
       residual_euler!, ghost_update_euler!, Transport, dynamic_viscosity, heat_conductivity, viscous_fluxes,
       shock_sensor, shear_rate, Ducros_sensor, pressure_coefficient, Accumulator, Interpolator, multigrid, FAS!,
       euler_step_host!, euler_step_host_end, step_euler!, march_euler!, local_step_update!, halo_begin!, halo_end!, wall_function, Smagorinsky_νSGS, standard_kϵ,
       Wray_Agarwal, WALE_νSGS

const libibx = get(ENV, "LIBIBX", joinpath(@__DIR__, "..", "libibx.so"))

struct IBXError <: Exception
    msg::String
end
Base.showerror(io::IO, e::IBXError) = print(io, "libibx: ", e.msg)

"Turn a non-zero status into a Julia exception (the reference throws plain exceptions, src/cfd.jl:289)."
@inline function check(status::Cint)
    status == 0 || throw(IBXError(unsafe_string(ccall((:ibx_last_error, libibx), Cstring, ()))))
    nothing
end

# ------------------------------------------------------------------ context + device arrays
const CTX = Ref{Ptr{Cvoid}}(C_NULL)
function context(device::Integer = 0)
    if CTX[] == C_NULL
        check(ccall((:ibx_init, libibx), Cint, (Cint, Ref{Ptr{Cvoid}}), device, CTX))
        atexit(() -> ccall((:ibx_finalize, libibx), Cint, (Ptr{Cvoid},), CTX[]))
    end
    CTX[]
end

"Device-resident column-major Float32 array: what `conv_to_backend` returns (src/ImmersedBoundary.jl:846-849)."
mutable struct IBXArray{N} <: AbstractArray{Float32, N}
    h::Int64
    dims::NTuple{N, Int}
    function IBXArray{N}(dims::NTuple{N, Int}) where {N}
        h = Ref{Int64}(0)
        check(ccall((:ibx_array_alloc, libibx), Cint, (Ptr{Cvoid}, Int64, Int64, Ref{Int64}),
                    context(), dims[1], N == 1 ? 1 : prod(dims[2:end]), h))
        a = new{N}(h[], dims)
        finalizer(x -> ccall((:ibx_array_free, libibx), Cint, (Ptr{Cvoid}, Int64), CTX[], x.h), a)
    end
end
Base.size(a::IBXArray) = a.dims
IBXArray(a::Array{Float32, N}) where {N} = (d = IBXArray{N}(size(a)); GC.@preserve a check(ccall(
    (:ibx_array_upload, libibx), Cint, (Ptr{Cvoid}, Int64, Ptr{Float32}), context(), d.h, a)); d)
Base.Array(d::IBXArray{N}) where {N} = (a = Array{Float32, N}(undef, d.dims); GC.@preserve a check(ccall(
    (:ibx_array_download, libibx), Cint, (Ptr{Cvoid}, Int64, Ptr{Float32}), context(), d.h, a)); a)
Base.similar(d::IBXArray{N}) where {N} = IBXArray{N}(d.dims)
ncols(d::IBXArray) = length(d.dims) == 1 ? 1 : prod(d.dims[2:end])

# elementwise glue: broadcasting `a .+ b`, `a .* s`, `abs.(a)`, `max.(a, b)` lowers to ibx_ew_* kernels
const _BINOPS = Dict(:+ => 0, :- => 1, :* => 2, :/ => 3, :max => 4, :min => 5)
for (f, op) in _BINOPS
    @eval function Base.broadcasted(::typeof($f), a::IBXArray, b::IBXArray)
        out = ncols(a) >= ncols(b) ? similar(a) : similar(b)
        check(ccall((:ibx_ew_binary, libibx), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Int64), context(), $op, a.h, b.h, out.h)); out
    end
    @eval function Base.broadcasted(::typeof($f), a::IBXArray, s::Real)
        out = similar(a)
        check(ccall((:ibx_ew_scalar, libibx), Cint, (Ptr{Cvoid}, Cint, Int64, Cfloat, Cint, Int64), context(), $op, a.h, s, 0, out.h)); out
    end
    @eval function Base.broadcasted(::typeof($f), s::Real, a::IBXArray)
        out = similar(a)
        check(ccall((:ibx_ew_scalar, libibx), Cint, (Ptr{Cvoid}, Cint, Int64, Cfloat, Cint, Int64), context(), $op, a.h, s, 1, out.h)); out
    end
end
for (f, op) in ((:abs, 0), (:-, 1), (:sqrt, 2), (:sign, 3), (:inv, 4))
    @eval function Base.broadcasted(::typeof($f), a::IBXArray)
        out = similar(a)
        check(ccall((:ibx_ew_unary, libibx), Cint, (Ptr{Cvoid}, Cint, Int64, Int64), context(), $op, a.h, out.h)); out
    end
end
Base.copyto!(dst::IBXArray, src::IBXArray) = (check(ccall((:ibx_array_copy, libibx), Cint,
    (Ptr{Cvoid}, Int64, Int64), context(), dst.h, src.h)); dst)
Base.materialize!(dst::IBXArray, src::IBXArray) = copyto!(dst, src)
Base.fill!(a::IBXArray, v::Real) = (check(ccall((:ibx_array_fill, libibx), Cint, (Ptr{Cvoid}, Int64, Cfloat), context(), a.h, v)); a)
function _reduce(op::Integer, a::IBXArray)
    out = Ref{Float64}(0)
    check(ccall((:ibx_reduce, libibx), Cint, (Ptr{Cvoid}, Cint, Int64, Cint, Ref{Float64}), context(), op, a.h, 0, out)); out[]
end
Base.sum(a::IBXArray) = Float32(_reduce(0, a))
Base.maximum(a::IBXArray) = Float32(_reduce(1, a))
Base.minimum(a::IBXArray) = Float32(_reduce(2, a))
import LinearAlgebra
LinearAlgebra.norm(a::IBXArray) = Float32(sqrt(_reduce(4, a)))

# ------------------------------------------------------------------ geometry + mesh (src/mesher.jl)
mutable struct Stereolitography
    h::Ptr{Cvoid}
end
function Stereolitography(fname::String)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ibx_stl_read, libibx), Cint, (Cstring, Ref{Ptr{Cvoid}}), fname, h))
    Stereolitography(h[])
end
function Stereolitography(points::AbstractMatrix; closed::Bool = true)       # points: (nd, npoints) like the reference
    n = size(points, 2)
    simp = closed ? [collect(0:n-1)'; circshift(collect(0:n-1), -1)'] : [collect(0:n-2)'; collect(1:n-1)']
    Stereolitography(points, simp .+ 1)
end
function Stereolitography(points::AbstractMatrix, simplices::AbstractMatrix)
    p = Float64.(permutedims(points)) |> permutedims      # (nd, np) column-major == np x nd row-major
    s = Int64.(simplices .- 1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ibx_stl_create, libibx), Cint, (Cint, Int64, Ptr{Float64}, Int64, Ptr{Int64}, Cint, Ref{Ptr{Cvoid}}),
                size(points, 1), size(points, 2), p, size(s, 2), s, eltype(points) == Float32, h))
    Stereolitography(h[])
end
function merge_points(stls::Stereolitography...; tolerance::Real = 1e-7, clean_degenerate::Bool = true)
    hs = [s.h for s in stls]; out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ibx_stl_merge_points, libibx), Cint, (Cint, Ptr{Ptr{Cvoid}}, Float64, Cint, Cint, Ref{Ptr{Cvoid}}),
                length(hs), hs, tolerance, tolerance isa Float32, clean_degenerate, out))
    Stereolitography(out[])
end
function feature_regions(stl::Stereolitography; angle::Real = 15.0, radius::Real = Inf64, include_boundaries::Bool = false)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ibx_stl_feature_regions, libibx), Cint, (Ptr{Cvoid}, Float64, Float64, Cint, Ref{Ptr{Cvoid}}),
                stl.h, angle, radius, include_boundaries, out))
    Stereolitography(out[])
end
mutable struct DistanceField
    h::Ptr{Cvoid}
end
function DistanceField(stl::Stereolitography)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ibx_dfield_create, libibx), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), stl.h, out)); DistanceField(out[])
end
struct Ball; center::Vector{Float64}; radius::Float64; end
struct Box; origin::Vector{Float64}; widths::Vector{Float64}; end
struct Line; p1::Vector{Float64}; p2::Vector{Float64}; end

"Mirror of `ibx_region` (include/ibx.h)."
struct CRegion
    kind::Cint; h_is_f32::Cint; c::NTuple{3, Float64}; a::NTuple{3, Float64}; dfield::Ptr{Cvoid}; h::Float64
end
_t3(v) = ntuple(i -> i <= length(v) ? Float64(v[i]) : 0.0, 3)
region(b::Ball, h) = CRegion(0, h isa Float32, _t3(b.center), (b.radius, 0.0, 0.0), C_NULL, h)
region(b::Box, h) = CRegion(1, h isa Float32, _t3(b.origin), _t3(b.widths), C_NULL, h)
region(l::Line, h) = CRegion(2, h isa Float32, _t3(l.p1), _t3(l.p2), C_NULL, h)
region(d::DistanceField, h) = CRegion(3, h isa Float32, _t3(()), _t3(()), d.h, h)

"Mirror of `ibx_surface` (include/ibx.h)."
struct CSurface
    name::Cstring; stl::Ptr{Cvoid}; h::Float64; h_is_f32::Cint; sphere_c::NTuple{3, Float64}; sphere_r::Float64
end

mutable struct Mesh
    h::Ptr{Cvoid}
    block_size::Int32
    nd::Int
    ncells::Int
end
"`Mesh(origin, widths, (name, stl, h)...; refinement_regions, growth_ratio, block_size)` (src/mesher.jl:972-1046)"
function Mesh(origin::AbstractVector, widths::AbstractVector, surfaces::Tuple...;
              growth_ratio::Real = 2.0f0, tolerance::Real = 1f-7, block_size::Int = 8,
              refinement_regions::AbstractVector = [], verbose::Bool = false)
    names = [String(s[1]) for s in surfaces]
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve names begin
        surf = [CSurface(Base.unsafe_convert(Cstring, names[i]), s[2].h, s[3], s[3] isa Float32, (0.0, 0.0, 0.0), 0.0)
                for (i, s) in enumerate(surfaces)]
        regs = CRegion[region(first(r), last(r)) for r in refinement_regions]
        check(ccall((:ibx_mesh_create, libibx), Cint,
                    (Cint, Ptr{Float32}, Ptr{Float32}, Cint, Ptr{CSurface}, Cint, Ptr{CRegion}, Float64, Float64, Cint, Cint,
                     Ref{Ptr{Cvoid}}),
                    length(origin), Float32.(origin), Float32.(widths), length(surf), surf, length(regs), regs,
                    growth_ratio, tolerance, tolerance isa Float32, block_size, out))
    end
    nd = Ref{Cint}(0); bs = Ref{Cint}(0); nb = Ref{Int64}(0); nc = Ref{Int64}(0); ns = Ref{Cint}(0)
    check(ccall((:ibx_mesh_info, libibx), Cint, (Ptr{Cvoid}, Ref{Cint}, Ref{Cint}, Ref{Int64}, Ref{Int64}, Ref{Cint}),
                out[], nd, bs, nb, nc, ns))
    Mesh(out[], bs[], nd[], nc[])
end
Base.length(m::Mesh) = m.ncells

# ------------------------------------------------------------------ Domain + partition runtime
struct Partition
    dom::Any
    p::Int                      # 0-based partition index
    n_domain::Int
    n_image::Int
    nfaces::Vector{Int}
end
Base.ndims(part::Partition) = part.dom.nd

mutable struct Domain
    h::Ptr{Cvoid}
    nd::Int
    ncells::Int
    partitions::Dict{Int64, Partition}
    boundary_index::Dict{String, Int}
    boundary_parts::Dict{String, Int}
    mesh::Mesh
end
Base.length(d::Domain) = d.ncells
Base.ndims(d::Domain) = d.nd

"`Domain(msh; max_partition_size, partition_skirt_depth, ghost_layer_ratio, hypercube_families)` (src/ImmersedBoundary.jl:536-786)"
function Domain(msh::Mesh; max_partition_size::Int = 100_000, partition_skirt_depth::Int = 2,
                ghost_layer_ratio::Real = 1.5f0, hypercube_families = [], verbose::Bool = false)
    names = [String(first(f)) for f in hypercube_families]
    fptr = Cint[0]; dims = Cint[]; fronts = Cint[]
    for f in hypercube_families
        for (d, front) in last(f); push!(dims, d - 1); push!(fronts, front); end
        push!(fptr, length(dims))
    end
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ibx_domain_build, libibx), Cint,
                (Ptr{Cvoid}, Int64, Cint, Cfloat, Cint, Ptr{Cstring}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Cint, Cint, Ref{Ptr{Cvoid}}),
                msh.h, max_partition_size, partition_skirt_depth, ghost_layer_ratio, length(names), names, fptr, dims, fronts, 1, 1, out))
    check(ccall((:ibx_domain_upload, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), context(), out[]))
    nd = Ref{Cint}(0); nc = Ref{Int64}(0); nf = Ref{Int64}(0); np = Ref{Cint}(0); nb = Ref{Cint}(0); ns = Ref{Cint}(0)
    check(ccall((:ibx_domain_info, libibx), Cint, (Ptr{Cvoid}, Ref{Cint}, Ref{Int64}, Ref{Int64}, Ref{Cint}, Ref{Cint}, Ref{Cint}),
                out[], nd, nc, nf, np, nb, ns))
    dom = Domain(out[], nd[], nc[], Dict{Int64, Partition}(), Dict{String, Int}(), Dict{String, Int}(), msh)
    for p = 0:(np[] - 1)
        a = Ref{Int64}(0); b = Ref{Int64}(0); s = Ref{Int64}(0); nfd = zeros(Int64, nd[])
        check(ccall((:ibx_partition_info, libibx), Cint, (Ptr{Cvoid}, Cint, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ptr{Int64}),
                    out[], p, a, b, s, nfd))
        dom.partitions[p + 1] = Partition(dom, p, a[], b[], nfd)
    end
    for b = 0:(nb[] - 1)
        nm = Ref{Cstring}(C_NULL); k = Ref{Cint}(0)
        check(ccall((:ibx_boundary_name, libibx), Cint, (Ptr{Cvoid}, Cint, Ref{Cstring}, Ref{Cint}), out[], b, nm, k))
        dom.boundary_index[unsafe_string(nm[])] = b; dom.boundary_parts[unsafe_string(nm[])] = k[]
    end
    dom
end

"`dom(f, args...)` (src/ImmersedBoundary.jl:820-864): gather -> f -> scatter of the image rows, all on the device."
function (dom::Domain)(f, args::AbstractArray{Float32}...; kwargs...)
    gl = map(a -> a isa IBXArray ? a : IBXArray(Array(a)), args)
    res = map(sort(collect(keys(dom.partitions)))) do i
        part = dom.partitions[i]
        dargs = map(gl) do g
            d = IBXArray{ndims(g)}((part.n_domain, size(g)[2:end]...))
            check(ccall((:ibx_gather_domain, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Int64), context(), dom.h, part.p, g.h, d.h)); d
        end
        r = f(part, dargs...; kwargs...)
        for (g, d) in zip(gl, dargs)
            check(ccall((:ibx_scatter_image, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Int64), context(), dom.h, part.p, d.h, g.h))
        end
        r
    end
    for (a, g) in zip(args, gl); a isa IBXArray || copyto!(a, Array(g)); end
    res
end

# ------------------------------------------------------------------ grid operators (src/ImmersedBoundary.jl:879-1157)
_out(part, rows, u) = IBXArray{ndims(u)}((rows, size(u)[2:end]...))
macro op(jname, cname, rows)
    quote
        function $(esc(jname))(part::Partition, u::IBXArray, dim::Int)
            out = _out(part, $(esc(rows))(part, dim), u)
            check(ccall(($(QuoteNode(cname)), libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64),
                        context(), part.dom.h, part.p, dim - 1, u.h, out.h)); out
        end
    end
end
_nf(part, dim) = part.nfaces[dim]
_nc(part, dim) = part.n_domain
@op at_owners ibx_at_owners _nf
@op at_neighbors ibx_at_neighbors _nf
@op at_faces ibx_at_faces _nf
@op green_gauss ibx_green_gauss _nc
@op unsigned_green_gauss ibx_unsigned_green_gauss _nc
@op cell_gradient ibx_cell_gradient _nc
@op face_gradient ibx_face_gradient _nf
cell_gradient(part::Partition, u::IBXArray) = tuple((cell_gradient(part, u, d) for d = 1:ndims(part))...)
for (jn, cn) in ((:face_distance, :ibx_face_distance), (:owner_distance, :ibx_owner_distance), (:neighbor_distance, :ibx_neighbor_distance))
    @eval function $jn(part::Partition, dim::Int)
        out = IBXArray{1}((part.nfaces[dim],))
        check(ccall(($(QuoteNode(cn)), libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64), context(), part.dom.h, part.p, dim - 1, out.h)); out
    end
end
function JST_sensor(part::Partition, p::IBXArray, dim::Int = 0)
    out = similar(p)
    check(ccall((:ibx_jst_sensor, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64), context(), part.dom.h, part.p, dim - 1, p.h, out.h)); out
end
function MUSCL(part::Partition, u::IBXArray, δu::IBXArray, dim::Int; D::Union{IBXArray{1}, Nothing} = nothing, high_order::Bool = false)
    uL = _out(part, part.nfaces[dim], u); uR = _out(part, part.nfaces[dim], u)
    check(ccall((:ibx_muscl, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64, Cint, Int64, Int64),
                context(), part.dom.h, part.p, dim - 1, u.h, δu.h, isnothing(D) ? 0 : D.h, high_order, uL.h, uR.h))
    (uL, uR)
end

# ------------------------------------------------------------------ impose_bc! (src/ImmersedBoundary.jl:1197-1247)
struct Boundary
    dom::Domain
    b::Int
    part::Int
    nghost::Int
    normals::IBXArray{2}
end
function impose_bc!(f, dom::Domain, bname::String, args::AbstractArray...; kwargs...)
    b = dom.boundary_index[bname]
    gl = map(a -> a isa IBXArray ? a : IBXArray(Array(a)), args)
    pending = []
    for k = 0:(dom.boundary_parts[bname] - 1)
        g = Ref{Int64}(0); m = Ref{Int64}(0); z = Ref{Int64}(0)
        check(ccall((:ibx_boundary_info, libibx), Cint, (Ptr{Cvoid}, Cint, Cint, Ref{Int64}, Ref{Int64}, Ref{Int64}), dom.h, b, k, g, m, z))
        nrm = IBXArray{2}((Int(g[]), dom.nd))
        check(ccall((:ibx_bc_normals, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64), context(), dom.h, b, k, nrm.h))
        bd = Boundary(dom, b, k, g[], nrm)
        iargs = map(gl) do a
            ia = IBXArray{ndims(a)}((Int(g[]), size(a)[2:end]...))
            check(ccall((:ibx_bc_image_values, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64), context(), dom.h, b, k, a.h, ia.h)); ia
        end
        r = f(bd, iargs...; kwargs...)
        r isa Tuple || (r = (r,))
        push!(pending, (k, iargs, r))
    end
    for (k, iargs, r) in pending, (a, ba, ia) in zip(gl, r, iargs)   # Jacobi: all reads above precede all writes
        if ba isa Real
            check(ccall((:ibx_bc_blend_scalar, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64, Cfloat), context(), dom.h, b, k, a.h, ia.h, ba))
        else
            dba = ba isa IBXArray ? ba : IBXArray(Float32.(ba))
            check(ccall((:ibx_bc_blend, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64), context(), dom.h, b, k, a.h, ia.h, dba.h))
        end
    end
    for (a, g) in zip(args, gl); a isa IBXArray || copyto!(a, Array(g)); end
end

function volume_integral(dom::Domain, A::AbstractArray)
    dA = A isa IBXArray ? A : IBXArray(Float32.(A))
    out = zeros(Float32, ncols(dA))
    check(ccall((:ibx_volume_integral, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Float32}), context(), dom.h, dA.h, out))
    ndims(A) == 1 ? out[1] : out
end

# ------------------------------------------------------------------ cfd.jl subset + fused entry points
struct Fluid; R::Float32; γ::Float32; end
Fluid(; R::Real = 283.0f0, γ::Real = 1.4f0, kwargs...) = Fluid(R, γ)
struct FlowBC; fluid::Fluid; P::Vector{Float32}; normal_flow::Bool; end
FlowBC(fluid::Fluid, P::AbstractVector; normal_flow::Bool = false) = FlowBC(fluid, Float32.(P), normal_flow)
function (bc::FlowBC)(P::IBXArray, normals::IBXArray; image_distances::Union{Nothing, IBXArray} = nothing,
                      du!dn::Union{Nothing, IBXArray} = nothing, transpiration::Union{Real, IBXArray} = 0.0f0)
    out = similar(P)
    h(x) = x isa IBXArray ? x.h : Int64(0)
    check(ccall((:ibx_flowbc_ex, libibx), Cint, (Ptr{Cvoid}, Fluid, Ptr{Float32}, Cint, Cint, Int64, Int64, Int64, Int64, Int64, Cfloat, Int64),
                context(), bc.fluid, bc.P, length(bc.P), bc.normal_flow, P.h, normals.h, h(image_distances), h(du!dn), h(transpiration),
                transpiration isa Real ? Float32(transpiration) : 0.0f0, out.h)); out
end
for (jn, cn) in ((:state2primitive, :ibx_state2primitive), (:primitive2state, :ibx_primitive2state), (:speed_of_sound, :ibx_speed_of_sound))
    @eval function $jn(fluid::Fluid, A::IBXArray)
        out = similar(A)
        check(ccall(($(QuoteNode(cn)), libibx), Cint, (Ptr{Cvoid}, Fluid, Int64, Int64), context(), fluid, A.h, out.h)); out
    end
end
function inviscid_fluxes(fluid::Fluid, PL::IBXArray, PR::IBXArray, dim::Int)
    F = similar(PL)
    check(ccall((:ibx_inviscid_fluxes_hll, libibx), Cint, (Ptr{Cvoid}, Fluid, Int64, Int64, Cint, Int64), context(), fluid, PL.h, PR.h, dim - 1, F.h)); F
end
function inviscid_fluxes(fluid::Fluid, PL::IBXArray, PR::IBXArray, νL::IBXArray, νR::IBXArray, dim::Int)
    F = similar(PL)
    check(ccall((:ibx_inviscid_fluxes_sensor, libibx), Cint, (Ptr{Cvoid}, Fluid, Int64, Int64, Int64, Int64, Cint, Int64),
                context(), fluid, PL.h, PR.h, νL.h, νR.h, dim - 1, F.h)); F
end

"Whole-domain fused Euler residual (block-structured kernels): `R, cfl` from the conservative state `Q`."
residual_euler!(dom::Domain, fluid::Fluid, Q::IBXArray, R::IBXArray, cfl::IBXArray; flux_kind::Int = 0) = check(ccall(
    (:ibx_residual_euler, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Fluid, Cint, Int64, Int64, Int64), context(), dom.h, fluid, flux_kind, Q.h, R.h, cfl.h))
"Fused IB ghost update of `Q` for boundary `bname` with `bc`."
ghost_update_euler!(dom::Domain, fluid::Fluid, bname::String, bc::FlowBC, Q::IBXArray) = check(ccall(
    (:ibx_ghost_update_euler, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Fluid, Ptr{Float32}, Cint, Cint, Int64),
    context(), dom.h, dom.boundary_index[bname], fluid, bc.P, length(bc.P), bc.normal_flow, Q.h))

# ------------------------------------------------------------------ transport, viscous fluxes, pointwise sensors (csrc/closures.cu)
"Sutherland / conductivity constants as `ibx_transport` (isbits, passed by value)."
struct Transport; μref::Float32; Tref::Float32; S::Float32; nk::Cint; k::NTuple{4, Float32}; end
Transport(; μref = 1.716f-5, Tref = 273.15f0, S = 110.4f0, k = (0.00646f0, 6.468f-5)) =
    Transport(μref, Tref, S, length(k), ntuple(i -> i <= length(k) ? Float32(k[i]) : 0.0f0, 4))
for (jn, cn) in ((:dynamic_viscosity, :ibx_dynamic_viscosity), (:heat_conductivity, :ibx_heat_conductivity))
    @eval function $jn(t::Transport, T::IBXArray)
        out = similar(T)
        check(ccall(($(QuoteNode(cn)), libibx), Cint, (Ptr{Cvoid}, Transport, Int64, Int64), context(), t, T.h, out.h)); out
    end
end
"`viscous_fluxes(fluid, P, Pgrad, dim; μₜ)` (src/cfd.jl:664-736); `dim::Int` (1-based) or an N x nd direction matrix."
function viscous_fluxes(t::Transport, P::IBXArray, Pgrad, dim::Union{Int, IBXArray}; μₜ::Union{Real, IBXArray} = 0.0f0)
    F = similar(P); hs = Int64[g.h for g in Pgrad]
    GC.@preserve hs check(ccall((:ibx_viscous_fluxes, libibx), Cint, (Ptr{Cvoid}, Transport, Int64, Ptr{Int64}, Cint, Int64, Int64, Cfloat, Int64),
        context(), t, P.h, hs, dim isa Int ? dim - 1 : -1, dim isa Int ? 0 : dim.h, μₜ isa Real ? 0 : μₜ.h, μₜ isa Real ? Float32(μₜ) : 0.0f0, F.h)); F
end
function JST_sensor(Pim1::IBXArray, Pi::IBXArray, Pip1::IBXArray)
    out = similar(Pi)
    check(ccall((:ibx_jst_sensor3, libibx), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Int64), context(), Pim1.h, Pi.h, Pip1.h, out.h)); out
end
"velocity_gradients[i, j] = ∂uᵢ/∂xⱼ (matrix of vectors) -> row-major handle table"
_grad_table(g::AbstractMatrix) = Int64[g[i, j].h for i in axes(g, 1) for j in axes(g, 2)]
for (jn, cn) in ((:shock_sensor, :ibx_shock_sensor), (:shear_rate, :ibx_shear_rate), (:Ducros_sensor, :ibx_ducros_sensor))
    @eval function $jn(g::AbstractMatrix)
        out = similar(g[1, 1]); hs = _grad_table(g)
        GC.@preserve hs check(ccall(($(QuoteNode(cn)), libibx), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Int64), context(), size(g, 1), hs, out.h)); out
    end
end
function pressure_coefficient(fluid::Fluid, p::IBXArray, p∞::Real, M∞::Real)
    out = similar(p)
    check(ccall((:ibx_pressure_coefficient, libibx), Cint, (Ptr{Cvoid}, Cfloat, Int64, Cfloat, Cfloat, Int64), context(), fluid.γ, p.h, p∞, M∞, out.h)); out
end
# ---- src/turbulence.jl
"Mirror of `ibx_wall_params` (include/ibx.h): the keyword constants of `wall_function` (src/turbulence.jl:27-33)."
struct WallParams; κ::Float32; C::Float32; A::Float32; β::Float32; βstar::Float32; D::Float32; A⁺::Float32; ω::Float32; n_iter::Cint; end
WallParams(; κ = 0.41f0, C = 4.9f0, A = 19.0f0, β = 0.075f0, βstar = 0.09f0, D = 4.2f0, A⁺ = 360.0f0, ω_fixed_point = 0.5f0, n_iter = 20) =
    WallParams(κ, C, A, β, βstar, D, A⁺, ω_fixed_point, n_iter)
function wall_function(Rey::IBXArray; kwargs...)
    o = ntuple(_ -> similar(Rey), 5)
    check(ccall((:ibx_wall_function_rey, libibx), Cint, (Ptr{Cvoid}, WallParams, Int64, Int64, Int64, Int64, Int64, Int64),
                context(), WallParams(; kwargs...), Rey.h, o[1].h, o[2].h, o[3].h, o[4].h, o[5].h))
    (y⁺ = o[1], u⁺ = o[2], μ⁺ = o[3], k⁺ = o[4], du⁺!dy⁺ = o[5])
end
function wall_function(y::IBXArray, u::IBXArray, ν::IBXArray; kwargs...)
    o = ntuple(_ -> similar(y), 6)
    check(ccall((:ibx_wall_function, libibx), Cint, (Ptr{Cvoid}, WallParams, Int64, Int64, Int64, Int64, Int64, Int64, Int64, Int64, Int64),
                context(), WallParams(; kwargs...), y.h, u.h, ν.h, o[1].h, o[2].h, o[3].h, o[4].h, o[5].h, o[6].h))
    (uτ = o[1], νₜ = o[2], k = o[3], ω = o[4], ϵ = o[5], du!dn = o[6])
end
function Smagorinsky_νSGS(Δ::IBXArray, S::IBXArray; Cₛ::Real = 0.17f0)
    out = similar(S)
    check(ccall((:ibx_smagorinsky, libibx), Cint, (Ptr{Cvoid}, Int64, Int64, Cfloat, Int64), context(), Δ.h, S.h, Cₛ, out.h)); out
end
function standard_kϵ(k::IBXArray, ϵ::IBXArray, S::IBXArray; Cμ = 0.09f0, σk = 1.0f0, σϵ = 1.3f0, C1ϵ = 1.44f0, C2ϵ = 1.92f0)
    o = ntuple(_ -> similar(k), 5)
    check(ccall((:ibx_standard_keps, libibx), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Cfloat, Cfloat, Cfloat, Cfloat, Cfloat, Int64, Int64, Int64, Int64, Int64),
                context(), k.h, ϵ.h, S.h, Cμ, σk, σϵ, C1ϵ, C2ϵ, o[1].h, o[2].h, o[3].h, o[4].h, o[5].h))
    (νk = o[1], νϵ = o[2], Sk = o[3], Sϵ = o[4], νₜ = o[5])
end
function Wray_Agarwal(R::IBXArray, S::IBXArray, ∇R::IBXArray, ∇S::IBXArray; σR = 0.72f0, C₁ = 0.0829f0, κ = 0.41f0)
    νR = similar(R); So = similar(R)
    check(ccall((:ibx_wray_agarwal, libibx), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Cfloat, Cfloat, Cfloat, Int64, Int64),
                context(), R.h, S.h, ∇R.h, ∇S.h, σR, C₁, κ, νR.h, So.h))
    (νₜ = R, νR = νR, S = So)
end
function WALE_νSGS(Δ::IBXArray, g::AbstractMatrix; Cw::Real = 0.325f0)
    @assert size(g, 1) == 3 "WALE model only implemented for 3D"
    out = similar(Δ); hs = _grad_table(g)
    GC.@preserve hs check(ccall((:ibx_wale, libibx), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Cfloat, Int64), context(), Δ.h, hs, Cw, out.h)); out
end

# ------------------------------------------------------------------ accumulators, multigrid, FAS!, host-buffer evaluation, halo exchange
"Device-resident `Accumulator` (src/accumulator.jl:12-130): CSR tables owned by the library; `acc(v)` is one kernel."
mutable struct Accumulator
    h::Ptr{Cvoid}
    n_output::Int
    uploaded::Bool
end
function Accumulator(h::Ptr{Cvoid})
    n = Ref{Int64}(0); nnz = Ref{Int64}(0); w = Ref{Cint}(0)
    check(ccall((:ibx_accum_info, libibx), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Cint}), h, n, nnz, w))
    acc = Accumulator(h, n[], false)
    finalizer(a -> ccall((:ibx_accum_free, libibx), Cint, (Ptr{Cvoid},), a.h), acc)
end
# `f` / `op` of src/accumulator.jl:78-81: a closure cannot cross the C ABI, so the functions the package itself passes are
# mapped to codes (ibx_accumulate_ex); any other f must be elementwise and Δ = false (then f(v[stencil]) == f.(v)[stencil]).
const _ACC_F = IdDict{Any, Cint}(identity => 0, abs => 1, abs2 => 2, sign => 3)
const _ACC_OP = IdDict{Any, Cint}((+) => 0, max => 1, min => 2, (*) => 3)
function (acc::Accumulator)(v::IBXArray; Δ::Bool = false, f = identity, op = +)
    acc.uploaded || (check(ccall((:ibx_accum_upload, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), context(), acc.h)); acc.uploaded = true)
    out = length(v.dims) == 1 ? IBXArray{1}((acc.n_output,)) : IBXArray{2}((acc.n_output, ncols(v)))
    haskey(_ACC_OP, op) || error("Accumulator: op must be one of +, max, min, * on the B200 backend")
    if !haskey(_ACC_F, f)
        Δ && error("Accumulator: an arbitrary f with Δ = true cannot be evaluated on the device (use abs, abs2 or sign)")
        v = f(v); f = identity          # elementwise f applied on the device array first
    end
    check(ccall((:ibx_accumulate_ex, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Cint, Cint, Int64),
                context(), acc.h, v.h, Δ, _ACC_F[f], _ACC_OP[op], out.h)); out
end

"""Run-time options of the fused Euler residual (`ibx_set_option`): `:arithmetic` (0 reference-exact, 1 fast),
`:path` (0 marching kernels, 1 tile kernels, 2 gather kernels), `:sensor` (1 JST blend, 0 `D = nothing`)."""
set_option(name::Union{Symbol, String}, value::Integer) =
    check(ccall((:ibx_set_option, libibx), Cint, (Ptr{Cvoid}, Cstring, Cint), context(), String(name), value))
function get_option(name::Union{Symbol, String})
    v = Ref{Cint}(0)
    check(ccall((:ibx_get_option, libibx), Cint, (Ptr{Cvoid}, Cstring, Ref{Cint}), context(), String(name), v)); Int(v[])
end
"`Interpolator(X, Xc; linear, k)` (src/nninterp.jl:85-138); X, Xc are (nd, npoints) like the reference and are passed row-major."
function Interpolator(X::AbstractMatrix, Xc::AbstractMatrix; linear::Bool = true, k::Int = 0)
    Xr = Matrix{Float32}(X); Xcr = Matrix{Float32}(Xc)          # column-major (nd, n) == row-major (n, nd)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ibx_interpolator_build, libibx), Cint, (Cint, Int64, Ptr{Float32}, Int64, Ptr{Float32}, Ptr{Float32}, Cint, Cint, Ref{Ptr{Cvoid}}),
                size(Xr, 1), size(Xr, 2), Xr, size(Xcr, 2), Xcr, C_NULL, linear, k, out))
    Accumulator(out[])
end
function cell_centers(dom::Domain)
    X = Matrix{Float32}(undef, dom.nd, dom.ncells)              # (nd, ncells) column-major == the ABI's ncells x nd row-major
    check(ccall((:ibx_domain_cells, libibx), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}), dom.h, X, C_NULL)); X
end
"`multigrid(dom)` (src/ImmersedBoundary.jl:1355-1407): coarse Domains from `ibx_mesh_from_blocks`, IDW transfer accumulators from
`ibx_interpolator_build`; returns `(coarse_doms, prolongators, coarseners)` like the reference code (:1406)."
function multigrid(dom::Domain; max_levels::Int = 0, factor::Int = 2, domain_kwargs...)
    msh = dom.mesh
    max_levels = max_levels == 0 ? floor(Int, log2(msh.block_size)) : max_levels
    coarse_doms = Domain[]; coarseners = Accumulator[]; prolongators = Accumulator[]
    Xold = cell_centers(dom); bsize = Int(msh.block_size)
    for _ = 1:max_levels
        bsize ÷= factor
        mh = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ibx_mesh_from_blocks, libibx), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}), msh.h, bsize, mh))
        cmsh = Mesh(mh[], Int32(bsize), msh.nd, length(dom) ÷ (Int(msh.block_size) ÷ bsize)^msh.nd)
        cdom = Domain(cmsh; domain_kwargs...)
        X = cell_centers(cdom)
        push!(coarseners, Interpolator(Xold, X; linear = false)); push!(prolongators, Interpolator(X, Xold; linear = false))
        push!(coarse_doms, cdom); Xold = X
    end
    (coarse_doms, prolongators, coarseners)
end

"`FAS!(f, Q, coarseners, prolongators; ...)` (src/solver.jl:39-91) on device arrays; keeps the `length(coarseners) > 1` condition (:60)."
function FAS!(f, Q::IBXArray, coarseners = (), prolongators = (); prescribed_f = nothing, multigrid_level::Int = 0,
              n_iter::Int = 50, rtol::Real = 1f-1, atol::Real = 1f-7)
    fQ, ω = f(multigrid_level, Q)
    source = isnothing(prescribed_f) ? nothing : prescribed_f .- fQ
    r = isnothing(source) ? fQ : fQ .+ source
    nr0 = norm(r); nr = nr0
    if length(coarseners) > 1
        Qc = coarseners[1](Q); Qcold = copy(Qc)
        FAS!(f, Qc, coarseners[2:end], prolongators[2:end]; prescribed_f = coarseners[1](r), multigrid_level = multigrid_level + 1,
             n_iter = n_iter, rtol = rtol, atol = atol)
        Q .+= prolongators[1](Qc .- Qcold)
    end
    for _ = 1:n_iter
        r, ω = f(multigrid_level, Q)
        isnothing(source) || (r .+= source)
        om = ω isa IBXArray ? ω : (o = IBXArray{1}((size(Q, 1),)); check(ccall((:ibx_array_fill, libibx), Cint, (Ptr{Cvoid}, Int64, Cfloat), context(), o.h, ω)); o)
        check(ccall((:ibx_clamped_update, libibx), Cint, (Ptr{Cvoid}, Int64, Int64, Int64), context(), Q.h, om.h, r.h))
        nr = norm(r)
        nr < nr0 * rtol + atol && break
    end
    nr / (nr0 + eps(Float32))
end

"One evaluation with HOST arrays (pinned for asynchronous copies): H2D(Q) -> ghost updates in `bcs` order -> residual -> D2H(R, cfl)."
struct BCSpec; boundary::Cint; normal_flow::Cint; n_pinf::Cint; Pinf::NTuple{5, Float32}; end
BCSpec(dom::Domain, name::String, bc::FlowBC) = BCSpec(dom.boundary_index[name], bc.normal_flow, length(bc.P),
                                                       ntuple(i -> i <= length(bc.P) ? bc.P[i] : 0.0f0, 5))
function euler_step_host!(dom::Domain, fluid::Fluid, bcs::Vector{Pair{String, FlowBC}}, Q::Matrix{Float32}, R::Matrix{Float32},
                          cfl::Vector{Float32}; flux_kind::Int = 0, slot::Union{Nothing, Int} = nothing)
    specs = [BCSpec(dom, n, bc) for (n, bc) in bcs]
    GC.@preserve specs Q R cfl begin
        if isnothing(slot)
            check(ccall((:ibx_euler_step_host, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Fluid, Cint, Cint, Ptr{BCSpec}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
                        context(), dom.h, fluid, flux_kind, length(specs), specs, Q, R, cfl))
        else    # two evaluations in flight: the caller keeps Q, R, cfl alive until euler_step_host_end(slot)
            check(ccall((:ibx_euler_step_host_begin, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Fluid, Cint, Cint, Ptr{BCSpec}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Cint),
                        context(), dom.h, fluid, flux_kind, length(specs), specs, Q, R, cfl, slot))
        end
    end
end
euler_step_host_end(slot::Int) = check(ccall((:ibx_euler_step_host_end, libibx), Cint, (Ptr{Cvoid}, Cint), context(), slot))

"Halo exchange of a rank-local (sharded) domain; `halo_begin!` alone may precede `residual_euler!` on the same array (it completes it)."
halo_begin!(dom::Domain, a::IBXArray) = check(ccall((:ibx_halo_begin, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), context(), dom.h, a.h))
halo_end!(dom::Domain, a::IBXArray) = check(ccall((:ibx_halo_end, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), context(), dom.h, a.h))

"""One step of a solver loop: ghost updates of `bcs` in order, then the residual, with the ghost update (and, on a rank-local
shard, both halo exchanges) hidden behind the residual of the blocks that read neither a ghost nor a halo cell
(`ibx_step_euler` / `ibx_step_euler_sharded`).  Same bits as `ghost_update_euler!` + `residual_euler!`."""
function step_euler!(dom::Domain, fluid::Fluid, bcs::Vector{Pair{String, FlowBC}}, Q::IBXArray, R::IBXArray, cfl::IBXArray;
                     flux_kind::Int = 0, sharded::Bool = false, exchange_between_families::Bool = false)
    specs = [BCSpec(dom, n, bc) for (n, bc) in bcs]
    GC.@preserve specs begin
        if sharded
            check(ccall((:ibx_step_euler_sharded, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Fluid, Cint, Cint, Ptr{BCSpec}, Cint, Int64, Int64, Int64),
                        context(), dom.h, fluid, flux_kind, length(specs), specs, exchange_between_families, Q.h, R.h, cfl.h))
        else
            check(ccall((:ibx_step_euler, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Fluid, Cint, Cint, Ptr{BCSpec}, Int64, Int64, Int64),
                        context(), dom.h, fluid, flux_kind, length(specs), specs, Q.h, R.h, cfl.h))
        end
    end
end

"`Q = Q0 + ((alpha / cfl) * R) * mask` in one kernel (`ibx_local_step_update`); `Q` may be `Q0`."
local_step_update!(Q::IBXArray, Q0::IBXArray, R::IBXArray, cfl::IBXArray, alpha::Real; mask::Union{Nothing, IBXArray} = nothing) =
    check(ccall((:ibx_local_step_update, libibx), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Cfloat, Int64),
                context(), Q0.h, R.h, cfl.h, isnothing(mask) ? 0 : mask.h, Float32(alpha), Q.h))

"""Pseudo-time march with local time steps on the device: per step, ghost update in place, `Q0 = Q`, then for every
multistage coefficient `a`: `step_euler!` and `Q = Q0 + a CFL R / cfl * live`.  `live` (0/1 per cell) freezes the ghost cells
between residual evaluations (they only take boundary values)."""
function march_euler!(dom::Domain, fluid::Fluid, bcs::Vector{Pair{String, FlowBC}}, Q::IBXArray, n_steps::Int;
                      CFL::Real = 0.8f0, stages = (0.1481f0, 0.4f0, 1.0f0), live::Union{Nothing, IBXArray} = nothing, flux_kind::Int = 0,
                      graph::Bool = true)
    # the whole loop runs inside the library (`ibx_march_euler`); with `graph` one step is replayed from a CUDA graph
    specs = [BCSpec(dom, n, bc) for (n, bc) in bcs]
    al = Float32[stages...]
    used = Ref{Cint}(0)
    GC.@preserve specs al begin
        check(ccall((:ibx_march_euler, libibx), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Fluid, Cint, Cint, Ptr{BCSpec}, Int64, Int64, Int64, Cfloat, Cint, Ptr{Cfloat}, Cint, Ref{Cint}),
                    context(), dom.h, fluid, flux_kind, length(specs), specs, Q.h, isnothing(live) ? 0 : live.h, n_steps, Float32(CFL),
                    length(al), al, graph, used))
    end
    Q
end

"Configuration C5: canonical RANS residual (`ibx_residual_rans`) and the ghost update of the transported variable."
residual_rans!(dom::Domain, fluid::Fluid, Q::IBXArray, qR::IBXArray, R::IBXArray, RR::IBXArray, cfl::IBXArray;
               transport::Transport = Transport(), σR::Real = 0.72f0, C₁::Real = 0.0829f0, κ::Real = 0.41f0) =
    check(ccall((:ibx_residual_rans, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Fluid, Transport, Cfloat, Cfloat, Cfloat, Int64, Int64, Int64, Int64, Int64),
                context(), dom.h, fluid, transport, σR, C₁, κ, Q.h, qR.h, R.h, RR.h, cfl.h))
ghost_update_rans!(dom::Domain, name::String, Q::IBXArray, qR::IBXArray, R_bc::Real) =
    check(ccall((:ibx_ghost_update_rans, libibx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Int64, Cfloat), context(), dom.h,
                dom.boundary_index[name], Q.h, qR.h, R_bc))

end # module
