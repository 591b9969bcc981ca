#=
Pinning kit for a Julia owner (VERDICT r1, item 9): run the REAL ImmersedBoundary.jl on the three configurations the
parity tests use and dump its tables and operator outputs, so that tests/test_reference_fixtures.py can pin the oracle
(and through it the sm_100a kernels) to the reference itself instead of to a restatement of it.

    julia --project=/path/to/ImmersedBoundary.jl tools/gen_reference_fixtures.jl [outdir = tests/golden/reference]

This image has no `julia` (SURVEY.md F2), so this script has never been executed here; it only uses the package's public
API and struct fields as of v1.1.0 (src/ImmersedBoundary.jl:383-414, 483-490; src/mesher.jl:926-933, 1064; src/accumulator.jl:137).
Output: one raw little-endian `.bin` per array (Julia memory order = column-major) + manifest.json (name -> eltype, size).
Indices are written as the package holds them (1-based); the Python side converts.
=#
using ImmersedBoundary
const IB = ImmersedBoundary

outdir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "reference")
mkpath(outdir)
manifest = String[]

function dump!(name::String, A::AbstractArray)
    A = collect(A)
    T = eltype(A)
    open(joinpath(outdir, name * ".bin"), "w") do io
        write(io, A)
    end
    push!(manifest, "\"$name\": {\"eltype\": \"$(T)\", \"size\": [$(join(size(A), ", "))]}")
end
dump!(name::String, x::Number) = dump!(name, [x])

# lists of unequal length -> CSR (ptr 0-based offsets, values as stored)
function dump_lists!(name::String, lists)
    ptr = Int64[0]
    vals = eltype(first(lists))[]
    for l in lists
        append!(vals, l)
        push!(ptr, length(vals))
    end
    dump!(name * "_ptr", ptr)
    dump!(name * "_val", vals)
end

function dump_domain!(tag::String, msh, dom)
    dump!(tag * "_block_origins", msh.block_origins)
    dump!(tag * "_block_widths", msh.block_widths)
    centers, widths = IB.get_cells(msh)
    dump!(tag * "_centers", centers)
    dump!(tag * "_widths", widths)
    nd = ndims(dom)
    pids = sort(collect(keys(dom.partitions)))
    dump!(tag * "_partition_ids", Int64.(pids))
    for pid in pids
        part = dom.partitions[pid]
        p = "$(tag)_p$(pid)"
        dump!(p * "_domain", Int64.(part.domain))
        dump!(p * "_image", Int64.(part.image))
        dump!(p * "_image_in_domain", Int64.(part.image_in_domain))
        for dim in 1:nd
            o, n = part.face_owners_neighbors[dim]
            dump!(p * "_own$(dim)", Int64.(o))
            dump!(p * "_nei$(dim)", Int64.(n))
            for side in (false, true)
                dec = IB.decompose(part.face_accumulators[(dim, side)])
                inds = dec isa Tuple ? dec[1] : dec
                dump_lists!(p * "_faces$(dim)$(side ? 'r' : 'l')", [Int64.(i) for i in inds])
            end
        end
    end
    for (bname, chunks) in dom.boundaries
        cids = sort(collect(keys(chunks)))
        dump!("$(tag)_b_$(bname)_chunks", Int64.(cids))
        for cid in cids
            b = chunks[cid]
            p = "$(tag)_b_$(bname)_$(cid)"
            dump!(p * "_ghost", Int64.(b.ghost_indices))
            dump!(p * "_proj", b.projections)
            dump!(p * "_normals", b.normals)
            dump!(p * "_image_dist", b.image_distances)
            dump!(p * "_ghost_dist", b.ghost_distances)
            dump!(p * "_image_domain", Int64.(b.image_domain))
            inds, ws = IB.decompose(b.image_interpolator)
            dump_lists!(p * "_donors", [Int64.(i) for i in inds])
            dump_lists!(p * "_weights", [Float64.(w) for w in ws])
        end
    end
end

# deterministic smooth field of the cell centres (Float32 arithmetic on Float32 centres)
field(X) = sin.(3.0f0 .* X[:, 1]) .+ cos.(2.0f0 .* X[:, 2]) .+ (size(X, 2) > 2 ? 0.5f0 .* X[:, 3] : 0.0f0)

function dump_operators!(tag::String, dom)
    N, nd = length(dom), ndims(dom)
    X = zeros(Float32, N, nd)
    dom(X) do part, X
        X .= part.centers
    end
    u = Float32.(field(X))
    dump!(tag * "_u", u)
    D = zeros(Float32, N)
    dom(u, D) do part, u, D
        D .= JST_sensor(part, u)
    end
    dump!(tag * "_jst", D)
    for dim in 1:nd
        g = zeros(Float32, N); uL = zeros(Float32, N); uR = zeros(Float32, N); gg = zeros(Float32, N); ugg = zeros(Float32, N)
        dom(u, g, gg, ugg) do part, u, g, gg, ugg
            du = cell_gradient(part, u, dim)
            g .= du
            l, r = MUSCL(part, u, du, dim; D = JST_sensor(part, u), high_order = true)
            gg .= green_gauss(part, (l .+ r) ./ 2, dim)
            ugg .= unsigned_green_gauss(part, at_faces(part, u, dim), dim)
        end
        dump!(tag * "_grad$(dim)", g)
        dump!(tag * "_gg_muscl$(dim)", gg)
        dump!(tag * "_ugg_faces$(dim)", ugg)
    end
    # Euler pieces (src/cfd.jl): state conversions and the HLL flux of the canonical residual (SURVEY.md A.10)
    fluid = IB.CFD.Fluid()
    a∞ = sqrt(1.4f0 * 283.0f0 * 288.15f0)
    P = zeros(Float32, N, nd + 2)
    P[:, 1] .= 101325.0f0 .* (1 .+ 0.02f0 .* u ./ 3)
    P[:, 2] .= 288.15f0 .* (1 .+ 0.01f0 .* cos.(X[:, 1]))
    P[:, 3] .= 0.5f0 * a∞ .* (1 .+ 0.05f0 .* sin.(X[:, 2]))
    Q = IB.CFD.primitive2state(fluid, P)
    dump!(tag * "_P", P)
    dump!(tag * "_Q", Q)
    R = zeros(Float32, N, nd + 2)
    cfl = zeros(Float32, N)
    dom(Q, R, cfl) do part, Q, R, cfl
        Pp = IB.CFD.state2primitive(fluid, Q)
        Dp = JST_sensor(part, Pp[:, 1])
        a = IB.CFD.speed_of_sound(fluid, Pp[:, 2])
        R .= 0
        cfl .= 0
        for dim in 1:ndims(part)
            ∇P = cell_gradient(part, Pp, dim)
            PL, PR = MUSCL(part, Pp, ∇P, dim; D = Dp, high_order = false)
            F = IB.CFD.inviscid_fluxes(fluid, PL, PR, dim)
            R .-= green_gauss(part, F, dim)
            cfl .+= unsigned_green_gauss(part, abs.(at_faces(part, Pp[:, 2 + dim], dim)) .+ at_faces(part, a, dim), dim)
        end
    end
    dump!(tag * "_R", R)
    dump!(tag * "_cfl", cfl)
end

# ---------------------------------------------------------------- C1: test/advection.jl mesh
let
    lower = Stereolitography([0.0 1.0; 0.0 0.0])
    upper = Stereolitography([0.0 0.0; 0.0 1.0])
    msh = Mesh([0.0, 0.0], [1.0, 1.0], ("lower", lower, 1f-2), ("upper", upper, 1f-2);
               refinement_regions = [Line([0.0, 0.0], [1.0, 1.0]) => 2f-2, Line([0.0, 0.0], [0.5, 0.5]) => 1f-2])
    dom = Domain(msh; hypercube_families = ["outlet" => [(1, true), (2, true)]])
    dump_domain!("advection", msh, dom)
    dump_operators!("advection", dom)
end

# ---------------------------------------------------------------- C3: test/rae2822.jl mesh (max_partition_size 10 000 -> several partitions)
let
    stl = Stereolitography(joinpath(@__DIR__, "..", "tests", "golden", "rae2822.dat")) |> merge_points
    features = feature_regions(stl; radius = 0.05) |> DistanceField
    msh = Mesh([-25.0f0, -25.0f0], [50.0f0, 50.0f0], ("wall", stl, 1f-2); refinement_regions = [features => 5f-3])
    dom = Domain(msh; max_partition_size = 10_000, hypercube_families = ["farfield" => [(1, false), (1, true), (2, false), (2, true)]])
    dump_domain!("rae2822", msh, dom)
    dump_operators!("rae2822", dom)
    X = zeros(Float32, length(dom), ndims(dom))
    dom(X) do part, X
        X .= part.centers
    end
    dump!("rae2822_CG", volume_integral(dom, X) ./ 2500.0f0)
end

# ---------------------------------------------------------------- 3-D: icosahedron-based sphere STL (vertices written by the Python side)
let
    pts_file = joinpath(outdir, "icosphere1_points.f32")     # written by tests/golden/make_fixtures.py: 3 x npts Float32
    tri_file = joinpath(outdir, "icosphere1_triangles.i64")  # 3 x ntri Int64, 1-based
    if isfile(pts_file) && isfile(tri_file)
        pts = reshape(reinterpret(Float32, read(pts_file)), 3, :)
        tri = reshape(reinterpret(Int64, read(tri_file)), 3, :)
        stl = Stereolitography(collect(pts), collect(tri))
        msh = Mesh([-2.0, -2.0, -2.0], [4.0, 4.0, 4.0], ("wall", stl, 0.12f0); refinement_regions = [Ball([0.0, 0.0, 0.0], 0.9) => 0.12f0])
        dom = Domain(msh; hypercube_families = ["farfield" => [(d, s) for d in 1:3 for s in (false, true)]])
        dump_domain!("sphere3d_stl", msh, dom)
        dump_operators!("sphere3d_stl", dom)
    else
        @warn "icosphere input files not found; run `python tests/golden/make_fixtures.py --icosphere` first" pts_file
    end
end

# ---------------------------------------------------------------- the docstring KAT (src/accumulator.jl:25-34)
let
    acc = IB.Accumulator([[1, 2], [2, 3, 4]], [[-1.0, 2.0], [3.0, 4.0, 5.0]])
    dump!("accumulator_kat", acc([1, 2, 3, 4]))
end

open(joinpath(outdir, "manifest.json"), "w") do io
    write(io, "{\n" * join(manifest, ",\n") * "\n}\n")
end
@info "wrote $(length(manifest)) arrays to $outdir"
