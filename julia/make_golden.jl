# make_golden.jl -- dump outputs of the REAL CameraCalibrations.jl for the parity tests.
#
#   julia --project=<env with CameraCalibrations v0.7.3> julia/make_golden.jl
#
# NOT RUN IN THIS REPOSITORY'S IMAGE (no Julia, SURVEY.md F1).  It reads the parameters the Python
# fixtures use (tests/golden/example_fit.json, written by tests/golden/make_golden.py), evaluates the
# reference's own functions on them and writes tests/golden/julia_*.json.  tests/test_julia_golden.py
# consumes those files when they exist (oracle on CPU, CUDA path on the GPU box) and reports an
# expected failure ("parity unpinned") while they do not.  One command turns the pin green.
#
# What is dumped, and the reference lines it exercises:
#   julia_points.json      c(::RowCol, i) src/meta.jl:82, c(::XYZ, i) :88, rectification :99-103
#   julia_errors.json      calculate_errors src/buildcalibrations.jl:37-67
#   julia_warp.json        get_ratio/get_axes/image_transformations src/plot_calibration.jl:1-22 and
#                          warp(img, tform, axs) :40 on a Float32 and an RGB{N0f8} image, plus probes of
#                          the interpolation rule at x == 1, x == n, just outside, and exact .5 ties
#   julia_calibration.json CameraCalibrations.save of the object (src/io.jl:28-32): the JSON3 schema
using CameraCalibrations, JSON3, StaticArrays, Random, LinearAlgebra
using ImageTransformations: warp
using CoordinateTransformations, Rotations
using ColorTypes: RGB, Gray
using FixedPointNumbers: N0f8
using CameraCalibrations: RowCol, XYZ, Calibration, obj2img, calculate_errors, get_ratio, get_axes,
                          image_transformations, rectification

const ROOT = normpath(joinpath(@__DIR__, ".."))
const GOLD = joinpath(ROOT, "tests", "golden")
fit = JSON3.read(read(joinpath(GOLD, "example_fit.json"), String))

intr = fit.intr
Rs = [collect(Float64, v.rvec) for v in fit.views]
ts = [collect(Float64, v.tvec) for v in fit.views]
cs = Float64(fit.checker_size)
intrinsic, extrinsics, scale = obj2img(Rs, ts, intr.frow, intr.fcol, intr.crow, intr.ccol, cs)
files = String.(fit.files)
c = Calibration(intrinsic, extrinsics, scale, Float64(intr.k), files)
sz = (Int(fit.sz[1]), Int(fit.sz[2]))
n_corners = (Int(fit.n_corners[1]), Int(fit.n_corners[2]))
nviews = length(files)
wr(name, x) = open(io -> JSON3.write(io, x; allow_inf = true), joinpath(GOLD, name), "w")

# ---- point maps ------------------------------------------------------------------------------
rng = MersenneTwister(20240601)
pix = vcat([RowCol(1.2, 3.4), RowCol(1.0, 2.0), RowCol(187.0, 250.0), RowCol(375.0, 500.0)],
           [RowCol(1 + (sz[1] - 1) * rand(rng), 1 + (sz[2] - 1) * rand(rng)) for _ in 1:200])
points = map(1:nviews) do i
    xyz = [c(p, i) for p in pix]                                   # src/meta.jl:82
    back = [c(q, i) for q in xyz]                                  # src/meta.jl:88
    rect = [rectification(c, i)(p) for p in pix]                   # src/meta.jl:99
    (; view = i, xyz = [collect(q) for q in xyz], back = [collect(q) for q in back],
       rect = [collect(q) for q in rect])
end
wr("julia_points.json", (; pix = [collect(p) for p in pix], views = points))

# ---- calculate_errors ---------------------------------------------------------------------------
# imgpoints as the reference holds them: a n1 x n2 Matrix{RowCol} per view (src/detect_fit.jl:19)
imgpointss = [reshape([RowCol(Float64(p[1]), Float64(p[2])) for p in fit.corners[i]], n_corners) for i in 1:nviews]
objpoints = reshape([XYZ(Float64(p[1]), Float64(p[2]), Float64(p[3])) for p in fit.obj], n_corners) .* cs
Random.seed!(1)
eps = calculate_errors(c, imgpointss, objpoints, cs, sz, files, n_corners, 100)
wr("julia_errors.json", (; n = eps.n, reprojection = eps.reprojection, projection = eps.projection,
                           distance = eps.distance, inverse = eps.inverse, inverse_samples = 100))

# ---- rectification ------------------------------------------------------------------------------
# (a) the reference's own plot path on a small frame: every view, Float32 gray and RGB{N0f8}
szw = (61, 47)
scale_px = szw[1] / sz[1]                       # shrink the camera with the frame
intr_s, ext_s, scale_s = obj2img(Rs, ts, intr.frow * scale_px, intr.fcol * scale_px, intr.crow * scale_px,
                                 intr.ccol * scale_px, cs)
cw = Calibration(intr_s, ext_s, scale_s, Float64(intr.k), files)
rngw = MersenneTwister(7)
img32 = rand(rngw, Float32, szw...)
img8 = rand(rngw, UInt8, 3, szw...)
imgrgb = collect(reinterpret(reshape, RGB{N0f8}, img8))
imgpts_s = [[p .* scale_px for p in ip] for ip in imgpointss]
warps = map(1:nviews) do i
    t = image_transformations(cw, i, imgpts_s, cs, n_corners, szw)           # src/plot_calibration.jl:15-22
    ratio = get_ratio(imgpts_s[i], cs)
    w32 = warp(img32, t.tform, t.axs)                                         # :40 (fill NaN for floats)
    w8 = warp(imgrgb, t.tform, t.axs)                                         # fill = black
    (; view = i, ratio, axs_min = [first(t.axs[1]), first(t.axs[2])],
       out32 = collect(vec(parent(w32))),
       out8 = collect(vec(reinterpret(reshape, UInt8, collect(parent(w8))))))
end
# (b) probes of the interpolation rule: identity camera, so tform(I) = I / ratio exactly
ident_i, ident_e, ident_s = obj2img([[0.0, 0.0, 0.0]], [[0.0, 0.0, 1.0]], 1.0, 1.0, 0.0, 0.0, 1.0)
ci = Calibration(ident_i, ident_e, ident_s, 0.0, ["extrinsic.png"])
probe = rand(MersenneTwister(9), Float32, 12, 9)
probe8 = collect(reinterpret(reshape, RGB{N0f8}, rand(MersenneTwister(10), UInt8, 3, 12, 9)))
probes = map([(1.0, (1, 1)), (1.0, (0, 0)), (2.0, (2, 2)), (2.0, (1, 1)), (4.0, (3, 5)), (0.5, (0, 0))]) do (ratio, amin)
    s = LinearMap(SDiagonal{2}(ratio * I))
    tform = ci.real2image[1] ∘ Base.Fix2(push, 0) ∘ inv(s)
    axs = (amin[1]:amin[1] + 11, amin[2]:amin[2] + 8)
    w = warp(probe, tform, axs)
    w8 = warp(probe8, tform, axs)
    (; ratio, axs_min = collect(amin), out32 = collect(vec(parent(w))),
       out8 = collect(vec(reinterpret(reshape, UInt8, collect(parent(w8))))))
end
wr("julia_warp.json", (; sz = collect(szw), scale_px, img32 = collect(vec(img32)), img8 = collect(vec(img8)),
                         views = warps, probe_sz = [12, 9], probe32 = collect(vec(probe)),
                         probe8 = collect(vec(reinterpret(reshape, UInt8, probe8))), probes))

# ---- the JSON the reference itself writes (schema of RotationVec / AffineMap / SDiagonal) -------
CameraCalibrations.save(joinpath(GOLD, "julia_calibration.json"), c)
println("wrote julia_points.json julia_errors.json julia_warp.json julia_calibration.json to ", GOLD)
