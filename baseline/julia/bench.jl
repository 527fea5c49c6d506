# bench.jl -- the reference's own CPU path for the benchmark workloads, for a box that has Julia
# and CameraCalibrations.jl installed (this repo's image has neither: bench.py --impl reference
# times the C restatement under oracle/ instead and says so).
#
#   JULIA_NUM_THREADS=$(nproc) julia --project=<env with CameraCalibrations> baseline/julia/bench.jl
#
# Prints one JSON line per workload: rectified Mpix/s (warp of src/plot_calibration.jl:40 on the
# bench view) and points/s of c.(pts, 1) (src/meta.jl:82,88), with the thread and core counts.
using CameraCalibrations, ImageTransformations, CoordinateTransformations, StaticArrays, Rotations
using CameraCalibrations: RowCol, XYZ

const INTR = (f = 1400.0, c = (540.0, 960.0), k = -0.12)
const VIEW = (rvec = (0.05, -0.04, 0.02), tvec = (-9.3, -6.4, 30.0))
const SZ = (1080, 1920)

function calibration()
    # the constructor of src/meta.jl:27-33 through the helper of src/buildcalibrations.jl:1-6
    intrinsic, extrinsics, scale = CameraCalibrations.obj2img([collect(VIEW.rvec)], [collect(VIEW.tvec)],
                                                              INTR.f, INTR.f, INTR.c..., 1.0)
    Calibration(intrinsic, extrinsics, scale, INTR.k, ["extrinsic.png"])
end

function main()
    nthreads, ncores = Threads.nthreads(), Sys.CPU_THREADS
    c = calibration()
    # --- point maps: threaded chunks of the broadcast the reference uses (src/buildcalibrations.jl:29,46)
    n = 2_000_000
    pts = [RowCol(SZ[1] * rand(), SZ[2] * rand()) for _ in 1:n]
    out = Vector{XYZ}(undef, n)
    f!() = Threads.@threads for chunk in collect(Iterators.partition(1:n, cld(n, 4nthreads)))
        @inbounds for i in chunk
            out[i] = c(pts[i], 1)
        end
    end
    f!(); t = @elapsed f!()
    println("{\"metric\": \"img2world_points_per_s\", \"value\": $(n / t), \"threads\": $nthreads, \"cores\": $ncores}")
    # --- rectification: warp, one frame per task (warp! itself is single-threaded)
    ratio = 1.0
    tform = c.real2image[1] ∘ (v -> push(v, 0)) ∘ inv(LinearMap(SDiagonal(ratio, ratio)))
    frames = [rand(Float32, SZ...) for _ in 1:2nthreads]
    g!() = Threads.@threads for i in eachindex(frames)
        warp(frames[i], tform, axes(frames[i]))
    end
    g!(); t = @elapsed g!()
    println("{\"metric\": \"rectified_mpix_per_s\", \"value\": $(length(frames) * prod(SZ) / t / 1e6), \"threads\": $nthreads, \"cores\": $ncores}")
end

main()
