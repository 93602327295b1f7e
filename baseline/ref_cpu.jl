# baseline/ref_cpu.jl — times the UNMODIFIED reference (Hikari.jl VolPath on KernelAbstractions.CPU()) on the C1 workload
# (SURVEY 8d), for machines that have Julia + Hikari.jl + Raycore.jl installed.  It cannot run in the build image of this
# repository (no Julia, no network; Raycore.jl is an un-vendored git dependency): there `bench.py --impl reference` times the
# CPU oracle port instead and labels it "port".  Usage, from a Hikari.jl checkout:
#
#     julia -t auto --project=. /path/to/baseline/ref_cpu.jl [spp] [width] [height]
#
# Prints one JSON line in the format of bench.py's reference arm, with "kind": "reference".
using Hikari, GeometryBasics
using GeometryBasics: normal_mesh, Tesselation

spp    = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 4
width  = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 512
height = length(ARGS) >= 3 ? parse(Int, ARGS[3]) : 512

# C1(b): floor + three matte UV-spheres under a directional light (examples/sphere_normals_test.jl:41-98), current API
# (test/volpath_integration.jl:36-82: Scene(), push!, sync!, PerspectiveCamera(eye, lookat, film; fov), VolPath(samples, max_depth))
scene = Hikari.Scene()
push!(scene, normal_mesh(Rect3f(Vec3f(-5, -1, -5), Vec3f(10, 0.1f0, 10))), Hikari.MatteMaterial(Kd=Hikari.RGBSpectrum(0.7f0, 0.7f0, 0.7f0)))
for (x, kd) in ((-1.5f0, (0.8f0, 0.2f0, 0.2f0)), (0f0, (0.2f0, 0.8f0, 0.2f0)), (1.5f0, (0.2f0, 0.2f0, 0.8f0)))
    push!(scene, normal_mesh(Tesselation(Sphere(Point3f(x, 0.5f0, 0f0), 0.8f0), 64)), Hikari.MatteMaterial(Kd=Hikari.RGBSpectrum(kd...)))
end
push!(scene, Hikari.DirectionalLight(Hikari.RGBSpectrum(3f0, 3f0, 3f0), normalize(Vec3f(-1, -1.5, -0.5))))
Hikari.sync!(scene)

film = Hikari.Film(Point2f(width, height))
camera = Hikari.PerspectiveCamera(Point3f(0f0, 1.5f0, 4f0), Point3f(0f0, 0.5f0, 0f0), film; fov=40f0)

warm = Hikari.VolPath(samples=1, max_depth=5)
warm(scene, film, camera)                       # compile + first touch, untimed
Hikari.clear!(film)

integrator = Hikari.VolPath(samples=spp, max_depth=5)
t = @elapsed integrator(scene, film, camera)
msamples = width * height * spp / t / 1e6
println("""{"impl": "reference", "metric": "VolPath throughput", "value": $(msamples), "unit": "Msamples/s", "n_gpus": 0, "steps": $(spp), "ms_per_step": $(1e3 * t / spp), "higher_is_better": true, "dtype": "f32", "data": "synthetic", "config": {"workload": "C1 sphere-normals scene, $(width)x$(height), VolPath max_depth=5"}, "cpu_baseline": {"value": $(msamples), "unit": "Msamples/s", "cores": $(Threads.nthreads()), "kind": "reference", "sample": "$(spp) spp"}}""")
