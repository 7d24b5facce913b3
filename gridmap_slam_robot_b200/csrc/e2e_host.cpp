// e2e_host — the SLAM step through the C-ABI from a compiled host, the way the Java shim's FFM downcalls reach it
// (GridMapApp.java:178-192: update -> Neff -> resample -> getStrongestParticle -> getWeightedPose), with HOST beam
// arrays: every H2D / D2H copy is inside the timed region.  bench.py runs it for the `e2e_native` figure; its own
// `e2e` goes through the Python binding and carries the interpreter's per-call cost.
//
//   e2e_host <libgms.so> <scans.bin> <particles> <grid_m> <shared 0|1> <steps> <warmup>
//   scans.bin: int32 nscan, int32 B, then per scan: f64 d_center, f64 d_theta, f64 xy[2B], f64 dist[B], u8 hit[B]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <dlfcn.h>

#include "../../include/gms.h"

#define SYM(name) auto p_##name = reinterpret_cast<decltype(&name)>(dlsym(lib, #name)); if (!p_##name) { std::fprintf(stderr, "missing %s\n", #name); return 2; }

int main(int argc, char** argv) {
    if (argc < 8) { std::fprintf(stderr, "usage: e2e_host lib scans particles grid_m shared steps warmup\n"); return 2; }
    void* lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) { std::fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    SYM(gms_config_default) SYM(gms_create) SYM(gms_destroy) SYM(gms_update) SYM(gms_resample) SYM(gms_get_strongest)
    SYM(gms_get_weighted_pose) SYM(gms_last_error) SYM(gms_sync)
    FILE* f = std::fopen(argv[2], "rb");
    if (!f) { std::perror("scans"); return 2; }
    int32_t nscan = 0, B = 0;
    if (std::fread(&nscan, 4, 1, f) != 1 || std::fread(&B, 4, 1, f) != 1) return 2;
    struct Scan { double dc, dt; std::vector<double> xy, dist; std::vector<uint8_t> hit; int hits; };
    std::vector<Scan> scans(nscan);
    for (auto& s : scans) {
        s.xy.resize(2 * (size_t)B); s.dist.resize(B); s.hit.resize(B);
        if (std::fread(&s.dc, 8, 1, f) != 1 || std::fread(&s.dt, 8, 1, f) != 1 ||
            std::fread(s.xy.data(), 8, 2 * (size_t)B, f) != 2 * (size_t)B || std::fread(s.dist.data(), 8, B, f) != (size_t)B ||
            std::fread(s.hit.data(), 1, B, f) != (size_t)B) return 2;
        s.hits = 0;
        for (uint8_t v : s.hit) s.hits += v != 0;
    }
    std::fclose(f);
    const int P = std::atoi(argv[3]);
    const float grid = (float)std::atof(argv[4]);
    const int shared = std::atoi(argv[5]), steps = std::atoi(argv[6]), warm = std::atoi(argv[7]);
    gms_config c;
    p_gms_config_default(&c);
    c.num_particles = P; c.map_width_m = grid; c.map_height_m = grid; c.origin_x = -grid / 2; c.origin_y = -grid / 2;
    c.map_mode = shared ? GMS_MAP_SHARED : GMS_MAP_PER_PARTICLE; c.seed = 20260101;
    gms_handle* h = nullptr;
    if (p_gms_create(&c, &h) != GMS_OK) { std::fprintf(stderr, "gms_create: %s\n", p_gms_last_error(nullptr)); return 3; }
    double scored = 0, neff = 0, w = 0;
    float pose[3], wp[3] = {0, 0, 0};
    int32_t idx = 0;
    auto step = [&](int i) -> int {
        const Scan& s = scans[i % nscan];
        int rc = p_gms_update(h, s.xy.data(), s.dist.data(), s.hit.data(), B, s.dc, s.dt, nullptr, &neff);
        if (!rc) rc = p_gms_resample(h, -1.0);
        if (!rc) rc = p_gms_get_strongest(h, &idx, pose, &w);
        if (!rc) rc = p_gms_get_weighted_pose(h, wp);
        return rc;
    };
    for (int i = 0; i < warm; i++)
        if (step(i)) { std::fprintf(stderr, "step: %s\n", p_gms_last_error(h)); return 4; }
    p_gms_sync(h);
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < steps; i++) {
        if (step(warm + i)) { std::fprintf(stderr, "step: %s\n", p_gms_last_error(h)); return 4; }
        scored += (double)P * scans[(warm + i) % nscan].hits;
    }
    p_gms_sync(h);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("{\"ms_per_step\": %.6f, \"value\": %.6e, \"steps\": %d, \"neff_last\": %.6f, \"weighted_pose\": [%.6f, %.6f, %.6f]}\n",
                1e3 * sec / steps, scored / sec, steps, neff, wp[0], wp[1], wp[2]);
    p_gms_destroy(h);
    return 0;
}
