// C++ host-side mirror (include/gms.hpp) exercised the way the reference's callers use its classes
// (GridMapApp.java:123,178-192).  Linked against libgms_ref.so on CPU boxes and libgms.so on the GPU box.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "gms.hpp"

#define CHECK(c)                                                         \
    do {                                                                 \
        if (!(c)) {                                                      \
            std::fprintf(stderr, "%s:%d: CHECK failed: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                    \
        }                                                                \
    } while (0)

#define CHECK_VOID(c)                                                    \
    do {                                                                 \
        if (!(c)) std::fprintf(stderr, "%s:%d: CHECK failed: %s\n", __FILE__, __LINE__, #c); \
    } while (0)

int main() {
    // ray (1,1)->(4,2) on a 10x10 grid, hit: F F O O O 0 0
    {
        gms::SLAM slam([](gms_config& c) {
            c.num_particles = 1;
            c.map_width_m = 0.5f; c.map_height_m = 0.5f; c.origin_x = 0.f; c.origin_y = 0.f;
        });
        CHECK(slam.gridWidth() == 10 && slam.gridHeight() == 10);
        gms::Particle p = slam.getParticles()[0];
        slam.getGridMap().applyMeasurement(p.m, 1, 1, 4, 2, std::sqrt(10.0f), true);
        std::vector<double> log = p.m.logData();
        const double lf = std::log((double)0.30f / (1.0 - (double)0.30f)), lo = std::log((double)0.9f / (1.0 - (double)0.9f));
        CHECK(std::fabs(log[1 + 1 * 10] - lf) < 1e-12 && std::fabs(log[2 + 1 * 10] - lf) < 1e-12);
        CHECK(std::fabs(log[3 + 1 * 10] - lo) < 1e-12 && std::fabs(log[3 + 2 * 10] - lo) < 1e-12 && std::fabs(log[4 + 2 * 10] - lo) < 1e-12);
        CHECK(log[5 + 2 * 10] == 0.0 && log[6 + 2 * 10] == 0.0 && log[0] == 0.0);
    }
    // the SLAM loop of GridMapApp.onHandleData on a small synthetic scan (a 2 m square room)
    {
        gms::SLAM slam([](gms_config& c) { c.num_particles = 40; });
        CHECK(slam.getParticles().size() == 40);
        CHECK(std::fabs(slam.calculateNeff() - 40.0) < 1e-9);
        for (int step = 0; step < 4; step++) {
            gms::Observation z;
            for (int b = 0; b < 90; b++) {
                const double a = 2 * M_PI * b / 90;
                const double d = 1.0 / std::fmax(std::fabs(std::cos(a)), std::fabs(std::sin(a)));  // square, half-width 1 m
                z.addMeasurement(gms::Measurement(a, d, true));
            }
            gms::Odometry u(0.0, 0.0);
            const double neff = slam.update(z, u);
            CHECK(neff >= 1.0 && neff <= 40.0 + 1e-9);
            if (neff < slam.getParticles().size() / 2) slam.resample();
            gms::Pose wp = slam.getWeightedPose();
            CHECK(std::isfinite(wp.x) && std::isfinite(wp.y) && std::isfinite(wp.theta));
            gms::Particle best = slam.getStrongestParticle();
            CHECK(best.weight > 0.0 && best.weight <= 1.0);
        }
        slam.resample(0.5);
        std::vector<int32_t> parents = slam.getParents();
        for (size_t i = 1; i < parents.size(); i++) CHECK(parents[i] >= parents[i - 1] && parents[i] < 40);
        gms::Particle p0 = slam.getParticles()[0];
        std::vector<double> lik = p0.m.likelihoodData();
        double mx = 0;
        for (double v : lik) mx = std::fmax(mx, v);
        CHECK(mx > 0.5 && mx <= 1.0 + 1e-12);  // walls have been integrated and blurred
        slam.reset();
        CHECK(std::fabs(slam.calculateNeff() - 40.0) < 1e-9);
    }
    // A4: GridMap.findBestPoseOptim (SLAM.java:97) as a hook between the motion sample and the weight
    {
        gms::SLAM slam([](gms_config& c) { c.num_particles = 8; });
        int calls = 0;
        slam.setPoseOptimizer([&](int first, std::vector<gms::Pose>& poses) {
            calls++;
            CHECK_VOID(first == 0 && poses.size() == 8);
            for (gms::Pose& p : poses) p.x = 0.25f;  // what the hook returns is what gets weighted and integrated
        });
        gms::Observation z;
        for (int b = 0; b < 30; b++) z.addMeasurement(gms::Measurement(2 * M_PI * b / 30, 1.0, true));
        slam.update(z, gms::Odometry(0.01, 0.0));
        CHECK(calls == 1);
        for (const gms::Particle& p : slam.getParticles()) CHECK(p.pose.x == 0.25f);
        slam.setPoseOptimizer(nullptr);  // back to the identity default
        slam.update(z, gms::Odometry(0.0, 0.0));
        CHECK(calls == 1);
    }
    // error behaviour: bad configuration is reported, not aborted on
    try {
        gms::SLAM bad([](gms_config& c) { c.num_particles = 0; });
        return 1;
    } catch (const gms::Error& e) {
        CHECK(e.code == GMS_ERR_INVALID_ARG);
    }
    gms::Odometry o(960, 960);  // one wheel revolution on both sides
    CHECK(std::fabs(o.dCenter - (double)(float)M_PI * 0.063) < 1e-15 && o.dTheta == 0.0);
    std::puts("mirror_test OK");
    return 0;
}
