// gms.hpp — header-only C++17 host-side mirror of the reference's SLAM classes over the C-ABI (gms.h).
//
// Same class and member names, argument meaning and error behaviour as
// java/GridMapGL/src/main/java/com/fmsz/gridmapgl/slam/{SLAM,GridMap,Observation,Odometry,Pose}.java, so code
// written against the reference reads the same:
//
//     gms::SLAM slam;                                        // new SLAM()            SLAM.java:56-62
//     double neff = slam.update(z, u);                       // SLAM.update           SLAM.java:80-131
//     if (neff < slam.getParticles().size() / 2) slam.resample();   //                GridMapApp.java:185-186
//     gms::Pose p = slam.getWeightedPose();                  //                       SLAM.java:165-178
//
// The arithmetic runs in whichever library implements gms.h and is linked in (libgms.so: CUDA, no CPU
// fallback).  Failures surface as gms::Error (the Java code throws unchecked exceptions).
#pragma once
#include <cmath>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "gms.h"

namespace gms {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("gms error " + std::to_string(c) + ": " + m), code(c) {}
};

struct Pose {  // Pose.java:21-34
    float x = 0, y = 0, theta = 0;
    Pose() = default;
    Pose(float x_, float y_, float t_) : x(x_), y(y_), theta(t_) {}
};

struct Measurement {  // Observation.Measurement Observation.java:37-78
    double angle, distance;
    bool wasHit;
    double localX, localY;
    Measurement(double angle_, double distance_, bool hit)  // Observation.java:44-51
        : angle(angle_), distance(distance_), wasHit(hit), localX(distance_ * std::cos(angle_)),
          localY(distance_ * std::sin(angle_)) {}
    static Measurement fromLocal(double x, double y, bool hit) {  // Measurement(x, y, wasHit, dummy) :69-76
        Measurement m(0, 0, hit);
        m.angle = std::atan2(y, x);
        m.distance = std::sqrt(x * x + y * y);
        m.localX = x;
        m.localY = y;
        return m;
    }
};

class Observation {  // Observation.java:29-106
    std::vector<Measurement> measurements;

   public:
    void addMeasurement(float angle, float distance, bool wasHit) { measurements.emplace_back(angle, distance, wasHit); }
    void addMeasurement(const Measurement& m) { measurements.push_back(m); }
    const std::vector<Measurement>& getMeasurements() const { return measurements; }
    int getNumberOfMeasurements() const { return (int)measurements.size(); }
    void reset() { measurements.clear(); }
};

struct Odometry {  // Odometry.java:25-104
    double dCenter, dTheta;
    Odometry(double dCenter_, double dTheta_) : dCenter(dCenter_), dTheta(dTheta_) {}
    Odometry(int leftCount, int rightCount) : dCenter(0), dTheta(0) {  // Odometry.java:41-55
        gms_odometry_from_counts(leftCount, rightCount, &dCenter, &dTheta);
    }
};

class SLAM;

struct GridMapData {  // GridMap.GridMapData GridMap.java:72-74, fetched from the device on access
    SLAM* owner = nullptr;
    int particle = 0;
    std::vector<double> logData() const;
    std::vector<double> likelihoodData() const;
};

struct Particle {  // SLAM.Particle SLAM.java:30-46
    double weight;
    Pose pose;
    GridMapData m;
};

class GridMap {  // GridMap.java — the per-map operators, bound to one handle
    friend class SLAM;
    SLAM* slam = nullptr;

   public:
    float getResolution() const;
    void computeLikelihoodMap(const GridMapData& map);                                        // GridMap.java:233-250
    double probabilityOf(const GridMapData& map, const Observation& obs, const Pose& p);     // GridMap.java:261-294
    void integrateObservation(const GridMapData& map, const Observation& obs, const Pose& p); // GridMap.java:173-191
    void applyMeasurement(const GridMapData& map, float startX, float startY, float endX, float endY,
                          float measuredDistance, bool wasHit);                              // GridMap.java:194-228
};

class SLAM {  // SLAM.java:26-204
    gms_handle* h = nullptr;
    gms_config cfg{};
    gms_info info{};
    GridMap gridMap;
    friend struct GridMapData;
    friend class GridMap;

    std::function<void(int, std::vector<Pose>&)> optimizer;

    void ck(int rc) const {
        if (rc) throw Error(rc, gms_last_error(h));
    }
    static int optimizerTrampoline(void* user, gms_handle*, int32_t first, int32_t count, float* xyt, const double*,
                                   const double*, const uint8_t*, int32_t, double, double) {
        SLAM* self = static_cast<SLAM*>(user);
        try {
            std::vector<Pose> poses;
            poses.reserve(count);
            for (int i = 0; i < count; i++) poses.emplace_back(xyt[3 * i], xyt[3 * i + 1], xyt[3 * i + 2]);
            self->optimizer(first, poses);
            for (int i = 0; i < count; i++) { xyt[3 * i] = poses[i].x; xyt[3 * i + 1] = poses[i].y; xyt[3 * i + 2] = poses[i].theta; }
            return 0;
        } catch (...) {
            return 1;  // no C++ exception may unwind through the C frames
        }
    }
    static void pack(const Observation& z, std::vector<double>& xy, std::vector<double>& d, std::vector<uint8_t>& hit) {
        for (const Measurement& m : z.getMeasurements()) {
            xy.push_back(m.localX);
            xy.push_back(m.localY);
            d.push_back(m.distance);
            hit.push_back(m.wasHit ? 1 : 0);
        }
    }

   public:
    // Defaults are the reference's (500 particles, 6 m x 6 m at 0.05 m, origin (-3,-3)); `tune` may edit the config.
    template <typename F>
    explicit SLAM(F tune) {
        gms_config_default(&cfg);
        tune(cfg);
        int rc = gms_create(&cfg, &h);
        if (rc) throw Error(rc, gms_last_error(nullptr));
        gms_get_info(h, &info);
        gridMap.slam = this;
    }
    SLAM() : SLAM([](gms_config&) {}) {}
    SLAM(const SLAM&) = delete;
    SLAM& operator=(const SLAM&) = delete;
    ~SLAM() { gms_destroy(h); }

    void reset() { ck(gms_reset(h)); }  // SLAM.java:65-77

    // `normals`: the 2*N standard normal draws of Odometry.apply (Odometry.java:80-81), or nullptr for the device's
    double update(const Observation& z, const Odometry& u, const double* normals = nullptr) {  // SLAM.java:80-131
        std::vector<double> xy, d;
        std::vector<uint8_t> hit;
        pack(z, xy, d, hit);
        double neff = 0;
        ck(gms_update(h, xy.data(), d.data(), hit.data(), (int32_t)d.size(), u.dCenter, u.dTheta, normals, &neff));
        return neff;
    }
    void resample(double u01 = -1.0) { ck(gms_resample(h, u01)); }  // SLAM.java:133-153 (u01 = Math.random())
    double calculateNeff() {                                         // SLAM.java:180-190
        double v = 0;
        ck(gms_calculate_neff(h, &v));
        return v;
    }
    Pose getWeightedPose() {  // SLAM.java:165-178
        float p[3];
        ck(gms_get_weighted_pose(h, p));
        return Pose(p[0], p[1], p[2]);
    }
    Particle getStrongestParticle() {  // SLAM.java:196-198
        int32_t idx = 0;
        float p[3];
        double w = 0;
        ck(gms_get_strongest(h, &idx, p, &w));
        return Particle{w, Pose(p[0], p[1], p[2]), GridMapData{this, idx < 0 ? 0 : idx}};
    }
    std::vector<Particle> getParticles() {  // SLAM.java:192-194
        std::vector<float> xyt(3 * (size_t)info.num_particles);
        std::vector<double> w(info.num_particles);
        ck(gms_get_poses(h, xyt.data()));
        ck(gms_get_weights(h, w.data()));
        std::vector<Particle> out;
        out.reserve(w.size());
        for (int i = 0; i < info.num_particles; i++)
            out.push_back(Particle{w[i], Pose(xyt[3 * i], xyt[3 * i + 1], xyt[3 * i + 2]), GridMapData{this, i}});
        return out;
    }
    GridMap& getGridMap() { return gridMap; }  // SLAM.java:200-202
    // GridMap.findBestPoseOptim (GridMap.java:348-369, called per particle at SLAM.java:97) as a batch hook: `f(first,
    // poses)` may replace the poses of the particles [first, first + poses.size()) between the motion sample and the
    // weight.  The reference's own optimiser returns its start pose, so no hook (the default) is the identity.
    void setPoseOptimizer(std::function<void(int, std::vector<Pose>&)> f) {
        optimizer = std::move(f);
        ck(gms_set_pose_optimizer(h, optimizer ? &SLAM::optimizerTrampoline : nullptr, this));
    }
    std::vector<int32_t> getParents() {        // the index i chosen for each m (SLAM.java:147)
        std::vector<int32_t> p(info.num_particles);
        ck(gms_get_parents(h, p.data()));
        return p;
    }
    int gridWidth() const { return info.grid_w; }
    int gridHeight() const { return info.grid_h; }
    gms_handle* handle() { return h; }
};

inline std::vector<double> GridMapData::logData() const {
    std::vector<double> v((size_t)owner->info.grid_w * owner->info.grid_h);
    owner->ck(gms_get_map(owner->h, particle, GMS_MAP_LOG, v.data(), v.size() * 8));
    return v;
}
inline std::vector<double> GridMapData::likelihoodData() const {
    std::vector<double> v((size_t)owner->info.grid_w * owner->info.grid_h);
    owner->ck(gms_get_map(owner->h, particle, GMS_MAP_LIKELIHOOD, v.data(), v.size() * 8));
    return v;
}
inline float GridMap::getResolution() const { return slam->cfg.resolution; }
inline void GridMap::computeLikelihoodMap(const GridMapData& map) { slam->ck(gms_map_compute_likelihood(slam->h, map.particle)); }
inline double GridMap::probabilityOf(const GridMapData& map, const Observation& obs, const Pose& p) {
    std::vector<double> xy, d;
    std::vector<uint8_t> hit;
    SLAM::pack(obs, xy, d, hit);
    const float pose[3] = {p.x, p.y, p.theta};
    double prob = 0;
    slam->ck(gms_map_probability_of(slam->h, map.particle, pose, xy.data(), hit.data(), (int32_t)d.size(), nullptr, &prob));
    return prob;
}
inline void GridMap::integrateObservation(const GridMapData& map, const Observation& obs, const Pose& p) {
    std::vector<double> xy, d;
    std::vector<uint8_t> hit;
    SLAM::pack(obs, xy, d, hit);
    const float pose[3] = {p.x, p.y, p.theta};
    slam->ck(gms_map_integrate_observation(slam->h, map.particle, pose, xy.data(), d.data(), hit.data(), (int32_t)d.size()));
}
inline void GridMap::applyMeasurement(const GridMapData& map, float sx, float sy, float ex, float ey, float meas, bool hit) {
    slam->ck(gms_map_apply_measurement(slam->h, map.particle, sx, sy, ex, ey, meas, hit ? 1 : 0));
}

}  // namespace gms
