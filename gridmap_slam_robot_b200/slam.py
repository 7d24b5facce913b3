"""Host-side mirror of the reference's SLAM classes over the C-ABI (include/gms.h).

Same class and member names, argument meaning and error behaviour as
java/GridMapGL/src/main/java/com/fmsz/gridmapgl/slam/{SLAM,GridMap,Observation,Odometry,Pose}.java, so
code written against the reference reads the same here:

    slam = SLAM()                                  # new SLAM()                    SLAM.java:56-62
    neff = slam.update(observation, odometry)      # SLAM.update                   SLAM.java:80-131
    if neff < len(slam.getParticles()) / 2:        # GridMapApp.java:185-186
        slam.resample()
    pose = slam.getWeightedPose()                  # SLAM.java:165-178

The arithmetic runs in libgms.so (CUDA); there is no Python or CPU fallback.  `lib=` lets the tests run
the same mirror over the oracle library.
"""
from __future__ import annotations

import math

import numpy as np

from . import binding as B


class Pose:
    """Pose.java:21-34 — three f32."""

    __slots__ = ("x", "y", "theta")

    def __init__(self, x=0.0, y=0.0, theta=0.0):
        if isinstance(x, Pose):
            x, y, theta = x.x, x.y, x.theta
        self.x, self.y, self.theta = float(np.float32(x)), float(np.float32(y)), float(np.float32(theta))

    def __repr__(self):  # Pose.toString
        return "[%.2f, %.2f] @ %.2f" % (self.x, self.y, math.degrees(self.theta))


class Measurement:
    """Observation.Measurement, Observation.java:37-78."""

    __slots__ = ("angle", "distance", "wasHit", "localX", "localY")

    def __init__(self, angle, distance, wasHit):  # Observation.java:44-51
        self.angle, self.distance, self.wasHit = float(angle), float(distance), bool(wasHit)
        self.localX = self.distance * math.cos(self.angle)
        self.localY = self.distance * math.sin(self.angle)

    @classmethod
    def fromLocal(cls, x, y, wasHit):  # Measurement(double x, double y, boolean wasHit, int dummy) :69-76
        m = cls.__new__(cls)
        m.angle, m.distance, m.wasHit = math.atan2(y, x), math.sqrt(x * x + y * y), bool(wasHit)
        m.localX, m.localY = float(x), float(y)
        return m


class Observation:
    """Observation.java:29-106 — one lidar revolution."""

    def __init__(self):
        self._m = []
        self._arrays = None

    def addMeasurement(self, a, distance=None, wasHit=None):
        if isinstance(a, Measurement):
            self._m.append(a)
        else:  # addMeasurement(float angle, float distance, boolean wasHit) Observation.java:87-89
            self._m.append(Measurement(float(np.float32(a)), float(np.float32(distance)), wasHit))
        self._arrays = None

    def getMeasurements(self):
        return self._m

    def getNumberOfMeasurements(self):
        return len(self._m)

    def reset(self):
        self._m.clear()
        self._arrays = None

    @classmethod
    def fromArrays(cls, beam_xy, beam_dist, beam_hit):
        """Bulk constructor (not in the reference): localX/localY, distance, wasHit as arrays."""
        o = cls()
        xy = np.ascontiguousarray(beam_xy, np.float64).reshape(-1, 2)
        d = np.ascontiguousarray(beam_dist, np.float64)
        h = np.ascontiguousarray(beam_hit, np.uint8)
        o._arrays = (xy, d, h)
        o._m = None
        return o

    def arrays(self):
        """(localX/localY [B,2] f64, distance [B] f64, wasHit [B] u8): the layout of gms_update."""
        if self._arrays is None:
            xy = np.array([[m.localX, m.localY] for m in self._m], np.float64).reshape(-1, 2)
            d = np.array([m.distance for m in self._m], np.float64)
            h = np.array([m.wasHit for m in self._m], np.uint8)
            self._arrays = (xy, d, h)
        return self._arrays


class Odometry:
    """Odometry.java:25-104.  `rng` supplies the standard normal draws of Odometry.apply
    (NormalDistribution.sample() = sd*z + mean); None lets the device draw them (Philox)."""

    def __init__(self, dCenter, dTheta=None, rng=None):
        if dTheta is None:
            raise TypeError("use Odometry(dCenter, dTheta) or Odometry.fromCounts(left, right)")
        self.dCenter, self.dTheta, self.rng = float(dCenter), float(dTheta), rng

    @classmethod
    def fromCounts(cls, leftCount, rightCount, rng=None, lib=None):  # Odometry(int, int) Odometry.java:41-55
        dc, dt = (lib or B.load()).odometry_from_counts(int(leftCount), int(rightCount))
        return cls(dc, dt, rng)


class GridMapData:
    """GridMap.GridMapData GridMap.java:72-74: logData / likelihoodData, fetched from the device on access."""

    def __init__(self, handle, particle):
        self._h, self._p = handle, particle

    @property
    def logData(self):
        return self._h.get_map(self._p, B.MAP_LOG).ravel()

    @property
    def likelihoodData(self):
        return self._h.get_map(self._p, B.MAP_LIKELIHOOD).ravel()

    @property
    def hitCounts(self):
        """(nFree, nOcc) u32 planes — the exact integer form of logData (not in the reference)."""
        return self._h.get_map(self._p, B.MAP_FREE_COUNT), self._h.get_map(self._p, B.MAP_OCC_COUNT)


class GridMap:
    """GridMap.java — geometry + the per-map operators, bound to one gms handle."""

    def __init__(self, handle):
        self._h = handle

    def getResolution(self):
        return float(self._h.cfg.resolution)

    def getWorldSize(self):
        return (self._h.info.world_w, self._h.info.world_h)

    def getPosition(self):
        return (float(self._h.cfg.origin_x), float(self._h.cfg.origin_y))

    def getGridSize(self):
        return (self._h.W, self._h.H)

    def computeLikelihoodMap(self, map: GridMapData):  # GridMap.java:233-250
        self._h.map_compute_likelihood(map._p)

    def probabilityOf(self, map: GridMapData, obs: Observation, p: Pose):  # GridMap.java:261-294
        xy, _, hit = obs.arrays()
        _, prob = self._h.map_probability_of(map._p, (p.x, p.y, p.theta), xy, hit)
        return prob

    def logProbabilityOf(self, map: GridMapData, obs: Observation, p: Pose):
        xy, _, hit = obs.arrays()
        return self._h.map_probability_of(map._p, (p.x, p.y, p.theta), xy, hit)[0]

    def integrateObservation(self, map: GridMapData, obs: Observation, p: Pose):  # GridMap.java:173-191
        xy, d, hit = obs.arrays()
        self._h.map_integrate_observation(map._p, (p.x, p.y, p.theta), xy, d, hit)

    def applyMeasurement(self, map, startX, startY, endX, endY, measuredDistance, wasHit):  # GridMap.java:194-228
        self._h.map_apply_measurement(map._p, startX, startY, endX, endY, measuredDistance, wasHit)


class Particle:
    """SLAM.Particle SLAM.java:30-46: public weight, pose, m."""

    __slots__ = ("weight", "pose", "m")

    def __init__(self, weight, pose, m):
        self.weight, self.pose, self.m = weight, pose, m


class SLAM:
    """SLAM.java:26-204.  Defaults are the reference's (500 particles, 6 m x 6 m at 0.05 m, origin
    (-3,-3)); keyword arguments override gms_config fields (e.g. num_particles=, map_mode=)."""

    def __init__(self, lib: B.Library | None = None, **config):
        self._lib = lib or B.load()
        self._h = self._lib.create(**config)
        self.gridMap = GridMap(self._h)
        self._particles = None

    # -- the step --
    def update(self, z: Observation, u: Odometry) -> float:  # SLAM.java:80-131
        xy, d, hit = z.arrays()
        normals = None
        if u.rng is not None:
            normals = u.rng.standard_normal(size=(self._h.info.local_count, 2))
        self._particles = None
        return self._h.update(xy, d, hit, u.dCenter, u.dTheta, normals)

    def resample(self, u01: float = -1.0):  # SLAM.java:133-153 (u01 replaces Math.random())
        self._particles = None
        self._h.resample(u01)

    def reset(self):  # SLAM.java:65-77
        self._particles = None
        self._h.reset()

    def calculateNeff(self) -> float:  # SLAM.java:180-190
        return self._h.calculate_neff()

    def getWeightedPose(self) -> Pose:  # SLAM.java:165-178
        return Pose(*self._h.weighted_pose())

    def getStrongestParticle(self) -> Particle:  # SLAM.java:196-198
        idx, pose, w = self._h.strongest()
        return Particle(w, Pose(*pose), GridMapData(self._h, max(idx, 0)))

    def getParticles(self):  # SLAM.java:192-194 — poses/weights in one D2H each, maps lazily
        if self._particles is None:
            poses, w = self._h.poses(), self._h.weights()
            self._particles = [Particle(float(w[i]), Pose(*poses[i]), GridMapData(self._h, i)) for i in range(self._h.P)]
        return self._particles

    def getGridMap(self) -> GridMap:  # SLAM.java:200-202
        return self.gridMap

    def getParents(self):
        """Indices chosen by the last resample (not in the reference; SLAM.java:147's `i` per m)."""
        return self._h.parents()

    @property
    def handle(self) -> B.Handle:
        return self._h
