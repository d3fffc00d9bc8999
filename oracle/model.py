"""Python side of the CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

PARITY UNPINNED -- see oracle/vsr_oracle.h.  This module restates, for the oracle only, the integer /
naming work that surrounds the hot path:

* ``java_random``      java.util.Random (LCG + polar nextGaussian), used by 2dhmsr terrain builders
* ``build_shape``      2dhmsr RobotUtils.buildShape(String)             [U]  (call site Starter.java:405)
* ``create_terrain``   2dhmsr Locomotion.createTerrain(String)          [U]  (call site Starter.java:87,91)
* ``sensor_config``    2dhmsr RobotUtils.buildSensorizingFunction       [U]  (call site Starter.java:404)
* ``inner_neurons``    /root/reference/.../builder/function/MLP.java:28-42   [V]
* ``count_weights``    2dhmsr MultiLayerPerceptron.countWeights         [U]
* ``episode_steps``    the ``while (t < finalT) t += dt`` loop of Locomotion.apply and
                       Outcome.subOutcome's key selection, emulated in IEEE double [U]

and binds ``liboracle`` through ctypes.  Nothing under ``vsrevolution_b200/`` imports this file.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvsr_oracle.so")


# --------------------------------------------------------------------------------------------------
# java.util.Random
# --------------------------------------------------------------------------------------------------
class JavaRandom:
    """Bit-exact java.util.Random (nextInt/nextDouble/nextGaussian)."""

    _M = (1 << 48) - 1

    def __init__(self, seed: int):
        self.seed = (seed ^ 0x5DEECE66D) & self._M
        self._have = False
        self._next_g = 0.0

    def _next(self, bits: int) -> int:
        self.seed = (self.seed * 0x5DEECE66D + 0xB) & self._M
        v = self.seed >> (48 - bits)
        if v >= 1 << (bits - 1) and bits == 32:
            v -= 1 << 32
        return v

    def next_int(self, bound: int | None = None) -> int:
        if bound is None:
            return self._next(32)
        r = self._next(31)
        m = bound - 1
        if (bound & m) == 0:
            return (bound * r) >> 31
        u = r
        while True:
            r = u % bound
            if u - r + m < (1 << 31):
                return r
            u = self._next(31)

    def next_double(self) -> float:
        return ((self._next(26) << 27) + self._next(27)) * (1.0 / (1 << 53))

    def next_gaussian(self) -> float:
        if self._have:
            self._have = False
            return self._next_g
        while True:
            v1 = 2 * self.next_double() - 1
            v2 = 2 * self.next_double() - 1
            s = v1 * v1 + v2 * v2
            if 0 < s < 1:
                break
        mul = math.sqrt(-2 * math.log(s) / s)
        self._next_g = v2 * mul
        self._have = True
        return v1 * mul


# --------------------------------------------------------------------------------------------------
# shapes / terrains / sensors / MLP sizing
# --------------------------------------------------------------------------------------------------
def build_shape(name: str) -> np.ndarray:
    """name -> uint8 mask [h][w], row y=0 is the BOTTOM row (voxel (x,y) sits at (x*L, y*L))."""
    kind, dims = name.split("-")
    w, h = (int(v) for v in dims.split("x"))
    m = np.zeros((h, w), dtype=np.uint8)
    for y in range(h):
        for x in range(w):
            if kind in ("box", "worm"):
                v = True
            elif kind == "biped":
                v = not (y < h // 2 and x >= w // 4 and x < w * 3 // 4)
            elif kind == "tripod":
                v = not (y < h // 2 and x != 0 and x != w - 1 and x != w // 2)
            elif kind == "comb":
                v = y >= h // 2 or x % 2 == 0
            else:
                raise ValueError(f"unknown shape {name}")
            m[y, x] = 1 if v else 0
    return m


TERRAIN_BORDER_HEIGHT = 100.0
TERRAIN_LENGTH = 2000.0
TERRAIN_BORDER_WIDTH = 10.0


def create_terrain(name: str) -> tuple[np.ndarray, np.ndarray]:
    p = name.split("-")
    if name == "flat":
        xs = [0.0, TERRAIN_BORDER_WIDTH, TERRAIN_LENGTH - TERRAIN_BORDER_WIDTH, TERRAIN_LENGTH]
        ys = [TERRAIN_BORDER_HEIGHT, 5.0, 5.0, TERRAIN_BORDER_HEIGHT]
    elif p[0] in ("hilly", "steppy"):
        h, w, seed = float(p[1]), float(p[2]), int(p[3])
        rnd = JavaRandom(seed)
        xs = [0.0, TERRAIN_BORDER_WIDTH]
        ys = [TERRAIN_BORDER_HEIGHT, 0.0]
        while xs[-1] < TERRAIN_LENGTH - TERRAIN_BORDER_WIDTH:
            xs.append(xs[-1] + max(1.0, (rnd.next_gaussian() * 0.25 + 1) * w))
            if p[0] == "hilly":
                ys.append(ys[-1] + rnd.next_gaussian() * h)
            else:
                xs.append(xs[-1] + 0.5)
                ys.append(ys[-1])
                ys.append(ys[-1] + rnd.next_gaussian() * h)
        xs.append(xs[-1] + TERRAIN_BORDER_WIDTH)
        ys.append(TERRAIN_BORDER_HEIGHT)
    elif p[0] in ("downhill", "uphill"):
        angle = float(p[1])
        dy = (TERRAIN_LENGTH - 2 * TERRAIN_BORDER_WIDTH) * math.sin(angle / 180 * math.pi)
        xs = [0.0, TERRAIN_BORDER_WIDTH, TERRAIN_LENGTH - TERRAIN_BORDER_WIDTH, TERRAIN_LENGTH]
        if p[0] == "downhill":
            ys = [TERRAIN_BORDER_HEIGHT + dy, 5 + dy, 5.0, TERRAIN_BORDER_HEIGHT]
        else:
            ys = [TERRAIN_BORDER_HEIGHT, 5.0, 5 + dy, TERRAIN_BORDER_HEIGHT + dy]
    else:
        raise ValueError(f"unknown terrain {name}")
    return np.asarray(xs, dtype=np.float64), np.asarray(ys, dtype=np.float64)


SENSOR_AREA_RATIO, SENSOR_TOUCH, SENSOR_VELOCITY, SENSOR_ANGLE = 0, 1, 2, 3


SENSOR_CONSTANT, SENSOR_TIME_SIN, SENSOR_LIDAR = 4, 5, 6


class VoSensor(C.Structure):
    _fields_ = [("kind", C.c_int), ("axes", C.c_int), ("rotated", C.c_int), ("soft_norm", C.c_int),
                ("dom_min", C.c_double), ("dom_max", C.c_double), ("avg_window", C.c_double),
                ("noise_sigma", C.c_double), ("param", C.c_double), ("n_rays", C.c_int), ("side", C.c_int)]


def _sensor(kind, axes=0, rotated=0, soft_norm=0, dom=(0.0, 1.0), avg=0.0, noise=0.0, param=0.0, n_rays=0, side=0):
    return VoSensor(kind, axes, rotated, soft_norm, dom[0], dom[1], avg, noise, param, n_rays, side)


def sensor_config(name: str) -> list[VoSensor]:
    """'uniform-a+t+vxy-0.01' -> per-voxel sensor list (PREDEFINED_SENSORS of RobotUtils [U])."""
    kind, names, sigma = name.split("-")
    if kind != "uniform":
        raise ValueError(f"unsupported sensor config {name}")
    sigma = float(sigma)
    out = []
    for s in names.split("+"):
        if s == "a":
            out.append(_sensor(SENSOR_AREA_RATIO, soft_norm=1, dom=(0.5, 1.5), noise=sigma))
        elif s == "t":
            out.append(_sensor(SENSOR_TOUCH, avg=0.25, noise=sigma))
        elif s == "r":
            out.append(_sensor(SENSOR_ANGLE, soft_norm=1, dom=(-math.pi, math.pi), noise=sigma))
        elif s in ("vx", "vy", "vxy"):
            axes = {"vx": 1, "vy": 2, "vxy": 3}[s]
            out.append(_sensor(SENSOR_VELOCITY, axes=axes, rotated=1, soft_norm=1, dom=(-8.0, 8.0), avg=0.5, noise=sigma))
        else:
            raise ValueError(f"unsupported sensor {s}")
    return out


def sensor_dim(s: VoSensor) -> int:
    if s.kind == SENSOR_LIDAR:
        return s.n_rays
    return bin(s.axes & 3).count("1") if s.kind == SENSOR_VELOCITY else 1


def _lidar_side(x: float, y: float) -> int:
    """RobotUtils.lidarSide [U]: 0 N, 1 E, 2 S, 3 W."""
    if y > x and y > 1 - x:
        return 2
    if y >= x and y <= 1 - x:
        return 3
    if y < x and y < 1 - x:
        return 0
    return 1


def predefined_sensor(name: str, x: int, y: int, w: int, h: int, sigma: float) -> VoSensor:
    """One entry of RobotUtils' PREDEFINED_SENSORS for the voxel at (x, y) of a w x h grid [U]."""
    rx = x / (w - 1) if w > 1 else float("nan")
    ry = y / (h - 1) if h > 1 else float("nan")
    if name == "a":
        return _sensor(SENSOR_AREA_RATIO, soft_norm=1, dom=(0.5, 1.5), noise=sigma)
    if name == "t":
        return _sensor(SENSOR_TOUCH, avg=0.25, noise=sigma)
    if name == "r":
        return _sensor(SENSOR_ANGLE, soft_norm=1, dom=(-math.pi, math.pi), noise=sigma)
    if name in ("vx", "vy", "vxy"):
        return _sensor(SENSOR_VELOCITY, axes={"vx": 1, "vy": 2, "vxy": 3}[name], rotated=1, soft_norm=1, dom=(-8.0, 8.0), avg=0.5, noise=sigma)
    if name == "px":
        return _sensor(SENSOR_CONSTANT, noise=sigma, param=rx)
    if name == "py":
        return _sensor(SENSOR_CONSTANT, noise=sigma, param=ry)
    if name == "cpg":
        return _sensor(SENSOR_TIME_SIN, soft_norm=2, dom=(-1.0, 1.0), noise=sigma, param=-1.0)
    if name in ("l1", "l5", "lf5"):
        return _sensor(SENSOR_LIDAR, soft_norm=2, dom=(0.0, 10.0), noise=sigma, param=10.0, n_rays=1 if name == "l1" else 5,
                       side=1 if name == "lf5" else _lidar_side(rx, ry))
    raise ValueError(f"unsupported sensor {name}")


def sensor_config_voxels(name: str, mask: np.ndarray) -> tuple[list[VoSensor], list[int]]:
    """RobotUtils.buildSensorizingFunction(name).apply(body) [U]: (the voxels' sensor lists concatenated in grid order, offsets)."""
    h, w = mask.shape
    p = name.split("-")
    sensors, off = [], [0]
    for y in range(h):
        for x in range(w):
            if not mask[y, x]:
                continue
            if name == "empty":
                pass
            elif p[0] == "uniform":
                sigma = float(p[2])
                sensors += [predefined_sensor(s, x, y, w, h, sigma) for s in p[1].split("+")]
            elif p[0] in ("spinedTouch", "spinedTouchSighted"):
                cpg, malfunction, sigma = p[1] == "t", p[2] == "t", float(p[3])
                if malfunction:
                    raise ValueError("the Malfunction sensor is not built")
                if y == 0:
                    sensors.append(predefined_sensor("t", x, y, w, h, sigma))
                if y == h - 1:
                    v = predefined_sensor("vxy", x, y, w, h, sigma)
                    v.soft_norm, v.avg_window = 2, 0.0
                    sensors.append(v)
                a = predefined_sensor("a", x, y, w, h, sigma)
                a.soft_norm = 2
                sensors.append(a)
                if cpg and x == w - 1 and y == h - 1:
                    sensors.append(predefined_sensor("cpg", x, y, w, h, sigma))
                if p[0] == "spinedTouchSighted" and x == w - 1:
                    sensors.append(predefined_sensor("lf5", x, y, w, h, sigma))
            else:
                raise ValueError(f"unsupported sensor config {name}")
            off.append(len(sensors))
    return sensors, off


def java_round(x: float) -> int:
    """Java Math.round(double): floor(x + 0.5)."""
    return int(math.floor(x + 0.5))


def inner_neurons(n_in: int, n_out: int, ratio: float, n_layers: int) -> list[int]:
    """builder/function/MLP.java:28-42 (Java int division truncates toward zero)."""
    def idiv(a, b):
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q
    inner = [0] * n_layers
    center = int(max(2, java_round(n_in * ratio)))
    if n_layers > 1:
        for i in range(n_layers // 2):
            inner[i] = n_in + idiv(center - n_in, n_layers // 2 + 1) * (i + 1)
        for i in range(n_layers // 2, n_layers):
            inner[i] = center + idiv(n_out - center, n_layers // 2 + 1) * (i - n_layers // 2)
    elif n_layers > 0:
        inner[0] = center
    return inner


def count_weights(neurons: list[int]) -> int:
    return sum(neurons[i + 1] * (neurons[i] + 1) for i in range(len(neurons) - 1))


def episode_steps(final_t: float, transient_t: float, dt: float = 1.0 / 60.0) -> tuple[int, int, int]:
    """(n_steps, k_first, k_last): emulates t accumulation in double; kept keys are transient<=t<final."""
    t, k, k_first, k_last = 0.0, 0, 0, 0
    while t < final_t:
        t += dt
        k += 1
        if transient_t <= t < final_t:
            if not k_first:
                k_first = k
            k_last = k
    return k, k_first, k_last


# --------------------------------------------------------------------------------------------------
# ctypes binding
# --------------------------------------------------------------------------------------------------
class VoParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("dt",)] + [("vel_iters", C.c_int), ("pos_iters", C.c_int)] + \
        [(n, C.c_double) for n in ("gravity_y", "baumgarte", "linear_tol", "angular_tol", "max_linear_corr",
                                   "max_translation", "max_rotation", "restitution_velocity", "side_length",
                                   "mass_side_ratio", "spring_f", "spring_d", "lin_damping", "ang_damping",
                                   "voxel_mass", "area_ratio_max_delta", "voxel_friction", "voxel_restitution")] + \
        [("limit_contraction", C.c_int), ("scaffolding_mask", C.c_int)] + \
        [(n, C.c_double) for n in ("ground_friction", "ground_restitution", "ground_thickness", "place_x_gap",
                                   "place_y_gap", "final_t", "transient_t")]


class VoCtrl(C.Structure):
    _fields_ = [("kind", C.c_int), ("freq", C.c_double), ("amp", C.c_double), ("n_layers", C.c_int),
                ("layers", C.c_int * 8), ("signals", C.c_int), ("input_activation", C.c_int)]


class VoResult(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("velocity", "distance", "time", "first_x", "last_x")] + \
        [("n_steps", C.c_int), ("k_first", C.c_int), ("k_last", C.c_int)] + \
        [(n, C.c_int64) for n in ("n_spring_vel", "n_rope_vel", "n_weld_vel", "n_contact_pt_vel", "n_block_vel",
                                  "n_rope_pos", "n_weld_pos", "n_contact_pt_pos", "n_detect_pairs", "n_pos_iters",
                                  "n_manifolds")] + \
        [(n, C.c_double) for n in ("control_energy", "area_ratio_abs", "area_ratio_sq")]


class VoExtras(C.Structure):
    _fields_ = [("frozen_mask", C.c_uint32), ("posture_n", C.c_int), ("posture_out", C.c_void_p), ("posture_counts", C.c_void_p),
                ("footprint_n", C.c_int * 2), ("footprints_out", C.c_void_p * 2), ("centers_out", C.c_void_p),
                ("vox_sensors", C.c_void_p), ("vox_sensor_off", C.c_void_p)]


CTRL_PHASE_SIN, CTRL_MLP_CENTRALIZED, CTRL_MLP_DISTRIBUTED_HOMO, CTRL_MLP_DISTRIBUTED_HETERO, CTRL_TIME_SIN = 0, 1, 2, 3, 4

_lib = None


def build_lib(force: bool = False) -> str:
    src = os.path.join(_HERE, "vsr_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build_lib()
        _lib = C.CDLL(_LIB_PATH)
        _lib.vo_default_params.argtypes = [C.POINTER(VoParams)]
        _lib.vo_genome_len.restype = C.c_int
        _lib.vo_run.restype = C.c_int
        _lib.vo_run_ex.restype = C.c_int
        _lib.vo_eval_batch_vox.restype = C.c_int
        _lib.vo_genome_len_vox.restype = C.c_int
        _lib.vo_central_voxel.restype = C.c_int
        _lib.vo_eval_batch.restype = C.c_int
    return _lib


def default_params(final_t: float = 30.0, transient_t: float = 1.0) -> VoParams:
    p = VoParams()
    lib().vo_default_params(C.byref(p))
    p.final_t = final_t
    p.transient_t = transient_t
    return p


@dataclass
class Model:
    """Everything the oracle needs for one (shape, sensors, controller, terrain) problem."""
    shape: str
    terrain: str
    sensors_name: str
    ctrl_kind: int
    params: VoParams
    freq: float = 1.0
    amp: float = 1.0
    ratio: float = 0.65
    n_inner: int = 1
    signals: int = 1
    input_activation: int = 1
    mask: np.ndarray = field(init=False)
    tx: np.ndarray = field(init=False)
    ty: np.ndarray = field(init=False)
    sensors: list = field(init=False)
    ctrl: VoCtrl = field(init=False)

    def __post_init__(self):
        self.mask = np.ascontiguousarray(build_shape(self.shape))
        self.tx, self.ty = create_terrain(self.terrain)
        nv = int(self.mask.sum())
        # every voxel's own sensor list; `sensors` stays the common list when all voxels agree (the uniform C entry points take it)
        self.vox_sensors, self.vox_off = sensor_config_voxels(self.sensors_name, self.mask) if self.sensors_name else ([], [0] * (nv + 1))
        key = lambda s: (s.kind, s.axes, s.rotated, s.soft_norm, s.dom_min, s.dom_max, s.avg_window, s.noise_sigma, s.param, s.n_rays, s.side)
        lists = [[key(s) for s in self.vox_sensors[self.vox_off[v]:self.vox_off[v + 1]]] for v in range(nv)]
        self.uniform = all(l == lists[0] for l in lists)
        self.sensors = self.vox_sensors[:self.vox_off[1]] if self.uniform else None
        dims = [sum(sensor_dim(s) for s in self.vox_sensors[self.vox_off[v]:self.vox_off[v + 1]]) for v in range(nv)]
        dim, total_dim = dims[0], sum(dims)
        c = VoCtrl()
        c.kind = self.ctrl_kind
        c.freq, c.amp = self.freq, self.amp
        c.signals = self.signals
        c.input_activation = self.input_activation
        if self.ctrl_kind == CTRL_MLP_CENTRALIZED:
            n_in, n_out = total_dim, nv
        elif self.ctrl_kind in (CTRL_MLP_DISTRIBUTED_HOMO, CTRL_MLP_DISTRIBUTED_HETERO):
            n_in, n_out = dim + 4 * self.signals, 1 + 4 * self.signals
        else:
            n_in = n_out = 0
        if self.ctrl_kind not in (CTRL_PHASE_SIN, CTRL_TIME_SIN):
            neurons = [n_in] + inner_neurons(n_in, n_out, self.ratio, self.n_inner) + [n_out]
            c.n_layers = len(neurons)
            for i, n in enumerate(neurons):
                c.layers[i] = n
        self.ctrl = c

    @property
    def n_voxels(self) -> int:
        return int(self.mask.sum())

    @property
    def h(self) -> int:
        return self.mask.shape[0]

    @property
    def w(self) -> int:
        return self.mask.shape[1]

    def _sensor_array(self):
        src = self.sensors or []
        arr = (VoSensor * max(1, len(src)))()
        for i, s in enumerate(src):
            arr[i] = s
        return arr

    def _n_uniform(self) -> int:
        return len(self.sensors) if self.sensors is not None else 0

    def _vox_arrays(self):
        arr = (VoSensor * max(1, len(self.vox_sensors)))()
        for i, s in enumerate(self.vox_sensors):
            arr[i] = s
        off = (C.c_int * len(self.vox_off))(*self.vox_off)
        return arr, off

    def genome_len(self) -> int:
        if not self.uniform:
            arr, off = self._vox_arrays()
            return lib().vo_genome_len_vox(self.w, self.h, self.mask.ctypes.data_as(C.c_void_p), arr, off, C.byref(self.ctrl))
        return lib().vo_genome_len(self.w, self.h, self.mask.ctypes.data_as(C.c_void_p), self._sensor_array(),
                                   self._n_uniform(), C.byref(self.ctrl))

    def run(self, genome: np.ndarray, trace_steps: int = 0):
        g = np.ascontiguousarray(genome, dtype=np.float64)
        nv = self.n_voxels
        tv = np.zeros((trace_steps, nv, 4, 2)) if trace_steps else None
        tb = np.zeros((trace_steps, nv * 4, 6)) if trace_steps else None
        tc = np.zeros((trace_steps, nv)) if trace_steps else None
        res = VoResult()
        ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        ex = VoExtras()
        keep = self._vox_arrays()
        if not self.uniform:
            ex.vox_sensors, ex.vox_sensor_off = C.addressof(keep[0]), C.addressof(keep[1])
        rc = lib().vo_run_ex(C.byref(self.params), self.w, self.h, ptr(self.mask), self._sensor_array(), self._n_uniform(),
                             C.byref(self.ctrl), ptr(self.tx), ptr(self.ty), len(self.tx), ptr(g), len(g),
                             trace_steps, ptr(tv), ptr(tb), ptr(tc), C.byref(ex), C.byref(res))
        if rc:
            raise RuntimeError(f"vo_run failed: {rc}")
        return res, tv, tb, tc

    def run_ex(self, genome: np.ndarray, frozen_mask: int = 0, posture_n: int = 0, footprint_n=(0, 0), centers: bool = False):
        """One episode with the extras of vo_run_ex: returns (result, dict with posture [n*n] uint8, posture_counts, footprints
        (two uint8 [K] arrays or None), centers [K][3] or None)."""
        g = np.ascontiguousarray(genome, dtype=np.float64)
        _, kf, kl = episode_steps(self.params.final_t, self.params.transient_t, self.params.dt)
        K = kl - kf + 1
        ex = VoExtras()
        keep = self._vox_arrays()
        if not self.uniform:
            ex.vox_sensors, ex.vox_sensor_off = C.addressof(keep[0]), C.addressof(keep[1])
        ex.frozen_mask, ex.posture_n = frozen_mask, posture_n
        post = np.zeros(max(1, posture_n * posture_n), dtype=np.uint8)
        cnt = np.zeros(max(1, posture_n * posture_n), dtype=np.int32)
        ex.posture_out, ex.posture_counts = post.ctypes.data, cnt.ctypes.data
        fps = [None, None]
        for q in range(2):
            ex.footprint_n[q] = footprint_n[q]
            if footprint_n[q]:
                fps[q] = np.zeros(K, dtype=np.uint8)
                ex.footprints_out[q] = fps[q].ctypes.data
        cen = np.zeros((K, 3)) if centers else None
        if centers:
            ex.centers_out = cen.ctypes.data
        res = VoResult()
        ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        rc = lib().vo_run_ex(C.byref(self.params), self.w, self.h, ptr(self.mask), self._sensor_array(), self._n_uniform(),
                             C.byref(self.ctrl), ptr(self.tx), ptr(self.ty), len(self.tx), ptr(g), len(g),
                             0, None, None, None, C.byref(ex), C.byref(res))
        if rc:
            raise RuntimeError(f"vo_run_ex failed: {rc}")
        return res, {"posture": post[:posture_n * posture_n], "posture_counts": cnt[:posture_n * posture_n], "footprints": fps,
                     "centers": cen, "kept": K}

    def central_voxel(self) -> int:
        return lib().vo_central_voxel(self.w, self.h, self.mask.ctypes.data_as(C.c_void_p))

    def eval_batch(self, genomes: np.ndarray, n_threads: int = 1, want_results: bool = False):
        g = np.ascontiguousarray(genomes, dtype=np.float64)
        n = g.shape[0]
        fit = np.zeros(n)
        results = (VoResult * n)() if want_results else None
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        if self.uniform:
            rc = lib().vo_eval_batch(C.byref(self.params), self.w, self.h, ptr(self.mask), self._sensor_array(),
                                     self._n_uniform(), C.byref(self.ctrl), ptr(self.tx), ptr(self.ty), len(self.tx),
                                     ptr(g), g.shape[1], n, n_threads, ptr(fit), results)
        else:
            arr, off = self._vox_arrays()
            rc = lib().vo_eval_batch_vox(C.byref(self.params), self.w, self.h, ptr(self.mask), arr, off, C.byref(self.ctrl),
                                         ptr(self.tx), ptr(self.ty), len(self.tx), ptr(g), g.shape[1], n, n_threads, ptr(fit), results)
        if rc:
            raise RuntimeError(f"vo_eval_batch failed: {rc}")
        return (fit, results) if want_results else fit


def config_model(i: int, final_t: float | None = None) -> Model:
    """The BASELINE.json configs."""
    if i in (1, 2, 5):
        return Model("biped-4x3", "flat", "uniform-a-0", CTRL_MLP_CENTRALIZED, default_params(final_t or 30.0))
    if i == 3:
        return Model("worm-8x2", "hilly-1-10-0", "", CTRL_PHASE_SIN, default_params(final_t or 60.0), freq=1.0)
    if i == 4:
        return Model("comb-7x2", "steppy-0.25-1-0", "uniform-t+vxy+a-0", CTRL_MLP_DISTRIBUTED_HOMO,
                     default_params(final_t or 30.0), signals=1)
    raise ValueError(i)
