"""ctypes binding of libvsr_eval.so (include/vsr_eval.h).  No fallback: importing the evaluator without the
built CUDA library raises, and creating a context without a CUDA device raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VSR_LIB") or os.path.join(_HERE, "libvsr_eval.so")  # VSR_LIB: A/B builds (tools/)

VSR_OK, VSR_EINVAL, VSR_ECUDA, VSR_ENOMEM, VSR_EDIM, VSR_EUNSUPPORTED, VSR_ESTATE = 0, -1, -2, -3, -4, -5, -6
CTRL_PHASE_SIN, CTRL_MLP_CENTRALIZED, CTRL_MLP_DISTRIBUTED_HOMO, CTRL_MLP_DISTRIBUTED_HETERO, CTRL_TIME_SIN = 0, 1, 2, 3, 4
SENSOR_AREA_RATIO, SENSOR_TOUCH, SENSOR_VELOCITY, SENSOR_ANGLE, SENSOR_CONSTANT, SENSOR_TIME_SIN, SENSOR_LIDAR = 0, 1, 2, 3, 4, 5, 6


class EnvDesc(C.Structure):
    _fields_ = (
        [("dt", C.c_double), ("vel_iters", C.c_int32), ("pos_iters", C.c_int32)]
        + [(n, C.c_double) for n in ("gravity_y", "baumgarte", "linear_tol", "angular_tol", "max_linear_corr",
                                     "max_translation", "max_rotation", "restitution_velocity", "side_length",
                                     "mass_side_ratio", "spring_f", "spring_d", "lin_damping", "ang_damping",
                                     "voxel_mass", "area_ratio_max_delta", "voxel_friction", "voxel_restitution")]
        + [("limit_contraction", C.c_int32), ("scaffolding_mask", C.c_int32)]
        + [(n, C.c_double) for n in ("ground_friction", "ground_restitution", "ground_thickness", "place_x_gap",
                                     "place_y_gap", "final_t", "transient_t")]
        + [("terrain_x", C.POINTER(C.c_double)), ("terrain_y", C.POINTER(C.c_double)), ("n_terrain", C.c_int32),
           ("precision", C.c_int32), ("n_devices", C.c_int32), ("devices", C.c_int32 * 8)]
    )


class SensorDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("axes", C.c_int32), ("rotated", C.c_int32), ("soft_norm", C.c_int32),
                ("dom_min", C.c_double), ("dom_max", C.c_double), ("avg_window", C.c_double),
                ("noise_sigma", C.c_double), ("param", C.c_double), ("n_rays", C.c_int32), ("side", C.c_int32)]

    def key(self) -> tuple:
        return (self.kind, self.axes, self.rotated, self.soft_norm, self.dom_min, self.dom_max, self.avg_window, self.noise_sigma,
                self.param, self.n_rays, self.side)


class Result(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("velocity", "distance", "time", "first_x", "last_x")] + \
        [(n, C.c_int32) for n in ("n_steps", "k_first", "k_last", "contact_overflow")] + \
        [(n, C.c_double) for n in ("control_energy", "area_ratio_abs", "area_ratio_sq", "reserved")]


RESULT_DTYPE = np.dtype([("velocity", "f8"), ("distance", "f8"), ("time", "f8"), ("first_x", "f8"), ("last_x", "f8"),
                         ("n_steps", "i4"), ("k_first", "i4"), ("k_last", "i4"), ("contact_overflow", "i4"),
                         ("control_energy", "f8"), ("area_ratio_abs", "f8"), ("area_ratio_sq", "f8"), ("reserved", "f8")])
assert RESULT_DTYPE.itemsize == C.sizeof(Result)


class EvalOpts(C.Structure):
    """vsr_eval_opts (include/vsr_eval.h)."""
    _fields_ = [("frozen_masks", C.c_void_p), ("terrain_x", C.c_void_p), ("terrain_y", C.c_void_p), ("terrain_off", C.c_void_p),
                ("posture_n", C.c_int32), ("footprint_n", C.c_int32 * 2), ("posture_out", C.c_void_p),
                ("footprints_out", C.c_void_p * 2), ("centers_out", C.c_void_p)]


class VsrError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libvsr_eval error {code}: {message}")
        self.code = code


class VsrIllegalArgument(VsrError, ValueError):
    """Maps VSR_EINVAL / VSR_EDIM, like the IllegalArgumentException the reference mappers throw."""


class GaDesc(C.Structure):
    _fields_ = [("n_pop", C.c_int32), ("xover_prob", C.c_double), ("tournament_rate", C.c_double),
                ("min_tournament", C.c_int32), ("sigma_mut", C.c_double), ("init_min", C.c_double),
                ("init_max", C.c_double), ("seed", C.c_uint64), ("remap", C.c_int32), ("solver", C.c_int32),
                ("n_parents", C.c_int32), ("n_elites", C.c_int32)]


SOLVER_NUM_GA, SOLVER_ES = 0, 1


_lib = None

_SIGS = {
    "vsr_env_default": (C.c_int, [C.POINTER(EnvDesc)]),
    "vsr_ctx_create": (C.c_int, [C.POINTER(EnvDesc), C.POINTER(C.c_void_p)]),
    "vsr_ctx_destroy": (None, [C.c_void_p]),
    "vsr_last_error": (C.c_char_p, [C.c_void_p]),
    "vsr_body_set": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(SensorDesc), C.c_int32]),
    "vsr_body_set_voxels": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(SensorDesc), C.c_void_p]),
    "vsr_sensor_config_voxels": (C.c_int, [C.c_char_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(SensorDesc), C.c_int32, C.c_void_p]),
    "vsr_controller_set": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_double,
                                     C.c_double, C.c_int32]),
    "vsr_genome_len": (C.c_int, [C.c_void_p]),
    "vsr_n_voxels": (C.c_int, [C.c_void_p]),
    "vsr_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "vsr_eval_centers": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsr_eval_terrains": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsr_eval_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsr_eval_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(EvalOpts), C.c_void_p, C.c_void_p]),
    "vsr_stage_eval_opts": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(EvalOpts)]),
    "vsr_broken_masks": (C.c_int, [C.c_void_p, C.c_double, C.c_int64, C.c_int32, C.c_void_p]),
    "vsr_broken_masks_stream": (C.c_int, [C.c_void_p, C.c_double, C.POINTER(C.c_uint64), C.c_int32, C.c_void_p]),
    "vsr_robot_transformation": (C.c_int, [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "vsr_central_voxel": (C.c_int, [C.c_void_p]),
    "vsr_ctx_episode": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "vsr_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsr_last_kernel_ms": (C.c_double, [C.c_void_p]),
    "vsr_kernel_launches": (C.c_int64, [C.c_void_p]),
    "vsr_contact_overflows": (C.c_int64, [C.c_void_p]),
    "vsr_kernel_info": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int32)] * 5),
    "vsr_shape_mask": (C.c_int, [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p, C.c_int32]),
    "vsr_terrain": (C.c_int, [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "vsr_sensor_config": (C.c_int, [C.c_char_p, C.POINTER(SensorDesc), C.c_int32]),
    "vsr_mlp_inner_neurons": (C.c_int, [C.c_int32, C.c_int32, C.c_double, C.c_int32, C.POINTER(C.c_int32)]),
    "vsr_mlp_count_weights": (C.c_int, [C.POINTER(C.c_int32), C.c_int32]),
    "vsr_mlp_weight_index": (C.c_int, [C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "vsr_episode_steps": (C.c_int, [C.c_double, C.c_double, C.c_double] + [C.POINTER(C.c_int32)] * 3),
    "vsr_input_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "vsr_voxel_cells": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "vsr_counts": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int32)] * 4),
    "vsr_java_random_gaussians": (C.c_int, [C.c_int64, C.c_void_p, C.c_int32]),
    "vsr_java_random_ints": (C.c_int, [C.c_int64, C.c_int32, C.c_void_p, C.c_int32]),
}

# include/vsr_ga.h
_GA_SIGS = {
    "vsr_ga_default": (C.c_int, [C.POINTER(GaDesc)]),
    "vsr_ga_tournament_size": (C.c_int, [C.POINTER(GaDesc)]),
    "vsr_es_default": (C.c_int, [C.POINTER(GaDesc), C.c_int32]),
    "vsr_ga_means": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vsr_ga_create": (C.c_int, [C.c_void_p, C.POINTER(GaDesc), C.POINTER(C.c_void_p)]),
    "vsr_ga_destroy": (None, [C.c_void_p]),
    "vsr_ga_last_error": (C.c_char_p, [C.c_void_p]),
    "vsr_ga_set_shard": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "vsr_ga_init": (C.c_int, [C.c_void_p]),
    "vsr_ga_init_begin": (C.c_int, [C.c_void_p]),
    "vsr_ga_init_end": (C.c_int, [C.c_void_p]),
    "vsr_ga_step": (C.c_int, [C.c_void_p]),
    "vsr_ga_step_begin": (C.c_int, [C.c_void_p]),
    "vsr_ga_step_end": (C.c_int, [C.c_void_p]),
    "vsr_ga_pending": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32), C.POINTER(C.c_void_p)]),
    "vsr_ga_sync": (C.c_int, [C.c_void_p]),
    "vsr_ga_best": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsr_ga_population": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsr_ga_offspring": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsr_ga_survivors": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vsr_ga_set_population": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsr_ga_generation": (C.c_int64, [C.c_void_p]),
    "vsr_ga_evaluations": (C.c_int64, [C.c_void_p]),
    "vsr_ga_kernel_launches": (C.c_int64, [C.c_void_p]),
    "vsr_ga_n_pop": (C.c_int32, [C.c_void_p]),
    "vsr_ga_genome_len": (C.c_int32, [C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)
GA_EXPORTED_SYMBOLS = tuple(_GA_SIGS)


def lib():
    """Load libvsr_eval.so.  Raises (never falls back) when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in {**_SIGS, **_GA_SIGS}.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, ctx=None):
    if rc >= 0:
        return rc
    msg = lib().vsr_last_error(ctx)
    msg = msg.decode() if msg else ""
    if rc in (VSR_EINVAL, VSR_EDIM):
        raise VsrIllegalArgument(rc, msg)
    raise VsrError(rc, msg)


def env_default() -> EnvDesc:
    e = EnvDesc()
    check(lib().vsr_env_default(C.byref(e)))
    return e


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """Thin RAII wrapper of vsr_ctx."""

    def __init__(self, env: EnvDesc, terrain_x: np.ndarray, terrain_y: np.ndarray):
        self._h = C.c_void_p()
        tx = np.ascontiguousarray(terrain_x, dtype=np.float64)
        ty = np.ascontiguousarray(terrain_y, dtype=np.float64)
        env.terrain_x = tx.ctypes.data_as(C.POINTER(C.c_double))
        env.terrain_y = ty.ctypes.data_as(C.POINTER(C.c_double))
        env.n_terrain = len(tx)
        check(lib().vsr_ctx_create(C.byref(env), C.byref(self._h)))
        self.env = env

    def close(self):
        if self._h:
            lib().vsr_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def body_set(self, mask: np.ndarray, sensors: list[SensorDesc]):
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        arr = (SensorDesc * max(1, len(sensors)))(*sensors)
        check(lib().vsr_body_set(self._h, m.shape[1], m.shape[0], _ptr(m), arr, len(sensors)), self._h)

    def body_set_voxels(self, mask: np.ndarray, sensors: list[SensorDesc], offsets):
        """vsr_body_set_voxels: voxel v (grid order) carries sensors[offsets[v]:offsets[v + 1]]."""
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        arr = (SensorDesc * max(1, len(sensors)))()
        for i, sd in enumerate(sensors):
            arr[i] = sd
        off = np.ascontiguousarray(offsets, dtype=np.int32)
        check(lib().vsr_body_set_voxels(self._h, m.shape[1], m.shape[0], _ptr(m), arr, _ptr(off)), self._h)

    def controller_set(self, kind: int, layers=(), signals=0, freq=1.0, amp=1.0, input_activation=1):
        arr = (C.c_int32 * max(1, len(layers)))(*layers)
        check(lib().vsr_controller_set(self._h, kind, arr, len(layers), signals, freq, amp, input_activation), self._h)

    @property
    def genome_len(self) -> int:
        return check(lib().vsr_genome_len(self._h), self._h)

    @property
    def n_voxels(self) -> int:
        return check(lib().vsr_n_voxels(self._h), self._h)

    def eval(self, genomes: np.ndarray, want_results: bool = False):
        g = np.ascontiguousarray(genomes, dtype=np.float64)
        if g.ndim != 2 or g.shape[1] != self.genome_len:
            raise VsrIllegalArgument(VSR_EDIM, f"Wrong number of values: {self.genome_len} expected, {g.shape[-1]} found")
        fit = np.empty(g.shape[0], dtype=np.float64)
        res = np.empty(g.shape[0], dtype=RESULT_DTYPE) if want_results else None
        check(lib().vsr_eval(self._h, _ptr(g), g.shape[0], _ptr(fit), _ptr(res)), self._h)
        return (fit, res) if want_results else fit

    def contact_overflows(self) -> int:
        return check(lib().vsr_contact_overflows(self._h), self._h)

    def eval_centers(self, genomes: np.ndarray):
        """(fitness [n], results [n], centers [n][K][2]): vsr_eval plus the robot-centre (x, y) at every kept observation."""
        g = np.ascontiguousarray(genomes, dtype=np.float64)
        if g.ndim != 2 or g.shape[1] != self.genome_len:
            raise VsrIllegalArgument(VSR_EDIM, f"Wrong number of values: {self.genome_len} expected, {g.shape[-1]} found")
        n_steps, k_first, k_last = episode_steps(self.env.final_t, self.env.transient_t, self.env.dt)
        fit = np.empty(g.shape[0], dtype=np.float64)
        res = np.empty(g.shape[0], dtype=RESULT_DTYPE)
        cen = np.empty((g.shape[0], k_last - k_first + 1, 2), dtype=np.float64)
        check(lib().vsr_eval_centers(self._h, _ptr(g), g.shape[0], _ptr(fit), _ptr(res), _ptr(cen)), self._h)
        return fit, res, cen

    def eval_terrains(self, genomes: np.ndarray, terrains, want_results: bool = False):
        """One terrain per robot (`-rnd`): terrains = sequence of (xs, ys) polylines, len == number of genomes."""
        g = np.ascontiguousarray(genomes, dtype=np.float64)
        if g.ndim != 2 or g.shape[1] != self.genome_len:
            raise VsrIllegalArgument(VSR_EDIM, f"Wrong number of values: {self.genome_len} expected, {g.shape[-1]} found")
        if len(terrains) != g.shape[0]:
            raise VsrIllegalArgument(VSR_EINVAL, f"{g.shape[0]} terrains expected, {len(terrains)} found")
        off = np.zeros(g.shape[0] + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(t[0]) for t in terrains])
        xs = np.ascontiguousarray(np.concatenate([np.asarray(t[0], dtype=np.float64) for t in terrains])) if len(terrains) else np.zeros(0)
        ys = np.ascontiguousarray(np.concatenate([np.asarray(t[1], dtype=np.float64) for t in terrains])) if len(terrains) else np.zeros(0)
        fit = np.empty(g.shape[0], dtype=np.float64)
        res = np.empty(g.shape[0], dtype=RESULT_DTYPE) if want_results else None
        check(lib().vsr_eval_terrains(self._h, _ptr(g), g.shape[0], _ptr(xs), _ptr(ys), _ptr(off), _ptr(fit), _ptr(res)), self._h)
        return (fit, res) if want_results else fit

    def eval_ex(self, genomes: np.ndarray, frozen_masks=None, terrains=None, posture_n: int = 0, footprint_n=(0, 0),
                centers: bool = False):
        """vsr_eval_ex: (fitness [n], results [n], extras) with extras = {"posture": [n][p*p] uint8 or None, "footprints": [two [n][K]
        uint8 arrays or None], "centers": [n][K][3] or None}.  frozen_masks: [n] uint32 (robot transformation `broken-...`);
        terrains: per-robot (xs, ys) polylines as in eval_terrains."""
        g = np.ascontiguousarray(genomes, dtype=np.float64)
        if g.ndim != 2 or g.shape[1] != self.genome_len:
            raise VsrIllegalArgument(VSR_EDIM, f"Wrong number of values: {self.genome_len} expected, {g.shape[-1]} found")
        n = g.shape[0]
        _, k_first, k_last = episode_steps(self.env.final_t, self.env.transient_t, self.env.dt)
        K = k_last - k_first + 1
        o = EvalOpts()
        keep = []
        if frozen_masks is not None:
            fm = np.ascontiguousarray(frozen_masks, dtype=np.uint32)
            if fm.shape != (n,):
                raise VsrIllegalArgument(VSR_EINVAL, f"{n} frozen masks expected, {fm.shape} found")
            keep.append(fm); o.frozen_masks = fm.ctypes.data
        if terrains is not None:
            if len(terrains) != n:
                raise VsrIllegalArgument(VSR_EINVAL, f"{n} terrains expected, {len(terrains)} found")
            off = np.zeros(n + 1, dtype=np.int32)
            off[1:] = np.cumsum([len(t[0]) for t in terrains])
            xs = np.ascontiguousarray(np.concatenate([np.asarray(t[0], dtype=np.float64) for t in terrains])) if n else np.zeros(1)
            ys = np.ascontiguousarray(np.concatenate([np.asarray(t[1], dtype=np.float64) for t in terrains])) if n else np.zeros(1)
            keep += [off, xs, ys]
            o.terrain_x, o.terrain_y, o.terrain_off = xs.ctypes.data, ys.ctypes.data, off.ctypes.data
        post = fps = cen = None
        if posture_n:
            post = np.zeros((n, posture_n * posture_n), dtype=np.uint8)
            o.posture_n, o.posture_out = posture_n, post.ctypes.data
        fps = [None, None]
        for q in range(2):
            if footprint_n[q]:
                fps[q] = np.zeros((n, K), dtype=np.uint8)
                o.footprint_n[q] = footprint_n[q]
                o.footprints_out[q] = fps[q].ctypes.data
        if centers:
            cen = np.zeros((n, K, 3), dtype=np.float64)
            o.centers_out = cen.ctypes.data
        fit = np.empty(n, dtype=np.float64)
        res = np.empty(n, dtype=RESULT_DTYPE)
        check(lib().vsr_eval_ex(self._h, _ptr(g), n, C.byref(o), _ptr(fit), _ptr(res)), self._h)
        return fit, res, {"posture": post, "footprints": fps, "centers": cen, "kept": K}

    def stage_eval_opts(self, n: int, frozen_masks=None, terrains=None):
        """vsr_stage_eval_opts: per-robot inputs of the next device-resident batches (eval_device, generation loop); n = 0 clears."""
        if n == 0 or (frozen_masks is None and terrains is None):
            check(lib().vsr_stage_eval_opts(self._h, 0, None), self._h)
            return
        o = EvalOpts()
        keep = []
        if frozen_masks is not None:
            fm = np.ascontiguousarray(frozen_masks, dtype=np.uint32)
            if fm.shape != (n,):
                raise VsrIllegalArgument(VSR_EINVAL, f"{n} frozen masks expected, {fm.shape} found")
            keep.append(fm); o.frozen_masks = fm.ctypes.data
        if terrains is not None:
            if len(terrains) != n:
                raise VsrIllegalArgument(VSR_EINVAL, f"{n} terrains expected, {len(terrains)} found")
            off = np.zeros(n + 1, dtype=np.int32)
            off[1:] = np.cumsum([len(t[0]) for t in terrains])
            xs = np.ascontiguousarray(np.concatenate([np.asarray(t[0], dtype=np.float64) for t in terrains]))
            ys = np.ascontiguousarray(np.concatenate([np.asarray(t[1], dtype=np.float64) for t in terrains]))
            keep += [off, xs, ys]
            o.terrain_x, o.terrain_y, o.terrain_off = xs.ctypes.data, ys.ctypes.data, off.ctypes.data
        check(lib().vsr_stage_eval_opts(self._h, n, C.byref(o)), self._h)

    def broken_masks(self, ratio: float, seed: int, n: int) -> np.ndarray:
        out = np.zeros(n, dtype=np.uint32)
        check(lib().vsr_broken_masks(self._h, ratio, seed, n, _ptr(out)), self._h)
        return out

    def broken_masks_stream(self, ratio: float, state: "JavaRandomState", n: int) -> np.ndarray:
        """broken_masks on a running java.util.Random (state.value is advanced)."""
        out = np.zeros(n, dtype=np.uint32)
        st = C.c_uint64(state.value)
        check(lib().vsr_broken_masks_stream(self._h, ratio, C.byref(st), n, _ptr(out)), self._h)
        state.value = st.value
        return out

    @property
    def central_voxel(self) -> int:
        return check(lib().vsr_central_voxel(self._h), self._h)

    def eval_into(self, genomes: np.ndarray, fitness_out: np.ndarray):
        """Zero-allocation variant for timing loops: genomes [n][G] float64 C-contiguous, fitness_out [n]."""
        check(lib().vsr_eval(self._h, _ptr(genomes), genomes.shape[0], _ptr(fitness_out), None), self._h)

    def eval_device(self, d_genomes: int, n: int, d_fitness: int, d_results: int = 0, stream: int = 0):
        check(lib().vsr_eval_device(self._h, C.c_void_p(d_genomes), n, C.c_void_p(d_fitness),
                                    C.c_void_p(d_results) if d_results else None,
                                    C.c_void_p(stream) if stream else None), self._h)

    def trace(self, genome: np.ndarray, n_steps: int):
        g = np.ascontiguousarray(genome, dtype=np.float64)
        nv = self.n_voxels
        tv = np.zeros((n_steps, nv, 4, 2))
        tb = np.zeros((n_steps, nv * 4, 6))
        tc = np.zeros((n_steps, nv))
        res = np.zeros(1, dtype=RESULT_DTYPE)
        check(lib().vsr_trace(self._h, _ptr(g), n_steps, _ptr(tv), _ptr(tb), _ptr(tc), _ptr(res)), self._h)
        return res[0], tv, tb, tc

    def last_kernel_ms(self) -> float:
        return lib().vsr_last_kernel_ms(self._h)

    def kernel_launches(self) -> int:
        return lib().vsr_kernel_launches(self._h)

    def kernel_info(self) -> dict:
        v = [C.c_int32() for _ in range(5)]
        check(lib().vsr_kernel_info(self._h, *[C.byref(x) for x in v]), self._h)
        return dict(zip(("regs_per_thread", "smem_bytes", "threads_per_cta", "robots_per_warp", "ctas_per_sm"),
                        [x.value for x in v]))

    def counts(self) -> tuple[int, int, int, int]:
        v = [C.c_int32() for _ in range(4)]
        check(lib().vsr_counts(self._h, *[C.byref(x) for x in v]), self._h)
        return tuple(x.value for x in v)

    def input_map(self):
        n = check(lib().vsr_input_map(self._h, None, None, None, 0), self._h)
        a, b, c = (np.zeros(n, dtype=np.int32) for _ in range(3))
        check(lib().vsr_input_map(self._h, _ptr(a), _ptr(b), _ptr(c), n), self._h)
        return a, b, c

    def voxel_cells(self):
        n = check(lib().vsr_voxel_cells(self._h, None, 0), self._h)
        a = np.zeros(n, dtype=np.int32)
        check(lib().vsr_voxel_cells(self._h, _ptr(a), n), self._h)
        return a


class GeneticAlgorithm:
    """Thin RAII wrapper of vsr_ga (include/vsr_ga.h): the device-resident `numGA` generation loop."""

    def __init__(self, ctx: Context, n_pop: int, seed: int = 0, remap: bool = False, solver: int = SOLVER_NUM_GA, **params):
        self._h = C.c_void_p()
        self.ctx = ctx  # the GA borrows the context: keep it alive
        d = GaDesc()
        if solver == SOLVER_ES:  # `ES` = SimpleES(0.35, 0.4) defaults for this population size (Starter.java:179)
            rc = lib().vsr_es_default(C.byref(d), n_pop)
            if rc < 0:
                raise VsrIllegalArgument(rc, "bad ES parameters")
        else:
            check(lib().vsr_ga_default(C.byref(d)))
        d.n_pop, d.seed, d.remap = n_pop, seed, int(remap)
        for k, v in params.items():
            if not hasattr(d, k):
                raise VsrIllegalArgument(VSR_EINVAL, f"unknown GA parameter {k}")
            setattr(d, k, v)
        self.desc = d
        rc = lib().vsr_ga_create(ctx._h, C.byref(d), C.byref(self._h))
        if rc == VSR_EINVAL:
            raise VsrIllegalArgument(rc, "bad GA parameters")
        if rc < 0:
            raise VsrError(rc, "vsr_ga_create failed")
        self.n_pop, self.genome_len = n_pop, ctx.genome_len
        self.tournament_size = lib().vsr_ga_tournament_size(C.byref(d))

    def _ck(self, rc: int):
        if rc >= 0:
            return rc
        msg = lib().vsr_ga_last_error(self._h)
        msg = msg.decode() if msg else ""
        raise (VsrIllegalArgument if rc in (VSR_EINVAL, VSR_EDIM) else VsrError)(rc, msg)

    def close(self):
        if self._h:
            lib().vsr_ga_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_shard(self, rank: int, world: int):
        self._ck(lib().vsr_ga_set_shard(self._h, rank, world))

    def init(self):
        self._ck(lib().vsr_ga_init(self._h))

    def step(self):
        self._ck(lib().vsr_ga_step(self._h))

    def init_begin(self):
        self._ck(lib().vsr_ga_init_begin(self._h))

    def init_end(self):
        self._ck(lib().vsr_ga_init_end(self._h))

    def step_begin(self):
        self._ck(lib().vsr_ga_step_begin(self._h))

    def step_end(self):
        self._ck(lib().vsr_ga_step_end(self._h))

    def pending(self):
        """(device pointer, count, lo, hi, stream) of the fitness vector the generation in flight is filling."""
        p, st = C.c_void_p(), C.c_void_p()
        cnt, lo, hi = C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(lib().vsr_ga_pending(self._h, C.byref(p), C.byref(cnt), C.byref(lo), C.byref(hi), C.byref(st)))
        return p.value, cnt.value, lo.value, hi.value, st.value or 0

    def sync(self):
        self._ck(lib().vsr_ga_sync(self._h))

    def best(self):
        g, f = np.empty(self.genome_len), np.empty(1)
        self._ck(lib().vsr_ga_best(self._h, _ptr(g), _ptr(f)))
        return g, float(f[0])

    def best_fitness(self) -> float:
        f = np.empty(1)
        self._ck(lib().vsr_ga_best(self._h, None, _ptr(f)))
        return float(f[0])

    def population(self):
        g, f = np.empty((self.n_pop, self.genome_len)), np.empty(self.n_pop)
        self._ck(lib().vsr_ga_population(self._h, _ptr(g), _ptr(f)))
        return g, f

    def offspring(self):
        g, f = np.empty((self.n_pop, self.genome_len)), np.empty(self.n_pop)
        p = np.empty((self.n_pop, 2), dtype=np.int32)
        self._ck(lib().vsr_ga_offspring(self._h, _ptr(g), _ptr(p), _ptr(f)))
        return g, p, f

    def means(self) -> np.ndarray:
        m = np.empty(self.genome_len)
        self._ck(lib().vsr_ga_means(self._h, _ptr(m)))
        return m

    def survivors(self) -> np.ndarray:
        o = np.empty(self.n_pop, dtype=np.int32)
        self._ck(lib().vsr_ga_survivors(self._h, _ptr(o)))
        return o

    def set_population(self, genomes: np.ndarray, fitness: np.ndarray):
        g = np.ascontiguousarray(genomes, dtype=np.float64)
        f = np.ascontiguousarray(fitness, dtype=np.float64)
        if g.shape != (self.n_pop, self.genome_len) or f.shape != (self.n_pop,):
            raise VsrIllegalArgument(VSR_EDIM, f"population must be [{self.n_pop}][{self.genome_len}] + [{self.n_pop}]")
        self._ck(lib().vsr_ga_set_population(self._h, _ptr(g), _ptr(f)))

    @property
    def generation(self) -> int:
        return lib().vsr_ga_generation(self._h)

    @property
    def evaluations(self) -> int:
        return lib().vsr_ga_evaluations(self._h)

    @property
    def kernel_launches(self) -> int:
        return lib().vsr_ga_kernel_launches(self._h)


# ---- naming / integer layer ------------------------------------------------------------------------
def shape_mask(name: str) -> np.ndarray:
    w, h = C.c_int32(), C.c_int32()
    check(lib().vsr_shape_mask(name.encode(), C.byref(w), C.byref(h), None, 0))
    m = np.zeros((h.value, w.value), dtype=np.uint8)
    check(lib().vsr_shape_mask(name.encode(), C.byref(w), C.byref(h), _ptr(m), m.size))
    return m


def terrain(name: str):
    n = check(lib().vsr_terrain(name.encode(), None, None, 0))
    xs, ys = np.zeros(n), np.zeros(n)
    check(lib().vsr_terrain(name.encode(), _ptr(xs), _ptr(ys), n))
    return xs, ys


def sensor_config(name: str) -> list[SensorDesc]:
    n = check(lib().vsr_sensor_config(name.encode(), None, 0))
    arr = (SensorDesc * max(1, n))()
    check(lib().vsr_sensor_config(name.encode(), arr, n))
    return [arr[i] for i in range(n)]


def sensor_config_voxels(name: str, mask: np.ndarray):
    """RobotUtils.buildSensorizingFunction(name).apply(body): (per-voxel sensor lists concatenated in grid order, offsets [n_vox + 1])."""
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    nv = int(m.sum())
    n = check(lib().vsr_sensor_config_voxels(name.encode(), m.shape[1], m.shape[0], _ptr(m), None, 0, None))
    arr = (SensorDesc * max(1, n))()
    off = np.zeros(nv + 1, dtype=np.int32)
    check(lib().vsr_sensor_config_voxels(name.encode(), m.shape[1], m.shape[0], _ptr(m), arr, n, _ptr(off)))
    return [arr[i] for i in range(n)], off


def mlp_inner_neurons(n_in: int, n_out: int, ratio: float, n_inner: int) -> list[int]:
    arr = (C.c_int32 * max(1, n_inner))()
    check(lib().vsr_mlp_inner_neurons(n_in, n_out, ratio, n_inner, arr))
    return [arr[i] for i in range(n_inner)]


def mlp_count_weights(neurons) -> int:
    arr = (C.c_int32 * len(neurons))(*neurons)
    return check(lib().vsr_mlp_count_weights(arr, len(neurons)))


def mlp_weight_index(neurons, l: int, j: int, k: int) -> int:
    arr = (C.c_int32 * len(neurons))(*neurons)
    return check(lib().vsr_mlp_weight_index(arr, len(neurons), l, j, k))


def episode_steps(final_t: float, transient_t: float, dt: float = 1.0 / 60.0):
    v = [C.c_int32() for _ in range(3)]
    check(lib().vsr_episode_steps(final_t, transient_t, dt, *[C.byref(x) for x in v]))
    return tuple(x.value for x in v)


class JavaRandomState:
    """Raw 48-bit LCG state of a java.util.Random shared by a run (`new Random(seed)`), advanced by the library calls that draw from it."""

    def __init__(self, seed: int):
        self.value = (seed ^ 0x5DEECE66D) & ((1 << 48) - 1)

    def _next(self, bits: int) -> int:
        self.value = (self.value * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
        v = self.value >> (48 - bits)
        return v - (1 << bits) if bits == 32 and v >= 1 << 31 else v

    def next_double(self) -> float:
        return ((self._next(26) << 27) + self._next(27)) * (1.0 / (1 << 53))

    def next_int(self, bound: int | None = None) -> int:
        """java.util.Random.nextInt() / nextInt(bound)."""
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


def robot_transformation(name: str):
    """RobotUtils.buildRobotTransformation name -> ("identity",) | ("broken", ratio, seed or None for `rnd`)."""
    kind, ratio, seed, rnd = C.c_int32(), C.c_double(), C.c_int64(), C.c_int32()
    rc = lib().vsr_robot_transformation(name.encode(), C.byref(kind), C.byref(ratio), C.byref(seed), C.byref(rnd))
    if rc == VSR_EUNSUPPORTED:
        raise VsrError(rc, f"robot transformation {name!r} is outside this build (malfunctions triggered during the episode)")
    if rc:
        raise VsrIllegalArgument(rc, f"Unknown transformation name: {name}")
    return ("identity",) if kind.value == 0 else ("broken", ratio.value, None if rnd.value else seed.value)


def java_random_ints(seed: int, bound: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.int32)
    check(lib().vsr_java_random_ints(seed, bound, _ptr(out), n))
    return out


def java_random_gaussians(seed: int, n: int) -> np.ndarray:
    out = np.zeros(n)
    check(lib().vsr_java_random_gaussians(seed, _ptr(out), n))
    return out
