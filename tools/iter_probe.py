"""Where does a step's time go?  Kernel time of config 2 (4096 robots) as a function of the solver iteration counts (runtime parameters
of vsr_env_desc): T = fixed + vel_iters * per_velocity_iteration + position phase.  Dev tool; results are not physical for counts != 10."""
import sys
import numpy as np
sys.path.insert(0, ".")
from vsrevolution_b200 import _ffi as F
from bench import WORKLOADS
w = WORKLOADS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
from vsrevolution_b200.robot import RobotUtils, Robot
from vsrevolution_b200.builder import build_mapper
from vsrevolution_b200.locomotion import Locomotion, Settings
for vi, pi in ((10, 10), (10, 6), (10, 4), (10, 3), (10, 2), (10, 1), (10, 0)):
    task = Locomotion(w["T"], Locomotion.create_terrain(w["terrain"]), Settings(1 / 60.0, vi, pi), 1.0)
    body = RobotUtils.build_sensorizing_function(w["sensors"])(RobotUtils.build_shape(w["shape"]))
    mapper = build_mapper(w["mapper"])
    t = Robot(None, body)
    G = len(mapper.example_for(t))
    ctx = task._context(mapper.build_for(t)(np.linspace(-0.5, 0.5, G)))
    g = np.random.default_rng(w["seed"]).uniform(-1, 1, (n, G))
    ctx.eval(g[:256])
    ms = []
    for _ in range(3):
        ctx.eval(g)
        ms.append(ctx.last_kernel_ms())
    print(f"vel_iters {vi:2d} pos_iters {pi:2d}: kernel {min(ms):7.2f} ms", flush=True)
    ctx.close()
