"""GPU box: the PRODUCTION step (s2b_step: pair update, narrow phase, persistent solver kernel, finalize — what s2World_Step
runs) on a pyramid too large for L2, and what the persistent kernel achieves there against the HBM roofline.
Usage: python tools/production_at_scale.py <base> [steps]    (base 2000 -> 2 001 000 boxes, ~6 M contact constraints)"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solver2d_b200 import device, scenes  # noqa: E402


def main():
    base = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    dev = device.Device()
    L = dev.lib
    L.s2b_last_solve_kernel_ms.restype = C.c_float
    L.s2b_last_solve_kernel_ms.argtypes = [C.c_void_p]
    t0 = time.perf_counter()
    bodies, shapes = scenes.pyramid_rows(base)
    dw = dev.create_world(7)
    dw.upload_bodies(bodies, len(bodies))
    dw.upload_shapes(shapes, len(shapes))
    ctx = device.make_context("TGS_Soft", 1.0 / 60.0, 4, 2, True)
    for _ in range(4):
        dw.step(ctx)
    dw.sync()
    build_s = time.perf_counter() - t0
    kernel_ms, step_ms, stages = [], [], []
    for _ in range(steps):
        dw.flush_l2()
        step_ms.append(dw.timed_steps(ctx, 1))
        kernel_ms.append(float(L.s2b_last_solve_kernel_ms(dw.h)))
        stages.append(dw.stage_ms())
    c = dw.counters()
    nb = len(bodies)
    C_, S, relax = c.constraintCount, 4, True
    alg = S * (C_ * (152.0 + 208.0 * 2) + nb * 100.0) + C_ * 266.0 + nb * 32.0  # SURVEY §8d, as bench.py
    k = float(np.median(kernel_ms))
    peak = 6478.6
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as fh:
            peak = float(json.load(fh)["hbm_gbs"])
    except Exception:
        pass
    out = {"workload": f"pyramid{base}_tgs_soft_s4_e2 through s2b_step (production path)", "boxes": nb - 1, "contact_constraints": C_,
           "colours": c.groupCount, "regions": c.regionCount, "step_ms_median": float(np.median(step_ms)),
           "stage_ms_median": [float(x) for x in np.median(np.array(stages), axis=0)],
           "persistent_kernel_ms_median": k, "algorithmic_bytes_per_launch": alg, "achieved_GBps": alg / (k * 1e-3) / 1e9,
           "peak_GBps": peak, "frac": alg / (k * 1e-3) / 1e9 / peak, "constraint_iters_per_s": C_ * 8 / (float(np.median(step_ms)) * 1e-3),
           "graph_replays": c.graphReplays, "build_and_settle_s": build_s}
    print(json.dumps(out))
    dw.destroy()


if __name__ == "__main__":
    main()
