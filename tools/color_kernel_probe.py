"""GPU box: HBM roofline of the per-colour contact impulse kernel at a size where one colour no longer fits in L2.
Builds a Pyramid of `base` rows straight from device rows, steps it a few times, then times the TGS_Soft relax pass of
its largest colour with L2 evicted before every launch (s2b_time_color_kernel)."""
import ctypes as C
import json
import sys
import time

from solver2d_b200 import device, scenes


def probe(base: int, reps: int = 10, settle_steps: int = 3):
    dev = device.Device()
    L = dev.lib
    t0 = time.perf_counter()
    bodies, shapes = scenes.pyramid_rows(base)
    dw = dev.create_world(7)
    dw.upload_bodies(bodies, len(bodies))
    dw.upload_shapes(shapes, len(shapes))
    ctx = device.make_context("TGS_Soft", 1.0 / 60.0, 4, 2, True)
    for _ in range(settle_steps):
        dw.step(ctx)
    dw.sync()
    build_s = time.perf_counter() - t0
    n = C.c_int(0)
    ms = float(L.s2b_time_color_kernel(dw.h, C.byref(ctx), reps, C.byref(n)))
    c = dw.counters()
    out = {"boxes": len(bodies) - 1, "contact_constraints": c.constraintCount, "colours": c.groupCount,
           "largest_colour_constraints": n.value, "kernel_ms": ms, "build_and_settle_s": build_s}
    if ms > 0:
        alg = 208.0 * n.value  # bytes per 2-point TGS_Soft constraint-iteration
        out["algorithmic_bytes"] = alg
        out["achieved_GBps"] = alg / (ms * 1e-3) / 1e9
    dw.destroy()
    return out


if __name__ == "__main__":
    base = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    out = probe(base)
    import os
    out["variant"] = os.environ.get("S2B_COLOR_KERNEL", "plain (coalesced loads)")
    print(json.dumps(out))
