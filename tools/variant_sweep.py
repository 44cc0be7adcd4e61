"""GPU box: every solver variant on SURVEY §8d config 3 (10 k-box motorised tumbler) and config 4 (4 k joints + 16 k
contacts): ms/step and constraint-iterations/s of this library and of the unmodified reference (one host core)."""
import ctypes as C
import json
import sys
import time

from oracle import ref as refmod
from solver2d_b200 import capi, device, scenes

DT = 1.0 / 60.0
VARIANTS = ["Jacobi", "PGS", "PGS_NGS", "PGS_NGS_Block", "PGS_Soft", "SoftStep", "TGS_Sticky", "TGS_Soft", "TGS_NGS", "XPBD"]


def passes(solver, s, e):
    if solver in ("TGS_Soft", "SoftStep"):
        return s * (1 + (1 if e > 0 else 0))
    if solver == "PGS":
        return s
    if solver in ("TGS_NGS", "XPBD"):
        return 2 * s
    return s + e


def main():
    P = capi.Solver2D(device.LIB_PATH)
    dev = device.Device()
    R = refmod.load()
    L = P.lib
    L.s2World_TimedSteps.restype = C.c_float
    L.s2World_TimedSteps.argtypes = [capi.WorldId, C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_bool, C.c_int32]
    rows = []
    # warm-up: until the lattice has fallen and piled up (tumbler ~4 s) / landed on the bridges (config 4 ~7 s); the actual
    # constraint counts are reported with every row
    configs = [("config3_tumbler10k", scenes.tumbler, dict(grid=100), 260),
               ("config4_joints4k_boxes5329", scenes.joint_contact_stress, dict(), 420)]
    only = sys.argv[1].split(",") if len(sys.argv) > 1 else VARIANTS
    which = sys.argv[2] if len(sys.argv) > 2 else ""
    for name, recipe, kw, warm in configs:
        if which not in name:
            continue
        for solver in only:
            so = recipe(P, solver, **kw)
            dw = device.DeviceWorld.attach(dev, so.world)
            for _ in range(warm):
                so.step(DT, 4, 2, True)
            steps = 10
            ms = float(L.s2World_TimedSteps(so.world, steps, DT, 4, 2, True, 1)) / steps
            c = dw.counters()
            so.destroy()
            sr = recipe(R, solver, **kw)
            for _ in range(warm):
                sr.step(DT, 4, 2, True)
            rsteps = 10
            rt = R.timed_steps(sr.world, rsteps, DT, 4, 2, True) / rsteps
            rc, rj = R.constraint_counts(sr.world)
            sr.destroy()
            p = passes(solver, 4, 2)
            row = {"config": name, "solver": solver, "ours_ms": ms, "ours_constraints": c.constraintCount, "ours_joints": c.jointCount,
                   "colours": c.groupCount, "overflow": c.overflowCount, "regions": c.regionCount, "cut": c.cutCount, "cut_colours": c.cutGroupCount,
                   "ours_ci_per_s": (c.constraintCount + c.jointCount) * p / (ms * 1e-3),
                   "ref_ms": rt * 1e3, "ref_constraints": rc, "ref_joints": rj, "ref_ci_per_s": (rc + rj) * p / rt,
                   "speedup_step_time": rt * 1e3 / ms}
            rows.append(row)
            print(json.dumps(row), flush=True)
    print("| config | variant | contact constraints / joints | ours ms/step | reference ms/step | step speed-up | ours c-i/s | colours |")
    print("|---|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['config']} | {r['solver']} | {r['ours_constraints']} / {r['ours_joints']} | {r['ours_ms']:.3f} | {r['ref_ms']:.1f} | "
              f"{r['speedup_step_time']:.0f} x | {r['ours_ci_per_s'] / 1e6:.0f} M | "
              f"{r['colours']}{' +overflow ' + str(r['overflow']) if r['overflow'] else ''}"
              f"{' regions ' + str(r['regions']) + ' cut ' + str(r['cut']) if r['regions'] else ''} |")


if __name__ == "__main__":
    main()
