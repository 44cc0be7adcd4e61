"""Golden fixtures (tests/golden/*.npz, generated from the unmodified reference by tests/golden/make_golden.py):

* CPU: the reference built here still reproduces them bit for bit (a stale fixture or a changed reference build fails);
* GPU: the product, stepped through its public API with the recorded Gauss-Seidel order imposed (validation schedule),
  reproduces them bit for bit — a parity check that needs nothing but the committed files (no reference library)."""
import os

import numpy as np
import pytest

from solver2d_b200 import capi, device, scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VARIANTS = ["Jacobi", "PGS", "PGS_NGS", "PGS_NGS_Block", "PGS_Soft", "SoftStep", "TGS_Sticky", "TGS_Soft", "TGS_NGS", "XPBD"]
DT = 1.0 / 60.0


def _load(solver):
    return np.load(os.path.join(GOLDEN, f"pyramid10_{solver.lower()}_90.npz"))


@pytest.mark.parametrize("solver", ["TGS_Soft", "PGS_NGS_Block", "XPBD"])
def test_reference_reproduces_golden(reference, solver):
    g = _load(solver)
    R = reference
    sc = scenes.pyramid(R, solver, base_count=10)
    marks = set(int(m) for m in g["marks"])
    for step in range(1, max(marks) + 1):
        sc.step(DT, 4, 2, True)
        if step in marks:
            bf, _ = R.bodies(sc.world)
            idx = [b.index for b in sc.bodies]
            assert np.array_equal(bf[idx, 0:2].astype(np.float32).view(np.uint32), g[f"origin_step{step}"].view(np.uint32)), step
            assert np.array_equal(bf[idx, 6:9].astype(np.float32).view(np.uint32), g[f"velocity_step{step}"].view(np.uint32)), step
    sc.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("solver", VARIANTS)
def test_product_reproduces_golden(dev, solver):
    g = _load(solver)
    P = capi.Solver2D(device.LIB_PATH)
    sc = scenes.pyramid(P, solver, base_count=10)
    dw = device.DeviceWorld.attach(dev, sc.world)
    dw.set_schedule(device.SCHEDULE_WAVEFRONT)
    marks = set(int(m) for m in g["marks"])
    keys, off = g["order_keys"], g["order_offsets"]
    for step in range(1, max(marks) + 1):
        dw.set_contact_order(np.ascontiguousarray(keys[off[step - 1]:off[step]]))
        sc.step(DT, 4, 2, True)
        if step in marks:
            pos = np.array([tuple(P.s2Body_GetPosition(b)) for b in sc.bodies], dtype=np.float32)
            assert np.array_equal(pos.view(np.uint32), g[f"origin_step{step}"].view(np.uint32)), f"{solver} step {step}"
    sc.destroy()
