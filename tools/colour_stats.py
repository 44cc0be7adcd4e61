"""GPU box: colour sizes and first-step cost of a pyramid world."""
import sys, time
from solver2d_b200 import capi, device, scenes
base = int(sys.argv[1]) if len(sys.argv) > 1 else 447
P = capi.Solver2D(device.LIB_PATH); dev = device.Device()
sc = scenes.pyramid(P, "TGS_Soft", base_count=base)
dw = device.DeviceWorld.attach(dev, sc.world)
t0 = time.perf_counter(); sc.step(1/60, 4, 2, True); dw.sync(); t1 = time.perf_counter()
print("first step ms", 1e3*(t1-t0), "stage ms", dw.stage_ms())
for _ in range(11): sc.step(1/60, 4, 2, True)
order, sizes = dw.solve_order(400000)
print("colour sizes", list(map(int, sizes)))
