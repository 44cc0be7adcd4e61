"""GPU box, under ncu: step a scene to a steady state, then run a few steps inside a cudaProfilerStart/Stop region, so that
`ncu --profile-from-start off` lists (or captures) exactly those launches. Usage:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/profile_region.py <pyramid|tumbler|joints|field> <size> <warm steps> <profiled steps> [solver] [graph 0/1]"""
import ctypes as C
import sys

import torch

from solver2d_b200 import capi, device, scenes

scene = sys.argv[1] if len(sys.argv) > 1 else "pyramid"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 447
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 12
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
solver = sys.argv[5] if len(sys.argv) > 5 else "TGS_Soft"
graph = int(sys.argv[6]) if len(sys.argv) > 6 else 0  # graphs hide the kernels of the solver stage from a launch list

P = capi.Solver2D(device.LIB_PATH)
dev = device.Device()
if scene == "pyramid":
    sc = scenes.pyramid(P, solver, base_count=size)
elif scene == "tumbler":
    sc = scenes.tumbler(P, solver, grid=size)
elif scene == "joints":
    sc = scenes.joint_contact_stress(P, solver)
else:
    sc = scenes.pyramid_field(P, solver, count=size, base_count=45)
dw = device.DeviceWorld.attach(dev, sc.world)
dw.set_graph(bool(graph))
for _ in range(warm):
    sc.step(1 / 60, 4, 2, True)
dw.sync()
torch.cuda.profiler.start()
for _ in range(steps):
    sc.step(1 / 60, 4, 2, True)
dw.sync()
torch.cuda.profiler.stop()
c = dw.counters()
print("constraints", c.constraintCount, "joints", c.jointCount, "colours", c.groupCount, "overflow", c.overflowCount, "regions", c.regionCount,
      "cut", c.cutCount, "stage ms", dw.stage_ms())
