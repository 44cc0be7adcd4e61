"""GPU box: dump the constraint graph of a scene as the solver coloured it (one row per contact constraint: contact slot,
bodyA, bodyB, invMassA > 0, invMassB > 0, colour group) for colouring experiments on the CPU.  Output: gpurun_out/graph_<scene>.npy"""
import sys
import numpy as np
from solver2d_b200 import capi, device, scenes

base = int(sys.argv[1]) if len(sys.argv) > 1 else 447
scene = sys.argv[2] if len(sys.argv) > 2 else "pyramid"
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 12
P = capi.Solver2D(device.LIB_PATH)
dev = device.Device()
sc = scenes.pyramid(P, "TGS_Soft", base_count=base) if scene == "pyramid" else scenes.tumbler(P, "TGS_Soft", grid=base)
dw = device.DeviceWorld.attach(dev, sc.world)
hist = []
for i in range(warm):
    sc.step(1 / 60, 4, 2, True)
    c = dw.counters()
    hist.append((c.constraintCount, c.groupCount))
print("constraints, colours per step:", hist)
c = dw.counters()
rows = dw.download_contacts(c.contactCount + 16)
bodies = dw.download_all_bodies(c.bodyCapacity)
items, sizes = dw.solve_order(c.contactCount + c.jointCapacity + 16)
group = np.full(len(rows), -1, dtype=np.int32)
pos = 0
for g, n in enumerate(sizes):
    sl = items[pos:pos + n]
    group[sl[sl >= 0]] = g
    pos += n
movable = (bodies["invMass"] != 0) | (bodies["invI"] != 0)
out = np.stack([np.arange(len(rows), dtype=np.int32), rows["bodyA"], rows["bodyB"], movable[rows["bodyA"]].astype(np.int32),
                movable[rows["bodyB"]].astype(np.int32), group, rows["pointCount"]], axis=1)
np.save(f"gpurun_out/graph_{scene}{base}.npy", out)
print("group sizes", sizes.tolist()[:70], "rows", out.shape)
