import sys
import numpy as np
from oracle import ref as refmod
from solver2d_b200 import capi, device, scenes
DT = 1 / 60
solver = sys.argv[1] if len(sys.argv) > 1 else "TGS_Soft"
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 100
R = refmod.load()
P = capi.Solver2D(device.LIB_PATH)
dev = device.Device()
sr = scenes.tumbler(R, solver, grid=grid)
sp = scenes.tumbler(P, solver, grid=grid)
dw = device.DeviceWorld.attach(dev, sp.world)
for step in range(61):
    sr.step(DT, 4, 2, True)
    sp.step(DT, 4, 2, True)
    if step % 10 == 0 or step < 3:
        c = dw.counters()
        rc, rj = R.constraint_counts(sr.world)
        pr = [tuple(R.s2Body_GetPosition(b)) for b in sr.bodies[:4]]
        pp = [tuple(P.s2Body_GetPosition(b)) for b in sp.bodies[:4]]
        ar, ap = R.s2Body_GetAngle(sr.bodies[1]), P.s2Body_GetAngle(sp.bodies[1])
        allr = np.array([tuple(R.s2Body_GetPosition(b)) for b in sr.bodies])
        allp = np.array([tuple(P.s2Body_GetPosition(b)) for b in sp.bodies])
        print(step, "ref C,J", rc, rj, "ours C,J", c.constraintCount, c.jointCount, "contacts(pairs)", c.contactCount,
              "container ref", pr[1], ar, "ours", pp[1], ap, "max|dpos|", np.abs(allr - allp).max())
