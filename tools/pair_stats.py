"""GPU box: per-step moved-proxy counts and pair-pass cost on the bench workload."""
import sys
from solver2d_b200 import capi, device, scenes
base = int(sys.argv[1]) if len(sys.argv) > 1 else 447
P = capi.Solver2D(device.LIB_PATH)
dev = device.Device()
sc = scenes.pyramid(P, "TGS_Soft", base_count=base)
dw = device.DeviceWorld.attach(dev, sc.world)
for step in range(60):
    sc.step(1 / 60, 4, 2, True)
    c = dw.counters()
    st = dw.stage_ms()
    print(step, "moved", c.movedCount, "contacts", c.contactCount, "constraints", c.constraintCount, "pairPasses", c.pairPassCount,
          "pairs_ms", round(st[0], 3), "contacts_ms", round(st[1], 3), "solve_ms", round(st[2], 3), "replays", c.graphReplays)
