"""Debug helper (GPU box): colour schedule vs permuted oracle for one scene / variant under several settings."""
import sys
import numpy as np
sys.path.insert(0, "tests")
from helpers import body_rows_from_ref, contact_rows_from_ref, joint_rows_from_ref
from oracle import port, ref as refmod
from solver2d_b200 import capi, device, scenes

DT = 1.0 / 60.0
solver = sys.argv[1] if len(sys.argv) > 1 else "TGS_Sticky"
R = refmod.load()
O = port.load()
dev = device.Device()
sc = scenes.limited_chains(R, solver)
for _ in range(40):
    sc.step(DT, 4, 2, True)
R.step_collide(sc.world)
bodies = body_rows_from_ref(*R.bodies(sc.world))
contacts, _ = contact_rows_from_ref(*R.contacts(sc.world))
joints = joint_rows_from_ref(*R.joints(sc.world))
live = contacts["pointCount"] > 0
print("contacts", int(live.sum()), "pairs", [(int(a), int(b)) for a, b in zip(contacts["bodyA"][live], contacts["bodyB"][live])])
for (vel, pos, persistent, sched) in ((4, 2, True, 0), (4, 2, False, 0), (1, 0, True, 0), (1, 1, True, 0), (2, 0, True, 0), (4, 0, True, 0),
                                      (4, 2, True, 1)):
    ctx = device.make_context(solver, DT, vel, pos, True)
    dw = dev.create_world(capi.SOLVER[solver])
    dw.upload_bodies(bodies.copy(), len(bodies))
    dw.upload_joints(joints.copy(), len(joints))
    dw.upload_contacts(contacts.copy())
    dw.set_schedule(sched)
    dw.set_persistent(persistent)
    dw.solve(ctx)
    got = dw.download_all_bodies(len(bodies))
    order, group_sizes = dw.solve_order(len(contacts) + len(joints))
    c = dw.counters()
    dw.destroy()
    ob, oc, oj = O.solve(capi.SOLVER[solver], bodies.copy(), contacts.copy(), joints.copy(), ctx, order=order)
    valid = (bodies["flags"] & 1) == 1
    rep = {}
    for name in ("position", "rot", "linearVelocity", "angularVelocity"):
        g = np.ascontiguousarray(got[name][valid]).reshape(int(valid.sum()), -1)
        o = np.ascontiguousarray(ob[name][valid]).reshape(int(valid.sum()), -1)
        bad = np.nonzero((g.view(np.uint32) != o.view(np.uint32)).any(axis=1))[0]
        if len(bad):
            rep[name] = (len(bad), np.nonzero(valid)[0][bad[:8]].tolist())
    print(f"vel={vel} pos={pos} persistent={persistent} sched={sched} groups={c.groupCount} overflow={c.overflowCount} "
          f"group_sizes={list(group_sizes)[:12]} order={list(order)} -> {rep}")
