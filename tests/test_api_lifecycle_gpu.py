"""GPU: the object model behind the public API against the unmodified reference — bodies, shapes and joints created and
destroyed mid-run (slot and proxy-id reuse), state edits, impulses, mouse-joint targets, s2World_QueryAABB, s2Shape_TestPoint
and s2World_Draw. The same scripted session is played on both libraries; with the reference's Gauss-Seidel order imposed
(validation schedule) every body must stay bit-identical after every phase."""
import ctypes as C

import numpy as np
import pytest

from oracle import ref as refmod
from solver2d_b200 import capi, device, scenes
from solver2d_b200.capi import Capsule, Circle, Vec2, default_body_def, default_mouse_def, default_revolute_def, default_shape_def

pytestmark = pytest.mark.gpu
DT = 1.0 / 60.0


class Color(C.Structure):
    _fields_ = [("r", C.c_float), ("g", C.c_float), ("b", C.c_float), ("a", C.c_float)]


class Box(C.Structure):
    _fields_ = [("lowerBound", Vec2), ("upperBound", Vec2)]


VP = C.POINTER(Vec2)
DRAW_SIGS = [
    ("DrawPolygon", C.CFUNCTYPE(None, VP, C.c_int, Color, C.c_void_p)),
    ("DrawSolidPolygon", C.CFUNCTYPE(None, VP, C.c_int, Color, C.c_void_p)),
    ("DrawRoundedPolygon", C.CFUNCTYPE(None, VP, C.c_int, C.c_float, Color, Color, C.c_void_p)),
    ("DrawCircle", C.CFUNCTYPE(None, Vec2, C.c_float, Color, C.c_void_p)),
    ("DrawSolidCircle", C.CFUNCTYPE(None, Vec2, C.c_float, Vec2, Color, C.c_void_p)),
    ("DrawCapsule", C.CFUNCTYPE(None, Vec2, Vec2, C.c_float, Color, C.c_void_p)),
    ("DrawSolidCapsule", C.CFUNCTYPE(None, Vec2, Vec2, C.c_float, Color, C.c_void_p)),
    ("DrawSegment", C.CFUNCTYPE(None, Vec2, Vec2, Color, C.c_void_p)),
    ("DrawTransform", C.CFUNCTYPE(None, capi.Transform, C.c_void_p)),
    ("DrawPoint", C.CFUNCTYPE(None, Vec2, C.c_float, Color, C.c_void_p)),
    ("DrawString", C.CFUNCTYPE(None, Vec2, C.c_char_p, C.c_void_p)),
]


class DebugDraw(C.Structure):
    _fields_ = [(n, t) for n, t in DRAW_SIGS] + [("dynamicBodyColor", Color)] + \
        [(n, C.c_bool) for n in ("drawShapes", "drawJoints", "drawAABBs", "drawMass", "drawContactPoints", "drawContactNormals",
                                 "drawContactImpulses", "drawFrictionImpulses")] + [("context", C.c_void_p)]


def _draw_log(lib, world):
    """Play s2World_Draw into a list of (primitive, rounded numbers)."""
    log = []

    def rec(name):
        def f(*args):
            vals = []
            for a in args[:-1]:
                if isinstance(a, Vec2):
                    vals += [a.x, a.y]
                elif isinstance(a, (float, int)):
                    vals.append(float(a))
                elif isinstance(a, capi.Transform):
                    vals += [a.p.x, a.p.y, a.q.s, a.q.c]
                elif isinstance(a, Color):
                    vals += [a.r, a.g, a.b]
                elif hasattr(a, "contents"):
                    n = args[1]
                    for k in range(n):
                        vals += [a[k].x, a[k].y]
            log.append((name, tuple(np.float32(v).item() for v in vals)))
        return f

    dd = DebugDraw()
    keep = []
    for name, sig in DRAW_SIGS:
        cb = sig(rec(name))
        keep.append(cb)
        setattr(dd, name, cb)
    dd.dynamicBodyColor = Color(0.5, 0.6, 0.7, 1.0)
    dd.drawShapes = True
    dd.drawJoints = True
    dd.drawAABBs = True
    dd.drawMass = True
    lib.lib.s2World_Draw.argtypes = [capi.WorldId, C.POINTER(DebugDraw)]
    lib.lib.s2World_Draw.restype = None
    lib.lib.s2World_Draw(world, C.byref(dd))
    return log


QUERY_CB = C.CFUNCTYPE(C.c_bool, capi.ShapeId, C.c_void_p)


def _query(lib, world, lo, hi):
    found = []

    def cb(shape_id, ctx):
        found.append((shape_id.index, shape_id.revision))
        return True
    fn = QUERY_CB(cb)
    lib.lib.s2World_QueryAABB.argtypes = [capi.WorldId, Box, QUERY_CB, C.c_void_p]
    lib.lib.s2World_QueryAABB.restype = None
    lib.lib.s2World_QueryAABB(world, Box(Vec2(*lo), Vec2(*hi)), fn, None)
    return sorted(found)


class Session:
    """One library's side of the scripted session."""

    def __init__(self, lib, solver):
        self.lib = lib
        self.world = lib.create_world(solver)
        self.bodies = {}
        self.shapes = {}
        self.joints = {}

    def body(self, name, kind, pos, shape, angle=0.0):
        lib = self.lib
        bd = default_body_def()
        bd.type = kind
        bd.position = Vec2(*pos)
        bd.angle = angle
        bid = lib.s2CreateBody(self.world, C.byref(bd))
        sd = default_shape_def()
        if shape[0] == "box":
            poly = lib.s2MakeBox(shape[1], shape[2])
            sid = lib.s2CreatePolygonShape(bid, C.byref(sd), C.byref(poly))
        elif shape[0] == "circle":
            c = Circle(Vec2(0.0, 0.0), shape[1])
            sid = lib.s2CreateCircleShape(bid, C.byref(sd), C.byref(c))
        else:
            cap = Capsule(Vec2(-shape[1], 0.0), Vec2(shape[1], 0.0), shape[2])
            sid = lib.s2CreateCapsuleShape(bid, C.byref(sd), C.byref(cap))
        self.bodies[name] = bid
        self.shapes[name] = sid
        return bid

    def revolute(self, name, a, b, pivot, **kw):
        lib = self.lib
        jd = default_revolute_def()
        jd.bodyIdA = self.bodies[a]
        jd.bodyIdB = self.bodies[b]
        jd.localAnchorA = lib.s2Body_GetLocalPoint(self.bodies[a], Vec2(*pivot))
        jd.localAnchorB = lib.s2Body_GetLocalPoint(self.bodies[b], Vec2(*pivot))
        for k, v in kw.items():
            setattr(jd, k, v)
        self.joints[name] = lib.s2CreateRevoluteJoint(self.world, C.byref(jd))

    def mouse(self, name, a, b, target):
        md = default_mouse_def()
        md.bodyIdA = self.bodies[a]
        md.bodyIdB = self.bodies[b]
        md.target = Vec2(*target)
        md.hertz = 4.0
        md.dampingRatio = 0.8
        self.joints[name] = self.lib.s2CreateMouseJoint(self.world, C.byref(md))

    def state(self):
        names = sorted(self.bodies)
        return np.array([tuple(self.lib.s2Body_GetPosition(self.bodies[n])) + (self.lib.s2Body_GetAngle(self.bodies[n]),)
                         for n in names], dtype=np.float64)


def _script(s: Session, phase: int):
    """The edits of each phase (identical calls on both libraries)."""
    L = s.lib
    if phase == 0:
        s.body("ground", capi.STATIC_BODY, (0.0, -1.0), ("box", 30.0, 1.0))
        for i in range(6):
            for j in range(5):
                kind = ("box", 0.5, 0.4) if (i + j) % 3 == 0 else (("circle", 0.45) if (i + j) % 3 == 1 else ("capsule", 0.4, 0.25))
                s.body(f"b{i}_{j}", capi.DYNAMIC_BODY, (-6.0 + 2.2 * j + 0.3 * (i % 2), 0.6 + 1.3 * i), kind, angle=0.1 * (i - j))
        s.body("arm", capi.DYNAMIC_BODY, (8.0, 4.0), ("box", 1.5, 0.2))
        s.revolute("hinge", "ground", "arm", (6.5, 4.0), enableLimit=True, lowerAngle=-0.6, upperAngle=0.9, enableMotor=True,
                   motorSpeed=1.0, maxMotorTorque=50.0)
        s.body("kin", capi.KINEMATIC_BODY, (-10.0, 3.0), ("box", 1.0, 0.3))
        L.s2Body_SetLinearVelocity(s.bodies["kin"], Vec2(1.5, 0.0))
    elif phase == 1:
        # destroy bodies in the middle of the pile (their contacts, shapes and proxies go), then reuse the slots
        for n in ("b1_2", "b2_2", "b3_1"):
            L.s2DestroyBody(s.bodies.pop(n))
            s.shapes.pop(n)
        s.body("new0", capi.DYNAMIC_BODY, (-1.5, 9.0), ("box", 0.6, 0.6))
        s.body("new1", capi.DYNAMIC_BODY, (1.0, 10.0), ("circle", 0.5))
        s.mouse("drag", "ground", "b5_4", (4.0, 9.0))
        L.s2Body_ApplyLinearImpulse(s.bodies["b4_0"], Vec2(3.0, 1.0), L.s2Body_GetPosition(s.bodies["b4_0"]))
    elif phase == 2:
        L.s2MouseJoint_SetTarget(s.joints["drag"], Vec2(-3.0, 8.0))
        L.s2RevoluteJoint_SetMotorSpeed(s.joints["hinge"], -2.0)
        L.s2Body_SetAngularVelocity(s.bodies["new0"], 3.0)
        s.body("late", capi.DYNAMIC_BODY, (8.0, 6.0), ("capsule", 0.5, 0.2))
        s.revolute("link", "arm", "late", (9.5, 4.0), collideConnected=False)
    elif phase == 3:
        L.s2DestroyJoint(s.joints.pop("drag"))
        L.s2RevoluteJoint_EnableLimit(s.joints["hinge"], False)
        L.s2DestroyJoint(s.joints.pop("link"))  # joints must go before their bodies (reference src/body.c:82-83)
        L.s2DestroyBody(s.bodies.pop("late"))
        s.shapes.pop("late")
        s.body("again", capi.DYNAMIC_BODY, (0.0, 12.0), ("box", 0.4, 0.4))


@pytest.mark.parametrize("solver", ["TGS_Soft", "PGS_NGS_Block"])
def test_scripted_session_matches_reference(reference, dev, solver):
    from test_e2e_gpu import _ref_pair_table
    R = reference
    P = capi.Solver2D(device.LIB_PATH)
    sr, sp = Session(R, solver), Session(P, solver)
    dw = None
    for phase in range(4):
        _script(sr, phase)
        _script(sp, phase)
        if dw is None:
            dw = device.DeviceWorld.attach(dev, sp.world)
            dw.set_schedule(device.SCHEDULE_WAVEFRONT)
        for step in range(45):
            R.step_collide(sr.world)
            keys, *_ = _ref_pair_table(R, sr.world)
            dw.set_contact_order(keys)
            R.step_solve(sr.world, DT, 4, 2, True)
            R.step_finalize(sr.world)
            P.step(sp.world, DT, 4, 2, True)
        a, b = sr.state(), sp.state()
        assert np.array_equal(a, b), f"phase {phase}: bodies differ, max {np.abs(a - b).max()}"
        assert R.s2World_GetStatistics(sr.world).contactCount == P.s2World_GetStatistics(sp.world).contactCount
        # queries and debug draw read the same state back
        for lo, hi in (((-7.0, -0.5), (0.0, 4.0)), ((-30.0, -3.0), (30.0, 30.0)), ((7.0, 3.0), (10.0, 7.0))):
            assert _query(R, sr.world, lo, hi) == _query(P, sp.world, lo, hi)
        for name in sorted(sr.shapes)[:12]:
            for pt in ((0.1, 0.5), (-3.0, 1.0), (8.0, 4.0)):
                assert R.s2Shape_TestPoint(sr.shapes[name], Vec2(*pt)) == P.s2Shape_TestPoint(sp.shapes[name], Vec2(*pt))
        dr, dp = _draw_log(R, sr.world), _draw_log(P, sp.world)
        assert len(dr) == len(dp) and len(dr) > 20
        assert sorted(dr) == sorted(dp), "s2World_Draw output differs"
    R.s2DestroyWorld(sr.world)
    P.s2DestroyWorld(sp.world)


def test_edge_cases_match_reference(reference, dev):
    """Empty world, a paused step (dt = 0: reference src/world.c:176-183 still runs the solver with inv_dt = 0), growth of
    every pool after the world has been stepped (device columns re-allocated, solver graph re-captured), and a step with
    zero relax iterations — all against the reference, reference order imposed."""
    from test_e2e_gpu import _ref_pair_table
    R = reference
    P = capi.Solver2D(device.LIB_PATH)
    # 1. an empty world steps (and reads back) without complaint
    we = P.create_world("TGS_Soft")
    P.step(we, DT, 4, 2, True)
    assert P.s2World_GetStatistics(we).bodyCount == 0
    P.s2DestroyWorld(we)

    sr, sp = Session(R, "TGS_Soft"), Session(P, "TGS_Soft")
    for s in (sr, sp):
        s.body("ground", capi.STATIC_BODY, (0.0, -1.0), ("box", 40.0, 1.0))
        for i in range(8):
            s.body(f"a{i}", capi.DYNAMIC_BODY, (-3.0 + 0.9 * i, 0.6 + 0.05 * i), ("box", 0.4, 0.4))
    dw = device.DeviceWorld.attach(dev, sp.world)
    dw.set_schedule(device.SCHEDULE_WAVEFRONT)

    def run(steps, dt, vel, pos):
        for _ in range(steps):
            R.step_collide(sr.world)
            keys, *_ = _ref_pair_table(R, sr.world)
            dw.set_contact_order(keys)
            R.step_solve(sr.world, dt, vel, pos, True)
            R.step_finalize(sr.world)
            P.step(sp.world, dt, vel, pos, True)
        a, b = sr.state(), sp.state()
        assert np.array_equal(a, b), f"max {np.abs(a - b).max()}"

    run(20, DT, 4, 2)
    run(3, 0.0, 4, 2)      # paused
    run(10, DT, 4, 0)      # no relax iterations
    # 2. pools grow well past their initial capacity after stepping
    for s in (sr, sp):
        for k in range(300):
            s.body(f"g{k}", capi.DYNAMIC_BODY, (-15.0 + 0.11 * k, 3.0 + 1.1 * (k % 7)), ("circle", 0.3) if k % 2 else ("box", 0.3, 0.3))
    run(40, DT, 4, 2)
    assert P.s2World_GetStatistics(sp.world).bodyCount == R.s2World_GetStatistics(sr.world).bodyCount == 309
    R.s2DestroyWorld(sr.world)
    P.s2DestroyWorld(sp.world)


def test_body_recreated_in_the_same_slot_at_the_same_place(reference, dev):
    """A resting box is destroyed and a new one is created where it stood: the new body re-uses the body and shape slots
    (and the proxy id) of the old one, so the contact table still holds the OLD shape's pairs under the same keys when the
    next pair pass runs. The pass has to drop those and report the pairs of the new shape in the same pass (the reference
    removes the keys on destroy and re-creates the contacts on its next update); otherwise the new box has no contacts
    until it leaves its fat AABB and sinks into its neighbours."""
    R = reference
    P = capi.Solver2D(device.LIB_PATH)

    def script(lib):
        sc = scenes.vertical_stack(lib, "TGS_Soft", count=4, columns=3)
        for _ in range(40):
            sc.step(DT, 4, 2, True)
        # replace the second box of the middle column by a new one at the same place, at rest
        victim = sc.bodies[1 + 4 + 1]
        pos = lib.s2Body_GetPosition(victim)
        lib.s2DestroyBody(victim)
        bd = default_body_def()
        bd.type = capi.DYNAMIC_BODY
        bd.position = Vec2(pos.x, pos.y)
        bid = lib.s2CreateBody(sc.world, C.byref(bd))
        sd = default_shape_def()
        box = lib.s2MakeSquare(0.5)
        lib.s2CreatePolygonShape(bid, C.byref(sd), C.byref(box))
        sc.bodies[1 + 4 + 1] = bid
        counts, ys = [], []
        for _ in range(30):
            sc.step(DT, 4, 2, True)
            counts.append(lib.s2World_GetStatistics(sc.world).contactCount)
            ys.append(lib.s2Body_GetPosition(bid).y)
        return sc, counts, ys

    sr, cr, yr = script(R)
    sp, cp, yp = script(P)
    assert cp == cr, f"contact counts after re-creating the body: device {cp[:6]}... reference {cr[:6]}..."
    assert max(abs(a - b) for a, b in zip(yr, yp)) < 2e-3, "the re-created box does not rest where the reference's does"
    assert min(yp) > yr[0] - 0.02, "the re-created box sank into its neighbour"
    sr.destroy()
    sp.destroy()
