"""Headless scene recipes, written only against the public C API (so they run on either library, see capi.py).

Each recipe restates one of the reference's sample scenes or one of the synthetic configs of SURVEY.md §8(d):

* :func:`pyramid`         — ``Pyramid`` (reference samples/collection/sample_contact.cpp:499-561), ground widened
                            for baseCount > 200 (SURVEY §8d config 2 / headline).
* :func:`vertical_stack`  — small stack of boxes (quick contact test).
* :func:`bridge`          — ``Bridge`` (reference samples/collection/sample_joints.cpp:15-90).
* :func:`joint_contact_stress` — config 4: stacked bridges + a grid of boxes dropped on them.
* :func:`tumbler`         — config 3: motorised hollow container with a grid of small boxes.
* :func:`mixed_shapes`    — circles/capsules/polygons/segments on a ground, exercising all nine manifold functions.

All scenes are deterministic lattices (no RNG). A recipe returns a :class:`Scene` holding the world id and the
body ids in creation order.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

from . import capi
from .capi import (BodyId, Capsule, Circle, Segment, Vec2, default_body_def, default_mouse_def, default_revolute_def,
                   default_shape_def)


@dataclass
class Scene:
    lib: capi.Solver2D
    world: capi.WorldId
    bodies: list = field(default_factory=list)
    joints: list = field(default_factory=list)
    name: str = ""
    piles: dict = field(default_factory=dict)

    def step(self, dt=1.0 / 60.0, vel_iters=4, pos_iters=2, warm_start=True):
        self.lib.s2World_Step(self.world, dt, vel_iters, pos_iters, warm_start)

    def destroy(self):
        self.lib.s2DestroyWorld(self.world)


def _ground(lib, world, half_width, half_height=1.0, y=-1.0):
    bd = default_body_def()
    bd.position = Vec2(0.0, y)
    gid = lib.s2CreateBody(world, C.byref(bd))
    box = lib.s2MakeBox(half_width, half_height)
    sd = default_shape_def()
    lib.s2CreatePolygonShape(gid, C.byref(sd), C.byref(box))
    return gid


def pyramid(lib: capi.Solver2D, solver="TGS_Soft", base_count=10) -> Scene:
    """Pyramid recipe, reference samples/collection/sample_contact.cpp:511-552.

    The reference ground is ``s2MakeBox(100, 1)``; for base_count > 200 the ground half-width is widened to
    ``0.5*N + 50`` so the base row still rests on it.
    """
    world = lib.create_world(solver)
    sc = Scene(lib, world, name=f"pyramid{base_count}")
    half_width = 100.0 if base_count <= 200 else 0.5 * base_count + 50.0
    sc.bodies.append(_ground(lib, world, half_width))

    bd = default_body_def()
    bd.type = capi.DYNAMIC_BODY
    sd = default_shape_def()
    sd.density = 1.0
    h = 0.5
    box = lib.s2MakeSquare(h)
    shift = 1.0 * h
    for i in range(base_count):
        # all terms are exactly representable in float32, so Python floats reproduce the sample's float math
        y = (2.0 * i + 1.0) * shift
        for j in range(i, base_count):
            x = (i + 1.0) * shift + 2.0 * (j - i) * shift - h * base_count
            bd.position = Vec2(x, y)
            bid = lib.s2CreateBody(world, C.byref(bd))
            lib.s2CreatePolygonShape(bid, C.byref(sd), C.byref(box))
            sc.bodies.append(bid)
    return sc


def vertical_stack(lib, solver="TGS_Soft", count=8, columns=1) -> Scene:
    world = lib.create_world(solver)
    sc = Scene(lib, world, name=f"stack{count}x{columns}")
    sc.bodies.append(_ground(lib, world, 40.0))
    bd = default_body_def()
    bd.type = capi.DYNAMIC_BODY
    sd = default_shape_def()
    box = lib.s2MakeSquare(0.5)
    for c in range(columns):
        for i in range(count):
            bd.position = Vec2(-2.0 * (columns - 1) + 4.0 * c, 0.5 + 1.0 * i)
            bid = lib.s2CreateBody(world, C.byref(bd))
            lib.s2CreatePolygonShape(bid, C.byref(sd), C.byref(box))
            sc.bodies.append(bid)
    return sc


def _add_bridge(lib, sc, ground_id, count, xbase, y):
    box = lib.s2MakeBox(0.5, 0.125)
    sd = default_shape_def()
    sd.density = 20.0
    jd = default_revolute_def()
    jd.drawSize = 0.1
    prev = ground_id
    for i in range(count):
        bd = default_body_def()
        bd.type = capi.DYNAMIC_BODY
        bd.position = Vec2(xbase + 0.5 + 1.0 * i, y)
        bd.linearDamping = 0.1
        bd.angularDamping = 0.1
        bid = lib.s2CreateBody(sc.world, C.byref(bd))
        lib.s2CreatePolygonShape(bid, C.byref(sd), C.byref(box))
        sc.bodies.append(bid)
        pivot = Vec2(xbase + 1.0 * i, y)
        jd.bodyIdA = prev
        jd.bodyIdB = bid
        jd.localAnchorA = lib.s2Body_GetLocalPoint(prev, pivot)
        jd.localAnchorB = lib.s2Body_GetLocalPoint(bid, pivot)
        sc.joints.append(lib.s2CreateRevoluteJoint(sc.world, C.byref(jd)))
        prev = bid
    pivot = Vec2(xbase + 1.0 * count, y)
    jd.bodyIdA = prev
    jd.bodyIdB = ground_id
    jd.localAnchorA = lib.s2Body_GetLocalPoint(prev, pivot)
    jd.localAnchorB = lib.s2Body_GetLocalPoint(ground_id, pivot)
    sc.joints.append(lib.s2CreateRevoluteJoint(sc.world, C.byref(jd)))


def bridge(lib, solver="TGS_Soft", count=160) -> Scene:
    """Bridge recipe, reference samples/collection/sample_joints.cpp:39-80."""
    world = lib.create_world(solver)
    sc = Scene(lib, world, name=f"bridge{count}")
    bd = default_body_def()
    gid = lib.s2CreateBody(world, C.byref(bd))
    sc.bodies.append(gid)
    _add_bridge(lib, sc, gid, count, -0.5 * count, 20.0)
    return sc


def joint_contact_stress(lib, solver="TGS_Soft", bridges=25, planks=160, grid=73) -> Scene:
    """``bridges`` suspension bridges stacked 4 m apart plus a grid x grid lattice of boxes
    dropped from above (≈4k revolute joints, ≈16k contacts in steady state at the defaults)."""
    world = lib.create_world(solver)
    sc = Scene(lib, world, name=f"jointstress{bridges}x{planks}")
    gid = _ground(lib, world, 0.5 * planks + 50.0)
    sc.bodies.append(gid)
    for b in range(bridges):
        _add_bridge(lib, sc, gid, planks, -0.5 * planks, 4.0 + 4.0 * b)
    bd = default_body_def()
    bd.type = capi.DYNAMIC_BODY
    sd = default_shape_def()
    box = lib.s2MakeSquare(0.25)
    top = 4.0 + 4.0 * bridges
    pitch = 0.75
    for i in range(grid):
        for j in range(grid):
            bd.position = Vec2((j - 0.5 * (grid - 1)) * pitch, top + 1.0 + i * pitch)
            bid = lib.s2CreateBody(world, C.byref(bd))
            lib.s2CreatePolygonShape(bid, C.byref(sd), C.byref(box))
            sc.bodies.append(bid)
    return sc


def tumbler(lib, solver="TGS_Soft", grid=100, half_extent=30.0) -> Scene:
    """hollow motorised container (four offset-box walls on one dynamic body, pattern of
    reference samples/collection/sample_contact.cpp:323-336 and sample_joints.cpp:266-276) holding a grid x grid
    lattice of 0.25 m boxes."""
    world = lib.create_world(solver)
    sc = Scene(lib, world, name=f"tumbler{grid}")
    bd = default_body_def()
    gid = lib.s2CreateBody(world, C.byref(bd))
    sc.bodies.append(gid)

    bd = default_body_def()
    bd.type = capi.DYNAMIC_BODY
    bd.position = Vec2(0.0, half_extent)
    cid = lib.s2CreateBody(world, C.byref(bd))
    sc.bodies.append(cid)
    sd = default_shape_def()
    sd.density = 5.0
    t = 0.5
    e = half_extent
    for (hx, hy, cx, cy) in ((t, e, e, 0.0), (t, e, -e, 0.0), (e, t, 0.0, e), (e, t, 0.0, -e)):
        wall = lib.s2MakeOffsetBox(hx, hy, Vec2(cx, cy), 0.0)
        lib.s2CreatePolygonShape(cid, C.byref(sd), C.byref(wall))

    jd = default_revolute_def()
    jd.bodyIdA = gid
    jd.bodyIdB = cid
    jd.localAnchorA = Vec2(0.0, half_extent)
    jd.localAnchorB = Vec2(0.0, 0.0)
    jd.enableMotor = True
    jd.motorSpeed = 0.05 * math.pi
    jd.maxMotorTorque = 1.0e8
    sc.joints.append(lib.s2CreateRevoluteJoint(world, C.byref(jd)))

    bd = default_body_def()
    bd.type = capi.DYNAMIC_BODY
    sd = default_shape_def()
    box = lib.s2MakeSquare(0.125)
    pitch = 0.3
    for i in range(grid):
        for j in range(grid):
            bd.position = Vec2((j - 0.5 * (grid - 1)) * pitch, half_extent + (i - 0.5 * (grid - 1)) * pitch)
            bid = lib.s2CreateBody(world, C.byref(bd))
            lib.s2CreatePolygonShape(bid, C.byref(sd), C.byref(box))
            sc.bodies.append(bid)
    return sc


def limited_chains(lib, solver="TGS_Soft", chains=3, links=6) -> Scene:
    """Joint feature coverage: hanging capsule chains whose revolute joints have angle limits (every joint), a motor
    (every other joint), collideConnected on one chain, and a mouse joint dragging the last link of the first chain
    (s2DefaultMouseJointDef pattern of reference samples/sample.cpp). The chains swing into a pile of boxes on the
    ground so that joints and contacts share bodies."""
    world = lib.create_world(solver)
    sc = Scene(lib, world, name=f"chains{chains}x{links}")
    gid = _ground(lib, world, 30.0)
    sc.bodies.append(gid)
    sd = default_shape_def()
    sd.density = 2.0
    cap = Capsule(Vec2(-0.4, 0.0), Vec2(0.4, 0.0), 0.12)
    for c in range(chains):
        x0 = -6.0 + 6.0 * c
        y0 = 6.0
        prev = gid
        for i in range(links):
            bd = default_body_def()
            bd.type = capi.DYNAMIC_BODY
            bd.position = Vec2(x0 + 0.5 + 1.0 * i, y0)
            bid = lib.s2CreateBody(world, C.byref(bd))
            lib.s2CreateCapsuleShape(bid, C.byref(sd), C.byref(cap))
            sc.bodies.append(bid)
            jd = default_revolute_def()
            pivot = Vec2(x0 + 1.0 * i, y0)
            jd.bodyIdA = prev
            jd.bodyIdB = bid
            jd.localAnchorA = lib.s2Body_GetLocalPoint(prev, pivot)
            jd.localAnchorB = lib.s2Body_GetLocalPoint(bid, pivot)
            jd.enableLimit = True
            jd.lowerAngle = -0.25 * math.pi if i > 0 else -0.6 * math.pi
            jd.upperAngle = 0.15 * math.pi if i > 0 else 0.1 * math.pi
            if i % 2 == 1:
                jd.enableMotor = True
                jd.motorSpeed = 0.5 if c % 2 == 0 else -0.5
                jd.maxMotorTorque = 30.0
            jd.collideConnected = c == 1
            sc.joints.append(lib.s2CreateRevoluteJoint(world, C.byref(jd)))
            prev = bid
        if c == 0:
            md = default_mouse_def()
            md.bodyIdA = gid
            md.bodyIdB = prev
            md.target = Vec2(x0 + links + 1.0, y0 - 1.0)
            md.hertz = 5.0
            md.dampingRatio = 0.7
            sc.joints.append(lib.s2CreateMouseJoint(world, C.byref(md)))
    bd = default_body_def()
    bd.type = capi.DYNAMIC_BODY
    box = lib.s2MakeSquare(0.3)
    for i in range(4):
        for j in range(10):
            bd.position = Vec2(-8.0 + 1.7 * j + 0.1 * i, 0.35 + 0.65 * i)
            bid = lib.s2CreateBody(world, C.byref(bd))
            lib.s2CreatePolygonShape(bid, C.byref(sd), C.byref(box))
            sc.bodies.append(bid)
    return sc


def mixed_shapes(lib, solver="TGS_Soft", rows=6, cols=8) -> Scene:
    """Every shape-pair manifold function of reference src/contact.c:139-154 in one pile: polygons, rounded
    polygons, circles and capsules dropped on a polygon ground bordered by two segment walls."""
    world = lib.create_world(solver)
    sc = Scene(lib, world, name=f"mixed{rows}x{cols}")
    gid = _ground(lib, world, 30.0)
    sc.bodies.append(gid)
    sd = default_shape_def()
    for x in (-14.0, 14.0):
        seg = Segment(Vec2(x, 0.0), Vec2(x * 1.2, 12.0))
        lib.s2CreateSegmentShape(gid, C.byref(sd), C.byref(seg))
    bd = default_body_def()
    bd.type = capi.DYNAMIC_BODY
    k = 0
    for i in range(rows):
        for j in range(cols):
            bd.position = Vec2(-10.0 + 2.6 * j + 0.35 * (i % 3), 1.0 + 2.2 * i)
            bd.angle = 0.17 * ((i * cols + j) % 7) - 0.5
            bid = lib.s2CreateBody(world, C.byref(bd))
            kind = k % 4
            if kind == 0:
                box = lib.s2MakeBox(0.6, 0.4)
                lib.s2CreatePolygonShape(bid, C.byref(sd), C.byref(box))
            elif kind == 1:
                circle = Circle(Vec2(0.1, 0.0), 0.5)
                lib.s2CreateCircleShape(bid, C.byref(sd), C.byref(circle))
            elif kind == 2:
                cap = Capsule(Vec2(-0.5, 0.0), Vec2(0.5, 0.1), 0.3)
                lib.s2CreateCapsuleShape(bid, C.byref(sd), C.byref(cap))
            else:
                pts = (Vec2 * 5)(Vec2(-0.6, -0.4), Vec2(0.5, -0.5), Vec2(0.8, 0.1), Vec2(0.1, 0.7), Vec2(-0.7, 0.3))
                hull = lib.s2ComputeHull(pts, 5)
                poly = lib.s2MakePolygon(C.byref(hull))
                poly.radius = 0.05 if (k % 8) == 3 else 0.0
                lib.s2CreatePolygonShape(bid, C.byref(sd), C.byref(poly))
            sc.bodies.append(bid)
            k += 1
    return sc


RECIPES = {
    "pyramid": pyramid,
    "vertical_stack": vertical_stack,
    "bridge": bridge,
    "joint_contact_stress": joint_contact_stress,
    "tumbler": tumbler,
    "mixed_shapes": mixed_shapes,
}


def pyramid_rows(base_count: int):
    """The Pyramid recipe (see ``pyramid``) as device rows built with numpy instead of one API call per box: bodies and
    shapes for ``Device.create_world().upload_*``. For worlds too large to author through ctypes in reasonable time
    (the per-colour kernel roofline probe builds 3 M boxes this way). Values are what the host library would produce for
    s2MakeSquare(0.5) boxes of density 1 on a static ground."""
    import numpy as np
    from . import device as dv
    n = base_count
    nb = n * (n + 1) // 2
    i_idx = np.repeat(np.arange(n), n - np.arange(n))
    j_idx = np.concatenate([np.arange(i, n) for i in range(n)])
    h = np.float32(0.5)
    x = ((i_idx + 1) * 0.5 + 2.0 * (j_idx - i_idx) * 0.5 - 0.5 * n).astype(np.float32)
    y = ((2.0 * i_idx + 1.0) * 0.5).astype(np.float32)
    bodies = np.zeros(nb + 1, dtype=dv.BODY_ROW)
    bodies["index"] = np.arange(nb + 1)
    bodies["flags"] = dv.ROW_VALID | (capi.DYNAMIC_BODY << 1)
    bodies["flags"][0] = dv.ROW_VALID | (capi.STATIC_BODY << 1)
    bodies["origin"][1:, 0] = x
    bodies["origin"][1:, 1] = y
    bodies["origin"][0] = (0.0, -1.0)
    bodies["position"] = bodies["origin"]
    bodies["rot"][:, 1] = 1.0
    bodies["mass"][1:] = 1.0
    bodies["invMass"][1:] = 1.0
    bodies["I"][1:] = np.float32(1.0 / 6.0)
    bodies["invI"][1:] = np.float32(6.0)
    bodies["gravityScale"] = 1.0
    shapes = np.zeros(nb + 1, dtype=dv.SHAPE_ROW)
    shapes["index"] = np.arange(nb + 1)
    polygon = 2  # s2_polygonShape
    shapes["flags"] = dv.ROW_VALID | (polygon << 1) | dv.SHAPE_MOVED | 0x20  # + FRESH
    shapes["flags"][0] = dv.ROW_VALID | (polygon << 1) | 0x20
    shapes["body"] = np.arange(nb + 1)
    shapes["proxyKey"] = np.arange(nb + 1)
    shapes["categoryBits"] = 1
    shapes["maskBits"] = 0xFFFFFFFF
    shapes["friction"] = 0.6
    shapes["count"] = 4
    hw = np.full(nb + 1, h, dtype=np.float32)
    hh = np.full(nb + 1, h, dtype=np.float32)
    hw[0] = max(100.0, 0.5 * n + 50.0)
    hh[0] = 1.0
    verts = np.stack([-hw, -hh, hw, -hh, hw, hh, -hw, hh], axis=1)
    shapes["vertices"][:, :8] = verts
    shapes["normals"][:, :8] = np.array([0.0, -1.0, 1.0, 0.0, 0.0, 1.0, -1.0, 0.0], dtype=np.float32)
    cx, cy = bodies["origin"][:, 0], bodies["origin"][:, 1]
    shapes["aabb"] = np.stack([cx - hw, cy - hh, cx + hw, cy + hh], axis=1)
    m = np.float32(0.1)  # s2_aabbMargin
    shapes["fatAABB"] = np.stack([cx - hw - m, cy - hh - m, cx + hw + m, cy + hh + m], axis=1)
    return bodies, shapes


def pyramid_field_interleaved(lib: capi.Solver2D, solver="TGS_Soft", count=12, base_count=8, pitch=None, only=None) -> Scene:
    """The same piles as :func:`pyramid_field`, but created ROUND-ROBIN — body k of every pile before body k + 1 of any — so
    that the islands are thoroughly interleaved in every pool and table (slots, proxy ids, contact keys). Within one pile
    the creation order is the pyramid recipe's, so each pile evolves bit for bit like a world that holds it alone
    (``only=k`` builds exactly that world, at the same position). ``Scene.bodies`` lists pile 0's bodies first (ground,
    then boxes), then pile 1's ... — ``Scene.piles[k]`` is the slice of pile k."""
    world = lib.create_world(solver)
    sc = Scene(lib, world, name=f"field{count}x{base_count}i")
    h = 0.5
    if pitch is None:
        pitch = float(base_count + 6)
    box = lib.s2MakeSquare(h)
    ground = lib.s2MakeBox(0.5 * base_count + 2.0, 1.0)
    sd = default_shape_def()
    sd.density = 1.0
    which = list(range(count)) if only is None else [only]
    cells = [(i, j) for i in range(base_count) for j in range(i, base_count)]
    per_pile = {k: [] for k in which}
    for k in which:
        bd = default_body_def()
        bd.position = Vec2(k * pitch, -1.0)
        gid = lib.s2CreateBody(world, C.byref(bd))
        lib.s2CreatePolygonShape(gid, C.byref(sd), C.byref(ground))
        per_pile[k].append(gid)
    bd = default_body_def()
    bd.type = capi.DYNAMIC_BODY
    for (i, j) in cells:
        y = (2.0 * i + 1.0) * h
        x = (i + 1.0) * h + 2.0 * (j - i) * h - h * base_count
        for k in which:
            bd.position = Vec2(k * pitch + x, y)
            bid = lib.s2CreateBody(world, C.byref(bd))
            lib.s2CreatePolygonShape(bid, C.byref(sd), C.byref(box))
            per_pile[k].append(bid)
    sc.piles = {}
    for k in which:
        sc.piles[k] = slice(len(sc.bodies), len(sc.bodies) + len(per_pile[k]))
        sc.bodies.extend(per_pile[k])
    return sc


def pyramid_field(lib: capi.Solver2D, solver="TGS_Soft", count=256, base_count=45, first=0, pitch=None) -> Scene:
    """SURVEY §8d config 5, batched: ``count`` independent Pyramid worlds (each with its own static ground) laid side by
    side in ONE s2World, ``pitch`` metres apart, so that a single set of launches steps them all — independent worlds are
    disconnected islands of one constraint graph, which is what makes small worlds big enough for the GPU. ``first`` offsets
    the world indices (rank r of an N-rank run builds worlds r, r + N, ... by passing the right slice)."""
    world = lib.create_world(solver)
    sc = Scene(lib, world, name=f"field{count}x{base_count}")
    h = 0.5
    if pitch is None:
        pitch = float(base_count + 6)
    box = lib.s2MakeSquare(h)
    sd = default_shape_def()
    sd.density = 1.0
    for k in range(count):
        x0 = (first + k) * pitch
        bd = default_body_def()
        bd.position = Vec2(x0, -1.0)
        gid = lib.s2CreateBody(world, C.byref(bd))
        ground = lib.s2MakeBox(0.5 * base_count + 2.0, 1.0)
        lib.s2CreatePolygonShape(gid, C.byref(sd), C.byref(ground))
        sc.bodies.append(gid)
        bd = default_body_def()
        bd.type = capi.DYNAMIC_BODY
        for i in range(base_count):
            y = (2.0 * i + 1.0) * h
            for j in range(i, base_count):
                x = (i + 1.0) * h + 2.0 * (j - i) * h - h * base_count
                bd.position = Vec2(x0 + x, y)
                bid = lib.s2CreateBody(world, C.byref(bd))
                lib.s2CreatePolygonShape(bid, C.byref(sd), C.byref(box))
                sc.bodies.append(bid)
    return sc
