"""TEST INFRASTRUCTURE — never imported by the product (only tests/, __graft_entry__.smoke(), bench.py baselines).

Python handle on the UNMODIFIED reference compiled by ``oracle/Makefile`` (``make ref``) into
``oracle/_ref/libsolver2d_ref.so`` together with our stage-tap shim ``oracle/ref_taps.c``.

The public API is bound through the very same ctypes mirror the product uses (solver2d_b200/capi.py), so one scene
script drives both libraries. The taps expose the reference's internal state as numpy records and let a test run
one step split at the stage boundaries of ``s2World_Step`` (reference src/world.c:120-306).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from solver2d_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libsolver2d_ref.so")
REFERENCE_ROOT = os.environ.get("S2_REFERENCE_ROOT", "/root/reference")


def build_ref(force: bool = False) -> bool:
    """Compile the reference where its sources exist (this container). Returns True if the .so is available."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "src")) and (force or not os.path.exists(REF_SO)):
        subprocess.run(["make", "-C", HERE, "ref", f"REF={REFERENCE_ROOT}"], check=True, capture_output=True)
    return os.path.exists(REF_SO)


def available() -> bool:
    return os.path.exists(REF_SO)


# float-record field offsets (must match oracle/ref_taps.c)
BODY_F = dict(origin=0, position=2, rot=4, v=6, w=8, dp=9, localCenter=11, mass=13, invMass=14, I=15, invI=16,
              force=17, torque=19, linearDamping=20, angularDamping=21, gravityScale=22, rot0=23, dp0=25)
BODY_I = dict(valid=0, type=1, revision=2, shapeList=3)
SHAPE_F = dict(aabb=0, fat=4, friction=8, density=9, restitution=10)
SHAPE_I = dict(valid=0, body=1, type=2, proxyKey=3, next=4, revision=5)
CONTACT_I = dict(valid=0, shapeA=1, shapeB=2, bodyA=3, bodyB=4, pointCount=5, id0=6, id1=7, persisted0=8, persisted1=9,
                 frictionPersisted=10, cacheCount=11, cacheA=12, cacheB=13)
CONTACT_F = dict(friction=0, normal=1, cacheMetric=3, points=4)  # 2 x 20 floats per point from `points`
POINT_F = dict(localAnchorA=0, localAnchorB=2, separation=4, normalImpulse=5, tangentImpulse=6, frictionAnchorA=7,
               frictionAnchorB=9, frictionNormalA=11, frictionNormalB=13)
POINT_STRIDE = 20
JOINT_I = dict(valid=0, type=1, bodyA=2, bodyB=3, collideConnected=4, enableMotor=5, enableLimit=6)
JOINT_F = dict(localOriginAnchorA=0, localOriginAnchorB=2, impulse=4, motorImpulse=6, lowerImpulse=7, upperImpulse=8,
               maxMotorTorque=9, motorSpeed=10, referenceAngle=11, lowerAngle=12, upperAngle=13, hertz=14,
               dampingRatio=15, target=16)


class Reference(capi.Solver2D):
    """The reference library + taps."""

    def __init__(self, path: str = REF_SO):
        super().__init__(path)
        L = self.lib
        sizes = (C.c_int * 8)()
        L.s2ref_record_sizes(sizes)
        (self.body_f, self.body_i, self.shape_f, self.shape_i, self.contact_f, self.contact_i, self.joint_f,
         self.joint_i) = list(sizes)
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int)
        for name in ("s2ref_dump_bodies", "s2ref_dump_shapes", "s2ref_dump_contacts", "s2ref_dump_joints"):
            fn = getattr(L, name)
            fn.restype = None
            fn.argtypes = [C.c_int, fp, ip]
        L.s2ref_load_body_state.restype = None
        L.s2ref_load_body_state.argtypes = [C.c_int, fp]
        L.s2ref_capacities.restype = None
        L.s2ref_capacities.argtypes = [C.c_int, ip]
        L.s2ref_step_collide.restype = None
        L.s2ref_step_collide.argtypes = [C.c_int]
        L.s2ref_step_solve.restype = None
        L.s2ref_step_solve.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.s2ref_step_finalize.restype = None
        L.s2ref_step_finalize.argtypes = [C.c_int]
        L.s2ref_set_contact_impulses.restype = None
        L.s2ref_set_contact_impulses.argtypes = [C.c_int, C.c_int] + [C.c_float] * 4
        for name in ("s2ref_load_contact_impulses", "s2ref_load_joint_impulses"):
            fn = getattr(L, name)
            fn.restype = None
            fn.argtypes = [C.c_int, fp]
        L.s2ref_timed_steps.restype = C.c_double
        L.s2ref_timed_steps.argtypes = [capi.WorldId, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.s2ref_timed_e2e_steps.restype = C.c_double
        L.s2ref_timed_e2e_steps.argtypes = [capi.WorldId, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_int, C.c_void_p]
        L.s2ref_constraint_counts.restype = None
        L.s2ref_constraint_counts.argtypes = [C.c_int, ip]

    # -- taps ---------------------------------------------------------------------------------------------------
    def capacities(self, wid) -> dict:
        out = (C.c_int * 8)()
        self.lib.s2ref_capacities(wid.index, out)
        keys = ("bodyCap", "shapeCap", "contactCap", "jointCap", "bodyCount", "shapeCount", "contactCount", "jointCount")
        return dict(zip(keys, list(out)))

    def _dump(self, fn, wid, cap, nf, ni):
        f = np.zeros((cap, nf), dtype=np.float32)
        n = np.zeros((cap, ni), dtype=np.int32)
        fn(wid.index, f.ctypes.data_as(C.POINTER(C.c_float)), n.ctypes.data_as(C.POINTER(C.c_int)))
        return f, n

    def bodies(self, wid):
        return self._dump(self.lib.s2ref_dump_bodies, wid, self.capacities(wid)["bodyCap"], self.body_f, self.body_i)

    def shapes(self, wid):
        return self._dump(self.lib.s2ref_dump_shapes, wid, self.capacities(wid)["shapeCap"], self.shape_f, self.shape_i)

    def contacts(self, wid):
        return self._dump(self.lib.s2ref_dump_contacts, wid, self.capacities(wid)["contactCap"], self.contact_f,
                          self.contact_i)

    def joints(self, wid):
        return self._dump(self.lib.s2ref_dump_joints, wid, self.capacities(wid)["jointCap"], self.joint_f, self.joint_i)

    def load_body_state(self, wid, f):
        f = np.ascontiguousarray(f, dtype=np.float32)
        self.lib.s2ref_load_body_state(wid.index, f.ctypes.data_as(C.POINTER(C.c_float)))

    def load_contact_impulses(self, wid, f):
        """f: (contactCap, 4) = normal / tangent impulse of both points, by pool slot."""
        f = np.ascontiguousarray(f, dtype=np.float32)
        self.lib.s2ref_load_contact_impulses(wid.index, f.ctypes.data_as(C.POINTER(C.c_float)))

    def load_joint_impulses(self, wid, f):
        """f: (jointCap, 5) = impulse.xy, motor, lower, upper impulse, by pool slot."""
        f = np.ascontiguousarray(f, dtype=np.float32)
        self.lib.s2ref_load_joint_impulses(wid.index, f.ctypes.data_as(C.POINTER(C.c_float)))

    def step_collide(self, wid):
        self.lib.s2ref_step_collide(wid.index)

    def step_solve(self, wid, dt, vel_iters, pos_iters, warm_start=True):
        self.lib.s2ref_step_solve(wid.index, dt, vel_iters, pos_iters, 1 if warm_start else 0)

    def step_finalize(self, wid):
        self.lib.s2ref_step_finalize(wid.index)

    def split_step(self, wid, dt, vel_iters, pos_iters, warm_start=True):
        self.step_collide(wid)
        self.step_solve(wid, dt, vel_iters, pos_iters, warm_start)
        self.step_finalize(wid)

    def constraint_counts(self, wid):
        out = (C.c_int * 2)()
        self.lib.s2ref_constraint_counts(wid.index, out)
        return out[0], out[1]

    def timed_e2e_steps(self, wid, steps, dt, vel_iters, pos_iters, warm_start, body_indices, forces_xy, transforms) -> float:
        return float(self.lib.s2ref_timed_e2e_steps(wid, steps, dt, vel_iters, pos_iters, 1 if warm_start else 0,
                                                    body_indices.ctypes.data, forces_xy.ctypes.data, len(body_indices),
                                                    transforms.ctypes.data))

    def timed_steps(self, wid, steps, dt, vel_iters, pos_iters, warm_start=True) -> float:
        return float(self.lib.s2ref_timed_steps(wid, steps, dt, vel_iters, pos_iters, 1 if warm_start else 0))


_cached = None


def load() -> Reference:
    global _cached
    if _cached is None:
        if not available():
            build_ref()
        _cached = Reference()
    return _cached
