"""CPU: the product library loads without a GPU, exports every symbol the headers declare, and its host-side geometry and
narrow-phase code (csrc/shared/s2_collide.h — the SAME source the CUDA narrow-phase kernel compiles) is bit-identical to
the unmodified reference on seeded inputs. No world is created here: s2CreateWorld needs a CUDA device by design."""
import ctypes as C
import re

import numpy as np
import pytest

import os

from solver2d_b200 import capi, device

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def product():
    return capi.Solver2D(device.LIB_PATH)


def test_library_exports_every_declared_symbol(product):
    assert product.missing == []
    lib = C.CDLL(device.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "s2b_device.h")).read()
    declared = re.findall(r"^S2B_API [^;(]*?(s2b_\w+)\(", header, flags=re.M)
    assert len(declared) >= 35
    for name in declared + device.ABI_SYMBOLS:
        assert hasattr(lib, name), name
    ext = open(os.path.join(ROOT, "include", "solver2d_b200.h")).read()
    for name in re.findall(r"^S2B_API [^;(]*?(s2World_\w+)\(", ext, flags=re.M):
        assert hasattr(lib, name), name
    for variant in ("Jacobi", "PGS", "PGS_NGS", "PGS_NGS_Block", "PGS_Soft", "TGS_Soft", "TGS_Sticky", "TGS_NGS", "XPBD",
                    "SoftStep"):
        assert hasattr(lib, f"s2Solve_{variant}")


def test_row_struct_sizes_match_numpy_mirrors():
    lib = C.CDLL(device.LIB_PATH)
    out = (C.c_int32 * 6)()
    lib.s2b_abi_sizes(out)
    assert list(out)[:4] == [device.BODY_ROW.itemsize, device.SHAPE_ROW.itemsize, device.JOINT_ROW.itemsize,
                             device.CONTACT_ROW.itemsize]
    assert out[4] == C.sizeof(device.StepContext) and out[5] == C.sizeof(device.Counters)


def _same_bytes(a, b):
    return bytes(a) == bytes(b)


def _poly_eq(p, q):
    n = p.count
    return (p.count == q.count and p.radius == q.radius
            and all(p.vertices[i].x == q.vertices[i].x and p.vertices[i].y == q.vertices[i].y for i in range(n))
            and all(p.normals[i].x == q.normals[i].x and p.normals[i].y == q.normals[i].y for i in range(n)))


def test_polygon_factories_and_mass_match_reference(reference, product):
    R, P = reference, product
    rng = np.random.default_rng(7)
    for _ in range(50):
        hx, hy = rng.uniform(0.05, 3.0, 2)
        ang = rng.uniform(-3, 3)
        c = capi.Vec2(*rng.uniform(-2, 2, 2))
        assert _poly_eq(R.s2MakeBox(hx, hy), P.s2MakeBox(hx, hy))
        a, b = R.s2MakeOffsetBox(hx, hy, c, ang), P.s2MakeOffsetBox(hx, hy, c, ang)
        assert _poly_eq(a, b)
        for dens in (1.0, 20.0):
            ma, mb = R.s2ComputePolygonMass(C.byref(a), dens), P.s2ComputePolygonMass(C.byref(b), dens)
            assert _same_bytes(ma, mb)
        rb = P.s2MakeRoundedBox(hx, hy, 0.1)
        ra = R.s2MakeBox(hx, hy)
        ra.radius = 0.1
        assert _same_bytes(R.s2ComputePolygonMass(C.byref(ra), 2.0), P.s2ComputePolygonMass(C.byref(rb), 2.0))
        p1, p2 = capi.Vec2(*rng.uniform(-1, 1, 2)), capi.Vec2(*rng.uniform(1.5, 3, 2))
        assert _poly_eq(R.s2MakeCapsule(p1, p2, 0.3), P.s2MakeCapsule(p1, p2, 0.3))
        cap = capi.Capsule(p1, p2, 0.3)
        assert _same_bytes(R.s2ComputeCapsuleMass(C.byref(cap), 1.5), P.s2ComputeCapsuleMass(C.byref(cap), 1.5))
        xf = capi.Transform(c, capi.Rot(np.float32(np.sin(ang)), np.float32(np.cos(ang))))
        assert _same_bytes(R.s2ComputePolygonAABB(C.byref(a), xf), P.s2ComputePolygonAABB(C.byref(b), xf))


def test_hull_matches_reference(reference, product):
    R, P = reference, product
    rng = np.random.default_rng(11)
    for trial in range(200):
        n = int(rng.integers(3, 9))
        pts = (capi.Vec2 * n)(*[capi.Vec2(*rng.uniform(-1.5, 1.5, 2)) for _ in range(n)])
        if trial % 5 == 0:  # welded / collinear inputs
            pts[1] = capi.Vec2(pts[0].x + 0.001, pts[0].y)
        ha, hb = R.s2ComputeHull(pts, n), P.s2ComputeHull(pts, n)
        assert ha.count == hb.count
        for i in range(ha.count):
            assert ha.points[i].x == hb.points[i].x and ha.points[i].y == hb.points[i].y
        if ha.count >= 3:
            assert _poly_eq(R.s2MakePolygon(C.byref(ha)), P.s2MakePolygon(C.byref(hb)))


def _manifold_eq(a, b):
    if a.pointCount != b.pointCount:
        return False
    if a.pointCount == 0:
        return True
    if (a.normal.x, a.normal.y) != (b.normal.x, b.normal.y):
        return False
    for i in range(a.pointCount):
        pa, pb = a.points[i], b.points[i]
        if (pa.localAnchorA.x, pa.localAnchorA.y, pa.localAnchorB.x, pa.localAnchorB.y, pa.separation, pa.id) != \
                (pb.localAnchorA.x, pb.localAnchorA.y, pb.localAnchorB.x, pb.localAnchorB.y, pb.separation, pb.id):
            return False
    return True


def test_manifold_functions_match_reference_bitwise(reference, product):
    """All nine shape-pair manifold functions (reference src/contact.c:139-154) on random near-contact configurations,
    with the GJK cache carried over several perturbed calls as a persistent contact does."""
    R, P = reference, product
    rng = np.random.default_rng(2024)

    def xf(x, y, ang):
        return capi.Transform(capi.Vec2(x, y), capi.Rot(np.float32(np.sin(ang)), np.float32(np.cos(ang))))

    hits = 0
    for trial in range(400):
        boxA = R.s2MakeBox(*rng.uniform(0.3, 1.2, 2))
        pts = (capi.Vec2 * 6)(*[capi.Vec2(*rng.uniform(-0.8, 0.8, 2)) for _ in range(6)])
        hull = R.s2ComputeHull(pts, 6)
        polyB = R.s2MakePolygon(C.byref(hull)) if hull.count >= 3 else R.s2MakeBox(0.5, 0.4)
        polyB.radius = 0.05 if trial % 3 == 0 else 0.0
        circ = capi.Circle(capi.Vec2(*rng.uniform(-0.2, 0.2, 2)), float(rng.uniform(0.2, 0.6)))
        circ2 = capi.Circle(capi.Vec2(0.0, 0.0), float(rng.uniform(0.2, 0.6)))
        cap = capi.Capsule(capi.Vec2(-0.5, 0.0), capi.Vec2(0.5, float(rng.uniform(-0.2, 0.2))), float(rng.uniform(0.1, 0.4)))
        cap2 = capi.Capsule(capi.Vec2(0.0, -0.4), capi.Vec2(0.1, 0.5), 0.25)
        seg = capi.Segment(capi.Vec2(-1.0, 0.0), capi.Vec2(1.0, float(rng.uniform(-0.3, 0.3))))
        d = rng.uniform(0.6, 1.9)
        th = rng.uniform(0, 2 * np.pi)
        cacheR = [capi.DistanceCache() for _ in range(5)]
        cacheP = [capi.DistanceCache() for _ in range(5)]
        for it in range(3):  # persistent cache across slightly moved poses
            A = xf(*rng.uniform(-0.01, 0.01, 2), rng.uniform(-0.02, 0.02) + 0.3 * trial)
            B = xf(d * np.cos(th) + rng.uniform(-0.01, 0.01), d * np.sin(th) + rng.uniform(-0.01, 0.01), rng.uniform(-3, 3) if it == 0 else 0.1 * it)
            pairs = [
                (R.s2CollideCircles(C.byref(circ), A, C.byref(circ2), B), P.s2CollideCircles(C.byref(circ), A, C.byref(circ2), B)),
                (R.s2CollideCapsuleAndCircle(C.byref(cap), A, C.byref(circ), B), P.s2CollideCapsuleAndCircle(C.byref(cap), A, C.byref(circ), B)),
                (R.s2CollideSegmentAndCircle(C.byref(seg), A, C.byref(circ), B), P.s2CollideSegmentAndCircle(C.byref(seg), A, C.byref(circ), B)),
                (R.s2CollidePolygonAndCircle(C.byref(boxA), A, C.byref(circ), B), P.s2CollidePolygonAndCircle(C.byref(boxA), A, C.byref(circ), B)),
                (R.s2CollidePolygons(C.byref(boxA), A, C.byref(polyB), B, C.byref(cacheR[0])), P.s2CollidePolygons(C.byref(boxA), A, C.byref(polyB), B, C.byref(cacheP[0]))),
                (R.s2CollideCapsules(C.byref(cap), A, C.byref(cap2), B, C.byref(cacheR[1])), P.s2CollideCapsules(C.byref(cap), A, C.byref(cap2), B, C.byref(cacheP[1]))),
                (R.s2CollidePolygonAndCapsule(C.byref(boxA), A, C.byref(cap), B, C.byref(cacheR[2])), P.s2CollidePolygonAndCapsule(C.byref(boxA), A, C.byref(cap), B, C.byref(cacheP[2]))),
                (R.s2CollideSegmentAndCapsule(C.byref(seg), A, C.byref(cap), B, C.byref(cacheR[3])), P.s2CollideSegmentAndCapsule(C.byref(seg), A, C.byref(cap), B, C.byref(cacheP[3]))),
                (R.s2CollideSegmentAndPolygon(C.byref(seg), A, C.byref(polyB), B, C.byref(cacheR[4])), P.s2CollideSegmentAndPolygon(C.byref(seg), A, C.byref(polyB), B, C.byref(cacheP[4]))),
            ]
            for k, (mr, mp) in enumerate(pairs):
                assert _manifold_eq(mr, mp), f"trial {trial} iter {it} function {k}"
                hits += mr.pointCount > 0
            for cr, cp in zip(cacheR, cacheP):
                assert cr.count == cp.count and bytes(cr.indexA) == bytes(cp.indexA) and bytes(cr.indexB) == bytes(cp.indexB)
    assert hits > 1500  # the sweep actually produced contacts


def _atan2_inputs(n, seed):
    rng = np.random.default_rng(seed)
    # (sin, cos)-like pairs as s2RelativeAngle produces them, raw bit patterns, and the special values
    a = rng.uniform(-np.pi, np.pi, n)
    y = [np.sin(a).astype(np.float32), rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32),
         np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, 1e-40, -1e-40, 3e38, 0.4375, 0.6875, 1.1875, 2.4375] * 13, np.float32)]
    x = [np.cos(a).astype(np.float32), rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32),
         np.repeat(np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, 1e-40, -1e-40, 3e38, 0.4375, 0.6875, 1.1875, 2.4375],
                            np.float32), 13)]
    return np.concatenate(y), np.concatenate(x)


def test_device_atan2_restatement_equals_libm_on_host():
    """include/solver2d/atan2_f32.h (what the kernels call) against the C library atan2f the reference is linked with."""
    import ctypes.util
    lib = C.CDLL(device.LIB_PATH)
    lib.s2Atan2Device.restype = C.c_float
    lib.s2Atan2Device.argtypes = [C.c_float, C.c_float]
    libm = C.CDLL(ctypes.util.find_library("m"))
    libm.atan2f.restype = C.c_float
    libm.atan2f.argtypes = [C.c_float, C.c_float]
    y, x = _atan2_inputs(150000, 11)
    mine = np.array([lib.s2Atan2Device(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    want = np.array([libm.atan2f(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(mine), nan)
    assert np.array_equal(mine[~nan].view(np.uint32), want[~nan].view(np.uint32))
