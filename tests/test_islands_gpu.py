"""GPU: the island builder (the reference reserves an island pool it never fills, src/world.h:31).

* islands = connected components of the constraint graph over the movable bodies: a field of piles built as ONE world in
  interleaved creation order yields one island per pile;
* every island evolves bit for bit like a world that holds it alone (the colouring, the region schedule and the warm-start
  gather are island-local: nothing leaks between islands), through the public API, free running;
* with islands kept whole inside one region the field needs NO device-wide constraint step at all (cut set empty)."""
import numpy as np
import pytest

from solver2d_b200 import capi, device, scenes

pytestmark = pytest.mark.gpu
DT = 1.0 / 60.0


@pytest.fixture(scope="module")
def product():
    return capi.Solver2D(device.LIB_PATH)


def _state(lib, bodies):
    pos = np.array([tuple(lib.s2Body_GetPosition(b)) for b in bodies], dtype=np.float32)
    ang = np.array([lib.s2Body_GetAngle(b) for b in bodies], dtype=np.float32)
    return pos, ang


@pytest.mark.parametrize("solver", ["TGS_Soft", "PGS_NGS_Block", "XPBD"])
def test_interleaved_field_islands_match_standalone_worlds(product, dev, solver):
    P = product
    count, base, steps = 12, 8, 40
    field = scenes.pyramid_field_interleaved(P, solver, count=count, base_count=base)
    for _ in range(steps):
        field.step(DT, 4, 2, True)
    dw = device.DeviceWorld.attach(dev, field.world)
    labels, n_islands = dw.islands()
    c = dw.counters()

    # one island per pile (its boxes), one singleton per static ground
    boxes_per_pile = base * (base + 1) // 2
    box_labels = []
    for k in range(count):
        ids = [b.index for b in field.bodies[field.piles[k]]]
        ground, boxes = ids[0], ids[1:]
        assert labels[ground] == ground, "a static body is an island of its own"
        lab = set(int(labels[i]) for i in boxes)
        assert len(lab) == 1, f"pile {k} was split into islands {lab}"
        assert lab.pop() == min(boxes), "the label of an island is its smallest body slot"
        box_labels.append(int(labels[boxes[0]]))
    assert len(set(box_labels)) == count and n_islands == 2 * count
    assert len(boxes) == boxes_per_pile

    # islands stay whole inside a region: no cut set, no device-wide constraint step
    assert c.regionCount >= 1 and c.cutCount == 0 and c.cutGroupCount == 0

    fpos, fang = _state(P, field.bodies)
    for k in range(count):
        alone = scenes.pyramid_field_interleaved(P, solver, count=count, base_count=base, only=k)
        for _ in range(steps):
            alone.step(DT, 4, 2, True)
        apos, aang = _state(P, alone.bodies)
        sl = field.piles[k]
        assert np.array_equal(fpos[sl].view(np.uint32), apos.view(np.uint32)), f"pile {k}: positions differ from the standalone world"
        assert np.array_equal(fang[sl].view(np.uint32), aang.view(np.uint32)), f"pile {k}: angles differ from the standalone world"
        alone.destroy()
    field.destroy()


def test_islands_of_jointed_scene(product, dev):
    """Bridges hang from the static ground: every bridge is one island (joints connect), boxes in the air are singletons
    until they land."""
    P = product
    sc = scenes.joint_contact_stress(P, "TGS_Soft", bridges=3, planks=24, grid=4)
    sc.step(DT, 4, 2, True)
    dw = device.DeviceWorld.attach(dev, sc.world)
    labels, n = dw.islands()
    ids = [b.index for b in sc.bodies]
    planks = ids[1:1 + 3 * 24]
    per_bridge = [set(int(labels[i]) for i in planks[24 * b:24 * (b + 1)]) for b in range(3)]
    assert all(len(s) == 1 for s in per_bridge) and len(set.union(*per_bridge)) == 3
    falling = ids[1 + 3 * 24:]
    assert all(labels[i] == i for i in falling)
    assert n == 1 + 3 + len(falling)
    sc.destroy()
