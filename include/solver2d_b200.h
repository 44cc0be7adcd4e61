// solver2d-b200 — extensions to the solver2d C API that only make sense for the GPU library.
// A client that sticks to include/solver2d/solver2d.h never needs this header.
#pragma once

#include "s2b_device.h"
#include "solver2d/id.h"

#ifdef __cplusplus
extern "C"
{
#endif

// The device world behind a public world id, for the stage-level entry points of s2b_device.h (schedule selection,
// counters, timed steps, NCCL packing). The handle is owned by the world.
S2B_API s2bWorld* s2World_GetDevice(s2WorldId worldId);

// Push every pending host-side edit to the device now (s2World_Step does this itself).
S2B_API void s2World_Flush(s2WorldId worldId);

// Bulk read-back for clients that want all transforms each frame (the per-body getters are one call each):
// writes {origin.x, origin.y, rot.s, rot.c} per body slot into out[4 * capacity]; returns the body capacity.
S2B_API int32_t s2World_GetBodyTransforms(s2WorldId worldId, float* out, int32_t capacity);

#ifdef __cplusplus
}
#endif
