// solver2d-b200 — opaque handles handed to clients (ABI of reference include/solver2d/id.h:12-47).
// A handle is {slot index, owning world, revision}; the revision is bumped each time a pool slot is reused so a
// stale handle can be detected. Handles are plain values: pass and store them by value.
#pragma once

#include <stdint.h>

typedef struct s2WorldId
{
	int16_t index;
	uint16_t revision;
} s2WorldId;

#define S2_DECLARE_OBJECT_ID(NAME)                                                                                     \
	typedef struct NAME                                                                                                \
	{                                                                                                                  \
		int32_t index;                                                                                                 \
		int16_t world;                                                                                                 \
		uint16_t revision;                                                                                             \
	} NAME

S2_DECLARE_OBJECT_ID(s2BodyId);
S2_DECLARE_OBJECT_ID(s2ShapeId);
S2_DECLARE_OBJECT_ID(s2JointId);

static const s2WorldId s2_nullWorldId = {-1, 0};
static const s2BodyId s2_nullBodyId = {-1, -1, 0};
static const s2ShapeId s2_nullShapeId = {-1, -1, 0};
static const s2JointId s2_nullJointId = {-1, -1, 0};

#define S2_IS_NULL(ID) ((ID).index == -1)
#define S2_NON_NULL(ID) ((ID).index != -1)
