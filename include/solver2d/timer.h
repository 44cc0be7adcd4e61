// solver2d-b200 — per-world counters (ABI of reference include/solver2d/timer.h:8-17).
#pragma once

#include "solver2d/types.h"

typedef struct s2Statistics
{
	int32_t bodyCount;
	int32_t contactCount;
	int32_t jointCount;
	int32_t proxyCount;
	// depth of the device BVH built by the last broad-phase pass (the reference reports its dynamic-tree height)
	int32_t treeHeight;
	// bytes of per-step device scratch reserved / used (the reference reports its per-step stack arena)
	int32_t stackCapacity;
	int32_t stackUsed;
} s2Statistics;
