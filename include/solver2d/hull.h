// solver2d-b200 — convex hull of a small point cloud (API of reference include/solver2d/hull.h).
#pragma once

#include "solver2d/constants.h"
#include "solver2d/types.h"

typedef struct s2Hull
{
	s2Vec2 points[s2_maxPolygonVertices];
	int32_t count;
} s2Hull;

#ifdef __cplusplus
extern "C"
{
#endif

// Quickhull with collinear-point removal; returns count == 0 on failure (fewer than 3 usable points).
s2Hull s2ComputeHull(const s2Vec2* points, int32_t count);

// Debug check that a hull is convex, counter-clockwise and free of collinear points.
bool s2ValidateHull(const s2Hull* hull);

#ifdef __cplusplus
}
#endif
