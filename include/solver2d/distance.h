// solver2d-b200 — closest-point queries (API of reference include/solver2d/distance.h).
#pragma once

#include "solver2d/constants.h"
#include "solver2d/types.h"

#ifdef __cplusplus
extern "C"
{
#endif

typedef struct s2SegmentDistanceResult
{
	s2Vec2 closest1;
	s2Vec2 closest2;
	float fraction1;
	float fraction2;
	float distanceSquared;
} s2SegmentDistanceResult;

s2SegmentDistanceResult s2SegmentDistance(s2Vec2 p1, s2Vec2 q1, s2Vec2 p2, s2Vec2 q2);

// convex point cloud + radius: the shape abstraction GJK works on
typedef struct s2DistanceProxy
{
	s2Vec2 vertices[s2_maxPolygonVertices];
	int32_t count;
	float radius;
} s2DistanceProxy;

// GJK simplex carried from one step to the next (zero `count` on first use)
typedef struct s2DistanceCache
{
	float metric;
	uint16_t count;
	uint8_t indexA[3];
	uint8_t indexB[3];
} s2DistanceCache;

static const s2DistanceCache s2_emptyDistanceCache = S2_ZERO_INIT;

typedef struct s2DistanceInput
{
	s2DistanceProxy proxyA;
	s2DistanceProxy proxyB;
	s2Transform transformA;
	s2Transform transformB;
	bool useRadii;
} s2DistanceInput;

typedef struct s2DistanceOutput
{
	s2Vec2 pointA;
	s2Vec2 pointB;
	float distance;
	int32_t iterations;
} s2DistanceOutput;

s2DistanceOutput s2ShapeDistance(s2DistanceCache* cache, const s2DistanceInput* input);
s2DistanceProxy s2MakeProxy(const s2Vec2* vertices, int32_t count, float radius);

#ifdef __cplusplus
}
#endif
