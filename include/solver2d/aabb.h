// solver2d-b200 — axis-aligned box helpers (API of reference include/solver2d/aabb.h).
#pragma once

#include "solver2d/constants.h"
#include "solver2d/math.h"
#include "solver2d/types.h"

#ifdef __cplusplus
extern "C"
{
#endif

bool s2AABB_IsValid(s2Box a);
s2RayCastOutput s2AABB_RayCast(s2Box a, s2Vec2 p1, s2Vec2 p2);

#ifdef __cplusplus
}
#endif

S2_INLINE s2Vec2 s2AABB_Center(s2Box a)
{
	s2Vec2 c = {0.5f * (a.lowerBound.x + a.upperBound.x), 0.5f * (a.lowerBound.y + a.upperBound.y)};
	return c;
}

S2_INLINE s2Vec2 s2AABB_Extents(s2Box a)
{
	s2Vec2 e = {0.5f * (a.upperBound.x - a.lowerBound.x), 0.5f * (a.upperBound.y - a.lowerBound.y)};
	return e;
}

S2_INLINE float s2AABB_Perimeter(s2Box a)
{
	float wx = a.upperBound.x - a.lowerBound.x;
	float wy = a.upperBound.y - a.lowerBound.y;
	return 2.0f * (wx + wy);
}

S2_INLINE s2Box s2AABB_Union(s2Box a, s2Box b)
{
	s2Box c;
	c.lowerBound.x = S2_MIN(a.lowerBound.x, b.lowerBound.x);
	c.lowerBound.y = S2_MIN(a.lowerBound.y, b.lowerBound.y);
	c.upperBound.x = S2_MAX(a.upperBound.x, b.upperBound.x);
	c.upperBound.y = S2_MAX(a.upperBound.y, b.upperBound.y);
	return c;
}

// grow *a to cover b; true if *a changed
S2_INLINE bool s2AABB_Enlarge(s2Box* a, s2Box b)
{
	bool changed = false;
	if (b.lowerBound.x < a->lowerBound.x)
	{
		a->lowerBound.x = b.lowerBound.x;
		changed = true;
	}
	if (b.lowerBound.y < a->lowerBound.y)
	{
		a->lowerBound.y = b.lowerBound.y;
		changed = true;
	}
	if (a->upperBound.x < b.upperBound.x)
	{
		a->upperBound.x = b.upperBound.x;
		changed = true;
	}
	if (a->upperBound.y < b.upperBound.y)
	{
		a->upperBound.y = b.upperBound.y;
		changed = true;
	}
	return changed;
}

// a fully contains b (closed comparison, reference aabb.h:93-101)
S2_INLINE bool s2AABB_Contains(s2Box a, s2Box b)
{
	return a.lowerBound.x <= b.lowerBound.x && a.lowerBound.y <= b.lowerBound.y && b.upperBound.x <= a.upperBound.x &&
		   b.upperBound.y <= a.upperBound.y;
}

S2_INLINE bool s2AABB_ContainsWithMargin(s2Box a, s2Box b, float margin)
{
	return (a.lowerBound.x <= b.lowerBound.x - margin) & (a.lowerBound.y <= b.lowerBound.y - margin) &
		   (b.upperBound.x + margin <= a.upperBound.x) & (b.upperBound.y + margin <= a.upperBound.y);
}

// closed overlap test: touching boxes overlap (reference aabb.h:111-123)
S2_INLINE bool s2AABB_Overlaps(s2Box a, s2Box b)
{
	float d1x = b.lowerBound.x - a.upperBound.x, d1y = b.lowerBound.y - a.upperBound.y;
	float d2x = a.lowerBound.x - b.upperBound.x, d2y = a.lowerBound.y - b.upperBound.y;
	if (d1x > 0.0f || d1y > 0.0f)
	{
		return false;
	}
	if (d2x > 0.0f || d2y > 0.0f)
	{
		return false;
	}
	return true;
}
