// solver2d-b200 — convex hull for polygon authoring (behaviour of reference src/hull.c): quickhull over at most
// s2_maxPolygonVertices points with point welding and collinear-vertex removal at the s2_linearSlop scale. The output
// vertex order is part of the contract (it fixes polygon feature ids), so the selection rules are the reference's:
// start at the point farthest from the AABB centre, then the point farthest from it, split, recurse on the right side
// of each edge, finally drop nearly collinear vertices.
#include "s2_host.h"

#include <float.h>

// points of `ps` strictly right of edge p1->p2, recursively: hull chain from p1 to p2 (exclusive of both)
static s2Hull s2HullChain(s2Vec2 p1, s2Vec2 p2, const s2Vec2* ps, int32_t count)
{
	s2Hull chain;
	chain.count = 0;
	if (count == 0)
	{
		return chain;
	}

	s2Vec2 e = s2Normalize(s2Sub(p2, p1));
	s2Vec2 right[s2_maxPolygonVertices];
	int32_t rightCount = 0;
	int32_t best = 0;
	float bestDistance = s2Cross(s2Sub(ps[0], p1), e);
	if (bestDistance > 0.0f)
	{
		right[rightCount++] = ps[0];
	}
	for (int32_t i = 1; i < count; ++i)
	{
		float distance = s2Cross(s2Sub(ps[i], p1), e);
		if (distance > bestDistance)
		{
			best = i;
			bestDistance = distance;
		}
		if (distance > 0.0f)
		{
			right[rightCount++] = ps[i];
		}
	}
	if (bestDistance < 2.0f * s2_linearSlop)
	{
		return chain;
	}

	s2Vec2 apex = ps[best];
	s2Hull before = s2HullChain(p1, apex, right, rightCount);
	s2Hull after = s2HullChain(apex, p2, right, rightCount);
	for (int32_t i = 0; i < before.count; ++i)
	{
		chain.points[chain.count++] = before.points[i];
	}
	chain.points[chain.count++] = apex;
	for (int32_t i = 0; i < after.count; ++i)
	{
		chain.points[chain.count++] = after.points[i];
	}
	return chain;
}

static int32_t s2FarthestFrom(s2Vec2 origin, const s2Vec2* ps, int32_t n)
{
	int32_t best = 0;
	float bestSq = s2DistanceSquared(origin, ps[0]);
	for (int32_t i = 1; i < n; ++i)
	{
		float dsq = s2DistanceSquared(origin, ps[i]);
		if (dsq > bestSq)
		{
			best = i;
			bestSq = dsq;
		}
	}
	return best;
}

s2Hull s2ComputeHull(const s2Vec2* points, int32_t count)
{
	s2Hull hull;
	hull.count = 0;
	if (count < 3 || count > s2_maxPolygonVertices)
	{
		return hull;
	}

	// weld points closer than 4 * linearSlop (the first of a cluster survives) and bound the input
	s2Box aabb = {{FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX}};
	s2Vec2 ps[s2_maxPolygonVertices];
	int32_t n = 0;
	const float tolSqr = 16.0f * s2_linearSlop * s2_linearSlop;
	for (int32_t i = 0; i < count; ++i)
	{
		aabb.lowerBound = s2Min(aabb.lowerBound, points[i]);
		aabb.upperBound = s2Max(aabb.upperBound, points[i]);
		bool unique = true;
		for (int32_t j = 0; j < i; ++j)
		{
			if (s2DistanceSquared(points[i], points[j]) < tolSqr)
			{
				unique = false;
				break;
			}
		}
		if (unique)
		{
			ps[n++] = points[i];
		}
	}
	if (n < 3)
	{
		return hull;
	}

	// two extreme points, removed from the working set by swapping in the last element
	int32_t f1 = s2FarthestFrom(s2AABB_Center(aabb), ps, n);
	s2Vec2 p1 = ps[f1];
	ps[f1] = ps[--n];
	int32_t f2 = s2FarthestFrom(p1, ps, n);
	s2Vec2 p2 = ps[f2];
	ps[f2] = ps[--n];

	s2Vec2 rightPoints[s2_maxPolygonVertices - 2], leftPoints[s2_maxPolygonVertices - 2];
	int32_t rightCount = 0, leftCount = 0;
	s2Vec2 e = s2Normalize(s2Sub(p2, p1));
	for (int32_t i = 0; i < n; ++i)
	{
		float d = s2Cross(s2Sub(ps[i], p1), e);
		if (d >= 2.0f * s2_linearSlop)
		{
			rightPoints[rightCount++] = ps[i];
		}
		else if (d <= -2.0f * s2_linearSlop)
		{
			leftPoints[leftCount++] = ps[i];
		}
	}

	s2Hull chain1 = s2HullChain(p1, p2, rightPoints, rightCount);
	s2Hull chain2 = s2HullChain(p2, p1, leftPoints, leftCount);
	if (chain1.count == 0 && chain2.count == 0)
	{
		return hull; // collinear input
	}

	hull.points[hull.count++] = p1;
	for (int32_t i = 0; i < chain1.count; ++i)
	{
		hull.points[hull.count++] = chain1.points[i];
	}
	hull.points[hull.count++] = p2;
	for (int32_t i = 0; i < chain2.count; ++i)
	{
		hull.points[hull.count++] = chain2.points[i];
	}

	// drop a vertex whose distance to the chord of its neighbours is within 2 * linearSlop; restart after each removal
	bool searching = true;
	while (searching && hull.count > 2)
	{
		searching = false;
		for (int32_t i = 0; i < hull.count; ++i)
		{
			int32_t i2 = (i + 1) % hull.count;
			int32_t i3 = (i + 2) % hull.count;
			s2Vec2 a = hull.points[i], b = hull.points[i2], c = hull.points[i3];
			s2Vec2 chord = s2Normalize(s2Sub(c, a));
			if (s2Cross(s2Sub(b, a), chord) <= 2.0f * s2_linearSlop)
			{
				for (int32_t j = i2; j < hull.count - 1; ++j)
				{
					hull.points[j] = hull.points[j + 1];
				}
				hull.count -= 1;
				searching = true;
				break;
			}
		}
	}
	if (hull.count < 3)
	{
		hull.count = 0;
	}
	return hull;
}

bool s2ValidateHull(const s2Hull* hull)
{
	if (hull->count < 3 || s2_maxPolygonVertices < hull->count)
	{
		return false;
	}
	// convex and counter-clockwise: every other vertex is strictly left of every edge
	for (int32_t i = 0; i < hull->count; ++i)
	{
		int32_t next = i < hull->count - 1 ? i + 1 : 0;
		s2Vec2 p = hull->points[i];
		s2Vec2 e = s2Normalize(s2Sub(hull->points[next], p));
		for (int32_t j = 0; j < hull->count; ++j)
		{
			if (j == i || j == next)
			{
				continue;
			}
			if (s2Cross(s2Sub(hull->points[j], p), e) >= 0.0f)
			{
				return false;
			}
		}
	}
	// no vertex within linearSlop of the chord of its neighbours
	for (int32_t i = 0; i < hull->count; ++i)
	{
		s2Vec2 a = hull->points[i];
		s2Vec2 b = hull->points[(i + 1) % hull->count];
		s2Vec2 c = hull->points[(i + 2) % hull->count];
		s2Vec2 e = s2Normalize(s2Sub(c, a));
		if (s2Cross(s2Sub(b, a), e) <= s2_linearSlop)
		{
			return false;
		}
	}
	return true;
}
