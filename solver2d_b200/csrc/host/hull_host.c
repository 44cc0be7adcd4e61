// solver2d-b200 — convex hull for polygon authoring (behaviour of reference src/hull.c): quickhull over at most
// s2_maxPolygonVertices points with point welding and collinear-vertex removal at the s2_linearSlop scale. The output
// vertex order is part of the contract (it fixes polygon feature ids), so the selection rules are the reference's:
// start at the point farthest from the AABB centre, then the point farthest from it, split, refine the right side of each
// edge, finally drop nearly collinear vertices.
//
// Shape of this implementation: no recursion and no intermediate hulls. The hull is grown in place as a list of
// vertices; every edge that may still hide a vertex is a JOB (edge start, edge end, candidate points) on a small explicit
// stack. A job either closes its edge (no candidate far enough out) or inserts the apex between the two ends and leaves
// two jobs behind. Jobs are taken so that vertices come out in the order a depth-first refinement would visit them:
// counter-clockwise from the first extreme point.
#include "s2_host.h"

#include <float.h>
#include <string.h>

#define S2_HULL_MAX s2_maxPolygonVertices

// candidate sets live in one pool: a job's candidates are a slice of it (every refinement only ever shrinks a set, and
// the depth is bounded by the vertex count, so S2_HULL_MAX slices of S2_HULL_MAX points are plenty)
typedef struct s2HullJob
{
	s2Vec2 from, to;
	int32_t first, count; // slice of the candidate pool
} s2HullJob;

typedef struct s2HullBuilder
{
	s2Vec2 pool[(2 * S2_HULL_MAX + 2) * S2_HULL_MAX];
	int32_t poolUsed;
	s2HullJob jobs[2 * S2_HULL_MAX + 4];
	int32_t jobCount;
	s2Vec2 out[2 * S2_HULL_MAX];
	int32_t outCount;
} s2HullBuilder;

static float s2SideOfEdge(s2Vec2 p, s2Vec2 from, s2Vec2 unitEdge)
{
	// > 0: to the right of the directed edge
	return s2Cross(s2Sub(p, from), unitEdge);
}

static void s2PushHullJob(s2HullBuilder* hb, s2Vec2 from, s2Vec2 to, const s2Vec2* candidates, int32_t count)
{
	s2HullJob* job = hb->jobs + hb->jobCount++;
	job->from = from;
	job->to = to;
	job->first = hb->poolUsed;
	job->count = count;
	if (count > 0)
	{
		memcpy(hb->pool + hb->poolUsed, candidates, sizeof(s2Vec2) * (size_t)count);
		hb->poolUsed += count;
	}
}

// Emits the vertices strictly between `from` and `to` (both excluded) that lie on the hull, in order along the edge.
static void s2RefineHullEdge(s2HullBuilder* hb, s2Vec2 from, s2Vec2 to, const s2Vec2* candidates, int32_t count)
{
	int32_t base = hb->jobCount;
	int32_t poolBase = hb->poolUsed;
	s2PushHullJob(hb, from, to, candidates, count);
	while (hb->jobCount > base)
	{
		s2HullJob job = hb->jobs[--hb->jobCount];
		if (job.count == 0)
		{
			// nothing outside this edge: its END vertex is next on the hull, unless it is the end of the whole refinement
			if (hb->jobCount > base)
			{
				hb->out[hb->outCount++] = job.to;
			}
			continue;
		}

		// the farthest candidate to the right of the edge is the apex; candidates on the right survive into both halves
		const s2Vec2* ps = hb->pool + job.first;
		s2Vec2 unitEdge = s2Normalize(s2Sub(job.to, job.from));
		s2Vec2 kept[S2_HULL_MAX];
		int32_t keptCount = 0;
		int32_t apexIndex = 0;
		float apexDistance = -FLT_MAX;
		for (int32_t i = 0; i < job.count; ++i)
		{
			float d = s2SideOfEdge(ps[i], job.from, unitEdge);
			if (i == 0 || d > apexDistance)
			{
				apexIndex = i;
				apexDistance = d;
			}
			if (d > 0.0f)
			{
				kept[keptCount++] = ps[i];
			}
		}
		if (apexDistance < 2.0f * s2_linearSlop)
		{
			if (hb->jobCount > base)
			{
				hb->out[hb->outCount++] = job.to;
			}
			continue;
		}
		s2Vec2 apex = ps[apexIndex];
		// the half that ends at job.to goes on the stack first, so the half that starts at job.from is refined first
		s2PushHullJob(hb, apex, job.to, kept, keptCount);
		s2PushHullJob(hb, job.from, apex, kept, keptCount);
	}
	hb->poolUsed = poolBase;
}

static int32_t s2FarthestPoint(s2Vec2 origin, const s2Vec2* ps, int32_t n)
{
	int32_t winner = 0;
	float winnerSq = s2DistanceSquared(origin, ps[0]);
	for (int32_t i = 1; i < n; ++i)
	{
		float dsq = s2DistanceSquared(origin, ps[i]);
		if (dsq > winnerSq)
		{
			winner = i;
			winnerSq = dsq;
		}
	}
	return winner;
}

// take point `index` out of the working set (the last point moves into its place)
static s2Vec2 s2TakePoint(s2Vec2* ps, int32_t* n, int32_t index)
{
	s2Vec2 p = ps[index];
	*n -= 1;
	ps[index] = ps[*n];
	return p;
}

s2Hull s2ComputeHull(const s2Vec2* points, int32_t count)
{
	s2Hull hull;
	hull.count = 0;
	if (count < 3 || count > S2_HULL_MAX)
	{
		return hull;
	}

	// working set: the input without points closer than 4 * linearSlop to an EARLIER input point; bounds over all of it
	const float weldSq = 16.0f * s2_linearSlop * s2_linearSlop;
	s2Vec2 lo = {FLT_MAX, FLT_MAX}, hi = {-FLT_MAX, -FLT_MAX};
	s2Vec2 work[S2_HULL_MAX];
	int32_t n = 0;
	for (int32_t i = 0; i < count; ++i)
	{
		lo = s2Min(lo, points[i]);
		hi = s2Max(hi, points[i]);
		int32_t j = 0;
		while (j < i && s2DistanceSquared(points[i], points[j]) >= weldSq)
		{
			j += 1;
		}
		if (j == i)
		{
			work[n++] = points[i];
		}
	}
	if (n < 3)
	{
		return hull;
	}

	// the two extreme points the hull starts from
	s2Box bounds = {lo, hi};
	s2Vec2 first = s2TakePoint(work, &n, s2FarthestPoint(s2AABB_Center(bounds), work, n));
	s2Vec2 second = s2TakePoint(work, &n, s2FarthestPoint(first, work, n));

	// what lies clearly right / clearly left of the line through them
	s2Vec2 rightSide[S2_HULL_MAX], leftSide[S2_HULL_MAX];
	int32_t rightCount = 0, leftCount = 0;
	s2Vec2 axis = s2Normalize(s2Sub(second, first));
	for (int32_t i = 0; i < n; ++i)
	{
		float d = s2SideOfEdge(work[i], first, axis);
		if (d >= 2.0f * s2_linearSlop)
		{
			rightSide[rightCount++] = work[i];
		}
		else if (d <= -2.0f * s2_linearSlop)
		{
			leftSide[leftCount++] = work[i];
		}
	}

	s2HullBuilder hb;
	hb.poolUsed = 0;
	hb.jobCount = 0;
	hb.outCount = 0;
	hb.out[hb.outCount++] = first;
	s2RefineHullEdge(&hb, first, second, rightSide, rightCount);
	int32_t afterFirstChain = hb.outCount;
	hb.out[hb.outCount++] = second;
	s2RefineHullEdge(&hb, second, first, leftSide, leftCount);
	if (afterFirstChain == 1 && hb.outCount == 2)
	{
		return hull; // collinear input: neither side has a vertex
	}

	// drop a vertex whose distance to the chord of its neighbours is within 2 * linearSlop; start over after each removal
	int32_t m = hb.outCount;
	s2Vec2* v = hb.out;
	for (int32_t i = 0; m > 2 && i < m;)
	{
		int32_t mid = (i + 1) % m;
		s2Vec2 a = v[i], b = v[mid], c = v[(i + 2) % m];
		s2Vec2 chord = s2Normalize(s2Sub(c, a));
		if (s2Cross(s2Sub(b, a), chord) <= 2.0f * s2_linearSlop)
		{
			memmove(v + mid, v + mid + 1, sizeof(s2Vec2) * (size_t)(m - 1 - mid));
			m -= 1;
			i = 0;
		}
		else
		{
			i += 1;
		}
	}
	if (m < 3)
	{
		return hull;
	}
	hull.count = m;
	memcpy(hull.points, v, sizeof(s2Vec2) * (size_t)m);
	return hull;
}

bool s2ValidateHull(const s2Hull* hull)
{
	int32_t n = hull->count;
	if (n < 3 || n > S2_HULL_MAX)
	{
		return false;
	}
	for (int32_t i = 0; i < n; ++i)
	{
		int32_t next = (i + 1 == n) ? 0 : i + 1;
		s2Vec2 p = hull->points[i];

		// convex and counter-clockwise: every other vertex is strictly left of the edge that starts here
		s2Vec2 edge = s2Normalize(s2Sub(hull->points[next], p));
		for (int32_t j = 0; j < n; ++j)
		{
			if (j != i && j != next && s2SideOfEdge(hull->points[j], p, edge) >= 0.0f)
			{
				return false;
			}
		}
	}
	for (int32_t i = 0; i < n; ++i)
	{
		// no vertex within linearSlop of the chord of its neighbours
		int32_t next = (i + 1) % n;
		int32_t afterNext = (i + 2) % n;
		s2Vec2 p = hull->points[i];
		s2Vec2 chord = s2Normalize(s2Sub(hull->points[afterNext], p));
		if (s2SideOfEdge(hull->points[next], p, chord) <= s2_linearSlop)
		{
			return false;
		}
	}
	return true;
}
