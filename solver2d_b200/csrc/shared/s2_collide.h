// solver2d-b200 — narrow-phase geometry: GJK closest features, SAT, edge clipping and the per-shape-pair manifold
// functions, written once as plain-C inline functions that compile for the host (public s2Collide* / s2ShapeDistance
// API, csrc/host/collide.c) and for the device (one thread per contact, csrc/device/narrowphase.cu).
//
// Behaviour contract = the reference narrow phase (src/manifold.c, src/distance.c): same feature ids, same
// contact points, and — because both sides are compiled without FMA contraction — the same float results
// bit for bit on the same inputs. Citations on each function.
#pragma once

#include "solver2d/distance.h"
#include "solver2d/geometry.h"
#include "solver2d/manifold.h"
#include "solver2d/math.h"

#include <float.h>

#if defined(__CUDACC__)
	#define S2C_FN static inline __host__ __device__
	#define S2C_UNROLL _Pragma("unroll")
#else
	#define S2C_FN static inline
	#define S2C_UNROLL
#endif

// ---- small helpers ----------------------------------------------------------------------------------------------

S2C_FN s2Vec2 s2cVec(float x, float y)
{
	s2Vec2 v;
	v.x = x;
	v.y = y;
	return v;
}

// s2Normalize (reference src/math.c:40-51)
S2C_FN s2Vec2 s2cNormalize(s2Vec2 v)
{
	float length = s2Length(v);
	if (length < 0.001f * FLT_EPSILON)
	{
		return s2cVec(0.0f, 0.0f);
	}
	float invLength = 1.0f / length;
	return s2cVec(invLength * v.x, invLength * v.y);
}

// s2NormalizeChecked (reference src/math.c:53-65)
S2C_FN s2Vec2 s2cNormalizeChecked(s2Vec2 v)
{
	float length = s2Length(v);
	if (length < FLT_EPSILON)
	{
		return s2cVec(0.0f, 0.0f);
	}
	float invLength = 1.0f / length;
	return s2cVec(invLength * v.x, invLength * v.y);
}

// s2GetLengthAndNormalize (reference src/math.c:67-79)
S2C_FN s2Vec2 s2cGetLengthAndNormalize(float* length, s2Vec2 v)
{
	*length = s2Length(v);
	if (*length < FLT_EPSILON)
	{
		return s2cVec(0.0f, 0.0f);
	}
	float invLength = 1.0f / *length;
	return s2cVec(invLength * v.x, invLength * v.y);
}

S2C_FN void s2cClearManifold(s2Manifold* m)
{
	// every byte zero, like the reference's `s2Manifold manifold = {0}`
	unsigned char* p = (unsigned char*)m;
	for (unsigned i = 0; i < sizeof(s2Manifold); ++i)
	{
		p[i] = 0;
	}
}

// s2MakeCapsule (reference src/geometry.c:96-111): a 2-gon with the two side normals
S2C_FN void s2cMakeCapsule(s2Polygon* shape, s2Vec2 p1, s2Vec2 p2, float radius)
{
	for (int i = 0; i < s2_maxPolygonVertices; ++i)
	{
		shape->vertices[i] = s2cVec(0.0f, 0.0f);
		shape->normals[i] = s2cVec(0.0f, 0.0f);
	}
	shape->vertices[0] = p1;
	shape->vertices[1] = p2;
	s2Vec2 axis = s2cNormalizeChecked(s2Sub(p2, p1));
	s2Vec2 normal = s2RightPerp(axis);
	shape->normals[0] = normal;
	shape->normals[1] = s2Neg(normal);
	shape->count = 2;
	shape->radius = radius;
}

// ---- GJK (reference src/distance.c) -----------------------------------------------------------------------------

typedef struct s2cSimplexVertex
{
	s2Vec2 wA, wB, w; // support points and their difference wB - wA
	float a;		  // barycentric weight
	int indexA, indexB;
} s2cSimplexVertex;

typedef struct s2cSimplex
{
	s2cSimplexVertex v[3];
	int count;
} s2cSimplex;

#define S2C_SET_SIMPLEX_VERTEX(dst, src) \
	do \
	{ \
		(dst).indexA = (src).indexA; \
		(dst).indexB = (src).indexB; \
		(dst).wA = (src).wA; \
		(dst).wB = (src).wB; \
		(dst).w = (src).w; \
	} while (0)

// s2FindSupport (reference src/distance.c:117-132)
S2C_FN int s2cFindSupport(const s2Vec2* vertices, int count, s2Vec2 direction)
{
	int bestIndex = 0;
	float bestValue = s2Dot(vertices[0], direction);
	for (int i = 1; i < count; ++i)
	{
		float value = s2Dot(vertices[i], direction);
		if (value > bestValue)
		{
			bestIndex = i;
			bestValue = value;
		}
	}
	return bestIndex;
}

S2C_FN s2Vec2 s2cWeight2(float a1, s2Vec2 w1, float a2, s2Vec2 w2)
{
	return s2cVec(a1 * w1.x + a2 * w2.x, a1 * w1.y + a2 * w2.y);
}

S2C_FN s2Vec2 s2cWeight3(float a1, s2Vec2 w1, float a2, s2Vec2 w2, float a3, s2Vec2 w3)
{
	return s2cVec(a1 * w1.x + a2 * w2.x + a3 * w3.x, a1 * w1.y + a2 * w2.y + a3 * w3.y);
}

// s2SolveSimplex2 (reference src/distance.c:304-334): closest point of a segment to the origin
S2C_FN void s2cSolveSimplex2(s2cSimplex* s)
{
	s2Vec2 w1 = s->v[0].w;
	s2Vec2 w2 = s->v[1].w;
	s2Vec2 e12 = s2Sub(w2, w1);

	float d12_2 = -s2Dot(w1, e12);
	if (d12_2 <= 0.0f)
	{
		s->v[0].a = 1.0f;
		s->count = 1;
		return;
	}

	float d12_1 = s2Dot(w2, e12);
	if (d12_1 <= 0.0f)
	{
		s->v[1].a = 1.0f;
		s->count = 1;
		s->v[0] = s->v[1];
		return;
	}

	float inv_d12 = 1.0f / (d12_1 + d12_2);
	s->v[0].a = d12_1 * inv_d12;
	s->v[1].a = d12_2 * inv_d12;
	s->count = 2;
}

// Un-normalised barycentric coordinates of the origin's projection on the line through p and q: {weight of p, weight of q}.
// Both positive: the projection lies between them; weight of q <= 0: at or before p; weight of p <= 0: at or past q.
typedef struct s2cLineWeights
{
	float atFirst, atSecond;
} s2cLineWeights;

S2C_FN s2cLineWeights s2cProjectOriginOnLine(s2Vec2 p, s2Vec2 q)
{
	s2Vec2 along = s2Sub(q, p);
	s2cLineWeights w;
	w.atFirst = s2Dot(q, along);
	w.atSecond = -s2Dot(p, along);
	return w;
}

// the reduced simplex is an edge: vertices `first` and `second` of the triangle end up in slots 0 and 1
#define S2C_REDUCE_TO_EDGE(s, first, second, weights)                                                                      \
	do                                                                                                                    \
	{                                                                                                                     \
		float scale__ = 1.0f / ((weights).atFirst + (weights).atSecond);                                                 \
		(s)->v[first].a = (weights).atFirst * scale__;                                                                   \
		(s)->v[second].a = (weights).atSecond * scale__;                                                                 \
		(s)->count = 2;                                                                                                   \
	} while (0)

// Closest point of the triangle (w1, w2, w3) of a GJK simplex to the origin, by Voronoi region (behaviour and region
// order of reference src/distance.c:336-446: a point on a region boundary goes to the first region listed). The triangle's
// vertices are addressed with literal indices only — on the device the simplex lives in registers.
S2C_FN void s2cSolveSimplex3(s2cSimplex* s)
{
	s2Vec2 p1 = s->v[0].w, p2 = s->v[1].w, p3 = s->v[2].w;
	s2cLineWeights side12 = s2cProjectOriginOnLine(p1, p2);
	s2cLineWeights side13 = s2cProjectOriginOnLine(p1, p3);
	s2cLineWeights side23 = s2cProjectOriginOnLine(p2, p3);

	// signed areas of the three sub-triangles the origin spans with the sides, scaled by the triangle's own signed area so
	// that their signs do not depend on the winding
	float winding = s2Cross(s2Sub(p2, p1), s2Sub(p3, p1));
	float area1 = winding * s2Cross(p2, p3);
	float area2 = winding * s2Cross(p3, p1);
	float area3 = winding * s2Cross(p1, p2);

	if (side12.atSecond <= 0.0f && side13.atSecond <= 0.0f)
	{
		// corner p1
		s->v[0].a = 1.0f;
		s->count = 1;
	}
	else if (side12.atFirst > 0.0f && side12.atSecond > 0.0f && area3 <= 0.0f)
	{
		S2C_REDUCE_TO_EDGE(s, 0, 1, side12);
	}
	else if (side13.atFirst > 0.0f && side13.atSecond > 0.0f && area2 <= 0.0f)
	{
		S2C_REDUCE_TO_EDGE(s, 0, 2, side13);
		s->v[1] = s->v[2];
	}
	else if (side12.atFirst <= 0.0f && side23.atSecond <= 0.0f)
	{
		// corner p2
		s->v[1].a = 1.0f;
		s->count = 1;
		s->v[0] = s->v[1];
	}
	else if (side13.atFirst <= 0.0f && side23.atFirst <= 0.0f)
	{
		// corner p3
		s->v[2].a = 1.0f;
		s->count = 1;
		s->v[0] = s->v[2];
	}
	else if (side23.atFirst > 0.0f && side23.atSecond > 0.0f && area1 <= 0.0f)
	{
		S2C_REDUCE_TO_EDGE(s, 1, 2, side23);
		s->v[0] = s->v[2];
	}
	else
	{
		// the origin is inside
		float scale = 1.0f / (area1 + area2 + area3);
		s->v[0].a = area1 * scale;
		s->v[1].a = area2 * scale;
		s->v[2].a = area3 * scale;
		s->count = 3;
	}
}

// s2ShapeDistance (reference src/distance.c:485-636) over raw vertex arrays. The simplex cache is input/output.
S2C_FN s2DistanceOutput s2cShapeDistance(s2DistanceCache* cache, const s2Vec2* vertsA, int countA, float radiusA,
										 s2Transform xfA, const s2Vec2* vertsB, int countB, float radiusB, s2Transform xfB,
										 bool useRadii)
{
	s2DistanceOutput output;
	output.pointA = s2cVec(0.0f, 0.0f);
	output.pointB = s2cVec(0.0f, 0.0f);
	output.distance = 0.0f;
	output.iterations = 0;

	// simplex from the cache (reference s2MakeSimplexFromCache, distance.c:174-218)
	// Every access to simplex.v[] and the cache index arrays below uses a literal index (loops over the at most three
	// vertices are unrolled with a guard): on the device the simplex then lives in registers instead of the local stack.
	s2cSimplex simplex;
	simplex.count = cache->count;
	S2C_UNROLL
	for (int i = 0; i < 3; ++i)
	{
		if (i < simplex.count)
		{
			s2cSimplexVertex* v = simplex.v + i;
			v->indexA = cache->indexA[i];
			v->indexB = cache->indexB[i];
			v->wA = s2TransformPoint(xfA, vertsA[v->indexA]);
			v->wB = s2TransformPoint(xfB, vertsB[v->indexB]);
			v->w = s2Sub(v->wB, v->wA);
			v->a = -1.0f;
		}
	}
	if (simplex.count == 0)
	{
		s2cSimplexVertex* v = simplex.v + 0;
		v->indexA = 0;
		v->indexB = 0;
		v->wA = s2TransformPoint(xfA, vertsA[0]);
		v->wB = s2TransformPoint(xfB, vertsB[0]);
		v->w = s2Sub(v->wB, v->wA);
		v->a = 1.0f;
		simplex.count = 1;
	}

	const int maxIters = 20;
	int saveA[3], saveB[3];
	int iter = 0;
	while (iter < maxIters)
	{
		int saveCount = simplex.count;
		S2C_UNROLL
		for (int i = 0; i < 3; ++i)
		{
			if (i < saveCount)
			{
				saveA[i] = simplex.v[i].indexA;
				saveB[i] = simplex.v[i].indexB;
			}
		}

		if (simplex.count == 2)
		{
			s2cSolveSimplex2(&simplex);
		}
		else if (simplex.count == 3)
		{
			s2cSolveSimplex3(&simplex);
		}

		// the origin is inside the triangle: overlap
		if (simplex.count == 3)
		{
			break;
		}

		// search direction (reference s2ComputeSimplexSearchDirection, distance.c:232-258)
		s2Vec2 d;
		if (simplex.count == 1)
		{
			d = s2Neg(simplex.v[0].w);
		}
		else
		{
			s2Vec2 e12 = s2Sub(simplex.v[1].w, simplex.v[0].w);
			float sgn = s2Cross(e12, s2Neg(simplex.v[0].w));
			if (sgn > 0.0f)
			{
				d = s2CrossSV(1.0f, e12);
			}
			else
			{
				d = s2CrossVS(e12, 1.0f);
			}
		}

		if (s2Dot(d, d) < FLT_EPSILON * FLT_EPSILON)
		{
			// the origin is (numerically) on the simplex: overlap, stop here
			break;
		}

		// the new vertex; it joins the simplex at v[count] (count is 1 or 2 here) unless it repeats a support point
		s2cSimplexVertex fresh;
		s2cSimplexVertex* vertex = &fresh;
		vertex->indexA = s2cFindSupport(vertsA, countA, s2InvRotateVector(xfA.q, s2Neg(d)));
		vertex->wA = s2TransformPoint(xfA, vertsA[vertex->indexA]);
		vertex->indexB = s2cFindSupport(vertsB, countB, s2InvRotateVector(xfB.q, d));
		vertex->wB = s2TransformPoint(xfB, vertsB[vertex->indexB]);
		vertex->w = s2Sub(vertex->wB, vertex->wA);

		++iter;

		// a repeated support point terminates the iteration
		bool duplicate = false;
		S2C_UNROLL
		for (int i = 0; i < 3; ++i)
		{
			if (i < saveCount && vertex->indexA == saveA[i] && vertex->indexB == saveB[i])
			{
				duplicate = true;
			}
		}
		if (duplicate)
		{
			break;
		}

		// (the weight `a` of the slot is left as it is, like the reference: the next s2cSolveSimplex* sets it)
		if (simplex.count == 1)
		{
			S2C_SET_SIMPLEX_VERTEX(simplex.v[1], fresh);
		}
		else
		{
			S2C_SET_SIMPLEX_VERTEX(simplex.v[2], fresh);
		}
		++simplex.count;
	}

	// witness points (reference s2ComputeSimplexWitnessPoints, distance.c:284-316)
	if (simplex.count == 1)
	{
		output.pointA = simplex.v[0].wA;
		output.pointB = simplex.v[0].wB;
	}
	else if (simplex.count == 2)
	{
		output.pointA = s2cWeight2(simplex.v[0].a, simplex.v[0].wA, simplex.v[1].a, simplex.v[1].wA);
		output.pointB = s2cWeight2(simplex.v[0].a, simplex.v[0].wB, simplex.v[1].a, simplex.v[1].wB);
	}
	else
	{
		output.pointA = s2cWeight3(simplex.v[0].a, simplex.v[0].wA, simplex.v[1].a, simplex.v[1].wA, simplex.v[2].a,
								   simplex.v[2].wA);
		output.pointB = output.pointA;
	}
	output.distance = s2Distance(output.pointA, output.pointB);
	output.iterations = iter;

	// cache the simplex (reference s2MakeSimplexCache + s2Simplex_Metric, distance.c:148-172, 220-230)
	if (simplex.count == 1)
	{
		cache->metric = 0.0f;
	}
	else if (simplex.count == 2)
	{
		cache->metric = s2Distance(simplex.v[0].w, simplex.v[1].w);
	}
	else
	{
		cache->metric = s2Cross(s2Sub(simplex.v[1].w, simplex.v[0].w), s2Sub(simplex.v[2].w, simplex.v[0].w));
	}
	cache->count = (uint16_t)simplex.count;
	S2C_UNROLL
	for (int i = 0; i < 3; ++i)
	{
		if (i < simplex.count)
		{
			cache->indexA[i] = (uint8_t)simplex.v[i].indexA;
			cache->indexB[i] = (uint8_t)simplex.v[i].indexB;
		}
	}

	if (useRadii)
	{
		if (output.distance < FLT_EPSILON)
		{
			s2Vec2 p = s2cVec(0.5f * (output.pointA.x + output.pointB.x), 0.5f * (output.pointA.y + output.pointB.y));
			output.pointA = p;
			output.pointB = p;
			output.distance = 0.0f;
		}
		else
		{
			float rA = radiusA;
			float rB = radiusB;
			output.distance = S2_MAX(0.0f, output.distance - rA - rB);
			s2Vec2 normal = s2cNormalize(s2Sub(output.pointB, output.pointA));
			s2Vec2 offsetA = s2cVec(rA * normal.x, rA * normal.y);
			s2Vec2 offsetB = s2cVec(rB * normal.x, rB * normal.y);
			output.pointA = s2Add(output.pointA, offsetA);
			output.pointB = s2Sub(output.pointB, offsetB);
		}
	}

	return output;
}

S2C_FN float s2cClampToUnit(float x)
{
	return S2_CLAMP(x, 0.0f, 1.0f);
}

// Closest points of the segments a0-a1 and b0-b1 (behaviour and arithmetic of reference src/distance.c:14-98, the
// clamped-projection scheme of Ericson 5.1.9). A: a0 + s u, B: b0 + t v with s, t in [0, 1] and w = a0 - b0; minimising
// |w + s u - t v|^2 gives s = (uv wv - wu vv) / (uu vv - uv^2), t = (uv s + wv) / vv; whenever t leaves [0, 1] it is
// pinned to the end it left through and s becomes the projection of that end point on A.
S2C_FN s2SegmentDistanceResult s2cSegmentDistance(s2Vec2 a0, s2Vec2 a1, s2Vec2 b0, s2Vec2 b1)
{
	s2Vec2 u = s2Sub(a1, a0), v = s2Sub(b1, b0), w = s2Sub(a0, b0);
	float uu = s2Dot(u, u), vv = s2Dot(v, v);
	float wv = s2Dot(w, v), wu = s2Dot(w, u);
	const float tiny = FLT_EPSILON * FLT_EPSILON;
	bool aIsPoint = uu < tiny, bIsPoint = vv < tiny;

	float sParam = 0.0f, tParam = 0.0f;
	if (aIsPoint || bIsPoint)
	{
		// point against segment (or point against point: both parameters stay 0)
		if (aIsPoint == false)
		{
			sParam = s2cClampToUnit(-wu / uu);
		}
		else if (bIsPoint == false)
		{
			tParam = s2cClampToUnit(wv / vv);
		}
	}
	else
	{
		float uv = s2Dot(u, v);
		float det = uu * vv - uv * uv; // zero: parallel, every s is as good, keep 0
		if (det != 0.0f)
		{
			sParam = s2cClampToUnit((uv * wv - wu * vv) / det);
		}
		tParam = (uv * sParam + wv) / vv;
		if (tParam < 0.0f)
		{
			tParam = 0.0f;
			sParam = s2cClampToUnit(-wu / uu);
		}
		else if (tParam > 1.0f)
		{
			tParam = 1.0f;
			sParam = s2cClampToUnit((uv - wu) / uu);
		}
	}

	s2SegmentDistanceResult out;
	out.fraction1 = sParam;
	out.fraction2 = tParam;
	out.closest1 = s2MulAdd(a0, sParam, u);
	out.closest2 = s2MulAdd(b0, tParam, v);
	out.distanceSquared = s2DistanceSquared(out.closest1, out.closest2);
	return out;
}

// ---- manifolds (reference src/manifold.c) -----------------------------------------------------------------------

// s2CollideCircles (reference src/manifold.c:16-49)
S2C_FN void s2cCollideCircles(s2Manifold* manifold, s2Vec2 centerA, float radiusA, s2Transform xfA, s2Vec2 centerB,
							  float radiusB, s2Transform xfB)
{
	s2cClearManifold(manifold);
	s2Transform xf = s2InvMulTransforms(xfA, xfB);
	s2Vec2 pointA = centerA;
	s2Vec2 pointB = s2TransformPoint(xf, centerB);

	float distance;
	s2Vec2 normal = s2cGetLengthAndNormalize(&distance, s2Sub(pointB, pointA));

	float separation = distance - radiusA - radiusB;
	if (separation > s2_speculativeDistance)
	{
		return;
	}

	s2Vec2 cA = s2MulAdd(pointA, radiusA, normal);
	s2Vec2 cB = s2MulAdd(pointB, -radiusB, normal);
	s2Vec2 contactPointA = s2Lerp(cA, cB, 0.5f);

	manifold->normal = s2RotateVector(xfA.q, normal);
	manifold->points[0].localAnchorA = contactPointA;
	manifold->points[0].localAnchorB = s2InvTransformPoint(xf, contactPointA);
	manifold->points[0].separation = separation;
	manifold->points[0].id = 0;
	manifold->pointCount = 1;
}

// s2CollideCapsuleAndCircle (reference src/manifold.c:52-112); a segment is a capsule of radius 0 (:656-660)
S2C_FN void s2cCollideCapsuleAndCircle(s2Manifold* manifold, s2Vec2 p1, s2Vec2 p2, float radiusA, s2Transform xfA,
									   s2Vec2 centerB, float radiusB, s2Transform xfB)
{
	s2cClearManifold(manifold);
	s2Transform xf = s2InvMulTransforms(xfA, xfB);
	s2Vec2 pB = s2TransformPoint(xf, centerB);

	s2Vec2 e = s2Sub(p2, p1);
	s2Vec2 pA;
	float s1 = s2Dot(s2Sub(pB, p1), e);
	float s2 = s2Dot(s2Sub(p2, pB), e);
	if (s1 < 0.0f)
	{
		pA = p1;
	}
	else if (s2 < 0.0f)
	{
		pA = p2;
	}
	else
	{
		float s = s1 / s2Dot(e, e);
		pA = s2MulAdd(p1, s, e);
	}

	float distance;
	s2Vec2 normal = s2cGetLengthAndNormalize(&distance, s2Sub(pB, pA));

	float separation = distance - radiusA - radiusB;
	if (separation > s2_speculativeDistance)
	{
		return;
	}

	s2Vec2 cA = s2MulAdd(pA, radiusA, normal);
	s2Vec2 cB = s2MulAdd(pB, -radiusB, normal);
	s2Vec2 contactPointA = s2Lerp(cA, cB, 0.5f);

	manifold->normal = s2RotateVector(xfA.q, normal);
	manifold->points[0].localAnchorA = contactPointA;
	manifold->points[0].localAnchorB = s2InvTransformPoint(xf, contactPointA);
	manifold->points[0].separation = separation;
	manifold->points[0].id = 0;
	manifold->pointCount = 1;
}

// s2CollidePolygonAndCircle (reference src/manifold.c:114-222)
S2C_FN void s2cCollidePolygonAndCircle(s2Manifold* manifold, const s2Vec2* vertices, const s2Vec2* normals, int vertexCount,
									   float radiusA, s2Transform xfA, s2Vec2 centerB, float radiusB, s2Transform xfB)
{
	s2cClearManifold(manifold);
	s2Transform xf = s2InvMulTransforms(xfA, xfB);
	s2Vec2 c = s2TransformPoint(xf, centerB);
	float radius = radiusA + radiusB;

	int normalIndex = 0;
	float separation = -FLT_MAX;
	for (int i = 0; i < vertexCount; ++i)
	{
		float s = s2Dot(normals[i], s2Sub(c, vertices[i]));
		if (s > separation)
		{
			separation = s;
			normalIndex = i;
		}
	}

	if (separation > radius + s2_speculativeDistance)
	{
		return;
	}

	int vertIndex1 = normalIndex;
	int vertIndex2 = vertIndex1 + 1 < vertexCount ? vertIndex1 + 1 : 0;
	s2Vec2 v1 = vertices[vertIndex1];
	s2Vec2 v2 = vertices[vertIndex2];

	float u1 = s2Dot(s2Sub(c, v1), s2Sub(v2, v1));
	float u2 = s2Dot(s2Sub(c, v2), s2Sub(v1, v2));

	if (u1 < 0.0f && separation > FLT_EPSILON)
	{
		// closest to v1, outside
		s2Vec2 normal = s2cNormalize(s2Sub(c, v1));
		separation = s2Dot(s2Sub(c, v1), normal);
		if (separation > radius + s2_speculativeDistance)
		{
			return;
		}
		s2Vec2 cA = s2MulAdd(v1, radiusA, normal);
		s2Vec2 cB = s2MulSub(c, radiusB, normal);
		s2Vec2 contactPointA = s2Lerp(cA, cB, 0.5f);
		manifold->normal = s2RotateVector(xfA.q, normal);
		manifold->points[0].localAnchorA = contactPointA;
		manifold->points[0].localAnchorB = s2InvTransformPoint(xf, contactPointA);
		manifold->points[0].separation = s2Dot(s2Sub(cB, cA), normal);
		manifold->points[0].id = 0;
		manifold->pointCount = 1;
	}
	else if (u2 < 0.0f && separation > FLT_EPSILON)
	{
		// closest to v2, outside
		s2Vec2 normal = s2cNormalize(s2Sub(c, v2));
		separation = s2Dot(s2Sub(c, v2), normal);
		if (separation > radius + s2_speculativeDistance)
		{
			return;
		}
		s2Vec2 cA = s2MulAdd(v2, radiusA, normal);
		s2Vec2 cB = s2MulSub(c, radiusB, normal);
		s2Vec2 contactPointA = s2Lerp(cA, cB, 0.5f);
		manifold->normal = s2RotateVector(xfA.q, normal);
		manifold->points[0].localAnchorA = contactPointA;
		manifold->points[0].localAnchorB = s2InvTransformPoint(xf, contactPointA);
		manifold->points[0].separation = s2Dot(s2Sub(cB, cA), normal);
		manifold->points[0].id = 0;
		manifold->pointCount = 1;
	}
	else
	{
		// face region (the centre may be inside)
		s2Vec2 normal = normals[normalIndex];
		manifold->normal = s2RotateVector(xfA.q, normal);
		s2Vec2 cA = s2MulAdd(c, radiusA - s2Dot(s2Sub(c, v1), normal), normal);
		s2Vec2 cB = s2MulSub(c, radiusB, normal);
		s2Vec2 contactPointA = s2Lerp(cA, cB, 0.5f);
		manifold->points[0].localAnchorA = contactPointA;
		manifold->points[0].localAnchorB = s2InvTransformPoint(xf, contactPointA);
		manifold->points[0].separation = separation - radius;
		manifold->points[0].id = 0;
		manifold->pointCount = 1;
	}
}

// anchors of the manifold's points expressed in shape B's frame (xf = frame of B seen from A). Literal point indices: on
// the device the manifold then stays in registers.
S2C_FN void s2cAnchorsIntoFrameB(s2Manifold* manifold, s2Transform xf)
{
	if (manifold->pointCount > 0)
	{
		manifold->points[0].localAnchorB = s2InvTransformPoint(xf, manifold->points[0].localAnchorA);
	}
	if (manifold->pointCount > 1)
	{
		manifold->points[1].localAnchorB = s2InvTransformPoint(xf, manifold->points[1].localAnchorA);
	}
}

// next vertex of a polygon, cyclically
S2C_FN int s2cNextIndex(const s2Polygon* poly, int i)
{
	return i + 1 < poly->count ? i + 1 : 0;
}

// Two-point manifold of a REFERENCE edge (the face that separates best) and the INCIDENT edge of the other polygon (the
// one most anti-parallel to it), everything in polyA's frame (behaviour, arithmetic and feature ids of reference
// src/manifold.c:248-399). `flip`: polyB owns the reference edge. The incident edge runs against the reference edge, so its
// END vertex is the one near the reference edge's start; each end is cut back to the side plane through the corresponding
// end of the reference edge if it sticks out, then moved to the mid surface of the two (rounded) polygons.
S2C_FN void s2cClipPolygons(s2Manifold* manifold, const s2Polygon* polyA, const s2Polygon* polyB, int edgeA, int edgeB, bool flip)
{
	s2cClearManifold(manifold);

	const s2Polygon* refPoly = flip ? polyB : polyA;
	const s2Polygon* incPoly = flip ? polyA : polyB;
	int refStart = flip ? edgeB : edgeA, incStart = flip ? edgeA : edgeB;
	int refEnd = s2cNextIndex(refPoly, refStart), incEnd = s2cNextIndex(incPoly, incStart);

	s2Vec2 normal = refPoly->normals[refStart];
	s2Vec2 along = s2CrossSV(1.0f, normal);
	s2Vec2 origin = refPoly->vertices[refStart];
	s2Vec2 incFirst = incPoly->vertices[incStart], incSecond = incPoly->vertices[incEnd];

	// positions along the reference edge (its start is 0)
	float refSpan = s2Dot(s2Sub(refPoly->vertices[refEnd], origin), along);
	float firstAt = s2Dot(s2Sub(incFirst, origin), along);
	float secondAt = s2Dot(s2Sub(incSecond, origin), along);
	bool cuttable = firstAt - secondAt > FLT_EPSILON; // a degenerate incident edge is not interpolated

	s2Vec2 nearStart = incSecond, nearEnd = incFirst;
	if (secondAt < 0.0f && cuttable)
	{
		nearStart = s2Lerp(incSecond, incFirst, (0.0f - secondAt) / (firstAt - secondAt));
	}
	if (firstAt > refSpan && cuttable)
	{
		nearEnd = s2Lerp(incSecond, incFirst, (refSpan - secondAt) / (firstAt - secondAt));
	}

	float gapAtStart = s2Dot(s2Sub(nearStart, origin), normal);
	float gapAtEnd = s2Dot(s2Sub(nearEnd, origin), normal);

	float refRadius = refPoly->radius, incRadius = incPoly->radius;
	nearStart = s2MulAdd(nearStart, 0.5f * (refRadius - incRadius - gapAtStart), normal);
	nearEnd = s2MulAdd(nearEnd, 0.5f * (refRadius - incRadius - gapAtEnd), normal);
	float bothRadii = refRadius + incRadius;

	// points are listed in polyA's edge direction; an id names (vertex of A, vertex of B)
	s2ManifoldPoint* first = manifold->points + 0;
	s2ManifoldPoint* second = manifold->points + 1;
	if (flip)
	{
		manifold->normal = s2Neg(normal);
		first->localAnchorA = nearEnd;
		first->separation = gapAtEnd - bothRadii;
		first->id = (uint16_t)S2_MAKE_ID(incStart, refEnd);
		second->localAnchorA = nearStart;
		second->separation = gapAtStart - bothRadii;
		second->id = (uint16_t)S2_MAKE_ID(incEnd, refStart);
	}
	else
	{
		manifold->normal = normal;
		first->localAnchorA = nearStart;
		first->separation = gapAtStart - bothRadii;
		first->id = (uint16_t)S2_MAKE_ID(refStart, incEnd);
		second->localAnchorA = nearEnd;
		second->separation = gapAtEnd - bothRadii;
		second->id = (uint16_t)S2_MAKE_ID(refEnd, incStart);
	}
	manifold->pointCount = 2;
}

// How far `other` stays outside the face (point `onFace`, outward normal `outward`): the signed distance of its deepest vertex.
S2C_FN float s2cClearanceOfFace(s2Vec2 outward, s2Vec2 onFace, const s2Polygon* other)
{
	float deepest = FLT_MAX;
	for (int k = 0; k < other->count; ++k)
	{
		float d = s2Dot(outward, s2Sub(other->vertices[k], onFace));
		deepest = d < deepest ? d : deepest;
	}
	return deepest;
}

// The face of `poly` that keeps `other` furthest out (the first one on ties) and that clearance (behaviour of reference
// src/manifold.c:402-438: the separating-axis search over the face normals of one polygon).
S2C_FN float s2cFindMaxSeparation(int* edgeIndex, const s2Polygon* poly, const s2Polygon* other)
{
	int bestFace = 0;
	float bestClearance = -FLT_MAX;
	for (int face = 0; face < poly->count; ++face)
	{
		float clearance = s2cClearanceOfFace(poly->normals[face], poly->vertices[face], other);
		if (clearance > bestClearance)
		{
			bestClearance = clearance;
			bestFace = face;
		}
	}
	*edgeIndex = bestFace;
	return bestClearance;
}

// the edge of `poly` whose normal opposes `direction` most (the first one on ties)
S2C_FN int s2cMostOpposedEdge(const s2Polygon* poly, s2Vec2 direction)
{
	int best = 0;
	float least = FLT_MAX;
	for (int k = 0; k < poly->count; ++k)
	{
		float alignment = s2Dot(direction, poly->normals[k]);
		if (alignment < least)
		{
			least = alignment;
			best = k;
		}
	}
	return best;
}

// Manifold of two polygons known to overlap (or nearly): reference face = the face of either polygon with the largest
// clearance (polyA's on ties), incident edge = the other polygon's edge most opposed to it (behaviour of reference
// src/manifold.c:441-493).
S2C_FN void s2cPolygonSAT(s2Manifold* manifold, const s2Polygon* polyA, const s2Polygon* polyB)
{
	int faceA = 0, faceB = 0;
	float clearanceA = s2cFindMaxSeparation(&faceA, polyA, polyB);
	float clearanceB = s2cFindMaxSeparation(&faceB, polyB, polyA);
	bool bOwnsReference = clearanceB > clearanceA;
	if (bOwnsReference)
	{
		faceA = s2cMostOpposedEdge(polyA, polyB->normals[faceB]);
	}
	else
	{
		faceB = s2cMostOpposedEdge(polyB, polyA->normals[faceA]);
	}
	s2cClipPolygons(manifold, polyA, polyB, faceA, faceB, bOwnsReference);
}

// s2CollidePolygons (reference src/manifold.c:509-650): GJK closest features, SAT when (nearly) overlapping,
// vertex-vertex or edge clipping otherwise. Capsules and segments arrive here as 2-gons.
// This is the part after polyB has been moved into polyA's frame: `localB` = polyB under xf = xfA^-1 * xfB. The device
// kernel builds localB itself, straight from the shape columns into shared memory (narrowphase.cu).
S2C_FN void s2cCollidePolygonsLocal(s2Manifold* manifold, const s2Polygon* polyA, s2Transform xfA, const s2Polygon* localB,
									s2Transform xf, s2DistanceCache* cache)
{
	s2cClearManifold(manifold);
	float radius = polyA->radius + localB->radius;

	s2Transform identity;
	identity.p = s2cVec(0.0f, 0.0f);
	identity.q.s = 0.0f;
	identity.q.c = 1.0f;
	int countA = S2_MIN(polyA->count, s2_maxPolygonVertices);
	int countB = S2_MIN(localB->count, s2_maxPolygonVertices);
	s2DistanceOutput output =
		s2cShapeDistance(cache, polyA->vertices, countA, 0.0f, identity, localB->vertices, countB, 0.0f, identity, false);

	if (output.distance > radius + s2_speculativeDistance)
	{
		return;
	}

	if (output.distance < 0.1f * s2_linearSlop)
	{
		s2cPolygonSAT(manifold, polyA, localB);
		if (manifold->pointCount > 0)
		{
			manifold->normal = s2RotateVector(xfA.q, manifold->normal);
			s2cAnchorsIntoFrameB(manifold, xf);
		}
		return;
	}

	if (cache->count == 1)
	{
		// vertex-vertex
		s2Vec2 pA = output.pointA;
		s2Vec2 pB = output.pointB;
		float distance = output.distance;
		s2Vec2 normal = s2cNormalize(s2Sub(pB, pA));
		s2Vec2 contactPointA = s2MulAdd(pB, 0.5f * (polyA->radius - localB->radius - distance), normal);

		manifold->normal = s2RotateVector(xfA.q, normal);
		manifold->points[0].localAnchorA = contactPointA;
		manifold->points[0].localAnchorB = s2InvTransformPoint(xf, contactPointA);
		manifold->points[0].separation = distance - radius;
		manifold->points[0].id = (uint16_t)S2_MAKE_ID(cache->indexA[0], cache->indexB[0]);
		manifold->pointCount = 1;
		return;
	}

	// vertex-edge: pick reference and incident edges around the closest features
	bool flip;
	int edgeA, edgeB;
	int a1 = cache->indexA[0];
	int a2 = cache->indexA[1];
	int b1 = cache->indexB[0];
	int b2 = cache->indexB[1];

	if (a1 == a2)
	{
		// one vertex of A against an edge of B
		s2Vec2 axis = s2Sub(output.pointA, output.pointB);
		float dot1 = s2Dot(axis, localB->normals[b1]);
		float dot2 = s2Dot(axis, localB->normals[b2]);
		edgeB = dot1 > dot2 ? b1 : b2;
		flip = true;

		axis = localB->normals[edgeB];
		int edgeA1 = a1;
		int edgeA2 = edgeA1 == 0 ? polyA->count - 1 : edgeA1 - 1;
		dot1 = s2Dot(axis, polyA->normals[edgeA1]);
		dot2 = s2Dot(axis, polyA->normals[edgeA2]);
		edgeA = dot1 < dot2 ? edgeA1 : edgeA2;
	}
	else
	{
		s2Vec2 axis = s2Sub(output.pointB, output.pointA);
		float dot1 = s2Dot(axis, polyA->normals[a1]);
		float dot2 = s2Dot(axis, polyA->normals[a2]);
		edgeA = dot1 > dot2 ? a1 : a2;
		flip = false;

		axis = polyA->normals[edgeA];
		int edgeB1 = b1;
		int edgeB2 = edgeB1 == 0 ? localB->count - 1 : edgeB1 - 1;
		dot1 = s2Dot(axis, localB->normals[edgeB1]);
		dot2 = s2Dot(axis, localB->normals[edgeB2]);
		edgeB = dot1 < dot2 ? edgeB1 : edgeB2;
	}

	s2cClipPolygons(manifold, polyA, localB, edgeA, edgeB, flip);
	if (manifold->pointCount > 0)
	{
		manifold->normal = s2RotateVector(xfA.q, manifold->normal);
		s2cAnchorsIntoFrameB(manifold, xf);
	}
}

// polyB in polyA's frame (reference src/manifold.c:515-527)
S2C_FN void s2cMakeLocalPolygon(s2Polygon* localB, const s2Polygon* polyB, s2Transform xf)
{
	localB->count = polyB->count;
	localB->radius = polyB->radius;
	for (int i = 0; i < localB->count; ++i)
	{
		localB->vertices[i] = s2TransformPoint(xf, polyB->vertices[i]);
		localB->normals[i] = s2RotateVector(xf.q, polyB->normals[i]);
	}
}

S2C_FN void s2cCollidePolygons(s2Manifold* manifold, const s2Polygon* polyA, s2Transform xfA, const s2Polygon* polyB,
							   s2Transform xfB, s2DistanceCache* cache)
{
	s2Transform xf = s2InvMulTransforms(xfA, xfB);
	s2Polygon localPolyB;
	s2cMakeLocalPolygon(&localPolyB, polyB, xf);
	s2cCollidePolygonsLocal(manifold, polyA, xfA, &localPolyB, xf, cache);
}
