// solver2d-b200 — collision geometry (API of reference include/solver2d/geometry.h).
// Geometry is authored on the host, copied into the shape at creation, and mirrored once into the device shape
// table; only its AABB refit and the narrow phase touch it on the step path.
#pragma once

#include "solver2d/constants.h"
#include "solver2d/types.h"

typedef struct s2Hull s2Hull;

typedef struct s2MassData
{
	float mass;
	s2Vec2 center; // centroid relative to the shape origin
	float I;	   // rotational inertia about the shape origin
} s2MassData;

typedef struct s2Circle
{
	s2Vec2 point;
	float radius;
} s2Circle;

typedef struct s2Capsule
{
	s2Vec2 point1, point2;
	float radius;
} s2Capsule;

// Convex polygon, counter-clockwise, at most s2_maxPolygonVertices vertices, optionally rounded by `radius`.
typedef struct s2Polygon
{
	s2Vec2 vertices[s2_maxPolygonVertices];
	s2Vec2 normals[s2_maxPolygonVertices];
	float radius;
	int32_t count;
} s2Polygon;

// two-sided line segment
typedef struct s2Segment
{
	s2Vec2 point1, point2;
} s2Segment;

// one-sided segment with ghost vertices (declared for API completeness; no solver2d shape uses it)
typedef struct s2SmoothSegment
{
	s2Vec2 ghost1;
	s2Vec2 point1, point2;
	s2Vec2 ghost2;
} s2SmoothSegment;

#ifdef __cplusplus
extern "C"
{
#endif

bool s2IsValidRay(const s2RayCastInput* input);

s2Polygon s2MakePolygon(const s2Hull* hull);
s2Polygon s2MakeSquare(float h);
s2Polygon s2MakeBox(float hx, float hy);
s2Polygon s2MakeRoundedBox(float hx, float hy, float radius);
s2Polygon s2MakeOffsetBox(float hx, float hy, s2Vec2 center, float angle);
s2Polygon s2MakeCapsule(s2Vec2 p1, s2Vec2 p2, float radius);

s2MassData s2ComputeCircleMass(const s2Circle* shape, float density);
s2MassData s2ComputeCapsuleMass(const s2Capsule* shape, float density);
s2MassData s2ComputePolygonMass(const s2Polygon* shape, float density);

s2Box s2ComputeCircleAABB(const s2Circle* shape, s2Transform xf);
s2Box s2ComputeCapsuleAABB(const s2Capsule* shape, s2Transform xf);
s2Box s2ComputePolygonAABB(const s2Polygon* shape, s2Transform xf);
s2Box s2ComputeSegmentAABB(const s2Segment* shape, s2Transform xf);

// point tests in shape-local space
bool s2PointInCircle(s2Vec2 point, const s2Circle* shape);
bool s2PointInCapsule(s2Vec2 point, const s2Capsule* shape);
bool s2PointInPolygon(s2Vec2 point, const s2Polygon* shape);

// ray casts in shape-local space; an initial overlap is a miss
s2RayCastOutput s2RayCastCircle(const s2RayCastInput* input, const s2Circle* shape);
s2RayCastOutput s2RayCastCapsule(const s2RayCastInput* input, const s2Capsule* shape);
s2RayCastOutput s2RayCastSegment(const s2RayCastInput* input, const s2Segment* shape);
s2RayCastOutput s2RayCastPolygon(const s2RayCastInput* input, const s2Polygon* shape);

#ifdef __cplusplus
}
#endif
