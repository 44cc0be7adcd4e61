// solver2d-b200 — contact manifolds (API of reference include/solver2d/manifold.h).
// On the step path manifolds live in the device contact table (one SoA row per shape pair, persistent across
// steps); the structs below are the host-visible form used by s2World_Draw and by the host-callable s2Collide*.
#pragma once

#include "solver2d/types.h"

#define s2_nullFeature UCHAR_MAX
#define S2_MAKE_ID(A, B) ((uint8_t)(A) << 8 | (uint8_t)(B))

typedef struct s2Circle s2Circle;
typedef struct s2Capsule s2Capsule;
typedef struct s2DistanceCache s2DistanceCache;
typedef struct s2Polygon s2Polygon;
typedef struct s2Segment s2Segment;
typedef struct s2SmoothSegment s2SmoothSegment;

typedef struct s2ManifoldPoint
{
	// contact location relative to each body origin, in that body's frame
	s2Vec2 localAnchorA;
	s2Vec2 localAnchorB;

	// persistent friction anchors / normals (TGS_Sticky)
	s2Vec2 frictionAnchorA;
	s2Vec2 frictionAnchorB;
	s2Vec2 frictionNormalA;
	s2Vec2 frictionNormalB;

	float separation;
	float normalImpulse;
	float tangentImpulse;
	uint16_t id; // feature pair, matches points across steps

	bool persisted;
} s2ManifoldPoint;

typedef struct s2Manifold
{
	s2ManifoldPoint points[2];
	s2Vec2 normal; // world space, from A to B
	int32_t pointCount;
	int32_t constraintIndex;
	bool frictionPersisted;
} s2Manifold;

static const s2Manifold s2_emptyManifold = S2_ZERO_INIT;

#ifdef __cplusplus
extern "C"
{
#endif

s2Manifold s2CollideCircles(const s2Circle* circleA, s2Transform xfA, const s2Circle* circleB, s2Transform xfB);
s2Manifold s2CollideCapsuleAndCircle(const s2Capsule* capsuleA, s2Transform xfA, const s2Circle* circleB, s2Transform xfB);
s2Manifold s2CollideSegmentAndCircle(const s2Segment* segmentA, s2Transform xfA, const s2Circle* circleB, s2Transform xfB);
s2Manifold s2CollidePolygonAndCircle(const s2Polygon* polygonA, s2Transform xfA, const s2Circle* circleB, s2Transform xfB);
s2Manifold s2CollideCapsules(const s2Capsule* capsuleA, s2Transform xfA, const s2Capsule* capsuleB, s2Transform xfB,
							 s2DistanceCache* cache);
s2Manifold s2CollideSegmentAndCapsule(const s2Segment* segmentA, s2Transform xfA, const s2Capsule* capsuleB, s2Transform xfB,
									  s2DistanceCache* cache);
s2Manifold s2CollidePolygonAndCapsule(const s2Polygon* polygonA, s2Transform xfA, const s2Capsule* capsuleB, s2Transform xfB,
									  s2DistanceCache* cache);
s2Manifold s2CollidePolygons(const s2Polygon* polyA, s2Transform xfA, const s2Polygon* polyB, s2Transform xfB,
							 s2DistanceCache* cache);
s2Manifold s2CollideSegmentAndPolygon(const s2Segment* segmentA, s2Transform xfA, const s2Polygon* polygonB, s2Transform xfB,
									  s2DistanceCache* cache);

#ifdef __cplusplus
}
#endif
