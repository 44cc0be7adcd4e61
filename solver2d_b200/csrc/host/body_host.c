// solver2d-b200 — bodies and shapes behind the public API (behaviour of reference src/body.c, src/shape.c).
// All of this is host bookkeeping: it edits the host copy of an object and marks its row dirty; the next s2World_Step
// uploads the row. Reads of simulation state first make the host copy current (s2SyncStateToHost).
#include "s2_host.h"

#include <string.h>

s2Body* s2GetBody(s2World* world, s2BodyId id)
{
	S2_ASSERT(0 <= id.index && id.index < world->bodyPool.capacity);
	s2Body* body = world->bodies + id.index;
	S2_ASSERT(s2ObjectValid(&body->object));
	S2_ASSERT(id.revision == body->object.revision);
	return body;
}

// reference src/body.c:17-63
s2BodyId s2CreateBody(s2WorldId worldId, const s2BodyDef* def)
{
	s2World* world = s2GetWorldFromId(worldId);
	s2Body* b = (s2Body*)s2AllocObject(&world->bodyPool);
	world->bodies = (s2Body*)world->bodyPool.memory;
	if (b->object.index + 1 > world->bodyHighWater)
	{
		world->bodyHighWater = b->object.index + 1;
	}
	b->onDevice = false;

	b->type = def->type;
	b->origin = def->position;
	b->position = def->position;
	b->rot = s2MakeRot(def->angle);
	b->localCenter = s2Vec2_zero;
	b->linearVelocity = def->linearVelocity;
	b->angularVelocity = def->angularVelocity;
	b->force = s2Vec2_zero;
	b->torque = 0.0f;
	b->shapeList = S2_NULL_INDEX;
	b->jointCount = 0;
	b->mass = 0.0f;
	b->invMass = 0.0f;
	b->I = 0.0f;
	b->invI = 0.0f;
	b->linearDamping = def->linearDamping;
	b->angularDamping = def->angularDamping;
	b->gravityScale = def->gravityScale;
	b->userData = def->userData;
	b->world = worldId.index;
	// rowDirty / forceDirty keep their value: a reused slot may already sit in the dirty list
	s2MarkBodyDirty(world, b);

	s2BodyId id = {b->object.index, worldId.index, b->object.revision};
	return id;
}

static void s2DestroyShapeProxy(s2World* world, s2Shape* shape)
{
	int type = shape->proxyKey & 0xF;
	s2FreeProxyId(world->proxyIds + type, shape->proxyKey >> 4);
	shape->proxyKey = S2_NULL_INDEX;
}

// reference src/body.c:75-150. Contacts of the body live on the device: invalidating its shape rows makes the next pair
// pass drop them.
void s2DestroyBody(s2BodyId bodyId)
{
	s2World* world = s2GetWorldFromIndex(bodyId.world);
	s2Body* body = world->bodies + bodyId.index;
	S2_ASSERT(body->jointCount == 0);

	int32_t shapeIndex = body->shapeList;
	while (shapeIndex != S2_NULL_INDEX)
	{
		s2Shape* shape = world->shapes + shapeIndex;
		shapeIndex = shape->nextShapeIndex;
		s2DestroyShapeProxy(world, shape);
		s2FreeObject(&world->shapePool, &shape->object);
		s2MarkShapeDirty(world, shape);
	}
	s2FreeObject(&world->bodyPool, &body->object);
	body->onDevice = false;
	s2MarkBodyDirty(world, body);
}

s2Box s2Shape_ComputeAABB(const s2Shape* shape, s2Transform xf)
{
	switch (shape->type)
	{
		case s2_capsuleShape:
			return s2ComputeCapsuleAABB(&shape->capsule, xf);
		case s2_circleShape:
			return s2ComputeCircleAABB(&shape->circle, xf);
		case s2_polygonShape:
			return s2ComputePolygonAABB(&shape->polygon, xf);
		case s2_segmentShape:
			return s2ComputeSegmentAABB(&shape->segment, xf);
		default:
		{
			s2Box empty = {xf.p, xf.p};
			return empty;
		}
	}
}

s2MassData s2Shape_ComputeMass(const s2Shape* shape)
{
	switch (shape->type)
	{
		case s2_capsuleShape:
			return s2ComputeCapsuleMass(&shape->capsule, shape->density);
		case s2_circleShape:
			return s2ComputeCircleMass(&shape->circle, shape->density);
		case s2_polygonShape:
			return s2ComputePolygonMass(&shape->polygon, shape->density);
		default:
		{
			s2MassData zero = {0};
			return zero;
		}
	}
}

// reference src/body.c:152-218: total mass, centre of mass and inertia about it from the attached shapes
static void s2UpdateBodyMass(s2World* world, s2Body* b)
{
	b->mass = 0.0f;
	b->invMass = 0.0f;
	b->I = 0.0f;
	b->invI = 0.0f;
	b->localCenter = s2Vec2_zero;

	if (b->type != s2_dynamicBody)
	{
		b->position = b->origin;
		return;
	}

	s2Vec2 localCenter = s2Vec2_zero;
	for (int32_t si = b->shapeList; si != S2_NULL_INDEX;)
	{
		const s2Shape* s = world->shapes + si;
		si = s->nextShapeIndex;
		if (s->density == 0.0f)
		{
			continue;
		}
		s2MassData md = s2Shape_ComputeMass(s);
		b->mass += md.mass;
		localCenter = s2MulAdd(localCenter, md.mass, md.center);
		b->I += md.I;
	}

	if (b->mass > 0.0f)
	{
		b->invMass = 1.0f / b->mass;
		localCenter = s2MulSV(b->invMass, localCenter);
	}

	if (b->I > 0.0f)
	{
		// inertia about the centre of mass
		b->I -= b->mass * s2Dot(localCenter, localCenter);
		b->invI = 1.0f / b->I;
	}
	else
	{
		b->I = 0.0f;
		b->invI = 0.0f;
	}

	s2Vec2 oldCenter = b->position;
	b->localCenter = localCenter;
	b->position = s2Add(s2RotateVector(b->rot, b->localCenter), b->origin);

	// the velocity refers to the centre of mass
	s2Vec2 deltaLinear = s2CrossSV(b->angularVelocity, s2Sub(b->position, oldCenter));
	b->linearVelocity = s2Add(b->linearVelocity, deltaLinear);
}

// reference src/body.c:220-280 + s2Shape_CreateProxy (src/shape.c:48-67)
static s2ShapeId s2CreateShape(s2BodyId bodyId, const s2ShapeDef* def, const void* geometry, s2ShapeType shapeType)
{
	s2World* world = s2GetWorldFromIndex(bodyId.world);
	s2SyncStateToHost(world); // the body's transform and velocity are read below

	s2Shape* shape = (s2Shape*)s2AllocObject(&world->shapePool);
	world->shapes = (s2Shape*)world->shapePool.memory;
	s2Body* body = world->bodies + bodyId.index;

	switch (shapeType)
	{
		case s2_capsuleShape:
			shape->capsule = *(const s2Capsule*)geometry;
			break;
		case s2_circleShape:
			shape->circle = *(const s2Circle*)geometry;
			break;
		case s2_polygonShape:
			shape->polygon = *(const s2Polygon*)geometry;
			break;
		case s2_segmentShape:
			shape->segment = *(const s2Segment*)geometry;
			break;
		default:
			break;
	}

	shape->bodyIndex = body->object.index;
	shape->type = shapeType;
	shape->density = def->density;
	shape->friction = def->friction;
	shape->restitution = def->restitution;
	shape->userData = def->userData;
	shape->filter = def->filter;
	shape->fresh = true; // rowDirty keeps its value: a reused slot may already sit in the dirty list

	// tight AABB + speculative margin; fat AABB adds the move margin for movable bodies
	s2Transform xf = {body->origin, body->rot};
	shape->aabb = s2Shape_ComputeAABB(shape, xf);
	shape->aabb.lowerBound.x -= s2_speculativeDistance;
	shape->aabb.lowerBound.y -= s2_speculativeDistance;
	shape->aabb.upperBound.x += s2_speculativeDistance;
	shape->aabb.upperBound.y += s2_speculativeDistance;
	float margin = body->type == s2_staticBody ? s2_speculativeDistance : s2_aabbMargin + s2_speculativeDistance;
	shape->fatAABB.lowerBound.x = shape->aabb.lowerBound.x - margin;
	shape->fatAABB.lowerBound.y = shape->aabb.lowerBound.y - margin;
	shape->fatAABB.upperBound.x = shape->aabb.upperBound.x + margin;
	shape->fatAABB.upperBound.y = shape->aabb.upperBound.y + margin;

	// proxy key of the reference broad phase: (tree node id << 4) | body type (reference src/broad_phase.h:18-20)
	int32_t proxyId = s2AllocProxyId(world->proxyIds + body->type);
	shape->proxyKey = (proxyId << 4) | (int32_t)body->type;

	shape->nextShapeIndex = body->shapeList;
	body->shapeList = shape->object.index;

	if (shape->density)
	{
		s2UpdateBodyMass(world, body);
	}

	s2MarkShapeDirty(world, shape);
	s2MarkBodyDirty(world, body);

	s2ShapeId id = {shape->object.index, bodyId.world, shape->object.revision};
	return id;
}

s2ShapeId s2CreateCircleShape(s2BodyId bodyId, const s2ShapeDef* def, const s2Circle* circle)
{
	return s2CreateShape(bodyId, def, circle, s2_circleShape);
}

s2ShapeId s2CreatePolygonShape(s2BodyId bodyId, const s2ShapeDef* def, const s2Polygon* polygon)
{
	return s2CreateShape(bodyId, def, polygon, s2_polygonShape);
}

s2ShapeId s2CreateSegmentShape(s2BodyId bodyId, const s2ShapeDef* def, const s2Segment* segment)
{
	if (s2DistanceSquared(segment->point1, segment->point2) <= s2_linearSlop * s2_linearSlop)
	{
		return s2_nullShapeId;
	}
	return s2CreateShape(bodyId, def, segment, s2_segmentShape);
}

s2ShapeId s2CreateCapsuleShape(s2BodyId bodyId, const s2ShapeDef* def, const s2Capsule* capsule)
{
	if (s2DistanceSquared(capsule->point1, capsule->point2) <= s2_linearSlop * s2_linearSlop)
	{
		return s2_nullShapeId;
	}
	return s2CreateShape(bodyId, def, capsule, s2_capsuleShape);
}

// ---- accessors (reference src/body.c:313-384) ---------------------------------------------------------------------

s2Vec2 s2Body_GetPosition(s2BodyId bodyId)
{
	s2World* world = s2GetWorldFromIndex(bodyId.world);
	s2SyncStateToHost(world);
	return world->bodies[bodyId.index].origin;
}

float s2Body_GetAngle(s2BodyId bodyId)
{
	s2World* world = s2GetWorldFromIndex(bodyId.world);
	s2SyncStateToHost(world);
	return s2Rot_GetAngle(world->bodies[bodyId.index].rot);
}

s2Vec2 s2Body_GetLocalPoint(s2BodyId bodyId, s2Vec2 globalPoint)
{
	s2World* world = s2GetWorldFromIndex(bodyId.world);
	s2SyncStateToHost(world);
	s2Body* body = s2GetBody(world, bodyId);
	s2Transform xf = {body->origin, body->rot};
	return s2InvTransformPoint(xf, globalPoint);
}

// Declared by the reference API (solver2d.h:38) but never defined there; provided here as a teleport.
void s2Body_SetTransform(s2BodyId bodyId, s2Vec2 position, float angle)
{
	s2World* world = s2GetWorldFromIndex(bodyId.world);
	s2SyncStateToHost(world);
	s2SyncBoxesToHost(world);
	s2Body* body = s2GetBody(world, bodyId);
	body->origin = position;
	body->rot = s2MakeRot(angle);
	body->position = s2Add(s2RotateVector(body->rot, body->localCenter), body->origin);
	s2MarkBodyDirty(world, body);
	s2Transform xf = {body->origin, body->rot};
	for (int32_t si = body->shapeList; si != S2_NULL_INDEX;)
	{
		s2Shape* shape = world->shapes + si;
		si = shape->nextShapeIndex;
		shape->aabb = s2Shape_ComputeAABB(shape, xf);
		shape->aabb.lowerBound.x -= s2_speculativeDistance;
		shape->aabb.lowerBound.y -= s2_speculativeDistance;
		shape->aabb.upperBound.x += s2_speculativeDistance;
		shape->aabb.upperBound.y += s2_speculativeDistance;
		float margin = body->type == s2_staticBody ? s2_speculativeDistance : s2_aabbMargin + s2_speculativeDistance;
		shape->fatAABB.lowerBound.x = shape->aabb.lowerBound.x - margin;
		shape->fatAABB.lowerBound.y = shape->aabb.lowerBound.y - margin;
		shape->fatAABB.upperBound.x = shape->aabb.upperBound.x + margin;
		shape->fatAABB.upperBound.y = shape->aabb.upperBound.y + margin;
		shape->fresh = true; // re-buffer the proxy as moved
		s2MarkShapeDirty(world, shape);
	}
}

void s2Body_SetLinearVelocity(s2BodyId bodyId, s2Vec2 linearVelocity)
{
	s2World* world = s2GetWorldFromIndex(bodyId.world);
	s2SyncStateToHost(world);
	s2Body* body = world->bodies + bodyId.index;
	body->linearVelocity = linearVelocity;
	s2MarkBodyDirty(world, body);
}

void s2Body_SetAngularVelocity(s2BodyId bodyId, float angularVelocity)
{
	s2World* world = s2GetWorldFromIndex(bodyId.world);
	s2SyncStateToHost(world);
	s2Body* body = world->bodies + bodyId.index;
	body->angularVelocity = angularVelocity;
	s2MarkBodyDirty(world, body);
}

// Forces are host-authoritative between steps (the device zeroes them at the end of every step, reference
// src/world.c:275-276), so accumulating one needs no read-back: only the 16-byte force row travels.
void s2Body_ApplyForceToCenter(s2BodyId bodyId, s2Vec2 force)
{
	s2World* world = s2GetWorldFromIndex(bodyId.world);
	s2Body* body = s2GetBody(world, bodyId);
	// the host copy is zeroed whenever it is uploaded (s2FlushToDevice), mirroring the reset at the end of the step
	body->force = s2Add(body->force, force);
	s2MarkBodyForceDirty(world, body);
}

void s2Body_ApplyLinearImpulse(s2BodyId bodyId, s2Vec2 impulse, s2Vec2 point)
{
	s2World* world = s2GetWorldFromIndex(bodyId.world);
	s2SyncStateToHost(world);
	s2Body* body = s2GetBody(world, bodyId);
	if (body->type != s2_dynamicBody)
	{
		return;
	}
	body->linearVelocity = s2MulAdd(body->linearVelocity, body->invMass, impulse);
	body->angularVelocity += body->invI * s2Cross(s2Sub(point, body->position), impulse);
	s2MarkBodyDirty(world, body);
}

s2BodyType s2Body_GetType(s2BodyId bodyId)
{
	s2World* world = s2GetWorldFromIndex(bodyId.world);
	return world->bodies[bodyId.index].type;
}

float s2Body_GetMass(s2BodyId bodyId)
{
	s2World* world = s2GetWorldFromIndex(bodyId.world);
	return world->bodies[bodyId.index].mass;
}

// ---- shapes (reference src/shape.c:91-137) ------------------------------------------------------------------------

s2BodyId s2Shape_GetBody(s2ShapeId shapeId)
{
	s2World* world = s2GetWorldFromIndex(shapeId.world);
	s2Shape* shape = world->shapes + shapeId.index;
	s2Body* body = world->bodies + shape->bodyIndex;
	s2BodyId bodyId = {body->object.index, shapeId.world, body->object.revision};
	return bodyId;
}

bool s2Shape_TestPoint(s2ShapeId shapeId, s2Vec2 point)
{
	s2World* world = s2GetWorldFromIndex(shapeId.world);
	s2SyncStateToHost(world);
	s2Shape* shape = world->shapes + shapeId.index;
	s2Body* body = world->bodies + shape->bodyIndex;
	s2Transform xf = {body->origin, body->rot};
	s2Vec2 localPoint = s2InvTransformPoint(xf, point);
	switch (shape->type)
	{
		case s2_capsuleShape:
			return s2PointInCapsule(localPoint, &shape->capsule);
		case s2_circleShape:
			return s2PointInCircle(localPoint, &shape->circle);
		case s2_polygonShape:
			return s2PointInPolygon(localPoint, &shape->polygon);
		default:
			return false;
	}
}
