// TEST INFRASTRUCTURE — not part of the product.
//
// Stage taps for the *unmodified* reference (erincatto/solver2d) compiled from /root/reference by
// oracle/Makefile into oracle/_ref/libsolver2d_ref.so. This file is our own code: it only *includes* the
// reference's internal headers at build time (-I/root/reference/src) so tests can
//   (1) read the reference's internal state (bodies, shapes, contacts + manifolds, joints) as flat float/int
//       records, and
//   (2) run one reference step split at the stage boundaries of s2World_Step (reference src/world.c:120-306):
//       collide (stages 1-3, world.c:125-168) | solve (context + dispatch, world.c:171-256) | finalize (stage 4,
//       world.c:260-305), by calling the reference's own internal functions in the reference's order.
// tests/test_oracle_ref.py checks that the split step is bit-identical to the reference's own s2World_Step.
//
// Nothing here is linked into, imported by, or reachable from the product library.

#include "body.h"
#include "broad_phase.h"
#include "contact.h"
#include "core.h"
#include "joint.h"
#include "shape.h"
#include "solvers.h"
#include "stack_allocator.h"
#include "world.h"

#include "solver2d/aabb.h"
#include "solver2d/solver2d.h"

#include <string.h>

#define TAP_EXPORT __attribute__((visibility("default")))

// ---------------------------------------------------------------------------------------------------------------
// record layouts (floats unless noted; ints are stored bit-cast into the float slots via memcpy on the Python side
// we instead use two parallel arrays: one float array and one int array per dump)
// ---------------------------------------------------------------------------------------------------------------

enum
{
	TAP_BODY_F = 28, // floats per body
	TAP_BODY_I = 4,	 // ints per body
	TAP_SHAPE_F = 12,
	TAP_SHAPE_I = 8,
	TAP_CONTACT_F = 44,
	TAP_CONTACT_I = 16,
	TAP_JOINT_F = 40,
	TAP_JOINT_I = 8,
};

TAP_EXPORT void s2ref_record_sizes(int* out)
{
	out[0] = TAP_BODY_F;
	out[1] = TAP_BODY_I;
	out[2] = TAP_SHAPE_F;
	out[3] = TAP_SHAPE_I;
	out[4] = TAP_CONTACT_F;
	out[5] = TAP_CONTACT_I;
	out[6] = TAP_JOINT_F;
	out[7] = TAP_JOINT_I;
}

TAP_EXPORT void s2ref_capacities(int worldIndex, int* out)
{
	s2World* w = s2GetWorldFromIndex((int16_t)worldIndex);
	out[0] = w->bodyPool.capacity;
	out[1] = w->shapePool.capacity;
	out[2] = w->contactPool.capacity;
	out[3] = w->jointPool.capacity;
	out[4] = w->bodyPool.count;
	out[5] = w->shapePool.count;
	out[6] = w->contactPool.count;
	out[7] = w->jointPool.count;
}

TAP_EXPORT void s2ref_dump_bodies(int worldIndex, float* f, int* n)
{
	s2World* w = s2GetWorldFromIndex((int16_t)worldIndex);
	int cap = w->bodyPool.capacity;
	for (int i = 0; i < cap; ++i)
	{
		const s2Body* b = w->bodies + i;
		float* o = f + i * TAP_BODY_F;
		int* k = n + i * TAP_BODY_I;
		k[0] = s2IsFree(&b->object) ? 0 : 1;
		k[1] = (int)b->type;
		k[2] = b->object.revision;
		k[3] = b->shapeList;
		o[0] = b->origin.x;
		o[1] = b->origin.y;
		o[2] = b->position.x;
		o[3] = b->position.y;
		o[4] = b->rot.s;
		o[5] = b->rot.c;
		o[6] = b->linearVelocity.x;
		o[7] = b->linearVelocity.y;
		o[8] = b->angularVelocity;
		o[9] = b->deltaPosition.x;
		o[10] = b->deltaPosition.y;
		o[11] = b->localCenter.x;
		o[12] = b->localCenter.y;
		o[13] = b->mass;
		o[14] = b->invMass;
		o[15] = b->I;
		o[16] = b->invI;
		o[17] = b->force.x;
		o[18] = b->force.y;
		o[19] = b->torque;
		o[20] = b->linearDamping;
		o[21] = b->angularDamping;
		o[22] = b->gravityScale;
		o[23] = b->rot0.s;
		o[24] = b->rot0.c;
		o[25] = b->deltaPosition0.x;
		o[26] = b->deltaPosition0.y;
		o[27] = 0.0f;
	}
}

// Overwrite the dynamic state of every live body (used to start the reference solver from an arbitrary state).
TAP_EXPORT void s2ref_load_body_state(int worldIndex, const float* f)
{
	s2World* w = s2GetWorldFromIndex((int16_t)worldIndex);
	int cap = w->bodyPool.capacity;
	for (int i = 0; i < cap; ++i)
	{
		s2Body* b = w->bodies + i;
		if (s2IsFree(&b->object))
		{
			continue;
		}
		const float* o = f + i * TAP_BODY_F;
		b->origin = (s2Vec2){o[0], o[1]};
		b->position = (s2Vec2){o[2], o[3]};
		b->rot = (s2Rot){o[4], o[5]};
		b->linearVelocity = (s2Vec2){o[6], o[7]};
		b->angularVelocity = o[8];
		b->deltaPosition = (s2Vec2){o[9], o[10]};
		b->force = (s2Vec2){o[17], o[18]};
		b->torque = o[19];
	}
}

TAP_EXPORT void s2ref_dump_shapes(int worldIndex, float* f, int* n)
{
	s2World* w = s2GetWorldFromIndex((int16_t)worldIndex);
	int cap = w->shapePool.capacity;
	for (int i = 0; i < cap; ++i)
	{
		const s2Shape* s = w->shapes + i;
		float* o = f + i * TAP_SHAPE_F;
		int* k = n + i * TAP_SHAPE_I;
		k[0] = s2IsFree(&s->object) ? 0 : 1;
		k[1] = s->bodyIndex;
		k[2] = (int)s->type;
		k[3] = s->proxyKey;
		k[4] = s->nextShapeIndex;
		k[5] = s->object.revision;
		k[6] = 0;
		k[7] = 0;
		o[0] = s->aabb.lowerBound.x;
		o[1] = s->aabb.lowerBound.y;
		o[2] = s->aabb.upperBound.x;
		o[3] = s->aabb.upperBound.y;
		o[4] = s->fatAABB.lowerBound.x;
		o[5] = s->fatAABB.lowerBound.y;
		o[6] = s->fatAABB.upperBound.x;
		o[7] = s->fatAABB.upperBound.y;
		o[8] = s->friction;
		o[9] = s->density;
		o[10] = s->restitution;
		o[11] = 0.0f;
	}
}

TAP_EXPORT void s2ref_dump_contacts(int worldIndex, float* f, int* n)
{
	s2World* w = s2GetWorldFromIndex((int16_t)worldIndex);
	int cap = w->contactPool.capacity;
	for (int i = 0; i < cap; ++i)
	{
		const s2Contact* c = w->contacts + i;
		float* o = f + i * TAP_CONTACT_F;
		int* k = n + i * TAP_CONTACT_I;
		memset(o, 0, sizeof(float) * TAP_CONTACT_F);
		memset(k, 0, sizeof(int) * TAP_CONTACT_I);
		if (s2IsFree(&c->object))
		{
			continue;
		}
		const s2Manifold* m = &c->manifold;
		k[0] = 1;
		k[1] = c->shapeIndexA;
		k[2] = c->shapeIndexB;
		k[3] = c->edges[0].bodyIndex;
		k[4] = c->edges[1].bodyIndex;
		k[5] = m->pointCount;
		k[6] = m->points[0].id;
		k[7] = m->points[1].id;
		k[8] = m->points[0].persisted ? 1 : 0;
		k[9] = m->points[1].persisted ? 1 : 0;
		k[10] = m->frictionPersisted ? 1 : 0;
		k[11] = c->cache.count;
		k[12] = c->cache.indexA[0] | (c->cache.indexA[1] << 8) | (c->cache.indexA[2] << 16);
		k[13] = c->cache.indexB[0] | (c->cache.indexB[1] << 8) | (c->cache.indexB[2] << 16);
		o[0] = c->friction;
		o[1] = m->normal.x;
		o[2] = m->normal.y;
		o[3] = c->cache.metric;
		for (int j = 0; j < 2; ++j)
		{
			const s2ManifoldPoint* p = m->points + j;
			float* q = o + 4 + 20 * j;
			q[0] = p->localAnchorA.x;
			q[1] = p->localAnchorA.y;
			q[2] = p->localAnchorB.x;
			q[3] = p->localAnchorB.y;
			q[4] = p->separation;
			q[5] = p->normalImpulse;
			q[6] = p->tangentImpulse;
			q[7] = p->frictionAnchorA.x;
			q[8] = p->frictionAnchorA.y;
			q[9] = p->frictionAnchorB.x;
			q[10] = p->frictionAnchorB.y;
			q[11] = p->frictionNormalA.x;
			q[12] = p->frictionNormalA.y;
			q[13] = p->frictionNormalB.x;
			q[14] = p->frictionNormalB.y;
		}
	}
}

TAP_EXPORT void s2ref_dump_joints(int worldIndex, float* f, int* n)
{
	s2World* w = s2GetWorldFromIndex((int16_t)worldIndex);
	int cap = w->jointPool.capacity;
	for (int i = 0; i < cap; ++i)
	{
		const s2Joint* j = w->joints + i;
		float* o = f + i * TAP_JOINT_F;
		int* k = n + i * TAP_JOINT_I;
		memset(o, 0, sizeof(float) * TAP_JOINT_F);
		memset(k, 0, sizeof(int) * TAP_JOINT_I);
		if (s2IsFree(&j->object))
		{
			continue;
		}
		k[0] = 1;
		k[1] = (int)j->type;
		k[2] = j->edges[0].bodyIndex;
		k[3] = j->edges[1].bodyIndex;
		k[4] = j->collideConnected ? 1 : 0;
		o[0] = j->localOriginAnchorA.x;
		o[1] = j->localOriginAnchorA.y;
		o[2] = j->localOriginAnchorB.x;
		o[3] = j->localOriginAnchorB.y;
		if (j->type == s2_revoluteJoint)
		{
			const s2RevoluteJoint* r = &j->revoluteJoint;
			k[5] = r->enableMotor ? 1 : 0;
			k[6] = r->enableLimit ? 1 : 0;
			o[4] = r->impulse.x;
			o[5] = r->impulse.y;
			o[6] = r->motorImpulse;
			o[7] = r->lowerImpulse;
			o[8] = r->upperImpulse;
			o[9] = r->maxMotorTorque;
			o[10] = r->motorSpeed;
			o[11] = r->referenceAngle;
			o[12] = r->lowerAngle;
			o[13] = r->upperAngle;
		}
		else
		{
			const s2MouseJoint* m = &j->mouseJoint;
			o[4] = m->impulse.x;
			o[5] = m->impulse.y;
			o[6] = m->motorImpulse;
			o[14] = m->hertz;
			o[15] = m->dampingRatio;
			o[16] = m->targetA.x;
			o[17] = m->targetA.y;
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// split step
// ---------------------------------------------------------------------------------------------------------------

// Stages 1-3 of s2World_Step (reference src/world.c:123-168): pair update, tree rebuild, contact update/destroy.
TAP_EXPORT void s2ref_step_collide(int worldIndex)
{
	s2World* world = s2GetWorldFromIndex((int16_t)worldIndex);
	world->stepId += 1;

	s2UpdateBroadPhasePairs(world);
	s2BroadPhase_RebuildTrees(&world->broadPhase);

	int contactCapacity = world->contactPool.capacity;
	for (int i = 0; i < contactCapacity; ++i)
	{
		s2Contact* contact = world->contacts + i;
		if (s2IsFree(&contact->object))
		{
			continue;
		}
		s2Shape* shapeA = world->shapes + contact->shapeIndexA;
		s2Shape* shapeB = world->shapes + contact->shapeIndexB;
		if (s2AABB_Overlaps(shapeA->fatAABB, shapeB->fatAABB))
		{
			s2UpdateContact(world, contact, shapeA, world->bodies + shapeA->bodyIndex, shapeB, world->bodies + shapeB->bodyIndex);
		}
		else
		{
			s2DestroyContact(world, contact);
		}
	}
}

// Step context + solver dispatch of s2World_Step (reference src/world.c:171-256).
TAP_EXPORT void s2ref_step_solve(int worldIndex, float timeStep, int velIters, int posIters, int warmStart)
{
	s2World* world = s2GetWorldFromIndex((int16_t)worldIndex);

	s2StepContext context = {0};
	context.dt = timeStep;
	context.iterations = velIters;
	context.extraIterations = posIters;
	context.warmStart = warmStart != 0;
	context.inv_dt = timeStep > 0.0f ? 1.0f / timeStep : 0.0f;

	s2SolverType type = world->solverType;
	bool substepping = type == s2_solverXPBD || type == s2_solverTGS_Soft || type == s2_solverTGS_Sticky ||
					   type == s2_solverTGS_NGS || type == s2_solverSoftStep;
	if (substepping)
	{
		context.h = context.dt / context.iterations;
		context.inv_h = context.inv_dt * context.iterations;
	}
	else
	{
		context.h = context.dt;
		context.inv_h = context.inv_dt;
	}
	context.bodies = world->bodies;
	context.bodyCapacity = world->bodyPool.capacity;

	switch (type)
	{
		case s2_solverJacobi:
			s2Solve_Jacobi(world, &context);
			break;
		case s2_solverPGS:
			s2Solve_PGS(world, &context);
			break;
		case s2_solverPGS_NGS:
			s2Solve_PGS_NGS(world, &context);
			break;
		case s2_solverPGS_NGS_Block:
			s2Solve_PGS_NGS_Block(world, &context);
			break;
		case s2_solverPGS_Soft:
			s2Solve_PGS_Soft(world, &context);
			break;
		case s2_solverTGS_Sticky:
			s2Solve_TGS_Sticky(world, &context);
			break;
		case s2_solverTGS_Soft:
			s2Solve_TGS_Soft(world, &context);
			break;
		case s2_solverTGS_NGS:
			s2Solve_TGS_NGS(world, &context);
			break;
		case s2_solverXPBD:
			s2Solve_XPBD(world, &context);
			break;
		case s2_solverSoftStep:
			s2Solve_SoftStep(world, &context);
			break;
		default:
			break;
	}
}

// Stage 4 of s2World_Step (reference src/world.c:258-305): transforms, force reset, AABB refit / proxy enlarge.
TAP_EXPORT void s2ref_step_finalize(int worldIndex)
{
	s2World* world = s2GetWorldFromIndex((int16_t)worldIndex);
	s2BroadPhase* broadPhase = &world->broadPhase;
	const s2Vec2 margin = {s2_aabbMargin, s2_aabbMargin};
	int bodyCapacity = world->bodyPool.capacity;
	for (int i = 0; i < bodyCapacity; ++i)
	{
		s2Body* body = world->bodies + i;
		if (s2IsFree(&body->object) || body->type == s2_staticBody)
		{
			continue;
		}
		body->origin = s2Sub(body->position, s2RotateVector(body->rot, body->localCenter));
		body->force = s2Vec2_zero;
		body->torque = 0.0f;
		s2Transform xf = {body->origin, body->rot};
		for (int si = body->shapeList; si != S2_NULL_INDEX;)
		{
			s2Shape* shape = world->shapes + si;
			shape->aabb = s2Shape_ComputeAABB(shape, xf);
			shape->aabb.lowerBound.x -= s2_speculativeDistance;
			shape->aabb.lowerBound.y -= s2_speculativeDistance;
			shape->aabb.upperBound.x += s2_speculativeDistance;
			shape->aabb.upperBound.y += s2_speculativeDistance;
			if (s2AABB_Contains(shape->fatAABB, shape->aabb) == false)
			{
				shape->fatAABB.lowerBound = s2Sub(shape->aabb.lowerBound, margin);
				shape->fatAABB.upperBound = s2Add(shape->aabb.upperBound, margin);
				s2BroadPhase_EnlargeProxy(broadPhase, shape->proxyKey, shape->fatAABB);
			}
			si = shape->nextShapeIndex;
		}
	}
	s2GrowStack(world->stackAllocator);
}

// Overwrite the warm-start impulses stored in one contact's manifold (for injected-state solver tests).
TAP_EXPORT void s2ref_set_contact_impulses(int worldIndex, int contactIndex, float n0, float t0, float n1, float t1)
{
	s2World* w = s2GetWorldFromIndex((int16_t)worldIndex);
	s2Manifold* m = &w->contacts[contactIndex].manifold;
	m->points[0].normalImpulse = n0;
	m->points[0].tangentImpulse = t0;
	m->points[1].normalImpulse = n1;
	m->points[1].tangentImpulse = t1;
}

// Bulk forms of the impulse setters: what a solver stage leaves in the manifolds / joints (warm-start state of the next
// step), written from arrays indexed by pool slot. Used by the multi-step parity test, which lets the ORDER-PERMUTED
// oracle advance the reference world so that every step of the free-running device schedule can be checked bit for bit.
// f: contactCapacity x 4 floats {normalImpulse0, tangentImpulse0, normalImpulse1, tangentImpulse1}
TAP_EXPORT void s2ref_load_contact_impulses(int worldIndex, const float* f)
{
	s2World* w = s2GetWorldFromIndex((int16_t)worldIndex);
	int cap = w->contactPool.capacity;
	for (int i = 0; i < cap; ++i)
	{
		s2Contact* c = w->contacts + i;
		if (s2IsFree(&c->object))
		{
			continue;
		}
		c->manifold.points[0].normalImpulse = f[4 * i + 0];
		c->manifold.points[0].tangentImpulse = f[4 * i + 1];
		c->manifold.points[1].normalImpulse = f[4 * i + 2];
		c->manifold.points[1].tangentImpulse = f[4 * i + 3];
	}
}

// f: jointCapacity x 5 floats {impulse.x, impulse.y, motorImpulse, lowerImpulse, upperImpulse}
TAP_EXPORT void s2ref_load_joint_impulses(int worldIndex, const float* f)
{
	s2World* w = s2GetWorldFromIndex((int16_t)worldIndex);
	int cap = w->jointPool.capacity;
	for (int i = 0; i < cap; ++i)
	{
		s2Joint* j = w->joints + i;
		if (s2IsFree(&j->object))
		{
			continue;
		}
		const float* o = f + 5 * i;
		if (j->type == s2_revoluteJoint)
		{
			j->revoluteJoint.impulse = (s2Vec2){o[0], o[1]};
			j->revoluteJoint.motorImpulse = o[2];
			j->revoluteJoint.lowerImpulse = o[3];
			j->revoluteJoint.upperImpulse = o[4];
		}
		else
		{
			j->mouseJoint.impulse = (s2Vec2){o[0], o[1]};
			j->mouseJoint.motorImpulse = o[2];
		}
	}
}

// Time N reference steps with CLOCK_MONOTONIC inside C (bench.py's reference arm; avoids ctypes overhead in the loop).
#include <time.h>
TAP_EXPORT double s2ref_timed_steps(s2WorldId worldId, int steps, float timeStep, int velIters, int posIters, int warmStart)
{
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int i = 0; i < steps; ++i)
	{
		s2World_Step(worldId, timeStep, velIters, posIters, warmStart != 0);
	}
	clock_gettime(CLOCK_MONOTONIC, &t1);
	return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

// The reference-side "end to end" step used by bench.py's reference arm: apply a force to every listed body through the
// reference's own API, step, read every transform back — the same client pattern the product's e2e number measures.
TAP_EXPORT double s2ref_timed_e2e_steps(s2WorldId worldId, int steps, float timeStep, int velIters, int posIters, int warmStart,
										const int* bodyIndices, const float* forcesXY, int forceCount, float* transforms)
{
	s2World* world = s2GetWorldFromId(worldId);
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int s = 0; s < steps; ++s)
	{
		float sign = (s & 1) ? -1.0f : 1.0f;
		for (int i = 0; i < forceCount; ++i)
		{
			s2Body* b = world->bodies + bodyIndices[i];
			s2BodyId id = {b->object.index, world->index, b->object.revision};
			s2Body_ApplyForceToCenter(id, (s2Vec2){sign * forcesXY[2 * i], sign * forcesXY[2 * i + 1]});
		}
		s2World_Step(worldId, timeStep, velIters, posIters, warmStart != 0);
		int cap = world->bodyPool.capacity;
		for (int i = 0; i < cap; ++i)
		{
			const s2Body* b = world->bodies + i;
			transforms[4 * i + 0] = b->origin.x;
			transforms[4 * i + 1] = b->origin.y;
			transforms[4 * i + 2] = b->rot.s;
			transforms[4 * i + 3] = b->rot.c;
		}
	}
	clock_gettime(CLOCK_MONOTONIC, &t1);
	return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

// constraint count of the next solve: manifolds with >= 1 point (what every s2Solve_* gathers, e.g. reference
// src/solve_tgs_soft.c:162-179) + live joints
TAP_EXPORT void s2ref_constraint_counts(int worldIndex, int* out)
{
	s2World* w = s2GetWorldFromIndex((int16_t)worldIndex);
	int manifolds = 0;
	for (int i = 0; i < w->contactPool.capacity; ++i)
	{
		const s2Contact* c = w->contacts + i;
		if (s2IsFree(&c->object) == false && c->manifold.pointCount > 0)
		{
			manifolds += 1;
		}
	}
	out[0] = manifolds;
	out[1] = w->jointPool.count;
}
