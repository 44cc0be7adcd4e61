// solver2d-b200 — joints behind the public API (behaviour of reference src/joint.c:154-292, src/revolute_joint.c:825-888,
// src/mouse_joint.c:18-29). Host bookkeeping only: the joint solvers run on the device.
#include "s2_host.h"

#include <string.h>

static s2Joint* s2AllocJoint(s2World* world, s2Body* bodyA, s2Body* bodyB)
{
	s2Joint* joint = (s2Joint*)s2AllocObject(&world->jointPool);
	world->joints = (s2Joint*)world->jointPool.memory;
	bool rowDirty = joint->rowDirty;
	s2Object object = joint->object;
	memset(joint, 0, sizeof(*joint));
	joint->object = object;
	joint->rowDirty = rowDirty; // a reused slot may already sit in the dirty list
	joint->bodyIndexA = bodyA->object.index;
	joint->bodyIndexB = bodyB->object.index;
	bodyA->jointCount += 1;
	bodyB->jointCount += 1;
	world->jointPairsDirty = true;
	s2MarkJointDirty(world, joint);
	return joint;
}

// reference src/joint.c:154-179
s2JointId s2CreateMouseJoint(s2WorldId worldId, const s2MouseJointDef* def)
{
	s2World* world = s2GetWorldFromId(worldId);
	s2SyncStateToHost(world);
	s2Body* bodyA = world->bodies + def->bodyIdA.index;
	s2Body* bodyB = world->bodies + def->bodyIdB.index;

	s2Joint* joint = s2AllocJoint(world, bodyA, bodyB);
	joint->type = s2_mouseJoint;
	joint->drawSize = 1.0f;
	s2Transform xfA = {bodyA->origin, bodyA->rot};
	s2Transform xfB = {bodyB->origin, bodyB->rot};
	joint->localOriginAnchorA = s2InvTransformPoint(xfA, def->target);
	joint->localOriginAnchorB = s2InvTransformPoint(xfB, def->target);
	joint->targetA = def->target;
	joint->hertz = def->hertz;
	joint->dampingRatio = def->dampingRatio;
	// the reference never initialises collideConnected for mouse joints and never removes existing contacts for them
	joint->collideConnected = true;

	s2JointId jointId = {joint->object.index, world->index, joint->object.revision};
	return jointId;
}

// reference src/joint.c:181-223
s2JointId s2CreateRevoluteJoint(s2WorldId worldId, const s2RevoluteJointDef* def)
{
	s2World* world = s2GetWorldFromId(worldId);
	s2Body* bodyA = world->bodies + def->bodyIdA.index;
	s2Body* bodyB = world->bodies + def->bodyIdB.index;

	s2Joint* joint = s2AllocJoint(world, bodyA, bodyB);
	joint->type = s2_revoluteJoint;
	joint->drawSize = def->drawSize;
	joint->localOriginAnchorA = def->localAnchorA;
	joint->localOriginAnchorB = def->localAnchorB;
	joint->referenceAngle = def->referenceAngle;
	joint->lowerAngle = def->lowerAngle;
	joint->upperAngle = def->upperAngle;
	joint->maxMotorTorque = def->maxMotorTorque;
	joint->motorSpeed = def->motorSpeed;
	joint->enableLimit = def->enableLimit;
	joint->enableMotor = def->enableMotor;
	// collideConnected == false removes the contacts that already exist between the two bodies (the pair pass does it on
	// the device); new contacts between jointed bodies are never created either way
	joint->collideConnected = def->collideConnected;

	s2JointId jointId = {joint->object.index, world->index, joint->object.revision};
	return jointId;
}

// reference src/joint.c:225-292
void s2DestroyJoint(s2JointId jointId)
{
	s2World* world = s2GetWorldFromIndex(jointId.world);
	s2Joint* joint = world->joints + jointId.index;
	world->bodies[joint->bodyIndexA].jointCount -= 1;
	world->bodies[joint->bodyIndexB].jointCount -= 1;
	s2FreeObject(&world->jointPool, &joint->object);
	s2MarkJointDirty(world, joint);
	world->jointPairsDirty = true;
}

static s2Joint* s2GetJointForEdit(s2JointId jointId, s2World** worldOut)
{
	s2World* world = s2GetWorldFromIndex(jointId.world);
	// the row carries the accumulated impulses: make the host copy current before it is re-uploaded
	s2SyncStateToHost(world);
	s2Joint* joint = world->joints + jointId.index;
	S2_ASSERT(s2ObjectValid(&joint->object));
	S2_ASSERT(joint->object.revision == jointId.revision);
	*worldOut = world;
	return joint;
}

void s2MouseJoint_SetTarget(s2JointId jointId, s2Vec2 target)
{
	s2World* world;
	s2Joint* joint = s2GetJointForEdit(jointId, &world);
	joint->targetA = target;
	s2MarkJointDirty(world, joint);
}

void s2RevoluteJoint_EnableLimit(s2JointId jointId, bool enableLimit)
{
	s2World* world;
	s2Joint* joint = s2GetJointForEdit(jointId, &world);
	joint->enableLimit = enableLimit;
	s2MarkJointDirty(world, joint);
}

void s2RevoluteJoint_EnableMotor(s2JointId jointId, bool enableMotor)
{
	s2World* world;
	s2Joint* joint = s2GetJointForEdit(jointId, &world);
	joint->enableMotor = enableMotor;
	s2MarkJointDirty(world, joint);
}

void s2RevoluteJoint_SetMotorSpeed(s2JointId jointId, float motorSpeed)
{
	s2World* world;
	s2Joint* joint = s2GetJointForEdit(jointId, &world);
	joint->motorSpeed = motorSpeed;
	s2MarkJointDirty(world, joint);
}

float s2RevoluteJoint_GetMotorTorque(s2JointId jointId, float inverseTimeStep)
{
	s2World* world = s2GetWorldFromIndex(jointId.world);
	s2SyncStateToHost(world);
	return inverseTimeStep * world->joints[jointId.index].motorImpulse;
}
