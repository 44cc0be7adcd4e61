// solver2d-b200 — joint definitions (ABI of reference include/solver2d/joint_types.h:9-97).
#pragma once

#include "solver2d/types.h"

// Soft point constraint dragging body B towards a world-space target (body A is only a bookkeeping anchor).
typedef struct s2MouseJointDef
{
	s2BodyId bodyIdA;
	s2BodyId bodyIdB;
	s2Vec2 target;		// initial world target
	float hertz;		// stiffness
	float dampingRatio; // non-dimensional
} s2MouseJointDef;

S2_INLINE struct s2MouseJointDef s2DefaultMouseJointDef(void)
{
	s2MouseJointDef def = S2_ZERO_INIT;
	def.bodyIdA = s2_nullBodyId;
	def.bodyIdB = s2_nullBodyId;
	def.hertz = 15.0f;
	def.dampingRatio = 1.0f;
	return def;
}

// Pin joint with optional angular limit and motor. Anchors are relative to each body's origin.
typedef struct s2RevoluteJointDef
{
	s2BodyId bodyIdA;
	s2BodyId bodyIdB;
	s2Vec2 localAnchorA;
	s2Vec2 localAnchorB;
	float referenceAngle; // angle(B) - angle(A) that counts as zero for the limit
	bool enableLimit;
	float lowerAngle;
	float upperAngle;
	bool enableMotor;
	float motorSpeed;	  // rad/s
	float maxMotorTorque; // N-m
	float drawSize;
	bool collideConnected; // keep contacts between the two bodies
} s2RevoluteJointDef;

S2_INLINE struct s2RevoluteJointDef s2DefaultRevoluteJointDef(void)
{
	s2RevoluteJointDef def = S2_ZERO_INIT;
	def.bodyIdA = s2_nullBodyId;
	def.bodyIdB = s2_nullBodyId;
	def.drawSize = 1.0f;
	return def;
}
