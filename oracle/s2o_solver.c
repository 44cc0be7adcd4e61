// TEST INFRASTRUCTURE — the oracle. Not part of the product: only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline leg may load this library; the product never links, imports or calls it.
//
// Plain-C, sequential, array-of-structs restatement of the solver stage of solver2d's s2World_Step
// (reference src/solve_*.c, src/solve_common.c, src/revolute_joint.c, src/mouse_joint.c), with ONE addition the
// reference does not have: the Gauss-Seidel visiting order of the constraints is a parameter (`order`).
//   * order == NULL        : joints in slot order, then contact constraints in row order — the reference's own order.
//                            In this mode the oracle is PINNED: tests/test_oracle_cpu.py checks it bit for bit against the
//                            unmodified reference (oracle/_ref) for every variant it restates.
//   * order == device order: the colour-major order the GPU schedule used (s2b_download_solve_order); the GPU result
//                            must then match this oracle bit for bit, which is how the production schedule is verified.
// It shares the row structs of include/s2b_device.h (types only) and the float inlines of include/solver2d/math.h; it is
// compiled with the reference's flags (gcc -std=gnu17 -O2, no FMA contraction).
//
// Restated: all ten variants s2World_Step dispatches (reference src/world.c:206-256) — s2Solve_PGS, _PGS_NGS,
// _PGS_NGS_Block, _PGS_Soft, _TGS_Soft, _TGS_NGS, _TGS_Sticky, _SoftStep, _XPBD, _Jacobi — with revolute and mouse joints.

#include "s2b_device.h"
#include "solver2d/constants.h"
#include "solver2d/math.h"

#include <stdlib.h>
#include <string.h>

#define S2O_EXPORT __attribute__((visibility("default")))

typedef struct s2oBody
{
	int valid, type;
	s2Vec2 position, dp, v, localCenter, force;
	s2Vec2 dv, dp0; // Jacobi accumulator; XPBD previous delta position
	float dw;
	s2Rot q, q0;
	float w, mass, invMass, I, invI, torque, linearDamping, angularDamping, gravityScale;
} s2oBody;

typedef struct s2oPoint
{
	s2Vec2 localAnchorA, localAnchorB; // relative to the centres of mass, body frames
	s2Vec2 rA0, rB0;				   // world anchors at prepare time
	s2Vec2 localFrictionAnchorA, localFrictionAnchorB;
	float separation, tangentSeparation;
	float adjustedSeparation, normalImpulse, tangentImpulse, normalMass, tangentMass;
	float biasCoefficient, massCoefficient, impulseCoefficient;
} s2oPoint;

typedef struct s2oConstraint
{
	int row, indexA, indexB, pointCount;
	int unpersist; // TGS_Sticky: friction hit its limit in some pass
	// PGS_NGS_Block
	s2Mat22 K, normalMass;
	float velocityBias[2];
	s2Vec2 normal;
	float friction;
	s2oPoint points[2];
} s2oConstraint;

typedef struct s2oJoint
{
	int slot, type, indexA, indexB, enableLimit, enableMotor;
	s2Vec2 localAnchorA, localAnchorB, centerDiff0, impulse;
	float invMassA, invMassB, invIA, invIB, axialMass;
	float biasCoefficient, massCoefficient, impulseCoefficient;
	float referenceAngle, lowerAngle, upperAngle, maxMotorTorque, motorSpeed;
	float motorImpulse, lowerImpulse, upperImpulse;
	float bodyBInertia;
	s2Mat22 pivotMass;
} s2oJoint;

// ---- bodies ------------------------------------------------------------------------------------------------------

// s2IntegrateVelocities (reference src/solve_common.c:10-45)
static void s2oIntegrateVelocities(s2oBody* bodies, int count, s2Vec2 gravity, float h)
{
	for (int i = 0; i < count; ++i)
	{
		s2oBody* b = bodies + i;
		if (b->valid == 0 || b->type != s2_dynamicBody)
		{
			continue;
		}
		s2Vec2 v = b->v;
		float w = b->w;
		v = s2Add(v, s2MulSV(h * b->invMass, s2MulAdd(b->force, b->mass * b->gravityScale, gravity)));
		w = w + h * b->invI * b->torque;
		v = s2MulSV(1.0f / (1.0f + h * b->linearDamping), v);
		w *= 1.0f / (1.0f + h * b->angularDamping);
		b->v = v;
		b->w = w;
	}
}

// s2IntegratePositions (reference src/solve_common.c:47-68)
static void s2oIntegratePositions(s2oBody* bodies, int count, float h)
{
	for (int i = 0; i < count; ++i)
	{
		s2oBody* b = bodies + i;
		if (b->valid == 0 || b->type == s2_staticBody)
		{
			continue;
		}
		b->dp = s2MulAdd(b->dp, h, b->v);
		b->q = s2IntegrateRot(b->q, h * b->w);
	}
}

// s2FinalizePositions (reference src/solve_common.c:70-91)
static void s2oFinalizePositions(s2oBody* bodies, int count)
{
	for (int i = 0; i < count; ++i)
	{
		s2oBody* b = bodies + i;
		if (b->valid == 0 || b->type == s2_staticBody)
		{
			continue;
		}
		b->position = s2Add(b->position, b->dp);
		b->dp = s2Vec2_zero;
	}
}

// ---- contacts ----------------------------------------------------------------------------------------------------

// s2PrepareContacts_Soft, one constraint (reference src/solve_common.c:188-274)
static void s2oPrepareContactSoft(s2oConstraint* c, const s2bContactRow* row, int rowIndex, const s2oBody* bodies, int warmStart,
								  float h, float hertz)
{
	c->row = rowIndex;
	c->indexA = row->bodyA;
	c->indexB = row->bodyB;
	c->normal = s2MakeVec2(row->normal[0], row->normal[1]);
	c->friction = row->friction;
	c->pointCount = row->pointCount;

	const s2oBody* bodyA = bodies + c->indexA;
	const s2oBody* bodyB = bodies + c->indexB;
	float mA = bodyA->invMass, iA = bodyA->invI, mB = bodyB->invMass, iB = bodyB->invI;
	float contactHertz = (mA == 0.0f || mB == 0.0f) ? 2.0f * hertz : hertz;
	s2Rot qA = bodyA->q, qB = bodyB->q;
	s2Vec2 normal = c->normal;
	s2Vec2 tangent = s2RightPerp(normal);

	for (int j = 0; j < c->pointCount; ++j)
	{
		const s2bContactPoint* mp = row->points + j;
		s2oPoint* cp = c->points + j;
		cp->normalImpulse = warmStart ? mp->normalImpulse : 0.0f;
		cp->tangentImpulse = warmStart ? mp->tangentImpulse : 0.0f;
		cp->localAnchorA = s2Sub(s2MakeVec2(mp->localAnchorA[0], mp->localAnchorA[1]), bodyA->localCenter);
		cp->localAnchorB = s2Sub(s2MakeVec2(mp->localAnchorB[0], mp->localAnchorB[1]), bodyB->localCenter);
		s2Vec2 rA = s2RotateVector(qA, cp->localAnchorA);
		s2Vec2 rB = s2RotateVector(qB, cp->localAnchorB);
		cp->rA0 = rA;
		cp->rB0 = rB;
		cp->separation = mp->separation;
		cp->adjustedSeparation = mp->separation - s2Dot(s2Sub(rB, rA), normal);

		float rnA = s2Cross(rA, normal);
		float rnB = s2Cross(rB, normal);
		float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
		cp->normalMass = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;

		float rtA = s2Cross(rA, tangent);
		float rtB = s2Cross(rB, tangent);
		float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
		cp->tangentMass = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;

		const float zeta = 10.0f;
		float omega = 2.0f * s2_pi * contactHertz;
		float cc = h * omega * (2.0f * zeta + h * omega);
		cp->biasCoefficient = omega / (2.0f * zeta + h * omega);
		cp->impulseCoefficient = 1.0f / (1.0f + cc);
		cp->massCoefficient = cc * cp->impulseCoefficient;
	}
}

// s2WarmStartContacts, one constraint (reference src/solve_common.c:276-326)
static void s2oWarmStartContact(s2oConstraint* c, s2oBody* bodies)
{
	s2oBody* bodyA = bodies + c->indexA;
	s2oBody* bodyB = bodies + c->indexB;
	float mA = bodyA->invMass, iA = bodyA->invI, mB = bodyB->invMass, iB = bodyB->invI;
	s2Vec2 vA = bodyA->v, vB = bodyB->v;
	float wA = bodyA->w, wB = bodyB->w;
	s2Rot qA = bodyA->q, qB = bodyB->q;
	s2Vec2 normal = c->normal;
	s2Vec2 tangent = s2RightPerp(normal);
	for (int j = 0; j < c->pointCount; ++j)
	{
		s2oPoint* cp = c->points + j;
		s2Vec2 rA = s2RotateVector(qA, cp->localAnchorA);
		s2Vec2 rB = s2RotateVector(qB, cp->localAnchorB);
		s2Vec2 P = s2Add(s2MulSV(cp->normalImpulse, normal), s2MulSV(cp->tangentImpulse, tangent));
		wA -= iA * s2Cross(rA, P);
		vA = s2MulAdd(vA, -mA, P);
		wB += iB * s2Cross(rB, P);
		vB = s2MulAdd(vB, mB, P);
	}
	bodyA->v = vA;
	bodyA->w = wA;
	bodyB->v = vB;
	bodyB->w = wB;
}

// s2SolveContacts_TGS_Soft, one constraint (reference src/solve_tgs_soft.c:17-135)
static void s2oSolveContactTgsSoft(s2oConstraint* c, s2oBody* bodies, float inv_h, int useBias)
{
	s2oBody* bodyA = bodies + c->indexA;
	s2oBody* bodyB = bodies + c->indexB;
	float mA = bodyA->invMass, iA = bodyA->invI, mB = bodyB->invMass, iB = bodyB->invI;
	s2Vec2 vA = bodyA->v, vB = bodyB->v;
	float wA = bodyA->w, wB = bodyB->w;
	s2Vec2 dcA = bodyA->dp, dcB = bodyB->dp;
	s2Rot qA = bodyA->q, qB = bodyB->q;
	s2Vec2 normal = c->normal;
	s2Vec2 tangent = s2RightPerp(normal);
	float friction = c->friction;

	for (int j = 0; j < c->pointCount; ++j)
	{
		s2oPoint* cp = c->points + j;
		s2Vec2 rA = s2RotateVector(qA, cp->localAnchorA);
		s2Vec2 rB = s2RotateVector(qB, cp->localAnchorB);
		s2Vec2 ds = s2Add(s2Sub(dcB, dcA), s2Sub(rB, rA));
		float s = s2Dot(ds, normal) + cp->adjustedSeparation;

		float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
		if (s > 0.0f)
		{
			bias = s * inv_h;
		}
		else if (useBias)
		{
			bias = S2_MAX(cp->biasCoefficient * s, -s2_maxBaumgarteVelocity);
			massScale = cp->massCoefficient;
			impulseScale = cp->impulseCoefficient;
		}

		s2Vec2 vrB = s2Add(vB, s2CrossSV(wB, rB));
		s2Vec2 vrA = s2Add(vA, s2CrossSV(wA, rA));
		float vn = s2Dot(s2Sub(vrB, vrA), normal);
		float impulse = -cp->normalMass * massScale * (vn + bias) - impulseScale * cp->normalImpulse;
		float newImpulse = S2_MAX(cp->normalImpulse + impulse, 0.0f);
		impulse = newImpulse - cp->normalImpulse;
		cp->normalImpulse = newImpulse;

		s2Vec2 P = s2MulSV(impulse, normal);
		vA = s2MulSub(vA, mA, P);
		wA -= iA * s2Cross(rA, P);
		vB = s2MulAdd(vB, mB, P);
		wB += iB * s2Cross(rB, P);
	}

	for (int j = 0; j < c->pointCount; ++j)
	{
		s2oPoint* cp = c->points + j;
		s2Vec2 rA = s2RotateVector(qA, cp->localAnchorA);
		s2Vec2 rB = s2RotateVector(qB, cp->localAnchorB);
		s2Vec2 vrB = s2Add(vB, s2CrossSV(wB, rB));
		s2Vec2 vrA = s2Add(vA, s2CrossSV(wA, rA));
		float vt = s2Dot(s2Sub(vrB, vrA), tangent);
		float impulse = -cp->tangentMass * vt;
		float maxFriction = friction * cp->normalImpulse;
		float newImpulse = S2_CLAMP(cp->tangentImpulse + impulse, -maxFriction, maxFriction);
		impulse = newImpulse - cp->tangentImpulse;
		cp->tangentImpulse = newImpulse;

		s2Vec2 P = s2MulSV(impulse, tangent);
		vA = s2MulSub(vA, mA, P);
		wA -= iA * s2Cross(rA, P);
		vB = s2MulAdd(vB, mB, P);
		wB += iB * s2Cross(rB, P);
	}

	bodyA->v = vA;
	bodyA->w = wA;
	bodyB->v = vB;
	bodyB->w = wB;
}

// ---- joints ------------------------------------------------------------------------------------------------------

static int s2oJointType(const s2bJointRow* row)
{
	return (row->flags >> 1) & 0x7;
}

// s2PrepareJoint_Soft (reference src/joint.c:365-381) -> s2PrepareRevolute_Soft (src/revolute_joint.c:421-506) /
// s2PrepareMouse (src/mouse_joint.c:31-83)
static void s2oPrepareJointSoft(s2oJoint* j, const s2bJointRow* row, const s2oBody* bodies, float contextH, float h, float hertz,
								int warmStart)
{
	memset(j, 0, sizeof(*j));
	j->slot = row->index;
	j->type = s2oJointType(row);
	j->indexA = row->bodyA;
	j->indexB = row->bodyB;
	j->enableLimit = (row->flags & S2B_JOINT_ENABLE_LIMIT) != 0;
	j->enableMotor = (row->flags & S2B_JOINT_ENABLE_MOTOR) != 0;
	j->impulse = s2MakeVec2(row->impulse[0], row->impulse[1]);
	j->motorImpulse = row->motorImpulse;
	j->lowerImpulse = row->lowerImpulse;
	j->upperImpulse = row->upperImpulse;
	const s2oBody* bodyA = bodies + j->indexA;
	const s2oBody* bodyB = bodies + j->indexB;

	if (j->type == S2B_JOINT_MOUSE)
	{
		float mB = bodyB->invMass, iB = bodyB->invI;
		j->localAnchorB = s2Sub(s2MakeVec2(row->localOriginAnchorB[0], row->localOriginAnchorB[1]), bodyB->localCenter);
		j->invMassB = mB;
		j->invIB = iB;
		j->bodyBInertia = bodyB->I;
		float zeta = row->dampingRatio;
		float omega = 2.0f * s2_pi * row->hertz;
		j->biasCoefficient = omega / (2.0f * zeta + contextH * omega);
		float c = contextH * omega * (2.0f * zeta + contextH * omega);
		j->impulseCoefficient = 1.0f / (1.0f + c);
		j->massCoefficient = c * j->impulseCoefficient;
		s2Vec2 rB = s2RotateVector(bodyB->q, j->localAnchorB);
		s2Mat22 K;
		K.cx.x = mB + iB * rB.y * rB.y;
		K.cx.y = -iB * rB.x * rB.y;
		K.cy.x = K.cx.y;
		K.cy.y = mB + iB * rB.x * rB.x;
		j->pivotMass = s2GetInverse22(K);
		j->centerDiff0 = s2Sub(bodyB->position, s2MakeVec2(row->target[0], row->target[1]));
		return;
	}

	j->localAnchorA = s2Sub(s2MakeVec2(row->localOriginAnchorA[0], row->localOriginAnchorA[1]), bodyA->localCenter);
	j->invMassA = bodyA->invMass;
	j->invIA = bodyA->invI;
	j->localAnchorB = s2Sub(s2MakeVec2(row->localOriginAnchorB[0], row->localOriginAnchorB[1]), bodyB->localCenter);
	j->invMassB = bodyB->invMass;
	j->invIB = bodyB->invI;
	j->centerDiff0 = s2Sub(bodyB->position, bodyA->position);
	j->referenceAngle = row->referenceAngle;
	j->lowerAngle = row->lowerAngle;
	j->upperAngle = row->upperAngle;
	j->maxMotorTorque = row->maxMotorTorque;
	j->motorSpeed = row->motorSpeed;

	const float zeta = 10.0f;
	float omega = 2.0f * s2_pi * hertz;
	j->biasCoefficient = omega / (2.0f * zeta + h * omega);
	float c = h * omega * (2.0f * zeta + h * omega);
	j->impulseCoefficient = 1.0f / (1.0f + c);
	j->massCoefficient = c * j->impulseCoefficient;

	float iA = j->invIA, iB = j->invIB;
	j->axialMass = iA + iB;
	int fixedRotation;
	if (j->axialMass > 0.0f)
	{
		j->axialMass = 1.0f / j->axialMass;
		fixedRotation = 0;
	}
	else
	{
		fixedRotation = 1;
	}
	if (j->enableLimit == 0 || fixedRotation || warmStart == 0)
	{
		j->lowerImpulse = 0.0f;
		j->upperImpulse = 0.0f;
	}
	if (j->enableMotor == 0 || fixedRotation || warmStart == 0)
	{
		j->motorImpulse = 0.0f;
	}
	if (warmStart == 0)
	{
		j->impulse = s2Vec2_zero;
	}
}

// s2WarmStartRevolute (reference src/revolute_joint.c:107-150) / s2WarmStartMouse (src/mouse_joint.c:85-107)
static void s2oWarmStartJoint(s2oJoint* j, s2oBody* bodies)
{
	s2oBody* bodyB = bodies + j->indexB;
	if (j->type == S2B_JOINT_MOUSE)
	{
		s2Vec2 rB = s2RotateVector(bodyB->q, j->localAnchorB);
		bodyB->v = s2MulAdd(bodyB->v, j->invMassB, j->impulse);
		bodyB->w += j->invIB * (s2Cross(rB, j->impulse) + j->motorImpulse);
		return;
	}
	s2oBody* bodyA = bodies + j->indexA;
	s2Vec2 rA = s2RotateVector(bodyA->q, j->localAnchorA);
	s2Vec2 rB = s2RotateVector(bodyB->q, j->localAnchorB);
	float axialImpulse = j->motorImpulse + j->lowerImpulse - j->upperImpulse;
	s2Vec2 P = j->impulse;
	s2Vec2 vA = bodyA->v, vB = bodyB->v;
	float wA = bodyA->w, wB = bodyB->w;
	vA = s2MulSub(vA, j->invMassA, P);
	wA -= j->invIA * (s2Cross(rA, P) + axialImpulse);
	vB = s2MulAdd(vB, j->invMassB, P);
	wB += j->invIB * (s2Cross(rB, P) + axialImpulse);
	bodyA->v = vA;
	bodyA->w = wA;
	bodyB->v = vB;
	bodyB->w = wB;
}

// s2SolveMouse (reference src/mouse_joint.c:109-167)
static void s2oSolveMouse(s2oJoint* j, s2oBody* bodies, float contextH)
{
	s2oBody* bodyB = bodies + j->indexB;
	s2Vec2 vB = bodyB->v;
	float wB = bodyB->w;
	float mB = j->invMassB, iB = j->invIB;
	{
		float zeta = 0.1f;
		float omega = 2.0f * s2_pi * 0.5f;
		float c = contextH * omega * (2.0f * zeta + contextH * omega);
		float impulseScale = 1.0f / (1.0f + c);
		float massScale = c * impulseScale;
		float impulse = -massScale * j->bodyBInertia * wB - impulseScale * j->motorImpulse;
		j->motorImpulse += impulse;
		wB += iB * impulse;
	}
	{
		s2Vec2 rB = s2RotateVector(bodyB->q, j->localAnchorB);
		s2Vec2 Cdot = s2Add(vB, s2CrossSV(wB, rB));
		s2Vec2 separation = s2Add(s2Add(bodyB->dp, rB), j->centerDiff0);
		s2Vec2 bias = s2MulSV(j->biasCoefficient, separation);
		s2Vec2 b = s2MulMV(j->pivotMass, s2Add(Cdot, bias));
		s2Vec2 impulse;
		impulse.x = -j->massCoefficient * b.x - j->impulseCoefficient * j->impulse.x;
		impulse.y = -j->massCoefficient * b.y - j->impulseCoefficient * j->impulse.y;
		j->impulse.x += impulse.x;
		j->impulse.y += impulse.y;
		vB = s2MulAdd(vB, mB, impulse);
		wB += iB * s2Cross(rB, impulse);
	}
	bodyB->v = vB;
	bodyB->w = wB;
}

// s2SolveRevolute_Soft (reference src/revolute_joint.c:508-657), including the lowerImpulse quirk at :595
static void s2oSolveRevoluteSoft(s2oJoint* j, s2oBody* bodies, float h, float inv_h, int useBias)
{
	s2oBody* bodyA = bodies + j->indexA;
	s2oBody* bodyB = bodies + j->indexB;
	s2Vec2 vA = bodyA->v, vB = bodyB->v;
	float wA = bodyA->w, wB = bodyB->w;
	float mA = j->invMassA, mB = j->invMassB, iA = j->invIA, iB = j->invIB;
	int fixedRotation = (iA + iB == 0.0f);

	if (j->enableMotor && fixedRotation == 0)
	{
		float Cdot = wB - wA - j->motorSpeed;
		float impulse = -j->axialMass * Cdot;
		float oldImpulse = j->motorImpulse;
		float maxImpulse = h * j->maxMotorTorque;
		j->motorImpulse = S2_CLAMP(j->motorImpulse + impulse, -maxImpulse, maxImpulse);
		impulse = j->motorImpulse - oldImpulse;
		wA -= iA * impulse;
		wB += iB * impulse;
	}

	if (j->enableLimit && fixedRotation == 0)
	{
		float jointAngle = s2RelativeAngle(bodyB->q, bodyA->q) - j->referenceAngle;
		{
			float C = jointAngle - j->lowerAngle;
			float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
			if (C > 0.0f)
			{
				bias = C * inv_h;
			}
			else if (useBias)
			{
				bias = j->biasCoefficient * C;
				massScale = j->massCoefficient;
				impulseScale = j->impulseCoefficient;
			}
			float Cdot = wB - wA;
			float impulse = -j->axialMass * massScale * (Cdot + bias) - impulseScale * j->lowerImpulse;
			float oldImpulse = j->lowerImpulse;
			j->lowerImpulse = S2_MAX(j->lowerImpulse + impulse, 0.0f);
			impulse = j->lowerImpulse - oldImpulse;
			wA -= iA * impulse;
			wB += iB * impulse;
		}
		{
			float C = j->upperAngle - jointAngle;
			float bias = 0.0f, massScale = 1.0f, impulseScale = 0.0f;
			if (C > 0.0f)
			{
				bias = C * inv_h;
			}
			else if (useBias)
			{
				bias = j->biasCoefficient * C;
				massScale = j->massCoefficient;
				impulseScale = j->impulseCoefficient;
			}
			float Cdot = wA - wB;
			float impulse = -j->axialMass * massScale * (Cdot + bias) - impulseScale * j->lowerImpulse;
			float oldImpulse = j->upperImpulse;
			j->upperImpulse = S2_MAX(j->upperImpulse + impulse, 0.0f);
			impulse = j->upperImpulse - oldImpulse;
			wA += iA * impulse;
			wB -= iB * impulse;
		}
	}

	{
		s2Vec2 rA = s2RotateVector(bodyA->q, j->localAnchorA);
		s2Vec2 rB = s2RotateVector(bodyB->q, j->localAnchorB);
		s2Vec2 Cdot = s2Sub(s2Add(vB, s2CrossSV(wB, rB)), s2Add(vA, s2CrossSV(wA, rA)));
		s2Vec2 bias = s2Vec2_zero;
		float massScale = 1.0f, impulseScale = 0.0f;
		if (useBias)
		{
			s2Vec2 separation = s2Add(s2Add(s2Sub(bodyB->dp, bodyA->dp), s2Sub(rB, rA)), j->centerDiff0);
			bias = s2MulSV(j->biasCoefficient, separation);
			massScale = j->massCoefficient;
			impulseScale = j->impulseCoefficient;
		}
		s2Mat22 K;
		K.cx.x = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
		K.cy.x = -rA.y * rA.x * iA - rB.y * rB.x * iB;
		K.cx.y = K.cy.x;
		K.cy.y = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
		s2Vec2 b = s2Solve22(K, s2Add(Cdot, bias));
		s2Vec2 impulse;
		impulse.x = -massScale * b.x - impulseScale * j->impulse.x;
		impulse.y = -massScale * b.y - impulseScale * j->impulse.y;
		j->impulse.x += impulse.x;
		j->impulse.y += impulse.y;
		vA = s2MulSub(vA, mA, impulse);
		wA -= iA * s2Cross(rA, impulse);
		vB = s2MulAdd(vB, mB, impulse);
		wB += iB * s2Cross(rB, impulse);
	}

	bodyA->v = vA;
	bodyA->w = wA;
	bodyB->v = vB;
	bodyB->w = wB;
}

// s2SolveJoint_Soft (reference src/joint.c:385-405)
static void s2oSolveJointSoft(s2oJoint* j, s2oBody* bodies, float contextH, float h, float inv_h, int useBias)
{
	if (j->type == S2B_JOINT_MOUSE)
	{
		if (useBias)
		{
			s2oSolveMouse(j, bodies, contextH);
		}
		return;
	}
	s2oSolveRevoluteSoft(j, bodies, h, inv_h, useBias);
}

#include "s2o_variants.inc"

#include "s2o_driver.inc"
