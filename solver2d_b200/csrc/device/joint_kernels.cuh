// solver2d-b200 — per-joint-constraint device functions (revolute + mouse), one thread per joint.
// Joints take part in the same colouring as contacts: a body appears at most once per group across both kinds.
// Float expressions follow the reference op for op; citations on each function. Reference quirks that parity
// depends on are kept on purpose.
#pragma once

#include "contact_kernels.cuh"

#define S2B_JOINT_TYPE(flags) (((flags) >> 1) & 0x7)

struct JointBodies
{
	int ia, ib;
	float4 velA, velB;
};

__device__ __forceinline__ void s2bStoreJointVelocities(const SolveArgs& a, int ia, int ib, float4 velA, float4 velB,
														s2Vec2 vA, float wA, s2Vec2 vB, float wB, float mA, float iA,
														float mB, float iB)
{
	if ((mA != 0.0f) | (iA != 0.0f))
	{
		a.bodies.vel[ia] = make_float4(vA.x, vA.y, wA, velA.w);
	}
	if ((mB != 0.0f) | (iB != 0.0f))
	{
		a.bodies.vel[ib] = make_float4(vB.x, vB.y, wB, velB.w);
	}
}

// ---------------------------------------------------------------------------------------------------------------
// prepare
// ---------------------------------------------------------------------------------------------------------------

enum JointPrepareKind
{
	JPREP_RIGID = 0, // s2PrepareRevolute       (reference src/revolute_joint.c:30-105)
	JPREP_SOFT = 1,	 // s2PrepareRevolute_Soft  (reference src/revolute_joint.c:421-506)
	JPREP_XPBD = 2,	 // s2PrepareRevolute_XPBD  (reference src/revolute_joint.c:792-823)
};

// Builds row t of the joint-constraint stream from joint slot `slot`. `warmStart` is the flag the *variant* passes
// (TGS_Soft and SoftStep pass true regardless of the step flag, TGS_Sticky false).
template <int KIND> __device__ __forceinline__ void s2bPrepareJoint(const SolveArgs& a, int t, int slot, bool warmStart)
{
	const JointConstraintView& jc = a.jc;
	int4 head = a.joints.head[slot];
	int flags = head.x;
	int ia = head.y, ib = head.z;
	float4 anchors = a.joints.anchors[slot];
	float4 imp = a.joints.imp[slot];
	float4 limp = a.joints.limp[slot];

	float4 orgA = a.bodies.org[ia], orgB = a.bodies.org[ib];
	float4 posA = a.bodies.pos[ia], posB = a.bodies.pos[ib];
	float4 poseA = a.bodies.pose[ia], poseB = a.bodies.pose[ib];
	float mA = a.bodies.vel[ia].w, mB = a.bodies.vel[ib].w;
	float iA = posA.z, iB = posB.z;

	jc.head[t] = make_int4(flags, ia, ib, slot);

	if (S2B_JOINT_TYPE(flags) == S2B_JOINT_MOUSE)
	{
		// s2PrepareMouse (reference src/mouse_joint.c:31-83); the same function serves every variant
		float4 motor = a.joints.motor[slot];
		float4 target = a.joints.target[slot];
		s2Vec2 lB = s2Sub(V2(anchors.z, anchors.w), V2(orgB.z, orgB.w));
		float h = a.ctx.h;
		float zeta = motor.w;
		float omega = 2.0f * s2_pi * motor.z;
		float biasCoefficient = omega / (2.0f * zeta + h * omega);
		float c = h * omega * (2.0f * zeta + h * omega);
		float impulseCoefficient = 1.0f / (1.0f + c);
		float massCoefficient = c * impulseCoefficient;

		s2Vec2 rB = s2RotateVector(R2(poseB.z, poseB.w), lB);
		s2Mat22 K;
		K.cx.x = mB + iB * rB.y * rB.y;
		K.cx.y = -iB * rB.x * rB.y;
		K.cy.x = K.cx.y;
		K.cy.y = mB + iB * rB.x * rB.x;
		s2Mat22 pivot = s2GetInverse22(K);
		s2Vec2 centerDiff0 = s2Sub(V2(posB.x, posB.y), V2(target.x, target.y));

		jc.anchor[t] = make_float4(0.0f, 0.0f, lB.x, lB.y);
		jc.mass[t] = make_float4(0.0f, 0.0f, mB, iB);
		jc.d0ax[t] = make_float4(centerDiff0.x, centerDiff0.y, 0.0f, posB.w /* I of body B */);
		jc.lim[t] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		jc.motor[t] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		jc.coef[t] = make_float4(biasCoefficient, massCoefficient, impulseCoefficient, 0.0f);
		jc.pivot[t] = make_float4(pivot.cx.x, pivot.cx.y, pivot.cy.x, pivot.cy.y);
		jc.imp[t] = imp;
		jc.limp[t] = limp;
		return;
	}

	// revolute
	float4 lim = a.joints.lim[slot];
	float4 motor = a.joints.motor[slot];
	s2Vec2 lA = s2Sub(V2(anchors.x, anchors.y), V2(orgA.z, orgA.w));
	s2Vec2 lB = s2Sub(V2(anchors.z, anchors.w), V2(orgB.z, orgB.w));
	s2Vec2 centerDiff0 = s2Sub(V2(posB.x, posB.y), V2(posA.x, posA.y));

	jc.anchor[t] = make_float4(lA.x, lA.y, lB.x, lB.y);
	jc.mass[t] = make_float4(mA, iA, mB, iB);
	jc.lim[t] = lim;
	jc.motor[t] = make_float4(motor.x, motor.y, 0.0f, 0.0f);
	jc.coef[t] = make_float4(a.softJoint.bias, a.softJoint.mass, a.softJoint.impulse, 0.0f);

	if (KIND == JPREP_XPBD)
	{
		jc.d0ax[t] = make_float4(centerDiff0.x, centerDiff0.y, 0.0f, 0.0f);
		jc.pivot[t] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		jc.imp[t] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		jc.limp[t] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		return;
	}

	s2Vec2 rA = s2RotateVector(R2(poseA.z, poseA.w), lA);
	s2Vec2 rB = s2RotateVector(R2(poseB.z, poseB.w), lB);
	s2Mat22 K;
	K.cx.x = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
	K.cy.x = -rA.y * rA.x * iA - rB.y * rB.x * iB;
	K.cx.y = K.cy.x;
	K.cy.y = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
	s2Mat22 pivot = s2GetInverse22(K);
	jc.pivot[t] = make_float4(pivot.cx.x, pivot.cx.y, pivot.cy.x, pivot.cy.y);

	float axialMass = iA + iB;
	bool fixedRotation;
	if (axialMass > 0.0f)
	{
		axialMass = 1.0f / axialMass;
		fixedRotation = false;
	}
	else
	{
		fixedRotation = true;
	}
	jc.d0ax[t] = make_float4(centerDiff0.x, centerDiff0.y, axialMass, 0.0f);

	bool enableLimit = (flags & S2B_JOINT_ENABLE_LIMIT) != 0;
	bool enableMotor = (flags & S2B_JOINT_ENABLE_MOTOR) != 0;
	if (enableLimit == false || fixedRotation || warmStart == false)
	{
		limp.x = 0.0f;
		limp.y = 0.0f;
	}
	if (enableMotor == false || fixedRotation || warmStart == false)
	{
		imp.z = 0.0f;
	}
	if (warmStart == false)
	{
		imp.x = 0.0f;
		imp.y = 0.0f;
	}
	jc.imp[t] = imp;
	jc.limp[t] = limp;
}

// write the accumulated joint impulses back to the persistent joint columns
__device__ __forceinline__ void s2bStoreJointImpulses(const SolveArgs& a, int t)
{
	int slot = a.jc.head[t].w;
	a.joints.imp[slot] = a.jc.imp[t];
	a.joints.limp[slot] = a.jc.limp[t];
}

// ---------------------------------------------------------------------------------------------------------------
// warm start: s2WarmStartRevolute (reference src/revolute_joint.c:107-150), s2WarmStartMouse (mouse_joint.c:85-107)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void s2bWarmStartJoint(const SolveArgs& a, int t)
{
	const JointConstraintView& jc = a.jc;
	int4 head = jc.head[t];
	int ia = head.y, ib = head.z;
	float4 anchor = jc.anchor[t];
	float4 mass = jc.mass[t];
	float4 imp = jc.imp[t];

	if (S2B_JOINT_TYPE(head.x) == S2B_JOINT_MOUSE)
	{
		float4 velB = a.bodies.vel[ib];
		float4 poseB = a.bodies.pose[ib];
		s2Vec2 rB = s2RotateVector(R2(poseB.z, poseB.w), V2(anchor.z, anchor.w));
		s2Vec2 vB = V2(velB.x, velB.y);
		float wB = velB.z;
		s2Vec2 impulse = V2(imp.x, imp.y);
		vB = s2MulAdd(vB, mass.z, impulse);
		wB += mass.w * (s2Cross(rB, impulse) + imp.z);
		if ((mass.z != 0.0f) | (mass.w != 0.0f))
		{
			a.bodies.vel[ib] = make_float4(vB.x, vB.y, wB, velB.w);
		}
		return;
	}

	float4 limp = jc.limp[t];
	float4 velA = a.bodies.vel[ia], velB = a.bodies.vel[ib];
	float4 poseA = a.bodies.pose[ia], poseB = a.bodies.pose[ib];
	s2Vec2 vA = V2(velA.x, velA.y), vB = V2(velB.x, velB.y);
	float wA = velA.z, wB = velB.z;
	s2Vec2 rA = s2RotateVector(R2(poseA.z, poseA.w), V2(anchor.x, anchor.y));
	s2Vec2 rB = s2RotateVector(R2(poseB.z, poseB.w), V2(anchor.z, anchor.w));
	float mA = mass.x, iA = mass.y, mB = mass.z, iB = mass.w;

	float axialImpulse = imp.z + limp.x - limp.y;
	s2Vec2 P = V2(imp.x, imp.y);

	vA = s2MulSub(vA, mA, P);
	wA -= iA * (s2Cross(rA, P) + axialImpulse);
	vB = s2MulAdd(vB, mB, P);
	wB += iB * (s2Cross(rB, P) + axialImpulse);

	s2bStoreJointVelocities(a, ia, ib, velA, velB, vA, wA, vB, wB, mA, iA, mB, iB);
}

// ---------------------------------------------------------------------------------------------------------------
// s2SolveMouse (reference src/mouse_joint.c:109-167): ad-hoc angular damper + soft point constraint on body B.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void s2bSolveMouse(const SolveArgs& a, int t)
{
	const JointConstraintView& jc = a.jc;
	int ib = jc.head[t].z;
	float4 anchor = jc.anchor[t];
	float4 mass = jc.mass[t];
	float4 d0 = jc.d0ax[t];
	float4 coef = jc.coef[t];
	float4 pv = jc.pivot[t];
	float4 imp = jc.imp[t];
	float4 velB = a.bodies.vel[ib];
	float4 poseB = a.bodies.pose[ib];

	s2Vec2 vB = V2(velB.x, velB.y);
	float wB = velB.z;
	float mB = mass.z, iB = mass.w;

	{
		float h = a.ctx.h;
		float zeta = 0.1f;
		float omega = 2.0f * s2_pi * 0.5f;
		float c = h * omega * (2.0f * zeta + h * omega);
		float impulseScale = 1.0f / (1.0f + c);
		float massScale = c * impulseScale;
		float impulse = -massScale * d0.w * wB - impulseScale * imp.z;
		imp.z += impulse;
		wB += iB * impulse;
	}

	{
		s2Vec2 rB = s2RotateVector(R2(poseB.z, poseB.w), V2(anchor.z, anchor.w));
		s2Vec2 Cdot = s2Add(vB, s2CrossSV(wB, rB));
		s2Vec2 dcB = V2(poseB.x, poseB.y);
		s2Vec2 separation = s2Add(s2Add(dcB, rB), V2(d0.x, d0.y));
		s2Vec2 bias = s2MulSV(coef.x, separation);
		float massScale = coef.y;
		float impulseScale = coef.z;
		s2Mat22 pivot;
		pivot.cx = V2(pv.x, pv.y);
		pivot.cy = V2(pv.z, pv.w);
		s2Vec2 b = s2MulMV(pivot, s2Add(Cdot, bias));
		s2Vec2 impulse;
		impulse.x = -massScale * b.x - impulseScale * imp.x;
		impulse.y = -massScale * b.y - impulseScale * imp.y;
		imp.x += impulse.x;
		imp.y += impulse.y;
		vB = s2MulAdd(vB, mB, impulse);
		wB += iB * s2Cross(rB, impulse);
	}

	jc.imp[t] = imp;
	if ((mB != 0.0f) | (iB != 0.0f))
	{
		a.bodies.vel[ib] = make_float4(vB.x, vB.y, wB, velB.w);
	}
}

// ---------------------------------------------------------------------------------------------------------------
// s2SolveRevolute_Soft (reference src/revolute_joint.c:508-657) and s2SolveRevolute_Baumgarte (:660-790):
// motor, lower / upper limit, point-to-point with a fresh 2x2 effective mass (S2_FRESH_PIVOT_MASS == 1).
// ---------------------------------------------------------------------------------------------------------------
enum JointSolveKind
{
	JSOLVE_SOFT = 0,
	JSOLVE_BAUMGARTE = 1,
};

template <int KIND> __device__ __forceinline__ void s2bSolveRevoluteVelocity(const SolveArgs& a, int t, float h, float inv_h, bool useBias)
{
	const JointConstraintView& jc = a.jc;
	int4 head = jc.head[t];
	int ia = head.y, ib = head.z;
	float4 anchor = jc.anchor[t];
	float4 mass = jc.mass[t];
	float4 d0 = jc.d0ax[t];
	float4 lim = jc.lim[t];
	float4 motor = jc.motor[t];
	float4 coef = jc.coef[t];
	float4 imp = jc.imp[t];
	float4 limp = jc.limp[t];

	float4 velA = a.bodies.vel[ia], velB = a.bodies.vel[ib];
	float4 poseA = a.bodies.pose[ia], poseB = a.bodies.pose[ib];
	s2Vec2 vA = V2(velA.x, velA.y), vB = V2(velB.x, velB.y);
	float wA = velA.z, wB = velB.z;
	float mA = mass.x, iA = mass.y, mB = mass.z, iB = mass.w;
	float axialMass = d0.z;
	s2Rot qA = R2(poseA.z, poseA.w), qB = R2(poseB.z, poseB.w);

	bool fixedRotation = (iA + iB == 0.0f);
	bool enableLimit = (head.x & S2B_JOINT_ENABLE_LIMIT) != 0;
	bool enableMotor = (head.x & S2B_JOINT_ENABLE_MOTOR) != 0;

	if (enableMotor && fixedRotation == false)
	{
		float Cdot = wB - wA - motor.y;
		float impulse = -axialMass * Cdot;
		float oldImpulse = imp.z;
		float maxImpulse = h * motor.x;
		imp.z = S2_CLAMP(imp.z + impulse, -maxImpulse, maxImpulse);
		impulse = imp.z - oldImpulse;
		wA -= iA * impulse;
		wB += iB * impulse;
	}

	if (enableLimit && fixedRotation == false)
	{
		float jointAngle = s2RelativeAngle(qB, qA) - lim.x;

		// lower limit
		{
			float C = jointAngle - lim.y;
			float bias = 0.0f;
			float massScale = 1.0f;
			float impulseScale = 0.0f;
			if (C > 0.0f)
			{
				bias = C * inv_h;
			}
			else if (useBias)
			{
				if (KIND == JSOLVE_SOFT)
				{
					bias = coef.x * C;
					massScale = coef.y;
					impulseScale = coef.z;
				}
				else
				{
					bias = s2_baumgarte * inv_h * C;
				}
			}
			float Cdot = wB - wA;
			float impulse = KIND == JSOLVE_SOFT ? -axialMass * massScale * (Cdot + bias) - impulseScale * limp.x
												: -axialMass * (Cdot + bias);
			float oldImpulse = limp.x;
			limp.x = S2_MAX(limp.x + impulse, 0.0f);
			impulse = limp.x - oldImpulse;
			wA -= iA * impulse;
			wB += iB * impulse;
		}

		// upper limit (signs flipped so C stays positive when satisfied)
		{
			float C = lim.z - jointAngle;
			float bias = 0.0f;
			float massScale = 1.0f;
			float impulseScale = 0.0f;
			if (C > 0.0f)
			{
				bias = C * inv_h;
			}
			else if (useBias)
			{
				if (KIND == JSOLVE_SOFT)
				{
					bias = coef.x * C;
					massScale = coef.y;
					impulseScale = coef.z;
				}
				else
				{
					bias = s2_baumgarte * inv_h * C;
				}
			}
			float Cdot = wA - wB;
			// the soft flavour relaxes with the *lower* impulse here: a reference quirk kept for parity
			// (reference src/revolute_joint.c:595)
			float impulse = KIND == JSOLVE_SOFT ? -axialMass * massScale * (Cdot + bias) - impulseScale * limp.x
												: -axialMass * (Cdot + bias);
			float oldImpulse = limp.y;
			limp.y = S2_MAX(limp.y + impulse, 0.0f);
			impulse = limp.y - oldImpulse;
			wA += iA * impulse;
			wB -= iB * impulse;
		}
	}

	// point-to-point
	{
		s2Vec2 rA = s2RotateVector(qA, V2(anchor.x, anchor.y));
		s2Vec2 rB = s2RotateVector(qB, V2(anchor.z, anchor.w));
		s2Vec2 Cdot = s2Sub(s2Add(vB, s2CrossSV(wB, rB)), s2Add(vA, s2CrossSV(wA, rA)));

		s2Vec2 bias = V2(0.0f, 0.0f);
		float massScale = 1.0f;
		float impulseScale = 0.0f;
		if (KIND == JSOLVE_SOFT)
		{
			if (useBias)
			{
				s2Vec2 dcA = V2(poseA.x, poseA.y), dcB = V2(poseB.x, poseB.y);
				s2Vec2 separation = s2Add(s2Add(s2Sub(dcB, dcA), s2Sub(rB, rA)), V2(d0.x, d0.y));
				bias = s2MulSV(coef.x, separation);
				massScale = coef.y;
				impulseScale = coef.z;
			}
		}
		else
		{
			// the Baumgarte flavour always applies its bias (reference src/revolute_joint.c:757-758)
			s2Vec2 dcA = V2(poseA.x, poseA.y), dcB = V2(poseB.x, poseB.y);
			s2Vec2 separation = s2Add(s2Add(s2Sub(dcB, dcA), s2Sub(rB, rA)), V2(d0.x, d0.y));
			bias = s2MulSV(s2_baumgarte * inv_h, separation);
		}

		s2Mat22 K;
		K.cx.x = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
		K.cy.x = -rA.y * rA.x * iA - rB.y * rB.x * iB;
		K.cx.y = K.cy.x;
		K.cy.y = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
		s2Vec2 b = s2Solve22(K, s2Add(Cdot, bias));

		s2Vec2 impulse;
		if (KIND == JSOLVE_SOFT)
		{
			impulse.x = -massScale * b.x - impulseScale * imp.x;
			impulse.y = -massScale * b.y - impulseScale * imp.y;
		}
		else
		{
			impulse.x = -b.x;
			impulse.y = -b.y;
		}
		imp.x += impulse.x;
		imp.y += impulse.y;

		vA = s2MulSub(vA, mA, impulse);
		wA -= iA * s2Cross(rA, impulse);
		vB = s2MulAdd(vB, mB, impulse);
		wB += iB * s2Cross(rB, impulse);
	}

	jc.imp[t] = imp;
	jc.limp[t] = limp;
	s2bStoreJointVelocities(a, ia, ib, velA, velB, vA, wA, vB, wB, mA, iA, mB, iB);
}

// s2SolveJoint_Soft dispatch (reference src/joint.c:385-405): the mouse joint is solved only in the biased pass.
__device__ __forceinline__ void s2bSolveJointSoft(const SolveArgs& a, int t, float h, float inv_h, bool useBias)
{
	int flags = a.jc.head[t].x;
	if (S2B_JOINT_TYPE(flags) == S2B_JOINT_MOUSE)
	{
		if (useBias)
		{
			s2bSolveMouse(a, t);
		}
		return;
	}
	s2bSolveRevoluteVelocity<JSOLVE_SOFT>(a, t, h, inv_h, useBias);
}

// s2SolveJoint_Baumgarte dispatch (reference src/joint.c:409-425)
__device__ __forceinline__ void s2bSolveJointBaumgarte(const SolveArgs& a, int t, float h, float inv_h, bool useBias)
{
	int flags = a.jc.head[t].x;
	if (S2B_JOINT_TYPE(flags) == S2B_JOINT_MOUSE)
	{
		s2bSolveMouse(a, t);
		return;
	}
	s2bSolveRevoluteVelocity<JSOLVE_BAUMGARTE>(a, t, h, inv_h, useBias);
}

// ---------------------------------------------------------------------------------------------------------------
// s2SolveRevolute (reference src/revolute_joint.c:152-303): rigid velocity solve used by PGS_NGS and TGS_NGS — motor,
// limits with a speculative max(C, 0)/h term, point-to-point with the pivot mass computed at prepare time.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void s2bSolveRevoluteRigid(const SolveArgs& a, int t, float h)
{
	const JointConstraintView& jc = a.jc;
	int4 head = jc.head[t];
	int ia = head.y, ib = head.z;
	float4 anchor = jc.anchor[t];
	float4 mass = jc.mass[t];
	float4 d0 = jc.d0ax[t];
	float4 lim = jc.lim[t];
	float4 motor = jc.motor[t];
	float4 pv = jc.pivot[t];
	float4 imp = jc.imp[t];
	float4 limp = jc.limp[t];
	float4 velA = a.bodies.vel[ia], velB = a.bodies.vel[ib];
	float4 poseA = a.bodies.pose[ia], poseB = a.bodies.pose[ib];
	s2Vec2 vA = V2(velA.x, velA.y), vB = V2(velB.x, velB.y);
	float wA = velA.z, wB = velB.z;
	float mA = mass.x, iA = mass.y, mB = mass.z, iB = mass.w;
	float axialMass = d0.z;
	s2Rot qA = R2(poseA.z, poseA.w), qB = R2(poseB.z, poseB.w);
	bool fixedRotation = (iA + iB == 0.0f);

	if ((head.x & S2B_JOINT_ENABLE_MOTOR) && fixedRotation == false)
	{
		float Cdot = wB - wA - motor.y;
		float impulse = -axialMass * Cdot;
		float oldImpulse = imp.z;
		float maxImpulse = h * motor.x;
		imp.z = S2_CLAMP(imp.z + impulse, -maxImpulse, maxImpulse);
		impulse = imp.z - oldImpulse;
		wA -= iA * impulse;
		wB += iB * impulse;
	}

	if ((head.x & S2B_JOINT_ENABLE_LIMIT) && fixedRotation == false)
	{
		float angle = s2RelativeAngle(qB, qA) - lim.x;
		{
			float C = angle - lim.y;
			float Cdot = wB - wA;
			float impulse = -axialMass * (Cdot + S2_MAX(C, 0.0f) / h);
			float oldImpulse = limp.x;
			limp.x = S2_MAX(limp.x + impulse, 0.0f);
			impulse = limp.x - oldImpulse;
			wA -= iA * impulse;
			wB += iB * impulse;
		}
		{
			float C = lim.z - angle;
			float Cdot = wA - wB;
			float impulse = -axialMass * (Cdot + S2_MAX(C, 0.0f) / h);
			float oldImpulse = limp.y;
			limp.y = S2_MAX(limp.y + impulse, 0.0f);
			impulse = limp.y - oldImpulse;
			wA += iA * impulse;
			wB -= iB * impulse;
		}
	}

	{
		s2Vec2 rA = s2RotateVector(qA, V2(anchor.x, anchor.y));
		s2Vec2 rB = s2RotateVector(qB, V2(anchor.z, anchor.w));
		s2Vec2 Cdot = s2Sub(s2Add(vB, s2CrossSV(wB, rB)), s2Add(vA, s2CrossSV(wA, rA)));
		s2Mat22 pivot;
		pivot.cx = V2(pv.x, pv.y);
		pivot.cy = V2(pv.z, pv.w);
		s2Vec2 impulse = s2MulMV(pivot, s2Neg(Cdot));
		imp.x += impulse.x;
		imp.y += impulse.y;
		vA = s2MulSub(vA, mA, impulse);
		wA -= iA * s2Cross(rA, impulse);
		vB = s2MulAdd(vB, mB, impulse);
		wB += iB * s2Cross(rB, impulse);
	}

	jc.imp[t] = imp;
	jc.limp[t] = limp;
	s2bStoreJointVelocities(a, ia, ib, velA, velB, vA, wA, vB, wB, mA, iA, mB, iB);
}

// s2SolveJoint dispatch (reference src/joint.c:329-345)
__device__ __forceinline__ void s2bSolveJointRigid(const SolveArgs& a, int t, float h)
{
	if (S2B_JOINT_TYPE(a.jc.head[t].x) == S2B_JOINT_MOUSE)
	{
		s2bSolveMouse(a, t);
		return;
	}
	s2bSolveRevoluteRigid(a, t, h);
}

// ---------------------------------------------------------------------------------------------------------------
// s2SolveRevolutePosition (reference src/revolute_joint.c:305-419): NGS position pass for the revolute joint (angular
// limit + point-to-point with a fresh 2x2 mass); s2SolveJointPosition does nothing for a mouse joint (joint.c:349-361).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void s2bSolveJointPosition(const SolveArgs& a, int t)
{
	const JointConstraintView& jc = a.jc;
	int4 head = jc.head[t];
	if (S2B_JOINT_TYPE(head.x) == S2B_JOINT_MOUSE)
	{
		return;
	}
	int ia = head.y, ib = head.z;
	float4 anchor = jc.anchor[t];
	float4 mass = jc.mass[t];
	float4 d0 = jc.d0ax[t];
	float4 lim = jc.lim[t];
	float4 poseA = a.bodies.pose[ia], poseB = a.bodies.pose[ib];
	s2Vec2 dcA = V2(poseA.x, poseA.y), dcB = V2(poseB.x, poseB.y);
	s2Rot qA = R2(poseA.z, poseA.w), qB = R2(poseB.z, poseB.w);
	float mA = mass.x, iA = mass.y, mB = mass.z, iB = mass.w;
	float axialMass = d0.z;
	bool fixedRotation = (iA + iB == 0.0f);

	if ((head.x & S2B_JOINT_ENABLE_LIMIT) && fixedRotation == false)
	{
		float angle = s2RelativeAngle(qB, qA) - lim.x;
		float C = 0.0f;
		if (S2_ABS(lim.z - lim.y) < 2.0f * s2_angularSlop)
		{
			C = S2_CLAMP(angle - lim.y, -s2_maxAngularCorrection, s2_maxAngularCorrection);
		}
		else if (angle <= lim.y)
		{
			C = S2_CLAMP(angle - lim.y + s2_angularSlop, -s2_maxAngularCorrection, 0.0f);
		}
		else if (angle >= lim.z)
		{
			C = S2_CLAMP(angle - lim.z - s2_angularSlop, 0.0f, s2_maxAngularCorrection);
		}
		float limitImpulse = -axialMass * C;
		qA = s2IntegrateRot(qA, -iA * limitImpulse);
		qB = s2IntegrateRot(qB, iB * limitImpulse);
	}

	{
		s2Vec2 rA = s2RotateVector(qA, V2(anchor.x, anchor.y));
		s2Vec2 rB = s2RotateVector(qB, V2(anchor.z, anchor.w));
		s2Vec2 C = s2Add(s2Add(s2Sub(dcB, dcA), s2Sub(rB, rA)), V2(d0.x, d0.y));
		s2Mat22 K;
		K.cx.x = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
		K.cx.y = -iA * rA.x * rA.y - iB * rB.x * rB.y;
		K.cy.x = K.cx.y;
		K.cy.y = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
		s2Vec2 impulse = s2Solve22(K, s2Neg(C));
		dcA = s2MulSub(dcA, mA, impulse);
		qA = s2IntegrateRot(qA, -iA * s2Cross(rA, impulse));
		dcB = s2MulAdd(dcB, mB, impulse);
		qB = s2IntegrateRot(qB, iB * s2Cross(rB, impulse));
	}

	if ((mA != 0.0f) || (iA != 0.0f))
	{
		a.bodies.pose[ia] = make_float4(dcA.x, dcA.y, qA.s, qA.c);
	}
	if ((mB != 0.0f) || (iB != 0.0f))
	{
		a.bodies.pose[ib] = make_float4(dcB.x, dcB.y, qB.s, qB.c);
	}
}

// ---------------------------------------------------------------------------------------------------------------
// s2SolveJoint_XPBD (reference src/joint.c:447-463): mouse -> s2SolveMouse; revolute -> s2SolveRevolute_XPBD
// (src/revolute_joint.c:825-888): distance constraint along the current separation, zero compliance.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void s2bSolveJointXpbd(const SolveArgs& a, int t)
{
	const JointConstraintView& jc = a.jc;
	int4 head = jc.head[t];
	if (S2B_JOINT_TYPE(head.x) == S2B_JOINT_MOUSE)
	{
		s2bSolveMouse(a, t);
		return;
	}
	const float compliance = 0.0f;
	int ia = head.y, ib = head.z;
	float4 anchor = jc.anchor[t];
	float4 mass = jc.mass[t];
	float4 d0 = jc.d0ax[t];
	float4 poseA = a.bodies.pose[ia], poseB = a.bodies.pose[ib];
	s2Vec2 dcA = V2(poseA.x, poseA.y), dcB = V2(poseB.x, poseB.y);
	s2Rot qA = R2(poseA.z, poseA.w), qB = R2(poseB.z, poseB.w);
	s2Vec2 rA = s2RotateVector(qA, V2(anchor.x, anchor.y));
	s2Vec2 rB = s2RotateVector(qB, V2(anchor.z, anchor.w));
	s2Vec2 separation = s2Add(s2Add(s2Sub(dcB, dcA), s2Sub(rB, rA)), V2(d0.x, d0.y));
	float c = s2Length(separation);
	// s2Normalize (reference src/math.c:40-51)
	s2Vec2 n = V2(0.0f, 0.0f);
	if (c >= 0.001f * 1.1920929e-07f)
	{
		float invLength = 1.0f / c;
		n = V2(invLength * separation.x, invLength * separation.y);
	}
	float mA = mass.x, iA = mass.y, mB = mass.z, iB = mass.w;
	if (mA == 0.0f && mB == 0.0f)
	{
		return;
	}
	float rnA = s2Cross(rA, n);
	float rnB = s2Cross(rB, n);
	float kA = mA + iA * rnA * rnA;
	float kB = mB + iB * rnB * rnB;
	float lambda = -c / (kA + kB + compliance);
	s2Vec2 p = s2MulSV(lambda, n);
	dcA = s2MulSub(dcA, mA, p);
	qA = s2IntegrateRot(qA, -iA * s2Cross(rA, p));
	dcB = s2MulAdd(dcB, mB, p);
	qB = s2IntegrateRot(qB, iB * s2Cross(rB, p));
	if ((mA != 0.0f) || (iA != 0.0f))
	{
		a.bodies.pose[ia] = make_float4(dcA.x, dcA.y, qA.s, qA.c);
	}
	if ((mB != 0.0f) || (iB != 0.0f))
	{
		a.bodies.pose[ib] = make_float4(dcB.x, dcB.y, qB.s, qB.c);
	}
}
