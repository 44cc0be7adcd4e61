// solver2d-b200 — per-body and per-contact-constraint device functions of the solver stage.
//
// One thread owns one constraint (both manifold points are solved sequentially by that thread, as the points of a
// manifold are coupled through the two bodies). Within a group (colour / wavefront level) no two constraints share a
// movable body, so the read-modify-write of the body velocity columns is race-free without atomics.
//
// Float expressions follow the reference op for op (compiled with -fmad=false), see the citations on each function.
#pragma once

#include "solver_state.cuh"

__device__ __forceinline__ s2Vec2 V2(float x, float y)
{
	s2Vec2 v = {x, y};
	return v;
}

__device__ __forceinline__ s2Rot R2(float s, float c)
{
	s2Rot q = {s, c};
	return q;
}

// ---------------------------------------------------------------------------------------------------------------
// body phases
// ---------------------------------------------------------------------------------------------------------------

// s2IntegrateVelocities (reference src/solve_common.c:10-45): dynamic bodies only; gravity, forces, implicit damping.
// Algorithmic traffic: read vel, frc, prm (48 B) + flag, write vel (16 B).
__device__ __forceinline__ void s2bIntegrateVelocity(const SolveArgs& a, int i, float h)
{
	unsigned f = a.bodies.flags[i];
	if ((f & S2B_BODY_VALID) == 0 || S2B_BODY_TYPE(f) != S2B_BODY_DYNAMIC)
	{
		return;
	}
	float4 vel = a.bodies.vel[i];
	float4 frc = a.bodies.frc[i];
	float4 prm = a.bodies.prm[i];
	float invMass = vel.w, invI = prm.w, mass = frc.w;
	s2Vec2 v = V2(vel.x, vel.y);
	float w = vel.z;
	s2Vec2 gravity = V2(a.gravity.x, a.gravity.y);

	v = s2Add(v, s2MulSV(h * invMass, s2MulAdd(V2(frc.x, frc.y), mass * prm.z, gravity)));
	w = w + h * invI * frc.z;

	v = s2MulSV(1.0f / (1.0f + h * prm.x), v);
	w *= 1.0f / (1.0f + h * prm.y);

	a.bodies.vel[i] = make_float4(v.x, v.y, w, invMass);
}

// s2IntegratePositions (reference src/solve_common.c:47-68): every non-static body (kinematic bodies move too).
__device__ __forceinline__ void s2bIntegratePosition(const SolveArgs& a, int i, float h)
{
	unsigned f = a.bodies.flags[i];
	if ((f & S2B_BODY_VALID) == 0 || S2B_BODY_TYPE(f) == S2B_BODY_STATIC)
	{
		return;
	}
	float4 vel = a.bodies.vel[i];
	float4 pose = a.bodies.pose[i];
	s2Vec2 dp = s2MulAdd(V2(pose.x, pose.y), h, V2(vel.x, vel.y));
	s2Rot q = s2IntegrateRot(R2(pose.z, pose.w), h * vel.z);
	a.bodies.pose[i] = make_float4(dp.x, dp.y, q.s, q.c);
}

// s2FinalizePositions (reference src/solve_common.c:70-91)
__device__ __forceinline__ void s2bFinalizePosition(const SolveArgs& a, int i)
{
	unsigned f = a.bodies.flags[i];
	if ((f & S2B_BODY_VALID) == 0 || S2B_BODY_TYPE(f) == S2B_BODY_STATIC)
	{
		return;
	}
	float4 pos = a.bodies.pos[i];
	float4 pose = a.bodies.pose[i];
	s2Vec2 p = s2Add(V2(pos.x, pos.y), V2(pose.x, pose.y));
	a.bodies.pos[i] = make_float4(p.x, p.y, pos.z, pos.w);
	a.bodies.pose[i] = make_float4(0.0f, 0.0f, pose.z, pose.w);
}

// ---------------------------------------------------------------------------------------------------------------
// contact constraints: prepare
// ---------------------------------------------------------------------------------------------------------------

enum PrepareKind
{
	PREPARE_PGS = 0,  // s2PrepareContacts_PGS  (reference src/solve_common.c:93-168)
	PREPARE_SOFT = 1, // s2PrepareContacts_Soft (reference src/solve_common.c:188-274)
};

// Builds row t of the constraint stream from contact slot src[t]. Writes are coalesced (row t), reads gather the
// persistent manifold and the two bodies.
template <int KIND> __device__ __forceinline__ void s2bPrepareContact(const SolveArgs& a, int t)
{
	const ConstraintView& cc = a.cc;
	int slot = cc.src[t];
	int2 bodies = a.contacts.bodies[slot];
	int4 info = a.contacts.info[slot];
	float4 mnf = a.contacts.nf[slot];
	int pointCount = S2B_CI_COUNT(info.x);

	float4 velA = a.bodies.vel[bodies.x], velB = a.bodies.vel[bodies.y];
	float4 poseA = a.bodies.pose[bodies.x], poseB = a.bodies.pose[bodies.y];
	float4 orgA = a.bodies.org[bodies.x], orgB = a.bodies.org[bodies.y];
	float mA = velA.w, mB = velB.w;
	float iA = a.bodies.prm[bodies.x].w, iB = a.bodies.prm[bodies.y].w;
	s2Rot qA = R2(poseA.z, poseA.w), qB = R2(poseB.z, poseB.w);
	s2Vec2 normal = V2(mnf.x, mnf.y);
	s2Vec2 tangent = s2RightPerp(normal);
	bool warmStart = a.ctx.warmStart != 0;

	// contact stiffness is doubled against a body of infinite mass (reference solve_common.c:219)
	unsigned flags = 0;
	if (mA == 0.0f || mB == 0.0f)
	{
		flags |= S2B_CF_STATIC_SOFT;
	}
	if (pointCount == 2)
	{
		flags |= S2B_CF_TWO_POINTS;
	}
	cc.idx[t] = make_int2(bodies.x, (int)((unsigned)bodies.y | flags));
	cc.nf[t] = make_float4(normal.x, normal.y, mnf.z, iA);

#pragma unroll
	for (int j = 0; j < 2; ++j)
	{
		if (j < pointCount)
		{
			float4 la = a.contacts.anchor[j][slot];
			float4 mi = a.contacts.impulse[j][slot];
			float separation = mi.x;
			// PGS prepare tests cp->separation <= 0 on the zero-filled scratch, i.e. always true, so
			// both prepare flavours take the stored impulses whenever warm starting is on.
			float normalImpulse = warmStart ? mi.y : 0.0f;
			float tangentImpulse = warmStart ? mi.z : 0.0f;

			s2Vec2 lA = s2Sub(V2(la.x, la.y), V2(orgA.z, orgA.w));
			s2Vec2 lB = s2Sub(V2(la.z, la.w), V2(orgB.z, orgB.w));
			s2Vec2 rA = s2RotateVector(qA, lA);
			s2Vec2 rB = s2RotateVector(qB, lB);

			float adjustedSeparation = separation - s2Dot(s2Sub(rB, rA), normal);

			float rnA = s2Cross(rA, normal);
			float rnB = s2Cross(rB, normal);
			float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
			float normalMass = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;

			float rtA = s2Cross(rA, tangent);
			float rtB = s2Cross(rB, tangent);
			float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
			float tangentMass = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;

			cc.anchor[j][t] = make_float4(lA.x, lA.y, lB.x, lB.y);
			cc.pm[j][t] = make_float4(adjustedSeparation, normalMass, tangentMass, j == 0 ? iB : 0.0f);
			cc.lambda[j][t] = make_float2(normalImpulse, tangentImpulse);
			if (cc.r0[j] != nullptr)
			{
				cc.r0[j][t] = make_float4(rA.x, rA.y, rB.x, rB.y);
			}
			if (cc.sep[j] != nullptr)
			{
				cc.sep[j][t] = separation;
			}
		}
		else
		{
			cc.anchor[j][t] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			cc.pm[j][t] = make_float4(0.0f, 0.0f, 0.0f, j == 0 ? iB : 0.0f);
			cc.lambda[j][t] = make_float2(0.0f, 0.0f);
			if (cc.r0[j] != nullptr)
			{
				cc.r0[j][t] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			}
			if (cc.sep[j] != nullptr)
			{
				cc.sep[j][t] = 0.0f;
			}
		}
	}
}

// s2StoreContactImpulses (reference src/solve_common.c:396-410): scatter accumulated impulses back to the manifolds.
__device__ __forceinline__ void s2bStoreContactImpulses(const SolveArgs& a, int t, float scale)
{
	const ConstraintView& cc = a.cc;
	int slot = cc.src[t];
	int pointCount = (cc.idx[t].y & S2B_CF_TWO_POINTS) ? 2 : 1;
#pragma unroll
	for (int j = 0; j < 2; ++j)
	{
		if (j < pointCount)
		{
			float2 l = cc.lambda[j][t];
			float4 mi = a.contacts.impulse[j][slot];
			mi.y = scale == 1.0f ? l.x : l.x * scale;
			mi.z = scale == 1.0f ? l.y : l.y * scale;
			a.contacts.impulse[j][slot] = mi;
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// contact constraints: shared load / store of the two bodies
// ---------------------------------------------------------------------------------------------------------------

struct BodyPair
{
	int ia, ib;
	float4 velA, velB;
	bool movA, movB;
};

__device__ __forceinline__ void s2bStoreVelocities(const SolveArgs& a, const BodyPair& bp, s2Vec2 vA, float wA, s2Vec2 vB,
												   float wB)
{
	// bodies of infinite mass and inertia (static, kinematic) are never written: the reference writes them back
	// unchanged, and skipping the store keeps every group free of write conflicts on them.
	if (bp.movA)
	{
		a.bodies.vel[bp.ia] = make_float4(vA.x, vA.y, wA, bp.velA.w);
	}
	if (bp.movB)
	{
		a.bodies.vel[bp.ib] = make_float4(vB.x, vB.y, wB, bp.velB.w);
	}
}

// ---------------------------------------------------------------------------------------------------------------
// s2WarmStartContacts (reference src/solve_common.c:276-326): impulses applied at the current anchors.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void s2bWarmStartContact(const SolveArgs& a, int t)
{
	const ConstraintView& cc = a.cc;
	int2 idx = cc.idx[t];
	int ia = idx.x, ib = idx.y & S2B_CF_INDEX_MASK;
	int pointCount = (idx.y & S2B_CF_TWO_POINTS) ? 2 : 1;
	float4 nf = cc.nf[t];
	float4 velA = a.bodies.vel[ia], velB = a.bodies.vel[ib];
	float4 poseA = a.bodies.pose[ia], poseB = a.bodies.pose[ib];
	float mA = velA.w, mB = velB.w, iA = nf.w, iB = cc.pm[0][t].w;
	s2Vec2 vA = V2(velA.x, velA.y), vB = V2(velB.x, velB.y);
	float wA = velA.z, wB = velB.z;
	s2Rot qA = R2(poseA.z, poseA.w), qB = R2(poseB.z, poseB.w);
	s2Vec2 normal = V2(nf.x, nf.y);
	s2Vec2 tangent = s2RightPerp(normal);

#pragma unroll
	for (int j = 0; j < 2; ++j)
	{
		if (j < pointCount)
		{
			float4 la = cc.anchor[j][t];
			float2 l = cc.lambda[j][t];
			s2Vec2 rA = s2RotateVector(qA, V2(la.x, la.y));
			s2Vec2 rB = s2RotateVector(qB, V2(la.z, la.w));
			s2Vec2 P = s2Add(s2MulSV(l.x, normal), s2MulSV(l.y, tangent));
			wA -= iA * s2Cross(rA, P);
			vA = s2MulAdd(vA, -mA, P);
			wB += iB * s2Cross(rB, P);
			vB = s2MulAdd(vB, mB, P);
		}
	}

	BodyPair bp = {ia, ib, velA, velB, (mA != 0.0f) || (iA != 0.0f), (mB != 0.0f) || (iB != 0.0f)};
	s2bStoreVelocities(a, bp, vA, wA, vB, wB);
}

// ---------------------------------------------------------------------------------------------------------------
// s2SolveContacts_TGS_Soft (reference src/solve_tgs_soft.c:17-135): THE inner kernel of the headline variant.
// Soft normal constraint evaluated at the current (sub-stepped) anchors, then Coulomb friction.
// Algorithmic traffic per 2-point constraint: stream idx 8 + nf 16 + 2 x (anchor 16 + pm 16 + lambda 8 r + 8 w)
// = 120 B, bodies 2 x (vel 16 + pose 16 r, vel 16 w) = 96 B.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void s2bSolveContactTgsSoft(const SolveArgs& a, int t, float inv_h, bool useBias)
{
	const ConstraintView& cc = a.cc;
	int2 idx = cc.idx[t];
	int ia = idx.x, ib = idx.y & S2B_CF_INDEX_MASK;
	int pointCount = (idx.y & S2B_CF_TWO_POINTS) ? 2 : 1;
	const SoftCoef soft = ((unsigned)idx.y & S2B_CF_STATIC_SOFT) ? a.softStatic : a.softDynamic;

	float4 nf = cc.nf[t];
	float4 la0 = cc.anchor[0][t], pm0 = cc.pm[0][t];
	float2 l0 = cc.lambda[0][t];
	float4 la1 = cc.anchor[1][t], pm1 = cc.pm[1][t];
	float2 l1 = cc.lambda[1][t];

	float4 velA = a.bodies.vel[ia], velB = a.bodies.vel[ib];
	float4 poseA = a.bodies.pose[ia], poseB = a.bodies.pose[ib];

	float mA = velA.w, mB = velB.w, iA = nf.w, iB = pm0.w;
	s2Vec2 vA = V2(velA.x, velA.y), vB = V2(velB.x, velB.y);
	float wA = velA.z, wB = velB.z;
	s2Vec2 dcA = V2(poseA.x, poseA.y), dcB = V2(poseB.x, poseB.y);
	s2Rot qA = R2(poseA.z, poseA.w), qB = R2(poseB.z, poseB.w);

	s2Vec2 normal = V2(nf.x, nf.y);
	s2Vec2 tangent = s2RightPerp(normal);
	float friction = nf.z;

	float4 la[2] = {la0, la1};
	float4 pm[2] = {pm0, pm1};
	float2 lam[2] = {l0, l1};

	// non-penetration
#pragma unroll
	for (int j = 0; j < 2; ++j)
	{
		if (j < pointCount)
		{
			s2Vec2 rA = s2RotateVector(qA, V2(la[j].x, la[j].y));
			s2Vec2 rB = s2RotateVector(qB, V2(la[j].z, la[j].w));

			// separation at the current sub-step pose
			s2Vec2 ds = s2Add(s2Sub(dcB, dcA), s2Sub(rB, rA));
			float s = s2Dot(ds, normal) + pm[j].x;

			float bias = 0.0f;
			float massScale = 1.0f;
			float impulseScale = 0.0f;
			if (s > 0.0f)
			{
				// speculative: allow approach up to the gap
				bias = s * inv_h;
			}
			else if (useBias)
			{
				bias = S2_MAX(soft.bias * s, -s2_maxBaumgarteVelocity);
				massScale = soft.mass;
				impulseScale = soft.impulse;
			}

			s2Vec2 vrB = s2Add(vB, s2CrossSV(wB, rB));
			s2Vec2 vrA = s2Add(vA, s2CrossSV(wA, rA));
			float vn = s2Dot(s2Sub(vrB, vrA), normal);

			float impulse = -pm[j].y * massScale * (vn + bias) - impulseScale * lam[j].x;

			float newImpulse = S2_MAX(lam[j].x + impulse, 0.0f);
			impulse = newImpulse - lam[j].x;
			lam[j].x = newImpulse;

			s2Vec2 P = s2MulSV(impulse, normal);
			vA = s2MulSub(vA, mA, P);
			wA -= iA * s2Cross(rA, P);
			vB = s2MulAdd(vB, mB, P);
			wB += iB * s2Cross(rB, P);
		}
	}

	// friction
#pragma unroll
	for (int j = 0; j < 2; ++j)
	{
		if (j < pointCount)
		{
			s2Vec2 rA = s2RotateVector(qA, V2(la[j].x, la[j].y));
			s2Vec2 rB = s2RotateVector(qB, V2(la[j].z, la[j].w));

			s2Vec2 vrB = s2Add(vB, s2CrossSV(wB, rB));
			s2Vec2 vrA = s2Add(vA, s2CrossSV(wA, rA));
			float vt = s2Dot(s2Sub(vrB, vrA), tangent);

			float impulse = -pm[j].z * vt;

			float maxFriction = friction * lam[j].x;
			float newImpulse = S2_CLAMP(lam[j].y + impulse, -maxFriction, maxFriction);
			impulse = newImpulse - lam[j].y;
			lam[j].y = newImpulse;

			s2Vec2 P = s2MulSV(impulse, tangent);
			vA = s2MulSub(vA, mA, P);
			wA -= iA * s2Cross(rA, P);
			vB = s2MulAdd(vB, mB, P);
			wB += iB * s2Cross(rB, P);
		}
	}

	cc.lambda[0][t] = lam[0];
	if (pointCount == 2)
	{
		cc.lambda[1][t] = lam[1];
	}
	BodyPair bp = {ia, ib, velA, velB, (mA != 0.0f) || (iA != 0.0f), (mB != 0.0f) || (iB != 0.0f)};
	s2bStoreVelocities(a, bp, vA, wA, vB, wB);
}
