// solver2d-b200 — the solver stage of s2World_Step on the device.
//
// Replaces the per-variant entry points s2Solve_* (reference src/solvers.h:70-79, dispatched at world.c:206-256).
// Pipeline per step, all on the world's stream, no host synchronisation on the production path:
//
//   gather      live joints and manifolds with >= 1 point, compacted in slot order      (reference gather loop,
//               e.g. src/solve_tgs_soft.c:162-179: this *is* the sequential Gauss-Seidel order of the reference)
//   schedule    partition joints + contact constraints into groups with no shared movable body:
//                 COLOR      Jones-Plassmann colouring of the constraint graph on the device (<= 64 colours, the
//                            rest spills to a serial overflow group), solve order = colour-major;
//                 WAVEFRONT  order-preserving levels (validation path, bit-exact vs the sequential reference)
//   prepare     build the SoA constraint streams in solve order (coalesced 128-bit rows)
//   iterate     the variant's schedule of body passes and per-group constraint passes, either as ONE persistent
//               cooperative kernel (grid barrier between groups) or as one launch per group (profiling / cross-check)
//   store       accumulated impulses back to the persistent manifolds / joints
#include "joint_kernels.cuh"

#include <cooperative_groups.h>
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>

namespace cg = cooperative_groups;

#define S2B_MAX_COLORS 64
#define S2B_OVERFLOW_KEY 255
#define S2B_BLOCK 128

// ---------------------------------------------------------------------------------------------------------------
// scratch management
// ---------------------------------------------------------------------------------------------------------------

SolverScratch* s2bGetSolverScratch(s2bWorld* w)
{
	if (w->scratch == nullptr)
	{
		w->scratch = new SolverScratch();
	}
	return w->scratch;
}

void s2bFreeSolverScratch(s2bWorld* w)
{
	SolverScratch* s = w->scratch;
	if (s == nullptr)
	{
		return;
	}
	s->counts.release();
	s->activeFlag.release();
	s->activeSlots.release();
	s->jointFlag.release();
	s->jointSlots.release();
	s->itemBodies.release();
	s->degree.release();
	s->adjStart.release();
	s->adjCursor.release();
	s->adj.release();
	s->colorA.release();
	s->colorB.release();
	s->sortKeyIn.release();
	s->sortKeyOut.release();
	s->sortValIn.release();
	s->sortValOut.release();
	s->cGroupOff.release();
	s->jGroupOff.release();
	s->cPerm.release();
	s->jPerm.release();
	s->cubTemp.release();
	s->idx.release();
	s->nf.release();
	for (int p = 0; p < 2; ++p)
	{
		s->anchor[p].release();
		s->pm[p].release();
		s->r0[p].release();
		s->fanchor[p].release();
		s->lambda[p].release();
		s->tsep[p].release();
		s->sep[p].release();
	}
	s->src.release();
	s->jhead.release();
	s->janchor.release();
	s->jmass.release();
	s->jd0ax.release();
	s->jlim.release();
	s->jmotor.release();
	s->jcoef.release();
	s->jpivot.release();
	s->jimp.release();
	s->jlimp.release();
	delete s;
	w->scratch = nullptr;
}

// which optional constraint columns a variant needs
struct VariantColumns
{
	bool r0;   // prepare-time world anchors
	bool sep;  // prepare-time separation
	bool sticky;
};

static VariantColumns columnsFor(int solverType)
{
	VariantColumns c = {false, false, false};
	switch (solverType)
	{
		case 0: // Jacobi
		case 1: // PGS
		case 2: // PGS_NGS
		case 3: // PGS_NGS_Block
		case 4: // PGS_Soft
			c.r0 = true;
			c.sep = true;
			break;
		case 5: // SoftStep
			c.r0 = true;
			break;
		case 6: // TGS_Sticky
			c.sticky = true;
			break;
		case 7: // TGS_Soft
			break;
		case 8: // TGS_NGS
			c.sep = true;
			break;
		case 9: // XPBD
			c.r0 = true;
			c.sep = true;
			break;
	}
	return c;
}

// ---------------------------------------------------------------------------------------------------------------
// gather
// ---------------------------------------------------------------------------------------------------------------

__global__ void s2bFlagActive(ContactView contacts, int contactCount, JointView joints, int jointCap, int* activeFlag,
							  int* jointFlag)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < contactCount)
	{
		activeFlag[i] = S2B_CI_COUNT(contacts.info[i].x) > 0 ? 1 : 0;
	}
	if (i < jointCap)
	{
		jointFlag[i] = (joints.head[i].x & S2B_ROW_VALID) ? 1 : 0;
	}
}

// Conflict endpoints of every item (joints first, then contacts, both in natural order): a body index if the
// constraint can change that body's velocity/position, else -1 (static and kinematic bodies are
// excluded from conflict detection and from write-back).
__global__ void s2bItemEndpoints(const int* counts, const int* jointSlots, const int* activeSlots, JointView joints,
								 ContactView contacts, BodyView bodies, int2* itemBodies, int* degree)
{
	int nJ = counts[CNT_JOINTS], nC = counts[CNT_CONTACTS];
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nJ + nC)
	{
		return;
	}
	int a, b;
	if (i < nJ)
	{
		int4 head = joints.head[jointSlots[i]];
		a = head.y;
		b = head.z;
		if (((head.x >> 1) & 0x7) == S2B_JOINT_MOUSE)
		{
			a = -1; // the mouse joint only acts on body B
		}
	}
	else
	{
		int2 bo = contacts.bodies[activeSlots[i - nJ]];
		a = bo.x;
		b = bo.y;
	}
	if (a >= 0)
	{
		bool movable = bodies.vel[a].w != 0.0f || bodies.prm[a].w != 0.0f;
		a = movable ? a : -1;
	}
	if (b >= 0)
	{
		bool movable = bodies.vel[b].w != 0.0f || bodies.prm[b].w != 0.0f;
		b = movable ? b : -1;
	}
	if (a == b)
	{
		b = -1;
	}
	itemBodies[i] = make_int2(a, b);
	if (degree != nullptr)
	{
		if (a >= 0)
		{
			atomicAdd(degree + a, 1);
		}
		if (b >= 0)
		{
			atomicAdd(degree + b, 1);
		}
	}
}

__global__ void s2bFillAdjacency(const int* counts, const int2* itemBodies, const int* adjStart, int* adjCursor, int* adj)
{
	int n = counts[CNT_JOINTS] + counts[CNT_CONTACTS];
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
	{
		return;
	}
	int2 e = itemBodies[i];
	if (e.x >= 0)
	{
		adj[adjStart[e.x] + atomicAdd(adjCursor + e.x, 1)] = i;
	}
	if (e.y >= 0)
	{
		adj[adjStart[e.y] + atomicAdd(adjCursor + e.y, 1)] = i;
	}
}

// ---------------------------------------------------------------------------------------------------------------
// Jones-Plassmann colouring of the constraint graph (items = nodes, shared movable body = edge)
// ---------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ unsigned s2bPriority(unsigned i)
{
	// integer hash (murmur3 finaliser); ties are broken by the index itself
	unsigned h = i * 0x9E3779B1u + 0x7F4A7C15u;
	h ^= h >> 16;
	h *= 0x85EBCA6Bu;
	h ^= h >> 13;
	h *= 0xC2B2AE35u;
	h ^= h >> 16;
	return h;
}

__device__ __forceinline__ bool s2bHigherPriority(unsigned j, unsigned i)
{
	unsigned pj = s2bPriority(j), pi = s2bPriority(i);
	return pj > pi || (pj == pi && j > i);
}

// One round for item i. colorIn is read-only in the round, colorOut is written (ping-pong) so the result does not
// depend on thread scheduling. Colour codes: -1 uncoloured, 0..maxColors-1, S2B_OVERFLOW_KEY overflow.
__device__ __forceinline__ int s2bColorRound(int i, const int2* itemBodies, const int* adjStart, const int* adj,
											 const int* colorIn, int maxColors)
{
	int c = colorIn[i];
	if (c != -1)
	{
		return c;
	}
	int2 e = itemBodies[i];
	unsigned long long forbidden = 0ull;
	bool isMax = true;
#pragma unroll
	for (int side = 0; side < 2; ++side)
	{
		int body = side == 0 ? e.x : e.y;
		if (body < 0)
		{
			continue;
		}
		int begin = adjStart[body], end = adjStart[body + 1];
		for (int k = begin; k < end; ++k)
		{
			int j = adj[k];
			if (j == i)
			{
				continue;
			}
			int cj = colorIn[j];
			if (cj == -1)
			{
				if (s2bHigherPriority((unsigned)j, (unsigned)i))
				{
					isMax = false;
				}
			}
			else if (cj < S2B_MAX_COLORS)
			{
				forbidden |= 1ull << cj;
			}
		}
	}
	if (isMax == false)
	{
		return -1;
	}
	unsigned long long freeMask = ~forbidden;
	int pick = freeMask == 0ull ? S2B_MAX_COLORS : (__ffsll((long long)freeMask) - 1);
	return pick < maxColors ? pick : S2B_OVERFLOW_KEY;
}

// Cooperative colouring: rounds separated by grid barriers until no item is left uncoloured.
// The "remaining" counter rotates over three slots so that resetting a slot never races with the adds of a round.
__global__ void s2bColorKernel(int* counts, const int2* itemBodies, const int* adjStart, const int* adj, int* colorA,
							   int* colorB, int maxColors)
{
	cg::grid_group grid = cg::this_grid();
	int n = counts[CNT_JOINTS] + counts[CNT_CONTACTS];
	int tid = blockIdx.x * blockDim.x + threadIdx.x;
	int stride = gridDim.x * blockDim.x;
	for (int i = tid; i < n; i += stride)
	{
		colorA[i] = -1;
	}
	grid.sync();
	int* in = colorA;
	int* out = colorB;
	for (int round = 0; round < 8192; ++round)
	{
		int* counter = counts + CNT_REMAINING + (round % 3);
		int localRemaining = 0;
		for (int i = tid; i < n; i += stride)
		{
			int c = s2bColorRound(i, itemBodies, adjStart, adj, in, maxColors);
			out[i] = c;
			localRemaining += (c == -1) ? 1 : 0;
		}
		if (localRemaining > 0)
		{
			atomicAdd(counter, localRemaining);
		}
		grid.sync();
		int remaining = *((volatile int*)counter);
		if (tid == 0)
		{
			counts[CNT_REMAINING + ((round + 2) % 3)] = 0;
			counts[CNT_ROUNDS] = round + 1;
		}
		int* tmp = in;
		in = out;
		out = tmp;
		if (remaining == 0)
		{
			break;
		}
	}
	// make colorA the final array
	if (in != colorA)
	{
		for (int i = tid; i < n; i += stride)
		{
			colorA[i] = in[i];
		}
	}
}

// sort keys: colour per item, split into the joint and the contact key arrays (values = natural index)
__global__ void s2bMakeSortKeys(const int* counts, const int* color, unsigned char* jKeys, int* jVals, unsigned char* cKeys,
								int* cVals)
{
	int nJ = counts[CNT_JOINTS], nC = counts[CNT_CONTACTS];
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nJ + nC)
	{
		return;
	}
	unsigned char key = (unsigned char)color[i];
	if (i < nJ)
	{
		jKeys[i] = key;
		jVals[i] = i;
	}
	else
	{
		cKeys[i - nJ] = key;
		cVals[i - nJ] = i - nJ;
	}
}

// group offsets from sorted colour keys: off[c] = first position with key >= c for c in [0, 64]; off[65] = n.
// Group 64 is the serial overflow group.
__global__ void s2bGroupOffsets(const int* counts, int which, const unsigned char* sortedKeys, int* off)
{
	int n = counts[which];
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > n)
	{
		return;
	}
	int prev = i == 0 ? -1 : (sortedKeys[i - 1] == S2B_OVERFLOW_KEY ? S2B_MAX_COLORS : sortedKeys[i - 1]);
	int cur = i == n ? S2B_MAX_COLORS + 1 : (sortedKeys[i] == S2B_OVERFLOW_KEY ? S2B_MAX_COLORS : sortedKeys[i]);
	for (int c = prev + 1; c <= cur; ++c)
	{
		off[c] = i;
	}
}

__global__ void s2bFinishGroups(int* counts, const int* cOff, const int* jOff)
{
	// number of colour groups actually used (largest non-empty colour + 1)
	int groups = 0;
	for (int c = 0; c < S2B_MAX_COLORS; ++c)
	{
		if (cOff[c + 1] > cOff[c] || jOff[c + 1] > jOff[c])
		{
			groups = c + 1;
		}
	}
	counts[CNT_GROUPS] = groups;
	counts[CNT_OVERFLOW_C] = cOff[S2B_MAX_COLORS + 1] - cOff[S2B_MAX_COLORS];
	counts[CNT_OVERFLOW_J] = jOff[S2B_MAX_COLORS + 1] - jOff[S2B_MAX_COLORS];
}

// src[t] = contact slot of the constraint at solve position t
__global__ void s2bBuildSources(const int* counts, const int* cPerm, const int* activeSlots, int* src)
{
	int n = counts[CNT_CONTACTS];
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t < n)
	{
		src[t] = activeSlots[cPerm[t]];
	}
}

// ---------------------------------------------------------------------------------------------------------------
// passes: ops applied to ranges of bodies / contact constraints / joint constraints
// ---------------------------------------------------------------------------------------------------------------

enum PassOp
{
	// body ops
	OP_INTEGRATE_VELOCITIES,
	OP_INTEGRATE_POSITIONS,
	OP_FINALIZE_POSITIONS,
	// contact ops
	OP_PREPARE_SOFT,
	OP_PREPARE_PGS,
	OP_WARM_START,
	OP_SOLVE_TGS_SOFT_BIAS,
	OP_SOLVE_TGS_SOFT_RELAX,
	OP_STORE,
};

template <int OP> __device__ __forceinline__ void s2bBodyOp(const SolveArgs& a, int i)
{
	if (OP == OP_INTEGRATE_VELOCITIES)
	{
		s2bIntegrateVelocity(a, i, a.ctx.h);
	}
	else if (OP == OP_INTEGRATE_POSITIONS)
	{
		s2bIntegratePosition(a, i, a.ctx.h);
	}
	else if (OP == OP_FINALIZE_POSITIONS)
	{
		s2bFinalizePosition(a, i);
	}
}

template <int OP> __device__ __forceinline__ void s2bContactOp(const SolveArgs& a, int t)
{
	if (OP == OP_PREPARE_SOFT)
	{
		s2bPrepareContact<PREPARE_SOFT>(a, t);
	}
	else if (OP == OP_PREPARE_PGS)
	{
		s2bPrepareContact<PREPARE_PGS>(a, t);
	}
	else if (OP == OP_WARM_START)
	{
		s2bWarmStartContact(a, t);
	}
	else if (OP == OP_SOLVE_TGS_SOFT_BIAS)
	{
		s2bSolveContactTgsSoft(a, t, a.ctx.inv_h, true);
	}
	else if (OP == OP_SOLVE_TGS_SOFT_RELAX)
	{
		s2bSolveContactTgsSoft(a, t, a.ctx.inv_h, false);
	}
	else if (OP == OP_STORE)
	{
		s2bStoreContactImpulses(a, t, 1.0f);
	}
}

enum JointOp
{
	JOP_NONE,
	JOP_PREPARE_SOFT_WARM,	// s2PrepareJoint_Soft(..., warmStart = true)
	JOP_PREPARE_SOFT_COLD,	// s2PrepareJoint_Soft(..., warmStart = false)
	JOP_PREPARE_RIGID_FLAG, // s2PrepareJoint(..., context->warmStart)
	JOP_WARM_START,
	JOP_SOLVE_SOFT_BIAS,
	JOP_SOLVE_SOFT_RELAX,
	JOP_SOLVE_BAUMGARTE_BIAS,
	JOP_SOLVE_BAUMGARTE_RELAX,
	JOP_STORE,
};

template <int JOP> __device__ __forceinline__ void s2bJointOp(const SolveArgs& a, int t, const int* jointSlots, const int* jPerm)
{
	if (JOP == JOP_PREPARE_SOFT_WARM)
	{
		s2bPrepareJoint<JPREP_SOFT>(a, t, jointSlots[jPerm[t]], true);
	}
	else if (JOP == JOP_PREPARE_SOFT_COLD)
	{
		s2bPrepareJoint<JPREP_SOFT>(a, t, jointSlots[jPerm[t]], false);
	}
	else if (JOP == JOP_PREPARE_RIGID_FLAG)
	{
		s2bPrepareJoint<JPREP_RIGID>(a, t, jointSlots[jPerm[t]], a.ctx.warmStart != 0);
	}
	else if (JOP == JOP_WARM_START)
	{
		s2bWarmStartJoint(a, t);
	}
	else if (JOP == JOP_SOLVE_SOFT_BIAS)
	{
		s2bSolveJointSoft(a, t, a.ctx.h, a.ctx.inv_h, true);
	}
	else if (JOP == JOP_SOLVE_SOFT_RELAX)
	{
		s2bSolveJointSoft(a, t, a.ctx.h, a.ctx.inv_h, false);
	}
	else if (JOP == JOP_SOLVE_BAUMGARTE_BIAS)
	{
		s2bSolveJointBaumgarte(a, t, a.ctx.h, a.ctx.inv_h, true);
	}
	else if (JOP == JOP_SOLVE_BAUMGARTE_RELAX)
	{
		s2bSolveJointBaumgarte(a, t, a.ctx.h, a.ctx.inv_h, false);
	}
	else if (JOP == JOP_STORE)
	{
		s2bStoreJointImpulses(a, t);
	}
}

struct PassPtrs
{
	const int* jointSlots;
	const int* jPerm;
};

// ---- multi-launch kernels -------------------------------------------------------------------------------------

template <int OP> __global__ void s2bBodyPassKernel(SolveArgs a)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < a.bodies.capacity)
	{
		s2bBodyOp<OP>(a, i);
	}
}

// one group (or the whole range when begin/end span everything): joints first, then contacts
template <int JOP, int OP> __global__ void s2bRangePassKernel(SolveArgs a, PassPtrs p, int jBegin, int jEnd, int cBegin, int cEnd)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	int nj = jEnd - jBegin;
	if (t < nj)
	{
		if (JOP != JOP_NONE)
		{
			s2bJointOp<JOP>(a, jBegin + t, p.jointSlots, p.jPerm);
		}
	}
	else if (t - nj < cEnd - cBegin)
	{
		s2bContactOp<OP>(a, cBegin + (t - nj));
	}
}

// serial overflow group: one thread walks the items in order
template <int JOP, int OP> __global__ void s2bSerialPassKernel(SolveArgs a, PassPtrs p, int jBegin, int jEnd, int cBegin, int cEnd)
{
	if (blockIdx.x == 0 && threadIdx.x == 0)
	{
		if (JOP != JOP_NONE)
		{
			for (int t = jBegin; t < jEnd; ++t)
			{
				s2bJointOp<JOP>(a, t, p.jointSlots, p.jPerm);
			}
		}
		for (int t = cBegin; t < cEnd; ++t)
		{
			s2bContactOp<OP>(a, t);
		}
	}
}

// ---- persistent kernel helpers --------------------------------------------------------------------------------

template <int OP> __device__ __forceinline__ void s2bGridBodyPass(const SolveArgs& a)
{
	int stride = gridDim.x * blockDim.x;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.bodies.capacity; i += stride)
	{
		s2bBodyOp<OP>(a, i);
	}
}

// whole-range pass (prepare / store): no ordering constraints between items
template <int JOP, int OP> __device__ __forceinline__ void s2bGridFlatPass(const SolveArgs& a, const PassPtrs& p)
{
	int nJ = a.counts[CNT_JOINTS], nC = a.counts[CNT_CONTACTS];
	int stride = gridDim.x * blockDim.x;
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nJ + nC; t += stride)
	{
		if (t < nJ)
		{
			if (JOP != JOP_NONE)
			{
				s2bJointOp<JOP>(a, t, p.jointSlots, p.jPerm);
			}
		}
		else
		{
			s2bContactOp<OP>(a, t - nJ);
		}
	}
}

// Gauss-Seidel pass: groups in order with a grid barrier after each, then the serial overflow group
template <int JOP, int OP> __device__ __forceinline__ void s2bGridGroupPass(const SolveArgs& a, const PassPtrs& p, cg::grid_group& grid)
{
	int groups = a.counts[CNT_GROUPS];
	int stride = gridDim.x * blockDim.x;
	int tid = blockIdx.x * blockDim.x + threadIdx.x;
	for (int g = 0; g < groups; ++g)
	{
		int jBegin = a.jGroupOff[g], jEnd = a.jGroupOff[g + 1];
		int cBegin = a.cGroupOff[g], cEnd = a.cGroupOff[g + 1];
		int nj = jEnd - jBegin, n = nj + (cEnd - cBegin);
		for (int t = tid; t < n; t += stride)
		{
			if (t < nj)
			{
				if (JOP != JOP_NONE)
				{
					s2bJointOp<JOP>(a, jBegin + t, p.jointSlots, p.jPerm);
				}
			}
			else
			{
				s2bContactOp<OP>(a, cBegin + (t - nj));
			}
		}
		grid.sync();
	}
	int ovC = a.counts[CNT_OVERFLOW_C], ovJ = a.counts[CNT_OVERFLOW_J];
	if (ovC + ovJ > 0)
	{
		if (tid == 0)
		{
			int jBegin = a.jGroupOff[S2B_MAX_COLORS], cBegin = a.cGroupOff[S2B_MAX_COLORS];
			if (JOP != JOP_NONE)
			{
				for (int t = 0; t < ovJ; ++t)
				{
					s2bJointOp<JOP>(a, jBegin + t, p.jointSlots, p.jPerm);
				}
			}
			for (int t = 0; t < ovC; ++t)
			{
				s2bContactOp<OP>(a, cBegin + t);
			}
		}
		grid.sync();
	}
}

// ---------------------------------------------------------------------------------------------------------------
// s2Solve_TGS_Soft (reference src/solve_tgs_soft.c:138-280) as one persistent cooperative kernel
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(S2B_BLOCK) s2bPersistentTgsSoft(SolveArgs a, PassPtrs p)
{
	cg::grid_group grid = cg::this_grid();

	// prepare (joints always warm start here: reference solve_tgs_soft.c:204-205)
	s2bGridFlatPass<JOP_PREPARE_SOFT_WARM, OP_PREPARE_SOFT>(a, p);
	grid.sync();

	int substeps = a.ctx.iterations;
	for (int s = 0; s < substeps; ++s)
	{
		s2bGridBodyPass<OP_INTEGRATE_VELOCITIES>(a);
		grid.sync();
		if (a.ctx.warmStart)
		{
			s2bGridGroupPass<JOP_WARM_START, OP_WARM_START>(a, p, grid);
		}
		s2bGridGroupPass<JOP_SOLVE_SOFT_BIAS, OP_SOLVE_TGS_SOFT_BIAS>(a, p, grid);
		s2bGridBodyPass<OP_INTEGRATE_POSITIONS>(a);
		grid.sync();
		if (a.ctx.extraIterations > 0)
		{
			s2bGridGroupPass<JOP_SOLVE_SOFT_RELAX, OP_SOLVE_TGS_SOFT_RELAX>(a, p, grid);
		}
	}

	s2bGridBodyPass<OP_FINALIZE_POSITIONS>(a);
	s2bGridFlatPass<JOP_STORE, OP_STORE>(a, p);
}

__global__ void s2bMeterWork(const int* counts, int passes, unsigned long long* work)
{
	work[0] += (unsigned long long)(counts[CNT_CONTACTS] + counts[CNT_JOINTS]) * (unsigned long long)passes;
	work[1] += 1ull;
}

// ---------------------------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------------------------

static SoftCoef makeSoft(float h, float hertz, float zeta)
{
	// reference src/solve_common.c:264-271 (contacts), src/revolute_joint.c:470-476 (joints)
	SoftCoef c;
	float omega = 2.0f * s2_pi * hertz;
	float cc = h * omega * (2.0f * zeta + h * omega);
	c.bias = omega / (2.0f * zeta + h * omega);
	c.impulse = 1.0f / (1.0f + cc);
	c.mass = cc * c.impulse;
	return c;
}

struct HostPlan
{
	// multi-launch mode only
	int groups = 0;
	std::vector<int> cOff, jOff;
};

template <int OP> static void launchBodyPass(s2bWorld* w, const SolveArgs& a)
{
	if (a.bodies.capacity > 0)
	{
		auto kernel = s2bBodyPassKernel<OP>;
		S2B_LAUNCH(w, kernel, gridFor(a.bodies.capacity, S2B_BLOCK), S2B_BLOCK, 0, a);
	}
}

template <int JOP, int OP> static void launchFlatPass(s2bWorld* w, const SolveArgs& a, const PassPtrs& p, int nJ, int nC)
{
	if (nJ + nC > 0)
	{
		auto kernel = s2bRangePassKernel<JOP, OP>;
		S2B_LAUNCH(w, kernel, gridFor(nJ + nC, S2B_BLOCK), S2B_BLOCK, 0, a, p, 0, nJ, 0, nC);
	}
}

template <int JOP, int OP> static void launchGroupPass(s2bWorld* w, const SolveArgs& a, const PassPtrs& p, const HostPlan& plan)
{
	for (int g = 0; g < plan.groups; ++g)
	{
		int jb = plan.jOff[g], je = plan.jOff[g + 1], cb = plan.cOff[g], ce = plan.cOff[g + 1];
		int n = (je - jb) + (ce - cb);
		if (n > 0)
		{
			auto kernel = s2bRangePassKernel<JOP, OP>;
			S2B_LAUNCH(w, kernel, gridFor(n, S2B_BLOCK), S2B_BLOCK, 0, a, p, jb, je, cb, ce);
		}
	}
	int G = (int)plan.cOff.size() - 2; // index of the overflow group
	int jb = plan.jOff[G], je = plan.jOff[G + 1], cb = plan.cOff[G], ce = plan.cOff[G + 1];
	if ((je - jb) + (ce - cb) > 0)
	{
		auto kernel = s2bSerialPassKernel<JOP, OP>;
		S2B_LAUNCH(w, kernel, 1, 32, 0, a, p, jb, je, cb, ce);
	}
}

static void runTgsSoftMultiLaunch(s2bWorld* w, const SolveArgs& a, const PassPtrs& p, const HostPlan& plan, int nJ, int nC)
{
	launchFlatPass<JOP_PREPARE_SOFT_WARM, OP_PREPARE_SOFT>(w, a, p, nJ, nC);
	for (int s = 0; s < a.ctx.iterations; ++s)
	{
		launchBodyPass<OP_INTEGRATE_VELOCITIES>(w, a);
		if (a.ctx.warmStart)
		{
			launchGroupPass<JOP_WARM_START, OP_WARM_START>(w, a, p, plan);
		}
		launchGroupPass<JOP_SOLVE_SOFT_BIAS, OP_SOLVE_TGS_SOFT_BIAS>(w, a, p, plan);
		launchBodyPass<OP_INTEGRATE_POSITIONS>(w, a);
		if (a.ctx.extraIterations > 0)
		{
			launchGroupPass<JOP_SOLVE_SOFT_RELAX, OP_SOLVE_TGS_SOFT_RELAX>(w, a, p, plan);
		}
	}
	launchBodyPass<OP_FINALIZE_POSITIONS>(w, a);
	launchFlatPass<JOP_STORE, OP_STORE>(w, a, p, nJ, nC);
}

// wavefront levels on the host (validation schedule): level(i) = 1 + max level of earlier items sharing a movable body
static void buildWavefront(s2bWorld* w, SolverScratch* s, int nJ, int nC, HostPlan& plan)
{
	cudaStream_t st = w->stream;
	int n = nJ + nC;
	std::vector<int2> ends((size_t)std::max(n, 1));
	std::vector<int> activeSlots((size_t)std::max(nC, 1));
	std::vector<unsigned long long> keys;
	S2B_CHECK(cudaMemcpyAsync(ends.data(), s->itemBodies.p, sizeof(int2) * (size_t)n, cudaMemcpyDeviceToHost, st));
	if (nC > 0)
	{
		S2B_CHECK(cudaMemcpyAsync(activeSlots.data(), s->activeSlots.p, sizeof(int) * (size_t)nC, cudaMemcpyDeviceToHost, st));
	}
	bool hinted = w->orderHint.empty() == false && nC > 0;
	if (hinted)
	{
		keys.resize((size_t)w->contactCount);
		S2B_CHECK(cudaMemcpyAsync(keys.data(), w->contacts[w->cur].key.p, sizeof(unsigned long long) * keys.size(),
								  cudaMemcpyDeviceToHost, st));
	}
	S2B_CHECK(cudaStreamSynchronize(st));

	// sequential order of the contact constraints: natural (slot) order, or the imposed order
	std::vector<int> order((size_t)nC);
	for (int i = 0; i < nC; ++i)
	{
		order[i] = i;
	}
	if (hinted)
	{
		std::vector<std::pair<unsigned long long, int>> hint(w->orderHint.size());
		for (size_t k = 0; k < hint.size(); ++k)
		{
			hint[k] = {w->orderHint[k], (int)k};
		}
		std::sort(hint.begin(), hint.end());
		std::vector<long long> rank((size_t)nC);
		for (int i = 0; i < nC; ++i)
		{
			unsigned long long key = keys[(size_t)activeSlots[i]];
			auto it = std::lower_bound(hint.begin(), hint.end(), std::make_pair(key, -1));
			rank[i] = (it != hint.end() && it->first == key) ? (long long)it->second : (long long)hint.size() + i;
		}
		std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return rank[x] < rank[y]; });
	}

	std::vector<int> lastLevel((size_t)std::max(w->bodyCap, 1), 0);
	std::vector<int> level((size_t)std::max(n, 1), 0);
	int maxLevel = 0;
	auto place = [&](int item) {
		int2 e = ends[(size_t)item];
		int l = 0;
		if (e.x >= 0)
		{
			l = std::max(l, lastLevel[e.x]);
		}
		if (e.y >= 0)
		{
			l = std::max(l, lastLevel[e.y]);
		}
		l += 1;
		level[item] = l;
		if (e.x >= 0)
		{
			lastLevel[e.x] = l;
		}
		if (e.y >= 0)
		{
			lastLevel[e.y] = l;
		}
		maxLevel = std::max(maxLevel, l);
	};
	for (int i = 0; i < nJ; ++i)
	{
		place(i);
	}
	for (int k = 0; k < nC; ++k)
	{
		place(nJ + order[k]);
	}

	int G = maxLevel;
	plan.groups = G;
	plan.cOff.assign((size_t)G + 2, 0);
	plan.jOff.assign((size_t)G + 2, 0);
	for (int i = 0; i < nJ; ++i)
	{
		plan.jOff[(size_t)level[i]] += 1; // level l (1-based) stored at index l, shifted below
	}
	for (int i = 0; i < nC; ++i)
	{
		plan.cOff[(size_t)level[nJ + i]] += 1;
	}
	// exclusive scan with the 1-based shift: off[g] = number of items with level <= g
	for (int g = 1; g <= G; ++g)
	{
		plan.jOff[g] += plan.jOff[g - 1];
		plan.cOff[g] += plan.cOff[g - 1];
	}
	plan.jOff[G + 1] = plan.jOff[G];
	plan.cOff[G + 1] = plan.cOff[G];
	// now off[g-1]..off[g] is the range of level g  (off[0] = 0)
	std::vector<int> jPerm((size_t)std::max(nJ, 1)), cPerm((size_t)std::max(nC, 1));
	{
		std::vector<int> jc(plan.jOff.begin(), plan.jOff.end()), ccur(plan.cOff.begin(), plan.cOff.end());
		for (int i = 0; i < nJ; ++i)
		{
			jPerm[(size_t)jc[(size_t)level[i] - 1]++] = i;
		}
		for (int k = 0; k < nC; ++k)
		{
			int i = order[k];
			cPerm[(size_t)ccur[(size_t)level[nJ + i] - 1]++] = i;
		}
	}

	s->cGroupOff.reserve((size_t)G + 2, st, false);
	s->jGroupOff.reserve((size_t)G + 2, st, false);
	S2B_CHECK(cudaMemcpyAsync(s->cGroupOff.p, plan.cOff.data(), sizeof(int) * ((size_t)G + 2), cudaMemcpyHostToDevice, st));
	S2B_CHECK(cudaMemcpyAsync(s->jGroupOff.p, plan.jOff.data(), sizeof(int) * ((size_t)G + 2), cudaMemcpyHostToDevice, st));
	if (nC > 0)
	{
		S2B_CHECK(cudaMemcpyAsync(s->cPerm.p, cPerm.data(), sizeof(int) * (size_t)nC, cudaMemcpyHostToDevice, st));
	}
	if (nJ > 0)
	{
		S2B_CHECK(cudaMemcpyAsync(s->jPerm.p, jPerm.data(), sizeof(int) * (size_t)nJ, cudaMemcpyHostToDevice, st));
	}
	int counts[CNT_SIZE] = {0};
	counts[CNT_CONTACTS] = nC;
	counts[CNT_JOINTS] = nJ;
	counts[CNT_GROUPS] = G;
	S2B_CHECK(cudaMemcpyAsync(s->counts.p, counts, sizeof(counts), cudaMemcpyHostToDevice, st));
	S2B_CHECK(cudaStreamSynchronize(st));
}

void s2bSolve(s2bWorld* w, int solverType, const s2bStepContext* ctxIn)
{
	SolverScratch* s = s2bGetSolverScratch(w);
	cudaStream_t st = w->stream;
	s2bStepContext ctx = *ctxIn;

	if (solverType != 7)
	{
		fprintf(stderr, "solver2d-b200: solver type %d is not implemented on the device yet\n", solverType);
		abort();
	}

	int contactCount = w->contactCount;
	int jointCap = w->jointCap;
	int bodyCap = w->bodyCap;
	int maxItems = contactCount + jointCap;
	VariantColumns cols = columnsFor(solverType);

	// ---- reserve scratch (sizes are upper bounds known on the host: no synchronisation) ----
	size_t nC = (size_t)std::max(contactCount, 1), nJ = (size_t)std::max(jointCap, 1), nI = (size_t)std::max(maxItems, 1);
	s->counts.reserve(CNT_SIZE, st, false);
	s->activeFlag.reserve(nC, st, false);
	s->activeSlots.reserve(nC, st, false);
	s->jointFlag.reserve(nJ, st, false);
	s->jointSlots.reserve(nJ, st, false);
	s->itemBodies.reserve(nI, st, false);
	s->degree.reserve((size_t)bodyCap + 1, st, false);
	s->adjStart.reserve((size_t)bodyCap + 2, st, false);
	s->adjCursor.reserve((size_t)bodyCap + 1, st, false);
	s->adj.reserve(2 * nI, st, false);
	s->colorA.reserve(nI, st, false);
	s->colorB.reserve(nI, st, false);
	s->sortKeyIn.reserve(2 * nI, st, false);
	s->sortKeyOut.reserve(2 * nI, st, false);
	s->sortValIn.reserve(2 * nI, st, false);
	s->sortValOut.reserve(2 * nI, st, false);
	s->cGroupOff.reserve(S2B_MAX_COLORS + 2, st, false);
	s->jGroupOff.reserve(S2B_MAX_COLORS + 2, st, false);
	s->cPerm.reserve(nC, st, false);
	s->jPerm.reserve(nJ, st, false);
	s->idx.reserve(nC, st, false);
	s->nf.reserve(nC, st, false);
	s->src.reserve(nC, st, false);
	for (int p = 0; p < 2; ++p)
	{
		s->anchor[p].reserve(nC, st, false);
		s->pm[p].reserve(nC, st, false);
		s->lambda[p].reserve(nC, st, false);
		if (cols.r0)
		{
			s->r0[p].reserve(nC, st, false);
		}
		if (cols.sep)
		{
			s->sep[p].reserve(nC, st, false);
		}
		if (cols.sticky)
		{
			s->fanchor[p].reserve(nC, st, false);
			s->tsep[p].reserve(nC, st, false);
		}
	}
	s->jhead.reserve(nJ, st, false);
	s->janchor.reserve(nJ, st, false);
	s->jmass.reserve(nJ, st, false);
	s->jd0ax.reserve(nJ, st, false);
	s->jlim.reserve(nJ, st, false);
	s->jmotor.reserve(nJ, st, false);
	s->jcoef.reserve(nJ, st, false);
	s->jpivot.reserve(nJ, st, false);
	s->jimp.reserve(nJ, st, false);
	s->jlimp.reserve(nJ, st, false);

	// cub temp storage (compaction, scan, 8-bit sort) sized for the largest use
	size_t tempBytes = 0, need = 0;
	cub::DeviceSelect::Flagged(nullptr, need, thrust::counting_iterator<int>(0), (int*)nullptr, (int*)nullptr, (int*)nullptr,
							   (int)nI, st);
	tempBytes = std::max(tempBytes, need);
	cub::DeviceScan::ExclusiveSum(nullptr, need, (int*)nullptr, (int*)nullptr, bodyCap + 1, st);
	tempBytes = std::max(tempBytes, need);
	cub::DeviceRadixSort::SortPairs(nullptr, need, (unsigned char*)nullptr, (unsigned char*)nullptr, (int*)nullptr,
									(int*)nullptr, (int)nI, 0, 8, st);
	tempBytes = std::max(tempBytes, need);
	s->cubTemp.reserve(tempBytes + 256, st, false, false);

	// ---- gather ----
	S2B_CHECK(cudaMemsetAsync(s->counts.p, 0, sizeof(int) * CNT_SIZE, st));
	{
		int n = std::max(contactCount, jointCap);
		if (n > 0)
		{
			S2B_LAUNCH(w, s2bFlagActive, gridFor(n, 256), 256, 0, makeView(w->contacts[w->cur]), contactCount, jointView(w),
					   jointCap, s->activeFlag.p, s->jointFlag.p);
		}
		size_t tb = s->cubTemp.cap;
		if (contactCount > 0)
		{
			cub::DeviceSelect::Flagged(s->cubTemp.p, tb, thrust::counting_iterator<int>(0), s->activeFlag.p, s->activeSlots.p,
									   s->counts.p + CNT_CONTACTS, contactCount, st);
			w->kernelLaunches += 2;
		}
		if (jointCap > 0)
		{
			tb = s->cubTemp.cap;
			cub::DeviceSelect::Flagged(s->cubTemp.p, tb, thrust::counting_iterator<int>(0), s->jointFlag.p, s->jointSlots.p,
									   s->counts.p + CNT_JOINTS, jointCap, st);
			w->kernelLaunches += 2;
		}
	}

	// ---- argument block ----
	SolveArgs a;
	memset(&a, 0, sizeof(a));
	a.bodies = bodyView(w);
	a.contacts = makeView(w->contacts[w->cur]);
	a.joints = jointView(w);
	a.cc.idx = s->idx.p;
	a.cc.nf = s->nf.p;
	a.cc.src = s->src.p;
	for (int p = 0; p < 2; ++p)
	{
		a.cc.anchor[p] = s->anchor[p].p;
		a.cc.pm[p] = s->pm[p].p;
		a.cc.lambda[p] = s->lambda[p].p;
		a.cc.r0[p] = cols.r0 ? s->r0[p].p : nullptr;
		a.cc.sep[p] = cols.sep ? s->sep[p].p : nullptr;
		a.cc.fanchor[p] = cols.sticky ? s->fanchor[p].p : nullptr;
		a.cc.tsep[p] = cols.sticky ? s->tsep[p].p : nullptr;
	}
	a.jc.head = s->jhead.p;
	a.jc.anchor = s->janchor.p;
	a.jc.mass = s->jmass.p;
	a.jc.d0ax = s->jd0ax.p;
	a.jc.lim = s->jlim.p;
	a.jc.motor = s->jmotor.p;
	a.jc.coef = s->jcoef.p;
	a.jc.pivot = s->jpivot.p;
	a.jc.imp = s->jimp.p;
	a.jc.limp = s->jlimp.p;
	a.counts = s->counts.p;
	a.ctx = ctx;
	a.gravity = w->gravity;
	a.solverType = solverType;
	a.sticky = w->sticky ? 1 : 0;

	// hertz clamps of the variant (reference src/solve_tgs_soft.c:185-186)
	float contactHertz = S2_MIN(s2_contactHertz, 0.25f * ctx.inv_h);
	float jointHertz = S2_MIN(s2_jointHertz, 0.125f * ctx.inv_h);
	a.contactHertz = contactHertz;
	a.jointHertz = jointHertz;
	a.softDynamic = makeSoft(ctx.h, contactHertz, 10.0f);
	a.softStatic = makeSoft(ctx.h, 2.0f * contactHertz, 10.0f);
	a.softJoint = makeSoft(ctx.h, jointHertz, 10.0f);

	PassPtrs pp = {s->jointSlots.p, s->jPerm.p};

	// ---- schedule ----
	HostPlan plan;
	bool needHostCounts = (w->schedule == S2B_SCHEDULE_WAVEFRONT) || (w->persistent == 0) || (w->coopSupported == 0);
	int hostNC = 0, hostNJ = 0;

	if (maxItems > 0)
	{
		bool wantAdj = w->schedule == S2B_SCHEDULE_COLOR;
		if (wantAdj)
		{
			S2B_CHECK(cudaMemsetAsync(s->degree.p, 0, sizeof(int) * ((size_t)bodyCap + 1), st));
			S2B_CHECK(cudaMemsetAsync(s->adjCursor.p, 0, sizeof(int) * ((size_t)bodyCap + 1), st));
		}
		S2B_LAUNCH(w, s2bItemEndpoints, gridFor(maxItems, 256), 256, 0, s->counts.p, s->jointSlots.p, s->activeSlots.p,
				   jointView(w), makeView(w->contacts[w->cur]), bodyView(w), s->itemBodies.p, wantAdj ? s->degree.p : nullptr);

		if (w->schedule == S2B_SCHEDULE_COLOR)
		{
			size_t tb = s->cubTemp.cap;
			cub::DeviceScan::ExclusiveSum(s->cubTemp.p, tb, s->degree.p, s->adjStart.p, bodyCap + 1, st);
			w->kernelLaunches += 2;
			S2B_LAUNCH(w, s2bFillAdjacency, gridFor(maxItems, 256), 256, 0, s->counts.p, s->itemBodies.p, s->adjStart.p,
					   s->adjCursor.p, s->adj.p);

			if (w->coopSupported)
			{
				int blocksPerSm = 0;
				S2B_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, s2bColorKernel, 256, 0));
				int grid = std::min(w->smCount * std::max(blocksPerSm, 1), std::max(1, gridFor(maxItems, 256)));
				int* countsPtr = s->counts.p;
				const int2* ib = s->itemBodies.p;
				const int* as = s->adjStart.p;
				const int* ad = s->adj.p;
				int* ca = s->colorA.p;
				int* cb = s->colorB.p;
				int mc = w->maxColors;
				void* args[] = {&countsPtr, &ib, &as, &ad, &ca, &cb, &mc};
				S2B_CHECK(cudaLaunchCooperativeKernel((void*)s2bColorKernel, dim3(grid), dim3(256), args, 0, st));
				w->kernelLaunches += 1;
			}
			else
			{
				fprintf(stderr, "solver2d-b200: cooperative launch unsupported on this device\n");
				abort();
			}

			// colour-major order: 8-bit stable radix sort of (colour, natural index), joints and contacts separately
			unsigned char* jKeysIn = s->sortKeyIn.p;
			unsigned char* cKeysIn = s->sortKeyIn.p + nI;
			unsigned char* jKeysOut = s->sortKeyOut.p;
			unsigned char* cKeysOut = s->sortKeyOut.p + nI;
			int* jValsIn = s->sortValIn.p;
			int* cValsIn = s->sortValIn.p + nI;
			// entries beyond the live counts keep key 255: they sort behind every live entry and are never read
			S2B_CHECK(cudaMemsetAsync(s->sortKeyIn.p, 0xFF, 2 * nI, st));
			S2B_LAUNCH(w, s2bMakeSortKeys, gridFor(maxItems, 256), 256, 0, s->counts.p, s->colorA.p, jKeysIn, jValsIn, cKeysIn,
					   cValsIn);
			// the sorts run over the host-known upper-bound sizes, so no device count has to be read back
			if (contactCount > 0)
			{
				tb = s->cubTemp.cap;
				cub::DeviceRadixSort::SortPairs(s->cubTemp.p, tb, cKeysIn, cKeysOut, cValsIn, s->cPerm.p, contactCount, 0, 8, st);
				w->kernelLaunches += 3;
			}
			if (jointCap > 0)
			{
				tb = s->cubTemp.cap;
				cub::DeviceRadixSort::SortPairs(s->cubTemp.p, tb, jKeysIn, jKeysOut, jValsIn, s->jPerm.p, jointCap, 0, 8, st);
				w->kernelLaunches += 3;
			}
			S2B_LAUNCH(w, s2bGroupOffsets, gridFor(contactCount + 1, 256), 256, 0, s->counts.p, (int)CNT_CONTACTS, cKeysOut,
					   s->cGroupOff.p);
			S2B_LAUNCH(w, s2bGroupOffsets, gridFor(jointCap + 1, 256), 256, 0, s->counts.p, (int)CNT_JOINTS, jKeysOut,
					   s->jGroupOff.p);
			S2B_LAUNCH(w, s2bFinishGroups, 1, 1, 0, s->counts.p, s->cGroupOff.p, s->jGroupOff.p);
		}

		if (needHostCounts)
		{
			int hostCounts[CNT_SIZE];
			S2B_CHECK(cudaMemcpyAsync(hostCounts, s->counts.p, sizeof(hostCounts), cudaMemcpyDeviceToHost, st));
			S2B_CHECK(cudaStreamSynchronize(st));
			hostNC = hostCounts[CNT_CONTACTS];
			hostNJ = hostCounts[CNT_JOINTS];
			if (w->schedule == S2B_SCHEDULE_WAVEFRONT)
			{
				buildWavefront(w, s, hostNJ, hostNC, plan);
			}
			else
			{
				plan.groups = hostCounts[CNT_GROUPS];
				plan.cOff.resize(S2B_MAX_COLORS + 2);
				plan.jOff.resize(S2B_MAX_COLORS + 2);
				S2B_CHECK(cudaMemcpy(plan.cOff.data(), s->cGroupOff.p, sizeof(int) * (S2B_MAX_COLORS + 2), cudaMemcpyDeviceToHost));
				S2B_CHECK(cudaMemcpy(plan.jOff.data(), s->jGroupOff.p, sizeof(int) * (S2B_MAX_COLORS + 2), cudaMemcpyDeviceToHost));
			}
			s->hostContacts = hostNC;
			s->hostJoints = hostNJ;
			s->hostGroups = plan.groups;
			s->hostCGroupOff = plan.cOff;
			s->hostJGroupOff = plan.jOff;
			s->hostCountsValid = true;
		}
		else
		{
			s->hostCountsValid = false;
		}

		if (contactCount > 0)
		{
			S2B_LAUNCH(w, s2bBuildSources, gridFor(contactCount, 256), 256, 0, s->counts.p, s->cPerm.p, s->activeSlots.p,
					   s->src.p);
		}
	}
	else
	{
		// no constraints at all: bodies still integrate
		plan.groups = 0;
		plan.cOff.assign(S2B_MAX_COLORS + 2, 0);
		plan.jOff.assign(S2B_MAX_COLORS + 2, 0);
		S2B_CHECK(cudaMemsetAsync(s->cGroupOff.p, 0, sizeof(int) * (S2B_MAX_COLORS + 2), st));
		S2B_CHECK(cudaMemsetAsync(s->jGroupOff.p, 0, sizeof(int) * (S2B_MAX_COLORS + 2), st));
		needHostCounts = true;
		s->hostContacts = s->hostJoints = s->hostGroups = 0;
		s->hostCGroupOff = plan.cOff;
		s->hostJGroupOff = plan.jOff;
		s->hostCountsValid = true;
	}

	// ---- iterate ----
	// (the group tables may have been re-allocated by the schedule step: take the pointers now)
	a.cGroupOff = s->cGroupOff.p;
	a.jGroupOff = s->jGroupOff.p;
	pp.jPerm = s->jPerm.p;
	bool usePersistent = w->persistent != 0 && w->coopSupported != 0;
	if (usePersistent)
	{
		// wavefront tables hold CNT_GROUPS levels and no overflow group; the overflow counts are zero so the kernel
		// never indexes past them
		int blocksPerSm = 0;
		S2B_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, s2bPersistentTgsSoft, S2B_BLOCK, 0));
		blocksPerSm = std::min(std::max(blocksPerSm, 1), 4);
		int wanted = std::max(gridFor(std::max(maxItems, bodyCap), S2B_BLOCK), 1);
		int grid = std::min(w->smCount * blocksPerSm, wanted);
		void* args[] = {&a, &pp};
		if (w->solveKernelStart == nullptr)
		{
			S2B_CHECK(cudaEventCreate(&w->solveKernelStart));
			S2B_CHECK(cudaEventCreate(&w->solveKernelEnd));
		}
		S2B_CHECK(cudaEventRecord(w->solveKernelStart, st));
		S2B_CHECK(cudaLaunchCooperativeKernel((void*)s2bPersistentTgsSoft, dim3(grid), dim3(S2B_BLOCK), args, 0, st));
		S2B_CHECK(cudaEventRecord(w->solveKernelEnd, st));
		w->solveKernelTimed = true;
		w->kernelLaunches += 1;
	}
	else
	{
		runTgsSoftMultiLaunch(w, a, pp, plan, hostNJ, hostNC);
	}

	// work meter: constraint-iterations of this step = (contact constraints + joints) x solve passes
	{
		w->dWork.reserve(4, st, true);
		int passes = ctx.iterations * (1 + (ctx.extraIterations > 0 ? 1 : 0));
		S2B_LAUNCH(w, s2bMeterWork, 1, 1, 0, s->counts.p, passes, w->dWork.p);
	}
}

extern "C" void s2b_get_work(s2bWorld* w, uint64_t out[2], int reset)
{
	S2B_CHECK(cudaSetDevice(w->device));
	S2B_CHECK(cudaStreamSynchronize(w->stream));
	out[0] = out[1] = 0;
	if (w->dWork.p != nullptr)
	{
		S2B_CHECK(cudaMemcpy(out, w->dWork.p, sizeof(uint64_t) * 2, cudaMemcpyDeviceToHost));
		if (reset)
		{
			S2B_CHECK(cudaMemset(w->dWork.p, 0, sizeof(uint64_t) * 2));
		}
	}
}

extern "C" float s2b_last_solve_kernel_ms(s2bWorld* w)
{
	S2B_CHECK(cudaSetDevice(w->device));
	if (w->solveKernelTimed == false)
	{
		return 0.0f;
	}
	S2B_CHECK(cudaEventSynchronize(w->solveKernelEnd));
	float ms = 0.0f;
	S2B_CHECK(cudaEventElapsedTime(&ms, w->solveKernelStart, w->solveKernelEnd));
	return ms;
}

extern "C" int s2b_download_solve_order(s2bWorld* w, int32_t* items, int maxItems, int32_t* groupSizes, int maxGroups,
										int32_t* groupCount)
{
	S2B_CHECK(cudaSetDevice(w->device));
	SolverScratch* s = s2bGetSolverScratch(w);
	S2B_CHECK(cudaStreamSynchronize(w->stream));
	int counts[CNT_SIZE];
	S2B_CHECK(cudaMemcpy(counts, s->counts.p, sizeof(counts), cudaMemcpyDeviceToHost));
	int nC = counts[CNT_CONTACTS], nJ = counts[CNT_JOINTS], groups = counts[CNT_GROUPS];
	bool wavefront = w->schedule == S2B_SCHEDULE_WAVEFRONT;
	int tableLen = (wavefront ? groups : S2B_MAX_COLORS) + 2;
	std::vector<int> cOff((size_t)tableLen), jOff((size_t)tableLen), src((size_t)std::max(nC, 1)), jPerm((size_t)std::max(nJ, 1)),
		jointSlots((size_t)std::max(nJ, 1));
	S2B_CHECK(cudaMemcpy(cOff.data(), s->cGroupOff.p, sizeof(int) * (size_t)tableLen, cudaMemcpyDeviceToHost));
	S2B_CHECK(cudaMemcpy(jOff.data(), s->jGroupOff.p, sizeof(int) * (size_t)tableLen, cudaMemcpyDeviceToHost));
	if (nC > 0)
	{
		S2B_CHECK(cudaMemcpy(src.data(), s->src.p, sizeof(int) * (size_t)nC, cudaMemcpyDeviceToHost));
	}
	if (nJ > 0)
	{
		S2B_CHECK(cudaMemcpy(jPerm.data(), s->jPerm.p, sizeof(int) * (size_t)nJ, cudaMemcpyDeviceToHost));
		S2B_CHECK(cudaMemcpy(jointSlots.data(), s->jointSlots.p, sizeof(int) * (size_t)nJ, cudaMemcpyDeviceToHost));
	}
	// group g = joints [jOff[g], jOff[g+1]) then contacts [cOff[g], cOff[g+1]); the serial overflow group (colour
	// schedule only) sits at table index S2B_MAX_COLORS and is visited last
	int written = 0, groupsOut = 0;
	auto emit = [&](int g) {
		int size = 0;
		for (int t = jOff[(size_t)g]; t < jOff[(size_t)g + 1]; ++t, ++size)
		{
			if (written < maxItems && items != nullptr)
			{
				items[written] = -1 - jointSlots[(size_t)jPerm[(size_t)t]];
			}
			written += 1;
		}
		for (int t = cOff[(size_t)g]; t < cOff[(size_t)g + 1]; ++t, ++size)
		{
			if (written < maxItems && items != nullptr)
			{
				items[written] = src[(size_t)t];
			}
			written += 1;
		}
		if (size > 0)
		{
			if (groupSizes != nullptr && groupsOut < maxGroups)
			{
				groupSizes[groupsOut] = size;
			}
			groupsOut += 1;
		}
	};
	for (int g = 0; g < groups; ++g)
	{
		emit(g);
	}
	if (wavefront == false)
	{
		emit(S2B_MAX_COLORS);
	}
	if (groupCount != nullptr)
	{
		*groupCount = groupsOut;
	}
	return written;
}

extern "C" void s2b_get_counters(s2bWorld* w, s2bCounters* out)
{
	S2B_CHECK(cudaSetDevice(w->device));
	S2B_CHECK(cudaStreamSynchronize(w->stream));
	memset(out, 0, sizeof(*out));
	out->bodyCapacity = w->bodyCap;
	out->shapeCapacity = w->shapeCap;
	out->jointCapacity = w->jointCap;
	out->contactCount = w->contactCount;
	if (w->scratch != nullptr && w->scratch->counts.p != nullptr)
	{
		int counts[CNT_SIZE];
		S2B_CHECK(cudaMemcpy(counts, w->scratch->counts.p, sizeof(counts), cudaMemcpyDeviceToHost));
		out->constraintCount = counts[CNT_CONTACTS];
		out->jointCount = counts[CNT_JOINTS];
		out->groupCount = counts[CNT_GROUPS];
		out->overflowCount = counts[CNT_OVERFLOW_C] + counts[CNT_OVERFLOW_J];
	}
	out->treeHeight = w->treeHeight;
	out->movedCount = w->hostMail[MAIL_MOVED];
	out->pairPassCount = w->pairPassCount;
	out->kernelLaunches = w->kernelLaunches;
}

extern "C" float s2b_time_color_kernel(s2bWorld* w, const s2bStepContext* context, int reps, int* constraints)
{
	(void)w;
	(void)context;
	(void)reps;
	if (constraints)
	{
		*constraints = 0;
	}
	return 0.0f;
}
