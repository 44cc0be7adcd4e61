// solver2d-b200 — plain-data types of the public API (ABI of reference include/solver2d/types.h:31-163).
#pragma once

#include "solver2d/color.h"
#include "solver2d/constants.h"
#include "solver2d/id.h"

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

// Inline helpers in the public headers are shared verbatim by host C, the CUDA kernels and the test oracle so that
// all three evaluate identical float expressions. Under nvcc they become __host__ __device__.
#if defined(__CUDACC__)
	#define S2_INLINE static inline __host__ __device__
#else
	#define S2_INLINE static inline
#endif

#ifdef __cplusplus
	#define S2_LITERAL(T) T
	#define S2_ZERO_INIT {}
#else
	#define S2_LITERAL(T) (T)
	#define S2_ZERO_INIT {0}
#endif

#define S2_ARRAY_COUNT(A) (int)(sizeof(A) / sizeof(A[0]))
#define S2_MAYBE_UNUSED(x) ((void)(x))
#define S2_NULL_INDEX (-1)

// ---- linear algebra -------------------------------------------------------------------------------------------

typedef struct s2Vec2
{
	float x, y;
} s2Vec2;

// rotation stored as (sin, cos)
typedef struct s2Rot
{
	float s, c;
} s2Rot;

typedef struct s2Transform
{
	s2Vec2 p;
	s2Rot q;
} s2Transform;

// column-major 2x2
typedef struct s2Mat22
{
	s2Vec2 cx, cy;
} s2Mat22;

typedef struct s2Box
{
	s2Vec2 lowerBound;
	s2Vec2 upperBound;
} s2Box;

typedef struct s2RayCastInput
{
	s2Vec2 p1, p2;
	float maxFraction;
} s2RayCastInput;

typedef struct s2RayCastOutput
{
	s2Vec2 normal;
	s2Vec2 point;
	float fraction;
	int32_t iterations;
	bool hit;
} s2RayCastOutput;

// ---- world ------------------------------------------------------------------------------------------------------

// Solver variant of a world. The numeric order is ABI: the samples harness indexes arrays with it.
typedef enum s2SolverType
{
	s2_solverJacobi,
	s2_solverPGS,
	s2_solverPGS_NGS,
	s2_solverPGS_NGS_Block,
	s2_solverPGS_Soft,
	s2_solverSoftStep,
	s2_solverTGS_Sticky,
	s2_solverTGS_Soft,
	s2_solverTGS_NGS,
	s2_solverXPBD,
	s2_solverTypeCount,
} s2SolverType;

typedef struct s2WorldDef
{
	enum s2SolverType solverType;
} s2WorldDef;

static const s2WorldDef s2_defaultWorldDef = {s2_solverPGS_NGS_Block};

S2_INLINE s2WorldDef s2DefaultWorldDef(void)
{
	s2WorldDef def = S2_ZERO_INIT;
	def.solverType = s2_solverPGS_NGS_Block;
	return def;
}

// ---- bodies -----------------------------------------------------------------------------------------------------

typedef enum s2BodyType
{
	s2_staticBody = 0,
	s2_kinematicBody = 1,
	s2_dynamicBody = 2,
	s2_bodyTypeCount
} s2BodyType;

typedef struct s2BodyDef
{
	s2BodyType type;
	s2Vec2 position;
	float angle;
	s2Vec2 linearVelocity;
	float angularVelocity;
	float linearDamping;
	float angularDamping;
	float gravityScale;
	void* userData;
} s2BodyDef;

static const s2BodyDef s2_defaultBodyDef = {
	s2_staticBody, {0.0f, 0.0f}, 0.0f, {0.0f, 0.0f}, 0.0f, 0.0f, 0.0f, 1.0f, NULL,
};

// ---- shapes -----------------------------------------------------------------------------------------------------

typedef struct s2Filter
{
	uint32_t categoryBits;
	uint32_t maskBits;
	int32_t groupIndex;
} s2Filter;

static const s2Filter s2_defaultFilter = {0x00000001, 0xFFFFFFFF, 0};

typedef struct s2ShapeDef
{
	void* userData;
	float friction;
	float restitution;
	float density;
	s2Filter filter;
} s2ShapeDef;

static const s2ShapeDef s2_defaultShapeDef = {
	NULL, 0.6f, 0.0f, 1.0f, {0x00000001, 0xFFFFFFFF, 0},
};
