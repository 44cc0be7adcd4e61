// solver2d-b200 — debug-draw callback table filled in by the client (ABI of reference
// include/solver2d/debug_draw.h:9-55). s2World_Draw reads simulation state back from the GPU lazily and feeds these.
#pragma once

#include "solver2d/types.h"

typedef struct s2DebugDraw
{
	void (*DrawPolygon)(const s2Vec2* vertices, int vertexCount, s2Color color, void* context);
	void (*DrawSolidPolygon)(const s2Vec2* vertices, int vertexCount, s2Color color, void* context);
	void (*DrawRoundedPolygon)(const s2Vec2* vertices, int vertexCount, float radius, s2Color lineColor, s2Color fillColor,
							   void* context);
	void (*DrawCircle)(s2Vec2 center, float radius, s2Color color, void* context);
	void (*DrawSolidCircle)(s2Vec2 center, float radius, s2Vec2 axis, s2Color color, void* context);
	void (*DrawCapsule)(s2Vec2 p1, s2Vec2 p2, float radius, s2Color color, void* context);
	void (*DrawSolidCapsule)(s2Vec2 p1, s2Vec2 p2, float radius, s2Color color, void* context);
	void (*DrawSegment)(s2Vec2 p1, s2Vec2 p2, s2Color color, void* context);
	void (*DrawTransform)(s2Transform xf, void* context);
	void (*DrawPoint)(s2Vec2 p, float size, s2Color color, void* context);
	void (*DrawString)(s2Vec2 p, const char* s, void* context);

	s2Color dynamicBodyColor;
	bool drawShapes;
	bool drawJoints;
	bool drawAABBs;
	bool drawMass;
	bool drawContactPoints;
	bool drawContactNormals;
	bool drawContactImpulses;
	bool drawFrictionImpulses;
	void* context;
} s2DebugDraw;
