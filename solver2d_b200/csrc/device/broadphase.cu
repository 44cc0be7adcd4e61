// solver2d-b200 — broad-phase pair update on the device (placeholder until the BVH pass lands).
#include "s2b_internal.cuh"

void s2bFreeBroadScratch(s2bWorld* w)
{
	(void)w;
}

void s2bBroadphaseUpdatePairs(s2bWorld* w)
{
	(void)w;
}
