// solver2d-b200 — narrow phase on the device (placeholder until the manifold kernels land).
#include "s2b_internal.cuh"

void s2bNarrowphaseUpdate(s2bWorld* w)
{
	(void)w;
}
