// solver2d-b200 — instantiations of the persistent solver kernel, one per variant (split over three translation units
// so that they compile in parallel). See persistent.cuh.
#include "persistent.cuh"

void* s2bPersistentKernelA(int solverType)
{
	switch (solverType)
	{
		case 7:
			return (void*)s2bPersistentSolveT<7>; // TGS_Soft
		case 5:
			return (void*)s2bPersistentSolveT<5>; // SoftStep
		case 8:
			return (void*)s2bPersistentSolveT<8>; // TGS_NGS
		default:
			return nullptr;
	}
}
