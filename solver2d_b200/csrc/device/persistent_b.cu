// solver2d-b200 — instantiations of the persistent solver kernel (see persistent_a.cu)
#include "persistent.cuh"

void* s2bPersistentKernelB(int solverType)
{
	switch (solverType)
	{
		case 0:
			return (void*)s2bPersistentSolveT<0>; // Jacobi
		case 1:
			return (void*)s2bPersistentSolveT<1>; // PGS
		case 2:
			return (void*)s2bPersistentSolveT<2>; // PGS_NGS
		case 4:
			return (void*)s2bPersistentSolveT<4>; // PGS_Soft
		default:
			return nullptr;
	}
}
