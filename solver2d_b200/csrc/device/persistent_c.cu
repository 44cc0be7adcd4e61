// solver2d-b200 — instantiations of the persistent solver kernel (see persistent_a.cu)
#include "persistent.cuh"

void* s2bPersistentKernelA(int solverType);
void* s2bPersistentKernelB(int solverType);

static void* s2bPersistentKernelC(int solverType)
{
	switch (solverType)
	{
		case 3:
			return (void*)s2bPersistentSolveT<3>; // PGS_NGS_Block
		case 6:
			return (void*)s2bPersistentSolveT<6>; // TGS_Sticky
		case 9:
			return (void*)s2bPersistentSolveT<9>; // XPBD
		default:
			return nullptr;
	}
}

void* s2bPersistentKernel(int solverType)
{
	void* k = s2bPersistentKernelA(solverType);
	if (k == nullptr)
	{
		k = s2bPersistentKernelB(solverType);
	}
	if (k == nullptr)
	{
		k = s2bPersistentKernelC(solverType);
	}
	return k;
}
