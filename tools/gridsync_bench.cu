// Micro-benchmark: cost of a grid-wide barrier on B200 — cooperative_groups grid.sync() vs a hand-rolled
// monotonic-counter barrier — for the grid shapes the persistent solver kernel can use. Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=false -o tools/gridsync_bench tools/gridsync_bench.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void cgSyncKernel(int iters, float* sink)
{
	cg::grid_group grid = cg::this_grid();
	float acc = 0.0f;
	for (int i = 0; i < iters; ++i)
	{
		acc += 1.0f;
		grid.sync();
	}
	if (threadIdx.x == 0 && blockIdx.x == 0)
	{
		*sink = acc;
	}
}

__device__ __forceinline__ unsigned ldAcquire(const unsigned* p)
{
	unsigned v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void customBarrier(unsigned* bar, unsigned gen, unsigned blocks)
{
	__syncthreads();
	if (threadIdx.x == 0)
	{
		__threadfence();
		unsigned arrived = atomicAdd(bar, 1u);
		if (arrived == blocks * gen - 1u)
		{
			asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 32), "r"(gen) : "memory");
		}
		else
		{
			while (ldAcquire(bar + 32) < gen)
			{
			}
		}
	}
	__syncthreads();
}

__global__ void customSyncKernel(int iters, unsigned* bar, float* sink)
{
	float acc = 0.0f;
	for (int i = 0; i < iters; ++i)
	{
		acc += 1.0f;
		customBarrier(bar, (unsigned)(i + 1), gridDim.x);
	}
	if (threadIdx.x == 0 && blockIdx.x == 0)
	{
		*sink = acc;
	}
}

int main()
{
	float* sink;
	unsigned* bar;
	cudaMalloc(&sink, 4);
	cudaMalloc(&bar, 256);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	int iters = 2000;
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, 0);
	int sms = prop.multiProcessorCount;
	printf("device %s, %d SMs\n", prop.name, sms);
	int perSm[] = {1, 2, 4};
	int blocks[] = {128, 256, 512};
	for (int a = 0; a < 3; ++a)
	{
		for (int b = 0; b < 3; ++b)
		{
			int grid = sms * perSm[a], block = blocks[b];
			void* args1[] = {&iters, &sink};
			for (int rep = 0; rep < 2; ++rep)
			{
				cudaEventRecord(e0);
				cudaError_t err = cudaLaunchCooperativeKernel((void*)cgSyncKernel, dim3(grid), dim3(block), args1, 0, 0);
				cudaEventRecord(e1);
				cudaEventSynchronize(e1);
				float ms = 0;
				cudaEventElapsedTime(&ms, e0, e1);
				if (rep == 1)
				{
					printf("cg grid.sync  grid=%4d x %3d : %7.3f us/barrier (%s)\n", grid, block, 1e3f * ms / iters, cudaGetErrorString(err));
				}
			}
			void* args2[] = {&iters, &bar, &sink};
			for (int rep = 0; rep < 2; ++rep)
			{
				cudaMemset(bar, 0, 256);
				cudaEventRecord(e0);
				cudaError_t err = cudaLaunchCooperativeKernel((void*)customSyncKernel, dim3(grid), dim3(block), args2, 0, 0);
				cudaEventRecord(e1);
				cudaEventSynchronize(e1);
				float ms = 0;
				cudaEventElapsedTime(&ms, e0, e1);
				if (rep == 1)
				{
					printf("custom barrier grid=%4d x %3d : %7.3f us/barrier (%s)\n", grid, block, 1e3f * ms / iters, cudaGetErrorString(err));
				}
			}
		}
	}
	return 0;
}
