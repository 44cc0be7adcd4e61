// Micro-benchmark lab for the persistent solver kernel's dependent "colour step" on B200:
//   (A) cost of a bare grid-wide barrier for several designs and grid shapes;
//   (B) cost of one colour step = barrier + gather of two body rows (L2 resident) + arithmetic + scatter, which is what
//       the solver pays 56 times per TGS_Soft step at 100 k boxes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/barrier_lab tools/barrier_lab.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
namespace cg = cooperative_groups;

#define CHECK(x)                                                                                                       \
	do                                                                                                                 \
	{                                                                                                                  \
		cudaError_t e = (x);                                                                                           \
		if (e != cudaSuccess)                                                                                          \
		{                                                                                                              \
			printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__);                                     \
			exit(1);                                                                                                   \
		}                                                                                                              \
	} while (0)

__device__ __forceinline__ unsigned ldAcquire(const unsigned* p)
{
	unsigned v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ unsigned ldRelaxed(const unsigned* p)
{
	unsigned v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void redRelease(unsigned* p, unsigned v)
{
	asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ldCg(const float4* p)
{
	float4 v;
	asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
	return v;
}
__device__ __forceinline__ void stCg(float4* p, float4 v)
{
	asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// design 0: cooperative groups
// design 1: fence + atomicAdd(return) + last arriver releases a flag, others poll the flag (acquire)
// design 2: red.release on a monotonic counter, everyone polls the counter (acquire)
// design 3: like 2 but relaxed polling + one fence.acquire at the end
// design 4: like 2, polled by every warp's lane 0 instead of thread 0 + second __syncthreads (no trailing bar.sync)
template <int DESIGN> __device__ __forceinline__ void gridBarrier(unsigned* bar, unsigned gen, unsigned blocks, cg::grid_group& grid)
{
	if (DESIGN == 0)
	{
		grid.sync();
		return;
	}
	if (DESIGN == 1)
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
		return;
	}
	if (DESIGN == 2)
	{
		__syncthreads();
		if (threadIdx.x == 0)
		{
			redRelease(bar, 1u);
			unsigned target = blocks * gen;
			while (ldAcquire(bar) < target)
			{
			}
		}
		__syncthreads();
		return;
	}
	if (DESIGN == 3)
	{
		__syncthreads();
		if (threadIdx.x == 0)
		{
			redRelease(bar, 1u);
			unsigned target = blocks * gen;
			while (ldRelaxed(bar) < target)
			{
			}
			asm volatile("fence.acq_rel.gpu;" ::: "memory");
		}
		__syncthreads();
		return;
	}
	if (DESIGN == 4)
	{
		__syncthreads();
		if (threadIdx.x == 0)
		{
			redRelease(bar, 1u);
		}
		if ((threadIdx.x & 31) == 0)
		{
			unsigned target = blocks * gen;
			while (ldAcquire(bar) < target)
			{
			}
		}
		__syncwarp();
		return;
	}
}

template <int DESIGN> __global__ void barrierOnly(int iters, unsigned* bar, float* sink)
{
	cg::grid_group grid = cg::this_grid();
	float acc = 0.0f;
	for (int i = 0; i < iters; ++i)
	{
		acc += 1.0f;
		gridBarrier<DESIGN>(bar, (unsigned)(i + 1), gridDim.x, grid);
	}
	if (threadIdx.x == 0 && blockIdx.x == 0)
	{
		*sink = acc;
	}
}

// colour step: n "constraints", thread t handles constraint t (grid-stride), gathers rows ia[t], ib[t] of two body columns,
// does `flops` dependent multiply-adds, writes one column of both rows back. MODE 0: default loads/stores; MODE 1: .cg
template <int DESIGN, int MODE> __global__ void colourStep(int iters, unsigned* bar, int n, const int2* __restrict__ pairs, float4* vel,
														   float4* pose, int flops, int colours)
{
	cg::grid_group grid = cg::this_grid();
	int tid = blockIdx.x * blockDim.x + threadIdx.x;
	int stride = gridDim.x * blockDim.x;
	unsigned gen = 0;
	for (int i = 0; i < iters; ++i)
	{
		int c = i % colours;
		for (int t = tid; t < n; t += stride)
		{
			int2 p = pairs[c * n + t];
			float4 va, vb, pa, pb;
			if (MODE == 0)
			{
				va = vel[p.x];
				vb = vel[p.y];
				pa = pose[p.x];
				pb = pose[p.y];
			}
			else
			{
				va = ldCg(vel + p.x);
				vb = ldCg(vel + p.y);
				pa = ldCg(pose + p.x);
				pb = ldCg(pose + p.y);
			}
			float x = va.x + vb.y, y = pa.z * pb.w;
			for (int k = 0; k < flops; ++k)
			{
				x = x * 0.999f + y;
				y = y * 1.001f - x * 1e-3f;
			}
			va.x = x;
			vb.y = y;
			if (MODE == 0)
			{
				vel[p.x] = va;
				vel[p.y] = vb;
			}
			else
			{
				stCg(vel + p.x, va);
				stCg(vel + p.y, vb);
			}
		}
		gen += 1;
		gridBarrier<DESIGN>(bar, gen, gridDim.x, grid);
	}
}

template <typename K, typename... Args> static float timeCoop(K kernel, int grid, int block, int reps, unsigned* bar, Args... args)
{
	cudaEvent_t e0, e1;
	CHECK(cudaEventCreate(&e0));
	CHECK(cudaEventCreate(&e1));
	float best = 1e30f;
	for (int r = 0; r < reps; ++r)
	{
		CHECK(cudaMemset(bar, 0, 1024));
		void* argv[] = {(void*)&args...};
		CHECK(cudaEventRecord(e0));
		cudaError_t err = cudaLaunchCooperativeKernel((void*)kernel, dim3(grid), dim3(block), argv, 0, 0);
		CHECK(cudaEventRecord(e1));
		CHECK(cudaEventSynchronize(e1));
		if (err != cudaSuccess)
		{
			printf("launch failed: %s\n", cudaGetErrorString(err));
			(void)cudaGetLastError();
			return -1.0f;
		}
		float ms = 0;
		CHECK(cudaEventElapsedTime(&ms, e0, e1));
		if (r > 0 && ms < best)
		{
			best = ms;
		}
	}
	return best;
}

int main()
{
	cudaDeviceProp prop;
	CHECK(cudaGetDeviceProperties(&prop, 0));
	int sms = prop.multiProcessorCount;
	printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
	float* sink;
	unsigned* bar;
	CHECK(cudaMalloc(&sink, 4));
	CHECK(cudaMalloc(&bar, 1024));
	int iters = 4000;

	struct Shape
	{
		int perSm, block;
	} shapes[] = {{1, 256}, {1, 512}, {1, 1024}, {2, 256}, {2, 512}, {4, 256}};
	printf("== (A) bare barrier, us per barrier ==\n");
	for (auto s : shapes)
	{
		int grid = sms * s.perSm;
		float t0 = timeCoop(barrierOnly<0>, grid, s.block, 3, bar, iters, bar, sink);
		float t1 = timeCoop(barrierOnly<1>, grid, s.block, 3, bar, iters, bar, sink);
		float t2 = timeCoop(barrierOnly<2>, grid, s.block, 3, bar, iters, bar, sink);
		float t3 = timeCoop(barrierOnly<3>, grid, s.block, 3, bar, iters, bar, sink);
		float t4 = timeCoop(barrierOnly<4>, grid, s.block, 3, bar, iters, bar, sink);
		printf("grid %4d x %4d : cg %.3f | fence+atomic+flag %.3f | red.release+poll %.3f | relaxed poll+fence %.3f | per-warp poll %.3f\n", grid,
			   s.block, 1e3f * t0 / iters, 1e3f * t1 / iters, 1e3f * t2 / iters, 1e3f * t3 / iters, 1e3f * t4 / iters);
	}

	// (B) colour step on a pyramid-like graph: bodies 100k, 7 colours of n constraints, constraint t of colour c touches
	// bodies (perm_c[2t], perm_c[2t+1]) — a random perfect matching per colour: every body in at most one constraint per colour
	int nBodies = 100128;
	int colours = 7;
	float4 *vel, *pose;
	CHECK(cudaMalloc(&vel, sizeof(float4) * nBodies));
	CHECK(cudaMalloc(&pose, sizeof(float4) * nBodies));
	CHECK(cudaMemset(vel, 0, sizeof(float4) * nBodies));
	CHECK(cudaMemset(pose, 0, sizeof(float4) * nBodies));
	int counts[] = {36, 5000, 50000};
	for (int locality = 0; locality < 2; ++locality)
	{
		for (int n : counts)
		{
			std::vector<int2> pairs((size_t)colours * n);
			srand(1234);
			for (int c = 0; c < colours; ++c)
			{
				std::vector<int> perm(nBodies);
				for (int i = 0; i < nBodies; ++i)
				{
					perm[i] = i;
				}
				if (locality == 0)
				{
					for (int i = nBodies - 1; i > 0; --i)
					{
						int j = rand() % (i + 1);
						std::swap(perm[i], perm[j]);
					}
				}
				else
				{
					// neighbouring bodies, shifted per colour (what shape-pair key order gives on a pyramid)
					for (int i = 0; i < nBodies; ++i)
					{
						perm[i] = (i + c) % nBodies;
					}
				}
				for (int t = 0; t < n; ++t)
				{
					pairs[(size_t)c * n + t] = make_int2(perm[2 * t], perm[2 * t + 1]);
				}
			}
			int2* dPairs;
			CHECK(cudaMalloc(&dPairs, sizeof(int2) * pairs.size()));
			CHECK(cudaMemcpy(dPairs, pairs.data(), sizeof(int2) * pairs.size(), cudaMemcpyHostToDevice));
			printf("== (B) colour step, %s pairs, n = %d constraints per colour, us per step ==\n", locality ? "local" : "random", n);
			for (auto s : shapes)
			{
				int grid = sms * s.perSm;
				for (int flops = 0; flops <= 100; flops += 100)
				{
					int it = 2100;
					float a0 = timeCoop(colourStep<0, 0>, grid, s.block, 3, bar, it, bar, n, (const int2*)dPairs, vel, pose, flops, colours);
					float a1 = timeCoop(colourStep<1, 0>, grid, s.block, 3, bar, it, bar, n, (const int2*)dPairs, vel, pose, flops, colours);
					float a2 = timeCoop(colourStep<2, 0>, grid, s.block, 3, bar, it, bar, n, (const int2*)dPairs, vel, pose, flops, colours);
					float a3 = timeCoop(colourStep<2, 1>, grid, s.block, 3, bar, it, bar, n, (const int2*)dPairs, vel, pose, flops, colours);
					float a4 = timeCoop(colourStep<3, 1>, grid, s.block, 3, bar, it, bar, n, (const int2*)dPairs, vel, pose, flops, colours);
					printf("grid %4d x %4d flops %3d : cg %.3f | flag %.3f | red+poll %.3f | red+poll .cg %.3f | relaxed+fence .cg %.3f\n", grid, s.block,
						   flops, 1e3f * a0 / it, 1e3f * a1 / it, 1e3f * a2 / it, 1e3f * a3 / it, 1e3f * a4 / it);
				}
			}
			CHECK(cudaFree(dPairs));
		}
	}
	return 0;
}
