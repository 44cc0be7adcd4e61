// Experiment: a CUDA graph with an IF conditional node whose body holds (a) plain kernels, (b) cub device-wide calls and
// (c) a cooperative kernel, all stream-captured into the body graph; the condition is set by an upstream kernel from a
// device flag. Answers: does it build without -rdc, does capture-to-graph of cub + cooperative launches work inside a
// conditional body on this driver, and what does a skipped body cost per replay.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/condgraph_lab tools/condgraph_lab.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cub/cub.cuh>
#include <cuda_runtime.h>
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

__global__ void setCondition(cudaGraphConditionalHandle handle, const int* dirty)
{
	cudaGraphSetConditional(handle, *dirty != 0 ? 1u : 0u);
}

__global__ void bump(int* counter)
{
	if (threadIdx.x == 0 && blockIdx.x == 0)
	{
		counter[0] += 1;
	}
}

__global__ void coopKernel(int* counter)
{
	cg::grid_group grid = cg::this_grid();
	grid.sync();
	if (threadIdx.x == 0 && blockIdx.x == 0)
	{
		counter[1] += 1;
	}
	grid.sync();
}

__global__ void always(int* counter)
{
	if (threadIdx.x == 0 && blockIdx.x == 0)
	{
		counter[2] += 1;
	}
}

int main()
{
	cudaStream_t st;
	CHECK(cudaStreamCreate(&st));
	int *dirty, *counter;
	CHECK(cudaMalloc(&dirty, 4));
	CHECK(cudaMalloc(&counter, 64));
	CHECK(cudaMemset(counter, 0, 64));
	int n = 300000;
	unsigned short *kIn, *kOut;
	int *vIn, *vOut;
	CHECK(cudaMalloc(&kIn, 2 * n));
	CHECK(cudaMalloc(&kOut, 2 * n));
	CHECK(cudaMalloc(&vIn, 4 * n));
	CHECK(cudaMalloc(&vOut, 4 * n));
	CHECK(cudaMemset(kIn, 1, 2 * n));
	size_t tb = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, tb, kIn, kOut, vIn, vOut, n, 0, 16, st);
	void* temp;
	CHECK(cudaMalloc(&temp, tb));

	cudaGraph_t graph;
	CHECK(cudaGraphCreate(&graph, 0));
	cudaGraphConditionalHandle handle;
	CHECK(cudaGraphConditionalHandleCreate(&handle, graph, 0, cudaGraphCondAssignDefault));

	// node 1: set the condition from the device flag
	cudaGraphNode_t setNode;
	{
		cudaKernelNodeParams kp = {};
		void* args[] = {&handle, &dirty};
		kp.func = (void*)setCondition;
		kp.gridDim = dim3(1);
		kp.blockDim = dim3(1);
		kp.kernelParams = args;
		CHECK(cudaGraphAddKernelNode(&setNode, graph, nullptr, 0, &kp));
	}
	// node 2: IF
	cudaGraphNode_t ifNode;
	cudaGraph_t body;
	{
		cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
		p.conditional.handle = handle;
		p.conditional.type = cudaGraphCondTypeIf;
		p.conditional.size = 1;
		CHECK(cudaGraphAddNode(&ifNode, graph, &setNode, 1, &p));
		body = p.conditional.phGraph_out[0];
	}
	// body: captured
	{
		CHECK(cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
		bump<<<1, 32, 0, st>>>(counter);
		size_t t2 = tb;
		cub::DeviceRadixSort::SortPairs(temp, t2, kIn, kOut, vIn, vOut, n, 0, 16, st);
		void* args[] = {&counter};
		cudaError_t e = cudaLaunchCooperativeKernel((void*)coopKernel, dim3(148), dim3(256), args, 0, st);
		printf("cooperative launch inside capture-to-conditional-body: %s\n", cudaGetErrorString(e));
		for (int i = 0; i < 20; ++i)
		{
			bump<<<1, 32, 0, st>>>(counter);
		}
		cudaGraph_t out = nullptr;
		e = cudaStreamEndCapture(st, &out);
		printf("end capture: %s (graph %s body)\n", cudaGetErrorString(e), out == body ? "==" : "!=");
		if (e != cudaSuccess)
		{
			return 1;
		}
	}
	// node 3: runs always, after the IF
	{
		cudaGraphNode_t node;
		cudaKernelNodeParams kp = {};
		void* args[] = {&counter};
		kp.func = (void*)always;
		kp.gridDim = dim3(1);
		kp.blockDim = dim3(32);
		kp.kernelParams = args;
		CHECK(cudaGraphAddKernelNode(&node, graph, &ifNode, 1, &kp));
	}
	cudaGraphExec_t exec;
	cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
	printf("instantiate: %s\n", cudaGetErrorString(e));
	if (e != cudaSuccess)
	{
		return 1;
	}
	cudaEvent_t e0, e1;
	CHECK(cudaEventCreate(&e0));
	CHECK(cudaEventCreate(&e1));
	for (int mode = 0; mode < 2; ++mode)
	{
		int flag = mode;
		CHECK(cudaMemcpy(dirty, &flag, 4, cudaMemcpyHostToDevice));
		CHECK(cudaMemset(counter, 0, 64));
		for (int i = 0; i < 5; ++i)
		{
			CHECK(cudaGraphLaunch(exec, st));
		}
		CHECK(cudaStreamSynchronize(st));
		CHECK(cudaEventRecord(e0, st));
		int reps = 200;
		for (int i = 0; i < reps; ++i)
		{
			CHECK(cudaGraphLaunch(exec, st));
		}
		CHECK(cudaEventRecord(e1, st));
		CHECK(cudaEventSynchronize(e1));
		float ms;
		CHECK(cudaEventElapsedTime(&ms, e0, e1));
		int h[3];
		CHECK(cudaMemcpy(h, counter, 12, cudaMemcpyDeviceToHost));
		printf("dirty=%d: %.2f us per replay; counters body-bumps %d coop %d always %d (of %d launches)\n", mode, 1e3f * ms / reps, h[0], h[1], h[2],
			   reps + 5);
	}
	return 0;
}
