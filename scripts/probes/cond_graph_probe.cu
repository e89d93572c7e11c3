// Probe (run once on the GPU box, not part of the product): does a CUDA-graph WHILE node accept a COOPERATIVE
// kernel in its body, and what does one trip of the loop cost?  nvcc -arch=sm_100a -o cond_graph_probe cond_graph_probe.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
namespace cg = cooperative_groups;
__global__ void body(int *ctr, int trips, cudaGraphConditionalHandle h)
{
    cg::grid_group g = cg::this_grid();
    g.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int v = atomicAdd(ctr, 1);
        if (h != 0) cudaGraphSetConditional(h, v + 1 < trips ? 1u : 0u);
    }
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
int main()
{
    cudaStream_t s, s2;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    int *ctr;
    CK(cudaMalloc(&ctr, 4));
    CK(cudaMemset(ctr, 0, 4));
    int trips = 1000;
    cudaGraph_t g;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    cudaStreamCaptureStatus st;
    unsigned long long id;
    cudaGraph_t cgr;
    const cudaGraphNode_t *deps;
    size_t nd;
    CK(cudaStreamGetCaptureInfo(s, &st, &id, &cgr, &deps, &nd));
    cudaGraphConditionalHandle h;
    CK(cudaGraphConditionalHandleCreate(&h, cgr, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = h;
    p.conditional.type = cudaGraphCondTypeWhile;
    p.conditional.size = 1;
    cudaGraphNode_t node;
    CK(cudaGraphAddNode(&node, cgr, deps, nd, &p));
    cudaGraph_t bg = p.conditional.phGraph_out[0];
    CK(cudaStreamBeginCaptureToGraph(s2, bg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    void *args[] = {(void *)&ctr, (void *)&trips, (void *)&h};
    CK(cudaLaunchCooperativeKernel((const void *)body, dim3(148), dim3(256), args, 0, s2));
    CK(cudaStreamEndCapture(s2, nullptr));
    CK(cudaStreamUpdateCaptureDependencies(s, &node, 1, cudaStreamSetCaptureDependencies));
    CK(cudaStreamEndCapture(s, &g));
    cudaGraphExec_t ex;
    CK(cudaGraphInstantiate(&ex, g, 0));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaMemsetAsync(ctr, 0, 4, s));
        cudaEventRecord(e0, s);
        CK(cudaGraphLaunch(ex, s));
        cudaEventRecord(e1, s);
        CK(cudaStreamSynchronize(s));
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        int hc;
        cudaMemcpy(&hc, ctr, 4, cudaMemcpyDeviceToHost);
        printf("while-node with a cooperative body: %d trips in %.3f ms = %.2f us per trip\n", hc, ms, 1e3 * ms / hc);
    }
    // the same kernel launched back to back without a graph
    int zero = 0;
    cudaGraphConditionalHandle h0 = 0;
    void *args2[] = {(void *)&ctr, (void *)&zero, (void *)&h0};
    cudaEventRecord(e0, s);
    for (int k = 0; k < trips; k++) CK(cudaLaunchCooperativeKernel((const void *)body, dim3(148), dim3(256), args2, 0, s));
    cudaEventRecord(e1, s);
    CK(cudaStreamSynchronize(s));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("stream launches: %.2f us per launch\n", 1e3 * ms / trips);
    return 0;
}
