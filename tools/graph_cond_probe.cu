#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_set(cudaGraphConditionalHandle h, const int *flag) { cudaGraphSetConditional(h, *flag != 0 ? 1u : 0u); }
__global__ void k_body(int *c) { atomicAdd(c, 1); }
int main() {
    cudaStream_t st; cudaStreamCreate(&st);
    int *flag, *cnt; cudaMalloc(&flag, 4); cudaMalloc(&cnt, 4); cudaMemset(cnt, 0, 4);
    cudaGraph_t g; 
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    cudaStreamCaptureStatus status; cudaGraph_t cg; const cudaGraphNode_t *deps; size_t ndeps;
    cudaStreamGetCaptureInfo(st, &status, nullptr, &cg, &deps, &ndeps);
    cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, cg, 0, cudaGraphCondAssignDefault);
    k_set<<<1,1,0,st>>>(h, flag);
    cudaStreamGetCaptureInfo(st, &status, nullptr, &cg, &deps, &ndeps);
    cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeIf; p.conditional.size = 1;
    cudaGraphNode_t node; cudaError_t e = cudaGraphAddNode(&node, cg, deps, ndeps, &p);
    printf("addnode %s\n", cudaGetErrorString(e));
    cudaGraph_t body = p.conditional.phGraph_out[0];
    cudaStream_t st2; cudaStreamCreate(&st2);
    e = cudaStreamBeginCaptureToGraph(st2, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    printf("begin2 %s\n", cudaGetErrorString(e));
    k_body<<<1,1,0,st2>>>(cnt);
    e = cudaStreamEndCapture(st2, nullptr); printf("end2 %s\n", cudaGetErrorString(e));
    e = cudaStreamUpdateCaptureDependencies(st, &node, 1, cudaStreamSetCaptureDependencies); printf("upd %s\n", cudaGetErrorString(e));
    k_body<<<1,1,0,st>>>(cnt + 0);
    e = cudaStreamEndCapture(st, &g); printf("end %s\n", cudaGetErrorString(e));
    cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0); printf("inst %s\n", cudaGetErrorString(e));
    int one = 1, zero = 0, c;
    cudaMemcpy(flag, &zero, 4, cudaMemcpyHostToDevice); cudaGraphLaunch(ex, st); cudaStreamSynchronize(st);
    cudaMemcpy(&c, cnt, 4, cudaMemcpyDeviceToHost); printf("flag0 cnt=%d (expect 1)\n", c);
    cudaMemcpy(flag, &one, 4, cudaMemcpyHostToDevice); cudaGraphLaunch(ex, st); cudaStreamSynchronize(st);
    cudaMemcpy(&c, cnt, 4, cudaMemcpyDeviceToHost); printf("flag1 cnt=%d (expect 3)\n", c);
    return 0;
}
