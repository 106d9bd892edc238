#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_body(int* cnt, cudaGraphConditionalHandle h){ int c = ++(*cnt); cudaGraphSetConditional(h, c < 5 ? 1u : 0u); }
int main(){
  cudaStream_t st; cudaStreamCreate(&st);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
  cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node; cudaError_t e = cudaGraphAddNode(&node, g, nullptr, 0, &p); printf("addnode %s\n", cudaGetErrorString(e));
  cudaGraph_t body = p.conditional.phGraph_out[0];
  int* cnt; cudaMalloc(&cnt, 4); cudaMemset(cnt, 0, 4);
  e = cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeGlobal); printf("begin %s\n", cudaGetErrorString(e));
  k_body<<<1,1,0,st>>>(cnt, h);
  cudaGraph_t out; e = cudaStreamEndCapture(st, &out); printf("end %s\n", cudaGetErrorString(e));
  cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0); printf("inst %s\n", cudaGetErrorString(e));
  cudaGraphLaunch(ex, st); cudaStreamSynchronize(st);
  int hc; cudaMemcpy(&hc, cnt, 4, cudaMemcpyDeviceToHost); printf("count %d (expect 5)\n", hc);
}
