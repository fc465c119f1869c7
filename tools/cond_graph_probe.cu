#include <cuda_runtime.h>
#include <stdio.h>
__global__ void body(int* c, cudaGraphConditionalHandle h) {
  if (threadIdx.x == 0) { int v = --(*c); cudaGraphSetConditional(h, v > 0 ? 1u : 0u); }
}
int main() {
  cudaStream_t s; cudaStreamCreate(&s);
  int* c; cudaMalloc(&c, 4); int v = 5; cudaMemcpy(c, &v, 4, cudaMemcpyHostToDevice);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
  cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
  p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node; cudaGraphAddNode(&node, g, nullptr, 0, &p);
  cudaGraph_t bg = p.conditional.phGraph_out[0];
  cudaStreamBeginCaptureToGraph(s, bg, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed);
  body<<<1, 32, 0, s>>>(c, h);
  cudaStreamEndCapture(s, nullptr);
  cudaGraphExec_t ex; cudaGraphInstantiate(&ex, g, 0);
  cudaGraphLaunch(ex, s); cudaStreamSynchronize(s);
  cudaMemcpy(&v, c, 4, cudaMemcpyDeviceToHost); printf("%d %s\n", v, cudaGetErrorString(cudaGetLastError()));
}
