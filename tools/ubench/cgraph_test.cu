// Prototype: nested WHILE conditional graph nodes whose bodies are stream-captured (with a forked side stream and a memset inside a body),
// the structure the device-side LM controller uses.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/cgraph_test tools/ubench/cgraph_test.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return 1; } } while (0)
struct St { int it, trials_total, inner, side, work; };
__global__ void k_begin(St *s, cudaGraphConditionalHandle outer) { s->it = 0, s->trials_total = 0, s->side = 0, s->work = 0; cudaGraphSetConditional(outer, 1); }
__global__ void k_iter_begin(St *s, cudaGraphConditionalHandle inner) { s->inner = 0; cudaGraphSetConditional(inner, 1); }
__global__ void k_work(St *s) { atomicAdd(&s->work, 1); }
__global__ void k_side(St *s) { atomicAdd(&s->side, 1); }
__global__ void k_decide(St *s, cudaGraphConditionalHandle inner, cudaGraphConditionalHandle outer, int iters) {
  s->inner++;
  s->trials_total++;
  const bool again = s->inner < 1 + (s->it % 3);  // 1, 2, 3, 1, ... trials per iteration
  cudaGraphSetConditional(inner, again);
  if (!again) {
    s->it++;
    cudaGraphSetConditional(outer, s->it < iters);
  }
}
static int add_while(cudaGraph_t parent, cudaGraphNode_t *dep, int ndep, cudaGraphConditionalHandle *h, cudaGraph_t *body, cudaGraphNode_t *node) {
  CK(cudaGraphConditionalHandleCreate(h, parent, 0, 0));
  cudaGraphNodeParams p = {};
  p.type = cudaGraphNodeTypeConditional;
  p.conditional.handle = *h;
  p.conditional.type = cudaGraphCondTypeWhile;
  p.conditional.size = 1;
  CK(cudaGraphAddNode(node, parent, dep, ndep, &p));
  *body = p.conditional.phGraph_out[0];
  return 0;
}
int main() {
  St *s;
  int *scratch;
  CK(cudaMalloc(&s, sizeof(St)));
  CK(cudaMalloc(&scratch, 1024));
  cudaStream_t st, side;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
  cudaEvent_t ef, ej;
  CK(cudaEventCreateWithFlags(&ef, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ej, cudaEventDisableTiming));
  const int iters = 7;
  cudaGraph_t G;
  CK(cudaGraphCreate(&G, 0));
  // the outer handle must exist before the kernel that sets it is added
  cudaGraphConditionalHandle hout, hin;
  CK(cudaGraphConditionalHandleCreate(&hout, G, 0, 0));
  cudaGraphNode_t nbegin;
  {
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeKernel;
    void *args[2] = {&s, &hout};
    p.kernel.func = (void *)k_begin;
    p.kernel.gridDim = dim3(1), p.kernel.blockDim = dim3(1);
    p.kernel.kernelParams = args;
    CK(cudaGraphAddNode(&nbegin, G, nullptr, 0, &p));
  }
  cudaGraph_t body_out;
  cudaGraphNode_t nwhile;
  {
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = hout;
    p.conditional.type = cudaGraphCondTypeWhile;
    p.conditional.size = 1;
    CK(cudaGraphAddNode(&nwhile, G, &nbegin, 1, &p));
    body_out = p.conditional.phGraph_out[0];
  }
  // outer body: captured from the stream: work kernels with a fork/join, iter_begin, then the inner WHILE (added explicitly), whose body is captured again
  CK(cudaGraphConditionalHandleCreate(&hin, body_out, 0, 0));
  CK(cudaStreamBeginCaptureToGraph(st, body_out, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  CK(cudaMemsetAsync(scratch, 0, 1024, st));
  CK(cudaEventRecord(ef, st));
  CK(cudaStreamWaitEvent(side, ef, 0));
  k_work<<<1, 1, 0, st>>>(s);
  k_side<<<1, 1, 0, side>>>(s);
  CK(cudaEventRecord(ej, side));
  CK(cudaStreamWaitEvent(st, ej, 0));
  k_iter_begin<<<1, 1, 0, st>>>(s, hin);
  cudaGraph_t tmp;
  // leaf nodes of the capture so far = dependencies of the inner while
  cudaStreamCaptureStatus cs;
  const cudaGraphNode_t *deps = nullptr;
  size_t ndeps = 0;
  CK(cudaStreamGetCaptureInfo(st, &cs, nullptr, nullptr, &deps, &ndeps));
  cudaGraphNode_t ninner;
  cudaGraph_t body_in;
  {
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = hin;
    p.conditional.type = cudaGraphCondTypeWhile;
    p.conditional.size = 1;
    CK(cudaGraphAddNode(&ninner, body_out, deps, ndeps, &p));
    body_in = p.conditional.phGraph_out[0];
  }
  CK(cudaStreamUpdateCaptureDependencies(st, &ninner, 1, cudaStreamSetCaptureDependencies));
  CK(cudaStreamEndCapture(st, &tmp));
  // inner body
  CK(cudaStreamBeginCaptureToGraph(st, body_in, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  CK(cudaMemsetAsync(scratch, 0, 1024, st));
  k_work<<<1, 1, 0, st>>>(s);
  k_decide<<<1, 1, 0, st>>>(s, hin, hout, iters);
  CK(cudaStreamEndCapture(st, &tmp));
  cudaGraphExec_t X;
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0), cudaEventCreate(&t1);
  CK(cudaGraphInstantiate(&X, G, 0));
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(t0, st);
    CK(cudaGraphLaunch(X, st));
    cudaEventRecord(t1, st);
    CK(cudaStreamSynchronize(st));
    St h;
    CK(cudaMemcpy(&h, s, sizeof h, cudaMemcpyDeviceToHost));
    float ms;
    cudaEventElapsedTime(&ms, t0, t1);
    const int expect_trials = 1 + 2 + 3 + 1 + 2 + 3 + 1;
    printf("rep %d: iterations %d (expect %d), trials %d (expect %d), side %d (expect %d), work %d (expect %d), %.3f ms  %s\n", rep, h.it, iters, h.trials_total,
           expect_trials, h.side, iters, h.work, iters + expect_trials, ms,
           (h.it == iters && h.trials_total == expect_trials && h.side == iters && h.work == iters + expect_trials) ? "OK" : "FAIL");
  }
  return 0;
}
