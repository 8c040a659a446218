// Does cudaGraphExecUpdate accept a re-captured LM-loop graph (nested WHILE nodes, NEW conditional handles, new kernel arguments)?
// If it does, a new window of the same topology costs an update instead of a cudaGraphInstantiate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/cgraph_update_test tools/ubench/cgraph_update_test.cu
#include <chrono>
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
  const bool again = s->inner < 1 + (s->it % 3);
  cudaGraphSetConditional(inner, again);
  if (!again) {
    s->it++;
    cudaGraphSetConditional(outer, s->it < iters);
  }
}
static cudaStream_t st, side;
static cudaEvent_t ef, ej;
static int add_while(cudaGraph_t parent, cudaGraphConditionalHandle hd, cudaGraph_t *body) {
  cudaStreamCaptureStatus cs;
  const cudaGraphNode_t *deps = nullptr;
  size_t ndeps = 0;
  CK(cudaStreamGetCaptureInfo(st, &cs, nullptr, nullptr, &deps, &ndeps));
  cudaGraphNodeParams p = {};
  p.type = cudaGraphNodeTypeConditional;
  p.conditional.handle = hd;
  p.conditional.type = cudaGraphCondTypeWhile;
  p.conditional.size = 1;
  cudaGraphNode_t node;
  CK(cudaGraphAddNode(&node, parent, deps, ndeps, &p));
  *body = p.conditional.phGraph_out[0];
  CK(cudaStreamUpdateCaptureDependencies(st, &node, 1, cudaStreamSetCaptureDependencies));
  return 0;
}
// the structure of ppo_engine.cu: build_lm_graph
static int build(St *s, int *scratch, int iters, int grid, cudaGraph_t *out) {
  cudaGraph_t G, tmp, body_out, body_in;
  CK(cudaGraphCreate(&G, 0));
  cudaGraphConditionalHandle hout, hin;
  CK(cudaGraphConditionalHandleCreate(&hout, G, 0, 0));
  CK(cudaStreamBeginCaptureToGraph(st, G, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  k_begin<<<1, 1, 0, st>>>(s, hout);
  if (add_while(G, hout, &body_out)) return 1;
  CK(cudaStreamEndCapture(st, &tmp));
  CK(cudaGraphConditionalHandleCreate(&hin, body_out, 0, 0));
  CK(cudaStreamBeginCaptureToGraph(st, body_out, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  CK(cudaMemsetAsync(scratch, 0, 1024, st));
  CK(cudaEventRecord(ef, st));
  CK(cudaStreamWaitEvent(side, ef, 0));
  k_work<<<grid, 1, 0, st>>>(s);
  k_side<<<1, 1, 0, side>>>(s);
  CK(cudaEventRecord(ej, side));
  CK(cudaStreamWaitEvent(st, ej, 0));
  k_iter_begin<<<1, 1, 0, st>>>(s, hin);
  if (add_while(body_out, hin, &body_in)) return 1;
  CK(cudaStreamEndCapture(st, &tmp));
  CK(cudaStreamBeginCaptureToGraph(st, body_in, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  CK(cudaMemsetAsync(scratch, 0, 1024, st));
  k_work<<<grid, 1, 0, st>>>(s);
  k_decide<<<1, 1, 0, st>>>(s, hin, hout, iters);
  CK(cudaStreamEndCapture(st, &tmp));
  *out = G;
  return 0;
}
static int run_check(cudaGraphExec_t X, St *s, int iters, int grid, const char *what) {
  CK(cudaMemset(s, 0xff, sizeof(St)));
  CK(cudaGraphLaunch(X, st));
  CK(cudaStreamSynchronize(st));
  St h;
  CK(cudaMemcpy(&h, s, sizeof h, cudaMemcpyDeviceToHost));
  int trials = 0;
  for (int i = 0; i < iters; i++) trials += 1 + i % 3;
  const bool ok = h.it == iters && h.trials_total == trials && h.side == iters && h.work == grid * (iters + trials);
  printf("%s: iterations %d (expect %d), trials %d (expect %d), side %d, work %d (expect %d)  %s\n", what, h.it, iters, h.trials_total, trials, h.side, h.work,
         grid * (iters + trials), ok ? "OK" : "FAIL");
  return ok ? 0 : 1;
}
int main() {
  St *s1, *s2;
  int *scratch1, *scratch2;
  CK(cudaMalloc(&s1, sizeof(St)));
  CK(cudaMalloc(&s2, sizeof(St)));
  CK(cudaMalloc(&scratch1, 1024));
  CK(cudaMalloc(&scratch2, 1024));
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&ef, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ej, cudaEventDisableTiming));
  cudaGraph_t G1, G2, G3;
  if (build(s1, scratch1, 7, 1, &G1)) return 1;
  cudaGraphExec_t X;
  auto t0 = std::chrono::steady_clock::now();
  CK(cudaGraphInstantiate(&X, G1, 0));
  auto t1 = std::chrono::steady_clock::now();
  printf("instantiate: %.3f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count());
  int bad = run_check(X, s1, 7, 1, "graph 1");
  // a second window: other buffers, other iteration count, other grid -- same topology, NEW conditional handles
  if (build(s2, scratch2, 5, 3, &G2)) return 1;
  cudaGraphExecUpdateResultInfo info;
  t0 = std::chrono::steady_clock::now();
  cudaError_t e = cudaGraphExecUpdate(X, G2, &info);
  t1 = std::chrono::steady_clock::now();
  printf("cudaGraphExecUpdate: %s (result %d), %.3f ms\n", cudaGetErrorString(e), (int)info.result, std::chrono::duration<double, std::milli>(t1 - t0).count());
  if (e == cudaSuccess) {
    CK(cudaMemset(s1, 0, sizeof(St)));
    bad += run_check(X, s2, 5, 3, "updated exec, graph 2");
    St h;
    CK(cudaMemcpy(&h, s1, sizeof h, cudaMemcpyDeviceToHost));
    printf("old buffers untouched: %s\n", (h.it == 0 && h.work == 0) ? "yes" : "NO");
    bad += !(h.it == 0 && h.work == 0);
    // the source graph of the update may be destroyed afterwards?
    CK(cudaGraphDestroy(G2));
    bad += run_check(X, s2, 5, 3, "after destroying graph 2");
    // and once more
    if (build(s1, scratch1, 9, 2, &G3)) return 1;
    e = cudaGraphExecUpdate(X, G3, &info);
    printf("second update: %s (result %d)\n", cudaGetErrorString(e), (int)info.result);
    if (e == cudaSuccess) bad += run_check(X, s1, 9, 2, "updated exec, graph 3");
  } else {
    cudaGetLastError();
  }
  printf(bad ? "FAILED\n" : "ALL OK\n");
  return bad;
}
