// Launch-chain diagnostics on this GPU: what one link of a dependent kernel chain costs inside a CUDA graph
// (plain stream order vs programmatic dependent launch), and what the first touch of kernel parameters costs.
// Diagnostic only; results are quoted in DESIGN.md.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

struct Big { long long v[120]; };  // 960 B of parameters

// a link: records its start / end wall time (ns), spins `work` cycles in between
__global__ void link_plain(unsigned long long* ts, int i, int work) {
  if (blockIdx.x == 0 && threadIdx.x == 0) ts[2 * i] = gtime();
  long long t0 = clock64();
  while (clock64() - t0 < work) { }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) ts[2 * i + 1] = gtime();
}
__global__ void link_pdl(unsigned long long* ts, int i, int work) {
  asm volatile("griddepcontrol.launch_dependents;");  // let the next link's CTAs get scheduled as soon as possible
  asm volatile("griddepcontrol.wait;" ::: "memory");  // ... and wait here until the previous link has fully completed
  if (blockIdx.x == 0 && threadIdx.x == 0) ts[2 * i] = gtime();
  long long t0 = clock64();
  while (clock64() - t0 < work) { }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) ts[2 * i + 1] = gtime();
}

// cost of touching parameters: cycles from entry to having read k distinct 64 B-apart parameter words
__global__ void param_touch(Big b, long long* out) {
  long long t0 = clock64();
  long long a = b.v[0];
  long long t1 = clock64();
  a += b.v[16];
  long long t2 = clock64();
  a += b.v[32] + b.v[48] + b.v[64] + b.v[80] + b.v[96] + b.v[112];
  long long t3 = clock64();
  a += b.v[1] + b.v[17];
  long long t4 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = a; }
}

static void run_chain(cudaStream_t st, bool pdl, int grid, int block, int work, int n, unsigned long long* d_ts) {
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n; ++i) {
    if (!pdl) link_plain<<<grid, block, 0, st>>>(d_ts, i, work);
    else {
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = st;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, link_pdl, d_ts, i, work);
    }
  }
  cudaError_t e = cudaStreamEndCapture(st, &g);
  if (e != cudaSuccess) { printf("capture failed: %s\n", cudaGetErrorString(e)); return; }
  e = cudaGraphInstantiate(&ge, g, 0);
  if (e != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(e)); return; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
  cudaEventRecord(e0, st); cudaGraphLaunch(ge, st); cudaEventRecord(e1, st); cudaStreamSynchronize(st);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<unsigned long long> ts(2 * n);
  cudaMemcpy(ts.data(), d_ts, 16 * n, cudaMemcpyDeviceToHost);
  double gap = 0, dur = 0; int cnt = 0;
  for (int i = 10; i + 1 < n; ++i) { gap += (double)ts[2 * (i + 1)] - (double)ts[2 * i + 1]; dur += (double)ts[2 * i + 1] - (double)ts[2 * i]; ++cnt; }
  printf("graph chain %-5s grid=%4d x %3d work=%5d cyc: %.2f us per link (event), in-kernel %.2f us, gap end->next start %.2f us\n",
         pdl ? "PDL" : "plain", grid, block, work, ms * 1000.0 / n, dur / cnt / 1000.0, gap / cnt / 1000.0);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
}

int main() {
  cudaStream_t st; cudaStreamCreate(&st);
  const int n = 500;
  unsigned long long* d_ts; cudaMalloc(&d_ts, 16 * n);
  for (int work : {0, 4000, 16000})
    for (int grid : {148, 512, 2048})
      for (int pdl = 0; pdl < 2; ++pdl) run_chain(st, pdl, grid, 64, work, n, d_ts);
  // two alternating chains on two streams inside one graph are covered by the product bench (pipelines)
  long long* d_out; cudaMalloc(&d_out, 64); long long h[8];
  Big big = {};
  for (int rep = 0; rep < 3; ++rep) {
    param_touch<<<512, 64, 0, st>>>(big, d_out);
    cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost);
    printf("param touch (960 B struct): first word %lld cyc, second line %lld, six more lines %lld, two already-touched lines %lld\n", h[0], h[1], h[2], h[3]);
  }
  return 0;
}
