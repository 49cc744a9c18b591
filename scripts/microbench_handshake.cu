// Bounded experiment ("next" 6): what does one simulation's hand-over cost when the search kernel stays
// RESIDENT for the whole move and exchanges flags with the user's leaf kernels, instead of being re-launched per simulation?
//
//   resident kernel R (grid x 64 threads, like k_sim):  for s in 0..S-1:  work(w_search)  ->  grid-wide arrive, the last
//       CTA stores flagA = s+1 (release)  ->  every CTA polls flagB >= s+1 (acquire)
//   leaf stream:  for s in 0..S-1:  cuStreamWaitValue32(flagA >= s+1)  ->  leaf kernel (ordinary launch, work(w_leaf))
//       ->  cuStreamWriteValue32(flagB = s+1)
//
// compared with the chain the product uses today: leaf kernel -> search kernel -> leaf kernel ... in one stream
// (ordinary launches and programmatic dependent launches), all inside CUDA graphs where the API allows it.
// The per-simulation period minus (w_search + w_leaf) is the hand-over cost of each scheme.
// Every spin loop is bounded (2 s by %globaltimer), so a scheme that cannot make progress ends instead of hanging the GPU.
//
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench_handshake microbench_handshake.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
#define CKD(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_ = nullptr; cuGetErrorString(r_, &s_); printf("driver error %s at %s:%d\n", s_ ? s_ : "?", __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void spin_cycles(int work) { const long long t0 = clock64(); while (clock64() - t0 < work) { } }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

struct Flags {
  unsigned flagA;      // resident kernel -> leaf stream: "select of simulation s is out"
  unsigned pad0[31];
  unsigned flagB;      // leaf stream -> resident kernel: "leaf results of simulation s are in"
  unsigned pad1[31];
  unsigned arrive;     // grid-wide arrival counter of the resident kernel
  unsigned timed_out;
  unsigned pad2[30];
  unsigned go;         // host -> resident kernel: the leaf stream's schedule is enqueued, start
};

// the resident search kernel
__global__ void __launch_bounds__(64) resident(Flags* f, int S, int work, unsigned long long* ts) {
  const unsigned long long deadline = gtime() + 300000000ull;  // 0.3 s
  if (threadIdx.x == 0) {
    while (ld_acquire(&f->go) == 0) {
      if (gtime() > deadline) { f->timed_out = 2; break; }
    }
  }
  __syncthreads();
  if (f->timed_out) return;
  for (int s = 0; s < S; ++s) {
    spin_cycles(work);  // select of simulation s (plus expand + backprop of s-1)
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned prev = atomicAdd(&f->arrive, 1u);
      if (prev == (unsigned)(s + 1) * gridDim.x - 1) {  // last CTA of this round
        ts[2 * s] = gtime();
        st_release(&f->flagA, (unsigned)(s + 1));
      }
      while (ld_acquire(&f->flagB) < (unsigned)(s + 1)) {
        if (gtime() > deadline) { f->timed_out = 1; break; }
      }
      if (blockIdx.x == 0) ts[2 * s + 1] = gtime();
    }
    __syncthreads();
    if (f->timed_out) return;
  }
}

__global__ void __launch_bounds__(64) leaf(int work, volatile unsigned* sink) {
  spin_cycles(work);
  if (sink && threadIdx.x == 0 && blockIdx.x == 0) *sink = 1;
}
__global__ void __launch_bounds__(64) leaf_pdl(int work, volatile unsigned* sink) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
  spin_cycles(work);
  if (sink && threadIdx.x == 0 && blockIdx.x == 0) *sink = 1;
}
__global__ void __launch_bounds__(64) search_pdl(int work_pre, int work, volatile unsigned* sink) {
  spin_cycles(work_pre);  // what k_sim does before griddepcontrol.wait (overlaps the leaf)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  spin_cycles(work);
  asm volatile("griddepcontrol.launch_dependents;");
  if (sink && threadIdx.x == 0 && blockIdx.x == 0) *sink = 2;
}

static void launch_pdl(cudaStream_t st, int grid, void (*k)(int, volatile unsigned*), int work, unsigned* sink) {
  cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.stream = st;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, k, work, (volatile unsigned*)sink));
}

static double time_graph(cudaStream_t st, cudaGraph_t g, int reps) {
  cudaGraphExec_t ge; CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(e0, st));
  for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGraphExecDestroy(ge));
  return ms / reps;
}

// after a run of the resident scheme: if the resident kernel gave up, the leaf stream still waits for flags that will never
// come -- release it from the host so that the streams drain
static void drain(Flags* f, cudaStream_t sa, cudaStream_t sb) {
  CK(cudaStreamSynchronize(sa));
  const unsigned big = 0x7fffffffu;
  CK(cudaMemcpy(&f->flagA, &big, 4, cudaMemcpyHostToDevice));
  CK(cudaStreamSynchronize(sb));
}

int main(int argc, char** argv) {
  setvbuf(stdout, nullptr, _IONBF, 0);
  const int S = 128, grid = 512;
  CK(cudaSetDevice(0));
  CKD(cuInit(0));
  cudaStream_t sa, sb, sc; CK(cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking));
  Flags* f; CK(cudaMalloc(&f, sizeof(Flags)));
  unsigned* sink; CK(cudaMalloc(&sink, 64));
  unsigned long long* ts; CK(cudaMalloc(&ts, 16 * S));
  int clock_khz = 0; CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  printf("device clock %d kHz; S = %d simulations, grid %d x 64 threads\n", clock_khz, S, grid);

  for (int wl : {0, 4000}) {        // leaf work in cycles (4000 ~ 2 us, the synthetic leaf of configs[1])
    for (int ws : {0, 6000}) {      // search work in cycles (6000 ~ 3 us)
      // ---- (1) today's chain: ordinary launches in one stream, in a graph -------------------------------------
      {
        cudaGraph_t g; CK(cudaStreamBeginCapture(sa, cudaStreamCaptureModeThreadLocal));
        for (int s = 0; s < S; ++s) { leaf<<<grid, 64, 0, sa>>>(wl, sink); leaf<<<grid, 64, 0, sa>>>(ws, sink); }
        CK(cudaStreamEndCapture(sa, &g));
        const double ms = time_graph(sa, g, 5);
        printf("leaf %5d cyc, search %5d cyc | ordinary chain (graph)      : %6.2f us per simulation\n", wl, ws, ms * 1e3 / S);
        CK(cudaGraphDestroy(g));
      }
      // ---- (2) programmatic dependent launches, in a graph ------------------------------------------------------
      {
        cudaGraph_t g; CK(cudaStreamBeginCapture(sa, cudaStreamCaptureModeThreadLocal));
        for (int s = 0; s < S; ++s) {
          launch_pdl(sa, grid, leaf_pdl, wl, sink);
          cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.stream = sa;
          cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          CK(cudaLaunchKernelEx(&cfg, search_pdl, 0, ws, (volatile unsigned*)sink));
        }
        CK(cudaStreamEndCapture(sa, &g));
        const double ms = time_graph(sa, g, 5);
        printf("leaf %5d cyc, search %5d cyc | programmatic chain (graph)  : %6.2f us per simulation\n", wl, ws, ms * 1e3 / S);
        CK(cudaGraphDestroy(g));
      }
      // ---- (3) resident search kernel + stream memory operations (plain streams; enqueued ahead of time) --------
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(f, 0, sizeof(Flags)));
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        // the resident kernel is launched FIRST and the leaf stream's schedule enqueued while it runs (a stream whose head is a
        // blocked wait, queued ahead of the kernel, kept the kernel from starting at all on this driver: first attempt hung)
        CK(cudaEventRecord(e0, sa));
        resident<<<grid, 64, 0, sa>>>(f, S, ws, ts);
        CK(cudaEventRecord(e1, sa));
        for (int s = 0; s < S; ++s) {
          CKD(cuStreamWaitValue32((CUstream)sb, (CUdeviceptr)&f->flagA, (cuuint32_t)(s + 1), CU_STREAM_WAIT_VALUE_GEQ));
          leaf<<<grid, 64, 0, sb>>>(wl, sink);
          CKD(cuStreamWriteValue32((CUstream)sb, (CUdeviceptr)&f->flagB, (cuuint32_t)(s + 1), CU_STREAM_WRITE_VALUE_DEFAULT));
        }
        { const unsigned one = 1; CK(cudaMemcpyAsync(&f->go, &one, 4, cudaMemcpyHostToDevice, sc)); CK(cudaStreamSynchronize(sc)); }
        CK(cudaStreamSynchronize(sa));
        Flags hf; CK(cudaMemcpy(&hf, f, sizeof(Flags), cudaMemcpyDeviceToHost));
        drain(f, sa, sb);
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (hf.timed_out && rep == 2) printf("    (gave up: flagA %u flagB %u arrivals %u)\n", hf.flagA, hf.flagB, hf.arrive);
        std::vector<unsigned long long> h(2 * S);
        CK(cudaMemcpy(h.data(), ts, 16 * S, cudaMemcpyDeviceToHost));
        double wait = 0; int cnt = 0;
        for (int s = 16; s < S; ++s) { wait += (double)h[2 * s + 1] - (double)h[2 * s]; ++cnt; }
        const double period = ((double)h[2 * (S - 1)] - (double)h[2 * 16]) / (S - 1 - 16) / 1e3;
        (void)ms;
        if (rep == 2)
          printf("leaf %5d cyc, search %5d cyc | resident + stream memops    : %6.2f us per simulation; flagA out -> flagB seen %.2f us%s\n",
                 wl, ws, period, wait / cnt / 1e3, hf.timed_out ? "  [TIMED OUT]" : "");
        CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
      }
    }
  }
  // ---- (4) the same resident scheme with the leaf stream's schedule captured in a graph (batch mem-op nodes) --------
  {
    const int wl = 4000, ws = 6000;
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamBeginCapture(sb, cudaStreamCaptureModeThreadLocal);
    bool ok = e == cudaSuccess;
    for (int s = 0; s < S && ok; ++s) {
      ok = ok && cuStreamWaitValue32((CUstream)sb, (CUdeviceptr)&f->flagA, (cuuint32_t)(s + 1), CU_STREAM_WAIT_VALUE_GEQ) == CUDA_SUCCESS;
      leaf<<<grid, 64, 0, sb>>>(wl, sink);
      ok = ok && cuStreamWriteValue32((CUstream)sb, (CUdeviceptr)&f->flagB, (cuuint32_t)(s + 1), CU_STREAM_WRITE_VALUE_DEFAULT) == CUDA_SUCCESS;
    }
    e = cudaStreamEndCapture(sb, &g);
    if (!ok || e != cudaSuccess || !g) {
      printf("stream memory operations could not be captured into a graph (%s)\n", cudaGetErrorString(e));
      cudaGetLastError();
    } else {
      cudaGraphExec_t ge; CK(cudaGraphInstantiate(&ge, g, 0));
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(f, 0, sizeof(Flags)));
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0, sa));
        resident<<<grid, 64, 0, sa>>>(f, S, ws, ts);
        CK(cudaEventRecord(e1, sa));
        CK(cudaGraphLaunch(ge, sb));
        { const unsigned one = 1; CK(cudaMemcpyAsync(&f->go, &one, 4, cudaMemcpyHostToDevice, sc)); CK(cudaStreamSynchronize(sc)); }
        CK(cudaStreamSynchronize(sa));
        Flags hf; CK(cudaMemcpy(&hf, f, sizeof(Flags), cudaMemcpyDeviceToHost));
        drain(f, sa, sb);
        std::vector<unsigned long long> h(2 * S);
        CK(cudaMemcpy(h.data(), ts, 16 * S, cudaMemcpyDeviceToHost));
        const double period = ((double)h[2 * (S - 1)] - (double)h[2 * 16]) / (S - 1 - 16) / 1e3;
        if (rep == 2)
          printf("leaf %5d cyc, search %5d cyc | resident + memop graph nodes: %6.2f us per simulation%s\n", wl, ws, period,
                 hf.timed_out ? "  [TIMED OUT]" : "");
      }
      CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
    }
  }
  return 0;
}
