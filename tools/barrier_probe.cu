// barrier_probe.cu — how long does a device-wide hand-off cost on this GPU?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/barrier_probe tools/barrier_probe.cu
// Measures, for a persistent grid of one block per SM:
//   (a) a counter barrier (every block: fence + atomicAdd, then spin on the counter) per round;
//   (b) the same plus an L2 round trip of payload (each block reads a value another block wrote
//       before the barrier) — the pattern of the persistent neuron chain;
//   (c) producer kernel -> consumer kernel flag hand-off between two co-resident kernels.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}
// every spin gives up after 2 s: a probe must never hang the box
__device__ unsigned g_abort = 0;
#define SPIN_WHILE(cond) { const unsigned long long t0_ = gtime(); while (cond) { \
  if (*(volatile unsigned*)&g_abort) break; if (gtime() - t0_ > 500000000ull) { g_abort = 1u; break; } } }

__global__ void barrier_rounds(unsigned* ctr, float* payload, int rounds, int with_payload, float* sink) {
  const unsigned nb = gridDim.x;
  float acc = 0.f;
  for (int r = 0; r < rounds; ++r) {
    if (with_payload) payload[(size_t)r % 2 * nb * blockDim.x + blockIdx.x * blockDim.x + threadIdx.x] = acc + r;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(ctr, 1u);
      SPIN_WHILE(ld_acquire(ctr) < nb * (unsigned)(r + 1))
    }
    __syncthreads();
    if (with_payload) {
      const unsigned ob = (blockIdx.x + 37u) % nb;
      acc += __ldcg(&payload[(size_t)r % 2 * nb * blockDim.x + ob * blockDim.x + threadIdx.x]);
    }
  }
  if (with_payload) sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// two kernels, co-resident: ping raises flag[0] = r+1, pong answers flag[1] = r+1
__global__ void ping(unsigned* flag, int rounds) {
  if (threadIdx.x == 0)
    for (int r = 0; r < rounds; ++r) {
      __threadfence();
      atomicExch(&flag[0], (unsigned)(r + 1));
      SPIN_WHILE(ld_acquire(&flag[1]) < (unsigned)(r + 1))
    }
}
__global__ void pong(unsigned* flag, int rounds) {
  if (threadIdx.x == 0)
    for (int r = 0; r < rounds; ++r) {
      SPIN_WHILE(ld_acquire(&flag[0]) < (unsigned)(r + 1))
      __threadfence();
      atomicExch(&flag[1], (unsigned)(r + 1));
    }
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* ctr; float *payload, *sink;
  cudaMalloc(&ctr, 64);
  cudaMalloc(&payload, sizeof(float) * 2 * sms * 1024);
  cudaMalloc(&sink, sizeof(float) * sms * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int rounds = 1000;
  for (int threads : {128, 512, 1024})
    for (int wp = 0; wp < 2; ++wp)
      for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(ctr, 0, 64);
        cudaEventRecord(e0);
        barrier_rounds<<<sms, threads>>>(ctr, payload, rounds, wp, sink);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("barrier blocks=%d threads=%d payload=%d: %.3f us per round (%s)\n", sms, threads, wp,
                        1e3 * ms / rounds, cudaGetErrorString(e));
      }
  // two blocks per SM
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(ctr, 0, 64);
    cudaEventRecord(e0);
    barrier_rounds<<<2 * sms, 256>>>(ctr, payload, rounds, 0, sink);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("barrier blocks=%d threads=256 payload=0: %.3f us per round\n", 2 * sms, 1e3 * ms / rounds);
  }
  cudaStream_t s1, s2;
  cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(ctr, 0, 64);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, s1);
    ping<<<1, 32, 0, s1>>>(ctr, rounds);
    pong<<<1, 32, 0, s2>>>(ctr, rounds);
    cudaEventRecord(e1, s1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("ping-pong between two kernels: %.3f us per round trip (2 hand-offs)\n", 1e3 * ms / rounds);
  }
  unsigned ab = 0;
  cudaMemcpyFromSymbol(&ab, g_abort, sizeof(ab));
  printf("abort flag (a spin timed out): %u\n", ab);
  return 0;
}
