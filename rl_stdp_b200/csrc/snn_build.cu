// snn_build.cu — one-time, type-independent construction of the synapse topology on the device.
//
// Replaces the dense `connection`/`s` matrices of Old_net/new_net.jl:14-16,27-56 and the
// post/pre/delays tables of Old_net/DASTDP.jl:32-49 by
//   - CSR by presynaptic neuron with every row sorted by axonal delay (segment table `seg`), and
//   - the incoming (transpose) index sorted by presynaptic id, which gives both the LTP walk
//     (new_net.jl:126 `sd[:,neuron] .+= STDP`) and the bit-reproducible pull order of
//     `I .+= s[neuron,:]` (new_net.jl:116-118).
// cub radix sort is used here only: construction is not on the per-millisecond path.
#include <cub/cub.cuh>

#include "snn_common.h"

namespace snn {

void Topo::release() {
  cudaFree(seg); cudaFree(post); cudaFree(perm); cudaFree(colptr); cudaFree(col_npl);
  cudaFree(in_syn); cudaFree(in_pre);
  seg = perm = colptr = col_npl = in_syn = in_pre = nullptr;
  post = nullptr;
}

namespace {

__device__ __forceinline__ int64_t upper_row(const int64_t* rowptr, int64_t n, int64_t e) {
  // largest i with rowptr[i] <= e
  int64_t lo = 0, hi = n;  // invariant rowptr[lo] <= e < rowptr[hi]
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (rowptr[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void make_row_keys(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ post,
                              const uint8_t* __restrict__ delay, int64_t N, int64_t E, int64_t Nper,
                              int Dmax, uint64_t* __restrict__ key, uint32_t* __restrict__ val,
                              int* __restrict__ bad) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E;
       e += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = upper_row(rowptr, N, e);
    int d = delay ? delay[e] : 1;
    int64_t j = post[e];
    if (d < 1 || d > Dmax) atomicOr(bad, 1);
    if (j < 0 || j >= N || (j / Nper) != (i / Nper)) atomicOr(bad, 2);
    key[e] = ((uint64_t)i << 8) | (uint64_t)(d & 255);
    val[e] = (uint32_t)e;
  }
}

__global__ void gather_rows(const uint64_t* __restrict__ key, const uint32_t* __restrict__ perm,
                            const int32_t* __restrict__ post_in, int64_t E, int bitsN,
                            int32_t* __restrict__ post_out, uint64_t* __restrict__ key2,
                            uint32_t* __restrict__ val2) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E;
       e += (int64_t)gridDim.x * blockDim.x) {
    int32_t j = post_in[perm[e]];
    post_out[e] = j;
    uint64_t pre = key[e] >> 8;
    key2[e] = ((uint64_t)(uint32_t)j << bitsN) | pre;
    val2[e] = (uint32_t)e;
  }
}

__device__ __forceinline__ uint32_t lower_bound_u64(const uint64_t* __restrict__ a, int64_t n, uint64_t x) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return (uint32_t)lo;
}

__global__ void make_seg(const uint64_t* __restrict__ key, int64_t E, int64_t N, int Dmax,
                         uint32_t* __restrict__ seg) {
  int64_t total = N * (Dmax + 1);
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total;
       q += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = q / (Dmax + 1);
    int a = (int)(q - i * (Dmax + 1));
    uint64_t x = a == Dmax ? ((uint64_t)(i + 1) << 8) : (((uint64_t)i << 8) | (uint64_t)(a + 1));
    seg[q] = lower_bound_u64(key, E, x);
  }
}

__global__ void make_incoming(const uint64_t* __restrict__ key2, const uint32_t* __restrict__ in_syn,
                              const uint64_t* __restrict__ key_row, int64_t E, int bitsN,
                              uint32_t* __restrict__ in_pre) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < E;
       k += (int64_t)gridDim.x * blockDim.x) {
    uint32_t pre = (uint32_t)(key2[k] & ((1ull << bitsN) - 1ull));
    uint32_t d = (uint32_t)(key_row[in_syn[k]] & 255ull);
    in_pre[k] = ((d - 1u) << kPreBits) | pre;
  }
}

__global__ void make_colptr(const uint64_t* __restrict__ key2, int64_t E, int64_t N, int64_t Nper,
                            int64_t Ne, int plastic_ei, int bitsN, uint32_t* __restrict__ colptr,
                            uint32_t* __restrict__ col_npl) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j <= N;
       j += (int64_t)gridDim.x * blockDim.x) {
    uint32_t c0 = lower_bound_u64(key2, E, (uint64_t)j << bitsN);
    colptr[j] = c0;
    if (j < N) {
      int64_t base = (j / Nper) * Nper;
      bool post_exc = (j - base) < Ne;
      uint32_t npl = 0;
      if (post_exc || plastic_ei)
        npl = lower_bound_u64(key2, E, ((uint64_t)j << bitsN) | (uint64_t)(base + Ne)) - c0;
      col_npl[j] = npl;
    }
  }
}

int ilog2_ceil(int64_t x) {
  int b = 0;
  while ((1ll << b) < x) ++b;
  return b;
}

}  // namespace

int build_topology(const snn_config& cfg, const int64_t* h_rowptr, const int32_t* h_post,
                   const uint8_t* h_delay, Topo* out, cudaStream_t stream, std::string& err) {
  const int64_t N = cfg.n_neurons;
  const int64_t E = h_rowptr[N];
  if (h_rowptr[0] != 0) { err = "rowptr[0] must be 0"; return SNN_ERR_INVALID; }
  if (E < 0 || E >= (int64_t)0xFFFFFFF0ll) { err = "synapse count must be < 2^32"; return SNN_ERR_INVALID; }
  if (N > (1ll << kPreBits)) { err = "n_neurons must be <= 2^26"; return SNN_ERR_INVALID; }
  for (int64_t i = 0; i < N; ++i)
    if (h_rowptr[i + 1] < h_rowptr[i]) { err = "rowptr must be non-decreasing"; return SNN_ERR_INVALID; }
  const int Dmax = cfg.max_delay;
  const int bitsN = ilog2_ceil(N) < 1 ? 1 : ilog2_ceil(N);

  Topo t;
  t.N = N; t.E = E; t.Nper = cfg.n_per_instance; t.Ne = cfg.n_exc; t.Dmax = Dmax; t.R = Dmax + 1;
  t.B = (int32_t)(N / cfg.n_per_instance);
  const int64_t Ea = E > 0 ? E : 1;

  int64_t* d_rowptr = nullptr; int32_t* d_post_in = nullptr; uint8_t* d_delay = nullptr;
  uint64_t *d_key = nullptr, *d_key_alt = nullptr; uint32_t *d_val = nullptr, *d_val_alt = nullptr;
  int* d_bad = nullptr; void* d_tmp = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_rowptr); cudaFree(d_post_in); cudaFree(d_delay); cudaFree(d_key); cudaFree(d_key_alt);
    cudaFree(d_val); cudaFree(d_val_alt); cudaFree(d_bad); cudaFree(d_tmp);
  };
#define BUILD_OK(expr)                                                     \
  do {                                                                     \
    cudaError_t _e = (expr);                                               \
    if (_e != cudaSuccess) {                                               \
      err = std::string(#expr) + ": " + cudaGetErrorString(_e);            \
      cleanup(); t.release();                                              \
      return SNN_ERR_CUDA;                                                 \
    }                                                                      \
  } while (0)

  BUILD_OK(cudaMalloc(&d_rowptr, (N + 1) * sizeof(int64_t)));
  BUILD_OK(cudaMalloc(&d_post_in, Ea * sizeof(int32_t)));
  BUILD_OK(cudaMalloc(&d_key, Ea * sizeof(uint64_t)));
  BUILD_OK(cudaMalloc(&d_key_alt, Ea * sizeof(uint64_t)));
  BUILD_OK(cudaMalloc(&d_val, Ea * sizeof(uint32_t)));
  BUILD_OK(cudaMalloc(&d_val_alt, Ea * sizeof(uint32_t)));
  BUILD_OK(cudaMalloc(&d_bad, sizeof(int)));
  BUILD_OK(cudaMemsetAsync(d_bad, 0, sizeof(int), stream));
  BUILD_OK(cudaMemcpyAsync(d_rowptr, h_rowptr, (N + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
  BUILD_OK(cudaMemcpyAsync(d_post_in, h_post, E * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  if (h_delay) {
    BUILD_OK(cudaMalloc(&d_delay, Ea));
    BUILD_OK(cudaMemcpyAsync(d_delay, h_delay, E, cudaMemcpyHostToDevice, stream));
  }
  BUILD_OK(cudaMalloc(&t.seg, N * (Dmax + 1) * sizeof(uint32_t)));
  BUILD_OK(cudaMalloc(&t.post, Ea * sizeof(int32_t)));
  BUILD_OK(cudaMalloc(&t.perm, Ea * sizeof(uint32_t)));
  BUILD_OK(cudaMalloc(&t.colptr, (N + 1) * sizeof(uint32_t)));
  BUILD_OK(cudaMalloc(&t.col_npl, N * sizeof(uint32_t)));
  BUILD_OK(cudaMalloc(&t.in_syn, Ea * sizeof(uint32_t)));
  BUILD_OK(cudaMalloc(&t.in_pre, Ea * sizeof(uint32_t)));

  const int threads = 256;
  const int blocks = 148 * 8;
  make_row_keys<<<blocks, threads, 0, stream>>>(d_rowptr, d_post_in, d_delay, N, E, cfg.n_per_instance,
                                                Dmax, d_key, d_val, d_bad);
  BUILD_OK(cudaGetLastError());
  int h_bad = 0;
  BUILD_OK(cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
  BUILD_OK(cudaStreamSynchronize(stream));
  if (h_bad) {
    err = (h_bad & 1) ? "delay outside 1..max_delay" : "post id outside the presynaptic neuron's instance";
    cleanup(); t.release();
    return SNN_ERR_INVALID;
  }

  // 1) rows sorted by (pre, delay), stable in the caller's order
  {
    cub::DoubleBuffer<uint64_t> kb(d_key, d_key_alt);
    cub::DoubleBuffer<uint32_t> vb(d_val, d_val_alt);
    size_t tmp_bytes = 0;
    BUILD_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int64_t)E, 0, 8 + bitsN, stream));
    BUILD_OK(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 16));
    BUILD_OK(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, kb, vb, (int64_t)E, 0, 8 + bitsN, stream));
    BUILD_OK(cudaFree(d_tmp)); d_tmp = nullptr;
    if (kb.Current() != d_key) { std::swap(d_key, d_key_alt); }
    if (vb.Current() != d_val) { std::swap(d_val, d_val_alt); }
  }
  BUILD_OK(cudaMemcpyAsync(t.perm, d_val, E * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
  make_seg<<<blocks, threads, 0, stream>>>(d_key, E, N, Dmax, t.seg);
  BUILD_OK(cudaGetLastError());

  // 2) incoming index sorted by (post, pre), stable in the internal order
  uint64_t *d_key2 = nullptr, *d_key2_alt = nullptr;
  BUILD_OK(cudaMalloc(&d_key2, Ea * sizeof(uint64_t)));
  BUILD_OK(cudaMalloc(&d_key2_alt, Ea * sizeof(uint64_t)));
  auto cleanup2 = [&]() { cudaFree(d_key2); cudaFree(d_key2_alt); };
  gather_rows<<<blocks, threads, 0, stream>>>(d_key, t.perm, d_post_in, E, bitsN, t.post, d_key2, d_val_alt);
  {
    cudaError_t e0 = cudaGetLastError();
    if (e0 != cudaSuccess) { err = cudaGetErrorString(e0); cleanup(); cleanup2(); t.release(); return SNN_ERR_CUDA; }
    cub::DoubleBuffer<uint64_t> kb(d_key2, d_key2_alt);
    cub::DoubleBuffer<uint32_t> vb(d_val_alt, d_val);
    size_t tmp_bytes = 0;
    cudaError_t e1 = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int64_t)E, 0, 2 * bitsN, stream);
    if (e1 == cudaSuccess) e1 = cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 16);
    if (e1 == cudaSuccess) e1 = cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, kb, vb, (int64_t)E, 0, 2 * bitsN, stream);
    if (e1 == cudaSuccess)
      e1 = cudaMemcpyAsync(t.in_syn, vb.Current(), E * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream);
    if (e1 != cudaSuccess) { err = cudaGetErrorString(e1); cleanup(); cleanup2(); t.release(); return SNN_ERR_CUDA; }
    const uint64_t* k2 = kb.Current();
    make_incoming<<<blocks, threads, 0, stream>>>(k2, t.in_syn, d_key, E, bitsN, t.in_pre);
    make_colptr<<<blocks, threads, 0, stream>>>(k2, E, N, cfg.n_per_instance, cfg.n_exc, cfg.plastic_ei,
                                                bitsN, t.colptr, t.col_npl);
    e1 = cudaGetLastError();
    if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(stream);
    if (e1 != cudaSuccess) { err = cudaGetErrorString(e1); cleanup(); cleanup2(); t.release(); return SNN_ERR_CUDA; }
  }
  cleanup2();
  cleanup();
#undef BUILD_OK
  *out = t;
  return SNN_OK;
}

void SortScratch::release() {
  if (tmp) cudaFree(tmp);
  if (out) cudaFree(out);
  tmp = nullptr; out = nullptr; tmp_cap = out_cap = 0;
}

// Spike-record formatting: ascending ids inside every step.  The scratch buffers are grow-only and
// owned by the engine: a cudaMalloc/cudaFree pair per window costs more than the sort itself.
int sort_segments(int32_t* d_keys, int64_t n_items, const int32_t* d_offsets, int32_t n_segments,
                  cudaStream_t stream, SortScratch* sc, std::string& err) {
  if (n_items <= 0 || n_segments <= 0) return SNN_OK;
  const size_t out_bytes = (size_t)n_items * sizeof(int32_t);
  if (out_bytes > sc->out_cap) {
    if (sc->out) cudaFree(sc->out);
    sc->out = nullptr; sc->out_cap = 0;
    SNN_CUDA_OK(cudaMalloc(&sc->out, out_bytes + out_bytes / 2 + 256));
    sc->out_cap = out_bytes + out_bytes / 2 + 256;
  }
  size_t tmp_bytes = 0;
  cudaError_t e = cub::DeviceSegmentedSort::SortKeys(nullptr, tmp_bytes, d_keys, sc->out, (int)n_items,
                                                     n_segments, d_offsets, d_offsets + 1, stream);
  if (e == cudaSuccess && tmp_bytes > sc->tmp_cap) {
    if (sc->tmp) cudaFree(sc->tmp);
    sc->tmp = nullptr; sc->tmp_cap = 0;
    e = cudaMalloc(&sc->tmp, tmp_bytes + tmp_bytes / 2 + 256);
    if (e == cudaSuccess) sc->tmp_cap = tmp_bytes + tmp_bytes / 2 + 256;
  }
  if (e == cudaSuccess) {
    tmp_bytes = sc->tmp_cap;
    e = cub::DeviceSegmentedSort::SortKeys(sc->tmp, tmp_bytes, d_keys, sc->out, (int)n_items, n_segments,
                                           d_offsets, d_offsets + 1, stream);
  }
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d_keys, sc->out, out_bytes, cudaMemcpyDeviceToDevice, stream);
  if (e != cudaSuccess) { err = std::string("sort_segments: ") + cudaGetErrorString(e); return SNN_ERR_CUDA; }
  return SNN_OK;
}

}  // namespace snn
