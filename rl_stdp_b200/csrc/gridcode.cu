// gridcode.cu — the grid-cell state encoders of Julia/Modules/Gridcells on the device:
//   input_spikes(obs, grid_cells::Array{HexGrid})    StateSpikeGaussGrid.jl:83-109
//   latency_spikes(obs, grid_cells, win_size)        StateSpikeGaussGrid.jl:111-139
// for a BATCH of observations (a population of nets evaluated together, Latencysim_rewardv0.jl:97-152 per net).
// The grids themselves (HexGrid.jl, GridGenerator.jl) are built by the caller — Julia stays the host language — and
// uploaded once.  For every (observation, grid, position | velocity) one warp finds the FIRST node of the grid, in the
// grid's own order, closer than reach * node_size (`filter(...)[1]`, :95-96,100,105) and emits
// gauss(d) = 0.999999 * exp(-2 d / node_size) (:9).  fp64, no contraction; exp is the fdlibm algorithm written out
// in IEEE operations, the same ones as oracle/lif_oracle.c::ora_exp, so the checker and the device agree bit for bit.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/snn_b200.h"

#define DMUL(a, b) __dmul_rn((a), (b))
#define DADD(a, b) __dadd_rn((a), (b))
#define DSUB(a, b) __dsub_rn((a), (b))

namespace snn {

// exp(x) for |x| < 700, fdlibm e_exp.c (k = round(x / ln2), r = x - k ln2 in two pieces, degree-5 rational kernel)
__device__ __forceinline__ double exp_fdlibm(double x) {
  const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10, invln2 = 1.44269504088896338700e+00,
               P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
               P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
  const double ax = fabs(x);
  double hi = 0.0, lo = 0.0;
  int k = 0;
  if (ax > 0.34657359027997264) {  // 0.5 ln2
    if (ax < 1.0397207708399179) { // 1.5 ln2
      if (x > 0) { hi = DSUB(x, ln2HI); lo = ln2LO; k = 1; } else { hi = DADD(x, ln2HI); lo = -ln2LO; k = -1; }
    } else {
      k = (int)DADD(DMUL(invln2, x), x < 0 ? -0.5 : 0.5);
      const double t = (double)k;
      hi = DSUB(x, DMUL(t, ln2HI));
      lo = DMUL(t, ln2LO);
    }
    x = DSUB(hi, lo);
  } else if (ax < 3.725290298461914e-09) {  // 2^-28
    return DADD(1.0, x);
  }
  const double t = DMUL(x, x);
  const double c = DSUB(x, DMUL(t, DADD(P1, DMUL(t, DADD(P2, DMUL(t, DADD(P3, DMUL(t, DADD(P4, DMUL(t, P5))))))))));
  if (k == 0) return DSUB(1.0, DSUB(__ddiv_rn(DMUL(x, c), DSUB(c, 2.0)), x));
  const double y = DSUB(1.0, DSUB(DSUB(lo, __ddiv_rn(DMUL(x, c), DSUB(2.0, c))), hi));
  return __longlong_as_double(__double_as_longlong(y) + ((long long)k << 52));  // y * 2^k, |k| small here
}

struct GridDev {
  int32_t n_grids;
  const int64_t* node_off;  // [n_grids + 1]
  const double* node_xy;    // [n_nodes][2]
  const double* node_size;  // [n_grids]
};

// one warp per (observation, grid, position | velocity); out [n_obs][2 * n_grids]: intensity (0 = silent) and/or
// the firing millisecond inside the window (-1 = silent)
__global__ void __launch_bounds__(256)
gridcode_kernel(GridDev g, const double* __restrict__ obs, int n_obs, double reach, int win_size,
                double* __restrict__ inten_out, int32_t* __restrict__ slot_out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long total = (long long)n_obs * g.n_grids * 2;
  if (warp >= total) return;
  const int which = (int)(warp % 2), gi = (int)((warp / 2) % g.n_grids), o = (int)(warp / (2LL * g.n_grids));
  const double* ob = obs + (size_t)o * 4;
  // pos = [obs[1], obs[3] % pi], vel = [obs[2], obs[4]] (:85-86), both scaled from [-15, 15] to the unit square (:88-89)
  const double a = which == 0 ? ob[0] : ob[1];
  const double b = which == 0 ? fmod(ob[2], 3.14159265358979323846) : ob[3];
  const double px = __ddiv_rn(DADD(a, 15.0), 30.0), py = __ddiv_rn(DADD(b, 15.0), 30.0);
  const double d = g.node_size[gi], radius = DMUL(reach, d);
  const int64_t n0 = g.node_off[gi], n1 = g.node_off[gi + 1];
  double found = -1.0;
  for (int64_t base = n0; base < n1; base += 32) {   // warp-uniform
    const int64_t q = base + lane;
    double dist = 1e300;
    if (q < n1) {
      const double dx = DSUB(g.node_xy[2 * q], px), dy = DSUB(g.node_xy[2 * q + 1], py);
      dist = __dsqrt_rn(DADD(DMUL(dx, dx), DMUL(dy, dy)));
    }
    const unsigned m = __ballot_sync(0xffffffffu, dist < radius);
    if (m) { found = __shfl_sync(0xffffffffu, dist, __ffs(m) - 1); break; }
  }
  if (lane != 0) return;
  const size_t idx = (size_t)o * 2 * g.n_grids + (size_t)which * g.n_grids + gi;  // neuron i / n_grids + i (:101,106)
  double x = 0.0;
  if (found >= 0.0) x = DMUL(0.999999, exp_fdlibm(DMUL(-2.0, __ddiv_rn(found, d))));
  if (inten_out) inten_out[idx] = x;
  if (slot_out) slot_out[idx] = found >= 0.0 ? win_size - (int)floor(DMUL((double)win_size, x)) - 1 : -1;  // :131-134
}

}  // namespace snn

struct snn_gridcode {
  snn::GridDev d{};
  int device = 0;
  int64_t n_nodes = 0;
  double* obs = nullptr; double* inten = nullptr; int32_t* slot = nullptr; size_t cap_obs = 0;
};

namespace snn { int fail_with(int code, const std::string& msg); }  // capi.cu: status + thread-local snn_last_error text
#define GC_FAIL(code, msg) return snn::fail_with((code), (msg))
#define GC_OK(expr)                                                                                   \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) return snn::fail_with(SNN_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

extern "C" int32_t snn_gridcode_create(int32_t device, int32_t n_grids, const int64_t* node_off, const double* node_xy,
                                       const double* node_size, snn_gridcode** out) {
  if (!out || !node_off || !node_xy || !node_size || n_grids <= 0) GC_FAIL(SNN_ERR_INVALID, "null argument / no grids");
  for (int g = 0; g < n_grids; ++g)
    if (node_off[g + 1] < node_off[g] || !(node_size[g] > 0.0)) GC_FAIL(SNN_ERR_INVALID, "node_off must ascend and node_size be positive");
  if (node_off[0] != 0) GC_FAIL(SNN_ERR_INVALID, "node_off[0] must be 0");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) { cudaGetLastError(); GC_FAIL(SNN_ERR_NODEVICE, "no CUDA device (there is no CPU fallback)"); }
  GC_OK(cudaSetDevice(device));
  snn_gridcode* h = new snn_gridcode;
  h->device = device; h->n_nodes = node_off[n_grids]; h->d.n_grids = n_grids;
  int64_t* d_off = nullptr; double *d_xy = nullptr, *d_sz = nullptr;
  cudaError_t e = cudaMalloc(&d_off, (size_t)(n_grids + 1) * sizeof(int64_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_xy, (size_t)std::max<int64_t>(1, h->n_nodes) * 2 * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_sz, (size_t)n_grids * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(d_off, node_off, (size_t)(n_grids + 1) * sizeof(int64_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && h->n_nodes) e = cudaMemcpy(d_xy, node_xy, (size_t)h->n_nodes * 2 * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_sz, node_size, (size_t)n_grids * sizeof(double), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(d_off); cudaFree(d_xy); cudaFree(d_sz); delete h;
    GC_FAIL(SNN_ERR_CUDA, std::string("gridcode create: ") + cudaGetErrorString(e));
  }
  h->d.node_off = d_off; h->d.node_xy = d_xy; h->d.node_size = d_sz;
  *out = h;
  return SNN_OK;
}

extern "C" int32_t snn_gridcode_destroy(snn_gridcode* h) {
  if (!h) return SNN_OK;
  cudaSetDevice(h->device);
  cudaFree(const_cast<int64_t*>(h->d.node_off)); cudaFree(const_cast<double*>(h->d.node_xy));
  cudaFree(const_cast<double*>(h->d.node_size)); cudaFree(h->obs); cudaFree(h->inten); cudaFree(h->slot);
  delete h;
  return SNN_OK;
}

extern "C" int32_t snn_gridcode_encode(snn_gridcode* h, const double* obs, int32_t n_obs, double reach, int32_t win_size,
                                       double* intensity_out, int32_t* slot_out) {
  if (!h || !obs || n_obs <= 0 || (!intensity_out && !slot_out)) GC_FAIL(SNN_ERR_INVALID, "null argument");
  if (slot_out && win_size <= 0) GC_FAIL(SNN_ERR_INVALID, "latency code needs win_size > 0");
  if (!(reach > 0.0)) GC_FAIL(SNN_ERR_INVALID, "reach must be positive");
  GC_OK(cudaSetDevice(h->device));
  const size_t per = (size_t)2 * h->d.n_grids;
  if ((size_t)n_obs > h->cap_obs) {   // staging lives with the handle (grow-only)
    cudaFree(h->obs); cudaFree(h->inten); cudaFree(h->slot);
    h->obs = nullptr; h->inten = nullptr; h->slot = nullptr; h->cap_obs = 0;
    const size_t cap = (size_t)n_obs + (size_t)n_obs / 4 + 16;
    GC_OK(cudaMalloc(&h->obs, cap * 4 * sizeof(double)));
    GC_OK(cudaMalloc(&h->inten, cap * per * sizeof(double)));
    GC_OK(cudaMalloc(&h->slot, cap * per * sizeof(int32_t)));
    h->cap_obs = cap;
  }
  GC_OK(cudaMemcpy(h->obs, obs, (size_t)n_obs * 4 * sizeof(double), cudaMemcpyHostToDevice));
  const long long warps = (long long)n_obs * (long long)per;
  const int blocks = (int)((warps * 32 + 255) / 256);
  snn::gridcode_kernel<<<blocks, 256>>>(h->d, h->obs, n_obs, reach, win_size, intensity_out ? h->inten : nullptr,
                                        slot_out ? h->slot : nullptr);
  GC_OK(cudaGetLastError());
  if (intensity_out) GC_OK(cudaMemcpy(intensity_out, h->inten, (size_t)n_obs * per * sizeof(double), cudaMemcpyDeviceToHost));
  if (slot_out) GC_OK(cudaMemcpy(slot_out, h->slot, (size_t)n_obs * per * sizeof(int32_t), cudaMemcpyDeviceToHost));
  return SNN_OK;
}
