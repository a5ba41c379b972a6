// snn_engine.cuh — the per-millisecond hot path: Izhikevich neuron step, spike propagation
// along axonal delays and dopamine-modulated STDP, for sm_100a.
//
// Instantiated twice: snn_f32.cu (production, FMA allowed) and snn_f64.cu (parity build:
// explicit __dmul_rn/__dadd_rn so no multiply-add is ever contracted, matching Julia's
// unfused broadcasts; compiled with -fmad=false as well).
//
// One simulated millisecond = one call of step!(net, train) (Old_net/new_net.jl:146-152):
//
//   reference order (new_net.jl:109-142)          kernels here
//   ------------------------------------------    ------------------------------------------
//   I = noise                          :111       neuron_kernel<DETECT>(t):  [Euler(t-1) first]
//   spiked = v > 30 ; v=-65 ; u+=d     :112-114     threshold/reset, trace set/decay, ballot
//   for n in spiked: I += s[n,:]       :116-118     bitmap + spike list, dopamine decay
//   v,v,u Euler                        :120-122   synapse_kernel(t): arrivals (I += w, LTD) and
//   STDP[spiked]=0.1                   :123         LTP over the incoming index
//   sd[:,n]+=STDP ; sd[n,:]-=1.5STDP   :125-128   [exact mode: gather_kernel(t) pulls I in
//   STDP*=0.95 ; da*=0.995             :130-131     ascending presynaptic order]
//   learn! (if train)                  :135-142   learn_kernel(t): w=clamp(w+(b+da)*sd), sd*=.99
//
// Detection at t only needs the Euler result of t-1, so Euler(t-1) and detect(t) are one pass
// over the neuron arrays (36 B/neuron fp32).  Trace values do not enter the Euler update, so the
// result is identical to the reference order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

#include "snn_common.h"

namespace snn {

template <typename T> struct Real2;
template <> struct Real2<float> { typedef float2 type; };
template <> struct Real2<double> { typedef double2 type; };

// Arithmetic that is never contracted in the fp64 parity build.
template <typename T> struct Ops;
template <> struct Ops<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
  static __device__ __forceinline__ float add(float a, float b) { return a + b; }
  static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
};
template <> struct Ops<double> {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
};

// Philox4x32-10, identical to oracle/snn_oracle.c:ora_philox4x32.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ double philox_noise(uint64_t seed, int64_t t, int64_t j, double amp) {
  uint4 o = philox4x32_10(make_uint4((uint32_t)(j >> 2), 0u, (uint32_t)t, (uint32_t)((uint64_t)t >> 32)),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t x = (j & 3) == 0 ? o.x : ((j & 3) == 1 ? o.y : ((j & 3) == 2 ? o.z : o.w));
  double U = __dmul_rn(__dadd_rn((double)x, 0.5), 1.0 / 4294967296.0);
  return __dmul_rn(amp, __dsub_rn(U, 0.5));
}

// Everything the kernels need, passed by value.
template <typename T>
struct DevNet {
  typedef typename Real2<T>::type T2;
  int64_t N, Nld, Nper, Ne, E;  // Nld = N rounded up to 128: leading dimension of ring slots
  int32_t Dmax, R, B, W;  // W = bitmap words per slot
  // neuron state
  T *v, *u, *ho, *Iacc;
  T* trace;          // [R][N]
  uint32_t* bits;    // [R][W]
  int32_t* spk_list; // [R][N]
  int32_t* spk_cnt;  // [R]
  T* da;             // [B]
  // synapses
  const uint32_t* seg;
  const int32_t* post;
  T2* wsd;           // [E] (w, sd) interleaved: one 8/16-byte access serves propagate + LTD
  const uint32_t *colptr, *col_npl, *in_syn, *in_pre;
  // counters: [0]=spikes [1]=synaptic events (exact mode) ; blk_counters[2*block+{0,1}] = events / ltp
  unsigned long long* counters;
  unsigned long long* blk_counters;
  // spike record
  int32_t* rec_ids; int32_t* rec_off; int64_t rec_cap;
  // parameters
  T vthresh, v_reset, a_exc, a_inh, d_exc, d_inh, trace_set, trace_decay, apost, apre, da_decay,
      learn_base, eta, wmin, wmax, sd_decay, theta, ttheta;
  double noise_amp; uint64_t noise_seed;
  int64_t ho_from;
  int32_t threshold_ge, sd_decay_mode, rate_mode, plastic_ei, noise_mode, exact, homeostasis;
};

constexpr int kMaxSynBlocks = 148 * 16 * 2;

struct DaEvents {  // CSR by window step
  const int32_t* off; const int32_t* inst; const double* amt;
};

__device__ __forceinline__ int slot_of(int64_t t, int R) {
  int64_t m = t % R;
  return (int)(m < 0 ? m + R : m);
}

// ---------------------------------------------------------------------------------------------
// K1 neuron_kernel: [Euler(t-1)] + [detect(t)]   (new_net.jl:111-114,120-123,130-131;
// homeostasis net_res.jl:167-171,204)
//   EULER : I = noise(t-1) + Iacc ; two half-steps of v ; u update ; Iacc cleared
//   DETECT: spike test, reset, homeostasis, trace set/decay into ring slot t, spike bitmap,
//           spike list append, dopamine events of step t-1 + decay (block 0)
// Each thread owns 4 consecutive neurons: every array is read/written with one 16-byte (fp32)
// access, all loads are issued before the first use (the kernel is latency-bound otherwise: ncu
// r1a showed 1 x 128 B in flight per warp = 0.5 TB/s), and one Philox4x32 call yields the four
// noise values.  A warp covers 128 neurons = 4 bitmap words.  Arrays are padded to Nld.
// ---------------------------------------------------------------------------------------------
template <typename T> struct Vec4 { T x[4]; };
__device__ __forceinline__ Vec4<float> ld4(const float* p) {
  float4 r = *reinterpret_cast<const float4*>(p);
  Vec4<float> o; o.x[0] = r.x; o.x[1] = r.y; o.x[2] = r.z; o.x[3] = r.w; return o;
}
__device__ __forceinline__ Vec4<double> ld4(const double* p) {
  double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
  Vec4<double> o; o.x[0] = a.x; o.x[1] = a.y; o.x[2] = b.x; o.x[3] = b.y; return o;
}
__device__ __forceinline__ void st4(float* p, const Vec4<float>& v) {
  *reinterpret_cast<float4*>(p) = make_float4(v.x[0], v.x[1], v.x[2], v.x[3]);
}
__device__ __forceinline__ void st4(double* p, const Vec4<double>& v) {
  *reinterpret_cast<double2*>(p) = make_double2(v.x[0], v.x[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v.x[2], v.x[3]);
}

template <typename T, bool EULER, bool DETECT>
__global__ void __launch_bounds__(256)
neuron_kernel(DevNet<T> p, int64_t t, int32_t k_rel, int slot, int prev, const T* __restrict__ noise_prev,
              DaEvents ev) {
  typedef Ops<T> O;
  const int lane = threadIdx.x & 31;
  T* __restrict__ tr = p.trace + (size_t)slot * p.Nld;
  const T* __restrict__ trp = p.trace + (size_t)prev * p.Nld;
  uint32_t* __restrict__ bits = p.bits + (size_t)slot * p.W;
  int32_t* __restrict__ list = p.spk_list + (size_t)slot * p.Nld;
  const int32_t rec_base = (DETECT && p.rec_ids) ? p.rec_off[k_rel] : 0;
  const uint32_t N = (uint32_t)p.N, Nper = (uint32_t)p.Nper, Ne = (uint32_t)p.Ne;

  // dopamine: da += events(step k_rel-1) ; da *= decay (new_net.jl:131, newnet_DASTDP.jl:54-56)
  if (blockIdx.x == 0) {
    if (ev.off && k_rel > 0 && threadIdx.x == 0)
      for (int q = ev.off[k_rel - 1]; q < ev.off[k_rel]; ++q)
        p.da[ev.inst[q]] = O::add(p.da[ev.inst[q]], (T)ev.amt[q]);
    __syncthreads();
    if (DETECT)
      for (int b = threadIdx.x; b < p.B; b += blockDim.x) p.da[b] = O::mul(p.da[b], p.da_decay);
  }

  const uint32_t nquad = (uint32_t)(p.Nld >> 2);
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < nquad; q += gridDim.x * blockDim.x) {
    const uint32_t j0 = q << 2;
    // ---- all loads first
    Vec4<T> v = ld4(p.v + j0), u = ld4(p.u + j0), I, tp, ho;
    if (EULER) I = ld4(p.Iacc + j0);
    if (DETECT) {
      tp = ld4(trp + j0);
      if (p.homeostasis) ho = ld4(p.ho + j0);
    }
    uint32_t l0 = p.B == 1 ? j0 : j0 % Nper;
    bool exc[4], act[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t l = l0 + c;
      if (l >= Nper) l -= Nper;
      exc[c] = l < Ne;
      act[c] = j0 + c < N;
    }
    if (EULER) {
      if (!p.exact) {
        if (p.noise_mode == SNN_NOISE_REPLAY) {
#pragma unroll
          for (int c = 0; c < 4; ++c) if (act[c]) I.x[c] = O::add(noise_prev[j0 + c], I.x[c]);
        } else if (p.noise_mode == SNN_NOISE_PHILOX) {
          const int64_t tn = t - 1;
          uint4 o = philox4x32_10(make_uint4(q, 0u, (uint32_t)tn, (uint32_t)((uint64_t)tn >> 32)),
                                  make_uint2((uint32_t)p.noise_seed, (uint32_t)(p.noise_seed >> 32)));
          const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            double U = __dmul_rn(__dadd_rn((double)w[c], 0.5), 1.0 / 4294967296.0);
            I.x[c] = O::add((T)__dmul_rn(p.noise_amp, __dsub_rn(U, 0.5)), I.x[c]);
          }
        }
        Vec4<T> z; z.x[0] = z.x[1] = z.x[2] = z.x[3] = (T)0;
        st4(p.Iacc + j0, z);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (!act[c]) continue;
        T vv = v.x[c];
        // v = v + 0.5*((0.04*v+5)*v + 140 - u + I), twice (new_net.jl:120-121)
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          T x = O::add(O::mul((T)0.04, vv), (T)5);
          x = O::add(O::mul(x, vv), (T)140);
          x = O::add(O::sub(x, u.x[c]), I.x[c]);
          vv = O::add(vv, O::mul((T)0.5, x));
        }
        v.x[c] = vv;
        // u = u + a*(0.2*v - u) (new_net.jl:122)
        u.x[c] = O::add(u.x[c], O::mul(exc[c] ? p.a_exc : p.a_inh, O::sub(O::mul((T)0.2, vv), u.x[c])));
      }
    }
    if (DETECT) {
      unsigned nib = 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        bool s = false;
        if (act[c]) {
          T x = p.homeostasis ? O::sub(v.x[c], ho.x[c]) : v.x[c];
          s = p.threshold_ge ? (x >= p.vthresh) : (x > p.vthresh);
          if (s) {
            v.x[c] = p.v_reset;
            u.x[c] = O::add(u.x[c], exc[c] ? p.d_exc : p.d_inh);
          }
        }
        if (p.homeostasis) {
          uint32_t l = l0 + c;
          if (l >= Nper) l -= Nper;
          T h = ho.x[c];
          if (s && exc[c] && (int64_t)l >= p.ho_from) h = O::add(h, p.theta);
          ho.x[c] = O::mul(h, p.ttheta);
        }
        tp.x[c] = s ? p.trace_set : (t > 0 ? O::mul(tp.x[c], p.trace_decay) : (T)0);
        nib |= (s ? 1u : 0u) << c;
      }
      st4(tr + j0, tp);
      if (p.homeostasis) st4(p.ho + j0, ho);
      // 8 lanes x 4 neurons = one 32-bit bitmap word
      unsigned word = nib << ((lane & 7) << 2);
      word |= __shfl_xor_sync(0xffffffffu, word, 1);
      word |= __shfl_xor_sync(0xffffffffu, word, 2);
      word |= __shfl_xor_sync(0xffffffffu, word, 4);
      if ((lane & 7) == 0) bits[j0 >> 5] = word;
      if (__any_sync(0xffffffffu, nib != 0)) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const bool s = (nib >> c) & 1u;
          const unsigned m = __ballot_sync(0xffffffffu, s);
          if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&p.spk_cnt[slot], __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (s) {
              const int pos = base + __popc(m & ((1u << lane) - 1u));
              list[pos] = (int32_t)(j0 + c);
              if (p.rec_ids && (int64_t)rec_base + pos < p.rec_cap) p.rec_ids[rec_base + pos] = (int32_t)(j0 + c);
            }
          }
        }
      }
    }
    st4(p.v + j0, v);
    st4(p.u + j0, u);
  }
}

// v[idx] = 40 (input! new_net.jl:104-106)
template <typename T>
__global__ void force_kernel(DevNet<T> p, const int32_t* __restrict__ idx, int n) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) p.v[idx[q]] = (T)40.0;
}

// ---------------------------------------------------------------------------------------------
// exact mode: I[j] = noise[j] + sum over spiking sources in ascending presynaptic id — the
// accumulation order of `for neuron in spiked: I .+= s[neuron,:]` (new_net.jl:116-118).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gather_kernel(DevNet<T> p, int64_t t, const T* __restrict__ noise) {
  typedef Ops<T> O;
  unsigned long long ev = 0;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < p.N;
       j += (int64_t)gridDim.x * blockDim.x) {
    T acc = (T)0;
    if (p.noise_mode == SNN_NOISE_REPLAY) acc = noise[j];
    else if (p.noise_mode == SNN_NOISE_PHILOX) acc = (T)philox_noise(p.noise_seed, t, j, p.noise_amp);
    const uint32_t c0 = p.colptr[j], c1 = p.colptr[j + 1];
    for (uint32_t k = c0; k < c1; ++k) {
      const uint32_t pk = p.in_pre[k];
      const uint32_t i = pk & kPreMask;
      const int a = (int)(pk >> kPreBits);
      const uint32_t word = p.bits[(size_t)slot_of(t - a, p.R) * p.W + (i >> 5)];
      if ((word >> (i & 31)) & 1u) { acc = O::add(acc, p.wsd[p.in_syn[k]].x); ++ev; }
    }
    p.Iacc[j] = acc;
  }
  if (ev) atomicAdd(&p.counters[1], ev);
}

// ---------------------------------------------------------------------------------------------
// K2+K3 synapse_kernel(t): persistent grid, sub-warp groups of G lanes per work item.
//   phase 1, arrivals: for every age a=0..Dmax-1 and every neuron i that spiked at t-a, walk the
//     delay-(a+1) segment of row i:  I[post] += w (PUSH), and for plastic rows LTD with the post
//     trace of this call (new_net.jl:127 `sd[n,:] .-= 1.5 .* STDP`).  If the target also spiked
//     at t the same thread applies its LTP too, in the reference's loop order
//     (post<=pre: + then -, else - then +; new_net.jl:125-128).
//   phase 2, LTP: for every neuron j that spiked at t walk the plastic prefix of its incoming
//     list (new_net.jl:126 `sd[:,n] .+= STDP`), skipping synapses owned by phase 1.
// Every synapse is written by exactly one thread per call: no atomics on sd, bit-reproducible.
// ---------------------------------------------------------------------------------------------
template <typename T, bool PUSH>
__global__ void __launch_bounds__(256)
synapse_kernel(DevNet<T> p, int64_t t, int32_t k_rel, int slot, int g_arr_log2, int g_ltp_log2) {
  typedef Ops<T> O;
  __shared__ int cum[kMaxDelay + 2];
  __shared__ int slots[kMaxDelay + 1];
  __shared__ unsigned long long blk_ev, blk_ltp;
  // prologue: the Dmax list lengths are loaded in parallel (one round trip), scanned in smem
  if (threadIdx.x < (unsigned)p.Dmax) {
    int sl = slot - (int)threadIdx.x;
    if (sl < 0) sl += p.R;
    slots[threadIdx.x] = sl;
    cum[threadIdx.x + 1] = p.spk_cnt[sl];
  }
  if (threadIdx.x == 0) { blk_ev = 0ull; blk_ltp = 0ull; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int a = 0; a < p.Dmax; ++a) { int c = cum[a + 1]; cum[a] = acc; acc += c; }
    cum[p.Dmax] = acc;
    if (blockIdx.x == 0) {
      const int c0 = cum[p.Dmax > 1 ? 1 : p.Dmax];
      p.counters[0] += (unsigned long long)c0;  // single writer
      if (p.rec_ids) {
        int64_t nxt = (int64_t)p.rec_off[k_rel] + c0;
        p.rec_off[k_rel + 1] = (int32_t)(nxt < p.rec_cap ? nxt : p.rec_cap);
      }
      int nx = slot + 1; if (nx >= p.R) nx -= p.R;
      p.spk_cnt[nx] = 0;  // next call's list; not read by this launch (R = Dmax+1)
    }
  }
  __syncthreads();
  const int total_arr = cum[p.Dmax];
  const int n_now = p.Dmax > 1 ? cum[1] : total_arr;  // = spk_cnt[slot]
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  const uint32_t Nper = (uint32_t)p.Nper, Ne = (uint32_t)p.Ne;
  const size_t Nld = (size_t)p.Nld;
  const T* __restrict__ tr_now = p.trace + (size_t)slot * Nld;
  const uint32_t* __restrict__ bits_now = p.bits + (size_t)slot * p.W;
  unsigned long long n_ev = 0, n_ltp = 0;

  {  // LTP over the incoming index of this call's spikers.  Items are handed out from the LAST
     // group downwards so they run concurrently with the arrivals of the low groups.
    const uint32_t G = 1u << g_ltp_log2;
    const uint32_t gl = gtid & (G - 1);
    const uint32_t ngroups = nthreads >> g_ltp_log2;
    const int32_t* __restrict__ list = p.spk_list + (size_t)slot * Nld;
    for (uint32_t item = ngroups - 1 - (gtid >> g_ltp_log2); item < (uint32_t)n_now; item += ngroups) {
      const int32_t j = list[item];
      const uint32_t c0 = p.colptr[j], c1 = c0 + p.col_npl[j];
      if (gl == 0) n_ltp += (c1 - c0);
      for (uint32_t kb = c0 + gl; kb < c1; kb += 3 * G) {
        uint32_t pk[3], e[3]; bool ok[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const uint32_t k = kb + q * G;
          ok[q] = k < c1;
          pk[q] = ok[q] ? p.in_pre[k] : 0u;
          e[q] = ok[q] ? p.in_syn[k] : 0u;
        }
        size_t toff[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const uint32_t i = pk[q] & kPreMask;
          const int sl = slots[pk[q] >> kPreBits];
          toff[q] = (size_t)sl * Nld + i;
          if (ok[q]) {
            const uint32_t word = p.bits[(size_t)sl * p.W + (i >> 5)];
            ok[q] = !((word >> (i & 31)) & 1u);  // else the spike arrives now: arrivals own the synapse
          }
        }
        T trv[3], sdv[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          trv[q] = ok[q] ? p.trace[toff[q]] : (T)0;
          sdv[q] = ok[q] ? p.wsd[e[q]].y : (T)0;
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
          if (ok[q]) p.wsd[e[q]].y = O::add(sdv[q], O::mul(p.apost, trv[q]));
      }
    }
  }
  {  // arrivals
    const uint32_t G = 1u << g_arr_log2;
    const uint32_t gl = gtid & (G - 1);
    const uint32_t ngroups = nthreads >> g_arr_log2;
    for (uint32_t item = gtid >> g_arr_log2; item < (uint32_t)total_arr; item += ngroups) {
      int a = 0;
      while (a + 1 < p.Dmax && cum[a + 1] <= (int)item) ++a;
      const int sl = slots[a];
      const uint32_t i = (uint32_t)p.spk_list[(size_t)sl * Nld + (item - cum[a])];
      const uint32_t s0 = p.seg[(size_t)i * (p.Dmax + 1) + a];
      const uint32_t s1 = p.seg[(size_t)i * (p.Dmax + 1) + a + 1];
      const uint32_t ibase = p.B == 1 ? 0u : (i / Nper) * Nper;
      const bool row_plastic = (i - ibase) < Ne;
      const T pre_up = row_plastic ? O::mul(p.apost, p.trace[(size_t)sl * Nld + i]) : (T)0;
      if (gl == 0) n_ev += (s1 - s0);
      for (uint32_t e = s0 + gl; e < s1; e += G) {
        const uint32_t j = (uint32_t)p.post[e];
        typename DevNet<T>::T2 ws = p.wsd[e];
        if (PUSH) atomicAdd(&p.Iacc[j], ws.x);
        if (row_plastic && (p.plastic_ei || (j - ibase) < Ne)) {
          const T dn = O::mul(p.apre, tr_now[j]);
          const bool jsp = (bits_now[j >> 5] >> (j & 31)) & 1u;
          T sd = ws.y;
          if (jsp) sd = (j <= i) ? O::sub(O::add(sd, pre_up), dn) : O::add(O::sub(sd, dn), pre_up);
          else sd = O::sub(sd, dn);
          p.wsd[e].y = sd;
        }
      }
    }
  }
  // block-local reduction, then one plain RMW on this block's own counter slot (no contended atomics)
  for (int o = 16; o > 0; o >>= 1) {
    n_ev += __shfl_down_sync(0xffffffffu, n_ev, o);
    n_ltp += __shfl_down_sync(0xffffffffu, n_ltp, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (n_ev) atomicAdd(&blk_ev, n_ev);
    if (n_ltp) atomicAdd(&blk_ltp, n_ltp);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (PUSH) p.blk_counters[2 * blockIdx.x] += blk_ev;
    p.blk_counters[2 * blockIdx.x + 1] += blk_ltp;
  }
}

// ---------------------------------------------------------------------------------------------
// K4 learn_kernel: one streaming pass over the plastic synapses of every instance
//   w = clamp(w + rate*sd, wmin, wmax) ; sd *= sd_decay      (new_net.jl:135-142)
// rate = (learn_base+da) | da | eta*da.  16 B/synapse fp32 (read w,sd; write w,sd).
// blockIdx.y = instance; exc rows of an instance are contiguous in the CSR.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void learn_one(typename Real2<T>::type& ws, T g, T wmin, T wmax, T sdk,
                                          bool do_learn, bool do_decay) {
  typedef Ops<T> O;
  if (do_learn) {
    T x = O::add(ws.x, O::mul(g, ws.y));
    ws.x = x > wmax ? wmax : (x < wmin ? wmin : x);
  }
  if (do_decay) ws.y = O::mul(ws.y, sdk);
}

template <typename T>
__global__ void __launch_bounds__(256)
learn_kernel(DevNet<T> p, int do_learn, int do_decay) {
  typedef Ops<T> O;
  typedef typename DevNet<T>::T2 T2;
  const int b = blockIdx.y;
  const int64_t r0 = (int64_t)b * p.Nper;
  const uint32_t lo = p.seg[(size_t)r0 * (p.Dmax + 1)];
  const uint32_t hi = p.seg[(size_t)(r0 + p.Ne - 1) * (p.Dmax + 1) + p.Dmax];
  const T da = p.da[b];
  const T g = p.rate_mode == SNN_RATE_BASE_PLUS_DA ? O::add(p.learn_base, da)
              : (p.rate_mode == SNN_RATE_DA ? da : O::mul(p.eta, da));
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  if (p.plastic_ei) {
    if constexpr (sizeof(T) == 4) {
      // 16-byte accesses: two synapses per load
      const uint32_t lo2 = (lo + 1u) & ~1u, hi2 = hi & ~1u;
      if (tid == 0 && lo < lo2 && lo < hi) {
        T2 ws = p.wsd[lo]; learn_one<T>(ws, g, p.wmin, p.wmax, p.sd_decay, do_learn, do_decay); p.wsd[lo] = ws;
      }
      if (tid == 1 && hi2 < hi && hi2 >= lo2) {
        T2 ws = p.wsd[hi2]; learn_one<T>(ws, g, p.wmin, p.wmax, p.sd_decay, do_learn, do_decay); p.wsd[hi2] = ws;
      }
      if (hi2 > lo2) {
        float4* __restrict__ q = reinterpret_cast<float4*>(p.wsd + lo2);
        const int64_t n4 = (hi2 - lo2) >> 1;
        int64_t x = tid;
        for (; x + 3 * nth < n4; x += 4 * nth) {
          float4 r[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) r[c] = q[x + c * nth];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float2 s0 = make_float2(r[c].x, r[c].y), s1 = make_float2(r[c].z, r[c].w);
            learn_one<float>(s0, (float)g, (float)p.wmin, (float)p.wmax, (float)p.sd_decay, do_learn, do_decay);
            learn_one<float>(s1, (float)g, (float)p.wmin, (float)p.wmax, (float)p.sd_decay, do_learn, do_decay);
            q[x + c * nth] = make_float4(s0.x, s0.y, s1.x, s1.y);
          }
        }
        for (; x < n4; x += nth) {
          float4 r = q[x];
          float2 s0 = make_float2(r.x, r.y), s1 = make_float2(r.z, r.w);
          learn_one<float>(s0, (float)g, (float)p.wmin, (float)p.wmax, (float)p.sd_decay, do_learn, do_decay);
          learn_one<float>(s1, (float)g, (float)p.wmin, (float)p.wmax, (float)p.sd_decay, do_learn, do_decay);
          q[x] = make_float4(s0.x, s0.y, s1.x, s1.y);
        }
      }
    } else {
      int64_t e = (int64_t)lo + tid;
      for (; e + 3 * nth < hi; e += 4 * nth) {
        T2 r[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) r[c] = p.wsd[e + c * nth];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          learn_one<T>(r[c], g, p.wmin, p.wmax, p.sd_decay, do_learn, do_decay);
          p.wsd[e + c * nth] = r[c];
        }
      }
      for (; e < hi; e += nth) {
        T2 ws = p.wsd[e];
        learn_one<T>(ws, g, p.wmin, p.wmax, p.sd_decay, do_learn, do_decay);
        p.wsd[e] = ws;
      }
    }
  } else {
    // only exc->exc synapses are plastic (net_res.jl:213-216): needs the post id
    for (int64_t e = (int64_t)lo + tid; e < hi; e += nth) {
      if (((int64_t)p.post[e] - r0) >= p.Ne) continue;
      T2 ws = p.wsd[e];
      learn_one<T>(ws, g, p.wmin, p.wmax, p.sd_decay, do_learn, do_decay);
      p.wsd[e] = ws;
    }
  }
}

// ----------------------------------------------------------------------------- small helpers
template <typename T>
__global__ void init_state_kernel(DevNet<T> p) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < p.Nld; j += (int64_t)gridDim.x * blockDim.x) {
    p.v[j] = p.v_reset;
    p.u[j] = Ops<T>::mul((T)0.2, p.v_reset);
  }
}
template <typename T>
__global__ void fill_kernel(T* a, int64_t n, T x) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) a[j] = x;
}
template <typename T>
__global__ void from_double_kernel(const double* __restrict__ in, T* __restrict__ out, int64_t n) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) out[j] = (T)in[j];
}
template <typename T>
__global__ void to_double_kernel(const T* __restrict__ in, double* __restrict__ out, int64_t n, double scale) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
    out[j] = scale == 1.0 ? (double)in[j] : (double)(Ops<T>::mul(in[j], (T)scale));
}
// caller order <-> internal order through perm; comp 0 = w, 1 = sd
template <typename T>
__global__ void wsd_from_caller(const double* __restrict__ in, const uint32_t* __restrict__ perm,
                                typename Real2<T>::type* wsd, int64_t E, int comp, int zero_other) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    T x = (T)in[perm[e]];
    if (comp == 0) { wsd[e].x = x; if (zero_other) wsd[e].y = (T)0; } else wsd[e].y = x;
  }
}
template <typename T>
__global__ void wsd_to_caller(const typename Real2<T>::type* __restrict__ wsd, const uint32_t* __restrict__ perm,
                              double* __restrict__ out, int64_t E, int comp) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x)
    out[perm[e]] = (double)(comp == 0 ? wsd[e].x : wsd[e].y);
}

// ---------------------------------------------------------------------------------------------
template <typename T>
class Engine : public IEngine {
 public:
  explicit Engine(const snn_config& c) : cfg_(c) {}
  ~Engine() override { release(); }

  int init(std::string& err) override {
    SNN_CUDA_OK(cudaSetDevice(cfg_.device));
    SNN_CUDA_OK(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking));
    stream_ = own_stream_;
    SNN_CUDA_OK(cudaEventCreate(&ev_begin_));
    SNN_CUDA_OK(cudaEventCreate(&ev_end_));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg_.device);
    n_sm_ = sms;
    DevNet<T>& p = p_;
    memset(&p, 0, sizeof(p));
    p.N = cfg_.n_neurons; p.Nld = (p.N + 127) / 128 * 128; p.Nper = cfg_.n_per_instance; p.Ne = cfg_.n_exc;
    p.Dmax = cfg_.max_delay; p.R = p.Dmax + 1; p.B = (int32_t)(p.N / p.Nper); p.W = (int32_t)(p.Nld / 32);
    const size_t N = (size_t)p.Nld;  // padded: vector accesses never leave the allocation
    SNN_CUDA_OK(cudaMalloc(&p.v, N * sizeof(T)));
    SNN_CUDA_OK(cudaMalloc(&p.u, N * sizeof(T)));
    SNN_CUDA_OK(cudaMalloc(&p.ho, N * sizeof(T)));
    SNN_CUDA_OK(cudaMalloc(&p.Iacc, N * sizeof(T)));
    SNN_CUDA_OK(cudaMalloc(&p.trace, (size_t)p.R * N * sizeof(T)));
    SNN_CUDA_OK(cudaMalloc(&p.bits, (size_t)p.R * p.W * sizeof(uint32_t)));
    SNN_CUDA_OK(cudaMalloc(&p.spk_list, (size_t)p.R * N * sizeof(int32_t)));
    SNN_CUDA_OK(cudaMalloc(&p.spk_cnt, (size_t)p.R * sizeof(int32_t)));
    SNN_CUDA_OK(cudaMalloc(&p.da, (size_t)p.B * sizeof(T)));
    SNN_CUDA_OK(cudaMalloc(&p.counters, 4 * sizeof(unsigned long long)));
    SNN_CUDA_OK(cudaMalloc(&p.blk_counters, 2 * kMaxSynBlocks * sizeof(unsigned long long)));
    SNN_CUDA_OK(cudaMemsetAsync(p.blk_counters, 0, 2 * kMaxSynBlocks * sizeof(unsigned long long), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.ho, 0, N * sizeof(T), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.Iacc, 0, N * sizeof(T), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.trace, 0, (size_t)p.R * N * sizeof(T), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.bits, 0, (size_t)p.R * p.W * sizeof(uint32_t), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.spk_cnt, 0, (size_t)p.R * sizeof(int32_t), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.da, 0, (size_t)p.B * sizeof(T), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.counters, 0, 4 * sizeof(unsigned long long), stream_));
    p.vthresh = (T)cfg_.vthresh; p.v_reset = (T)cfg_.v_reset; p.a_exc = (T)cfg_.a_exc; p.a_inh = (T)cfg_.a_inh;
    p.d_exc = (T)cfg_.d_exc; p.d_inh = (T)cfg_.d_inh; p.trace_set = (T)cfg_.trace_set;
    p.trace_decay = (T)cfg_.trace_decay; p.apost = (T)cfg_.apost; p.apre = (T)cfg_.apre;
    p.da_decay = (T)cfg_.da_decay; p.learn_base = (T)cfg_.learn_base; p.eta = (T)cfg_.eta;
    p.wmin = (T)cfg_.wmin; p.wmax = (T)cfg_.wmax; p.sd_decay = (T)cfg_.sd_decay; p.theta = (T)cfg_.theta;
    p.ttheta = (T)cfg_.ttheta; p.noise_amp = cfg_.noise_amp; p.noise_seed = cfg_.noise_seed;
    p.ho_from = cfg_.ho_from; p.threshold_ge = cfg_.threshold_ge; p.sd_decay_mode = cfg_.sd_decay_mode;
    p.rate_mode = cfg_.rate_mode; p.plastic_ei = cfg_.plastic_ei; p.noise_mode = cfg_.noise_mode;
    p.exact = cfg_.mode == SNN_MODE_EXACT;
    p.homeostasis = !(cfg_.theta == 0.0 && cfg_.ttheta == 1.0);
    init_state_kernel<T><<<grid_for(p.Nld), 256, 0, stream_>>>(p);
    SNN_CUDA_OK(cudaGetLastError());
    SNN_CUDA_OK(cudaStreamSynchronize(stream_));
    return SNN_OK;
  }

  int set_topology(Topo&& topo, const double* h_w, std::string& err) override {
    SNN_CUDA_OK(cudaSetDevice(cfg_.device));
    topo_.release();
    topo_ = topo;
    DevNet<T>& p = p_;
    p.E = topo_.E;
    p.seg = topo_.seg; p.post = topo_.post; p.colptr = topo_.colptr; p.col_npl = topo_.col_npl;
    p.in_syn = topo_.in_syn; p.in_pre = topo_.in_pre;
    cudaFree(p.wsd); p.wsd = nullptr;
    const int64_t Ea = p.E > 0 ? p.E : 1;
    SNN_CUDA_OK(cudaMalloc(&p.wsd, Ea * sizeof(typename DevNet<T>::T2)));
    int rc = upload_wsd(h_w, 0, 1, err);
    if (rc) return rc;
    // group sizes from the mean work-item length
    const double mean_row = p.N ? (double)p.E / (double)p.N : 0.0;
    int n_delays = p.Dmax;
    double mean_seg = mean_row / (n_delays > 0 ? n_delays : 1);
    g_arr_log2_ = pick_group(mean_seg);
    g_ltp_log2_ = pick_group(mean_row);
    have_topo_ = true;
    return SNN_OK;
  }

  int step(const StepArgs& a, std::string& err) override;

  int get_array(int field, double* host, int64_t count, std::string& err) override {
    SNN_CUDA_OK(cudaSetDevice(cfg_.device));
    DevNet<T>& p = p_;
    const T* src = nullptr; int64_t n = p.N; double scale = 1.0;
    switch (field) {
      case SNN_FIELD_V: src = p.v; break;
      case SNN_FIELD_U: src = p.u; break;
      case SNN_FIELD_HO: src = p.ho; break;
      case SNN_FIELD_TRACE:
        if (t_ == 0) { if (count != p.N) { err = "count mismatch"; return SNN_ERR_INVALID; } std::fill(host, host + count, 0.0); return SNN_OK; }
        src = p.trace + (size_t)slot_host(t_ - 1) * p.Nld; scale = cfg_.trace_decay; break;
      case SNN_FIELD_DA: src = p.da; n = p.B; break;
      case SNN_FIELD_I:
        if (!p.exact) { err = "FIELD_I is only kept in SNN_MODE_EXACT"; return SNN_ERR_INVALID; }
        src = p.Iacc; break;
      case SNN_FIELD_W: case SNN_FIELD_SD: {
        if (!have_topo_) { err = "synapses not set"; return SNN_ERR_STATE; }
        if (count != p.E) { err = "count mismatch"; return SNN_ERR_INVALID; }
        if (p.E == 0) return SNN_OK;
        double* d = nullptr;
        SNN_CUDA_OK(cudaMalloc(&d, p.E * sizeof(double)));
        wsd_to_caller<T><<<grid_for(p.E), 256, 0, stream_>>>(p.wsd, topo_.perm, d, p.E, field == SNN_FIELD_W ? 0 : 1);
        cudaError_t e = cudaMemcpyAsync(host, d, p.E * sizeof(double), cudaMemcpyDeviceToHost, stream_);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
        cudaFree(d);
        if (e != cudaSuccess) { err = cudaGetErrorString(e); return SNN_ERR_CUDA; }
        return SNN_OK;
      }
      default: err = "unknown field"; return SNN_ERR_INVALID;
    }
    if (count != n) { err = "count mismatch"; return SNN_ERR_INVALID; }
    double* d = nullptr;
    SNN_CUDA_OK(cudaMalloc(&d, n * sizeof(double)));
    to_double_kernel<T><<<grid_for(n), 256, 0, stream_>>>(src, d, n, scale);
    cudaError_t e = cudaMemcpyAsync(host, d, n * sizeof(double), cudaMemcpyDeviceToHost, stream_);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    cudaFree(d);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return SNN_ERR_CUDA; }
    return SNN_OK;
  }

  int set_array(int field, const double* host, int64_t count, std::string& err) override {
    SNN_CUDA_OK(cudaSetDevice(cfg_.device));
    DevNet<T>& p = p_;
    T* dst = nullptr; int64_t n = p.N;
    switch (field) {
      case SNN_FIELD_V: dst = p.v; break;
      case SNN_FIELD_U: dst = p.u; break;
      case SNN_FIELD_HO: dst = p.ho; break;
      case SNN_FIELD_DA: dst = p.da; n = p.B; break;
      case SNN_FIELD_W: case SNN_FIELD_SD:
        if (!have_topo_) { err = "synapses not set"; return SNN_ERR_STATE; }
        if (count != p.E) { err = "count mismatch"; return SNN_ERR_INVALID; }
        return upload_wsd(host, field == SNN_FIELD_W ? 0 : 1, 0, err);
      default: err = "field is not settable"; return SNN_ERR_INVALID;
    }
    if (count != n) { err = "count mismatch"; return SNN_ERR_INVALID; }
    double* d = nullptr;
    SNN_CUDA_OK(cudaMalloc(&d, n * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d, host, n * sizeof(double), cudaMemcpyHostToDevice, stream_);
    from_double_kernel<T><<<grid_for(n), 256, 0, stream_>>>(d, dst, n);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    cudaFree(d);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return SNN_ERR_CUDA; }
    return SNN_OK;
  }

  int add_dopamine(int instance, double amount, std::string& err) override {
    if (instance < 0 || instance >= p_.B) { err = "instance out of range"; return SNN_ERR_INVALID; }
    std::vector<double> da(p_.B);
    int rc = get_array(SNN_FIELD_DA, da.data(), p_.B, err);
    if (rc) return rc;
    // same rounding as the device type: da (T) + amount (T)
    da[instance] = (double)((T)da[instance] + (T)amount);
    return set_array(SNN_FIELD_DA, da.data(), p_.B, err);
  }

  int reset_v(std::string& err) override {
    SNN_CUDA_OK(cudaSetDevice(cfg_.device));
    fill_kernel<T><<<grid_for(p_.N), 256, 0, stream_>>>(p_.v, p_.N, p_.v_reset);
    SNN_CUDA_OK(cudaGetLastError());
    SNN_CUDA_OK(cudaStreamSynchronize(stream_));
    return SNN_OK;
  }

  int64_t n_synapses() const override { return p_.E; }
  int64_t step_index() const override { return t_; }
  void set_profiling(bool on) override { profiling_ = on; }
  void set_stream(cudaStream_t s) override { stream_ = s ? s : own_stream_; }
  cudaStream_t stream() const override { return stream_; }

 private:
  int grid_for(int64_t n) const {
    int64_t b = (n + 255) / 256;
    int64_t cap = (int64_t)n_sm_ * 8;
    return (int)std::max<int64_t>(1, std::min(b, cap));
  }
  int slot_host(int64_t t) const { int64_t m = t % p_.R; return (int)(m < 0 ? m + p_.R : m); }
  static int pick_group(double mean_len) {
    int g = 0;
    while (g < 5 && (double)(1 << g) < mean_len) ++g;
    return g;
  }
  int upload_wsd(const double* h, int comp, int zero_other, std::string& err) {
    DevNet<T>& p = p_;
    if (p.E == 0) return SNN_OK;
    double* d = nullptr;
    SNN_CUDA_OK(cudaMalloc(&d, p.E * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d, h, p.E * sizeof(double), cudaMemcpyHostToDevice, stream_);
    wsd_from_caller<T><<<grid_for(p.E), 256, 0, stream_>>>(d, topo_.perm, p.wsd, p.E, comp, zero_other);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    cudaFree(d);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return SNN_ERR_CUDA; }
    return SNN_OK;
  }
  void release() {
    DevNet<T>& p = p_;
    cudaFree(p.v); cudaFree(p.u); cudaFree(p.ho); cudaFree(p.Iacc); cudaFree(p.trace); cudaFree(p.bits);
    cudaFree(p.spk_list); cudaFree(p.spk_cnt); cudaFree(p.da); cudaFree(p.counters); cudaFree(p.blk_counters); cudaFree(p.wsd);
    topo_.release();
    if (ev_begin_) cudaEventDestroy(ev_begin_);
    if (ev_end_) cudaEventDestroy(ev_end_);
    for (auto& e : prof_events_) cudaEventDestroy(e);
    prof_events_.clear();
    if (own_stream_) cudaStreamDestroy(own_stream_);
    memset(&p, 0, sizeof(p));
    own_stream_ = nullptr; ev_begin_ = ev_end_ = nullptr;
  }

  // profiling: event pairs around launches, tagged by kernel class
  void prof_begin(int cls) {
    if (!profiling_) return;
    if (prof_used_ + 2 > prof_events_.size()) {
      size_t old = prof_events_.size();
      prof_events_.resize(old + 256);
      for (size_t q = old; q < prof_events_.size(); ++q) cudaEventCreate(&prof_events_[q]);
    }
    cudaEventRecord(prof_events_[prof_used_], stream_);
    prof_cls_.push_back(cls);
  }
  void prof_end() {
    if (!profiling_) return;
    cudaEventRecord(prof_events_[prof_used_ + 1], stream_);
    prof_used_ += 2;
  }

  snn_config cfg_;
  DevNet<T> p_;
  Topo topo_;
  bool have_topo_ = false, profiling_ = false;
  int64_t t_ = 0;
  int n_sm_ = 148, g_arr_log2_ = 5, g_ltp_log2_ = 5, gS_ = 0;
  cudaStream_t own_stream_ = nullptr, stream_ = nullptr;
  cudaEvent_t ev_begin_ = nullptr, ev_end_ = nullptr;
  std::vector<cudaEvent_t> prof_events_;
  std::vector<int> prof_cls_;
  size_t prof_used_ = 0;
};

template <typename T>
int Engine<T>::step(const StepArgs& a, std::string& err) {
  if (!have_topo_) { err = "snn_step before synapses were set"; return SNN_ERR_STATE; }
  if (a.n_steps <= 0) { err = "n_steps must be > 0"; return SNN_ERR_INVALID; }
  if (cfg_.noise_mode == SNN_NOISE_REPLAY && !a.noise) { err = "noise_mode REPLAY needs a noise array"; return SNN_ERR_INVALID; }
  if (a.spikes_out && (!a.spike_offsets || a.spikes_cap <= 0 || a.spikes_cap > 0x7fffffff)) {
    err = "spikes_out needs spike_offsets and 0 < spikes_cap < 2^31"; return SNN_ERR_INVALID;
  }
  SNN_CUDA_OK(cudaSetDevice(cfg_.device));
  DevNet<T>& p = p_;
  const int n = a.n_steps;
  const int64_t N = p.N;

  // ---- stage the window's inputs on the device
  T* d_noise = nullptr; double* d_noise64 = nullptr;
  int32_t *d_ev_off = nullptr, *d_ev_inst = nullptr, *d_force = nullptr, *d_rec = nullptr, *d_rec_off = nullptr;
  double* d_ev_amt = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_noise); cudaFree(d_noise64); cudaFree(d_ev_off); cudaFree(d_ev_inst); cudaFree(d_ev_amt);
    cudaFree(d_force); cudaFree(d_rec); cudaFree(d_rec_off);
    p.rec_ids = nullptr; p.rec_off = nullptr; p.rec_cap = 0;
  };
#define STEP_OK(expr)                                              \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) {                                       \
      err = std::string(#expr) + ": " + cudaGetErrorString(_e);    \
      cleanup();                                                   \
      return SNN_ERR_CUDA;                                         \
    }                                                              \
  } while (0)

  STEP_OK(cudaEventRecord(ev_begin_, stream_));
  if (cfg_.noise_mode == SNN_NOISE_REPLAY) {
    const size_t cnt = (size_t)n * N;
    if (sizeof(T) == 8) {
      STEP_OK(cudaMalloc(&d_noise, cnt * sizeof(T)));
      STEP_OK(cudaMemcpyAsync(d_noise, a.noise, cnt * sizeof(double), cudaMemcpyHostToDevice, stream_));
    } else {
      STEP_OK(cudaMalloc(&d_noise64, cnt * sizeof(double)));
      STEP_OK(cudaMalloc(&d_noise, cnt * sizeof(T)));
      STEP_OK(cudaMemcpyAsync(d_noise64, a.noise, cnt * sizeof(double), cudaMemcpyHostToDevice, stream_));
      from_double_kernel<T><<<grid_for((int64_t)cnt), 256, 0, stream_>>>(d_noise64, d_noise, (int64_t)cnt);
    }
  }
  DaEvents ev{nullptr, nullptr, nullptr};
  if (a.n_da > 0) {
    std::vector<int32_t> off(n + 1, 0), inst(a.n_da);
    std::vector<double> amt(a.n_da);
    int prev = -1;
    for (int q = 0; q < a.n_da; ++q) {
      const snn_da_event& e = a.da[q];
      if (e.step < 0 || e.step >= n || e.step < prev || e.instance < 0 || e.instance >= p.B) {
        err = "da events must be sorted by step, inside the window, instance in range"; cleanup(); return SNN_ERR_INVALID;
      }
      prev = e.step; off[e.step + 1]++; inst[q] = e.instance; amt[q] = e.amount;
    }
    for (int k = 0; k < n; ++k) off[k + 1] += off[k];
    STEP_OK(cudaMalloc(&d_ev_off, (n + 1) * sizeof(int32_t)));
    STEP_OK(cudaMalloc(&d_ev_inst, a.n_da * sizeof(int32_t)));
    STEP_OK(cudaMalloc(&d_ev_amt, a.n_da * sizeof(double)));
    STEP_OK(cudaMemcpyAsync(d_ev_off, off.data(), (n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
    STEP_OK(cudaMemcpyAsync(d_ev_inst, inst.data(), a.n_da * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
    STEP_OK(cudaMemcpyAsync(d_ev_amt, amt.data(), a.n_da * sizeof(double), cudaMemcpyHostToDevice, stream_));
    STEP_OK(cudaStreamSynchronize(stream_));  // host vectors go out of scope
    ev = DaEvents{d_ev_off, d_ev_inst, d_ev_amt};
  }
  std::vector<int32_t> force_off(n + 1, 0);
  if (a.n_force > 0) {
    std::vector<int32_t> ids(a.n_force);
    int prev = -1;
    for (int q = 0; q < a.n_force; ++q) {
      const snn_force_event& e = a.force[q];
      if (e.step < 0 || e.step >= n || e.step < prev || e.neuron < 0 || e.neuron >= N) {
        err = "force events must be sorted by step, inside the window, neuron in range"; cleanup(); return SNN_ERR_INVALID;
      }
      prev = e.step; force_off[e.step + 1]++; ids[q] = e.neuron;
    }
    for (int k = 0; k < n; ++k) force_off[k + 1] += force_off[k];
    STEP_OK(cudaMalloc(&d_force, a.n_force * sizeof(int32_t)));
    STEP_OK(cudaMemcpyAsync(d_force, ids.data(), a.n_force * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
    STEP_OK(cudaStreamSynchronize(stream_));
  }
  if (a.spikes_out) {
    STEP_OK(cudaMalloc(&d_rec, a.spikes_cap * sizeof(int32_t)));
    STEP_OK(cudaMalloc(&d_rec_off, (n + 1) * sizeof(int32_t)));
    STEP_OK(cudaMemsetAsync(d_rec_off, 0, (n + 1) * sizeof(int32_t), stream_));
    p.rec_ids = d_rec; p.rec_off = d_rec_off; p.rec_cap = a.spikes_cap;
  }
  unsigned long long c_before[4];
  STEP_OK(cudaMemcpyAsync(c_before, p.counters, sizeof(c_before), cudaMemcpyDeviceToHost, stream_));
  STEP_OK(cudaMemsetAsync(p.blk_counters, 0, 2 * kMaxSynBlocks * sizeof(unsigned long long), stream_));

  prof_used_ = 0; prof_cls_.clear();
  // one quad of neurons per thread; persistent synapse grid sized to what is resident
  const int gN = (int)std::max<int64_t>(1, std::min<int64_t>((p.Nld / 4 + 255) / 256, (int64_t)n_sm_ * 8));
  if (gS_ == 0) {
    int per_sm = 0;
    if (p.exact) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, synapse_kernel<T, false>, 256, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, synapse_kernel<T, true>, 256, 0);
    gS_ = std::min(kMaxSynBlocks, n_sm_ * std::max(1, per_sm));
  }
  const int gS = gS_;
  // learn grid: x blocks per instance
  int64_t plastic_per_inst = p.B ? (p.E / p.B) : 0;
  int gLx = (int)std::max<int64_t>(1, std::min<int64_t>((plastic_per_inst / 2 + 256 * 4 - 1) / (256 * 4),
                                                         std::max<int64_t>(1, (int64_t)n_sm_ * 8 / p.B)));
  dim3 gL(gLx, p.B);

  for (int k = 0; k <= n; ++k) {
    const int64_t t = t_ + k;
    const int slot = slot_host(t), prev = slot_host(t - 1);
    const T* noise_prev = (d_noise && k > 0) ? d_noise + (size_t)(k - 1) * N : nullptr;
    const bool forced = k < n && force_off[k + 1] > force_off[k];
    prof_begin(0);
    if (k == 0) {
      if (forced) force_kernel<T><<<(force_off[1] - force_off[0] + 255) / 256, 256, 0, stream_>>>(p, d_force + force_off[0], force_off[1] - force_off[0]);
      neuron_kernel<T, false, true><<<gN, 256, 0, stream_>>>(p, t, k, slot, prev, noise_prev, ev);
    } else if (k == n) {
      neuron_kernel<T, true, false><<<gN, 256, 0, stream_>>>(p, t, k, slot, prev, noise_prev, ev);
    } else if (forced) {
      neuron_kernel<T, true, false><<<gN, 256, 0, stream_>>>(p, t, k, slot, prev, noise_prev, ev);
      const int nf = force_off[k + 1] - force_off[k];
      force_kernel<T><<<(nf + 255) / 256, 256, 0, stream_>>>(p, d_force + force_off[k], nf);
      neuron_kernel<T, false, true><<<gN, 256, 0, stream_>>>(p, t, k, slot, prev, noise_prev, DaEvents{nullptr, nullptr, nullptr});
    } else {
      neuron_kernel<T, true, true><<<gN, 256, 0, stream_>>>(p, t, k, slot, prev, noise_prev, ev);
    }
    prof_end();
    if (k == n) break;
    prof_begin(1);
    if (p.exact) {
      const T* noise_now = d_noise ? d_noise + (size_t)k * N : nullptr;
      gather_kernel<T><<<grid_for(N), 256, 0, stream_>>>(p, t, noise_now);
      synapse_kernel<T, false><<<gS, 256, 0, stream_>>>(p, t, k, slot, g_arr_log2_, g_ltp_log2_);
    } else {
      synapse_kernel<T, true><<<gS, 256, 0, stream_>>>(p, t, k, slot, g_arr_log2_, g_ltp_log2_);
    }
    prof_end();
    const bool train = a.train_flags && a.train_flags[k];
    const bool decay_step = cfg_.sd_decay_mode == SNN_SD_DECAY_EVERY_STEP;
    if ((train || decay_step) && p.E > 0 && p.Ne > 0) {
      prof_begin(2);
      learn_kernel<T><<<gL, 256, 0, stream_>>>(p, train ? 1 : 0,
                                                (decay_step || (train && cfg_.sd_decay_mode == SNN_SD_DECAY_ON_LEARN)) ? 1 : 0);
      prof_end();
    }
  }
  STEP_OK(cudaGetLastError());
  t_ += n;

  // ---- results
  int rc = SNN_OK;
  std::vector<int32_t> h_off;
  if (a.spikes_out) {
    h_off.resize(n + 1);
    STEP_OK(cudaMemcpyAsync(h_off.data(), d_rec_off, (n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
    STEP_OK(cudaStreamSynchronize(stream_));
    const int64_t total = h_off[n];
    int src = sort_segments(d_rec, total, d_rec_off, n, stream_, err);
    if (src) { cleanup(); return src; }
    if (total > 0) STEP_OK(cudaMemcpyAsync(a.spikes_out, d_rec, total * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
    for (int k = 0; k <= n; ++k) a.spike_offsets[k] = h_off[k];
  }
  unsigned long long c_after[4];
  STEP_OK(cudaMemcpyAsync(c_after, p.counters, sizeof(c_after), cudaMemcpyDeviceToHost, stream_));
  std::vector<unsigned long long> h_blk(2 * (size_t)gS);
  STEP_OK(cudaMemcpyAsync(h_blk.data(), p.blk_counters, h_blk.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream_));
  STEP_OK(cudaEventRecord(ev_end_, stream_));
  STEP_OK(cudaStreamSynchronize(stream_));
  if (a.spikes_out && (int64_t)(c_after[0] - c_before[0]) > a.spikes_cap) {
    err = "spikes_cap too small: record truncated"; rc = SNN_ERR_CAPACITY;
  }
  if (a.stats) {
    snn_step_stats& s = *a.stats;
    memset(&s, 0, sizeof(s));
    s.n_spikes = (int64_t)(c_after[0] - c_before[0]);
    s.n_syn_events = (int64_t)(c_after[1] - c_before[1]);
    for (int b = 0; b < gS; ++b) { s.n_syn_events += (int64_t)h_blk[2 * b]; s.n_ltp_events += (int64_t)h_blk[2 * b + 1]; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev_begin_, ev_end_);
    s.device_ms = ms;
    for (size_t q = 0; q + 1 < prof_used_ + 1 && q / 2 < prof_cls_.size(); q += 2) {
      float m = 0.f;
      cudaEventElapsedTime(&m, prof_events_[q], prof_events_[q + 1]);
      int cls = prof_cls_[q / 2];
      if (cls == 0) { s.neuron_ms += m; s.neuron_launches++; }
      else if (cls == 1) { s.synapse_ms += m; s.synapse_launches++; }
      else { s.learn_ms += m; s.learn_launches++; }
    }
  }
  cleanup();
#undef STEP_OK
  return rc;
}

}  // namespace snn
