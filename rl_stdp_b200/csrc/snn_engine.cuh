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
//   I = noise                          :111       neuron_kernel(t):  Euler(t-1) then detect(t):
//   spiked = v > 30 ; v=-65 ; u+=d     :112-114     threshold/reset, trace set/decay, bitmap,
//   for n in spiked: I += s[n,:]       :116-118     spike list, dopamine decay
//   v,v,u Euler                        :120-122   synapse_kernel(t): LTD for every spike arriving now
//   STDP[spiked]=0.1                   :123         + I += w for the arrivals of t+1 (delay ring of lists)
//   sd[:,n]+=STDP ; sd[n,:]-=1.5STDP   :125-128   ltp_kernel(t): LTP over the incoming index
//   STDP*=0.95 ; da*=0.995             :130-131   [exact mode: gather_kernel(t) pulls I in
//   learn! (if train)                  :135-142     ascending presynaptic order]
//                                                 learn_kernel(t): w=clamp(w+(b+da)*sd), sd*=.99
//
// Detection at t only needs the Euler result of t-1, so Euler(t-1) and detect(t) are one pass
// over the neuron arrays (36 B/neuron fp32).  Trace values do not enter the Euler update, so the
// result is identical to the reference order.
//
// Scheduling (production mode): only neuron(k) -> neuron(k+1) is a true serial chain.  The
// currents consumed by Euler(k) are pushed ahead of time into a parity-double-buffered Iacc:
//   - delay-1 synapses of step-k spikers: inside neuron_kernel(k), right after detection;
//   - every other arrival of step k+1 (spikes of steps <= k): by synapse_kernel(k), which walks
//     the delay-(a+1) segment for LTD (arrival now) and the adjacent delay-(a+2) segment for the
//     push (arrival at k+1) in one pass.
// sd updates (LTD in synapse_kernel, LTP in ltp_kernel) never feed the neuron update, so they run
// on side streams.  A window is captured once as a CUDA graph over three streams:
//   s1: neuron(0) -> neuron(1) -> neuron(2) ...           (waits synapse(k-2), learn(k-1))
//   s2: synapse(k) after neuron(k), ltp(k-1)              [learn steps: LTD | learn | push]
//   s3: ltp(k) after neuron(k), synapse(k-1) ; learn(k) after synapse(k)
// All step-dependent values (absolute step, ring slot, input pointers) come from a device-side
// WindowCtx, so one instantiated graph serves every window with the same length and train-flag
// pattern.  Exact mode, forced spikes and per-kernel profiling use plain stream-ordered launches.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

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
__device__ __forceinline__ double philox_to_noise(uint32_t x, double amp) {
  double U = __dmul_rn(__dadd_rn((double)x, 0.5), 1.0 / 4294967296.0);
  return __dmul_rn(amp, __dsub_rn(U, 0.5));
}
__device__ __forceinline__ double philox_noise(uint64_t seed, int64_t t, int64_t j, double amp) {
  uint4 o = philox4x32_10(make_uint4((uint32_t)(j >> 2), 0u, (uint32_t)t, (uint32_t)((uint64_t)t >> 32)),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t x = (j & 3) == 0 ? o.x : ((j & 3) == 1 ? o.y : ((j & 3) == 2 ? o.z : o.w));
  return philox_to_noise(x, amp);
}

constexpr int kMaxSynBlocks = 148 * 16 * 2;

// Network state + parameters, passed to the kernels by value.
template <typename T>
struct DevNet {
  typedef typename Real2<T>::type T2;
  int64_t N, Nld, Nper, Ne, E;  // Nld = N rounded up to 128: leading dimension of ring slots
  int32_t Dmax, R, B, W;        // R = Dmax+2 ring slots ; W = bitmap words per slot
  // neuron state
  T *v, *u, *ho;
  T* Iacc;           // [2][Nld]  Iacc[t&1] = currents consumed by the Euler update of step t
  T* trace;          // [R][Nld]  trace after the set of that step (pre-decay)
  uint32_t* bits;    // [R][W]    spike bitmap per step
  int32_t* spk_list; // [R][Nld]  spike list per step (unordered)
  int32_t* spk_cnt;  // [R]
  T* da;             // [2][B]    da[t&1] = dopamine after the decay of step t-1
  // synapses (internal order: ascending pre, delay, caller position)
  const uint32_t* seg;   // [N][Dmax+1] delay-segment table
  const int32_t* post;   // [E]
  T2* wsd;               // [E] (w, sd) interleaved: one 8-byte access serves propagate + LTD
  const uint32_t *colptr, *col_npl, *in_syn, *in_pre;
  // counters[0] = spikes, [1] = synaptic events (exact mode); per-block partial sums otherwise
  unsigned long long* counters;
  unsigned long long* blk_ev;   // [kMaxSynBlocks] synaptic events (synapse_kernel)
  unsigned long long* blk_ltp;  // [kMaxSynBlocks] LTP visits (ltp_kernel)
  // parameters
  T vthresh, v_reset, a_exc, a_inh, d_exc, d_inh, trace_set, trace_decay, apost, apre, da_decay,
      learn_base, eta, wmin, wmax, sd_decay, theta, ttheta;
  double noise_amp; uint64_t noise_seed;
  int64_t ho_from;
  int32_t threshold_ge, sd_decay_mode, rate_mode, plastic_ei, noise_mode, exact, homeostasis;
  int32_t exp_flags;  // experiments only (snn_config.reserved[1]); 0 in production
  // neuron partition (multi-GPU): this rank integrates neurons [part_lo, part_hi)
  // and holds every synapse whose POST is in that range; spike bitmaps are exchanged per step
  int64_t part_lo, part_hi;
};

// Per-window inputs/outputs.  Lives in device memory so that a captured graph never has to be
// re-parameterised: kernels receive (net, ctx pointer, k) and derive everything else.
template <typename T>
struct WindowCtx {
  int64_t t0;        // absolute step of window step 0
  int32_t slot0;     // t0 mod R
  int32_t n_steps;
  const T* noise;    // [n_steps][N] replayed thalamic input, or null
  const int32_t* ev_off; const int32_t* ev_inst; const double* ev_amt;  // dopamine events, CSR by step
  int32_t* rec_ids; int32_t* rec_off; int64_t rec_cap;                  // spike record
};

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
__device__ __forceinline__ int wrap_slot(int s, int R) { return s >= R ? s - R : (s < 0 ? s + R : s); }

// ---------------------------------------------------------------------------------------------
// K1 neuron_kernel: [Euler(t-1)] + [detect(t)]   (new_net.jl:111-114,120-123,130-131;
// homeostasis net_res.jl:167-171,204)
//   EULER : I = noise(t-1) + Iacc ; two half-steps of v ; u update ; Iacc cleared
//   DETECT: spike test, reset, homeostasis, trace set/decay into ring slot t, spike bitmap,
//           spike list append; block 0: da[(t+1)&1] = (da[t&1] + events(t-1)) * decay
// Each thread owns 4 consecutive neurons: every array is read/written with one 16-byte (fp32)
// access, all loads are issued before the first use (the scalar version had one 128-B request in
// flight per warp and ran at 0.5 TB/s — profiles/r1a), one Philox4x32 call yields the four noise
// values.  A warp covers 128 neurons = 4 bitmap words.  Arrays are padded to Nld.
// ---------------------------------------------------------------------------------------------
template <typename T, bool EULER, bool DETECT>
__global__ void __launch_bounds__(256)
neuron_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k, int32_t apply_ev) {
  typedef Ops<T> O;
  const int lane = threadIdx.x & 31;
  const int64_t t = cx->t0 + k;
  const int slot = (cx->slot0 + k) % p.R, prev = wrap_slot(slot - 1, p.R);
  T* __restrict__ tr = p.trace + (size_t)slot * p.Nld;
  const T* __restrict__ trp = p.trace + (size_t)prev * p.Nld;
  uint32_t* __restrict__ bits = p.bits + (size_t)slot * p.W;
  int32_t* __restrict__ list = p.spk_list + (size_t)slot * p.Nld;
  int32_t* __restrict__ rec_ids = cx->rec_ids;
  const int64_t rec_cap = cx->rec_cap;
  // spike-record offset of this step = offset of the previous step + its (final) spike count
  int32_t rec_base = 0;
  const int n_prev = (k > 0 && (rec_ids || blockIdx.x == 0)) ? p.spk_cnt[prev] : 0;
  if (rec_ids && k > 0) {
    int64_t nxt = (int64_t)cx->rec_off[k - 1] + n_prev;
    rec_base = (int32_t)(nxt < rec_cap ? nxt : rec_cap);
  }
  T* __restrict__ I_prev = p.Iacc + (size_t)((t - 1) & 1) * p.Nld;  // consumed by Euler(t-1)
  T* __restrict__ I_now = p.Iacc + (size_t)(t & 1) * p.Nld;         // receives the pushes of step t
  const T* __restrict__ noise_prev = (EULER && cx->noise) ? cx->noise + (size_t)(k - 1) * p.N : nullptr;
  const uint32_t N = (uint32_t)p.N, Nper = (uint32_t)p.Nper, Ne = (uint32_t)p.Ne;

  // dopamine (new_net.jl:131 da *= 0.995 ; newnet_DASTDP.jl:54-56 da += 0.5 after step!)
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      if (EULER && k > 0) {  // exactly one EULER launch per step k >= 1 (and the window-end launch)
        p.counters[0] += (unsigned long long)n_prev;
        if (rec_ids) cx->rec_off[k] = rec_base;
      }
      // next step's list; R = Dmax+2 so no kernel still in flight reads that slot
      if (DETECT) p.spk_cnt[wrap_slot(slot + 1, p.R)] = 0;
    }
    const T* __restrict__ din = p.da + (size_t)(t & 1) * p.B;
    T* __restrict__ dout = p.da + (size_t)((t + 1) & 1) * p.B;
    const bool evs = apply_ev && cx->ev_off && k > 0;
    const int q0 = evs ? cx->ev_off[k - 1] : 0, q1 = evs ? cx->ev_off[k] : 0;
    for (int b = threadIdx.x; b < p.B; b += blockDim.x) {
      T d = din[b];
      for (int q = q0; q < q1; ++q)
        if (cx->ev_inst[q] == b) d = O::add(d, (T)cx->ev_amt[q]);
      if (DETECT) dout[b] = O::mul(d, p.da_decay);
      else if (q1 > q0) p.da[(size_t)(t & 1) * p.B + b] = d;  // Euler-only launch: events only, in place
    }
  }

  // whole warps only: the last rank's range is padded up to Nld (padding neurons are inactive)
  const uint32_t q_lo = (uint32_t)(p.part_lo >> 2), q_hi = (uint32_t)((p.part_hi >= p.N ? p.Nld : p.part_hi) >> 2);
  for (uint32_t q = q_lo + blockIdx.x * blockDim.x + threadIdx.x; q < q_hi; q += gridDim.x * blockDim.x) {
    const uint32_t j0 = q << 2;
    // ---- all loads first
    Vec4<T> v = ld4(p.v + j0), u = ld4(p.u + j0), I, tp, ho;
    if (EULER) I = ld4(I_prev + j0);
    if (DETECT) {
      tp = ld4(trp + j0);
      if (p.homeostasis) ho = ld4(p.ho + j0);
    }
    const uint32_t l0 = p.B == 1 ? j0 : j0 % Nper;
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
          for (int c = 0; c < 4; ++c) I.x[c] = O::add((T)philox_to_noise(w[c], p.noise_amp), I.x[c]);
        }
        Vec4<T> z; z.x[0] = z.x[1] = z.x[2] = z.x[3] = (T)0;
        st4(I_prev + j0, z);
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
              if (rec_ids && (int64_t)rec_base + pos < rec_cap) rec_ids[rec_base + pos] = (int32_t)(j0 + c);
            }
            if (!p.exact && !(p.exp_flags & 1)) {
              // delay-1 synapses of this step's spikers feed the Euler update of this very step
              // (new_net.jl:116-121): push them now, warp-cooperatively, into Iacc[t&1]
              unsigned mm = m;
              while (mm) {
                const int src = __ffs(mm) - 1; mm &= mm - 1;
                const uint32_t i = __shfl_sync(0xffffffffu, j0, src) + c;
                const uint32_t s0 = p.seg[(size_t)i * (p.Dmax + 1)], s1 = p.seg[(size_t)i * (p.Dmax + 1) + 1];
                // four loads in flight per lane before the first reduction (an inhibitory row has
                // ~100 delay-1 synapses; one round trip instead of four)
                for (uint32_t e0 = s0 + lane; e0 < s1; e0 += 128) {
                  int32_t pj[4]; T pw[4];
#pragma unroll
                  for (int q4 = 0; q4 < 4; ++q4) {
                    const uint32_t e = e0 + 32 * q4;
                    pj[q4] = e < s1 ? p.post[e] : -1;
                    pw[q4] = e < s1 ? p.wsd[e].x : (T)0;
                  }
#pragma unroll
                  for (int q4 = 0; q4 < 4; ++q4) if (pj[q4] >= 0) atomicAdd(&I_now[pj[q4]], pw[q4]);
                }
              }
            }
          }
        }
      }
    }
    st4(p.v + j0, v);
    st4(p.u + j0, u);
  }
}

// ---------------------------------------------------------------------------------------------
// remote_kernel(k) — neuron-partitioned runs only.  After the spike bitmap of step k has been
// all-gathered, every rank replays detect(k) for the neurons it does NOT own: trace set/decay
// (the replicated pre-trace ring that LTP gathers from), spike-list append, spike record and the
// delay-1 current push of remote spikers into local targets.  Same quad/warp mapping as
// neuron_kernel; no v/u traffic (16 B/neuron).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
remote_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k) {
  typedef Ops<T> O;
  const int lane = threadIdx.x & 31;
  const int64_t t = cx->t0 + k;
  const int slot = (cx->slot0 + k) % p.R, prev = wrap_slot(slot - 1, p.R);
  T* __restrict__ tr = p.trace + (size_t)slot * p.Nld;
  const T* __restrict__ trp = p.trace + (size_t)prev * p.Nld;
  const uint32_t* __restrict__ bits = p.bits + (size_t)slot * p.W;
  int32_t* __restrict__ list = p.spk_list + (size_t)slot * p.Nld;
  int32_t* __restrict__ rec_ids = cx->rec_ids;
  const int64_t rec_cap = cx->rec_cap;
  int32_t rec_base = 0;
  if (rec_ids && k > 0) {
    int64_t nxt = (int64_t)cx->rec_off[k - 1] + p.spk_cnt[prev];
    rec_base = (int32_t)(nxt < rec_cap ? nxt : rec_cap);
  }
  T* __restrict__ I_now = p.Iacc + (size_t)(t & 1) * p.Nld;
  // whole warps only: the last rank's range is padded up to Nld (padding neurons are inactive)
  const uint32_t q_lo = (uint32_t)(p.part_lo >> 2), q_hi = (uint32_t)((p.part_hi >= p.N ? p.Nld : p.part_hi) >> 2);
  const uint32_t nquad = (uint32_t)(p.Nld >> 2), n_remote = nquad - (q_hi - q_lo);
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_remote; r += gridDim.x * blockDim.x) {
    const uint32_t q = r < q_lo ? r : r + (q_hi - q_lo);  // part bounds are multiples of 128: warps stay whole
    const uint32_t j0 = q << 2;
    Vec4<T> tp = ld4(trp + j0);
    const unsigned nib = (bits[j0 >> 5] >> (j0 & 31)) & 0xFu;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const bool s = (nib >> c) & 1u;
      tp.x[c] = s ? p.trace_set : (t > 0 ? O::mul(tp.x[c], p.trace_decay) : (T)0);
    }
    st4(tr + j0, tp);
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
            if (rec_ids && (int64_t)rec_base + pos < rec_cap) rec_ids[rec_base + pos] = (int32_t)(j0 + c);
          }
          if (!p.exact) {
            unsigned mm = m;
            while (mm) {
              const int src = __ffs(mm) - 1; mm &= mm - 1;
              const uint32_t i = __shfl_sync(0xffffffffu, j0, src) + c;
              const uint32_t s0 = p.seg[(size_t)i * (p.Dmax + 1)], s1 = p.seg[(size_t)i * (p.Dmax + 1) + 1];
              for (uint32_t e = s0 + lane; e < s1; e += 32) atomicAdd(&I_now[p.post[e]], p.wsd[e].x);
            }
          }
        }
      }
    }
  }
}

// v[idx] = 40 (input! new_net.jl:104-106)
template <typename T>
__global__ void force_kernel(DevNet<T> p, const int32_t* __restrict__ idx, int n) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) p.v[idx[q]] = (T)40.0;
}

// input!(net, ::Array{Tuple{Int64,Float64}}) net_arch.jl:316-326: two Euler half-steps of v[idx]
// with the injected current only (I is zeroed again by spike!, :330)
template <typename T>
__global__ void current_kernel(DevNet<T> p, const int32_t* __restrict__ idx, const double* __restrict__ cur, int n) {
  typedef Ops<T> O;
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int32_t j = idx[q];
  const T I = (T)cur[q], u = p.u[j];
  T v = p.v[j];
#pragma unroll
  for (int rep = 0; rep < 2; ++rep) {
    T x = O::add(O::mul((T)0.04, v), (T)5);
    x = O::add(O::mul(x, v), (T)140);
    x = O::add(O::sub(x, u), I);
    v = O::add(v, O::mul((T)0.5, x));
  }
  p.v[j] = v;
}

// ---------------------------------------------------------------------------------------------
// exact mode: I[j] = noise[j] + sum over spiking sources in ascending presynaptic id — the
// accumulation order of `for neuron in spiked: I .+= s[neuron,:]` (new_net.jl:116-118).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gather_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k) {
  typedef Ops<T> O;
  const int64_t t = cx->t0 + k;
  const int slot = (cx->slot0 + k) % p.R;
  const T* __restrict__ noise = cx->noise ? cx->noise + (size_t)k * p.N : nullptr;
  const int64_t j_hi = p.part_hi < p.N ? p.part_hi : p.N;
  for (int64_t j = p.part_lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < j_hi;
       j += (int64_t)gridDim.x * blockDim.x) {
    T acc = (T)0;
    if (p.noise_mode == SNN_NOISE_REPLAY) acc = noise[j];
    else if (p.noise_mode == SNN_NOISE_PHILOX) acc = (T)philox_noise(p.noise_seed, t, j, p.noise_amp);
    const uint32_t c0 = p.colptr[j], c1 = p.colptr[j + 1];
    for (uint32_t q = c0; q < c1; ++q) {
      const uint32_t pk = p.in_pre[q];
      const uint32_t i = pk & kPreMask;
      const int sl = wrap_slot(slot - (int)(pk >> kPreBits), p.R);
      const uint32_t word = p.bits[(size_t)sl * p.W + (i >> 5)];
      if ((word >> (i & 31)) & 1u) acc = O::add(acc, p.wsd[p.in_syn[q]].x);
    }
    p.Iacc[(size_t)(t & 1) * p.Nld + j] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// K2 synapse_kernel(k): persistent grid, sub-warp groups of G lanes per work item.  A work item is
// (age a, neuron i that spiked at step k-a), a = 0..Dmax-1; its CSR row is sorted by delay, so the
// delay-(a+1) segment (arriving NOW) and the delay-(a+2) segment (arriving at k+1) are adjacent:
//   LTD      on the first: sd -= apre*trace_post(k) (new_net.jl:127 `sd[n,:] .-= 1.5 .* STDP`); if
//            the target also spiked at k the same thread applies its LTP too, in the reference's
//            loop order (post<=pre: + then -, else - then +; new_net.jl:125-128).  ltp_kernel skips
//            those synapses, so every synapse is written by exactly one thread per step: no
//            atomics on sd, bit-reproducible (exact mode).  Production mode (ATOMIC) uses
//            red.global.add on sd instead: no load of sd, no ordering between the LTD and LTP
//            kernels of neighbouring steps, so they pipeline freely on their side streams.
//   PUSHNEXT on the second: Iacc[(k+1)&1][post] += w — the currents of step k+1 are complete one
//            step early, which takes propagation off the neuron(k)->neuron(k+1) critical path.
// (delay-1 pushes of step-k spikers happen inside neuron_kernel(k).)
// ---------------------------------------------------------------------------------------------
template <typename T, bool LTD, bool PUSHNEXT, bool ATOMIC>
__global__ void __launch_bounds__(256)
synapse_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k, int g_log2) {
  typedef Ops<T> O;
  __shared__ int cum[kMaxDelay + 2];
  __shared__ int slots[kMaxDelay + 1];
  __shared__ unsigned long long blk_ev;
  const int64_t t = cx->t0 + k;
  const int slot = (int)(((int64_t)cx->slot0 + k + p.R) % p.R);  // k may be -1 (window-start push)
  // prologue: the Dmax list lengths are loaded in parallel (one round trip), scanned in smem
  if (threadIdx.x < (unsigned)p.Dmax) {
    const int sl = wrap_slot(slot - (int)threadIdx.x, p.R);
    slots[threadIdx.x] = sl;
    cum[threadIdx.x + 1] = p.spk_cnt[sl];
  }
  if (threadIdx.x == 0) blk_ev = 0ull;
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int a = 0; a < p.Dmax; ++a) { int c = cum[a + 1]; cum[a] = acc; acc += c; }
    cum[p.Dmax] = acc;
  }
  __syncthreads();
  const int total = cum[p.Dmax];
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  const uint32_t Nper = (uint32_t)p.Nper, Ne = (uint32_t)p.Ne;
  const size_t Nld = (size_t)p.Nld;
  const T* __restrict__ tr_now = p.trace + (size_t)slot * Nld;
  const uint32_t* __restrict__ bits_now = p.bits + (size_t)slot * p.W;
  T* __restrict__ I_next = p.Iacc + (size_t)((t + 1) & 1) * Nld;
  unsigned long long n_ev = 0;
  const uint32_t G = 1u << g_log2;
  const uint32_t gl = gtid & (G - 1);
  const uint32_t ngroups = nthreads >> g_log2;
  for (uint32_t item = gtid >> g_log2; item < (uint32_t)total; item += ngroups) {
    int a = 0;
    while (a + 1 < p.Dmax && cum[a + 1] <= (int)item) ++a;
    const int sl = slots[a];
    const uint32_t i = (uint32_t)p.spk_list[(size_t)sl * Nld + (item - cum[a])];
    const uint32_t* __restrict__ sg = p.seg + (size_t)i * (p.Dmax + 1) + a;
    const uint32_t s0 = sg[0], s1 = sg[1];
    const uint32_t s2 = (PUSHNEXT && a + 1 < p.Dmax) ? sg[2] : s1;
    const uint32_t ibase = p.B == 1 ? 0u : (i / Nper) * Nper;
    const bool row_plastic = LTD && (i - ibase) < Ne;
    const T pre_up = (row_plastic && !ATOMIC) ? O::mul(p.apost, p.trace[(size_t)sl * Nld + i]) : (T)0;
    if (LTD && gl == 0) n_ev += (s1 - s0);
    for (uint32_t e = (LTD ? s0 : s1) + gl; e < s2; e += G) {
      const uint32_t j = (uint32_t)p.post[e];
      if (e >= s1) {                       // arrives at k+1: push only
        atomicAdd(&I_next[j], p.wsd[e].x);
      } else if (row_plastic && (p.plastic_ei || (j - ibase) < Ne)) {
        const T dn = O::mul(p.apre, tr_now[j]);
        if (ATOMIC) {  // production: fire-and-forget reduction, LTP of this step comes from ltp_kernel
          atomicAdd(&p.wsd[e].y, -dn);
          continue;
        }
        const bool jsp = (bits_now[j >> 5] >> (j & 31)) & 1u;
        T sd = p.wsd[e].y;
        if (jsp) sd = (j <= i) ? O::sub(O::add(sd, pre_up), dn) : O::add(O::sub(sd, dn), pre_up);
        else sd = O::sub(sd, dn);
        p.wsd[e].y = sd;
      }
    }
  }
  if (LTD) {
    // block-local reduction, then one plain RMW on this block's own slot (no contended atomics)
    for (int o = 16; o > 0; o >>= 1) n_ev += __shfl_down_sync(0xffffffffu, n_ev, o);
    if ((threadIdx.x & 31) == 0 && n_ev) atomicAdd(&blk_ev, n_ev);
    __syncthreads();
    if (threadIdx.x == 0 && blk_ev) p.blk_ev[blockIdx.x] += blk_ev;
  }
}

// ---------------------------------------------------------------------------------------------
// K3 ltp_kernel(t): for every neuron j that spiked at t walk the plastic prefix of its incoming
// list (new_net.jl:126 `sd[:,n] .+= STDP`): sd += apost * trace_pre as seen at the synapse
// (trace of step t-(d-1), Old_net/DASTDP.jl:62-68).  Synapses whose presynaptic spike arrives
// at this very call are owned by synapse_kernel (see there).  Three gathers are kept in flight
// per lane.
// ---------------------------------------------------------------------------------------------
template <typename T, bool ATOMIC>
__global__ void __launch_bounds__(256)
ltp_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k, int g_log2) {
  typedef Ops<T> O;
  __shared__ unsigned long long blk_ltp;
  const int slot = (cx->slot0 + k) % p.R;
  const int n_now = p.spk_cnt[slot];
  if (threadIdx.x == 0) blk_ltp = 0ull;
  __syncthreads();
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  const size_t Nld = (size_t)p.Nld;
  const uint32_t G = 1u << g_log2;
  const uint32_t gl = gtid & (G - 1);
  const uint32_t ngroups = nthreads >> g_log2;
  const int32_t* __restrict__ list = p.spk_list + (size_t)slot * Nld;
  unsigned long long n_ltp = 0;
  for (uint32_t item = gtid >> g_log2; item < (uint32_t)n_now; item += ngroups) {
    const int32_t j = list[item];
    const uint32_t c0 = p.colptr[j], c1 = c0 + p.col_npl[j];
    if (gl == 0) n_ltp += (c1 - c0);
    for (uint32_t kb = c0 + gl; kb < c1; kb += 3 * G) {
      uint32_t pk[3], e[3]; bool ok[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const uint32_t kk = kb + q * G;
        ok[q] = kk < c1;
        pk[q] = ok[q] ? p.in_pre[kk] : 0u;
        e[q] = ok[q] ? p.in_syn[kk] : 0u;
      }
      size_t toff[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const uint32_t i = pk[q] & kPreMask;
        const int sl = wrap_slot(slot - (int)(pk[q] >> kPreBits), p.R);
        toff[q] = (size_t)sl * Nld + i;
        if (!ATOMIC && ok[q]) {
          const uint32_t word = p.bits[(size_t)sl * p.W + (i >> 5)];
          ok[q] = !((word >> (i & 31)) & 1u);
        }
      }
      T trv[3], sdv[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        trv[q] = ok[q] ? p.trace[toff[q]] : (T)0;
        sdv[q] = (!ATOMIC && ok[q]) ? p.wsd[e[q]].y : (T)0;
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (!ok[q]) continue;
        if (ATOMIC) atomicAdd(&p.wsd[e[q]].y, O::mul(p.apost, trv[q]));
        else p.wsd[e[q]].y = O::add(sdv[q], O::mul(p.apost, trv[q]));
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) n_ltp += __shfl_down_sync(0xffffffffu, n_ltp, o);
  if ((threadIdx.x & 31) == 0 && n_ltp) atomicAdd(&blk_ltp, n_ltp);
  __syncthreads();
  if (threadIdx.x == 0 && blk_ltp) p.blk_ltp[blockIdx.x] += blk_ltp;
}

// ---------------------------------------------------------------------------------------------
// K4 learn_kernel: one streaming pass over the plastic synapses of every instance
//   w = clamp(w + rate*sd, wmin, wmax) ; sd *= sd_decay      (new_net.jl:135-142)
// rate = (learn_base+da) | da | eta*da with da as decayed by this step (new_net.jl:131 runs
// before learn!).  16 B/synapse fp32 (read w,sd; write w,sd); four 16-byte loads in flight per
// thread.  blockIdx.y = instance; exc rows of an instance are contiguous in the CSR.
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
learn_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k, int do_learn, int do_decay) {
  typedef Ops<T> O;
  typedef typename DevNet<T>::T2 T2;
  const int b = blockIdx.y;
  const int64_t t = cx->t0 + k;
  const int64_t r0 = (int64_t)b * p.Nper;
  const uint32_t lo = p.seg[(size_t)r0 * (p.Dmax + 1)];
  const uint32_t hi = p.seg[(size_t)(r0 + p.Ne - 1) * (p.Dmax + 1) + p.Dmax];
  const T da = p.da[(size_t)((t + 1) & 1) * p.B + b];
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

// grow-only device buffer
struct DevBuf {
  void* ptr = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
  template <typename U> U* as() const { return reinterpret_cast<U*>(ptr); }
};

// ---------------------------------------------------------------------------------------------
template <typename T>
class Engine : public IEngine {
 public:
  explicit Engine(const snn_config& c) : cfg_(c) {}
  ~Engine() override { release(); }

  int init(std::string& err) override {
    SNN_CUDA_OK(cudaSetDevice(cfg_.device));
    SNN_CUDA_OK(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking));
    SNN_CUDA_OK(cudaStreamCreateWithFlags(&side_stream_, cudaStreamNonBlocking));
    stream_ = own_stream_;
    SNN_CUDA_OK(cudaEventCreate(&ev_begin_));
    SNN_CUDA_OK(cudaEventCreate(&ev_end_));
    for (auto& e : dep_ev_) SNN_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    SNN_CUDA_OK(cudaStreamCreateWithFlags(&side2_stream_, cudaStreamNonBlocking));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg_.device);
    n_sm_ = sms;
    use_graph_ = cfg_.reserved[0] == 0;
    DevNet<T>& p = p_;
    memset(&p, 0, sizeof(p));
    p.N = cfg_.n_neurons; p.Nld = (p.N + 127) / 128 * 128; p.Nper = cfg_.n_per_instance; p.Ne = cfg_.n_exc;
    p.Dmax = cfg_.max_delay; p.R = p.Dmax + 2; p.B = (int32_t)(p.N / p.Nper); p.W = (int32_t)(p.Nld / 32);
    const size_t N = (size_t)p.Nld;  // padded: vector accesses never leave the allocation
    SNN_CUDA_OK(cudaMalloc(&p.v, N * sizeof(T)));
    SNN_CUDA_OK(cudaMalloc(&p.u, N * sizeof(T)));
    SNN_CUDA_OK(cudaMalloc(&p.ho, N * sizeof(T)));
    SNN_CUDA_OK(cudaMalloc(&p.Iacc, 2 * N * sizeof(T)));
    SNN_CUDA_OK(cudaMalloc(&p.trace, (size_t)p.R * N * sizeof(T)));
    SNN_CUDA_OK(cudaMalloc(&p.bits, (size_t)p.R * p.W * sizeof(uint32_t)));
    SNN_CUDA_OK(cudaMalloc(&p.spk_list, (size_t)p.R * N * sizeof(int32_t)));
    SNN_CUDA_OK(cudaMalloc(&p.spk_cnt, (size_t)p.R * sizeof(int32_t)));
    SNN_CUDA_OK(cudaMalloc(&p.da, 2 * (size_t)p.B * sizeof(T)));
    SNN_CUDA_OK(cudaMalloc(&p.counters, 4 * sizeof(unsigned long long)));
    SNN_CUDA_OK(cudaMalloc(&p.blk_ev, kMaxSynBlocks * sizeof(unsigned long long)));
    SNN_CUDA_OK(cudaMalloc(&p.blk_ltp, kMaxSynBlocks * sizeof(unsigned long long)));
    SNN_CUDA_OK(cudaMalloc(&d_ctx_, sizeof(WindowCtx<T>)));
    SNN_CUDA_OK(cudaMallocHost(&h_ctx_, sizeof(WindowCtx<T>)));
    SNN_CUDA_OK(cudaMallocHost(&h_blk_, 2 * kMaxSynBlocks * sizeof(unsigned long long)));
    SNN_CUDA_OK(cudaMemsetAsync(p.ho, 0, N * sizeof(T), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.Iacc, 0, 2 * N * sizeof(T), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.trace, 0, (size_t)p.R * N * sizeof(T), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.bits, 0, (size_t)p.R * p.W * sizeof(uint32_t), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.spk_cnt, 0, (size_t)p.R * sizeof(int32_t), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.da, 0, 2 * (size_t)p.B * sizeof(T), stream_));
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
    p.exp_flags = cfg_.reserved[1];
    p.part_lo = 0; p.part_hi = p.N;
    p.homeostasis = !(cfg_.theta == 0.0 && cfg_.ttheta == 1.0);
    init_state_kernel<T><<<grid_for(p.Nld), 256, 0, stream_>>>(p);
    SNN_CUDA_OK(cudaGetLastError());
    SNN_CUDA_OK(cudaStreamSynchronize(stream_));
    return SNN_OK;
  }

  int set_topology(Topo&& topo, const double* h_w, std::string& err) override {
    SNN_CUDA_OK(cudaSetDevice(cfg_.device));
    drop_graph();
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
    // lanes per work item from the mean item length
    const double mean_row = p.N ? (double)p.E / (double)p.N : 0.0;
    g_arr_log2_ = pick_group(mean_row / p.Dmax);
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
      case SNN_FIELD_DA: src = p.da + (size_t)(t_ & 1) * p.B; n = p.B; break;
      case SNN_FIELD_I:
        if (!p.exact) { err = "FIELD_I is only kept in SNN_MODE_EXACT"; return SNN_ERR_INVALID; }
        src = p.Iacc + (size_t)((t_ - 1) & 1) * p.Nld; break;
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
      case SNN_FIELD_DA: dst = p.da + (size_t)(t_ & 1) * p.B; n = p.B; break;
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

  int set_partition(int64_t lo, int64_t hi, snn_exchange_fn fn, void* user, std::string& err) override {
    if (lo < 0 || hi > p_.N || lo >= hi || (lo % 128) != 0 || (hi % 128 != 0 && hi != p_.N)) {
      err = "partition bounds must satisfy 0 <= lo < hi <= n_neurons, lo and hi multiples of 128 (hi may be n_neurons)";
      return SNN_ERR_INVALID;
    }
    if (p_.B != 1) { err = "neuron partitioning applies to a single network instance"; return SNN_ERR_INVALID; }
    const bool whole = lo == 0 && hi == p_.N;
    if (!whole && !fn) { err = "a partial partition needs an exchange function"; return SNN_ERR_INVALID; }
    if (t_ != 0) { err = "set the partition before the first step"; return SNN_ERR_STATE; }
    drop_graph();
    p_.part_lo = lo; p_.part_hi = hi;
    exchange_ = whole ? nullptr : fn; exchange_user_ = user;
    return SNN_OK;
  }

  int64_t n_synapses() const override { return p_.E; }
  int64_t step_index() const override { return t_; }
  void set_profiling(int mode) override { profiling_ = mode; }
  void set_stream(cudaStream_t s) override {
    cudaStream_t ns = s ? s : own_stream_;
    if (ns != stream_) drop_graph();
    stream_ = ns;
  }
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
  struct CachedGraph {
    std::string key; cudaGraphExec_t exec = nullptr; size_t prof_used = 0; std::vector<int> prof_cls;
    std::vector<cudaEvent_t> prof_events;  // event record nodes are baked into the graph
  };
  void drop_graph() {
    for (auto& g : graphs_) {
      if (g.exec) cudaGraphExecDestroy(g.exec);
      for (auto& e : g.prof_events) cudaEventDestroy(e);
    }
    graphs_.clear();
  }
  void release() {
    DevNet<T>& p = p_;
    drop_graph();
    cudaFree(p.v); cudaFree(p.u); cudaFree(p.ho); cudaFree(p.Iacc); cudaFree(p.trace); cudaFree(p.bits);
    cudaFree(p.spk_list); cudaFree(p.spk_cnt); cudaFree(p.da); cudaFree(p.counters); cudaFree(p.blk_ev);
    cudaFree(p.blk_ltp); cudaFree(p.wsd); cudaFree(d_ctx_);
    if (h_ctx_) cudaFreeHost(h_ctx_);
    if (h_blk_) cudaFreeHost(h_blk_);
    b_noise_.release(); b_noise64_.release(); b_ev_off_.release(); b_ev_inst_.release(); b_ev_amt_.release();
    b_force_.release(); b_rec_.release(); b_rec_off_.release(); b_cur_idx_.release(); b_cur_val_.release();
    topo_.release();
    if (ev_begin_) cudaEventDestroy(ev_begin_);
    if (ev_end_) cudaEventDestroy(ev_end_);
    for (auto& e : dep_ev_) if (e) cudaEventDestroy(e);
    for (auto& e : prof_events_) cudaEventDestroy(e);
    prof_events_.clear();
    if (own_stream_) cudaStreamDestroy(own_stream_);
    if (side_stream_) cudaStreamDestroy(side_stream_);
    if (side2_stream_) cudaStreamDestroy(side2_stream_);
    side2_stream_ = nullptr;
    memset(&p, 0, sizeof(p));
    own_stream_ = side_stream_ = nullptr; ev_begin_ = ev_end_ = nullptr; d_ctx_ = nullptr; h_ctx_ = nullptr; h_blk_ = nullptr;
  }

  // event pairs around launches, tagged by kernel class (0 neuron, 1 synapse, 2 learn)
  void prof_reserve(size_t n_events) {
    if (n_events > prof_events_.size()) {
      size_t old = prof_events_.size();
      prof_events_.resize(n_events);
      for (size_t q = old; q < prof_events_.size(); ++q) cudaEventCreate(&prof_events_[q]);
    }
  }
  // `captured`: inside stream capture an event only becomes a real record node with the
  // External flag (otherwise it is just a dependency edge and carries no timestamp)
  void prof_begin(int cls, cudaStream_t s, bool captured = false) {
    prof_reserve(prof_used_ + 2);
    if (captured) cudaEventRecordWithFlags(prof_events_[prof_used_], s, cudaEventRecordExternal);
    else cudaEventRecord(prof_events_[prof_used_], s);
    prof_cls_.push_back(cls);
  }
  void prof_end(cudaStream_t s, bool captured = false) {
    if (captured) cudaEventRecordWithFlags(prof_events_[prof_used_ + 1], s, cudaEventRecordExternal);
    else cudaEventRecord(prof_events_[prof_used_ + 1], s);
    prof_used_ += 2;
  }

  struct Launch {
    int gN, gA, gP, gR; dim3 gL;
  };
  Launch grids() {
    DevNet<T>& p = p_;
    Launch l;
    const int64_t local_quads = ((p.part_hi >= p.N ? p.Nld : p.part_hi) >> 2) - (p.part_lo >> 2);
    l.gN = (int)std::max<int64_t>(1, std::min<int64_t>((local_quads + 255) / 256, (int64_t)n_sm_ * 8));
    l.gR = (int)std::max<int64_t>(1, std::min<int64_t>((p.Nld / 4 - local_quads + 255) / 256, (int64_t)n_sm_ * 8));
    if (gA_ == 0) {
      int per_sm = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, synapse_kernel<T, true, true, true>, 256, 0);
      gA_ = std::min(kMaxSynBlocks, n_sm_ * std::max(1, per_sm));
    }
    l.gA = gA_;
    l.gP = std::min(kMaxSynBlocks, n_sm_ * 4);
    int64_t plastic_per_inst = p.B ? (p.E / p.B) : 0;
    int gLx = (int)std::max<int64_t>(1, std::min<int64_t>((plastic_per_inst / 2 + 256 * 4 - 1) / (256 * 4),
                                                           std::max<int64_t>(1, (int64_t)n_sm_ * 8 / p.B)));
    l.gL = dim3(gLx, p.B);
    return l;
  }
  bool wants_learn(const StepArgs& a, int k, int* do_learn, int* do_decay) const {
    const bool train = a.train_flags && a.train_flags[k];
    const bool decay_step = cfg_.sd_decay_mode == SNN_SD_DECAY_EVERY_STEP;
    *do_learn = train ? 1 : 0;
    *do_decay = (decay_step || (train && cfg_.sd_decay_mode == SNN_SD_DECAY_ON_LEARN)) ? 1 : 0;
    return (train || decay_step) && p_.E > 0 && p_.Ne > 0;
  }
  void launch_serial(const StepArgs& a, const std::vector<int32_t>& force_off, const int32_t* d_force,
                     const std::vector<int32_t>& cur_off, const int32_t* d_cur_idx, const double* d_cur_val, bool prof);
  int launch_graph(const StepArgs& a, bool prof, std::string& err);

  snn_config cfg_;
  DevNet<T> p_;
  Topo topo_;
  bool have_topo_ = false, use_graph_ = true;
  int profiling_ = 0;
  int64_t t_ = 0;
  int n_sm_ = 148, g_arr_log2_ = 5, g_ltp_log2_ = 5, gA_ = 0;
  cudaStream_t own_stream_ = nullptr, side_stream_ = nullptr, stream_ = nullptr;
  cudaEvent_t ev_begin_ = nullptr, ev_end_ = nullptr;
  cudaEvent_t dep_ev_[12] = {};
  cudaStream_t side2_stream_ = nullptr;
  std::vector<cudaEvent_t> prof_events_;
  std::vector<int> prof_cls_;
  size_t prof_used_ = 0;
  WindowCtx<T>* d_ctx_ = nullptr; WindowCtx<T>* h_ctx_ = nullptr;
  unsigned long long* h_blk_ = nullptr;
  DevBuf b_noise_, b_noise64_, b_ev_off_, b_ev_inst_, b_ev_amt_, b_force_, b_rec_, b_rec_off_, b_cur_idx_, b_cur_val_;
  snn_exchange_fn exchange_ = nullptr; void* exchange_user_ = nullptr; int exchange_failed_ = 0;
  const std::vector<cudaEvent_t>* graph_prof_events_ = nullptr;
  std::vector<CachedGraph> graphs_;  // a few (length, train pattern, profiling) variants
};

// Plain stream-ordered launches: exact mode, forced spikes, per-kernel profiling.
template <typename T>
void Engine<T>::launch_serial(const StepArgs& a, const std::vector<int32_t>& force_off, const int32_t* d_force,
                              const std::vector<int32_t>& cur_off, const int32_t* d_cur_idx, const double* d_cur_val,
                              bool prof) {
  DevNet<T>& p = p_;
  const int n = a.n_steps;
  const Launch g = grids();
  cudaStream_t s = stream_;
  // production mode: arrivals of window step 0 that are already in flight (spikes of earlier windows)
  if (!p.exact) synapse_kernel<T, false, true, true><<<g.gA, 256, 0, s>>>(p, d_ctx_, -1, g_arr_log2_);
  for (int k = 0; k <= n; ++k) {
    const int nf = k < n ? force_off[k + 1] - force_off[k] : 0, nc = k < n ? cur_off[k + 1] - cur_off[k] : 0;
    const bool forced = nf > 0 || nc > 0;
    auto inputs = [&]() {  // input! of this step: forced spikes, then injected currents
      if (nf > 0) force_kernel<T><<<(nf + 255) / 256, 256, 0, s>>>(p, d_force + force_off[k], nf);
      if (nc > 0) current_kernel<T><<<(nc + 255) / 256, 256, 0, s>>>(p, d_cur_idx + cur_off[k], d_cur_val + cur_off[k], nc);
    };
    if (prof) prof_begin(0, s);
    if (k == 0) {
      if (forced) inputs();
      neuron_kernel<T, false, true><<<g.gN, 256, 0, s>>>(p, d_ctx_, k, 1);
    } else if (k == n) {
      neuron_kernel<T, true, false><<<g.gN, 256, 0, s>>>(p, d_ctx_, k, 1);
    } else if (forced) {
      // events of step k-1 must enter before the decay of step k: fold them into da[t&1] in
      // place (no learn kernel is in flight on this serial path), then detect with decay only
      neuron_kernel<T, true, false><<<g.gN, 256, 0, s>>>(p, d_ctx_, k, 1);
      inputs();
      neuron_kernel<T, false, true><<<g.gN, 256, 0, s>>>(p, d_ctx_, k, 0);
    } else {
      neuron_kernel<T, true, true><<<g.gN, 256, 0, s>>>(p, d_ctx_, k, 1);
    }
    if (k < n && exchange_) {
      // neuron-partitioned run: all-gather this step's spike bitmap, then replay detection for
      // the neurons owned by the other ranks (traces, spike list, delay-1 pushes)
      uint32_t* slot_bits = p.bits + (size_t)slot_host(t_ + k) * p.W;
      const int rc = exchange_(exchange_user_, slot_bits, p.part_lo / 32, (p.part_hi - p.part_lo + 31) / 32, p.W, (void*)s);
      if (rc != 0) { exchange_failed_ = rc; return; }
      remote_kernel<T><<<g.gR, 256, 0, s>>>(p, d_ctx_, k);
    }
    if (prof) prof_end(s);
    if (k == n) break;
    int dl, dd;
    const bool learn = wants_learn(a, k, &dl, &dd);
    const bool push_next = !p.exact && k + 1 < n;  // the last step leaves its pushes to the next window
    if (prof) prof_begin(1, s);
    if (p.exact) {
      gather_kernel<T><<<grid_for(p.N), 256, 0, s>>>(p, d_ctx_, k);
      synapse_kernel<T, true, false, false><<<g.gA, 256, 0, s>>>(p, d_ctx_, k, g_arr_log2_);
      ltp_kernel<T, false><<<g.gP, 256, 0, s>>>(p, d_ctx_, k, g_ltp_log2_);
    } else {
      if (push_next && !learn) synapse_kernel<T, true, true, true><<<g.gA, 256, 0, s>>>(p, d_ctx_, k, g_arr_log2_);
      else synapse_kernel<T, true, false, true><<<g.gA, 256, 0, s>>>(p, d_ctx_, k, g_arr_log2_);
      ltp_kernel<T, true><<<g.gP, 256, 0, s>>>(p, d_ctx_, k, g_ltp_log2_);
    }
    if (prof) prof_end(s);
    if (learn) {
      if (prof) prof_begin(2, s);
      learn_kernel<T><<<g.gL, 256, 0, s>>>(p, d_ctx_, k, dl, dd);
      if (prof) prof_end(s);
      // the pushes of step k+1 must see the weights after learn!(k)
      if (push_next) synapse_kernel<T, false, true, true><<<g.gA, 256, 0, s>>>(p, d_ctx_, k, g_arr_log2_);
    }
  }
}

// Whole window as one CUDA graph (captured once per (length, train pattern)), three streams.
//   needs(neuron(k))  : neuron(k-1) [stream], pushes for step k-1 = synapse(k-2) | push(k-2),
//                       learn(k-1) (its delay-1 pushes read w)
//   needs(synapse(k)) : neuron(k), learn(k-1), synapse(k-1) [stream]   (sd updates are reductions:
//   needs(ltp(k))     : neuron(k), learn(k-1) [stream], ltp(k-1) [stream]  no LTD/LTP ordering)
//   needs(learn(k))   : synapse(k) LTD, ltp(k) [stream]
//   needs(push(k))    : learn(k)            (learn steps only; otherwise fused into synapse(k))
template <typename T>
int Engine<T>::launch_graph(const StepArgs& a, bool prof, std::string& err) {
  DevNet<T>& p = p_;
  const int n = a.n_steps;
  std::string key(reinterpret_cast<const char*>(&n), sizeof(n));
  key.push_back(prof ? 'p' : '-');
  if (a.train_flags) key.append(reinterpret_cast<const char*>(a.train_flags), n);
  CachedGraph* hit = nullptr;
  for (auto& cg : graphs_) if (cg.key == key) hit = &cg;
  if (!hit) {
    if (graphs_.size() >= 6) {  // evict the oldest
      if (graphs_.front().exec) cudaGraphExecDestroy(graphs_.front().exec);
      for (auto& e : graphs_.front().prof_events) cudaEventDestroy(e);
      graphs_.erase(graphs_.begin());
    }
    // profiling events of a captured graph are baked into it: give this graph its own set
    std::vector<cudaEvent_t> saved_events;
    saved_events.swap(prof_events_);
    const Launch g = grids();
    cudaStream_t s1 = stream_, s2 = side_stream_, s3 = side2_stream_;
    cudaGraph_t graph = nullptr;
    prof_used_ = 0; prof_cls_.clear();
    if (prof) prof_reserve(2 * (size_t)n + 2);  // no allocation while capturing
    // events: [0] neuron done, [1..3] pushes-for-step ready (ring of 3), [4..5] synapse LTD done
    // (ring of 2), [6..7] ltp done (ring of 2), [8] learn done
    cudaEvent_t evN = dep_ev_[0], evL = dep_ev_[8];
    auto evPush = [&](int step) { return dep_ev_[1 + ((step % 3) + 3) % 3]; };  // pushes consumed by Euler(step)
    auto evS = [&](int step) { return dep_ev_[4 + (step & 1)]; };
    auto evP = [&](int step) { return dep_ev_[6 + (step & 1)]; };
    SNN_CUDA_OK(cudaStreamBeginCapture(s1, cudaStreamCaptureModeThreadLocal));
    cudaEventRecord(evN, s1);                      // fork point
    cudaStreamWaitEvent(s2, evN, 0);
    synapse_kernel<T, false, true, true><<<g.gA, 256, 0, s2>>>(p, d_ctx_, -1, g_arr_log2_);  // pushes for step 0
    cudaEventRecord(evPush(0), s2);
    bool learn_prev = false, s3_used = false;
    for (int k = 0; k < n; ++k) {
      int dl, dd;
      const bool learn = wants_learn(a, k, &dl, &dd);
      const bool push_next = k + 1 < n;
      if (k >= 1) cudaStreamWaitEvent(s1, evPush(k - 1), 0);  // Euler(k-1) consumes them
      if (learn_prev) cudaStreamWaitEvent(s1, evL, 0);
      if (k == 0) neuron_kernel<T, false, true><<<g.gN, 256, 0, s1>>>(p, d_ctx_, k, 1);
      else neuron_kernel<T, true, true><<<g.gN, 256, 0, s1>>>(p, d_ctx_, k, 1);
      cudaEventRecord(evN, s1);
      // s2: synapse(k)
      cudaStreamWaitEvent(s2, evN, 0);
      if (learn_prev) cudaStreamWaitEvent(s2, evL, 0);  // sd reductions must not race the learn pass
      if (push_next && !learn) {
        synapse_kernel<T, true, true, true><<<g.gA, 256, 0, s2>>>(p, d_ctx_, k, g_arr_log2_);
        cudaEventRecord(evPush(k + 1), s2);
      } else {
        synapse_kernel<T, true, false, true><<<g.gA, 256, 0, s2>>>(p, d_ctx_, k, g_arr_log2_);
      }
      cudaEventRecord(evS(k), s2);
      // s3: ltp(k) [-> learn(k)]
      cudaStreamWaitEvent(s3, evN, 0);
      ltp_kernel<T, true><<<g.gP, 256, 0, s3>>>(p, d_ctx_, k, g_ltp_log2_);
      s3_used = true;
      if (learn) {
        cudaStreamWaitEvent(s3, evS(k), 0);
        if (prof) prof_begin(2, s3, true);
        learn_kernel<T><<<g.gL, 256, 0, s3>>>(p, d_ctx_, k, dl, dd);
        if (prof) prof_end(s3, true);
        cudaEventRecord(evL, s3);
        if (push_next) {
          cudaStreamWaitEvent(s2, evL, 0);
          synapse_kernel<T, false, true, true><<<g.gA, 256, 0, s2>>>(p, d_ctx_, k, g_arr_log2_);
          cudaEventRecord(evPush(k + 1), s2);
        }
      }
      cudaEventRecord(evP(k), s3);
      learn_prev = learn;
    }
    // join everything, then the window-end Euler
    cudaEventRecord(evS(n), s2);
    cudaStreamWaitEvent(s1, evS(n), 0);
    if (s3_used) { cudaEventRecord(evP(n), s3); cudaStreamWaitEvent(s1, evP(n), 0); }
    neuron_kernel<T, true, false><<<g.gN, 256, 0, s1>>>(p, d_ctx_, n, 1);
    cudaError_t e = cudaStreamEndCapture(s1, &graph);
    if (e != cudaSuccess || !graph) {
      for (auto& ev : prof_events_) cudaEventDestroy(ev);
      prof_events_.clear();
      prof_events_.swap(saved_events);
      err = std::string("graph capture: ") + cudaGetErrorString(e); return SNN_ERR_CUDA;
    }
    CachedGraph cg;
    cg.key = key; cg.prof_used = prof_used_; cg.prof_cls = prof_cls_;
    cg.prof_events.swap(prof_events_);
    prof_events_.swap(saved_events);
    e = cudaGraphInstantiate(&cg.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
      for (auto& ev : cg.prof_events) cudaEventDestroy(ev);
      err = std::string("graph instantiate: ") + cudaGetErrorString(e); return SNN_ERR_CUDA;
    }
    graphs_.push_back(std::move(cg));
    hit = &graphs_.back();
  }
  prof_used_ = hit->prof_used;
  prof_cls_ = hit->prof_cls;
  graph_prof_events_ = &hit->prof_events;
  SNN_CUDA_OK(cudaGraphLaunch(hit->exec, stream_));
  return SNN_OK;
}

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

  // ---- validate + stage the window's inputs (persistent grow-only buffers)
  std::vector<int32_t> ev_off, ev_inst, force_off(n + 1, 0), force_ids;
  std::vector<double> ev_amt;
  if (a.n_da > 0) {
    ev_off.assign(n + 1, 0); ev_inst.resize(a.n_da); ev_amt.resize(a.n_da);
    int prev = -1;
    for (int q = 0; q < a.n_da; ++q) {
      const snn_da_event& e = a.da[q];
      if (e.step < 0 || e.step >= n || e.step < prev || e.instance < 0 || e.instance >= p.B) {
        err = "da events must be sorted by step, inside the window, instance in range"; return SNN_ERR_INVALID;
      }
      prev = e.step; ev_off[e.step + 1]++; ev_inst[q] = e.instance; ev_amt[q] = e.amount;
    }
    for (int k = 0; k < n; ++k) ev_off[k + 1] += ev_off[k];
  }
  if (a.n_force > 0) {
    force_ids.resize(a.n_force);
    int prev = -1;
    for (int q = 0; q < a.n_force; ++q) {
      const snn_force_event& e = a.force[q];
      if (e.step < 0 || e.step >= n || e.step < prev || e.neuron < 0 || e.neuron >= N) {
        err = "force events must be sorted by step, inside the window, neuron in range"; return SNN_ERR_INVALID;
      }
      prev = e.step; force_off[e.step + 1]++; force_ids[q] = e.neuron;
    }
    for (int k = 0; k < n; ++k) force_off[k + 1] += force_off[k];
  }
  std::vector<int32_t> cur_off(n + 1, 0), cur_ids;
  std::vector<double> cur_val;
  if (a.n_cur > 0) {
    cur_ids.resize(a.n_cur); cur_val.resize(a.n_cur);
    int prev = -1;
    for (int q = 0; q < a.n_cur; ++q) {
      const snn_current_event& e = a.cur[q];
      if (e.step < 0 || e.step >= n || e.step < prev || e.neuron < 0 || e.neuron >= N) {
        err = "current events must be sorted by step, inside the window, neuron in range"; return SNN_ERR_INVALID;
      }
      prev = e.step; cur_off[e.step + 1]++; cur_ids[q] = e.neuron; cur_val[q] = e.current;
    }
    for (int k = 0; k < n; ++k) cur_off[k + 1] += cur_off[k];
  }
  SNN_CUDA_OK(cudaEventRecord(ev_begin_, stream_));
  WindowCtx<T>& cx = *h_ctx_;
  memset(&cx, 0, sizeof(cx));
  cx.t0 = t_; cx.slot0 = slot_host(t_); cx.n_steps = n;
  if (cfg_.noise_mode == SNN_NOISE_REPLAY) {
    const size_t cnt = (size_t)n * N;
    SNN_CUDA_OK(b_noise_.reserve(cnt * sizeof(T)));
    if (sizeof(T) == 8) {
      SNN_CUDA_OK(cudaMemcpyAsync(b_noise_.ptr, a.noise, cnt * sizeof(double), cudaMemcpyHostToDevice, stream_));
    } else {
      SNN_CUDA_OK(b_noise64_.reserve(cnt * sizeof(double)));
      SNN_CUDA_OK(cudaMemcpyAsync(b_noise64_.ptr, a.noise, cnt * sizeof(double), cudaMemcpyHostToDevice, stream_));
      from_double_kernel<T><<<grid_for((int64_t)cnt), 256, 0, stream_>>>(b_noise64_.as<double>(), b_noise_.as<T>(), (int64_t)cnt);
    }
    cx.noise = b_noise_.as<T>();
  }
  if (a.n_da > 0) {
    SNN_CUDA_OK(b_ev_off_.reserve((n + 1) * sizeof(int32_t)));
    SNN_CUDA_OK(b_ev_inst_.reserve(a.n_da * sizeof(int32_t)));
    SNN_CUDA_OK(b_ev_amt_.reserve(a.n_da * sizeof(double)));
    SNN_CUDA_OK(cudaMemcpyAsync(b_ev_off_.ptr, ev_off.data(), (n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
    SNN_CUDA_OK(cudaMemcpyAsync(b_ev_inst_.ptr, ev_inst.data(), a.n_da * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
    SNN_CUDA_OK(cudaMemcpyAsync(b_ev_amt_.ptr, ev_amt.data(), a.n_da * sizeof(double), cudaMemcpyHostToDevice, stream_));
    cx.ev_off = b_ev_off_.as<int32_t>(); cx.ev_inst = b_ev_inst_.as<int32_t>(); cx.ev_amt = b_ev_amt_.as<double>();
  }
  if (a.n_force > 0) {
    SNN_CUDA_OK(b_force_.reserve(a.n_force * sizeof(int32_t)));
    SNN_CUDA_OK(cudaMemcpyAsync(b_force_.ptr, force_ids.data(), a.n_force * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
  }
  if (a.n_cur > 0) {
    SNN_CUDA_OK(b_cur_idx_.reserve(a.n_cur * sizeof(int32_t)));
    SNN_CUDA_OK(b_cur_val_.reserve(a.n_cur * sizeof(double)));
    SNN_CUDA_OK(cudaMemcpyAsync(b_cur_idx_.ptr, cur_ids.data(), a.n_cur * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
    SNN_CUDA_OK(cudaMemcpyAsync(b_cur_val_.ptr, cur_val.data(), a.n_cur * sizeof(double), cudaMemcpyHostToDevice, stream_));
  }
  if (a.spikes_out) {
    SNN_CUDA_OK(b_rec_.reserve(a.spikes_cap * sizeof(int32_t)));
    SNN_CUDA_OK(b_rec_off_.reserve((n + 1) * sizeof(int32_t)));
    SNN_CUDA_OK(cudaMemsetAsync(b_rec_off_.ptr, 0, (n + 1) * sizeof(int32_t), stream_));
    cx.rec_ids = b_rec_.as<int32_t>(); cx.rec_off = b_rec_off_.as<int32_t>(); cx.rec_cap = a.spikes_cap;
  }
  SNN_CUDA_OK(cudaMemcpyAsync(d_ctx_, h_ctx_, sizeof(cx), cudaMemcpyHostToDevice, stream_));
  unsigned long long c_before[4], c_after[4];
  SNN_CUDA_OK(cudaMemcpyAsync(c_before, p.counters, sizeof(c_before), cudaMemcpyDeviceToHost, stream_));
  SNN_CUDA_OK(cudaMemsetAsync(p.blk_ev, 0, kMaxSynBlocks * sizeof(unsigned long long), stream_));
  SNN_CUDA_OK(cudaMemsetAsync(p.blk_ltp, 0, kMaxSynBlocks * sizeof(unsigned long long), stream_));

  // ---- the window
  const bool graphable = use_graph_ && !p.exact && a.n_force == 0 && a.n_cur == 0 && profiling_ != 1 && !exchange_;
  if (graphable) {
    int rc = launch_graph(a, profiling_ == 2, err);
    if (rc) return rc;
  } else {
    prof_used_ = 0; prof_cls_.clear();
    exchange_failed_ = 0;
    launch_serial(a, force_off, b_force_.as<int32_t>(), cur_off, b_cur_idx_.as<int32_t>(), b_cur_val_.as<double>(),
                  profiling_ == 1);
    if (exchange_failed_) {
      cudaStreamSynchronize(stream_);
      err = "bitmap exchange callback failed with code " + std::to_string(exchange_failed_);
      return SNN_ERR_STATE;
    }
  }
  SNN_CUDA_OK(cudaGetLastError());
  t_ += n;

  // ---- results
  int rc = SNN_OK;
  std::vector<int32_t> h_off;
  if (a.spikes_out) {
    h_off.resize(n + 1);
    SNN_CUDA_OK(cudaMemcpyAsync(h_off.data(), b_rec_off_.ptr, (n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
    SNN_CUDA_OK(cudaStreamSynchronize(stream_));
    const int64_t total = h_off[n];
    int src = sort_segments(b_rec_.as<int32_t>(), total, b_rec_off_.as<int32_t>(), n, stream_, err);
    if (src) return src;
    if (total > 0) SNN_CUDA_OK(cudaMemcpyAsync(a.spikes_out, b_rec_.ptr, total * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
    for (int k = 0; k <= n; ++k) a.spike_offsets[k] = h_off[k];
  }
  SNN_CUDA_OK(cudaMemcpyAsync(c_after, p.counters, sizeof(c_after), cudaMemcpyDeviceToHost, stream_));
  SNN_CUDA_OK(cudaMemcpyAsync(h_blk_, p.blk_ev, kMaxSynBlocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream_));
  SNN_CUDA_OK(cudaMemcpyAsync(h_blk_ + kMaxSynBlocks, p.blk_ltp, kMaxSynBlocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream_));
  SNN_CUDA_OK(cudaEventRecord(ev_end_, stream_));
  SNN_CUDA_OK(cudaStreamSynchronize(stream_));
  if (a.spikes_out && (int64_t)(c_after[0] - c_before[0]) > a.spikes_cap) {
    err = "spikes_cap too small: record truncated"; rc = SNN_ERR_CAPACITY;
  }
  if (a.stats) {
    snn_step_stats& s = *a.stats;
    memset(&s, 0, sizeof(s));
    s.n_spikes = (int64_t)(c_after[0] - c_before[0]);
    s.n_syn_events = (int64_t)(c_after[1] - c_before[1]);
    for (int b = 0; b < kMaxSynBlocks; ++b) { s.n_syn_events += (int64_t)h_blk_[b]; s.n_ltp_events += (int64_t)h_blk_[kMaxSynBlocks + b]; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev_begin_, ev_end_);
    s.device_ms = ms;
    const std::vector<cudaEvent_t>& pev = (graphable && graph_prof_events_) ? *graph_prof_events_ : prof_events_;
    for (size_t q = 0; q + 1 < pev.size() && q < prof_used_ && q / 2 < prof_cls_.size(); q += 2) {
      float m = 0.f;
      if (cudaEventElapsedTime(&m, pev[q], pev[q + 1]) != cudaSuccess) { cudaGetLastError(); continue; }
      int cls = prof_cls_[q / 2];
      if (cls == 0) { s.neuron_ms += m; s.neuron_launches++; }
      else if (cls == 1) { s.synapse_ms += m; s.synapse_launches++; }
      else { s.learn_ms += m; s.learn_launches++; }
    }
    if (graphable) {  // launches inside the graph (events exist only around learn, if at all)
      s.neuron_launches = n + 1; s.synapse_launches = 2 * n + 1;
      if (profiling_ != 2) {
        int nl = 0, dl, dd;
        for (int k = 0; k < n; ++k) nl += wants_learn(a, k, &dl, &dd) ? 1 : 0;
        s.learn_launches = nl;
      }
    } else if (profiling_ != 1) {
      s.neuron_launches = n + 1; s.synapse_launches = (p.exact ? 3 : 2) * n;
      int nl = 0, dl, dd;
      for (int k = 0; k < n; ++k) nl += wants_learn(a, k, &dl, &dd) ? 1 : 0;
      s.learn_launches = nl;
    }
  }
  return rc;
}

}  // namespace snn
