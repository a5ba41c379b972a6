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
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <unistd.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "snn_common.h"

#ifndef NEURON_MIN_BLOCKS
#define NEURON_MIN_BLOCKS 4
#endif

namespace snn {

template <typename T> struct LogEntry { uint32_t e; T d; };
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
constexpr int kMaxRanks = 8;

// State of the asynchronous learn pass.  `state` packs (cur << 31) | frontier, frontier in units
// of 32 synapses: synapses with (e >> 5) < frontier have been processed by the pass in flight
// (their post-learn (w, sd) is in wsdX[cur]); the others are still only valid in wsdX[cur^1],
// which stays frozen until the pass and its log replay are over.  kFrontAll = no pass in flight.
struct LearnCtl {
  uint32_t state;
  uint32_t epoch;      // pass counter: tile_done[t] == epoch <=> tile t of the pass in flight is written
  uint32_t next_tile;  // tile dispenser of the persistent pass kernel
  uint32_t spilled;    // some sd update went to the dense spill (a log region was full)
  uint32_t ticket;     // block counter of learn_replay_kernel (its last block begins the next pass)
};
constexpr uint32_t kFrontAll = 0x7fffffffu;
constexpr uint32_t kTileSyn = 32768;   // default synapses per tile of the learn pass (multiple of 1024)
constexpr uint32_t kLogStride = 32;    // log counters are 128 B apart (per-block addresses, no contention)

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
  T2* wsd;               // = wsdX[0]; [E] (w, sd) interleaved: one 8-byte access serves propagate + LTD
  // asynchronous learn pass (production mode): the pass reads wsdX[cur^1] (frozen) and writes
  // wsdX[cur] while the following steps keep running; see LearnCtl
  T2* wsdX[2];
  LearnCtl* lc;          // device-side pass state
  T* lg;                 // [B] learn rate (base+da | da | eta*da) of the pass in flight
  // sd updates that hit not-yet-processed synapses while a pass is in flight: one private log
  // region per (kernel, block) — a single shared append counter serialises in L2 at ~1 atomic/clk
  LogEntry<T>* dlog;     // [log_regions][log_region_cap]
  uint32_t* log_cnt;     // [log_regions * kLogStride]
  T* dsd;                // [E] dense spill of the same when a region is full (never in practice)
  uint32_t log_regions, log_region_cap, log_ltp_base;
  uint32_t* tile_done;   // [n_tiles]
  const uint32_t *inst_lo, *inst_hi;  // [B] plastic synapse range of every instance
  uint32_t n_tiles, tile_syn;
  int32_t static_cur;    // >= 0: no pass can be in flight while this launch runs and wsdX[static_cur] is
                         // current — kernels skip the LearnCtl load (in-place windows, exact mode)
  // event-driven ("lazy") learning, production mode, opt-in: between two
  // re-basing passes wsd holds the two accumulators (sd' = sd / r^k, C = sum_events d' * H[k_e]) —
  // adjacent, so that one 8-byte vector reduction serves an STDP event — and cacc the base weight s0
  // of EVERY synapse; the weight is s = clamp(s0 + H*sd' - C) with H = sum_j rate_j * r^(j-1) over the
  // learn steps since the last re-basing; learn! touches no synapse except every `lazy`-th time (or
  // while da > lazy_da), when a dense pass writes s0 = s, sd' = sd, C = 0.  The API layout (w, sd) is
  // restored around snn_get_array / snn_set_array.
  int32_t lazy;          // 0 = off, else M = learn steps per re-basing pass
  T lazy_da;
  T* cacc;               // [E] s0
  T *lzH, *lzR, *lzInv;  // [B] H, r^k, 1/r^k
  int32_t* lzK;          // [B] learn steps since the last re-basing
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
  int32_t ho_inh, variant_g;  // see snn_config
  int32_t exp_flags;  // snn_config.reserved[1]: 2 = kernel timeline, 4 = nothing overlaps the learn pass, 8 = 256-thread neuron blocks, 512 = LSU pass
  // neuron partition (multi-GPU): this rank integrates neurons [part_lo, part_hi)
  // and holds every synapse whose POST is in that range; spike bitmaps are exchanged per step
  int64_t part_lo, part_hi;
  // peer-to-peer spike exchange (snn_partition_connect): right after detection every warp stores
  // its 4 bitmap words straight into every peer's exchange ring over NVLink / NVSwitch.  Each word
  // travels in one 8-byte store together with the step it belongs to (word, step+1) — the
  // low-latency flag-in-data protocol: no fence, no separate flag, no end-of-kernel rendezvous on
  // the sending side (a __threadfence_system after peer stores cost ~12 us per step); the receiver
  // (remote_kernel) spins on the tag of exactly the words it needs.
  uint2* ll;                  // [R][W] own exchange ring, written by the peers
  uint2* peer_ll[kMaxRanks];  // the peers' rings (CUDA-IPC mapped / same-process pointers)
  uint32_t* xdone;            // [1] = a wait timed out
  int32_t n_ranks, rank;
  // diagnostics (exp_flags & 2): per-node first-block-start / last-block-end %globaltimer stamps
  unsigned long long* tl;
};

// timeline classes: 0 neuron, 1 synapse, 2 ltp, 3 learn, 4 remote, 5 other
constexpr int kTlSteps = 258, kTlClasses = 6;
__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void tl_begin(unsigned long long* tl, int cls, int k) {
  if (tl && threadIdx.x == 0 && k + 1 >= 0 && k + 1 < kTlSteps) atomicMin(&tl[cls * kTlSteps + k + 1], gtime_ns());
}
__device__ __forceinline__ void tl_end(unsigned long long* tl, int cls, int k) {
  if (tl && threadIdx.x == 0 && k + 1 >= 0 && k + 1 < kTlSteps)
    atomicMax(&tl[kTlClasses * kTlSteps + cls * kTlSteps + k + 1], gtime_ns());
}

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
// View of the (w, sd) storage while a learn pass may be in flight.  Read ONCE per kernel: a stale
// (smaller) frontier is always safe — the synapse is then simply treated as not yet processed.
// All accesses go to L2 (ld.cg / red): the L1 of this SM may not see another kernel's writes.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct WView {
  typename Real2<T>::type* cur;        // post-learn values of processed synapses; all non-plastic ones
  const typename Real2<T>::type* old;  // frozen pre-learn values (valid for every synapse)
  uint32_t front;                      // in units of 32 synapses
  bool lazy;                           // event-driven learning: cur holds (s0, sd'), see DevNet::lazy
};
template <typename T>
__device__ __forceinline__ WView<T> wview(const DevNet<T>& p) {
  WView<T> v;
  v.lazy = p.lazy != 0;
  if (p.static_cur >= 0) {
    v.cur = p.wsdX[p.static_cur]; v.old = p.wsdX[p.static_cur ^ 1]; v.front = kFrontAll;
    return v;
  }
  const uint32_t s = *reinterpret_cast<const volatile uint32_t*>(&p.lc->state);
  v.cur = p.wsdX[s >> 31];
  v.old = p.wsdX[(s >> 31) ^ 1u];
  v.front = s & kFrontAll;
  return v;
}
__device__ __forceinline__ float ldcg_x(const float2* q) { return __ldcg(reinterpret_cast<const float*>(q)); }
__device__ __forceinline__ double ldcg_x(const double2* q) { return __ldcg(reinterpret_cast<const double*>(q)); }
// weight of synapse e as learn! leaves it: already written by the pass, or formed on the fly from
// the frozen (w, sd) with the pass's rate (new_net.jl:137-138).  Two phases so that callers can
// keep several independent loads in flight: w_fetch issues the load, w_resolve finishes.
// `row_exc`: the presynaptic neuron is excitatory (only those rows can be plastic).
template <typename T>
__device__ __forceinline__ typename Real2<T>::type w_fetch(const WView<T>& vw, uint32_t e, bool row_exc, bool& fly) {
  fly = row_exc && (vw.lazy || (e >> 5) >= vw.front);
  typename Real2<T>::type ws;
  if (vw.lazy) {  // (sd', C) of a possibly plastic synapse; the base weight comes from c_fetch
    if (fly) ws = __ldcg(vw.cur + e);
    else { ws.x = (T)0; ws.y = (T)0; }
  } else if (fly) ws = __ldcg(vw.old + e);
  else { ws.x = ldcg_x(vw.cur + e); ws.y = (T)0; }
  return ws;
}
// event-driven learning: the base weight s0 (issued together with w_fetch)
template <typename T>
__device__ __forceinline__ T c_fetch(const DevNet<T>& p, const WView<T>& vw, uint32_t e, bool fly) {
  return vw.lazy ? __ldcg(p.cacc + e) : (T)0;
}
template <typename T>
__device__ __forceinline__ T w_resolve(const DevNet<T>& p, const typename Real2<T>::type& ws, bool fly, bool plastic,
                                       uint32_t inst, T c) {
  if (!(fly && plastic)) return p.lazy ? c : ws.x;  // non-plastic weights never change: both buffers hold them
  const T x = p.lazy ? Ops<T>::sub(Ops<T>::add(c, Ops<T>::mul(p.lzH[inst], ws.x)), ws.y)
                     : Ops<T>::add(ws.x, Ops<T>::mul(p.lg[inst], ws.y));
  return x > p.wmax ? p.wmax : (x < p.wmin ? p.wmin : x);
}
// sd[e] += delta (production mode: unordered reductions).  A synapse the pass in flight has not
// reached keeps its frozen sd: the update is logged (region = this block's private log) and
// replayed onto the post-learn value.
template <typename T>
__device__ __forceinline__ void sd_add(const DevNet<T>& p, const WView<T>& vw, uint32_t e, T delta, uint32_t region,
                                       uint32_t inst) {
  if (vw.lazy) {  // sd' += d', C += d' * H: one fire-and-forget 8-byte vector reduction in fp32
    const T d = Ops<T>::mul(delta, p.lzInv[inst]);
    const T dh = Ops<T>::mul(d, p.lzH[inst]);
    if constexpr (sizeof(T) == 4) {
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(vw.cur + e), "f"(d), "f"(dh) : "memory");
    } else {
      atomicAdd(&vw.cur[e].x, d);
      atomicAdd(&vw.cur[e].y, dh);
    }
    return;
  }
  if ((e >> 5) < vw.front) {
    atomicAdd(&vw.cur[e].y, delta);
    return;
  }
  namespace cg = cooperative_groups;
  cg::coalesced_group g = cg::coalesced_threads();
  uint32_t base = 0;
  if (g.thread_rank() == 0) base = atomicAdd(&p.log_cnt[(size_t)region * kLogStride], (uint32_t)g.size());
  base = g.shfl(base, 0);
  const uint32_t pos = base + g.thread_rank();
  if (pos < p.log_region_cap) {
    LogEntry<T> le; le.e = e; le.d = delta;
    p.dlog[(size_t)region * p.log_region_cap + pos] = le;
  } else {
    atomicAdd(&p.dsd[e], delta);
    p.lc->spilled = 1u;
  }
}

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
// `bid` / `nblk`: this block's index and the block count of the neuron part (the partitioned runs
// launch it fused with remote_body in one grid)
template <typename T, bool EULER, bool DETECT>
__device__ __forceinline__ void neuron_body(const DevNet<T>& p, const WindowCtx<T>* __restrict__ cx, int32_t k, int32_t apply_ev,
                                            const uint32_t bid, const uint32_t nblk) {
  typedef Ops<T> O;
  tl_begin(p.tl, 0, k);
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
  const int n_prev = (k > 0 && (rec_ids || bid == 0)) ? p.spk_cnt[prev] : 0;
  if (rec_ids && k > 0) {
    int64_t nxt = (int64_t)cx->rec_off[k - 1] + n_prev;
    rec_base = (int32_t)(nxt < rec_cap ? nxt : rec_cap);
  }
  T* __restrict__ I_prev = p.Iacc + (size_t)((t - 1) & 1) * p.Nld;  // consumed by Euler(t-1)
  T* __restrict__ I_now = p.Iacc + (size_t)(t & 1) * p.Nld;         // receives the pushes of step t
  const T* __restrict__ noise_prev = (EULER && cx->noise) ? cx->noise + (size_t)(k - 1) * p.N : nullptr;
  const uint32_t N = (uint32_t)p.N, Nper = (uint32_t)p.Nper, Ne = (uint32_t)p.Ne;
  WView<T> vw;
  if (DETECT) vw = wview(p);

  // dopamine (new_net.jl:131 da *= 0.995 ; newnet_DASTDP.jl:54-56 da += 0.5 after step!)
  if (bid == 0) {
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
  for (uint32_t q = q_lo + bid * blockDim.x + threadIdx.x; q < q_hi; q += nblk * blockDim.x) {
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
          if (s && (exc[c] || p.ho_inh) && (int64_t)l >= p.ho_from) h = O::add(h, p.theta);
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
      if (p.n_ranks > 1) {
        // fused exchange: a warp owns 4 consecutive bitmap words = two 16-byte stores of
        // (word, step tag) pairs into every peer's ring over NVLink
        const unsigned w1 = __shfl_sync(0xffffffffu, word, 8), w2 = __shfl_sync(0xffffffffu, word, 16),
                       w3 = __shfl_sync(0xffffffffu, word, 24);
        if (lane < p.n_ranks && lane != p.rank) {
          const uint32_t tag = (uint32_t)(t + 1);
          uint4* dst = reinterpret_cast<uint4*>(p.peer_ll[lane] + (size_t)slot * p.W + (j0 >> 5));
          dst[0] = make_uint4(word, tag, w1, tag);
          dst[1] = make_uint4(w2, tag, w3, tag);
        }
      }
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
              // delay-1 synapses of this step's spikers feed the Euler update of this very step
              // (new_net.jl:116-121): push them now, warp-cooperatively, into Iacc[t&1]
              unsigned mm = m;
              while (mm) {
                const int src = __ffs(mm) - 1; mm &= mm - 1;
                const uint32_t i = __shfl_sync(0xffffffffu, j0, src) + c;
                const uint32_t s0 = p.seg[(size_t)i * (p.Dmax + 1)], s1 = p.seg[(size_t)i * (p.Dmax + 1) + 1];
                const uint32_t inst = p.B == 1 ? 0u : i / Nper, ibase = inst * Nper;
                const bool row_exc = (i - ibase) < Ne;
                // four loads in flight per lane before the first reduction (an inhibitory row has
                // ~100 delay-1 synapses; one round trip instead of four)
                for (uint32_t e0 = s0 + lane; e0 < s1; e0 += 128) {
                  int32_t pj[4]; typename DevNet<T>::T2 pw[4]; bool fly[4]; T pc[4];
#pragma unroll
                  for (int q4 = 0; q4 < 4; ++q4) {
                    const uint32_t e = e0 + 32 * q4;
                    pj[q4] = e < s1 ? p.post[e] : -1;
                    fly[q4] = false; pc[q4] = (T)0;
                    if (e < s1) { pw[q4] = w_fetch<T>(vw, e, row_exc, fly[q4]); pc[q4] = c_fetch<T>(p, vw, e, fly[q4]); }
                  }
#pragma unroll
                  for (int q4 = 0; q4 < 4; ++q4)
                    if (pj[q4] >= 0)
                      atomicAdd(&I_now[pj[q4]],
                                w_resolve<T>(p, pw[q4], fly[q4], p.plastic_ei || ((uint32_t)pj[q4] - ibase) < Ne, inst, pc[q4]));
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
  if (p.tl) __syncthreads();  // timeline: stamp when the block's last warp (spike pushes) is done
  tl_end(p.tl, 0, k);
}

template <typename T, bool EULER, bool DETECT>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? NEURON_MIN_BLOCKS : 2)
neuron_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k, int32_t apply_ev) {
  neuron_body<T, EULER, DETECT>(p, cx, k, apply_ev, blockIdx.x, gridDim.x);
}

// The bitmap word `wi` of ring slot `slot` as sent by its owner for the step tagged `tag`: spins on
// the (word, tag) pair until the tag matches.  Gives up after ~4 s (flag in xdone[1]) instead of
// hanging the device.
template <typename T>
__device__ __forceinline__ uint32_t ll_word(const DevNet<T>& p, int slot, uint32_t wi, uint32_t tag) {
  const uint2* q = p.ll + (size_t)slot * p.W + wi;
  uint32_t w, g;
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(w), "=r"(g) : "l"(q) : "memory");
  if (g == tag) return w;
  if (*reinterpret_cast<volatile uint32_t*>(&p.xdone[1]) != 0u) return 0u;
  const unsigned long long t_start = gtime_ns();
  for (;;) {
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(w), "=r"(g) : "l"(q) : "memory");
    if (g == tag) return w;
    if (gtime_ns() - t_start > 4000000000ull) { p.xdone[1] = 1u; return 0u; }
    __nanosleep(64);
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
__device__ __forceinline__ void remote_body(const DevNet<T>& p, const WindowCtx<T>* __restrict__ cx, int32_t k, int32_t wait_peers,
                                            const uint32_t bid, const uint32_t nblk) {
  typedef Ops<T> O;
  tl_begin(p.tl, 4, k);
  const bool from_ll = wait_peers != 0;
  const int lane = threadIdx.x & 31;
  const int64_t t = cx->t0 + k;
  const int slot = (cx->slot0 + k) % p.R, prev = wrap_slot(slot - 1, p.R);
  T* __restrict__ tr = p.trace + (size_t)slot * p.Nld;
  const T* __restrict__ trp = p.trace + (size_t)prev * p.Nld;
  uint32_t* __restrict__ bits = p.bits + (size_t)slot * p.W;
  int32_t* __restrict__ list = p.spk_list + (size_t)slot * p.Nld;
  int32_t* __restrict__ rec_ids = cx->rec_ids;
  const int64_t rec_cap = cx->rec_cap;
  int32_t rec_base = 0;
  if (rec_ids && k > 0) {
    int64_t nxt = (int64_t)cx->rec_off[k - 1] + p.spk_cnt[prev];
    rec_base = (int32_t)(nxt < rec_cap ? nxt : rec_cap);
  }
  T* __restrict__ I_now = p.Iacc + (size_t)(t & 1) * p.Nld;
  const WView<T> vw = wview(p);
  // whole warps only: the last rank's range is padded up to Nld (padding neurons are inactive)
  const uint32_t q_lo = (uint32_t)(p.part_lo >> 2), q_hi = (uint32_t)((p.part_hi >= p.N ? p.Nld : p.part_hi) >> 2);
  const uint32_t nquad = (uint32_t)(p.Nld >> 2), n_remote = nquad - (q_hi - q_lo);
  for (uint32_t r = bid * blockDim.x + threadIdx.x; r < n_remote; r += nblk * blockDim.x) {
    const uint32_t q = r < q_lo ? r : r + (q_hi - q_lo);  // part bounds are multiples of 128: warps stay whole
    const uint32_t j0 = q << 2;
    Vec4<T> tp = ld4(trp + j0);
    uint32_t word;
    if (from_ll) {  // peer-to-peer exchange: wait for the owner's (word, step) pair, keep `bits` complete
      word = ll_word<T>(p, slot, j0 >> 5, (uint32_t)(t + 1));
      if ((lane & 7) == 0) bits[j0 >> 5] = word;
    } else {
      word = bits[j0 >> 5];
    }
    const unsigned nib = (word >> (j0 & 31)) & 0xFu;
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
              const bool row_exc = i < (uint32_t)p.Ne;  // partitioned runs are single-instance
              for (uint32_t e = s0 + lane; e < s1; e += 32) {
                const uint32_t j = (uint32_t)p.post[e];
                bool fly;
                const typename DevNet<T>::T2 ws = w_fetch<T>(vw, e, row_exc, fly);
                const T wc = c_fetch<T>(p, vw, e, fly);
                atomicAdd(&I_now[j], w_resolve<T>(p, ws, fly, p.plastic_ei || j < (uint32_t)p.Ne, 0u, wc));
              }
            }
          }
        }
      }
    }
  }
  tl_end(p.tl, 4, k);
}

template <typename T>
__global__ void __launch_bounds__(256)
remote_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k, int32_t wait_peers) {
  remote_body<T>(p, cx, k, wait_peers, blockIdx.x, gridDim.x);
}

// Partitioned runs: detection of the local neurons and the replay of the peers' detection in ONE
// grid (blocks [0, n_local) / [n_local, gridDim)).  Both halves of a step start together — the
// peers detect at the same time, so their (word, step) pairs arrive while the local half
// integrates — and the step is one graph node instead of two kernels and a join.
template <typename T, bool EULER>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? NEURON_MIN_BLOCKS : 2)
neuron_remote_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k, int32_t apply_ev, uint32_t n_local) {
  if (blockIdx.x < n_local) neuron_body<T, EULER, true>(p, cx, k, apply_ev, blockIdx.x, n_local);
  else remote_body<T>(p, cx, k, 1, blockIdx.x - n_local, gridDim.x - n_local);
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
  const typename DevNet<T>::T2* __restrict__ wsd = wview(p).cur;
  const int64_t j_hi = p.part_hi < p.N ? p.part_hi : p.N;
  for (int64_t j = p.part_lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < j_hi;
       j += (int64_t)gridDim.x * blockDim.x) {
    T acc = (T)0;
    if (p.noise_mode == SNN_NOISE_REPLAY) acc = noise[j];
    else if (p.noise_mode == SNN_NOISE_PHILOX) acc = (T)philox_noise(p.noise_seed, t, j, p.noise_amp);
    const uint32_t c0 = p.colptr[j], c1 = p.colptr[j + 1];
    for (uint32_t qq = c0; qq < c1; ++qq) {
      const uint32_t q = p.variant_g ? c1 - 1u - (qq - c0) : qq;  // DASTDP.jl:72 walks the spikers backwards
      const uint32_t pk = p.in_pre[q];
      const uint32_t i = pk & kPreMask;
      const int sl = wrap_slot(slot - (int)(pk >> kPreBits), p.R);
      const uint32_t word = p.bits[(size_t)sl * p.W + (i >> 5)];
      if ((word >> (i & 31)) & 1u) acc = O::add(acc, wsd[p.in_syn[q]].x);
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
  tl_begin(p.tl, LTD ? 1 : 5, k);
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
  const WView<T> vw = wview(p);
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
    const uint32_t inst = p.B == 1 ? 0u : i / Nper, ibase = inst * Nper;
    const bool row_exc = (i - ibase) < Ne;
    const bool row_plastic = LTD && row_exc;
    // variant_g: the LTP term of a simultaneous post spike reads the pre trace of the step before
    const T pre_up = (row_plastic && !ATOMIC)
                         ? O::mul(p.apost, p.trace[(size_t)(p.variant_g ? wrap_slot(sl - 1, p.R) : sl) * Nld + i]) : (T)0;
    if (LTD && gl == 0) n_ev += (s1 - s0);
    for (uint32_t e = (LTD ? s0 : s1) + gl; e < s2; e += G) {
      const uint32_t j = (uint32_t)p.post[e];
      if (e >= s1) {                       // arrives at k+1: push only
        bool fly;
        const typename DevNet<T>::T2 ws = w_fetch<T>(vw, e, row_exc, fly);
        const T wc = c_fetch<T>(p, vw, e, fly);
        atomicAdd(&I_next[j], w_resolve<T>(p, ws, fly, p.plastic_ei || (j - ibase) < Ne, inst, wc));
      } else if (row_plastic && (p.plastic_ei || (j - ibase) < Ne)) {
        const T dn = O::mul(p.apre, tr_now[j]);
        if (ATOMIC) {  // production: fire-and-forget reduction, LTP of this step comes from ltp_kernel
          sd_add<T>(p, vw, e, -dn, blockIdx.x, inst);
          continue;
        }
        const bool jsp = (bits_now[j >> 5] >> (j & 31)) & 1u;
        T sd = vw.cur[e].y;
        if (jsp) sd = (p.variant_g || j <= i) ? O::sub(O::add(sd, pre_up), dn) : O::add(O::sub(sd, dn), pre_up);
        else sd = O::sub(sd, dn);
        vw.cur[e].y = sd;
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
  tl_end(p.tl, LTD ? 1 : 5, k);
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
  tl_begin(p.tl, 2, k);
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
  const WView<T> vw = wview(p);
  unsigned long long n_ltp = 0;
  for (uint32_t item = gtid >> g_log2; item < (uint32_t)n_now; item += ngroups) {
    const int32_t j = list[item];
    const uint32_t inst = p.B == 1 ? 0u : (uint32_t)j / (uint32_t)p.Nper;
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
        toff[q] = (size_t)(p.variant_g ? wrap_slot(sl - 1, p.R) : sl) * Nld + i;  // DASTDP.jl:65 STDP[pre,msec]
        if (!ATOMIC && ok[q]) {
          const uint32_t word = p.bits[(size_t)sl * p.W + (i >> 5)];
          ok[q] = !((word >> (i & 31)) & 1u);
        }
      }
      T trv[3], sdv[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        trv[q] = ok[q] ? p.trace[toff[q]] : (T)0;
        sdv[q] = (!ATOMIC && ok[q]) ? vw.cur[e[q]].y : (T)0;
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (!ok[q]) continue;
        if (ATOMIC) sd_add<T>(p, vw, e[q], O::mul(p.apost, trv[q]), p.log_ltp_base + blockIdx.x, inst);
        else vw.cur[e[q]].y = O::add(sdv[q], O::mul(p.apost, trv[q]));
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) n_ltp += __shfl_down_sync(0xffffffffu, n_ltp, o);
  if ((threadIdx.x & 31) == 0 && n_ltp) atomicAdd(&blk_ltp, n_ltp);
  __syncthreads();
  if (threadIdx.x == 0 && blk_ltp) p.blk_ltp[blockIdx.x] += blk_ltp;
  tl_end(p.tl, 2, k);
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

// synapses [lo, hi) of one instance (first neuron r0): dst = learn(src); src == dst for the in-place pass
template <typename T, int UNR, bool STREAM>
__device__ __forceinline__ void learn_range(const DevNet<T>& p, const typename Real2<T>::type* src,
                                            typename Real2<T>::type* dst, int64_t lo, int64_t hi, int64_t r0, T g,
                                            int do_learn, int do_decay, int64_t tid, int64_t nth) {
  typedef typename DevNet<T>::T2 T2;
  if (hi <= lo) return;
  if (p.plastic_ei) {
    if constexpr (sizeof(T) == 4) {
      // 16-byte accesses: two synapses per load
      const int64_t lo2 = (lo + 1) & ~(int64_t)1, hi2 = hi & ~(int64_t)1;
      if (tid == 0 && lo < lo2) {
        T2 ws = src[lo]; learn_one<T>(ws, g, p.wmin, p.wmax, p.sd_decay, do_learn, do_decay); dst[lo] = ws;
      }
      if (tid == 1 % nth && hi2 < hi && hi2 >= lo2) {
        T2 ws = src[hi2]; learn_one<T>(ws, g, p.wmin, p.wmax, p.sd_decay, do_learn, do_decay); dst[hi2] = ws;
      }
      if (hi2 > lo2) {
        const float4* __restrict__ qs = reinterpret_cast<const float4*>(src + lo2);
        float4* __restrict__ qd = reinterpret_cast<float4*>(dst + lo2);
        const int64_t n4 = (hi2 - lo2) >> 1;
        int64_t x = tid;
        for (; x + (UNR - 1) * nth < n4; x += UNR * nth) {
          float4 r[UNR];
#pragma unroll
          for (int c = 0; c < UNR; ++c) r[c] = STREAM ? __ldcs(qs + x + c * nth) : qs[x + c * nth];
#pragma unroll
          for (int c = 0; c < UNR; ++c) {
            float2 s0 = make_float2(r[c].x, r[c].y), s1 = make_float2(r[c].z, r[c].w);
            learn_one<float>(s0, (float)g, (float)p.wmin, (float)p.wmax, (float)p.sd_decay, do_learn, do_decay);
            learn_one<float>(s1, (float)g, (float)p.wmin, (float)p.wmax, (float)p.sd_decay, do_learn, do_decay);
            if (STREAM) __stcs(qd + x + c * nth, make_float4(s0.x, s0.y, s1.x, s1.y));
            else qd[x + c * nth] = make_float4(s0.x, s0.y, s1.x, s1.y);
          }
        }
        for (; x < n4; x += nth) {
          float4 r = STREAM ? __ldcs(qs + x) : qs[x];
          float2 s0 = make_float2(r.x, r.y), s1 = make_float2(r.z, r.w);
          learn_one<float>(s0, (float)g, (float)p.wmin, (float)p.wmax, (float)p.sd_decay, do_learn, do_decay);
          learn_one<float>(s1, (float)g, (float)p.wmin, (float)p.wmax, (float)p.sd_decay, do_learn, do_decay);
          if (STREAM) __stcs(qd + x, make_float4(s0.x, s0.y, s1.x, s1.y));
          else qd[x] = make_float4(s0.x, s0.y, s1.x, s1.y);
        }
      }
    } else {
      int64_t e = lo + tid;
      for (; e + (UNR - 1) * nth < hi; e += UNR * nth) {
        T2 r[UNR];
#pragma unroll
        for (int c = 0; c < UNR; ++c) r[c] = src[e + c * nth];
#pragma unroll
        for (int c = 0; c < UNR; ++c) {
          learn_one<T>(r[c], g, p.wmin, p.wmax, p.sd_decay, do_learn, do_decay);
          dst[e + c * nth] = r[c];
        }
      }
      for (; e < hi; e += nth) {
        T2 ws = src[e];
        learn_one<T>(ws, g, p.wmin, p.wmax, p.sd_decay, do_learn, do_decay);
        dst[e] = ws;
      }
    }
  } else {
    // only exc->exc synapses are plastic (net_res.jl:213-216): needs the post id
    for (int64_t e = lo + tid; e < hi; e += nth) {
      if (((int64_t)p.post[e] - r0) >= p.Ne) continue;
      T2 ws = src[e];
      learn_one<T>(ws, g, p.wmin, p.wmax, p.sd_decay, do_learn, do_decay);
      dst[e] = ws;
    }
  }
}

template <typename T>
__device__ __forceinline__ T learn_rate(const DevNet<T>& p, T da) {
  typedef Ops<T> O;
  return p.rate_mode == SNN_RATE_BASE_PLUS_DA ? O::add(p.learn_base, da)
         : (p.rate_mode == SNN_RATE_DA ? da : O::mul(p.eta, da));
}

// in-place pass (exact mode, serial launches, windows whose learn steps are too dense to overlap)
template <typename T>
__global__ void __launch_bounds__(256)
learn_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k, int do_learn, int do_decay) {
  tl_begin(p.tl, 3, k);
  const int b = blockIdx.y;
  const int64_t t = cx->t0 + k;
  const int64_t r0 = (int64_t)b * p.Nper;
  const uint32_t lo = p.seg[(size_t)r0 * (p.Dmax + 1)];
  const uint32_t hi = p.seg[(size_t)(r0 + p.Ne - 1) * (p.Dmax + 1) + p.Dmax];
  const T g = learn_rate<T>(p, p.da[(size_t)((t + 1) & 1) * p.B + b]);
  typename DevNet<T>::T2* w = wview(p).cur;
  learn_range<T, 4, false>(p, w, w, lo, hi, r0, g, do_learn, do_decay, blockIdx.x * (int64_t)blockDim.x + threadIdx.x,
                    (int64_t)gridDim.x * blockDim.x);
  tl_end(p.tl, 3, k);
}

// ---------------------------------------------------------------------------------------------
// Event-driven learning (DevNet::lazy): learn!(k) of instance b only advances (H, r^k, count) —
// lazy_commit_kernel — unless this is a re-basing step, where lazy_learn_kernel first runs the dense
// pass  s0 = clamp(s0 + H*sd' - C), sd' = sd' * r^k, C = 0  over the plastic synapses with the
// advanced values (both kernels recompute them from the same, not yet committed, state).
// force: materialise only (snn_get_array / snn_set_array), no learn step is counted.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ bool lazy_next(const DevNet<T>& p, const WindowCtx<T>* __restrict__ cx, int32_t k, int b,
                                          int do_learn, int do_decay, int force, T& Hn, T& Rn, int& Kn) {
  typedef Ops<T> O;
  const int64_t t = cx->t0 + k;
  const T da = p.da[(size_t)((t + 1) & 1) * p.B + b];
  const T H = p.lzH[b], R = p.lzR[b];
  Hn = do_learn ? O::add(H, O::mul(learn_rate<T>(p, da), R)) : H;
  Rn = do_decay ? O::mul(R, p.sd_decay) : R;
  Kn = p.lzK[b] + (do_learn ? 1 : 0);
  return force || Kn >= p.lazy || (do_learn && da > p.lazy_da);
}
template <typename T>
__global__ void __launch_bounds__(256)
lazy_learn_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k, int do_learn, int do_decay, int force) {
  typedef Ops<T> O;
  tl_begin(p.tl, 3, k);
  const int b = blockIdx.y;
  T Hn, Rn; int Kn;
  if (lazy_next<T>(p, cx, k, b, do_learn, do_decay, force, Hn, Rn, Kn)) {
    const int64_t r0 = (int64_t)b * p.Nper;
    const int64_t lo = p.seg[(size_t)r0 * (p.Dmax + 1)];
    const int64_t hi = p.seg[(size_t)(r0 + p.Ne - 1) * (p.Dmax + 1) + p.Dmax];
    typename DevNet<T>::T2* w = wview(p).cur;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x, tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    auto one = [&](int64_t e) {
      typename DevNet<T>::T2 ac = w[e];  // (sd', C)
      const T x = O::sub(O::add(p.cacc[e], O::mul(Hn, ac.x)), ac.y);
      p.cacc[e] = x > p.wmax ? p.wmax : (x < p.wmin ? p.wmin : x);
      ac.x = O::mul(ac.x, Rn);
      ac.y = (T)0;
      w[e] = ac;
    };
    bool done = false;
    if constexpr (sizeof(T) == 4) {
      if (p.plastic_ei) {  // two synapses per 16-byte access, four accesses in flight (24 B per synapse in all)
        done = true;
        const int64_t lo2 = (lo + 1) & ~(int64_t)1, hi2 = hi & ~(int64_t)1;
        if (tid == 0 && lo < lo2 && lo < hi) one(lo);
        if (tid == 1 % nth && hi2 < hi && hi2 >= lo2) one(hi2);
        if (hi2 > lo2) {
          float4* __restrict__ q = reinterpret_cast<float4*>(w + lo2);
          float2* __restrict__ c2 = reinterpret_cast<float2*>(p.cacc + lo2);
          const int64_t n2 = (hi2 - lo2) >> 1;
          const float H = (float)Hn, R = (float)Rn, wmin = (float)p.wmin, wmax = (float)p.wmax;
          // r = (sd'0, C0, sd'1, C1), c = (s0_0, s0_1)
          auto two = [&](float4 r, float2 c) {
            float x0 = c.x + H * r.x - r.y, x1 = c.y + H * r.z - r.w;
            x0 = x0 > wmax ? wmax : (x0 < wmin ? wmin : x0);
            x1 = x1 > wmax ? wmax : (x1 < wmin ? wmin : x1);
            return make_float2(x0, x1);
          };
          int64_t x = tid;
          for (; x + 3 * nth < n2; x += 4 * nth) {
            float4 r[4]; float2 c[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { r[u] = q[x + u * nth]; c[u] = c2[x + u * nth]; }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              c2[x + u * nth] = two(r[u], c[u]);
              q[x + u * nth] = make_float4(r[u].x * R, 0.f, r[u].z * R, 0.f);
            }
          }
          for (; x < n2; x += nth) {
            const float4 r = q[x]; const float2 c = c2[x];
            c2[x] = two(r, c);
            q[x] = make_float4(r.x * R, 0.f, r.z * R, 0.f);
          }
        }
      }
    }
    if (!done)
      for (int64_t e = lo + tid; e < hi; e += nth) {
        if (!p.plastic_ei && ((int64_t)p.post[e] - r0) >= p.Ne) continue;
        one(e);
      }
  }
  tl_end(p.tl, 3, k);
}
// API layout (w, sd) <-> event-driven layout (wsd = (sd', C = 0), cacc = s0); only right after a re-basing
template <typename T>
__global__ void lazy_layout_kernel(typename Real2<T>::type* __restrict__ wsd, T* __restrict__ cacc, int64_t E, int to_lazy) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    typename Real2<T>::type a = wsd[e];
    if (to_lazy) { cacc[e] = a.x; a.x = a.y; a.y = (T)0; }
    else { a.y = a.x; a.x = cacc[e]; }
    wsd[e] = a;
  }
}
template <typename T>
__global__ void lazy_commit_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k, int do_learn, int do_decay,
                                   int force) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  T Hn, Rn; int Kn;
  if (lazy_next<T>(p, cx, k, b, do_learn, do_decay, force, Hn, Rn, Kn)) { Hn = (T)0; Rn = (T)1; Kn = 0; }
  p.lzH[b] = Hn; p.lzR[b] = Rn; p.lzInv[b] = (T)1 / Rn; p.lzK[b] = Kn;
}

// ---------------------------------------------------------------------------------------------
// Asynchronous learn pass (production mode).  learn!(k) runs as ONE persistent kernel on a
// low-priority stream while the neuron/synapse/ltp kernels of steps k+1.. run beside it.
//   learn_ctl(begin)  : g[b] from da as decayed by step k, flip cur, frontier = 0, new epoch
//   learn_pass        : blocks take 32K-synapse tiles in ascending order: wsdX[cur] =
//                       learn(wsdX[cur^1]); a finished tile is flagged and the frontier (the
//                       contiguous prefix of finished tiles) advanced with one atomicMax on `state`
//   learn_ctl(all)    : frontier = all
//   learn_replay      : (before the next begin / at the window end) sd[e] += logged deltas
// Readers (w_fetch/w_resolve, sd_add) never wait: an unprocessed synapse's post-learn weight is
// formed on the fly from the frozen buffer, its sd updates are logged.
// ---------------------------------------------------------------------------------------------
// learn_pass_kernel: the register-staged (LSU) version of the pass, kept as the A/B alternative of
// learn_pass_tma_kernel (snn_config.reserved[1] & 512).  <= 64 registers: two resident blocks must
// leave half of the SM's register file to the step kernels (at 128 registers two blocks owned the
// whole file and nothing ran beside the pass).
template <typename T>
__global__ void __launch_bounds__(256, 4)
learn_pass_kernel(DevNet<T> p, int32_t k, int do_decay) {
  tl_begin(p.tl, 3, k);
  __shared__ uint32_t s_tile;
  const uint32_t st0 = *reinterpret_cast<const volatile uint32_t*>(&p.lc->state);
  const uint32_t cur = st0 >> 31, curbit = st0 & 0x80000000u;
  const uint32_t epoch = p.lc->epoch;
  const typename DevNet<T>::T2* __restrict__ src = p.wsdX[cur ^ 1u];
  typename DevNet<T>::T2* __restrict__ dst = p.wsdX[cur];
  const uint32_t TS = p.tile_syn, U = TS >> 5;  // frontier units per tile
  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(&p.lc->next_tile, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    __syncthreads();
    if (tile >= p.n_tiles) break;
    const int64_t a = (int64_t)tile * TS, z = a + TS;
    // instances whose plastic range meets [a, z): first b with inst_hi[b] > a
    int b = 0;
    if (p.B > 1) {
      int lo = 0, hi = p.B;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int64_t)p.inst_hi[mid] > a) hi = mid; else lo = mid + 1; }
      b = lo;
    }
    bool done = false;
    if constexpr (sizeof(T) == 4) {
      // whole tile inside one instance's plastic range: 16-byte accesses, 8 in flight per thread.
      // Every tile starts at its own offset (the blocks walk tiles that are TS*8 bytes apart in
      // lockstep).  Measured neutral — the slowness this was written against was the per-tile CAS
      // chain of the frontier, see learn_pass_tma_kernel — kept because it costs nothing.
      if (p.plastic_ei && b < p.B && (int64_t)p.inst_lo[b] <= a && (int64_t)p.inst_hi[b] >= z) {
        const float4* __restrict__ qs = reinterpret_cast<const float4*>(src + a);
        float4* __restrict__ qd = reinterpret_cast<float4*>(dst + a);
        const uint32_t n4 = TS >> 1;  // multiple of 512
        const uint32_t rot = ((tile * 2654435761u) >> 8) % (n4 >> 5) << 5;
        const float g = p.lg[b];
        for (uint32_t x0 = threadIdx.x; x0 < n4; x0 += 8 * 256) {
          float4 r[8]; uint32_t ix[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint32_t x = x0 + c * 256 + rot;
            if (x >= n4) x -= n4;
            ix[c] = x;
            if (x0 + c * 256 < n4) r[c] = __ldcs(qs + x);
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (x0 + c * 256 >= n4) continue;
            float2 s0 = make_float2(r[c].x, r[c].y), s1 = make_float2(r[c].z, r[c].w);
            learn_one<float>(s0, g, (float)p.wmin, (float)p.wmax, (float)p.sd_decay, 1, do_decay);
            learn_one<float>(s1, g, (float)p.wmin, (float)p.wmax, (float)p.sd_decay, 1, do_decay);
            __stcs(qd + ix[c], make_float4(s0.x, s0.y, s1.x, s1.y));
          }
        }
        done = true;
      }
    }
    if (!done) {
      for (; b < p.B && (int64_t)p.inst_lo[b] < z; ++b) {
        const int64_t lo = max((int64_t)p.inst_lo[b], a), hi = min((int64_t)p.inst_hi[b], z);
        learn_range<T, 8, true>(p, src, dst, lo, hi, (int64_t)b * p.Nper, p.lg[b], 1, do_decay, threadIdx.x, blockDim.x);
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      const uint32_t lane = threadIdx.x;
      if (lane == 0) {
        __threadfence();  // the tile's stores are visible before its flag
        *reinterpret_cast<volatile uint32_t*>(&p.tile_done[tile]) = epoch;
        __threadfence();
      }
      __syncwarp();
      // advance the frontier over the contiguous prefix of finished tiles: the warp scans 32 flags
      // per round trip, then ONE atomicMax (the cur bit is constant during a pass, so the packed
      // word is monotonic).  A serial CAS per tile would bound the whole pass by n_tiles L2 round
      // trips.  Losing a race only leaves the frontier behind until the next tile finishes.
      const uint32_t f0 = (*reinterpret_cast<volatile uint32_t*>(&p.lc->state) & kFrontAll) / U;
      uint32_t f = f0;
      for (;;) {
        const uint32_t idx = f + lane;
        const bool d = idx < p.n_tiles && *reinterpret_cast<volatile uint32_t*>(&p.tile_done[idx]) == epoch;
        const unsigned m = __ballot_sync(0xffffffffu, d);
        const int run = m == 0xffffffffu ? 32 : __ffs(~m) - 1;
        f += run;
        if (run < 32) break;
      }
      if (lane == 0 && f > f0) {
        __threadfence();
        atomicMax(&p.lc->state, curbit | (f * U));
      }
    }
  }
  tl_end(p.tl, 3, k);
}

// ---- TMA (bulk async copy) + mbarrier primitives, sm_100a
__device__ __forceinline__ uint32_t smem_u32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
// shared -> global, tracked by bulk async-groups
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// learn_pass_tma_kernel — the asynchronous pass, staged through shared memory by the TMA engine.
// The pass has to stream 16 B per plastic synapse at HBM speed for a whole learn period WHILE the
// step kernels run on the same SMs.  An LSU version needs ~10 MB of loads in flight = hundreds of
// threads x 8 x 16-B registers per SM (it took half of every SM's register file and halved the
// step kernels' occupancy).  Here one elected thread keeps kPassStages-1 16-KB bulk loads in
// flight per block (cp.async.bulk + mbarrier complete_tx), 128 threads apply learn! in place in
// shared memory, and the result leaves through a bulk store: ~4 K registers and 1/16 of the thread
// slots per SM.  Loads and stores carry an L2 evict-first policy so that the pass does not flush
// the neuron state (36 MB per step, L2-resident) out of the 126 MB L2.
// ---------------------------------------------------------------------------------------------
constexpr int kPassThreads = 256, kPassStages = 4, kPassMaxStages = 12;
constexpr uint32_t kPassChunkBytes = 16384;

template <typename T>
__global__ void __launch_bounds__(kPassThreads)
learn_pass_tma_kernel(DevNet<T> p, int32_t k, int do_decay, int n_stages) {
  typedef typename DevNet<T>::T2 T2;
  tl_begin(p.tl, 3, k);
  extern __shared__ __align__(128) unsigned char pass_smem[];
  __shared__ __align__(8) uint64_t mbar[kPassMaxStages];
  const uint32_t S = (uint32_t)n_stages;
  __shared__ uint32_t s_tile;
  constexpr uint32_t CH = kPassChunkBytes / sizeof(T2);  // synapses per chunk
  T2* const buf = reinterpret_cast<T2*>(pass_smem);      // [kPassStages][CH]
  const uint32_t tid = threadIdx.x;
  if (tid == 0) {
    for (uint32_t s_ = 0; s_ < S; ++s_) mbar_init(&mbar[s_], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint64_t pol = l2_evict_first_policy();
  const uint32_t st0 = *reinterpret_cast<const volatile uint32_t*>(&p.lc->state);
  const uint32_t cur = st0 >> 31, curbit = st0 & 0x80000000u;
  const uint32_t epoch = p.lc->epoch;
  const T2* __restrict__ src = p.wsdX[cur ^ 1u];
  T2* __restrict__ dst = p.wsdX[cur];
  const uint32_t TS = p.tile_syn, U = TS >> 5;
  uint32_t q = 0;  // chunks consumed so far by this block: stage = q % S, parity = (q / S) & 1
  for (;;) {
    if (tid == 0) s_tile = atomicAdd(&p.lc->next_tile, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    __syncthreads();
    if (tile >= p.n_tiles) break;
    const int64_t a = (int64_t)tile * TS, z = a + TS;
    int b = 0;
    if (p.B > 1) {  // first instance whose plastic range ends after a
      int lo = 0, hi = p.B;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int64_t)p.inst_hi[mid] > a) hi = mid; else lo = mid + 1; }
      b = lo;
    }
    for (; b < p.B && (int64_t)p.inst_lo[b] < z; ++b) {
      int64_t lo = max((int64_t)p.inst_lo[b], a), hi = min((int64_t)p.inst_hi[b], z);
      if (hi <= lo) continue;
      const T g = p.lg[b];
      const int64_t r0 = (int64_t)b * p.Nper;
      auto one = [&](T2& ws, int64_t e) {
        if (!p.plastic_ei && ((int64_t)p.post[e] - r0) >= p.Ne) return;  // exc->inh: not plastic (net_res.jl:213-216)
        learn_one<T>(ws, g, p.wmin, p.wmax, p.sd_decay, 1, do_decay);
      };
      // bulk copies need 16-byte alignment: odd fp32 ends are done by hand
      if (sizeof(T2) == 8) {
        if ((lo & 1) && tid == 0) { T2 ws = src[lo]; one(ws, lo); dst[lo] = ws; }
        lo = (lo + 1) & ~(int64_t)1;
        if ((hi & 1) && hi > lo && tid == 1) { T2 ws = src[hi - 1]; one(ws, hi - 1); dst[hi - 1] = ws; }
        hi &= ~(int64_t)1;
      }
      if (hi <= lo) continue;
      const uint32_t n = (uint32_t)(hi - lo), nch = (n + CH - 1) / CH;
      auto chunk_bytes = [&](uint32_t i) { return (uint32_t)((i + 1 < nch ? CH : n - i * CH) * sizeof(T2)); };
      if (tid == 0) {
        for (uint32_t i = 0; i < nch && i < S; ++i) {
          const uint32_t sg = (q + i) % S;
          mbar_expect_tx(&mbar[sg], chunk_bytes(i));
          tma_load_1d(buf + (size_t)sg * CH, src + lo + (size_t)i * CH, chunk_bytes(i), &mbar[sg], pol);
        }
      }
      for (uint32_t i = 0; i < nch; ++i) {
        const uint32_t sg = (q + i) % S, par = ((q + i) / S) & 1u;
        mbar_wait(&mbar[sg], par);
        T2* const bs = buf + (size_t)sg * CH;
        const uint32_t cnt = chunk_bytes(i) / (uint32_t)sizeof(T2);
        const int64_t e0 = lo + (int64_t)i * CH;
        if constexpr (sizeof(T2) == 8) {
          // 16-byte shared-memory accesses: two synapses per thread and iteration (cnt is even)
          float4* const b4 = reinterpret_cast<float4*>(bs);
#pragma unroll 4
          for (uint32_t x = tid; x < (cnt >> 1); x += kPassThreads) {
            float4 r = b4[x];
            float2 s0 = make_float2(r.x, r.y), s1 = make_float2(r.z, r.w);
            one(s0, e0 + 2 * x);
            one(s1, e0 + 2 * x + 1);
            b4[x] = make_float4(s0.x, s0.y, s1.x, s1.y);
          }
        } else {
          for (uint32_t x = tid; x < cnt; x += kPassThreads) {
            T2 ws = bs[x];
            one(ws, e0 + x);
            bs[x] = ws;
          }
        }
        fence_proxy_async();  // generic-proxy writes to smem -> visible to the bulk store
        __syncthreads();
        if (tid == 0) {
          tma_store_1d(dst + e0, bs, chunk_bytes(i), pol);
          tma_commit();
          if (i >= 1 && i - 1 + S < nch) {
            tma_wait_read<1>();  // store(i-1) has read its stage: refill it with chunk i-1+S
            const uint32_t c = i - 1 + S, sc = (q + c) % S;
            mbar_expect_tx(&mbar[sc], chunk_bytes(c));
            tma_load_1d(buf + (size_t)sc * CH, src + lo + (size_t)c * CH, chunk_bytes(c), &mbar[sc], pol);
          }
        }
      }
      if (tid == 0) tma_wait_read<0>();  // every stage is free again before the next range
      __syncthreads();
      q += nch;
    }
    // the tile's bulk stores are complete (written, not merely read) before its flag goes up
    if (tid == 0) tma_wait_all();
    __syncthreads();
    if (tid < 32) {
      const uint32_t lane = tid;
      if (lane == 0) {
        __threadfence();
        *reinterpret_cast<volatile uint32_t*>(&p.tile_done[tile]) = epoch;
        __threadfence();
      }
      __syncwarp();
      const uint32_t f0 = (*reinterpret_cast<volatile uint32_t*>(&p.lc->state) & kFrontAll) / U;
      uint32_t f = f0;
      for (;;) {
        const uint32_t idx = f + lane;
        const bool d = idx < p.n_tiles && *reinterpret_cast<volatile uint32_t*>(&p.tile_done[idx]) == epoch;
        const unsigned m = __ballot_sync(0xffffffffu, d);
        const int run = m == 0xffffffffu ? 32 : __ffs(~m) - 1;
        f += run;
        if (run < 32) break;
      }
      if (lane == 0 && f > f0) {
        __threadfence();
        atomicMax(&p.lc->state, curbit | (f * U));
      }
    }
  }
  tl_end(p.tl, 3, k);
}

enum { kCtlBegin = 0, kCtlAll = 1 };
// begin of a pass, by one block: g[b] from da as decayed by step k, flip cur, frontier = 0, new epoch
template <typename T>
__device__ __forceinline__ void learn_begin(const DevNet<T>& p, const WindowCtx<T>* __restrict__ cx, int32_t k) {
  const uint32_t st = p.lc->state;
  const int64_t t = cx->t0 + k;
  for (int b = threadIdx.x; b < p.B; b += blockDim.x) p.lg[b] = learn_rate<T>(p, p.da[(size_t)((t + 1) & 1) * p.B + b]);
  if (threadIdx.x == 0) {
    p.lc->state = ((st >> 31) ^ 1u) << 31;
    p.lc->epoch += 1u;
    p.lc->next_tile = 0u;
    p.lc->spilled = 0u;  // the replay before this has merged the spill
  }
}
template <typename T>
__global__ void learn_ctl_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k, int mode) {
  if (mode == kCtlBegin) learn_begin<T>(p, cx, k);
  else if (threadIdx.x == 0) p.lc->state = (p.lc->state & 0x80000000u) | kFrontAll;
}

// one block per log region (block-stride); empties the regions it replays
template <typename T>
__global__ void __launch_bounds__(256)
learn_replay_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k, int begin_next) {
  typename DevNet<T>::T2* w = p.wsdX[p.lc->state >> 31];
  for (uint32_t r = blockIdx.x; r < p.log_regions; r += gridDim.x) {
    const uint32_t cnt = p.log_cnt[(size_t)r * kLogStride];
    const uint32_t n = cnt < p.log_region_cap ? cnt : p.log_region_cap;
    const LogEntry<T>* __restrict__ lg = p.dlog + (size_t)r * p.log_region_cap;
    for (uint32_t q = threadIdx.x; q < n; q += blockDim.x) {
      const LogEntry<T> le = lg[q];
      atomicAdd(&w[le.e].y, le.d);
    }
    __syncthreads();
    if (threadIdx.x == 0 && cnt) p.log_cnt[(size_t)r * kLogStride] = 0u;
  }
  if (p.lc->spilled) {
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = tid; e < p.E; e += nth) {
      const T d = p.dsd[e];
      if (d != (T)0) { atomicAdd(&w[e].y, d); p.dsd[e] = (T)0; }
    }
  }
  if (begin_next) {
    // the last block to finish begins the next pass: one node less on the learn step's critical path
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&p.lc->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {
      __threadfence();
      learn_begin<T>(p, cx, k);
      if (threadIdx.x == 0) p.lc->ticket = 0u;
    }
  }
}
template <typename T>
__global__ void learn_spill_reset_kernel(DevNet<T> p) { p.lc->spilled = 0u; }

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

// ---- synapse probes: (w, sd) of a few chosen synapses after every step — the reference's
// `shist[time,:] = [net.s[n1,syn], net.sd[n1,syn]]` (Old_net/Julia_DA_STDPv2.jl:91, newnet_DASTDP.jl:57-60)
static __global__ void probe_locate_kernel(const uint32_t* __restrict__ perm, int64_t E, const int64_t* __restrict__ want,
                                    int n, uint32_t* __restrict__ pos) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = perm[e];
    for (int q = 0; q < n; ++q) if (want[q] == c) pos[q] = (uint32_t)e;
  }
}
template <typename T>
__global__ void probe_kernel(DevNet<T> p, const uint32_t* __restrict__ pos, int n, double* __restrict__ out, int32_t k) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const typename DevNet<T>::T2 ws = wview(p).cur[pos[q]];
  out[((size_t)k * n + q) * 2] = (double)ws.x;
  out[((size_t)k * n + q) * 2 + 1] = (double)ws.y;
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
    int prio_lo = 0, prio_hi = 0;  // numerically: lowest priority = prio_lo, highest = prio_hi
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    int prio_mid = prio_hi < prio_lo ? prio_hi + 1 : prio_lo;
    SNN_CUDA_OK(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking));
    SNN_CUDA_OK(cudaStreamCreateWithPriority(&side_stream_, cudaStreamNonBlocking, prio_mid));
    SNN_CUDA_OK(cudaStreamCreateWithPriority(&chain_stream_, cudaStreamNonBlocking, prio_hi));   // neuron chain
    SNN_CUDA_OK(cudaStreamCreateWithPriority(&learn_stream_, cudaStreamNonBlocking, prio_lo));   // async learn pass
    stream_ = own_stream_;
    SNN_CUDA_OK(cudaEventCreate(&ev_begin_));
    SNN_CUDA_OK(cudaEventCreate(&ev_end_));
    for (auto& e : dep_ev_) SNN_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    SNN_CUDA_OK(cudaStreamCreateWithPriority(&side2_stream_, cudaStreamNonBlocking, prio_mid));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg_.device);
    n_sm_ = sms;
    use_graph_ = cfg_.reserved[0] == 0;
    DevNet<T>& p = p_;
    memset(&p, 0, sizeof(p));
    p.N = cfg_.n_neurons; p.Nld = (p.N + 127) / 128 * 128; p.Nper = cfg_.n_per_instance; p.Ne = cfg_.n_exc;
    // ring: R-1 steps of history + the slot being prepared; variant_g reads one step further back
    p.Dmax = cfg_.max_delay; p.R = p.Dmax + 2 + (cfg_.variant_g ? 1 : 0); p.B = (int32_t)(p.N / p.Nper); p.W = (int32_t)(p.Nld / 32);
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
    SNN_CUDA_OK(cudaMalloc(&p.xdone, 2 * sizeof(uint32_t)));
    SNN_CUDA_OK(cudaMemsetAsync(p.xdone, 0, 2 * sizeof(uint32_t), stream_));
    SNN_CUDA_OK(cudaMalloc(&p.lc, sizeof(LearnCtl)));
    SNN_CUDA_OK(cudaMalloc(&p.lg, (size_t)p.B * sizeof(T)));
    SNN_CUDA_OK(cudaMemsetAsync(p.lg, 0, (size_t)p.B * sizeof(T), stream_));
    {
      LearnCtl lc0; lc0.state = kFrontAll; lc0.epoch = 0; lc0.next_tile = 0; lc0.spilled = 0; lc0.ticket = 0;
      SNN_CUDA_OK(cudaMemcpyAsync(p.lc, &lc0, sizeof(lc0), cudaMemcpyHostToDevice, stream_));
      SNN_CUDA_OK(cudaStreamSynchronize(stream_));
    }
    SNN_CUDA_OK(cudaMalloc(&d_ctx_, sizeof(WindowCtx<T>)));
    SNN_CUDA_OK(cudaMallocHost(&h_ctx_, sizeof(WindowCtx<T>)));
    SNN_CUDA_OK(cudaMallocHost(&h_blk_, 2 * kMaxSynBlocks * sizeof(unsigned long long)));
    SNN_CUDA_OK(cudaMemsetAsync(p.ho, 0, N * sizeof(T), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.Iacc, 0, 2 * N * sizeof(T), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.trace, 0, (size_t)p.R * N * sizeof(T), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.bits, 0, (size_t)p.R * p.W * sizeof(uint32_t), stream_));
    SNN_CUDA_OK(cudaMalloc(&p.ll, (size_t)p.R * p.W * sizeof(uint2)));
    SNN_CUDA_OK(cudaMemsetAsync(p.ll, 0xFF, (size_t)p.R * p.W * sizeof(uint2), stream_));  // tag 0xffffffff: never a step
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
    p.ho_inh = cfg_.ho_inh; p.variant_g = cfg_.variant_g;
    p.exp_flags = cfg_.reserved[1];
    p.part_lo = 0; p.part_hi = p.N;
    p.static_cur = -1;
    if (cfg_.reserved[1] & 2) SNN_CUDA_OK(cudaMalloc(&p.tl, 2 * kTlClasses * kTlSteps * sizeof(unsigned long long)));
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
    free_synapse_state();
    const int64_t Ea = p.E > 0 ? p.E : 1;
    typedef typename DevNet<T>::T2 T2;
    SNN_CUDA_OK(cudaMalloc(&p.wsdX[0], Ea * sizeof(T2)));
    p.wsd = p.wsdX[0]; p.wsdX[1] = p.wsdX[0];
    cur_ = 0;
    {
      LearnCtl lc0; lc0.state = kFrontAll; lc0.epoch = 0; lc0.next_tile = 0; lc0.spilled = 0; lc0.ticket = 0;
      SNN_CUDA_OK(cudaMemcpyAsync(p.lc, &lc0, sizeof(lc0), cudaMemcpyHostToDevice, stream_));
    }
    // plastic synapse range of every instance (exc rows are contiguous in the internal order)
    inst_lo_.assign(p.B, 0); inst_hi_.assign(p.B, 0);
    if (p.E > 0 && p.Ne > 0) {
      const size_t pitch = (size_t)p.Nper * (p.Dmax + 1) * sizeof(uint32_t);
      SNN_CUDA_OK(cudaMemcpy2D(inst_lo_.data(), sizeof(uint32_t), p.seg, pitch, sizeof(uint32_t), p.B, cudaMemcpyDeviceToHost));
      SNN_CUDA_OK(cudaMemcpy2D(inst_hi_.data(), sizeof(uint32_t), p.seg + (size_t)(p.Ne - 1) * (p.Dmax + 1) + p.Dmax, pitch,
                               sizeof(uint32_t), p.B, cudaMemcpyDeviceToHost));
    }
    n_plastic_ = 0;
    for (int b = 0; b < p.B; ++b) n_plastic_ += (int64_t)inst_hi_[b] - (int64_t)inst_lo_[b];
    // production mode: second (w, sd) buffer + update log for the asynchronous learn pass
    // reserved[4] >= 16: event-driven learning, re-basing every reserved[4]-16 learn steps (and at every
    // learn step while da > reserved[5]/1000 when that is > 0); production mode only
    p.lazy = 0;
    if (cfg_.reserved[4] >= 16) {
      if (p.exact) { err = "event-driven learning (reserved[4] >= 16) needs SNN_MODE_FAST"; return SNN_ERR_INVALID; }
      if (cfg_.sd_decay_mode == SNN_SD_DECAY_EVERY_STEP) {
        err = "event-driven learning does not support SNN_SD_DECAY_EVERY_STEP"; return SNN_ERR_INVALID;
      }
      if (n_plastic_ > 0) {
        p.lazy = std::max(1, cfg_.reserved[4] - 16);
        p.lazy_da = cfg_.reserved[5] > 0 ? (T)(cfg_.reserved[5] * 1e-3) : (T)1e30;
        SNN_CUDA_OK(cudaMalloc(&p.cacc, Ea * sizeof(T)));
        SNN_CUDA_OK(cudaMemsetAsync(p.cacc, 0, Ea * sizeof(T), stream_));
        SNN_CUDA_OK(cudaMalloc(&p.lzH, 3 * (size_t)p.B * sizeof(T)));
        p.lzR = p.lzH + p.B; p.lzInv = p.lzR + p.B;
        SNN_CUDA_OK(cudaMalloc(&p.lzK, (size_t)p.B * sizeof(int32_t)));
        SNN_CUDA_OK(cudaMemsetAsync(p.lzK, 0, (size_t)p.B * sizeof(int32_t), stream_));
        std::vector<T> init(3 * (size_t)p.B, (T)1);
        std::fill(init.begin(), init.begin() + p.B, (T)0);
        SNN_CUDA_OK(cudaMemcpyAsync(p.lzH, init.data(), init.size() * sizeof(T), cudaMemcpyHostToDevice, stream_));
        SNN_CUDA_OK(cudaStreamSynchronize(stream_));
      }
    }
    have_async_ = !p.lazy && !p.exact && use_graph_ && cfg_.reserved[4] != 1 && n_plastic_ > 0 &&
                  cfg_.sd_decay_mode != SNN_SD_DECAY_EVERY_STEP;
    if (have_async_) {
      SNN_CUDA_OK(cudaMalloc(&p.wsdX[1], Ea * sizeof(T2)));
      const Launch g = grids();
      p.log_regions = (uint32_t)(g.gA + g.gP); p.log_ltp_base = (uint32_t)g.gA;
      const int64_t log_cap = cfg_.reserved[5] > 0 ? cfg_.reserved[5]
                                                  : std::min<int64_t>(std::max<int64_t>(1 << 20, n_plastic_ / 8), 1ll << 28);
      p.log_region_cap = (uint32_t)std::max<int64_t>(1, log_cap / p.log_regions);
      SNN_CUDA_OK(cudaMalloc(&p.dlog, (size_t)p.log_regions * p.log_region_cap * sizeof(LogEntry<T>)));
      SNN_CUDA_OK(cudaMalloc(&p.log_cnt, (size_t)p.log_regions * kLogStride * sizeof(uint32_t)));
      SNN_CUDA_OK(cudaMemsetAsync(p.log_cnt, 0, (size_t)p.log_regions * kLogStride * sizeof(uint32_t), stream_));
      SNN_CUDA_OK(cudaMalloc(&p.dsd, Ea * sizeof(T)));
      SNN_CUDA_OK(cudaMemsetAsync(p.dsd, 0, Ea * sizeof(T), stream_));
      p.tile_syn = cfg_.reserved[3] > 0 ? (uint32_t)cfg_.reserved[3] * 1024u : kTileSyn;
      p.n_tiles = (uint32_t)(((int64_t)inst_hi_[p.B - 1] + p.tile_syn - 1) / p.tile_syn);
      SNN_CUDA_OK(cudaMalloc(&p.tile_done, (size_t)std::max<uint32_t>(1, p.n_tiles) * sizeof(uint32_t)));
      SNN_CUDA_OK(cudaMemsetAsync(p.tile_done, 0, (size_t)std::max<uint32_t>(1, p.n_tiles) * sizeof(uint32_t), stream_));
      uint32_t* d_lohi = nullptr;
      SNN_CUDA_OK(cudaMalloc(&d_lohi, 2 * (size_t)p.B * sizeof(uint32_t)));
      SNN_CUDA_OK(cudaMemcpyAsync(d_lohi, inst_lo_.data(), (size_t)p.B * sizeof(uint32_t), cudaMemcpyHostToDevice, stream_));
      SNN_CUDA_OK(cudaMemcpyAsync(d_lohi + p.B, inst_hi_.data(), (size_t)p.B * sizeof(uint32_t), cudaMemcpyHostToDevice, stream_));
      p.inst_lo = d_lohi; p.inst_hi = d_lohi + p.B;
    }
    lazy_layout_ = false;
    int rc = upload_wsd(h_w, 0, 1, err);
    if (rc) return rc;
    rc = lazy_from_api(err);  // event-driven learning keeps its own layout between API calls
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
        { const int rcf = lazy_to_api(err); if (rcf) return rcf; }
        double* d = nullptr;
        SNN_CUDA_OK(cudaMalloc(&d, p.E * sizeof(double)));
        wsd_to_caller<T><<<grid_for(p.E), 256, 0, stream_>>>(p.wsdX[cur_], topo_.perm, d, p.E, field == SNN_FIELD_W ? 0 : 1);
        cudaError_t e = cudaMemcpyAsync(host, d, p.E * sizeof(double), cudaMemcpyDeviceToHost, stream_);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
        cudaFree(d);
        { const int rcf = lazy_from_api(err); if (rcf) return rcf; }
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
        {
          int rcf = lazy_to_api(err);
          if (rcf) return rcf;
          rcf = upload_wsd(host, field == SNN_FIELD_W ? 0 : 1, 0, err);
          if (rcf) return rcf;
          return lazy_from_api(err);
        }
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
    if (t_ != 0) { err = "set the partition before the first step"; return SNN_ERR_STATE; }
    if (!whole && (p_.lazy || cfg_.reserved[4] >= 16)) {
      err = "event-driven learning (reserved[4] >= 16) has not been validated on neuron-partitioned runs"; return SNN_ERR_INVALID;
    }
    drop_graph();
    p_.part_lo = lo; p_.part_hi = hi;
    exchange_ = whole ? nullptr : fn; exchange_user_ = user;
    return SNN_OK;
  }

  int get_timeline(unsigned long long* host, int64_t count, std::string& err) override {
    if (!p_.tl) { err = "timeline recording is off (snn_config.reserved[1] & 2)"; return SNN_ERR_STATE; }
    if (count != 2 * kTlClasses * kTlSteps) { err = "timeline count mismatch"; return SNN_ERR_INVALID; }
    SNN_CUDA_OK(cudaSetDevice(cfg_.device));
    SNN_CUDA_OK(cudaMemcpy(host, p_.tl, count * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return SNN_OK;
  }
  // ---- peer-to-peer exchange set-up (snn_partition_export / snn_partition_connect)
  int export_peer(PeerBlob* b, std::string& err) override {
    SNN_CUDA_OK(cudaSetDevice(cfg_.device));
    memset(b, 0, sizeof(*b));
    b->magic = kPeerMagic; b->pid = (int32_t)getpid(); b->device = cfg_.device;
    b->proc_nonce = process_nonce();
    b->raw_ring = (uint64_t)(uintptr_t)p_.ll;
    b->words_per_slot = p_.W; b->ring = p_.R;
    SNN_CUDA_OK(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(b->ipc_ring), p_.ll));
    return SNN_OK;
  }
  int connect_peers(int rank, int world, const PeerBlob* blobs, std::string& err) override {
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) { err = "rank / world out of range (world <= 8)"; return SNN_ERR_INVALID; }
    if (t_ != 0) { err = "connect the peers before the first step"; return SNN_ERR_STATE; }
    SNN_CUDA_OK(cudaSetDevice(cfg_.device));
    drop_graph();
    close_peers();
    DevNet<T>& p = p_;
    for (int r = 0; r < world; ++r) {
      const PeerBlob& b = blobs[r];
      if (b.magic != kPeerMagic || b.words_per_slot != p.W || b.ring != p.R) {
        err = "peer blob of rank " + std::to_string(r) + " does not match this network (size / max_delay)"; return SNN_ERR_INVALID;
      }
      if (r == rank) { p.peer_ll[r] = p.ll; continue; }
      // several handles in ONE process: plain pointers.  The pid alone does not identify a process (ranks in
      // different containers / pid namespaces can share one), so the blob carries a per-process random nonce
      if (b.pid == (int32_t)getpid() && b.proc_nonce == process_nonce()) {
        if (b.device != cfg_.device) {
          cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return SNN_ERR_CUDA; }
          cudaGetLastError();
        }
        p.peer_ll[r] = reinterpret_cast<uint2*>((uintptr_t)b.raw_ring);
      } else {
        void* pb = nullptr;
        SNN_CUDA_OK(cudaIpcOpenMemHandle(&pb, *reinterpret_cast<const cudaIpcMemHandle_t*>(b.ipc_ring), cudaIpcMemLazyEnablePeerAccess));
        ipc_opened_.push_back(pb);
        p.peer_ll[r] = reinterpret_cast<uint2*>(pb);
      }
    }
    p.n_ranks = world; p.rank = rank;
    return SNN_OK;
  }
  void close_peers() {
    for (void* q : ipc_opened_) cudaIpcCloseMemHandle(q);
    ipc_opened_.clear();
    p_.n_ranks = 0; p_.rank = 0;
    for (int r = 0; r < kMaxRanks; ++r) p_.peer_ll[r] = nullptr;
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
    // weights go to both buffers (non-plastic ones are read from either); sd lives in the current one
    wsd_from_caller<T><<<grid_for(p.E), 256, 0, stream_>>>(d, topo_.perm, p.wsdX[cur_], p.E, comp, zero_other);
    if (p.wsdX[1] != p.wsdX[0] && comp == 0)
      wsd_from_caller<T><<<grid_for(p.E), 256, 0, stream_>>>(d, topo_.perm, p.wsdX[cur_ ^ 1], p.E, comp, zero_other);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    cudaFree(d);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return SNN_ERR_CUDA; }
    return SNN_OK;
  }
  struct CachedGraph {
    std::string key; cudaGraphExec_t exec = nullptr; size_t prof_used = 0; std::vector<int> prof_cls;
    int pass_launches = 0;
    std::vector<cudaEvent_t> prof_events;  // event record nodes are baked into the graph
  };
  void drop_graph() {
    for (auto& g : graphs_) {
      if (g.exec) cudaGraphExecDestroy(g.exec);
      for (auto& e : g.prof_events) cudaEventDestroy(e);
    }
    graphs_.clear();
  }
  void free_synapse_state() {
    DevNet<T>& p = p_;
    if (p.wsdX[1] != p.wsdX[0]) cudaFree(p.wsdX[1]);
    cudaFree(p.wsdX[0]); cudaFree(p.dlog); cudaFree(p.dsd); cudaFree(p.log_cnt); cudaFree(p.tile_done);
    cudaFree(p.cacc); cudaFree(p.lzH); cudaFree(p.lzK);
    cudaFree(const_cast<uint32_t*>(p.inst_lo));
    p.wsd = p.wsdX[0] = p.wsdX[1] = nullptr; p.dlog = nullptr; p.dsd = nullptr; p.log_cnt = nullptr;
    p.cacc = nullptr; p.lzH = p.lzR = p.lzInv = nullptr; p.lzK = nullptr; p.lazy = 0;
    p.tile_done = nullptr; p.inst_lo = p.inst_hi = nullptr; p.log_regions = p.log_region_cap = p.n_tiles = 0;
  }
  void release() {
    DevNet<T>& p = p_;
    drop_graph();
    cudaFree(p.v); cudaFree(p.u); cudaFree(p.ho); cudaFree(p.Iacc); cudaFree(p.trace); cudaFree(p.bits);
    cudaFree(p.spk_list); cudaFree(p.spk_cnt); cudaFree(p.da); cudaFree(p.counters); cudaFree(p.blk_ev);
    cudaFree(p.blk_ltp); free_synapse_state(); cudaFree(d_ctx_); cudaFree(p.tl); cudaFree(p.lc); cudaFree(p.lg);
    close_peers();
    cudaFree(p.ll); cudaFree(p.xdone);
    if (h_ctx_) cudaFreeHost(h_ctx_);
    if (h_blk_) cudaFreeHost(h_blk_);
    b_noise_.release(); b_noise64_.release(); b_ev_off_.release(); b_ev_inst_.release(); b_ev_amt_.release();
    sort_scratch_.release(); b_probe_want_.release(); b_probe_pos_.release(); b_probe_out_.release();
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
    if (chain_stream_) cudaStreamDestroy(chain_stream_);
    if (learn_stream_) cudaStreamDestroy(learn_stream_);
    side2_stream_ = chain_stream_ = learn_stream_ = nullptr;
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
      // reserved[7]: side-kernel footprint, tens = synapse blocks per SM, units = ltp blocks per SM
      // default 6 + 2: the side kernels are persistent (their blocks live as long as the kernel), so
      // whatever they hold is not available to the next neuron kernel when it becomes ready
      const int syn_per_sm = cfg_.reserved[7] >= 10 ? cfg_.reserved[7] / 10 : 6;
      gA_ = std::min(kMaxSynBlocks, n_sm_ * std::max(1, std::min(per_sm, syn_per_sm)));
    }
    l.gA = gA_;
    l.gP = std::min(kMaxSynBlocks, n_sm_ * ((cfg_.reserved[7] % 10) > 0 ? cfg_.reserved[7] % 10 : 2));
    int64_t plastic_per_inst = p.B ? (p.E / p.B) : 0;
    int gLx = (int)std::max<int64_t>(1, std::min<int64_t>((plastic_per_inst / 2 + 256 * 4 - 1) / (256 * 4),
                                                           std::max<int64_t>(1, (int64_t)n_sm_ * 8 / p.B)));
    l.gL = dim3(gLx, p.B);
    return l;
  }
  // event-driven learning: learn step k (or, force = 1, materialise (w, sd) for the host)
  void launch_lazy(const DevNet<T>& p, const Launch& g, cudaStream_t s, int k, int dl, int dd, int force) {
    lazy_learn_kernel<T><<<g.gL, 256, 0, s>>>(p, d_ctx_, k, dl, dd, force);
    lazy_commit_kernel<T><<<(p.B + 127) / 128, 128, 0, s>>>(p, d_ctx_, k, dl, dd, force);
  }
  // materialise and switch wsdX[cur_] to the API layout (w, sd) / back to the event-driven layout
  int lazy_to_api(std::string& err) {
    if (!p_.lazy || !lazy_layout_) return SNN_OK;
    DevNet<T> p = p_;
    p.static_cur = cur_;
    launch_lazy(p, grids(), stream_, 0, 0, 0, 1);
    lazy_layout_kernel<T><<<grid_for(p.E), 256, 0, stream_>>>(p.wsdX[cur_], p.cacc, p.E, 0);
    SNN_CUDA_OK(cudaStreamSynchronize(stream_));
    lazy_layout_ = false;
    return SNN_OK;
  }
  int lazy_from_api(std::string& err) {
    if (!p_.lazy || lazy_layout_ || p_.E == 0) return SNN_OK;
    lazy_layout_kernel<T><<<grid_for(p_.E), 256, 0, stream_>>>(p_.wsdX[cur_], p_.cacc, p_.E, 1);
    SNN_CUDA_OK(cudaStreamSynchronize(stream_));
    lazy_layout_ = true;
    return SNN_OK;
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
  bool async_ok(const StepArgs& a) const;
  void launch_pass(cudaStream_t s4, int k, int do_decay, bool prof);

  snn_config cfg_;
  DevNet<T> p_;
  Topo topo_;
  bool have_topo_ = false, use_graph_ = true;
  bool lazy_layout_ = false;  // wsdX[cur_] / cacc hold the event-driven layout (DevNet::lazy)
  int profiling_ = 0;
  int64_t t_ = 0;
  int n_sm_ = 148, g_arr_log2_ = 5, g_ltp_log2_ = 5, gA_ = 0;
  cudaStream_t own_stream_ = nullptr, side_stream_ = nullptr, stream_ = nullptr;
  cudaEvent_t ev_begin_ = nullptr, ev_end_ = nullptr;
  cudaEvent_t dep_ev_[16] = {};
  cudaStream_t side2_stream_ = nullptr, chain_stream_ = nullptr, learn_stream_ = nullptr;
  // asynchronous learn pass
  bool have_async_ = false, pass_attr_set_ = false;
  int cur_ = 0;                  // host mirror of LearnCtl.state >> 31 (read back after every window)
  int64_t n_plastic_ = 0;
  std::vector<uint32_t> inst_lo_, inst_hi_;
  int last_pass_launches_ = 0;   // kernels of the async passes of the last window (chunks + control)
  std::vector<cudaEvent_t> prof_events_;
  std::vector<int> prof_cls_;
  size_t prof_used_ = 0;
  WindowCtx<T>* d_ctx_ = nullptr; WindowCtx<T>* h_ctx_ = nullptr;
  unsigned long long* h_blk_ = nullptr;
  DevBuf b_noise_, b_noise64_, b_ev_off_, b_ev_inst_, b_ev_amt_, b_force_, b_rec_, b_rec_off_, b_cur_idx_, b_cur_val_;
  SortScratch sort_scratch_;
  DevBuf b_probe_want_, b_probe_pos_, b_probe_out_;
  std::vector<void*> ipc_opened_;
  snn_exchange_fn exchange_ = nullptr; void* exchange_user_ = nullptr; int exchange_failed_ = 0;
  const std::vector<cudaEvent_t>* graph_prof_events_ = nullptr;
  std::vector<CachedGraph> graphs_;  // a few (length, train pattern, profiling) variants
};

// Plain stream-ordered launches: exact mode, forced spikes, per-kernel profiling.
template <typename T>
void Engine<T>::launch_serial(const StepArgs& a, const std::vector<int32_t>& force_off, const int32_t* d_force,
                              const std::vector<int32_t>& cur_off, const int32_t* d_cur_idx, const double* d_cur_val,
                              bool prof) {
  DevNet<T> p = p_;          // by-value copy handed to the kernels
  p.static_cur = cur_;       // plain launches: learn! runs in place, no pass is ever in flight
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
      remote_kernel<T><<<g.gR, 256, 0, s>>>(p, d_ctx_, k, 0);
    } else if (k < n && p.n_ranks > 1) {
      // peer-to-peer exchange: neuron_kernel has stored this rank's words into the peers' rings
      remote_kernel<T><<<g.gR, 256, 0, s>>>(p, d_ctx_, k, 1);
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
      if (p.lazy) launch_lazy(p, g, s, k, dl, dd, 0);
      else learn_kernel<T><<<g.gL, 256, 0, s>>>(p, d_ctx_, k, dl, dd);
      if (prof) prof_end(s);
      // the pushes of step k+1 must see the weights after learn!(k)
      if (push_next) synapse_kernel<T, false, true, true><<<g.gA, 256, 0, s>>>(p, d_ctx_, k, g_arr_log2_);
    }
    if (a.n_probes > 0)
      probe_kernel<T><<<(a.n_probes + 63) / 64, 64, 0, s>>>(p, b_probe_pos_.as<uint32_t>(), a.n_probes, b_probe_out_.as<double>(), k);
  }
}

// Whole window as one CUDA graph (captured once per (length, train pattern)).
//   needs(neuron(k))  : neuron(k-1) [stream], pushes for step k-1 = synapse(k-2) | push(k-2),
//                       learn(k-1) begun (async) / done (in place): its delay-1 pushes read w
//   needs(synapse(k)) : neuron(k), synapse(k-1) [stream]        (sd updates are reductions:
//   needs(ltp(k))     : neuron(k), ltp(k-1) [stream]             no LTD/LTP ordering)
//   needs(learn(k))   : synapse(k) LTD, ltp(k); async: the previous pass [stream] + its log replay
//   needs(push(k))    : learn(k) begun / done   (learn steps only; otherwise fused into synapse(k))
// Streams: sN neuron chain (highest priority), s2 synapse, s3 ltp (+ in-place learn), s4 the
// asynchronous learn pass (lowest priority): learn_replay, learn_ctl(begin), chunks, learn_ctl(all).
template <typename T>
bool Engine<T>::async_ok(const StepArgs& a) const {
  if (p_.lazy) return false;  // event-driven learning: the (rare) dense pass runs in place
  if (!have_async_) return false;
  int last = -1000, n_learn = 0, dl, dd;
  for (int k = 0; k < a.n_steps; ++k) {
    if (!wants_learn(a, k, &dl, &dd)) continue;
    if (k - last < 3) return false;  // passes back to back: nothing to overlap with
    last = k; ++n_learn;
  }
  if (n_learn == 0) return false;
  if (cfg_.reserved[4] == 2) return true;               // forced (tests on small nets)
  return n_plastic_ * 2 * (int64_t)sizeof(typename DevNet<T>::T2) >= (32ll << 20);
}

template <typename T>
void Engine<T>::launch_pass(cudaStream_t s4, int k, int do_decay, bool prof) {
  DevNet<T>& p = p_;
  // persistent pass: LB blocks per SM stay resident for the whole pass and leave the other
  // thread slots to the step kernels (LB x 256 threads x 8 x 16 B in flight per SM)
  const bool lsu = (cfg_.reserved[1] & 512) != 0;  // experiment 512: register-staged (LSU) pass instead of TMA
  const int LB = cfg_.reserved[2] > 0 ? cfg_.reserved[2] : (lsu ? 2 : 3);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)n_sm_ * LB, p.n_tiles));
  if (prof) prof_begin(2, s4, true);
  if (lsu) {
    learn_pass_kernel<T><<<grid, 256, 0, s4>>>(p, k, do_decay);
  } else {
    const int S = cfg_.reserved[6] >= 2 && cfg_.reserved[6] <= kPassMaxStages ? cfg_.reserved[6] : kPassStages;
    const size_t smem = (size_t)S * kPassChunkBytes;
    if (!pass_attr_set_) {  // per engine: the attribute belongs to the (function, device) pair
      cudaFuncSetAttribute(learn_pass_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           kPassMaxStages * (int)kPassChunkBytes);
      pass_attr_set_ = true;
    }
    learn_pass_tma_kernel<T><<<grid, kPassThreads, smem, s4>>>(p, k, do_decay, S);
  }
  learn_ctl_kernel<T><<<1, 32, 0, s4>>>(p, d_ctx_, k, kCtlAll);
  last_pass_launches_ += 2;
  if (prof) prof_end(s4, true);
}

template <typename T>
int Engine<T>::launch_graph(const StepArgs& a, bool prof, std::string& err) {
  const int n = a.n_steps;
  const bool async = async_ok(a);
  DevNet<T> p = p_;                  // by-value copy baked into the captured kernel nodes
  p.static_cur = async ? -1 : cur_;  // in-place windows: the (w, sd) view is fixed, skip the LearnCtl load
  std::string key(reinterpret_cast<const char*>(&n), sizeof(n));
  key.push_back(prof ? 'p' : '-');
  key.push_back(async ? 'a' : (cur_ ? 't' : 's'));
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
    cudaStream_t s0 = stream_, sN = chain_stream_, s2 = side_stream_, s3 = side2_stream_, s4 = learn_stream_;
    cudaGraph_t graph = nullptr;
    prof_used_ = 0; prof_cls_.clear();
    last_pass_launches_ = 0;
    if (prof) prof_reserve(2 * (size_t)n + 2);  // no allocation while capturing
    // events: [0] neuron done, [1..3] pushes-for-step ready (ring of 3), [4..5] synapse LTD done
    // (ring of 2), [6..7] ltp done (ring of 2), [8] learn done / begun, [9] fork, [10] pass+replay done
    cudaEvent_t evL = dep_ev_[8], evFork = dep_ev_[9], evFin = dep_ev_[10];
    auto evPush = [&](int step) { return dep_ev_[1 + ((step % 3) + 3) % 3]; };  // pushes consumed by Euler(step)
    auto evS = [&](int step) { return dep_ev_[4 + (step & 1)]; };
    auto evP = [&](int step) { return dep_ev_[12 + (step & 3)]; };  // ltp done (ring of 4)
    SNN_CUDA_OK(cudaStreamBeginCapture(s0, cudaStreamCaptureModeThreadLocal));
    cudaEventRecord(evFork, s0);                    // fork point
    cudaStreamWaitEvent(sN, evFork, 0);
    cudaStreamWaitEvent(s2, evFork, 0);
    synapse_kernel<T, false, true, true><<<g.gA, 256, 0, s2>>>(p, d_ctx_, -1, g_arr_log2_);  // pushes for step 0
    cudaEventRecord(evPush(0), s2);
    bool s3_used = false, pass_open = false;
    cudaEvent_t evNk[2] = {dep_ev_[0], dep_ev_[11]};
    // neuron blocks of 64 threads (4 K registers each): they slip into an SM as soon as any side block
    // retires instead of waiting for 16 K registers to come free (+0.8 %; reserved[1] & 8: 256 threads)
    const int nbm = (cfg_.reserved[1] & 8) ? 1 : 4;
    auto cap_neuron = [&](int k) {
      int dl, dd;
      const bool learn_prev = k >= 1 && wants_learn(a, k - 1, &dl, &dd);
      if (k >= 1) cudaStreamWaitEvent(sN, evPush(k - 1), 0);  // Euler(k-1) consumes them
      if (learn_prev) cudaStreamWaitEvent(sN, evL, 0);
      if (p.n_ranks > 1) {  // neuron-partitioned: local detection + replay of the peers' detection, one node
        if (k == 0) neuron_remote_kernel<T, false><<<g.gN + g.gR, 256, 0, sN>>>(p, d_ctx_, k, 1, (uint32_t)g.gN);
        else neuron_remote_kernel<T, true><<<g.gN + g.gR, 256, 0, sN>>>(p, d_ctx_, k, 1, (uint32_t)g.gN);
      } else if (k == 0) {
        neuron_kernel<T, false, true><<<g.gN * nbm, 256 / nbm, 0, sN>>>(p, d_ctx_, k, 1);
      } else {
        neuron_kernel<T, true, true><<<g.gN * nbm, 256 / nbm, 0, sN>>>(p, d_ctx_, k, 1);
      }
      cudaEventRecord(evNk[k & 1], sN);
    };
    auto cap_side = [&](int k) {
      int dl, dd;
      const bool learn = wants_learn(a, k, &dl, &dd);
      const bool learn_prev = k >= 1 && wants_learn(a, k - 1, &dl, &dd);
      wants_learn(a, k, &dl, &dd);
      const bool push_next = k + 1 < n;
      // s2: synapse(k).  It also keeps ltp from falling behind: neuron(m) overwrites the trace-ring
      // slot of step m-R (R = Dmax+2), which ltp(j) still reads for j <= m-3.  neuron(m) waits for the
      // pushes of synapse(m-2), so making synapse(k) wait for ltp(k-1) closes the chain.  (Without this
      // edge a delayed ltp stream read pre traces that had already been overwritten — seen once in ~10
      // runs as one synapse with a spurious LTP term.)
      if (k >= 1) cudaStreamWaitEvent(s2, evP(k - 1), 0);
      cudaStreamWaitEvent(s2, evNk[k & 1], 0);
      if (learn_prev && !async) cudaStreamWaitEvent(s2, evL, 0);  // sd reductions must not race the in-place pass
      if (push_next && !learn) {
        synapse_kernel<T, true, true, true><<<g.gA, 256, 0, s2>>>(p, d_ctx_, k, g_arr_log2_);
        cudaEventRecord(evPush(k + 1), s2);
      } else {
        synapse_kernel<T, true, false, true><<<g.gA, 256, 0, s2>>>(p, d_ctx_, k, g_arr_log2_);
      }
      cudaEventRecord(evS(k), s2);
      // s3: ltp(k) [-> in-place learn(k)]
      cudaStreamWaitEvent(s3, evNk[k & 1], 0);
      ltp_kernel<T, true><<<g.gP, 256, 0, s3>>>(p, d_ctx_, k, g_ltp_log2_);
      s3_used = true;
      cudaEventRecord(evP(k), s3);
      if (learn && async) {
        // s4: [previous pass, stream order] -> replay its log -> begin -> pass -> frontier = all
        cudaStreamWaitEvent(s4, evS(k), 0);
        cudaStreamWaitEvent(s4, evP(k), 0);
        if (pass_open) learn_replay_kernel<T><<<n_sm_ * 2, 256, 0, s4>>>(p, d_ctx_, k, 1);  // replay + begin
        else learn_ctl_kernel<T><<<1, 256, 0, s4>>>(p, d_ctx_, k, kCtlBegin);
        ++last_pass_launches_;
        cudaEventRecord(evL, s4);
        launch_pass(s4, k, dd, prof);
        if (cfg_.reserved[1] & 4) cudaEventRecord(evL, s4);  // experiment: nothing overlaps the pass
        pass_open = true;
        if (push_next) {
          cudaStreamWaitEvent(s2, evL, 0);
          synapse_kernel<T, false, true, true><<<g.gA, 256, 0, s2>>>(p, d_ctx_, k, g_arr_log2_);
          cudaEventRecord(evPush(k + 1), s2);
        }
      } else if (learn) {
        cudaStreamWaitEvent(s3, evS(k), 0);
        if (prof) prof_begin(2, s3, true);
        if (p.lazy) launch_lazy(p, g, s3, k, dl, dd, 0);
        else learn_kernel<T><<<g.gL, 256, 0, s3>>>(p, d_ctx_, k, dl, dd);
        if (prof) prof_end(s3, true);
        cudaEventRecord(evL, s3);
        if (push_next) {
          cudaStreamWaitEvent(s2, evL, 0);
          synapse_kernel<T, false, true, true><<<g.gA, 256, 0, s2>>>(p, d_ctx_, k, g_arr_log2_);
          cudaEventRecord(evPush(k + 1), s2);
        }
        cudaEventRecord(evP(k), s3);
      }
    };
    // Capture order = launch order of sibling nodes: the next neuron kernel (the only serial chain)
    // goes in before the side kernels of the step it does not depend on.
    cap_neuron(0);
    for (int k = 0; k < n; ++k) {
      int dl, dd;
      const bool learn = wants_learn(a, k, &dl, &dd);
      if (k + 1 < n && !learn) { cap_neuron(k + 1); cap_side(k); }
      else { cap_side(k); if (k + 1 < n) cap_neuron(k + 1); }
    }
    cudaEvent_t evN = evNk[0];
    // join everything, then the window-end Euler
    cudaEventRecord(evS(n), s2);
    cudaStreamWaitEvent(sN, evS(n), 0);
    if (s3_used) { cudaEventRecord(evP(n), s3); cudaStreamWaitEvent(sN, evP(n), 0); }
    if (pass_open) {
      // the pass still in flight ends with the window: replay its log onto the post-learn sd
      cudaStreamWaitEvent(s4, evS(n), 0);
      if (s3_used) cudaStreamWaitEvent(s4, evP(n), 0);
      learn_replay_kernel<T><<<n_sm_ * 2, 256, 0, s4>>>(p, d_ctx_, n, 0);
      learn_spill_reset_kernel<T><<<1, 1, 0, s4>>>(p);
      last_pass_launches_ += 2;
      cudaEventRecord(evFin, s4);
      cudaStreamWaitEvent(sN, evFin, 0);
    }
    neuron_kernel<T, true, false><<<g.gN, 256, 0, sN>>>(p, d_ctx_, n, 1);
    cudaEventRecord(evN, sN);
    cudaStreamWaitEvent(s0, evN, 0);
    cudaError_t e = cudaStreamEndCapture(s0, &graph);
    if (e != cudaSuccess || !graph) {
      for (auto& ev : prof_events_) cudaEventDestroy(ev);
      prof_events_.clear();
      prof_events_.swap(saved_events);
      err = std::string("graph capture: ") + cudaGetErrorString(e); return SNN_ERR_CUDA;
    }
    CachedGraph cg;
    cg.key = key; cg.prof_used = prof_used_; cg.prof_cls = prof_cls_; cg.pass_launches = last_pass_launches_;
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
  last_pass_launches_ = hit->pass_launches;
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
  if (!(p.part_lo == 0 && p.part_hi == p.N) && !exchange_ && p.n_ranks < 2) {
    err = "a partial neuron partition needs an exchange: snn_partition_connect or an exchange callback"; return SNN_ERR_STATE;
  }

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
  if (p.tl) {
    SNN_CUDA_OK(cudaMemsetAsync(p.tl, 0xFF, kTlClasses * kTlSteps * sizeof(unsigned long long), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.tl + kTlClasses * kTlSteps, 0, kTlClasses * kTlSteps * sizeof(unsigned long long), stream_));
  }

  // ---- the window
  if (a.n_probes > 0 && p_.lazy) { err = "synapse probes are not available with event-driven learning"; return SNN_ERR_INVALID; }
  if (a.n_probes > 0) {  // per-step probes: plain launches, learn! in place (diagnostics path)
    if (a.n_probes > 4096 || !a.probe_syn || !a.probe_out) { err = "probes: 1..4096 synapse indices and an output buffer"; return SNN_ERR_INVALID; }
    for (int q = 0; q < a.n_probes; ++q)
      if (a.probe_syn[q] < 0 || a.probe_syn[q] >= p.E) { err = "probe synapse index out of range"; return SNN_ERR_INVALID; }
    SNN_CUDA_OK(b_probe_want_.reserve(a.n_probes * sizeof(int64_t)));
    SNN_CUDA_OK(b_probe_pos_.reserve(a.n_probes * sizeof(uint32_t)));
    SNN_CUDA_OK(b_probe_out_.reserve((size_t)n * a.n_probes * 2 * sizeof(double)));
    SNN_CUDA_OK(cudaMemcpyAsync(b_probe_want_.ptr, a.probe_syn, a.n_probes * sizeof(int64_t), cudaMemcpyHostToDevice, stream_));
    probe_locate_kernel<<<grid_for(p.E), 256, 0, stream_>>>(topo_.perm, p.E, b_probe_want_.as<int64_t>(), a.n_probes,
                                                            b_probe_pos_.as<uint32_t>());
  }
  const bool graphable = use_graph_ && !p.exact && a.n_force == 0 && a.n_cur == 0 && profiling_ != 1 && !exchange_ &&
                         a.n_probes == 0;
  last_pass_launches_ = 0;
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
    int src = sort_segments(b_rec_.as<int32_t>(), total, b_rec_off_.as<int32_t>(), n, stream_, &sort_scratch_, err);
    if (src) return src;
    if (total > 0) SNN_CUDA_OK(cudaMemcpyAsync(a.spikes_out, b_rec_.ptr, total * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
    for (int k = 0; k <= n; ++k) a.spike_offsets[k] = h_off[k];
  }
  if (a.n_probes > 0)
    SNN_CUDA_OK(cudaMemcpyAsync(a.probe_out, b_probe_out_.ptr, (size_t)n * a.n_probes * 2 * sizeof(double),
                                cudaMemcpyDeviceToHost, stream_));
  LearnCtl lc_h;
  uint32_t xdone_h[2] = {0, 0};
  SNN_CUDA_OK(cudaMemcpyAsync(&lc_h, p.lc, sizeof(lc_h), cudaMemcpyDeviceToHost, stream_));
  SNN_CUDA_OK(cudaMemcpyAsync(xdone_h, p.xdone, sizeof(xdone_h), cudaMemcpyDeviceToHost, stream_));
  SNN_CUDA_OK(cudaMemcpyAsync(c_after, p.counters, sizeof(c_after), cudaMemcpyDeviceToHost, stream_));
  SNN_CUDA_OK(cudaMemcpyAsync(h_blk_, p.blk_ev, kMaxSynBlocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream_));
  SNN_CUDA_OK(cudaMemcpyAsync(h_blk_ + kMaxSynBlocks, p.blk_ltp, kMaxSynBlocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream_));
  SNN_CUDA_OK(cudaEventRecord(ev_end_, stream_));
  SNN_CUDA_OK(cudaStreamSynchronize(stream_));
  cur_ = (int)(lc_h.state >> 31);
  if (xdone_h[1]) {
    cudaMemsetAsync(p.xdone, 0, 2 * sizeof(uint32_t), stream_);
    err = "peer exchange timed out: a rank of the partition did not publish its spikes"; return SNN_ERR_STATE;
  }
  if ((lc_h.state & kFrontAll) != kFrontAll) {
    err = "internal: learn pass still open at the end of the window"; return SNN_ERR_STATE;
  }
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
      s.reserved = last_pass_launches_;  // async pass: chunk + control kernels (learn_launches counts passes)
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
