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
#include <stdlib.h>
#include <stdio.h>

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

// x += v as a fire-and-forget reduction (red.global: no value comes back, the warp does not wait for L2).  Written as
// PTX because inside the persistent kernels — loops around acquire loads and fences — nvcc keeps atomicAdd() with an
// unused result as a RETURNING atomic (ATOMG), which stalls the issuing warp for a full L2 round trip per update.
__device__ __forceinline__ void red_add(float* q, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(q), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add(double* q, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(q), "d"(v) : "memory"); }
__device__ __forceinline__ void red_add(uint32_t* q, uint32_t v) { asm volatile("red.global.add.u32 [%0], %1;" ::"l"(q), "r"(v) : "memory"); }

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
constexpr int32_t kNever = -(1 << 30);  // "no spike yet" in hist / old_last
// How far the neuron chain may run ahead of the slowest reader of spike history (ltp): neuron(m) waits for the
// pushes of synapse(m-2), which waited for ltp(m-3).  Ring depth R = Dmax + 2*kLag + 1 and retire age
// G = Dmax + kLag (+1 each for variant_g) make every slot a reader may scan valid and every older spike
// already retired into old_last, whatever the interleaving inside that bound.
constexpr int kLag = 3;
// The persistent chain pushes currents up to kMaxLook steps ahead of their arrival (Engine::launch_chain): Iacc is a
// ring of kIaccRing buffers there (the other schedulers use the first two, indexed by step parity).
constexpr int kMaxLook = 2, kIaccRing = 4;  // (a lookahead of 3 was measured 3 % slower than 1; every step of it costs 1 KB of shared memory per wave block)
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
  T* Iacc;           // [kIaccRing][Nld]  Iacc[t&1] (chain: Iacc[t % ring]) = currents consumed by the Euler update of step t
  // STDP traces are not stored: between two spikes of a neuron the reference's trace is the FIXED sequence
  // set, set*decay, (set*decay)*decay, ... (new_net.jl:123,130), so the trace of neuron i after the set of step
  // s is tseq[s - (last spike of i at or before s)], 0 before the first spike — bit for bit what the iterated
  // multiplication gives.  State per neuron: the last two spike steps (one 8-byte word, rewritten only when
  // the neuron spikes) and, for the rare neuron that spiked more than twice inside the delay window, the
  // bitmap ring + the last spike that has left it.  (The dense trace ring this replaces cost 8 B per neuron
  // and step in the neuron kernel, made LTP gather from an 88 MB array, and had to be replayed for every
  // foreign neuron of a partitioned run.)
  int2* hist;            // [Nld] (last, prev) spike step, kNever if none
  int32_t* old_last;     // [Nld] last spike step <= t - G (retired from the bitmap ring G steps after it happened)
  const T* tseq;         // [tseq_len] set * decay^n by iterated multiplication, up to its fixed point
  int32_t tseq_len, G;   // G = retire age, see kLag
  uint32_t* bits;    // [R][W]    spike bitmap per step
  int32_t* spk_list; // [R][Nld]  spike list per step (unordered)
  int32_t* spk_cnt;  // [R]
  T* da;             // [2][B]    da[t&1] = dopamine after the decay of step t-1
  // synapses (internal order: ascending pre, delay, caller position)
  const uint4* segL;     // [N] bounds of the delay-1, -2 and -3 segments of every row (seg[i][0..3]): the 16 MB the neuron
                         // role needs of the 84 MB table, L2-resident
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
  int32_t exp_flags;  // diagnostics (Options): 2 = kernel timeline, 4 = nothing overlaps the learn pass, 8 = 256-thread neuron blocks, 512 = LSU pass
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

// timeline classes: 0 neuron, 1 synapse, 2 ltp, 3 learn, 4 remote (one GPU: the learn role's early replay), 5 push; phases
// of the persistent chain's roles: 6 neuron wait, 7 currents (phase 1), 8 Euler + detect (phase 2), 9 spikers (phase 3),
// 10 the fence + signal of the hand-over (timeline = 2 only; 5 then is retire + block barrier),
// 11 / 12 synapse wave stage A / B (timeline = 2: the two halves of the neuron role's spiker wave), 13 - 15 per-SM diagnostics
constexpr int kTlSteps = 258, kTlClasses = 16;
__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void tl_begin(unsigned long long* tl, int cls, int k) {
  if (tl && threadIdx.x == 0 && k + 1 >= 0 && k + 1 < kTlSteps)
    asm volatile("red.global.min.u64 [%0], %1;" ::"l"(&tl[cls * kTlSteps + k + 1]), "l"(gtime_ns()) : "memory");
}
__device__ __forceinline__ void tl_end(unsigned long long* tl, int cls, int k) {
  if (tl && threadIdx.x == 0 && k + 1 >= 0 && k + 1 < kTlSteps)
    asm volatile("red.global.max.u64 [%0], %1;" ::"l"(&tl[kTlClasses * kTlSteps + cls * kTlSteps + k + 1]), "l"(gtime_ns()) : "memory");
}

__device__ __forceinline__ uint32_t smid() { uint32_t r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
// timeline diagnostics of the persistent chain: which SM hosts how many blocks of a role (end half of class `cls`,
// indexed by SM), and how long the neuron block of every SM needs for two chosen steps (class 15: begin half = step
// kTlProbeA, end half = step kTlProbeB)
constexpr int kTlProbeA = 27, kTlProbeB = 33;
__device__ __forceinline__ void tl_count_sm(unsigned long long* tl, int cls) {
  if (tl && threadIdx.x == 0 && smid() < (uint32_t)kTlSteps) atomicAdd(&tl[kTlClasses * kTlSteps + cls * kTlSteps + smid()], 1ull);
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
// Spike history -> STDP trace (see DevNet::hist).  `t` is the step the calling kernel works on, `s` the
// step whose trace is wanted (t - Dmax <= s <= t).  Two phases so that callers keep several gathers
// in flight: hist_fetch issues the 8-byte load (L2: the kernel that writes hist runs concurrently),
// trace_resolve finishes.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int2 hist_fetch(const int2* __restrict__ hist, uint32_t i) { return __ldcg(hist + i); }
// more than two spikes of i after step s (the neuron chain is up to kLag steps ahead): walk the bitmap
// ring back from s; anything older than the ring's guaranteed part has been retired into old_last
// (arguments by value: a reference to the by-value kernel parameter DevNet would force the whole struct
// into a local-memory copy in every kernel that can reach this call)
static __device__ __noinline__ int32_t last_spike_slow(const uint32_t* __restrict__ bits, const int32_t* __restrict__ old_last,
                                                       int R, int W, int G, uint32_t i, int32_t s, int32_t t) {
  const int32_t x_lo = t - G + 1;
  for (int32_t x = s; x >= x_lo && x >= 0; --x) {
    const uint32_t word = __ldcg(bits + (size_t)(x % R) * W + (i >> 5));
    if ((word >> (i & 31)) & 1u) return x;
  }
  return __ldcg(old_last + i);
}
template <typename T>
__device__ __forceinline__ T trace_resolve(const DevNet<T>& p, const int2 h, uint32_t i, int32_t s, int32_t t) {
  const int32_t ls = h.x <= s ? h.x : (h.y <= s ? h.y : last_spike_slow(p.bits, p.old_last, p.R, p.W, p.G, i, s, t));
  if (ls == kNever) return (T)0;
  const int32_t n = s - ls;
  return __ldg(p.tseq + (n < p.tseq_len ? n : p.tseq_len - 1));
}
template <typename T>
__device__ __forceinline__ T trace_at(const DevNet<T>& p, uint32_t i, int32_t s, int32_t t) {
  return trace_resolve<T>(p, hist_fetch(p.hist, i), i, s, t);
}

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
                     : Ops<T>::add(ws.x, Ops<T>::mul(__ldcg(p.lg + inst), ws.y));
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
      red_add(&vw.cur[e].x, d);
      red_add(&vw.cur[e].y, dh);
    }
    return;
  }
  if ((e >> 5) < vw.front) {
    red_add(&vw.cur[e].y, delta);
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
    red_add(&p.dsd[e], delta);
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
// The spikes of step t - G leave the part of the bitmap ring a reader may rely on: remember them per neuron
// (detect(t) of every step, all threads of the neuron grid share the ~N*rate entries of that step's list).
template <typename T>
__device__ __forceinline__ void retire_spikes(const DevNet<T>& p, int64_t t, uint32_t gtid, uint32_t nth) {
  const int64_t xr = t - p.G;
  if (xr < 0) return;
  const int rs = (int)(xr % p.R);
  const int cnt = __ldcg(p.spk_cnt + rs);
  const int32_t* __restrict__ rl = p.spk_list + (size_t)rs * p.Nld;
  for (uint32_t q = gtid; q < (uint32_t)cnt; q += nth) p.old_last[__ldcg(rl + q)] = (int32_t)xr;
}

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
    Vec4<T> v = ld4(p.v + j0), u = ld4(p.u + j0), I, ho;
    if (EULER) I = ld4(I_prev + j0);
    if (DETECT && p.homeostasis) ho = ld4(p.ho + j0);
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
      for (int c = 0; c < 4; ++c) {  // selects, no branches: four independent dependency chains the compiler interleaves
        T vv = v.x[c];
        // v = v + 0.5*((0.04*v+5)*v + 140 - u + I), twice (new_net.jl:120-121)
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          T x = O::add(O::mul((T)0.04, vv), (T)5);
          x = O::add(O::mul(x, vv), (T)140);
          x = O::add(O::sub(x, u.x[c]), I.x[c]);
          vv = O::add(vv, O::mul((T)0.5, x));
        }
        // u = u + a*(0.2*v - u) (new_net.jl:122)
        const T un = O::add(u.x[c], O::mul(exc[c] ? p.a_exc : p.a_inh, O::sub(O::mul((T)0.2, vv), u.x[c])));
        v.x[c] = act[c] ? vv : v.x[c];
        u.x[c] = act[c] ? un : u.x[c];
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
            p.hist[j0 + c] = make_int2((int32_t)t, __ldcg(p.hist + j0 + c).x);  // trace set (new_net.jl:123), see DevNet::hist
          }
        }
        if (p.homeostasis) {
          uint32_t l = l0 + c;
          if (l >= Nper) l -= Nper;
          T h = ho.x[c];
          if (s && (exc[c] || p.ho_inh) && (int64_t)l >= p.ho_from) h = O::add(h, p.theta);
          ho.x[c] = O::mul(h, p.ttheta);
        }
        nib |= (s ? 1u : 0u) << c;
      }
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
                      red_add(&I_now[pj[q4]],
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
  // (after the neurons: the list length it needs is a dependent load nobody should wait for up front)
  if (DETECT) retire_spikes<T>(p, t, bid * blockDim.x + threadIdx.x, nblk * blockDim.x);
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
    if (gtime_ns() - t_start > 10000000000ull) { p.xdone[1] = 1u; return 0u; }
    __nanosleep(64);
  }
}

// ---------------------------------------------------------------------------------------------
// remote_kernel(k) — neuron-partitioned runs only.  After the spike bitmap of step k has arrived from the
// peers, every rank replays detect(k) for the neurons it does NOT own.  The work is O(spikes), not
// O(neurons): one thread per FOREIGN BITMAP WORD (N/32 of them, ~4 % non-zero at C5) waits for the
// word, keeps `bits` complete, and for every set bit appends to the spike list / record, rewrites the
// neuron's spike history (the traces LTP reads are derived from it, see DevNet::hist) and pushes the
// spiker's delay-1 synapses into local targets.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void remote_body(const DevNet<T>& p, const WindowCtx<T>* __restrict__ cx, int32_t k, int32_t wait_peers,
                                            const uint32_t bid, const uint32_t nblk, const int nd = 1, const int ring = 2) {
  // nd / ring: the persistent chain pushes the delays 1..nd of a spiker here, into a ring of `ring` current buffers
  tl_begin(p.tl, 4, k);
  const bool from_ll = wait_peers != 0;
  const int lane = threadIdx.x & 31;
  const int64_t t = cx->t0 + k;
  const int slot = (cx->slot0 + k) % p.R, prev = wrap_slot(slot - 1, p.R);
  uint32_t* __restrict__ bits = p.bits + (size_t)slot * p.W;
  int32_t* __restrict__ list = p.spk_list + (size_t)slot * p.Nld;
  int32_t* __restrict__ rec_ids = cx->rec_ids;
  const int64_t rec_cap = cx->rec_cap;
  int32_t rec_base = 0;
  if (rec_ids && k > 0) {
    int64_t nxt = (int64_t)__ldcg(cx->rec_off + (k - 1)) + __ldcg(p.spk_cnt + prev);
    rec_base = (int32_t)(nxt < rec_cap ? nxt : rec_cap);
  }
  const WView<T> vw = wview(p);
  // partition bounds are multiples of 128 neurons = 4 words (the last rank's range is padded up to Nld)
  const uint32_t w_lo = (uint32_t)(p.part_lo >> 5), w_hi = (uint32_t)((p.part_hi >= p.N ? p.Nld : p.part_hi) >> 5);
  const uint32_t n_remote = (uint32_t)p.W - (w_hi - w_lo);
  const uint32_t tag = (uint32_t)(t + 1);
  for (uint32_t rb = bid * blockDim.x + (threadIdx.x & ~31u); rb < n_remote; rb += nblk * blockDim.x) {
    const uint32_t r = rb + lane;
    uint32_t word = 0, wi = 0;
    if (r < n_remote) {
      wi = r < w_lo ? r : r + (w_hi - w_lo);
      if (from_ll) {  // peer-to-peer exchange: wait for the owner's (word, step) pair, keep `bits` complete
        word = ll_word<T>(p, slot, wi, tag);
        bits[wi] = word;
      } else {
        word = __ldcg(bits + wi);
      }
    }
    const unsigned busy = __ballot_sync(0xffffffffu, word != 0u);
    if (!busy) continue;
    // list positions: warp-wide exclusive scan of the per-lane spike counts, one atomic per warp
    const int cnt = __popc(word);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 0) base = atomicAdd(&p.spk_cnt[slot], total);
    base = __shfl_sync(0xffffffffu, base, 0);
    int pos = base + incl - cnt;
    for (uint32_t wb = word; wb; wb &= wb - 1u) {
      const uint32_t i = (wi << 5) + (uint32_t)(__ffs(wb) - 1);
      list[pos] = (int32_t)i;
      if (rec_ids && (int64_t)rec_base + pos < rec_cap) rec_ids[rec_base + pos] = (int32_t)i;
      p.hist[i] = make_int2((int32_t)t, __ldcg(p.hist + i).x);
      ++pos;
    }
    if (!p.exact) {
      // delay-1 synapses of the remote spikers feed this very step's Euler update: warp-cooperative push
      for (unsigned mm = busy; mm; mm &= mm - 1u) {
        const int src = __ffs(mm) - 1;
        const uint32_t wsrc = __shfl_sync(0xffffffffu, word, src), isrc = __shfl_sync(0xffffffffu, wi, src) << 5;
        for (uint32_t wb = wsrc; wb; wb &= wb - 1u) {
          const uint32_t i = isrc + (uint32_t)(__ffs(wb) - 1);
          const uint4 sg = p.segL[i];
          const uint32_t s_end = nd == 1 ? sg.y : (nd == 2 ? sg.z : sg.w);
          const bool row_exc = i < (uint32_t)p.Ne;  // partitioned runs are single-instance
          for (uint32_t e = sg.x + lane; e < s_end; e += 32) {
            const uint32_t j = (uint32_t)p.post[e];
            bool fly;
            const typename DevNet<T>::T2 ws = w_fetch<T>(vw, e, row_exc, fly);
            const T wc = c_fetch<T>(p, vw, e, fly);
            const uint32_t dl = (e >= sg.y ? 1u : 0u) + (e >= sg.z ? 1u : 0u);  // delay - 1
            red_add(p.Iacc + (size_t)((t + dl) % ring) * p.Nld + j,
                      w_resolve<T>(p, ws, fly, p.plastic_ei || j < (uint32_t)p.Ne, 0u, wc));
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
__device__ __forceinline__ void synapse_body(const DevNet<T>& p, const WindowCtx<T>* __restrict__ cx, int32_t k, int g_log2,
                                             const uint32_t bid, const uint32_t nblk) {
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
    cum[threadIdx.x + 1] = __ldcg(p.spk_cnt + sl);
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
  const uint32_t gtid = bid * blockDim.x + threadIdx.x;
  const uint32_t nthreads = nblk * blockDim.x;
  const uint32_t Nper = (uint32_t)p.Nper, Ne = (uint32_t)p.Ne;
  const size_t Nld = (size_t)p.Nld;
  const int32_t t32 = (int32_t)t;
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
    const uint32_t i = (uint32_t)__ldcg(p.spk_list + (size_t)sl * Nld + (item - cum[a]));
    const uint32_t* __restrict__ sg = p.seg + (size_t)i * (p.Dmax + 1) + a;
    const uint32_t s0 = sg[0], s1 = sg[1];
    const uint32_t s2 = (PUSHNEXT && a + 1 < p.Dmax) ? sg[2] : s1;
    const uint32_t inst = p.B == 1 ? 0u : i / Nper, ibase = inst * Nper;
    const bool row_exc = (i - ibase) < Ne;
    const bool row_plastic = LTD && row_exc;
    // variant_g: the LTP term of a simultaneous post spike reads the pre trace of the step before
    const T pre_up = (row_plastic && !ATOMIC)
                         ? O::mul(p.apost, trace_at<T>(p, i, t32 - a - (p.variant_g ? 1 : 0), t32)) : (T)0;
    if (LTD && gl == 0) n_ev += (s1 - s0);
    for (uint32_t e = (LTD ? s0 : s1) + gl; e < s2; e += G) {
      const uint32_t j = (uint32_t)p.post[e];
      if (e >= s1) {                       // arrives at k+1: push only
        bool fly;
        const typename DevNet<T>::T2 ws = w_fetch<T>(vw, e, row_exc, fly);
        const T wc = c_fetch<T>(p, vw, e, fly);
        red_add(&I_next[j], w_resolve<T>(p, ws, fly, p.plastic_ei || (j - ibase) < Ne, inst, wc));
      } else if (row_plastic && (p.plastic_ei || (j - ibase) < Ne)) {
        const T dn = O::mul(p.apre, trace_at<T>(p, j, t32, t32));
        if (ATOMIC) {  // production: fire-and-forget reduction, LTP of this step comes from ltp_kernel
          sd_add<T>(p, vw, e, -dn, bid, inst);
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
    if (threadIdx.x == 0 && blk_ev) p.blk_ev[bid] += blk_ev;
  }
  tl_end(p.tl, LTD ? 1 : 5, k);
}
template <typename T, bool LTD, bool PUSHNEXT, bool ATOMIC>
__global__ void __launch_bounds__(256)
synapse_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k, int g_log2) {
  synapse_body<T, LTD, PUSHNEXT, ATOMIC>(p, cx, k, g_log2, blockIdx.x, gridDim.x);
}

// ---------------------------------------------------------------------------------------------
// K3 ltp_kernel(t): for every neuron j that spiked at t walk the plastic prefix of its incoming
// list (new_net.jl:126 `sd[:,n] .+= STDP`): sd += apost * trace_pre as seen at the synapse
// (trace of step t-(d-1), Old_net/DASTDP.jl:62-68).  Synapses whose presynaptic spike arrives
// at this very call are owned by synapse_kernel (see there).  Three gathers are kept in flight
// per lane.
// ---------------------------------------------------------------------------------------------
template <typename T, bool ATOMIC>
__device__ __forceinline__ void ltp_body(const DevNet<T>& p, const WindowCtx<T>* __restrict__ cx, int32_t k, int g_log2,
                                         const uint32_t bid, const uint32_t nblk) {
  typedef Ops<T> O;
  __shared__ unsigned long long blk_ltp;
  tl_begin(p.tl, 2, k);
  const int slot = (cx->slot0 + k) % p.R;
  const int32_t t32 = (int32_t)(cx->t0 + k);
  const int n_now = __ldcg(p.spk_cnt + slot);
  if (threadIdx.x == 0) blk_ltp = 0ull;
  __syncthreads();
  const uint32_t gtid = bid * blockDim.x + threadIdx.x;
  const uint32_t nthreads = nblk * blockDim.x;
  const size_t Nld = (size_t)p.Nld;
  const uint32_t G = 1u << g_log2;
  const uint32_t gl = gtid & (G - 1);
  const uint32_t ngroups = nthreads >> g_log2;
  const int32_t* __restrict__ list = p.spk_list + (size_t)slot * Nld;
  const WView<T> vw = wview(p);
  unsigned long long n_ltp = 0;
  for (uint32_t item = gtid >> g_log2; item < (uint32_t)n_now; item += ngroups) {
    const int32_t j = __ldcg(list + item);
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
      // pre trace as seen at the synapse: of step t-(d-1) (DASTDP.jl:62-68; variant_g one step earlier, :65)
      int2 hq[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const uint32_t i = pk[q] & kPreMask;
        if (!ATOMIC && ok[q]) {
          const int sl = wrap_slot(slot - (int)(pk[q] >> kPreBits), p.R);
          const uint32_t word = p.bits[(size_t)sl * p.W + (i >> 5)];
          ok[q] = !((word >> (i & 31)) & 1u);
        }
        hq[q] = ok[q] ? hist_fetch(p.hist, i) : make_int2(kNever, kNever);
      }
      T trv[3], sdv[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        sdv[q] = (!ATOMIC && ok[q]) ? vw.cur[e[q]].y : (T)0;
        trv[q] = ok[q] ? trace_resolve<T>(p, hq[q], pk[q] & kPreMask, t32 - (int32_t)(pk[q] >> kPreBits) - (p.variant_g ? 1 : 0), t32)
                       : (T)0;
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (!ok[q]) continue;
        if (ATOMIC) sd_add<T>(p, vw, e[q], O::mul(p.apost, trv[q]), p.log_ltp_base + bid, inst);
        else vw.cur[e[q]].y = O::add(sdv[q], O::mul(p.apost, trv[q]));
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) n_ltp += __shfl_down_sync(0xffffffffu, n_ltp, o);
  if ((threadIdx.x & 31) == 0 && n_ltp) atomicAdd(&blk_ltp, n_ltp);
  __syncthreads();
  if (threadIdx.x == 0 && blk_ltp) p.blk_ltp[bid] += blk_ltp;
  tl_end(p.tl, 2, k);
}
template <typename T, bool ATOMIC>
__global__ void __launch_bounds__(256)
ltp_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k, int g_log2) {
  ltp_body<T, ATOMIC>(p, cx, k, g_log2, blockIdx.x, gridDim.x);
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
// learn_pass_tma_kernel (snn_set_param "lsu_pass").  <= 64 registers: two resident blocks must
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

// one pass: blocks take tiles until the dispenser runs dry.  `q` = chunks consumed so far by this block (stage and
// mbarrier parity continue across passes in the persistent learn role)
template <typename T, int NT>
__device__ __forceinline__ void learn_pass_tma_body(const DevNet<T>& p, int do_decay, const uint32_t S, uint64_t* mbar,
                                                    uint32_t* s_tile, typename DevNet<T>::T2* const buf, uint32_t& q) {
  typedef typename DevNet<T>::T2 T2;
  constexpr uint32_t CH = kPassChunkBytes / sizeof(T2);  // synapses per chunk
  const uint32_t tid = threadIdx.x;
  const uint64_t pol = l2_evict_first_policy();
  const uint32_t st0 = *reinterpret_cast<const volatile uint32_t*>(&p.lc->state);
  const uint32_t cur = st0 >> 31, curbit = st0 & 0x80000000u;
  const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(&p.lc->epoch);
  const T2* __restrict__ src = p.wsdX[cur ^ 1u];
  T2* __restrict__ dst = p.wsdX[cur];
  const uint32_t TS = p.tile_syn, U = TS >> 5;
  for (;;) {
    if (tid == 0) *s_tile = atomicAdd(&p.lc->next_tile, 1u);
    __syncthreads();
    const uint32_t tile = *s_tile;
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
      const T g = __ldcg(p.lg + b);
      const int64_t r0 = (int64_t)b * p.Nper;
      auto one = [&](T2& ws, int64_t e) {
        if (!p.plastic_ei && ((int64_t)p.post[e] - r0) >= p.Ne) return;  // exc->inh: not plastic (net_res.jl:213-216)
        learn_one<T>(ws, g, p.wmin, p.wmax, p.sd_decay, 1, do_decay);
      };
      // bulk copies need 16-byte alignment: odd fp32 ends are done by hand
      if (sizeof(T2) == 8) {
        if ((lo & 1) && tid == 0) { T2 ws = __ldcg(src + lo); one(ws, lo); dst[lo] = ws; }
        lo = (lo + 1) & ~(int64_t)1;
        if ((hi & 1) && hi > lo && tid == 1) { T2 ws = __ldcg(src + hi - 1); one(ws, hi - 1); dst[hi - 1] = ws; }
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
          for (uint32_t x = tid; x < (cnt >> 1); x += NT) {
            float4 r = b4[x];
            float2 s0 = make_float2(r.x, r.y), s1 = make_float2(r.z, r.w);
            one(s0, e0 + 2 * x);
            one(s1, e0 + 2 * x + 1);
            b4[x] = make_float4(s0.x, s0.y, s1.x, s1.y);
          }
        } else {
          for (uint32_t x = tid; x < cnt; x += NT) {
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
}

template <typename T>
__global__ void __launch_bounds__(kPassThreads)
learn_pass_tma_kernel(DevNet<T> p, int32_t k, int do_decay, int n_stages) {
  typedef typename DevNet<T>::T2 T2;
  tl_begin(p.tl, 3, k);
  extern __shared__ __align__(128) unsigned char pass_smem[];
  __shared__ __align__(8) uint64_t mbar[kPassMaxStages];
  __shared__ uint32_t s_tile;
  const uint32_t S = (uint32_t)n_stages;
  if (threadIdx.x == 0) {
    for (uint32_t s_ = 0; s_ < S; ++s_) mbar_init(&mbar[s_], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t q = 0;  // chunks consumed so far by this block: stage = q % S, parity = (q / S) & 1
  learn_pass_tma_body<T, kPassThreads>(p, do_decay, S, mbar, &s_tile, reinterpret_cast<T2*>(pass_smem), q);
  tl_end(p.tl, 3, k);
}

enum { kCtlBegin = 0, kCtlAll = 1 };
// begin of a pass, by one block: g[b] from da as decayed by step k, flip cur, frontier = 0, new epoch
template <typename T>
__device__ __forceinline__ void learn_begin(const DevNet<T>& p, const WindowCtx<T>* __restrict__ cx, int32_t k) {
  const uint32_t st = *reinterpret_cast<const volatile uint32_t*>(&p.lc->state);
  const int64_t t = cx->t0 + k;
  for (int b = threadIdx.x; b < p.B; b += blockDim.x) p.lg[b] = learn_rate<T>(p, __ldcg(p.da + (size_t)((t + 1) & 1) * p.B + b));
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

__device__ __forceinline__ Vec4<float> ld4cg(const float* p) {
  float4 r = __ldcg(reinterpret_cast<const float4*>(p));
  Vec4<float> o; o.x[0] = r.x; o.x[1] = r.y; o.x[2] = r.z; o.x[3] = r.w; return o;
}
__device__ __forceinline__ Vec4<double> ld4cg(const double* p) {
  double2 a = __ldcg(reinterpret_cast<const double2*>(p)), b = __ldcg(reinterpret_cast<const double2*>(p + 2));
  Vec4<double> o; o.x[0] = a.x; o.x[1] = a.y; o.x[2] = b.x; o.x[3] = b.y; return o;
}

// Dense spill of the update log (a region was full): dsd holds a few thousand non-zeros among E entries.  16-byte
// loads, four of them in flight per thread (an entry at a time made the scan a chain of E / threads dependent round
// trips: 360 us for the 12.5 M synapses of an 8-way partition, on the critical path of every learn period).
template <typename T>
static __device__ __noinline__ void spill_merge(T* __restrict__ dsd, typename Real2<T>::type* __restrict__ w, const int64_t E,
                                                const int64_t tid, const int64_t nth) {
  const int64_t n4 = E >> 2;
  for (int64_t q0 = tid; q0 < n4; q0 += 4 * nth) {
    Vec4<T> d[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t q = q0 + u * nth;
      if (q < n4) d[u] = ld4cg(dsd + (q << 2));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t q = q0 + u * nth;
      if (q >= n4) break;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (d[u].x[c] != (T)0) { red_add(&w[(q << 2) + c].y, d[u].x[c]); dsd[(q << 2) + c] = (T)0; }
    }
  }
  for (int64_t e = (n4 << 2) + tid; e < E; e += nth) {
    const T dv = __ldcg(dsd + e);
    if (dv != (T)0) { red_add(&w[e].y, dv); dsd[e] = (T)0; }
  }
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
      red_add(&w[le.e].y, le.d);
    }
    __syncthreads();
    if (threadIdx.x == 0 && cnt) p.log_cnt[(size_t)r * kLogStride] = 0u;
  }
  if (p.lc->spilled)
    spill_merge<T>(p.dsd, w, p.E, blockIdx.x * (int64_t)blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
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

// =============================================================================================
// The persistent chain (production mode, one window = FOUR kernel launches).
//
// A window used to be a CUDA graph of ~4 kernels per simulated millisecond whose grids took turns on
// the SMs (round-1 timeline: neuron start->start 20.5 us for a 15 us kernel; side kernels starting
// 5-10 us after their dependencies were met because the neuron grid owned the register file).  Here
// every role is ONE kernel that lives for the whole window, all four co-resident (sized on the host
// from their register / shared-memory footprints), and steps are handed over through counters in
// global memory (ChainCtl) instead of kernel boundaries:
//
//   chain_neuron_kernel   one block per SM; v and u of its neurons stay in SHARED MEMORY for the
//                         whole window, so a step touches only Iacc (read + clear, L2-resident) and
//                         the spike bookkeeping: 8 B per neuron and ms instead of 36.  Iacc is staged
//                         into shared memory with cp.async (every quad of a thread in flight, no
//                         registers), the Philox draw is formed inside the Euler loop, and the four
//                         neurons of a quad are written as selects so that their dependency chains
//                         interleave (a block has 8 warps: the role is bound by latency, not issue).
//   chain_synapse_kernel  LTD + next-step pushes of step k as soon as neuron(k) is complete
//   chain_ltp_kernel      LTP of step k
//   chain_learn_kernel    the asynchronous TMA pass of every learn step of the window (begin, tiles,
//                         frontier = all, early replay of the log), same double-buffer / frontier /
//                         log protocol as the graph path
//
// Dependencies (the same as the graph path, as counters; `done` counters count BLOCKS, so "step k
// complete" reads  counter >= blocks * (k + 1)):
//   neuron(k)  : neuron(k-1) on every block [pushes of delay-1 synapses land in other blocks' Iacc],
//                push phases 0..k-1 [Euler(k-1) consumes them], learn(k-1) begun if step k-1 learns
//                [delay-1 pushes read w], ltp(k-kLag-1) [bounds the readers' lag, see kLag]
//   synapse(k) : neuron(k); on a learn step the pushes for k+1 wait for learn(k) begun
//   ltp(k)     : neuron(k)
//   learn(k)   : LTD(k) and ltp(k) on every block; the previous pass complete and replayed
// Every wait gives up after 10 s and raises ChainCtl::abort (the host then reports SNN_ERR_STATE):
// a scheduling accident can cost a window, never hang the device.
// Mutable data is read through L2 (ld.cg / volatile / acquire): nothing invalidates L1 inside a kernel.
// =============================================================================================
enum { kcNeuron = 0, kcPush = 1, kcLtd = 2, kcLtp = 3, kcBegun = 4, kcLearnBar = 5, kcAbort = 6, kcPassNs = 7, kcPassCnt = 8,
       kcRemote = 9, kcPush1 = 10, kcLtd1 = 11, kcCounters = 12 };  // kcPush1 / kcLtd1: the synapse role's second step group
struct ChainCtl { uint32_t w[kcCounters * 32]; };  // counter c at w[c * 32]: one 128-byte line each
struct ChainDims {
  uint32_t NB, SB, PB, LB, RB;   // blocks of the neuron / synapse / ltp / learn / remote (partitioned runs) role
  uint32_t SG, SBg[2];           // the synapse role works in SG (1 or 2) groups of SBg[g] blocks; group k & 1 takes step k
  uint32_t qpb;              // quads (4 neurons) per neuron block, multiple of 32
  int32_t n_learn;           // learn steps in this window
  int32_t do_decay;          // sd *= sd_decay inside learn! (SNN_SD_DECAY_ON_LEARN)
  int32_t g_arr_log2, g_ltp_log2, pass_stages;
  int32_t thrN, thrP, thrL;  // threads per block of the neuron / ltp / learn role (synapse: 256)
  int32_t la, ring;          // push lookahead (1..kMaxLook) and Iacc ring depth (la + 1)
  int32_t o_lo, o_hi, o_nd, o_need;  // offsets of the push schedule inside `sched`, see Engine::launch_chain
};
// sched[0..n+1]: learn steps among steps < k ; sched[n+2 ..]: the learn steps themselves
__device__ __forceinline__ int sched_before(const int32_t* __restrict__ sched, int k) { return sched[k]; }
__device__ __forceinline__ bool sched_learn(const int32_t* __restrict__ sched, int k) { return sched[k + 1] != sched[k]; }

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* q) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(q) : "memory");
  return v;
}
static __device__ __noinline__ bool chain_wait_slow(ChainCtl* c, int which, uint32_t target, uint32_t sleep_ns) {
  const unsigned long long t0 = gtime_ns();
  for (;;) {
    if (ld_acquire_u32(&c->w[which * 32]) >= target) return true;
    if (*reinterpret_cast<volatile uint32_t*>(&c->w[kcAbort * 32]) != 0u) return false;
    if (gtime_ns() - t0 > 10000000000ull) { atomicExch(&c->w[kcAbort * 32], 1u + (uint32_t)which); return false; }
    if (sleep_ns) __nanosleep(sleep_ns);
  }
}
__device__ __forceinline__ bool chain_wait(ChainCtl* c, int which, uint32_t target, uint32_t sleep_ns) {
  if (ld_acquire_u32(&c->w[which * 32]) >= target) return true;
  return chain_wait_slow(c, which, target, sleep_ns);
}
// Several conditions at once: ALL the counters are loaded before the first one is looked at (one L2 round trip for
// the lot — waiting for them one after the other cost the neuron role three to four round trips per step), then a
// single acquire fence.  target 0 = nothing to wait for.
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* q) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(q) : "memory");
  return v;
}
template <int N>
__device__ __forceinline__ bool chain_wait_all(ChainCtl* c, const int (&which)[N], const uint32_t (&target)[N], uint32_t sleep_ns) {
  unsigned long long t0 = 0;
  for (;;) {
    uint32_t v[N];
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = target[i] ? ld_relaxed_u32(&c->w[which[i] * 32]) : 0u;
    bool ok = true;
    int first_bad = 0;
#pragma unroll
    for (int i = N - 1; i >= 0; --i) if (v[i] < target[i]) { ok = false; first_bad = which[i]; }
    if (ok) { asm volatile("fence.acq_rel.gpu;" ::: "memory"); return true; }
    if (*reinterpret_cast<volatile uint32_t*>(&c->w[kcAbort * 32]) != 0u) return false;
    if (!t0) t0 = gtime_ns();
    else if (gtime_ns() - t0 > 10000000000ull) { atomicExch(&c->w[kcAbort * 32], 1u + (uint32_t)first_bad); return false; }
    if (sleep_ns) __nanosleep(sleep_ns);
  }
}
// phases of the synapse role's group g among the steps -1 .. k_last (group k & 1 takes step k; one group: all of them)
__device__ __forceinline__ uint32_t syn_phases(const ChainDims& d, int g, int k_last) {
  if (d.SG == 1) return g == 0 ? (uint32_t)(k_last + 2) : 0u;
  if (g == 1) return (uint32_t)((k_last + 1) / 2 + 1);        // -1, 1, 3, ...
  return k_last >= 0 ? (uint32_t)(k_last / 2 + 1) : 0u;       // 0, 2, 4, ...
}
// LTD phases (steps 0 .. k) of group g
__device__ __forceinline__ uint32_t syn_ltd_phases(const ChainDims& d, int g, int k) {
  if (d.SG == 1) return g == 0 ? (uint32_t)(k + 1) : 0u;
  return g == 0 ? (uint32_t)(k / 2 + 1) : (uint32_t)((k + 1) / 2);
}

// all threads of the block have finished their part (call after __syncthreads, from one thread)
__device__ __forceinline__ void chain_arrive(ChainCtl* c, int which) {
  // release: an acq_rel fence, not __threadfence() — that one is a sequentially consistent fence (MEMBAR.SC), which
  // the hand-off does not need and which waits longer under the learn pass's traffic
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
  red_add(&c->w[which * 32], 1u);
}


constexpr int kChainNeuronThreads = 512, kChainLearnThreads = 128, kSpkWave = 128;

// Delay-1 synapses of spiker i pushed warp-cooperatively into Iacc[t&1] — they feed the Euler update of this very
// step (new_net.jl:116-121).  Weights of synapses the learn pass in flight has not reached are formed on the fly.
template <typename T>
__device__ __forceinline__ void chain_push_delay1(const DevNet<T>& p, const WView<T>& vw, const uint32_t i, const uint32_t lane,
                                                  T* __restrict__ I_now) {
  const uint32_t Ne = (uint32_t)p.Ne, Nper = (uint32_t)p.Nper;
  const uint32_t s0 = p.seg[(size_t)i * (p.Dmax + 1)], s1 = p.seg[(size_t)i * (p.Dmax + 1) + 1];
  const uint32_t inst = p.B == 1 ? 0u : i / Nper, ibase = inst * Nper;
  const bool row_exc = (i - ibase) < Ne;
  for (uint32_t eb = s0 + lane; eb < s1; eb += 128) {  // four loads in flight per lane
    int32_t pj[4]; typename DevNet<T>::T2 pw[4]; bool fly[4];
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const uint32_t e = eb + 32 * q4;
      pj[q4] = e < s1 ? p.post[e] : -1;
      fly[q4] = false;
      if (e < s1) pw[q4] = w_fetch<T>(vw, e, row_exc, fly[q4]);
    }
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4)
      if (pj[q4] >= 0)
        red_add(&I_now[pj[q4]], w_resolve<T>(p, pw[q4], fly[q4], p.plastic_ei || ((uint32_t)pj[q4] - ibase) < Ne, inst, (T)0));
  }
}

// ---------------------------------------------------------------------------------------------
// chain_neuron_kernel: Euler(k-1) + detect(k) for k = 0..n over a block's own neurons (new_net.jl:111-114,
// 120-123,130-131), state resident in shared memory: sv / su [4*qpb], sn [4*qpb] the currents of this Euler.
// ---------------------------------------------------------------------------------------------
// MINB: resident blocks per SM the register allocation is capped for (2: 64 registers, for blocks of <= 256
// threads; 3: 40 registers with some spills, for 384-512 threads) — the role shares the SM with three others
template <typename T, int MINB>
__global__ void __launch_bounds__(kChainNeuronThreads, MINB)
chain_neuron_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, ChainCtl* __restrict__ ctl,
                    const int32_t* __restrict__ sched, const ChainDims d) {
  typedef Ops<T> O;
  extern __shared__ __align__(16) unsigned char nsm_raw[];
  __shared__ int s_ok;
  __shared__ int s_nspk[2], s_base;
  __shared__ uint32_t s_ps0[kSpkWave], s_pb1[kSpkWave], s_pb2[kSpkWave], s_poff[kSpkWave + 1];  // spikers of this block in step k: s_nspk[k & 1] (the other one is being cleared)
  T* const sv = reinterpret_cast<T*>(nsm_raw);
  T* const su = sv + (size_t)4 * d.qpb;
  T* const sn = su + (size_t)4 * d.qpb;
  uint16_t* const s_spk = reinterpret_cast<uint16_t*>(sn + (size_t)4 * d.qpb);  // [4*qpb] this step's spikers (local ids)
  const uint32_t tid = threadIdx.x, lane = tid & 31u, bid = blockIdx.x, TN = blockDim.x;
  const int n = cx->n_steps;
  const int64_t t0 = cx->t0;
  const uint32_t N = (uint32_t)p.N, Nper = (uint32_t)p.Nper, Ne = (uint32_t)p.Ne;
  const uint32_t q_lo = (uint32_t)(p.part_lo >> 2), q_hi = (uint32_t)((p.part_hi >= p.N ? p.Nld : p.part_hi) >> 2);
  const uint32_t qb0 = q_lo + bid * d.qpb;
  const uint32_t nq = qb0 < q_hi ? min(d.qpb, q_hi - qb0) : 0u;  // multiple of 32: warps stay whole
  const bool philox = p.noise_mode == SNN_NOISE_PHILOX;
  int32_t* __restrict__ rec_ids = cx->rec_ids;
  const int64_t rec_cap = cx->rec_cap;

  for (uint32_t ql = tid; ql < nq; ql += TN) {
    st4(sv + 4 * ql, ld4(p.v + ((size_t)(qb0 + ql) << 2)));
    st4(su + 4 * ql, ld4(p.u + ((size_t)(qb0 + ql) << 2)));
  }
  if (tid == 0) s_nspk[0] = s_nspk[1] = 0;
  __syncthreads();
  unsigned long long spikes_acc = 0;  // block 0, thread 0

  const T k_a_exc = p.a_exc, k_a_inh = p.a_inh, k_d_exc = p.d_exc, k_d_inh = p.d_inh, k_vthresh = p.vthresh, k_v_reset = p.v_reset;
  const bool k_thr_ge = p.threshold_ge != 0, k_homeo = p.homeostasis != 0;
  // timeline = 2: the phase classes 6-12 are stamped by ONE block (the middle one) — a block's own step, not the
  // first-begin / last-end envelope over all blocks; 11 / 12 then are the two halves of this role's spiker wave
  const bool solo = (p.exp_flags & 1024) != 0;
  unsigned long long* const tq = (!solo || bid == d.NB / 2) ? p.tl : nullptr;
  unsigned long long* const ts = solo ? tq : nullptr;
  for (int k = 0; k <= n; ++k) {
    const bool euler = k > 0, detect = k < n;
    tl_begin(tq, 6, k);
    if (tid == 0) {
      // neuron(k-1) on every block | the push phases Euler(k-1) consumes | learn!(k-1) begun | ltp not more than kLag
      // steps behind | (partitioned) the peers' spikes of step k-1 pushed and listed
      const int kl = k >= 1 ? sched[d.o_need + k] - 2 : -2;  // last push phase needed
      const int which[6] = {kcNeuron, kcPush, kcPush1, kcBegun, kcLtp, kcRemote};
      const uint32_t target[6] = {d.NB * (uint32_t)k,
                                  kl >= -1 ? d.SBg[0] * syn_phases(d, 0, kl) : 0u,
                                  kl >= -1 ? d.SBg[1] * syn_phases(d, 1, kl) : 0u,
                                  (k >= 1 && sched_learn(sched, k - 1)) ? (uint32_t)sched_before(sched, k) : 0u,
                                  k > kLag ? d.PB * (uint32_t)(k - kLag) : 0u,
                                  (k >= 1 && d.RB) ? d.RB * (uint32_t)k : 0u};
      s_ok = chain_wait_all<6>(ctl, which, target, 0) ? 1 : 0;
      s_nspk[(k + 1) & 1] = 0;  // last read in phase 3 of step k-1, which every thread has left
    }
    __syncthreads();
    if (!s_ok) break;
    tl_end(tq, 6, k);
    tl_begin(p.tl, 0, k);
    tl_begin(tq, 7, k);
    const unsigned long long t_step0 = (p.tl && (k == kTlProbeA || k == kTlProbeB)) ? gtime_ns() : 0ull;
    const int64_t t = t0 + k;
    const int slot = (cx->slot0 + k) % p.R, prev = wrap_slot(slot - 1, p.R);
    uint32_t* __restrict__ bits = p.bits + (size_t)slot * p.W;
    int32_t* __restrict__ list = p.spk_list + (size_t)slot * p.Nld;
    T* __restrict__ I_prev = p.Iacc + (size_t)((t + d.ring - 1) % d.ring) * p.Nld;  // consumed by Euler(t-1)
    const T* __restrict__ noise_prev = (euler && cx->noise) ? cx->noise + (size_t)(k - 1) * p.N : nullptr;
    int32_t rec_base = 0;
    const int n_prev = (k > 0 && (rec_ids || bid == 0)) ? __ldcg(p.spk_cnt + prev) : 0;
    if (rec_ids && k > 0) {
      const int64_t nxt = (int64_t)__ldcg(cx->rec_off + (k - 1)) + n_prev;
      rec_base = (int32_t)(nxt < rec_cap ? nxt : rec_cap);
    }
    // (retire_spikes: after phase 2 — its list length is a dependent load nobody should wait for up front)
    const int64_t xr = t - p.G;
    const int retire_cnt = (detect && xr >= 0) ? __ldcg(p.spk_cnt + (int)(xr % p.R)) : 0;
    WView<T> vw;
    if (detect) vw = wview(p);
    // dopamine (new_net.jl:131 da *= 0.995 ; newnet_DASTDP.jl:54-56 da += 0.5 after step!)
    if (bid == 0) {
      if (tid == 0) {
        if (euler) {
          spikes_acc += (unsigned long long)n_prev;
          if (rec_ids) cx->rec_off[k] = rec_base;
        }
        if (detect) p.spk_cnt[wrap_slot(slot + 1, p.R)] = 0;  // next step's list (no reader is R-1 steps behind)
      }
      const T* __restrict__ din = p.da + (size_t)(t & 1) * p.B;
      T* __restrict__ dout = p.da + (size_t)((t + 1) & 1) * p.B;
      const bool evs = cx->ev_off && k > 0;
      const int e0 = evs ? cx->ev_off[k - 1] : 0, e1 = evs ? cx->ev_off[k] : 0;
      for (int b = tid; b < p.B; b += TN) {
        T dv = __ldcg(din + b);
        for (int q = e0; q < e1; ++q)
          if (cx->ev_inst[q] == b) dv = O::add(dv, (T)cx->ev_amt[q]);
        if (detect) dout[b] = O::mul(dv, p.da_decay);
        else if (e1 > e0) p.da[(size_t)(t & 1) * p.B + b] = dv;  // window end: events only, in place
      }
    }

    // ---- phase 1: the Iacc quads of this block, global -> shared memory (sn) as asynchronous 16-byte copies: every
    // quad of a thread is in flight at once and none of them occupies a register (the register version had to go in
    // two batches of four — twice the L2 latency under the learn pass's traffic — and a batch of eight spilled).
    // A thread waits for its own copies only and consumes the same quads in phase 2: no barrier in between.
    if (euler) {
      for (uint32_t ql = tid; ql < nq; ql += TN) {
        const T* src = I_prev + ((size_t)(qb0 + ql) << 2);
        const uint32_t dst = smem_u32(sn + 4 * ql);
#pragma unroll
        for (uint32_t o = 0; o < 4 * sizeof(T); o += 16)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + o), "l"(reinterpret_cast<const char*>(src) + o) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (p.tl) { __syncthreads(); tl_end(tq, 7, k); tl_begin(tq, 8, k); }
    // ---- phase 2: Euler(t-1) and detect(t) out of shared memory
    for (uint32_t ql = tid; ql < nq; ql += TN) {  // warp-uniform trip count (nq and TN are multiples of 32)
      const uint32_t q = qb0 + ql, j0 = q << 2;
      Vec4<T> v = ld4(sv + 4 * ql), u = ld4(su + 4 * ql), ho;
      if (detect && k_homeo) ho = ld4cg(p.ho + j0);
      const uint32_t l0 = p.B == 1 ? j0 : j0 % Nper;
      unsigned excm = 0, actm = 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t l = l0 + c;
        if (l >= Nper) l -= Nper;
        excm |= (l < Ne ? 1u : 0u) << c;
        actm |= (j0 + c < N ? 1u : 0u) << c;
      }
      // the four neurons of a quad are four independent dependency chains: selects instead of branches, so that the
      // compiler interleaves them (a block has few warps: the role is bound by instruction latency, not by issue slots)
      if (euler) {
        // total current of Euler(t-1) = noise + Iacc (new_net.jl:111,116-118); Iacc is cleared for the pushes that
        // arrive `ring` steps later
        Vec4<T> I = ld4(sn + 4 * ql);
        if (p.noise_mode == SNN_NOISE_REPLAY) {
#pragma unroll
          for (int c = 0; c < 4; ++c) if (j0 + c < N) I.x[c] = O::add(noise_prev[j0 + c], I.x[c]);
        } else if (philox) {  // 13*(U-0.5) (new_net.jl:111): a counter-based draw keyed by (seed, step, quad)
          const int64_t tp = t - 1;
          const uint4 o = philox4x32_10(make_uint4(q, 0u, (uint32_t)tp, (uint32_t)((uint64_t)tp >> 32)),
                                        make_uint2((uint32_t)p.noise_seed, (uint32_t)(p.noise_seed >> 32)));
          I.x[0] = O::add((T)philox_to_noise(o.x, p.noise_amp), I.x[0]);
          I.x[1] = O::add((T)philox_to_noise(o.y, p.noise_amp), I.x[1]);
          I.x[2] = O::add((T)philox_to_noise(o.z, p.noise_amp), I.x[2]);
          I.x[3] = O::add((T)philox_to_noise(o.w, p.noise_amp), I.x[3]);
        }
        {
          Vec4<T> z; z.x[0] = z.x[1] = z.x[2] = z.x[3] = (T)0;
          st4(I_prev + j0, z);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const bool act = (actm >> c) & 1u, exc = (excm >> c) & 1u;
          T vv = v.x[c];
          // v = v + 0.5*((0.04*v+5)*v + 140 - u + I), twice (new_net.jl:120-121)
#pragma unroll
          for (int rep = 0; rep < 2; ++rep) {
            T x = O::add(O::mul((T)0.04, vv), (T)5);
            x = O::add(O::mul(x, vv), (T)140);
            x = O::add(O::sub(x, u.x[c]), I.x[c]);
            vv = O::add(vv, O::mul((T)0.5, x));
          }
          const T un = O::add(u.x[c], O::mul(exc ? k_a_exc : k_a_inh, O::sub(O::mul((T)0.2, vv), u.x[c])));  // :122
          v.x[c] = act ? vv : v.x[c];
          u.x[c] = act ? un : u.x[c];
        }
      }
      if (detect) {
        unsigned nib = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const bool act = (actm >> c) & 1u, exc = (excm >> c) & 1u;
          const T x = k_homeo ? O::sub(v.x[c], ho.x[c]) : v.x[c];
          const bool s = act && (k_thr_ge ? (x >= k_vthresh) : (x > k_vthresh));
          const T ur = O::add(u.x[c], exc ? k_d_exc : k_d_inh);
          v.x[c] = s ? k_v_reset : v.x[c];
          u.x[c] = s ? ur : u.x[c];
          if (k_homeo) {
            uint32_t l = l0 + c;
            if (l >= Nper) l -= Nper;
            T h = ho.x[c];
            if (s && (exc || p.ho_inh) && (int64_t)l >= p.ho_from) h = O::add(h, p.theta);
            ho.x[c] = O::mul(h, p.ttheta);
          }
          nib |= (s ? 1u : 0u) << c;
        }
        if (k_homeo) st4(p.ho + j0, ho);
        st4(sv + 4 * ql, v);
        st4(su + 4 * ql, u);
        unsigned word = nib << ((lane & 7) << 2);  // 8 lanes x 4 neurons = one 32-bit bitmap word
        word |= __shfl_xor_sync(0xffffffffu, word, 1);
        word |= __shfl_xor_sync(0xffffffffu, word, 2);
        word |= __shfl_xor_sync(0xffffffffu, word, 4);
        if ((lane & 7) == 0) bits[j0 >> 5] = word;
        if (p.n_ranks > 1) {
          // fused exchange (neuron-partitioned runs): a warp row owns 4 consecutive bitmap words = two 16-byte stores
          // of (word, step tag) pairs into every peer's ring over NVLink
          const unsigned w1 = __shfl_sync(0xffffffffu, word, 8), w2 = __shfl_sync(0xffffffffu, word, 16),
                         w3 = __shfl_sync(0xffffffffu, word, 24);
          if (lane < (uint32_t)p.n_ranks && lane != (uint32_t)p.rank) {
            const uint32_t tag = (uint32_t)(t + 1);
            uint4* dst = reinterpret_cast<uint4*>(p.peer_ll[lane] + (size_t)slot * p.W + (j0 >> 5));
            dst[0] = make_uint4(word, tag, w1, tag);
            dst[1] = make_uint4(w2, tag, w3, tag);
          }
        }
        // spikers (rare) only go on the block's list here: their global bookkeeping and delay-1 pushes are a chain
        // of dependent memory round trips, done for all of them at once in phase 3
        for (unsigned nb = nib; nb; nb &= nb - 1u) s_spk[atomicAdd(&s_nspk[k & 1], 1)] = (uint16_t)(4 * ql + (__ffs(nb) - 1));
      } else {
        st4(sv + 4 * ql, v);
        st4(su + 4 * ql, u);
      }
    }
    if (ts) { __syncthreads(); tl_end(ts, 8, k); tl_begin(ts, 5, k); }  // (solo timeline: class 5 = retire + block barrier)
    if (retire_cnt > 0) {  // the spikes of step t - G leave the part of the bitmap ring a reader may rely on
      const int32_t* __restrict__ rl = p.spk_list + (size_t)(xr % p.R) * p.Nld;
      for (uint32_t q = bid * TN + tid; q < (uint32_t)retire_cnt; q += d.NB * TN) p.old_last[__ldcg(rl + q)] = (int32_t)xr;
    }
    __syncthreads();
    if (ts) tl_end(ts, 5, k); else tl_end(tq, 8, k);
    tl_begin(tq, 9, k);
    // ---- phase 3: this block's spikers, all in flight together (a chain of dependent memory round trips per
    // spiker: done as a small wave — one thread per spiker fetches the bounds of its delay-1 segment, then one thread
    // per SYNAPSE pushes it: delay 1 feeds the Euler update of this very step (new_net.jl:116-121), delays up to the
    // lookahead the following ones).
    // Thread 0 reserves the block's stretch of the global spike list meanwhile.
    if (detect) {
      const int nspk = s_nspk[k & 1];
      const int nd = sched[d.o_nd + k];  // delays 1..nd of this step's spikers are pushed here (the rest by the synapse role)
      if (nspk > 0) {
        if (tid == 0) s_base = atomicAdd(&p.spk_cnt[slot], nspk);
        // previous spike of this thread's spiker(s), for the history word written at the end: in flight during the pushes
        const int32_t h_prev = (int)tid < nspk ? __ldcg(p.hist + ((qb0 << 2) + s_spk[tid])).x : 0;
        for (int r0 = 0; r0 < nspk; r0 += kSpkWave) {
          const int nr = min(kSpkWave, nspk - r0);
          tl_begin(ts, 11, k);
          for (int x = (int)tid; x < nr; x += (int)TN) {
            const uint4 sg = p.segL[(qb0 << 2) + s_spk[r0 + x]];
            s_ps0[x] = sg.x; s_pb1[x] = sg.y; s_pb2[x] = sg.z;
            s_poff[x + 1] = (nd == 1 ? sg.y : (nd == 2 ? sg.z : sg.w)) - sg.x;
          }
          __syncthreads();
          if (tid < 32) {
            constexpr int PER = kSpkWave / 32;
            const int b0 = (int)lane * PER;
            uint32_t acc = 0;
#pragma unroll
            for (int q = 0; q < PER; ++q) if (b0 + q < nr) acc += s_poff[b0 + q + 1];
            uint32_t inc = acc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
              if (lane >= (uint32_t)o) inc += up;
            }
            uint32_t run = inc - acc;
            if (lane == 0) s_poff[0] = 0u;
#pragma unroll
            for (int q = 0; q < PER; ++q)
              if (b0 + q < nr) { run += s_poff[b0 + q + 1]; s_poff[b0 + q + 1] = run; }
          }
          __syncthreads();
          const uint32_t V = s_poff[nr];
          tl_end(ts, 11, k); tl_begin(ts, 12, k);
          for (uint32_t v0 = tid; v0 < V; v0 += 4 * TN) {
            int32_t pj[4]; typename DevNet<T>::T2 pw[4]; bool fly[4]; uint32_t pi[4], pd[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint32_t v = v0 + u * TN;
              pj[u] = -1; fly[u] = false; pi[u] = 0; pd[u] = 0;
              if (v >= V) continue;
              uint32_t a = 0, b = (uint32_t)nr;
              while (b - a > 1) { const uint32_t m = (a + b) >> 1; if (s_poff[m] <= v) a = m; else b = m; }
              const uint32_t e = s_ps0[a] + (v - s_poff[a]);
              pi[u] = (qb0 << 2) + s_spk[r0 + a];
              pd[u] = (e >= s_pb1[a] ? 1u : 0u) + (e >= s_pb2[a] ? 1u : 0u);  // delay - 1 = steps until it arrives
              const uint32_t ib = p.B == 1 ? 0u : (pi[u] / Nper) * Nper;
              pj[u] = p.post[e];
              pw[u] = w_fetch<T>(vw, e, (pi[u] - ib) < Ne, fly[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (pj[u] < 0) continue;
              const uint32_t inst = p.B == 1 ? 0u : pi[u] / Nper;
              red_add(p.Iacc + (size_t)((t + pd[u]) % d.ring) * p.Nld + pj[u],
                        w_resolve<T>(p, pw[u], fly[u], p.plastic_ei || ((uint32_t)pj[u] - inst * Nper) < Ne, inst, (T)0));
            }
          }
          __syncthreads();
          tl_end(ts, 12, k);
        }
        const int base = s_base;
        for (int w = (int)tid; w < nspk; w += (int)TN) {
          const uint32_t j = (qb0 << 2) + s_spk[w];
          list[base + w] = (int32_t)j;
          if (rec_ids && (int64_t)rec_base + base + w < rec_cap) rec_ids[rec_base + base + w] = (int32_t)j;
          p.hist[j] = make_int2((int32_t)t, w == (int)tid ? h_prev : __ldcg(p.hist + j).x);  // trace set (new_net.jl:123)
        }
        __syncthreads();
      }
    }
    tl_end(tq, 9, k);
    if (t_step0 && tid == 0 && smid() < (uint32_t)kTlSteps)
      p.tl[(k == kTlProbeA ? 0 : kTlClasses * kTlSteps) + 15 * kTlSteps + smid()] = gtime_ns() - t_step0;
    tl_begin(ts, 10, k);  // (solo timeline: class 10 = the fence + counter update of the hand-over)
    if (tid == 0) chain_arrive(ctl, kcNeuron);
    tl_end(ts, 10, k);
    tl_end(p.tl, 0, k);
  }
  __syncthreads();
  if (bid == 0 && tid == 0) p.counters[0] += spikes_acc;
  for (uint32_t ql = tid; ql < nq; ql += TN) {
    st4(p.v + ((size_t)(qb0 + ql) << 2), ld4(sv + 4 * ql));
    st4(p.u + ((size_t)(qb0 + ql) << 2), ld4(su + 4 * ql));
  }
}

// ---------------------------------------------------------------------------------------------
// synapse_wave (persistent chain, production mode): the same work as synapse_body — LTD on the delay-(a+1)
// segment of every (age a, spiker) item, pushes of the adjacent delay-(a+2) segment into Iacc[(k+1)&1] — laid out
// for memory-level parallelism instead of thread count.  A resident role has a fixed, small number of threads;
// one group of lanes walking one item at a time made a step a serial chain of ~5 dependent DRAM round trips per
// item (55-90 us per step at C5 while the learn pass keeps HBM busy).  Here a block takes a contiguous slice of
// the step's items and
//   stage A  one thread per item: spike-list entry -> its three segment bounds (all items' loads in flight at
//            once), written to shared memory; block-wide scan of the segment lengths;
//   stage B  one thread per SYNAPSE VISIT (balanced over the block whatever the row lengths), two visits per
//            thread in flight: post id + weight, then the post neuron's spike history, then the reduction.
// A step costs each block ~4 dependent round trips in total.
// ---------------------------------------------------------------------------------------------
constexpr int kWaveItems = 256;
struct WaveSmem {
  int cum[kMaxDelay + 2];
  int slots[kMaxDelay + 1];
  uint32_t s0[kWaveItems], n1[kWaveItems], src[kWaveItems];
  uint32_t pb[kWaveItems][kMaxLook + 1];  // synapse wave: bounds of the delay segments an item pushes
  uint32_t off[kWaveItems + 1];
  unsigned long long blk_ev;
};

// r_lo..r_hi: the pushes of this call are the arrivals of steps k+r_lo .. k+r_hi (none if r_hi < r_lo), i.e. for an
// item of age a the delay segments a+r_lo+1 .. a+r_hi+1, each into the ring buffer of its arrival step.
template <typename T, bool LTD>
__device__ __forceinline__ void synapse_wave(const DevNet<T>& p, const WindowCtx<T>* __restrict__ cx, int32_t k, const int r_lo,
                                             const int r_hi, const int ring, const uint32_t bid, const uint32_t nblk,
                                             const uint32_t region, WaveSmem& sm) {
  // bid / nblk: this block's share of the step's items; region: its private sd-update log (unique per block);
  // sm.blk_ev accumulates the delivered events over the calls (the caller zeroes it first and stores it last)
  typedef Ops<T> O;
  tl_begin(p.tl, LTD ? 1 : 5, k);
  const uint32_t tid = threadIdx.x, TB = blockDim.x, lane = tid & 31u;
  const int64_t t = cx->t0 + k;
  const int32_t t32 = (int32_t)t;
  const int slot = (int)(((int64_t)cx->slot0 + k + p.R) % p.R);  // k may be -1 (window-start push)
  const bool push = r_hi >= r_lo;
  const int nseg = push ? r_hi - r_lo + 1 : 0;                   // <= kMaxLook
  if (tid < (uint32_t)p.Dmax) {
    const int sl = wrap_slot(slot - (int)tid, p.R);
    sm.slots[tid] = sl;
    sm.cum[tid + 1] = __ldcg(p.spk_cnt + sl);
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int a = 0; a < p.Dmax; ++a) { const int c = sm.cum[a + 1]; sm.cum[a] = acc; acc += c; }
    sm.cum[p.Dmax] = acc;
  }
  __syncthreads();
  const uint32_t total = (uint32_t)sm.cum[p.Dmax];
  const uint32_t per = (total + nblk - 1) / nblk;
  const uint32_t lo = min(total, bid * per), hi = min(total, lo + per);
  const uint32_t Nper = (uint32_t)p.Nper, Ne = (uint32_t)p.Ne;
  const size_t Nld = (size_t)p.Nld;
  const WView<T> vw = wview(p);
  unsigned long long n_ev = 0;
  for (uint32_t r0 = lo; r0 < hi; r0 += kWaveItems) {
    const uint32_t nitem = min((uint32_t)kWaveItems, hi - r0);
    if (LTD && !(p.exp_flags & 1024)) tl_begin(p.tl, 11, k);
    // ---- stage A: item -> (LTD segment, push segments, spiker)
    for (uint32_t x = tid; x < nitem; x += TB) {
      const uint32_t item = r0 + x;
      int a = 0;
      while (a + 1 < p.Dmax && sm.cum[a + 1] <= (int)item) ++a;
      const uint32_t i = (uint32_t)__ldcg(p.spk_list + (size_t)sm.slots[a] * Nld + (item - sm.cum[a]));
      const uint32_t* __restrict__ sg = p.seg + (size_t)i * (p.Dmax + 1);
      const uint32_t g0 = sg[a], g1 = sg[a + 1];
      uint32_t pbv[kMaxLook + 1], p_end = g1;  // (no run-time index into pbv: it would live in local memory)
#pragma unroll
      for (int q = 0; q <= kMaxLook; ++q) {
        pbv[q] = (push && q <= nseg) ? sg[min(a + r_lo + q, p.Dmax)] : g1;
        if (q == nseg) p_end = pbv[q];
      }
      const uint32_t ibase = p.B == 1 ? 0u : (i / Nper) * Nper;
      const bool ltd_row = LTD && (i - ibase) < Ne;  // only excitatory rows are plastic: the others have no LTD visits
      if (LTD) n_ev += (g1 - g0);
      const uint32_t b0 = ltd_row ? g0 : g1;
      sm.s0[x] = b0; sm.n1[x] = g1 - b0; sm.src[x] = i;
#pragma unroll
      for (int q = 0; q <= kMaxLook; ++q) sm.pb[x][q] = pbv[q];
      sm.off[x + 1] = (g1 - b0) + (push ? p_end - pbv[0] : 0u);  // length; scanned below
    }
    __syncthreads();
    if (tid < 32) {  // inclusive scan of off[1..nitem] by one warp, consecutive items per lane
      constexpr int PER = kWaveItems / 32;
      const uint32_t b = lane * PER;
      uint32_t acc = 0;
#pragma unroll
      for (int q = 0; q < PER; ++q) if (b + q < nitem) acc += sm.off[b + q + 1];
      uint32_t inc = acc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += up;
      }
      uint32_t run = inc - acc;
      if (lane == 0) sm.off[0] = 0u;
#pragma unroll
      for (int q = 0; q < PER; ++q)
        if (b + q < nitem) { run += sm.off[b + q + 1]; sm.off[b + q + 1] = run; }
    }
    __syncthreads();
    const uint32_t V = sm.off[nitem];
    if (LTD && !(p.exp_flags & 1024)) { tl_end(p.tl, 11, k); tl_begin(p.tl, 12, k); }
    // ---- stage B: one thread per synapse visit, two in flight
    for (uint32_t v0 = tid; v0 < V; v0 += 2 * TB) {
      uint32_t e[2], j[2], inst[2], ibs[2], ahead[2]; bool ok[2], isl[2], pl[2], fly[2]; typename DevNet<T>::T2 ws[2]; int2 hq[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const uint32_t v = v0 + u * TB;
        ok[u] = v < V; isl[u] = false; pl[u] = false; fly[u] = false; e[u] = 0; j[u] = 0; inst[u] = 0; ibs[u] = 0; ahead[u] = 0;
        if (!ok[u]) continue;
        uint32_t a = 0, b = nitem;  // off[a] <= v < off[b]
        while (b - a > 1) { const uint32_t m = (a + b) >> 1; if (sm.off[m] <= v) a = m; else b = m; }
        const uint32_t rel = v - sm.off[a], n1 = sm.n1[a];
        isl[u] = rel < n1;                   // inside the segment that arrives now (LTD), else a push for a later step
        if (isl[u]) e[u] = sm.s0[a] + rel;
        else {
          e[u] = sm.pb[a][0] + (rel - n1);
          uint32_t seg_i = 0;
#pragma unroll
          for (int q = 1; q < kMaxLook; ++q) seg_i += (q < nseg && e[u] >= sm.pb[a][q]) ? 1u : 0u;
          ahead[u] = (uint32_t)r_lo + seg_i;  // arrives at step k + ahead
        }
        const uint32_t i = sm.src[a];
        inst[u] = p.B == 1 ? 0u : i / Nper; ibs[u] = inst[u] * Nper;
        j[u] = (uint32_t)p.post[e[u]];
        if (!isl[u]) ws[u] = w_fetch<T>(vw, e[u], (i - ibs[u]) < Ne, fly[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        pl[u] = ok[u] && isl[u] && (p.plastic_ei || (j[u] - ibs[u]) < Ne);  // exc -> inh may be non-plastic (net_res.jl:213-216)
        if (pl[u]) hq[u] = hist_fetch(p.hist, j[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (!ok[u]) continue;
        if (pl[u])             // arrives now: sd -= apre * trace_post(t)   (new_net.jl:127)
          sd_add<T>(p, vw, e[u], -O::mul(p.apre, trace_resolve<T>(p, hq[u], j[u], t32, t32)), region, inst[u]);
        else if (!isl[u])      // arrives later: I += w into the buffer of its arrival step   (new_net.jl:116-118)
          red_add(p.Iacc + (size_t)((t + ahead[u]) % ring) * Nld + j[u],
                    w_resolve<T>(p, ws[u], fly[u], p.plastic_ei || (j[u] - ibs[u]) < Ne, inst[u], (T)0));
      }
    }
    __syncthreads();
    if (LTD && !(p.exp_flags & 1024)) tl_end(p.tl, 12, k);
  }
  if (LTD) {
    for (int o = 16; o > 0; o >>= 1) n_ev += __shfl_down_sync(0xffffffffu, n_ev, o);
    if (lane == 0 && n_ev) atomicAdd(&sm.blk_ev, n_ev);
  }
  tl_end(p.tl, LTD ? 1 : 5, k);
}

// the waves as kernels of their own (captured-graph scheduler): one launch per step, as many blocks as the step has work for
template <typename T, bool LTD>
__global__ void __launch_bounds__(256)
synapse_wave_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k, int r_lo, int r_hi) {
  __shared__ WaveSmem wsm;
  if (threadIdx.x == 0) wsm.blk_ev = 0ull;
  synapse_wave<T, LTD>(p, cx, k, r_lo, r_hi, 2, blockIdx.x, gridDim.x, blockIdx.x, wsm);
  __syncthreads();
  if (threadIdx.x == 0 && wsm.blk_ev) p.blk_ev[blockIdx.x] += wsm.blk_ev;
}

// ---------------------------------------------------------------------------------------------
// chain_synapse_kernel: k = -1 (arrivals of window step 0 already in flight), then per step LTD + the pushes of
// the next step; on a learn step the pushes wait until learn!(k) has begun (they must see its weights)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256, 5)
chain_synapse_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, ChainCtl* __restrict__ ctl,
                     const int32_t* __restrict__ sched, const ChainDims d) {
  __shared__ int s_ok;
  __shared__ WaveSmem wsm;
  const int n = cx->n_steps;
  // two groups of blocks take alternate steps (group k & 1 takes step k): a step is a chain of dependent memory
  // round trips that neither group can shorten, but two steps in flight double the role's throughput
  tl_count_sm(p.tl, 13);
  const uint32_t grp = d.SG == 2 ? (blockIdx.x & 1u) : 0u;
  // (no run-time index into d.SBg: the whole by-value ChainDims would be copied to local memory)
  const uint32_t bid = d.SG == 2 ? (blockIdx.x >> 1) : blockIdx.x, nblk = grp ? d.SBg[1] : d.SBg[0], region = blockIdx.x;
  const int cPush = grp ? kcPush1 : kcPush, cLtd = grp ? kcLtd1 : kcLtd;
  auto arrive = [&](int which) {
    __syncthreads();
    if (threadIdx.x == 0) chain_arrive(ctl, which);
  };
  if (threadIdx.x == 0) wsm.blk_ev = 0ull;
  // push schedule (Engine::launch_chain): at step k the arrivals of steps lo[k]..hi[k] are pushed — one step of
  // lookahead further than a step ago, everything up to the lookahead right after a learn step has begun
  const int32_t* __restrict__ plo = sched + d.o_lo + 1;  // indexed by k = -1 .. n-1
  const int32_t* __restrict__ phi = sched + d.o_hi + 1;
  for (int k = -1; k < n; ++k) {
    if (d.SG == 2 && (uint32_t)(k & 1) != grp) continue;
    const bool learn = k >= 0 && sched_learn(sched, k);
    const int r_lo = plo[k] - k, r_hi = phi[k] - k;
    if (k >= 0) {
      if (threadIdx.x == 0) {  // neuron(k) on every block; partitioned: the step's spike list complete
        const int which[2] = {kcNeuron, kcRemote};
        const uint32_t target[2] = {d.NB * (uint32_t)(k + 1), d.RB * (uint32_t)(k + 1)};
        s_ok = chain_wait_all<2>(ctl, which, target, 40) ? 1 : 0;
      }
      __syncthreads();
      if (!s_ok) break;
    }
    // (each instantiation of the wave appears once: four inlined copies cost 200 bytes of spills per thread)
    if (k >= 0) {  // LTD of the arrivals of step k, fused with the look-ahead pushes unless this is a learn step
      // (the log region rotates with the step: a block whose slice of the items carries most of the plastic visits —
      // in a partition the local spikers sit together in the list — does not fill one region on its own)
      synapse_wave<T, true>(p, cx, k, learn ? 1 : r_lo, learn ? 0 : r_hi, d.ring, bid, nblk, (region + (uint32_t)k) % d.SB, wsm);
      arrive(cLtd);
      if (!learn && threadIdx.x == 0) red_add(&ctl->w[cPush * 32], 1u);  // ordered after the fence of the arrive above
    }
    if (k < 0 || learn) {  // spikes of earlier windows / the pushes that must see learn!(k)
      if (learn && r_hi >= r_lo) {
        if (threadIdx.x == 0) s_ok = chain_wait(ctl, kcBegun, (uint32_t)sched_before(sched, k) + 1u, 40) ? 1 : 0;
        __syncthreads();
        if (!s_ok) break;
      }
      if (r_hi >= r_lo) synapse_wave<T, false>(p, cx, k, r_lo, r_hi, d.ring, bid, nblk, region, wsm);
      arrive(cPush);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && wsm.blk_ev) p.blk_ev[blockIdx.x] += wsm.blk_ev;
}

// ltp_wave (persistent chain): the work of ltp_body — for every neuron j that spiked at t, sd += apost * pre trace
// as seen at the synapse over the plastic prefix of j's incoming list (new_net.jl:126, DASTDP.jl:62-68) — as a wave:
// one thread per spiker fetches the bounds of its list, then one thread per INCOMING SYNAPSE, three in flight.
template <typename T>
__device__ __forceinline__ void ltp_wave(const DevNet<T>& p, const WindowCtx<T>* __restrict__ cx, int32_t k,
                                         const uint32_t bid, const uint32_t nblk, const uint32_t region, WaveSmem& sm) {
  // region: the sd-update log region this block appends to in this step (one block per region and step)
  typedef Ops<T> O;
  tl_begin(p.tl, 2, k);
  const uint32_t tid = threadIdx.x, TB = blockDim.x, lane = tid & 31u;
  const int slot = (cx->slot0 + k) % p.R;
  const int32_t t32 = (int32_t)(cx->t0 + k);
  const uint32_t total = (uint32_t)__ldcg(p.spk_cnt + slot);
  const uint32_t per = (total + nblk - 1) / nblk;
  const uint32_t lo = min(total, bid * per), hi = min(total, lo + per);
  const int32_t* __restrict__ list = p.spk_list + (size_t)slot * (size_t)p.Nld;
  const WView<T> vw = wview(p);
  unsigned long long n_ltp = 0;
  for (uint32_t r0 = lo; r0 < hi; r0 += kWaveItems) {
    const uint32_t nitem = min((uint32_t)kWaveItems, hi - r0);
    for (uint32_t x = tid; x < nitem; x += TB) {
      const uint32_t j = (uint32_t)__ldcg(list + r0 + x);
      const uint32_t c0 = p.colptr[j], npl = p.col_npl[j];
      sm.s0[x] = c0; sm.src[x] = j; sm.off[x + 1] = npl;
      n_ltp += npl;
    }
    __syncthreads();
    if (tid < 32) {
      constexpr int PER = kWaveItems / 32;
      const uint32_t b = lane * PER;
      uint32_t acc = 0;
#pragma unroll
      for (int q = 0; q < PER; ++q) if (b + q < nitem) acc += sm.off[b + q + 1];
      uint32_t inc = acc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += up;
      }
      uint32_t run = inc - acc;
      if (lane == 0) sm.off[0] = 0u;
#pragma unroll
      for (int q = 0; q < PER; ++q)
        if (b + q < nitem) { run += sm.off[b + q + 1]; sm.off[b + q + 1] = run; }
    }
    __syncthreads();
    const uint32_t V = sm.off[nitem];
    for (uint32_t v0 = tid; v0 < V; v0 += 3 * TB) {
      uint32_t pk[3], e[3], inst[3]; bool ok[3]; int2 hq[3];
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const uint32_t v = v0 + u * TB;
        ok[u] = v < V; pk[u] = 0; e[u] = 0; inst[u] = 0;
        if (!ok[u]) continue;
        uint32_t a = 0, b = nitem;
        while (b - a > 1) { const uint32_t m = (a + b) >> 1; if (sm.off[m] <= v) a = m; else b = m; }
        const uint32_t kk = sm.s0[a] + (v - sm.off[a]);
        inst[u] = p.B == 1 ? 0u : sm.src[a] / (uint32_t)p.Nper;
        pk[u] = p.in_pre[kk];
        e[u] = p.in_syn[kk];
      }
#pragma unroll
      for (int u = 0; u < 3; ++u) if (ok[u]) hq[u] = hist_fetch(p.hist, pk[u] & kPreMask);
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        if (!ok[u]) continue;
        // pre trace as seen at the synapse: of step t-(d-1) (variant_g: one step earlier)
        const T tr = trace_resolve<T>(p, hq[u], pk[u] & kPreMask, t32 - (int32_t)(pk[u] >> kPreBits) - (p.variant_g ? 1 : 0), t32);
        sd_add<T>(p, vw, e[u], O::mul(p.apost, tr), region, inst[u]);
      }
    }
    __syncthreads();
  }
  for (int o = 16; o > 0; o >>= 1) n_ltp += __shfl_down_sync(0xffffffffu, n_ltp, o);
  if (lane == 0 && n_ltp) atomicAdd(&sm.blk_ev, n_ltp);  // (accumulates over the calls; the caller stores it)
  tl_end(p.tl, 2, k);
}

template <typename T>
__global__ void __launch_bounds__(256)
ltp_wave_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, int32_t k) {
  __shared__ WaveSmem wsm;
  if (threadIdx.x == 0) wsm.blk_ev = 0ull;
  __syncthreads();
  ltp_wave<T>(p, cx, k, blockIdx.x, gridDim.x, p.log_ltp_base + blockIdx.x, wsm);
  __syncthreads();
  if (threadIdx.x == 0 && wsm.blk_ev) p.blk_ltp[blockIdx.x] += wsm.blk_ev;
}

template <typename T>
__global__ void __launch_bounds__(256, 6)
chain_ltp_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, ChainCtl* __restrict__ ctl, const ChainDims d) {
  __shared__ int s_ok;
  __shared__ WaveSmem wsm;
  const int n = cx->n_steps;
  if (threadIdx.x == 0) wsm.blk_ev = 0ull;
  for (int k = 0; k < n; ++k) {
    if (threadIdx.x == 0) {
      const int which[2] = {kcNeuron, kcRemote};
      const uint32_t target[2] = {d.NB * (uint32_t)(k + 1), d.RB * (uint32_t)(k + 1)};
      s_ok = chain_wait_all<2>(ctl, which, target, 40) ? 1 : 0;
    }
    __syncthreads();
    if (!s_ok) break;
    ltp_wave<T>(p, cx, k, blockIdx.x, d.PB, p.log_ltp_base + (blockIdx.x + (uint32_t)k) % d.PB, wsm);
    __syncthreads();
    if (threadIdx.x == 0) chain_arrive(ctl, kcLtp);
  }
  __syncthreads();
  if (threadIdx.x == 0 && wsm.blk_ev) p.blk_ltp[blockIdx.x] += wsm.blk_ev;
}

// chain_remote_kernel (neuron-partitioned runs): replay of the peers' detection, step by step.  remote(k) may start
// as soon as the local neuron role has finished step k-1 (Iacc[t&1] cleared, the step's list slot reset); it then
// waits for the peers' (word, step) pairs themselves.
template <typename T>
__global__ void __launch_bounds__(256)
chain_remote_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, ChainCtl* __restrict__ ctl,
                    const int32_t* __restrict__ sched, const ChainDims d) {
  __shared__ int s_ok;
  const int n = cx->n_steps;
  for (int k = 0; k < n; ++k) {
    if (threadIdx.x == 0) {
      // neuron(k-1) on every block (Iacc cleared, the step's list slot reset) and remote(k-1) on every block (the
      // spike record's offset of step k is the final count of step k-1)
      const int which[2] = {kcNeuron, kcRemote};
      const uint32_t target[2] = {d.NB * (uint32_t)k, d.RB * (uint32_t)k};
      s_ok = chain_wait_all<2>(ctl, which, target, 0) ? 1 : 0;
    }
    __syncthreads();
    if (!s_ok) return;
    remote_body<T>(p, cx, k, 1, blockIdx.x, d.RB, sched[d.o_nd + k], d.ring);
    __syncthreads();
    if (threadIdx.x == 0) chain_arrive(ctl, kcRemote);
  }
}

// log replay of the pass that has just ended (all LB blocks, block-stride over the regions)
template <typename T>
__device__ __forceinline__ void chain_replay(const DevNet<T>& p, const ChainDims& d, const uint32_t bid, const uint32_t nblk) {
  typename DevNet<T>::T2* w = p.wsdX[*reinterpret_cast<const volatile uint32_t*>(&p.lc->state) >> 31];
  // only the regions the chain's blocks write: [0, SB) synapse role, [log_ltp_base, log_ltp_base + PB) ltp role.  This
  // block's share is looked at all at once (one round trip for the counts), then the non-empty ones are replayed.
  __shared__ uint32_t s_cnt[256];
  const uint32_t n_reg = d.SB + d.PB, TB = min(blockDim.x, 256u);
  // block b replays the regions b, b + nblk, b + 2 nblk, ... (the work stays spread over all blocks); thread x looks
  // at the count of the x-th of them, so a block sees all its counts after one round trip
  for (uint32_t base = bid; base < n_reg; base += nblk * TB) {
    const uint32_t mine = base + threadIdx.x * nblk;
    if (threadIdx.x < TB)
      s_cnt[threadIdx.x] = mine < n_reg ? __ldcg(p.log_cnt + (size_t)(mine < d.SB ? mine : p.log_ltp_base + (mine - d.SB)) * kLogStride) : 0u;
    __syncthreads();
    for (uint32_t x = 0; x < TB && base + x * nblk < n_reg; ++x) {
      const uint32_t cnt = s_cnt[x];
      if (!cnt) continue;
      const uint32_t rr = base + x * nblk, r = rr < d.SB ? rr : p.log_ltp_base + (rr - d.SB);
      const uint32_t nn = cnt < p.log_region_cap ? cnt : p.log_region_cap;
      const LogEntry<T>* __restrict__ lg = p.dlog + (size_t)r * p.log_region_cap;
      // four entries of a thread in flight at once (an entry at a time made the replay a chain of ~15 dependent
      // round trips per region: 33 us per learn period at C5)
      for (uint32_t q0 = threadIdx.x; q0 < nn; q0 += 4 * blockDim.x) {
        LogEntry<T> le[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t q = q0 + u * blockDim.x;
          if (q < nn) { le[u].e = __ldcg(&lg[q].e); le[u].d = __ldcg(&lg[q].d); }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (q0 + u * blockDim.x < nn) red_add(&w[le[u].e].y, le[u].d);
      }
      if (threadIdx.x == 0) p.log_cnt[(size_t)r * kLogStride] = 0u;
    }
    __syncthreads();
  }
  if (*reinterpret_cast<const volatile uint32_t*>(&p.lc->spilled)) {
    // (the plain loop, not spill_merge: a call in this role costs the 40-register learn kernel spills in its hot
    // loops; a spill is rare — the log holds >= 4 M entries and the regions rotate with the step)
    const int64_t tid = bid * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)nblk * blockDim.x;
    for (int64_t e = tid; e < p.E; e += nth) {
      const T dv = __ldcg(p.dsd + e);
      if (dv != (T)0) { red_add(&w[e].y, dv); p.dsd[e] = (T)0; }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// chain_learn_kernel: learn!(k) (new_net.jl:135-142) of every learn step of the window, as the asynchronous
// double-buffered TMA pass.  The role's own barrier (kcLearnBar, counts blocks) separates replay | begin |
// tiles | frontier = all.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kChainLearnThreads, 12)  // <= 40 registers: two blocks per SM fit beside the other roles
chain_learn_kernel(DevNet<T> p, const WindowCtx<T>* __restrict__ cx, ChainCtl* __restrict__ ctl,
                   const int32_t* __restrict__ sched, const ChainDims d) {
  typedef typename DevNet<T>::T2 T2;
  extern __shared__ __align__(128) unsigned char pass_smem[];
  __shared__ __align__(8) uint64_t mbar[kPassMaxStages];
  __shared__ uint32_t s_tile;
  __shared__ int s_ok;
  const uint32_t S = (uint32_t)d.pass_stages, bid = blockIdx.x, LB = d.LB;
  const int n = cx->n_steps;
  tl_count_sm(p.tl, 14);
  if (threadIdx.x == 0) {
    for (uint32_t s_ = 0; s_ < S; ++s_) mbar_init(&mbar[s_], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t q = 0, bar = 0;  // chunks consumed by this block ; barriers of the role passed so far
  auto wait1 = [&](int which, uint32_t target) {
    if (threadIdx.x == 0) s_ok = chain_wait(ctl, which, target, 100) ? 1 : 0;
    __syncthreads();
    const bool ok = s_ok != 0;
    __syncthreads();
    return ok;
  };
  auto role_barrier = [&]() {
    __syncthreads();
    ++bar;
    if (threadIdx.x == 0) {
      chain_arrive(ctl, kcLearnBar);
      s_ok = chain_wait(ctl, kcLearnBar, LB * bar, 100) ? 1 : 0;
    }
    __syncthreads();
    const bool ok = s_ok != 0;
    __syncthreads();
    return ok;
  };
  const int32_t* __restrict__ steps = sched + n + 2;
  unsigned long long t_begin = 0;
  for (int i = 0; i <= d.n_learn; ++i) {
    const bool last = i == d.n_learn;           // window end: only the replay of the final pass
    const int k = last ? n - 1 : steps[i];
    // every sd update of the steps up to k has been issued (and none of step k+1 yet: they wait for `begun`)
    if (threadIdx.x == 0) {
      const int which[3] = {kcLtd, kcLtd1, kcLtp};
      const uint32_t target[3] = {d.SBg[0] * syn_ltd_phases(d, 0, k), d.SBg[1] * syn_ltd_phases(d, 1, k), d.PB * (uint32_t)(k + 1)};
      s_ok = chain_wait_all<3>(ctl, which, target, 100) ? 1 : 0;
    }
    __syncthreads();
    if (!s_ok) return;
    __syncthreads();
    // (i > 0: the log is empty — the early replay below has folded it in and was followed by the role's barrier, and
    // nothing is appended between `frontier = all` and the next `begun`)
    if (i == 0) {
      chain_replay<T>(p, d, bid, LB);  // whatever an earlier window's scheduler may have left
      if (!role_barrier()) return;
    }
    if (last) {
      if (bid == 0 && threadIdx.x == 0) p.lc->spilled = 0u;
      break;
    }
    if (bid == 0) {
      tl_begin(p.tl, 3, k);
      learn_begin<T>(p, cx, k);
      __syncthreads();
      if (threadIdx.x == 0) {
        t_begin = gtime_ns();
        __threadfence();
        atomicExch(&ctl->w[kcBegun * 32], (uint32_t)(i + 1));
      }
    }
    if (!wait1(kcBegun, (uint32_t)(i + 1))) return;
    learn_pass_tma_body<T, kChainLearnThreads>(p, d.do_decay, S, mbar, &s_tile, reinterpret_cast<T2*>(pass_smem), q);
    if (!role_barrier()) return;
    if (bid == 0 && threadIdx.x == 0) {
      p.lc->state = (*reinterpret_cast<volatile uint32_t*>(&p.lc->state) & 0x80000000u) | kFrontAll;
      __threadfence();
      atomicAdd(reinterpret_cast<unsigned long long*>(&ctl->w[kcPassNs * 32]), gtime_ns() - t_begin);
      atomicAdd(&ctl->w[kcPassCnt * 32], 1u);
    }
    if (bid == 0) tl_end(p.tl, 3, k);
    // ---- early replay.  The sd updates logged while the pass was in flight are folded in NOW, in the idle time
    // before the next learn step, instead of on its critical path (the replay sat between the last LTD / LTP of a
    // learn step and `begun`: ~20 us of every learn period at C5 during which the neuron role waits).  Only waves
    // that took their view of the frontier before it read `all` can still append: those of the steps whose neuron
    // role had completed when every block of this role knows the frontier is published.
    if (!role_barrier()) return;
    {
      __shared__ uint32_t s_done;
      if (threadIdx.x == 0) s_done = min(ld_acquire_u32(&ctl->w[kcNeuron * 32]) / d.NB, (uint32_t)n);  // steps 0 .. s_done-1
      __syncthreads();
      const int sd_ = (int)s_done;
      if (sd_ >= 1) {
        if (threadIdx.x == 0) {
          const int which[3] = {kcLtd, kcLtd1, kcLtp};
          const uint32_t target[3] = {d.SBg[0] * syn_ltd_phases(d, 0, sd_ - 1), d.SBg[1] * syn_ltd_phases(d, 1, sd_ - 1), d.PB * (uint32_t)sd_};
          s_ok = chain_wait_all<3>(ctl, which, target, 100) ? 1 : 0;
        }
        __syncthreads();
        if (!s_ok) return;
        __syncthreads();
      }
      // (timeline class 4 on one GPU: the early replay — the envelope over the role's blocks, or block 0's own with timeline = 2)
      unsigned long long* const t4 = ((p.exp_flags & 1024) && bid != 0) ? nullptr : p.tl;
      tl_begin(t4, 4, k);
      chain_replay<T>(p, d, bid, LB);
      tl_end(t4, 4, k);
      if (!role_barrier()) return;  // every block's share is in before the next learn! flips the buffers
    }
  }
}

// ----------------------------------------------------------------------------- small helpers
// set, set*decay, (set*decay)*decay, ... until the value no longer changes (0, a denormal, inf): out[0] = length,
// out[1] = converged within cap.  seq == nullptr: count only.
constexpr int32_t kTseqCap = 1 << 24;
template <typename T>
__global__ void tseq_build_kernel(T* seq, int32_t cap, T set, T decay, int32_t* out) {
  T x = set;
  int32_t n = 0, conv = 0;
  while (n < cap) {
    if (seq) seq[n] = x;
    ++n;
    const T y = Ops<T>::mul(x, decay);
    if (y == x) { conv = 1; break; }
    x = y;
  }
  out[0] = n; out[1] = conv;
}
// SNN_FIELD_TRACE: trace after the set of step s, times decay
template <typename T>
__global__ void trace_to_double_kernel(DevNet<T> p, int32_t s, double* __restrict__ out) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < p.N; j += (int64_t)gridDim.x * blockDim.x)
    out[j] = (double)Ops<T>::mul(trace_at<T>(p, (uint32_t)j, s, s), p.trace_decay);
}
template <typename T>
__global__ void init_state_kernel(DevNet<T> p) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < p.Nld; j += (int64_t)gridDim.x * blockDim.x) {
    p.v[j] = p.v_reset;
    p.u[j] = Ops<T>::mul((T)0.2, p.v_reset);
  }
}
template <typename T>
__global__ void add_scalar_kernel(T* a, T x) { *a = Ops<T>::add(*a, x); }
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
static __global__ void segL_kernel(const uint32_t* __restrict__ seg, int64_t N, int Dmax, uint4* __restrict__ out) {
  const int stride = Dmax + 1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = make_uint4(seg[i * stride], seg[i * stride + 1], seg[i * stride + (Dmax < 2 ? Dmax : 2)],
                        seg[i * stride + (Dmax < 3 ? Dmax : 3)]);
}
// The spike record of a window: every step's ids are appended in arrival order (atomics); the API returns them
// ascending inside a step.  One block per step sorts its segment in shared memory (bitonic network, padded to a power
// of two) — a step of C5 has ~1 300 spikes; segments longer than kRecSortMax go through the library sort instead.
constexpr int kRecSortMax = 8192, kRecSortThreads = 1024;
static __global__ void __launch_bounds__(kRecSortThreads)
record_sort_kernel(int32_t* __restrict__ rec, const int32_t* __restrict__ off) {
  __shared__ int32_t s_key[kRecSortMax];
  const int32_t lo = off[blockIdx.x], len = off[blockIdx.x + 1] - lo;
  if (len <= 1 || len > kRecSortMax) return;
  int P = 2;
  while (P < len) P <<= 1;
  for (int i = threadIdx.x; i < P; i += kRecSortThreads) s_key[i] = i < len ? rec[lo + i] : 0x7fffffff;
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int c = threadIdx.x; c < (P >> 1); c += kRecSortThreads) {
        const int i = ((c / stride) * stride << 1) + (c % stride), j = i + stride;  // comparator (i, j)
        const bool up = (i & size) == 0;
        const int32_t a = s_key[i], b = s_key[j];
        if ((a > b) == up) { s_key[i] = b; s_key[j] = a; }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < len; i += kRecSortThreads) rec[lo + i] = s_key[i];
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
// (w, sd) of a few chosen synapses, read between two windows (snn_get_synapses)
template <typename T>
__global__ void gather_syn_kernel(const typename Real2<T>::type* __restrict__ wsd, const uint32_t* __restrict__ pos, int n,
                                  double* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const typename Real2<T>::type ws = wsd[pos[q]];
  out[2 * q] = (double)ws.x;
  out[2 * q + 1] = (double)ws.y;
}
template <typename T>
__global__ void probe_kernel(DevNet<T> p, const uint32_t* __restrict__ pos, int n, double* __restrict__ out, int32_t k) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const typename DevNet<T>::T2 ws = wview(p).cur[pos[q]];
  out[((size_t)k * n + q) * 2] = (double)ws.x;
  out[((size_t)k * n + q) * 2 + 1] = (double)ws.y;
}

// Everything that is not part of the model: scheduling, footprints, diagnostics, the experimental learning mode.
// Set by name through snn_set_param (the names are the quoted strings of Engine::set_param).
struct Options {
  int graph = 1;             // captured window graph / persistent chain (1) or plain stream-ordered launches (0)
  int persistent = -1;       // production-mode windows: -1 auto (the persistent chain of resident role kernels unless the
                             // batch has more than 256 instances), 0 the captured graph of per-step kernels, 1 the chain.
                             // DESIGN.md section 5 has the measurements.  Under a profiler / CUDA_LAUNCH_BLOCKING the chain is
                             // never used: serialised kernels cannot hand steps over to each other.
  int learn_async = 0;       // 0 auto, 1 learn! always in place, 2 always the asynchronous double-buffered pass
  int64_t learn_log_cap = 0; // asynchronous pass: total entries of the sd-update log (0 = by size)
  int pass_blocks = 0, pass_tile_k = 0, pass_stages = 0;  // asynchronous pass of the graph path: blocks per SM, tile (K synapses), TMA stages
  int syn_blocks = 0, ltp_blocks = 0;                     // graph path: side-kernel blocks per SM
  int timeline = 0, serialise_pass = 0, neuron_blocks_256 = 0, lsu_pass = 0;  // diagnostics / A-B switches
  int learn_lazy = 0;        // EXPERIMENTAL event-driven learning: dense re-basing every M learn steps (0 = the reference's rule)
  double learn_lazy_da = 0;  //   ... and at every learn step while da exceeds this (0 = never)
  int chain[7] = {0, 0, 0, 0, 0, 0, 0};  // persistent chain footprints, see Engine::chain_plan
  int chain_look = 0;        // persistent chain: steps of push lookahead (0 = default 2; 1..3)
  int chain_groups = 0;      // persistent chain: 2 = the synapse role works on alternate steps in two block groups
  int waves = 0;             // captured graph: 1 = the chain's load-balanced wave kernels for LTD / pushes / LTP instead of the
                             // group-per-item kernels (measured slower as kernels of their own: 19.9x vs 21.0x real time on C5)
};

// grow-only device buffer
struct DevBuf {
  void* ptr = nullptr; size_t cap = 0;
  // stream-ordered allocation (cudaMallocAsync / cudaFreeAsync on `s`): cudaMalloc and cudaFree wait for the whole
  // device, and a handle that grows a staging buffer while the resident kernels of ANOTHER handle of the process wait
  // for its spikes (several ranks in one process) would never get its turn
  cudaError_t reserve(size_t bytes, cudaStream_t s) {
    if (bytes <= cap) return cudaSuccess;
    if (ptr) cudaFreeAsync(ptr, s);
    ptr = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocAsync(&ptr, want, s);
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
    SNN_CUDA_OK(cudaStreamCreateWithPriority(&remote_stream_, cudaStreamNonBlocking, prio_hi));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg_.device);
    n_sm_ = sms;
    DevNet<T>& p = p_;
    memset(&p, 0, sizeof(p));
    p.N = cfg_.n_neurons; p.Nld = (p.N + 127) / 128 * 128; p.Nper = cfg_.n_per_instance; p.Ne = cfg_.n_exc;
    // ring: R-1 steps of history + the slot being prepared; variant_g reads one step further back
    // ring: see kLag; variant_g reads one step further back
    p.Dmax = cfg_.max_delay; p.G = p.Dmax + kLag + (cfg_.variant_g ? 1 : 0); p.R = p.G + kLag + 1;
    p.B = (int32_t)(p.N / p.Nper); p.W = (int32_t)(p.Nld / 32);
    const size_t N = (size_t)p.Nld;  // padded: vector accesses never leave the allocation
    SNN_CUDA_OK(cudaMalloc(&p.v, N * sizeof(T)));
    SNN_CUDA_OK(cudaMalloc(&p.u, N * sizeof(T)));
    SNN_CUDA_OK(cudaMalloc(&p.ho, N * sizeof(T)));
    SNN_CUDA_OK(cudaMalloc(&p.Iacc, kIaccRing * N * sizeof(T)));
    SNN_CUDA_OK(cudaMalloc(&p.hist, N * sizeof(int2)));
    SNN_CUDA_OK(cudaMalloc(&p.old_last, N * sizeof(int32_t)));
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
    SNN_CUDA_OK(cudaMalloc(&d_chain_, sizeof(ChainCtl)));
    SNN_CUDA_OK(cudaMallocHost(&h_chain_, sizeof(ChainCtl)));
    memset(h_chain_, 0, sizeof(ChainCtl));
    for (auto& e : chain_ev_) SNN_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    SNN_CUDA_OK(cudaMalloc(&d_ctx_, sizeof(WindowCtx<T>)));
    SNN_CUDA_OK(cudaMallocHost(&h_ctx_, sizeof(WindowCtx<T>)));
    SNN_CUDA_OK(cudaMallocHost(&h_blk_, 2 * kMaxSynBlocks * sizeof(unsigned long long)));
    SNN_CUDA_OK(cudaMemsetAsync(p.ho, 0, N * sizeof(T), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.Iacc, 0, kIaccRing * N * sizeof(T), stream_));
    fill_kernel<int32_t><<<grid_for(2 * (int64_t)N), 256, 0, stream_>>>(reinterpret_cast<int32_t*>(p.hist), 2 * (int64_t)N, kNever);
    fill_kernel<int32_t><<<grid_for((int64_t)N), 256, 0, stream_>>>(p.old_last, (int64_t)N, kNever);
    SNN_CUDA_OK(cudaMemsetAsync(p.bits, 0, (size_t)p.R * p.W * sizeof(uint32_t), stream_));
    SNN_CUDA_OK(cudaMalloc(&p.ll, (size_t)p.R * p.W * sizeof(uint2)));
    SNN_CUDA_OK(cudaMemsetAsync(p.ll, 0xFF, (size_t)p.R * p.W * sizeof(uint2), stream_));  // tag 0xffffffff: never a step
    SNN_CUDA_OK(cudaMemsetAsync(p.spk_cnt, 0, (size_t)p.R * sizeof(int32_t), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.da, 0, 2 * (size_t)p.B * sizeof(T), stream_));
    SNN_CUDA_OK(cudaMemsetAsync(p.counters, 0, 4 * sizeof(unsigned long long), stream_));
    load_model_scalars();
    p.ho_from = cfg_.ho_from; p.threshold_ge = cfg_.threshold_ge; p.sd_decay_mode = cfg_.sd_decay_mode;
    p.rate_mode = cfg_.rate_mode; p.plastic_ei = cfg_.plastic_ei; p.noise_mode = cfg_.noise_mode;
    p.exact = cfg_.mode == SNN_MODE_EXACT;
    p.ho_inh = cfg_.ho_inh; p.variant_g = cfg_.variant_g;
    p.part_lo = 0; p.part_hi = p.N;
    p.static_cur = -1;
    { const int rc = apply_options(err); if (rc) return rc; }
    { const int rc = build_tseq(err); if (rc) return rc; }
    (void)chain_kernel_info();  // now, while none of the chain's kernels can be running in this process
    init_state_kernel<T><<<grid_for(p.Nld), 256, 0, stream_>>>(p);
    SNN_CUDA_OK(cudaGetLastError());
    SNN_CUDA_OK(cudaStreamSynchronize(stream_));
    return SNN_OK;
  }

  // trace table: length first (one thread iterates to the fixed point of x -> x*decay), then the values
  int build_tseq(std::string& err) {
    DevNet<T>& p = p_;
    int32_t* d_len = nullptr; int32_t h_len[2] = {0, 0};
    SNN_CUDA_OK(cudaMalloc(&d_len, 2 * sizeof(int32_t)));
    tseq_build_kernel<T><<<1, 1, 0, stream_>>>(nullptr, kTseqCap, p.trace_set, p.trace_decay, d_len);
    SNN_CUDA_OK(cudaMemcpyAsync(h_len, d_len, sizeof(h_len), cudaMemcpyDeviceToHost, stream_));
    SNN_CUDA_OK(cudaStreamSynchronize(stream_));
    if (!h_len[1]) {
      cudaFree(d_len);
      err = "trace_decay: set*decay^n does not reach a fixed point within 2^24 steps (use 0 <= trace_decay < 1, not within 1e-6 of 1)";
      return SNN_ERR_INVALID;
    }
    T* seq = nullptr;
    SNN_CUDA_OK(cudaMalloc(&seq, (size_t)h_len[0] * sizeof(T)));
    tseq_build_kernel<T><<<1, 1, 0, stream_>>>(seq, h_len[0], p.trace_set, p.trace_decay, d_len);
    SNN_CUDA_OK(cudaStreamSynchronize(stream_));
    cudaFree(d_len);
    cudaFree(const_cast<T*>(p.tseq));
    p.tseq = seq; p.tseq_len = h_len[0];
    return SNN_OK;
  }
  int apply_options(std::string& err) {
    DevNet<T>& p = p_;
    use_graph_ = opt_.graph != 0;
    p.exp_flags = (opt_.timeline ? 2 : 0) | (opt_.timeline == 2 ? 1024 : 0) | (opt_.serialise_pass ? 4 : 0) | (opt_.neuron_blocks_256 ? 8 : 0) | (opt_.lsu_pass ? 512 : 0);
    if (opt_.timeline && !p.tl) SNN_CUDA_OK(cudaMalloc(&p.tl, 2 * kTlClasses * kTlSteps * sizeof(unsigned long long)));
    if (!opt_.timeline && p.tl) { cudaFree(p.tl); p.tl = nullptr; }
    return SNN_OK;
  }

  // snn_set_param / snn_get_param: options by name; the model's scalars (snn_config fields) by their field name
  int param_ref(const std::string& n, int** pi, int64_t** pl, double** pd, bool* structural) {
    *pi = nullptr; *pl = nullptr; *pd = nullptr; *structural = false;
    struct E { const char* name; int* i; bool st; };
    const E ints[] = {
        {"graph", &opt_.graph, false}, {"persistent", &opt_.persistent, false}, {"learn_async", &opt_.learn_async, true},
        {"learn_pass_blocks", &opt_.pass_blocks, false}, {"learn_pass_tile_k", &opt_.pass_tile_k, true},
        {"learn_pass_stages", &opt_.pass_stages, false}, {"synapse_blocks_per_sm", &opt_.syn_blocks, true},
        {"ltp_blocks_per_sm", &opt_.ltp_blocks, true}, {"timeline", &opt_.timeline, false},
        {"serialise_learn_pass", &opt_.serialise_pass, false}, {"neuron_blocks_256", &opt_.neuron_blocks_256, false},
        {"lsu_pass", &opt_.lsu_pass, false}, {"learn_lazy", &opt_.learn_lazy, true},
        {"chain_neuron_threads", &opt_.chain[0], false}, {"chain_synapse_blocks", &opt_.chain[1], false},
        {"chain_ltp_blocks", &opt_.chain[2], false}, {"chain_ltp_threads", &opt_.chain[3], false},
        {"chain_learn_blocks", &opt_.chain[4], false}, {"chain_neuron_registers", &opt_.chain[5], false},
        {"chain_learn_stages", &opt_.chain[6], false}, {"chain_lookahead", &opt_.chain_look, false}, {"wave_kernels", &opt_.waves, false}, {"chain_synapse_groups", &opt_.chain_groups, false}};
    for (const E& e : ints) if (n == e.name) { *pi = e.i; *structural = e.st; return 1; }
    if (n == "learn_log_cap") { *pl = &opt_.learn_log_cap; *structural = true; return 1; }
    if (n == "learn_lazy_da") { *pd = &opt_.learn_lazy_da; *structural = true; return 1; }
    return 0;
  }
  double* model_ref(const std::string& n) {
    struct E { const char* name; double* d; };
    const E ds[] = {{"vthresh", &cfg_.vthresh}, {"v_reset", &cfg_.v_reset}, {"a_exc", &cfg_.a_exc}, {"a_inh", &cfg_.a_inh},
                    {"d_exc", &cfg_.d_exc}, {"d_inh", &cfg_.d_inh}, {"trace_set", &cfg_.trace_set},
                    {"trace_decay", &cfg_.trace_decay}, {"apost", &cfg_.apost}, {"apre", &cfg_.apre},
                    {"da_decay", &cfg_.da_decay}, {"learn_base", &cfg_.learn_base}, {"eta", &cfg_.eta},
                    {"wmin", &cfg_.wmin}, {"wmax", &cfg_.wmax}, {"sd_decay", &cfg_.sd_decay}, {"theta", &cfg_.theta},
                    {"ttheta", &cfg_.ttheta}, {"noise_amp", &cfg_.noise_amp}};
    for (const E& e : ds) if (n == e.name) return e.d;
    return nullptr;
  }
  void load_model_scalars() {
    DevNet<T>& p = p_;
    p.vthresh = (T)cfg_.vthresh; p.v_reset = (T)cfg_.v_reset; p.a_exc = (T)cfg_.a_exc; p.a_inh = (T)cfg_.a_inh;
    p.d_exc = (T)cfg_.d_exc; p.d_inh = (T)cfg_.d_inh; p.trace_set = (T)cfg_.trace_set;
    p.trace_decay = (T)cfg_.trace_decay; p.apost = (T)cfg_.apost; p.apre = (T)cfg_.apre;
    p.da_decay = (T)cfg_.da_decay; p.learn_base = (T)cfg_.learn_base; p.eta = (T)cfg_.eta;
    p.wmin = (T)cfg_.wmin; p.wmax = (T)cfg_.wmax; p.sd_decay = (T)cfg_.sd_decay; p.theta = (T)cfg_.theta;
    p.ttheta = (T)cfg_.ttheta; p.noise_amp = cfg_.noise_amp; p.noise_seed = cfg_.noise_seed;
    p.homeostasis = !(cfg_.theta == 0.0 && cfg_.ttheta == 1.0);
  }
  int set_param(const char* name, double value, std::string& err) override {
    if (!name) { err = "null parameter name"; return SNN_ERR_INVALID; }
    SNN_CUDA_OK(cudaSetDevice(cfg_.device));
    const std::string n(name);
    int* pi; int64_t* pl; double* pd; bool structural;
    if (param_ref(n, &pi, &pl, &pd, &structural)) {
      if (structural && have_topo_) { err = "parameter '" + n + "' shapes the synapse storage: set it before the synapses"; return SNN_ERR_STATE; }
      if (value < 0 && n != "persistent") { err = "parameter '" + n + "' must be >= 0"; return SNN_ERR_INVALID; }
      if (pi) *pi = (int)value; else if (pl) *pl = (int64_t)value; else *pd = value;
      drop_graph();
      return apply_options(err);
    }
    if (n == "noise_seed") { cfg_.noise_seed = (uint64_t)value; load_model_scalars(); drop_graph(); return SNN_OK; }
    if (double* d = model_ref(n)) {
      // the reference's scalars are plain mutable fields / Dict entries (net.param["..."], cfg.yml): same here,
      // between windows.  The values are baked into the captured window: it is captured again.
      const double old = *d;
      *d = value;
      if (!(cfg_.wmin <= cfg_.wmax)) { *d = old; err = "wmin must be <= wmax"; return SNN_ERR_INVALID; }
      load_model_scalars();
      drop_graph();
      if (n == "trace_set" || n == "trace_decay") {
        const int rc = build_tseq(err);
        if (rc) { *d = old; load_model_scalars(); build_tseq(err); return rc; }
      }
      return SNN_OK;
    }
    err = "unknown parameter '" + n + "'"; return SNN_ERR_INVALID;
  }
  int get_param(const char* name, double* value, std::string& err) override {
    if (!name || !value) { err = "null argument"; return SNN_ERR_INVALID; }
    const std::string n(name);
    int* pi; int64_t* pl; double* pd; bool structural;
    if (param_ref(n, &pi, &pl, &pd, &structural)) { *value = pi ? (double)*pi : (pl ? (double)*pl : *pd); return SNN_OK; }
    if (n == "noise_seed") { *value = (double)cfg_.noise_seed; return SNN_OK; }
    if (n == "window_path") { *value = last_window_chain_ ? 2.0 : (last_window_graph_ ? 1.0 : 0.0); return SNN_OK; }  // 2 chain, 1 graph, 0 plain launches
    if (double* d = model_ref(n)) { *value = *d; return SNN_OK; }
    err = "unknown parameter '" + n + "'"; return SNN_ERR_INVALID;
  }

  int set_topology(Topo&& topo, const double* h_w, std::string& err) override {
    SNN_CUDA_OK(cudaSetDevice(cfg_.device));
    drop_graph();
    { const int rc = apply_options(err); if (rc) return rc; }
    topo_.release();
    topo_ = topo;
    DevNet<T>& p = p_;
    p.E = topo_.E;
    p.seg = topo_.seg; p.post = topo_.post; p.colptr = topo_.colptr; p.col_npl = topo_.col_npl;
    p.in_syn = topo_.in_syn; p.in_pre = topo_.in_pre;
    free_synapse_state();
    {
      uint4* sl = nullptr;
      SNN_CUDA_OK(cudaMalloc(&sl, (size_t)std::max<int64_t>(1, p.N) * sizeof(uint4)));
      segL_kernel<<<grid_for(p.N), 256, 0, stream_>>>(p.seg, p.N, p.Dmax, sl);
      p.segL = sl;
    }
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
    // "learn_lazy" = M > 0: event-driven learning, re-basing every M learn steps (and at every learn step while
    // da > "learn_lazy_da" when that is > 0); production mode only
    p.lazy = 0;
    if (opt_.learn_lazy > 0) {
      if (p.exact) { err = "event-driven learning (learn_lazy) needs SNN_MODE_FAST"; return SNN_ERR_INVALID; }
      if (cfg_.sd_decay_mode == SNN_SD_DECAY_EVERY_STEP) {
        err = "event-driven learning does not support SNN_SD_DECAY_EVERY_STEP"; return SNN_ERR_INVALID;
      }
      if (n_plastic_ > 0) {
        p.lazy = std::max(1, opt_.learn_lazy);
        p.lazy_da = opt_.learn_lazy_da > 0 ? (T)opt_.learn_lazy_da : (T)1e30;
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
    have_async_ = !p.lazy && !p.exact && opt_.learn_async != 1 && n_plastic_ > 0 &&
                  cfg_.sd_decay_mode != SNN_SD_DECAY_EVERY_STEP;
    if (have_async_) {
      SNN_CUDA_OK(cudaMalloc(&p.wsdX[1], Ea * sizeof(T2)));
      const Launch g = grids();
      p.log_regions = (uint32_t)(g.gA + g.gP); p.log_ltp_base = (uint32_t)g.gA;
      const int64_t log_cap = opt_.learn_log_cap > 0 ? opt_.learn_log_cap
                                                  : std::min<int64_t>(std::max<int64_t>(1 << 22, n_plastic_ / 8), 1ll << 28);
      p.log_region_cap = (uint32_t)std::max<int64_t>(1, log_cap / p.log_regions);
      SNN_CUDA_OK(cudaMalloc(&p.dlog, (size_t)p.log_regions * p.log_region_cap * sizeof(LogEntry<T>)));
      SNN_CUDA_OK(cudaMalloc(&p.log_cnt, (size_t)p.log_regions * kLogStride * sizeof(uint32_t)));
      SNN_CUDA_OK(cudaMemsetAsync(p.log_cnt, 0, (size_t)p.log_regions * kLogStride * sizeof(uint32_t), stream_));
      SNN_CUDA_OK(cudaMalloc(&p.dsd, Ea * sizeof(T)));
      SNN_CUDA_OK(cudaMemsetAsync(p.dsd, 0, Ea * sizeof(T), stream_));
      p.tile_syn = opt_.pass_tile_k > 0 ? (uint32_t)opt_.pass_tile_k * 1024u : kTileSyn;
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
    syn_want_.clear();  // positions cached by snn_get_synapses belong to the previous synapse order
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
      case SNN_FIELD_TRACE: {
        // the reference holds the post-decay value between calls (new_net.jl:130): trace after step t-1, times decay
        if (count != p.N) { err = "count mismatch"; return SNN_ERR_INVALID; }
        if (t_ == 0) { std::fill(host, host + count, 0.0); return SNN_OK; }
        SNN_CUDA_OK(b_stage_.reserve(p.N * sizeof(double), stream_));
        double* d = b_stage_.as<double>();
        trace_to_double_kernel<T><<<grid_for(p.N), 256, 0, stream_>>>(p, (int32_t)(t_ - 1), d);
        SNN_CUDA_OK(cudaMemcpyAsync(host, d, p.N * sizeof(double), cudaMemcpyDeviceToHost, stream_));
        SNN_CUDA_OK(cudaStreamSynchronize(stream_));
        return SNN_OK;
      }
      case SNN_FIELD_DA: src = p.da + (size_t)(t_ & 1) * p.B; n = p.B; break;
      case SNN_FIELD_I:
        if (!p.exact) { err = "FIELD_I is only kept in SNN_MODE_EXACT"; return SNN_ERR_INVALID; }
        src = p.Iacc + (size_t)((t_ - 1) & 1) * p.Nld; break;
      case SNN_FIELD_W: case SNN_FIELD_SD: {
        if (!have_topo_) { err = "synapses not set"; return SNN_ERR_STATE; }
        if (count != p.E) { err = "count mismatch"; return SNN_ERR_INVALID; }
        if (p.E == 0) return SNN_OK;
        { const int rcf = lazy_to_api(err); if (rcf) return rcf; }
        SNN_CUDA_OK(b_stage_.reserve(p.E * sizeof(double), stream_));
        double* d = b_stage_.as<double>();
        wsd_to_caller<T><<<grid_for(p.E), 256, 0, stream_>>>(p.wsdX[cur_], topo_.perm, d, p.E, field == SNN_FIELD_W ? 0 : 1);
        cudaError_t e = cudaMemcpyAsync(host, d, p.E * sizeof(double), cudaMemcpyDeviceToHost, stream_);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
        { const int rcf = lazy_from_api(err); if (rcf) return rcf; }
        if (e != cudaSuccess) { err = cudaGetErrorString(e); return SNN_ERR_CUDA; }
        return SNN_OK;
      }
      default: err = "unknown field"; return SNN_ERR_INVALID;
    }
    if (count != n) { err = "count mismatch"; return SNN_ERR_INVALID; }
    // persistent staging buffer: the closed-loop callers (reward / dopamine experiments) read a few scalars per
    // window, and a cudaMalloc + cudaFree pair per call cost more than the window itself
    SNN_CUDA_OK(b_stage_.reserve(n * sizeof(double), stream_));
    double* d = b_stage_.as<double>();
    to_double_kernel<T><<<grid_for(n), 256, 0, stream_>>>(src, d, n, scale);
    SNN_CUDA_OK(cudaMemcpyAsync(host, d, n * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    SNN_CUDA_OK(cudaStreamSynchronize(stream_));
    return SNN_OK;
  }

  // `net.s[n1, syn]`, `net.sd[n1, syn]` (newnet_DASTDP.jl:57-60): a few synapses by their index in the caller's CSR
  // order, without the 2 x E-element round trip of get_array.  The positions of the last query are remembered: the
  // reference reads the same synapses every 100 ms.
  int get_synapses(const int64_t* idx, int32_t n, double* w, double* sd, std::string& err) override {
    if (!have_topo_) { err = "synapses not set"; return SNN_ERR_STATE; }
    if (n < 1 || n > 4096 || !idx || (!w && !sd)) { err = "get_synapses: 1..4096 indices and at least one output"; return SNN_ERR_INVALID; }
    if (p_.lazy) { err = "get_synapses is not available with event-driven learning (use snn_get_array)"; return SNN_ERR_INVALID; }
    for (int q = 0; q < n; ++q)
      if (idx[q] < 0 || idx[q] >= p_.E) { err = "synapse index out of range"; return SNN_ERR_INVALID; }
    SNN_CUDA_OK(cudaSetDevice(cfg_.device));
    SNN_CUDA_OK(b_syn_pos_.reserve(n * sizeof(uint32_t), stream_));
    SNN_CUDA_OK(b_stage_.reserve((size_t)n * 2 * sizeof(double) + n * sizeof(int64_t), stream_));
    double* d_out = b_stage_.as<double>();
    if (syn_want_.size() != (size_t)n || !std::equal(syn_want_.begin(), syn_want_.end(), idx)) {
      int64_t* d_want = reinterpret_cast<int64_t*>(d_out + 2 * (size_t)n);
      SNN_CUDA_OK(cudaMemcpyAsync(d_want, idx, n * sizeof(int64_t), cudaMemcpyHostToDevice, stream_));
      probe_locate_kernel<<<grid_for(p_.E), 256, 0, stream_>>>(topo_.perm, p_.E, d_want, n, b_syn_pos_.as<uint32_t>());
      syn_want_.assign(idx, idx + n);
    }
    gather_syn_kernel<T><<<(n + 63) / 64, 64, 0, stream_>>>(p_.wsdX[cur_], b_syn_pos_.as<uint32_t>(), n, d_out);
    std::vector<double> h(2 * (size_t)n);
    SNN_CUDA_OK(cudaMemcpyAsync(h.data(), d_out, h.size() * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    SNN_CUDA_OK(cudaStreamSynchronize(stream_));
    for (int q = 0; q < n; ++q) { if (w) w[q] = h[2 * q]; if (sd) sd[q] = h[2 * q + 1]; }
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
    SNN_CUDA_OK(b_stage_.reserve(n * sizeof(double), stream_));
    double* d = b_stage_.as<double>();
    SNN_CUDA_OK(cudaMemcpyAsync(d, host, n * sizeof(double), cudaMemcpyHostToDevice, stream_));
    from_double_kernel<T><<<grid_for(n), 256, 0, stream_>>>(d, dst, n);
    SNN_CUDA_OK(cudaStreamSynchronize(stream_));  // `host` is pageable memory of the caller
    return SNN_OK;
  }

  // net.da += amount between two step! calls (newnet_DASTDP.jl:54-56): one tiny kernel, stream-ordered with the
  // windows, no host round trip (same rounding as the device type: da (T) + amount (T))
  int add_dopamine(int instance, double amount, std::string& err) override {
    if (instance < 0 || instance >= p_.B) { err = "instance out of range"; return SNN_ERR_INVALID; }
    SNN_CUDA_OK(cudaSetDevice(cfg_.device));
    add_scalar_kernel<T><<<1, 1, 0, stream_>>>(p_.da + (size_t)(t_ & 1) * p_.B + instance, (T)amount);
    SNN_CUDA_OK(cudaGetLastError());
    return SNN_OK;
  }

  // ---- checkpoint / resume: EVERYTHING a window depends on, delays in flight included — neuron state, dopamine,
  // (w, sd), the spike history the traces derive from, and the delay ring (per-step spike lists + bitmaps + the
  // currents already pushed).  The blob is only meaningful for a handle with the same configuration and synapses.
  struct CkptHeader {
    uint64_t magic; int32_t version, dtype;
    int64_t N, Nld, E, t;
    int32_t Dmax, R, B, W, cur, pad;
    int64_t list_total;   // entries of the R spike lists, concatenated
    uint64_t topo_sum;    // cheap fingerprint of the synapse layout (sum of post ids and segment table)
  };
  static constexpr uint64_t kCkptMagic = 0x534e4e434b503033ull;  // "SNNCKP03"
  int64_t ckpt_bytes_for(int64_t list_total) const {
    const DevNet<T>& p = p_;
    const int64_t Nld = p.Nld;
    return (int64_t)sizeof(CkptHeader) + Nld * (int64_t)sizeof(T) * 5                 // v, u, ho, Iacc[2]
           + Nld * (int64_t)(sizeof(int2) + sizeof(int32_t))                           // hist, old_last
           + (int64_t)p.R * p.W * (int64_t)sizeof(uint32_t) + (int64_t)p.R * (int64_t)sizeof(int32_t)  // bits, spk_cnt
           + list_total * (int64_t)sizeof(int32_t) + 2 * (int64_t)p.B * (int64_t)sizeof(T)              // lists, da
           + p.E * (int64_t)sizeof(typename DevNet<T>::T2);
  }
  int ckpt_counts(std::vector<int32_t>& cnt, int64_t& total, std::string& err) {
    cnt.assign(p_.R, 0);
    SNN_CUDA_OK(cudaMemcpyAsync(cnt.data(), p_.spk_cnt, (size_t)p_.R * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
    SNN_CUDA_OK(cudaStreamSynchronize(stream_));
    total = 0;
    for (int32_t c : cnt) total += c;
    return SNN_OK;
  }
  uint64_t topo_fingerprint() const { return (uint64_t)p_.E * 0x9E3779B97F4A7C15ull ^ (uint64_t)p_.N * 0xC2B2AE3D27D4EB4Full ^ (uint64_t)n_plastic_; }
  int checkpoint_bytes(int64_t* bytes, std::string& err) override {
    if (!have_topo_) { err = "synapses not set"; return SNN_ERR_STATE; }
    SNN_CUDA_OK(cudaSetDevice(cfg_.device));
    std::vector<int32_t> cnt; int64_t total = 0;
    const int rc = ckpt_counts(cnt, total, err);
    if (rc) return rc;
    *bytes = ckpt_bytes_for(total);
    return SNN_OK;
  }
  int checkpoint(void* buf, int64_t bytes, std::string& err) override {
    if (!have_topo_) { err = "synapses not set"; return SNN_ERR_STATE; }
    SNN_CUDA_OK(cudaSetDevice(cfg_.device));
    { const int rcf = lazy_to_api(err); if (rcf) return rcf; }
    DevNet<T>& p = p_;
    std::vector<int32_t> cnt; int64_t total = 0;
    int rc = ckpt_counts(cnt, total, err);
    if (rc) return rc;
    if (bytes < ckpt_bytes_for(total)) { err = "checkpoint buffer too small (ask snn_checkpoint_bytes right before)"; return SNN_ERR_CAPACITY; }
    unsigned char* o = static_cast<unsigned char*>(buf);
    CkptHeader h; memset(&h, 0, sizeof(h));
    h.magic = kCkptMagic; h.version = 3; h.dtype = cfg_.dtype; h.N = p.N; h.Nld = p.Nld; h.E = p.E; h.t = t_;
    h.Dmax = p.Dmax; h.R = p.R; h.B = p.B; h.W = p.W; h.cur = cur_; h.list_total = total; h.topo_sum = topo_fingerprint();
    memcpy(o, &h, sizeof(h)); o += sizeof(h);
    auto put = [&](const void* d, size_t n) {
      cudaError_t e = n ? cudaMemcpyAsync(o, d, n, cudaMemcpyDeviceToHost, stream_) : cudaSuccess;
      o += n;
      return e;
    };
    const size_t Nld = (size_t)p.Nld;
    SNN_CUDA_OK(put(p.v, Nld * sizeof(T))); SNN_CUDA_OK(put(p.u, Nld * sizeof(T))); SNN_CUDA_OK(put(p.ho, Nld * sizeof(T)));
    SNN_CUDA_OK(put(p.Iacc, 2 * Nld * sizeof(T)));
    SNN_CUDA_OK(put(p.hist, Nld * sizeof(int2))); SNN_CUDA_OK(put(p.old_last, Nld * sizeof(int32_t)));
    SNN_CUDA_OK(put(p.bits, (size_t)p.R * p.W * sizeof(uint32_t)));
    memcpy(o, cnt.data(), (size_t)p.R * sizeof(int32_t)); o += (size_t)p.R * sizeof(int32_t);
    for (int sl = 0; sl < p.R; ++sl) SNN_CUDA_OK(put(p.spk_list + (size_t)sl * Nld, (size_t)cnt[sl] * sizeof(int32_t)));
    SNN_CUDA_OK(put(p.da, 2 * (size_t)p.B * sizeof(T)));
    SNN_CUDA_OK(put(p.wsdX[cur_], (size_t)p.E * sizeof(typename DevNet<T>::T2)));
    SNN_CUDA_OK(cudaStreamSynchronize(stream_));
    return lazy_from_api(err);
  }
  int restore(const void* buf, int64_t bytes, std::string& err) override {
    if (!have_topo_) { err = "set the synapses (same topology) before snn_restore"; return SNN_ERR_STATE; }
    SNN_CUDA_OK(cudaSetDevice(cfg_.device));
    DevNet<T>& p = p_;
    if (bytes < (int64_t)sizeof(CkptHeader)) { err = "not a checkpoint"; return SNN_ERR_INVALID; }
    const unsigned char* in = static_cast<const unsigned char*>(buf);
    CkptHeader h; memcpy(&h, in, sizeof(h)); in += sizeof(h);
    if (h.magic != kCkptMagic || h.version != 3) { err = "not a checkpoint of this library version"; return SNN_ERR_INVALID; }
    if (h.dtype != cfg_.dtype || h.N != p.N || h.Nld != p.Nld || h.E != p.E || h.Dmax != p.Dmax || h.R != p.R || h.B != p.B ||
        h.W != p.W || h.topo_sum != topo_fingerprint()) {
      err = "checkpoint belongs to a different network (dtype / size / delays / synapse count)"; return SNN_ERR_INVALID;
    }
    if (bytes < ckpt_bytes_for(h.list_total)) { err = "checkpoint truncated"; return SNN_ERR_INVALID; }
    { const int rcf = lazy_to_api(err); if (rcf) return rcf; }
    drop_graph();
    auto get = [&](void* d, size_t n) {
      cudaError_t e = n ? cudaMemcpyAsync(d, in, n, cudaMemcpyHostToDevice, stream_) : cudaSuccess;
      in += n;
      return e;
    };
    const size_t Nld = (size_t)p.Nld;
    SNN_CUDA_OK(get(p.v, Nld * sizeof(T))); SNN_CUDA_OK(get(p.u, Nld * sizeof(T))); SNN_CUDA_OK(get(p.ho, Nld * sizeof(T)));
    SNN_CUDA_OK(get(p.Iacc, 2 * Nld * sizeof(T)));
    SNN_CUDA_OK(get(p.hist, Nld * sizeof(int2))); SNN_CUDA_OK(get(p.old_last, Nld * sizeof(int32_t)));
    SNN_CUDA_OK(get(p.bits, (size_t)p.R * p.W * sizeof(uint32_t)));
    std::vector<int32_t> cnt(p.R);
    memcpy(cnt.data(), in, (size_t)p.R * sizeof(int32_t));
    int64_t total = 0;
    for (int32_t c : cnt) { if (c < 0 || c > p.Nld) { err = "checkpoint corrupt (spike-list length)"; return SNN_ERR_INVALID; } total += c; }
    if (total != h.list_total) { err = "checkpoint corrupt (spike-list total)"; return SNN_ERR_INVALID; }
    SNN_CUDA_OK(get(p.spk_cnt, (size_t)p.R * sizeof(int32_t)));
    for (int sl = 0; sl < p.R; ++sl) SNN_CUDA_OK(get(p.spk_list + (size_t)sl * Nld, (size_t)cnt[sl] * sizeof(int32_t)));
    SNN_CUDA_OK(get(p.da, 2 * (size_t)p.B * sizeof(T)));
    const unsigned char* wsd_src = in;
    SNN_CUDA_OK(get(p.wsdX[cur_], (size_t)p.E * sizeof(typename DevNet<T>::T2)));
    if (p.wsdX[1] != p.wsdX[0])  // the other buffer holds the same weights (non-plastic ones are read from either)
      SNN_CUDA_OK(cudaMemcpyAsync(p.wsdX[cur_ ^ 1], wsd_src, (size_t)p.E * sizeof(typename DevNet<T>::T2), cudaMemcpyHostToDevice, stream_));
    SNN_CUDA_OK(cudaStreamSynchronize(stream_));
    t_ = h.t;
    return lazy_from_api(err);
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
    if (!whole && (p_.lazy || opt_.learn_lazy > 0)) {
      err = "event-driven learning (learn_lazy) has not been validated on neuron-partitioned runs"; return SNN_ERR_INVALID;
    }
    drop_graph();
    p_.part_lo = lo; p_.part_hi = hi;
    exchange_ = whole ? nullptr : fn; exchange_user_ = user;
    return SNN_OK;
  }

  int get_timeline(unsigned long long* host, int64_t count, std::string& err) override {
    if (!p_.tl) { err = "timeline recording is off (snn_set_param \"timeline\" = 1 before the synapses are set)"; return SNN_ERR_STATE; }
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
    SNN_CUDA_OK(b_stage_.reserve(p.E * sizeof(double), stream_));
    double* d = b_stage_.as<double>();
    cudaError_t e = cudaMemcpyAsync(d, h, p.E * sizeof(double), cudaMemcpyHostToDevice, stream_);
    // weights go to both buffers (non-plastic ones are read from either); sd lives in the current one
    wsd_from_caller<T><<<grid_for(p.E), 256, 0, stream_>>>(d, topo_.perm, p.wsdX[cur_], p.E, comp, zero_other);
    if (p.wsdX[1] != p.wsdX[0] && comp == 0)
      wsd_from_caller<T><<<grid_for(p.E), 256, 0, stream_>>>(d, topo_.perm, p.wsdX[cur_ ^ 1], p.E, comp, zero_other);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return SNN_ERR_CUDA; }
    return SNN_OK;
  }
  struct CachedGraph {
    std::string key; cudaGraphExec_t exec = nullptr; size_t prof_used = 0; std::vector<int> prof_cls;
    int pass_launches = 0;
    std::vector<cudaEvent_t> prof_events;  // event record nodes are baked into the graph
  };
  void drop_graph() {
    chain_state_[0] = chain_state_[1] = 0;  // topology / partition / stream changed: size the persistent chain again
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
    cudaFree(p.cacc); cudaFree(p.lzH); cudaFree(p.lzK); cudaFree(const_cast<uint4*>(p.segL)); p.segL = nullptr;
    cudaFree(const_cast<uint32_t*>(p.inst_lo));
    p.wsd = p.wsdX[0] = p.wsdX[1] = nullptr; p.dlog = nullptr; p.dsd = nullptr; p.log_cnt = nullptr;
    p.cacc = nullptr; p.lzH = p.lzR = p.lzInv = nullptr; p.lzK = nullptr; p.lazy = 0;
    p.tile_done = nullptr; p.inst_lo = p.inst_hi = nullptr; p.log_regions = p.log_region_cap = p.n_tiles = 0;
  }
  void release() {
    DevNet<T>& p = p_;
    drop_graph();
    cudaFree(p.v); cudaFree(p.u); cudaFree(p.ho); cudaFree(p.Iacc); cudaFree(p.hist); cudaFree(p.old_last);
    cudaFree(const_cast<T*>(p.tseq)); cudaFree(p.bits);
    cudaFree(p.spk_list); cudaFree(p.spk_cnt); cudaFree(p.da); cudaFree(p.counters); cudaFree(p.blk_ev);
    cudaFree(p.blk_ltp); free_synapse_state(); cudaFree(d_ctx_); cudaFree(p.tl); cudaFree(p.lc); cudaFree(p.lg);
    close_peers();
    cudaFree(p.ll); cudaFree(p.xdone);
    if (h_ctx_) cudaFreeHost(h_ctx_);
    if (h_blk_) cudaFreeHost(h_blk_);
    cudaFree(d_chain_); d_chain_ = nullptr;
    if (h_chain_) cudaFreeHost(h_chain_);
    h_chain_ = nullptr;
    for (auto& e : chain_ev_) { if (e) cudaEventDestroy(e); e = nullptr; }
    b_sched_.release();
    b_noise_.release(); b_noise64_.release(); b_ev_off_.release(); b_ev_inst_.release(); b_ev_amt_.release();
    sort_scratch_.release(); b_stage_.release(); b_probe_want_.release(); b_probe_pos_.release(); b_probe_out_.release(); b_syn_pos_.release(); syn_want_.clear();
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
    if (remote_stream_) cudaStreamDestroy(remote_stream_);
    remote_stream_ = nullptr;
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
  cudaStream_t own_side_stream() { return remote_stream_; }
  int gA_cap() { return grids().gA; }
  int gP_cap() { return grids().gP; }
  Launch grids() {
    DevNet<T>& p = p_;
    Launch l;
    const int64_t local_quads = ((p.part_hi >= p.N ? p.Nld : p.part_hi) >> 2) - (p.part_lo >> 2);
    l.gN = (int)std::max<int64_t>(1, std::min<int64_t>((local_quads + 255) / 256, (int64_t)n_sm_ * 8));
    l.gR = (int)std::max<int64_t>(1, std::min<int64_t>(((p.Nld / 4 - local_quads) / 8 + 255) / 256, (int64_t)n_sm_ * 8));  // a thread per foreign bitmap word
    if (gA_ == 0) {
      int per_sm = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, synapse_kernel<T, true, true, true>, 256, 0);
      // "synapse_blocks_per_sm" / "ltp_blocks_per_sm": footprint of the side kernels of the graph path
      // default 6 + 2: the side kernels are persistent (their blocks live as long as the kernel), so
      // whatever they hold is not available to the next neuron kernel when it becomes ready
      const int syn_per_sm = opt_.syn_blocks > 0 ? opt_.syn_blocks : 6;
      gA_ = std::min(kMaxSynBlocks, n_sm_ * std::max(1, std::min(per_sm, syn_per_sm)));
    }
    l.gA = gA_;
    l.gP = std::min(kMaxSynBlocks, n_sm_ * (opt_.ltp_blocks > 0 ? opt_.ltp_blocks : 2));
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
  // Attributes of the chain's kernels, read and set ONCE per process and precision, before any of them runs: querying
  // or changing a function's attributes loads it / may wait for the device, and a handle that does so while another
  // handle's resident kernels are waiting for it (several ranks in one process) would deadlock the exchange.
  struct ChainKernelInfo { bool ok = false; cudaFuncAttributes fa[6]; };
  static const ChainKernelInfo& chain_kernel_info() {
    static const ChainKernelInfo info = [] {
      ChainKernelInfo k;
      int dev = 0, optin = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
      bool ok = cudaFuncGetAttributes(&k.fa[0], chain_neuron_kernel<T, 2>) == cudaSuccess &&
                cudaFuncGetAttributes(&k.fa[1], chain_neuron_kernel<T, 3>) == cudaSuccess &&
                cudaFuncGetAttributes(&k.fa[2], chain_synapse_kernel<T>) == cudaSuccess &&
                cudaFuncGetAttributes(&k.fa[3], chain_ltp_kernel<T>) == cudaSuccess &&
                cudaFuncGetAttributes(&k.fa[4], chain_learn_kernel<T>) == cudaSuccess &&
                cudaFuncGetAttributes(&k.fa[5], chain_remote_kernel<T>) == cudaSuccess;
      // An SM's L1 / shared-memory split cannot change while a block is resident on it: if the first role to arrive
      // configured it for its own (smaller) need, the blocks of a role that needs more would wait for the SM to
      // drain — forever, since the resident role waits for them.  Every role asks for the maximum split.
      const int co = cudaSharedmemCarveoutMaxShared;
      ok = ok && cudaFuncSetAttribute(chain_neuron_kernel<T, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, co) == cudaSuccess;
      ok = ok && cudaFuncSetAttribute(chain_neuron_kernel<T, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, co) == cudaSuccess;
      ok = ok && cudaFuncSetAttribute(chain_synapse_kernel<T>, cudaFuncAttributePreferredSharedMemoryCarveout, co) == cudaSuccess;
      ok = ok && cudaFuncSetAttribute(chain_ltp_kernel<T>, cudaFuncAttributePreferredSharedMemoryCarveout, co) == cudaSuccess;
      ok = ok && cudaFuncSetAttribute(chain_learn_kernel<T>, cudaFuncAttributePreferredSharedMemoryCarveout, co) == cudaSuccess;
      ok = ok && cudaFuncSetAttribute(chain_remote_kernel<T>, cudaFuncAttributePreferredSharedMemoryCarveout, co) == cudaSuccess;
      ok = ok && cudaFuncSetAttribute(chain_neuron_kernel<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)k.fa[0].sharedSizeBytes) == cudaSuccess;
      ok = ok && cudaFuncSetAttribute(chain_neuron_kernel<T, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)k.fa[1].sharedSizeBytes) == cudaSuccess;
      ok = ok && cudaFuncSetAttribute(chain_learn_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPassMaxStages * (int)kPassChunkBytes) == cudaSuccess;
      if (!ok) cudaGetLastError();
      k.ok = ok;
      return k;
    }();
    return info;
  }
  // ncu / nsys / compute-sanitizer inject themselves through these variables; CUDA_LAUNCH_BLOCKING makes launches synchronous
  static bool kernels_are_serialised() {
    static const bool s = [] {
      const char* lb = getenv("CUDA_LAUNCH_BLOCKING");
      return (lb && lb[0] == '1') || getenv("CUDA_INJECTION64_PATH") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") ||
             getenv("NVTX_INJECTION64_PATH") || getenv("NV_NSIGHT_INJECTION_TRANSPORT_TYPE");
    }();
    return s;
  }
  bool chain_plan(bool with_learn, std::string* why);
  bool chain_ok(const StepArgs& a);
  int launch_chain(const StepArgs& a, std::string& err);

  snn_config cfg_;
  Options opt_;
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
  cudaStream_t side2_stream_ = nullptr, chain_stream_ = nullptr, learn_stream_ = nullptr, remote_stream_ = nullptr;
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
  DevBuf b_probe_want_, b_probe_pos_, b_probe_out_, b_stage_, b_syn_pos_;
  std::vector<int64_t> syn_want_;  // the synapses b_syn_pos_ holds the internal positions of (snn_get_synapses)
  std::vector<void*> ipc_opened_;
  snn_exchange_fn exchange_ = nullptr; void* exchange_user_ = nullptr; int exchange_failed_ = 0;
  const std::vector<cudaEvent_t>* graph_prof_events_ = nullptr;
  std::vector<CachedGraph> graphs_;  // a few (length, train pattern, profiling) variants
  // persistent chain
  ChainCtl* d_chain_ = nullptr; ChainCtl* h_chain_ = nullptr;
  DevBuf b_sched_;
  ChainDims chain_{};
  int chain_state_[2] = {0, 0};  // per plan (without / with the learn role): 0 = not planned yet, 1 = fits, -1 = does not fit
  ChainDims chain_plans_[2] = {};
  size_t chain_smem_[2][2] = {};
  size_t chain_smem_neuron_ = 0, chain_smem_learn_ = 0;
  bool last_window_chain_ = false, last_window_graph_ = false;
  cudaEvent_t chain_ev_[4] = {};
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
  if (opt_.learn_async == 2) return true;               // forced (tests on small nets)
  return n_plastic_ * 2 * (int64_t)sizeof(typename DevNet<T>::T2) >= (32ll << 20);
}

template <typename T>
void Engine<T>::launch_pass(cudaStream_t s4, int k, int do_decay, bool prof) {
  DevNet<T>& p = p_;
  // persistent pass: LB blocks per SM stay resident for the whole pass and leave the other
  // thread slots to the step kernels (LB x 256 threads x 8 x 16 B in flight per SM)
  const bool lsu = opt_.lsu_pass != 0;  // experiment: register-staged (LSU) pass instead of TMA
  const int LB = opt_.pass_blocks > 0 ? opt_.pass_blocks : (lsu ? 2 : 3);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)n_sm_ * LB, p.n_tiles));
  if (prof) prof_begin(2, s4, true);
  if (lsu) {
    learn_pass_kernel<T><<<grid, 256, 0, s4>>>(p, k, do_decay);
  } else {
    const int S = opt_.pass_stages >= 2 && opt_.pass_stages <= kPassMaxStages ? opt_.pass_stages : kPassStages;
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
    // side kernels: the load-balanced waves (default) or the group-per-item kernels ("wave_kernels" = 0)
    const bool wv = opt_.waves != 0 && !p.lazy;  // (the event-driven mode forms weights from three arrays: old kernels)
    auto syn = [&](cudaStream_t st, int k, bool ltd, bool push) {
      if (wv) {
        if (ltd) synapse_wave_kernel<T, true><<<g.gA, 256, 0, st>>>(p, d_ctx_, k, 1, push ? 1 : 0);
        else synapse_wave_kernel<T, false><<<g.gA, 256, 0, st>>>(p, d_ctx_, k, 1, 1);
      } else if (ltd && push) synapse_kernel<T, true, true, true><<<g.gA, 256, 0, st>>>(p, d_ctx_, k, g_arr_log2_);
      else if (ltd) synapse_kernel<T, true, false, true><<<g.gA, 256, 0, st>>>(p, d_ctx_, k, g_arr_log2_);
      else synapse_kernel<T, false, true, true><<<g.gA, 256, 0, st>>>(p, d_ctx_, k, g_arr_log2_);
    };
    syn(s2, -1, false, true);  // pushes for step 0
    cudaEventRecord(evPush(0), s2);
    bool s3_used = false, pass_open = false;
    cudaEvent_t evNk[2] = {dep_ev_[0], dep_ev_[11]};
    // neuron blocks of 64 threads (4 K registers each): they slip into an SM as soon as any side block
    // retires instead of waiting for 16 K registers to come free (+0.8 %; "neuron_blocks_256": 256 threads)
    const int nbm = opt_.neuron_blocks_256 ? 1 : 4;
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
        syn(s2, k, true, true);
        cudaEventRecord(evPush(k + 1), s2);
      } else {
        syn(s2, k, true, false);
      }
      cudaEventRecord(evS(k), s2);
      // s3: ltp(k) [-> in-place learn(k)]
      cudaStreamWaitEvent(s3, evNk[k & 1], 0);
      if (wv) ltp_wave_kernel<T><<<g.gP, 256, 0, s3>>>(p, d_ctx_, k);
      else ltp_kernel<T, true><<<g.gP, 256, 0, s3>>>(p, d_ctx_, k, g_ltp_log2_);
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
        if (opt_.serialise_pass) cudaEventRecord(evL, s4);  // experiment: nothing overlaps the pass
        pass_open = true;
        if (push_next) {
          cudaStreamWaitEvent(s2, evL, 0);
          syn(s2, k, false, true);
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
          syn(s2, k, false, true);
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

// ---------------------------------------------------------------------------------------------
// Persistent chain, host side.  chain_plan sizes the four roles so that ALL their blocks are resident
// at once (registers, shared memory, threads and block slots per SM, from the compiled kernels' own
// attributes); if that cannot be had the window falls back to the captured graph.
// ---------------------------------------------------------------------------------------------
template <typename T>
bool Engine<T>::chain_plan(bool with_learn, std::string* why) {
  // two plans: windows without a learn step do not launch the learn role and need not make room for it
  const int pi = with_learn ? 1 : 0;
  if (chain_state_[pi] != 0) { chain_ = chain_plans_[pi]; chain_smem_neuron_ = chain_smem_[pi][0]; chain_smem_learn_ = chain_smem_[pi][1]; return chain_state_[pi] > 0; }
  chain_state_[pi] = -1;
  DevNet<T>& p = p_;
  auto fail = [&](const char* m) { if (why) *why = m; return false; };
  if (p.exact || !use_graph_ || opt_.persistent == 0 || p.lazy) return fail("mode");
  if (kernels_are_serialised()) return fail("a profiler or CUDA_LAUNCH_BLOCKING serialises kernels: resident roles could not hand over");
  if (cfg_.sd_decay_mode == SNN_SD_DECAY_EVERY_STEP) return fail("sd decays every step");
  const bool parted = !(p.part_lo == 0 && p.part_hi == p.N);
  if (exchange_ || (parted && p.n_ranks < 2)) return fail("partitioned with a host-side exchange");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg_.device) != cudaSuccess) return fail("device properties");
  const int64_t local_quads = ((p.part_hi >= p.N ? p.Nld : p.part_hi) >> 2) - (p.part_lo >> 2);
  ChainDims d{};
  d.NB = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(n_sm_, (local_quads + 255) / 256));
  d.qpb = (uint32_t)(((local_quads + d.NB - 1) / d.NB + 31) / 32 * 32);
  d.NB = (uint32_t)((local_quads + d.qpb - 1) / d.qpb);  // no empty blocks
  const size_t smem_n = (size_t)3 * 4 * d.qpb * sizeof(T) + (size_t)4 * d.qpb * sizeof(uint16_t);
  // work-proportional block counts: small nets get small grids (every waiting block polls the counters)
  const int64_t ev_guess = std::max<int64_t>(1, p.E / 64);  // synapse visits per step at ~1.5 % of the neurons firing
  // role footprints per SM (defaults measured on C5; snn_set_param "chain_*" overrides them for sweeps)
  int tune[7] = {256, 2, 1, 128, 2, 0, 3};
  for (int q = 0; q < 7; ++q) if (opt_.chain[q] > 0) tune[q] = opt_.chain[q];
  const int thrN = std::min(kChainNeuronThreads, std::max(32, tune[0] / 32 * 32));
  int syn_per_sm = tune[1], ltp_per_sm = tune[2], learn_per_sm = tune[4];
  const int thrP = std::min(256, std::max(32, tune[3] / 32 * 32)), thrL = kChainLearnThreads;
  int S = std::min(kPassMaxStages, std::max(2, tune[6]));
  size_t smem_l = (size_t)S * kPassChunkBytes;
  const bool lean = thrN > 256 && tune[5] != 64;  // more than 256 threads: the 40-register build of the neuron role, unless "chain_neuron_registers" = 64 (sweeps)
  cudaFuncAttributes fa[5];
  {
    const ChainKernelInfo& ki = chain_kernel_info();
    if (!ki.ok) return fail("kernel attributes");
    fa[0] = ki.fa[lean ? 1 : 0]; fa[1] = ki.fa[2]; fa[2] = ki.fa[3]; fa[3] = ki.fa[4]; fa[4] = ki.fa[5];
  }
  if (smem_n + fa[0].sharedSizeBytes > (size_t)prop.sharedMemPerBlockOptin) return fail("neuron state does not fit in shared memory");
  auto fits = [&](int sp, int lp, int le) {
    const int thr[5] = {thrN, 256, thrP, thrL, 256};
    const int cnt[5] = {1, sp, lp, with_learn ? le : 0, parted ? 1 : 0};
    const size_t dyn[5] = {smem_n, 0, 0, smem_l, 0};
    size_t regs = 0, smem = 0; int threads = 0, blocks = 0;
    for (int r = 0; r < 5; ++r) {
      regs += (size_t)cnt[r] * (size_t)((fa[r].numRegs + 7) / 8 * 8) * thr[r];
      smem += (size_t)cnt[r] * ((dyn[r] + fa[r].sharedSizeBytes + 1024 + 127) / 128 * 128);
      threads += cnt[r] * thr[r]; blocks += cnt[r];
    }
    return regs + 4096 <= (size_t)prop.regsPerMultiprocessor && smem <= (size_t)prop.sharedMemPerMultiprocessor &&
           threads <= prop.maxThreadsPerMultiProcessor && blocks <= 32;
  };
  while (!fits(syn_per_sm, ltp_per_sm, learn_per_sm)) {
    if (S > 2) { --S; smem_l = (size_t)S * kPassChunkBytes; }  // two blocks of two stages stream faster than one block of three (c1_x1024: 375 against 421 us per pass)
    else if (learn_per_sm > 1) --learn_per_sm;
    else if (syn_per_sm > 1) --syn_per_sm;
    else return fail("the four roles do not fit on an SM together");
  }
  d.SB = (uint32_t)std::max<int64_t>(1, std::min<int64_t>((int64_t)n_sm_ * syn_per_sm, (ev_guess + 255) / 256));
  d.PB = (uint32_t)std::max<int64_t>(1, std::min<int64_t>((int64_t)n_sm_ * ltp_per_sm, (ev_guess + 255) / 256));
  d.LB = (uint32_t)std::max<int64_t>(1, std::min<int64_t>((int64_t)n_sm_ * learn_per_sm, std::max<uint32_t>(1, p.n_tiles)));
  if (d.SB > (uint32_t)gA_cap() || d.PB > (uint32_t)gP_cap()) return fail("log regions");
  // two step groups were measured slower on C5 (22.7x vs 24.4x real time: a step with half the blocks takes twice as
  // long, and the pushes right after a learn step are on the neuron role's critical path): one group unless asked
  d.SG = (opt_.chain_groups == 2 && d.SB >= 2) ? 2u : 1u;
  d.SBg[0] = d.SG == 2 ? (d.SB + 1) / 2 : d.SB; d.SBg[1] = d.SG == 2 ? d.SB / 2 : 0u;
  d.RB = 0;
  if (parted) {
    const int64_t remote_words = p.W - local_quads / 8;
    d.RB = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(n_sm_, (remote_words + 255) / 256));
  }
  d.g_arr_log2 = g_arr_log2_; d.g_ltp_log2 = g_ltp_log2_; d.pass_stages = S;
  d.thrN = thrN; d.thrP = thrP; d.thrL = thrL;
  d.la = std::min(kMaxLook, std::max(1, opt_.chain_look > 0 ? opt_.chain_look : 1));
  d.ring = d.la + 1;
  d.do_decay = cfg_.sd_decay_mode == SNN_SD_DECAY_ON_LEARN ? 1 : 0;
  chain_ = d; chain_smem_neuron_ = smem_n; chain_smem_learn_ = smem_l;
  chain_plans_[pi] = d; chain_smem_[pi][0] = smem_n; chain_smem_[pi][1] = smem_l;
  chain_state_[pi] = 1;
  return true;
}

template <typename T>
bool Engine<T>::chain_ok(const StepArgs& a) {
  if (a.n_force > 0 || a.n_cur > 0 || a.n_probes > 0 || profiling_ == 1) return false;
  int dl, dd, last = -1000;
  bool any = false;
  for (int k = 0; k < a.n_steps; ++k)
    if (wants_learn(a, k, &dl, &dd)) {
      if (!have_async_) return false;   // no second (w, sd) buffer: the in-place pass of the graph path
      if (k - last < 2 && opt_.learn_async != 2) return false;  // learn! at every step: nothing to run beside the pass
      last = k; any = true;
    }
  return chain_plan(any, nullptr);
}

template <typename T>
int Engine<T>::launch_chain(const StepArgs& a, std::string& err) {
  DevNet<T> p = p_;
  p.static_cur = -1;  // the (w, sd) view follows LearnCtl
  const int n = a.n_steps;
  std::vector<int32_t> sched(n + 2, 0), steps;
  int dl, dd;
  for (int k = 0; k < n; ++k) {
    const bool learn = wants_learn(a, k, &dl, &dd);
    sched[k + 1] = sched[k] + (learn ? 1 : 0);
    if (learn) steps.push_back(k);
  }
  sched[n + 1] = sched[n];
  sched.insert(sched.end(), steps.begin(), steps.end());
  // ---- push schedule.  A current that arrives at step m is pushed into Iacc[m % ring] as early as the lookahead
  // LA allows, but never before a learn! that precedes its arrival has begun (the weight applies at arrival time):
  //   neuron(s), not a learn step : the spikers' own delays 1..nd[s]   (arrivals s .. s+nd-1)
  //   synapse(k), not a learn step: arrival k+LA                        (one step further than synapse(k-1) got)
  //   synapse(k), learn step / k=-1: arrivals k+1 .. k+LA, after learn!(k) has begun
  // each clipped to the window and to the next learn step.  need[m] = push phases neuron(m) has to see complete
  // before Euler(m-1): with LA = 2 the cycle neuron(k) -> synapse(k) -> neuron(k+3) spans three steps instead of two.
  ChainDims d = chain_;
  const int LA = d.la;
  std::vector<char> is_learn(n + 1, 0);
  for (int k : steps) is_learn[k] = 1;
  std::vector<int> next_learn(n + 2, n);  // smallest learn step > k, for k = -1 .. n-1 (index k+1)
  {
    int nxt = n;
    for (int k = n - 1; k >= -1; --k) { next_learn[k + 1] = nxt; if (k >= 0 && is_learn[k]) nxt = k; }
  }
  std::vector<int32_t> s_lo(n + 1), s_hi(n + 1), s_nd(std::max(1, n)), s_need(n + 1, 1);
  for (int k = -1; k < n; ++k) {
    const int nl = next_learn[k + 1];
    int lo, hi;
    if (k == -1 || is_learn[k]) { lo = k + 1; hi = std::min(std::min(k + LA, n - 1), nl); }
    else { lo = hi = k + LA; if (hi > n - 1 || nl < hi) hi = lo - 1; }
    s_lo[k + 1] = lo; s_hi[k + 1] = hi;
    for (int m = lo; m <= hi; ++m) s_need[m + 1] = std::max(s_need[m + 1], k + 2);  // phases complete after synapse(k)
  }
  for (int k = 0; k < n; ++k) {
    int nd = 1;
    if (!is_learn[k])
      for (int dd2 = 2; dd2 <= LA; ++dd2) { const int m = k + dd2 - 1; if (m <= n - 1 && next_learn[k + 1] >= m) nd = dd2; }
    s_nd[k] = nd;
  }
  // (pushes into a ring buffer must also come after the Euler update that cleared it: they happen at synapse step
  // >= m - LA = (m - ring) + 1, which waits for neuron(m - ring + 1), the one that consumed arrival m - ring)
  d.o_lo = (int32_t)sched.size(); sched.insert(sched.end(), s_lo.begin(), s_lo.end());
  d.o_hi = (int32_t)sched.size(); sched.insert(sched.end(), s_hi.begin(), s_hi.end());
  d.o_nd = (int32_t)sched.size(); sched.insert(sched.end(), s_nd.begin(), s_nd.end());
  d.o_need = (int32_t)sched.size(); sched.insert(sched.end(), s_need.begin(), s_need.end());
  SNN_CUDA_OK(b_sched_.reserve(sched.size() * sizeof(int32_t), stream_));
  SNN_CUDA_OK(cudaMemcpyAsync(b_sched_.ptr, sched.data(), sched.size() * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
  SNN_CUDA_OK(cudaMemsetAsync(d_chain_, 0, sizeof(ChainCtl), stream_));
  d.n_learn = (int32_t)steps.size();
  cudaStream_t s0 = stream_, sN = chain_stream_, s2 = side_stream_, s3 = side2_stream_, s4 = learn_stream_;
  cudaEvent_t evFork = dep_ev_[9];
  // the copy of the schedule comes from pageable memory: it has been staged before the call returns
  SNN_CUDA_OK(cudaEventRecord(evFork, s0));
  SNN_CUDA_OK(cudaStreamWaitEvent(sN, evFork, 0));
  SNN_CUDA_OK(cudaStreamWaitEvent(s2, evFork, 0));
  SNN_CUDA_OK(cudaStreamWaitEvent(s3, evFork, 0));
  // the neuron role first: its blocks (one per SM) hand a step over among themselves and must all be resident
  if (d.thrN > 256 && opt_.chain[5] != 64) chain_neuron_kernel<T, 3><<<d.NB, d.thrN, chain_smem_neuron_, sN>>>(p, d_ctx_, d_chain_, b_sched_.as<int32_t>(), d);
  else chain_neuron_kernel<T, 2><<<d.NB, d.thrN, chain_smem_neuron_, sN>>>(p, d_ctx_, d_chain_, b_sched_.as<int32_t>(), d);
  chain_synapse_kernel<T><<<d.SB, 256, 0, s2>>>(p, d_ctx_, d_chain_, b_sched_.as<int32_t>(), d);
  chain_ltp_kernel<T><<<d.PB, d.thrP, 0, s3>>>(p, d_ctx_, d_chain_, d);
  if (d.RB) {  // neuron-partitioned: the replay of the peers' detection (own stream: it spins on NVLink traffic)
    SNN_CUDA_OK(cudaStreamWaitEvent(own_side_stream(), evFork, 0));
    chain_remote_kernel<T><<<d.RB, 256, 0, own_side_stream()>>>(p, d_ctx_, d_chain_, b_sched_.as<int32_t>(), d);
    SNN_CUDA_OK(cudaEventRecord(dep_ev_[0], own_side_stream()));
    SNN_CUDA_OK(cudaStreamWaitEvent(s0, dep_ev_[0], 0));
  }
  if (d.n_learn > 0) {
    SNN_CUDA_OK(cudaStreamWaitEvent(s4, evFork, 0));
    chain_learn_kernel<T><<<d.LB, d.thrL, chain_smem_learn_, s4>>>(p, d_ctx_, d_chain_, b_sched_.as<int32_t>(), d);
    SNN_CUDA_OK(cudaEventRecord(chain_ev_[3], s4));
    SNN_CUDA_OK(cudaStreamWaitEvent(s0, chain_ev_[3], 0));
  }
  SNN_CUDA_OK(cudaEventRecord(chain_ev_[0], sN));
  SNN_CUDA_OK(cudaEventRecord(chain_ev_[1], s2));
  SNN_CUDA_OK(cudaEventRecord(chain_ev_[2], s3));
  for (int q = 0; q < 3; ++q) SNN_CUDA_OK(cudaStreamWaitEvent(s0, chain_ev_[q], 0));
  SNN_CUDA_OK(cudaMemcpyAsync(h_chain_, d_chain_, sizeof(ChainCtl), cudaMemcpyDeviceToHost, s0));
  last_pass_launches_ = d.n_learn > 0 ? 1 : 0;
  return SNN_OK;
}

template <typename T>
int Engine<T>::step(const StepArgs& a, std::string& err) {
  if (!have_topo_) { err = "snn_step before synapses were set"; return SNN_ERR_STATE; }
  if (a.n_steps <= 0) { err = "n_steps must be > 0"; return SNN_ERR_INVALID; }
  if (t_ + a.n_steps >= (1ll << 30)) { err = "step counter would pass 2^30 (spike steps are kept as 32-bit values): ~12 days simulated"; return SNN_ERR_STATE; }
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
    SNN_CUDA_OK(b_noise_.reserve(cnt * sizeof(T), stream_));
    if (sizeof(T) == 8) {
      SNN_CUDA_OK(cudaMemcpyAsync(b_noise_.ptr, a.noise, cnt * sizeof(double), cudaMemcpyHostToDevice, stream_));
    } else {
      SNN_CUDA_OK(b_noise64_.reserve(cnt * sizeof(double), stream_));
      SNN_CUDA_OK(cudaMemcpyAsync(b_noise64_.ptr, a.noise, cnt * sizeof(double), cudaMemcpyHostToDevice, stream_));
      from_double_kernel<T><<<grid_for((int64_t)cnt), 256, 0, stream_>>>(b_noise64_.as<double>(), b_noise_.as<T>(), (int64_t)cnt);
    }
    cx.noise = b_noise_.as<T>();
  }
  if (a.n_da > 0) {
    SNN_CUDA_OK(b_ev_off_.reserve((n + 1) * sizeof(int32_t), stream_));
    SNN_CUDA_OK(b_ev_inst_.reserve(a.n_da * sizeof(int32_t), stream_));
    SNN_CUDA_OK(b_ev_amt_.reserve(a.n_da * sizeof(double), stream_));
    SNN_CUDA_OK(cudaMemcpyAsync(b_ev_off_.ptr, ev_off.data(), (n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
    SNN_CUDA_OK(cudaMemcpyAsync(b_ev_inst_.ptr, ev_inst.data(), a.n_da * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
    SNN_CUDA_OK(cudaMemcpyAsync(b_ev_amt_.ptr, ev_amt.data(), a.n_da * sizeof(double), cudaMemcpyHostToDevice, stream_));
    cx.ev_off = b_ev_off_.as<int32_t>(); cx.ev_inst = b_ev_inst_.as<int32_t>(); cx.ev_amt = b_ev_amt_.as<double>();
  }
  if (a.n_force > 0) {
    SNN_CUDA_OK(b_force_.reserve(a.n_force * sizeof(int32_t), stream_));
    SNN_CUDA_OK(cudaMemcpyAsync(b_force_.ptr, force_ids.data(), a.n_force * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
  }
  if (a.n_cur > 0) {
    SNN_CUDA_OK(b_cur_idx_.reserve(a.n_cur * sizeof(int32_t), stream_));
    SNN_CUDA_OK(b_cur_val_.reserve(a.n_cur * sizeof(double), stream_));
    SNN_CUDA_OK(cudaMemcpyAsync(b_cur_idx_.ptr, cur_ids.data(), a.n_cur * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
    SNN_CUDA_OK(cudaMemcpyAsync(b_cur_val_.ptr, cur_val.data(), a.n_cur * sizeof(double), cudaMemcpyHostToDevice, stream_));
  }
  if (a.spikes_out) {
    SNN_CUDA_OK(b_rec_.reserve(a.spikes_cap * sizeof(int32_t), stream_));
    SNN_CUDA_OK(b_rec_off_.reserve((n + 1) * sizeof(int32_t), stream_));
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
    SNN_CUDA_OK(b_probe_want_.reserve(a.n_probes * sizeof(int64_t), stream_));
    SNN_CUDA_OK(b_probe_pos_.reserve(a.n_probes * sizeof(uint32_t), stream_));
    SNN_CUDA_OK(b_probe_out_.reserve((size_t)n * a.n_probes * 2 * sizeof(double), stream_));
    SNN_CUDA_OK(cudaMemcpyAsync(b_probe_want_.ptr, a.probe_syn, a.n_probes * sizeof(int64_t), cudaMemcpyHostToDevice, stream_));
    probe_locate_kernel<<<grid_for(p.E), 256, 0, stream_>>>(topo_.perm, p.E, b_probe_want_.as<int64_t>(), a.n_probes,
                                                            b_probe_pos_.as<uint32_t>());
  }
  const bool graphable = use_graph_ && !p.exact && a.n_force == 0 && a.n_cur == 0 && profiling_ != 1 && !exchange_ &&
                         a.n_probes == 0;
  last_pass_launches_ = 0;
  const bool chained = graphable && chain_ok(a);
  last_window_chain_ = chained; last_window_graph_ = graphable && !chained;
  if (chained) {
    int rc = launch_chain(a, err);
    if (rc) return rc;
  } else if (graphable) {
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
    int32_t longest = 0;
    for (int k = 0; k < n; ++k) longest = std::max(longest, h_off[k + 1] - h_off[k]);
    if (longest <= kRecSortMax) {
      if (longest > 1) record_sort_kernel<<<n, kRecSortThreads, 0, stream_>>>(b_rec_.as<int32_t>(), b_rec_off_.as<int32_t>());
    } else {  // a step with more spikes than a block sorts in shared memory (a burst): the library's segmented sort
      int src = sort_segments(b_rec_.as<int32_t>(), total, b_rec_off_.as<int32_t>(), n, stream_, &sort_scratch_, err);
      if (src) return src;
    }
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
  if (chained && h_chain_->w[kcAbort * 32]) {
    err = "persistent chain: a hand-off timed out (role counter " + std::to_string((int)h_chain_->w[kcAbort * 32] - 1) +
          "); the state of the network is undefined";
    return SNN_ERR_STATE;
  }
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
    if (chained) {  // one kernel per role; the passes are timed on the device (%globaltimer, begin -> frontier = all)
      unsigned long long ns = 0;
      memcpy(&ns, &h_chain_->w[kcPassNs * 32], sizeof(ns));
      s.neuron_ms = s.synapse_ms = 0.f;
      s.learn_ms = (float)((double)ns * 1e-6);
      s.learn_launches = (int32_t)h_chain_->w[kcPassCnt * 32];
      s.neuron_launches = 1; s.synapse_launches = 2;
      s.reserved = last_pass_launches_;
    } else if (graphable) {  // launches inside the graph (events exist only around learn, if at all)
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
