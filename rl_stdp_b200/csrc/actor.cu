// actor.cu — batched layered LIF actors (variant F): thousands of tiny networks, one warp each.
//
// Replaces, for a whole population at once, Julia/Modules/Networks/net_arch.jl
//   input_LIF! :142-152, spike_LIF! :154-211, learn! :213-220, learn_critic! :222-235,
//   step_LIF! :256-263, step_LIF_critic! :265-272
// and the decision / episode loop of Julia/Actors/cartpole_fitness.jl:14-114 (norm_cartpole,
// 40-ms spike-count window, left_right, tie -> replayed random action with -0.9 reward) with a
// RESTATED CartPole-v1 (the reference calls OpenAI gym through PyCall; not in the reference).
//
// Mapping: neuron j lives in lane j & 31, slot j >> 5 of its agent's warp (one neuron per lane for the
// [8,8,8] actor: 24 exc + 8 inh; two for nets of up to 64 neurons such as [8,8,8,8,8]).
// Weights e[l] and eligibility de[l] of the agent live in shared memory for the whole window /
// episode; state is read from and written back to HBM once per launch.  fp64 with explicit
// round-to-nearest mul/add (never contracted) so results are bit-identical to the oracle
// (oracle/lif_oracle.c); compiled with -fmad=false as well.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "actor.h"

namespace snn {
namespace {

#define DMUL(a, b) __dmul_rn((a), (b))
#define DADD(a, b) __dadd_rn((a), (b))
#define DSUB(a, b) __dsub_rn((a), (b))

struct ActorDev {
  int32_t n_agents, L, n_exc, n_inh, n, w_tot;
  int32_t arch[kActorMaxLayers], off[kActorMaxLayers + 1], ioff[kActorMaxLayers], woff[kActorMaxLayers];
  double vthresh, c, wei, wie, wmax, theta, ttheta, imin, apost, apre, tde, ttrace, taulif, eta, winh, da_inc;
  double *v, *ho, *trace, *e, *de, *da;  // [A][n], [A][n], [A][n_exc], [A][w_tot], [A][w_tot], [A]
};

struct Lane {          // per-slot view of one neuron (neuron j lives in lane j & 31, slot j >> 5)
  int layer;           // exc layer index, or -1
  int ilayer;          // hidden layer whose inh partner this neuron is, or -1
  int local;           // index inside the layer
};

__device__ __forceinline__ Lane lane_info(const ActorDev& p, int j) {
  Lane li; li.layer = -1; li.ilayer = -1; li.local = 0;
  for (int l = 0; l < p.L; ++l) {
    if (j >= p.off[l] && j < p.off[l + 1]) { li.layer = l; li.local = j - p.off[l]; }
    if (p.ioff[l] >= 0 && j >= p.ioff[l] && j < p.ioff[l] + p.arch[l]) { li.ilayer = l; li.local = j - p.ioff[l]; }
  }
  return li;
}
// the `cnt` (<= 32) spike bits of the neurons off .. off+cnt-1
__device__ __forceinline__ unsigned ext(unsigned long long mask, int off, int cnt) {
  return (unsigned)(mask >> off) & (cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u));
}

// One call of step_LIF!(net, spikes, train) / step_LIF_critic! for the agent owned by this warp.
// NS = neurons per lane (1: nets of <= 32 neurons, 2: <= 64).  se/sde: the agent's weights in shared
// memory, str: trace_e in shared memory.  Returns the spike mask (bit j = neuron j spiked).
template <int NS>
__device__ __forceinline__ unsigned long long lif_step(const ActorDev& p, const Lane (&li)[NS], int lane,
                                                       const double (&intensity)[NS], double (&v)[NS], double (&ho)[NS],
                                                       double& da, double* se, double* sde, double* str, int train,
                                                       int critic) {
  const unsigned FULL = 0xffffffffu;
  const double kk = 1.0 / p.taulif;  // (1/taulif) :151,:181
  bool alive[NS], is_exc[NS], s[NS];
  unsigned long long mask = 0ull;
#pragma unroll
  for (int sl = 0; sl < NS; ++sl) {
    const int j = lane + 32 * sl;
    alive[sl] = j < p.n;
    is_exc[sl] = j < p.n_exc;
    // input_LIF! :142-152 — I = (max-min)*intensity + min with max = imin + 1.0 computed literally
    if (alive[sl] && li[sl].layer == 0) {
      const double gain = DSUB(DADD(p.imin, 1.0), p.imin);
      const double I = DADD(0.0, DADD(DMUL(gain, intensity[sl]), p.imin));
      v[sl] = DADD(v[sl], DMUL(kk, DADD(DADD(-v[sl], I), p.c)));
    }
    // spike_LIF! :154-211
    s[sl] = alive[sl] && (DSUB(v[sl], ho[sl]) > p.vthresh);
    mask |= (unsigned long long)__ballot_sync(FULL, s[sl]) << (32 * sl);
  }
#pragma unroll
  for (int sl = 0; sl < NS; ++sl) {
    if (s[sl] && is_exc[sl]) ho[sl] = DADD(ho[sl], p.theta);  // :160 (every layer)
    if (s[sl]) v[sl] = p.c;
  }
#pragma unroll
  for (int sl = 0; sl < NS; ++sl) {
    // propagate :164-176 — ascending spiker id: exc sources first, then inh partners
    double I = 0.0;
    if (alive[sl]) {
      if (li[sl].layer >= 1) {
        const int l = li[sl].layer - 1;  // source layer
        unsigned m = ext(mask, p.off[l], p.arch[l]);
        while (m) {
          const int np = __ffs(m) - 1; m &= m - 1;
          I = DADD(I, se[p.woff[l] + np + p.arch[l] * li[sl].local]);
        }
        if (p.ioff[li[sl].layer] >= 0) {  // ie[l] = wie .* (ones - I) :29
          unsigned mi = ext(mask, p.ioff[li[sl].layer], p.arch[li[sl].layer]);
          while (mi) {
            const int np = __ffs(mi) - 1; mi &= mi - 1;
            I = DADD(I, DMUL(p.wie, DSUB(1.0, np == li[sl].local ? 1.0 : 0.0)));
          }
        }
      } else if (li[sl].ilayer >= 0) {    // ei[l] = wei .* I(arch[l]) :30
        unsigned m = ext(mask, p.off[li[sl].ilayer], p.arch[li[sl].ilayer]);
        while (m) {
          const int np = __ffs(m) - 1; m &= m - 1;
          I = DADD(I, DMUL(p.wei, np == li[sl].local ? 1.0 : 0.0));
        }
      }
    }
    // lr = arch[1]+1 : n_exc  :179-182 (inhibitory neurons are never integrated in LIF mode)
    if (alive[sl] && li[sl].layer >= 1) {
      double x = DADD(v[sl], I);
      x = DADD(x, DMUL(kk, DADD(-x, p.c)));
      v[sl] = fmax(x, 0.);
    }
    // trace_e[spiked_e] = 1.0 :185
    if (is_exc[sl] && s[sl]) str[lane + 32 * sl] = 1.0;
  }
  __syncwarp();
  // STDP :188-200 with the new traces
  if (mask & (p.n_exc >= 64 ? ~0ull : ((1ull << p.n_exc) - 1ull))) {
    for (int l = 0; l < p.L - 1; ++l) {
      const int nr = p.arch[l], nc = p.arch[l + 1];
      const unsigned mr = ext(mask, p.off[l], nr);
      const unsigned mc = ext(mask, p.off[l + 1], nc);
      if (!(mr | mc)) continue;
      for (int q = lane; q < nr * nc; q += 32) {
        const int i = q % nr, jj = q / nr;  // column-major [i + nr*jj]
        double x = sde[p.woff[l] + q];
        // pre id < post id always, so the reference loop applies LTD (row of i) before LTP (column of j)
        if ((mr >> i) & 1u) x = DSUB(x, DMUL(p.apre, str[p.off[l + 1] + jj]));
        if ((mc >> jj) & 1u) x = DADD(x, DMUL(p.apost, str[p.off[l] + i]));
        sde[p.woff[l] + q] = x;
      }
    }
  }
  __syncwarp();
  // time decay :204-209
#pragma unroll
  for (int sl = 0; sl < NS; ++sl) {
    if (alive[sl]) ho[sl] = DMUL(ho[sl], p.ttheta);
    if (is_exc[sl]) str[lane + 32 * sl] = DMUL(str[lane + 32 * sl], p.ttrace);
  }
  da = DMUL(da, p.tde);  // (tde, not tda)
  for (int q = lane; q < p.w_tot; q += 32) sde[q] = DMUL(sde[q], p.tde);
  __syncwarp();
  if (train) {
    const double g = critic ? DMUL(p.eta, da) : da;  // learn_critic! :226 / learn! :217
    for (int q = lane; q < p.w_tot; q += 32) {
      const double x = DADD(se[q], DMUL(g, sde[q]));
      se[q] = x > p.wmax ? p.wmax : (x < 0.0 ? 0.0 : x);
    }
    if (critic) {  // :229-234 last div(arch[l],4) rows of hidden e[l] <- winh
      __syncwarp();
      for (int l = 1; l < p.L - 1; ++l) {
        const int n_inh = p.arch[l] / 4, nr = p.arch[l], nc = p.arch[l + 1];
        for (int q = lane; q < n_inh * nc; q += 32) {
          const int i = nr - 1 - (q % n_inh), jj = q / n_inh;
          se[p.woff[l] + i + nr * jj] = p.winh;
        }
      }
    }
    __syncwarp();
  }
  return mask;
}

constexpr int kActorTraceSlots = 64;  // shared-memory slots of trace_e per agent

template <int NS>
__device__ __forceinline__ void load_agent(const ActorDev& p, int a, int lane, double (&v)[NS], double (&ho)[NS], double& da,
                                           double* se, double* sde, double* str) {
#pragma unroll
  for (int sl = 0; sl < NS; ++sl) {
    const int j = lane + 32 * sl;
    v[sl] = j < p.n ? p.v[(size_t)a * p.n + j] : 0.0;
    ho[sl] = j < p.n ? p.ho[(size_t)a * p.n + j] : 0.0;
    if (j < p.n_exc) str[j] = p.trace[(size_t)a * p.n_exc + j];
  }
  da = p.da[a];
  for (int q = lane; q < p.w_tot; q += 32) {
    se[q] = p.e[(size_t)a * p.w_tot + q];
    sde[q] = p.de[(size_t)a * p.w_tot + q];
  }
  __syncwarp();
}
template <int NS>
__device__ __forceinline__ void store_agent(const ActorDev& p, int a, int lane, const double (&v)[NS], const double (&ho)[NS],
                                            double da, const double* se, const double* sde, const double* str) {
  __syncwarp();
#pragma unroll
  for (int sl = 0; sl < NS; ++sl) {
    const int j = lane + 32 * sl;
    if (j < p.n) { p.v[(size_t)a * p.n + j] = v[sl]; p.ho[(size_t)a * p.n + j] = ho[sl]; }
    if (j < p.n_exc) p.trace[(size_t)a * p.n_exc + j] = str[j];
  }
  if (lane == 0) p.da[a] = da;
  for (int q = lane; q < p.w_tot; q += 32) {
    p.e[(size_t)a * p.w_tot + q] = se[q];
    p.de[(size_t)a * p.w_tot + q] = sde[q];
  }
}

// n_steps calls of step_LIF! with a constant input (cartpole_fitness.jl:70-79): optional v .= c
// first, last-layer spike counts accumulated from step index >= L-1 (1-based msec >= L-1).
template <int NS>
__global__ void __launch_bounds__(128)
actor_window_kernel(ActorDev p, const double* __restrict__ intensity, int n_steps, int train, int critic,
                    int reset_v, int32_t* __restrict__ counts, int32_t* __restrict__ spikes, int64_t spikes_stride) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int a = blockIdx.x * (blockDim.x >> 5) + wib;
  if (a >= p.n_agents) return;
  double* se = smem + (size_t)wib * (2 * p.w_tot + kActorTraceSlots);
  double* sde = se + p.w_tot;
  double* str = sde + p.w_tot;
  double v[NS], ho[NS], da, inten[NS];
  Lane li[NS];
  int cnt[NS];
  load_agent<NS>(p, a, lane, v, ho, da, se, sde, str);
#pragma unroll
  for (int sl = 0; sl < NS; ++sl) {
    const int j = lane + 32 * sl;
    li[sl] = lane_info(p, j);
    inten[sl] = (li[sl].layer == 0 && j < p.n) ? intensity[(size_t)a * p.arch[0] + li[sl].local] : 0.0;
    if (reset_v && j < p.n) v[sl] = p.c;  // net.v .= c  cartpole_fitness.jl:71
    cnt[sl] = 0;
  }
  for (int msec = 1; msec <= n_steps; ++msec) {
    const unsigned long long m = lif_step<NS>(p, li, lane, inten, v, ho, da, se, sde, str, train, critic);
#pragma unroll
    for (int sl = 0; sl < NS; ++sl)
      if (msec >= p.L - 1 && li[sl].layer == p.L - 1 && ((m >> (lane + 32 * sl)) & 1ull)) ++cnt[sl];
    if (spikes && lane == 0) spikes[(size_t)a * spikes_stride + (msec - 1)] = (int32_t)(unsigned)m;
  }
#pragma unroll
  for (int sl = 0; sl < NS; ++sl)
    if (counts && li[sl].layer == p.L - 1 && lane + 32 * sl < p.n)
      counts[(size_t)a * p.arch[p.L - 1] + li[sl].local] = cnt[sl];
  store_agent<NS>(p, a, lane, v, ho, da, se, sde, str);
}

// sin/cos restated with the fdlibm kernel polynomials (|x| <= pi/4 — the pole angle never leaves
// +-0.21 rad before termination), in plain IEEE operations so that host oracle and device agree
// bit for bit (libm results may differ in the last ulp between glibc and CUDA).
__device__ __forceinline__ double k_sin(double x) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
               S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  const double z = DMUL(x, x), vv = DMUL(z, x);
  const double r = DADD(S2, DMUL(z, DADD(S3, DMUL(z, DADD(S4, DMUL(z, DADD(S5, DMUL(z, S6))))))));
  return DADD(x, DMUL(vv, DADD(S1, DMUL(z, r))));
}
__device__ __forceinline__ double k_cos(double x) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
               C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  const double z = DMUL(x, x);
  const double r = DMUL(z, DADD(C1, DMUL(z, DADD(C2, DMUL(z, DADD(C3, DMUL(z, DADD(C4, DMUL(z, DADD(C5, DMUL(z, C6)))))))))));
  return DADD(DSUB(1.0, DMUL(0.5, z)), DMUL(z, r));
}

// CartPole-v1 step, restated (gym classic_control/cartpole.py, Euler integrator)
// gym wraps both environments in a TimeLimit (max_episode_steps of CartPole-v1 / MountainCar-v0): an
// episode ends there whatever n_decisions_max the caller passes
constexpr int kCartPoleMaxSteps = 500, kMountainCarMaxSteps = 200;
__device__ __forceinline__ int cartpole_step(double s[4], int action) {
  const double gravity = 9.8, masscart = 1.0, masspole = 0.1, length = 0.5, force_mag = 10.0, tau = 0.02;
  const double total_mass = DADD(masspole, masscart), polemass_length = DMUL(masspole, length);
  double x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
  const double force = action == 1 ? force_mag : -force_mag;
  const double costheta = k_cos(theta), sintheta = k_sin(theta);
  const double temp = __ddiv_rn(DADD(force, DMUL(DMUL(polemass_length, DMUL(theta_dot, theta_dot)), sintheta)), total_mass);
  const double thetaacc = __ddiv_rn(
      DSUB(DMUL(gravity, sintheta), DMUL(costheta, temp)),
      DMUL(length, DSUB(__ddiv_rn(4.0, 3.0), __ddiv_rn(DMUL(masspole, DMUL(costheta, costheta)), total_mass))));
  const double xacc = DSUB(temp, __ddiv_rn(DMUL(DMUL(polemass_length, thetaacc), costheta), total_mass));
  x = DADD(x, DMUL(tau, x_dot)); x_dot = DADD(x_dot, DMUL(tau, xacc));
  theta = DADD(theta, DMUL(tau, theta_dot)); theta_dot = DADD(theta_dot, DMUL(tau, thetaacc));
  s[0] = x; s[1] = x_dot; s[2] = theta; s[3] = theta_dot;
  const double th = __ddiv_rn(DMUL(DMUL(12.0, 2.0), 3.14159265358979323846), 360.0);
  return x < -2.4 || x > 2.4 || theta < -th || theta > th;
}

// MountainCar-v0 step, restated (gym classic_control/mountain_car.py); cos(3*position) by Cody-Waite
// reduction + the kernel polynomials, the same IEEE operations as oracle/lif_oracle.c::ora_cos_reduced
__device__ __forceinline__ double cos_reduced(double x) {
  const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00,
               pio2_1t = 6.07710050650619224932e-11;
  const double fn = rint(DMUL(x, invpio2));
  const double y = DSUB(DSUB(x, DMUL(fn, pio2_1)), DMUL(fn, pio2_1t));
  switch (((int)fn) & 3) {
    case 0: return k_cos(y);
    case 1: return -k_sin(y);
    case 2: return -k_cos(y);
    default: return k_sin(y);
  }
}
__device__ __forceinline__ int mountcar_step(double s[4], int action) {
  const double min_position = -1.2, max_position = 0.6, max_speed = 0.07, goal_position = 0.5, force = 0.001,
               gravity = 0.0025;
  double position = s[0], velocity = s[1];
  velocity = DADD(velocity, DADD(DMUL((double)(action - 1), force), DMUL(cos_reduced(DMUL(3.0, position)), -gravity)));
  velocity = velocity < -max_speed ? -max_speed : (velocity > max_speed ? max_speed : velocity);
  position = DADD(position, velocity);
  position = position < min_position ? min_position : (position > max_position ? max_position : position);
  if (position == min_position && velocity < 0) velocity = 0;
  s[0] = position; s[1] = velocity;
  return position >= goal_position && velocity >= 0;
}

// play_episode (cartpole_fitness.jl:51-114; ENV 1: mountcar_fitness.jl:48-107) for every agent; optional teacher-style dopamine
// (CartPole/Teacher/teacher_sim.jl:61-80,132-133: da += (action==draw ? 1 : -200) * da_inc).
template <int NS, int ENV>
__global__ void __launch_bounds__(128)
actor_episode_kernel(ActorDev p, const double* __restrict__ matalpha, const double* __restrict__ state0,
                     const int32_t* __restrict__ tie_actions, int tie_stride, int tie_agent_stride, const int32_t* __restrict__ teach_actions,
                     int n_decisions_max, int win_size, int train, int32_t* __restrict__ tie_overflow,
                     double* __restrict__ rewards,
                     int32_t* __restrict__ n_dec, int32_t* __restrict__ n_rand_out, double* __restrict__ state_out) {
  extern __shared__ double smem[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int a = blockIdx.x * (blockDim.x >> 5) + wib;
  if (a >= p.n_agents) return;
  double* se = smem + (size_t)wib * (2 * p.w_tot + kActorTraceSlots);
  double* sde = se + p.w_tot;
  double* str = sde + p.w_tot;
  double v[NS], ho[NS], da;
  Lane li[NS];
  load_agent<NS>(p, a, lane, v, ho, da, se, sde, str);
#pragma unroll
  for (int sl = 0; sl < NS; ++sl) li[sl] = lane_info(p, lane + 32 * sl);
  const int A0 = p.arch[0], Lout = p.arch[p.L - 1], n_left = Lout / 2;
  double st[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) st[c] = state0[(size_t)a * 4 + c];
  constexpr int NOBS = ENV == 0 ? 4 : 2;
  double ma[NS][NOBS];  // row `local` of matalpha (arch[0] x NOBS, column-major)
#pragma unroll
  for (int sl = 0; sl < NS; ++sl)
    for (int c = 0; c < NOBS; ++c)
      ma[sl][c] = (li[sl].layer == 0 && lane + 32 * sl < p.n) ? matalpha[(size_t)a * A0 * NOBS + li[sl].local + A0 * c] : 0.0;
  double tot_reward = 0.0;
  int n_rand = 0, n_tot = 0;
  const double pi = 3.14159265358979323846;
  for (int jdec = 0; jdec < n_decisions_max; ++jdec) {
    double on[NOBS];
    if (ENV == 0) {  // norm_cartpole :14-24
      double r = fmod(DADD(st[2], pi), DMUL(2.0, pi));
      if (r == 0) r = copysign(r, DMUL(2.0, pi)); else if (r < 0) r = DADD(r, DMUL(2.0, pi));
      const double theta = DSUB(r, pi);
      const double cone = __ddiv_rn(DMUL(12.0, pi), 180.0);
      on[0] = __ddiv_rn(DADD(st[0], 2.4), 4.8);
      on[1] = __ddiv_rn(DADD(st[1], 7.5), 15.0);
      on[NOBS - 2] = __ddiv_rn(DADD(theta, cone), DMUL(2.0, cone));
      on[NOBS - 1] = __ddiv_rn(DADD(st[3], 7.5), 15.0);
    } else {  // norm_mountcar mountcar_fitness.jl:13-20
      on[0] = __ddiv_rn(DADD(st[0], 1.2), 1.8);
      on[1] = __ddiv_rn(DADD(st[1], 0.07), 0.14);
    }
    double inten[NS];  // matalpha*norm_cartpole(obs) :67
    int cnt[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      inten[q] = 0.0;
#pragma unroll
      for (int c = 0; c < NOBS; ++c) inten[q] = DADD(inten[q], DMUL(ma[q][c], on[c]));
      if (lane + 32 * q < p.n) v[q] = p.c;  // :71
      cnt[q] = 0;
    }
    for (int msec = 1; msec <= win_size; ++msec) {
      const unsigned long long m = lif_step<NS>(p, li, lane, inten, v, ho, da, se, sde, str, train, 0);
#pragma unroll
      for (int q = 0; q < NS; ++q)
        if (msec >= p.L - 1 && li[q].layer == p.L - 1 && ((m >> (lane + 32 * q)) & 1ull)) ++cnt[q];
    }
    // left_right :36-41 / push_or_not mountcar_fitness.jl:33-39 — counts are small integers: any
    // summation order is exact.  ENV 1: pools of div(n,3), div(n,3) and the rest (:77-79)
    const int n_mid_end = ENV == 0 ? n_left : 2 * (Lout / 3), n_left_e = ENV == 0 ? n_left : Lout / 3;
    int sl = 0, sm = 0, sr = 0;
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      const bool outl = li[q].layer == p.L - 1 && lane + 32 * q < p.n;
      sl += (outl && li[q].local < n_left_e) ? cnt[q] : 0;
      sm += (outl && li[q].local >= n_left_e && li[q].local < n_mid_end) ? cnt[q] : 0;
      sr += (outl && li[q].local >= n_mid_end) ? cnt[q] : 0;
    }
    for (int o = 16; o > 0; o >>= 1) {
      sl += __shfl_xor_sync(FULL, sl, o); sr += __shfl_xor_sync(FULL, sr, o);
      if (ENV != 0) sm += __shfl_xor_sync(FULL, sm, o);
    }
    int action, done;
    if (ENV == 0) {
      action = (sl >= sr) ? 0 : 1;
      if (sl == sr) {  // :87-95
        if (n_rand >= tie_stride) *tie_overflow = 1;  // the replayed rand(rng_action, 0:1) stream ran out
        action = n_rand < tie_stride ? tie_actions[(size_t)a * tie_agent_stride + n_rand] : 0;
        tot_reward = DSUB(tot_reward, 0.9);
        ++n_rand;
      }
      ++n_tot;
      if (train && teach_actions) {  // teacher_sim.jl:92-99,132-133
        const double inc = (action == teach_actions[(size_t)a * n_decisions_max + jdec]) ? 1.0 : -200.0;
        da = DADD(da, DMUL(inc, p.da_inc));
      }
      done = cartpole_step(st, action) | (jdec + 1 >= kCartPoleMaxSteps);  // gym TimeLimit of CartPole-v1
      tot_reward = DADD(tot_reward, 1.0);
    } else {
      action = 0;  // argmax([sl, sm, sr]) - 1: the first maximum
      int best = sl;
      if (sm > best) { action = 1; best = sm; }
      if (sr > best) action = 2;
      double fac = 1.0;  // :83-93 every tied pair draws from its pool, the last draw wins
      const int dif[3] = {sl - sr, sl - sm, sm - sr};
      const int pool[3][2] = {{0, 2}, {0, 1}, {1, 2}};
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (dif[i] == 0) {
          if (n_rand >= tie_stride) *tie_overflow = 1;
          const int pick = n_rand < tie_stride ? tie_actions[(size_t)a * tie_agent_stride + n_rand] : 0;
          action = pool[i][pick & 1];
          fac = 2.0;
          ++n_rand;
        }
      ++n_tot;
      done = mountcar_step(st, action) | (jdec + 1 >= kMountainCarMaxSteps);  // gym TimeLimit of MountainCar-v0
      tot_reward = DADD(tot_reward, DMUL(-1.0, fac));  // :98
    }
    if (done) break;
  }
  if (lane == 0) {
    rewards[a] = tot_reward;
    if (n_dec) n_dec[a] = n_tot;
    if (n_rand_out) n_rand_out[a] = n_rand;
    if (state_out) for (int c = 0; c < 4; ++c) state_out[(size_t)a * 4 + c] = st[c];
  }
  store_agent<NS>(p, a, lane, v, ho, da, se, sde, str);
}

// learn!(net) / learn_critic!(net) on their own (net_arch.jl:213-235), for every agent: e += g * de, clamp to
// [0, wmax] (g = da, critic: eta*da), then for the critic the last div(arch[l],4) rows of the hidden e[l] <- winh
__global__ void actor_learn_kernel(ActorDev p, int critic) {
  const int a = blockIdx.x;
  double* __restrict__ e = p.e + (size_t)a * p.w_tot;
  const double* __restrict__ de = p.de + (size_t)a * p.w_tot;
  const double g = critic ? DMUL(p.eta, p.da[a]) : p.da[a];
  for (int q = threadIdx.x; q < p.w_tot; q += blockDim.x) {
    const double x = DADD(e[q], DMUL(g, de[q]));
    e[q] = x > p.wmax ? p.wmax : (x < 0.0 ? 0.0 : x);
  }
  if (critic) {
    __syncthreads();
    for (int l = 1; l < p.L - 1; ++l) {
      const int n_inh = p.arch[l] / 4, nr = p.arch[l], nc = p.arch[l + 1];
      for (int q = threadIdx.x; q < n_inh * nc; q += blockDim.x) {
        const int i = nr - 1 - (q % n_inh), jj = q / n_inh;
        e[p.woff[l] + i + nr * jj] = p.winh;
      }
    }
  }
}

__global__ void fill_d(double* a, int64_t n, double x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] = x;
}

}  // namespace

// grow-only device buffer
struct ActBuf {
  void* ptr = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&ptr, bytes + bytes / 4 + 256);
    if (e == cudaSuccess) cap = bytes + bytes / 4 + 256;
    return e;
  }
  void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
};

struct Actor::Impl {
  ActorDev d;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ActBuf b_ma, b_s0, b_so, b_rw, b_tie, b_teach, b_nd;  // staging of episodes()
};

#define ACT_OK(expr)                                                       \
  do {                                                                     \
    cudaError_t _e = (expr);                                               \
    if (_e != cudaSuccess) {                                               \
      err = std::string(#expr) + ": " + cudaGetErrorString(_e);            \
      return SNN_ERR_CUDA;                                                 \
    }                                                                      \
  } while (0)

Actor::Actor() : impl_(new Impl) { memset(&impl_->d, 0, sizeof(impl_->d)); }
Actor::~Actor() {
  ActorDev& d = impl_->d;
  cudaSetDevice(impl_->device);
  cudaFree(d.v); cudaFree(d.ho); cudaFree(d.trace); cudaFree(d.e); cudaFree(d.de); cudaFree(d.da);
  impl_->b_ma.release(); impl_->b_s0.release(); impl_->b_so.release(); impl_->b_rw.release();
  impl_->b_tie.release(); impl_->b_teach.release(); impl_->b_nd.release();
  if (impl_->e0) cudaEventDestroy(impl_->e0);
  if (impl_->e1) cudaEventDestroy(impl_->e1);
  if (impl_->stream) cudaStreamDestroy(impl_->stream);
  delete impl_;
}

int Actor::init(const snn_actor_config& c, std::string& err) {
  ActorDev& d = impl_->d;
  impl_->device = c.device;
  cfg_ = c;
  d.n_agents = c.n_agents; d.L = c.n_layers;
  int acc = 0;
  for (int l = 0; l < d.L; ++l) { d.arch[l] = c.arch[l]; d.off[l] = acc; acc += c.arch[l]; }
  d.off[d.L] = acc; d.n_exc = acc;
  int iacc = acc, wacc = 0;
  for (int l = 0; l < d.L; ++l) {
    if (l > 0 && l < d.L - 1) { d.ioff[l] = iacc; iacc += c.arch[l]; } else d.ioff[l] = -1;  // Ind :58-65
    if (l < d.L - 1) { d.woff[l] = wacc; wacc += c.arch[l] * c.arch[l + 1]; }
  }
  d.n = iacc; d.n_inh = iacc - acc; d.w_tot = wacc;
  if (d.n > 64) { err = "actor kernel maps at most two neurons per lane: n_exc + n_inh must be <= 64 (use the sparse engine for larger nets)"; return SNN_ERR_INVALID; }
  for (int l = 0; l < d.L; ++l) if (c.arch[l] < 1 || c.arch[l] > 32) { err = "layer sizes must be in 1..32"; return SNN_ERR_INVALID; }
  if ((size_t)(2 * wacc + kActorTraceSlots) * sizeof(double) > 48 * 1024) { err = "actor weights exceed 48 KB of shared memory per agent"; return SNN_ERR_INVALID; }
  d.vthresh = c.vthresh; d.c = c.c; d.wei = c.wei; d.wie = c.wie; d.wmax = c.wmax; d.theta = c.theta; d.ttheta = c.ttheta;
  d.imin = c.imin; d.apost = c.apost; d.apre = c.apre; d.tde = c.tde; d.ttrace = c.ttrace; d.taulif = c.taulif;
  d.eta = c.eta; d.winh = c.winh; d.da_inc = c.da_inc;
  ACT_OK(cudaSetDevice(c.device));
  ACT_OK(cudaStreamCreateWithFlags(&impl_->stream, cudaStreamNonBlocking));
  ACT_OK(cudaEventCreate(&impl_->e0));
  ACT_OK(cudaEventCreate(&impl_->e1));
  const size_t A = (size_t)d.n_agents;
  ACT_OK(cudaMalloc(&d.v, A * d.n * sizeof(double)));
  ACT_OK(cudaMalloc(&d.ho, A * d.n * sizeof(double)));
  ACT_OK(cudaMalloc(&d.trace, A * d.n_exc * sizeof(double)));
  ACT_OK(cudaMalloc(&d.e, A * d.w_tot * sizeof(double)));
  ACT_OK(cudaMalloc(&d.de, A * d.w_tot * sizeof(double)));
  ACT_OK(cudaMalloc(&d.da, A * sizeof(double)));
  fill_d<<<256, 256, 0, impl_->stream>>>(d.v, (int64_t)(A * d.n), c.c);  // v = c*ones :105
  ACT_OK(cudaMemsetAsync(d.ho, 0, A * d.n * sizeof(double), impl_->stream));
  ACT_OK(cudaMemsetAsync(d.trace, 0, A * d.n_exc * sizeof(double), impl_->stream));
  ACT_OK(cudaMemsetAsync(d.e, 0, A * d.w_tot * sizeof(double), impl_->stream));
  ACT_OK(cudaMemsetAsync(d.de, 0, A * d.w_tot * sizeof(double), impl_->stream));
  ACT_OK(cudaMemsetAsync(d.da, 0, A * sizeof(double), impl_->stream));
  ACT_OK(cudaStreamSynchronize(impl_->stream));
  return SNN_OK;
}

int Actor::field_ptr(int field, int layer, double** ptr, int64_t* per_agent, int64_t* stride, std::string& err) {
  ActorDev& d = impl_->d;
  switch (field) {
    case SNN_FIELD_V: *ptr = d.v; *per_agent = d.n; *stride = d.n; return SNN_OK;
    case SNN_FIELD_HO: *ptr = d.ho; *per_agent = d.n; *stride = d.n; return SNN_OK;
    case SNN_FIELD_TRACE: *ptr = d.trace; *per_agent = d.n_exc; *stride = d.n_exc; return SNN_OK;
    case SNN_FIELD_DA: *ptr = d.da; *per_agent = 1; *stride = 1; return SNN_OK;
    case SNN_FIELD_W: case SNN_FIELD_SD:
      if (layer < 0 || layer >= d.L - 1) { err = "layer out of range"; return SNN_ERR_INVALID; }
      *ptr = (field == SNN_FIELD_W ? d.e : d.de) + d.woff[layer];
      *per_agent = (int64_t)d.arch[layer] * d.arch[layer + 1]; *stride = d.w_tot; return SNN_OK;
    default: err = "unknown actor field"; return SNN_ERR_INVALID;
  }
}

int Actor::get_array(int field, int layer, double* host, int64_t count, std::string& err) {
  double* p; int64_t per, stride;
  int rc = field_ptr(field, layer, &p, &per, &stride, err);
  if (rc) return rc;
  if (count != per * impl_->d.n_agents) { err = "count mismatch"; return SNN_ERR_INVALID; }
  ACT_OK(cudaSetDevice(impl_->device));
  ACT_OK(cudaMemcpy2DAsync(host, per * sizeof(double), p, stride * sizeof(double), per * sizeof(double),
                           impl_->d.n_agents, cudaMemcpyDeviceToHost, impl_->stream));
  ACT_OK(cudaStreamSynchronize(impl_->stream));
  return SNN_OK;
}
int Actor::set_array(int field, int layer, const double* host, int64_t count, std::string& err) {
  double* p; int64_t per, stride;
  int rc = field_ptr(field, layer, &p, &per, &stride, err);
  if (rc) return rc;
  if (count != per * impl_->d.n_agents) { err = "count mismatch"; return SNN_ERR_INVALID; }
  ACT_OK(cudaSetDevice(impl_->device));
  ACT_OK(cudaMemcpy2DAsync(p, stride * sizeof(double), host, per * sizeof(double), per * sizeof(double),
                           impl_->d.n_agents, cudaMemcpyHostToDevice, impl_->stream));
  ACT_OK(cudaStreamSynchronize(impl_->stream));
  return SNN_OK;
}

// agents (warps) per block: four while their weights fit the default 48 KB of dynamic shared memory
static inline int actor_warps(const ActorDev& d) {
  for (int w = 4; w > 1; w >>= 1)
    if ((size_t)w * (2 * d.w_tot + kActorTraceSlots) * sizeof(double) <= 48 * 1024) return w;
  return 1;
}
static inline size_t actor_smem(const ActorDev& d, int warps) { return (size_t)warps * (2 * d.w_tot + kActorTraceSlots) * sizeof(double); }

int Actor::learn(int critic, std::string& err) {
  ACT_OK(cudaSetDevice(impl_->device));
  actor_learn_kernel<<<impl_->d.n_agents, 128, 0, impl_->stream>>>(impl_->d, critic);
  ACT_OK(cudaGetLastError());
  ACT_OK(cudaStreamSynchronize(impl_->stream));
  return SNN_OK;
}

int Actor::window(int n_steps, const double* intensity, int train, int critic, int reset_v, int32_t* counts,
                  int32_t* spike_masks, double* device_ms, std::string& err) {
  ActorDev& d = impl_->d;
  if (n_steps <= 0 || !intensity) { err = "n_steps must be > 0 and intensity non-null"; return SNN_ERR_INVALID; }
  if (spike_masks && d.n > 32) { err = "spike_masks hold 32 neurons: not available for nets with more than 32 neurons"; return SNN_ERR_INVALID; }
  ACT_OK(cudaSetDevice(impl_->device));
  const size_t A = (size_t)d.n_agents;
  double* d_int = nullptr; int32_t *d_cnt = nullptr, *d_spk = nullptr;
  const int Lout = d.arch[d.L - 1];
  cudaError_t e = cudaMalloc(&d_int, A * d.arch[0] * sizeof(double));
  if (e == cudaSuccess && counts) e = cudaMalloc(&d_cnt, A * Lout * sizeof(int32_t));
  if (e == cudaSuccess && spike_masks) e = cudaMalloc(&d_spk, A * n_steps * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_int, intensity, A * d.arch[0] * sizeof(double), cudaMemcpyHostToDevice, impl_->stream);
  if (e == cudaSuccess) {
    const int warps = actor_warps(d), blocks = (int)((A + warps - 1) / warps);
    cudaEventRecord(impl_->e0, impl_->stream);
    if (d.n <= 32)
      actor_window_kernel<1><<<blocks, warps * 32, actor_smem(d, warps), impl_->stream>>>(d, d_int, n_steps, train, critic,
                                                                                           reset_v, d_cnt, d_spk, n_steps);
    else
      actor_window_kernel<2><<<blocks, warps * 32, actor_smem(d, warps), impl_->stream>>>(d, d_int, n_steps, train, critic,
                                                                                           reset_v, d_cnt, d_spk, n_steps);
    cudaEventRecord(impl_->e1, impl_->stream);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && counts) e = cudaMemcpyAsync(counts, d_cnt, A * Lout * sizeof(int32_t), cudaMemcpyDeviceToHost, impl_->stream);
  if (e == cudaSuccess && spike_masks) e = cudaMemcpyAsync(spike_masks, d_spk, A * n_steps * sizeof(int32_t), cudaMemcpyDeviceToHost, impl_->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(impl_->stream);
  if (e == cudaSuccess && device_ms) { float ms = 0; cudaEventElapsedTime(&ms, impl_->e0, impl_->e1); *device_ms = ms; }
  cudaFree(d_int); cudaFree(d_cnt); cudaFree(d_spk);
  if (e != cudaSuccess) { err = std::string("actor window: ") + cudaGetErrorString(e); return SNN_ERR_CUDA; }
  return SNN_OK;
}

int Actor::episodes(int env, const double* matalpha, const double* state0, const int32_t* tie_actions, int tie_stride,
                    const int32_t* teach_actions, int n_decisions_max, int win_size, int train, double* rewards,
                    int32_t* n_dec, int32_t* n_rand, double* state_out, double* device_ms, std::string& err) {
  ActorDev& d = impl_->d;
  // tie_stride < 0: ONE stream of -tie_stride draws shared by all agents — the reference's default (`rng_seed = true`:
  // every episode draws from a fresh MersenneTwister(0), cartpole_fitness.jl:55), 0.8 KB instead of 3.3 MB per call
  const bool tie_shared = tie_stride < 0;
  if (tie_shared) tie_stride = -tie_stride;
  if (!matalpha || !state0 || !tie_actions || !rewards || n_decisions_max <= 0 || win_size <= 0 || tie_stride <= 0) {
    err = "episodes: null argument / non-positive sizes"; return SNN_ERR_INVALID;
  }
  if (env != SNN_ENV_CARTPOLE && env != SNN_ENV_MOUNTAINCAR) { err = "episodes: unknown environment"; return SNN_ERR_INVALID; }
  if (env == SNN_ENV_MOUNTAINCAR && (teach_actions || d.arch[d.L - 1] < 3)) {
    err = "episodes: MountainCar needs >= 3 output neurons and has no teacher signal"; return SNN_ERR_INVALID;
  }
  const size_t nobs = env == SNN_ENV_CARTPOLE ? 4 : 2;
  ACT_OK(cudaSetDevice(impl_->device));
  const size_t A = (size_t)d.n_agents;
  // staging buffers live with the population (grow-only): a generation of an evolutionary run calls
  // this once per episode batch, and eight cudaMalloc/cudaFree pairs per call were a third of the
  // end-to-end time of config 4
  Impl& m = *impl_;
  ACT_OK(m.b_ma.reserve(A * d.arch[0] * nobs * sizeof(double)));
  ACT_OK(m.b_s0.reserve(A * 4 * sizeof(double)));
  ACT_OK(m.b_so.reserve(A * 4 * sizeof(double)));
  ACT_OK(m.b_rw.reserve(A * sizeof(double)));
  const size_t tie_rows = tie_shared ? 1 : A;
  ACT_OK(m.b_tie.reserve(tie_rows * (size_t)tie_stride * sizeof(int32_t)));
  if (teach_actions) ACT_OK(m.b_teach.reserve(A * (size_t)n_decisions_max * sizeof(int32_t)));
  ACT_OK(m.b_nd.reserve((2 * A + 1) * sizeof(int32_t)));
  double *d_ma = (double*)m.b_ma.ptr, *d_s0 = (double*)m.b_s0.ptr, *d_rw = (double*)m.b_rw.ptr, *d_so = (double*)m.b_so.ptr;
  int32_t *d_tie = (int32_t*)m.b_tie.ptr, *d_teach = teach_actions ? (int32_t*)m.b_teach.ptr : nullptr;
  int32_t *d_nd = (int32_t*)m.b_nd.ptr, *d_nr = d_nd + A, *d_ovf = d_nd + 2 * A;
  cudaStream_t s = impl_->stream;
  ACT_OK(cudaMemsetAsync(d_ovf, 0, sizeof(int32_t), s));
  ACT_OK(cudaMemcpyAsync(d_ma, matalpha, A * d.arch[0] * nobs * sizeof(double), cudaMemcpyHostToDevice, s));
  ACT_OK(cudaMemcpyAsync(d_s0, state0, A * 4 * sizeof(double), cudaMemcpyHostToDevice, s));
  ACT_OK(cudaMemcpyAsync(d_tie, tie_actions, tie_rows * (size_t)tie_stride * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  if (teach_actions)
    ACT_OK(cudaMemcpyAsync(d_teach, teach_actions, A * (size_t)n_decisions_max * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  {
    const int warps = actor_warps(d), blocks = (int)((A + warps - 1) / warps);
    cudaEventRecord(impl_->e0, s);
#define EPISODE_LAUNCH(NS, ENV)                                                                                      \
  actor_episode_kernel<NS, ENV><<<blocks, warps * 32, actor_smem(d, warps), s>>>(d, d_ma, d_s0, d_tie, tie_stride, tie_shared ? 0 : tie_stride, d_teach, \
                                                                                 n_decisions_max, win_size, train, d_ovf, d_rw, \
                                                                                 d_nd, d_nr, d_so)
    if (env == SNN_ENV_CARTPOLE) { if (d.n <= 32) EPISODE_LAUNCH(1, 0); else EPISODE_LAUNCH(2, 0); }
    else                         { if (d.n <= 32) EPISODE_LAUNCH(1, 1); else EPISODE_LAUNCH(2, 1); }
#undef EPISODE_LAUNCH
    cudaEventRecord(impl_->e1, s);
    ACT_OK(cudaGetLastError());
  }
  int32_t overflow = 0;
  ACT_OK(cudaMemcpyAsync(rewards, d_rw, A * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (n_dec) ACT_OK(cudaMemcpyAsync(n_dec, d_nd, A * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  if (n_rand) ACT_OK(cudaMemcpyAsync(n_rand, d_nr, A * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  if (state_out) ACT_OK(cudaMemcpyAsync(state_out, d_so, A * 4 * sizeof(double), cudaMemcpyDeviceToHost, s));
  ACT_OK(cudaMemcpyAsync(&overflow, d_ovf, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  ACT_OK(cudaStreamSynchronize(s));
  if (device_ms) { float ms = 0; cudaEventElapsedTime(&ms, impl_->e0, impl_->e1); *device_ms = ms; }
  if (overflow) {
    // the reference draws rand(rng_action, ...) at every tie (cartpole_fitness.jl:87-95, up to three per
    // decision in mountcar_fitness.jl:83-93); an exhausted replay stream would silently diverge from it
    err = "episodes: an agent met more ties than tie_stride replayed draws (MountainCar can take 3 per decision)";
    return SNN_ERR_CAPACITY;
  }
  return SNN_OK;
}

}  // namespace snn
