/*
 * lif_oracle.c — CPU restatement (fp64) of the layered networks of
 * Julia/Modules/Networks/net_arch.jl: variant F (LIF actor: input_LIF! :142-152, spike_LIF!
 * :154-211, learn! :213-220, learn_critic! :222-235, step_LIF! :256-272) and variant E
 * (Izhikevich layered: input! :135-137,:316-326, spike! :328-379), plus the CartPole decision
 * loop of Julia/Actors/cartpole_fitness.jl:14-114 and a restated CartPole-v1.
 *
 * TEST INFRASTRUCTURE ONLY (see snn_oracle.c header).  The reference has no tests or golden
 * vectors; what it does ship is results.  Variant F with train = false (LIF step, encoder,
 * read-out, decision loop, CartPole) is PINNED by 60 known answers: the distinct sNES elites
 * saved in Actors/cartpoleElites/sNESelites_cartpole_{1..17}.jld carry the fitness computed from
 * gym's seed-0 reset state (59 x 200.0, one 199.1 = one tie), and ora_play_episode reproduces
 * all 60 (tests/test_oracle_cpu.py::test_reference_elites_known_answers; a 2 % change of vthresh
 * or 4 % of taulif leaves fewer than 10 of the 60), plus ten NON-saturated
 * answers: the MountainCar elites of Actors/sNESelites_mountcar.jld (-129 ... -117), reproduced by
 * ora_lay_step_lif with Julia's MersenneTwister(0) tie stream restated in rl_stdp_b200/julia_rng.py.  The
 * learning rules (learn!, learn_critic!) and variant E stay PARITY UNPINNED BY THE REFERENCE and
 * are checked against the independent numpy restatement oracle/np_restate.py.  CartPole-v1 and
 * MountainCar-v0 physics are NOT in the reference
 * (PyCall -> OpenAI gym, version unpinned, cartpole_fitness.jl:3-4,52,99): restated from the
 * public classic-control definition and labelled as such.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int32_t L;          /* number of layers */
  int32_t arch[8];
  int32_t off[9];     /* off[l] = first exc id of layer l (0-based), off[L] = n_exc */
  int32_t ioff[9];    /* ioff[l] = first inh id of hidden layer l, -1 otherwise */
  int32_t n_exc, n_inh, n;
  int32_t model;      /* 0 = LIF (variant F), 1 = Izhikevich (variant E) */
  /* param Dict (Modules/cfg.yml keys) */
  double vthresh, c, ae, ai, de_, di, wei, wie, wmax, theta, ttheta, imin, imax, apost, apre, tda,
      tde, ttrace, taulif, eta, winh;
  double *v, *u, *ho, *I, *trace_e;
  double* e[8];       /* e[l]: arch[l] x arch[l+1], column-major [i + arch[l]*j] */
  double* de[8];
  double da;
} ora_lay;

/* PARAM order for create: vthresh,c,ae,ai,de,di,wei,wie,wmax,theta,ttheta,imin,imax,apost,apre,
 * tda,tde,ttrace,taulif,eta,winh */
ora_lay* ora_lay_create(int32_t model, int32_t L, const int32_t* arch, const double* p) {
  ora_lay* o = (ora_lay*)calloc(1, sizeof(*o));
  o->model = model; o->L = L;
  int32_t acc = 0;
  for (int l = 0; l < L; ++l) { o->arch[l] = arch[l]; o->off[l] = acc; acc += arch[l]; }
  o->off[L] = acc; o->n_exc = acc;
  int32_t iacc = acc;
  for (int l = 0; l < L; ++l) {
    if (l > 0 && l < L - 1) { o->ioff[l] = iacc; iacc += arch[l]; } else o->ioff[l] = -1; /* Ind :58-65 */
  }
  o->n = iacc; o->n_inh = iacc - acc;
  o->vthresh = p[0]; o->c = p[1]; o->ae = p[2]; o->ai = p[3]; o->de_ = p[4]; o->di = p[5];
  o->wei = p[6]; o->wie = p[7]; o->wmax = p[8]; o->theta = p[9]; o->ttheta = p[10];
  o->imin = p[11]; o->imax = p[12]; o->apost = p[13]; o->apre = p[14]; o->tda = p[15];
  o->tde = p[16]; o->ttrace = p[17]; o->taulif = p[18]; o->eta = p[19]; o->winh = p[20];
  o->v = (double*)malloc(o->n * sizeof(double));
  o->u = (double*)malloc(o->n * sizeof(double));
  o->ho = (double*)calloc(o->n, sizeof(double));
  o->I = (double*)calloc(o->n, sizeof(double));
  o->trace_e = (double*)calloc(o->n_exc, sizeof(double));
  for (int j = 0; j < o->n; ++j) { o->v[j] = o->c * 1.0; o->u[j] = 0.2 * o->v[j]; } /* :105-106 */
  for (int l = 0; l < L - 1; ++l) {
    o->e[l] = (double*)calloc((size_t)arch[l] * arch[l + 1], sizeof(double));
    o->de[l] = (double*)calloc((size_t)arch[l] * arch[l + 1], sizeof(double));
  }
  o->da = 0.0;
  return o;
}
void ora_lay_destroy(ora_lay* o) {
  if (!o) return;
  free(o->v); free(o->u); free(o->ho); free(o->I); free(o->trace_e);
  for (int l = 0; l < o->L - 1; ++l) { free(o->e[l]); free(o->de[l]); }
  free(o);
}
int32_t ora_lay_n(ora_lay* o) { return o->n; }
double* ora_lay_field(ora_lay* o, int which, int l) {
  switch (which) {
    case 0: return o->v; case 1: return o->u; case 2: return o->trace_e; case 3: return o->ho;
    case 4: return o->e[l]; case 5: return o->de[l]; case 6: return &o->da; case 7: return o->I;
    default: return 0;
  }
}

static int layer_of_exc(const ora_lay* o, int n) {
  for (int l = 0; l < o->L; ++l) if (n >= o->off[l] && n < o->off[l + 1]) return l;
  return -1;
}
/* shared by spike_LIF! (:164-176) and spike! (:338-350) */
static void propagate(ora_lay* o, const int32_t* spiked, int k) {
  for (int q = 0; q < k; ++q) {
    int n = spiked[q];
    for (int l = 0; l < o->L - 1; ++l) {
      if (n >= o->off[l] && n < o->off[l + 1]) {
        int neur = n - o->off[l];
        for (int j = 0; j < o->arch[l + 1]; ++j) o->I[o->off[l + 1] + j] += o->e[l][neur + o->arch[l] * j];
        if (o->ioff[l] >= 0) /* ei[l] = wei .* I(arch[l]) :30 */
          for (int j = 0; j < o->arch[l]; ++j) o->I[o->ioff[l] + j] += o->wei * (j == neur ? 1.0 : 0.0);
      }
      if (o->ioff[l] >= 0 && n >= o->ioff[l] && n < o->ioff[l] + o->arch[l]) {
        int neur = n - o->ioff[l];
        for (int j = 0; j < o->arch[l]; ++j) /* ie[l] = wie .* (ones - I) :29 */
          o->I[o->off[l] + j] += o->wie * (1.0 - (j == neur ? 1.0 : 0.0));
      }
    }
  }
}
/* :188-200 / :359-371 */
static void stdp(ora_lay* o, const int32_t* spiked, int k) {
  for (int q = 0; q < k; ++q) {
    int n = spiked[q];
    if (n >= o->n_exc) continue;
    int l = layer_of_exc(o, n);
    int neur = n - o->off[l];
    if (l > 0)
      for (int i = 0; i < o->arch[l - 1]; ++i)
        o->de[l - 1][i + o->arch[l - 1] * neur] += o->apost * o->trace_e[o->off[l - 1] + i];
    if (l < o->L - 1)
      for (int j = 0; j < o->arch[l + 1]; ++j)
        o->de[l][neur + o->arch[l] * j] -= o->apre * o->trace_e[o->off[l + 1] + j];
  }
}

/* input_LIF! :142-152 */
void ora_lay_input_lif(ora_lay* o, const int32_t* idx, const double* intensity, int n_in) {
  memset(o->I, 0, o->n * sizeof(double));
  double max_current = o->imin + 1.0, min_current = o->imin;
  for (int q = 0; q < n_in; ++q) o->I[idx[q]] = 0.0 + ((max_current - min_current) * intensity[q] + min_current);
  double k = 1 / o->taulif;
  for (int j = 0; j < o->arch[0]; ++j) o->v[j] = o->v[j] + k * (-o->v[j] + o->I[j] + o->c);
}

/* spike_LIF! :154-211 ; returns number of spikes, ids ascending */
int ora_lay_spike_lif(ora_lay* o, int32_t* spiked) {
  int k = 0;
  memset(o->I, 0, o->n * sizeof(double));
  for (int j = 0; j < o->n; ++j) if (o->v[j] - o->ho[j] > o->vthresh) spiked[k++] = j;
  for (int q = 0; q < k; ++q) if (spiked[q] < o->n_exc) o->ho[spiked[q]] += o->theta; /* :160 */
  for (int q = 0; q < k; ++q) o->v[spiked[q]] = o->c;
  propagate(o, spiked, k);
  double kk = 1 / o->taulif;
  for (int j = o->arch[0]; j < o->n_exc; ++j) { /* lr :179-182 */
    double v = o->v[j] + o->I[j];
    v = v + kk * (-v + o->c);
    o->v[j] = fmax(v, 0.);
  }
  for (int q = 0; q < k; ++q) if (spiked[q] < o->n_exc) o->trace_e[spiked[q]] = 1.0; /* :185 */
  stdp(o, spiked, k);
  for (int j = 0; j < o->n; ++j) o->ho[j] *= o->ttheta;
  for (int j = 0; j < o->n_exc; ++j) o->trace_e[j] *= o->ttrace;
  o->da *= o->tde; /* :206 (tde, not tda) */
  for (int l = 0; l < o->L - 1; ++l)
    for (int e = 0; e < o->arch[l] * o->arch[l + 1]; ++e) o->de[l][e] *= o->tde;
  return k;
}

static inline double clampd(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }

/* learn! :213-220 */
void ora_lay_learn(ora_lay* o) {
  for (int l = 0; l < o->L - 1; ++l)
    for (int e = 0; e < o->arch[l] * o->arch[l + 1]; ++e)
      o->e[l][e] = clampd(o->e[l][e] + o->da * o->de[l][e], 0, o->wmax);
}
/* learn_critic! :222-235 */
void ora_lay_learn_critic(ora_lay* o) {
  for (int l = 0; l < o->L - 1; ++l)
    for (int e = 0; e < o->arch[l] * o->arch[l + 1]; ++e)
      o->e[l][e] = clampd(o->e[l][e] + o->eta * o->da * o->de[l][e], 0, o->wmax);
  for (int l = 0; l < o->L - 1; ++l) {
    int n_inh = (l == 0 || l == o->L - 1) ? 0 : o->arch[l] / 4; /* arch_inh :118-125 */
    for (int i = 1; i <= n_inh; ++i)
      for (int j = 0; j < o->arch[l + 1]; ++j) o->e[l][(o->arch[l] - i) + o->arch[l] * j] = o->winh;
  }
}

/* step_LIF! :256-263 (critic=0) / step_LIF_critic! :265-272 (critic=1) */
int ora_lay_step_lif(ora_lay* o, const int32_t* idx, const double* intensity, int n_in, int train,
                     int critic, int32_t* spiked) {
  ora_lay_input_lif(o, idx, intensity, n_in);
  int k = ora_lay_spike_lif(o, spiked);
  if (train) { if (critic) ora_lay_learn_critic(o); else ora_lay_learn(o); }
  return k;
}

/* ---- variant E: Izhikevich layered */
/* input!(::Array{Int64}) :135-137 */
void ora_lay_input_force(ora_lay* o, const int32_t* idx, int n_in) {
  for (int q = 0; q < n_in; ++q) o->v[idx[q]] = 40.0;
}
/* input!(::Array{Tuple{Int64,Float64}}) :316-326 */
void ora_lay_input_izh(ora_lay* o, const int32_t* idx, const double* intensity, int n_in) {
  memset(o->I, 0, o->n * sizeof(double));
  for (int q = 0; q < n_in; ++q) o->I[idx[q]] = 0.0 + ((o->imax - o->imin) * intensity[q] + o->imin);
  for (int rep = 0; rep < 2; ++rep)
    for (int q = 0; q < n_in; ++q) {
      int j = idx[q];
      double v = o->v[j];
      o->v[j] = v + 0.5 * ((0.04 * v + 5) * v + 140 - o->u[j] + o->I[j]);
    }
}
/* spike! :328-379 */
int ora_lay_spike_izh(ora_lay* o, int32_t* spiked) {
  int k = 0;
  memset(o->I, 0, o->n * sizeof(double));
  for (int j = 0; j < o->n; ++j) if (o->v[j] - o->ho[j] > o->vthresh) spiked[k++] = j;
  for (int q = 0; q < k; ++q) {
    int n = spiked[q];
    if (n < o->n_exc && n >= o->arch[0]) o->ho[n] += o->theta; /* spiked_ho :333-334 */
  }
  for (int q = 0; q < k; ++q) {
    int n = spiked[q];
    o->v[n] = o->c;
    o->u[n] += (n < o->n_exc) ? o->de_ : o->di;
  }
  propagate(o, spiked, k);
  for (int rep = 0; rep < 2; ++rep)
    for (int j = 0; j < o->n; ++j) {
      double v = o->v[j];
      o->v[j] = v + 0.5 * ((0.04 * v + 5) * v + 140 - o->u[j] + o->I[j]);
    }
  for (int j = 0; j < o->n; ++j) {
    double a = j < o->n_exc ? o->ae : o->ai;
    o->u[j] = o->u[j] + a * (0.2 * o->v[j] - o->u[j]);
  }
  for (int q = 0; q < k; ++q) if (spiked[q] < o->n_exc) o->trace_e[spiked[q]] = 0.1;
  stdp(o, spiked, k);
  for (int j = 0; j < o->n; ++j) o->ho[j] *= o->ttheta;
  for (int j = 0; j < o->n_exc; ++j) o->trace_e[j] *= o->ttrace;
  o->da *= o->tda;
  return k;
}

/* ---- CartPole-v1, restated (NOT pinned by the reference; gym classic_control/cartpole.py).
 * sin/cos are the fdlibm kernel polynomials (|x| <= pi/4; the pole angle never leaves +-0.21 rad
 * before termination) written in plain IEEE operations, so the device implementation
 * (rl_stdp_b200/csrc/actor.cu) can reproduce them bit for bit; libm results may differ in the last
 * ulp between glibc and CUDA. */
static double k_sin(double x) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
               S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double z = x * x, v = z * x;
  double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  return x + v * (S1 + z * r);
}
static double k_cos(double x) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
               C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double z = x * x;
  double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  return (1.0 - 0.5 * z) + z * r;
}
int ora_cartpole_step(double s[4], int action) {
  const double gravity = 9.8, masscart = 1.0, masspole = 0.1, length = 0.5, force_mag = 10.0, tau = 0.02;
  const double total_mass = masspole + masscart, polemass_length = masspole * length;
  double x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
  double force = action == 1 ? force_mag : -force_mag;
  double costheta = k_cos(theta), sintheta = k_sin(theta);
  double temp = (force + polemass_length * (theta_dot * theta_dot) * sintheta) / total_mass;
  double thetaacc = (gravity * sintheta - costheta * temp) /
                    (length * (4.0 / 3.0 - masspole * (costheta * costheta) / total_mass));
  double xacc = temp - polemass_length * thetaacc * costheta / total_mass;
  x = x + tau * x_dot; x_dot = x_dot + tau * xacc;
  theta = theta + tau * theta_dot; theta_dot = theta_dot + tau * thetaacc;
  s[0] = x; s[1] = x_dot; s[2] = theta; s[3] = theta_dot;
  const double th = 12 * 2 * 3.14159265358979323846 / 360;
  return x < -2.4 || x > 2.4 || theta < -th || theta > th;
}
double ora_ksin(double x) { return k_sin(x); }
double ora_kcos(double x) { return k_cos(x); }

/* norm_cartpole cartpole_fitness.jl:14-24 */
void ora_norm_cartpole(const double obs[4], double out[4]) {
  const double pi = M_PI;
  double pos = obs[0];
  /* Julia mod(x,y) for floats: r = rem(x,y) (exact); r==0 -> copysign(r,y); sign(r)!=sign(y) -> r+y */
  double r = fmod(obs[2] + pi, 2 * pi);
  if (r == 0) r = copysign(r, 2 * pi); else if (r < 0) r = r + 2 * pi;
  double theta = r - pi;
  double cone = 12 * pi / 180;
  out[0] = (pos + 2.4) / 4.8;
  out[1] = (obs[1] + 7.5) / 15;
  out[2] = (theta + cone) / (2 * cone);
  out[3] = (obs[3] + 7.5) / 15;
}

/* play_episode cartpole_fitness.jl:51-114 with the network kept by the caller.
 * matalpha: arch[0] x 4 column-major.  tie_actions: replayed stream of the rand(rng_action,0:1)
 * draws (:89).  state0: the env.reset() observation.  Returns tot_reward; n_decisions out. */
double ora_play_episode(ora_lay* o, const double* matalpha, const double state0[4],
                        const int32_t* tie_actions, const int32_t* teach_actions, double da_inc, int n_steps,
                        int train, int32_t* n_dec_out, int32_t* n_rand_out, double* state_out) {
  const int win_size = 40;
  double tot_reward = 0;
  double st[4] = {state0[0], state0[1], state0[2], state0[3]};
  int32_t idx[64]; double inten[64]; int32_t spiked[1024];
  int n_rand = 0, n_tot = 0;
  const int A0 = o->arch[0], Lout = o->arch[o->L - 1];
  for (int j = 0; j < n_steps; ++j) {
    double on[4];
    ora_norm_cartpole(st, on);
    for (int r = 0; r < A0; ++r) { /* matalpha*norm_cartpole(obs) :67 */
      double a = 0.0;
      for (int c = 0; c < 4; ++c) a += matalpha[r + A0 * c] * on[c];
      idx[r] = r; inten[r] = a;
    }
    double count_tot[64] = {0};
    for (int q = 0; q < o->n; ++q) o->v[q] = o->c; /* :71 */
    for (int msec = 1; msec <= win_size; ++msec) {
      int k = ora_lay_step_lif(o, idx, inten, A0, train, 0, spiked);
      if (msec >= o->L - 1) /* :76 */
        for (int q = 0; q < k; ++q) {
          int n = spiked[q];
          if (n >= o->off[o->L - 1] && n < o->off[o->L]) count_tot[n - o->off[o->L - 1]] += 1;
        }
    }
    int n_left = Lout / 2;
    double sl = 0, sr = 0;
    for (int q = 0; q < n_left; ++q) sl += count_tot[q];
    for (int q = n_left; q < Lout; ++q) sr += count_tot[q];
    int action = (sl >= sr) ? 0 : 1; /* :36-41 */
    if (sl == sr) { action = tie_actions[n_rand]; tot_reward -= 0.9; n_rand++; } /* :87-95 */
    n_tot++;
    if (train && teach_actions) { /* teacher_sim.jl:92-99,132-133 */
      double inc = (action == teach_actions[j]) ? 1.0 : -200.0;
      o->da += inc * da_inc;
    }
    int done = ora_cartpole_step(st, action);
    tot_reward += 1;
    if (done) break;
  }
  if (n_dec_out) *n_dec_out = n_tot;
  if (n_rand_out) *n_rand_out = n_rand;
  if (state_out) for (int q = 0; q < 4; ++q) state_out[q] = st[q];
  return tot_reward;
}

/* ---- MountainCar-v0, restated (NOT in the reference: PyCall -> gym classic_control/mountain_car.py).
 * cos(3*position) needs an argument reduction (|3*position| <= 3.6): Cody-Waite with the two leading
 * parts of pi/2 (fdlibm's pio2_1, pio2_1t) and the kernel polynomials above, in plain IEEE operations so
 * the device reproduces it bit for bit (within ~1 ulp of libm; the reference's ten saved MountainCar
 * results are reproduced with it, tests/test_oracle_cpu.py). */
double ora_cos_reduced(double x) {
  const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00,
               pio2_1t = 6.07710050650619224932e-11;
  double fn = rint(x * invpio2);
  double y = (x - fn * pio2_1) - fn * pio2_1t;
  switch (((int)fn) & 3) {
    case 0: return k_cos(y);
    case 1: return -k_sin(y);
    case 2: return -k_cos(y);
    default: return k_sin(y);
  }
}
/* st = (position, velocity); action 0 push left, 1 none, 2 push right; returns done */
int ora_mountcar_step(double st[2], int action) {
  const double min_position = -1.2, max_position = 0.6, max_speed = 0.07, goal_position = 0.5, force = 0.001,
               gravity = 0.0025;
  double position = st[0], velocity = st[1];
  velocity = velocity + ((double)(action - 1) * force + ora_cos_reduced(3 * position) * (-gravity));
  velocity = velocity < -max_speed ? -max_speed : (velocity > max_speed ? max_speed : velocity);
  position = position + velocity;
  position = position < min_position ? min_position : (position > max_position ? max_position : position);
  if (position == min_position && velocity < 0) velocity = 0;
  st[0] = position; st[1] = velocity;
  return position >= goal_position && velocity >= 0;
}

/* play_episode of Actors/mountcar_fitness.jl:48-107.  matalpha: arch[0] x 2 column-major; state0 =
 * (position, velocity) of env.reset(); tie_picks: replayed stream of the `rand(rng_action, pool)` draws
 * (:88) as 0-based indices into the two-element pool, n_picks entries (0 is used beyond). */
double ora_play_episode_mountcar(ora_lay* o, const double* matalpha, const double state0[2],
                                 const int32_t* tie_picks, int n_picks, int n_steps, int train,
                                 int32_t* n_dec_out, int32_t* n_rand_out, double* state_out) {
  const int win_size = 40;
  double tot_reward = 0;
  double st[2] = {state0[0], state0[1]};
  int32_t idx[64]; double inten[64]; int32_t spiked[1024];
  int n_rand = 0, n_tot = 0;
  const int A0 = o->arch[0], Lout = o->arch[o->L - 1];
  for (int j = 0; j < n_steps; ++j) {
    double on[2] = {(st[0] + 1.2) / 1.8, (st[1] + 0.07) / 0.14}; /* norm_mountcar :13-20 */
    for (int r = 0; r < A0; ++r) {
      double a = 0.0;
      for (int c = 0; c < 2; ++c) a += matalpha[r + A0 * c] * on[c];
      idx[r] = r; inten[r] = a;
    }
    double count_tot[64] = {0};
    for (int q = 0; q < o->n; ++q) o->v[q] = o->c; /* :66 */
    for (int msec = 1; msec <= win_size; ++msec) {
      int k = ora_lay_step_lif(o, idx, inten, A0, train, 0, spiked);
      if (msec >= o->L - 1) /* :71 */
        for (int q = 0; q < k; ++q) {
          int n = spiked[q];
          if (n >= o->off[o->L - 1] && n < o->off[o->L]) count_tot[n - o->off[o->L - 1]] += 1;
        }
    }
    const int n_left = Lout / 3; /* :77-79: left, middle = div(n,3), right = the rest */
    double s3[3] = {0, 0, 0};
    for (int q = 0; q < Lout; ++q) s3[q < n_left ? 0 : (q < 2 * n_left ? 1 : 2)] += count_tot[q];
    int action = 0; /* argmax([sl,sm,sr]) - 1: the first maximum :37 */
    if (s3[1] > s3[action]) action = 1;
    if (s3[2] > s3[action]) action = 2;
    double fac_reward = 1; /* :83-93: every tied pair draws, the last draw wins */
    const double difs[3] = {s3[0] - s3[2], s3[0] - s3[1], s3[1] - s3[2]};
    const int pools[3][2] = {{0, 2}, {0, 1}, {1, 2}};
    for (int i = 0; i < 3; ++i)
      if (difs[i] == 0.0) {
        action = pools[i][n_rand < n_picks ? tie_picks[n_rand] : 0];
        fac_reward = 2;
        n_rand++;
      }
    n_tot++;
    int done = ora_mountcar_step(st, action);
    tot_reward += -1.0 * fac_reward; /* :98 */
    if (done) break;
  }
  if (n_dec_out) *n_dec_out = n_tot;
  if (n_rand_out) *n_rand_out = n_rand;
  if (state_out) { state_out[0] = st[0]; state_out[1] = st[1]; }
  return tot_reward;
}

/* exp(x) for |x| < 700: fdlibm e_exp.c (k = round(x / ln2), r = x - k*ln2 in two pieces, degree-5 rational kernel),
 * written out operation by operation — the device encoders (rl_stdp_b200/csrc/gridcode.cu::exp_fdlibm) evaluate the
 * same IEEE operations, so gauss() of StateSpikeGaussGrid.jl:9 compares bit for bit between checker and device. */
double ora_exp(double x) {
  const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10, invln2 = 1.44269504088896338700e+00,
               P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
               P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
  const double ax = x < 0 ? -x : x;
  double hi = 0.0, lo = 0.0;
  int k = 0;
  if (ax > 0.34657359027997264) {
    if (ax < 1.0397207708399179) {
      if (x > 0) { hi = x - ln2HI; lo = ln2LO; k = 1; } else { hi = x + ln2HI; lo = -ln2LO; k = -1; }
    } else {
      k = (int)(invln2 * x + (x < 0 ? -0.5 : 0.5));
      const double t = (double)k;
      hi = x - t * ln2HI;
      lo = t * ln2LO;
    }
    x = hi - lo;
  } else if (ax < 3.725290298461914e-09) {
    return 1.0 + x;
  }
  const double t = x * x;
  const double c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
  const double y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
  union { double d; long long i; } u;
  u.d = y;
  u.i += (long long)k << 52;
  return u.d;
}
