/*
 * snn_oracle.c — CPU restatement (fp64, single thread) of RL-STDP's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or executed
 * from the product path (libsnn_b200.so / rl_stdp_b200/).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it — as the checker.
 *
 * PARITY UNPINNED BY THE REFERENCE: the reference (Julia) ships no tests, golden vectors or
 * fixtures for this path, and Julia is not installed here, so this restatement cannot be pinned
 * against reference outputs.  It is pinned against an INDEPENDENT numpy restatement written
 * from the same cited lines (oracle/np_restate.py, fixtures under tests/golden/).  (The LIF actor
 * of lif_oracle.c is the exception: 70 saved results of the reference pin it, see there.)
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (Julia/LLVM never contracts a*b+c without
 * @fastmath/muladd, neither is used by the reference), see oracle/Makefile.
 *
 * Three parts:
 *   1. ora_dense_*   variant A, literal: Julia/Izhikevic_dastdp/Old_net/new_net.jl:80-161
 *   2. ora_sparse_*  the sparse + axonal-delay generalisation that the CUDA kernels implement;
 *                    reduces bit-exactly to (1) when every delay is 1 (tests check this)
 *   3. ora_lif_*     variant F, layered LIF actor: Julia/Modules/Networks/net_arch.jl:142-272
 *   + Philox4x32-10 noise shared with the device build, and a restated CartPole-v1.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* threads the OpenMP build really uses (torchrun exports OMP_NUM_THREADS=1 to its workers, so the
 * count has to come from the runtime, not from the host's core count) */
int ora_omp_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void ora_omp_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------ Philox4x32-10 */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
  uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  uint32_t n0 = hi1 ^ c[1] ^ k[0];
  uint32_t n2 = hi0 ^ c[3] ^ k[1];
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
void ora_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                    uint32_t out[4]) {
  uint32_t c[4] = {c0, c1, c2, c3};
  uint32_t k[2] = {k0, k1};
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  memcpy(out, c, sizeof(c));
}
/* noise of neuron j at absolute step t: amp*(U-0.5), U=(x+0.5)/2^32, x = word (j&3) of
 * Philox(counter=(j>>2, 0, t_lo, t_hi), key=seed).  The spec of `13 .* (rand(N) .- 0.5)`
 * (new_net.jl:111) with the RNG replaced by a counter-based one both sides can replay. */
double ora_philox_noise(uint64_t seed, int64_t t, int64_t j, double amp) {
  uint32_t o[4];
  ora_philox4x32((uint32_t)(j >> 2), 0u, (uint32_t)t, (uint32_t)((uint64_t)t >> 32),
                 (uint32_t)seed, (uint32_t)(seed >> 32), o);
  double U = ((double)o[j & 3] + 0.5) * (1.0 / 4294967296.0);
  return amp * (U - 0.5);
}
void ora_philox_noise_fill(uint64_t seed, int64_t t, int64_t n, double amp, double* out) {
  for (int64_t j = 0; j < n; ++j) out[j] = ora_philox_noise(seed, t, j, amp);
}

static inline double clampd(double x, double lo, double hi) {
  /* Julia Base.clamp: ifelse(x > hi, hi, ifelse(x < lo, lo, x)) */
  return x > hi ? hi : (x < lo ? lo : x);
}

/* ------------------------------------------------------------------ 1. dense variant A */
typedef struct {
  int64_t N, Ne;
  double *v, *u, *d, *a, *STDP, *I;
  double *s, *sd;      /* column-major N x N, index [pre + N*post]  (s[pre,post]) */
  int64_t* connection; /* same layout */
  double da;
} ora_dense;

/* Network(n_neurons, connectivity, n_exc) new_net.jl:80-98, with `connection` supplied by the
 * caller (the reference draws rand(N,N) .< p, line 30: an input we record and replay). */
ora_dense* ora_dense_create(int64_t N, int64_t Ne, const int64_t* connection_colmajor) {
  ora_dense* o = (ora_dense*)calloc(1, sizeof(*o));
  o->N = N; o->Ne = Ne;
  o->v = (double*)malloc(N * sizeof(double));
  o->u = (double*)malloc(N * sizeof(double));
  o->d = (double*)malloc(N * sizeof(double));
  o->a = (double*)malloc(N * sizeof(double));
  o->STDP = (double*)calloc(N, sizeof(double));
  o->I = (double*)calloc(N, sizeof(double));
  o->s = (double*)malloc(N * N * sizeof(double));
  o->sd = (double*)calloc(N * N, sizeof(double));
  o->connection = (int64_t*)malloc(N * N * sizeof(int64_t));
  memcpy(o->connection, connection_colmajor, N * N * sizeof(int64_t));
  for (int64_t i = 0; i < N; ++i) {
    o->v[i] = -65.0 * 1.0;          /* :81 */
    o->u[i] = 0.2 * o->v[i];        /* :82 */
    o->d[i] = (i < Ne) ? 8.0 : 2.0; /* :83 */
    o->a[i] = (i < Ne) ? 0.02 : 0.1; /* :84 */
  }
  for (int64_t j = 0; j < N; ++j)
    for (int64_t i = 0; i < N; ++i) {
      double w = (i < Ne) ? 1.0 * 1.0 : -1.0 * 1.0; /* :89-90 */
      o->s[i + N * j] = w * (double)o->connection[i + N * j]; /* :91 */
    }
  o->da = 0.0;
  return o;
}
void ora_dense_destroy(ora_dense* o) {
  if (!o) return;
  free(o->v); free(o->u); free(o->d); free(o->a); free(o->STDP); free(o->I);
  free(o->s); free(o->sd); free(o->connection); free(o);
}
double* ora_dense_field(ora_dense* o, int which) {
  switch (which) {
    case 0: return o->v; case 1: return o->u; case 2: return o->STDP; case 4: return o->s;
    case 5: return o->sd; case 6: return &o->da; case 7: return o->I; default: return 0;
  }
}

/* input!(net, input_spikes) new_net.jl:104-106 */
void ora_dense_input(ora_dense* o, const int32_t* idx, int64_t n) {
  for (int64_t k = 0; k < n; ++k) o->v[idx[k]] = 40.0;
}

/* spike!(net) new_net.jl:109-133.  noise = the recorded 13 .* (rand(N) .- 0.5) of line 111.
 * Returns the number of spiking neurons; ids (0-based, ascending) in spiked[]. */
int64_t ora_dense_spike(ora_dense* o, const double* noise, int32_t* spiked) {
  const int64_t N = o->N;
  int64_t k = 0;
  for (int64_t j = 0; j < N; ++j) o->I[j] = noise ? noise[j] : 0.0; /* :111 */
  for (int64_t j = 0; j < N; ++j)
    if (o->v[j] > 30) spiked[k++] = (int32_t)j; /* :112 findall(x->x>30) */
  for (int64_t q = 0; q < k; ++q) {
    int64_t n = spiked[q];
    o->v[n] = -65.0;       /* :113 */
    o->u[n] += o->d[n];    /* :114 */
  }
  for (int64_t q = 0; q < k; ++q) { /* :116-118  I .+= s[neuron,:] */
    int64_t n = spiked[q];
    for (int64_t j = 0; j < N; ++j) o->I[j] += o->s[n + N * j];
  }
  for (int rep = 0; rep < 2; ++rep) /* :120-121 */
    for (int64_t j = 0; j < N; ++j) {
      double v = o->v[j];
      o->v[j] = v + 0.5 * ((0.04 * v + 5) * v + 140 - o->u[j] + o->I[j]);
    }
  for (int64_t j = 0; j < N; ++j) /* :122 */
    o->u[j] = o->u[j] + o->a[j] * (0.2 * o->v[j] - o->u[j]);
  for (int64_t q = 0; q < k; ++q) o->STDP[spiked[q]] = 0.1; /* :123 */
  for (int64_t q = 0; q < k; ++q) { /* :125-128 */
    int64_t n = spiked[q];
    for (int64_t i = 0; i < N; ++i) o->sd[i + N * n] += 1.0 * o->STDP[i]; /* sd[:,n] */
    for (int64_t j = 0; j < N; ++j) o->sd[n + N * j] -= 1.5 * o->STDP[j]; /* sd[n,:] */
  }
  for (int64_t j = 0; j < N; ++j) o->STDP[j] *= 0.95; /* :130 */
  o->da *= 0.995;                                     /* :131 */
  return k;
}

/* learn!(net) new_net.jl:135-142 */
void ora_dense_learn(ora_dense* o) {
  const int64_t N = o->N, Ne = o->Ne;
  const double g = 0.002 + o->da;
  for (int64_t j = 0; j < N; ++j)
    for (int64_t i = 0; i < Ne; ++i) {
      double x = o->s[i + N * j] + g * o->sd[i + N * j]; /* :137 */
      o->s[i + N * j] = clampd(x, 0, 4);                 /* :138 */
    }
  for (int64_t e = 0; e < N * N; ++e) o->s[e] *= (double)o->connection[e]; /* :139 */
  for (int64_t e = 0; e < N * N; ++e) o->sd[e] *= 0.99;                    /* :141 */
}

/* step!(net[, input_spikes], train) new_net.jl:146-161 */
int64_t ora_dense_step(ora_dense* o, const double* noise, const int32_t* force, int64_t n_force,
                       int train, int32_t* spiked) {
  if (n_force > 0) ora_dense_input(o, force, n_force);
  int64_t k = ora_dense_spike(o, noise, spiked);
  if (train) ora_dense_learn(o);
  return k;
}

/* ------------------------------------------------------------------ 2. sparse + delays */
typedef struct {
  int64_t N, Nper, Ne, E;
  int32_t Dmax, R;
  /* parameters (mirror of snn_config) */
  int32_t threshold_ge, sd_decay_mode, rate_mode, plastic_ei;
  int32_t ho_inh;    /* new_net_col_sep.jl:149-150: inhibitory spikers raise their ho too */
  int32_t variant_g; /* Old_net/DASTDP.jl ordering: LTP reads the pre trace of the previous ms
                        (:62-68 STDP[pre,msec]), every LTP precedes every LTD (Julia_DA_STDPv2.jl:78-80),
                        I accumulates in descending presynaptic id (:70-80 walks firings backwards) */
  double vthresh, v_reset, a_exc, a_inh, d_exc, d_inh, trace_set, trace_decay, apost, apre,
      da_decay, learn_base, eta, wmin, wmax, sd_decay, theta, ttheta;
  int64_t ho_from;
  /* synapses: caller's CSR order */
  int64_t* rowptr; int32_t* post; int32_t* pre; uint8_t* delay; double *w, *sd;
  uint8_t* plastic;
  /* incoming lists in canonical order: ascending pre, ties by (delay, caller order) */
  int64_t* colptr; int64_t* in_syn;
  /* state */
  double *v, *u, *ho, *I;
  double* trace;   /* R x N: trace[slot][j] = value after the set of that step (pre-decay) */
  uint8_t* spk;    /* R x N spike flags per slot */
  double* da;      /* per instance */
  int64_t t;       /* absolute step counter */
  int64_t n_syn_events, n_ltp_events;
  /* event-driven learning of the production build (snn_config.reserved[4] = 16 + M),
   * restated so that that path has a checker WITH the clamp: w holds s0, sd holds sd' = sd / r^k,
   * cacc the accumulator C; per instance H, r^k, 1/r^k and the learn steps since the last re-basing */
  int32_t lazy; double lazy_da;
  double *cacc, *lzH, *lzR, *lzInv; int32_t* lzK;
} ora_sparse;

typedef struct { int64_t key; int64_t e; } ora_sortrec;
static int ora_cmp(const void* a, const void* b) {
  const ora_sortrec* x = (const ora_sortrec*)a; const ora_sortrec* y = (const ora_sortrec*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->e < y->e ? -1 : (x->e > y->e ? 1 : 0);
}

/* params: array of doubles in the order documented in oracle/py_oracle.py (PARAM_ORDER) */
ora_sparse* ora_sparse_create(int64_t N, int64_t Nper, int64_t Ne, int32_t Dmax,
                              const int64_t* rowptr, const int32_t* post, const double* w,
                              const uint8_t* delay, const double* p) {
  ora_sparse* o = (ora_sparse*)calloc(1, sizeof(*o));
  o->N = N; o->Nper = Nper; o->Ne = Ne; o->Dmax = Dmax; o->R = Dmax + 1;
  const int64_t E = rowptr[N];
  o->E = E;
  o->threshold_ge = (int32_t)p[0]; o->vthresh = p[1]; o->v_reset = p[2];
  o->a_exc = p[3]; o->a_inh = p[4]; o->d_exc = p[5]; o->d_inh = p[6];
  o->trace_set = p[7]; o->trace_decay = p[8]; o->apost = p[9]; o->apre = p[10];
  o->da_decay = p[11]; o->learn_base = p[12]; o->eta = p[13]; o->wmin = p[14]; o->wmax = p[15];
  o->sd_decay = p[16]; o->sd_decay_mode = (int32_t)p[17]; o->rate_mode = (int32_t)p[18];
  o->plastic_ei = (int32_t)p[19]; o->theta = p[20]; o->ttheta = p[21]; o->ho_from = (int64_t)p[22];
  o->ho_inh = (int32_t)p[23]; o->variant_g = (int32_t)p[24];
  o->lazy = (int32_t)p[25]; o->lazy_da = p[26] > 0 ? p[26] : 1e30;
  o->rowptr = (int64_t*)malloc((N + 1) * sizeof(int64_t));
  memcpy(o->rowptr, rowptr, (N + 1) * sizeof(int64_t));
  o->post = (int32_t*)malloc((E + 1) * sizeof(int32_t));
  o->pre = (int32_t*)malloc((E + 1) * sizeof(int32_t));
  o->delay = (uint8_t*)malloc(E + 1);
  o->w = (double*)malloc((E + 1) * sizeof(double));
  o->sd = (double*)calloc(E + 1, sizeof(double));
  o->plastic = (uint8_t*)malloc(E + 1);
  memcpy(o->post, post, E * sizeof(int32_t));
  memcpy(o->w, w, E * sizeof(double));
  for (int64_t i = 0; i < N; ++i)
    for (int64_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
      o->pre[e] = (int32_t)i;
      o->delay[e] = delay ? delay[e] : 1;
      int pre_exc = (i % Nper) < Ne;
      int post_exc = (post[e] % Nper) < Ne;
      o->plastic[e] = (uint8_t)(pre_exc && (post_exc || o->plastic_ei));
    }
  /* incoming lists */
  ora_sortrec* rec = (ora_sortrec*)malloc((E + 1) * sizeof(ora_sortrec));
  for (int64_t e = 0; e < E; ++e) {
    /* key: post, pre, delay ; tie -> caller order e */
    rec[e].key = (((int64_t)o->post[e] * (int64_t)N) + o->pre[e]) * 256 + o->delay[e];
    rec[e].e = e;
  }
  qsort(rec, (size_t)E, sizeof(ora_sortrec), ora_cmp);
  o->colptr = (int64_t*)calloc(N + 1, sizeof(int64_t));
  o->in_syn = (int64_t*)malloc((E + 1) * sizeof(int64_t));
  for (int64_t k = 0; k < E; ++k) { o->in_syn[k] = rec[k].e; o->colptr[o->post[rec[k].e] + 1]++; }
  for (int64_t j = 0; j < N; ++j) o->colptr[j + 1] += o->colptr[j];
  free(rec);
  o->v = (double*)malloc(N * sizeof(double));
  o->u = (double*)malloc(N * sizeof(double));
  o->ho = (double*)calloc(N, sizeof(double));
  o->I = (double*)calloc(N, sizeof(double));
  o->trace = (double*)calloc((size_t)o->R * N, sizeof(double));
  o->spk = (uint8_t*)calloc((size_t)o->R * N, 1);
  o->da = (double*)calloc(N / Nper, sizeof(double));
  for (int64_t j = 0; j < N; ++j) { o->v[j] = o->v_reset; o->u[j] = 0.2 * o->v[j]; }
  o->t = 0;
  if (o->lazy) {
    const int64_t B = N / Nper;
    o->cacc = (double*)calloc(E + 1, sizeof(double));
    o->lzH = (double*)calloc(B, sizeof(double));
    o->lzR = (double*)malloc(B * sizeof(double));
    o->lzInv = (double*)malloc(B * sizeof(double));
    o->lzK = (int32_t*)calloc(B, sizeof(int32_t));
    for (int64_t b = 0; b < B; ++b) o->lzR[b] = o->lzInv[b] = 1.0;
  }
  return o;
}
/* weight of synapse e as the event-driven build delivers it */
static double ora_lazy_w(const ora_sparse* o, int64_t e) {
  if (!o->lazy || !o->plastic[e]) return o->w[e];
  return clampd((o->w[e] + o->lzH[o->pre[e] / o->Nper] * o->sd[e]) - o->cacc[e], o->wmin, o->wmax);
}
/* re-basing pass of instance b with the advanced (H, r^k): s0 = clamp(s0 + H*sd' - C), sd' *= r^k, C = 0 */
static void ora_lazy_rebase(ora_sparse* o, int64_t b, double H, double Rk) {
  for (int64_t e = 0; e < o->E; ++e) {
    if (!o->plastic[e] || o->pre[e] / o->Nper != b) continue;
    o->w[e] = clampd((o->w[e] + H * o->sd[e]) - o->cacc[e], o->wmin, o->wmax);
    o->sd[e] = o->sd[e] * Rk;
    o->cacc[e] = 0.0;
  }
  o->lzH[b] = 0.0; o->lzR[b] = 1.0; o->lzInv[b] = 1.0; o->lzK[b] = 0;
}
/* what snn_get_array / snn_set_array do first in that mode: materialise (w, sd) */
void ora_sparse_lazy_flush(ora_sparse* o) {
  if (!o->lazy) return;
  for (int64_t b = 0; b < o->N / o->Nper; ++b) ora_lazy_rebase(o, b, o->lzH[b], o->lzR[b]);
}
void ora_sparse_destroy(ora_sparse* o) {
  if (!o) return;
  free(o->rowptr); free(o->post); free(o->pre); free(o->delay); free(o->w); free(o->sd);
  free(o->plastic); free(o->colptr); free(o->in_syn); free(o->v); free(o->u); free(o->ho);
  free(o->I); free(o->trace); free(o->spk); free(o->da);
  free(o->cacc); free(o->lzH); free(o->lzR); free(o->lzInv); free(o->lzK); free(o);
}
/* field ids follow include/snn_b200.h SNN_FIELD_*; trace returns the post-decay value the
 * reference holds between calls (new_net.jl:130). */
void ora_sparse_get(ora_sparse* o, int which, double* out) {
  const int64_t N = o->N;
  if (which == 4 || which == 5) ora_sparse_lazy_flush(o);
  switch (which) {
    case 0: memcpy(out, o->v, N * sizeof(double)); break;
    case 1: memcpy(out, o->u, N * sizeof(double)); break;
    case 2: {
      const double* tr = o->trace + (size_t)((o->t + o->R - 1) % o->R) * N;
      for (int64_t j = 0; j < N; ++j) out[j] = o->t == 0 ? 0.0 : tr[j] * o->trace_decay;
      break;
    }
    case 3: memcpy(out, o->ho, N * sizeof(double)); break;
    case 4: memcpy(out, o->w, o->E * sizeof(double)); break;
    case 5: memcpy(out, o->sd, o->E * sizeof(double)); break;
    case 6: memcpy(out, o->da, (N / o->Nper) * sizeof(double)); break;
    case 7: memcpy(out, o->I, N * sizeof(double)); break;
  }
}
void ora_sparse_set(ora_sparse* o, int which, const double* in) {
  const int64_t N = o->N;
  if (which == 4 || which == 5) ora_sparse_lazy_flush(o);
  switch (which) {
    case 0: memcpy(o->v, in, N * sizeof(double)); break;
    case 1: memcpy(o->u, in, N * sizeof(double)); break;
    case 3: memcpy(o->ho, in, N * sizeof(double)); break;
    case 4: memcpy(o->w, in, o->E * sizeof(double)); break;
    case 5: memcpy(o->sd, in, o->E * sizeof(double)); break;
    case 6: memcpy(o->da, in, (N / o->Nper) * sizeof(double)); break;
  }
}
void ora_sparse_add_da(ora_sparse* o, int64_t inst, double x) { o->da[inst] += x; }
void ora_sparse_counters(ora_sparse* o, int64_t* out) { out[0] = o->n_syn_events; out[1] = o->n_ltp_events; }

/* One call of step!(net[, input_spikes], train), generalised to CSR synapses with axonal delays
 * (SURVEY §8a "delay semantics"):
 *   - a spike of pre i detected at call t0 arrives through a synapse of delay d at call
 *     t0+d-1: there it adds w to I (consumed by the same call's Euler update, as
 *     new_net.jl:116-121 does for d=1) and takes LTD with the post trace of that call;
 *   - a spike of post j at call t gives LTP with the pre trace as seen at the synapse, i.e. the
 *     trace i had at call t-(d-1) (Old_net/DASTDP.jl:62-68 time-indexed STDP[N,1001+D]);
 *   - per synapse the new_net.jl:125-128 loop order is kept: post<=pre -> LTP then LTD,
 *     otherwise LTD then LTP;
 *   - I[j] accumulates in ascending presynaptic id starting from the noise value (:111,:116-118).
 */
int64_t ora_sparse_step(ora_sparse* o, const double* noise, const int32_t* force, int64_t n_force,
                        int train, int32_t* spiked) {
  const int64_t N = o->N, Nper = o->Nper, Ne = o->Ne;
  const int32_t R = o->R;
  const int64_t t = o->t;
  const int32_t slot = (int32_t)(t % R), prev = (int32_t)((t + R - 1) % R);
  double* tr = o->trace + (size_t)slot * N;
  const double* trp = o->trace + (size_t)prev * N;
  uint8_t* sp = o->spk + (size_t)slot * N;
  int64_t k = 0;
  for (int64_t q = 0; q < n_force; ++q) o->v[force[q]] = 40.0; /* input! new_net.jl:104-106 */
  /* the OpenMP build (libsnn_oracle_omp.so, bench.py --impl reference) splits the loops whose
   * iterations are independent; every per-neuron / per-synapse expression is evaluated by one
   * thread in the same order, so its results are bit-identical to the serial build */
  #pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < N; ++j) {
    o->I[j] = noise ? noise[j] : 0.0;
    double x = o->v[j] - o->ho[j];
    int s = o->threshold_ge ? (x >= o->vthresh) : (x > o->vthresh);
    sp[j] = (uint8_t)s;
  }
  for (int64_t j = 0; j < N; ++j) if (sp[j]) spiked[k++] = (int32_t)j;
  for (int64_t q = 0; q < k; ++q) {
    int64_t n = spiked[q];
    int64_t l = n % Nper;
    if ((l < Ne || o->ho_inh) && l >= o->ho_from) o->ho[n] += o->theta; /* net_res.jl:170-171 */
    o->v[n] = o->v_reset;
    o->u[n] += (l < Ne) ? o->d_exc : o->d_inh;
  }
  /* traces: set for this call's spikers, decayed copy of the previous call otherwise */
  #pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < N; ++j) tr[j] = sp[j] ? o->trace_set : (t == 0 ? 0.0 : trp[j] * o->trace_decay);
  /* arrivals, pull order */
  int64_t n_ev = 0, n_ltp = 0;
  #pragma omp parallel for schedule(dynamic, 256) reduction(+ : n_ev)
  for (int64_t j = 0; j < N; ++j) {
    double acc = o->I[j];
    const int64_t c0 = o->colptr[j], c1 = o->colptr[j + 1];
    for (int64_t qq = c0; qq < c1; ++qq) {
      int64_t q = o->variant_g ? (c1 - 1 - (qq - c0)) : qq; /* DASTDP.jl:72 last_ walks backwards */
      int64_t e = o->in_syn[q];
      int64_t ts = t - (o->delay[e] - 1);
      if (ts < 0) continue;
      if (o->spk[(size_t)(ts % R) * N + o->pre[e]]) { acc += ora_lazy_w(o, e); n_ev++; }
    }
    o->I[j] = acc;
  }
  o->n_syn_events += n_ev;
  #pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < N; ++j) {
    double v = o->v[j], u = o->u[j], I = o->I[j];
    v = v + 0.5 * ((0.04 * v + 5) * v + 140 - u + I);
    v = v + 0.5 * ((0.04 * v + 5) * v + 140 - u + I);
    double a = ((j % Nper) < Ne) ? o->a_exc : o->a_inh;
    u = u + a * (0.2 * v - u);
    o->v[j] = v; o->u[j] = u;
  }
  /* STDP on plastic synapses */
  #pragma omp parallel for schedule(static) reduction(+ : n_ltp)
  for (int64_t e = 0; e < o->E; ++e) {
    if (!o->plastic[e]) continue;
    int64_t i = o->pre[e], j = o->post[e];
    int64_t ts = t - (o->delay[e] - 1);
    int ltd = ts >= 0 && o->spk[(size_t)(ts % R) * N + i];
    int ltp = sp[j];
    if (ltp) n_ltp++;
    if (!ltd && !ltp) continue;
    int64_t tsp = o->variant_g ? ts - 1 : ts; /* DASTDP.jl:65 STDP[pre_neurons,msec]: before this ms's spikes */
    double pre_tr = tsp >= 0 ? o->trace[(size_t)(tsp % R) * N + i] : 0.0;
    double up = o->apost * pre_tr;
    double dn = o->apre * tr[j];
    if (o->lazy) { /* the two additive accumulators; LTP and LTD are separate (unordered) reductions there */
      const int64_t b = i / Nper;
      if (ltp) { double d = up * o->lzInv[b]; o->sd[e] += d; o->cacc[e] += d * o->lzH[b]; }
      if (ltd) { double d = (-dn) * o->lzInv[b]; o->sd[e] += d; o->cacc[e] += d * o->lzH[b]; }
      continue;
    }
    double x = o->sd[e];
    if (ltd && ltp) x = (o->variant_g || j <= i) ? (x + up) - dn : (x - dn) + up;
    else if (ltp) x = x + up;
    else x = x - dn;
    o->sd[e] = x;
  }
  o->n_ltp_events += n_ltp;
  for (int64_t b = 0; b < N / Nper; ++b) o->da[b] *= o->da_decay;
  for (int64_t j = 0; j < N; ++j) o->ho[j] *= o->ttheta; /* net_res.jl:204 */
  if (o->sd_decay_mode == 1 && !o->lazy) {
    #pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < o->E; ++e) if (o->plastic[e]) o->sd[e] *= o->sd_decay;
  }
  if (o->lazy && (train || o->sd_decay_mode == 1)) {
    /* lazy_learn_kernel / lazy_commit_kernel: advance (H, r^k), re-base every M-th learn step or while da is high */
    for (int64_t b = 0; b < N / Nper; ++b) {
      double da = o->da[b];
      double g = o->rate_mode == 0 ? (o->learn_base + da) : (o->rate_mode == 1 ? da : o->eta * da);
      double Hn = train ? o->lzH[b] + g * o->lzR[b] : o->lzH[b];
      double Rn = (o->sd_decay_mode == 1 || (train && o->sd_decay_mode == 0)) ? o->lzR[b] * o->sd_decay : o->lzR[b];
      int Kn = o->lzK[b] + (train ? 1 : 0);
      if (Kn >= o->lazy || (train && da > o->lazy_da)) ora_lazy_rebase(o, b, Hn, Rn);
      else { o->lzH[b] = Hn; o->lzR[b] = Rn; o->lzInv[b] = 1.0 / Rn; o->lzK[b] = Kn; }
    }
  } else if (train) {
    #pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < o->E; ++e) {
      if (!o->plastic[e]) continue;
      double da = o->da[o->pre[e] / Nper];
      double g = o->rate_mode == 0 ? (o->learn_base + da) : (o->rate_mode == 1 ? da : o->eta * da);
      o->w[e] = clampd(o->w[e] + g * o->sd[e], o->wmin, o->wmax);
      if (o->sd_decay_mode == 0) o->sd[e] *= o->sd_decay;
    }
  }
  o->t = t + 1;
  return k;
}
