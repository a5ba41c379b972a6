/*
 * snn_b200.h — C ABI of libsnn_b200.so: the B200-native (sm_100a) replacement for the
 * per-millisecond spiking-neuron step + dopamine-modulated STDP update of SuReLI/RL-STDP.
 *
 * The reference has no FFI layer: its boundary is the Julia-level `Network` API
 * (constructor, step!/step_LIF!/learn!, public mutable fields).  Every entry point below
 * cites the reference interface it replaces (paths relative to the reference repo root).
 * A Julia `ccall` wrapper (julia/RLSTDPB200.jl) and a Python `ctypes` mirror
 * (rl_stdp_b200/) bind exactly these symbols; see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns int32 status: 0 = OK, <0 = error; text via snn_last_error()
 *     (thread-local).  No exception or abort crosses the boundary.
 *   - neuron / synapse ids are 0-based here (the Julia wrapper adds 1).
 *   - host buffers are caller-owned and only touched during the call.
 *   - there is NO CPU backend in this library: without a CUDA device snn_create fails.
 */
#ifndef SNN_B200_H
#define SNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNN_ABI_VERSION 3

/* status codes */
enum {
  SNN_OK = 0,
  SNN_ERR_INVALID = -1,   /* bad argument / config */
  SNN_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
  SNN_ERR_NODEVICE = -3,  /* no usable CUDA device: the library never falls back to CPU */
  SNN_ERR_STATE = -4,     /* call order violated (e.g. step before synapses were set) */
  SNN_ERR_CAPACITY = -5   /* an output buffer was too small (results truncated) */
};

enum { SNN_F32 = 0, SNN_F64 = 1 };

/* propagation mode */
enum {
  SNN_MODE_FAST = 0,  /* push: atomic scatter-add of weights into the current array */
  SNN_MODE_EXACT = 1  /* pull: every target sums its spiking sources in ascending presynaptic
                         id, the accumulation order of new_net.jl:116-118 — bit-reproducible */
};

enum { SNN_NOISE_NONE = 0, SNN_NOISE_REPLAY = 1, SNN_NOISE_PHILOX = 2 };

/* when the eligibility trace sd is decayed (SURVEY §8a switch list) */
enum {
  SNN_SD_DECAY_ON_LEARN = 0, /* new_net.jl:141  sd .*= 0.99 inside learn!            */
  SNN_SD_DECAY_EVERY_STEP = 1, /* net_arch.jl:207-209 de[l] .*= tde every ms (LIF)   */
  SNN_SD_DECAY_NEVER = 2     /* net_arch.jl:328-379 Izhikevich layered path          */
};

/* learn-rate form: w += rate * sd */
enum {
  SNN_RATE_BASE_PLUS_DA = 0, /* (learn_base + da)   new_net.jl:137, net_res.jl:213 */
  SNN_RATE_DA = 1,           /* da                  net_arch.jl:217                */
  SNN_RATE_ETA_DA = 2        /* eta*da              net_arch.jl:226                */
};

/* Plain-old-data configuration of one (possibly batched) sparse Izhikevich DA-STDP network.
 * Replaces: `Network(n_neurons, connectivity, n_exc)` Old_net/new_net.jl:80-98 and the
 * hard-coded constants of spike!/learn! (new_net.jl:109-142); the YAML `param` Dict of
 * net_res.jl / net_arch.jl (Modules/cfg.yml) maps onto the same fields. */
typedef struct snn_config {
  int32_t struct_size;     /* = sizeof(snn_config), checked */
  int32_t dtype;           /* SNN_F32 | SNN_F64 : arithmetic + storage type of state */
  int32_t mode;            /* SNN_MODE_FAST | SNN_MODE_EXACT */
  int32_t device;          /* CUDA device ordinal */
  int64_t n_neurons;       /* total neurons = n_instances * n_per_instance */
  int64_t n_per_instance;  /* neurons of one independent network instance */
  int64_t n_exc;           /* per instance: local ids [0,n_exc) excitatory, rest inhibitory */
  int32_t max_delay;       /* axonal delays are 1..max_delay steps (1 = reference semantics) */
  int32_t threshold_ge;    /* 0: spike iff v-ho >  vthresh (new_net.jl:112)
                              1: spike iff v-ho >= vthresh (NeuronModel.jl / Julia_DA_STDPv2.jl:75) */
  double vthresh;          /* 30   new_net.jl:112 */
  double v_reset;          /* -65  new_net.jl:113 (param "c") */
  double a_exc, a_inh;     /* 0.02 / 0.1  new_net.jl:84 */
  double d_exc, d_inh;     /* 8 / 2       new_net.jl:83 */
  double trace_set;        /* 0.1   new_net.jl:123 */
  double trace_decay;      /* 0.95  new_net.jl:130 */
  double apost;            /* 1.0   LTP amplitude new_net.jl:126 */
  double apre;             /* 1.5   LTD amplitude new_net.jl:127 */
  double da_decay;         /* 0.995 new_net.jl:131 */
  double learn_base;       /* 0.002 new_net.jl:137 */
  double eta;              /* only for SNN_RATE_ETA_DA */
  double wmin, wmax;       /* clamp bounds 0 / 4  new_net.jl:138 */
  double sd_decay;         /* 0.99  new_net.jl:141 */
  int32_t sd_decay_mode;   /* SNN_SD_DECAY_* */
  int32_t rate_mode;       /* SNN_RATE_* */
  int32_t noise_mode;      /* SNN_NOISE_* ; REPLAY = caller passes the recorded 13*(rand-0.5) */
  int32_t plastic_ei;      /* 1: exc->inh synapses are plastic too (variant A / net_res arch==[0]);
                              0: only exc->exc (net_res.jl layered) */
  double noise_amp;        /* 13.0 (PHILOX mode: I = noise_amp*(U-0.5))  new_net.jl:111 */
  uint64_t noise_seed;     /* PHILOX key */
  double theta;            /* homeostasis increment (0 = off)  net_res.jl:171 */
  double ttheta;           /* homeostasis decay                net_res.jl:204 */
  int64_t ho_from;         /* local ids >= ho_from (and < n_exc) get homeostasis (arch[1]) */
  int32_t ho_inh;          /* 1: inhibitory spikers raise their ho too (Old_net/new_net_col_sep.jl:149-150
                              filters on `x > arch[1]` only); 0: excitatory only (net_res.jl:169) */
  int32_t variant_g;       /* 1: the ordering of Old_net/DASTDP.jl + Julia_DA_STDPv2.jl:71-96 — LTP reads the
                              pre trace of the PREVIOUS ms (STDP[pre,msec], :62-68: a presynaptic spike of
                              the same ms is not seen), all LTP before all LTD, the input current
                              accumulates in descending presynaptic id (:70-80 walks `firings` backwards) */
  int32_t reserved[8];     /* must be zero.  (ABI 2 carried tuning switches here as magic numbers; since ABI 3
                              everything that is not part of the model is set BY NAME with snn_set_param.) */
} snn_config;

typedef struct snn_net snn_net;

/* dopamine injection event: after step `step` (0-based inside the window) of instance
 * `instance`, da += amount.  Replaces `n.da += 0.5` newnet_DASTDP.jl:54-56 (which runs after
 * step! returns) and Latencysim_rewardv0.jl:141-145. */
typedef struct snn_da_event {
  int32_t step;
  int32_t instance;
  double amount;
} snn_da_event;

/* forced-spike input: before step `step`, v[neuron] = 40.0.
 * Replaces input!(net, ::Array{Int64}) new_net.jl:104-106. */
typedef struct snn_force_event {
  int32_t step;
  int32_t neuron;
} snn_force_event;

/* injected-current input: before step `step`, neuron `neuron` integrates two Euler half-steps
 * with the external current `current` (= (imax-imin)*intensity+imin, formed by the caller).
 * Replaces input!(net, ::Array{Tuple{Int64,Float64}}) net_arch.jl:316-326 / net_res.jl:150-160. */
typedef struct snn_current_event {
  int32_t step;
  int32_t neuron;
  double current;
} snn_current_event;

/* fields for snn_get_array / snn_set_array */
enum {
  SNN_FIELD_V = 0,       /* [n_neurons] real  — Network.v  new_net.jl:9  */
  SNN_FIELD_U = 1,       /* [n_neurons] real  — Network.u  new_net.jl:10 */
  SNN_FIELD_TRACE = 2,   /* [n_neurons] real  — Network.STDP new_net.jl:17 (post-decay value) */
  SNN_FIELD_HO = 3,      /* [n_neurons] real  — Network.ho net_res.jl:27 */
  SNN_FIELD_W = 4,       /* [n_synapses] real, caller's CSR order — Network.s  new_net.jl:15 */
  SNN_FIELD_SD = 5,      /* [n_synapses] real, caller's CSR order — Network.sd new_net.jl:16
                            (only plastic synapses are maintained; others read 0) */
  SNN_FIELD_DA = 6,      /* [n_instances] real — Network.da new_net.jl:23 */
  SNN_FIELD_I = 7        /* [n_neurons] real — Network.I new_net.jl:22 (input current used by
                            the last Euler update) */
};

/* per-window statistics filled by snn_step */
typedef struct snn_step_stats {
  int64_t n_spikes;          /* spikes emitted in the window */
  int64_t n_syn_events;      /* spike deliveries (one spike across one synapse) */
  int64_t n_ltp_events;      /* incoming plastic synapses visited for LTP */
  double device_ms;          /* CUDA-event time of the window on the net's stream */
  double learn_ms;           /* summed CUDA-event time of the learn kernel launches (profiling on) */
  double neuron_ms;          /* same for the neuron kernel */
  double synapse_ms;         /* same for the propagate+STDP kernel */
  int32_t learn_launches, neuron_launches, synapse_launches;
  int32_t reserved;
} snn_step_stats;

const char* snn_last_error(void);
int32_t snn_abi_version(void);

/* Fill `cfg` with the constants of variant A (Old_net/new_net.jl:80-142). */
int32_t snn_config_default(snn_config* cfg);

/* Replaces the Network constructors (new_net.jl:58-98): allocates all device state,
 * v = v_reset, u = 0.2*v (new_net.jl:81-82), traces/sd/da = 0. */
int32_t snn_create(const snn_config* cfg, snn_net** out);
int32_t snn_destroy(snn_net* net);

/* Synapses in CSR by presynaptic neuron.  rowptr[n_neurons+1], post[E] (global 0-based ids, must
 * stay inside the presynaptic neuron's instance), w[E] (float64), delay[E] in 1..max_delay or
 * NULL (= all 1).  Replaces `s`/`connection` of new_net.jl:14-15,27-56 and post/delays of
 * Old_net/DASTDP.jl:32-49.  The library re-sorts each row by (delay, given order) and builds
 * the incoming-synapse index on the device. */
int32_t snn_set_synapses_csr(snn_net* net, const int64_t* rowptr, const int32_t* post,
                             const double* w, const uint8_t* delay);

/* Exactly what variant A holds: s[pre,post] and connection[pre,post], column-major N x N
 * (new_net.jl:14-15, 86-91); single instance only. */
int32_t snn_set_synapses_dense(snn_net* net, const double* s_colmajor,
                               const int64_t* connection_colmajor);

/* The transposed storage of variant B: s[post,pre] and connection[post,pre], column-major N x N, "read by
 * columns" (Old_net/new_net_col.jl:13-15, 93-99); single instance only. */
int32_t snn_set_synapses_dense_col(snn_net* net, const double* s_colmajor,
                                   const int64_t* connection_colmajor);

/* Named parameters.
 *  (1) The model's scalars, by their snn_config field name ("vthresh", "v_reset", "a_exc", "a_inh", "d_exc",
 *      "d_inh", "trace_set", "trace_decay", "apost", "apre", "da_decay", "learn_base", "eta", "wmin", "wmax",
 *      "sd_decay", "theta", "ttheta", "noise_amp", "noise_seed"): the reference keeps them as plain mutable
 *      fields / Dict entries (net.param["..."], Modules/cfg.yml) that a script may change between two step!
 *      calls; so may the caller here, between two windows.
 *  (2) Everything that is NOT the model — any time unless marked [S] = only before the synapses are set:
 *      "graph"                 1 (default) whole windows as device-side schedules; 0 plain launches per step
 *      "persistent"            how production-mode windows are scheduled: -1 (default) automatic; 0 a captured CUDA
 *                              graph of per-step kernels; 1 the persistent chain — one resident kernel per role for
 *                              the whole window, steps handed over through device-side counters.
 *                              Automatic = the chain, except whenever a
 *                              profiler (ncu, nsys, compute-sanitizer) or CUDA_LAUNCH_BLOCKING serialises kernels
 *                              (resident roles cannot hand over to kernels that are not running)
 *      "chain_lookahead"       persistent chain: steps of push lookahead, 1..2 (0 = default 1)
 *      "learn_async"       [S] 0 auto, 1 learn! always in place, 2 always the asynchronous double-buffered pass
 *      "learn_log_cap"     [S] entries of the sd-update log of the asynchronous pass (0 = by size)
 *      "learn_pass_blocks", "learn_pass_stages", "learn_pass_tile_k" [S]   asynchronous pass of the graph path
 *      "synapse_blocks_per_sm", "ltp_blocks_per_sm" [S]                    side kernels of the graph path
 *      "chain_neuron_threads", "chain_synapse_blocks", "chain_ltp_blocks", "chain_ltp_threads",
 *      "chain_learn_blocks", "chain_neuron_registers", "chain_learn_stages"   footprints of the chain's roles (sweeps)
 *      "timeline"              1: kernels stamp %globaltimer for snn_debug_timeline (diagnostic); 2: the chain's phase
 *                              classes are stamped by one block of the neuron role instead of the envelope over all
 *      "serialise_learn_pass", "neuron_blocks_256", "lsu_pass"             A/B switches (diagnostic)
 *      "learn_lazy"        [S] M > 0: EXPERIMENTAL event-driven learning (SNN_MODE_FAST only) — an
 *                              approximation of the reference's rule: the clamp wmin <= w <= wmax is applied
 *                              when a weight is read and at a dense re-basing pass every M learn steps
 *                              instead of at every learn step (M = 1 is the reference's rule; not with
 *                              snn_set_partition or synapse probes)
 *      "learn_lazy_da"     [S] with learn_lazy: re-base at every learn step while da exceeds this (0 = never)
 *      read-only: "window_path" = how the last window ran (2 persistent chain, 1 captured graph, 0 plain launches)
 * Unknown names are an error (SNN_ERR_INVALID). */
int32_t snn_set_param(snn_net* net, const char* name, double value);
int32_t snn_get_param(snn_net* net, const char* name, double* value);

/* Checkpoint / resume of a run in flight (the reference itself only saves genomes,
 * Modules/save_param.jl:4-47).  The blob holds everything the next window depends on: v, u, ho, da, (w, sd),
 * the spike history the STDP traces derive from, and the axonal-delay ring (per-step spike lists and bitmaps
 * of the last max_delay steps, the currents already pushed), plus the step counter.  It restores only into a
 * handle created with the same snn_config and given the same synapses.  A restored run continues bit for
 * bit (SNN_MODE_EXACT; in SNN_MODE_FAST up to the order of the atomic sums, as any two runs).
 *   snn_checkpoint_bytes: size the blob needs right now (it depends on the spikes in flight). */
int32_t snn_checkpoint_bytes(snn_net* net, int64_t* bytes);
int32_t snn_checkpoint(snn_net* net, void* buf, int64_t bytes);
int32_t snn_restore(snn_net* net, const void* buf, int64_t bytes);

/* Run n_steps calls of step!(net[, input_spikes], train) (new_net.jl:146-161) on the device
 * with no host round trip.
 *   noise       [n_steps][n_neurons] float64 (REPLAY mode) or NULL
 *   train_flags [n_steps] or NULL (= never learn); train_flags[k]!=0 -> learn! after step k
 *   force       forced spikes (input!), sorted by step, may be NULL
 *   da          dopamine events sorted by step, may be NULL
 *   spikes_out  ids of spiking neurons, ascending inside each step, concatenated; NULL = not
 *               recorded.  spike_offsets[n_steps+1] delimits the steps.
 * Returns SNN_ERR_CAPACITY (after finishing the window) if more than spikes_cap spikes occurred. */
int32_t snn_step(snn_net* net, int32_t n_steps, const double* noise, const uint8_t* train_flags,
                 const snn_force_event* force, int32_t n_force, const snn_da_event* da,
                 int32_t n_da, int32_t* spikes_out, int64_t* spike_offsets, int64_t spikes_cap,
                 snn_step_stats* stats);

/* Extensible form of snn_step: every input/output of a window in one struct (zero-initialise,
 * set struct_size).  Adds injected-current events to what snn_step takes. */
typedef struct snn_step_io {
  int32_t struct_size;
  int32_t n_steps;
  const double* noise;
  const uint8_t* train_flags;
  const snn_force_event* force;
  const snn_da_event* da;
  const snn_current_event* currents;  /* sorted by step */
  int32_t n_force, n_da, n_currents, reserved;
  int32_t* spikes_out;
  int64_t* spike_offsets;
  int64_t spikes_cap;
  snn_step_stats* stats;
  /* synapse probes: (w, sd) of `n_probes` synapses (indices in the caller's CSR order) after every
   * step of the window, probe_out[n_steps][n_probes][2] — `shist[time,:] = [net.s[n1,syn],
   * net.sd[n1,syn]]` of Old_net/Julia_DA_STDPv2.jl:91 / newnet_DASTDP.jl:57-60.  A window with probes
   * runs as plain launches with learn! in place (a diagnostics path). */
  const int64_t* probe_syn;
  double* probe_out;
  int32_t n_probes, reserved2;
} snn_step_io;
int32_t snn_step_ex(snn_net* net, const snn_step_io* io);

/* `net.da += x` (newnet_DASTDP.jl:55, teacher_sim.jl:133, critic_sim.jl:69) */
int32_t snn_add_dopamine(snn_net* net, int32_t instance, double amount);
/* `net.v .= c` (cartpole_fitness.jl:71) */
int32_t snn_reset_v(snn_net* net);

/* Field access (`net.s[n1,n2] = 0.0` newnet_DASTDP.jl:19, reads at :58-59).  Host buffers are
 * always float64 regardless of dtype; count = number of elements. */
int32_t snn_get_array(snn_net* net, int32_t field, double* host, int64_t count);
int32_t snn_set_array(snn_net* net, int32_t field, const double* host, int64_t count);
/* `shist[time,:] = [net.s[n1,syn], net.sd[n1,syn]]` (newnet_DASTDP.jl:57-60): (w, sd) of n (1..4096) synapses, by
 * their index in the CSR arrays passed to snn_set_synapses_*; either output may be NULL.  A gather of n elements
 * instead of two E-element copies; the positions of the last query are remembered.  Not with "learn_lazy". */
int32_t snn_get_synapses(snn_net* net, const int64_t* idx, int32_t n, double* w, double* sd);

int64_t snn_n_synapses(const snn_net* net);
int64_t snn_step_index(const snn_net* net);

/* Neuron partitioning of ONE network over several GPUs: this handle integrates
 * neurons [neuron_lo, neuron_hi) (multiples of 128) and must be given, through
 * snn_set_synapses_csr, exactly the synapses whose POST is in that range (rowptr still spans all
 * presynaptic neurons).  After every step's local detection the library calls `fn`, which must
 * all-gather the spike bitmap of the step: `words` points to the device bitmap (uint32[total_words]),
 * this rank owns [word_lo, word_lo+n_words), the other ranks' words must be valid when work
 * enqueued on `cuda_stream` after the call runs (e.g. ncclAllGather / torch.distributed on that
 * stream).  Traces of remote neurons, the spike list and `da` are replicated, so the union of the
 * ranks reproduces the single-GPU run.  Call before the first step. */
typedef int32_t (*snn_exchange_fn)(void* user, void* words, int64_t word_lo, int64_t n_words, int64_t total_words,
                                   void* cuda_stream);
int32_t snn_set_partition(snn_net* net, int64_t neuron_lo, int64_t neuron_hi, snn_exchange_fn fn, void* user);

/* In-library exchange for the neuron partition (no callback, the window runs as one CUDA graph):
 * every rank exports a blob describing its exchange ring (a CUDA IPC handle; a plain pointer when
 * the ranks are handles of one process), the blobs of all ranks are gathered by the host program
 * (e.g. torch.distributed.all_gather_object / MPI) and passed, in rank order, to
 * snn_partition_connect.  From then on neuron_kernel stores this rank's spike-bitmap words of every
 * step straight into the peers' rings over NVLink / NVSwitch, each word in one 8-byte store together
 * with its step number (flag-in-data: no fence, no separate flag); remote_kernel on the receiving
 * side spins on the step tag of exactly the words it needs.  Call after
 * snn_set_partition(lo, hi, NULL, NULL), before the first step; all ranks must have connected before
 * any of them steps (host barrier). */
#define SNN_PEER_BLOB_BYTES 192
int32_t snn_partition_export(snn_net* net, void* blob /* SNN_PEER_BLOB_BYTES */);
int32_t snn_partition_connect(snn_net* net, int32_t rank, int32_t world, const void* blobs /* world blobs */);

/* 0 (default): only the window events; the window runs as one CUDA graph when possible.
 * 1: CUDA events around every kernel class, plain stream-ordered launches (diagnostics).
 * 2: the CUDA-graph path with events around the learn_kernel launches only (bench.py roofline). */
int32_t snn_set_profiling(snn_net* net, int32_t enable);
/* Diagnostics (snn_config.reserved[1] & 2): %globaltimer stamps (ns) of the last window's kernels,
 * out[2][SNN_TL_CLASSES][SNN_TL_STEPS]: [0] = first block start, [1] = last block end; class
 * 0 neuron, 1 synapse, 2 ltp, 3 learn, 4 remote, 5 push-only synapse; index = window step + 1. */
#define SNN_TL_CLASSES 16
#define SNN_TL_STEPS 258
int32_t snn_debug_timeline(snn_net* net, uint64_t* out, int64_t count);
/* Use an externally owned CUDA stream (cudaStream_t) instead of the net's own. */
int32_t snn_set_stream(snn_net* net, void* cuda_stream);

/* ------------------------------------------------------------------------------------------
 * Batched layered LIF actors (variant F) — thousands of independent tiny networks, one warp each.
 * Replaces, for a whole population, Julia/Modules/Networks/net_arch.jl: Network(arch, param)
 * :98-128, input_LIF! :142-152, spike_LIF! :154-211, learn! :213-220, learn_critic! :222-235,
 * step_LIF! / step_LIF_critic! :256-272, and the decision/episode loop of
 * Julia/Actors/cartpole_fitness.jl:51-114 (with a restated CartPole-v1: gym is not part of the
 * reference).  Arrays carry a leading agent dimension; matrices are column-major per agent,
 * exactly the layout of the Julia `e[l]`.  fp64 only (bit-identical to the oracle). */
typedef struct snn_actor_config {
  int32_t struct_size;   /* = sizeof(snn_actor_config) */
  int32_t device;
  int32_t n_agents;
  int32_t n_layers;      /* length(arch) */
  int32_t arch[8];       /* e.g. {8,8,8}: sum(arch) + sum(arch[2:end-1]) must be <= 64, every layer <= 32 */
  double vthresh, c;     /* cartpole_cfglif.yml:7-41 keys */
  double wei, wie, wmax;
  double theta, ttheta;
  double imin;
  double apost, apre;
  double tde, ttrace;
  double taulif;
  double eta, winh;      /* learn_critic! only */
  double da_inc;         /* teacher_sim.jl:133 */
  int32_t reserved[8];
} snn_actor_config;

typedef struct snn_actor snn_actor;

/* defaults = Julia/Actors/cartpole_cfglif.yml, arch [8,8,8], 1 agent */
int32_t snn_actor_config_default(snn_actor_config* cfg);
int32_t snn_actor_create(const snn_actor_config* cfg, snn_actor** out);
int32_t snn_actor_destroy(snn_actor* a);
/* field: SNN_FIELD_V|HO [n_agents][n_neurons], TRACE [n_agents][n_exc], DA [n_agents],
 * W|SD with `layer` l: e[l] / de[l] as [n_agents][arch[l]*arch[l+1]] column-major
 * (`net.s.e[l][idx] = w` cartpole_fitness.jl:147). */
int32_t snn_actor_get_array(snn_actor* a, int32_t field, int32_t layer, double* host, int64_t count);
int32_t snn_actor_set_array(snn_actor* a, int32_t field, int32_t layer, const double* host, int64_t count);
/* n_steps calls of step_LIF!(net, spikes, train) (critic: step_LIF_critic!) with the constant
 * input spikes = [(i, intensity[agent][i]) for i in 1:arch[1]] (int_spikes,
 * PoissonStateSpike.jl:62-68).  reset_v: `net.v .= c` first (cartpole_fitness.jl:71).
 * counts [n_agents][arch[end]]: last-layer spike counts from msec >= length(arch)-1 (:76-78), may
 * be NULL.  spike_masks [n_agents][n_steps]: bit j = neuron j spiked at that step, may be NULL. */
int32_t snn_actor_window(snn_actor* a, int32_t n_steps, const double* intensity, int32_t train, int32_t critic,
                         int32_t reset_v, int32_t* counts, int32_t* spike_masks, double* device_ms);
/* learn!(net) (critic = 0, net_arch.jl:213-220) or learn_critic!(net) (critic != 0, :222-235) on its own, for
 * every agent: the scripts that call it outside step_LIF! (critic_sim.jl:107, Latencysim_rewardv0.jl:131). */
int32_t snn_actor_learn(snn_actor* a, int32_t critic);
/* play_episode (cartpole_fitness.jl:51-114) for every agent, environment on the device.
 * matalpha [n_agents][arch[1]*4] column-major, state0 [n_agents][4] (the env.reset() observation),
 * tie_actions [n_agents][n_decisions_max]: replayed rand(rng_action, 0:1) stream (:89),
 * teach_actions: NULL or [n_agents][n_decisions_max] replayed rand(0:1) of teacher_sim.jl:132 —
 * with train != 0, da += (action == draw ? 1 : -200) * da_inc after every decision.
 * Outputs: rewards [n_agents] (tot_reward), n_decisions, n_rand, final state — any may be NULL
 * except rewards. */
int32_t snn_actor_episodes(snn_actor* a, const double* matalpha, const double* state0, const int32_t* tie_actions,
                           const int32_t* teach_actions, int32_t n_decisions_max, int32_t win_size, int32_t train,
                           double* rewards, int32_t* n_decisions, int32_t* n_rand, double* state_out,
                           double* device_ms);

/* The same decision loop with a selectable restated environment (no teacher signal).
 *   SNN_ENV_CARTPOLE     exactly snn_actor_episodes (matalpha arch[1] x 4).
 *   SNN_ENV_MOUNTAINCAR  play_episode of Actors/mountcar_fitness.jl:48-107: matalpha
 *                        [n_agents][arch[1]*2] column-major over norm_mountcar(obs) (:13-20); state0
 *                        [n_agents][4] = (position, velocity, -, -); three output pools of div(n,3),
 *                        div(n,3) and the remaining neurons, action = argmax - 1 (:33-39, :77-80);
 *                        every tied pair of pools draws `rand(rng_action, pool)` (:83-93) and the
 *                        decision then costs 2 instead of 1.
 * tie_picks [n_agents][tie_stride]: the replayed draw stream as 0-based indices into the
 * two-element pool (CartPole: the action itself); a MountainCar decision can consume three.
 * tie_stride < 0: tie_picks is ONE stream of -tie_stride draws shared by all agents (the reference's default
 * `rng_seed = true`: every episode draws from a fresh MersenneTwister(0), cartpole_fitness.jl:55).  An agent
 * that meets more ties than the stream holds makes the call fail with SNN_ERR_CAPACITY.  An episode also ends at
 * gym's TimeLimit (500 decisions CartPole-v1, 200 MountainCar-v0) whatever n_decisions_max says. */
#define SNN_ENV_CARTPOLE 0
#define SNN_ENV_MOUNTAINCAR 1
int32_t snn_actor_episodes_env(snn_actor* a, int32_t env, const double* matalpha, const double* state0,
                               const int32_t* tie_picks, int32_t tie_stride, int32_t n_decisions_max, int32_t win_size,
                               int32_t train, double* rewards, int32_t* n_decisions, int32_t* n_rand,
                               double* state_out, double* device_ms);

/* ---------------------------------------------------------------------------------------------
 * Grid-cell state encoders on the device (Julia/Modules/Gridcells): the distance tests of
 *   input_spikes(obs, grid_cells::Array{HexGrid})      StateSpikeGaussGrid.jl:83-109
 *   latency_spikes(obs, grid_cells, win_size)          StateSpikeGaussGrid.jl:111-139
 * for a batch of observations.  The grids (HexGrid.jl, GridGenerator.jl: `grid.grid`, `grid.node_size`) are built
 * by the caller and uploaded once: grid g has the nodes node_xy[2*node_off[g] .. 2*node_off[g+1]) as (x, y) pairs
 * in the grid's own order (the encoder takes the FIRST node inside the radius, `filter(...)[1]`).
 * ------------------------------------------------------------------------------------------- */
typedef struct snn_gridcode snn_gridcode;
int32_t snn_gridcode_create(int32_t device, int32_t n_grids, const int64_t* node_off, const double* node_xy,
                            const double* node_size, snn_gridcode** out);
int32_t snn_gridcode_destroy(snn_gridcode* g);
/* obs [n_obs][4] CartPole observations.  Neuron i (0-based) = grid i on (x, theta % pi), neuron n_grids + i =
 * grid i on (x_dot, theta_dot), both scaled from [-15, 15] to the unit square (:85-89).
 *   intensity_out [n_obs][2*n_grids] or NULL: 0.999999 * exp(-2 d / node_size) of the first node closer than
 *     reach * node_size (reach 1.5 for input_spikes :95, 1.25 for latency_spikes :123), 0 where no node is;
 *   slot_out [n_obs][2*n_grids] or NULL: the 0-based millisecond of the win_size window at which the neuron fires,
 *     win_size - floor(win_size * intensity) - 1 (:131-134), -1 where it stays silent. */
int32_t snn_gridcode_encode(snn_gridcode* g, const double* obs, int32_t n_obs, double reach, int32_t win_size,
                            double* intensity_out, int32_t* slot_out);

#ifdef __cplusplus
}
#endif
#endif /* SNN_B200_H */
