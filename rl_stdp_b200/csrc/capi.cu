// capi.cu — the extern "C" boundary of libsnn_b200.so (declared in include/snn_b200.h).
// Plain pointers and sizes only; every error becomes a status code + thread-local message.
#include <string.h>
#include <stdio.h>
#include <time.h>
#include <unistd.h>

#include <memory>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: ranges cost a pointer test unless a tool (nsys, ncu --nvtx) is attached

#include "actor.h"
#include "snn_common.h"

namespace {
// one NVTX range per entry point that launches device work: a timeline shows which API call a kernel belongs to
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace

struct snn_actor {
  snn::Actor impl;
};

struct snn_net {
  snn_config cfg;
  std::unique_ptr<snn::IEngine> eng;
};

namespace snn {
// 64 random bits drawn once per process (/dev/urandom; clock + pid + an address if that fails): tells two
// processes with the same pid (different pid namespaces) apart when peers exchange raw device pointers
uint64_t process_nonce() {
  static const uint64_t nonce = [] {
    uint64_t x = 0;
    if (FILE* f = fopen("/dev/urandom", "rb")) {
      if (fread(&x, sizeof(x), 1, f) != 1) x = 0;
      fclose(f);
    }
    if (x == 0) {
      struct timespec ts; clock_gettime(CLOCK_REALTIME, &ts);
      x = (uint64_t)ts.tv_nsec * 0x9E3779B97F4A7C15ull ^ ((uint64_t)getpid() << 32) ^ (uint64_t)(uintptr_t)&ts ^ (uint64_t)ts.tv_sec;
      if (x == 0) x = 1;
    }
    return x;
  }();
  return nonce;
}
}  // namespace snn

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
}  // namespace
namespace snn { int fail_with(int code, const std::string& msg) { return fail(code, msg); } }
namespace {

int validate(const snn_config& c, std::string& err) {
  if (c.struct_size != (int32_t)sizeof(snn_config)) { err = "snn_config.struct_size mismatch (ABI)"; return SNN_ERR_INVALID; }
  if (c.dtype != SNN_F32 && c.dtype != SNN_F64) { err = "dtype must be SNN_F32 or SNN_F64"; return SNN_ERR_INVALID; }
  if (c.mode != SNN_MODE_FAST && c.mode != SNN_MODE_EXACT) { err = "bad mode"; return SNN_ERR_INVALID; }
  if (c.n_neurons <= 0 || c.n_per_instance <= 0 || c.n_neurons % c.n_per_instance != 0) {
    err = "n_neurons must be a positive multiple of n_per_instance"; return SNN_ERR_INVALID;
  }
  if (c.n_exc < 0 || c.n_exc > c.n_per_instance) { err = "0 <= n_exc <= n_per_instance"; return SNN_ERR_INVALID; }
  if (c.max_delay < 1 || c.max_delay > snn::kMaxDelay) { err = "max_delay must be in 1..63"; return SNN_ERR_INVALID; }
  if (c.n_neurons > (1ll << snn::kPreBits)) { err = "n_neurons must be <= 2^26"; return SNN_ERR_INVALID; }
  if (c.noise_mode < SNN_NOISE_NONE || c.noise_mode > SNN_NOISE_PHILOX) { err = "bad noise_mode"; return SNN_ERR_INVALID; }
  if (c.sd_decay_mode < 0 || c.sd_decay_mode > 2 || c.rate_mode < 0 || c.rate_mode > 2) { err = "bad sd_decay_mode / rate_mode"; return SNN_ERR_INVALID; }
  if (!(c.wmin <= c.wmax)) { err = "wmin must be <= wmax"; return SNN_ERR_INVALID; }
  for (int q = 0; q < 8; ++q)
    if (c.reserved[q] != 0) { err = "snn_config.reserved must be zero (ABI 3: tuning switches are set by name with snn_set_param)"; return SNN_ERR_INVALID; }
  return SNN_OK;
}
}  // namespace

extern "C" {

const char* snn_last_error(void) { return g_err.c_str(); }
int32_t snn_abi_version(void) { return SNN_ABI_VERSION; }

int32_t snn_config_default(snn_config* c) {
  if (!c) return fail(SNN_ERR_INVALID, "null config");
  memset(c, 0, sizeof(*c));
  c->struct_size = (int32_t)sizeof(snn_config);
  c->dtype = SNN_F32; c->mode = SNN_MODE_FAST; c->device = 0;
  c->n_neurons = 1000; c->n_per_instance = 1000; c->n_exc = 800; c->max_delay = 1;
  c->threshold_ge = 0; c->vthresh = 30.0; c->v_reset = -65.0;
  c->a_exc = 0.02; c->a_inh = 0.1; c->d_exc = 8.0; c->d_inh = 2.0;
  c->trace_set = 0.1; c->trace_decay = 0.95; c->apost = 1.0; c->apre = 1.5;
  c->da_decay = 0.995; c->learn_base = 0.002; c->eta = 0.0; c->wmin = 0.0; c->wmax = 4.0;
  c->sd_decay = 0.99; c->sd_decay_mode = SNN_SD_DECAY_ON_LEARN; c->rate_mode = SNN_RATE_BASE_PLUS_DA;
  c->noise_mode = SNN_NOISE_REPLAY; c->plastic_ei = 1; c->noise_amp = 13.0; c->noise_seed = 0;
  c->theta = 0.0; c->ttheta = 1.0; c->ho_from = 0; c->ho_inh = 0; c->variant_g = 0;
  return SNN_OK;
}

int32_t snn_create(const snn_config* cfg, snn_net** out) {
  if (!cfg || !out) return fail(SNN_ERR_INVALID, "null argument");
  *out = nullptr;
  std::string err;
  int rc = validate(*cfg, err);
  if (rc) return fail(rc, err);
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev <= 0)
    return fail(SNN_ERR_NODEVICE, std::string("no CUDA device (libsnn_b200 has no CPU fallback): ") +
                                      (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0"));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(SNN_ERR_INVALID, "device ordinal out of range");
  snn_net* n = new (std::nothrow) snn_net;
  if (!n) return fail(SNN_ERR_INVALID, "out of host memory");
  n->cfg = *cfg;
  n->eng.reset(cfg->dtype == SNN_F64 ? snn::make_engine_f64(*cfg) : snn::make_engine_f32(*cfg));
  rc = n->eng->init(err);
  if (rc) { delete n; return fail(rc, err); }
  *out = n;
  return SNN_OK;
}

int32_t snn_destroy(snn_net* net) {
  if (!net) return SNN_OK;
  cudaSetDevice(net->cfg.device);
  delete net;
  return SNN_OK;
}

int32_t snn_set_synapses_csr(snn_net* net, const int64_t* rowptr, const int32_t* post, const double* w,
                             const uint8_t* delay) {
  if (!net || !rowptr) return fail(SNN_ERR_INVALID, "null argument");
  const int64_t E = rowptr[net->cfg.n_neurons];
  if (E > 0 && (!post || !w)) return fail(SNN_ERR_INVALID, "null post / w");
  std::string err;
  NvtxRange nv("snn_set_synapses_csr");
  cudaSetDevice(net->cfg.device);
  snn::Topo topo;
  int rc = snn::build_topology(net->cfg, rowptr, post, delay, &topo, net->eng->stream(), err);
  if (rc) return fail(rc, err);
  rc = net->eng->set_topology(std::move(topo), w, err);
  if (rc) return fail(rc, err);
  return SNN_OK;
}

int32_t snn_set_synapses_dense(snn_net* net, const double* s, const int64_t* connection) {
  if (!net || !s || !connection) return fail(SNN_ERR_INVALID, "null argument");
  const int64_t N = net->cfg.n_neurons;
  if (net->cfg.n_per_instance != N) return fail(SNN_ERR_INVALID, "dense synapses: single instance only");
  if (net->cfg.max_delay != 1) return fail(SNN_ERR_INVALID, "dense synapses carry no delays: max_delay must be 1");
  std::vector<int64_t> rowptr(N + 1, 0);
  std::vector<int32_t> post;
  std::vector<double> w;
  for (int64_t i = 0; i < N; ++i) {
    for (int64_t j = 0; j < N; ++j)
      if (connection[i + N * j] != 0) { post.push_back((int32_t)j); w.push_back(s[i + N * j]); }
    rowptr[i + 1] = (int64_t)post.size();
  }
  return snn_set_synapses_csr(net, rowptr.data(), post.data(), w.data(), nullptr);
}

int32_t snn_set_synapses_dense_col(snn_net* net, const double* s, const int64_t* connection) {
  if (!net || !s || !connection) return fail(SNN_ERR_INVALID, "null argument");
  const int64_t N = net->cfg.n_neurons;
  if (net->cfg.n_per_instance != N) return fail(SNN_ERR_INVALID, "dense synapses: single instance only");
  if (net->cfg.max_delay != 1) return fail(SNN_ERR_INVALID, "dense synapses carry no delays: max_delay must be 1");
  std::vector<int64_t> rowptr(N + 1, 0);
  std::vector<int32_t> post;
  std::vector<double> w;
  for (int64_t i = 0; i < N; ++i) {       // presynaptic neuron = COLUMN i of the [post, pre] matrices
    for (int64_t j = 0; j < N; ++j)
      if (connection[j + N * i] != 0) { post.push_back((int32_t)j); w.push_back(s[j + N * i]); }
    rowptr[i + 1] = (int64_t)post.size();
  }
  return snn_set_synapses_csr(net, rowptr.data(), post.data(), w.data(), nullptr);
}

int32_t snn_step(snn_net* net, int32_t n_steps, const double* noise, const uint8_t* train_flags,
                 const snn_force_event* force, int32_t n_force, const snn_da_event* da, int32_t n_da,
                 int32_t* spikes_out, int64_t* spike_offsets, int64_t spikes_cap, snn_step_stats* stats) {
  if (!net) return fail(SNN_ERR_INVALID, "null net");
  if ((n_force > 0 && !force) || (n_da > 0 && !da) || n_force < 0 || n_da < 0)
    return fail(SNN_ERR_INVALID, "event list pointer / count mismatch");
  snn::StepArgs a;
  a.n_steps = n_steps; a.noise = noise; a.train_flags = train_flags; a.force = force; a.n_force = n_force;
  a.da = da; a.n_da = n_da; a.spikes_out = spikes_out; a.spike_offsets = spike_offsets; a.spikes_cap = spikes_cap;
  a.stats = stats;
  std::string err;
  NvtxRange nv("snn_step");
  int rc = net->eng->step(a, err);
  if (rc) return fail(rc, err);
  return SNN_OK;
}

int32_t snn_step_ex(snn_net* net, const snn_step_io* io) {
  if (!net || !io) return fail(SNN_ERR_INVALID, "null argument");
  if (io->struct_size != (int32_t)sizeof(snn_step_io)) return fail(SNN_ERR_INVALID, "snn_step_io.struct_size mismatch (ABI)");
  if ((io->n_force > 0 && !io->force) || (io->n_da > 0 && !io->da) || (io->n_currents > 0 && !io->currents) ||
      io->n_force < 0 || io->n_da < 0 || io->n_currents < 0)
    return fail(SNN_ERR_INVALID, "event list pointer / count mismatch");
  snn::StepArgs a;
  a.n_steps = io->n_steps; a.noise = io->noise; a.train_flags = io->train_flags; a.force = io->force;
  a.n_force = io->n_force; a.da = io->da; a.n_da = io->n_da; a.cur = io->currents; a.n_cur = io->n_currents;
  a.spikes_out = io->spikes_out; a.spike_offsets = io->spike_offsets; a.spikes_cap = io->spikes_cap; a.stats = io->stats;
  a.probe_syn = io->probe_syn; a.n_probes = io->n_probes; a.probe_out = io->probe_out;
  if (io->n_probes < 0) return fail(SNN_ERR_INVALID, "n_probes < 0");
  std::string err;
  NvtxRange nv("snn_step_ex");
  int rc = net->eng->step(a, err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_add_dopamine(snn_net* net, int32_t instance, double amount) {
  if (!net) return fail(SNN_ERR_INVALID, "null net");
  std::string err;
  int rc = net->eng->add_dopamine(instance, amount, err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_set_partition(snn_net* net, int64_t neuron_lo, int64_t neuron_hi, snn_exchange_fn fn, void* user) {
  if (!net) return fail(SNN_ERR_INVALID, "null net");
  std::string err;
  int rc = net->eng->set_partition(neuron_lo, neuron_hi, fn, user, err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_partition_export(snn_net* net, void* blob) {
  if (!net || !blob) return fail(SNN_ERR_INVALID, "null argument");
  std::string err;
  int rc = net->eng->export_peer(reinterpret_cast<snn::PeerBlob*>(blob), err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_partition_connect(snn_net* net, int32_t rank, int32_t world, const void* blobs) {
  if (!net || !blobs) return fail(SNN_ERR_INVALID, "null argument");
  std::string err;
  int rc = net->eng->connect_peers(rank, world, reinterpret_cast<const snn::PeerBlob*>(blobs), err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_reset_v(snn_net* net) {
  if (!net) return fail(SNN_ERR_INVALID, "null net");
  std::string err;
  int rc = net->eng->reset_v(err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_set_param(snn_net* net, const char* name, double value) {
  if (!net || !name) return fail(SNN_ERR_INVALID, "null argument");
  std::string err;
  int rc = net->eng->set_param(name, value, err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_get_param(snn_net* net, const char* name, double* value) {
  if (!net || !name || !value) return fail(SNN_ERR_INVALID, "null argument");
  std::string err;
  int rc = net->eng->get_param(name, value, err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_checkpoint_bytes(snn_net* net, int64_t* bytes) {
  if (!net || !bytes) return fail(SNN_ERR_INVALID, "null argument");
  std::string err;
  int rc = net->eng->checkpoint_bytes(bytes, err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_checkpoint(snn_net* net, void* buf, int64_t bytes) {
  if (!net || !buf) return fail(SNN_ERR_INVALID, "null argument");
  std::string err;
  int rc = net->eng->checkpoint(buf, bytes, err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_restore(snn_net* net, const void* buf, int64_t bytes) {
  if (!net || !buf) return fail(SNN_ERR_INVALID, "null argument");
  std::string err;
  int rc = net->eng->restore(buf, bytes, err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_get_array(snn_net* net, int32_t field, double* host, int64_t count) {
  if (!net || !host) return fail(SNN_ERR_INVALID, "null argument");
  std::string err;
  int rc = net->eng->get_array(field, host, count, err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_get_synapses(snn_net* net, const int64_t* idx, int32_t n, double* w, double* sd) {
  if (!net) return fail(SNN_ERR_INVALID, "null net");
  std::string err;
  int rc = net->eng->get_synapses(idx, n, w, sd, err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_set_array(snn_net* net, int32_t field, const double* host, int64_t count) {
  if (!net || !host) return fail(SNN_ERR_INVALID, "null argument");
  std::string err;
  int rc = net->eng->set_array(field, host, count, err);
  return rc ? fail(rc, err) : SNN_OK;
}

int64_t snn_n_synapses(const snn_net* net) { return net ? net->eng->n_synapses() : -1; }
int64_t snn_step_index(const snn_net* net) { return net ? net->eng->step_index() : -1; }

int32_t snn_set_profiling(snn_net* net, int32_t enable) {
  if (!net) return fail(SNN_ERR_INVALID, "null net");
  net->eng->set_profiling(enable);
  return SNN_OK;
}

int32_t snn_debug_timeline(snn_net* net, uint64_t* out, int64_t count) {
  if (!net || !out) return fail(SNN_ERR_INVALID, "null argument");
  std::string err;
  int rc = net->eng->get_timeline((unsigned long long*)out, count, err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_set_stream(snn_net* net, void* cuda_stream) {
  if (!net) return fail(SNN_ERR_INVALID, "null net");
  net->eng->set_stream((cudaStream_t)cuda_stream);
  return SNN_OK;
}

int32_t snn_actor_config_default(snn_actor_config* c) {
  if (!c) return fail(SNN_ERR_INVALID, "null config");
  memset(c, 0, sizeof(*c));
  c->struct_size = (int32_t)sizeof(snn_actor_config);
  c->device = 0; c->n_agents = 1; c->n_layers = 3; c->arch[0] = c->arch[1] = c->arch[2] = 8;
  c->vthresh = 1.0; c->c = 0.0; c->wei = 0.0; c->wie = 0.0; c->wmax = 1.1; c->theta = 0.0; c->ttheta = 0.998;
  c->imin = 1.0; c->apost = 0.1; c->apre = 0.15; c->tde = 0.99; c->ttrace = 0.9; c->taulif = 5.0;
  c->eta = 0.001; c->winh = 0.0; c->da_inc = 0.01;
  return SNN_OK;
}

int32_t snn_actor_create(const snn_actor_config* cfg, snn_actor** out) {
  if (!cfg || !out) return fail(SNN_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->struct_size != (int32_t)sizeof(snn_actor_config)) return fail(SNN_ERR_INVALID, "snn_actor_config.struct_size mismatch (ABI)");
  if (cfg->n_agents <= 0) return fail(SNN_ERR_INVALID, "n_agents must be > 0");
  if (cfg->n_layers < 2 || cfg->n_layers > snn::kActorMaxLayers) return fail(SNN_ERR_INVALID, "n_layers must be in 2..8");
  if (!(cfg->taulif != 0.0)) return fail(SNN_ERR_INVALID, "taulif must be non-zero");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev <= 0)
    return fail(SNN_ERR_NODEVICE, std::string("no CUDA device (libsnn_b200 has no CPU fallback): ") +
                                      (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0"));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(SNN_ERR_INVALID, "device ordinal out of range");
  snn_actor* a = new (std::nothrow) snn_actor;
  if (!a) return fail(SNN_ERR_INVALID, "out of host memory");
  std::string err;
  int rc = a->impl.init(*cfg, err);
  if (rc) { delete a; return fail(rc, err); }
  *out = a;
  return SNN_OK;
}

int32_t snn_actor_destroy(snn_actor* a) { delete a; return SNN_OK; }

int32_t snn_actor_get_array(snn_actor* a, int32_t field, int32_t layer, double* host, int64_t count) {
  if (!a || !host) return fail(SNN_ERR_INVALID, "null argument");
  std::string err;
  int rc = a->impl.get_array(field, layer, host, count, err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_actor_set_array(snn_actor* a, int32_t field, int32_t layer, const double* host, int64_t count) {
  if (!a || !host) return fail(SNN_ERR_INVALID, "null argument");
  std::string err;
  int rc = a->impl.set_array(field, layer, host, count, err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_actor_window(snn_actor* a, int32_t n_steps, const double* intensity, int32_t train, int32_t critic,
                         int32_t reset_v, int32_t* counts, int32_t* spike_masks, double* device_ms) {
  if (!a) return fail(SNN_ERR_INVALID, "null actor");
  std::string err;
  NvtxRange nv("snn_actor_window");
  int rc = a->impl.window(n_steps, intensity, train, critic, reset_v, counts, spike_masks, device_ms, err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_actor_learn(snn_actor* a, int32_t critic) {
  if (!a) return fail(SNN_ERR_INVALID, "null actor");
  std::string err;
  NvtxRange nv("snn_actor_learn");
  int rc = a->impl.learn(critic, err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_actor_episodes(snn_actor* a, const double* matalpha, const double* state0, const int32_t* tie_actions,
                           const int32_t* teach_actions, int32_t n_decisions_max, int32_t win_size, int32_t train,
                           double* rewards, int32_t* n_decisions, int32_t* n_rand, double* state_out,
                           double* device_ms) {
  if (!a) return fail(SNN_ERR_INVALID, "null actor");
  std::string err;
  NvtxRange nv("snn_actor_episodes");
  int rc = a->impl.episodes(SNN_ENV_CARTPOLE, matalpha, state0, tie_actions, n_decisions_max, teach_actions,
                            n_decisions_max, win_size, train, rewards, n_decisions, n_rand, state_out, device_ms, err);
  return rc ? fail(rc, err) : SNN_OK;
}

int32_t snn_actor_episodes_env(snn_actor* a, int32_t env, const double* matalpha, const double* state0,
                               const int32_t* tie_picks, int32_t tie_stride, int32_t n_decisions_max, int32_t win_size,
                               int32_t train, double* rewards, int32_t* n_decisions, int32_t* n_rand,
                               double* state_out, double* device_ms) {
  if (!a) return fail(SNN_ERR_INVALID, "null actor");
  std::string err;
  NvtxRange nv("snn_actor_episodes_env");
  int rc = a->impl.episodes(env, matalpha, state0, tie_picks, tie_stride, nullptr, n_decisions_max, win_size, train,
                            rewards, n_decisions, n_rand, state_out, device_ms, err);
  return rc ? fail(rc, err) : SNN_OK;
}

}  // extern "C"
