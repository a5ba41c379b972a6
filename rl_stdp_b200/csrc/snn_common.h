// snn_common.h — host-side types shared by the translation units of libsnn_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/snn_b200.h"

namespace snn {

constexpr int kMaxDelay = 63;        // delay-1 is packed into 6 bits of in_pre
constexpr int kPreBits = 26;         // neuron ids < 2^26 in the packed incoming entry
constexpr uint32_t kPreMask = (1u << kPreBits) - 1u;

#define SNN_CUDA_OK(expr)                                                        \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) {                                                     \
      err = std::string(#expr) + ": " + cudaGetErrorString(_e);                  \
      return SNN_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

// Type-independent synapse topology, built once on the device (snn_build.cu).
// Internal synapse order: ascending (pre, delay, caller position).
struct Topo {
  int64_t N = 0, E = 0, Nper = 0, Ne = 0;
  int32_t Dmax = 1, R = 2, B = 1;
  uint32_t* seg = nullptr;      // [N*(Dmax+1)] seg[i*(Dmax+1)+a] = first synapse of row i with delay a+1
  int32_t* post = nullptr;      // [E]
  uint32_t* perm = nullptr;     // [E] internal position -> caller position
  uint32_t* colptr = nullptr;   // [N+1] incoming lists, ascending (pre, internal position)
  uint32_t* col_npl = nullptr;  // [N] length of the plastic prefix of each incoming list
  uint32_t* in_syn = nullptr;   // [E] internal synapse position
  uint32_t* in_pre = nullptr;   // [E] (delay-1) << 26 | pre
  void release();
};

int build_topology(const snn_config& cfg, const int64_t* h_rowptr, const int32_t* h_post,
                   const uint8_t* h_delay, Topo* out, cudaStream_t stream, std::string& err);

// Sort every [off[k], off[k+1]) segment of keys ascending (spike record formatting).
struct SortScratch {
  void* tmp = nullptr; size_t tmp_cap = 0;
  int32_t* out = nullptr; size_t out_cap = 0;
  void release();
};
int sort_segments(int32_t* d_keys, int64_t n_items, const int32_t* d_offsets, int32_t n_segments,
                  cudaStream_t stream, SortScratch* scratch, std::string& err);

// What a rank publishes so that its peers can store spike bitmaps into its ring (snn_partition_export)
constexpr uint64_t kPeerMagic = 0x534e4e5032503031ull;  // "SNNP2P01"
struct PeerBlob {
  uint64_t magic;
  int32_t pid, device;
  uint64_t raw_ring;               // device pointer of the exchange ring, valid inside the exporting process
  unsigned char ipc_ring[64];      // cudaIpcMemHandle_t of the exchange ring ((word, step) pairs)
  int32_t words_per_slot, ring;
  uint64_t proc_nonce;             // random per process: same pid AND same nonce <=> raw_ring is usable
  unsigned char pad[88];
};
uint64_t process_nonce();          // capi.cu
static_assert(sizeof(PeerBlob) == SNN_PEER_BLOB_BYTES, "PeerBlob must match SNN_PEER_BLOB_BYTES");

struct StepArgs {
  int32_t n_steps;
  const double* noise;
  const uint8_t* train_flags;
  const snn_force_event* force;
  int32_t n_force;
  const snn_da_event* da;
  int32_t n_da;
  const snn_current_event* cur = nullptr;
  int32_t n_cur = 0;
  int32_t* spikes_out;
  int64_t* spike_offsets;
  int64_t spikes_cap;
  snn_step_stats* stats;
  const int64_t* probe_syn = nullptr;  // caller-order synapse indices
  int32_t n_probes = 0;
  double* probe_out = nullptr;         // [n_steps][n_probes][2] = (w, sd) after each step
};

class IEngine {
 public:
  virtual ~IEngine() {}
  virtual int init(std::string& err) = 0;
  virtual int set_topology(Topo&& topo, const double* h_w_caller, std::string& err) = 0;
  virtual int step(const StepArgs& a, std::string& err) = 0;
  virtual int get_array(int field, double* host, int64_t count, std::string& err) = 0;
  virtual int set_array(int field, const double* host, int64_t count, std::string& err) = 0;
  virtual int get_synapses(const int64_t* idx, int32_t n, double* w, double* sd, std::string& err) = 0;
  virtual int add_dopamine(int instance, double amount, std::string& err) = 0;
  virtual int reset_v(std::string& err) = 0;
  virtual int checkpoint_bytes(int64_t* bytes, std::string& err) = 0;
  virtual int checkpoint(void* buf, int64_t bytes, std::string& err) = 0;
  virtual int restore(const void* buf, int64_t bytes, std::string& err) = 0;
  virtual int set_param(const char* name, double value, std::string& err) = 0;
  virtual int get_param(const char* name, double* value, std::string& err) = 0;
  virtual int set_partition(int64_t lo, int64_t hi, snn_exchange_fn fn, void* user, std::string& err) = 0;
  virtual int export_peer(PeerBlob* b, std::string& err) = 0;
  virtual int connect_peers(int rank, int world, const PeerBlob* blobs, std::string& err) = 0;
  virtual int get_timeline(unsigned long long* host, int64_t count, std::string& err) = 0;
  virtual int64_t n_synapses() const = 0;
  virtual int64_t step_index() const = 0;
  virtual void set_profiling(int mode) = 0;
  virtual void set_stream(cudaStream_t s) = 0;
  virtual cudaStream_t stream() const = 0;
};

IEngine* make_engine_f32(const snn_config& cfg);
IEngine* make_engine_f64(const snn_config& cfg);

}  // namespace snn
