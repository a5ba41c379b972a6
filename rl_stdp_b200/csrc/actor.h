// actor.h — host-side class behind the snn_actor_* entry points (batched LIF actors, actor.cu).
#pragma once
#include <stdint.h>

#include <string>

#include "../../include/snn_b200.h"

namespace snn {

constexpr int kActorMaxLayers = 8;

class Actor {
 public:
  Actor();
  ~Actor();
  int init(const snn_actor_config& cfg, std::string& err);
  int get_array(int field, int layer, double* host, int64_t count, std::string& err);
  int set_array(int field, int layer, const double* host, int64_t count, std::string& err);
  int learn(int critic, std::string& err);
  int window(int n_steps, const double* intensity, int train, int critic, int reset_v, int32_t* counts,
             int32_t* spike_masks, double* device_ms, std::string& err);
  int episodes(int env, const double* matalpha, const double* state0, const int32_t* tie_actions, int tie_stride,
               const int32_t* teach_actions, int n_decisions_max, int win_size, int train, double* rewards, int32_t* n_dec, int32_t* n_rand,
               double* state_out, double* device_ms, std::string& err);
  const snn_actor_config& config() const { return cfg_; }

 private:
  int field_ptr(int field, int layer, double** ptr, int64_t* per_agent, int64_t* stride, std::string& err);
  struct Impl;
  Impl* impl_;
  snn_actor_config cfg_;
};

}  // namespace snn
