// snn_f32.cu — production instantiation (fp32 state and arithmetic, FMA contraction allowed).
#include "snn_engine.cuh"
namespace snn {
IEngine* make_engine_f32(const snn_config& cfg) { return new Engine<float>(cfg); }
}  // namespace snn
