// snn_f64.cu — parity instantiation (fp64, no contraction: Ops<double> uses __dmul_rn/__dadd_rn
// and this file is compiled with -fmad=false), reproduces the oracle's spike times exactly.
#include "snn_engine.cuh"
namespace snn {
IEngine* make_engine_f64(const snn_config& cfg) { return new Engine<double>(cfg); }
}  // namespace snn
