// Explicit instantiations of the pass-chain kernel (chain_kernel.cuh) for the pairs of wide shapes the plans produce
// (planner.cu: build_seq splits log2 N as evenly as possible, find_kernel picks ~256-thread CTAs in float and
// ~128-thread CTAs in double).  Compiled once per GENFFT_CSET so the sets build in parallel.
#include "registry.h"

namespace genfft_cuda {

#if GENFFT_CSET == 0
void register_chains_0(std::vector<ChainEntry>& v) {
  add_chain_shapes<float, 64, 32, 64, 32>(v);
  add_chain_shapes<float, 128, 32, 128, 32>(v);
  add_chain_shapes<float, 256, 16, 128, 32>(v);
}
#elif GENFFT_CSET == 1
void register_chains_1(std::vector<ChainEntry>& v) {
  add_chain_shapes<float, 256, 16, 256, 16>(v);
  add_chain_shapes<float, 512, 16, 512, 16>(v);
}
#elif GENFFT_CSET == 2
void register_chains_2(std::vector<ChainEntry>& v) {
  add_chain_shapes<double, 64, 32, 64, 32>(v);
  add_chain_shapes<double, 128, 16, 64, 32>(v);
  add_chain_shapes<double, 128, 16, 128, 16>(v);
}
#elif GENFFT_CSET == 3
void register_chains_3(std::vector<ChainEntry>& v) {
  add_chain_shapes<double, 256, 8, 128, 16>(v);
  add_chain_shapes<double, 256, 8, 256, 8>(v);
  add_chain_shapes<double, 512, 8, 512, 8>(v);
}
#else
#error "GENFFT_CSET must be 0..3"
#endif

}  // namespace genfft_cuda
