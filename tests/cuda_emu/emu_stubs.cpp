// TEST INFRASTRUCTURE: the tensor-engine entry points are not emulated (tcgen05 has no host meaning); the emulated
// library reports the tensor engine as unavailable so that every call takes the SIMT path.
#include "pnr_common.cuh"

namespace pnr {
bool tc_supported(const PnrScene&, const PnrMlp&) { return false; }
size_t tc_workspace_bytes(const PnrScene&, const PnrMlp&, int64_t) { return 0; }
int tc_field_eval(const PnrScene&, const PnrMlp&, const float*, const PointSource&, int64_t, float*, void*, size_t,
                  cudaStream_t) {
  set_error("tensor engine is not available in the emulator build");
  return PNR_ERR_UNSUPPORTED;
}
}  // namespace pnr

extern "C" {
size_t pnr_pack_mlp_bytes(const PnrMlp*) { return 0; }
int pnr_pack_mlp(const PnrMlp*, void*, size_t, void*) { return PNR_ERR_UNSUPPORTED; }
size_t pnr_project_latent_bytes(const PnrScene*, const PnrMlp*) { return 0; }
int pnr_project_latent(const PnrScene*, const PnrMlp*, float*, size_t, void*, size_t, void*) { return PNR_ERR_UNSUPPORTED; }
int pnr_tc_status(int*) { return 0; }
int pnr_tc_counters(unsigned long long* out8) {
  for (int i = 0; i < 8; ++i) out8[i] = 0;
  return 0;
}
}
