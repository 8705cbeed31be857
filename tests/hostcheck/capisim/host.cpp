// libapg_b200_sim.so, part 1: the C-ABI layer ITSELF (csrc/capi.cu, csrc/capi_prep.cu: argument checks, workspace
// plan, weight packing, kernel selection by configuration / environment) on the CUDA-runtime shim of gpu_sim.h.
// "Device" pointers are host pointers; launches run on the CPU models.  Test-only.
#define APG_SIM 1
#include "../te_sim.h"

#include "../../../apg_trajectory_tracking_b200/csrc/capi.cu"
#include "../../../apg_trajectory_tracking_b200/csrc/capi_prep.cu"

// violations recorded by the models since the last call (messages joined into buf)
extern "C" __attribute__((visibility("default"))) int apg_sim_take_errors(char* buf, int len) {
  std::vector<std::string>& e = sim::errors();
  std::string all;
  for (const std::string& s : e) all += s + "; ";
  if (buf && len > 0) { strncpy(buf, all.c_str(), (size_t)len - 1); buf[len - 1] = 0; }
  const int n = (int)e.size();
  e.clear();
  return n;
}
