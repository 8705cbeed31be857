// Test-only: the two kernels of the gradient exchange over peer memory THEMSELVES (csrc/p2p_kernels.cu, unchanged
// source, -DAPG_SIM) on the CPU thread model; every "rank" is a launch with its own symmetric buffer, pointer tables,
// ticket word and partials - the ticket / last-CTA signalling, the flag wait and the rank-ordered sum run as written.
#define APG_SIM 1
#include "te_sim.h"

#include "../../apg_trajectory_tracking_b200/csrc/p2p_kernels.cu"

using namespace apg;

// mem: all ranks' symmetric buffers back to back ([world][total_floats]); partials: [world][ncta][n]
// one step `epoch` of all ranks in the given launch order; grads / params / bufs: [world][n]
extern "C" int hc_p2psim_step(float* mem, int world, int n, unsigned epoch, const float* partials, int ncta,
                              float scale, const int* order, float* grads, float* params, float* bufs, float lr,
                              float momentum, unsigned* tickets, char* err, int err_len) {
  const GradCommLayout L{world, n};
  const int set = (int)(epoch & 1u);
  std::vector<float*> slots(world);
  std::vector<unsigned*> flags(world);
  for (int q = 0; q < world; ++q) {
    slots[q] = mem + (size_t)q * L.total_floats() + L.slot_off(set, 0);
    flags[q] = reinterpret_cast<unsigned*>(mem + (size_t)q * L.total_floats() + L.flag_off(set, 0));
  }
  const int blocks = (n + 127) / 128;
  try {
    for (int i = 0; i < world; ++i) {
      const int r = order[i];
      sim::launch((n + 31) / 32, 128, [&]() {               // 32 parameters x 4 CTA slices per block (the real launcher)
        apg_reduce_scatter_p2p_kernel(partials + (size_t)r * ncta * n, ncta, n, scale, 0, 0, 0, slots.data(),
                                      flags.data(), r, world, epoch, tickets + r);
      });
    }
    for (int r = 0; r < world; ++r)
      sim::launch(blocks, 128, [&]() {
        apg_gather_sgd_p2p_kernel(slots[r], flags[r], world, n, epoch, grads + (size_t)r * n,
                                  params ? params + (size_t)r * n : nullptr, bufs ? bufs + (size_t)r * n : nullptr, lr,
                                  momentum);
      });
  } catch (const std::exception& ex) {
    sim::fail(std::string("exception: ") + ex.what());
  }
  std::vector<std::string>& e = sim::errors();
  std::string all;
  for (const std::string& s : e) all += s + "; ";
  if (err && err_len > 0) { strncpy(err, all.c_str(), (size_t)err_len - 1); err[err_len - 1] = 0; }
  const int ne = (int)e.size();
  e.clear();
  return ne;
}
