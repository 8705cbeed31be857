// libapg_b200_sim.so, part 2: every kernel built on the tile engine / plain CUDA, with its real launcher.
#define APG_SIM 1
#include "../te_sim.h"

#include "../../../apg_trajectory_tracking_b200/csrc/misc_kernels.cu"
#include "../../../apg_trajectory_tracking_b200/csrc/prep_kernels.cu"
#include "../../../apg_trajectory_tracking_b200/csrc/hutter_kernels.cu"
#include "../../../apg_trajectory_tracking_b200/csrc/eval_kernels.cu"
#include "../../../apg_trajectory_tracking_b200/csrc/simple_kernels.cu"
#include "../../../apg_trajectory_tracking_b200/csrc/learnt_kernels.cu"
#include "../../../apg_trajectory_tracking_b200/csrc/p2p_kernels.cu"
#include "../../../apg_trajectory_tracking_b200/csrc/rec_kernels.cu"
#include "../../../apg_trajectory_tracking_b200/csrc/lstm_kernels.cu"
