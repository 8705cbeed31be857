// libapg_b200_sim.so, part 3: tcgen05 forward / dX chain (csrc/hutter_tc_kernels.cu) on the tcgen05 model.
#define APG_TC_SIM 1
#define APG_SIM 1
#include "../tc_sim.h"

#include "../../../apg_trajectory_tracking_b200/csrc/hutter_tc_kernels.cu"
