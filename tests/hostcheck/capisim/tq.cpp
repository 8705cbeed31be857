// libapg_b200_sim.so, part 5: second-generation tcgen05 path (csrc/tq_kernels.cu forward / dynamics / dX chain,
// csrc/tq_dw_kernels.cu streaming dW GEMM) on the tcgen05 model.
#define APG_TC_SIM 1
#define APG_SIM 1
#include "../tc_sim.h"

#include "../../../apg_trajectory_tracking_b200/csrc/tq_kernels.cu"
#include "../../../apg_trajectory_tracking_b200/csrc/tq_dw_kernels.cu"
