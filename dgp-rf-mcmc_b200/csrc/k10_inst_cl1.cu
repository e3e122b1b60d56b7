// K10 instantiations for clusters of 1 CTA (see k10_step_cluster.cuh)
#include "k10_step_cluster.cuh"
K10_DEFINE_CL(1)
