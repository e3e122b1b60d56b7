// K10 instantiations for clusters of 2 CTAs (see k10_step_cluster.cuh)
#include "k10_step_cluster.cuh"
K10_DEFINE_CL(2)
