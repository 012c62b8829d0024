// Single translation unit of librbslam.so: the kernels live in headers shared by
// the three source files below, so they are compiled once here.
#include "engine.cu"
#include "smoother.cu"
#include "ops.cu"
#include "sharded.cu"
#include "ekf.cu"
#include "localization.cu"
