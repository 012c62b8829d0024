#include <cstdio>
#include "../rao-blackwellized-slam-smoothing_b200/csrc/step_kernels.cuh"
using namespace rb;
int main() {
  ModelConsts mc{}; mc.family = FAM_DENSE_MAG3D; mc.n = 7; mc.nw = 6; mc.n_odo = 7;
  double hQ[36] = {0}; double dg[6] = {0.25,0.25,0.01,3e-8,3e-8,2.7e-5};
  for (int i = 0; i < 6; ++i) hQ[i*7] = dg[i];
  double* dQ; cudaMalloc(&dQ, sizeof hQ); cudaMemcpy(dQ, hQ, sizeof hQ, cudaMemcpyHostToDevice);
  double hx[7] = {0.01,0,0,1,0,0,0}, hxi[14] = {0,0,0,1,0,0,0, 0,0,0,1,0,0,0}, hdx[7] = {0,0,0,1,0,0,0};
  double *dxk, *dxn, *ddx, *dout; cudaMalloc(&dxk, 56); cudaMalloc(&dxn, 112); cudaMalloc(&ddx, 56); cudaMalloc(&dout, 16);
  cudaMemcpy(dxk, hx, 56, cudaMemcpyHostToDevice); cudaMemcpy(dxn, hxi, 112, cudaMemcpyHostToDevice); cudaMemcpy(ddx, hdx, 56, cudaMemcpyHostToDevice);
  k_dyn_logweight<<<1,128>>>(mc, 2, dxk, dxn, ddx, 0.01, dQ, 0, dout);
  double ho[2]; cudaMemcpy(ho, dout, 16, cudaMemcpyDeviceToHost);
  printf("kernel: %g %g %s\n", ho[0], ho[1], cudaGetErrorString(cudaDeviceSynchronize()));
}
