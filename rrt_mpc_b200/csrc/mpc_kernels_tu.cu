// mpc_kernels_tu.cu — one instantiation of K_solve / K_rollout per translation unit (nvcc -DMPC_TU=n; see mpc_kernels.cuh).
#include "mpc_kernels.cuh"

#if MPC_TU == 0
#define KS mpc_solve_kernel<256, 1>
cudaError_t solve_set_smem_0(int bytes) { return cudaFuncSetAttribute(KS, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); }
void solve_launch_0(int grid, int threads, int smem, cudaStream_t st, const Params& p, const Settings& s, const BatchArgs& a, int P, int F) { KS<<<grid, threads, smem, st>>>(p, s, a, P, F); }
#elif MPC_TU == 1
#define KS mpc_solve_kernel<128, 2>
cudaError_t solve_set_smem_1(int bytes) { return cudaFuncSetAttribute(KS, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); }
void solve_launch_1(int grid, int threads, int smem, cudaStream_t st, const Params& p, const Settings& s, const BatchArgs& a, int P, int F) { KS<<<grid, threads, smem, st>>>(p, s, a, P, F); }
#elif MPC_TU == 2
cudaError_t rollout_set_smem(int bytes) { return cudaFuncSetAttribute(mpc_rollout_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); }
cudaError_t rollout_occupancy(int bytes, int* n) { return cudaOccupancyMaxActiveBlocksPerMultiprocessor(n, mpc_rollout_kernel<1>, 32, bytes); }
void rollout_launch(int grid, int smem, cudaStream_t st, const Params& p, const Settings& s, const cudampc_rollout_cfg& cfg, const RolloutArgs& a) { mpc_rollout_kernel<1><<<grid, 32, smem, st>>>(p, s, cfg, a); }
#else
#error "MPC_TU must be 0..2"
#endif
