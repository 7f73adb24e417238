// The step kernel: one warp integrates one env for a whole env-step (all sub-steps on chip); WPB warps share one copy of
// the model constants that a single thread stages into shared memory with a TMA bulk copy (cp.async.bulk + mbarrier).
// Included by b200sim.cu (32-bit dof masks, NVP <= 30) and by b200sim_wide.cu (B200_WIDE: 64-bit dof masks, NVP = 36).
#pragma once
#include <cuda_runtime.h>
#include "fetch_task.cuh"

#ifdef B200_STAGE_TIMING
static __device__ unsigned long long g_stage_cycles[TM_COUNT];
#endif

static __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int WPB, int NVP>
__global__ void __launch_bounds__(WPB * 32) fetch_kernel(const uint32_t* __restrict__ model_g, FetchTask task, int mode, int nraw,
                                                         int N, float* __restrict__ state, const float* __restrict__ actions,
                                                         const unsigned char* __restrict__ mask, float* __restrict__ obs,
                                                         float* __restrict__ achieved, float* __restrict__ desired,
                                                         float* __restrict__ reward, float* __restrict__ success,
                                                         int* __restrict__ info) {
  extern __shared__ __align__(128) uint32_t smem[];
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- stage the model constants: TMA 1-D bulk copy global -> shared, completion on an mbarrier
  const int model_words = ((const DMHead*)model_g)->hot_words;  // header + HOT arrays (uniform scalar load)
  const uint32_t bytes = (uint32_t)model_words * 4u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)),
                 "l"(model_g), "r"(bytes), "r"(smem_u32(&bar))
                 : "memory");
  }
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                   : "=r"(done)
                   : "r"(smem_u32(&bar))
                   : "memory");
    }
  }
  const DMHead* h = (const DMHead*)smem;
  const int env = blockIdx.x * WPB + warp;
  const bool active = env < N && !(mask && !mask[env]);  // warp-uniform
  Ctx c;
#ifdef B200_STAGE_TIMING
  long long tim[TM_COUNT];
  for (int k = 0; k < TM_COUNT; k++) tim[k] = 0;
  c.tim = tim;
  const long long t_begin = clock64();
#endif
  c.mg = model_g; c.mw = smem; c.h = h; c.lane = lane;
  c.s = (float*)(smem + model_words) + (size_t)warp * h->scr_words;
  const size_t e = active ? (size_t)env : 0;
  const float* act = actions ? actions + e * task.nact : nullptr;  // only dereferenced in MODE_STEP by active warps
  fetch_env_step<NVP>(c, task, active, mode, nraw, state + e * task.st_stride, act, obs + e * task.nobs, achieved + e * task.ngoal,
                      desired + e * task.ngoal, reward + e, success + e, info ? info + e : nullptr);
#ifdef B200_STAGE_TIMING
  if (lane == 0 && active) {
    long long sum = 0;
    for (int k = 0; k < TM_OTHER; k++) sum += tim[k];
    tim[TM_OTHER] = clock64() - t_begin - sum;
    for (int k = 0; k < TM_COUNT; k++) atomicAdd(&g_stage_cycles[k], (unsigned long long)tim[k]);
  }
#endif
}

