// TEST-ONLY host emulation of the CUDA physics core: compiles gymnasium_robotics_b200/csrc/sim_core.cuh with
// WARP_W == 1 (one "lane" runs every strided loop) so the fp32 kernel arithmetic can be compared with the fp64 oracle
// on a machine without a GPU.  Never linked into the product library; the product path is CUDA only.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#ifdef B200_HOST_WARP
#include "hostwarp.h"   // WARP_W == 32: the lane-parallel code paths on 32 lock-step fibers
#endif
#include "../../gymnasium_robotics_b200/csrc/fetch_task.cuh"

// run f(ctx) once (WARP_W == 1) or on the 32 lanes of the emulated warp
#if defined(B200_HOST_WARP) && defined(B200_STAGE_TIMING)
static long long g_stage_rounds[TM_COUNT];   // lane 0's per-stage collective counts, accumulated over calls
extern "C" int hostsim_stage_rounds(long long* out, int reset) {
  for (int k = 0; k < TM_COUNT; k++) { out[k] = g_stage_rounds[k]; if (reset) g_stage_rounds[k] = 0; }
  return TM_COUNT;
}
#endif
template <class F> static void warp_call(const Ctx& base, F f) {
#if defined(B200_HOST_WARP) && defined(B200_STAGE_TIMING)
  hw_run([&](int lane) {
    long long tim[TM_COUNT] = {0};
    Ctx c = base; c.lane = lane; c.tim = tim; f(c);
    if (lane == 0) for (int k = 0; k < TM_COUNT; k++) g_stage_rounds[k] += tim[k];
  });
#elif defined(B200_HOST_WARP)
  hw_run([&](int lane) { Ctx c = base; c.lane = lane; f(c); });
#else
  f(base);
#endif
}

#ifdef B200_WIDE
static const bool kWide = true;   // 64-bit dof masks: the tables carry two words per mask
#else
static const bool kWide = false;
#endif

struct HostSim {
  std::vector<uint8_t> blob;
  b200_model_view view;
  std::vector<uint32_t> model;
  std::vector<float> scratch;
  Ctx ctx;
  std::string err;
};

extern "C" {
void* hostsim_create(const void* blob, size_t n, const double* eq_data, const float* ref, int penv_body, int ngrp_cap) {
  HostSim* s = new HostSim;
  s->blob.assign((const uint8_t*)blob, (const uint8_t*)blob + n);
  if (b200_model_parse(s->blob.data(), n, &s->view) != 0) { delete s; return nullptr; }
  if (dm_build(s->view, eq_data, ref, s->model, s->err, penv_body, kWide, ngrp_cap) != 0) { fprintf(stderr, "hostsim: %s\n", s->err.c_str()); delete s; return nullptr; }
  const DMHead* h = (const DMHead*)s->model.data();
  s->scratch.assign(h->scr_words, 0.f);
  s->ctx.mg = s->model.data(); s->ctx.mw = s->model.data(); s->ctx.h = h; s->ctx.s = s->scratch.data(); s->ctx.lane = 0;
  return s;
}
void hostsim_destroy(void* p) { delete (HostSim*)p; }
float* hostsim_scratch(void* p) { return ((HostSim*)p)->scratch.data(); }
int hostsim_scr_words(void* p) { return ((HostSim*)p)->ctx.h->scr_words; }
int hostsim_npair(void* p) { return ((HostSim*)p)->ctx.h->npair; }
#ifdef B200_KITCHEN_GROUPS
int hostsim_nbgrp(void* p) { return ((HostSim*)p)->ctx.h->nbgrp; }
#endif
int hostsim_offset(void* p, const char* name) {
  const DMHead* h = ((HostSim*)p)->ctx.h;
#define X(nm, words) if (strcmp(name, #nm) == 0) return h->s_##nm;
  DM_SCRATCH_PERSIST(X)
#undef X
#define X(nm) if (strcmp(name, #nm) == 0) return h->s_##nm;
  DM_SCRATCH_UNION(X)
#undef X
  return -1;
}
void hostsim_forward(void* p) { warp_call(((HostSim*)p)->ctx, [](const Ctx& c) { forward<32>(c, true); }); }
void hostsim_step(void* p, int n) { warp_call(((HostSim*)p)->ctx, [n](const Ctx& c) { for (int i = 0; i < n; i++) { forward<32>(c, true); euler_step<32>(c); } }); }
void hostsim_kinematics(void* p) { warp_call(((HostSim*)p)->ctx, [](const Ctx& c) { kinematics(c); com_quantities(c); mass_matrix(c); }); }
int hostsim_task_size() { return (int)sizeof(FetchTask); }
int hostsim_env_step(void* p, const FetchTask* t, int mode, int nraw, float* st, const float* action, float* obs, float* achieved,
                     float* desired, float* reward, float* success) {
  int it = 0;
  warp_call(((HostSim*)p)->ctx, [&](const Ctx& c) {
    fetch_env_step<32>(c, *t, true, mode, nraw, st, action, obs, achieved, desired, reward, success, &it);
  });
  return it;
}
}
// in-kernel reset sampling (csrc/reset_sample.cuh), the same code the CUDA reset kernel runs
#include "../../gymnasium_robotics_b200/csrc/reset_sample.cuh"
extern "C" void hostsim_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { rs_philox4x32_10(ctr, key, out); }
extern "C" void hostsim_fetch_reset_record(const b200sim_fetch_reset_t* p, unsigned long long seed, unsigned env, unsigned episode, const float* rest,
                                           int stride, int st_qpos, int st_goal, float* rec) {
  rs_fetch_reset_record(*p, seed, env, episode, rest, stride, st_qpos, st_goal, rec);
}
extern "C" void hostsim_uniform_reset_record(const b200sim_uniform_reset_t* p, unsigned long long seed, unsigned env, unsigned episode, const float* rest,
                                             int stride, float* rec) {
  rs_uniform_reset_record(*p, seed, env, episode, rest, stride, rec);
}
extern "C" void hostsim_maze_reset_record(const b200sim_maze_reset_t* p, const float* goal_xy, const float* reset_xy, unsigned long long seed, unsigned env,
                                          unsigned episode, const float* rest, int stride, int st_qpos, int st_goal, float* rec) {
  rs_maze_reset_record(*p, goal_xy, reset_xy, seed, env, episode, rest, stride, st_qpos, st_goal, rec);
}
extern "C" void hostsim_hand_pose_record(const b200sim_hand_reset_t* p, const float* parallel, unsigned long long seed, unsigned env, unsigned episode,
                                         unsigned attempt, const float* rest, int stride, int st_qpos, int st_goal, int ngoal, float* rec) {
  rs_hand_pose_record(*p, parallel, seed, env, episode, attempt, rest, stride, st_qpos, st_goal, ngoal, rec);
}
extern "C" void hostsim_hand_goal(const b200sim_hand_reset_t* p, const float* parallel, unsigned long long seed, unsigned env, unsigned episode,
                                  int st_qpos, int st_goal, float* rec) {
  rs_hand_goal(*p, parallel, seed, env, episode, st_qpos, st_goal, rec);
}
extern "C" void hostsim_reach_reset_record(const b200sim_reach_reset_t* p, unsigned long long seed, unsigned env, unsigned episode, const float* rest,
                                           int stride, int st_goal, float* rec) {
  rs_reach_reset_record(*p, seed, env, episode, rest, stride, st_goal, rec);
}
extern "C" int hostsim_check_record(float* rec, int stride, const float* rest, const b200sim_keep_t* keep) {
  b200sim_keep_t k;
  k.n = 0;
  if (keep) k = *keep;
  return rs_check_record(rec, stride, rest, k);
}
#ifdef B200_HOST_WARP
extern "C" long hostsim_collectives() { return g_hw.collectives; }   // scheduler rounds so far = warp collectives executed (2 per shuffle / ballot, 1 per __syncwarp)
#endif
extern "C" int hostsim_model_words(void* p) { return ((HostSim*)p)->ctx.h->nwords; }
extern "C" int hostsim_hot_words(void* p) { return ((HostSim*)p)->ctx.h->hot_words; }
