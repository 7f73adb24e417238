// TEST-ONLY: 32 lock-step fibers standing for the 32 lanes of a warp, so that the lane-parallel code paths of sim_core.cuh
// (B200_WARP_CODE: shuffle reductions, scans, ballots, the register Cholesky) run on the CPU exactly as written for the GPU.
// Every warp-collective (`__shfl_*_sync`, `__ballot_sync`, `__syncwarp`) is a rendezvous: a lane parks its value and yields to
// the scheduler, which resumes the lanes round-robin, so one scheduler round advances every lane by exactly one collective --
// the converged-warp execution model the kernels assume.  Deterministic, single OS thread (ucontext fibers).
// A lane that leaves the function while others still wait in a collective is reported (on the GPU that is a hang).
#pragma once
#include <ucontext.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <functional>

#define __device__
#define __noinline__
#define __forceinline__
#define __isShared(p) true
#ifndef __clang__
#define __builtin_assume(x) ((void)0)
#endif
struct float4 { float x, y, z, w; };

struct HostWarp {
  ucontext_t main, ctx[32];
  char* stack[32];
  bool done[32];
  uint32_t slot[32];
  int cur = -1, waiting = 0;
  long collectives = 0;
  std::function<void(int)> body;
};
static HostWarp g_hw;

static inline int hw_lane() { return g_hw.cur; }
// park until the scheduler's next round (every live lane has reached a collective by then)
static inline void hw_yield() {
  HostWarp& w = g_hw;
  int me = w.cur;
  w.waiting++;
  swapcontext(&w.ctx[me], &w.main);
}
static void hw_entry(int lane) {
  g_hw.body(lane);
  g_hw.done[lane] = true;
  swapcontext(&g_hw.ctx[lane], &g_hw.main);
}
// run body(lane) on 32 fibers in lock-step
static inline void hw_run(std::function<void(int)> body) {
  HostWarp& w = g_hw;
  w.body = body;
  const size_t SS = 1 << 20;
  for (int l = 0; l < 32; l++) {
    if (!w.stack[l]) w.stack[l] = (char*)malloc(SS);
    w.done[l] = false;
    getcontext(&w.ctx[l]);
    w.ctx[l].uc_stack.ss_sp = w.stack[l]; w.ctx[l].uc_stack.ss_size = SS; w.ctx[l].uc_link = &w.main;
    makecontext(&w.ctx[l], (void (*)())hw_entry, 1, l);
  }
  // lane order inside a scheduler round: ascending by default; HOSTWARP_ORDER=reverse | shuffle runs the lanes of every interval
  // between two collectives in another order -- a result that depends on it is a shared-memory race between lanes
  static int order_mode = -1;
  if (order_mode < 0) { const char* e = getenv("HOSTWARP_ORDER"); order_mode = !e ? 0 : (!strcmp(e, "reverse") ? 1 : (!strcmp(e, "shuffle") ? 2 : 0)); }
  uint32_t rng = 0x9E3779B9u;
  for (;;) {
    int live = 0, ndone = 0;
    w.waiting = 0;
    int perm[32];
    for (int l = 0; l < 32; l++) perm[l] = order_mode == 1 ? 31 - l : l;
    if (order_mode == 2) for (int l = 31; l > 0; l--) { rng = rng * 1664525u + 1013904223u; int j = (int)((rng >> 8) % (uint32_t)(l + 1)); int t = perm[l]; perm[l] = perm[j]; perm[j] = t; }
    for (int li = 0; li < 32; li++) {
      int l = perm[li];
      if (w.done[l]) { ndone++; continue; }
      w.cur = l;
      swapcontext(&w.main, &w.ctx[l]);
      if (!w.done[l]) live++; else ndone++;
    }
    w.cur = -1;
    if (live == 0) break;
    if (ndone > 0 && live > 0) { fprintf(stderr, "hostwarp: %d lane(s) left while %d wait in a warp collective (divergent exit)\n", ndone, live); abort(); }
    w.collectives++;
  }
}

// ---- the intrinsics, as rendezvous: write, round, read, round (nobody overwrites a slot before everyone has read it)
template <class T> static inline T hw_exchange(T v, int src) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  HostWarp& w = g_hw;
  int me = w.cur;
  memcpy(&w.slot[me], &v, 4);
  hw_yield();
  T r;
  memcpy(&r, &w.slot[src & 31], 4);
  hw_yield();
  return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return hw_exchange(v, src); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return hw_exchange(v, hw_lane() ^ m); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) { int me = hw_lane(); return hw_exchange(v, me >= d ? me - d : me); }
static inline unsigned __ballot_sync(unsigned, bool p) {
  HostWarp& w = g_hw;
  w.slot[w.cur] = p ? 1u : 0u;
  hw_yield();
  unsigned m = 0;
  for (int l = 0; l < 32; l++) if (!w.done[l] && w.slot[l]) m |= 1u << l;
  hw_yield();
  return m;
}
static inline void __syncwarp() { hw_yield(); }
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }

// -DB200_STAGE_TIMING in this build: the stage "clock" is the scheduler-round counter, so TIC / TOC attribute warp collectives to stages
static inline long long clock64() { return g_hw.collectives; }
