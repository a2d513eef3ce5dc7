// tests/emul/warp_emul.h -- TEST INFRASTRUCTURE ONLY.
//
// A single-warp SIMT emulator that lets the *same* device code that nvcc compiles for sm_100a
// (brotli_g_sdk_b200/csrc/page_decode.cuh) be compiled with g++ and executed on the CPU, one page per
// emulated warp. There is no GPU in the development container, so this is how kernel logic is
// exercised before it is sent to a B200; it is never part of the product path.
//
// Model: 32 lanes = 32 cooperative fibers (ucontext) on one OS thread. A lane runs until it reaches
// a warp collective (__shfl_sync, __ballot_sync, __syncwarp, ...), parks there, and the collective
// completes when all 32 lanes have arrived. Lanes therefore run far "out of lock-step" between
// collectives, which makes a missing __syncwarp() show up as a wrong answer instead of hiding it.
// Restrictions (checked): every collective uses the full mask and is reached by all 32 lanes.
#pragma once
#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define BGX_EMULATED 1

namespace wemu {

enum Op : int { OP_NONE = 0, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_BALLOT, OP_SYNC, OP_MATCH, OP_REDUCE, OP_BLOCKSYNC };

constexpr int kMaxWarps = 4;

struct Rendezvous {          // one per warp, plus one for the whole block (__syncthreads)
  uint64_t slot[32 * kMaxWarps];
  uint32_t aux[32 * kMaxWarps];
  int op[32 * kMaxWarps];
  int arrived = 0;
  uint64_t gen = 0;
  uint64_t result[2][32 * kMaxWarps];
  uint32_t result_aux[2][32 * kMaxWarps];
};

struct Block {
  ucontext_t sched;
  ucontext_t ctx[32 * kMaxWarps];
  char* stack[32 * kMaxWarps];
  bool done[32 * kMaxWarps];
  int nthreads = 32;
  int cur = 0;
  Rendezvous warp[kMaxWarps];
  Rendezvous block;
  struct NamedBar { int arrived = 0; uint64_t gen = 0; } nbar[16];   // bar.sync / bar.arrive with an id
  uint64_t collectives = 0;
  std::function<void()> body;
};

inline Block*& B() {
  static thread_local Block* b = nullptr;
  return b;
}
inline int lane() { return B()->cur & 31; }
inline int warp_id() { return B()->cur >> 5; }
inline int thread_id() { return B()->cur; }

inline void yield_to_sched() {
  Block* b = B();
  swapcontext(&b->ctx[b->cur], &b->sched);
}

// All `count` participants contribute (v, aux); afterwards each can see everybody's contribution.
// `me` is the participant index inside the rendezvous (lane for a warp, thread id for the block).
inline void rendezvous(Rendezvous& r, int count, int me, int op, uint64_t v, uint32_t aux, const uint64_t** vals,
                       const uint32_t** auxs) {
  Block* b = B();
  r.slot[me] = v;
  r.aux[me] = aux;
  r.op[me] = op;
  const uint64_t mygen = r.gen;
  if (++r.arrived == count) {
    for (int i = 0; i < count; ++i) {
      if (r.op[i] != op) {
        fprintf(stderr, "warp_emul: participants disagree on the collective (%d: op %d vs %d: op %d)\n", i, r.op[i], me, op);
        abort();
      }
      r.result[mygen & 1][i] = r.slot[i];
      r.result_aux[mygen & 1][i] = r.aux[i];
    }
    r.arrived = 0;
    b->collectives++;
    r.gen++;
  } else {
    while (r.gen == mygen) yield_to_sched();
  }
  *vals = r.result[mygen & 1];
  *auxs = r.result_aux[mygen & 1];
}
inline void rendezvous(int op, uint64_t v, uint32_t aux, const uint64_t** vals, const uint32_t** auxs) {
  Block* b = B();
  rendezvous(b->warp[b->cur >> 5], 32, b->cur & 31, op, v, aux, vals, auxs);
}

// Named CTA barriers (PTX bar.sync / bar.arrive id, count): a phase completes when `count` threads have
// arrived; bar.arrive does not wait, bar.sync waits for the phase it arrived in to complete.
inline void named_bar_arrive(int id, int count) {
  Block* b = B();
  Block::NamedBar& nb = b->nbar[id];
  if (++nb.arrived == count) {
    nb.arrived = 0;
    nb.gen++;
    b->collectives++;
  } else if (nb.arrived > count) {
    fprintf(stderr, "warp_emul: named barrier %d over-subscribed\n", id);
    abort();
  }
}
inline void named_bar_sync(int id, int count) {
  Block* b = B();
  Block::NamedBar& nb = b->nbar[id];
  const uint64_t g = nb.gen;
  if (++nb.arrived == count) {
    nb.arrived = 0;
    nb.gen++;
    b->collectives++;
  } else {
    while (nb.gen == g) yield_to_sched();
  }
}
// mbarrier objects (8 bytes in shared memory): low word = pending arrivals, high word = count << 1 | phase.
inline void mbar_init(uint64_t* m, uint32_t count) { *m = ((uint64_t)((count << 1) | 0u) << 32) | count; }
inline void mbar_arrive(uint64_t* m) {
  uint32_t pending = (uint32_t)*m, hi = (uint32_t)(*m >> 32);
  if (pending == 0 || (hi >> 1) == 0) { fprintf(stderr, "warp_emul: arrive on an uninitialised mbarrier\n"); abort(); }
  if (--pending == 0) {
    hi ^= 1u;             // next phase
    pending = hi >> 1;
    B()->collectives++;
  }
  *m = ((uint64_t)hi << 32) | pending;
}
inline void mbar_wait(uint64_t* m, uint32_t parity) {   // returns once the phase with this parity has completed
  while ((((uint32_t)(*(volatile uint64_t*)m >> 32)) & 1u) == (parity & 1u)) yield_to_sched();
}
inline bool named_bars_idle() {
  Block* b = B();
  for (int i = 0; i < 16; ++i)
    if (b->nbar[i].arrived) return false;
  return true;
}

inline void check_mask(unsigned mask) {
  if (mask != 0xffffffffu) {
    fprintf(stderr, "warp_emul: only full-mask collectives are supported (got %08x)\n", mask);
    abort();
  }
}

inline void fiber_entry() {
  Block* b = B();
  b->body();
  b->done[b->cur] = true;
  swapcontext(&b->ctx[b->cur], &b->sched);
}

// Runs body() once per thread of a block of `nwarps` warps to completion. Returns the number of collectives.
inline uint64_t run_block(int nwarps, const std::function<void()>& body) {
  Block* b = new Block();
  Block* saved = B();
  B() = b;
  b->body = body;
  b->nthreads = 32 * nwarps;
  const size_t kStack = 256 * 1024;
  for (int i = 0; i < b->nthreads; ++i) {
    b->stack[i] = (char*)malloc(kStack);
    b->done[i] = false;
    getcontext(&b->ctx[i]);
    b->ctx[i].uc_stack.ss_sp = b->stack[i];
    b->ctx[i].uc_stack.ss_size = kStack;
    b->ctx[i].uc_link = &b->sched;
    makecontext(&b->ctx[i], (void (*)())fiber_entry, 0);
  }
  int remaining = b->nthreads;
  int idle_sweeps = 0;
  while (remaining > 0) {
    const uint64_t before = b->collectives;
    const int rem_before = remaining;
    for (int i = 0; i < b->nthreads; ++i) {
      if (b->done[i]) continue;
      b->cur = i;
      swapcontext(&b->sched, &b->ctx[i]);
      if (b->done[i]) --remaining;
    }
    if (b->collectives == before && remaining == rem_before && remaining > 0) {
      // nobody completed a collective or finished in a full sweep: threads are parked at collectives
      // that can never complete -> deadlock in real hardware terms.
      if (++idle_sweeps > 2) {
        fprintf(stderr, "warp_emul: deadlock (%d threads alive; block barrier has %d arrivals", remaining, b->block.arrived);
        for (int w = 0; w < b->nthreads / 32; ++w) {
          int alive = 0;
          for (int t = 0; t < 32; ++t) alive += !b->done[w * 32 + t];
          fprintf(stderr, "; warp %d: %d alive, %d at a warp collective (op %d)", w, alive, b->warp[w].arrived, b->warp[w].op[0]);
        }
        fprintf(stderr, ")\n");
        abort();
      }
    } else {
      idle_sweeps = 0;
    }
  }
  const uint64_t n = b->collectives;
  for (int i = 0; i < b->nthreads; ++i) free(b->stack[i]);
  delete b;
  B() = saved;
  return n;
}
inline uint64_t run_warp(const std::function<void()>& body) { return run_block(1, body); }

}  // namespace wemu

// ---------------------------------------------------------------- CUDA intrinsics used by the kernels
template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src) {
  wemu::check_mask(mask);
  static_assert(sizeof(T) <= 8, "shfl value too wide");
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  const uint64_t* vals;
  const uint32_t* auxs;
  wemu::rendezvous(wemu::OP_SHFL, raw, 0, &vals, &auxs);
  T out;
  memcpy(&out, &vals[src & 31], sizeof(T));
  return out;
}
template <typename T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
  wemu::check_mask(mask);
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  const uint64_t* vals;
  const uint32_t* auxs;
  wemu::rendezvous(wemu::OP_SHFL_UP, raw, 0, &vals, &auxs);
  const int me = wemu::lane();
  const int src = me - (int)delta;
  T out;
  memcpy(&out, &vals[src < 0 ? me : src], sizeof(T));
  return out;
}
template <typename T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta) {
  wemu::check_mask(mask);
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  const uint64_t* vals;
  const uint32_t* auxs;
  wemu::rendezvous(wemu::OP_SHFL_DOWN, raw, 0, &vals, &auxs);
  const int me = wemu::lane();
  const int src = me + (int)delta;
  T out;
  memcpy(&out, &vals[src > 31 ? me : src], sizeof(T));
  return out;
}
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int lanemask) {
  wemu::check_mask(mask);
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  const uint64_t* vals;
  const uint32_t* auxs;
  wemu::rendezvous(wemu::OP_SHFL_XOR, raw, 0, &vals, &auxs);
  T out;
  memcpy(&out, &vals[(wemu::lane() ^ lanemask) & 31], sizeof(T));
  return out;
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
  wemu::check_mask(mask);
  const uint64_t* vals;
  const uint32_t* auxs;
  wemu::rendezvous(wemu::OP_BALLOT, pred ? 1 : 0, 0, &vals, &auxs);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= (unsigned)(vals[i] & 1) << i;
  return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }
inline void __syncwarp(unsigned mask = 0xffffffffu) {
  wemu::check_mask(mask);
  const uint64_t* vals;
  const uint32_t* auxs;
  wemu::rendezvous(wemu::OP_SYNC, 0, 0, &vals, &auxs);
}
inline void __syncthreads() {
  const uint64_t* vals;
  const uint32_t* auxs;
  wemu::Block* b = wemu::B();
  wemu::rendezvous(b->block, b->nthreads, b->cur, wemu::OP_BLOCKSYNC, 0, 0, &vals, &auxs);
}
inline unsigned __match_any_sync(unsigned mask, unsigned v) {
  wemu::check_mask(mask);
  const uint64_t* vals;
  const uint32_t* auxs;
  wemu::rendezvous(wemu::OP_MATCH, v, 0, &vals, &auxs);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= (unsigned)(vals[i] == (uint64_t)v) << i;
  return r;
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
  wemu::check_mask(mask);
  const uint64_t* vals;
  const uint32_t* auxs;
  wemu::rendezvous(wemu::OP_REDUCE, v, 1, &vals, &auxs);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r += (unsigned)vals[i];
  return r;
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
  wemu::check_mask(mask);
  const uint64_t* vals;
  const uint32_t* auxs;
  wemu::rendezvous(wemu::OP_REDUCE, v, 2, &vals, &auxs);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r = (unsigned)vals[i] > r ? (unsigned)vals[i] : r;
  return r;
}
inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
  wemu::check_mask(mask);
  const uint64_t* vals;
  const uint32_t* auxs;
  wemu::rendezvous(wemu::OP_REDUCE, v, 3, &vals, &auxs);
  unsigned r = 0xffffffffu;
  for (int i = 0; i < 32; ++i) r = (unsigned)vals[i] < r ? (unsigned)vals[i] : r;
  return r;
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
  wemu::check_mask(mask);
  const uint64_t* vals;
  const uint32_t* auxs;
  wemu::rendezvous(wemu::OP_REDUCE, v, 4, &vals, &auxs);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= (unsigned)vals[i];
  return r;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
  const uint64_t w = ((uint64_t)hi << 32) | lo;
  return (unsigned)(w >> (shift & 31));
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift) {
  const uint64_t w = ((uint64_t)hi << 32) | lo;
  return (unsigned)((w << (shift & 31)) >> 32);
}
template <typename T>
inline T atomicAdd(T* p, T v) { T old = *p; *p = old + v; return old; }
template <typename T>
inline T atomicOr(T* p, T v) { T old = *p; *p = old | v; return old; }
template <typename T>
inline T atomicMax(T* p, T v) { T old = *p; if (v > old) *p = v; return old; }
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
