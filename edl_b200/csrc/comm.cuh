// Device-side primitives for collectives over NVSwitch peer memory (sm_100a).
//
// Every rank maps every peer's symmetric buffer + signal pad (CUDA VMM / IPC handles exchanged
// through the rendezvous store at each elastic *stage*).  Kernels synchronise with per-block flag
// words written with st.release.sys and polled with ld.acquire.sys; flags carry a monotonically
// increasing epoch kept in device memory so that CUDA-graph replays need no host-side state.
// A rank that disappears mid-collective makes its peers spin: the spin is bounded by a
// globaltimer timeout that raises a device-visible error word instead of hanging the GPU
// (SURVEY 7.3 hard part #2) -- the host turns that into the elastic re-rendezvous path.
#pragma once
#include "common.cuh"

namespace edl {

constexpr int kMaxWorld = 16;
constexpr int kMaxCommBlocks = 64;
// signal pad layout in uint32 words
constexpr int kSigStart = 0;                                   // [blocks][world]
constexpr int kSigEnd = kMaxCommBlocks * kMaxWorld;            // [blocks][world]
constexpr int kSigEpoch = 2 * kMaxCommBlocks * kMaxWorld;      // [blocks]   (local use only)
constexpr int kSigScratch = kSigEpoch + kMaxCommBlocks;        // [world][8] floats: per-rank partials
constexpr int kSigError = kSigScratch + kMaxWorld * 8;         // [1] error word (local)
constexpr int kSigWords = kSigError + 8;
// how long a barrier still waits once this rank's error word is set (a peer already timed out)
constexpr unsigned long long kCommBrokenGraceNs = 200ull * 1000ull;

struct CommCtx {
  void* data[kMaxWorld];      // peer pointers to the symmetric payload buffer
  uint32_t* sig[kMaxWorld];   // peer pointers to the signal pads
  void* mc_data;              // NVLS multicast alias of the payload buffer (or nullptr)
  int rank;
  int world;
  unsigned long long timeout_ns;
};

EDL_DEVICE void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
EDL_DEVICE uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
EDL_DEVICE unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

EDL_DEVICE uint32_t comm_epoch_begin(const CommCtx& c) {
  // every thread reads the same word; the value is written back by comm_epoch_end
  return c.sig[c.rank][kSigEpoch + blockIdx.x] + 1;
}
EDL_DEVICE void comm_epoch_end(const CommCtx& c, uint32_t epoch) {
  if (threadIdx.x == 0) c.sig[c.rank][kSigEpoch + blockIdx.x] = epoch;
}

// Cross-rank barrier for block `blockIdx.x` of every rank.  `slot` is kSigStart or kSigEnd.
// With release=true all global writes of this block made before the call are visible to any peer
// that observes the flag.
template <bool kRelease>
EDL_DEVICE void comm_barrier(const CommCtx& c, int slot, uint32_t epoch) {
  if (kRelease) {
    __threadfence_system();
    __syncthreads();
  }
  if ((int)threadIdx.x < c.world) {
    const int peer = threadIdx.x;
    st_release_sys(c.sig[peer] + slot + blockIdx.x * kMaxWorld + c.rank, epoch);
    const uint32_t* mine = c.sig[c.rank] + slot + blockIdx.x * kMaxWorld + peer;
    // The error word is sticky for the lifetime of the pool (one elastic stage).  Once ANY barrier of this rank
    // has timed out, a peer is gone: every later barrier -- the second one of the same kernel, the other CTAs, the
    // next buckets, the next graph replays -- gives up after a short grace period instead of the full timeout, so
    // a dead peer costs one timeout in total and the host sees the error at its next poll.
    volatile uint32_t* errw = c.sig[c.rank] + kSigError;
    const unsigned long long t0 = globaltimer_ns();
    unsigned long long limit = (*errw != 0u) ? kCommBrokenGraceNs : c.timeout_ns;
    uint32_t spins = 0;
    while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
      if ((++spins & 63u) == 0u && limit != 0 && *errw != 0u && limit > kCommBrokenGraceNs) limit = kCommBrokenGraceNs;
      if (limit != 0 && globaltimer_ns() - t0 > limit) {
        atomicCAS(c.sig[c.rank] + kSigError, 0u, 1u + (uint32_t)peer);   // keep the first culprit
        break;
      }
    }
  }
  __syncthreads();
}

// ---- NVLS (multimem) helpers: in-switch reduction / broadcast over the multicast alias ----
EDL_DEVICE bf16x8 multimem_ld_reduce_bf16(const void* mc_ptr) {
  int4 r;
  asm volatile(
      "multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
      : "l"(mc_ptr)
      : "memory");
  return *reinterpret_cast<bf16x8*>(&r);
}
EDL_DEVICE void multimem_st_bf16(void* mc_ptr, const bf16x8& v) {
  const int4& r = *reinterpret_cast<const int4*>(&v);
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_ptr),
               "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w)
               : "memory");
}
EDL_DEVICE float4 multimem_ld_reduce_f32(const void* mc_ptr) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc_ptr)
               : "memory");
  return r;
}
EDL_DEVICE void multimem_st_f32(void* mc_ptr, const float4& r) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_ptr),
               "f"(r.x), "f"(r.y), "f"(r.z), "f"(r.w)
               : "memory");
}

// plain (non-.nc) 128-bit peer load: peer data changes between launches
EDL_DEVICE int4 ld_peer(const void* p) {
  int4 r;
  asm volatile("ld.global.relaxed.sys.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
EDL_DEVICE void st_peer(void* p, const int4& r) {
  asm volatile("st.global.relaxed.sys.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(r.x), "r"(r.y),
               "r"(r.z), "r"(r.w)
               : "memory");
}

}  // namespace edl
