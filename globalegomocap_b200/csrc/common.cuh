// Shared helpers for the gem_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/gem_b200.h"

namespace gem {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define GEM_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t _e = (call);                                                    \
        if (_e != cudaSuccess) return gem::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define GEM_CHECK_LAUNCH() GEM_CUDA(cudaGetLastError())

#define GEM_REQUIRE(cond, msg)                       \
    do {                                             \
        if (!(cond)) {                               \
            gem::set_error(std::string("invalid argument: ") + (msg)); \
            return GEM_ERR_INVALID;                  \
        }                                            \
    } while (0)

// cudaFuncSetAttribute applies to the current device only: one flag per device and call site
struct PerDeviceOnce {
    bool done[64] = {};
    bool* flag() {
        int d = 0;
        cudaGetDevice(&d);
        return &done[d & 63];
    }
};

constexpr int kMaxJoints = 32;
constexpr int kMaxPoly = 24;
constexpr int kNumSMs = 148;

struct CameraConst {
    float poly[kMaxPoly];
    int n_poly;
    float cx, cy;
};

struct SkeletonConst {
    int parent[kMaxJoints];
    int num_joints;
    int child_start[kMaxJoints + 1];   // CSR of each joint's children (joints whose parent it is, itself excluded)
    int child_list[kMaxJoints];
};

// ---- warp / block reductions -------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- mbarrier / bulk-copy (TMA engine) PTX -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU.  try_wait itself may
// suspend the thread for a while, so the bound is wall-clock time (4 s), checked every 256 polls.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    unsigned long long t0 = 0;
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
        if ((spin & 255u) == 255u) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000ull) __trap();
        }
    }
}
// global -> shared bulk copy (cp.async.bulk, SASS UBLKCP); 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// the same with an L2 eviction-priority hint (createpolicy): evict_last keeps rows that will be read again soon
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// shared -> global bulk copy
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// ---- per-joint texel window of the energy kernel's cache (energy_device.cuh) ------------------------------------------
constexpr int kPatchWd = 8;            // side of the per-joint texel window of the HWC cache (at most 8: 64 valid bits)
#ifndef GEM_PLANAR_W
#define GEM_PLANAR_W 16
#endif
#ifndef GEM_PLANAR_H
#define GEM_PLANAR_H 16
#endif
#ifndef GEM_PLANAR_ALIGN
#define GEM_PLANAR_ALIGN 8
#endif
constexpr int kPlanarW = GEM_PLANAR_W, kPlanarH = GEM_PLANAR_H, kPlanarAlign = GEM_PLANAR_ALIGN;
constexpr int kPatchFloats = kPlanarW * kPlanarH > kPatchWd * kPatchWd ? kPlanarW * kPlanarH : kPatchWd * kPatchWd;
static_assert(kPlanarW % kPlanarAlign == 0 && kPlanarW >= 2 * kPlanarAlign && (kPlanarAlign == 4 || kPlanarAlign == 8 || kPlanarAlign == 16),
              "window width / alignment");
static_assert(kPlanarH <= 64 && kPlanarW <= 32, "one valid bit per window row; a row is fetched by at most 8 lanes");
constexpr int kTileW = 8, kTileH = 4, kTileFloats = kTileW * kTileH;      // tiled maps (layout 2): one 128-byte line per tile
static_assert(kPlanarW == 2 * kTileW && kPlanarH == 4 * kTileH, "the tiled window is 2 x 4 tiles");

}  // namespace gem
