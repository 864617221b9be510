// cuda_emu.h -- just enough of the CUDA execution model to run the gather kernels of csrc/gradient.cu on the host.
// TEST INFRASTRUCTURE ONLY (tests/test_kernel_emulation.py): the kernels' source text is extracted verbatim from the .cu
// files and compiled with g++ against this header, so that kernel variants written without GPU access are executed --
// thread indexing, shared memory, warp shuffles, barriers, the ticketed grid reduction -- before GPU time is spent.
// One CTA runs at a time, each of its threads is an OS thread; __shared__ becomes a function-local static (one CTA at a
// time, so that is per-CTA storage); warp shuffles and barriers are std::barrier rendezvous.  Loads with cache hints are
// plain loads.  Nothing here is fast and nothing here ships.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

typedef long long i64;
struct double2 { double x, y; };
struct uint3_emu { unsigned x = 0, y = 0, z = 0; };

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)

namespace emu {
struct Warp {
    std::barrier<> bar{32};
    unsigned long long slot[32];
};
struct Cta {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<std::unique_ptr<Warp>> warps;
};
inline thread_local uint3_emu t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local Warp *t_warp = nullptr;
inline thread_local Cta *t_cta = nullptr;

template <typename T>
inline T exchange(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle operand");
    unsigned long long bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    const int lane = t_threadIdx.x & 31;
    t_warp->slot[lane] = bits;
    t_warp->bar.arrive_and_wait();
    const unsigned long long got = t_warp->slot[src & 31];
    t_warp->bar.arrive_and_wait();
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}

// run kernel(args...) on a grid x block launch; CTAs one after the other
template <typename K, typename... A>
void launch(K kernel, int grid, int block, A... args) {
    for (int b = 0; b < grid; b++) {
        Cta cta;
        cta.bar = std::make_unique<std::barrier<>>(block);
        for (int w = 0; w < (block + 31) / 32; w++) cta.warps.push_back(std::make_unique<Warp>());
        std::vector<std::thread> th;
        for (int t = 0; t < block; t++) {
            th.emplace_back([&, t]() {
                t_threadIdx.x = (unsigned)t; t_blockIdx.x = (unsigned)b; t_blockDim.x = (unsigned)block; t_gridDim.x = (unsigned)grid;
                t_cta = &cta; t_warp = cta.warps[t / 32].get();
                kernel(args...);
                t_warp->bar.arrive_and_drop();   // an exited thread no longer takes part in barriers
                cta.bar->arrive_and_drop();
            });
        }
        for (auto &x : th) x.join();
    }
}
}  // namespace emu

#define threadIdx emu::t_threadIdx
#define blockIdx emu::t_blockIdx
#define blockDim emu::t_blockDim
#define gridDim emu::t_gridDim

template <typename T> inline T __ldg(const T *p) { return *p; }
template <typename T> inline T __ldcg(const T *p) { return *p; }
inline void __syncthreads() { emu::t_cta->bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::t_warp->bar.arrive_and_wait(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return emu::exchange(v, src); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int mask) { return emu::exchange(v, (int)(emu::t_threadIdx.x & 31) ^ mask); }
inline unsigned atomicAdd(unsigned *p, unsigned v) {
    return std::atomic_ref<unsigned>(*p).fetch_add(v);
}

// the inline-PTX helpers at the top of gradient.cu (L2 policies, hinted loads): plain loads here
inline unsigned long long pol_evict_last() { return 1ull; }
inline unsigned long long pol_evict_first() { return 2ull; }
inline int ldg_i32_hint(const int *p, unsigned long long) { return *p; }
inline double ldg_f64_hint(const double *p, unsigned long long) { return *p; }
inline double2 ldg_f64x2_hint(const double *p, unsigned long long) { return *reinterpret_cast<const double2 *>(p); }
