// Storage policy shared by the one-env-per-thread QP solvers (qp_dual.cuh, qp_thread.cuh): a small per-env vector is
// either a register array or a column of a shared-memory store interleaved over the threads of the CTA (element c of
// thread t at base[c * TPB + t]: conflict-free).  Which vectors go where is a compile-time mask chosen from
// measurements: registers cap the number of resident warps, and these kernels are bound by
// dependent-issue latency, i.e. by how many warps an SM sub-partition can switch between.
#pragma once
#include "common.cuh"

namespace mrb {

template <int LEN, int TPB, bool SHARED>
struct MVec;
template <int LEN, int TPB>
struct MVec<LEN, TPB, false> {
    double r[LEN];
    __device__ __forceinline__ MVec(double *, int) {}
    __device__ __forceinline__ double get(int c) const { return r[c]; }
    __device__ __forceinline__ void set(int c, double v) { r[c] = v; }
};
template <int LEN, int TPB>
struct MVec<LEN, TPB, true> {
    double *base;
    // offset: position of this vector in the thread's column, in doubles
    __device__ __forceinline__ MVec(double *vs, int offset) : base(vs + (size_t)offset * TPB) {}
    // volatile: without it the compiler merges the repeated loads of one iteration into a single early load and
    // keeps the value live (then spills it to local memory) -- the opposite of what this store is for
    __device__ __forceinline__ double get(int c) const { return reinterpret_cast<const volatile double *>(base)[c * TPB]; }
    __device__ __forceinline__ void set(int c, double v) { reinterpret_cast<volatile double *>(base)[c * TPB] = v; }
};
__device__ __forceinline__ void lds_block(uint32_t addr, uint32_t stride, double (&o)[4])
{
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "r"(addr));
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(o[2]), "=d"(o[3]) : "r"(addr + stride));
}
__device__ __forceinline__ void sts_block(uint32_t addr, uint32_t stride, const double (&o)[4])
{
    asm volatile("st.volatile.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(o[0]), "d"(o[1]));
    asm volatile("st.volatile.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr + stride), "d"(o[2]), "d"(o[3]));
}

}  // namespace mrb
