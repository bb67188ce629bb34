// tcgen05 / TMEM version of the policy step for the common model shape (hidden 128, GRUCell): the reference's
// RNNAgent forward (utilities/rnn_agent.py:21-29) + greedy argmax (utilities/misc.py:170) for 128 envs x one agent
// index per CTA.  Same arithmetic contract as policy_act_kernel (FP16 operands, FP32 accumulation, FP32 gates).
//
//   phase 0  warp 0 allocates all 512 TMEM columns; mbarriers initialised; one thread starts the bulk copies
//            (cp.async.bulk, 1-D, no tensor map needed: the weights were laid out on the host in the UMMA
//            canonical K-major / no-swizzle layout) of fc1 / fc2 / biases and of the first three 32 KB GRU slabs
//   phase 1  all threads stage obs (+ one-hot agent id) and the hidden state as FP16 in the canonical layout
//   phase 2  fc1:  D[128 x 128] = obs * W1'        -> TMEM cols [0,128)     (one thread issues tcgen05.mma)
//   phase 3  x = relu(D + b1) -> FP16 -> shared (canonical), thread t owns row t (TMEM lane t)
//   phase 4  GRU: six slabs W_ir, W_hr, W_iz, W_hz, W_in, W_hn stream through a 3-deep ring of 32 KB buffers;
//            r -> cols [0,128), z -> [128,256), W_in x -> [256,384), W_hn h -> [384,512); 48 MMAs of 128x128x16
//   phase 5  epilogue: thread t reads its row of the four accumulators in 16-column chunks (tcgen05.ld 32x32b.x16),
//            gates on the SFU, h' -> global (FP32) and -> shared (FP16, canonical) for fc2
//   phase 6  fc2:  q[128 x 32] = h' * W2'           -> TMEM cols [0,32);  argmax per row, first maximum
//
// Canonical K-major layout, no swizzle (CUTLASS cute/atom/mma_traits_sm100.hpp "LayoutType::INTERLEAVE:
// ((8,n),2):((1,SBO),LBO)" in 16-byte units): element (row r, k) of an operand with R rows lives at
//     (k / 8) * LBO + (r / 8) * SBO + (r % 8) * 16 + (k % 8) * 2   bytes,   SBO = 128,
// i.e. 8 rows x 8 halves form one contiguous 128-byte core matrix.  Weights use LBO = R / 8 * 128 (dense);
// the activation buffers use LBO = 2064 so that the row-wise staging stores do not all hit one bank group.
// Every wait on an mbarrier is bounded and traps instead of hanging the GPU.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace mrb {
namespace tc {

constexpr int kRows = 128, kH = 128, kThreads = 128;
constexpr int kSlabBytes = kH * kH * 2;                    // one 128 x 128 FP16 gate slab (32 KB)
constexpr int kRing = 3;
constexpr int kNpad2 = 32;                                 // fc2 output width padded to one 32-column MMA
constexpr int kActLBO = 2048 + 16;                         // activation buffers: padded k-chunk stride
constexpr int kActBytes = 16 * kActLBO;                    // 128 rows x 128 k
constexpr int kMaxKp1 = 64;                                // fc1 input width (padded to 16)
constexpr int kBiasFloats = kH + 3 * kH + 3 * kH + 32;     // b1 | b_ih | b_hh | b2

struct Params {
    const uint8_t *img;       // packed weight sets (device)
    int64_t set_bytes;
    int64_t B;
    int32_t obs_dim, input_dim, n_actions, n_agents, obs_agent_id, non_shared;
    int32_t Kp1;              // fc1 k extent, multiple of 16
    int32_t w1_bytes, w2_bytes, head_bytes;   // head = W1 | W2 | biases
};

// ---- shared-memory map (bytes from the aligned base).  obs (fc1's A operand) aliases the x buffer: x is only
// written after fc1 has completed.
struct Smem {
    static constexpr int x = 0;
    static constexpr int obs = x;
    static constexpr int h = x + kActBytes;
    static constexpr int ring = h + kActBytes;
    static constexpr int bars = ring + kRing * kSlabBytes;          // mbarriers: head, full[3], empty[3], mma_done
    static constexpr int tmem_slot = bars + 8 * 8;
    static constexpr int head = tmem_slot + 64;                     // W1 | W2 | biases
    static constexpr int head_max = kMaxKp1 / 8 * (kH / 8 * 128) + kH / 8 * (kNpad2 / 8 * 128) + kBiasFloats * 4;
    static constexpr int total = head + head_max + 1024;            // + slack for the 1024-byte alignment of the base
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 22)) __trap();      // never hang the device on a protocol bug
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// shared-memory matrix descriptor: start address, LBO, SBO in 16-byte units, version 1 (sm_100), no swizzle
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}
// instruction descriptor, kind::f16: D = F32, A = B = F16, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr));
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t h2(float lo, float hi)
{
    const __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&v);
}
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigm(float x) { return rcpa(1.f + ex2a(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_(float x) { return fmaf(2.f, rcpa(1.f + ex2a(-2.8853900817779268f * x)), -1.f); }

__global__ void __launch_bounds__(kThreads, 1)
policy_act_tc_kernel(const Params p, const float *obs, float *hidden, int32_t *__restrict__ actions,
                     float *__restrict__ q_out, const uint8_t *__restrict__ fresh)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *sm = smem_raw;
    const uint32_t sbase = smem_u32(sm);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int agent = blockIdx.y;
    const uint8_t *img = p.img + (size_t)(p.non_shared ? agent : 0) * p.set_bytes;
    const uint32_t bar_head = sbase + Smem::bars, bar_full = bar_head + 8, bar_empty = bar_full + 8 * kRing,
                   bar_done = bar_empty + 8 * kRing;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + Smem::tmem_slot);

    // ---- phase 0
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(bar_head, 1);
        for (int i = 0; i < kRing; i++) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (tid == 32) {
        mbar_expect_tx(bar_head, (uint32_t)p.head_bytes);
        bulk_g2s(sbase + Smem::head, img, (uint32_t)p.head_bytes, bar_head);
        for (int s = 0; s < kRing; s++) {
            mbar_expect_tx(bar_full + 8 * s, kSlabBytes);
            bulk_g2s(sbase + Smem::ring + s * kSlabBytes, img + p.head_bytes + (size_t)s * kSlabBytes, kSlabBytes, bar_full + 8 * s);
        }
    }

    // ---- phase 1: stage obs and hidden (FP16, canonical layout); warp w takes rows w, w + 4, ...
    const int N = p.n_agents, D = p.obs_dim;
    const int64_t e0 = (int64_t)blockIdx.x * kRows;
    // eight rows per pass: all global loads of a pass are issued before the first shared store (latency overlap)
#pragma unroll 1
    for (int r0 = warp * 32; r0 < warp * 32 + 32; r0 += 8) {
        float4 hv[8];
        float ov[8][2];
        bool zr[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int64_t e = e0 + r0 + i;
            const bool valid = e < p.B;
            zr[i] = !valid || (fresh && fresh[e]);
            const int64_t row = (valid ? e : 0) * N + agent;
            hv[i] = *reinterpret_cast<const float4 *>(hidden + row * kH + 4 * lane);
            const int k2 = 2 * lane;                       // obs (+ one-hot agent id, misc.py:161-162): lane l holds k = 2l, 2l+1
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int c = k2 + j;
                ov[i][j] = (k2 < p.Kp1) ? obs[row * D + (c < D ? c : 0)] : 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int r = r0 + i;
            const uint32_t rbase = (uint32_t)((r >> 3) * 128 + (r & 7) * 16);
            if (zr[i]) hv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int k = 4 * lane;                        // hidden: lane l holds k = 4l .. 4l+3
            *reinterpret_cast<uint2 *>(sm + Smem::h + (k >> 3) * kActLBO + rbase + (k & 7) * 2) =
                make_uint2(h2(hv[i].x, hv[i].y), h2(hv[i].z, hv[i].w));
            const int k2 = 2 * lane;
            if (k2 < p.Kp1) {
                float v[2];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int c = k2 + j;
                    v[j] = c < D ? (zr[i] ? 0.f : ov[i][j]) : ((p.obs_agent_id && c - D == agent) ? 1.f : 0.f);
                }
                *reinterpret_cast<uint32_t *>(sm + Smem::obs + (k2 >> 3) * kActLBO + rbase + (k2 & 7) * 2) = h2(v[0], v[1]);
            }
        }
    }
    fence_async_smem();                                    // generic-proxy stores -> visible to the tensor core's async proxy
    __syncthreads();

    const float *bias = reinterpret_cast<const float *>(sm + Smem::head + p.w1_bytes + p.w2_bytes);
    const uint32_t w1_addr = sbase + Smem::head, w2_addr = w1_addr + p.w1_bytes;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32 TMEM lanes
    const int row = tid;                                                // thread t <-> tile row t <-> TMEM lane t
    const uint32_t my_rbase = (uint32_t)((row >> 3) * 128 + (row & 7) * 16);

    // ---- phase 2: fc1 on the tensor core
    if (tid == 0) {
        mbar_wait(bar_head, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc(kH);
        for (int ks = 0; ks < p.Kp1 / 16; ks++) {
            const uint64_t a = make_desc(sbase + Smem::obs + ks * 2 * kActLBO, kActLBO, 128);
            const uint64_t b = make_desc(w1_addr + ks * 2 * (kH / 8 * 128), kH / 8 * 128, 128);
            umma(tmem, a, b, idesc, ks > 0);
        }
        tc_commit(bar_done);
    }
    mbar_wait(bar_done, 0);
    mbar_wait(bar_head, 0);                                // biases are read below by every thread
    __syncwarp();                                          // tcgen05.ld is .sync.aligned: reconverge after the single-thread issue
    tc_fence_after();

    // ---- phase 3: x = relu(fc1 + b1) -> shared, FP16                              rnn_agent.py:22
#pragma unroll 1
    for (int c = 0; c < kH / 16; c++) {
        float v[16];
        tmem_ld16(lane_base + 16 * c, v);
        tmem_ld_wait();
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; i++)
            w[i] = h2(fmaxf(v[2 * i] + bias[16 * c + 2 * i], 0.f), fmaxf(v[2 * i + 1] + bias[16 * c + 2 * i + 1], 0.f));
        *reinterpret_cast<uint4 *>(sm + Smem::x + (2 * c) * kActLBO + my_rbase) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4 *>(sm + Smem::x + (2 * c + 1) * kActLBO + my_rbase) = make_uint4(w[4], w[5], w[6], w[7]);
    }
    tc_fence_before();
    fence_async_smem();
    __syncthreads();

    // ---- phase 4: GRU slabs (rnn_agent.py:24-25, torch.nn.GRUCell gate order r, z, n)
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc(kH);
        for (int s = 0; s < 6; s++) {
            const int buf = s % kRing;
            mbar_wait(bar_full + 8 * buf, (s / kRing) & 1);
            tc_fence_after();
            const uint32_t a_base = sbase + ((s & 1) ? Smem::h : Smem::x);          // even slabs multiply x, odd slabs h
            const uint32_t d_col = s < 4 ? (uint32_t)(s >> 1) * kH : (uint32_t)(s - 2) * kH;   // r, r, z, z, in, hn
            const uint32_t acc0 = (s == 1 || s == 3) ? 1u : 0u;                      // h-side of r and z accumulates
            for (int ks = 0; ks < kH / 16; ks++) {
                const uint64_t a = make_desc(a_base + ks * 2 * kActLBO, kActLBO, 128);
                const uint64_t b = make_desc(sbase + Smem::ring + buf * kSlabBytes + ks * 2 * (kH / 8 * 128), kH / 8 * 128, 128);
                umma(tmem + d_col, a, b, idesc, acc0 | (ks > 0));
            }
            tc_commit(bar_empty + 8 * buf);                 // arrives when the MMAs that read this buffer are done
        }
        tc_commit(bar_done);
    } else if (tid == 32) {
        for (int s = kRing; s < 6; s++) {                   // refill the ring behind the tensor core
            const int buf = s % kRing;
            mbar_wait(bar_empty + 8 * buf, 0);
            mbar_expect_tx(bar_full + 8 * buf, kSlabBytes);
            bulk_g2s(sbase + Smem::ring + buf * kSlabBytes, img + p.head_bytes + (size_t)s * kSlabBytes, kSlabBytes, bar_full + 8 * buf);
        }
    }
    mbar_wait(bar_done, 1);
    __syncwarp();
    tc_fence_after();

    // ---- phase 5: gates, h' -> global and -> shared (FP16) for fc2
    const int64_t e = e0 + row;
    const bool valid = e < p.B;
    const bool zero = !valid || (fresh && fresh[e]);
    const int64_t grow = (valid ? e : 0) * N + agent;
    float *hrow = hidden + grow * kH;
    const float *b_ih = bias + kH, *b_hh = bias + 4 * kH;
    float4 nxt[4];
#pragma unroll
    for (int i = 0; i < 4; i++) nxt[i] = *reinterpret_cast<const float4 *>(hrow + 4 * i);
#pragma unroll 1
    for (int c = 0; c < kH / 16; c++) {
        float ar[16], az[16], ai[16], ah[16];
        tmem_ld16(lane_base + 16 * c, ar);
        tmem_ld16(lane_base + kH + 16 * c, az);
        tmem_ld16(lane_base + 2 * kH + 16 * c, ai);
        tmem_ld16(lane_base + 3 * kH + 16 * c, ah);
        float old[16];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            old[4 * i] = nxt[i].x; old[4 * i + 1] = nxt[i].y; old[4 * i + 2] = nxt[i].z; old[4 * i + 3] = nxt[i].w;
        }
        if (c + 1 < kH / 16) {                              // old h of the next chunk is in flight while this one computes
#pragma unroll
            for (int i = 0; i < 4; i++) nxt[i] = *reinterpret_cast<const float4 *>(hrow + 16 * (c + 1) + 4 * i);
        }
        tmem_ld_wait();
        float hn[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int k = 16 * c + i;
            const float r = sigm(ar[i] + b_ih[k] + b_hh[k]);
            const float z = sigm(az[i] + b_ih[kH + k] + b_hh[kH + k]);
            const float n = tanh_(ai[i] + b_ih[2 * kH + k] + r * (ah[i] + b_hh[2 * kH + k]));
            hn[i] = (1.f - z) * n + z * (zero ? 0.f : old[i]);
        }
        if (valid) {
#pragma unroll
            for (int i = 0; i < 4; i++)
                *reinterpret_cast<float4 *>(hrow + 16 * c + 4 * i) = make_float4(hn[4 * i], hn[4 * i + 1], hn[4 * i + 2], hn[4 * i + 3]);
        }
        *reinterpret_cast<uint4 *>(sm + Smem::x + (2 * c) * kActLBO + my_rbase) =
            make_uint4(h2(hn[0], hn[1]), h2(hn[2], hn[3]), h2(hn[4], hn[5]), h2(hn[6], hn[7]));
        *reinterpret_cast<uint4 *>(sm + Smem::x + (2 * c + 1) * kActLBO + my_rbase) =
            make_uint4(h2(hn[8], hn[9]), h2(hn[10], hn[11]), h2(hn[12], hn[13]), h2(hn[14], hn[15]));
    }
    tc_fence_before();
    fence_async_smem();
    __syncthreads();

    // ---- phase 6: q = fc2(h')                                                      rnn_agent.py:28
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc(kNpad2);
        for (int ks = 0; ks < kH / 16; ks++) {
            const uint64_t a = make_desc(sbase + Smem::x + ks * 2 * kActLBO, kActLBO, 128);
            const uint64_t b = make_desc(w2_addr + ks * 2 * (kNpad2 / 8 * 128), kNpad2 / 8 * 128, 128);
            umma(tmem, a, b, idesc, ks > 0);
        }
        tc_commit(bar_head);                                // reuse: second phase of the head barrier
    }
    mbar_wait(bar_head, 1);
    __syncwarp();
    tc_fence_after();
    {
        const float *b2 = bias + 7 * kH;
        float best = -INFINITY;
        int idx = 0;
#pragma unroll
        for (int c = 0; c < kNpad2 / 16; c++) {
            float v[16];
            tmem_ld16(lane_base + 16 * c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int a = 16 * c + i;
                if (a < p.n_actions) {
                    const float qv = v[i] + b2[a];
                    if (qv > best) { best = qv; idx = a; }          // first maximum, like np.argmax (misc.py:170)
                    if (q_out && valid) q_out[grow * p.n_actions + a] = qv;
                }
            }
        }
        if (valid) actions[grow] = idx;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ---- host: FP16 canonical images.  W is torch layout [out][in]; element (row r = output unit, k = input unit).
inline void pack_canonical(uint8_t *dst, const float *W, int ld, int row0, int rows_valid, int cols_valid, int R, int K)
{
    const int lbo = R / 8 * 128;
    for (int k = 0; k < K; k++)
        for (int r = 0; r < R; r++) {
            const float v = (r < rows_valid && k < cols_valid) ? W[(size_t)(row0 + r) * ld + k] : 0.f;
            const uint16_t hbits = __half_as_ushort(__float2half_rn(v));
            std::memcpy(dst + (size_t)(k / 8) * lbo + (r / 8) * 128 + (r % 8) * 16 + (k % 8) * 2, &hbits, 2);
        }
}

}  // namespace tc
}  // namespace mrb
