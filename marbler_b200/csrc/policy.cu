// On-device policy step: the reference evaluates its trained agents on the host,
// once per env step, inside utilities/misc.py:134-221 run_env:
//     q_values, hs = model(torch.Tensor(obs), torch.Tensor(hs));  actions = np.argmax(q_values, axis=1)
// with model = utilities/rnn_agent.py:5-29 RNNAgent (fc1 -> ReLU -> GRUCell | Linear+ReLU -> fc2) or
// utilities/rnn_ns_agent.py:5-36 RNNNSAgent (one RNNAgent per agent).  Here ONE kernel does that for every
// agent of every env and writes the int32 actions buffer that mrb_step consumes.
//
// Mapping: a warp owns 16 rows = one agent index in 16 consecutive envs (rows of one agent share a weight
// set, so non-shared agents cost nothing extra).  All three layers are TF32 m16n8k8 tensor-core MMAs with
// FP32 accumulation.  The activations never leave registers: fc1's accumulator fragments are turned into
// the next layer's A fragments with 8 warp shuffles per 16x8 tile, the hidden state is loaded once as A
// fragments, and fc2 is accumulated tile by tile as the GRU produces h'.  Weights are rounded to TF32 and
// laid out in fragment order ONCE on the host (mrb_policy_create), so staging them is a linear cp.async copy
// and every B fragment is a single conflict-free 64/128-bit shared load; the GRU weights (6 H^2 floats,
// 393 KB for H = 128) stream through a double-buffered 2 x 24.6 KB window, one 8-unit column tile at a time.
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"

namespace mrb {

constexpr int kPolicyWarps = 4;          // 64 envs of one agent per CTA
constexpr int kMaxKS1 = 8;               // fc1 input width <= 64
constexpr int kMaxAT = 3;                // n_actions <= 24

struct PolicyParams {
    const float *wpack;                  // packed weight sets (device)
    int64_t set_floats;                  // floats per set
    int64_t B;
    int32_t obs_dim, input_dim, n_actions, n_agents, obs_agent_id, non_shared;
    int32_t KS1, AT;                     // fc1 k-steps, fc2 column tiles
    int32_t szW1, szW2, head_floats;     // packed section sizes (floats)
};

__device__ __forceinline__ uint32_t to_tf32(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// D = A (16x8, row) * B (8x8, col) + D, TF32 inputs, FP32 accumulate
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], float b0, float b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
// accumulator fragment of a 16x8 tile (rows g, g+8; columns 2t, 2t+1) -> A fragment of the same tile
// (rows g, g+8; columns t, t+4): column c of row g lives in lane 4g + c/2, element c & 1
__device__ __forceinline__ void c_to_a(const float (&c)[4], uint32_t (&a)[4], int lane)
{
    const unsigned full = 0xffffffffu;
    const int t = lane & 3, lo = (lane & ~3) | (t >> 1), hi = lo + 2;
    const float v00 = __shfl_sync(full, c[0], lo), v01 = __shfl_sync(full, c[1], lo);
    const float v10 = __shfl_sync(full, c[2], lo), v11 = __shfl_sync(full, c[3], lo);
    const float w00 = __shfl_sync(full, c[0], hi), w01 = __shfl_sync(full, c[1], hi);
    const float w10 = __shfl_sync(full, c[2], hi), w11 = __shfl_sync(full, c[3], hi);
    const bool odd = t & 1;
    a[0] = to_tf32(odd ? v01 : v00);
    a[1] = to_tf32(odd ? v11 : v10);
    a[2] = to_tf32(odd ? w01 : w00);
    a[3] = to_tf32(odd ? w11 : w10);
}
__device__ __forceinline__ void cp_async16(float *dst_smem, const float *src)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <int H, bool RNN>
__global__ void __launch_bounds__(kPolicyWarps * 32, 2)
policy_act_kernel(const PolicyParams p, const float *obs, float *hidden, int32_t *__restrict__ actions,
                  float *__restrict__ q_out, const uint8_t *__restrict__ fresh)
{
    extern __shared__ __align__(16) float sm[];
    constexpr int NT = H / 8, KS = H / 8, G = RNN ? 6 : 1;
    constexpr int CHUNK = G * KS * 64;                       // floats per 8-unit column tile of the recurrent layer
    constexpr int NB = RNN ? 6 * H : H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int agent = blockIdx.y;
    const float *wimg = p.wpack + (size_t)(p.non_shared ? agent : 0) * p.set_floats;
    float *sW1 = sm, *sB1 = sW1 + p.szW1, *sBias = sB1 + H, *sW2 = sBias + NB, *sB2 = sW2 + p.szW2;
    float *sChunk = sm + p.head_floats;
    const float *gchunks = wimg + p.head_floats;

    for (int i = threadIdx.x * 4; i < p.head_floats; i += kPolicyWarps * 128) cp_async16(sm + i, wimg + i);
    cp_commit();
    for (int i = threadIdx.x * 4; i < CHUNK; i += kPolicyWarps * 128) cp_async16(sChunk + i, gchunks + i);
    cp_commit();

    const int N = p.n_agents, D = p.obs_dim;
    const int64_t e0 = (int64_t)blockIdx.x * (kPolicyWarps * 16) + warp * 16;
    const int64_t eA = e0 + g, eB = e0 + g + 8;
    const bool vA = eA < p.B, vB = eB < p.B;
    const bool zA = !vA || (fresh && fresh[eA]), zB = !vB || (fresh && fresh[eB]);     // rows that read as zero
    const int64_t rA = ((vA ? eA : 0) * N + agent), rB = ((vB ? eB : 0) * N + agent);
    const float *oA = obs + rA * D, *oB = obs + rB * D;
    float *hA_ptr = hidden + rA * H, *hB_ptr = hidden + rB * H;

    // ---- observation (+ one-hot agent id, misc.py:161-162) as A fragments
    uint32_t oa[kMaxKS1][4];
    auto in_val = [&](const float *o, bool zero, int c) -> float {
        if (c < D) return zero ? 0.f : o[c];
        return (p.obs_agent_id && c - D == agent) ? 1.f : 0.f;
    };
#pragma unroll
    for (int s = 0; s < kMaxKS1; s++) {
        if (s < p.KS1) {
            oa[s][0] = to_tf32(in_val(oA, zA, 8 * s + t));
            oa[s][1] = to_tf32(in_val(oB, zB, 8 * s + t));
            oa[s][2] = to_tf32(in_val(oA, zA, 8 * s + t + 4));
            oa[s][3] = to_tf32(in_val(oB, zB, 8 * s + t + 4));
        } else {
            oa[s][0] = oa[s][1] = oa[s][2] = oa[s][3] = 0u;
        }
    }
    // ---- hidden state as A fragments (GRU only)
    uint32_t ha[RNN ? KS : 1][4];
    if (RNN) {
#pragma unroll
        for (int s = 0; s < KS; s++) {
            ha[s][0] = to_tf32(zA ? 0.f : hA_ptr[8 * s + t]);
            ha[s][1] = to_tf32(zB ? 0.f : hB_ptr[8 * s + t]);
            ha[s][2] = to_tf32(zA ? 0.f : hA_ptr[8 * s + t + 4]);
            ha[s][3] = to_tf32(zB ? 0.f : hB_ptr[8 * s + t + 4]);
        }
    }

    cp_wait<1>();                                            // head (W1, biases, W2) has landed
    __syncthreads();

    // ---- x = relu(fc1(obs))                                 rnn_agent.py:22
    uint32_t xa[KS][4];
#pragma unroll
    for (int j = 0; j < NT; j++) {
        const float2 b = *reinterpret_cast<const float2 *>(sB1 + 8 * j + 2 * t);
        float acc[4] = {b.x, b.y, b.x, b.y};
#pragma unroll
        for (int s = 0; s < kMaxKS1; s++)
            if (s < p.KS1) {
                const float2 w = reinterpret_cast<const float2 *>(sW1)[(j * p.KS1 + s) * 32 + lane];
                mma_tf32(acc, oa[s], w.x, w.y);
            }
#pragma unroll
        for (int i = 0; i < 4; i++) acc[i] = fmaxf(acc[i], 0.f);
        c_to_a(acc, xa[j], lane);
    }

    // ---- recurrent layer, one 8-unit column tile at a time; fc2 accumulated on the fly
    float qacc[kMaxAT][4];
#pragma unroll
    for (int jt = 0; jt < kMaxAT; jt++) {
        const float2 b = jt < p.AT ? *reinterpret_cast<const float2 *>(sB2 + 8 * jt + 2 * t) : make_float2(0.f, 0.f);
        qacc[jt][0] = b.x; qacc[jt][1] = b.y; qacc[jt][2] = b.x; qacc[jt][3] = b.y;
    }
#pragma unroll 1
    for (int j = 0; j < NT; j++) {
        if (j + 1 < NT) {
            float *dst = sChunk + ((j + 1) & 1) * CHUNK;
            const float *src = gchunks + (size_t)(j + 1) * CHUNK;
            for (int i = threadIdx.x * 4; i < CHUNK; i += kPolicyWarps * 128) cp_async16(dst + i, src + i);
        }
        cp_commit();
        cp_wait<1>();                                        // tile j has landed (tile j + 1 may be in flight)
        __syncthreads();
        const float *cb = sChunk + (j & 1) * CHUNK;
        const int c0 = 8 * j + 2 * t;
        float hn[4];
        if (RNN) {
            // torch.nn.GRUCell: r = s(W_ir x + b_ir + W_hr h + b_hr), z likewise, n = tanh(W_in x + b_in + r (W_hn h + b_hn)),
            // h' = (1 - z) n + z h; weight_ih / weight_hh rows are the (r, z, n) gates in that order
            const float *bi = sBias, *bh = sBias + 3 * H;
            float ar[4], az[4], ai[4], ah[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int c = c0 + (i & 1);
                ar[i] = bi[c] + bh[c];
                az[i] = bi[H + c] + bh[H + c];
                ai[i] = bi[2 * H + c];
                ah[i] = bh[2 * H + c];
            }
            const float4 *c4 = reinterpret_cast<const float4 *>(cb);
#pragma unroll
            for (int sp = 0; sp < KS / 2; sp++) {
                float4 w;
                w = c4[(0 * (KS / 2) + sp) * 32 + lane]; mma_tf32(ar, xa[2 * sp], w.x, w.y); mma_tf32(ar, xa[2 * sp + 1], w.z, w.w);
                w = c4[(1 * (KS / 2) + sp) * 32 + lane]; mma_tf32(az, xa[2 * sp], w.x, w.y); mma_tf32(az, xa[2 * sp + 1], w.z, w.w);
                w = c4[(2 * (KS / 2) + sp) * 32 + lane]; mma_tf32(ai, xa[2 * sp], w.x, w.y); mma_tf32(ai, xa[2 * sp + 1], w.z, w.w);
                w = c4[(3 * (KS / 2) + sp) * 32 + lane]; mma_tf32(ar, ha[2 * sp], w.x, w.y); mma_tf32(ar, ha[2 * sp + 1], w.z, w.w);
                w = c4[(4 * (KS / 2) + sp) * 32 + lane]; mma_tf32(az, ha[2 * sp], w.x, w.y); mma_tf32(az, ha[2 * sp + 1], w.z, w.w);
                w = c4[(5 * (KS / 2) + sp) * 32 + lane]; mma_tf32(ah, ha[2 * sp], w.x, w.y); mma_tf32(ah, ha[2 * sp + 1], w.z, w.w);
            }
            const float2 oldA = zA ? make_float2(0.f, 0.f) : *reinterpret_cast<const float2 *>(hA_ptr + c0);
            const float2 oldB = zB ? make_float2(0.f, 0.f) : *reinterpret_cast<const float2 *>(hB_ptr + c0);
            const float old[4] = {oldA.x, oldA.y, oldB.x, oldB.y};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float r = sigmoidf_(ar[i]), z = sigmoidf_(az[i]);
                const float n = tanhf(ai[i] + r * ah[i]);
                hn[i] = (1.f - z) * n + z * old[i];
            }
        } else {
            // h = relu(rnn(x)) with rnn = nn.Linear                rnn_agent.py:26-27
            const float2 b = *reinterpret_cast<const float2 *>(sBias + c0);
            hn[0] = b.x; hn[1] = b.y; hn[2] = b.x; hn[3] = b.y;
            const float4 *c4 = reinterpret_cast<const float4 *>(cb);
#pragma unroll
            for (int sp = 0; sp < KS / 2; sp++) {
                const float4 w = c4[sp * 32 + lane];
                mma_tf32(hn, xa[2 * sp], w.x, w.y);
                mma_tf32(hn, xa[2 * sp + 1], w.z, w.w);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) hn[i] = fmaxf(hn[i], 0.f);
        }
        if (vA) *reinterpret_cast<float2 *>(hA_ptr + c0) = make_float2(hn[0], hn[1]);
        if (vB) *reinterpret_cast<float2 *>(hB_ptr + c0) = make_float2(hn[2], hn[3]);
        // q += h'[:, tile j] * fc2.weight[:, tile j]'             rnn_agent.py:28
        uint32_t hp[4];
        c_to_a(hn, hp, lane);
#pragma unroll
        for (int jt = 0; jt < kMaxAT; jt++)
            if (jt < p.AT) {
                const float2 w = reinterpret_cast<const float2 *>(sW2)[(jt * KS + j) * 32 + lane];
                mma_tf32(qacc[jt], hp, w.x, w.y);
            }
        __syncthreads();                                     // everyone is done with buffer j & 1
    }

    // ---- greedy action: first maximum, like np.argmax (misc.py:170)
    float bestA = -INFINITY, bestB = -INFINITY;
    int idxA = 0, idxB = 0;
#pragma unroll
    for (int jt = 0; jt < kMaxAT; jt++)
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int c = 8 * jt + 2 * t + k;
            if (jt < p.AT && c < p.n_actions) {
                if (qacc[jt][k] > bestA) { bestA = qacc[jt][k]; idxA = c; }
                if (qacc[jt][2 + k] > bestB) { bestB = qacc[jt][2 + k]; idxB = c; }
                if (q_out) {
                    if (vA) q_out[rA * p.n_actions + c] = qacc[jt][k];
                    if (vB) q_out[rB * p.n_actions + c] = qacc[jt][2 + k];
                }
            }
        }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
        const float ovA = __shfl_xor_sync(0xffffffffu, bestA, o), ovB = __shfl_xor_sync(0xffffffffu, bestB, o);
        const int oiA = __shfl_xor_sync(0xffffffffu, idxA, o), oiB = __shfl_xor_sync(0xffffffffu, idxB, o);
        if (ovA > bestA || (ovA == bestA && oiA < idxA)) { bestA = ovA; idxA = oiA; }
        if (ovB > bestB || (ovB == bestB && oiB < idxB)) { bestB = ovB; idxB = oiB; }
    }
    if (t == 0) {
        if (vA) actions[rA] = idxA;
        if (vB) actions[rB] = idxB;
    }
}

}  // namespace mrb

// ------------------------------------------------------------------------------------------------ host side
using namespace mrb;

struct mrb_policy {
    mrb_policy_desc d;
    int device;
    float *wpack;
    PolicyParams p;
    size_t smem_bytes;
    std::string err;
};

static std::string g_policy_create_error;
namespace mrb { void count_launch(); }       // capi.cu: the library-wide launch counter

static int pfail(mrb_policy *p, int code, const std::string &msg)
{
    if (p) p->err = msg; else g_policy_create_error = msg;
    return code;
}

// cvt.rna.tf32.f32 on the host: round to 10 explicit mantissa bits, ties away from zero
static float tf32_round(float x)
{
    uint32_t u;
    std::memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) != 0x7f800000u) { u += 0x1000u; u &= 0xffffe000u; }
    std::memcpy(&x, &u, 4);
    return x;
}

// B fragments of W' for column tile j, k-step s (W in torch layout [out][in]; rows >= rows_valid and columns >=
// cols_valid read as zero): lane (g, t) holds b0 = W[row0 + 8j + g][8s + t], b1 = W[row0 + 8j + g][8s + t + 4]
static void frag(const float *W, int ld, int row0, int rows_valid, int cols_valid, int j, int s, int lane, float &b0, float &b1)
{
    const int g = lane >> 2, t = lane & 3, r = 8 * j + g;
    const int k0 = 8 * s + t, k1 = k0 + 4;
    b0 = (r < rows_valid && k0 < cols_valid) ? tf32_round(W[(size_t)(row0 + r) * ld + k0]) : 0.f;
    b1 = (r < rows_valid && k1 < cols_valid) ? tf32_round(W[(size_t)(row0 + r) * ld + k1]) : 0.f;
}

extern "C" const char *mrb_policy_last_error(const mrb_policy *p) { return p ? p->err.c_str() : g_policy_create_error.c_str(); }

extern "C" int mrb_policy_create(const mrb_policy_desc *desc, int device, const float *weights, int64_t num_weights,
                                 mrb_policy **out)
{
    if (!desc || !weights || !out) return pfail(nullptr, MRB_E_ARG, "mrb_policy_create: null argument");
    if (desc->struct_size != (int32_t)sizeof(mrb_policy_desc))
        return pfail(nullptr, MRB_E_ARG, "mrb_policy_create: mrb_policy_desc.struct_size does not match this library");
    const mrb_policy_desc &d = *desc;
    const int H = d.hidden_dim, Din = d.input_dim, A = d.n_actions, N = d.n_agents;
    if (H != 64 && H != 128) return pfail(nullptr, MRB_E_UNSUPPORTED, "mrb_policy_create: hidden_dim must be 64 or 128");
    if (A < 1 || A > 8 * kMaxAT) return pfail(nullptr, MRB_E_UNSUPPORTED, "mrb_policy_create: n_actions must be in [1, 24]");
    if (N < 1 || N > MRB_MAX_ROBOTS) return pfail(nullptr, MRB_E_ARG, "mrb_policy_create: n_agents must be in [1, 32]");
    if (Din < 1 || Din > 8 * kMaxKS1) return pfail(nullptr, MRB_E_UNSUPPORTED, "mrb_policy_create: input_dim must be in [1, 64]");
    if (Din != d.obs_dim + (d.obs_agent_id ? N : 0))
        return pfail(nullptr, MRB_E_ARG, "mrb_policy_create: input_dim must equal obs_dim (+ n_agents with obs_agent_id): "
                                         "these weights were trained on a different observation layout");
    const int sets = d.non_shared ? N : 1;
    const int64_t per_set = (int64_t)H * Din + H + (d.use_rnn ? 2LL * 3 * H * H + 2 * 3 * H : (int64_t)H * H + H) + (int64_t)A * H + A;
    if (num_weights != per_set * sets)
        return pfail(nullptr, MRB_E_ARG, "mrb_policy_create: weight count does not match the descriptor");
    int ndev = 0;
    cudaError_t st = cudaGetDeviceCount(&ndev);
    if (st != cudaSuccess || ndev == 0) return pfail(nullptr, MRB_E_CUDA, "mrb_policy_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return pfail(nullptr, MRB_E_ARG, "mrb_policy_create: bad device index");
    cudaDeviceProp prop;
    if ((st = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return pfail(nullptr, MRB_E_CUDA, cudaGetErrorString(st));
    if (prop.major != 10) return pfail(nullptr, MRB_E_UNSUPPORTED, "mrb_policy_create: kernels are built for sm_100a (B200) only");

    const int KS1 = (Din + 7) / 8, NT = H / 8, KS = H / 8, AT = (A + 7) / 8, G = d.use_rnn ? 6 : 1;
    const int szW1 = NT * KS1 * 64, szW2 = AT * KS * 64, NB = d.use_rnn ? 6 * H : H;
    const int head = szW1 + H + NB + szW2 + AT * 8;
    const int chunk = G * KS * 64;
    const int64_t set_floats = head + (int64_t)NT * chunk;
    std::vector<float> img((size_t)set_floats * sets, 0.f);
    for (int sidx = 0; sidx < sets; sidx++) {
        const float *w = weights + per_set * sidx;
        const float *fc1w = w, *fc1b = fc1w + (size_t)H * Din;
        const float *rw = fc1b + H;
        const float *wih = rw, *whh = nullptr, *bih = nullptr, *bhh = nullptr, *fc2w = nullptr;
        if (d.use_rnn) { whh = wih + (size_t)3 * H * H; bih = whh + (size_t)3 * H * H; bhh = bih + 3 * H; fc2w = bhh + 3 * H; }
        else { bih = wih + (size_t)H * H; fc2w = bih + H; }
        const float *fc2b = fc2w + (size_t)A * H;
        float *o = img.data() + (size_t)set_floats * sidx;
        float *oW1 = o, *oB1 = oW1 + szW1, *oBias = oB1 + H, *oW2 = oBias + NB, *oB2 = oW2 + szW2, *oCh = o + head;
        for (int j = 0; j < NT; j++)
            for (int s = 0; s < KS1; s++)
                for (int l = 0; l < 32; l++)
                    frag(fc1w, Din, 0, H, Din, j, s, l, oW1[((j * KS1 + s) * 32 + l) * 2], oW1[((j * KS1 + s) * 32 + l) * 2 + 1]);
        std::memcpy(oB1, fc1b, sizeof(float) * H);
        if (d.use_rnn) { std::memcpy(oBias, bih, sizeof(float) * 3 * H); std::memcpy(oBias + 3 * H, bhh, sizeof(float) * 3 * H); }
        else std::memcpy(oBias, bih, sizeof(float) * H);
        for (int jt = 0; jt < AT; jt++)
            for (int s = 0; s < KS; s++)
                for (int l = 0; l < 32; l++)
                    frag(fc2w, H, 0, A, H, jt, s, l, oW2[((jt * KS + s) * 32 + l) * 2], oW2[((jt * KS + s) * 32 + l) * 2 + 1]);
        for (int a = 0; a < A; a++) oB2[a] = fc2b[a];
        for (int j = 0; j < NT; j++)
            for (int gi = 0; gi < G; gi++) {
                const float *W = d.use_rnn ? (gi < 3 ? wih : whh) : wih;
                const int row0 = d.use_rnn ? (gi % 3) * H : 0;
                for (int sp = 0; sp < KS / 2; sp++)
                    for (int l = 0; l < 32; l++) {
                        float *q = oCh + (size_t)j * chunk + ((size_t)(gi * (KS / 2) + sp) * 32 + l) * 4;
                        frag(W, H, row0, H, H, j, 2 * sp, l, q[0], q[1]);
                        frag(W, H, row0, H, H, j, 2 * sp + 1, l, q[2], q[3]);
                    }
            }
    }
    mrb_policy *pol = new (std::nothrow) mrb_policy();
    if (!pol) return pfail(nullptr, MRB_E_ARG, "mrb_policy_create: out of host memory");
    pol->d = d;
    pol->device = device;
    pol->wpack = nullptr;
    if ((st = cudaSetDevice(device)) != cudaSuccess || (st = cudaMalloc(&pol->wpack, img.size() * sizeof(float))) != cudaSuccess ||
        (st = cudaMemcpy(pol->wpack, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) {
        const std::string msg = std::string("mrb_policy_create: ") + cudaGetErrorString(st);
        if (pol->wpack) cudaFree(pol->wpack);
        delete pol;
        return pfail(nullptr, MRB_E_CUDA, msg);
    }
    PolicyParams &p = pol->p;
    p.wpack = pol->wpack; p.set_floats = set_floats; p.B = 0;
    p.obs_dim = d.obs_dim; p.input_dim = Din; p.n_actions = A; p.n_agents = N;
    p.obs_agent_id = d.obs_agent_id; p.non_shared = d.non_shared;
    p.KS1 = KS1; p.AT = AT; p.szW1 = szW1; p.szW2 = szW2; p.head_floats = head;
    pol->smem_bytes = sizeof(float) * ((size_t)head + 2 * (size_t)chunk);
    *out = pol;
    return MRB_OK;
}

extern "C" int mrb_policy_destroy(mrb_policy *p)
{
    if (!p) return MRB_E_ARG;
    cudaSetDevice(p->device);
    if (p->wpack) cudaFree(p->wpack);
    delete p;
    return MRB_OK;
}

template <int H, bool RNN>
static cudaError_t launch_policy(const mrb_policy *pol, const PolicyParams &p, const float *obs, float *hidden, int32_t *actions,
                                 float *q, const uint8_t *fresh, cudaStream_t s)
{
    cudaError_t st = cudaFuncSetAttribute(policy_act_kernel<H, RNN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pol->smem_bytes);
    if (st != cudaSuccess) return st;
    const dim3 grid((unsigned)((p.B + kPolicyWarps * 16 - 1) / (kPolicyWarps * 16)), (unsigned)p.n_agents);
    policy_act_kernel<H, RNN><<<grid, kPolicyWarps * 32, pol->smem_bytes, s>>>(p, obs, hidden, actions, q, fresh);
    return cudaGetLastError();
}

extern "C" int mrb_policy_act(mrb_policy *pol, int64_t num_envs, const float *obs, float *hidden, int32_t *actions,
                              float *q, const uint8_t *fresh, void *stream)
{
    if (!pol || !obs || !hidden || !actions || num_envs < 1) return pfail(pol, MRB_E_ARG, "mrb_policy_act: bad argument");
    if (((uintptr_t)hidden & 7) != 0) return pfail(pol, MRB_E_ARG, "mrb_policy_act: hidden must be 8-byte aligned");
    cudaError_t st = cudaSetDevice(pol->device);
    if (st != cudaSuccess) return pfail(pol, MRB_E_CUDA, cudaGetErrorString(st));
    PolicyParams p = pol->p;
    p.B = num_envs;
    cudaStream_t s = (cudaStream_t)stream;
    const bool rnn = pol->d.use_rnn != 0;
    if (pol->d.hidden_dim == 128) st = rnn ? launch_policy<128, true>(pol, p, obs, hidden, actions, q, fresh, s)
                                            : launch_policy<128, false>(pol, p, obs, hidden, actions, q, fresh, s);
    else st = rnn ? launch_policy<64, true>(pol, p, obs, hidden, actions, q, fresh, s)
                  : launch_policy<64, false>(pol, p, obs, hidden, actions, q, fresh, s);
    if (st != cudaSuccess) return pfail(pol, MRB_E_CUDA, std::string("policy kernel launch: ") + cudaGetErrorString(st));
    count_launch();
    return MRB_OK;
}
