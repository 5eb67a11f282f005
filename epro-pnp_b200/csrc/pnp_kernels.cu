// pnp_kernels.cu -- sm_100a kernels of the EPro-PnP hot path and their C ABI (include/epropnp_b200.h).
//
// Execution model (all solve kernels):
//   * persistent grid: one 128-thread CTA works on one object at a time and strides over the batch;
//   * the object's correspondence set {x3d, x2d, w2d} is pulled from HBM exactly once by TMA bulk
//     copies (cp.async.bulk -> mbarrier complete_tx) into a 2-slot staging ring, re-packed into a
//     64-byte record per PAIR of points {X0 X1 Y0 Y1 | Z0 Z1 -u0 -u1 | -v0 -v1 wu0 wu1 | wv0 wv1 . .}
//     in shared memory (operands of the packed fp32x2 FFMA2 the AMIS sweep runs on), and every later pass (K+1 LM evaluations, I AMIS
//     cost sweeps over S samples) reads shared memory only; the next object's chunks are already in
//     flight while the current object is being solved;
//   * LM: threads stride over points, 28 partial sums (21 J^T J + 6 J^T r + cost) are reduced with
//     a transposed butterfly (31 shuffles per warp instead of 140) and one cross-warp pass; the 6x6
//     damped Cholesky solve, SE(3) retraction and trust-region logic run on one thread out of
//     shared memory;
//   * AMIS: one thread owns one sample of the iteration: draws it (injected noise or Philox), sweeps
//     all N points with the pre-multiplied projection K[R|t] (points are smem broadcasts), evaluates
//     the proposal densities it needs; the proposal refit is a handful of block reductions.
// No tensor cores: the only contraction is 6-deep, the work is FP32-pipe + MUFU bound.
//
// Build options (all OFF in the shipped library; DESIGN.md section 9.2, tools/variants.py, and the CPU emulation in
// tests/simt_emul run every one of them): EPNP_LM_PACKED, EPNP_LM_NOREFINE, EPNP_LM_COST_FIRST, EPNP_SWEEP_RSQ,
// EPNP_SWEEP_NOCLAMP, EPNP_SWEEP_SPLIT, EPNP_FAST_BLOCKSUM (candidate speed-ups awaiting their first GPU A/B),
// EPNP_TF32X3_NUMERICS (accuracy study), EPNP_PHASE_TIMERS (profiling), EPNP_SIMT_EMUL (g++ build for the emulator).
// The kernels of the default build are pinned by profiles/validated_sass.json (tools/sass_identity.py).
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cstdlib>

#include "pnp_math.cuh"

// Kernel launches and the dynamic shared-memory declaration are spelled through two macros so that this file also
// builds, unchanged, under the test-only SIMT emulator (tests/simt_emul, g++ -DEPNP_SIMT_EMUL) that lets the CPU
// suite execute the kernels' control flow.  In the nvcc build they expand to the plain CUDA forms.
#if defined(EPNP_SIMT_EMUL)
#define EPNP_LAUNCH(kern, grid, block, smem, stream, ...) simt::launch(grid, block, smem, [&] { kern(__VA_ARGS__); })
#define EPNP_DYN_SMEM(type, name, align) type* name = reinterpret_cast<type*>(simt::state().dyn_smem)
#else
#define EPNP_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define EPNP_DYN_SMEM(type, name, align) extern __shared__ __align__(align) type name[]
#endif

namespace {
using namespace pnp;

#if defined(EPNP_SWEEP_MMA) && !defined(EPNP_SWEEP_SPLIT)
#define EPNP_SWEEP_SPLIT 1              // the tensor-pipe sweep reuses the draw-then-sweep control flow
#endif
#if (defined(EPNP_SWEEP_NOCLAMP) || defined(EPNP_SWEEP_SPLIT) || defined(EPNP_TF32X3_NUMERICS)) && !defined(EPNP_SWEEP_RSQ)
#define EPNP_SWEEP_RSQ 1
#endif

#if defined(EPNP_NO_LW) && !defined(EPNP_AMIS_LSE)
#define EPNP_AMIS_LSE 1                 // recomputing a log-weight needs the mixture density as a single value
#endif
// EPNP_NO_LW (experiment): no per-sample log-weight buffer in shared memory (-4 M bytes).  A log-weight is
// -cost - (lse - log count); the refit passes recompute it where they need it, the last iteration writes it straight
// to the output, and the split sweep parks the second point half of a new sample's cost in the object's (not yet
// written) log-weight OUTPUT slot instead.
#if defined(EPNP_NO_LW)
#define LW_STORE(m, v)
#define LW_LOGWEIGHT(m) (-cst[m] - (logp[m] - log_cnt))
#define LW_E(m) expf(LW_LOGWEIGHT(m) - mx)
#define LW_E_STORE(m, e)
#define LW_W(m) (LW_E(m) * inv_sum)
#define LW_W_STORE(m, w)
#else
#define LW_STORE(m, v) lw[m] = (v)
#define LW_E(m) expf(lw[m] - mx)
#define LW_E_STORE(m, e) lw[m] = (e)
#define LW_W(m) (lw[m] * inv_sum)
#define LW_W_STORE(m, w) lw[m] = (w)
#endif

#if defined(EPNP_SWEEP_MMA)
#define EPNP_PTAB_PARAM , float* ptab
#define EPNP_PTAB_ARG(x) , (x)
#else
#define EPNP_PTAB_PARAM
#define EPNP_PTAB_ARG(x)
#endif

constexpr int NT = 128;                 // threads per CTA
constexpr int NW = NT / 32;
constexpr int CH = 128;                 // correspondences per TMA chunk (2 slots x 3.5 KB)
constexpr int STAGE_FLOATS = CH * 7;    // x3d (3) + x2d (2) + w2d (2)
constexpr int MAX_ITER = 8;             // AMIS iterations supported (reference default 4)
constexpr int PROP_FLOATS = 19;         // proposals dump: mu3, Lt6, Lr10
constexpr size_t SMEM_LIMIT = 227 * 1024;

thread_local int g_last_cuda_error = 0;

// Phase timers (profiling build: -DEPNP_PHASE_TIMERS; tools/phase_profile.py).  Thread 0 of every CTA adds
// the clock64() cycles it spent in each phase; the production build compiles them away.
enum Phase { PH_LOAD = 0, PH_LM_EVAL, PH_LM_SERIAL, PH_COV, PH_INIT_FIT, PH_DRAW_SWEEP, PH_LOGP_OLD, PH_WEIGHTS,
             PH_REFIT_SUMS, PH_REFIT_FINISH, PH_OUTPUT, PH_COUNT };
#ifdef EPNP_PHASE_TIMERS
#define PH_DECL long long ph_t = clock64()
#define PH_MARK(a_, which)                                                                     \
    do {                                                                                       \
        if ((int)threadIdx.x == serial_thread(a_) && (a_).prof) {                                                \
            const long long now_ = clock64();                                                  \
            atomicAdd((a_).prof + (which), (unsigned long long)(now_ - ph_t));                 \
            ph_t = now_;                                                                       \
        }                                                                                      \
    } while (0)
#else
#define PH_DECL
#define PH_MARK(a_, which)
#endif

// ------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + 1-D TMA bulk copy
#if defined(EPNP_SIMT_EMUL)
// emulator: an mbarrier word is {completed phases (low 32 bits), bytes still expected (high 32 bits)}; a bulk copy
// is a memcpy that retires its bytes and completes the phase when none are left; waiting on a parity yields to the
// other fibers until that phase has completed.  Exact libm stands in for the approximate special-function units.
inline void mbar_init(uint64_t* bar, uint32_t) { *bar = 0; }
inline void fence_barrier_init() {}
inline void fence_proxy_async() {}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { *bar += (uint64_t)bytes << 32; }
inline void mbar_wait(uint64_t* bar, uint32_t parity) { while (((uint32_t)*bar & 1u) == parity) simt::yield(); }
inline void tma_load_1d(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    std::memcpy(dst_smem, src, bytes);
    *bar -= (uint64_t)bytes << 32;
    if ((*bar >> 32) == 0) *bar = (uint32_t)*bar + 1u;
}
struct FastRcp { float operator()(float x) const { return 1.0f / x; } };
struct FastSqrt { float operator()(float x) const { return sqrtf(x); } };
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct FastRcp {
    __device__ __forceinline__ float operator()(float x) const { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
};
struct FastSqrt {
    __device__ __forceinline__ float operator()(float x) const { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
};
#endif

// ------------------------------------------------------------------------------------------------
struct KArgs {
    const float *x3d, *x2d, *w2d, *cam, *lb, *ub, *delta;
    const float *pose_init;                 // LM entry
    const float *pose_opt_in, *pose_cov_in; // AMIS-only entry
    const float *noise_n3, *noise_chi2, *noise_rot;
    const float *poses;                     // cost-only entry: (S, B, D)
    float *pose_opt, *pose_cov, *cost, *pose_plus, *cost_init;
    float *pose_samples, *logw, *proposals;
    float *cost_out;                        // cost-only entry: (S, B)
    int B, N, S_eval, use_tma, num_sms;
    uint32_t obj_offset;
    uint64_t seed;
    unsigned long long* prof;               // phase timers (profiling build only), else unused
    Params p;
};

// Static part of the shared-memory image (the dynamic arrays follow it).
template <int DOF> struct SmemHead {
    uint64_t bar[2];
    float red[2 * NW * 32];                 // two halves: consecutive reductions alternate, one barrier each
    float ev[32];                           // reduced evaluation: NV floats
    LMState<DOF> lm;
    float cov[DOF * DOF];
    Proposal6 prop[MAX_ITER];
    Proposal4 prop4[MAX_ITER];
};

struct SmemPlan {        // offsets in floats from the start of dynamic smem
    int stage, pts, smp, cost, logp, lw, total_bytes;
};

// alias_stage (build option EPNP_ALIAS_STAGE, one object per CTA only): the staging ring is dead once the object is
// packed and the sample buffer is not written before the AMIS loop, so the ring lives inside it.
template <int DOF>
__host__ __device__ inline bool can_alias_stage(int M, bool amis) {
#if defined(EPNP_ALIAS_STAGE)
    return amis && Dim<DOF>::POSE * M >= 2 * STAGE_FLOATS;
#else
    return false;
#endif
}
template <int DOF>
__host__ __device__ inline SmemPlan plan_smem(int N, int M, int I, bool amis, bool alias_stage = false) {
    SmemPlan s;
    int off = (int)((sizeof(SmemHead<DOF>) + 127) / 128 * 128 / 4);
    const bool alias = alias_stage && can_alias_stage<DOF>(M, amis);
    s.stage = off; if (!alias) off += 2 * STAGE_FLOATS;
    s.pts = off; off += 16 * ((N + 1) / 2);        // 64 B per pair of points
    s.smp = off; if (amis) off += Dim<DOF>::POSE * M;
    if (alias) s.stage = s.smp;
    s.cost = off; if (amis) off += M;
#if defined(EPNP_AMIS_LSE)
    s.logp = off; if (amis) off += M;               // running log-sum-exp of the proposal densities, one per sample
#else
    s.logp = off; if (amis) off += I * M;
#endif
#if defined(EPNP_NO_LW)
    s.lw = off;                                     // no log-weight buffer
#else
    s.lw = off; if (amis) off += M;
#endif
    s.total_bytes = off * 4;
    return s;
}

// Packed point store.  Pair j = points (2j, 2j+1) occupies 16 floats:
//   [0..3] X0 X1 Y0 Y1   [4..7] Z0 Z1 -u0 -u1   [8..11] -v0 -v1 wu0 wu1   [12..15] wv0 wv1 0 0
__device__ __forceinline__ void store_point(float* pts, int n, float X, float Y, float Z, float u, float v, float wu, float wv) {
    float* p = pts + (n >> 1) * 16 + (n & 1);
    p[0] = X; p[2] = Y; p[4] = Z; p[6] = -u; p[8] = -v; p[10] = wu; p[12] = wv;
#if defined(EPNP_SWEEP_MMA)
    p[14] = 1.0f;           // the homogeneous coordinate, so that a B fragment is one load for every lane
#endif
}
// odd N: the second half of the last pair is a zero-weight copy of the last point (contributes exactly 0);
// written by the SAME thread that stores point N-1, so no other thread's data is read
__device__ __forceinline__ void store_point_padded(float* pts, int n, int N, float X, float Y, float Z, float u, float v,
                                                   float wu, float wv) {
    store_point(pts, n, X, Y, Z, u, v, wu, wv);
    if ((N & 1) && n == N - 1) store_point(pts, n + 1, X, Y, Z, u, v, 0.f, 0.f);
}
// ------------------------------------------------------------------------------------------------
// Correspondence loader: TMA ring (or plain loads when pointers / N break the 16-byte rules).
struct Loader {
    const KArgs& a;
    uint64_t* bar;
    float* stage;
    int nch, n_my, total;

    __device__ Loader(const KArgs& a_, uint64_t* bar_, float* stage_) : a(a_), bar(bar_), stage(stage_) {
        nch = (a.N + CH - 1) / CH;
        n_my = ((int)blockIdx.x < a.B) ? (a.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        total = n_my * nch;
    }
    __device__ void issue(int c) {          // one thread
        const int obj = (int)blockIdx.x + (c / nch) * (int)gridDim.x;
        const int k = c % nch;
        const int npts = min(CH, a.N - k * CH);
        const size_t first = (size_t)obj * a.N + (size_t)k * CH;
        float* dst = stage + (c & 1) * STAGE_FLOATS;
        uint64_t* b = bar + (c & 1);
        mbar_expect_tx(b, (uint32_t)npts * 28u);
        tma_load_1d(dst, a.x3d + first * 3, (uint32_t)npts * 12u, b);
        tma_load_1d(dst + CH * 3, a.x2d + first * 2, (uint32_t)npts * 8u, b);
        tma_load_1d(dst + CH * 5, a.w2d + first * 2, (uint32_t)npts * 8u, b);
    }
    __device__ void prologue() {
        if (a.use_tma && threadIdx.x == 0) {
            mbar_init(bar + 0, 1);
            mbar_init(bar + 1, 1);
            fence_barrier_init();
        }
        __syncthreads();
        if (a.use_tma && threadIdx.x == 0) {
            if (total > 0) issue(0);
            if (total > 1) issue(1);
        }
    }
    // Bring object number `it` of this CTA into the packed point array. Ends with a __syncthreads.
    __device__ void load_object(int it, int obj, float* pts) {
        const int tid = threadIdx.x;
        if (a.use_tma) {
            for (int k = 0; k < nch; ++k) {
                const int c = it * nch + k;
                const float* st = stage + (c & 1) * STAGE_FLOATS;
                mbar_wait(bar + (c & 1), (uint32_t)((c >> 1) & 1));
                const int npts = min(CH, a.N - k * CH);
                for (int n = tid; n < npts; n += NT) {
                    const float2 uv = reinterpret_cast<const float2*>(st + CH * 3)[n];
                    const float2 w = reinterpret_cast<const float2*>(st + CH * 5)[n];
                    store_point_padded(pts, k * CH + n, a.N, st[3 * n], st[3 * n + 1], st[3 * n + 2], uv.x, uv.y, w.x, w.y);
                }
                __syncthreads();            // slot drained (and, after the last chunk, pts complete)
                if (tid == 0 && c + 2 < total) { fence_proxy_async(); issue(c + 2); }
            }
        } else {
            const float* g3 = a.x3d + (size_t)obj * a.N * 3;
            const float* g2 = a.x2d + (size_t)obj * a.N * 2;
            const float* gw = a.w2d + (size_t)obj * a.N * 2;
            for (int n = tid; n < a.N; n += NT)
                store_point_padded(pts, n, a.N, __ldg(g3 + 3 * n), __ldg(g3 + 3 * n + 1), __ldg(g3 + 3 * n + 2),
                                   __ldg(g2 + 2 * n), __ldg(g2 + 2 * n + 1), __ldg(gw + 2 * n), __ldg(gw + 2 * n + 1));
            __syncthreads();
        }
    }
};

__device__ __forceinline__ Cam load_cam(const KArgs& a, int obj) {
    Cam c;
#pragma unroll
    for (int i = 0; i < 9; ++i) c.k[i] = __ldg(a.cam + (size_t)obj * 9 + i);
    c.z_min = a.p.z_min;
    c.bounded = (a.lb != nullptr && a.ub != nullptr) ? 1 : 0;
    if (c.bounded) {
        c.lbx = __ldg(a.lb + 2 * obj); c.lby = __ldg(a.lb + 2 * obj + 1);
        c.ubx = __ldg(a.ub + 2 * obj); c.uby = __ldg(a.ub + 2 * obj + 1);
    } else {
        c.lbx = c.lby = -CUDART_INF_F; c.ubx = c.uby = CUDART_INF_F;
    }
    return c;
}

// The once-per-iteration serial work of a CTA (LM step solve, refit finish, first proposal) runs on lane 0
// of ONE warp.  Co-resident CTAs of an SM are typically blockIdx, blockIdx + #SM, ...; rotating the serial
// warp with blockIdx / #SM puts their serial chains on different SM sub-partitions (warp w -> SMSP w % 4)
// instead of all of them competing for the scheduler of warp 0.
__device__ __forceinline__ int serial_thread(const KArgs& a) {
    return 32 * (int)((blockIdx.x / (unsigned)max(a.num_sms, 1)) & (NW - 1));
}

// ------------------------------------------------------------------------------------------------
// Reductions
// 32 values per lane -> lane j holds the warp total of v[j]   (31 shuffles)
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int k = 0; k < half; ++k) {
            const float keep = up ? v[k + half] : v[k];
            const float send = up ? v[k] : v[k + half];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

#if defined(EPNP_FAST_BLOCKSUM)
// experiment (off by default): W values per lane (W = 8, 16 or 32) -> lane l holds the warp total of
// v[l >> (5 - log2 W)]: log2(W) transposed-butterfly stages (W - 1 shuffles) + plain butterfly adds for the rest,
// instead of 5 shuffles per value.
template <int W> __device__ __forceinline__ float warp_transpose_sum_w(float (&v)[W]) {
    static_assert(W == 8 || W == 16 || W == 32, "W must be 8, 16 or 32");
    const int lane = threadIdx.x & 31;
    int m = 16;
#pragma unroll
    for (int half = W / 2; half >= 1; half >>= 1, m >>= 1) {
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int k = 0; k < half; ++k) {
            const float keep = up ? v[k + half] : v[k];
            const float send = up ? v[k] : v[k + half];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
    }
    float r = v[0];
#pragma unroll
    for (int o = 16 / W; o >= 1; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    return r;
}
#endif

// Block-wide sums of K values per thread; every thread gets the totals.  `red` holds two halves of
// NW*32 floats: call sites alternate `half` so one __syncthreads per reduction is enough (a thread can be
// at most one reduction ahead of the slowest reader, and then it writes the other half).
template <int K> __device__ __forceinline__ void block_sum(float (&v)[K], float* red, int half) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* r = red + half * (NW * 32);
#if defined(EPNP_FAST_BLOCKSUM)
    if constexpr (K > 2) {
        constexpr int W = K <= 8 ? 8 : (K <= 16 ? 16 : 32);
        static_assert(K <= 32, "block_sum: at most 32 values");
        float w[W];
#pragma unroll
        for (int k = 0; k < W; ++k) w[k] = k < K ? v[k] : 0.f;
        const float tot = warp_transpose_sum_w<W>(w);
        constexpr int SH = W == 8 ? 2 : (W == 16 ? 1 : 0);
        if ((lane & ((1 << SH) - 1)) == 0 && (lane >> SH) < K) r[warp * 32 + (lane >> SH)] = tot;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = (r[k] + r[32 + k]) + (r[64 + k] + r[96 + k]);
        return;
    }
#endif
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) r[warp * 32 + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = (r[k] + r[32 + k]) + (r[64 + k] + r[96 + k]);
}

__device__ __forceinline__ float block_max(float v, float* red, int half) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* r = red + half * (NW * 32);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) r[warp * 32] = v;
    __syncthreads();
    return fmaxf(fmaxf(r[0], r[32]), fmaxf(r[64], r[96]));
}

// Evaluate the normal equations at `pose` (shared memory) over all points; result in ev[0..NV).
template <int DOF, bool CLIP>
__device__ void eval_normal_eq(const float* pts, int N, const float* pose, const Cam& cam, float delta,
                               float huber_eps, float* red, float* ev) {
    constexpr int NV = Dim<DOF>::NV;
    float R[9], t[3];
    {
        float ps[Dim<DOF>::POSE];
#pragma unroll
        for (int i = 0; i < Dim<DOF>::POSE; ++i) ps[i] = pose[i];
        pose_to_rot<DOF>(ps, R);
        t[0] = ps[0]; t[1] = ps[1]; t[2] = ps[2];
    }
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    // one PAIR record (4 x LDS.128) per thread and step: 4-way bank conflicts instead of the 16-way a
    // per-point scalar walk over the 64-byte records would cause
    const float4* p4 = reinterpret_cast<const float4*>(pts);
    const int npair = (N + 1) >> 1;
#if defined(EPNP_LM_PACKED)
    // experiment (off by default): u-row / v-row of the Jacobian in the two lanes of packed fp32x2 registers
    {
        constexpr int NP = Dim<DOF>::NA + DOF;
        pnp::V2 acc2[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) acc2[i] = pnp::v2splat(0.f);
        const pnp::V2 kuv[3] = {pnp::v2(cam.k[0], cam.k[3]), pnp::v2(cam.k[1], cam.k[4]), pnp::v2(cam.k[2], cam.k[5])};
        float cost = 0.f;
        for (int j = threadIdx.x; j < npair; j += NT) {
            const float4 q0 = p4[4 * j], q1 = p4[4 * j + 1], q2 = p4[4 * j + 2], q3 = p4[4 * j + 3];
            point_normal_eq_rows<DOF, CLIP>(R, t, cam, kuv, delta, huber_eps, q0.x, q0.z, q1.x, q1.z, q2.x, q2.z, q3.x,
                                            acc2, cost);
            if (2 * j + 1 < N)
                point_normal_eq_rows<DOF, CLIP>(R, t, cam, kuv, delta, huber_eps, q0.y, q0.w, q1.y, q1.w, q2.y, q2.w,
                                                q3.y, acc2, cost);
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) acc[i] = acc2[i].x + acc2[i].y;
        acc[NP] = cost;
    }
#else
    for (int j = threadIdx.x; j < npair; j += NT) {
        const float4 q0 = p4[4 * j], q1 = p4[4 * j + 1], q2 = p4[4 * j + 2], q3 = p4[4 * j + 3];
        point_normal_eq<DOF, CLIP>(R, t, cam, delta, huber_eps, q0.x, q0.z, q1.x, -q1.z, -q2.x, q2.z, q3.x, acc);
        if (2 * j + 1 < N)
            point_normal_eq<DOF, CLIP>(R, t, cam, delta, huber_eps, q0.y, q0.w, q1.y, -q1.w, -q2.y, q2.w, q3.y, acc);
    }
#endif
    const float tot = warp_transpose_sum(acc);
    red[(threadIdx.x >> 5) * 32 + (threadIdx.x & 31)] = tot;
    __syncthreads();
    if (threadIdx.x < NV) {
        const int j = threadIdx.x;
        ev[j] = (red[j] + red[32 + j]) + (red[64 + j] + red[96 + j]);
    }
    __syncthreads();
}

#if defined(EPNP_LM_COST_FIRST)
// experiment (off by default): Huber cost of ONE pose over all resident points, block-parallel (each thread its share
// of the pair records, packed fp32x2, rsqrt seed + one Newton step so the value tracks the normal-equation pass's
// cost to ~1 ulp), every thread gets the total.
struct RefinedRsqrt {
#if defined(EPNP_SIMT_EMUL)
    float operator()(float x) const { return 1.0f / sqrtf(x); }
#else
    __device__ __forceinline__ float operator()(float x) const {
        float y;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y * fmaf(-0.5f * x * y, y, 1.5f);
    }
#endif
};
template <int DOF>
__device__ float block_cost(const float* pts, int N, const float* pose_smem, const Cam& cam, float delta, float* red) {
    float R[9], P[12], ps[Dim<DOF>::POSE];
#pragma unroll
    for (int i = 0; i < Dim<DOF>::POSE; ++i) ps[i] = pose_smem[i];
    pose_to_rot<DOF>(ps, R);
    make_proj(cam.k, R, ps, P);
    float2 P2[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) P2[k] = make_float2(P[k], P[k]);
    const float4* p4 = reinterpret_cast<const float4*>(pts);
    const int npair = (N + 1) >> 1;
    float2 acc = make_float2(0.f, 0.f);
    for (int j = threadIdx.x; j < npair; j += NT) {
        const float4 q0 = p4[4 * j], q1 = p4[4 * j + 1], q2 = p4[4 * j + 2], q3 = p4[4 * j + 3];
        const float2 X = make_float2(q0.x, q0.y), Y = make_float2(q0.z, q0.w), Z = make_float2(q1.x, q1.y);
        const float2 nu = make_float2(q1.z, q1.w), nv = make_float2(q2.x, q2.y), wu = make_float2(q2.z, q2.w), wv = make_float2(q3.x, q3.y);
        acc = cam.bounded ? pnp::pair_cost_rsq<true, true>(P2, cam, delta, X, Y, Z, nu, nv, wu, wv, acc, RefinedRsqrt())
                          : pnp::pair_cost_rsq<false, true>(P2, cam, delta, X, Y, Z, nu, nv, wu, wv, acc, RefinedRsqrt());
    }
    float v[1] = {acc.x + acc.y};
    block_sum<1>(v, red, 1);
    return v[0];
}
#endif

// ------------------------------------------------------------------------------------------------
// LM / GN solve of the object resident in pts4.  Leaves the solution in sh.lm.pose, the covariance
// in sh.cov (when want_cov) and writes the requested outputs.
template <int DOF>
__device__ void lm_phase(const KArgs& a, SmemHead<DOF>& sh, const float* pts4, const Cam& cam, float delta,
                         int obj, bool want_cov) {
    constexpr int PD = Dim<DOF>::POSE;
    const Params& p = a.p;
    const int tid = threadIdx.x;
    const int st = serial_thread(a);
    PH_DECL;
    if (tid == st) {
#pragma unroll
        for (int i = 0; i < PD; ++i) sh.lm.pose[i] = __ldg(a.pose_init + (size_t)obj * PD + i);
        sh.lm.radius = p.initial_radius;
        sh.lm.shrink = 2.0f;
    }
    __syncthreads();
    if (!p.fast_mode) {
        eval_normal_eq<DOF, true>(pts4, a.N, sh.lm.pose, cam, delta, p.huber_eps, sh.red, sh.ev);
        PH_MARK(a, PH_LM_EVAL);
        if (tid == st) {
            lm_adopt<DOF>(sh.lm, sh.ev);
            if (a.cost_init) a.cost_init[obj] = sh.lm.cost;
            if (p.lm_iter > 0) lm_propose<DOF>(sh.lm, p);
        }
        PH_MARK(a, PH_LM_SERIAL);
        __syncthreads();
#if defined(EPNP_LM_COST_FIRST)
        for (int it = 0; it < p.lm_iter; ++it) {
            const float cost_new = block_cost<DOF>(pts4, a.N, sh.lm.pose_new, cam, delta, sh.red);
            if (tid == st) sh.ev[31] = lm_decide<DOF>(sh.lm, cost_new, p) ? 1.f : 0.f;
            __syncthreads();
            if (sh.ev[31] != 0.f) {             // accepted (CTA-uniform): linearise at the new pose
                eval_normal_eq<DOF, true>(pts4, a.N, sh.lm.pose, cam, delta, p.huber_eps, sh.red, sh.ev);
                if (tid == st) lm_adopt<DOF>(sh.lm, sh.ev);
            }
            PH_MARK(a, PH_LM_EVAL);
            if (tid == st && it + 1 < p.lm_iter) lm_propose<DOF>(sh.lm, p);
            PH_MARK(a, PH_LM_SERIAL);
            __syncthreads();
        }
#else
        for (int it = 0; it < p.lm_iter; ++it) {
            eval_normal_eq<DOF, true>(pts4, a.N, sh.lm.pose_new, cam, delta, p.huber_eps, sh.red, sh.ev);
            PH_MARK(a, PH_LM_EVAL);
            if (tid == st) {
                lm_update<DOF>(sh.lm, sh.ev, p);
                if (it + 1 < p.lm_iter) lm_propose<DOF>(sh.lm, p);
            }
            PH_MARK(a, PH_LM_SERIAL);
            __syncthreads();
        }
#endif
    } else {
        for (int it = 0; it < p.lm_iter; ++it) {
            eval_normal_eq<DOF, false>(pts4, a.N, sh.lm.pose, cam, delta, p.huber_eps, sh.red, sh.ev);
            if (tid == st) {
                lm_adopt<DOF>(sh.lm, sh.ev);                 // kept for covariance / cost (pre-step)
                if (it == 0 && a.cost_init) a.cost_init[obj] = sh.lm.cost;
                gn_advance<DOF>(sh.lm.pose, sh.ev, p.eps, sh.lm.pose);
            }
            __syncthreads();
        }
    }
    if (tid < PD) a.pose_opt[(size_t)obj * PD + tid] = sh.lm.pose[tid];
    if (tid == 32 && a.cost) a.cost[obj] = sh.lm.cost;
    if (want_cov) {
        if (tid < DOF) {                    // one covariance column per lane (fp64 Cholesky + one solve)
            float col[DOF];
            pose_covariance_column<DOF>(sh.lm.a, p.eps, tid, col);
#pragma unroll
            for (int i = 0; i < DOF; ++i) sh.cov[i * DOF + tid] = col[i];
        }
        __syncthreads();
        if (a.pose_cov && tid < DOF * DOF) a.pose_cov[(size_t)obj * DOF * DOF + tid] = sh.cov[tid];
    }
    PH_MARK(a, PH_COV);
    if (a.pose_plus) {      // y* (+) one undamped GN step, clip_jac always on (gn_step default)
        eval_normal_eq<DOF, true>(pts4, a.N, sh.lm.pose, cam, delta, p.huber_eps, sh.red, sh.ev);
        if (tid == st) {
            float plus[PD];
            gn_advance<DOF>(sh.lm.pose, sh.ev, p.eps, plus);
#pragma unroll
            for (int i = 0; i < PD; ++i) a.pose_plus[(size_t)obj * PD + i] = plus[i];
        }
    }
    __syncthreads();
}

// Huber cost of one pose over every resident point: thread-private sweep, every lane reads the same pair
// record (shared-memory broadcast, 4 x LDS.128 per 2 points) and evaluates two points per instruction with
// packed fp32x2 arithmetic (SASS FFMA2 / FMUL2 / FADD2).  P2[k] = (P[k], P[k]) is the pre-multiplied
// projection K[R|t] duplicated into both halves.  Same arithmetic per half as pnp::point_cost.
__device__ __forceinline__ float2 splat(float x) { return make_float2(x, x); }

template <bool BOUNDED>
__device__ __forceinline__ float2 pair_cost(const float2 (&P2)[12], const Cam& cam, float2 d2, float2 nhd2,
                                            const float4 q0, const float4 q1, const float4 q2, const float4 q3) {
    const float2 X = make_float2(q0.x, q0.y), Y = make_float2(q0.z, q0.w), Z = make_float2(q1.x, q1.y);
    const float2 nu = make_float2(q1.z, q1.w), nv = make_float2(q2.x, q2.y);
    const float2 wu = make_float2(q2.z, q2.w), wv = make_float2(q3.x, q3.y);
    const float2 xh = __ffma2_rn(P2[0], X, __ffma2_rn(P2[1], Y, __ffma2_rn(P2[2], Z, P2[3])));
    const float2 yh = __ffma2_rn(P2[4], X, __ffma2_rn(P2[5], Y, __ffma2_rn(P2[6], Z, P2[7])));
    const float2 zh = __ffma2_rn(P2[8], X, __ffma2_rn(P2[9], Y, __ffma2_rn(P2[10], Z, P2[11])));
    const float2 iz = make_float2(FastRcp()(fmaxf(zh.x, cam.z_min)), FastRcp()(fmaxf(zh.y, cam.z_min)));
    float2 tx, ty;
    if (BOUNDED) {
        float2 px = __fmul2_rn(xh, iz), py = __fmul2_rn(yh, iz);
        px.x = fminf(fmaxf(px.x, cam.lbx), cam.ubx); px.y = fminf(fmaxf(px.y, cam.lbx), cam.ubx);
        py.x = fminf(fmaxf(py.x, cam.lby), cam.uby); py.y = fminf(fmaxf(py.y, cam.lby), cam.uby);
        tx = __fadd2_rn(px, nu); ty = __fadd2_rn(py, nv);
    } else {
        tx = __ffma2_rn(xh, iz, nu); ty = __ffma2_rn(yh, iz, nv);
    }
    const float2 rx = __fmul2_rn(tx, wu), ry = __fmul2_rn(ty, wv);
    const float2 s2 = __ffma2_rn(rx, rx, __fmul2_rn(ry, ry));
    const float2 s = make_float2(FastSqrt()(s2.x), FastSqrt()(s2.y));
    const float2 inl = __fmul2_rn(s2, splat(0.5f));
    const float2 outl = __ffma2_rn(s, d2, nhd2);
    return make_float2(s.x <= d2.x ? inl.x : outl.x, s.y <= d2.x ? inl.y : outl.y);
}

#if defined(EPNP_SWEEP_RSQ)
// experiments (off by default), all on pnp::pair_cost_rsq -- one MUFU.RSQ per point instead of RCP + SQRT, Huber
// without selects, accumulation folded into the last FFMA2:
//   EPNP_SWEEP_RSQ      the formulation itself
//   EPNP_SWEEP_NOCLAMP  + drop the z clamp when pose_depth_margin() proves it idle for the whole warp
//   EPNP_SWEEP_SPLIT    + two samples per thread over half of the points each (pair-record loads amortised)
#if defined(EPNP_SIMT_EMUL)
struct SweepRsqrt { float operator()(float x) const { return 1.0f / sqrtf(x); } };
#else
struct SweepRsqrt {
    __device__ __forceinline__ float operator()(float x) const { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
};
#endif
template <bool BOUNDED, bool CLAMPZ>
__device__ __forceinline__ float2 pair_cost_acc(const float2 (&P2)[12], const Cam& cam, float delta, float2 acc,
                                                const float4 q0, const float4 q1, const float4 q2, const float4 q3) {
    return pnp::pair_cost_rsq<BOUNDED, CLAMPZ>(P2, cam, delta, make_float2(q0.x, q0.y), make_float2(q0.z, q0.w),
                                               make_float2(q1.x, q1.y), make_float2(q1.z, q1.w),
                                               make_float2(q2.x, q2.y), make_float2(q2.z, q2.w),
                                               make_float2(q3.x, q3.y), acc, SweepRsqrt());
}
template <bool BOUNDED, bool CLAMPZ = true>
__device__ __forceinline__ float sweep_cost(const float4* pts4, int N, const float* P, const Cam& cam, float delta) {
    float2 P2[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) P2[k] = splat(P[k]);
    float2 c0 = splat(0.f), c1 = splat(0.f), c2 = splat(0.f), c3 = splat(0.f);
    const int npair = (N + 1) >> 1;
    int j = 0;
    for (; j + 4 <= npair; j += 4) {
        const float4* q = pts4 + 4 * j;
        c0 = pair_cost_acc<BOUNDED, CLAMPZ>(P2, cam, delta, c0, q[0], q[1], q[2], q[3]);
        c1 = pair_cost_acc<BOUNDED, CLAMPZ>(P2, cam, delta, c1, q[4], q[5], q[6], q[7]);
        c2 = pair_cost_acc<BOUNDED, CLAMPZ>(P2, cam, delta, c2, q[8], q[9], q[10], q[11]);
        c3 = pair_cost_acc<BOUNDED, CLAMPZ>(P2, cam, delta, c3, q[12], q[13], q[14], q[15]);
    }
    for (; j < npair; ++j) {
        const float4* q = pts4 + 4 * j;
        c0 = pair_cost_acc<BOUNDED, CLAMPZ>(P2, cam, delta, c0, q[0], q[1], q[2], q[3]);
    }
    c0 = __fadd2_rn(c0, c2); c1 = __fadd2_rn(c1, c3);
    return (c0.x + c0.y) + (c1.x + c1.y);
}
#elif defined(EPNP_SWEEP_HUBER_M)
// experiment (off by default): the shipped sweep arithmetic (reciprocal + square root) with the select-free Huber
// m (s - m / 2), m = min(s, delta), accumulated by its last FFMA2: 17 instead of 18 packed FP ops per point pair
// (34 FMA-pipe cycles), no FSETP / FSEL, still 4 MUFU.
template <bool BOUNDED>
__device__ __forceinline__ float2 pair_cost_m(const float2 (&P2)[12], const Cam& cam, float delta, float2 acc,
                                              const float4 q0, const float4 q1, const float4 q2, const float4 q3) {
    const float2 X = make_float2(q0.x, q0.y), Y = make_float2(q0.z, q0.w), Z = make_float2(q1.x, q1.y);
    const float2 nu = make_float2(q1.z, q1.w), nv = make_float2(q2.x, q2.y);
    const float2 wu = make_float2(q2.z, q2.w), wv = make_float2(q3.x, q3.y);
    const float2 xh = __ffma2_rn(P2[0], X, __ffma2_rn(P2[1], Y, __ffma2_rn(P2[2], Z, P2[3])));
    const float2 yh = __ffma2_rn(P2[4], X, __ffma2_rn(P2[5], Y, __ffma2_rn(P2[6], Z, P2[7])));
    const float2 zh = __ffma2_rn(P2[8], X, __ffma2_rn(P2[9], Y, __ffma2_rn(P2[10], Z, P2[11])));
    const float2 iz = make_float2(FastRcp()(fmaxf(zh.x, cam.z_min)), FastRcp()(fmaxf(zh.y, cam.z_min)));
    float2 tx, ty;
    if (BOUNDED) {
        float2 px = __fmul2_rn(xh, iz), py = __fmul2_rn(yh, iz);
        px.x = fminf(fmaxf(px.x, cam.lbx), cam.ubx); px.y = fminf(fmaxf(px.y, cam.lbx), cam.ubx);
        py.x = fminf(fmaxf(py.x, cam.lby), cam.uby); py.y = fminf(fmaxf(py.y, cam.lby), cam.uby);
        tx = __fadd2_rn(px, nu); ty = __fadd2_rn(py, nv);
    } else {
        tx = __ffma2_rn(xh, iz, nu); ty = __ffma2_rn(yh, iz, nv);
    }
    const float2 rx = __fmul2_rn(tx, wu), ry = __fmul2_rn(ty, wv);
    const float2 s2 = __ffma2_rn(rx, rx, __fmul2_rn(ry, ry));
    const float2 s = make_float2(FastSqrt()(s2.x), FastSqrt()(s2.y));
    const float2 m = make_float2(fminf(s.x, delta), fminf(s.y, delta));
    return __ffma2_rn(m, __ffma2_rn(m, splat(-0.5f), s), acc);
}
template <bool BOUNDED>
__device__ __forceinline__ float sweep_cost(const float4* pts4, int N, const float* P, const Cam& cam, float delta) {
    float2 P2[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) P2[k] = splat(P[k]);
    float2 c0 = splat(0.f), c1 = splat(0.f), c2 = splat(0.f), c3 = splat(0.f);
    const int npair = (N + 1) >> 1;
    int j = 0;
    for (; j + 4 <= npair; j += 4) {
        const float4* q = pts4 + 4 * j;
        c0 = pair_cost_m<BOUNDED>(P2, cam, delta, c0, q[0], q[1], q[2], q[3]);
        c1 = pair_cost_m<BOUNDED>(P2, cam, delta, c1, q[4], q[5], q[6], q[7]);
        c2 = pair_cost_m<BOUNDED>(P2, cam, delta, c2, q[8], q[9], q[10], q[11]);
        c3 = pair_cost_m<BOUNDED>(P2, cam, delta, c3, q[12], q[13], q[14], q[15]);
    }
    for (; j < npair; ++j) {
        const float4* q = pts4 + 4 * j;
        c0 = pair_cost_m<BOUNDED>(P2, cam, delta, c0, q[0], q[1], q[2], q[3]);
    }
    c0 = __fadd2_rn(c0, c2); c1 = __fadd2_rn(c1, c3);
    return (c0.x + c0.y) + (c1.x + c1.y);
}
#else
template <bool BOUNDED>
__device__ __forceinline__ float sweep_cost(const float4* pts4, int N, const float* P, const Cam& cam, float delta) {
    float2 P2[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) P2[k] = splat(P[k]);
    const float2 d2 = splat(delta), nhd2 = splat(-0.5f * delta * delta);
    float2 c0 = splat(0.f), c1 = splat(0.f), c2 = splat(0.f), c3 = splat(0.f);
    const int npair = (N + 1) >> 1;
    int j = 0;
    for (; j + 4 <= npair; j += 4) {            // 8 points in flight per thread
        const float4* q = pts4 + 4 * j;
        c0 = __fadd2_rn(c0, pair_cost<BOUNDED>(P2, cam, d2, nhd2, q[0], q[1], q[2], q[3]));
        c1 = __fadd2_rn(c1, pair_cost<BOUNDED>(P2, cam, d2, nhd2, q[4], q[5], q[6], q[7]));
        c2 = __fadd2_rn(c2, pair_cost<BOUNDED>(P2, cam, d2, nhd2, q[8], q[9], q[10], q[11]));
        c3 = __fadd2_rn(c3, pair_cost<BOUNDED>(P2, cam, d2, nhd2, q[12], q[13], q[14], q[15]));
    }
    for (; j < npair; ++j) {
        const float4* q = pts4 + 4 * j;
        c0 = __fadd2_rn(c0, pair_cost<BOUNDED>(P2, cam, d2, nhd2, q[0], q[1], q[2], q[3]));
    }
    c0 = __fadd2_rn(c0, c2); c1 = __fadd2_rn(c1, c3);
    return (c0.x + c0.y) + (c1.x + c1.y);
}
#endif

template <int DOF>
__device__ __forceinline__ float pose_cost(const float* pts, int N, const float* pose, const Cam& cam, float delta) {
    float R[9], P[12];
    pose_to_rot<DOF>(pose, R);
    make_proj(cam.k, R, pose, P);
    const float4* pts4 = reinterpret_cast<const float4*>(pts);
    return cam.bounded ? sweep_cost<true>(pts4, N, P, cam, delta) : sweep_cost<false>(pts4, N, P, cam, delta);
}

#if defined(EPNP_SWEEP_NOCLAMP) || defined(EPNP_SWEEP_SPLIT)
// Largest |X| over the resident points (every thread gets it): the `radius` of pose_depth_margin.
__device__ __forceinline__ float object_radius(const float* pts, int N, float* red, int half) {
    const float4* p4 = reinterpret_cast<const float4*>(pts);
    const int npair = (N + 1) >> 1;
    float r2 = 0.f;
    for (int j = threadIdx.x; j < npair; j += NT) {
        const float4 q0 = p4[4 * j], q1 = p4[4 * j + 1];
        r2 = fmaxf(r2, fmaxf(fmaf(q0.x, q0.x, fmaf(q0.z, q0.z, q1.x * q1.x)), fmaf(q0.y, q0.y, fmaf(q0.w, q0.w, q1.y * q1.y))));
    }
    return sqrtf(block_max(r2, red, half));
}

// pose_cost with the clamp-free loop when the whole warp's poses keep every point in front of z_min
// (radius < 0: unknown, always clamp).
template <int DOF>
__device__ __forceinline__ float pose_cost(const float* pts, int N, const float* pose, const Cam& cam, float delta,
                                           float radius) {
    float R[9], P[12];
    pose_to_rot<DOF>(pose, R);
    make_proj(cam.k, R, pose, P);
    const float4* pts4 = reinterpret_cast<const float4*>(pts);
    bool free_z = false;
#if defined(EPNP_SWEEP_NOCLAMP)
    if (radius >= 0.f) free_z = __all_sync(__activemask(), pose_depth_margin(P, radius, cam.z_min) >= 0.f);
#endif
    if (free_z) return cam.bounded ? sweep_cost<true, false>(pts4, N, P, cam, delta) : sweep_cost<false, false>(pts4, N, P, cam, delta);
    return cam.bounded ? sweep_cost<true, true>(pts4, N, P, cam, delta) : sweep_cost<false, true>(pts4, N, P, cam, delta);
}
#endif

#if defined(EPNP_SWEEP_SPLIT)
// Two poses over the pair records [j0, j1): every record is loaded once and used for both.
template <bool BOUNDED, bool CLAMPZ>
__device__ __forceinline__ void sweep_cost2(const float4* pts4, int j0, int j1, const float* Pa, const float* Pb,
                                            const Cam& cam, float delta, float& ca, float& cb) {
    float2 A2[12], B2[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) { A2[k] = splat(Pa[k]); B2[k] = splat(Pb[k]); }
    float2 a0 = splat(0.f), a1 = splat(0.f), b0 = splat(0.f), b1 = splat(0.f);
    int j = j0;
    for (; j + 2 <= j1; j += 2) {
        const float4* q = pts4 + 4 * j;
        const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4], q5 = q[5], q6 = q[6], q7 = q[7];
        a0 = pair_cost_acc<BOUNDED, CLAMPZ>(A2, cam, delta, a0, q0, q1, q2, q3);
        b0 = pair_cost_acc<BOUNDED, CLAMPZ>(B2, cam, delta, b0, q0, q1, q2, q3);
        a1 = pair_cost_acc<BOUNDED, CLAMPZ>(A2, cam, delta, a1, q4, q5, q6, q7);
        b1 = pair_cost_acc<BOUNDED, CLAMPZ>(B2, cam, delta, b1, q4, q5, q6, q7);
    }
    for (; j < j1; ++j) {
        const float4* q = pts4 + 4 * j;
        const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
        a0 = pair_cost_acc<BOUNDED, CLAMPZ>(A2, cam, delta, a0, q0, q1, q2, q3);
        b0 = pair_cost_acc<BOUNDED, CLAMPZ>(B2, cam, delta, b0, q0, q1, q2, q3);
    }
    a0 = __fadd2_rn(a0, a1); b0 = __fadd2_rn(b0, b1);
    ca = a0.x + a0.y; cb = b0.x + b0.y;
}

// Costs of the S new samples smp[m0 .. m0+S) of an AMIS iteration.  Work item w = (sample pair k, point half h):
// the two halves of a sample's cost land in cst[m] (h = 0) and lw[m] (h = 1, free until the weights pass, which
// folds it in).  Odd S: whole sweeps, lw[m] = 0.
template <int DOF>
__device__ void sweep_new_samples(const float* pts, int N, const float* smp, float* cst, float* lw, int m0, int S,
                                  const Cam& cam, float delta, float radius) {
    constexpr int PD = Dim<DOF>::POSE;
    const float4* pts4 = reinterpret_cast<const float4*>(pts);
    if (S & 1) {
        for (int s = threadIdx.x; s < S; s += NT) {
            float pose[PD];
#pragma unroll
            for (int k = 0; k < PD; ++k) pose[k] = smp[(m0 + s) * PD + k];
            cst[m0 + s] = pose_cost<DOF>(pts, N, pose, cam, delta, radius);
            lw[m0 + s] = 0.f;
        }
        return;
    }
    const int npair = (N + 1) >> 1, H = S >> 1;
    const int jmid = min(npair, ((npair >> 1) + 1) & ~1);
    for (int w = threadIdx.x; w < S; w += NT) {
        const int k = w % H, h = w / H;
        const int ma = m0 + 2 * k, mb = ma + 1;
        float Pa[12], Pb[12];
        {
            float pose[PD], R[9];
#pragma unroll
            for (int c = 0; c < PD; ++c) pose[c] = smp[ma * PD + c];
            pose_to_rot<DOF>(pose, R);
            make_proj(cam.k, R, pose, Pa);
#pragma unroll
            for (int c = 0; c < PD; ++c) pose[c] = smp[mb * PD + c];
            pose_to_rot<DOF>(pose, R);
            make_proj(cam.k, R, pose, Pb);
        }
        bool free_z = false;
#if defined(EPNP_SWEEP_NOCLAMP)
        if (radius >= 0.f)
            free_z = __all_sync(__activemask(), fminf(pose_depth_margin(Pa, radius, cam.z_min),
                                                      pose_depth_margin(Pb, radius, cam.z_min)) >= 0.f);
#endif
        const int j0 = h ? jmid : 0, j1 = h ? npair : jmid;
        float ca, cb;
        if (free_z) {
            if (cam.bounded) sweep_cost2<true, false>(pts4, j0, j1, Pa, Pb, cam, delta, ca, cb);
            else sweep_cost2<false, false>(pts4, j0, j1, Pa, Pb, cam, delta, ca, cb);
        } else {
            if (cam.bounded) sweep_cost2<true, true>(pts4, j0, j1, Pa, Pb, cam, delta, ca, cb);
            else sweep_cost2<false, true>(pts4, j0, j1, Pa, Pb, cam, delta, ca, cb);
        }
        float* dst = h ? lw : cst;
        dst[ma] = ca; dst[mb] = cb;
    }
}
#endif

#if defined(EPNP_SWEEP_MMA)
// experiment (off by default): the sweep's 3x4 projection on the tensor pipe through the legacy warp-level
// mma.sync.m16n8k8 TF32 instruction (SASS HMMA.1688.F32.TF32), operands in registers -- no TMEM, no descriptors, no
// extra point storage.  Error-compensated 3xTF32: D = [P_hi | P_lo] [X_hi ; X_hi] + [P_hi | P_lo] [X_lo ; 0].
//   work item = (tile of 16 samples, half of the points); a warp takes items warp, warp + 4, ...
//   A fragments: P[s][4 r + t] of samples s0 = 16 ti + g and s0 + 8 from the table `ptab` (S x 12, built here)
//   B fragments: coordinate t of point 8 tl + g, one LDS.32 from the pair records, split with two ALU ops
//   D fragments: (s0 | s0 + 8) x points (2t, 2t + 1) -- the register pairs the packed Huber tail consumes
// Returns false (nothing done) when the shape does not fit; the caller then runs the CUDA-core path.
__device__ __forceinline__ float tf32_hi_bits(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const float (&a)[4], float b0, float b1) {
#if defined(EPNP_SIMT_EMUL)
    simt::mma_m16n8k8_tf32(d, a, b0, b1);
#else
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])),
                   "r"(__float_as_uint(a[3])), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
#endif
}

template <int DOF, bool BOUNDED>
__device__ __forceinline__ void sweep_mma_items(const float* pts, int N, const float* smp, const float* ptab, float* cst,
                                                float* lw, int m0, int S, const Cam& cam, float delta, float radius);

template <int DOF>
__device__ bool sweep_new_samples_mma(const float* pts, int N, const float* smp, float* ptab, float* cst, float* lw,
                                      int m0, int S, const Cam& cam, float delta, float radius) {
    if ((S & 15) != 0) return false;
    constexpr int PD = Dim<DOF>::POSE;
    if (ptab != nullptr) {                               // the new samples' K[R|t], one row of the table each
        for (int s = threadIdx.x; s < S; s += NT) {
            float pose[PD], R[9], P[12];
#pragma unroll
            for (int c = 0; c < PD; ++c) pose[c] = smp[(m0 + s) * PD + c];
            pose_to_rot<DOF>(pose, R);
            make_proj(cam.k, R, pose, P);
#pragma unroll
            for (int c = 0; c < 12; ++c) ptab[s * 12 + c] = P[c];
        }
        __syncthreads();
    }                                                    // no table (no idle shared memory): columns computed per item
    if (cam.bounded) sweep_mma_items<DOF, true>(pts, N, smp, ptab, cst, lw, m0, S, cam, delta, radius);
    else sweep_mma_items<DOF, false>(pts, N, smp, ptab, cst, lw, m0, S, cam, delta, radius);
    return true;
}

// Column t of K [R | t] of one sample (what a lane's A fragments hold), straight from the pose.
template <int DOF>
__device__ __forceinline__ void proj_column(const float* pose, const float* K, int t, float (&col)[3]) {
    float R[9];
    pose_to_rot<DOF>(pose, R);
    const float v0 = t == 0 ? R[0] : (t == 1 ? R[1] : (t == 2 ? R[2] : pose[0]));
    const float v1 = t == 0 ? R[3] : (t == 1 ? R[4] : (t == 2 ? R[5] : pose[1]));
    const float v2 = t == 0 ? R[6] : (t == 1 ? R[7] : (t == 2 ? R[8] : pose[2]));
#pragma unroll
    for (int r = 0; r < 3; ++r) col[r] = K[r * 3 + 0] * v0 + K[r * 3 + 1] * v1 + K[r * 3 + 2] * v2;
}

template <int DOF, bool BOUNDED>
__device__ __forceinline__ void sweep_mma_items(const float* pts, int N, const float* smp, const float* ptab, float* cst,
                                                float* lw, int m0, int S, const Cam& cam, float delta, float radius) {
    constexpr int PD = Dim<DOF>::POSE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int npair = (N + 1) >> 1, npts = 2 * npair;    // even-padded point count (the pad carries zero weights)
    const int ntile = (npts + 7) >> 3, tmid = (ntile + 1) >> 1, T = S >> 4;
    for (int item = warp; item < 2 * T; item += NW) {
        const int ti = item >> 1, h = item & 1;
        const int s0 = ti * 16 + g, s1 = s0 + 8;
        float ah[3][2], al[3][2], c0v[3], c1v[3];
        if (ptab != nullptr) {
#pragma unroll
            for (int r = 0; r < 3; ++r) { c0v[r] = ptab[s0 * 12 + 4 * r + t]; c1v[r] = ptab[s1 * 12 + 4 * r + t]; }
        } else {
            float pose[PD];
#pragma unroll
            for (int c = 0; c < PD; ++c) pose[c] = smp[(m0 + s0) * PD + c];
            proj_column<DOF>(pose, cam.k, t, c0v);
#pragma unroll
            for (int c = 0; c < PD; ++c) pose[c] = smp[(m0 + s1) * PD + c];
            proj_column<DOF>(pose, cam.k, t, c1v);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            ah[r][0] = tf32_hi_bits(c0v[r]); al[r][0] = c0v[r] - ah[r][0];
            ah[r][1] = tf32_hi_bits(c1v[r]); al[r][1] = c1v[r] - ah[r][1];
        }
        // ONE A quad per projection row, [P_hi | P_lo]: the second MMA multiplies it by [X_lo ; 0], so the unwanted
        // P_lo X_lo term drops out without a second set of fragments
        float a1[3][4];
#pragma unroll
        for (int r = 0; r < 3; ++r) { a1[r][0] = ah[r][0]; a1[r][1] = ah[r][1]; a1[r][2] = al[r][0]; a1[r][3] = al[r][1]; }
        V2 acc0 = v2splat(0.f), acc1 = v2splat(0.f);
        const int t0 = h ? tmid : 0, t1 = h ? ntile : tmid;
        const int full1 = min(t1, npts >> 3);            // tiles [t0, full1) hold 8 real points each
        bool free_z = false;
#if defined(EPNP_SWEEP_NOCLAMP)
        if (radius >= 0.f) {
            // pose_depth_margin of the lane's two samples from the third projection row, whose four entries sit in
            // the four lanes of the group: |P[8..10]| and P[11] by two exchanges
            float q0 = (t < 3) ? c0v[2] * c0v[2] : 0.f, q1 = (t < 3) ? c1v[2] * c1v[2] : 0.f;
            q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
            q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
            const float tz0 = __shfl_sync(0xffffffffu, c0v[2], (lane & ~3) | 3), tz1 = __shfl_sync(0xffffffffu, c1v[2], (lane & ~3) | 3);
            const float r0 = sqrtf(q0) * radius, r1 = sqrtf(q1) * radius;
            const float m0z = (tz0 - r0) - cam.z_min - 1e-5f * (fabsf(tz0) + r0 + cam.z_min);
            const float m1z = (tz1 - r1) - cam.z_min - 1e-5f * (fabsf(tz1) + r1 + cam.z_min);
            free_z = __all_sync(0xffffffffu, fminf(m0z, m1z) >= 0.f);
        }
#endif
        // lane-private cursors: coordinate t (pad slot 14 holds the homogeneous 1) of point 8 tl + g, record of pair 4 tl + t
        const float* bp = pts + (g >> 1) * 16 + (t < 3 ? 2 * t : 14) + (g & 1) + t0 * 64;
        const float* up = pts + t * 16 + t0 * 64;
        if (free_z) {
            for (int tl = t0; tl < full1; ++tl, bp += 64, up += 64) {
                const float c = *bp;
                const float bh = tf32_hi_bits(c), bl = c - bh;
                float d[3][4];
    #pragma unroll
                for (int r = 0; r < 3; ++r) {
                    d[r][0] = d[r][1] = d[r][2] = d[r][3] = 0.f;
                    mma_tf32_16x8x8(d[r], a1[r], bh, bh);
                    mma_tf32_16x8x8(d[r], a1[r], bl, 0.f);
                }
                const float2 nu2 = *reinterpret_cast<const float2*>(up + 6);
                const float4 mid = *reinterpret_cast<const float4*>(up + 8);
                const float2 wv2 = *reinterpret_cast<const float2*>(up + 12);
                const V2 nu = v2(nu2.x, nu2.y), nv = v2(mid.x, mid.y), wu = v2(mid.z, mid.w), wv = v2(wv2.x, wv2.y);
                acc0 = pair_cost_tail<BOUNDED, false>(v2(d[0][0], d[0][1]), v2(d[1][0], d[1][1]), v2(d[2][0], d[2][1]), cam, delta, nu, nv, wu, wv, acc0, SweepRsqrt());
                acc1 = pair_cost_tail<BOUNDED, false>(v2(d[0][2], d[0][3]), v2(d[1][2], d[1][3]), v2(d[2][2], d[2][3]), cam, delta, nu, nv, wu, wv, acc1, SweepRsqrt());
            }
        } else {
            for (int tl = t0; tl < full1; ++tl, bp += 64, up += 64) {
                const float c = *bp;
                const float bh = tf32_hi_bits(c), bl = c - bh;
                float d[3][4];
    #pragma unroll
                for (int r = 0; r < 3; ++r) {
                    d[r][0] = d[r][1] = d[r][2] = d[r][3] = 0.f;
                    mma_tf32_16x8x8(d[r], a1[r], bh, bh);
                    mma_tf32_16x8x8(d[r], a1[r], bl, 0.f);
                }
                const float2 nu2 = *reinterpret_cast<const float2*>(up + 6);
                const float4 mid = *reinterpret_cast<const float4*>(up + 8);
                const float2 wv2 = *reinterpret_cast<const float2*>(up + 12);
                const V2 nu = v2(nu2.x, nu2.y), nv = v2(mid.x, mid.y), wu = v2(mid.z, mid.w), wv = v2(wv2.x, wv2.y);
                acc0 = pair_cost_tail<BOUNDED>(v2(d[0][0], d[0][1]), v2(d[1][0], d[1][1]), v2(d[2][0], d[2][1]), cam, delta, nu, nv, wu, wv, acc0, SweepRsqrt());
                acc1 = pair_cost_tail<BOUNDED>(v2(d[0][2], d[0][3]), v2(d[1][2], d[1][3]), v2(d[2][2], d[2][3]), cam, delta, nu, nv, wu, wv, acc1, SweepRsqrt());
            }
        }
        for (int tl = max(t0, full1); tl < t1; ++tl) {   // the (at most one) partial tile: clamped loads, masked weights
            const int pt = min(tl * 8 + g, npts - 1);
            const float c = (t < 3) ? pts[(pt >> 1) * 16 + 2 * t + (pt & 1)] : 1.0f;
            const float bh = tf32_hi_bits(c), bl = c - bh;
            float d[3][4];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                d[r][0] = d[r][1] = d[r][2] = d[r][3] = 0.f;
                mma_tf32_16x8x8(d[r], a1[r], bh, bh);
                mma_tf32_16x8x8(d[r], a1[r], bl, 0.f);
            }
            const int pair = tl * 4 + t;
            const bool valid = pair < npair;
            const float* rec = pts + min(pair, npair - 1) * 16;
            const V2 nu = v2(rec[6], rec[7]), nv = v2(rec[8], rec[9]);
            const V2 wu = valid ? v2(rec[10], rec[11]) : v2splat(0.f), wv = valid ? v2(rec[12], rec[13]) : v2splat(0.f);
            acc0 = pair_cost_tail<BOUNDED>(v2(d[0][0], d[0][1]), v2(d[1][0], d[1][1]), v2(d[2][0], d[2][1]), cam, delta, nu, nv, wu, wv, acc0, SweepRsqrt());
            acc1 = pair_cost_tail<BOUNDED>(v2(d[0][2], d[0][3]), v2(d[1][2], d[1][3]), v2(d[2][2], d[2][3]), cam, delta, nu, nv, wu, wv, acc1, SweepRsqrt());
        }
        float c0 = acc0.x + acc0.y, c1 = acc1.x + acc1.y;
        c0 += __shfl_xor_sync(0xffffffffu, c0, 1); c0 += __shfl_xor_sync(0xffffffffu, c0, 2);
        c1 += __shfl_xor_sync(0xffffffffu, c1, 1); c1 += __shfl_xor_sync(0xffffffffu, c1, 2);
        if (t == 0) {
            float* dst = h ? lw : cst;
            dst[m0 + s0] = c0; dst[m0 + s1] = c1;
        }
    }
}
#endif

#if defined(EPNP_AMIS_LSE)
// experiment (off by default): the mixture density of a sample is kept as ONE running log-sum-exp over the proposals
// seen so far instead of one log-density per (proposal, sample) -- (I - 1) * M floats less shared memory.
struct RunningLse {
    float top, acc;
    __device__ __forceinline__ void start(float lp) { top = lp; acc = 1.f; }
    __device__ __forceinline__ void add(float lp) {
        if (lp > top) { acc = fmaf(acc, expf(top - lp), 1.f); top = lp; }
        else acc += expf(lp - top);
    }
    __device__ __forceinline__ float value() const { return top + logf(acc); }
};
__device__ __forceinline__ float log_add_exp(float a, float b) {
    const float hi = fmaxf(a, b), lo = fminf(a, b);
    return (lo == -CUDART_INF_F) ? hi : hi + log1pf(expf(lo - hi));
}
#endif

// ------------------------------------------------------------------------------------------------
// AMIS loop for the resident object (6DoF).  sh.prop[0] must not be set yet; pose / cov are read from
// pose_opt[7] / cov[36] (shared or registers of thread 0 -- passed as shared pointers).
__device__ void amis_phase6(const KArgs& a, SmemHead<6>& sh, const float* pts4, float* smp, float* cst,
                            float* logp, float* lw, const Cam& cam, float delta, int obj,
                            const float* pose_opt, const float* cov EPNP_PTAB_PARAM) {
    const Params& p = a.p;
    const int tid = threadIdx.x;
    const int M = p.mc_samples, I = p.mc_iter, S = M / I;
    const bool injected = a.noise_n3 != nullptr;
    const int st = serial_thread(a);
    PH_DECL;

    if (tid == st) initial_fit6(pose_opt, cov, p.acg_dispersion, sh.prop[0]);
    PH_MARK(a, PH_INIT_FIT);
#if defined(EPNP_SWEEP_NOCLAMP)
    const float radius = object_radius(pts4, a.N, sh.red, 0);
#elif defined(EPNP_SWEEP_SPLIT)
    const float radius = -1.f;
#endif
    __syncthreads();

    for (int i = 0; i < I; ++i) {
        // ---- draw, cost, densities of the new samples (one sample per thread and pass)
        for (int s = tid; s < S; s += NT) {
            const int m = i * S + s;
            float n3[3], n4[4], chi2;
            if (injected) {
                const size_t g = (size_t)obj * M + m;
                n3[0] = __ldg(a.noise_n3 + g * 3); n3[1] = __ldg(a.noise_n3 + g * 3 + 1); n3[2] = __ldg(a.noise_n3 + g * 3 + 2);
                chi2 = __ldg(a.noise_chi2 + g);
                n4[0] = __ldg(a.noise_rot + g * 4); n4[1] = __ldg(a.noise_rot + g * 4 + 1);
                n4[2] = __ldg(a.noise_rot + g * 4 + 2); n4[3] = __ldg(a.noise_rot + g * 4 + 3);
            } else {
                draw_base_noise(a.seed, a.obj_offset + (uint32_t)obj, (uint32_t)m, n3, chi2, n4);
            }
            float q[7];
            proposal_draw6(sh.prop[i], n3, chi2, n4, q);
#pragma unroll
            for (int k = 0; k < 7; ++k) smp[m * 7 + k] = q[k];
            float* out = a.pose_samples + ((size_t)obj * M + m) * 7;
#pragma unroll
            for (int k = 0; k < 7; ++k) out[k] = q[k];
#if defined(EPNP_SWEEP_SPLIT)
            // cost: sweep_new_samples below, once every new sample of the iteration is in shared memory
#elif defined(EPNP_SWEEP_NOCLAMP)
            cst[m] = pose_cost<6>(pts4, a.N, q, cam, delta, radius);
#else
            cst[m] = pose_cost<6>(pts4, a.N, q, cam, delta);
#endif
#if defined(EPNP_AMIS_LSE)
            {
                RunningLse l;
                l.start(proposal_logpdf6(sh.prop[0], q));
                for (int j = 1; j <= i; ++j) l.add(proposal_logpdf6(sh.prop[j], q));
                logp[m] = l.value();
            }
#else
            for (int j = 0; j <= i; ++j) logp[j * M + m] = proposal_logpdf6(sh.prop[j], q);
#endif
        }
#if defined(EPNP_SWEEP_SPLIT)
        __syncthreads();
#if defined(EPNP_NO_LW)
        float* const half1 = a.logw + (size_t)obj * M;
#else
        float* const half1 = lw;
#endif
#if defined(EPNP_SWEEP_MMA)
        if (!sweep_new_samples_mma<6>(pts4, a.N, smp, ptab, cst, half1, i * S, S, cam, delta, radius))
#endif
        sweep_new_samples<6>(pts4, a.N, smp, cst, half1, i * S, S, cam, delta, radius);
#endif
        PH_MARK(a, PH_DRAW_SWEEP);
        // ---- the new proposal on all earlier samples
        for (int m = tid; m < i * S; m += NT) {
            float q[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) q[k] = smp[m * 7 + k];
#if defined(EPNP_AMIS_LSE)
            logp[m] = log_add_exp(logp[m], proposal_logpdf6(sh.prop[i], q));
#else
            logp[i * M + m] = proposal_logpdf6(sh.prop[i], q);
#endif
        }
        __syncthreads();
        PH_MARK(a, PH_LOGP_OLD);
        // ---- mixture density and log-weights of all samples so far
        const int n = (i + 1) * S;
        const float log_cnt = logf((float)(i + 1));
        float mx = -CUDART_INF_F;
        for (int m = tid; m < n; m += NT) {
#if defined(EPNP_AMIS_LSE)
            const float top = logp[m], acc = 1.f;               // logp[m] already is the log-sum-exp
#else
            float top = logp[m];
            for (int j = 1; j <= i; ++j) top = fmaxf(top, logp[j * M + m]);
            float acc = 0.f;
            for (int j = 0; j <= i; ++j) acc += expf(logp[j * M + m] - top);
#endif
#if defined(EPNP_SWEEP_SPLIT)
            float cm = cst[m];
#if defined(EPNP_NO_LW)
            if (m >= i * S) { cm += a.logw[(size_t)obj * M + m]; cst[m] = cm; }   // second half parked in the output slot
#else
            if (m >= i * S) { cm += lw[m]; cst[m] = cm; }       // fold the second point half of a new sample in
#endif
            const float v = -cm - ((top + logf(acc)) - log_cnt);
#else
            const float v = -cst[m] - ((top + logf(acc)) - log_cnt);
#endif
#if defined(EPNP_NO_LW)
            if (i == I - 1) a.logw[(size_t)obj * M + m] = v;    // n == M on the last iteration: every slot gets its value
#else
            lw[m] = v;
#endif
            mx = fmaxf(mx, v);
        }
        PH_MARK(a, PH_WEIGHTS);
        if (i == I - 1) {
#if !defined(EPNP_NO_LW)
            for (int m = tid; m < M; m += NT) a.logw[(size_t)obj * M + m] = lw[m];
#endif
            PH_MARK(a, PH_OUTPUT);
            break;
        }
        // ---- refit proposal i+1 to the weighted samples (estimate_params, epropnp.py:317-342).
        // Four block reductions, one barrier each:
        //   A  max of the log-weights
        //   B  e = exp(lw - max): sum e, sum e t, and ACG fixed-point iteration 1 (Lambda_0 = I, so
        //      M = q.q) -- the normalisation of the weights cancels in Lambda
        //   C  translation covariance about the mean (+ ACG iteration 2)
        //   D+ remaining ACG iterations
        mx = block_max(mx, sh.red, 0);
        float lam10[10];
        float mean[3], inv_sum;
        {
            float acc[15];
#pragma unroll
            for (int r = 0; r < 15; ++r) acc[r] = 0.f;
            for (int m = tid; m < n; m += NT) {
                const float e = LW_E(m);
                LW_E_STORE(m, e);
                const float* s7 = smp + m * 7;
                acc[0] += e;
                acc[1] = fmaf(e, s7[0], acc[1]); acc[2] = fmaf(e, s7[1], acc[2]); acc[3] = fmaf(e, s7[2], acc[3]);
                const float* q = s7 + 3;
                const float mq = fmaxf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3], p.amis_eps);
                const float wm = e / mq;
                acc[4] += wm;
                int idx = 5;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = r; c < 4; ++c) { acc[idx] = fmaf(wm * q[r], q[c], acc[idx]); ++idx; }
            }
            block_sum<15>(acc, sh.red, 1);
            inv_sum = 1.0f / acc[0];
            mean[0] = acc[1] * inv_sum; mean[1] = acc[2] * inv_sum; mean[2] = acc[3] * inv_sum;
            const float inv0 = 1.0f / acc[4];
#pragma unroll
            for (int r = 0; r < 10; ++r) lam10[r] = acc[5 + r] * inv0;
            lam10[0] += p.amis_eps; lam10[4] += p.amis_eps; lam10[7] += p.amis_eps; lam10[9] += p.amis_eps;
        }
        if (p.acg_mle_iter == 0) {          // degenerate configuration: Lambda stays the identity
#pragma unroll
            for (int r = 0; r < 10; ++r) lam10[r] = 0.f;
            lam10[0] = lam10[4] = lam10[7] = lam10[9] = 1.f;
        }
        float tc[6];
        {
            const bool more = p.acg_mle_iter >= 2;
            float lam_inv[16];
            if (more) acg_scatter_inverse(lam10, lam_inv);
            float acc[17];
#pragma unroll
            for (int r = 0; r < 17; ++r) acc[r] = 0.f;
            for (int m = tid; m < n; m += NT) {
                const float w = LW_W(m);                                 // normalised softmax weight
                LW_W_STORE(m, w);
                const float* s7 = smp + m * 7;
                const float d0 = s7[0] - mean[0], d1 = s7[1] - mean[1], d2 = s7[2] - mean[2];
                acc[11] = fmaf(w * d0, d0, acc[11]); acc[12] = fmaf(w * d0, d1, acc[12]); acc[13] = fmaf(w * d0, d2, acc[13]);
                acc[14] = fmaf(w * d1, d1, acc[14]); acc[15] = fmaf(w * d1, d2, acc[15]); acc[16] = fmaf(w * d2, d2, acc[16]);
                if (more) {
                    const float* q = s7 + 3;
                    const float wm = w / fmaxf(quad4(lam_inv, q), p.amis_eps);
                    acc[0] += wm;
                    int idx = 1;
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = r; c < 4; ++c) { acc[idx] = fmaf(wm * q[r], q[c], acc[idx]); ++idx; }
                }
            }
            block_sum<17>(acc, sh.red, 0);
#pragma unroll
            for (int r = 0; r < 6; ++r) tc[r] = acc[11 + r];
            if (more) {
                const float inv0 = 1.0f / acc[0];
#pragma unroll
                for (int r = 0; r < 10; ++r) lam10[r] = acc[1 + r] * inv0;
                lam10[0] += p.amis_eps; lam10[4] += p.amis_eps; lam10[7] += p.amis_eps; lam10[9] += p.amis_eps;
            }
        }
        for (int itr = 2; itr < p.acg_mle_iter; ++itr) {
            float lam_inv[16];
            acg_scatter_inverse(lam10, lam_inv);
            float acc[11];
#pragma unroll
            for (int r = 0; r < 11; ++r) acc[r] = 0.f;
            for (int m = tid; m < n; m += NT) {
                const float* q = smp + m * 7 + 3;
#if defined(EPNP_NO_LW)
                const float wm = LW_W(m) / fmaxf(quad4(lam_inv, q), p.amis_eps);
#else
                const float wm = lw[m] / fmaxf(quad4(lam_inv, q), p.amis_eps);
#endif
                acc[0] += wm;
                int idx = 1;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = r; c < 4; ++c) { acc[idx] = fmaf(wm * q[r], q[c], acc[idx]); ++idx; }
            }
            block_sum<11>(acc, sh.red, (itr + 1) & 1);
            const float inv0 = 1.0f / acc[0];
#pragma unroll
            for (int r = 0; r < 10; ++r) lam10[r] = acc[1 + r] * inv0;
            lam10[0] += p.amis_eps; lam10[4] += p.amis_eps; lam10[7] += p.amis_eps; lam10[9] += p.amis_eps;
        }
        PH_MARK(a, PH_REFIT_SUMS);
        if (tid == st) {
            refit_finish6(mean, tc, lam10, p.acg_dispersion, sh.prop[i + 1]);
        }
        PH_MARK(a, PH_REFIT_FINISH);
        __syncthreads();
    }
    if (a.proposals && tid < I) {
        float* o = a.proposals + ((size_t)obj * I + tid) * PROP_FLOATS;
        const Proposal6& pr = sh.prop[tid];
        o[0] = pr.mu[0]; o[1] = pr.mu[1]; o[2] = pr.mu[2];
#pragma unroll
        for (int r = 0; r < 6; ++r) o[3 + r] = pr.lt[r];
#pragma unroll
        for (int r = 0; r < 10; ++r) o[9 + r] = pr.lr[r];
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// AMIS loop for the resident object, 4DoF (EProPnP4DoF, epropnp.py:199-260): same skeleton as amis_phase6
// with the yaw proposal 0.75 von Mises + 0.25 uniform.  Injected noise: noise_rot (B, M) holds the yaw
// draws themselves (the reference samples them with numpy on the host, distributions.py:61-72, so there
// is no base noise to replay).
__device__ void amis_phase4(const KArgs& a, SmemHead<4>& sh, const float* pts4, float* smp, float* cst,
                            float* logp, float* lw, const Cam& cam, float delta, int obj,
                            const float* pose_opt, const float* cov EPNP_PTAB_PARAM) {
    const Params& p = a.p;
    const int tid = threadIdx.x;
    const int M = p.mc_samples, I = p.mc_iter, S = M / I;
    const bool injected = a.noise_n3 != nullptr;
    const int st = serial_thread(a);
    PH_DECL;

    if (tid == st) initial_fit4(pose_opt, cov, p.amis_eps, sh.prop4[0]);
    PH_MARK(a, PH_INIT_FIT);
#if defined(EPNP_SWEEP_NOCLAMP)
    const float radius = object_radius(pts4, a.N, sh.red, 0);
#elif defined(EPNP_SWEEP_SPLIT)
    const float radius = -1.f;
#endif
    __syncthreads();

    for (int i = 0; i < I; ++i) {
        for (int s = tid; s < S; s += NT) {
            const int m = i * S + s;
            float n3[3], chi2, q[4];
            if (injected) {
                const size_t g = (size_t)obj * M + m;
                n3[0] = __ldg(a.noise_n3 + g * 3); n3[1] = __ldg(a.noise_n3 + g * 3 + 1); n3[2] = __ldg(a.noise_n3 + g * 3 + 2);
                chi2 = __ldg(a.noise_chi2 + g);
                q[3] = __ldg(a.noise_rot + g);
            } else {
                draw_base_noise_t(a.seed, a.obj_offset + (uint32_t)obj, (uint32_t)m, n3, chi2);
                q[3] = draw_yaw(a.seed, a.obj_offset + (uint32_t)obj, (uint32_t)m, s, S, sh.prop4[i].mode, sh.prop4[i].kappa);
            }
            draw_translation(sh.prop4[i].mu, sh.prop4[i].lt, n3, chi2, q);
            float* out = a.pose_samples + ((size_t)obj * M + m) * 4;
#pragma unroll
            for (int k = 0; k < 4; ++k) { smp[m * 4 + k] = q[k]; out[k] = q[k]; }
#if defined(EPNP_SWEEP_SPLIT)
            // cost: sweep_new_samples below, once every new sample of the iteration is in shared memory
#elif defined(EPNP_SWEEP_NOCLAMP)
            cst[m] = pose_cost<4>(pts4, a.N, q, cam, delta, radius);
#else
            cst[m] = pose_cost<4>(pts4, a.N, q, cam, delta);
#endif
#if defined(EPNP_AMIS_LSE)
            {
                RunningLse l;
                l.start(proposal_logpdf4(sh.prop4[0], q));
                for (int j = 1; j <= i; ++j) l.add(proposal_logpdf4(sh.prop4[j], q));
                logp[m] = l.value();
            }
#else
            for (int j = 0; j <= i; ++j) logp[j * M + m] = proposal_logpdf4(sh.prop4[j], q);
#endif
        }
#if defined(EPNP_SWEEP_SPLIT)
        __syncthreads();
#if defined(EPNP_NO_LW)
        float* const half1 = a.logw + (size_t)obj * M;
#else
        float* const half1 = lw;
#endif
#if defined(EPNP_SWEEP_MMA)
        if (!sweep_new_samples_mma<4>(pts4, a.N, smp, ptab, cst, half1, i * S, S, cam, delta, radius))
#endif
        sweep_new_samples<4>(pts4, a.N, smp, cst, half1, i * S, S, cam, delta, radius);
#endif
        PH_MARK(a, PH_DRAW_SWEEP);
#if defined(EPNP_AMIS_LSE)
        for (int m = tid; m < i * S; m += NT) logp[m] = log_add_exp(logp[m], proposal_logpdf4(sh.prop4[i], smp + m * 4));
#else
        for (int m = tid; m < i * S; m += NT) logp[i * M + m] = proposal_logpdf4(sh.prop4[i], smp + m * 4);
#endif
        __syncthreads();
        PH_MARK(a, PH_LOGP_OLD);
        const int n = (i + 1) * S;
        const float log_cnt = logf((float)(i + 1));
        float mx = -CUDART_INF_F;
        for (int m = tid; m < n; m += NT) {
#if defined(EPNP_AMIS_LSE)
            const float top = logp[m], acc = 1.f;               // logp[m] already is the log-sum-exp
#else
            float top = logp[m];
            for (int j = 1; j <= i; ++j) top = fmaxf(top, logp[j * M + m]);
            float acc = 0.f;
            for (int j = 0; j <= i; ++j) acc += expf(logp[j * M + m] - top);
#endif
#if defined(EPNP_SWEEP_SPLIT)
            float cm = cst[m];
#if defined(EPNP_NO_LW)
            if (m >= i * S) { cm += a.logw[(size_t)obj * M + m]; cst[m] = cm; }   // second half parked in the output slot
#else
            if (m >= i * S) { cm += lw[m]; cst[m] = cm; }       // fold the second point half of a new sample in
#endif
            const float v = -cm - ((top + logf(acc)) - log_cnt);
#else
            const float v = -cst[m] - ((top + logf(acc)) - log_cnt);
#endif
#if defined(EPNP_NO_LW)
            if (i == I - 1) a.logw[(size_t)obj * M + m] = v;    // n == M on the last iteration: every slot gets its value
#else
            lw[m] = v;
#endif
            mx = fmaxf(mx, v);
        }
        PH_MARK(a, PH_WEIGHTS);
        if (i == I - 1) {
#if !defined(EPNP_NO_LW)
            for (int m = tid; m < M; m += NT) a.logw[(size_t)obj * M + m] = lw[m];
#endif
            PH_MARK(a, PH_OUTPUT);
            break;
        }
        // ---- refit (estimate_params, epropnp.py:232-260): A max, B sums of e, e t, e sin, e cos, C covariance
        mx = block_max(mx, sh.red, 0);
        float accB[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int m = tid; m < n; m += NT) {
            const float e = LW_E(m);
            LW_E_STORE(m, e);
            const float* s4 = smp + m * 4;
            float sn, cs;
            sincosf(s4[3], &sn, &cs);
            accB[0] += e;
            accB[1] = fmaf(e, s4[0], accB[1]); accB[2] = fmaf(e, s4[1], accB[2]); accB[3] = fmaf(e, s4[2], accB[3]);
            accB[4] = fmaf(e, sn, accB[4]); accB[5] = fmaf(e, cs, accB[5]);
        }
        block_sum<6>(accB, sh.red, 1);
        const float inv_sum = 1.0f / accB[0];
        const float mean[3] = {accB[1] * inv_sum, accB[2] * inv_sum, accB[3] * inv_sum};
        float tc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int m = tid; m < n; m += NT) {
            const float w = LW_W(m);
            const float* s4 = smp + m * 4;
            const float d0 = s4[0] - mean[0], d1 = s4[1] - mean[1], d2 = s4[2] - mean[2];
            tc[0] = fmaf(w * d0, d0, tc[0]); tc[1] = fmaf(w * d0, d1, tc[1]); tc[2] = fmaf(w * d0, d2, tc[2]);
            tc[3] = fmaf(w * d1, d1, tc[3]); tc[4] = fmaf(w * d1, d2, tc[4]); tc[5] = fmaf(w * d2, d2, tc[5]);
        }
        block_sum<6>(tc, sh.red, 0);
        PH_MARK(a, PH_REFIT_SUMS);
        if (tid == st) refit_finish4(mean, tc, accB[4] * inv_sum, accB[5] * inv_sum, p.amis_eps, sh.prop4[i + 1]);
        PH_MARK(a, PH_REFIT_FINISH);
        __syncthreads();
    }
    if (a.proposals && tid < I) {       // (B, I, 19): mu3, Lt6, mode, kappa, 0...
        float* o = a.proposals + ((size_t)obj * I + tid) * PROP_FLOATS;
        const Proposal4& pr = sh.prop4[tid];
        o[0] = pr.mu[0]; o[1] = pr.mu[1]; o[2] = pr.mu[2];
#pragma unroll
        for (int r = 0; r < 6; ++r) o[3 + r] = pr.lt[r];
        o[9] = pr.mode; o[10] = pr.kappa;
#pragma unroll
        for (int r = 11; r < PROP_FLOATS; ++r) o[r] = 0.f;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Kernels
#if !defined(EPNP_CTAS_PER_SM)
#define EPNP_CTAS_PER_SM 4              // build option: 5 needs <= 96 registers and <= 44 KB of shared memory per CTA
#endif
template <int DOF, bool DO_LM, bool DO_AMIS>
__global__ void __launch_bounds__(NT, EPNP_CTAS_PER_SM) solve_kernel(const KArgs a) {
    EPNP_DYN_SMEM(unsigned char, smem_raw, 128);
    SmemHead<DOF>& sh = *reinterpret_cast<SmemHead<DOF>*>(smem_raw);
    float* dyn = reinterpret_cast<float*>(smem_raw);
#if defined(EPNP_ALIAS_STAGE)
    // one object per CTA (grid == B): nothing is prefetched during the solve, the ring may live in the sample buffer
    const SmemPlan pl = plan_smem<DOF>(a.N, a.p.mc_samples, a.p.mc_iter, DO_AMIS, gridDim.x >= (unsigned)a.B);
#else
    const SmemPlan pl = plan_smem<DOF>(a.N, a.p.mc_samples, a.p.mc_iter, DO_AMIS);
#endif
    float* pts4 = dyn + pl.pts;

    Loader ld(a, sh.bar, dyn + pl.stage);
    ld.prologue();
    for (int it = 0; it < ld.n_my; ++it) {
        const int obj = (int)blockIdx.x + it * (int)gridDim.x;
        PH_DECL;
        ld.load_object(it, obj, pts4);
        const Cam cam = load_cam(a, obj);
        const float delta = __ldg(a.delta + obj);
        PH_MARK(a, PH_LOAD);
        if constexpr (DO_LM) {
            lm_phase<DOF>(a, sh, pts4, cam, delta, obj, DO_AMIS || a.pose_cov != nullptr);
        }
        if constexpr (DO_AMIS) {
            if constexpr (!DO_LM) {
                if (threadIdx.x < Dim<DOF>::POSE) sh.lm.pose[threadIdx.x] = __ldg(a.pose_opt_in + (size_t)obj * Dim<DOF>::POSE + threadIdx.x);
                if (threadIdx.x < DOF * DOF) sh.cov[threadIdx.x] = __ldg(a.pose_cov_in + (size_t)obj * DOF * DOF + threadIdx.x);
                __syncthreads();
            }
#if defined(EPNP_SWEEP_MMA)
            // table of the current iteration's K[R|t] (12 floats per new sample): the staging ring, which is idle once
            // the object is packed -- unless this CTA prefetches its next object (persistent grid) or the ring is
            // aliased into the sample buffer, in which case the tensor-pipe sweep is not used
            float* ptab = nullptr;
            if (gridDim.x >= (unsigned)a.B && pl.stage != pl.smp && 12 * (a.p.mc_samples / a.p.mc_iter) <= 2 * STAGE_FLOATS)
                ptab = dyn + pl.stage;
#endif
            if constexpr (DOF == 6)
                amis_phase6(a, sh, pts4, dyn + pl.smp, dyn + pl.cost, dyn + pl.logp, dyn + pl.lw, cam, delta, obj,
                            sh.lm.pose, sh.cov EPNP_PTAB_ARG(ptab));
            else
                amis_phase4(a, sh, pts4, dyn + pl.smp, dyn + pl.cost, dyn + pl.logp, dyn + pl.lw, cam, delta, obj,
                            sh.lm.pose, sh.cov EPNP_PTAB_ARG(ptab));
        }
    }
}

// Fused solve + all-gather over peer memory (multi-GPU, one node).  Same solve as solve_kernel<DOF, true, true>; when an
// object is finished, its CTA also stores the object's pose and its M log-weights into row (obj_offset + obj) of the
// full-batch result buffers of up to EPNP_MAX_PEERS other GPUs (pointers into their memory, mapped through CUDA IPC):
// plain st.global over NVLink, 2 KB + 28 B per object and peer, issued object by object underneath the other CTAs'
// math.  No gather kernel and no copy afterwards; the caller only needs a rendezvous before reading (sharded.PushGather).
// The local outputs (a.pose_opt, a.logw) are normally the local slice of this rank's own full-batch buffers.
constexpr int EPNP_MAX_PEERS = 8;
struct PushArgs {
    float* logw[EPNP_MAX_PEERS];            // (B_total, M) on each peer
    float* pose[EPNP_MAX_PEERS];            // (B_total, D) on each peer
    int n;
};

template <int DOF>
__global__ void __launch_bounds__(NT, EPNP_CTAS_PER_SM) solve_push_kernel(const KArgs a, const PushArgs push) {
    EPNP_DYN_SMEM(unsigned char, smem_raw, 128);
    SmemHead<DOF>& sh = *reinterpret_cast<SmemHead<DOF>*>(smem_raw);
    float* dyn = reinterpret_cast<float*>(smem_raw);
    constexpr int PD = Dim<DOF>::POSE;
#if defined(EPNP_ALIAS_STAGE)
    const SmemPlan pl = plan_smem<DOF>(a.N, a.p.mc_samples, a.p.mc_iter, true, gridDim.x >= (unsigned)a.B);
#else
    const SmemPlan pl = plan_smem<DOF>(a.N, a.p.mc_samples, a.p.mc_iter, true);
#endif
    float* pts4 = dyn + pl.pts;
    const int M = a.p.mc_samples;
    Loader ld(a, sh.bar, dyn + pl.stage);
    ld.prologue();
    for (int it = 0; it < ld.n_my; ++it) {
        const int obj = (int)blockIdx.x + it * (int)gridDim.x;
        ld.load_object(it, obj, pts4);
        const Cam cam = load_cam(a, obj);
        const float delta = __ldg(a.delta + obj);
        lm_phase<DOF>(a, sh, pts4, cam, delta, obj, true);
#if defined(EPNP_SWEEP_MMA)
        float* ptab = nullptr;
        if (gridDim.x >= (unsigned)a.B && pl.stage != pl.smp && 12 * (a.p.mc_samples / a.p.mc_iter) <= 2 * STAGE_FLOATS)
            ptab = dyn + pl.stage;
#endif
        if constexpr (DOF == 6)
            amis_phase6(a, sh, pts4, dyn + pl.smp, dyn + pl.cost, dyn + pl.logp, dyn + pl.lw, cam, delta, obj,
                        sh.lm.pose, sh.cov EPNP_PTAB_ARG(ptab));
        else
            amis_phase4(a, sh, pts4, dyn + pl.smp, dyn + pl.cost, dyn + pl.logp, dyn + pl.lw, cam, delta, obj,
                        sh.lm.pose, sh.cov EPNP_PTAB_ARG(ptab));
        // ---- push: the object's outputs, as this CTA wrote them, to the same global row on every peer
        __syncthreads();                                        // the CTA's own global stores are visible to all its threads
        const size_t row = (size_t)a.obj_offset + (size_t)obj;
        const float* src = a.logw + (size_t)obj * M;
        if ((M & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
            for (int q = threadIdx.x; q < M / 4; q += NT) {
                const float4 v = reinterpret_cast<const float4*>(src)[q];
                for (int r = 0; r < push.n; ++r) {
                    float* dst = push.logw[r] + row * M;
                    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) reinterpret_cast<float4*>(dst)[q] = v;
                    else { dst[4 * q] = v.x; dst[4 * q + 1] = v.y; dst[4 * q + 2] = v.z; dst[4 * q + 3] = v.w; }
                }
            }
        } else {
            for (int m = threadIdx.x; m < M; m += NT) {
                const float v = src[m];
                for (int r = 0; r < push.n; ++r) push.logw[r][row * M + m] = v;
            }
        }
        if (threadIdx.x < PD) {
            const float v = a.pose_opt[(size_t)obj * PD + threadIdx.x];
            for (int r = 0; r < push.n; ++r) push.pose[r][row * PD + threadIdx.x] = v;
        }
    }
}

// cost of S poses per object: poses (S, B, D) -> cost (S, B)
template <int DOF>
__global__ void __launch_bounds__(NT, 4) cost_kernel(const KArgs a) {
    EPNP_DYN_SMEM(unsigned char, smem_raw, 128);
    SmemHead<DOF>& sh = *reinterpret_cast<SmemHead<DOF>*>(smem_raw);
    float* dyn = reinterpret_cast<float*>(smem_raw);
    const SmemPlan pl = plan_smem<DOF>(a.N, 0, 0, false);
    float* pts4 = dyn + pl.pts;
    constexpr int PD = Dim<DOF>::POSE;
    Loader ld(a, sh.bar, dyn + pl.stage);
    ld.prologue();
    for (int it = 0; it < ld.n_my; ++it) {
        const int obj = (int)blockIdx.x + it * (int)gridDim.x;
        ld.load_object(it, obj, pts4);
        const Cam cam = load_cam(a, obj);
        const float delta = __ldg(a.delta + obj);
        for (int s = threadIdx.x; s < a.S_eval; s += NT) {
            float pose[PD];
#pragma unroll
            for (int k = 0; k < PD; ++k) pose[k] = __ldg(a.poses + ((size_t)s * a.B + obj) * PD + k);
            a.cost_out[(size_t)s * a.B + obj] = pose_cost<DOF>(pts4, a.N, pose, cam, delta);
        }
        __syncthreads();        // pts4 is overwritten by the next object
    }
}

// ------------------------------------------------------------------------------------------------
// Random-sample LM initialiser (RSLMSolver.solve, levenberg_marquardt.py:300-353) for one object per CTA:
// thread <-> hypothesis.  A thread runs the whole LM / GN iteration of its hypothesis serially over that
// hypothesis' n sampled correspondences (read out of the object's resident pair records: the (P*B, n, .) gathered
// copies, the P-fold repeated cameras and the P*B tiny solves of the reference do not exist), scores the result
// on ALL N points with the same sweep the AMIS loop uses, and the CTA keeps the cheapest hypothesis.
struct RslmArgs {
    KArgs k;                    // correspondences, camera, bounds, delta, B, N, LM parameters
    const int* inds;            // (P, B, n) indices of the sampled correspondences, within the object
    const float* start;         // (P, B, D) starting poses
    float* pose_best;           // (B, D)
    float* cost_best;           // (B)
    float* pose_all;            // [opt] (P, B, D)
    float* cost_all;            // [opt] (P, B)
    int P, n;
};

// Order of torch.min over the hypotheses (levenberg_marquardt.py:350): a NaN cost wins (min propagates NaN), then the
// smaller cost, then the earlier hypothesis.
__device__ __forceinline__ bool cheaper_hypothesis(float c, int h, float wc, int wh) {
    const bool cn = (c != c), wn = (wc != wc);
    if (cn != wn) return cn;
    if (cn) return h < wh;
    return c < wc || (c == wc && h < wh);
}

template <int DOF, bool CLIP>
__device__ __forceinline__ void eval_subset(const float* pts, const int* idx, int n, const float* pose, const Cam& cam,
                                            float delta, float huber_eps, float* acc) {
    constexpr int NV = Dim<DOF>::NV;
    float R[9];
    pose_to_rot<DOF>(pose, R);
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
    for (int i = 0; i < n; ++i) {
        const int j = __ldg(idx + i);
        const float* q = pts + (j >> 1) * 16 + (j & 1);
        point_normal_eq<DOF, CLIP>(R, pose, cam, delta, huber_eps, q[0], q[2], q[4], -q[6], -q[8], q[10], q[12], acc);
    }
}

template <int DOF>
__global__ void __launch_bounds__(NT, 2) rslm_kernel(const RslmArgs r) {
    const KArgs& a = r.k;
    EPNP_DYN_SMEM(unsigned char, smem_raw, 128);
    SmemHead<DOF>& sh = *reinterpret_cast<SmemHead<DOF>*>(smem_raw);
    float* dyn = reinterpret_cast<float*>(smem_raw);
    const SmemPlan pl = plan_smem<DOF>(a.N, 0, 0, false);
    float* pts = dyn + pl.pts;
    constexpr int PD = Dim<DOF>::POSE;
    const Params& p = a.p;
    const int tid = threadIdx.x;
    float* best_cost = sh.red;                                  // [NT]
    int* best_hyp = reinterpret_cast<int*>(sh.red + NT);        // [NT]
    Loader ld(a, sh.bar, dyn + pl.stage);
    ld.prologue();
    for (int it = 0; it < ld.n_my; ++it) {
        const int obj = (int)blockIdx.x + it * (int)gridDim.x;
        ld.load_object(it, obj, pts);
        const Cam cam = load_cam(a, obj);
        const float delta = __ldg(a.delta + obj);
        float my_cost = CUDART_INF_F, my_pose[PD];
        int my_hyp = -1;
        for (int h = tid; h < r.P; h += NT) {
            const int* idx = r.inds + ((size_t)h * a.B + obj) * r.n;
            LMState<DOF> s;
#pragma unroll
            for (int i = 0; i < PD; ++i) s.pose[i] = __ldg(r.start + ((size_t)h * a.B + obj) * PD + i);
            s.radius = p.initial_radius;
            s.shrink = 2.0f;
            float acc[Dim<DOF>::NV];
            if (!p.fast_mode) {
                eval_subset<DOF, true>(pts, idx, r.n, s.pose, cam, delta, p.huber_eps, acc);
                lm_adopt<DOF>(s, acc);
                if (p.lm_iter > 0) lm_propose<DOF>(s, p);
                for (int k = 0; k < p.lm_iter; ++k) {
                    eval_subset<DOF, true>(pts, idx, r.n, s.pose_new, cam, delta, p.huber_eps, acc);
                    lm_update<DOF>(s, acc, p);
                    if (k + 1 < p.lm_iter) lm_propose<DOF>(s, p);
                }
            } else {
                for (int k = 0; k < p.lm_iter; ++k) {
                    eval_subset<DOF, false>(pts, idx, r.n, s.pose, cam, delta, p.huber_eps, acc);
                    gn_advance<DOF>(s.pose, acc, p.eps, s.pose);
                }
            }
            const float c = pose_cost<DOF>(pts, a.N, s.pose, cam, delta);       // score on the full set
            if (r.pose_all) {
#pragma unroll
                for (int i = 0; i < PD; ++i) r.pose_all[((size_t)h * a.B + obj) * PD + i] = s.pose[i];
            }
            if (r.cost_all) r.cost_all[(size_t)h * a.B + obj] = c;
            if (my_hyp < 0 || cheaper_hypothesis(c, h, my_cost, my_hyp)) {
                my_cost = c; my_hyp = h;
#pragma unroll
                for (int i = 0; i < PD; ++i) my_pose[i] = s.pose[i];
            }
        }
        best_cost[tid] = my_cost;
        best_hyp[tid] = my_hyp;
        __syncthreads();
        if (tid == 0) {
            int w = -1;
            for (int t = 0; t < NT; ++t) {
                if (best_hyp[t] < 0) continue;
                if (w < 0 || cheaper_hypothesis(best_cost[t], best_hyp[t], best_cost[w], best_hyp[w])) w = t;
            }
            best_hyp[0] = w;                                    // the winning THREAD (it still holds the pose)
        }
        __syncthreads();
        if (tid == best_hyp[0]) {
#pragma unroll
            for (int i = 0; i < PD; ++i) r.pose_best[(size_t)obj * PD + i] = my_pose[i];
            r.cost_best[obj] = my_cost;
        }
        __syncthreads();            // pts and the reduction arrays are reused by the next object
    }
}

// ------------------------------------------------------------------------------------------------
// Backward of pose_opt_plus = pose (+) gn_step(pose) w.r.t. the correspondences and the Huber delta (the
// derivative-regularisation branch of training, LMSolver.forward :66-68 with autograd on).  One CTA per object:
// re-evaluate the normal equations, let the serial lane form step, dL/dstep (through pose_add) and
// v = -(H + eps I)^-1 dL/dstep, then thread <-> correspondence with forward-mode duals (pnp::gn_step_point_backward).
struct GnBwArgs {
    KArgs k;                    // correspondences, camera, bounds, delta, poses = pose (B, D), p.eps / huber_eps / z_min
    const float* gplus;         // (B, D)  dL/d pose_opt_plus
    float *gx3d, *gx2d, *gw2d;  // [opt] (B, N, 3|2|2)
    float* gdelta;              // [opt] (B)
};

template <int DOF>
__global__ void __launch_bounds__(NT, 2) gn_plus_backward_kernel(const GnBwArgs g) {
    const KArgs& a = g.k;
    EPNP_DYN_SMEM(unsigned char, smem_raw, 128);
    SmemHead<DOF>& sh = *reinterpret_cast<SmemHead<DOF>*>(smem_raw);
    float* dyn = reinterpret_cast<float*>(smem_raw);
    const SmemPlan pl = plan_smem<DOF>(a.N, 0, 0, false);
    float* pts = dyn + pl.pts;
    constexpr int PD = Dim<DOF>::POSE, NA = Dim<DOF>::NA;
    const int tid = threadIdx.x;
    Loader ld(a, sh.bar, dyn + pl.stage);
    ld.prologue();
    for (int it = 0; it < ld.n_my; ++it) {
        const int obj = (int)blockIdx.x + it * (int)gridDim.x;
        ld.load_object(it, obj, pts);
        const Cam cam = load_cam(a, obj);
        const float delta = __ldg(a.delta + obj);
        if (tid < PD) sh.lm.pose[tid] = __ldg(a.poses + (size_t)obj * PD + tid);
        __syncthreads();
        eval_normal_eq<DOF, true>(pts, a.N, sh.lm.pose, cam, delta, a.p.huber_eps, sh.red, sh.ev);
        float* step = sh.cov;               // [DOF]
        float* vvec = sh.cov + DOF;         // [DOF]
        if (tid == 0) {
            float add[DOF], st[DOF], sbar[DOF], vv[DOF], gout[PD], pose[PD];
#pragma unroll
            for (int i = 0; i < DOF; ++i) add[i] = a.p.eps;
#pragma unroll
            for (int i = 0; i < PD; ++i) { pose[i] = sh.lm.pose[i]; gout[i] = __ldg(g.gplus + (size_t)obj * PD + i); }
            damped_step_refined<DOF>(sh.ev, sh.ev + NA, add, st);
            pose_add_backward<DOF>(pose, st, gout, sbar);
            damped_step_refined<DOF>(sh.ev, sbar, add, vv);
#pragma unroll
            for (int i = 0; i < DOF; ++i) { step[i] = st[i]; vvec[i] = vv[i]; }
        }
        __syncthreads();
        float R[9], t[3], sv[DOF], vv[DOF];
        {
            float ps[PD];
#pragma unroll
            for (int i = 0; i < PD; ++i) ps[i] = sh.lm.pose[i];
            pose_to_rot<DOF>(ps, R);
            t[0] = ps[0]; t[1] = ps[1]; t[2] = ps[2];
#pragma unroll
            for (int i = 0; i < DOF; ++i) { sv[i] = step[i]; vv[i] = vvec[i]; }
        }
        float gd[1] = {0.f};
        for (int n = tid; n < a.N; n += NT) {
            const float* q = pts + (n >> 1) * 16 + (n & 1);
            float gr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            gn_step_point_backward<DOF>(R, t, cam, delta, a.p.huber_eps, q[0], q[2], q[4], -q[6], -q[8], q[10], q[12], vv, sv, gr);
            const size_t o = (size_t)obj * a.N + n;
            if (g.gx3d) { g.gx3d[o * 3] = gr[0]; g.gx3d[o * 3 + 1] = gr[1]; g.gx3d[o * 3 + 2] = gr[2]; }
            if (g.gx2d) { g.gx2d[o * 2] = gr[3]; g.gx2d[o * 2 + 1] = gr[4]; }
            if (g.gw2d) { g.gw2d[o * 2] = gr[5]; g.gw2d[o * 2 + 1] = gr[6]; }
            gd[0] += gr[7];
        }
        block_sum<1>(gd, sh.red, 0);
        if (tid == 0 && g.gdelta) g.gdelta[obj] = gd[0];
        __syncthreads();            // pts, pose and the step vectors are reused by the next object
    }
}

// residual / Jacobian / cost written out per point (API parity with evaluate_pnp's out_* tensors)
template <int DOF>
__global__ void __launch_bounds__(NT) evaluate_full_kernel(const KArgs a, float* residual, float* jac, float* cost,
                                                         int clip, float huber_eps) {
    __shared__ float red[NW];
    const int obj = blockIdx.x;
    constexpr int PD = Dim<DOF>::POSE;
    const Cam cam = load_cam(a, obj);
    const float delta = __ldg(a.delta + obj);
    float pose[PD], R[9];
#pragma unroll
    for (int k = 0; k < PD; ++k) pose[k] = __ldg(a.poses + (size_t)obj * PD + k);
    pose_to_rot<DOF>(pose, R);
    float csum = 0.f;
    for (int n = threadIdx.x; n < a.N; n += NT) {
        const size_t g = (size_t)obj * a.N + n;
        float r[2], j[2 * DOF];
        csum += point_residual_jac<DOF>(R, pose, cam, delta, huber_eps, clip != 0,
                                        a.x3d[g * 3], a.x3d[g * 3 + 1], a.x3d[g * 3 + 2],
                                        a.x2d[g * 2], a.x2d[g * 2 + 1], a.w2d[g * 2], a.w2d[g * 2 + 1], r, j);
        if (residual) { residual[g * 2] = r[0]; residual[g * 2 + 1] = r[1]; }
        if (jac) {
#pragma unroll
            for (int k = 0; k < 2 * DOF; ++k) jac[g * 2 * DOF + k] = j[k];
        }
    }
    if (cost) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = csum;
        __syncthreads();
        if (threadIdx.x == 0) cost[obj] = (red[0] + red[1]) + (red[2] + red[3]);
    }
}

// ------------------------------------------------------------------------------------------------
// Backward of the Monte-Carlo cost: for every object, sum over its poses p (pose_init and all AMIS samples)
// of g[p] * d cost(pose p; x3d, x2d, w2d, delta) / d(x3d, x2d, w2d, delta).
// One CTA per object; thread <-> correspondence (up to BW_PPT per thread, accumulators in registers); the
// poses' pre-multiplied projections K[R|t] and upstream gradients are staged once in shared memory and read
// as broadcasts, so the (pose, point) loop touches no global memory.
constexpr int BW_PPT = 4;               // correspondences per thread and tile
constexpr int BW_POSE_TILE = 1024;      // poses staged per tile (13 floats each)

#if defined(EPNP_SIMT_EMUL)
struct FastRsqrt { float operator()(float x) const { return 1.0f / sqrtf(x); } };
#else
struct FastRsqrt {
    __device__ __forceinline__ float operator()(float x) const { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
};
#endif

// Reverse mode of the cost for TWO correspondences at once (packed fp32x2; same math per half as
// pnp::point_cost_backward, with s = s2 * rsqrt(s2) and delta / s = delta * rsqrt(s2) from one MUFU.RSQ).
// g[0..2] += dL/d(X,Y,Z), g[3..4] += dL/d(u,v), g[5..6] += dL/d(wu,wv); gd += dL/d delta (both halves).
template <bool BOUNDED>
__device__ __forceinline__ void pair_cost_backward(const float (&P)[12], const Cam& cam, float delta, float gj,
                                                   float2 X, float2 Y, float2 Z, float2 nu, float2 nv, float2 wu,
                                                   float2 wv, float2 (&g)[7], float2& gd) {
    const float2 xh = __ffma2_rn(splat(P[0]), X, __ffma2_rn(splat(P[1]), Y, __ffma2_rn(splat(P[2]), Z, splat(P[3]))));
    const float2 yh = __ffma2_rn(splat(P[4]), X, __ffma2_rn(splat(P[5]), Y, __ffma2_rn(splat(P[6]), Z, splat(P[7]))));
    const float2 zh = __ffma2_rn(splat(P[8]), X, __ffma2_rn(splat(P[9]), Y, __ffma2_rn(splat(P[10]), Z, splat(P[11]))));
    const float2 iz = make_float2(FastRcp()(fmaxf(zh.x, cam.z_min)), FastRcp()(fmaxf(zh.y, cam.z_min)));
    const float2 px = __fmul2_rn(xh, iz), py = __fmul2_rn(yh, iz);
    float2 pxc = px, pyc = py;
    if (BOUNDED) {
        pxc.x = fminf(fmaxf(px.x, cam.lbx), cam.ubx); pxc.y = fminf(fmaxf(px.y, cam.lbx), cam.ubx);
        pyc.x = fminf(fmaxf(py.x, cam.lby), cam.uby); pyc.y = fminf(fmaxf(py.y, cam.lby), cam.uby);
    }
    const float2 ex = __fadd2_rn(pxc, nu), ey = __fadd2_rn(pyc, nv);
    const float2 rx = __fmul2_rn(ex, wu), ry = __fmul2_rn(ey, wv);
    const float2 s2 = __ffma2_rn(rx, rx, __fmul2_rn(ry, ry));
    const float2 rs = make_float2(FastRsqrt()(fmaxf(s2.x, 1e-30f)), FastRsqrt()(fmaxf(s2.y, 1e-30f)));
    const float2 s = __fmul2_rn(s2, rs);
    const bool in0 = s.x <= delta, in1 = s.y <= delta;
    const float2 ko = __fmul2_rn(splat(gj * delta), rs);                 // outlier: g * delta / s
    const float2 k = make_float2(in0 ? gj : ko.x, in1 ? gj : ko.y);
    const float2 grx = __fmul2_rn(k, rx), gry = __fmul2_rn(k, ry);
    g[5] = __ffma2_rn(grx, ex, g[5]);
    g[6] = __ffma2_rn(gry, ey, g[6]);
    const float2 gex = __fmul2_rn(grx, wu), gey = __fmul2_rn(gry, wv);
    g[3] = __ffma2_rn(gex, splat(-1.0f), g[3]);
    g[4] = __ffma2_rn(gey, splat(-1.0f), g[4]);
    float2 gpx = gex, gpy = gey;
    if (BOUNDED) {
        gpx.x = (pxc.x == px.x) ? gex.x : 0.f; gpx.y = (pxc.y == px.y) ? gex.y : 0.f;
        gpy.x = (pyc.x == py.x) ? gey.x : 0.f; gpy.y = (pyc.y == py.y) ? gey.y : 0.f;
    }
    const float2 gxh = __fmul2_rn(gpx, iz), gyh = __fmul2_rn(gpy, iz);
    const float2 tz = __ffma2_rn(gxh, px, __fmul2_rn(gyh, py));
    const float2 gzh = make_float2(zh.x >= cam.z_min ? -tz.x : 0.f, zh.y >= cam.z_min ? -tz.y : 0.f);
    g[0] = __ffma2_rn(splat(P[0]), gxh, __ffma2_rn(splat(P[4]), gyh, __ffma2_rn(splat(P[8]), gzh, g[0])));
    g[1] = __ffma2_rn(splat(P[1]), gxh, __ffma2_rn(splat(P[5]), gyh, __ffma2_rn(splat(P[9]), gzh, g[1])));
    g[2] = __ffma2_rn(splat(P[2]), gxh, __ffma2_rn(splat(P[6]), gyh, __ffma2_rn(splat(P[10]), gzh, g[2])));
    const float2 over = __fadd2_rn(s, splat(-delta));
    gd = __ffma2_rn(make_float2(in0 ? 0.f : over.x, in1 ? 0.f : over.y), splat(gj), gd);
}

struct BwArgs {
    const float *x3d, *x2d, *w2d, *cam, *lb, *ub, *delta;
    const float *poses_a, *grad_a;      // (B, PA, D), (B, PA)
    const float *poses_b, *grad_b;      // [opt] (B, PB, D), (B, PB)
    float *gx3d, *gx2d, *gw2d, *gdelta;
    int B, N, PA, PB;
    float z_min;
};

template <int DOF>
__global__ void __launch_bounds__(NT, 4) cost_backward_kernel(const BwArgs a) {
    EPNP_DYN_SMEM(float, bw_smem, 16);
    float* Pm = bw_smem;                                 // [tile][12]
    float* gs = bw_smem + BW_POSE_TILE * 12;             // [tile]
    float* red = gs + BW_POSE_TILE;                      // 2 * NW * 32
    constexpr int PD = Dim<DOF>::POSE;
    const int tid = threadIdx.x;
    const int P_total = a.PA + a.PB;
    for (int obj = blockIdx.x; obj < a.B; obj += gridDim.x) {
        KArgs ka{};
        ka.cam = a.cam; ka.lb = a.lb; ka.ub = a.ub; ka.p.z_min = a.z_min;
        const Cam cam = load_cam(ka, obj);
        const float delta = __ldg(a.delta + obj);
        float2 gd2 = make_float2(0.f, 0.f);
        for (int base = 0; base < a.N; base += NT * BW_PPT) {
            // BW_PPT = 4 correspondences per thread = 2 packed pairs: pair p holds points base + (2p)*NT + tid (.x)
            // and base + (2p+1)*NT + tid (.y); out-of-range slots get zero weights (exactly zero contribution)
            constexpr int NP = BW_PPT / 2;
            float2 X[NP], Y[NP], Z[NP], nu[NP], nv[NP], wu[NP], wv[NP], g[NP][7];
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                float v7[2][7];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int n = base + (2 * k + h) * NT + tid;
                    const bool ok = n < a.N;
                    const size_t q = (size_t)obj * a.N + (ok ? n : 0);
                    v7[h][0] = __ldg(a.x3d + q * 3); v7[h][1] = __ldg(a.x3d + q * 3 + 1); v7[h][2] = __ldg(a.x3d + q * 3 + 2);
                    v7[h][3] = -__ldg(a.x2d + q * 2); v7[h][4] = -__ldg(a.x2d + q * 2 + 1);
                    v7[h][5] = ok ? __ldg(a.w2d + q * 2) : 0.f; v7[h][6] = ok ? __ldg(a.w2d + q * 2 + 1) : 0.f;
                }
                X[k] = make_float2(v7[0][0], v7[1][0]); Y[k] = make_float2(v7[0][1], v7[1][1]); Z[k] = make_float2(v7[0][2], v7[1][2]);
                nu[k] = make_float2(v7[0][3], v7[1][3]); nv[k] = make_float2(v7[0][4], v7[1][4]);
                wu[k] = make_float2(v7[0][5], v7[1][5]); wv[k] = make_float2(v7[0][6], v7[1][6]);
#pragma unroll
                for (int c = 0; c < 7; ++c) g[k][c] = make_float2(0.f, 0.f);
            }
            for (int p0 = 0; p0 < P_total; p0 += BW_POSE_TILE) {
                const int np = min(BW_POSE_TILE, P_total - p0);
                __syncthreads();                          // previous tile fully consumed
                for (int j = tid; j < np; j += NT) {
                    const int pi = p0 + j;
                    const bool in_a = pi < a.PA;
                    const float* src = in_a ? a.poses_a + ((size_t)obj * a.PA + pi) * PD
                                            : a.poses_b + ((size_t)obj * a.PB + (pi - a.PA)) * PD;
                    float pose[PD], R[9], Pj[12];
#pragma unroll
                    for (int c = 0; c < PD; ++c) pose[c] = __ldg(src + c);
                    pose_to_rot<DOF>(pose, R);
                    make_proj(cam.k, R, pose, Pj);
#pragma unroll
                    for (int c = 0; c < 12; ++c) Pm[j * 12 + c] = Pj[c];
                    gs[j] = in_a ? __ldg(a.grad_a + (size_t)obj * a.PA + pi) : __ldg(a.grad_b + (size_t)obj * a.PB + (pi - a.PA));
                }
                __syncthreads();
                for (int j = 0; j < np; ++j) {
                    const float4 r0 = reinterpret_cast<const float4*>(Pm)[3 * j];
                    const float4 r1 = reinterpret_cast<const float4*>(Pm)[3 * j + 1];
                    const float4 r2 = reinterpret_cast<const float4*>(Pm)[3 * j + 2];
                    const float Pj[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
                    const float gj = gs[j];
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        if (cam.bounded) pair_cost_backward<true>(Pj, cam, delta, gj, X[k], Y[k], Z[k], nu[k], nv[k], wu[k], wv[k], g[k], gd2);
                        else pair_cost_backward<false>(Pj, cam, delta, gj, X[k], Y[k], Z[k], nu[k], nv[k], wu[k], wv[k], g[k], gd2);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < NP; ++k) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int n = base + (2 * k + h) * NT + tid;
                    if (n < a.N) {
                        const size_t q = (size_t)obj * a.N + n;
                        auto pick = [&](int c) { return h == 0 ? g[k][c].x : g[k][c].y; };
                        if (a.gx3d) { a.gx3d[q * 3] = pick(0); a.gx3d[q * 3 + 1] = pick(1); a.gx3d[q * 3 + 2] = pick(2); }
                        if (a.gx2d) { a.gx2d[q * 2] = pick(3); a.gx2d[q * 2 + 1] = pick(4); }
                        if (a.gw2d) { a.gw2d[q * 2] = pick(5); a.gw2d[q * 2 + 1] = pick(6); }
                    }
                }
            }
        }
        const float gdelta = gd2.x + gd2.y;
        // padded (n >= N) lanes carry zero weights: their residual is 0 -> inlier -> no delta contribution
        float gd[1] = {gdelta};
        __syncthreads();
        block_sum<1>(gd, red, 0);
        if (tid == 0 && a.gdelta) a.gdelta[obj] = gd[0];
    }
}

// AdaptiveHuberPnPCost.set_param: delta = mean(w2d) * sqrt(var_x + var_y) * relative_delta
__global__ void __launch_bounds__(NT) adaptive_delta_kernel(const float* x2d, const float* w2d, float rel,
                                                          float* delta, int N) {
    __shared__ float red[2 * NW * 32];
    const int obj = blockIdx.x;
    const float2* x = reinterpret_cast<const float2*>(x2d) + (size_t)obj * N;
    const float2* w = reinterpret_cast<const float2*>(w2d) + (size_t)obj * N;
    float s[3] = {0.f, 0.f, 0.f};
    for (int n = threadIdx.x; n < N; n += NT) {
        const float2 a = x[n], b = w[n];
        s[0] += a.x; s[1] += a.y; s[2] += b.x + b.y;
    }
    block_sum<3>(s, red, 0);
    const float mx = s[0] / N, my = s[1] / N, mw = s[2] / (2.f * N);
    float v[2] = {0.f, 0.f};
    for (int n = threadIdx.x; n < N; n += NT) {
        const float2 a = x[n];
        v[0] = fmaf(a.x - mx, a.x - mx, v[0]); v[1] = fmaf(a.y - my, a.y - my, v[1]);
    }
    block_sum<2>(v, red, 1);
    if (threadIdx.x == 0) delta[obj] = mw * sqrtf((v[0] + v[1]) / (float)(N - 1)) * rel;
}

// Epilogue on the AMIS outputs (the step after the path).  One CTA per object reads the
// object's M log-weights (and, for the score, the x / z components of its M pose samples) once and writes
//   lse[b]      = logsumexp_m logw[b, m]                        (Monte-Carlo pose loss: loss_pred)
//   loss[b]     = cost_target[b] + lse[b], NaN -> 0              (the per-object loss before its mean)
//   weights     = softmax_m logw[b, :]                           (Det: pose_sample_logweights.softmax(dim=0))
//   score_te[b] = sum_m weights[m] * clamp((2.5 - log2 |(x,z)_m - (x,z)_opt|) / 4, 0, 1)   (Det 'te' MC score)
// with torch's conventions at the infinities (all -inf -> lse -inf, any NaN -> NaN).
template <int PD>
__global__ void __launch_bounds__(NT) mc_epilogue_kernel(const float* logw, const float* samples, const float* pose_opt,
                                                       const float* cost_target, float* lse_out, float* loss_out,
                                                       float* weights, float* score_out, int M) {
    __shared__ float red[2 * NW * 32];
    const int obj = blockIdx.x, tid = threadIdx.x;
    const float* l = logw + (size_t)obj * M;
    float mx = -CUDART_INF_F;
    bool any_nan = false;
    for (int m = tid; m < M; m += NT) { const float v = __ldg(l + m); mx = fmaxf(mx, v); any_nan |= (v != v); }
    mx = block_max(mx, red, 0);
    const float ref = (fabsf(mx) == CUDART_INF_F) ? 0.f : mx;          // torch.logsumexp: an infinite max is not subtracted
    const bool want_score = samples != nullptr && pose_opt != nullptr && score_out != nullptr;
    float ox = 0.f, oz = 0.f;
    if (want_score) { ox = __ldg(pose_opt + (size_t)obj * PD); oz = __ldg(pose_opt + (size_t)obj * PD + 2); }
    float acc[3] = {0.f, 0.f, any_nan ? 1.f : 0.f};                    // sum e, sum e * score, NaN seen
    for (int m = tid; m < M; m += NT) {
        const float e = expf(__ldg(l + m) - ref);
        acc[0] += e;
        if (want_score) {
            const float* sp = samples + ((size_t)obj * M + m) * PD;
            const float dx = __ldg(sp) - ox, dz = __ldg(sp + 2) - oz;
            const float dev = sqrtf(fmaf(dx, dx, dz * dz));
            const float sc = fminf(fmaxf((2.5f - log2f(dev)) * 0.25f, 0.f), 1.f);
            acc[1] = fmaf(e, sc, acc[1]);
        }
    }
    block_sum<3>(acc, red, 1);
    const float nan_in = acc[2] > 0.f ? CUDART_NAN_F : 0.f;            // fmaxf drops NaNs: put them back
    const float lse = logf(acc[0]) + ref + nan_in;
    const float inv = (mx == CUDART_INF_F) ? CUDART_NAN_F : 1.0f / acc[0];   // softmax: inf - inf poisons the whole object
    if (tid == 0) {
        if (lse_out) lse_out[obj] = lse;
        if (loss_out) {
            const float v = (cost_target ? __ldg(cost_target + obj) : 0.f) + lse;
            loss_out[obj] = (v != v) ? 0.f : v;
        }
        if (want_score) score_out[obj] = acc[1] * inv + nan_in;
    }
    if (weights) {
        float* w = weights + (size_t)obj * M;
        for (int m = tid; m < M; m += NT) w[m] = expf(__ldg(l + m) - ref) * inv;
    }
}

// Backward of lse (and of the NaN -> 0 mask): grad_logw[b, m] = g[b] * exp(logw[b, m] - lse[b]); g[b] == 0 gives exact zeros
// (masked objects have lse = NaN).
__global__ void __launch_bounds__(NT) mc_lse_backward_kernel(const float* logw, const float* lse, const float* g,
                                                           float* grad_logw, int M) {
    const int obj = blockIdx.x;
    const float gb = __ldg(g + obj), ls = __ldg(lse + obj);
    const float* l = logw + (size_t)obj * M;
    float* o = grad_logw + (size_t)obj * M;
    for (int m = threadIdx.x; m < M; m += NT) o[m] = (gb == 0.f) ? 0.f : gb * expf(__ldg(l + m) - ls);
}

// ------------------------------------------------------------------------------------------------
// Host side
int cuda_fail(cudaError_t e) { g_last_cuda_error = (int)e; return EPNP_ERR_CUDA; }

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_common(const KArgs& a) {
    if (!a.x3d || !a.x2d || !a.w2d || !a.cam || !a.delta) return EPNP_ERR_BAD_ARG;
    if ((a.lb == nullptr) != (a.ub == nullptr)) return EPNP_ERR_BAD_ARG;
    if (a.B < 0 || a.N <= 0) return EPNP_ERR_BAD_ARG;
    if (a.p.dof != 4 && a.p.dof != 6) return EPNP_ERR_BAD_ARG;
    return EPNP_OK;
}

template <class Kern, class... Extra>
int launch_persistent(Kern kern, KArgs& a, int smem_bytes, cudaStream_t stream, int smem_bytes_single = 0, Extra... extra) {
    (void)smem_bytes_single;            // EPNP_ALIAS_STAGE: dynamic shared memory when every CTA solves one object
    if (a.B == 0) return EPNP_OK;
    if ((size_t)smem_bytes > SMEM_LIMIT) return EPNP_ERR_TOO_MANY_POINTS;
    a.use_tma = (a.N % 4 == 0) && aligned16(a.x3d) && aligned16(a.x2d) && aligned16(a.w2d);
    int dev = 0, sms = 0, occ = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e);
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return cuda_fail(e);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return cuda_fail(e);
#if defined(EPNP_CTAS_PER_SM) && EPNP_CTAS_PER_SM > 4
    // five / six resident CTAs need (nearly) the whole 228 KB of the SM as shared memory: ask for the full carve-out
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return cuda_fail(e);
#endif
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem_bytes);
    if (e != cudaSuccess) return cuda_fail(e);
    if (occ < 1) return EPNP_ERR_TOO_MANY_POINTS;
    a.num_sms = sms;
    // Work distribution.  Default: ONE object per CTA (grid = B) and the hardware scheduler hands CTAs to SMs as
    // slots free up -- measured 5 % faster than a persistent grid at B = 4096 (2.93 vs 2.79 M objects/s): the CTAs'
    // serial and parallel phases de-synchronise and there is no lock-step tail, which outweighs losing the
    // cross-object TMA prefetch (~1.5 %); it also lets a concurrently enqueued kernel of another stream (the NCCL
    // gather of the previous batch in the multi-GPU pipeline) start at once instead of waiting for a resident grid
    // to drain.  EPNP_MAX_OBJECTS_PER_CTA = k (environment, read per call) sets the cap; 0 = fully persistent
    // (grid = resident slots, objects strided over CTAs, next object's chunks prefetched during the current solve).
    const int slots = sms * occ;
    int rounds = 1;
    if (const char* env = std::getenv("EPNP_MAX_OBJECTS_PER_CTA")) {
        const int k = std::atoi(env);
        const int persistent_rounds = (a.B + slots - 1) / slots;
        rounds = (k <= 0) ? persistent_rounds : (k < persistent_rounds ? k : persistent_rounds);
    }
    const int grid = (a.B + rounds - 1) / rounds;
#if defined(EPNP_ALIAS_STAGE)
    if (rounds == 1 && smem_bytes_single > 0) smem_bytes = smem_bytes_single;      // the kernel sees grid == B too
#endif
    EPNP_LAUNCH(kern, grid, NT, smem_bytes, stream, a, extra...);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e);
    return EPNP_OK;
}

unsigned long long* g_prof_buffer = nullptr;     // set by epnp_debug_set_phase_buffer (profiling build)

// Objects the device solves at once with the fused kernel: SMs x resident CTAs per SM (one CTA per object).  Used by
// the host-buffer entry point to cut the batch at whole waves.  0 when it cannot be determined.
template <class Kern>
int resident_objects(Kern kern, int smem_bytes) {
    int dev = 0, sms = 0, occ = 0;
    if ((size_t)smem_bytes > SMEM_LIMIT) return 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem_bytes) != cudaSuccess) return 0;
    return sms * occ;
}

int check_amis_params(const Params& p) {
    if (p.mc_iter <= 0 || p.mc_iter > MAX_ITER || p.mc_samples <= 0 || p.mc_samples % p.mc_iter != 0) return EPNP_ERR_BAD_ARG;
    if (p.acg_mle_iter < 0) return EPNP_ERR_BAD_ARG;
    return EPNP_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

#ifdef EPNP_PHASE_TIMERS
// profiling build only (not declared in the public header): device buffer of PH_COUNT uint64 counters
void epnp_debug_set_phase_buffer(unsigned long long* dev_buf) { g_prof_buffer = dev_buf; }
#endif

int epnp_abi_version(void) { return EPNP_ABI_VERSION; }

int epnp_last_cuda_error(void) { return g_last_cuda_error; }

const char* epnp_error_string(int code) {
    switch (code) {
        case EPNP_OK: return "ok";
        case EPNP_ERR_BAD_ARG: return "bad argument";
        case EPNP_ERR_TOO_MANY_POINTS: return "correspondence set (and sample buffers) exceed 227 KB of shared memory";
        case EPNP_ERR_UNSUPPORTED: return "combination not supported by this build";
        case EPNP_ERR_CUDA: return "CUDA runtime error (see epnp_last_cuda_error)";
        case EPNP_ERR_NO_DEVICE: return "no CUDA device";
        default: return "unknown error";
    }
}

void epnp_default_params(EpnpParams* p, int dof) {
    p->dof = dof; p->lm_iter = 10; p->fast_mode = 0; p->z_min = 0.1f;
    p->min_lm_diagonal = 1e-6f; p->max_lm_diagonal = 1e32f; p->min_relative_decrease = 1e-3f;
    p->initial_radius = 30.0f; p->max_radius = 1e16f; p->eps = 1e-5f; p->huber_eps = 1e-10f;
    p->mc_samples = 512; p->mc_iter = 4; p->amis_eps = 1e-5f; p->acg_mle_iter = 3; p->acg_dispersion = 1e-3f;
}

int epnp_max_points(int dof, int mc_samples, int mc_iter) {
    const bool amis = mc_samples > 0;
    int lo = 0, hi = 1 << 16;
    while (lo + 4 <= hi) {                   // largest multiple of 4 that fits
        const int mid = ((lo + hi) / 2) / 4 * 4;
        if (mid == lo) break;
        const int bytes = (dof == 6) ? plan_smem<6>(mid, mc_samples, mc_iter, amis).total_bytes
                                     : plan_smem<4>(mid, mc_samples, mc_iter, amis).total_bytes;
        if ((size_t)bytes <= SMEM_LIMIT) lo = mid; else hi = mid;
    }
    return lo;
}

int epnp_adaptive_delta_f32(const float* x2d, const float* w2d, float relative_delta, float* delta, int B, int N,
                            void* stream) {
    if (!x2d || !w2d || !delta || B < 0 || N <= 0) return EPNP_ERR_BAD_ARG;
    if (B == 0) return EPNP_OK;
    EPNP_LAUNCH(adaptive_delta_kernel, B, NT, 0, (cudaStream_t)stream, x2d, w2d, relative_delta, delta, N);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? EPNP_OK : cuda_fail(e);
}

int epnp_mc_epilogue_f32(const float* logw, const float* pose_samples, const float* pose_opt, const float* cost_target,
                         float* lse, float* loss, float* weights, float* score_te, int B, int M, int dof, void* stream) {
    if (!logw || B < 0 || M <= 0 || (dof != 4 && dof != 6)) return EPNP_ERR_BAD_ARG;
    if (!lse && !loss && !weights && !score_te) return EPNP_ERR_BAD_ARG;
    if (score_te && (!pose_samples || !pose_opt)) return EPNP_ERR_BAD_ARG;
    if (B == 0) return EPNP_OK;
    if (dof == 6)
        EPNP_LAUNCH(mc_epilogue_kernel<7>, B, NT, 0, (cudaStream_t)stream, logw, pose_samples, pose_opt, cost_target, lse, loss,
                    weights, score_te, M);
    else
        EPNP_LAUNCH(mc_epilogue_kernel<4>, B, NT, 0, (cudaStream_t)stream, logw, pose_samples, pose_opt, cost_target, lse, loss,
                    weights, score_te, M);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? EPNP_OK : cuda_fail(e);
}

int epnp_mc_lse_backward_f32(const float* logw, const float* lse, const float* grad_lse, float* grad_logw, int B, int M,
                             void* stream) {
    if (!logw || !lse || !grad_lse || !grad_logw || B < 0 || M <= 0) return EPNP_ERR_BAD_ARG;
    if (B == 0) return EPNP_OK;
    EPNP_LAUNCH(mc_lse_backward_kernel, B, NT, 0, (cudaStream_t)stream, logw, lse, grad_lse, grad_logw, M);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? EPNP_OK : cuda_fail(e);
}

int epnp_evaluate_cost_f32(const float* x3d, const float* x2d, const float* w2d, const float* cam_mats,
                           const float* lb, const float* ub, const float* delta, const float* poses, float* cost,
                           int S, int B, int N, int dof, float z_min, void* stream) {
    KArgs a{};
    a.x3d = x3d; a.x2d = x2d; a.w2d = w2d; a.cam = cam_mats; a.lb = lb; a.ub = ub; a.delta = delta;
    a.poses = poses; a.cost_out = cost; a.S_eval = S; a.B = B; a.N = N;
    epnp_default_params(&a.p, dof);
    a.p.z_min = z_min;
    int rc = check_common(a);
    if (rc != EPNP_OK) return rc;
    if (!poses || !cost || S < 0) return EPNP_ERR_BAD_ARG;
    if (S == 0) return EPNP_OK;
    if (dof == 6) return launch_persistent(cost_kernel<6>, a, plan_smem<6>(N, 0, 0, false).total_bytes, (cudaStream_t)stream);
    return launch_persistent(cost_kernel<4>, a, plan_smem<4>(N, 0, 0, false).total_bytes, (cudaStream_t)stream);
}

int epnp_evaluate_f32(const float* x3d, const float* x2d, const float* w2d, const float* cam_mats,
                      const float* lb, const float* ub, const float* delta, const float* pose,
                      float* residual, float* jac, float* cost, int clip_jac,
                      int B, int N, int dof, float z_min, float huber_eps, void* stream) {
    KArgs a{};
    a.x3d = x3d; a.x2d = x2d; a.w2d = w2d; a.cam = cam_mats; a.lb = lb; a.ub = ub; a.delta = delta;
    a.poses = pose; a.B = B; a.N = N;
    epnp_default_params(&a.p, dof);
    a.p.z_min = z_min;
    int rc = check_common(a);
    if (rc != EPNP_OK) return rc;
    if (!pose) return EPNP_ERR_BAD_ARG;
    if (B == 0) return EPNP_OK;
    if (dof == 6) EPNP_LAUNCH(evaluate_full_kernel<6>, B, NT, 0, (cudaStream_t)stream, a, residual, jac, cost, clip_jac, huber_eps);
    else EPNP_LAUNCH(evaluate_full_kernel<4>, B, NT, 0, (cudaStream_t)stream, a, residual, jac, cost, clip_jac, huber_eps);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? EPNP_OK : cuda_fail(e);
}

int epnp_lm_solve_f32(const float* x3d, const float* x2d, const float* w2d, const float* cam_mats,
                      const float* lb, const float* ub, const float* delta, const float* pose_init,
                      float* pose_opt, float* pose_cov, float* cost, float* pose_opt_plus, float* cost_init,
                      int B, int N, const EpnpParams* p, void* stream) {
    if (!p) return EPNP_ERR_BAD_ARG;
    KArgs a{};
    a.x3d = x3d; a.x2d = x2d; a.w2d = w2d; a.cam = cam_mats; a.lb = lb; a.ub = ub; a.delta = delta;
    a.pose_init = pose_init; a.pose_opt = pose_opt; a.pose_cov = pose_cov; a.cost = cost;
    a.pose_plus = pose_opt_plus; a.cost_init = cost_init; a.B = B; a.N = N; a.p = *p;
    int rc = check_common(a);
    if (rc != EPNP_OK) return rc;
    if (!pose_init || !pose_opt || p->lm_iter < 0) return EPNP_ERR_BAD_ARG;
    if (p->dof == 6) return launch_persistent(solve_kernel<6, true, false>, a, plan_smem<6>(N, 0, 0, false).total_bytes, (cudaStream_t)stream);
    return launch_persistent(solve_kernel<4, true, false>, a, plan_smem<4>(N, 0, 0, false).total_bytes, (cudaStream_t)stream);
}

int epnp_gn_plus_backward_f32(const float* x3d, const float* x2d, const float* w2d, const float* cam_mats,
                              const float* lb, const float* ub, const float* delta, const float* pose,
                              const float* grad_pose_plus, float* grad_x3d, float* grad_x2d, float* grad_w2d,
                              float* grad_delta, int B, int N, int dof, float z_min, float eps, float huber_eps,
                              void* stream) {
    GnBwArgs g{};
    KArgs& a = g.k;
    a.x3d = x3d; a.x2d = x2d; a.w2d = w2d; a.cam = cam_mats; a.lb = lb; a.ub = ub; a.delta = delta;
    a.poses = pose; a.B = B; a.N = N;
    epnp_default_params(&a.p, dof);
    a.p.z_min = z_min; a.p.eps = eps; a.p.huber_eps = huber_eps;
    g.gplus = grad_pose_plus; g.gx3d = grad_x3d; g.gx2d = grad_x2d; g.gw2d = grad_w2d; g.gdelta = grad_delta;
    int rc = check_common(a);
    if (rc != EPNP_OK) return rc;
    if (!pose || !grad_pose_plus) return EPNP_ERR_BAD_ARG;
    if (B == 0) return EPNP_OK;
    const int smem_bytes = (dof == 6) ? plan_smem<6>(N, 0, 0, false).total_bytes : plan_smem<4>(N, 0, 0, false).total_bytes;
    if ((size_t)smem_bytes > SMEM_LIMIT) return EPNP_ERR_TOO_MANY_POINTS;
    a.use_tma = (N % 4 == 0) && aligned16(x3d) && aligned16(x2d) && aligned16(w2d);
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e);
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return cuda_fail(e);
    a.num_sms = sms;
    if (dof == 6) {
        e = cudaFuncSetAttribute(gn_plus_backward_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return cuda_fail(e);
        EPNP_LAUNCH(gn_plus_backward_kernel<6>, B, NT, smem_bytes, (cudaStream_t)stream, g);
    } else {
        e = cudaFuncSetAttribute(gn_plus_backward_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return cuda_fail(e);
        EPNP_LAUNCH(gn_plus_backward_kernel<4>, B, NT, smem_bytes, (cudaStream_t)stream, g);
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? EPNP_OK : cuda_fail(e);
}

int epnp_rslm_f32(const float* x3d, const float* x2d, const float* w2d, const float* cam_mats,
                  const float* lb, const float* ub, const float* delta, const int* inds, const float* start,
                  float* pose_best, float* cost_best, float* pose_all, float* cost_all,
                  int P, int n, int B, int N, const EpnpParams* p, void* stream) {
    if (!p) return EPNP_ERR_BAD_ARG;
    RslmArgs r{};
    KArgs& a = r.k;
    a.x3d = x3d; a.x2d = x2d; a.w2d = w2d; a.cam = cam_mats; a.lb = lb; a.ub = ub; a.delta = delta;
    a.B = B; a.N = N; a.p = *p;
    r.inds = inds; r.start = start; r.pose_best = pose_best; r.cost_best = cost_best;
    r.pose_all = pose_all; r.cost_all = cost_all; r.P = P; r.n = n;
    int rc = check_common(a);
    if (rc != EPNP_OK) return rc;
    if (!inds || !start || !pose_best || !cost_best || P <= 0 || n <= 0 || p->lm_iter < 0) return EPNP_ERR_BAD_ARG;
    if (B == 0) return EPNP_OK;
    const int smem_bytes = (p->dof == 6) ? plan_smem<6>(N, 0, 0, false).total_bytes : plan_smem<4>(N, 0, 0, false).total_bytes;
    if ((size_t)smem_bytes > SMEM_LIMIT) return EPNP_ERR_TOO_MANY_POINTS;
    a.use_tma = (N % 4 == 0) && aligned16(x3d) && aligned16(x2d) && aligned16(w2d);
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e);
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return cuda_fail(e);
    a.num_sms = sms;
    if (p->dof == 6) {
        e = cudaFuncSetAttribute(rslm_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return cuda_fail(e);
        EPNP_LAUNCH(rslm_kernel<6>, B, NT, smem_bytes, (cudaStream_t)stream, r);
    } else {
        e = cudaFuncSetAttribute(rslm_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return cuda_fail(e);
        EPNP_LAUNCH(rslm_kernel<4>, B, NT, smem_bytes, (cudaStream_t)stream, r);
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? EPNP_OK : cuda_fail(e);
}

int epnp_amis_f32(const float* x3d, const float* x2d, const float* w2d, const float* cam_mats,
                  const float* lb, const float* ub, const float* delta, const float* pose_opt, const float* pose_cov,
                  const float* noise_normal, const float* noise_chi2, const float* noise_rot,
                  uint64_t seed, uint32_t obj_offset, float* pose_samples, float* logw, float* proposals,
                  int B, int N, const EpnpParams* p, void* stream) {
    if (!p) return EPNP_ERR_BAD_ARG;
    KArgs a{};
    a.x3d = x3d; a.x2d = x2d; a.w2d = w2d; a.cam = cam_mats; a.lb = lb; a.ub = ub; a.delta = delta;
    a.pose_opt_in = pose_opt; a.pose_cov_in = pose_cov;
    a.noise_n3 = noise_normal; a.noise_chi2 = noise_chi2; a.noise_rot = noise_rot;
    a.seed = seed; a.obj_offset = obj_offset;
    a.pose_samples = pose_samples; a.logw = logw; a.proposals = proposals; a.B = B; a.N = N; a.p = *p;
    int rc = check_common(a);
    if (rc != EPNP_OK) return rc;
    rc = check_amis_params(*p);
    if (rc != EPNP_OK) return rc;
    if (!pose_opt || !pose_cov || !pose_samples || !logw) return EPNP_ERR_BAD_ARG;
    const bool any = noise_normal || noise_chi2 || noise_rot, all = noise_normal && noise_chi2 && noise_rot;
    if (any && !all) return EPNP_ERR_BAD_ARG;
    if (p->dof == 6)
        return launch_persistent(solve_kernel<6, false, true>, a, plan_smem<6>(N, p->mc_samples, p->mc_iter, true).total_bytes,
                                 (cudaStream_t)stream, plan_smem<6>(N, p->mc_samples, p->mc_iter, true, true).total_bytes);
    return launch_persistent(solve_kernel<4, false, true>, a, plan_smem<4>(N, p->mc_samples, p->mc_iter, true).total_bytes,
                             (cudaStream_t)stream, plan_smem<4>(N, p->mc_samples, p->mc_iter, true, true).total_bytes);
}

int epnp_lm_amis_fused_f32(const float* x3d, const float* x2d, const float* w2d, const float* cam_mats,
                           const float* lb, const float* ub, const float* delta, const float* pose_init,
                           const float* noise_normal, const float* noise_chi2, const float* noise_rot,
                           uint64_t seed, uint32_t obj_offset,
                           float* pose_opt, float* pose_cov, float* cost, float* pose_opt_plus, float* cost_init,
                           float* pose_samples, float* logw, float* proposals,
                           int B, int N, const EpnpParams* p, void* stream) {
    if (!p) return EPNP_ERR_BAD_ARG;
    KArgs a{};
    a.x3d = x3d; a.x2d = x2d; a.w2d = w2d; a.cam = cam_mats; a.lb = lb; a.ub = ub; a.delta = delta;
    a.pose_init = pose_init;
    a.noise_n3 = noise_normal; a.noise_chi2 = noise_chi2; a.noise_rot = noise_rot;
    a.seed = seed; a.obj_offset = obj_offset;
    a.pose_opt = pose_opt; a.pose_cov = pose_cov; a.cost = cost; a.pose_plus = pose_opt_plus; a.cost_init = cost_init;
    a.pose_samples = pose_samples; a.logw = logw; a.proposals = proposals; a.B = B; a.N = N; a.p = *p;
    a.prof = g_prof_buffer;
    int rc = check_common(a);
    if (rc != EPNP_OK) return rc;
    rc = check_amis_params(*p);
    if (rc != EPNP_OK) return rc;
    if (!pose_init || !pose_opt || !pose_samples || !logw || p->lm_iter < 0) return EPNP_ERR_BAD_ARG;
    const bool any = noise_normal || noise_chi2 || noise_rot, all = noise_normal && noise_chi2 && noise_rot;
    if (any && !all) return EPNP_ERR_BAD_ARG;
    if (p->dof == 6)
        return launch_persistent(solve_kernel<6, true, true>, a, plan_smem<6>(N, p->mc_samples, p->mc_iter, true).total_bytes,
                                 (cudaStream_t)stream, plan_smem<6>(N, p->mc_samples, p->mc_iter, true, true).total_bytes);
    return launch_persistent(solve_kernel<4, true, true>, a, plan_smem<4>(N, p->mc_samples, p->mc_iter, true).total_bytes,
                             (cudaStream_t)stream, plan_smem<4>(N, p->mc_samples, p->mc_iter, true, true).total_bytes);
}

int epnp_lm_amis_fused_push_f32(const float* x3d, const float* x2d, const float* w2d, const float* cam_mats,
                                const float* lb, const float* ub, const float* delta, const float* pose_init,
                                uint64_t seed, uint32_t obj_offset,
                                float* pose_opt, float* pose_cov, float* cost, float* pose_samples, float* logw,
                                float* const* peer_logw, float* const* peer_pose, int n_peers,
                                int B, int N, const EpnpParams* p, void* stream) {
    if (!p) return EPNP_ERR_BAD_ARG;
    if (n_peers < 0 || n_peers > EPNP_MAX_PEERS || (n_peers > 0 && (!peer_logw || !peer_pose))) return EPNP_ERR_BAD_ARG;
    KArgs a{};
    a.x3d = x3d; a.x2d = x2d; a.w2d = w2d; a.cam = cam_mats; a.lb = lb; a.ub = ub; a.delta = delta;
    a.pose_init = pose_init;
    a.seed = seed; a.obj_offset = obj_offset;
    a.pose_opt = pose_opt; a.pose_cov = pose_cov; a.cost = cost;
    a.pose_samples = pose_samples; a.logw = logw; a.B = B; a.N = N; a.p = *p;
    a.prof = g_prof_buffer;
    int rc = check_common(a);
    if (rc != EPNP_OK) return rc;
    rc = check_amis_params(*p);
    if (rc != EPNP_OK) return rc;
    if (!pose_init || !pose_opt || !pose_samples || !logw || p->lm_iter < 0) return EPNP_ERR_BAD_ARG;
    PushArgs push{};
    push.n = n_peers;
    for (int r = 0; r < n_peers; ++r) {
        if (!peer_logw[r] || !peer_pose[r]) return EPNP_ERR_BAD_ARG;
        push.logw[r] = peer_logw[r]; push.pose[r] = peer_pose[r];
    }
    if (p->dof == 6)
        return launch_persistent(solve_push_kernel<6>, a, plan_smem<6>(N, p->mc_samples, p->mc_iter, true).total_bytes,
                                 (cudaStream_t)stream, plan_smem<6>(N, p->mc_samples, p->mc_iter, true, true).total_bytes, push);
    return launch_persistent(solve_push_kernel<4>, a, plan_smem<4>(N, p->mc_samples, p->mc_iter, true).total_bytes,
                             (cudaStream_t)stream, plan_smem<4>(N, p->mc_samples, p->mc_iter, true, true).total_bytes, push);
}

int epnp_cost_backward_f32(const float* x3d, const float* x2d, const float* w2d, const float* cam_mats,
                           const float* lb, const float* ub, const float* delta,
                           const float* poses_a, const float* grad_a, int PA,
                           const float* poses_b, const float* grad_b, int PB,
                           float* grad_x3d, float* grad_x2d, float* grad_w2d, float* grad_delta,
                           int B, int N, int dof, float z_min, void* stream) {
    if (!x3d || !x2d || !w2d || !cam_mats || !delta) return EPNP_ERR_BAD_ARG;
    if ((lb == nullptr) != (ub == nullptr)) return EPNP_ERR_BAD_ARG;
    if (B < 0 || N <= 0 || PA < 0 || PB < 0 || (dof != 4 && dof != 6)) return EPNP_ERR_BAD_ARG;
    if ((PA > 0 && (!poses_a || !grad_a)) || (PB > 0 && (!poses_b || !grad_b))) return EPNP_ERR_BAD_ARG;
    if (B == 0) return EPNP_OK;
    BwArgs a{};
    a.x3d = x3d; a.x2d = x2d; a.w2d = w2d; a.cam = cam_mats; a.lb = lb; a.ub = ub; a.delta = delta;
    a.poses_a = poses_a; a.grad_a = grad_a; a.poses_b = poses_b; a.grad_b = grad_b; a.PA = PA; a.PB = PB;
    a.gx3d = grad_x3d; a.gx2d = grad_x2d; a.gw2d = grad_w2d; a.gdelta = grad_delta;
    a.B = B; a.N = N; a.z_min = z_min;
    const int smem = (BW_POSE_TILE * 13 + 2 * NW * 32) * 4;
    cudaError_t e;
    int dev = 0, sms = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return cuda_fail(e);
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return cuda_fail(e);
    const int grid = B < sms * 4 ? B : sms * 4;
    if (dof == 6) {
        if ((e = cudaFuncSetAttribute(cost_backward_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return cuda_fail(e);
        EPNP_LAUNCH(cost_backward_kernel<6>, grid, NT, smem, (cudaStream_t)stream, a);
    } else {
        if ((e = cudaFuncSetAttribute(cost_backward_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return cuda_fail(e);
        EPNP_LAUNCH(cost_backward_kernel<4>, grid, NT, smem, (cudaStream_t)stream, a);
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? EPNP_OK : cuda_fail(e);
}

// ---- host-buffer entry point: chunked H2D -> fused kernel -> D2H on two internal streams
static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

struct WsLayout {
    size_t x3d, x2d, w2d, cam, lb, ub, delta, pose_init, pose_opt, pose_cov, cost, samples, logw, total;
};

static WsLayout ws_layout(int B, int N, const EpnpParams* p) {
    const size_t D = (p->dof == 6) ? 7 : 4, dof = p->dof, M = p->mc_samples;
    WsLayout w; size_t o = 0;
    w.x3d = o; o += align256((size_t)B * N * 3 * 4);
    w.x2d = o; o += align256((size_t)B * N * 2 * 4);
    w.w2d = o; o += align256((size_t)B * N * 2 * 4);
    w.cam = o; o += align256((size_t)B * 9 * 4);
    w.lb = o; o += align256((size_t)B * 2 * 4);
    w.ub = o; o += align256((size_t)B * 2 * 4);
    w.delta = o; o += align256((size_t)B * 4);
    w.pose_init = o; o += align256((size_t)B * D * 4);
    w.pose_opt = o; o += align256((size_t)B * D * 4);
    w.pose_cov = o; o += align256((size_t)B * dof * dof * 4);
    w.cost = o; o += align256((size_t)B * 4);
    w.samples = o; o += align256((size_t)B * M * D * 4);
    w.logw = o; o += align256((size_t)B * M * 4);
    w.total = o;
    return w;
}

size_t epnp_fused_workspace_bytes(int B, int N, const EpnpParams* p) {
    if (!p || B < 0 || N <= 0) return 0;
    return ws_layout(B, N, p).total;
}

// Helper streams of the host-buffer entry point: one copy-in, two compute, one copy-out stream per host
// thread, created on first use and kept for the life of the thread (the only resource the library owns;
// creating and destroying four streams per call costs more than a whole chunk of work).
struct HostPipe {
    cudaStream_t in = nullptr, k[2] = {nullptr, nullptr}, out = nullptr;
    int device = -1;
    cudaError_t ensure() {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev == device && in) return cudaSuccess;
        if (in) { cudaStreamDestroy(in); cudaStreamDestroy(k[0]); cudaStreamDestroy(k[1]); cudaStreamDestroy(out); in = nullptr; }
        if ((e = cudaStreamCreateWithFlags(&in, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&k[0], cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&k[1], cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&out, cudaStreamNonBlocking)) != cudaSuccess) return e;
        device = dev;
        return cudaSuccess;
    }
};
static thread_local HostPipe g_pipe;

int epnp_lm_amis_fused_host_f32(const float* x3d_host, const float* x2d_host, const float* w2d_host,
                                const float* cam_mats_host, const float* lb_host, const float* ub_host,
                                const float* delta_host, const float* pose_init_host,
                                uint64_t seed, uint32_t obj_offset,
                                float* pose_opt_host, float* pose_cov_host, float* cost_host,
                                float* pose_samples_host, float* logw_host,
                                void* workspace, size_t workspace_bytes, int n_chunks,
                                int B, int N, const EpnpParams* p, void* stream_) {
    if (!p || !workspace || !x3d_host || !x2d_host || !w2d_host || !cam_mats_host || !delta_host || !pose_init_host ||
        !pose_opt_host || !logw_host)
        return EPNP_ERR_BAD_ARG;
    if ((lb_host == nullptr) != (ub_host == nullptr)) return EPNP_ERR_BAD_ARG;
    if (B < 0 || N <= 0) return EPNP_ERR_BAD_ARG;
    if (B == 0) return EPNP_OK;
    const WsLayout w = ws_layout(B, N, p);
    if (workspace_bytes < w.total) return EPNP_ERR_BAD_ARG;
    // n_chunks = 0: cut the batch at whole waves of resident CTAs (a chunk of 1.7 waves costs 2), as few waves per chunk
    // as the 64-chunk limit allows; n_chunks >= 1: that many equal chunks
    int chunk_objects = 0;
    if (n_chunks == 0) {
        if (check_amis_params(*p) != EPNP_OK) return EPNP_ERR_BAD_ARG;
        const int wave = (p->dof == 6)
            ? resident_objects(solve_kernel<6, true, true>, plan_smem<6>(N, p->mc_samples, p->mc_iter, true).total_bytes)
            : resident_objects(solve_kernel<4, true, true>, plan_smem<4>(N, p->mc_samples, p->mc_iter, true).total_bytes);
        if (wave > 0) {
            const int waves_per_chunk = (B + 64 * wave - 1) / (64 * wave);
            chunk_objects = wave * waves_per_chunk;
            n_chunks = (B + chunk_objects - 1) / chunk_objects;
        }
    }
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > B) n_chunks = B;
    if (n_chunks > 64) n_chunks = 64;
    cudaStream_t stream = (cudaStream_t)stream_;
    char* ws = (char*)workspace;
    const size_t D = (p->dof == 6) ? 7 : 4, dof = p->dof, M = p->mc_samples;
    // Three-stage pipeline over object chunks: copy-in stream -> {compute 0, compute 1} -> copy-out stream, chained
    // by per-chunk events, so all H2D copies run back to back on one DMA engine, all D2H copies on the other, and
    // the solve of chunk c overlaps both (and the tail of chunk c-1 on the other compute stream).
    cudaError_t e = cudaSuccess;
    int rc = EPNP_OK;
    cudaEvent_t fork = nullptr, ev_in[64], ev_k[64], done[3] = {nullptr, nullptr, nullptr};
    for (int i = 0; i < 64; ++i) { ev_in[i] = nullptr; ev_k[i] = nullptr; }
#define EPNP_TRY(x) do { e = (x); if (e != cudaSuccess) { rc = cuda_fail(e); goto done; } } while (0)
    EPNP_TRY(g_pipe.ensure());
    EPNP_TRY(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    EPNP_TRY(cudaEventRecord(fork, stream));
    EPNP_TRY(cudaStreamWaitEvent(g_pipe.in, fork, 0));
    EPNP_TRY(cudaStreamWaitEvent(g_pipe.k[0], fork, 0));
    EPNP_TRY(cudaStreamWaitEvent(g_pipe.k[1], fork, 0));
    EPNP_TRY(cudaStreamWaitEvent(g_pipe.out, fork, 0));
    for (int c = 0; c < n_chunks; ++c) {
        const int b0 = chunk_objects ? c * chunk_objects : (int)((long long)B * c / n_chunks);
        const int b1 = chunk_objects ? (b0 + chunk_objects < B ? b0 + chunk_objects : B) : (int)((long long)B * (c + 1) / n_chunks);
        const int nb = b1 - b0;
        if (nb <= 0) continue;
        cudaStream_t sk = g_pipe.k[c & 1];
#define H2D(field, host, per) EPNP_TRY(cudaMemcpyAsync(ws + w.field + (size_t)b0 * (per) * 4, (host) + (size_t)b0 * (per), (size_t)nb * (per) * 4, cudaMemcpyHostToDevice, g_pipe.in))
#define D2H(field, host, per) EPNP_TRY(cudaMemcpyAsync((host) + (size_t)b0 * (per), ws + w.field + (size_t)b0 * (per) * 4, (size_t)nb * (per) * 4, cudaMemcpyDeviceToHost, g_pipe.out))
        H2D(x3d, x3d_host, (size_t)N * 3); H2D(x2d, x2d_host, (size_t)N * 2); H2D(w2d, w2d_host, (size_t)N * 2);
        H2D(cam, cam_mats_host, 9); H2D(delta, delta_host, 1); H2D(pose_init, pose_init_host, D);
        if (lb_host) { H2D(lb, lb_host, 2); H2D(ub, ub_host, 2); }
        EPNP_TRY(cudaEventCreateWithFlags(&ev_in[c], cudaEventDisableTiming));
        EPNP_TRY(cudaEventRecord(ev_in[c], g_pipe.in));
        EPNP_TRY(cudaStreamWaitEvent(sk, ev_in[c], 0));
        rc = epnp_lm_amis_fused_f32(
            (float*)(ws + w.x3d) + (size_t)b0 * N * 3, (float*)(ws + w.x2d) + (size_t)b0 * N * 2,
            (float*)(ws + w.w2d) + (size_t)b0 * N * 2, (float*)(ws + w.cam) + (size_t)b0 * 9,
            lb_host ? (float*)(ws + w.lb) + (size_t)b0 * 2 : nullptr, lb_host ? (float*)(ws + w.ub) + (size_t)b0 * 2 : nullptr,
            (float*)(ws + w.delta) + b0, (float*)(ws + w.pose_init) + (size_t)b0 * D,
            nullptr, nullptr, nullptr, seed, obj_offset + (uint32_t)b0,
            (float*)(ws + w.pose_opt) + (size_t)b0 * D, (float*)(ws + w.pose_cov) + (size_t)b0 * dof * dof,
            (float*)(ws + w.cost) + b0, nullptr, nullptr,
            (float*)(ws + w.samples) + (size_t)b0 * M * D, (float*)(ws + w.logw) + (size_t)b0 * M, nullptr,
            nb, N, p, sk);
        if (rc != EPNP_OK) goto done;
        EPNP_TRY(cudaEventCreateWithFlags(&ev_k[c], cudaEventDisableTiming));
        EPNP_TRY(cudaEventRecord(ev_k[c], sk));
        EPNP_TRY(cudaStreamWaitEvent(g_pipe.out, ev_k[c], 0));
        D2H(pose_opt, pose_opt_host, D); D2H(logw, logw_host, M);
        if (pose_cov_host) D2H(pose_cov, pose_cov_host, dof * dof);
        if (cost_host) D2H(cost, cost_host, 1);
        if (pose_samples_host) D2H(samples, pose_samples_host, M * D);
#undef H2D
#undef D2H
    }
    {   // join: the caller's stream waits for everything the helpers were given
        cudaStream_t all[3] = {g_pipe.out, g_pipe.k[0], g_pipe.k[1]};
        for (int i = 0; i < 3; ++i) {
            EPNP_TRY(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
            EPNP_TRY(cudaEventRecord(done[i], all[i]));
            EPNP_TRY(cudaStreamWaitEvent(stream, done[i], 0));
        }
    }
done:
#undef EPNP_TRY
    // events are released once recorded work has drained (cudaEventDestroy defers)
    for (int i = 0; i < 64; ++i) { if (ev_in[i]) cudaEventDestroy(ev_in[i]); if (ev_k[i]) cudaEventDestroy(ev_k[i]); }
    for (int i = 0; i < 3; ++i) if (done[i]) cudaEventDestroy(done[i]);
    if (fork) cudaEventDestroy(fork);
    return rc;
}

}  // extern "C"
