// tc_probe.cu -- standalone hardware probe for DESIGN.md section 9.3 (the AMIS sweep's 3x4 projection on tcgen05 with
// kind::tf32 operands).  Not part of the library; nothing in the product calls it.  It answers, in one short GPU run,
// the two questions that cannot be settled without hardware:
//
//   1. correctness / descriptor semantics: D[128 x N] (fp32 in TMEM) = A[128 x 8] * B[N x 8]^T with both operands
//      K-major, no swizzle, core matrices of 8 rows x 16 bytes.  The same shared-memory image is described twice --
//      hypothesis 0: LBO = stride between the two 16-byte K-chunks, SBO = stride between 8-row groups (my reading of
//      cute/arch/mma_sm100_desc.hpp + mma_traits_sm100.hpp); hypothesis 1: the two fields swapped -- and the result
//      read back with tcgen05.ld.32x32b is compared with an exact CPU product (inputs are TF32-representable).
//   2. throughput: the sweep pattern itself -- per tile of 32 points, 6 MMAs (3 projection rows x {[A_hi|A_hi],
//      [A_lo|0]} against one [X_hi;X_lo] descriptor) into a double-buffered 96-column TMEM region, then every thread
//      reads its sample's lane (3 x tcgen05.ld.32x32b.x32) and runs the packed Huber epilogue -- reported as SM
//      cycles per (sample, point pair) next to the same epilogue fed by the CUDA-core projection (9 extra FFMA2).
//
// Build / run (see tools/tc_probe/README):  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o tc_probe tc_probe.cu
//                                           timeout 60 ./tc_probe
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

namespace {

constexpr int M_ROWS = 128;        // MMA M = TMEM lanes = samples
constexpr int KDIM = 8;            // kind::tf32: K = 8 per instruction = two 16-byte chunks

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {       // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {     // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4   [32,46) stride byte offset >> 4
//   [46,48) version = 1 (Blackwell)   [61,64) layout type = 0 (SWIZZLE_NONE)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @[4,6), a/b_format TF32 = 2 @[7,10)/[10,13),
// a/b_major K = 0 @15/16, N >> 3 @[17,23), M >> 4 @[24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared image of a (rows x 8) K-major operand: element (m, k) at byte  (k / 4) * rows * 16 + m * 16 + (k % 4) * 4,
// i.e. K-chunk c is a dense array of `rows` float4; 8-row groups are 128 bytes apart, the two chunks rows*16 bytes.
__device__ __forceinline__ void fill_operand(float* dst, const float* src /*rows x 8 row-major*/, int rows) {
    for (int i = threadIdx.x; i < rows * KDIM; i += blockDim.x) {
        const int m = i / KDIM, k = i % KDIM;
        dst[(k / 4) * rows * 4 + m * 4 + (k % 4)] = src[i];
    }
}

// ---------------------------------------------------------------------------------------------- probe 1
__global__ void __launch_bounds__(128) correctness_kernel(const float* A, const float* B, float* D, int N, int hyp, int dup_b) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* sA = reinterpret_cast<float*>(smem);                     // 2 * 128 * 4 floats
    float* sB = sA + 2 * M_ROWS * 4;                                // 2 * N * 4 floats
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 2 * N * 4);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5;
    uint32_t cols = 32;
    while ((int)cols < N) cols <<= 1;
    fill_operand(sA, A, M_ROWS);
    fill_operand(sB, B, N);
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(slot, cols);
    fence_async_smem();                    // generic-proxy writes of the operands -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = *slot;
    if (threadIdx.x == 0) {
        // dup_b: describe B with a K-chunk stride of 0, i.e. both K-halves read chunk 0 (would save the duplicated
        // operand copy of the split-precision scheme if the hardware accepts it)
        const uint32_t chunkA = M_ROWS * 16, chunkB = dup_b ? 0u : (uint32_t)N * 16, group = 128;
        const uint64_t da = hyp == 0 ? make_desc(smem_u32(sA), chunkA, group) : make_desc(smem_u32(sA), group, chunkA);
        const uint64_t db = hyp == 0 ? make_desc(smem_u32(sB), chunkB, group) : make_desc(smem_u32(sB), group, chunkB);
        mma_tf32(taddr, da, db, make_idesc(M_ROWS, N), 0u);
        mma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (c0 + i < N) D[(size_t)threadIdx.x * N + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(taddr, cols);
}

// ---------------------------------------------------------------------------------------------- probe 2
// Packed Huber epilogue of the one-rsqrt sweep (pnp::pair_cost_rsq) for 16 point pairs whose xh / yh / zh are given.
__device__ __forceinline__ float2 huber_pairs(const float (&xh)[32], const float (&yh)[32], const float (&zh)[32],
                                              const float4* uvw /*16 records: -u0 -u1 -v0 -v1 | wu0 wu1 wv0 wv1*/,
                                              float z_min, float delta, float2 acc) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float4 q0 = uvw[2 * j], q1 = uvw[2 * j + 1];
        const float2 z = make_float2(fmaxf(zh[2 * j], z_min), fmaxf(zh[2 * j + 1], z_min));
        const float2 a = __fmul2_rn(__ffma2_rn(make_float2(q0.x, q0.y), z, make_float2(xh[2 * j], xh[2 * j + 1])), make_float2(q1.x, q1.y));
        const float2 b = __fmul2_rn(__ffma2_rn(make_float2(q0.z, q0.w), z, make_float2(yh[2 * j], yh[2 * j + 1])), make_float2(q1.z, q1.w));
        const float2 q = __ffma2_rn(a, a, __fmul2_rn(b, b));
        const float2 qz = __ffma2_rn(q, __fmul2_rn(z, z), make_float2(1e-30f, 1e-30f));
        float r0, r1;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(qz.x));
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(qz.y));
        const float2 s = __fmul2_rn(q, make_float2(r0, r1));
        const float2 m = make_float2(fminf(s.x, delta), fminf(s.y, delta));
        acc = __ffma2_rn(m, __ffma2_rn(m, make_float2(-0.5f, -0.5f), s), acc);
    }
    return acc;
}

// mode 0: projection on tcgen05 (6 MMAs per 32-point tile, double-buffered TMEM, epilogue from tcgen05.ld)
// mode 1: projection on the CUDA cores (9 FFMA2 per pair), same epilogue -- the shipped design's arithmetic
__global__ void __launch_bounds__(128, 4) sweep_kernel(const float* Aall /*6 x 128 x 8*/, const float* Xsplit /*N x 8*/,
                                                       const float* P /*128 x 12*/, const float4* pts /*N/2 x 4*/,
                                                       const float4* uvw /*N/2 x 2*/, float* out, long long* cycles,
                                                       int N, int reps, int mode, int hyp) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* sA = reinterpret_cast<float*>(smem);                     // 6 operands x (2 x 128 x 4)
    float* sX = sA + 6 * 2 * M_ROWS * 4;                            // 2 x N x 4   (chunk 0 = X_hi, chunk 1 = X_lo)
    float4* sU = reinterpret_cast<float4*>(sX + 2 * N * 4);         // N/2 x 2
    float4* sP = sU + N;                                            // N/2 x 4 pair records (mode 1)
    uint64_t* full = reinterpret_cast<uint64_t*>(sP + 2 * N);       // [2]
    uint64_t* empty = full + 2;                                     // [2]
    uint32_t* slot = reinterpret_cast<uint32_t*>(empty + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int o = 0; o < 6; ++o) fill_operand(sA + o * 2 * M_ROWS * 4, Aall + o * M_ROWS * KDIM, M_ROWS);
    fill_operand(sX, Xsplit, N);
    for (int i = tid; i < N; i += blockDim.x) sU[i] = uvw[i];
    for (int i = tid; i < 2 * N; i += blockDim.x) sP[i] = pts[i];
    if (tid == 0) {
        mbar_init(full + 0, 1); mbar_init(full + 1, 1);
        mbar_init(empty + 0, 128); mbar_init(empty + 1, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (mode == 0 && warp == 0) tmem_alloc(slot, 256);              // 2 buffers x 3 rows x 32 columns
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = (mode == 0) ? *slot : 0u;
    const int ntile = N / 32;
    float2 acc = make_float2(0.f, 0.f);
    const long long t0 = clock64();
    if (mode == 0) {
        const uint32_t idesc = make_idesc(M_ROWS, 32);
        int issued = 0;                                             // tiles issued (thread 0), over all reps
        const int total = ntile * reps;
        auto issue = [&](int g) {                                   // thread 0: 6 MMAs of global tile g into buffer g & 1
            const int buf = g & 1, t = g % ntile;
            if (g >= 2) mbar_wait(empty + buf, (uint32_t)(((g >> 1) - 1) & 1));   // epilogue of tile g-2 has drained it
            tc_fence_after();
            auto desc = [&](const float* base, uint32_t chunk) {    // field order as established by probe 1
                return hyp == 0 ? make_desc(smem_u32(base), chunk, 128) : make_desc(smem_u32(base), 128, chunk);
            };
            const uint64_t db = desc(sX + t * 32 * 4, (uint32_t)N * 16);
            for (int r = 0; r < 3; ++r) {
                const uint32_t d = taddr + (uint32_t)(buf * 96 + r * 32);
                mma_tf32(d, desc(sA + (2 * r) * 2 * M_ROWS * 4, M_ROWS * 16), db, idesc, 0u);
                mma_tf32(d, desc(sA + (2 * r + 1) * 2 * M_ROWS * 4, M_ROWS * 16), db, idesc, 1u);
            }
            mma_commit(full + buf);
        };
        if (tid == 0) { issue(0); if (total > 1) issue(1); issued = 2; }
        for (int g = 0; g < total; ++g) {
            const int buf = g & 1, t = g % ntile;
            mbar_wait(full + buf, (uint32_t)((g >> 1) & 1));
            tc_fence_after();
            float xh[32], yh[32], zh[32];
            const uint32_t base = taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * 96);
            tmem_ld32(base, xh); tmem_ld32(base + 32, yh); tmem_ld32(base + 64, zh);
            tc_fence_before();
            mbar_arrive(empty + buf);                               // this thread no longer needs the buffer
            acc = huber_pairs(xh, yh, zh, sU + t * 32, 0.1f, 3.0f, acc);
            if (tid == 0 && issued < total) { issue(issued); ++issued; }
        }
    } else {
        float2 P2[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) P2[k] = make_float2(P[tid * 12 + k], P[tid * 12 + k]);
        for (int rep = 0; rep < reps; ++rep)
            for (int t = 0; t < ntile; ++t) {
                float xh[32], yh[32], zh[32];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 q0 = sP[(t * 16 + j) * 2], q1 = sP[(t * 16 + j) * 2 + 1];      // X0 X1 Y0 Y1 | Z0 Z1 . .
                    const float2 X = make_float2(q0.x, q0.y), Y = make_float2(q0.z, q0.w), Z = make_float2(q1.x, q1.y);
                    const float2 x = __ffma2_rn(P2[0], X, __ffma2_rn(P2[1], Y, __ffma2_rn(P2[2], Z, P2[3])));
                    const float2 y = __ffma2_rn(P2[4], X, __ffma2_rn(P2[5], Y, __ffma2_rn(P2[6], Z, P2[7])));
                    const float2 z = __ffma2_rn(P2[8], X, __ffma2_rn(P2[9], Y, __ffma2_rn(P2[10], Z, P2[11])));
                    xh[2 * j] = x.x; xh[2 * j + 1] = x.y; yh[2 * j] = y.x; yh[2 * j + 1] = y.y; zh[2 * j] = z.x; zh[2 * j + 1] = z.y;
                }
                acc = huber_pairs(xh, yh, zh, sU + t * 32, 0.1f, 3.0f, acc);
            }
    }
    const long long t1 = clock64();
    out[(size_t)blockIdx.x * blockDim.x + tid] = acc.x + acc.y;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (mode == 0 && warp == 0) tmem_dealloc(taddr, 256);
}

// ---------------------------------------------------------------------------------------------- probe 3
// Legacy warp-level path: mma.sync.m16n8k8 TF32 (operands and accumulators in registers, no TMEM, no descriptors).
// Its rate decides whether the projection could move to the tensor pipe WITHOUT the shared-memory / TMEM cost of the
// tcgen05 plan: 12 such MMAs per warp give the 3x4 projection (3xTF32 split) of 32 samples x 8 points.
__global__ void __launch_bounds__(128) mma_sync_rate_kernel(float* out, int iters) {
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f900000u, 0x3fa00000u, 0x3fb00000u}, b[2] = {0x3f000000u, 0x3f100000u};
    float d[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)          // 8 independent accumulator tiles per warp: latency hidden
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                         "{%0, %1, %2, %3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Fragment layout of mma.sync.m16n8k8 (tf32) as I read the PTX ISA -- g = lane / 4, t = lane % 4:
//   A (16x8): a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)     B (8x8): b0 (k = t, n = g)  b1 (k = t+4, n = g)
//   D (16x8): d0 (g, 2t)  d1 (g, 2t+1)  d2 (g+8, 2t)  d3 (g+8, 2t+1)
__global__ void mma_sync_layout_kernel(const float* A /*16x8*/, const float* B /*8(k)x8(n)*/, float* D /*16x8*/) {
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    const uint32_t a0 = __float_as_uint(A[g * 8 + t]), a1 = __float_as_uint(A[(g + 8) * 8 + t]);
    const uint32_t a2 = __float_as_uint(A[g * 8 + t + 4]), a3 = __float_as_uint(A[(g + 8) * 8 + t + 4]);
    const uint32_t b0 = __float_as_uint(B[t * 8 + g]), b1 = __float_as_uint(B[(t + 4) * 8 + g]);
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    D[g * 8 + 2 * t] = d0; D[g * 8 + 2 * t + 1] = d1; D[(g + 8) * 8 + 2 * t] = d2; D[(g + 8) * 8 + 2 * t + 1] = d3;
}

float tf32_trunc(float x) { uint32_t u; std::memcpy(&u, &x, 4); u &= 0xFFFFE000u; std::memcpy(&x, &u, 4); return x; }

}  // namespace

int main(int argc, char** argv) {
    const int N = 64;
    std::printf("{\"probe\": \"tcgen05 kind::tf32, K-major no-swizzle operands\"");
    // ---------------- probe 1: exact product of TF32-representable inputs
    std::vector<float> A(M_ROWS * KDIM), B(N * KDIM), ref((size_t)M_ROWS * N), got((size_t)M_ROWS * N);
    uint32_t seed = 12345u;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return (float)((int)((seed >> 20) & 63) - 32) * 0.125f; };   // multiples of 1/8
    for (auto& v : A) v = rnd();
    for (auto& v : B) v = rnd();
    for (int m = 0; m < M_ROWS; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0.f;
            for (int k = 0; k < KDIM; ++k) s += A[m * KDIM + k] * B[n * KDIM + k];
            ref[(size_t)m * N + n] = s;
        }
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, got.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    int good_hyp = 0;
    for (int hyp = 0; hyp < 2; ++hyp) {
        CK(cudaMemset(dD, 0xFF, got.size() * 4));
        const size_t smem = (2 * M_ROWS * 4 + 2 * N * 4) * 4 + 64;
        correctness_kernel<<<1, 128, smem>>>(dA, dB, dD, N, hyp, 0);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost));
        double worst = 0;
        long bad = 0;
        for (size_t i = 0; i < got.size(); ++i) {
            const double e = std::fabs((double)got[i] - (double)ref[i]);
            if (!(e <= 1e-6)) ++bad;
            if (e > worst || e != e) worst = e;
        }
        if (bad == 0) good_hyp = hyp;
        std::printf(", \"hypothesis_%d\": {\"lbo_is\": \"%s\", \"max_abs_err\": %.4g, \"mismatches\": %ld, \"of\": %zu}", hyp,
                    hyp == 0 ? "K-chunk stride" : "8-row-group stride", worst, bad, got.size());
    }
    {   // zero K-chunk stride on B: D = A[:, 0:4] B[:, 0:4]^T + A[:, 4:8] B[:, 0:4]^T
        for (int m = 0; m < M_ROWS; ++m)
            for (int n = 0; n < N; ++n) {
                float s = 0.f;
                for (int k = 0; k < KDIM; ++k) s += A[m * KDIM + k] * B[n * KDIM + (k % 4)];
                ref[(size_t)m * N + n] = s;
            }
        CK(cudaMemset(dD, 0xFF, got.size() * 4));
        const size_t smem = (2 * M_ROWS * 4 + 2 * N * 4) * 4 + 64;
        correctness_kernel<<<1, 128, smem>>>(dA, dB, dD, N, good_hyp, 1);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost));
        long bad = 0;
        for (size_t i = 0; i < got.size(); ++i)
            if (!(std::fabs((double)got[i] - (double)ref[i]) <= 1e-6)) ++bad;
        std::printf(", \"zero_chunk_stride_on_b\": {\"mismatches\": %ld, \"of\": %zu}", bad, got.size());
    }
    // ---------------- probe 2: sweep throughput, tensor-core projection vs CUDA-core projection
    {
        const int NP = 512, reps = 64, blocks = 148 * 4;
        std::vector<float> Aall(6 * M_ROWS * KDIM, 0.f), X(NP * KDIM, 0.f), P(M_ROWS * 12), pts(NP / 2 * 16, 0.f), uvw(NP / 2 * 8);
        for (auto& v : P) v = rnd();
        for (int s = 0; s < M_ROWS; ++s)
            for (int r = 0; r < 3; ++r)
                for (int k = 0; k < 4; ++k) {
                    const float p = P[s * 12 + r * 4 + k], hi = tf32_trunc(p), lo = tf32_trunc(p - hi);
                    Aall[((2 * r) * M_ROWS + s) * KDIM + k] = hi;            // [A_hi | A_hi]
                    Aall[((2 * r) * M_ROWS + s) * KDIM + 4 + k] = hi;
                    Aall[((2 * r + 1) * M_ROWS + s) * KDIM + k] = lo;        // [A_lo | 0]
                }
        for (int n = 0; n < NP; ++n) {
            const float x[4] = {rnd(), rnd(), rnd() + 8.f, 1.f};
            for (int k = 0; k < 4; ++k) {
                const float hi = tf32_trunc(x[k]);
                X[n * KDIM + k] = hi; X[n * KDIM + 4 + k] = tf32_trunc(x[k] - hi);
            }
            float* rec = &pts[(n >> 1) * 16 + (n & 1)];
            rec[0] = x[0]; rec[2] = x[1]; rec[4] = x[2];
            float* u = &uvw[(n >> 1) * 8 + (n & 1)];
            u[0] = rnd(); u[2] = rnd(); u[4] = 1.f; u[6] = 1.f;
        }
        float *dAall, *dX, *dP, *dpts, *duvw, *dout;
        long long* dcyc;
        CK(cudaMalloc(&dAall, Aall.size() * 4)); CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dP, P.size() * 4));
        CK(cudaMalloc(&dpts, pts.size() * 4)); CK(cudaMalloc(&duvw, uvw.size() * 4));
        CK(cudaMalloc(&dout, (size_t)blocks * 128 * 4)); CK(cudaMalloc(&dcyc, blocks * sizeof(long long)));
        CK(cudaMemcpy(dAall, Aall.data(), Aall.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dpts, pts.data(), pts.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(duvw, uvw.data(), uvw.size() * 4, cudaMemcpyHostToDevice));
        const size_t smem = (6 * 2 * M_ROWS * 4 + 2 * NP * 4) * 4 + (size_t)NP * 16 + (size_t)2 * NP * 16 + 64;
        CK(cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        std::vector<float> outs[2];
        for (int mode = 0; mode < 2; ++mode) {
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            sweep_kernel<<<blocks, 128, smem>>>(dAall, dX, dP, (const float4*)dpts, (const float4*)duvw, dout, dcyc, NP, 2, mode, good_hyp);
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            sweep_kernel<<<blocks, 128, smem>>>(dAall, dX, dP, (const float4*)dpts, (const float4*)duvw, dout, dcyc, NP, reps, mode, good_hyp);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            outs[mode].resize((size_t)blocks * 128);
            CK(cudaMemcpy(outs[mode].data(), dout, outs[mode].size() * 4, cudaMemcpyDeviceToHost));
            std::vector<long long> cyc(blocks);
            CK(cudaMemcpy(cyc.data(), dcyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
            double mean = 0;
            for (auto c : cyc) mean += (double)c;
            mean /= blocks;
            const double pair_samples = (double)blocks * 128.0 * (NP / 2) * reps;
            std::printf(", \"sweep_%s\": {\"ms\": %.4f, \"pair_samples_per_s\": %.4g, \"cta_cycles_per_tile\": %.1f}",
                        mode == 0 ? "tcgen05" : "cuda_cores", ms, pair_samples / (ms * 1e-3), mean / ((double)(NP / 32) * reps));
        }
        double worst = 0;
        for (size_t i = 0; i < 128; ++i) {
            const double d = std::fabs((double)outs[0][i] - (double)outs[1][i]) / (std::fabs((double)outs[1][i]) + 1e-6);
            if (d > worst || d != d) worst = d;
        }
        std::printf(", \"sweep_rel_diff_tc_vs_cuda\": %.3g, \"sweep_note\": \"both modes: 128 threads, %zu B smem per CTA, descriptor hypothesis %d\"", worst, smem, good_hyp);
    }
    {   // probe 3a: do I have the mma.sync fragment layout right? (exact inputs, exact comparison)
        std::vector<float> a(16 * 8), b(8 * 8), want(16 * 8), have(16 * 8);
        for (auto& v : a) v = rnd();
        for (auto& v : b) v = rnd();
        for (int m = 0; m < 16; ++m)
            for (int n = 0; n < 8; ++n) {
                float acc = 0.f;
                for (int k = 0; k < 8; ++k) acc += a[m * 8 + k] * b[k * 8 + n];
                want[m * 8 + n] = acc;
            }
        float *da, *db, *dd;
        CK(cudaMalloc(&da, a.size() * 4)); CK(cudaMalloc(&db, b.size() * 4)); CK(cudaMalloc(&dd, have.size() * 4));
        CK(cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
        mma_sync_layout_kernel<<<1, 32>>>(da, db, dd);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(have.data(), dd, have.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (size_t i = 0; i < have.size(); ++i) bad += !(std::fabs(have[i] - want[i]) <= 1e-6f);
        std::printf(", \"mma_sync_fragment_layout_mismatches\": %d", bad);
    }
    {   // probe 3: mma.sync m16n8k8 TF32 rate, 4 warps per CTA, 4 CTAs per SM
        const int blocks = 148 * 4, iters = 4096;
        float* dout;
        CK(cudaMalloc(&dout, (size_t)blocks * 128 * 4));
        mma_sync_rate_kernel<<<blocks, 128>>>(dout, 64);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        mma_sync_rate_kernel<<<blocks, 128>>>(dout, iters);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double mmas = (double)blocks * 4 * 8 * iters;               // warp-level instructions
        const double macs = mmas * 16 * 8 * 8;
        std::printf(", \"mma_sync_m16n8k8_tf32\": {\"ms\": %.4f, \"tflops\": %.2f, \"warp_mma_per_us_per_sm\": %.1f, "
                    "\"note\": \"needs >= ~256 MAC/clk/SM-subpartition to carry the projection (12 MMAs per 32 samples x 8 points)\"}",
                    ms, 2.0 * macs / (ms * 1e-3) / 1e12, mmas / (ms * 1e3) / 148.0);
    }
    std::printf("}\n");
    return 0;
}
