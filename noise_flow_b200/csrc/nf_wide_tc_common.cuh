// Building blocks shared by the two tensor-core wide-net kernels (nf_wide_tc.cu: weights resident in shared memory,
// widths 32 / 64 / 128; nf_wide_tcs.cu: weights streamed through a ring, widths 256 / 512): PTX wrappers for tcgen05 /
// TMEM / mbarrier / the TMA engine, the bf16 (hi, lo) split, and the TMEM -> TMEM epilogues.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "nf_params.h"
#include "nf_wide.h"

namespace nf {
namespace wtc {

constexpr int COMPUTE_THREADS = 512;
constexpr int THREADS = COMPUTE_THREADS + 32;      // + the TMA producer warp

struct __align__(16) GroupSmem {
    float4 z[NF_PIXELS];        // the patch
    float4 pre[NF_PIXELS];      // conv-3 output being assembled: (shift0, shift1, raw log-scale0, raw log-scale1)
    float4 ex[2][4][32];        // conv-3 row exchange of the current tile: [0] dy = 0 partial sums (for row r + 1), [1] dy = 2 (row r - 1)
    float hdr[128];             // fp32 header of the current coupling (mix matrices, rescaling scale, edge-bias table)
    float red[64];
    float sacc[256];            // batch-statistics probes: per-channel sum [W] and sum of squares [W] of this patch
    uint64_t mbar;              // MMA completion
    uint64_t pad_;
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);   // version 1, SWIZZLE_NONE
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, K-major, M = 128
__host__ __device__ constexpr uint32_t idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t id, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(id), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tNFW_WAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra NFW_DONE;\n\tbra NFW_WAIT;\n\tNFW_DONE:\n\t}\n" :: "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
// TMA engine, 1-D bulk copy global -> shared, completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
// TMA engine, 2-D tensor copy global -> shared: box (NF_TMA_ROW_FLOATS x box_rows of the tensor map) at row `row`
__device__ __forceinline__ void tma_load_rows(uint32_t dst, const CUtensorMap* tmap, int row, uint32_t mbar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(0), "r"(row), "r"(mbar) : "memory");
}
__device__ __forceinline__ void group_barrier(int g, int threads) { asm volatile("bar.sync %0, %1;" :: "r"(g + 1), "r"(threads) : "memory"); }
// Wait for the group's MMAs: only the group's first warp polls the mbarrier; everybody else blocks on the group's named
// barrier, which costs no issue slots (128 spinning threads per waiting group took ~18 % of the SM's issued instructions).
__device__ __forceinline__ void group_wait_mma(uint32_t mbar, uint32_t& phase, int g, int i, int threads) {
    if (i < 32) mbar_wait(mbar, phase);
    phase ^= 1u;
    group_barrier(g, threads);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
                 "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                   "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

// packed fp32 arithmetic (one issue slot for two lanes of work)
__device__ __forceinline__ float2 wt_ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 wt_add2(float2 a, float2 b) { return wt_ffma2(a, make_float2(1.f, 1.f), b); }

// (hi, lo) bf16 split of two floats, packed as the TMEM A operand wants them (first value in the low half):
// hi = rn(v), lo = rn(v - hi); the subtraction is one packed FFMA2.
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float2 r = wt_ffma2(make_float2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xFFFF0000u)), make_float2(-1.f, -1.f), make_float2(v0, v1));
    const __nv_bfloat162 l = __floats2bfloat162_rn(r.x, r.y);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// ReLU fused into the split: hi = rz(max(v, 0)) (cvt.rz.relu: truncation keeps hi <= v, so the remainder of a positive v is
// never negative), lo = rn(max(v - hi, 0)) (cvt.rn.relu: a negative v has hi = 0 and its remainder v < 0 clamps to 0).
// Five instructions per pair instead of eight; dropped-term error 2^-17 relative as before.
__device__ __forceinline__ void relu_split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
    const float2 r = wt_ffma2(make_float2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xFFFF0000u)), make_float2(-1.f, -1.f), make_float2(v0, v1));
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r.y), "f"(r.x));
}
__device__ __forceinline__ float t_tanh(float v) { return 1.f - __fdividef(2.f, exp2f(v * 2.885390081777927f) + 1.f); }
__device__ __forceinline__ float t_exp(float v) { return exp2f(v * 1.4426950408889634f); }
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
__device__ __forceinline__ float4 mix4(float4 v, const float* m) {   // out[o] = sum_i v[i] * m[o*4 + i]
    float r[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) r[o] = fmaf(v.x, m[o * 4], fmaf(v.y, m[o * 4 + 1], fmaf(v.z, m[o * 4 + 2], v.w * m[o * 4 + 3])));
    return make_float4(r[0], r[1], r[2], r[3]);
}

// accumulator columns [c0, c0 + 32) of this thread's pixel -> ReLU -> (hi, lo) split -> A-operand columns of the next GEMM
// (hi words at a_hi, lo words at a_lo, 16 columns each).  ADD2: the accumulator is the sum of two column blocks.
template <bool ADD2>
__device__ __forceinline__ void relu_split_store(uint32_t d0, uint32_t d1, uint32_t a_hi, uint32_t a_lo) {
    uint32_t r[32];
    tmem_ld32(d0, r);
    if (ADD2) {
        uint32_t q[32];
        tmem_ld32(d1, q);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float2 t = wt_add2(make_float2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1])),
                                     make_float2(__uint_as_float(q[2 * k]), __uint_as_float(q[2 * k + 1])));
            r[2 * k] = __float_as_uint(t.x);
            r[2 * k + 1] = __float_as_uint(t.y);
        }
    } else {
        tmem_wait_ld();
    }
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int k = 0; k < 16; ++k)
        relu_split2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1]), hi[k], lo[k]);
    tmem_st16(a_hi, hi);
    tmem_st16(a_lo, lo);
    tmem_wait_st();
}

// Transposed butterfly: 32 values per lane -> lane l ends up with the warp-wide sum of value l (31 shuffles instead of 160).
__device__ __forceinline__ float warp_sum_to_lanes(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < off; ++k) {
            const float send = up ? v[k] : v[k + off], mine = up ? v[k + off] : v[k];
            v[k] = mine + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}
// Batch-statistics probe (batch_norm(training=True), layers.py:388-393): this warp's 32 pixels x this thread's 32 channels
// of a pre-BatchNorm activation -> per-channel sum / sum of squares; lane l accumulates channel (32 h + l).
template <bool ADD2>
__device__ __forceinline__ void probe_accumulate(uint32_t d0, uint32_t d1, int lane, float& acc_s, float& acc_q) {
    uint32_t r[32];
    tmem_ld32(d0, r);
    float v[32], q[32];
    if (ADD2) {
        uint32_t r2[32];
        tmem_ld32(d1, r2);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]) + __uint_as_float(r2[k]);
    } else {
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) q[k] = v[k] * v[k];
    acc_s += warp_sum_to_lanes(v, lane);
    acc_q += warp_sum_to_lanes(q, lane);
}


// Host: describe the weight blob to the TMA engine (driver entry point fetched through the runtime: no libcuda link).
inline cudaError_t make_blob_tensor_map(const float* blob, long long blob_floats, int box_rows, CUtensorMap* out) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (!p || q != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
        fn = (EncodeFn)p;
    }
    const cuuint64_t gdim[2] = {NF_TMA_ROW_FLOATS, (cuuint64_t)(blob_floats / NF_TMA_ROW_FLOATS)};
    const cuuint64_t gstride[1] = {NF_TMA_ROW_FLOATS * sizeof(float)};
    const cuuint32_t box[2] = {NF_TMA_ROW_FLOATS, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(blob), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace wtc
}  // namespace nf
