// Stand-alone probe (NOT part of the product): validates the tcgen05 shared-memory descriptor semantics the
// tensor-core coupling path relies on, and times small-N MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/tc_probe tools/tc_probe.cu && gpurun_out/tc_probe
//
// A operand = "pixel-major image": pixel p at byte p*16 holds 8 bf16.  One M=128,N=16,K=16 MMA reads rows = 128
// consecutive pixels starting at `off`, K chunk 0 = the pixel itself, K chunk 1 = the pixel `tap` pixels further
// (descriptor LBO = tap*16 bytes, SBO = 128 bytes): an implicit-GEMM 3x3 convolution needs no im2col.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // version = 1 (sm_100)
    return d;                 // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

// idesc for kind::f16: c=F32(1)<<4, a=BF16(1)<<7, b=BF16(1)<<10, a/b K-major, N>>3 at 17, M>>4 at 24
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" :: "r"(mbar), "r"(parity) : "memory");
}

struct Args {
    const __nv_bfloat16* a_img;   // [n_pix][8]
    const __nv_bfloat16* b_raw;   // 512 B in the canonical B layout (host builds it)
    float* out;                   // [128][16]
    long long* cycles;            // timing result
    int n_pix, off, lbo_a, sbo_a, lbo_b, sbo_b, reps, M, N, n_acc;
};

__global__ void __launch_bounds__(128, 1) tc_probe(Args g) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    unsigned char* a_s = smem;                        // n_pix * 16 bytes
    unsigned char* b_s = smem + ((g.n_pix * 16 + 127) / 128) * 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < g.n_pix * 4; i += 128) reinterpret_cast<uint32_t*>(a_s)[i] = reinterpret_cast<const uint32_t*>(g.a_img)[i];
    for (int i = tid; i < 128; i += 128) reinterpret_cast<uint32_t*>(b_s)[i] = reinterpret_cast<const uint32_t*>(g.b_raw)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // st.shared data -> visible to the tensor-core (async) proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const uint32_t idesc = make_idesc(g.M, g.N);
    const uint64_t adesc = make_desc(smem_u32(a_s) + g.off * 16, g.lbo_a, g.sbo_a);
    const uint64_t bdesc = make_desc(smem_u32(b_s), g.lbo_b, g.sbo_b);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        t0 = clock64();
        if (g.reps == 1) {
            mma_f16(tmem, adesc, bdesc, idesc, 0u);
        } else if (g.n_acc == 1) {          // dependent chain, unrolled x8 (no per-MMA index math)
            mma_f16(tmem, adesc, bdesc, idesc, 0u);
            for (int r = 0; r < g.reps / 8; ++r) {
#pragma unroll
                for (int u = 0; u < 8; ++u) mma_f16(tmem, adesc, bdesc, idesc, 1u);
            }
        } else if (g.n_acc == 8) {          // 8 independent accumulators round-robin, unrolled
            for (int r = 0; r < g.reps / 8; ++r) {
#pragma unroll
                for (int u = 0; u < 8; ++u) mma_f16(tmem + 16u * u, adesc, bdesc, idesc, r > 0 ? 1u : 0u);
            }
        } else {                            // 8 accumulators + a descriptor increment per MMA (row-tile walk)
            for (int r = 0; r < g.reps / 8; ++r) {
                uint64_t ad = adesc;
#pragma unroll
                for (int u = 0; u < 8; ++u) { mma_f16(tmem + 16u * u, ad, bdesc, idesc, r > 0 ? 1u : 0u); ad += 8; }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    }
    mbar_wait(smem_u32(&mbar), 0);
    if (tid == 0) { t1 = clock64(); g.cycles[0] = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (g.M == 128 || warp < 2 || true)
        for (int c = 0; c < 16; ++c) g.out[(warp * 32 + lane) * 16 + c] = __uint_as_float(r[c]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256u) : "memory");
}

static float bf(const __nv_bfloat16& v) { return __bfloat162float(v); }

int main() {
    const int n_pix = 400, off = 40, tap = 35, M = 128, N = 16;
    std::vector<__nv_bfloat16> a(n_pix * 8), braw(256);
    std::vector<float> B(16 * 16);
    srand(1);
    for (auto& v : a) v = __float2bfloat16((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : B) v = (rand() % 2001 - 1000) / 1000.f;
    // canonical K-major no-swizzle B: element (n,k) at bytes (n%8)*16 + (n/8)*SBO_b + (k/8)*LBO_b + (k%8)*2
    const int lbo_b = 128, sbo_b = 256;
    for (int n = 0; n < 16; ++n)
        for (int k = 0; k < 16; ++k) {
            const int byte = (n % 8) * 16 + (n / 8) * sbo_b + (k / 8) * lbo_b + (k % 8) * 2;
            braw[byte / 2] = __float2bfloat16(B[n * 16 + k]);
            B[n * 16 + k] = bf(braw[byte / 2]);
        }
    __nv_bfloat16 *da, *db; float* dout; long long* dcyc;
    CK(cudaMalloc(&da, a.size() * 2)); CK(cudaMalloc(&db, 512)); CK(cudaMalloc(&dout, 128 * 16 * 4)); CK(cudaMalloc(&dcyc, 8));
    CK(cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, braw.data(), 512, cudaMemcpyHostToDevice));
    const size_t smem = ((n_pix * 16 + 127) / 128) * 128 + 1024;   // B tile (512 B) + slack for the N=32 timing run
    for (int variant = 0; variant < 1; ++variant) {   // variant 1 (LBO/SBO swapped) faults: confirmed on B200, round 1
        Args g = {da, db, dout, dcyc, n_pix, off, 0, 0, 0, 0, 1, M, N, 1};
        // variant 0: LBO = K-chunk stride (tap), SBO = 8-row-group stride (128 B); variant 1: swapped
        if (variant == 0) { g.lbo_a = tap * 16; g.sbo_a = 128; g.lbo_b = lbo_b; g.sbo_b = sbo_b; }
        else              { g.lbo_a = 128; g.sbo_a = tap * 16; g.lbo_b = sbo_b; g.sbo_b = lbo_b; }
        CK(cudaMemset(dout, 0, 128 * 16 * 4));
        tc_probe<<<1, 128, smem>>>(g);
        CK(cudaDeviceSynchronize());
        std::vector<float> out(128 * 16);
        CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
        double err = 0, ref_max = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 16; ++n) {
                double acc = 0;
                for (int k = 0; k < 16; ++k) acc += (double)bf(a[(off + m + (k / 8) * tap) * 8 + (k % 8)]) * B[n * 16 + k];
                err = fmax(err, fabs(acc - out[m * 16 + n]));
                ref_max = fmax(ref_max, fabs(acc));
            }
        printf("variant %d (%s): max |D - expected| = %.3e (max |expected| %.3f)  out[0][0..3] = %.4f %.4f %.4f %.4f\n", variant,
               variant == 0 ? "LBO=K-chunk stride, SBO=8-row stride" : "swapped", err, ref_max, out[0], out[1], out[2], out[3]);
    }
    // timing: 4096 MMAs from one thread, round-robin over n_acc independent accumulators (TMEM column blocks)
    for (int n_acc = 1; n_acc <= 16; n_acc *= (n_acc == 1 ? 8 : 2)) {
        Args g = {da, db, dout, dcyc, n_pix, off, tap * 16, 128, lbo_b, sbo_b, 4096, 128, 16, n_acc};
        tc_probe<<<1, 128, smem>>>(g);
        CK(cudaDeviceSynchronize());
        long long cyc = 0;
        CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
        printf("timing M=128 N=16 K=16, mode %2d (1 = dependent chain, 8 = 8 accumulators, 16 = 8 acc + desc walk): %.2f cycles per MMA\n", n_acc, cyc / 4096.0);
    }
    return 0;
}
