// Stand-alone probe (NOT part of the product), round 4: what the wide-net tensor-core kernel (csrc/nf_wide_tc.cu)
// relies on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/tc_probe2 tools/tc_probe2.cu && gpurun_out/tc_probe2
//
// 1. tcgen05.mma with the A operand in TENSOR MEMORY (lane = row of A, one 32-bit column = two consecutive bf16 K
//    elements, written by the row's own thread with tcgen05.st) and B in shared memory in the no-swizzle K-major
//    layout [K/8][N][8]: numerical check against a host GEMM.
// 2. Cycles per MMA (M = 128, K = 16) as a function of N for A-in-TMEM vs A-in-shared-memory.
// 3. tcgen05.ld / tcgen05.st throughput (32x32b.x32) with 4 / 8 / 16 warps.
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
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" :: "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
#define LD32(taddr, r)                                                                                                     \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19," \
                 "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                   \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),   \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),       \
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),      \
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                    \
                 : "r"(taddr))
#define ST32(taddr, r)                                                                                                     \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20," \
                 "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"                                                          \
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), \
                   "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),   \
                   "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),  \
                   "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory")

struct Args {
    const uint32_t* a_rows;   // [128][8] packed bf16x2: row m, K elements 0..15
    const uint16_t* b_tile;   // [2][N][8] bf16 (K-major no-swizzle, [K/8][N][8])
    float* out;               // [128][N]
    long long* cycles;
    int N, reps, mode;        // mode 0: TS check + timing, 1: SS timing (A copied to smem canonical)
};

// One CTA, 128 threads.  A: 8 TMEM columns at column 256; D: columns 0..N-1.
__global__ void __launch_bounds__(128, 1) probe_mma(Args g) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char* b_s = smem;                              // 2 * N * 16 bytes
    unsigned char* a_s = smem + 2 * 256 * 16;               // SS mode: [2][128][8] bf16 canonical = 4 KB
    for (int i = tid; i < 2 * g.N * 4; i += 128) reinterpret_cast<uint32_t*>(b_s)[i] = reinterpret_cast<const uint32_t*>(g.b_tile)[i];
    // SS copy of A: element (m, k) at (k/8)*(128*16) + m*16 + (k%8)*2
    for (int i = tid; i < 128 * 8; i += 128) {
        const int m = i >> 3, c = i & 7;                    // column c = K elements 2c, 2c+1
        reinterpret_cast<uint32_t*>(a_s)[((c >> 2) * 128 * 16 + m * 16 + (c & 3) * 4) / 4] = g.a_rows[m * 8 + c];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    {   // A row of this thread -> TMEM columns 256..263
        uint32_t r[8];
        for (int c = 0; c < 8; ++c) r[c] = g.a_rows[tid * 8 + c];
        tmem_st8(lane_base + 256u, r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    const uint32_t idesc = make_idesc(128, g.N);
    const uint64_t bdesc = make_desc(smem_u32(b_s), (uint32_t)g.N * 16u, 128u);
    const uint64_t adesc = make_desc(smem_u32(a_s), 128u * 16u, 128u);
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const long long t0 = clock64();
        if (g.mode == 0) {
            mma_ts(tmem, tmem + 256u, bdesc, idesc, 0u);
            for (int r = 1; r < g.reps; ++r) mma_ts(tmem, tmem + 256u, bdesc, idesc, g.reps > 1 ? 0u : 1u);
        } else {
            mma_ss(tmem, adesc, bdesc, idesc, 0u);
            for (int r = 1; r < g.reps; ++r) mma_ss(tmem, adesc, bdesc, idesc, 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
        mbar_wait(smem_u32(&mbar), 0);
        g.cycles[0] = clock64() - t0;
    } else {
        mbar_wait(smem_u32(&mbar), 0);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < g.N; c0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(lane_base + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 8; ++c) g.out[tid * g.N + c0 + c] = __uint_as_float(r[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

// tcgen05.ld / st throughput: every warp moves `reps` x (32 lanes x 32 columns x 4 B = 4 KB)
__global__ void __launch_bounds__(512, 1) probe_ldst(long long* cycles, float* sink, int reps, int store) {
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32 % 512);
    uint32_t r[32];
    for (int k = 0; k < 32; ++k) r[k] = tid + k;
    ST32(base, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    float acc = 0.f;
    for (int i = 0; i < reps; ++i) {
        if (store) {
            ST32(base, r);
            if ((i & 3) == 3) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        } else {
            uint32_t q[32];
            LD32(base, q);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += __uint_as_float(q[i & 31]);
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    if (tid == 0) cycles[0] = clock64() - t0;
    if (acc == 123.456f) sink[tid] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
}

static float bf2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }
static uint16_t f2bf(float f) { __nv_bfloat16 b = __float2bfloat16(f); uint16_t u; memcpy(&u, &b, 2); return u; }

int main() {
    srand(3);
    long long* dcyc; float* dout; uint32_t* da; uint16_t* db;
    CK(cudaMalloc(&dcyc, 8)); CK(cudaMalloc(&dout, 128 * 256 * 4)); CK(cudaMalloc(&da, 128 * 8 * 4)); CK(cudaMalloc(&db, 2 * 256 * 16));
    CK(cudaFuncSetAttribute(probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * 16 + 4096));
    std::vector<uint16_t> A(128 * 16);
    for (auto& v : A) v = f2bf((rand() % 2001 - 1000) / 1000.f);
    std::vector<uint32_t> arows(128 * 8);
    for (int m = 0; m < 128; ++m)
        for (int c = 0; c < 8; ++c) arows[m * 8 + c] = (uint32_t)A[m * 16 + 2 * c] | ((uint32_t)A[m * 16 + 2 * c + 1] << 16);
    CK(cudaMemcpy(da, arows.data(), arows.size() * 4, cudaMemcpyHostToDevice));
    const int Ns[] = {16, 32, 48, 64, 128, 256};
    for (int N : Ns) {
        std::vector<uint16_t> B(N * 16), bt(2 * N * 8);
        for (auto& v : B) v = f2bf((rand() % 2001 - 1000) / 1000.f);
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < 16; ++k) bt[((k / 8) * N + n) * 8 + (k % 8)] = B[n * 16 + k];
        CK(cudaMemcpy(db, bt.data(), bt.size() * 2, cudaMemcpyHostToDevice));
        Args g = {da, db, dout, dcyc, N, 1, 0};
        CK(cudaMemset(dout, 0, 128 * 256 * 4));
        probe_mma<<<1, 128, 2 * 256 * 16 + 4096>>>(g);
        CK(cudaDeviceSynchronize());
        std::vector<float> out(128 * N);
        CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
        double err = 0, mx = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double acc = 0;
                for (int k = 0; k < 16; ++k) acc += (double)bf2f(A[m * 16 + k]) * bf2f(B[n * 16 + k]);
                err = fmax(err, fabs(acc - out[m * N + n]));
                mx = fmax(mx, fabs(acc));
            }
        long long c_ts = 0, c_ss = 0;
        g.reps = 1024; g.mode = 0;
        probe_mma<<<1, 128, 2 * 256 * 16 + 4096>>>(g); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&c_ts, dcyc, 8, cudaMemcpyDeviceToHost));
        g.mode = 1;
        probe_mma<<<1, 128, 2 * 256 * 16 + 4096>>>(g); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&c_ss, dcyc, 8, cudaMemcpyDeviceToHost));
        printf("N=%3d  A-in-TMEM check: max|D-expected| = %.3e (max|expected| %.2f)   cycles/MMA (M=128,K=16): A in TMEM %.1f, A in smem %.1f   (floor 128*N/256 = %.0f)\n",
               N, err, mx, c_ts / 1024.0, c_ss / 1024.0, 128.0 * N / 256.0);
    }
    for (int store = 0; store < 2; ++store)
        for (int threads = 128; threads <= 512; threads *= 2) {
            probe_ldst<<<1, threads>>>(dcyc, dout, 256, store);
            CK(cudaDeviceSynchronize());
            long long c = 0;
            CK(cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost));
            printf("tcgen05.%s 32x32b.x32, %2d warps: %.1f cycles per instruction per warp, %.1f B/cycle/SM\n", store ? "st" : "ld", threads / 32,
                   c / 256.0, (double)(threads / 32) * 256 * 4096 / c);
        }
    return 0;
}
