// Microbenchmark: issue rate of tcgen05.mma (kind::f16, M=128) from shared memory for the operand layouts the conv
// kernel uses.  One CTA per SM, one warp issues `reps` rounds of a fixed MMA pattern, cycles measured with clock64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_microbench tools/mma_microbench.cu && ./mma_microbench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_acc(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

struct Cfg { int N, ntaps, KK, J, win_rows, a_shift_rows, swz, reps; };

__global__ void __launch_bounds__(128, 1) k_bench(Cfg c, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // 1.0h
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tslot;
    if (warp == 1) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | (8u << 24);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 120 * 1024);
        uint64_t a0, b0; uint32_t a_kk, b_kk;
        if (c.swz) {       // SWIZZLE_128B K-major: rows of 128 B, 8-row atoms of 1024 B; K step = 32 B inside the row
            a0 = desc(a_base, 16, 1024, 2); b0 = desc(b_base, 16, 1024, 2); a_kk = 2; b_kk = 2;
        } else {           // no swizzle: [K/8][rows][16 B], LBO = rows*16, SBO = 128
            a0 = desc(a_base, c.win_rows * 16, 128, 0); b0 = desc(b_base, c.N * 16, 128, 0);
            a_kk = 2 * c.win_rows; b_kk = 2 * c.N;
        }
        const long long t0 = clock64();
        for (int r = 0; r < c.reps; ++r) {
            for (int tap = 0; tap < c.ntaps; ++tap) {
                const int shift = c.swz ? 0 : (tap * c.a_shift_rows) % 101;     // emulate the tap row shifts
                const uint64_t at = a0 + (uint64_t)shift, bt = b0 + (uint64_t)(c.swz ? 0 : (tap & 1) * 4 * c.N);
                for (int j = 0; j < c.J; ++j) {
                    const uint64_t aj = at + (uint64_t)(c.swz ? j * 1024 : j * 128);
                    const uint32_t d = tmem + j * c.N;
                    for (int kk = 0; kk < c.KK; ++kk)
                        if (elect_one()) mma(d, aj + (uint64_t)(kk * a_kk), bt + (uint64_t)(kk * b_kk), idesc, (r | tap | kk) != 0);
                }
            }
        }
        if (elect_one()) commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int J, int KK, int NT>
__global__ void __launch_bounds__(128, 1) k_bench_u(Cfg c, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tslot;
    if (warp == 1) {
        const uint32_t N = c.N;
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 120 * 1024);
        const uint64_t a0 = desc(a_base, c.win_rows * 16, 128, 0), b0 = desc(b_base, N * 16, 128, 0);
        const uint32_t a_kk = 2 * c.win_rows, b_kk = 2 * N;
        int offs[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) offs[t] = (t * c.a_shift_rows) % 101;
        const long long t0 = clock64();
        for (int r = 0; r < c.reps; ++r) {
#pragma unroll
            for (int tap = 0; tap < NT; ++tap) {
                const uint64_t at = a0 + (uint64_t)offs[tap], bt = b0 + (uint64_t)((tap & 1) * 4 * N);
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        if (tap == 0) mma(tmem + j * N, at + (uint64_t)(j * 128), bt, idesc, r != 0);
                        else mma_acc(tmem + j * N, at + (uint64_t)(j * 128), bt, idesc);
#pragma unroll
                        for (int kk = 1; kk < KK; ++kk)
                            mma_acc(tmem + j * N, at + (uint64_t)(j * 128 + kk * a_kk), bt + (uint64_t)(kk * b_kk), idesc);
                    }
                }
                __syncwarp();
            }
        }
        if (elect_one()) commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int J, int KK, int NT> void run_u(Cfg c, int sms, long long* d, long long* h) {
    cudaFuncSetAttribute(k_bench_u<J, KK, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_bench_u<J, KK, NT><<<sms, 128, 200 * 1024>>>(c, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
    cudaMemcpy(h, d, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    const double n = (double)c.reps * NT * J * KK;
    printf("unrolled N=%3d J=%d KK=%d taps=%d shift=%2d | %8.1f cyc/MMA (ideal %5.1f)\n", c.N, J, KK, NT, c.a_shift_rows, avg / n, 128.0 * c.N / 256.0);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long* d;
    cudaMalloc(&d, sms * sizeof(long long));
    long long* h = new long long[sms];
    const Cfg cfgs[] = {
        // N ntaps KK J win shift swz reps
        {32, 9, 2, 4, 612, 0, 0, 20},  {32, 9, 2, 4, 612, 1, 0, 20},  {32, 9, 2, 4, 612, 8, 0, 20},  {32, 9, 2, 4, 612, 49, 0, 20},
        {64, 9, 4, 4, 564, 0, 0, 10},  {64, 9, 4, 4, 564, 1, 0, 10},  {64, 9, 4, 4, 564, 25, 0, 10},
        {128, 9, 4, 2, 284, 0, 0, 10}, {128, 9, 4, 2, 284, 1, 0, 10}, {128, 9, 4, 2, 284, 13, 0, 10},
        {256, 9, 4, 1, 144, 0, 0, 10}, {256, 9, 4, 1, 144, 7, 0, 10},
        {32, 9, 2, 4, 0, 0, 1, 20}, {64, 9, 4, 4, 0, 0, 1, 10}, {128, 9, 4, 2, 0, 0, 1, 10}, {256, 9, 4, 1, 0, 0, 1, 10},
    };
    printf("%5s %5s %3s %2s %5s %5s %4s | %10s %10s %8s\n", "N", "taps", "KK", "J", "win", "shift", "swz", "cyc/MMA", "ideal", "TF/s/SM");
    for (const Cfg& c : cfgs) {
        k_bench<<<sms, 128, 200 * 1024>>>(c, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sms * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
        const double n = (double)c.reps * c.ntaps * c.J * c.KK;
        printf("%5d %5d %3d %2d %5d %5d %4d | %10.1f %10.1f\n", c.N, c.ntaps, c.KK, c.J, c.win_rows, c.a_shift_rows, c.swz, avg / n, 128.0 * c.N / 256.0);
    }
    run_u<4, 2, 9>(Cfg{32, 9, 2, 4, 612, 49, 0, 20}, sms, d, h);
    run_u<4, 4, 9>(Cfg{64, 9, 4, 4, 564, 25, 0, 10}, sms, d, h);
    run_u<2, 4, 9>(Cfg{128, 9, 4, 2, 284, 13, 0, 10}, sms, d, h);
    run_u<1, 4, 9>(Cfg{256, 9, 4, 1, 144, 7, 0, 10}, sms, d, h);
    run_u<4, 2, 1>(Cfg{32, 1, 2, 4, 612, 49, 0, 200}, sms, d, h);
    return 0;
}
