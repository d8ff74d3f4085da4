// Microbenchmark: the FP64 denominators of this path on B200.
//   (a) throughput of NON-FUSED binary64 arithmetic (independent DMUL/DADD streams, -fmad=false style):
//       every operation of the reference is a separately rounded multiply or add, so the attainable
//       peak of the contact kernels is the DMUL+DADD issue rate, half the DFMA flop rate.
//   (b) latency of one DEPENDENT operation: DADD, DMUL, DFMA, an IEEE division, a square root
//       (the solver's dataflow kernel and the SAT loop are chains of such operations).
//   (c) round-trip latency of a dependent L2 load (ld.cg pointer chase, 32 MB footprint), an L2
//       atomic on a private address, and a store + fence.acq_rel.gpu -- the three memory steps of one
//       hop of k_solve.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_tput(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-3 + 1.0, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int k = 0; k < iters; ++k) {
        a0 = __dadd_rn(__dmul_rn(a0, m), c); a1 = __dadd_rn(__dmul_rn(a1, m), c); a2 = __dadd_rn(__dmul_rn(a2, m), c); a3 = __dadd_rn(__dmul_rn(a3, m), c);
        a4 = __dadd_rn(__dmul_rn(a4, m), c); a5 = __dadd_rn(__dmul_rn(a5, m), c); a6 = __dadd_rn(__dmul_rn(a6, m), c); a7 = __dadd_rn(__dmul_rn(a7, m), c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int OP>
__global__ void k_lat(double *out, long long *cycles, int iters, double x, double y)
{
    double a = x;
    const long long t0 = clock64();
    for (int k = 0; k < iters; ++k) {
        if (OP == 0) a = __dadd_rn(a, y);
        if (OP == 1) a = __dmul_rn(a, y);
        if (OP == 2) a = __fma_rn(a, y, y);
        if (OP == 3) a = __ddiv_rn(y, a) + 1.0;
        if (OP == 4) a = __dsqrt_rn(a) + y;
    }
    const long long t1 = clock64();
    out[0] = a; cycles[0] = t1 - t0;
}

__global__ void k_chase(const unsigned *next, long long *cycles, unsigned *sink, int iters)
{
    unsigned p = 0;
    const long long t0 = clock64();
    for (int k = 0; k < iters; ++k) p = __ldcg(&next[p]);
    const long long t1 = clock64();
    *sink = p; cycles[0] = t1 - t0;
}
__global__ void k_atom(int *cnt, long long *cycles, int iters)
{
    int v = 0;
    const long long t0 = clock64();
    for (int k = 0; k < iters; ++k) v = atomicSub(&cnt[(v & 1023) * 32], 1);
    const long long t1 = clock64();
    cnt[0] = v; cycles[0] = t1 - t0;
}
__global__ void k_fence(double *buf, long long *cycles, int iters)
{
    const long long t0 = clock64();
    for (int k = 0; k < iters; ++k) {
        __stcg(&buf[(k & 1023) * 4], (double)k); __stcg(&buf[(k & 1023) * 4 + 1], (double)k);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    const long long t1 = clock64();
    cycles[0] = t1 - t0;
}

int main()
{
    double *out; long long *cyc; cudaMalloc(&out, 148 * 16 * 1024 * 8); cudaMalloc(&cyc, 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // (a)
    const int iters = 20000;
    for (int bps : { 2, 4, 8 }) {
        float best = 1e9f;
        for (int it = 0; it < 5; ++it) {
            cudaEventRecord(a); k_tput<<<148 * bps, 256>>>(out, iters); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (it > 0 && ms < best) best = ms;
        }
        const double ops = 148.0 * bps * 256 * iters * 16.0;
        std::printf("non-fused FP64 throughput, %d blocks/SM x 256 threads: %.2f T op/s (%.3f ms)\n", bps, ops / (best * 1e-3) / 1e12, best);
    }
    // (b)
    const char *names[] = { "DADD", "DMUL", "DFMA", "DDIV (+DADD)", "DSQRT (+DADD)" };
    const int n = 4000;
    for (int op = 0; op < 5; ++op) {
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
            if (op == 0) k_lat<0><<<1, 1>>>(out, cyc, n, 1.0, 1e-9);
            if (op == 1) k_lat<1><<<1, 1>>>(out, cyc, n, 1.0, 1.0000001);
            if (op == 2) k_lat<2><<<1, 1>>>(out, cyc, n, 1.0, 0.5);
            if (op == 3) k_lat<3><<<1, 1>>>(out, cyc, n, 1.5, 0.7);
            if (op == 4) k_lat<4><<<1, 1>>>(out, cyc, n, 1.5, 0.7);
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        }
        std::printf("dependent %s: %.1f cycles per step\n", names[op], (double)h / n);
    }
    // (c)
    {
        const unsigned N = 8u << 20;                      // 32 MB of indices: L2 resident, far beyond L1
        unsigned *h = new unsigned[N], *d, *sink;
        unsigned long long s = 88172645463325252ull;
        for (unsigned k = 0; k < N; ++k) h[k] = k;
        for (unsigned k = N - 1; k > 0; --k) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; unsigned r = (unsigned)(s % k); unsigned t = h[k]; h[k] = h[r]; h[r] = t; }   // Sattolo: one cycle
        cudaMalloc(&d, N * 4ull); cudaMalloc(&sink, 4); cudaMemcpy(d, h, N * 4ull, cudaMemcpyHostToDevice);
        long long c = 0;
        for (int rep = 0; rep < 2; ++rep) { k_chase<<<1, 1>>>(d, cyc, sink, 20000); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); }
        std::printf("dependent ld.cg (32 MB chase, second pass mostly L2): %.0f cycles\n", (double)c / 20000);
        int *cnt; cudaMalloc(&cnt, 1024 * 32 * 4); cudaMemset(cnt, 0, 1024 * 32 * 4);
        for (int rep = 0; rep < 2; ++rep) { k_atom<<<1, 1>>>(cnt, cyc, 4000); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); }
        std::printf("dependent L2 atomic (atomicSub with return): %.0f cycles\n", (double)c / 4000);
        for (int rep = 0; rep < 2; ++rep) { k_fence<<<1, 1>>>(out, cyc, 4000); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); }
        std::printf("2 x st.cg + fence.acq_rel.gpu: %.0f cycles\n", (double)c / 4000);
    }
    std::printf("SM clock attribute: %d kHz\n", clk);
    return 0;
}
