// Microbenchmark: cost of k_rows' store pattern on B200 -- 27 f64 + 4 i32 + 1 u8 columns,
// each warp storing 32 consecutive rows per instruction, with the warp's first row either
// aligned (multiple of 32) or shifted by `shift` rows (partial sectors at both ends).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int CD = 27, CI = 4;
struct Cols { double *d[CD]; int *i[CI]; uint8_t *b; };
template <bool NARROW>
__global__ void k_cols(Cols cols, long long rows, int shift, int chunk) {
    // each warp handles `chunk` consecutive rows starting at warp_id*chunk + shift (chunk <= 64)
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long w = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); w * chunk < rows; w += nw) {
        for (int r0 = 0; r0 < chunk; r0 += 32) {
            const long long r = w * chunk + shift + r0 + lane;
            if (r0 + lane >= chunk || r >= rows) continue;
#pragma unroll
            for (int k = 0; k < CD; ++k) cols.d[k][r] = (double)(r + k);
            if (NARROW) {
#pragma unroll
                for (int k = 0; k < CI; ++k) cols.i[k][r] = (int)(r + k);
                cols.b[r] = (uint8_t)r;
            }
        }
    }
}
int main() {
    const long long rows = 7507653;
    Cols cols;
    for (int k = 0; k < CD; ++k) cudaMalloc(&cols.d[k], rows * 8 + 1024);
    for (int k = 0; k < CI; ++k) cudaMalloc(&cols.i[k], rows * 4 + 1024);
    cudaMalloc(&cols.b, rows + 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int chunks[] = { 64, 60, 50 };
    for (int narrow = 0; narrow < 2; ++narrow)
        for (int chunk : chunks)
            for (int shift = 0; shift < 4; shift += 3) {
                float best = 1e9f;
                for (int it = 0; it < 6; ++it) {
                    cudaEventRecord(a);
                    if (narrow) k_cols<true><<<148 * 8, 256>>>(cols, rows, shift, chunk);
                    else k_cols<false><<<148 * 8, 256>>>(cols, rows, shift, chunk);
                    cudaEventRecord(b); cudaEventSynchronize(b);
                    float ms; cudaEventElapsedTime(&ms, a, b); if (it > 0 && ms < best) best = ms;
                }
                const double bytes = rows * (8.0 * CD + (narrow ? 17.0 : 0.0));
                printf("narrow %d chunk %2d shift %d: %.3f ms  %.0f GB/s\n", narrow, chunk, shift, best, bytes / best / 1e6);
            }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
