// Microbenchmark: achievable HBM write bandwidth for k_rows' store pattern on B200.
//   mode 0: one contiguous stream of C*R doubles
//   mode 1: C separate column arrays, each warp stores 256 B to every column per tile (k_rows pattern)
//   mode 2: as 1, but each lane stores two consecutive rows as one 16 B store per column
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o write_streams write_streams.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int C = 27;        // FP64 columns of a contact row (plus 4 int32 + 1 u8 columns in the real kernel)
struct Cols { double *c[C]; };
__global__ void k_one(double *out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = (double)i;
}
__global__ void k_cols(Cols cols, long long rows) {
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < C; ++k) cols.c[k][r] = (double)(r + k);
    }
}
__global__ void k_cols2(Cols cols, long long rows) {
    for (long long r = 2 * (blockIdx.x * (long long)blockDim.x + threadIdx.x); r + 1 < rows; r += 2LL * gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < C; ++k) *reinterpret_cast<double2 *>(&cols.c[k][r]) = make_double2((double)(r + k), (double)(r + k + 1));
    }
}
int main() {
    const long long rows = 7507653;
    Cols cols; double *one;
    for (int k = 0; k < C; ++k) cudaMalloc(&cols.c[k], rows * 8 + 64);
    cudaMalloc(&one, rows * 8 * C);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mode = 0; mode < 3; ++mode)
        for (int blocks_per_sm = 2; blocks_per_sm <= 16; blocks_per_sm *= 2) {
            float best = 1e9f;
            for (int it = 0; it < 6; ++it) {
                cudaEventRecord(a);
                if (mode == 0) k_one<<<148 * blocks_per_sm, 256>>>(one, rows * C);
                else if (mode == 1) k_cols<<<148 * blocks_per_sm, 256>>>(cols, rows);
                else k_cols2<<<148 * blocks_per_sm, 256>>>(cols, rows);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b); if (it > 0 && ms < best) best = ms;
            }
            printf("mode %d blocks/SM %2d: %.3f ms  %.0f GB/s\n", mode, blocks_per_sm, best, rows * 8.0 * C / best / 1e6);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
