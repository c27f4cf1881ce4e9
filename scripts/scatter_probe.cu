// scatter_probe.cu -- what bounds the counting-sort scatter: the returning atomics on the bucket cursors or the
// scattered 4-byte stores?  218M entries, 2^19 buckets (the 2^24-point commit at c = 20).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *cursor, uint32_t *sorted, uint32_t nb, uint32_t per_bucket, uint32_t n, uint32_t *sink) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, w = blockIdx.y;
    if (i >= n) return;
    const uint32_t b = mix(i * 31u + w * 0x9e3779b9u) & (nb - 1);
    uint32_t pos;
    if (MODE == 0) {  // returning atomic + scattered store (the real thing)
        pos = atomicAdd(&cursor[b], 1u);
        sorted[(size_t)b * per_bucket + (pos % per_bucket)] = i;
    } else if (MODE == 1) {  // returning atomic only
        pos = atomicAdd(&cursor[b], 1u);
        if (pos == 0xffffffffu) sink[0] = pos;
    } else if (MODE == 2) {  // scattered store only (position without an atomic)
        pos = mix(i + w) % per_bucket;
        sorted[(size_t)b * per_bucket + pos] = i;
    } else if (MODE == 3) {  // non-returning atomic (RED)
        atomicAdd(&cursor[b], 1u);
    } else if (MODE == 4) {  // gather of the cursor + scattered store (rank known beforehand)
        pos = __ldg(&cursor[b]) + (mix(i + w) % per_bucket);
        sorted[(size_t)b * per_bucket + (pos % per_bucket)] = i;
    }
}
int main() {
    const uint32_t n = 1u << 24, W = 13, nb = 1u << 19, per_bucket = (uint32_t)(((uint64_t)n * W) / nb) + 64;
    uint32_t *cursor, *sorted, *sink;
    cudaMalloc(&cursor, nb * 4);
    cudaMalloc(&sorted, (size_t)nb * per_bucket * 4);
    cudaMalloc(&sink, 4);
    const char *names[] = {"atomic(return) + scattered store", "atomic(return) only", "scattered store only", "atomic (no return)", "cursor gather + scattered store"};
    for (int m = 0; m < 5; m++) {
        float best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
            cudaMemset(cursor, 0, nb * 4);
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a);
            dim3 grid(n / 256, W);
            if (m == 0) k<0><<<grid, 256>>>(cursor, sorted, nb, per_bucket, n, sink);
            if (m == 1) k<1><<<grid, 256>>>(cursor, sorted, nb, per_bucket, n, sink);
            if (m == 2) k<2><<<grid, 256>>>(cursor, sorted, nb, per_bucket, n, sink);
            if (m == 3) k<3><<<grid, 256>>>(cursor, sorted, nb, per_bucket, n, sink);
            if (m == 4) k<4><<<grid, 256>>>(cursor, sorted, nb, per_bucket, n, sink);
            cudaEventRecord(b);
            cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, a, b);
            best = ms < best ? ms : best;
        }
        printf("%-36s %.3f ms  (%.1f G entries/s)\n", names[m], best, (double)n * W / best * 1e-6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
