// Microbenchmark behind DESIGN.md's choice of the per-particle counter-map layout: P x 360 rays, each a 4-connected
// DDA of ~L cells from the centre of its own map, one fire-and-forget RED per cell.  Variants: element size
// (8-byte pair / 4-byte plane), layout (row-major / 16x16 blocks row-major inside / 16x16 blocks Z-order inside),
// and whether the maps fit L2 (few distinct maps) or not.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

template <int LAYOUT>
__device__ __forceinline__ size_t cell_index(int x, int y, int W) {
    if (LAYOUT == 0) return (size_t)x + (size_t)y * W;
    const size_t blk = (size_t)(y >> 4) * (W >> 4) + (x >> 4);
    if (LAYOUT == 1) return blk * 256 + (y & 15) * 16 + (x & 15);
    unsigned xx = x & 15, yy = y & 15;  // Z-order inside the block
    xx = (xx | (xx << 2)) & 0x33; xx = (xx | (xx << 1)) & 0x55;
    yy = (yy | (yy << 2)) & 0x33; yy = (yy | (yy << 1)) & 0x55;
    return blk * 256 + (xx | (yy << 1));
}

template <typename T, int LAYOUT>
__global__ void k_rays(T* maps, int P, int B, int W, int L, int distinct, size_t map_elems) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)P * B) return;
    const int p = (int)(gid / B), b = (int)(gid - (long long)p * B);
    T* map = maps + (size_t)(p % distinct) * map_elems;
    const float ang = 6.2831853f * (b + 0.37f * (p % 7)) / B;
    const float dx = fabsf(cosf(ang)), dy = fabsf(sinf(ang));
    const int xi = cosf(ang) > 0 ? 1 : -1, yi = sinf(ang) > 0 ? 1 : -1;
    int x = W / 2 + (p % 5), y = W / 2 + (p % 3);
    float err = 0.5f * (dy - dx);
    int n = (int)(L * (dx + dy));
    for (int i = 0; i < n; i++) {
        atomicAdd(map + cell_index<LAYOUT>(x, y, W), (T)1);
        if (err > 0.f) { y += yi; err -= dx; } else { x += xi; err += dy; }
    }
}

template <typename T, int LAYOUT>
float run(T* buf, int P, int B, int W, int L, int distinct, int reps) {
    const size_t map_elems = (size_t)W * W;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const long long total = (long long)P * B;
    k_rays<T, LAYOUT><<<(unsigned)((total + 127) / 128), 128>>>(buf, P, B, W, L, distinct, map_elems);
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(a);
        k_rays<T, LAYOUT><<<(unsigned)((total + 127) / 128), 128>>>(buf, P, B, W, L, distinct, map_elems);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    const int P = 1000, B = 360, W = 1024, L = 170;  // ~216 cells per ray on average
    void* buf;
    const size_t bytes = (size_t)P * W * W * 8;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 0, bytes);
    printf("cells per launch ~ %.1f M\n", P * (double)B * L * 1.27 / 1e6);
    for (int distinct : {1000, 8}) {
        printf("distinct maps %d (%s)\n", distinct, distinct == 1000 ? "8 GB / 4 GB: DRAM" : "64 MB / 32 MB: L2 resident");
        printf("  u64 row-major   %.3f ms\n", run<unsigned long long, 0>((unsigned long long*)buf, P, B, W, L, distinct, 3));
        printf("  u64 blocked16   %.3f ms\n", run<unsigned long long, 1>((unsigned long long*)buf, P, B, W, L, distinct, 3));
        printf("  u64 blocked16-Z %.3f ms\n", run<unsigned long long, 2>((unsigned long long*)buf, P, B, W, L, distinct, 3));
        printf("  u32 row-major   %.3f ms\n", run<unsigned, 0>((unsigned*)buf, P, B, W, L, distinct, 3));
        printf("  u32 blocked16   %.3f ms\n", run<unsigned, 1>((unsigned*)buf, P, B, W, L, distinct, 3));
        printf("  u32 blocked16-Z %.3f ms\n", run<unsigned, 2>((unsigned*)buf, P, B, W, L, distinct, 3));
    }
    return 0;
}
