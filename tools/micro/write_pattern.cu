// Microbenchmark (GPU box): HBM write bandwidth of the HalfKP chain kernel's store pattern.
// One chain per lane, `len` rows of 128 B per chain in two arrays; the warp writes the rows of its 32
// chains P plies at a time (P = 1: the shipped kernel; larger P: P * 128 contiguous bytes per lane).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o write_pattern write_pattern.cu && ./write_pattern
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128) k_write(int4* __restrict__ white, int4* __restrict__ black, long long chains, int len, int P)
{
    const long long chain0 = ((long long)blockIdx.x * 128 + (threadIdx.x & ~31));
    const int lane = threadIdx.x & 31, sub = lane >> 3, chunk = lane & 7;
    if (chain0 >= chains) return;
    for (int k = 0; k < len; k += P) {
        for (int q = 0; q < 32 * P; q += 4) {
            const int row = q + sub, owner = row / P, ply = row % P;
            const long long r = (chain0 + owner) * len + k + ply;
            const int4 v = make_int4((int)r, k, q, lane);
            white[r * 8 + chunk] = v;
            black[r * 8 + chunk] = v;
        }
        __syncwarp();
    }
}

int main()
{
    const long long chains = 1000000;
    const int len = 100;
    const size_t bytes = (size_t)chains * len * 128;
    int4 *w, *b;
    cudaMalloc(&w, bytes);
    cudaMalloc(&b, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int Ps[] = {1, 2, 4, 5, 10, 20, 100};
    for (int P : Ps) {
        float best = 1e9f;
        for (int it = 0; it < 4; ++it) {
            cudaEventRecord(e0);
            k_write<<<(unsigned)((chains + 127) / 128), 128>>>(w, b, chains, len, P);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it > 0 && ms < best) best = ms;
        }
        printf("P=%3d  %.3f ms  %.0f GB/s\n", P, best, 2.0 * bytes / best / 1e6);
    }
    cudaMemsetAsync(w, 0, bytes);
    cudaEventRecord(e0);
    cudaMemsetAsync(w, 1, bytes);
    cudaMemsetAsync(b, 1, bytes);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("memset  %.3f ms  %.0f GB/s\n", ms, 2.0 * bytes / ms / 1e6);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
