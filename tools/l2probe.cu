// tools/l2probe.cu -- measures the ceilings the SpMM gather kernel lives under on this GPU:
//   (a) streaming read bandwidth of a buffer of S bytes re-read many times (L2-resident when
//       S << 126 MB, HBM when S >> L2), float4 coalesced loads;
//   (b) random ROW-GATHER bandwidth: each warp reads random `row_bytes`-byte rows of an
//       [nrows, row_bytes] matrix (the SpMM access pattern without the index stream or math),
//       U independent rows in flight per warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2probe tools/l2probe.cu
// Run:   tools/l2probe            (prints one line per configuration)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void stream_read(const float4* __restrict__ p, size_t n4, int reps, float* sink) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            const float4 v = __ldg(p + i);
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 123.456f) *sink = acc;
}

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// LANES lanes cover one row with float4 loads (row_bytes = LANES*16); 32/LANES rows per warp load
template <int LANES, int U>
__global__ void row_gather(const float4* __restrict__ p, unsigned nrows, int iters, float* sink) {
    const int lane = threadIdx.x & 31;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int sub = lane / LANES, l = lane % LANES;
    float acc = 0.f;
    for (int it = 0; it < iters; it += U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned r = hash32((warp * 131071u + (unsigned)(it + u)) * (32 / LANES) + sub) % nrows;
            v[u] = __ldg(p + (size_t)r * LANES + l);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 123.456f) *sink = acc;
}

// same with 256-bit loads (sm_100: LDG.E.256): LANES8 lanes x 32 B cover one row
template <int LANES8, int U>
__global__ void row_gather256(const float4* __restrict__ p, unsigned nrows, int iters, float* sink) {
    const int lane = threadIdx.x & 31;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int sub = lane / LANES8, l = lane % LANES8;
    float acc = 0.f;
    for (int it = 0; it < iters; it += U) {
        unsigned w[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned r = hash32((warp * 131071u + (unsigned)(it + u)) * (32 / LANES8) + sub) % nrows;
            const float4* a = p + ((size_t)r * LANES8 + l) * 2;
            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(w[u][0]), "=r"(w[u][1]), "=r"(w[u][2]), "=r"(w[u][3]), "=r"(w[u][4]), "=r"(w[u][5]),
                           "=r"(w[u][6]), "=r"(w[u][7]) : "l"(a));
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int q = 0; q < 8; ++q) acc += __uint_as_float(w[u][q]);
    }
    if (acc == 123.456f) *sink = acc;
}

template <int LANES8, int U>
void run_gather256(const float4* buf, size_t bytes, float* sink, int warps_per_sm_target);

template <typename F> float time_ms(F f, int n = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaEventRecord(a);
    for (int i = 0; i < n; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / n;
}

template <int LANES, int U>
void run_gather(const float4* buf, size_t bytes, float* sink, int warps_per_sm_target) {
    const unsigned nrows = (unsigned)(bytes / (LANES * 16));
    const int threads = 256, blocks = 148 * warps_per_sm_target / 8;
    const int iters = 4096;
    float ms = time_ms([&] { row_gather<LANES, U><<<blocks, threads>>>(buf, nrows, iters, sink); });
    CK(cudaGetLastError());
    double moved = (double)blocks * (threads / 32) * iters * 512.0;
    printf("gather  footprint=%7.1f MB row=%4d B  U=%d  warps/SM=%2d : %8.1f GB/s\n", bytes / 1e6, LANES * 16, U,
           warps_per_sm_target, moved / ms / 1e6);
}

template <int LANES8, int U>
void run_gather256(const float4* buf, size_t bytes, float* sink, int warps_per_sm_target) {
    const unsigned nrows = (unsigned)(bytes / (LANES8 * 32));
    const int threads = 256, blocks = 148 * warps_per_sm_target / 8;
    const int iters = 4096;
    float ms = time_ms([&] { row_gather256<LANES8, U><<<blocks, threads>>>(buf, nrows, iters, sink); });
    CK(cudaGetLastError());
    double moved = (double)blocks * (threads / 32) * iters * 1024.0;
    printf("gather256 footprint=%7.1f MB row=%4d B  U=%d  warps/SM=%2d : %8.1f GB/s\n", bytes / 1e6, LANES8 * 32, U,
           warps_per_sm_target, moved / ms / 1e6);
}

int main() {
    const size_t maxb = (size_t)2 << 30;
    float4* buf; float* sink;
    CK(cudaMalloc(&buf, maxb)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 0, maxb));
    const double mbs[] = {16, 30, 60, 90, 119, 160, 512, 2048};
    for (double mb : mbs) {
        const size_t bytes = (size_t)(mb * 1e6) / 16 * 16;
        const int reps = mb < 200 ? 40 : 4;
        float ms = time_ms([&] { stream_read<<<148 * 8, 512>>>(buf, bytes / 16, reps, sink); });
        CK(cudaGetLastError());
        printf("stream  footprint=%7.1f MB : %8.1f GB/s\n", mb, (double)bytes * reps / ms / 1e6);
    }
    for (double mb : {30.0, 60.0, 119.0, 1000.0}) {
        const size_t bytes = (size_t)(mb * 1e6) / 512 * 512;
        run_gather<32, 4>(buf, bytes, sink, 32);
        run_gather<32, 8>(buf, bytes, sink, 32);
        run_gather<32, 8>(buf, bytes, sink, 64);
        run_gather<16, 4>(buf, bytes, sink, 32);
        run_gather<16, 8>(buf, bytes, sink, 64);
        run_gather<8, 8>(buf, bytes, sink, 64);
        run_gather256<16, 4>(buf, bytes, sink, 16);
        run_gather256<16, 4>(buf, bytes, sink, 24);
        run_gather256<16, 4>(buf, bytes, sink, 32);
        run_gather256<16, 4>(buf, bytes, sink, 64);
        run_gather<32, 4>(buf, bytes, sink, 16);
        run_gather<32, 4>(buf, bytes, sink, 24);
        run_gather256<8, 4>(buf, bytes, sink, 32);
        run_gather256<8, 2>(buf, bytes, sink, 64);
    }
    return 0;
}
