// Micro-benchmark (2 GPUs, one process): what the building blocks of a cross-GPU barrier cost.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o p2p_fence p2p_fence.cu && ./p2p_fence
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// single thread: time `reps` repetitions of an operation mix; results in ns per repetition
__global__ void k_ops(unsigned long long* local, unsigned long long* remote, double* out, int reps) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long t0, t1;
    // (0) fence.sys, nothing outstanding
    t0 = gt(); for (int i = 0; i < reps; ++i) __threadfence_system(); t1 = gt(); out[0] = double(t1 - t0) / reps;
    // (1) fence.gpu
    t0 = gt(); for (int i = 0; i < reps; ++i) __threadfence(); t1 = gt(); out[1] = double(t1 - t0) / reps;
    // (2) local store + fence.sys
    t0 = gt(); for (int i = 0; i < reps; ++i) { *(volatile unsigned long long*)(local + 8) = i; __threadfence_system(); } t1 = gt(); out[2] = double(t1 - t0) / reps;
    // (3) remote store + fence.sys
    t0 = gt(); for (int i = 0; i < reps; ++i) { *(volatile unsigned long long*)(remote + 8) = i; __threadfence_system(); } t1 = gt(); out[3] = double(t1 - t0) / reps;
    // (4) st.release.sys remote (nothing else outstanding)
    t0 = gt(); for (int i = 0; i < reps; ++i) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote + 16), "l"((unsigned long long)i) : "memory"); t1 = gt(); out[4] = double(t1 - t0) / reps;
    // (5) red.release.sys local
    t0 = gt(); for (int i = 0; i < reps; ++i) asm volatile("red.release.sys.global.add.u64 [%0], 1;" ::"l"(local + 24) : "memory"); t1 = gt(); out[5] = double(t1 - t0) / reps;
    // (6) red.release.gpu local
    t0 = gt(); for (int i = 0; i < reps; ++i) asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(local + 24) : "memory"); t1 = gt(); out[6] = double(t1 - t0) / reps;
    // (7) ld.acquire.sys local
    unsigned long long v, acc = 0;
    t0 = gt(); for (int i = 0; i < reps; ++i) { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(local + 32) : "memory"); acc += v; } t1 = gt(); out[7] = double(t1 - t0) / reps;
    // (8) ld.relaxed.sys local
    t0 = gt(); for (int i = 0; i < reps; ++i) { asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(local + 32) : "memory"); acc += v; } t1 = gt(); out[8] = double(t1 - t0) / reps;
    // (9) remote relaxed store only (issue cost)
    t0 = gt(); for (int i = 0; i < reps; ++i) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(remote + 40), "l"((unsigned long long)i) : "memory"); t1 = gt(); out[9] = double(t1 - t0) / reps;
    // (10) remote load (round trip)
    t0 = gt(); for (int i = 0; i < reps; ++i) { asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(remote + 48 + (i & 1)) : "memory"); acc += v; } t1 = gt(); out[10] = double(t1 - t0) / reps;
    // (11) ld.acquire.gpu local
    t0 = gt(); for (int i = 0; i < reps; ++i) { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(local + 32) : "memory"); acc += v; } t1 = gt(); out[11] = double(t1 - t0) / reps;
    out[15] = (double)acc;
}
// same fences while the other 295 blocks of the GPU stream stores to local memory (a busy memory system)
__global__ void k_ops_busy(unsigned long long* local, unsigned long long* remote, double* out, int reps, double* sink, size_t n, volatile int* stop) {
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            unsigned long long t0, t1;
            t0 = gt(); for (int i = 0; i < reps; ++i) __threadfence_system(); t1 = gt(); out[0] = double(t1 - t0) / reps;
            t0 = gt(); for (int i = 0; i < reps; ++i) __threadfence(); t1 = gt(); out[1] = double(t1 - t0) / reps;
            t0 = gt(); for (int i = 0; i < reps; ++i) { *(volatile unsigned long long*)(remote + 8) = i; __threadfence_system(); } t1 = gt(); out[3] = double(t1 - t0) / reps;
            t0 = gt(); for (int i = 0; i < reps; ++i) asm volatile("red.release.sys.global.add.u64 [%0], 1;" ::"l"(local + 24) : "memory"); t1 = gt(); out[5] = double(t1 - t0) / reps;
            t0 = gt(); for (int i = 0; i < reps; ++i) asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(local + 24) : "memory"); t1 = gt(); out[6] = double(t1 - t0) / reps;
            *stop = 1;
            __threadfence();
        }
        return;
    }
    size_t i = (size_t)(blockIdx.x - 1) * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)(gridDim.x - 1) * blockDim.x;
    double x = 1.0;
    while (!*stop) {
        for (int r = 0; r < 64; ++r) { sink[i % n] = x; i += stride; x += 1.0; }
    }
}
// ping-pong: GPU0 block writes word to GPU1's flag, GPU1 answers into GPU0's flag; one-way latency = RTT / 2
__global__ void k_ping(unsigned long long* mine, unsigned long long* theirs, int reps, int first, double* out, int mode) {
    if (threadIdx.x != 0) return;
    unsigned long long t0 = gt();
    for (unsigned long long i = 1; i <= (unsigned long long)reps; ++i) {
        if (first) {
            if (mode == 0) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(i) : "memory");
            else asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(i) : "memory");
        }
        unsigned long long v;
        do {
            if (mode == 0) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
            else asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        } while (v < i);
        if (!first) {
            if (mode == 0) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(i) : "memory");
            else asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(i) : "memory");
        }
    }
    out[0] = double(gt() - t0) / reps;
}
int main() {
    int n = 0; CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    unsigned long long *a0, *a1; double *o0, *o1, *sink; int* stop;
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0)); CK(cudaMalloc(&a0, 4096)); CK(cudaMemset(a0, 0, 4096)); CK(cudaMalloc(&o0, 256)); CK(cudaMemset(o0, 0, 256));
    CK(cudaMalloc(&sink, 1 << 28)); CK(cudaMalloc(&stop, 4)); CK(cudaMemset(stop, 0, 4));
    CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0)); CK(cudaMalloc(&a1, 4096)); CK(cudaMemset(a1, 0, 4096)); CK(cudaMalloc(&o1, 256));
    CK(cudaSetDevice(0));
    double h[32];
    const char* names[12] = {"fence.sys idle", "fence.gpu idle", "local st + fence.sys", "remote st + fence.sys", "st.release.sys remote", "red.release.sys local",
                             "red.release.gpu local", "ld.acquire.sys local", "ld.relaxed.sys local", "st.relaxed.sys remote (issue)", "ld remote (round trip)", "ld.acquire.gpu local"};
    for (int rep = 0; rep < 2; ++rep) {
        k_ops<<<1, 32>>>(a0, a1, o0, 2000); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, o0, 128, cudaMemcpyDeviceToHost));
    }
    for (int i = 0; i < 12; ++i) printf("%-32s %8.1f ns\n", names[i], h[i]);
    k_ops_busy<<<296, 512>>>(a0, a1, o0, 500, sink, (size_t)(1 << 28) / 8, stop); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, o0, 128, cudaMemcpyDeviceToHost));
    printf("--- while 295 blocks stream stores to local HBM\n");
    for (int i : {0, 1, 3, 5, 6}) printf("%-32s %8.1f ns\n", names[i], h[i]);
    for (int mode = 0; mode < 2; ++mode) {
        CK(cudaMemset(a0, 0, 4096)); CK(cudaSetDevice(1)); CK(cudaMemset(a1, 0, 4096)); CK(cudaDeviceSynchronize());
        k_ping<<<1, 32>>>(a1 + 64, a0 + 64, 2000, 0, o1, mode);
        CK(cudaSetDevice(0));
        k_ping<<<1, 32>>>(a0 + 64, a1 + 64, 2000, 1, o0, mode);
        CK(cudaDeviceSynchronize()); CK(cudaSetDevice(1)); CK(cudaDeviceSynchronize()); CK(cudaSetDevice(0));
        CK(cudaMemcpy(h, o0, 8, cudaMemcpyDeviceToHost));
        printf("ping-pong %-22s %8.1f ns round trip\n", mode ? "(release / acquire)" : "(relaxed)", h[0]);
    }
    return 0;
}
