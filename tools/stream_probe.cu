// Do two streams of one process run concurrently on this box?  A polling kernel on stream A waits for a flag that only
// work submitted LATER on stream B sets.  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o stream_probe.bin stream_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include <time.h>

__global__ void spin_kernel(const unsigned* flag, unsigned target, int sys) {
    if (threadIdx.x == 0) {
        unsigned v;
        do {
            if (sys) asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            else v = *(volatile const unsigned*)flag;
            if (v < target) __nanosleep(200);
        } while (v < target);
    }
}
__global__ void set_kernel(unsigned* flag, unsigned v, int sys) {
    if (threadIdx.x == 0) {
        if (sys) { __threadfence_system(); asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(flag), "r"(v) : "memory"); }
        else *(volatile unsigned*)flag = v;
    }
}
static double now() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static bool poll(cudaStream_t s, double secs) {
    const double t0 = now();
    while (now() - t0 < secs) { if (cudaStreamQuery(s) == cudaSuccess) return true; usleep(1000); }
    return false;
}
int main(int argc, char** argv) {
    const int nonblocking = argc > 1 ? atoi(argv[1]) : 1;
    const int preload = argc > 2 ? atoi(argv[2]) : 0;        // 1: cudaFuncGetAttributes, 2: a warm-up launch of every kernel
    if (preload == 1) { cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, spin_kernel); cudaFuncGetAttributes(&fa, set_kernel); }
    cudaStream_t a, b, c;
    if (nonblocking) { cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&c, cudaStreamNonBlocking); }
    else { cudaStreamCreate(&a); cudaStreamCreate(&b); cudaStreamCreate(&c); }
    unsigned* flag; cudaMalloc(&flag, 256); cudaMemset(flag, 0, 256);
    unsigned* host; cudaMallocHost(&host, 256);
    if (preload == 2) { spin_kernel<<<1, 32>>>(flag, 0, 0); set_kernel<<<1, 32>>>(flag + 32, 1, 0); }
    cudaDeviceSynchronize();
    printf("preload = %d (lazy module loading: a first launch behind a polling kernel needs a context synchronisation)\n", preload);
    for (int sys = 0; sys < 2; ++sys) {
        cudaMemset(flag, 0, 256); cudaDeviceSynchronize();
        spin_kernel<<<1, 32, 0, a>>>(flag, 1, sys);
        printf("sys=%d: spin launched on A, done (expect 0): %d\n", sys, (int)poll(a, 0.2)); fflush(stdout);
        cudaMemcpyAsync(host, flag, 64, cudaMemcpyDeviceToHost, c);
        printf("  D2H copy on C next to the spinning kernel done (expect 1): %d\n", (int)poll(c, 1.0)); fflush(stdout);
        set_kernel<<<1, 32, 0, b>>>(flag, 1, sys);
        printf("  set kernel on B done (expect 1): %d, spin on A done (expect 1): %d\n", (int)poll(b, 2.0), (int)poll(a, 2.0)); fflush(stdout);
        if (cudaStreamQuery(a) != cudaSuccess) { printf("  STUCK: streams do not run concurrently here\n"); fflush(stdout); _exit(1); }
    }
    // a chain on A (kernel, spin, kernel) with the release coming from B in the middle
    cudaMemset(flag, 0, 256); cudaDeviceSynchronize();
    set_kernel<<<1, 32, 0, a>>>(flag + 8, 1, 0);
    spin_kernel<<<1, 32, 0, a>>>(flag, 1, 1);
    set_kernel<<<1, 32, 0, a>>>(flag + 16, 1, 0);
    set_kernel<<<1, 32, 0, b>>>(flag, 1, 1);
    printf("chain: A done (expect 1): %d\n", (int)poll(a, 2.0)); fflush(stdout);
    _exit(0);
}
