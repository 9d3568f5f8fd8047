// pdl_chain.cu -- what programmatic dependent launch guarantees along a CHAIN of kernels (checked on the B200 before the
// frame's kernels were chained this way):  nvcc -gencode arch=compute_100a,code=sm_100a -o pdl_chain pdl_chain.cu && ./pdl_chain
//   A (normal launch): triggers at once, spins ~80 us, then writes buf = val
//   B (PDL): wait; trigger; counts buf != val                      -> must be 0 (completion of the predecessor)
//   C (PDL): counts buf != val BEFORE its wait (A's data: B triggered only after A had completed), wait, trigger
//   D (PDL): same as C one link further
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__global__ void kA(int* buf, int n, int val, int triggerEarly)
{
    if (triggerEarly) pdl_trigger();
    const long long t0 = clock64();
    while (clock64() - t0 < 150000) {}
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) buf[i] = val;
}
__global__ void kB(const int* buf, int n, int val, int* out, int slot, int preWaitCheck, int triggerBeforeWait)
{
    int bad = 0;
    if (triggerBeforeWait) pdl_trigger();
    if (preWaitCheck)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) bad += (((volatile const int*)buf)[i] != val);
    pdl_wait();
    if (!triggerBeforeWait) pdl_trigger();
    if (!preWaitCheck)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) bad += (((volatile const int*)buf)[i] != val);
    if (bad) atomicAdd(out + slot, bad);
}
template <typename... KArgs, typename... Args>
void launch(void (*k)(KArgs...), int grid, int block, cudaStream_t s, bool pdl, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = s;
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    a[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = a; cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, static_cast<KArgs>(args)...);
    if (e != cudaSuccess) printf("launch: %s\n", cudaGetErrorString(e));
}
int main()
{
    const int n = 1 << 20;
    int *buf, *out;
    cudaMalloc(&buf, n * 4); cudaMalloc(&out, 64);
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    for (int mode = 0; mode < 2; mode++)  // 0: trigger after the wait (what the renderer does)  1: trigger before the wait
    {
        cudaMemset(out, 0, 64);
        for (int it = 1; it <= 200; it++)
        {
            launch(kA, 148 * 2, 256, s, false, buf, n, it, 1);
            launch(kB, 148 * 2, 256, s, true, (const int*)buf, n, it, out, 0, 0, mode);  // B: post-wait check
            launch(kB, 148 * 2, 256, s, true, (const int*)buf, n, it, out, 1, 1, mode);  // C: pre-wait check of A's data
            launch(kB, 148 * 2, 256, s, true, (const int*)buf, n, it, out, 2, 1, mode);  // D: same, one link further
        }
        cudaStreamSynchronize(s);
        int h[4];
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("mode %d (trigger %s the wait): B post-wait mismatches %d, C pre-wait %d, D pre-wait %d   [%s]\n", mode, mode ? "BEFORE" : "after", h[0], h[1], h[2],
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
