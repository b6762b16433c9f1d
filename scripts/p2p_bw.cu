// p2p_bw.cu -- micro-benchmark: how fast can SMs push data into a peer GPU's
// memory over NVLink?  (1) plain 16-byte stores, (2) TMA bulk stores from shared
// memory, (3) copy engine.  Single process, two devices, peer access enabled.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/p2p_bw.cu -o /tmp/p2p_bw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(256) StoreKernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __ldg(src + i);
}

// each warp: stage ROW bytes in smem, one lane issues a bulk store
template <int ROW>
__global__ void __launch_bounds__(256) BulkKernel(const uint4* __restrict__ src, char* __restrict__ dst, size_t rows) {
  extern __shared__ __align__(128) char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  char* buf = smem + warp * 2 * ROW;
  size_t gw = (size_t)blockIdx.x * 8 + warp, nw = (size_t)gridDim.x * 8;
  int ph = 0;
  for (size_t r = gw; r < rows; r += nw) {
    char* b = buf + ph * ROW;
    // make sure the bulk store that read this buffer two iterations ago is done
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    const uint4* s = src + r * (ROW / 16);
    for (int i = lane; i < ROW / 16; i += 32) reinterpret_cast<uint4*>(b)[i] = __ldg(s + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      uint32_t sa = (uint32_t)__cvta_generic_to_shared(b);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst + r * ROW), "r"(sa), "r"(ROW) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    ph ^= 1;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  int n = 0; CK(cudaGetDeviceCount(&n));
  if (n < 2) { printf("need 2 GPUs\n"); return 0; }
  const size_t bytes = 256ull << 20;
  CK(cudaSetDevice(1)); char* peer; CK(cudaMalloc(&peer, bytes));
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
  char *src, *local; CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&local, bytes));
  CK(cudaMemset(src, 1, bytes));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time = [&](auto fn, const char* name) {
    fn(); cudaDeviceSynchronize();
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) fn(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    printf("%-44s %8.3f ms  %8.1f GB/s  (%s)\n", name, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  for (int dsti = 0; dsti < 2; ++dsti) {
    char* dst = dsti ? peer : local;
    printf("--- destination: %s\n", dsti ? "PEER (NVLink)" : "local HBM");
    for (int ctas : {37, 74, 148, 296, 592, 1184}) {
      char name[64]; snprintf(name, 64, "st.v4 %d CTAs x 256", ctas);
      time([&] { StoreKernel<<<ctas, 256>>>((const uint4*)src, (uint4*)dst, bytes / 16); }, name);
    }
    for (int ctas : {74, 148, 296, 592}) {
      char name[64]; snprintf(name, 64, "TMA bulk 1 KB rows, %d CTAs", ctas);
      cudaFuncSetAttribute(BulkKernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 1024);
      time([&] { BulkKernel<1024><<<ctas, 256, 8 * 2 * 1024>>>((const uint4*)src, dst, bytes / 1024); }, name);
    }
    for (int ctas : {148, 296}) {
      char name[64]; snprintf(name, 64, "TMA bulk 512 B rows, %d CTAs", ctas);
      time([&] { BulkKernel<512><<<ctas, 256, 8 * 2 * 512>>>((const uint4*)src, dst, bytes / 512); }, name);
    }
    time([&] { cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault); }, "copy engine");
  }
  return 0;
}
