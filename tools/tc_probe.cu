// tc_probe.cu -- standalone probe of one tcgen05.mma.kind::tf32 (M=128, N=128, K=8) under several shared-memory
// operand layouts / descriptor encodings.  Prints the max error of D = A * B^T against the host reference per
// variant.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tc_probe tools/tc_probe.cu ; run on the B200.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Variant {
  int a_mn_major, b_mn_major;  // instruction descriptor major bits
  int layout;                  // 0: MN-major core matrices [g][r] (8 K-rows x 16 B), 1: K-major core matrices (8 rows x 16 B of K)
  uint32_t lbo, sbo;           // bytes
  uint32_t kstep;              // unused here (single K block)
};

__global__ void probe_kernel(const float* __restrict__ A /*128x8 row-major*/, const float* __restrict__ B /*128x8*/, Variant v,
                             float* __restrict__ D /*128x128*/, uint32_t* __restrict__ raw) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  float* sA = reinterpret_cast<float*>(smem);
  float* sB = reinterpret_cast<float*>(smem + 8192);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  __syncthreads();
  // element (p, k): p = row (M or N index), k = 0..7
  for (int e = threadIdx.x; e < 128 * 8; e += blockDim.x) {
    const int p = e >> 3, k = e & 7;
    int off;  // float index inside the 4 KB tile
    if (v.layout == 0) off = (p >> 2) * 32 + k * 4 + (p & 3);               // [g = p/4][r = k][e = p%4]
    else off = (k >> 2) * 512 + (p >> 3) * 32 + (p & 7) * 4 + (k & 3);      // [kc = k/4][rowgroup = p/8][row = p%8][e = k%4]
    sA[off] = A[e];
    sB[off] = B[e];
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&s_bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)v.a_mn_major << 15) | ((uint32_t)v.b_mn_major << 16) |
                           ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto mk = [&](uint32_t addr) {
      return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((v.lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((v.sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
    };
    const uint64_t ad = mk(smem_u32(sA)), bd = mk(smem_u32(sB));
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
        "l"(ad), "l"(bd), "r"(idesc), "r"(0u)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&s_bar)) : "memory");
  }
  uint32_t ok = 0;
  int spin = 0;
  while (!ok && spin < 2000000) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(&s_bar)), "r"(0u) : "memory");
    ++spin;
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp < 4) {
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t r[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + ch * 32;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
            "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
            "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      for (int c = 0; c < 32; ++c) D[(warp * 32 + lane) * 128 + ch * 32 + c] = __uint_as_float(r[c]);
    }
  }
  if (threadIdx.x == 0) { raw[0] = ok; raw[1] = spin; raw[2] = tmem; }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(128) : "memory");
}

int main() {
  std::vector<float> A(128 * 8), B(128 * 8), ref(128 * 128), D(128 * 128);
  srand(1);
  for (int i = 0; i < 128 * 8; ++i) {  // small integers: exact in TF32
    A[i] = (float)(rand() % 17 - 8);
    B[i] = (float)(rand() % 13 - 6);
  }
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 128; ++j) {
      float s = 0;
      for (int k = 0; k < 8; ++k) s += A[i * 8 + k] * B[j * 8 + k];
      ref[i * 128 + j] = s;
    }
  float *dA, *dB, *dD;
  uint32_t* draw;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&draw, 64);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const Variant vs[] = {
      {1, 1, 0, 4096, 128, 0}, {1, 1, 0, 128, 4096, 0}, {1, 1, 0, 128, 128, 0},   // MN-major, [g][r] core matrices
      {0, 0, 1, 2048, 128, 0}, {0, 0, 1, 128, 2048, 0},                           // K-major
      {0, 0, 0, 4096, 128, 0}, {1, 1, 1, 2048, 128, 0},                           // mismatched controls
  };
  for (size_t t = 0; t < sizeof(vs) / sizeof(vs[0]); ++t) {
    cudaMemset(dD, 0xFF, D.size() * 4);
    probe_kernel<<<1, 128, 16384>>>(dA, dB, vs[t], dD, draw);
    cudaError_t e = cudaDeviceSynchronize();
    uint32_t raw[3] = {0, 0, 0};
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(raw, draw, 12, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxerrT = 0;
    int nz = 0;
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < 128; ++j) {
        maxerr = fmax(maxerr, fabs((double)D[i * 128 + j] - ref[i * 128 + j]));
        maxerrT = fmax(maxerrT, fabs((double)D[j * 128 + i] - ref[i * 128 + j]));
        nz += D[i * 128 + j] != 0.f;
      }
    printf("variant %zu (amn=%d bmn=%d layout=%d lbo=%u sbo=%u): err=%s ok=%u spin=%u tmem=0x%x maxerr=%g maxerrT=%g nonzero=%d  D[0][0..3]=%g %g %g %g ref=%g %g %g %g\n",
           t, vs[t].a_mn_major, vs[t].b_mn_major, vs[t].layout, vs[t].lbo, vs[t].sbo, cudaGetErrorString(e), raw[0], raw[1], raw[2], maxerr,
           maxerrT, nz, D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3]);
  }
  return 0;
}
