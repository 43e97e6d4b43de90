// Instruction-throughput microbenchmark for the probe loop's instruction mix (development aid).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
template <int KIND>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, uint32_t p0, uint32_t p1, uint32_t p2) {
  __shared__ uint32_t sm[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 2654435761u;
  __syncthreads();
  uint32_t a[8];
#pragma unroll
  for (int j = 0; j < 8; j++) a[j] = threadIdx.x * 977 + j * 131 + p0;
  const uint32_t lane4 = (threadIdx.x & 31) << 2;
  const unsigned char* smb = (const unsigned char*)sm;
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (KIND == 0) a[j] = a[j] * p1 + p2;                                   // IMAD
      if (KIND == 1) a[j] = __umulhi(a[j], p1) + p2;                          // IMAD.HI (+IADD)
      if (KIND == 2) a[j] = __funnelshift_r(a[j], a[(j + 1) & 7], 8);         // SHF.R.W const
      if (KIND == 3) a[j] = __funnelshift_l(a[j], a[j], a[(j + 1) & 7]);      // SHF.L.W variable
      if (KIND == 4) a[j] = (a[j] & 0x1FF80u) | p1;                           // LOP3
      if (KIND == 5) a[j] = *(const uint32_t*)(smb + ((a[j] & 0x7F80u) | lane4)) + j;  // LDS bank-private (+LOP3+IADD)
      if (KIND == 6) a[j] = __byte_perm(a[j], a[(j + 1) & 7], 0x4321);        // PRMT
      if (KIND == 7) a[j] = __umulhi(a[j], p1);                               // IMAD.HI pure
      if (KIND == 8) a[j] = a[j] + a[(j + 1) & 7] + p2;                       // IADD3
      if (KIND == 9) a[j] = __popc(a[j]) + a[(j + 1) & 7];                    // POPC
      if (KIND == 10) a[j] = __shfl_sync(0xffffffffu, a[j], (threadIdx.x + 1) & 31);  // SHFL
      if (KIND == 11) { asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[j]) : "r"(p1)); }   // mul.hi via PTX
      if (KIND == 12) a[j] = *(const uint32_t*)(smb + ((a[j] & 0x7FC0u) | (lane4 & 0x3C))) + j;  // LDS 16 copies (2-way conflicts)
      if (KIND == 13) a[j] = *(const uint32_t*)(smb + (a[j] & 0x7FFCu)) + j;  // LDS fully shared random
      if (KIND == 14) { uint32_t y = (a[j] * 0x9E3779B1u) >> 15; uint32_t w = *(const uint32_t*)(smb + ((y & 0x7F80u) | lane4)); uint32_t t = __funnelshift_l(w, w, y); a[j] = __funnelshift_l(t, a[j], 1) + j; }  // full probe
      if (KIND == 16) { uint32_t x = a[j] * 0x9E3779B1u; uint32_t w = *(const uint32_t*)(smb + ((x >> 21) * p1 + lane4)); uint32_t t = __funnelshift_l(w, w, x); a[j] = __funnelshift_l(t, a[j], 1) + j; }  // WB probe: IMAD,SHF,IMAD,LDS,SHF,SHF
      if (KIND == 17) { unsigned long long xx = (unsigned long long)(a[j] * 0x9E3779B1u) * (unsigned long long)p2; uint32_t w = *(const uint32_t*)(smb + ((uint32_t)(xx >> 32) * p1 + lane4)); uint32_t t = __funnelshift_l(w, w, (uint32_t)xx); a[j] = __funnelshift_l(t, a[j], 1) + j; }  // row via IMAD.WIDE hi
      if (KIND == 15) { uint32_t y = __umulhi(a[j] * 0x9E3779B1u, p1); uint32_t w = *(const uint32_t*)(smb + ((y & 0x7F80u) | lane4)); uint32_t t = __funnelshift_l(w, w, y); a[j] = __funnelshift_l(t, a[j], 1) + j; }  // probe, y via IMAD.HI
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s ^= a[j];
  if (s == 0x12345678u) out[0] = s;
}

template <int KIND> void run(const char* name, int ops_per_inner) {
  uint32_t* d; cudaMalloc(&d, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<KIND><<<148, 1024>>>(d, 3, KIND >= 16 ? 4u : 1u << 17, KIND == 17 ? 16u : 5);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<KIND><<<148, 1024>>>(d, 3, KIND >= 16 ? 4u : 1u << 17, KIND == 17 ? 16u : 5);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_inner = 148.0 * 32 * ITERS * 8;   // warp-level executions of the inner statement
  printf("%-34s %8.3f ms  %6.2f inner-stmt/ns  (%d SASS ops each) -> %.2f warp-stmt/clk/SM @1.9GHz\n", name, ms, warp_inner / (ms * 1e6), ops_per_inner,
         warp_inner / 148 / (ms * 1e-3 * 1.9e9));
  cudaFree(d);
}
int main() {
  run<0>("IMAD", 1); run<1>("IMAD.HI+IADD", 2); run<7>("IMAD.HI", 1); run<11>("mul.hi PTX", 1);
  run<2>("SHF.R.W const", 1); run<3>("SHF.L.W var", 1); run<4>("LOP3", 1); run<6>("PRMT", 1); run<8>("IADD3", 1); run<9>("POPC+IADD", 2);
  run<10>("SHFL", 1); run<5>("LDS private +LOP3+IADD", 3); run<12>("LDS 16copies +LOP3+IADD", 3); run<13>("LDS shared rnd +LOP+IADD", 3);
  run<14>("probe (IMAD,SHF,LOP3,LDS,SHF,SHF,IADD)", 7); run<15>("probe y via IMAD.HI", 7);
  run<16>("WB probe (IMAD,SHF,IMAD,LDS,SHF,SHF)", 7); run<17>("WB probe, row via IMAD.WIDE", 7);
  return 0;
}
