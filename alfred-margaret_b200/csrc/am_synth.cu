// am_synth.cu -- counter-based synthetic haystack generator (bench / test tooling).
//
// BASELINE.json's configs name synthetic haystacks of up to 64 GiB; they are generated on the
// device, shard by shard.  Every byte is a pure function of (seed, absolute byte index), so a
// host implementation (alfred_margaret_b200/synth.py, numpy) reproduces any slice exactly and
// the CPU oracle can be run on the same bytes.
//
//   fill : byte i = alphabet[(r8 * alphabet_len) >> 8],  r8 = byte (i & 7) of mix64(seed ^ (i >> 3) * GOLDEN)
//   plant: text block k (of `block` bytes) gets needle (r mod n) written at offset 16 + (r >> 32) mod (block - 16),
//          r = mix64(seed ^ k * GOLDEN); a plant may spill into the first 15 bytes of block k + 1, which
//          no other plant touches (needles are at most 16 bytes in the named configs).
#include "am_kernels.h"

namespace am {

__host__ __device__ inline uint64_t synth_mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x;
}
constexpr uint64_t GOLDEN = 0x9E3779B97F4A7C15ull;

__global__ void synth_fill_kernel(uint8_t* buf, uint64_t len, uint64_t first, uint64_t seed, const uint8_t* alpha, uint32_t alpha_len) {
  __shared__ uint8_t sa[256];
  for (uint32_t i = threadIdx.x; i < alpha_len; i += blockDim.x) sa[i] = alpha[i];
  __syncthreads();
  // one thread per aligned 8-byte word of the ABSOLUTE index space
  const uint64_t w0 = first >> 3;
  const uint64_t nwords = ((first + len + 7) >> 3) - w0;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nwords; k += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t w = w0 + k;
    const uint64_t r = synth_mix(seed ^ (w * GOLDEN));
    uint64_t out = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const uint32_t r8 = (uint32_t)(r >> (8 * j)) & 0xFFu;
      out |= (uint64_t)sa[(r8 * alpha_len) >> 8] << (8 * j);
    }
    const uint64_t abs0 = w << 3;
    if (abs0 >= first && abs0 + 8 <= first + len && (((uintptr_t)(buf + (abs0 - first))) & 7) == 0) {
      *reinterpret_cast<uint64_t*>(buf + (abs0 - first)) = out;
    } else {
      for (int j = 0; j < 8; j++) {
        const uint64_t ai = abs0 + j;
        if (ai >= first && ai < first + len) buf[ai - first] = (uint8_t)(out >> (8 * j));
      }
    }
  }
}

__global__ void synth_plant_kernel(uint8_t* buf, uint64_t len, uint64_t first, uint64_t seed, const uint8_t* nb,
                                   const uint32_t* noff, uint32_t n, uint32_t block) {
  const uint64_t k0 = first / block;
  const uint64_t kfirst = k0 > 0 ? k0 - 1 : 0;
  const uint64_t klast = (first + len) / block;  // inclusive
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; kfirst + t <= klast; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = kfirst + t;
    const uint64_t r = synth_mix(seed ^ (k * GOLDEN));
    const uint32_t id = (uint32_t)(r & 0xFFFFFFFFu) % n;
    const uint64_t off = 16 + (r >> 32) % (block - 16);
    const uint64_t at = k * block + off;
    const uint32_t lo = noff[id], hi = noff[id + 1];
    for (uint32_t j = lo; j < hi; j++) {
      const uint64_t ai = at + (j - lo);
      if (ai >= first && ai < first + len) buf[ai - first] = nb[j];
    }
  }
}

cudaError_t launch_synth_fill(uint8_t* buf, uint64_t len, uint64_t first, uint64_t seed, const uint8_t* d_alpha, uint32_t alpha_len, cudaStream_t st) {
  if (len == 0) return cudaSuccess;
  g_kernel_launches++;
  synth_fill_kernel<<<148 * 8, 256, 0, st>>>(buf, len, first, seed, d_alpha, alpha_len);
  return cudaGetLastError();
}

cudaError_t launch_synth_plant(uint8_t* buf, uint64_t len, uint64_t first, uint64_t seed, const uint8_t* d_needle_bytes,
                               const uint32_t* d_needle_off, uint32_t n, uint32_t block, cudaStream_t st) {
  if (len == 0 || n == 0) return cudaSuccess;
  g_kernel_launches++;
  synth_plant_kernel<<<148 * 4, 256, 0, st>>>(buf, len, first, seed, d_needle_bytes, d_needle_off, n, block);
  return cudaGetLastError();
}

}  // namespace am
