// Probe of TMA tile::gather4 on sm_100a: (1) semantics -- does a tensor map over the embedding table with box {32, 1}
// and SWIZZLE_128B, driven with {column, row0..row3}, land four 128-byte table-row segments as four consecutive rows
// of a SWIZZLE_128B tile?  (2) throughput -- clocks per 128-row x 32-column tile (32 gather4 instructions) with the
// instructions spread over 1, 2 or 4 producer warps.  Decides whether the layer-1 A operand can be gathered by TMA
// ("gather fused into layer 1").   nvcc -gencode arch=compute_100a,code=sm_100a -o gather4_probe gather4_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try(bar, parity)) if (clock64() - t0 > 2000000000ll) { printf("timeout\n"); __trap(); }
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* map, uint32_t bar, int col, int r0, int r1, int r2, int r3) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

constexpr int ROWS = 128, BK = 32, TILE_BYTES = ROWS * BK * 4, STAGES = 4;

// producer warps: 0..NP-1; each lane of them issues 32 / (32 * NP) ... i.e. lane (w, l) issues instruction w*32/NP.. ; consumer: warp NP spins on the barriers
template <int NP>
__global__ void __launch_bounds__(32 * (NP + 1)) probe(const __grid_constant__ CUtensorMap map, const int* __restrict__ idx /*[iters][128]*/,
                                                      int iters, float* out /*[grid][128*32] last tile, un-swizzled*/, long long* clocks) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem + STAGES * TILE_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2 * STAGES; ++s) mbar_init(bars + 8 * s, s < STAGES ? 1 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp < NP) {
    // instruction j (0..31) covers tile rows 4j..4j+3; this warp takes j = warp*(32/NP) + lane for lane < 32/NP
    constexpr int PER = 32 / NP;
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(bars + 8 * (STAGES + s), ph ^ 1);               // stage free (first pass: passes immediately)
      if (warp == 0 && lane == 0) mbar_expect_tx(bars + 8 * s, TILE_BYTES);
      __syncwarp();
      if (NP > 1) asm volatile("bar.sync 1, %0;" ::"r"(32 * NP));   // expect_tx before any complete_tx
      if (lane < PER) {
        const int j = warp * PER + lane;
        const int4 r = *reinterpret_cast<const int4*>(idx + (size_t)it * ROWS + 4 * j);
        gather4(smem + s * TILE_BYTES + j * 512, &map, bars + 8 * s, (it & 3) * 32, r.x, r.y, r.z, r.w);
      }
      __syncwarp();
    }
  } else {
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(bars + 8 * s, ph);
      if (it == iters - 1) {                                     // copy the last tile out, undoing the 128B swizzle
        const float* tile = reinterpret_cast<const float*>(smem_raw + (smem - smem_u32(smem_raw)) + s * TILE_BYTES);
        for (int e = lane; e < ROWS * BK; e += 32) {
          const int row = e / BK, col = e % BK;
          const int chunk = (col / 4) ^ (row & 7);
          out[(size_t)blockIdx.x * ROWS * BK + e] = tile[row * BK + chunk * 4 + (col & 3)];
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 8 * (STAGES + s)) : "memory");
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) clocks[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int NP>
static void run(const CUtensorMap& map, const int* d_idx, const std::vector<int>& h_idx, const std::vector<float>& table, int iters, int n_items) {
  const int grid = 148;
  float* d_out; long long* d_clk;
  CK(cudaMalloc(&d_out, (size_t)grid * ROWS * BK * 4));
  CK(cudaMalloc(&d_clk, grid * 8));
  const int smem = STAGES * TILE_BYTES + 1024 + 256;
  CK(cudaFuncSetAttribute(probe<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int rep = 0; rep < 2; ++rep) probe<NP><<<grid, 32 * (NP + 1), smem>>>(map, d_idx, iters, d_out, d_clk);
  CK(cudaDeviceSynchronize());
  std::vector<float> out((size_t)grid * ROWS * BK);
  std::vector<long long> clk(grid);
  CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(clk.data(), d_clk, grid * 8, cudaMemcpyDeviceToHost));
  // check CTA 0's last tile: row r = table[idx[iters-1][r]][col0 + c]
  long long bad = 0;
  const int col0 = ((iters - 1) & 3) * 32;
  for (int r = 0; r < ROWS; ++r)
    for (int c = 0; c < BK; ++c) {
      const float want = table[(size_t)h_idx[(size_t)(iters - 1) * ROWS + r] * 128 + col0 + c];
      if (out[r * BK + c] != want) ++bad;
    }
  double mean = 0; long long mx = 0;
  for (auto c : clk) { mean += (double)c; if (c > mx) mx = c; }
  mean /= grid;
  printf("{\"producer_warps\": %d, \"iters\": %d, \"mismatches\": %lld, \"clk_per_tile_mean\": %.1f, \"clk_per_tile_max\": %.1f, \"bytes_per_clk_per_sm\": %.2f}\n",
         NP, iters, bad, mean / iters, (double)mx / iters, TILE_BYTES / (mean / iters));
  cudaFree(d_out); cudaFree(d_clk);
}

int main() {
  const int n_items = 26744, D = 128, iters = 400;
  std::vector<float> table((size_t)n_items * D);
  for (size_t i = 0; i < table.size(); ++i) table[i] = (float)(i % 1000003) * 0.25f;
  std::vector<int> idx((size_t)iters * ROWS);
  uint32_t s = 12345;
  for (auto& v : idx) { s = s * 1664525u + 1013904223u; v = (int)(s % n_items); }
  float* d_table; int* d_idx;
  CK(cudaMalloc(&d_table, table.size() * 4));
  CK(cudaMalloc(&d_idx, idx.size() * 4));
  CK(cudaMemcpy(d_table, table.data(), table.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice));
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  CUtensorMap map;
  const cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)n_items};
  const cuuint64_t strides[1] = {(cuuint64_t)D * 4};
  const cuuint32_t box[2] = {32, 1};                 // gather4: box rows = 1, four row coordinates per instruction
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_table, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("{\"encode_failed\": %d}\n", (int)r); return 1; }
  run<1>(map, d_idx, idx, table, iters, n_items);
  run<2>(map, d_idx, idx, table, iters, n_items);
  run<4>(map, d_idx, idx, table, iters, n_items);
  return 0;
}
