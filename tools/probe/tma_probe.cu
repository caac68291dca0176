// Probe: which ways of handing a CUtensorMap to cp.async.bulk.tensor work on this box.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <stdlib.h>
#include "../../webrender_b200/csrc/tma.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

template <bool FENCE>
__global__ void copy_box(const CUtensorMap* src, const CUtensorMap* dst, int sx, int sy, int dx, int dy) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    if (FENCE) {
      asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(src) : "memory");
      asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(dst) : "memory");
    }
    wr_mbar_init(&bar, 1);
    wr_fence_mbar_init();
    wr_mbar_expect_tx(&bar, WR_TMA_BOX_BYTES);
    wr_tma_load_2d(sm, src, sx, sy, &bar);
    wr_mbar_wait(&bar, 0);
    wr_tma_store_2d(dst, dx, dy, sm);
    wr_tma_commit();
    wr_tma_wait_all<0>();
  }
}
__global__ void copy_box_param(const __grid_constant__ CUtensorMap src, const __grid_constant__ CUtensorMap dst, int sx, int sy, int dx, int dy) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    wr_mbar_init(&bar, 1);
    wr_fence_mbar_init();
    wr_mbar_expect_tx(&bar, WR_TMA_BOX_BYTES);
    wr_tma_load_2d(sm, &src, sx, sy, &bar);
    wr_mbar_wait(&bar, 0);
    wr_tma_store_2d(&dst, dx, dy, sm);
    wr_tma_commit();
    wr_tma_wait_all<0>();
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn enc;
static int make(CUtensorMap* m, void* p, int w, int h, size_t pitch) {
  cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h};
  cuuint64_t strides[1] = {pitch};
  cuuint32_t box[2] = {WR_TMA_BOX_W, WR_TMA_BOX_H}, es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d (w=%d h=%d pitch=%zu)\n", (int)r, w, h, pitch); return 1; }
  return 0;
}
static int check(const char* what, uint32_t* d_dst, size_t dpitch, int dx, int dy, uint32_t* h_src, int spw, int sx, int sy) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-44s : CUDA ERROR %s\n", what, cudaGetErrorString(e)); return 1; }
  std::vector<uint32_t> row(WR_TMA_BOX_W);
  int bad = 0;
  for (int r = 0; r < WR_TMA_BOX_H; r++) {
    cudaMemcpy(row.data(), (uint8_t*)d_dst + (size_t)(dy + r) * dpitch + (size_t)dx * 4, WR_TMA_BOX_W * 4, cudaMemcpyDeviceToHost);
    for (int x = 0; x < WR_TMA_BOX_W; x++) bad += row[x] != h_src[(size_t)(sy + r) * spw + sx + x];
  }
  printf("%-44s : %s (%d bad)\n", what, bad ? "MISMATCH" : "ok", bad);
  return bad != 0;
}
int main(int argc, char** argv) {
  const int which = argc > 1 ? atoi(argv[1]) : -1;
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  enc = (EncodeFn)fn;
  const int SW = 1024, SH = 512, DW = 3840, DH = 2160;
  size_t sp = SW * 4, dp = 3840 * 4;
  uint32_t *d_src, *d_dst, *d_src2;
  CK(cudaMalloc(&d_src, sp * SH)); CK(cudaMalloc(&d_src2, sp * SH)); CK(cudaMalloc(&d_dst, dp * DH));
  std::vector<uint32_t> h(SW * SH), h2(SW * SH);
  for (size_t i = 0; i < h.size(); i++) { h[i] = (uint32_t)(i * 2654435761u); h2[i] = ~h[i]; }
  CK(cudaMemcpy(d_src, h.data(), sp * SH, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_src2, h2.data(), sp * SH, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_dst, 0, dp * DH));
  CUtensorMap ms, md, ms2;
  if (make(&ms, d_src, SW, SH, sp) || make(&md, d_dst, DW, DH, dp) || make(&ms2, d_src2, SW, SH, sp)) return 1;
  CUtensorMap* table;
  CK(cudaMalloc(&table, 16 * sizeof(CUtensorMap)));
  CK(cudaMemcpy(table + 1, &ms, sizeof ms, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(table + 2, &md, sizeof md, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(copy_box<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WR_TMA_BOX_BYTES));
  CK(cudaFuncSetAttribute(copy_box<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WR_TMA_BOX_BYTES));
  CK(cudaFuncSetAttribute(copy_box_param, cudaFuncAttributeMaxDynamicSharedMemorySize, WR_TMA_BOX_BYTES));
  int fails = 0;
  if (which == 0) { copy_box_param<<<1, 32, WR_TMA_BOX_BYTES>>>(ms, md, 0, 0, 0, 0);
  fails += check("param maps, aligned coords", d_dst, dp, 0, 0, h.data(), SW, 0, 0); }
  if (which == 1) { copy_box_param<<<1, 32, WR_TMA_BOX_BYTES>>>(ms, md, 17, 5, 1041, 33);
  fails += check("param maps, odd coords (17,5)->(1041,33)", d_dst, dp, 1041, 33, h.data(), SW, 17, 5); }
  if (which == 7) { copy_box_param<<<1, 32, WR_TMA_BOX_BYTES>>>(ms, md, 16, 5, 1040, 33);
  fails += check("param maps, 4-px aligned x, odd y", d_dst, dp, 1040, 33, h.data(), SW, 16, 5); }
  if (which == 8) { copy_box_param<<<1, 32, WR_TMA_BOX_BYTES>>>(ms, md, 17, 5, 1040, 33);
  fails += check("param maps, odd src x, aligned dst x", d_dst, dp, 1040, 33, h.data(), SW, 17, 5); }
  if (which == 9) { copy_box_param<<<1, 32, WR_TMA_BOX_BYTES>>>(ms, md, 16, 5, 1041, 33);
  fails += check("param maps, aligned src x, odd dst x", d_dst, dp, 1041, 33, h.data(), SW, 16, 5); }
  if (which == 2) { copy_box<false><<<1, 32, WR_TMA_BOX_BYTES>>>(table + 1, table + 2, 256, 16, 512, 64);
  fails += check("global maps, no fence", d_dst, dp, 512, 64, h.data(), SW, 256, 16); }
  if (which == 3) { copy_box<true><<<1, 32, WR_TMA_BOX_BYTES>>>(table + 1, table + 2, 256, 32, 512, 128);
  fails += check("global maps, acquire fence", d_dst, dp, 512, 128, h.data(), SW, 256, 32); }
  if (which == 4) {
  copy_box<false><<<1, 32, WR_TMA_BOX_BYTES>>>(table + 1, table + 2, 256, 16, 512, 64);
  fails += check("global maps, first use", d_dst, dp, 512, 64, h.data(), SW, 256, 16);
  // rewrite slot 1 with another texture's map: does the next kernel see the new descriptor?
  CK(cudaMemcpy(table + 1, &ms2, sizeof ms2, cudaMemcpyHostToDevice));
  copy_box<false><<<1, 32, WR_TMA_BOX_BYTES>>>(table + 1, table + 2, 0, 64, 768, 256);
  fails += check("global maps, slot rewritten, no fence", d_dst, dp, 768, 256, h2.data(), SW, 0, 64); }
  if (which == 5) {
  copy_box<true><<<1, 32, WR_TMA_BOX_BYTES>>>(table + 1, table + 2, 256, 16, 512, 64);
  fails += check("global maps, first use (fenced)", d_dst, dp, 512, 64, h.data(), SW, 256, 16);
  CK(cudaMemcpy(table + 1, &ms2, sizeof ms2, cudaMemcpyHostToDevice));
  copy_box<true><<<1, 32, WR_TMA_BOX_BYTES>>>(table + 1, table + 2, 0, 64, 768, 256);
  fails += check("global maps, slot rewritten, acquire fence", d_dst, dp, 768, 256, h2.data(), SW, 0, 64); }
  if (which == 6) {
  copy_box<true><<<1, 32, WR_TMA_BOX_BYTES>>>(table + 1, table + 2, 256, 16, 512, 64);
  fails += check("global maps, first use (fenced)", d_dst, dp, 512, 64, h.data(), SW, 256, 16);
  CK(cudaMemcpy(table + 1, &ms, sizeof ms, cudaMemcpyHostToDevice));
  copy_box<true><<<1, 32, WR_TMA_BOX_BYTES>>>(table + 1, table + 2, 0, 80, 768, 512);
  fails += check("global maps, slot rewritten, acquire fence", d_dst, dp, 768, 512, h.data(), SW, 0, 80);
  // freed-and-reallocated memory behind a rewritten slot
  CK(cudaFree(d_src2));
  uint32_t* d_src3; CK(cudaMalloc(&d_src3, sp * SH));
  CK(cudaMemcpy(d_src3, h2.data(), sp * SH, cudaMemcpyHostToDevice));
  CUtensorMap ms3; if (make(&ms3, d_src3, SW, SH, sp)) return 1;
  CK(cudaMemcpy(table + 1, &ms3, sizeof ms3, cudaMemcpyHostToDevice));
  copy_box<true><<<1, 32, WR_TMA_BOX_BYTES>>>(table + 1, table + 2, 0, 96, 1024, 600);
  fails += check("global maps, realloc'd, acquire fence", d_dst, dp, 1024, 600, h2.data(), SW, 0, 96); }
  printf("probe: %d failing case(s)\n", fails);
  return 0;
}
