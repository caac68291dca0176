// tma.cuh — the sm_100a bulk-tensor copy engine (TMA) and its mbarrier plumbing as
// thin inline-PTX wrappers, plus the copy-class kernels built on them.
//
// The bandwidth-bound rows of the hot path (`composite` of opaque
// picture-cache tiles — "pure bandwidth"; row 16 clears) move whole rectangles of pixels and
// compute nothing.  Routing them through the command-binning tile kernel costs barriers, ballots
// and per-pixel address arithmetic for what is a strided 2-D copy.  Here ONE elected thread per CTA
// drives the copy engine: 2-D `cp.async.bulk.tensor` loads of 256x16-pixel boxes (16 KB) from the tile
// texture into shared memory, completion signalled on an mbarrier, then a bulk-tensor store of the
// same shared-memory box into the framebuffer — a STAGES-deep ring, so every SM keeps
// STAGES x 16 KB of reads in flight and no register, LSU instruction or barrier is spent per pixel.
// Boxes that are not wholly inside the instance's rect (ragged right / bottom edges) are copied by
// the CTA's threads with plain vector accesses.
//
// SASS evidence (profiles/): UTMALDG.2D / UTMASTG.2D / SYNCS (mbarrier) in wr_composite_copy.
#pragma once
#ifndef WRCU_HOSTEMU
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <stdint.h>

#define WR_TMA_BOX_W 256  // pixels (u32 elements): 1 KB contiguous per box row
#define WR_TMA_BOX_H 16
#define WR_TMA_BOX_BYTES (WR_TMA_BOX_W * WR_TMA_BOX_H * 4)
#define WR_TMA_STAGES 4
#define WR_TMA_THREADS 128

__device__ __forceinline__ uint32_t wr_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void wr_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wr_smem_u32(bar)), "r"(count) : "memory");
}
// make the barrier initialisation visible to the async proxy (the copy engine arrives on it)
__device__ __forceinline__ void wr_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy writes to shared memory → visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void wr_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void wr_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wr_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wr_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(wr_smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// 2-D tiled load: box at element coordinates (x, y) of the tensor described by `map` → shared memory;
// the engine adds the box's bytes to the barrier's transaction count as they land.  Parts of the box
// outside the tensor are zero-filled (and still counted).
__device__ __forceinline__ void wr_tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(wr_smem_u32(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(wr_smem_u32(bar))
      : "memory");
}
// 2-D tiled store: shared memory box → tensor at (x, y); parts outside the tensor are dropped.
__device__ __forceinline__ void wr_tma_store_2d(const CUtensorMap* map, int x, int y, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(x),
               "r"(y), "r"(wr_smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void wr_tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's committed store groups still READ their shared-memory source
template <int N>
__device__ __forceinline__ void wr_tma_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void wr_tma_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// A tensor map that lives in global memory and was (re)written since the copy engine last fetched it
// — a texture handle recycled for another texture — must be re-acquired through the tensormap proxy
// before use, or the engine may run on its cached copy of the old descriptor.
__device__ __forceinline__ void wr_tma_acquire_map(const CUtensorMap* map) {
  asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(map) : "memory");
}
__device__ __forceinline__ void wr_tma_prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
#endif  // !WRCU_HOSTEMU
