// art_planner_b200/csrc/artp_tiles.cuh
// Stage B of the pose-validity pipeline: one WARP per queued box, the box's zone staged in shared memory by TMA.
//
// Every queued box needs the heights of its zone -- the sub-rectangle of the heightfield its AABB overlaps, ~9 x 9
// vertices for a reach box, ~40 x 40 for the torso (heightfield.cpp:1880-1892) -- several times: vertex scan, corner
// candidates, merge screen. The warp copies the zone into shared memory as ONE 2-D TMA tile
// (cp.async.bulk.tensor.2d with the tensor map of the layer; tile origin = zone origin rounded down to a multiple of 4
// columns because TMA needs 16-byte aligned row starts -- an unaligned inner coordinate traps as "illegal instruction",
// profiles/tma_probe.cu; out-of-map elements are zero-filled and never read) and double-buffers: while box i is decided
// out of tile slot i & 1, the tile of box i + 1 is already in flight into the other slot (one mbarrier per slot).
// No address arithmetic, L1 wavefronts or registers are spent on the gather, and all later reads are shared-memory reads.
//
// Two queues feed two launches of the same kernel: reach boxes (small tiles, 8 warps per CTA) and torso boxes (big
// tiles, 4 warps per CTA). A box whose zone does not fit its tile (cannot happen for the sizes the tiles are derived
// from) goes to the exact grouping stage. Round 2 also tried one THREAD per reach box over the staged tiles (no
// cross-lane traffic at all): SIMT divergence left 8-10 of 32 lanes busy and it was slower (profiles/r02_v1_reach_*).
#pragma once

#include <cuda.h>

#include "artp_kernels.cuh"

namespace artp {

constexpr uint32_t kDeferReachBit = 0x80000000u;   // defer-list entry refers to the reach-box queue
constexpr int kTileChunk = 8;                      // records claimed per atomic (the tile pipeline restarts per chunk)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "TWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra TWAIT_DONE;\n\t"
      "bra TWAIT_LOOP;\n\t"
      "TWAIT_DONE:\n\t"
      "}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_tile_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_addr(dst)),
      "l"(map), "r"(smem_addr(bar)), "r"(c0), "r"(c1)
      : "memory");
}

struct TileCfg {
  int tw, th;              // tile extent in floats (tw a multiple of 4; a zone may be at most tw - 3 wide, th high)
  uint32_t bytes, stride;  // tw * th * 4, and that rounded up to 128 bytes
  int x_off;               // first stored vertex column of the layer (map window), a multiple of 4
  int slots;               // tile slots per warp: 2 = the next box's tile is prefetched while this one is decided, 1 = none
};
constexpr int kMaxTileWarps = 8;

// map0 / map1: tiles of `elevation` / `elevation_masked` (same tile extent). queue_bit tags defer-list entries.
__global__ void __launch_bounds__(kMaxTileWarps * 32, 3)
box_tiles_warp_kernel(const Checker c, const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                      const TileCfg tc, const Work w, const BoxRec* __restrict__ recs, const uint32_t* __restrict__ rec_count,
                      uint32_t* __restrict__ work_counter, uint32_t* __restrict__ defer_count, uint32_t* __restrict__ defer_list,
                      uint32_t queue_bit, int force_defer) {
  extern __shared__ __align__(128) unsigned char tile_smem[];
  __shared__ WarpScratch ws_all[kMaxTileWarps];
  __shared__ uint64_t bars[kMaxTileWarps][2];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  WarpScratch& ws = ws_all[wid];
  unsigned char* slots = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tile_smem) + 127) & ~(uintptr_t)127) +
                         (size_t)wid * tc.slots * tc.stride;   // TMA destinations: 128-byte aligned
  if (lane == 0) { mbar_init(&bars[wid][0], 1); mbar_init(&bars[wid][1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  uint32_t phase[2] = {0u, 0u};
  const uint32_t total = *rec_count;
  // lane 0 starts the copy of record ri's zone into tile slot `slot`
  auto prefetch = [&](uint32_t ri, int slot) {
    const uint4 zr = __ldg(reinterpret_cast<const uint4*>(recs + ri) + 3);        // R.., minB, maxB | x0 x1 z0 z1 -> words 14..17
    const uint4 zr2 = __ldg(reinterpret_cast<const uint4*>(recs + ri) + 4);
    const int x0 = (int)zr.z, z0 = (int)zr2.x;                                    // BoxRec words: 14 x0, 15 x1, 16 z0, 17 z1
    const uint32_t fl = zr2.w;                                                    // word 19: flags
    if (lane == 0) {
      mbar_expect_tx(&bars[wid][slot], tc.bytes);
      tma_tile_2d(slots + (size_t)slot * tc.stride, (fl & 7u) ? &map1 : &map0, &bars[wid][slot], (x0 & ~3) - tc.x_off, z0);
    }
  };
  for (;;) {
    uint32_t r0 = 0;
    if (lane == 0) r0 = atomicAdd(work_counter, (uint32_t)kTileChunk);
    r0 = __shfl_sync(kFull, r0, 0);
    if (r0 >= total) break;
    const uint32_t r1 = min(r0 + (uint32_t)kTileChunk, total);
    if (force_defer) {
      if (lane == 0) for (uint32_t ri = r0; ri < r1; ++ri) defer_list[atomicAdd(defer_count, 1u)] = ri | queue_bit;
      continue;
    }
    for (uint32_t ri = r0; ri < r1; ++ri) {
      const int slot = (int)(ri - r0) & (tc.slots - 1);
      if (tc.slots == 1 || ri == r0) { __syncwarp(); prefetch(ri, slot); }   // every lane is done with the slot
      // every lane reads the whole 80-byte record itself: five 16-byte loads from one address per warp (broadcasts)
      BoxRec r;
      {
        const uint4* rp = reinterpret_cast<const uint4*>(recs + ri);
        uint4* dst = reinterpret_cast<uint4*>(&r);
#pragma unroll
        for (int i = 0; i < 5; ++i) dst[i] = __ldg(rp + i);
      }
      if (tc.slots == 2 && ri + 1 < r1) { __syncwarp(); prefetch(ri + 1, slot ^ 1); }   // the other slot's box (ri - 1) is finished
      const uint32_t slot_item = r.item;
      const bool foot = (r.flags & 7) != 0;
      // another box of the item (or state of the edge) already failed: nothing can change the verdict (perf only)
      int dead = 0;
      if (lane == 0) dead = (*(volatile const uint8_t*)(w.valid + slot_item) == 0);
      dead = __shfl_sync(kFull, dead, 0);
      mbar_wait(&bars[wid][slot], phase[slot]);       // the tile must land before its slot can be reused
      phase[slot] ^= 1u;
      if (dead) continue;
      BoxCtx b;
      rec_to_ctx(c, r, b);
      int res;
      if ((b.x1 - b.x0) + 4 > tc.tw || (b.z1 - b.z0) + 1 > tc.th) {
        res = R_DEFER;                                // zone larger than the tile: exact grouping stage
      } else {
        const ZoneView zv{reinterpret_cast<const float*>(slots + (size_t)slot * tc.stride) + (b.x0 & 3), tc.tw};
        res = box_collide_warp<true>(foot ? c.f[1] : c.f[0], b, zv, ws, lane, c.cell_margin, (r.flags & REC_NEEDS_REDUCE) != 0,
                                     (r.flags & REC_ALLFINITE) != 0, (r.flags & REC_MERGEFREE) != 0);
      }
      if (lane == 0) {
        if (res == R_DEFER) defer_list[atomicAdd(defer_count, 1u)] = ri | queue_bit;
        else if ((!foot && res == R_HIT) || (foot && res == R_FREE)) w.valid[slot_item] = 0;
      }
    }
  }
}

}  // namespace artp
