// art_planner_b200/csrc/artp_reach.cuh
// Reach-box stages of the pose-validity pipeline: one THREAD per queued reach box, zone tiles staged in shared memory by
// TMA.
//
// A reach box (ValidityCheckerFeet, validity_checker_feet.cpp:32-70) overlaps ~9 x 9 heightfield vertices at the
// shipped geometry: far too little work for a warp (round 1 ran one warp per box at ~1000 warp instructions each, 20 of
// 32 lanes busy, most of it cross-lane plumbing: ballots, shuffles, a shared-memory bloom filter). Here every lane owns a
// whole box and walks its zone serially -- no cross-lane traffic at all, and all 32 lanes do the same kind of work.
// A first version read the zones straight from global memory: 32 lanes x 32 different sectors per load instruction and
// a dependent load -> compare -> branch chain per vertex made it latency / L1-wavefront bound (0.49 ms for 360 k boxes,
// profiles/r02_v1_*). So the zone of every box is first copied into shared memory as ONE 2-D TMA tile
// (cp.async.bulk.tensor.2d; the tile origin is the zone origin rounded down to a multiple of 4 columns -- TMA needs
// 16-byte aligned row starts, an unaligned inner coordinate traps as "illegal instruction", profiles/tma_probe.cu --
// out-of-map elements are zero-filled and never read): 32 tiles in
// flight per warp, one mbarrier per warp, no registers and no L1 wavefronts spent on the gather. TMA tiles start on
// 128-byte boundaries, i.e. every lane's tile starts in bank 0; lanes therefore walk their tiles in a rotated order
// (lane i starts at word i) so that the 32 concurrent shared-memory reads fall into different banks.
//
//   F1  reach_vertex_kernel   the vertex-in-box scan of dCollideHeightfieldZone (heightfield.cpp:1306-1441). A hit
//                             decides the box (a touching reach box leaves the pose's provisional 1 alone); the others are
//                             compacted into a second list (one warp-aggregated atomic).
//   F2  reach_plane_kernel    the plane stage (heightfield.cpp:1474-1617) with the same exact shortcut as the warp stage
//                             (artp_kernels.cuh): only triangles under the 8 box corners can own a contact point; an
//                             earlier kept triangle that epsilon-matches a live candidate sends the box to the exact
//                             grouping stage (C). The bloom filter becomes a direct compare of the approximate normal with
//                             the <= 32 live candidates (same margin argument: kKeyMargin > eps + rsqrt.approx error).
//
// Zones larger than the tile (other robot / map resolution) or not covered by the range tables are routed to the generic
// warp stage by the classify kernel.
#pragma once

#include <cuda.h>

#include "artp_kernels.cuh"

namespace artp {

constexpr int kReachMaxCand = 32;          // live corner candidates kept per box (more -> grouping stage)
constexpr int kReachWarpsPerCta = 1;   // 32 tiles of shared memory per warp: small CTAs pack the SM best
constexpr uint32_t kDeferReachBit = 0x80000000u;   // defer-list entry refers to the reach-box queue

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
      "RWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra RWAIT_DONE;\n\t"
      "bra RWAIT_LOOP;\n\t"
      "RWAIT_DONE:\n\t"
      "}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_tile_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_addr(dst)),
      "l"(map), "r"(smem_addr(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void load_rec(const BoxRec* __restrict__ p, BoxRec& r) {
  const uint4* rp = reinterpret_cast<const uint4*>(p);
  uint4* dst = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < 5; ++i) dst[i] = __ldg(rp + i);
}

__device__ __forceinline__ void rec_to_ctx_side(const float side[3], const BoxRec& r, BoxCtx& b) {
#pragma unroll
  for (int i = 0; i < 9; ++i) b.R1[i] = r.R1[i];
  b.P[0] = r.P[0]; b.P[1] = r.P[1]; b.P[2] = r.P[2];
  b.minB = r.minB; b.maxB = r.maxB;
  b.x0 = r.x0; b.x1 = r.x1; b.z0 = r.z0; b.z1 = r.z1;
  b.side[0] = side[0]; b.side[1] = side[1]; b.side[2] = side[2];
}

// One warp's round of TMA staging: lane i (if `want`) gets the tw x th tile whose origin is its zone origin.
// Returns after all requested tiles have landed. tile stride = c.reach_tile_bytes (multiple of 128).
__device__ __forceinline__ void stage_tiles(const CUtensorMap* tmap, unsigned char* warp_tiles, uint32_t tile_stride,
                                            uint32_t tile_bytes, uint64_t* bar, uint32_t& phase, bool want, int x0, int z0,
                                            int lane) {
  unsigned m = __ballot_sync(kFull, want);
  if (m == 0u) return;
  if (lane == 0) mbar_expect_tx(bar, (uint32_t)__popc(m) * tile_bytes);
  __syncwarp();
  // The TMA instruction takes warp-uniform operands (UTMALDG reads uniform registers; issuing it from 32 lanes with
  // divergent coordinates traps as an illegal instruction): lane 0 issues one copy per requesting lane.
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    const int sx = __shfl_sync(kFull, x0, src), sz = __shfl_sync(kFull, z0, src);
    if (lane == 0) tma_tile_2d(warp_tiles + (size_t)src * tile_stride, tmap, bar, sx & ~3, sz);   // 16-byte aligned rows
  }
  mbar_wait(bar, phase);
  phase ^= 1u;
}

// Vertex stage for one box, one thread, zone in shared memory (tile[zi * tw + xi] = H(x0 + xi, z0 + zi)).
// All-finite zone: every vertex above the box bottom belongs to a kept triangle. Otherwise a vertex counts only if one
// of the (up to six) triangles around it has three finite vertices (heightfield.cpp:1329-1344): with E, W, N, S, NW, SE
// the finiteness of the neighbours (x+1,z), (x-1,z), (x,z+1), (x,z-1), (x-1,z+1), (x+1,z-1) inside the zone, that is
// E&N | W&NW | N&NW | S&SE | E&SE | S&W (Up / Down triangles of the cells (x,z), (x-1,z), (x,z-1), (x-1,z-1)).
// A vertex strictly inside the box lies within the box's vertical extent, so h >= maxB + slack cannot be inside
// (slack = 1e-4 + 4e-6 |maxB|: > 10x the rounding of the rotated coordinates and of maxB itself) and is skipped.
__device__ __forceinline__ bool reach_vertex_hit(const Field& f, const BoxCtx& b, bool allFinite, const float* tile, int tw,
                                                 int lane) {
  const int nX = b.x1 - b.x0 + 1, nZ = b.z1 - b.z0 + 1;
  const float top = b.maxB + (1e-4f + 4e-6f * fabsf(b.maxB));
  const int nWords = tw * nZ;
  int word = lane % nWords;                       // rotated start: concurrent lanes read different banks
  int zi = word / tw, xi = word - zi * tw;
  for (int t = 0; t < nWords; ++t) {
    if (xi < nX) {
      const float h = tile[zi * tw + xi];
      if (h > b.minB && h < top) {
        bool counts = allFinite;
        if (!allFinite) {
          const bool E = xi + 1 < nX && finitef(tile[zi * tw + xi + 1]), W = xi > 0 && finitef(tile[zi * tw + xi - 1]);
          const bool N = zi + 1 < nZ && finitef(tile[(zi + 1) * tw + xi]), S = zi > 0 && finitef(tile[(zi - 1) * tw + xi]);
          const bool NW = xi > 0 && zi + 1 < nZ && finitef(tile[(zi + 1) * tw + xi - 1]);
          const bool SE = xi + 1 < nX && zi > 0 && finitef(tile[(zi - 1) * tw + xi + 1]);
          counts = (E && N) || (W && NW) || (N && NW) || (S && SE) || (E && SE) || (S && W);
        }
        if (counts && vertex_inside(b, (b.x0 + xi) * f.sW, h, (b.z0 + zi) * f.sD)) return true;
      }
    }
    if (++xi == tw) { xi = 0; if (++zi == nZ) zi = 0; }
  }
  return false;
}

__device__ __forceinline__ void tile_cell(const float* tile, int tw, int lx, int lz, float& hA, float& hB, float& hC, float& hD) {
  const float* p = tile + lz * tw + lx;
  hA = p[0]; hB = p[1]; hC = p[tw]; hD = p[tw + 1];
}

// Plane stage for one box, one thread, zone in shared memory. Returns R_FREE / R_HIT / R_DEFER. Same decision
// procedure as box_collide_warp's stages (4)-(5): live corner candidates with their own-plane contacts, then the merge
// screen over the earlier kept triangles; nothing merges => every candidate is its own group base => its own-plane
// contacts are the reference's.
__device__ __forceinline__ int reach_plane_decide(const Field& f, const BoxCtx& b, float cell_margin, const float* tile, int tw,
                                                  int lane) {
  const int nX = b.x1 - b.x0 + 1, nZ = b.z1 - b.z0 + 1;
  const int nCX = nX - 1, nCZ = nZ - 1;
  float cn0[kReachMaxCand], cn2[kReachMaxCand], cn1[kReachMaxCand], cd[kReachMaxCand];
  int cidx[kReachMaxCand];
  int nLive = 0, max_live = -1;
  bool hit_own = false;
  const float h0 = 0.5f * b.side[0], h1 = 0.5f * b.side[1], h2 = 0.5f * b.side[2];
#pragma unroll 1
  for (int corner = 0; corner < 8; ++corner) {
    float px = b.P[0], pz = b.P[2];
    if (corner & 1) { px += h0 * b.R1[0]; pz += h0 * b.R1[6]; } else { px -= h0 * b.R1[0]; pz -= h0 * b.R1[6]; }
    if (corner & 2) { px += h1 * b.R1[1]; pz += h1 * b.R1[7]; } else { px -= h1 * b.R1[1]; pz -= h1 * b.R1[7]; }
    if (corner & 4) { px += h2 * b.R1[2]; pz += h2 * b.R1[8]; } else { px -= h2 * b.R1[2]; pz -= h2 * b.R1[8]; }
    const float gx = px * f.iW, gz = pz * f.iD;
    const int cxl = (int)floorf(gx - cell_margin), cxh = (int)floorf(gx + cell_margin);
    const int czl = (int)floorf(gz - cell_margin), czh = (int)floorf(gz + cell_margin);
    const int nsub = (cxh != cxl || czh != czl) ? 4 : 1;     // extra cells only within cell_margin of a cell boundary
#pragma unroll 1
    for (int sub = 0; sub < nsub; ++sub) {
      if (((sub & 1) && cxh == cxl) || ((sub & 2) && czh == czl)) continue;
      const int ccx = (sub & 1) ? cxh : cxl, ccz = (sub & 2) ? czh : czl;
      if (ccx < b.x0 || ccx >= b.x1 || ccz < b.z0 || ccz >= b.z1) continue;
      float hA, hB, hC, hD;
      tile_cell(tile, tw, ccx - b.x0, ccz - b.z0, hA, hB, hC, hD);
      const bool fA = finitef(hA), fB = finitef(hB), fC = finitef(hC), fD = finitef(hD);
      const bool cA = fA && hA > b.minB, cB = fB && hB > b.minB, cC = fC && hC > b.minB, cD = fD && hD > b.minB;
      const int cell_idx = ((ccx - b.x0) * nCZ + (ccz - b.z0)) * 2;   // emission order: x outer, z inner, Up, Down
#pragma unroll 1
      for (int u = 0; u < 2; ++u) {
        const bool isUp = (u == 0);
        const bool keep = isUp ? ((cA || cB || cC) && (fA && fB && fC)) : ((cB || cC || cD) && (fB && fC && fD));
        if (!keep) continue;
        const int idx = cell_idx + u;
        bool dup = false;                                  // an upright box projects two corners into the same cell
        for (int j = 0; j < nLive; ++j) dup = dup || (cidx[j] == idx);
        if (dup) continue;
        float pl[4];
        cell_plane(f, isUp, ccx, ccz, hA, hB, hC, hD, pl);
        // Liveness: any plane within eps of this one changes the box-plane depth by far less than tau.
        const float Q1 = pl[0] * b.R1[0] + pl[1] * b.R1[3] + pl[2] * b.R1[6];
        const float Q2 = pl[0] * b.R1[1] + pl[1] * b.R1[4] + pl[2] * b.R1[7];
        const float Q3 = pl[0] * b.R1[2] + pl[1] * b.R1[5] + pl[2] * b.R1[8];
        const float B1 = fabsf(b.side[0] * Q1), B2 = fabsf(b.side[1] * Q2), B3 = fabsf(b.side[2] * Q3);
        const float depth = pl[3] + 0.5f * (B1 + B2 + B3) - (pl[0] * b.P[0] + pl[1] * b.P[1] + pl[2] * b.P[2]);
        const float tau = 16.0f * ARTP_EPS * (1.0f + b.side[0] + b.side[1] + b.side[2] + fabsf(b.P[0]) + fabsf(b.P[1]) +
                                              fabsf(b.P[2]) + fabsf(pl[3]));
        if (!(depth >= -tau)) continue;                    // dead: no plane of its would-be group can touch the box
        if (nLive == kReachMaxCand) return R_DEFER;
        cn0[nLive] = pl[0]; cn1[nLive] = pl[1]; cn2[nLive] = pl[2]; cd[nLive] = pl[3];
        cidx[nLive] = idx;
        ++nLive;
        max_live = max(max_live, idx);
        float ctx[4], ctz[4];
        const int nc = box_plane(b, pl, 4, ctx, ctz);
        const int tcx = isUp ? ccx : ccx + 1, tcz = isUp ? ccz : ccz + 1;
        for (int i = 0; i < nc; ++i) hit_own = hit_own || on_tri(f, isUp, tcx, tcz, ctx[i], ctz[i]);
      }
    }
  }
  if (nLive == 0) return R_FREE;
  // Merge screen (greedy grouping, heightfield.cpp:1511-1556): a live candidate is absorbed only by an EARLIER kept
  // triangle whose plane matches it within eps. Emission index = (cxi * nCZ + czi) * 2 + u, so only columns
  // cxi <= max_live / (2 nCZ) can hold earlier triangles. Cells are visited in a lane-rotated order (bank spread).
  const int ncol = min(nCX - 1, max_live / (2 * nCZ)) + 1;
  const int nCells = ncol * nCZ;
  int cell = lane % nCells;
  int czi = cell / ncol, cxi = cell - czi * ncol;
  for (int t = 0; t < nCells; ++t) {
    const int cell_idx = (cxi * nCZ + czi) * 2;
    if (cell_idx < max_live) {
      float hA, hB, hC, hD;
      tile_cell(tile, tw, cxi, czi, hA, hB, hC, hD);
      const bool fA = finitef(hA), fB = finitef(hB), fC = finitef(hC), fD = finitef(hD);
      const bool cA = fA && hA > b.minB, cB = fB && hB > b.minB, cC = fC && hC > b.minB, cD = fD && hD > b.minB;
      const bool keepUp = (cA || cB || cC) && (fA && fB && fC), keepDn = (cB || cC || cD) && (fB && fC && fD);
      if (keepUp || keepDn) {
        const int cx = b.x0 + cxi, cz = b.z0 + czi;
        const float xA = cx * f.sW, xB = (cx + 1) * f.sW, zA = cz * f.sD, zC = (cz + 1) * f.sD;
#pragma unroll 1
        for (int u = 0; u < 2; ++u) {
          const int idx = cell_idx + u;
          if (!(u == 0 ? keepUp : keepDn) || idx >= max_live) continue;
          float c0, c1, c2;   // value-identical to tri_plane's cross product (zero terms dropped)
          if (u == 0) {   // Up (A,B,C): E1 = C-A, E2 = B-A; c = E1 x E2
            const float e1y = hC - hA, e1z = zC - zA, e2x = xB - xA, e2y = hB - hA;
            c0 = -(e1z * e2y); c1 = e1z * e2x; c2 = -(e1y * e2x);
          } else {        // Down (D,B,C): E1 = C-D, E2 = B-D; c = E2 x E1
            const float e1x = xA - xB, e1y = hC - hD, e2y = hB - hD, e2z = zA - zC;
            c0 = -(e2z * e1y); c1 = e2z * e1x; c2 = -(e2y * e1x);
          }
          const float r = rsqrtf(c0 * c0 + c1 * c1 + c2 * c2);
          const float n0 = c0 * r, n2 = c2 * r;
          for (int j = 0; j < nLive; ++j) {
            if (fabsf(n0 - cn0[j]) <= kKeyMargin && fabsf(n2 - cn2[j]) <= kKeyMargin && cidx[j] > idx) {
              float plm[4];
              cell_plane(f, u == 0, cx, cz, hA, hB, hC, hD, plm);
              const float plc[4] = {cn0[j], cn1[j], cn2[j], cd[j]};
              if (plane_match(plm, plc)) return R_DEFER;
            }
          }
        }
      }
    }
    if (++cxi == ncol) { cxi = 0; if (++czi == nCZ) czi = 0; }
  }
  return hit_own ? R_HIT : R_FREE;
}

// F1: vertex scan, one thread per record of the reach-box queue (persistent grid, one 32-record claim per warp).
__global__ void __launch_bounds__(kReachWarpsPerCta * 32)
reach_vertex_kernel(const Checker c, const __grid_constant__ CUtensorMap tmap, const Work w, const BoxRec* __restrict__ recs,
                    const uint32_t* __restrict__ rec_count, uint32_t* __restrict__ work_counter,
                    uint32_t* __restrict__ plane_list, uint32_t* __restrict__ plane_count) {
  extern __shared__ __align__(128) unsigned char reach_smem[];
  __shared__ uint64_t bars[kReachWarpsPerCta];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned char* warp_tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(reach_smem) + 127) & ~(uintptr_t)127) +
                              (size_t)wid * 32 * c.reach_tile_stride;   // TMA destinations: 128-byte aligned
  if (lane == 0) mbar_init(&bars[wid], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  uint32_t phase = 0;
  const uint32_t total = *rec_count;
  for (;;) {
    uint32_t r0 = 0;
    if (lane == 0) r0 = atomicAdd(work_counter, 32u);
    r0 = __shfl_sync(kFull, r0, 0);
    if (r0 >= total) break;
    const uint32_t ri = r0 + lane;
    BoxRec r;
    bool alive = false;
    if (ri < total) {
      load_rec(recs + ri, r);
      // another box of the item (or state of the edge) already failed: nothing can change the verdict (perf only)
      alive = *(volatile const uint8_t*)(w.valid + item_slot(w, r.item)) != 0;
    }
    __syncwarp();   // every lane is done with the previous round's tiles
    stage_tiles(&tmap, warp_tiles, c.reach_tile_stride, c.reach_tile_bytes, &bars[wid], phase, alive, r.x0, r.z0, lane);
    bool survive = false;
    if (alive) {
      BoxCtx b;
      rec_to_ctx_side(c.side[1], r, b);
      survive = !reach_vertex_hit(c.f[1], b, (r.flags & REC_ALLFINITE) != 0,
                                  reinterpret_cast<const float*>(warp_tiles + (size_t)lane * c.reach_tile_stride) + (r.x0 & 3),
                                  c.reach_tw, lane);
    }
    const unsigned m = __ballot_sync(kFull, survive);
    if (m) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(plane_count, (uint32_t)__popc(m));
      base = __shfl_sync(kFull, base, 0);
      if (survive) plane_list[base + __popc(m & ((1u << lane) - 1u))] = ri;
    }
  }
}

// F2: plane stage, one thread per entry of the plane list.
__global__ void __launch_bounds__(kReachWarpsPerCta * 32)
reach_plane_kernel(const Checker c, const __grid_constant__ CUtensorMap tmap, const Work w, const BoxRec* __restrict__ recs,
                   const uint32_t* __restrict__ plane_list, const uint32_t* __restrict__ plane_count,
                   uint32_t* __restrict__ work_counter, uint32_t* __restrict__ defer_count, uint32_t* __restrict__ defer_list) {
  extern __shared__ __align__(128) unsigned char reach_smem[];
  __shared__ uint64_t bars[kReachWarpsPerCta];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned char* warp_tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(reach_smem) + 127) & ~(uintptr_t)127) +
                              (size_t)wid * 32 * c.reach_tile_stride;   // TMA destinations: 128-byte aligned
  if (lane == 0) mbar_init(&bars[wid], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  uint32_t phase = 0;
  const uint32_t total = *plane_count;
  for (;;) {
    uint32_t q0 = 0;
    if (lane == 0) q0 = atomicAdd(work_counter, 32u);
    q0 = __shfl_sync(kFull, q0, 0);
    if (q0 >= total) break;
    const uint32_t q = q0 + lane;
    BoxRec r;
    bool alive = false;
    uint32_t ri = 0, slot = 0;
    if (q < total) {
      ri = __ldg(plane_list + q);
      load_rec(recs + ri, r);
      slot = item_slot(w, r.item);
      alive = *(volatile const uint8_t*)(w.valid + slot) != 0;
    }
    __syncwarp();
    stage_tiles(&tmap, warp_tiles, c.reach_tile_stride, c.reach_tile_bytes, &bars[wid], phase, alive, r.x0, r.z0, lane);
    if (alive) {
      BoxCtx b;
      rec_to_ctx_side(c.side[1], r, b);
      const int res = reach_plane_decide(c.f[1], b, c.cell_margin,
                                         reinterpret_cast<const float*>(warp_tiles + (size_t)lane * c.reach_tile_stride) + (r.x0 & 3),
                                         c.reach_tw, lane);
      if (res == R_DEFER) defer_list[atomicAdd(defer_count, 1u)] = ri | kDeferReachBit;
      else if (res == R_FREE) w.valid[slot] = 0;          // a reach box that does not touch: pose invalid
    }
  }
}

}  // namespace artp
