// weed_kernels.cuh — the per-frame kernels (sm_100a).  See DESIGN.md for the data flow.
//
//   id order   k_cell_key      K1  cell key + arrival rank (one L2 atomic per entity)
//   cells      k_cell_scan     K2  exclusive scan, single pass, decoupled look-back; lists the cells above BIG_CELL
//   id order   k_scatter_ids   K3a ids into their cell segment (arrival order)
//   cells      k_sort_big_cells    the id lists of piles (cells above 128 entities), a block each
//   id order   k_slot_rank     K3b stable position inside the cell (ascending id): count, or binary search in a sorted pile
//   id order   k_build_slots   K3c Verlet integration (K5) + derived speed/angle fused, one 32 B
//                                  slot record per entity (the only scattered write)
//   slot order k_slot_prep     K3d query positions, candidate records, scan windows, list heads,
//                                  first bounds pass, first sweep's input
//   slot order k_neighbors2    K4  capped ordered gather, thread per entity: float32 pre-filter into a queue, converged
//                                  binary64 pass, API rows as slot-major planes (k_neighbors_wide: warp per entity
//                                  for long rows)
//   slot order k_beyond_cap    K4b lower-id partners a capped row lost.  Few capped rows: a warp per capped entity
//                                  resumes its scan.  Many (a settled bed): reverse edges — every entity reports
//                                  itself to the partners of its own row (count here, k_back_alloc, k_back_write,
//                                  k_back_sort)
//              k_sort_lists    K4c explicit (asymmetric) pair lists into slot order; parallel branch
//   slot order k_sweep<F,L>    K6  circle-circle correction, J-order (+ next sweep's bounds); k_sweep_heavy beside it:
//                                  a warp per entity with an overflow-pool row or a resumed scan
//                                  (k_sweep_tile: the TMA-staged form, selectable, slower)
//   id order   k_writeback     WB  gather results by id; per-tile outgoing pair counts
//   tiles      k_pair_scan     K7a prefix of the tile counts, pair count
//   id order   k_pair_emit     K7b collisionData emission by the tiles below the cap
#pragma once
#include <cooperative_groups.h>
#include <type_traits>
#include <math_constants.h>

#include "weed_device.cuh"

namespace weed {
namespace cg = cooperative_groups;

static constexpr int SCAN_THREADS = 512;
static constexpr int SCAN_ITEMS = 16;   // per thread: the look-back chain is 1 / (512 * 16) of the cell count long
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
static constexpr int WB_THREADS = 256;

// ---- decoupled look-back ---------------------------------------------------------------
// status word: [63:34] epoch, [33:32] flag (1 = tile aggregate, 2 = inclusive prefix), [31:0] value.
// Entries of other epochs read as "not ready", so the array never needs clearing.
__device__ __forceinline__ unsigned long long lb_pack(uint32_t epoch, uint32_t flag, uint32_t v) {
  return ((unsigned long long)(epoch & 0x3FFFFFFFu) << 34) | ((unsigned long long)flag << 32) | v;
}
__device__ __forceinline__ unsigned long long lb_load(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void lb_store(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Called by all 32 lanes of warp 0.  Publishes this tile's aggregate, walks back over the
// predecessors, publishes the inclusive prefix and returns the exclusive prefix.
__device__ __forceinline__ uint32_t lb_exclusive(unsigned long long* status, uint32_t tile, uint32_t agg,
                                                 uint32_t epoch) {
  const uint32_t lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0) lb_store(status, lb_pack(epoch, 2, agg));
    return 0;
  }
  if (lane == 0) lb_store(status + tile, lb_pack(epoch, 1, agg));
  uint32_t exclusive = 0;
  int look = (int)tile - 1;
  const uint32_t ep = epoch & 0x3FFFFFFFu;
  while (true) {
    const int idx = look - (int)lane;
    uint32_t flag, val;
    if (idx >= 0) {
      unsigned long long w;
      do {
        w = lb_load(status + idx);
      } while ((uint32_t)(w >> 34) != ep || ((w >> 32) & 3u) == 0);
      flag = (uint32_t)(w >> 32) & 3u; val = (uint32_t)w;
    } else {
      flag = 2; val = 0;  // virtual tile before the first one
    }
    const uint32_t inc = __ballot_sync(0xffffffffu, flag == 2);
    if (inc) {
      const uint32_t first = __ffs(inc) - 1;
      uint32_t v = lane <= first ? val : 0;
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      exclusive += v;
      break;
    }
    uint32_t v = val;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    exclusive += v;
    look -= 32;
  }
  if (lane == 0) lb_store(status + tile, lb_pack(epoch, 2, exclusive + agg));
  return exclusive;
}

// ---- frame bookkeeping -----------------------------------------------------------------
__global__ void k_spatial_begin(Counters* ctr) {
  if (threadIdx.x == 0) {
    ctr->epoch++;
    ctr->anyCapped = 0;
    ctr->xoverRows = 0;
    ctr->xpoolUsed = 0;
    ctr->nCapped = 0;
    ctr->nSort = 0;
    ctr->nBigCells = 0;
    ctr->nHeavy = 0;
    ctr->explicitPairs = 0;
    ctr->maxCellFrame = 0;
    ctr->tBegin = global_timer_ns();
  }
}
__global__ void k_physics_end(Counters* ctr) {
  if (threadIdx.x == 0) {
    ctr->frame++; ctr->frames++;
    const unsigned long long dt = global_timer_ns() - ctr->tBegin;
    ctr->frameNs = dt > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)dt;
  }
}

// ---- K1: cell key + arrival rank (spatial_worker.js:146-169) ----------------------------
__global__ void __launch_bounds__(256)
k_cell_key(GridDims g, const float4* __restrict__ DP, const uint8_t* __restrict__ F,
           uint32_t* __restrict__ key, uint32_t* __restrict__ rank, uint32_t* __restrict__ cellCount) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.N) return;
  const uint32_t f = F[i];
  const float4 p = DP[i];
  // skip inactive (:148) and NaN positions (:153)
  if (!(f & F_T_ACTIVE) || p.x != p.x || p.y != p.y) { key[i] = KEY_INVALID; return; }
  int32_t col, row;
  cell_of(g, p.x, p.y, col, row);
  const uint32_t cell = (uint32_t)row * (uint32_t)g.cols + (uint32_t)col;
  // arrival rank inside the cell; the stable (ascending id) position is fixed up in K3b
  key[i] = cell;
  rank[i] = atomicAdd(&cellCount[cell], 1u);
}

// ---- K2: exclusive scan of the cell histogram, also clears it for the next frame ---------
// cells above BIG_CELL entities get their id lists sorted by a block (k_sort_big_cells, below)
static constexpr uint32_t BIG_CELL = 128, BIG_CELL_SORT_CAP = 4096;
__device__ __forceinline__ bool big_cell_sorted(uint32_t n) { return n > BIG_CELL && n <= BIG_CELL_SORT_CAP; }

__global__ void __launch_bounds__(SCAN_THREADS)
k_cell_scan(uint32_t* __restrict__ cellCount, uint32_t* __restrict__ cellStart, uint32_t numTiles,
            unsigned long long* status, Counters* ctr, const int32_t* __restrict__ slabCuts, uint32_t cols, int32_t halo,
            uint32_t* __restrict__ bigCells, uint32_t bigCap) {
  __shared__ uint32_t s_tile, s_excl, s_warp[SCAN_THREADS / 32], s_max[SCAN_THREADS / 32];
  if (threadIdx.x == 0) s_tile = atomicAdd(&ctr->scanTile, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t epoch = ctr->epoch;
  constexpr int V = SCAN_ITEMS / 4;
  const size_t i0 = (size_t)tile * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  uint4 c[V];                                                  // arrays are padded to whole tiles
  // a slab only has entities in its own rows + halo: the tiles outside hold zeros, no need to read or clear them
  bool populated = true;
  if (slabCuts) {
    const long long lo = (long long)(slabCuts[0] - halo) * cols, hi = (long long)(slabCuts[1] + halo) * cols;
    const long long t0 = (long long)tile * SCAN_TILE;
    populated = t0 + SCAN_TILE > lo && t0 < hi;
  }
  if (populated) {
#pragma unroll
    for (int v = 0; v < V; v++) c[v] = reinterpret_cast<const uint4*>(cellCount + i0)[v];
#pragma unroll
    for (int v = 0; v < V; v++) reinterpret_cast<uint4*>(cellCount + i0)[v] = make_uint4(0, 0, 0, 0);
  } else {
#pragma unroll
    for (int v = 0; v < V; v++) c[v] = make_uint4(0, 0, 0, 0);
  }
  uint32_t tsum = 0, tmax = 0;
#pragma unroll
  for (int v = 0; v < V; v++) {
    tsum += c[v].x + c[v].y + c[v].z + c[v].w;
    tmax = max(tmax, max(max(c[v].x, c[v].y), max(c[v].z, c[v].w)));
  }
  if (tmax > BIG_CELL) {                                       // piles: their lists are sorted by a block each
#pragma unroll
    for (int v = 0; v < V; v++) {
      const uint32_t cv[4] = {c[v].x, c[v].y, c[v].z, c[v].w};
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (cv[u] > BIG_CELL) {
          const uint32_t at = atomicAdd(&ctr->nBigCells, 1u);
          if (at < bigCap) bigCells[at] = (uint32_t)(i0 + (size_t)v * 4 + u);
        }
    }
  }
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = tsum;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += v;
  }
  for (int o = 16; o; o >>= 1) tmax = max(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
  if (lane == 31) s_warp[warp] = inc;
  if (lane == 0) s_max[warp] = tmax;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0;
    uint32_t winc = w;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= (uint32_t)o) winc += v;
    }
    if (lane < SCAN_THREADS / 32) s_warp[lane] = winc - w;  // exclusive warp offsets
    const uint32_t agg = __shfl_sync(0xffffffffu, winc, 31);
    const uint32_t excl = lb_exclusive(status, tile, agg, epoch);
    uint32_t m = lane < SCAN_THREADS / 32 ? s_max[lane] : 0;
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
      s_excl = excl;
      if (m > 0) atomicMax(&ctr->maxCellFrame, m);
      if (tile == numTiles - 1) {
        ctr->scanTile = 0;                 // every tile has drawn its index by now
        ctr->activeInGrid = excl + agg;
      }
    }
  }
  __syncthreads();
  uint32_t e = s_excl + s_warp[warp] + (inc - tsum);
#pragma unroll
  for (int v = 0; v < V; v++) {
    uint4 o4;
    o4.x = e; e += c[v].x; o4.y = e; e += c[v].y; o4.z = e; e += c[v].z; o4.w = e; e += c[v].w;
    reinterpret_cast<uint4*>(cellStart + i0)[v] = o4;
  }
}

// ---- K3a: ids into their cell segment, arrival order ------------------------------------
__global__ void __launch_bounds__(256)
k_scatter_ids(uint32_t N, const uint32_t* __restrict__ key, const uint32_t* __restrict__ rank,
              const uint32_t* __restrict__ cellStart, uint32_t* __restrict__ arrIds,
              const uint32_t* __restrict__ GID) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint32_t k = key[i];
  if (k == KEY_INVALID) return;
  arrIds[cellStart[k] + rank[i]] = GID ? GID[i] : i;   // the id that orders the cell list
}

// ---- K3b + K5: slot records, Verlet integration, derived properties ----------------------
// moveBallsVerlet (physics_worker.js:264-315), collisionCount reset (:174-177) and
// updateDerivedProperties (:591-603; it only reads the vx,vy stored by the integration, so it
// commutes with the constraint substeps).  INTEGRATE=false builds the slot records of the
// unchanged state (weed_spatial on its own).
//
// The only scattered write is ONE full 32-byte sector per entity (the slot record SA); every
// other slot-order array is produced by the coalesced k_slot_prep pass.
struct ById {
  float4* DP;        // x, y, px, py
  float2* ACC;       // ax, ay
  float4* AT;        // maxVel, radius, visualRange, velocityAngle
  float4* V;         // vx, vy, speed, -
  uint8_t* F;        // flag bits
  uint8_t* CC;       // collisionCount
  uint32_t* GID;     // slab mode: global entity id of each local slot (nullptr: id = index)
  uint8_t* ET;       // Transform.entityType (device-side systems only)
};
struct BySlot {
  float4* SA;        // slot record: SA[2s] = (x, y, radius, flagword), SA[2s+1] = (px, py, visualRange, id bits)
  float2* QXY;       // position at grid-build time (query position)
  float4* CXY;       // candidate record of the thread-per-entity scan: query x, y, visualRange, id bits;
                     // bit 31 of the id word = "centre cell outside the grid" (CX_EDGE)
  int4* WIN;         // clamped scan window of the entity: r0, r1, c0, c1 (r0 > r1: no scan)
  uint4* PW;         // what a neighbour needs about me in one 16 B gather: visualRange bits, id,
                     // r0 | r1 << 16, c0 | c1 << 16 (grid dimensions are limited to 65535)
  float4* GA;        // substep ping-pong: x, y, radius, flagword
  float4* GB;
  float2* PXY;       // px, py after the first substep
  uint32_t* NCNT;    // neighbor count of the API row | entries of the internal row << 16
  uint32_t* NST;     // internal rows, transposed: NST[k * Npad + slot]
  uint32_t* XHEAD;   // explicit incoming pairs: list head (0 = empty, else node + 1)
  uint32_t* XNEXT;   // next link, indexed by the OWNER's row position (k * Npad + slot)
  OutRec* OUT;       // last-substep result
  uint32_t* SLID;    // slab mode: local index of the entity in this slot (nullptr: id = index)
  uint32_t* LSLOT;   // capped rows: slot of the last listed entry (SLOT_NONE: row not capped)
  uint32_t* CAPLIST; // slots whose row hit the cap this frame (unordered)
  uint32_t* SORTLIST;// slots that received explicit pairs this frame (unordered)
  uint32_t* BCNT;    // capped rows, dense regime: lost lower-id partners reported to this slot; BCUR: the write cursor
  uint32_t* BCUR;
  uint32_t* HEAVY;   // slots marked F_XPOOL / F_XOVER this frame (unordered): a warp each in the sweeps (k_sweep_heavy)
  uint32_t* XPID;    // F_XPOOL: which row of the overflow pool continues this entity's internal row
  uint32_t* XR;      // overflow pool: XPOOL_ROW words per pool row (entity-major: one entity's words are consecutive)
  uint32_t* XRCNT;   // entries in each pool row
  TileDesc* TD;      // one descriptor per TILE slots (k_slot_prep), or nullptr when no tiled kernel runs
};

// The same two structs in device memory, for the out-of-line slow paths of the sweep: a function that is
// not inlined takes its arguments by address, and the address of a kernel parameter forces EVERY thread
// to copy the struct to local memory on entry (measured: 2 GB of extra DRAM writes per sweep at 16M).
struct FrameConst { GridDims g; BySlot s; };

// stable position: number of ids in my cell smaller than mine (cell lists are ascending in the
// reference because it inserts i = 0..N-1 in order, spatial_worker.js:146,168).  A kernel of its
// own: the dependent gathers key -> cellStart -> ids of the cell need occupancy, not registers.
// Piles.  A settled bed puts hundreds to thousands of entities into one cell (config 4 after 300 frames:
// 3.6 M entities in cells of 500-1900), and counting the smaller ids of a cell costs occupancy^2 per cell
// (6.6 ms per frame there).  Cells above BIG_CELL entities are listed by k_cell_scan; one block sorts each
// list in shared memory (bitonic), and k_slot_rank finds its rank by binary search.  Lists beyond
// BIG_CELL_SORT_CAP stay unsorted and keep the linear count (both kernels apply the same rule).

__global__ void __launch_bounds__(256)
k_sort_big_cells(const uint32_t* __restrict__ bigCells, uint32_t bigCap, const Counters* __restrict__ ctr,
                 const uint32_t* __restrict__ cellStart, uint32_t* __restrict__ arrIds) {
  __shared__ uint32_t sm[BIG_CELL_SORT_CAP];
  const uint32_t nBig = min(ctr->nBigCells, bigCap);
  for (uint32_t w = blockIdx.x; w < nBig; w += gridDim.x) {
    const uint32_t c = bigCells[w];
    const uint32_t s0 = cellStart[c], n = cellStart[c + 1] - s0;
    if (!big_cell_sorted(n)) continue;              // block-uniform
    uint32_t P = 256;
    while (P < n) P <<= 1;
    for (uint32_t k = threadIdx.x; k < P; k += blockDim.x) sm[k] = k < n ? arrIds[s0 + k] : 0xFFFFFFFFu;
    __syncthreads();
    for (uint32_t k = 2; k <= P; k <<= 1)
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        for (uint32_t idx = threadIdx.x; idx < P; idx += blockDim.x) {
          const uint32_t ixj = idx ^ j;
          if (ixj > idx) {
            const uint32_t a = sm[idx], b = sm[ixj];
            if ((a > b) == ((idx & k) == 0)) { sm[idx] = b; sm[ixj] = a; }
          }
        }
        __syncthreads();
      }
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) arrIds[s0 + k] = sm[k];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
k_slot_rank(GridDims g, const uint8_t* __restrict__ F, const uint32_t* __restrict__ GID,
            const uint32_t* __restrict__ key, const uint32_t* __restrict__ cellStart,
            const uint32_t* __restrict__ arrIds, uint32_t* __restrict__ slotOf) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.N) return;
  const uint32_t k = key[i];
  if (!(F[i] & F_T_ACTIVE) || k == KEY_INVALID) { slotOf[i] = SLOT_NONE; return; }
  const uint32_t gid = GID ? GID[i] : i;
  const uint32_t s0 = cellStart[k], s1 = cellStart[k + 1];
  uint32_t r = 0;
  if (big_cell_sorted(s1 - s0)) {                  // k_sort_big_cells left this list in id order
    uint32_t lo = s0, hi = s1;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (arrIds[mid] < gid) lo = mid + 1; else hi = mid;
    }
    r = lo - s0;
  } else {
    for (uint32_t t = s0; t < s1; t++) r += arrIds[t] < gid;
  }
  slotOf[i] = s0 + r;
}

#ifndef WEED_BUILD_MINBLOCKS
#define WEED_BUILD_MINBLOCKS 1
#endif
template <bool INTEGRATE>
__global__ void __launch_bounds__(256, WEED_BUILD_MINBLOCKS)
k_build_slots(GridDims g, const Params* __restrict__ pp, int subSteps, bool afterSpatial, ById d, BySlot s,
              const uint32_t* __restrict__ key, const uint32_t* __restrict__ cellStart,
              const uint32_t* __restrict__ arrIds, uint32_t* __restrict__ slotOf,
              const int32_t* __restrict__ slabCuts) {     // {begin, end} of the slab in device memory, or nullptr
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.N) return;
  const uint32_t f = d.F[i];
  const uint32_t k = key[i];
  if (!(f & F_T_ACTIVE)) { slotOf[i] = SLOT_NONE; return; }   // inactive: nothing reads or writes it
  const Params p = *pp;
  float4 dp = d.DP[i];
  const float4 at = d.AT[i];
  const float x0 = dp.x, y0 = dp.y;
  uint32_t cc = d.CC[i];
  uint32_t moved = 0;
  if (INTEGRATE) {
    if (f & F_RB_ACTIVE) cc = 0;                                   // :174-177
    if ((f & F_DYNAMIC_MASK) == F_DYNAMIC_VAL) {
      const float2 a = d.ACC[i];
      double ddx = dmul(dsub((double)dp.x, (double)dp.z), p.damping);          // :275
      double ddy = dmul(dsub((double)dp.y, (double)dp.w), p.damping);
      ddx = dadd(ddx, dadd(p.gravityScaleX, dmul((double)a.x, p.dtRatio)));    // :279
      ddy = dadd(ddy, dadd(p.gravityScaleY, dmul((double)a.y, p.dtRatio)));
      const double maxSpeed = at.x > 0 ? (double)at.x : 100.0;                 // :284
      ddx = js_max(-maxSpeed, js_min(maxSpeed, ddx));                          // :297-298
      ddy = js_max(-maxSpeed, js_min(maxSpeed, ddy));
      dp.z = dp.x; dp.w = dp.y;                                                // :305-306
      dp.x = fround(dadd((double)x0, ddx));                                    // :301-302
      dp.y = fround(dadd((double)y0, ddy));
      moved = F_MOVED;
      float4 v;
      // :309-310; x / 1 is x (a headless run has dtRatio = 1): skip the two binary64 divisions
      const bool unit = p.dtRatio == 1.0;
      v.x = fround(unit ? ddx : ddiv(ddx, p.dtRatio));
      v.y = fround(unit ? ddy : ddiv(ddy, p.dtRatio));
      const double sp = __dsqrt_rn(dadd(dmul((double)v.x, (double)v.x), dmul((double)v.y, (double)v.y)));
      v.z = fround(sp); v.w = 0.f;
      d.V[i] = v;
      if (a.x != 0.f || a.y != 0.f || a.x != a.x || a.y != a.y) d.ACC[i] = make_float2(0.f, 0.f);  // :313-314
      if (sp > p.minSpeedForRotation)                                          // :600-602
        reinterpret_cast<float*>(d.AT + i)[3] = fround(dadd(atan2((double)v.y, (double)v.x), 1.5707963267948966));
    } else if (f & F_RB_ACTIVE) {
      // static body: derived properties from whatever vx,vy the host stored
      float4 v = d.V[i];
      const double sp = __dsqrt_rn(dadd(dmul((double)v.x, (double)v.x), dmul((double)v.y, (double)v.y)));
      v.z = fround(sp);
      d.V[i] = v;
      if (sp > p.minSpeedForRotation)
        reinterpret_cast<float*>(d.AT + i)[3] = fround(dadd(atan2((double)v.y, (double)v.x), 1.5707963267948966));
    }
  }
  if (k == KEY_INVALID) {
    // active but NaN position: never in the grid, never collides (every comparison of the
    // sweep is false), but the integration and the boundary pass still apply per axis.
    if (INTEGRATE) {
      if ((f & F_DYNAMIC_MASK) == F_DYNAMIC_VAL)
        for (int st = 0; st < subSteps; st++) apply_bounds(g, p.boundaryElasticity, at.y, dp.x, dp.y, dp.z, dp.w);
      d.DP[i] = dp;
      d.CC[i] = (uint8_t)cc;
    }
    slotOf[i] = SLOT_NONE;
    return;
  }
  // stable position: number of ids in my cell smaller than mine (cell lists are ascending
  // in the reference because it inserts i = 0..N-1 in order, spatial_worker.js:146,168)
  const uint32_t gid = d.GID ? d.GID[i] : i;
  const uint32_t slot = slotOf[i];                               // k_slot_rank
  const int32_t crow = (int32_t)(k / (uint32_t)g.cols);
  const int32_t sb = slabCuts ? slabCuts[0] : g.slabBegin, se = slabCuts ? slabCuts[1] : g.slabEnd;
  const uint32_t owned = (crow >= sb && crow < se) ? F_OWNED : 0u;
  if (s.SLID) s.SLID[slot] = i;
  uint32_t keep = 0;
  if (afterSpatial) {        // rows of this frame already exist: keep the cap flag, and the
    keep = __float_as_uint(s.SA[2 * (size_t)slot].w) & (F_CAPPED | F_XOVER | F_XPOOL);   // query position stays the pre-move one
    // no k_slot_prep follows on this path: the first sweep's boundary pass happens here
    if (INTEGRATE && moved && !clear_of_walls(g, dp.x, dp.y, at.y)) apply_bounds(g, p.boundaryElasticity, at.y, dp.x, dp.y, dp.z, dp.w);
  }
  s.SA[2 * (size_t)slot] = make_float4(dp.x, dp.y, at.y, __uint_as_float(f | keep | moved | owned | (cc << F_CC_SHIFT)));
  s.SA[2 * (size_t)slot + 1] = make_float4(dp.z, dp.w, at.z, __uint_as_float(gid));
}

// ---- K3c: coalesced slot-order pass: query positions, scan windows, list heads, tile descriptors ----
static constexpr int PREP_THREADS = 256;     // two tiles per block
static_assert(PREP_THREADS % TILE == 0, "a prep block covers whole tiles");

__global__ void __launch_bounds__(PREP_THREADS)
k_slot_prep(GridDims g, const Params* __restrict__ pp, BySlot s, const uint32_t* __restrict__ cellStart) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = e < cellStart[g.cells];
  int4 wi = make_int4(1, 0, 1, 0);
  if (live) {
    float4 lo = s.SA[2 * (size_t)e], hi = s.SA[2 * (size_t)e + 1];
    const bool moved = (__float_as_uint(lo.w) & F_MOVED) != 0;     // integrated: px,py hold the pre-move position
    const float x0 = moved ? hi.x : lo.x, y0 = moved ? hi.y : lo.y;
    s.QXY[e] = make_float2(x0, y0);
    // The boundary pass of the FIRST sweep (physics_worker.js:344-376) is applied here, once the
    // query position is safe in QXY: every sweep then reads positions that already had their
    // pass, and the sweep kernel applies the NEXT sweep's pass to its own result before storing it.
    if (moved && !clear_of_walls(g, lo.x, lo.y, lo.z)) {
      apply_bounds(g, pp->boundaryElasticity, lo.z, lo.x, lo.y, hi.x, hi.y);
      s.SA[2 * (size_t)e] = lo;
      s.SA[2 * (size_t)e + 1] = hi;
    }
    s.GA[e] = lo;                                                  // what the first sweep gathers: x, y, radius, flags (16 B, not the 32 B slot record)
    Window w;
    if (query_window(g, x0, y0, hi.z, w)) wi = make_int4(w.r0, w.r1, w.c0, w.c1);
    s.WIN[e] = wi;
    if (s.PW)      // only the warp-per-entity scan gathers this record
      s.PW[e] = make_uint4(__float_as_uint(hi.z), __float_as_uint(hi.w), (uint32_t)wi.x | ((uint32_t)wi.y << 16),
                           (uint32_t)wi.z | ((uint32_t)wi.w << 16));
    // CX_EDGE clear: my unclamped centre cell is my (clamped) grid cell and trunc == floor.  Two
    // such entities with the same visualRange that pass 0 < d2 < vr^2 always lie in each other's
    // window: |x_a - x_b| < vr  =>  |floor(x_a inv) - floor(x_b inv)| <= ceil(vr inv).
    const int32_t ucol = js_toint32(dmul((double)x0, g.inv)), urow = js_toint32(dmul((double)y0, g.inv));
    const bool inside = x0 >= 0.f && y0 >= 0.f && ucol >= 0 && ucol < g.cols && urow >= 0 && urow < g.rows;
    s.CXY[e] = make_float4(x0, y0, hi.z, __uint_as_float(__float_as_uint(hi.w) | (inside ? 0u : CX_EDGE)));
    s.XHEAD[e] = 0;
  }
  if (!s.TD) return;                                               // uniform: no tiled kernel in this context
  // ---- tile descriptor: union of the tile's windows, one slot range per grid row --------------------
  __shared__ int32_t sWin[PREP_THREADS / 32][4];
  const bool has = wi.x <= wi.y;                                   // an empty window contributes nothing
  const int32_t r0 = __reduce_min_sync(0xffffffffu, has ? wi.x : 0x7fffffff);
  const int32_t r1 = __reduce_max_sync(0xffffffffu, has ? wi.y : -1);
  const int32_t c0 = __reduce_min_sync(0xffffffffu, has ? wi.z : 0x7fffffff);
  const int32_t c1 = __reduce_max_sync(0xffffffffu, has ? wi.w : -1);
  const uint32_t warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sWin[warp][0] = r0; sWin[warp][1] = r1; sWin[warp][2] = c0; sWin[warp][3] = c1; }
  __syncthreads();
  if (threadIdx.x % TILE == 0) {
    int32_t R0 = 0x7fffffff, R1 = -1, C0 = 0x7fffffff, C1 = -1;
    for (uint32_t w = warp; w < warp + TILE / 32; w++) {
      R0 = min(R0, sWin[w][0]); R1 = max(R1, sWin[w][1]); C0 = min(C0, sWin[w][2]); C1 = max(C1, sWin[w][3]);
    }
    TileDesc d;
    d.a[0] = d.a[1] = d.a[2] = 0; d.n[0] = d.n[1] = d.n[2] = 0; d.r0c0 = 0; d.shape = 0;
    if (R1 >= R0 && R1 - R0 < 3) {
      for (int32_t r = 0; r <= R1 - R0; r++) {
        const uint32_t a = cellStart[(uint32_t)(R0 + r) * g.cols + C0];
        d.a[r] = a;
        d.n[r] = cellStart[(uint32_t)(R0 + r) * g.cols + C1 + 1] - a;
      }
      d.r0c0 = (uint32_t)R0 | ((uint32_t)C0 << 16);
      d.shape = (uint32_t)min(C1 - C0 + 1, 0xFFFF) | ((uint32_t)(R1 - R0 + 1) << 16) | TD_OK;
    }
    s.TD[e / TILE] = d;
  }
}

// first sweep input after a separate weed_spatial + weed_physics: the slot records were rewritten
__global__ void __launch_bounds__(256)
k_slots_to_sweep_input(GridDims g, BySlot s, const uint32_t* __restrict__ cellStart) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < cellStart[g.cells]) s.GA[e] = s.SA[2 * (size_t)e];
}

// ---- K4: capped, ordered neighbor gather (spatial_worker.js:195-277) ---------------------
// One thread per entity, in slot order.  Candidates of one window row are ONE contiguous
// slot range because slots are sorted by (cell, id) and the cells of a grid row are
// consecutive; the reference's scan order (rows, then columns, then list order) is therefore
// ascending slot order, so a sequential scan reproduces row content and cap exactly.
//   * phase 1 (scan): fp32 pre-filter — a candidate whose float32 d2 exceeds vr2 by more than
//     1e-5 relative is certainly rejected by the binary64 predicate — then the exact binary64
//     predicate; accepted slots are staged in shared memory (15 per thread and round).
//   * phase 2 (per staged entry, all lanes in step): partner id, "does the partner's scan
//     accept me" bit (its precomputed window + its visualRange), explicit pushes.
//   * flush by the whole warp: API rows (scattered by entity id) are written 16 consecutive
//     words at a time, the internal transposed rows NST[k][slot] 32 consecutive slots at a time.
//
// An explicit pair is a row entry of its owner (the lower id), so the owner's row position
// (k * Npad + slot) is a unique pool index: incoming pairs of a target form a linked list
// threaded through XNEXT, with no capacity limit and no allocation.
// Pads the internal row of slot e (n entries) to a multiple of four with words that have no
// membership bit and point at the entity itself: k_sweep walks rows four entries at a time.
__device__ __forceinline__ void row_tail_fill(const GridDims& g, const BySlot& s, uint32_t e, uint32_t n) {
  for (uint32_t k = n; k < ((n + 3u) & ~3u); k++) s.NST[k * g.Npad + e] = e;
}

// Would the scan of the entity in slot tc accept me (ignoring its cap)?  Its precomputed window must hold
// my clamped cell and its predicate 0 < d2 < vr^2 must pass (d2 is bitwise symmetric; d2 > 0 is known).
__device__ __forceinline__ bool scan_accepts(const GridDims& g, const BySlot& s, uint32_t tc, double myX, double myY,
                                             int32_t myCol, int32_t myRow, uint32_t myVrBits) {
  const float4 c = s.CXY[tc];
  const int4 wt = s.WIN[tc];
  // An empty window is stored as r0 = 1 > r1 = 0, which no cell satisfies.
  bool back = myRow >= wt.x && myRow <= wt.y && myCol >= wt.z && myCol <= wt.w;
  if (back && __float_as_uint(c.z) != myVrBits) {
    const double dX = dsub((double)c.x, myX), dY = dsub((double)c.y, myY);
    const double d2 = dadd(dmul(dX, dX), dmul(dY, dY));
    back = d2 < dmul((double)c.z, (double)c.z);
  }
  return back;
}

// End of an entity's scan: counts (API row | internal row << 16), the slot that closes a capped row,
// flags, padding of the internal row.  A capped row is "everything the scan accepts up to LSLOT", so
// `am I in the row of the capped entity t` is one comparison: my slot <= LSLOT[t].
__device__ __forceinline__ void row_finish(const GridDims& g, const BySlot& s, Counters* ctr, uint32_t e, uint32_t n,
                                           uint32_t lastApi) {
  if (n >= g.M && g.M > 0) {
    reinterpret_cast<uint32_t*>(s.SA + 2 * (size_t)e)[3] |= F_CAPPED;
    reinterpret_cast<uint32_t*>(s.GA + e)[3] |= F_CAPPED;
    ctr->anyCapped = 1;
    s.CAPLIST[atomicAdd(&ctr->nCapped, 1u)] = e;
    s.BCNT[e] = 0;
    s.BCUR[e] = 0;
  }
  s.NCNT[e] = n | (n << 16);          // k_beyond_cap extends the internal part of a capped row
  s.LSLOT[e] = lastApi;
  row_tail_fill(g, s, e, n);
}

__device__ __forceinline__ void explicit_push(const BySlot& s, Counters* ctr, uint32_t dstSlot, uint32_t ownerRowPos) {
  const uint32_t prev = atomicExch(&s.XHEAD[dstSlot], ownerRowPos + 1u);
  s.XNEXT[ownerRowPos] = prev;
  atomicAdd(&ctr->explicitPairs, 1u);
  // exactly one pusher finds the list empty: it enters the slot in the list k_sort_lists walks
  if (prev == 0) s.SORTLIST[atomicAdd(&ctr->nSort, 1u)] = dstSlot;
}

// ---- K4, second form: survivor queue + converged exact pass ------------------------------------
// Same rows, same bits as k_neighbors; the control flow differs.
//   phase 1  streams the 8-byte query positions of the entity's own slot ranges (four in flight)
//            through the float32 pre-filter and only PUSHES the slots that survive onto a per-thread
//            shared-memory queue (a predicated store and an increment: nothing else runs under
//            divergence).  It ends when the queue is full or the window is exhausted.
//   phase 2  every lane drains its queue in step with the others: 16-byte candidate record, the
//            binary64 predicate of spatial_worker.js:252-257, id and float32 d2 straight to the
//            slot-major API planes, the internal row word staged in place (accepted <= drained),
//            cap of :264.
//   then     partners with another visualRange or on the rim resolve their NS_BACK bit, warp flush of
//            the row words.
// Measured at 16M (ms): rows scattered by entity id with a staged 16-lane flush 3.33 (round-1 layout);
// slot-major planes, everything staged 2.19; ids / d2 written directly from phase 2 (8.7 KB of shared
// memory instead of 26 KB) 1.82; row words written directly as well 1.88; queue of 12 / 8 instead of
// 16 entries 2.24 / 2.54 (more rounds).
#ifndef WEED_K4V2_Q
#define WEED_K4V2_Q 16
#endif
#ifndef WEED_K4V2_MINBLOCKS
#define WEED_K4V2_MINBLOCKS 8
#endif
static constexpr int K4V2_THREADS = 128;
static constexpr int K4V2_Q = WEED_K4V2_Q;                    // queue = stage entries per thread and round (at most 16)
static constexpr int K4V2_STRIDE = K4V2_Q | 1;                // odd stride: conflict-free smem both ways
static_assert(K4V2_Q >= 4 && K4V2_Q <= 16, "the flush writes at most 16 words of a row per round");

template <bool WRITE_ROWS>
__global__ void __launch_bounds__(K4V2_THREADS, WEED_K4V2_MINBLOCKS)
k_neighbors2(GridDims g, BySlot s, const uint32_t* __restrict__ cellStart, int32_t* __restrict__ nd,
             float* __restrict__ dd, Counters* ctr) {
  constexpr uint32_t PLANE = (K4V2_THREADS / 32) * 32 * K4V2_STRIDE;
  __shared__ uint32_t sStage[PLANE];           // survivor queue, then the accepted row words of the round
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* const myW = &sStage[warp * 32 * K4V2_STRIDE + lane * K4V2_STRIDE];
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t A = cellStart[g.cells];
  const uint32_t M = g.M;
  const bool live = e < A;
  float2 q = make_float2(0.f, 0.f);
  float vr = 0.f;
  uint32_t id = 0, edge = CX_EDGE;
  int4 win = make_int4(1, 0, 1, 0);
  if (live) {
    const float4 me = s.CXY[e];
    q = make_float2(me.x, me.y);
    vr = me.z; id = __float_as_uint(me.w) & ~CX_EDGE; edge = __float_as_uint(me.w) & CX_EDGE;
    win = s.WIN[e];
  }
  bool done = !(live && M > 0 && win.x <= win.y);
  const double myX = q.x, myY = q.y;
  const double vrSq = dmul((double)vr, (double)vr);
  const float vrSqF = vr * vr * 1.00001f;            // pre-filter threshold (NaN/Inf compare false)
  const uint32_t vrBits = __float_as_uint(vr);
  int32_t myCol = 0, myRow = 0;
  if (live) cell_of(g, q.x, q.y, myCol, myRow);      // my clamped cell (for partners' windows)
  const float2* __restrict__ QXY = s.QXY;
  uint32_t n = 0;
  uint32_t lastApi = SLOT_NONE;                      // slot of the M-th entry once the row is capped
  int32_t row = win.x;
  uint32_t t = 0, b = 0;
  if (!done) {
    t = cellStart[(uint32_t)row * g.cols + win.z];
    b = cellStart[(uint32_t)row * g.cols + win.w + 1];
  }
  do {
    // ---- phase 1: survivors of the float32 pre-filter -------------------------------------------
    uint32_t* wp = myW;                                     // next free queue entry
    uint32_t* const wfull = myW + (K4V2_Q - 3);             // room for four more below this
    while (!done && wp < wfull) {
      if (t >= b) {
        if (++row > win.y) { done = true; break; }
        t = cellStart[(uint32_t)row * g.cols + win.z];
        b = cellStart[(uint32_t)row * g.cols + win.w + 1];
        continue;
      }
      // four positions off one pointer; reads past the end of the range stay inside the (padded)
      // array and are masked by the range test
      const float2* cp = QXY + t;
      float2 c[4];
#pragma unroll
      for (int u = 0; u < 4; u++) c[u] = __ldg(cp + u);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const float fx = c[u].x - q.x, fy = c[u].y - q.y;
        const bool pass = (t + (uint32_t)u < b) && !(__fmaf_rn(fx, fx, fy * fy) > vrSqF);
        if (pass) *wp = t + (uint32_t)u;
        wp += pass ? 1 : 0;
      }
      t += 4;
    }
    const uint32_t qn = (uint32_t)(wp - myW);
    // ---- phase 2: exact predicate, staged in place -------------------------------------------------
    uint32_t cnt = 0;
    bool anySlow = false;
    uint32_t slowMask = 0;                               // staged entries whose NS_BACK bit is still open
    for (uint32_t k = 0; k < qn; k++) {
      const uint32_t tc = myW[k];
      const float4 c = __ldg(s.CXY + tc);
      const double dX = dsub((double)c.x, myX);           // :252-254
      const double dY = dsub((double)c.y, myY);
      const double d2 = dadd(dmul(dX, dX), dmul(dY, dY));
      if (!(d2 < vrSq && d2 > 0)) continue;               // :257 (d2 > 0 also skips myself, :249)
      const uint32_t jw = __float_as_uint(c.w);
      const uint32_t jid = jw & ~CX_EDGE;
      const bool sure = __float_as_uint(c.z) == vrBits && !((jw | edge) & CX_EDGE);
      // the API row entry goes straight to its slot-major planes (lanes of a warp sit at similar row
      // positions, so these stores touch a few lines each); the internal row word is staged, because its
      // NS_BACK bit may still be open, and leaves with the warp flush below
      if (WRITE_ROWS) {
        const uint32_t ix = (n + cnt) * g.Npad + e;
        __stcs(nd + ix, (int32_t)jid);                     // :259
        __stcs(dd + ix, fround(d2));                       // :260
      }
      myW[cnt] = tc | (jid > id ? NS_OUT : 0u) | (sure ? NS_BACK : 0u);
      if (!sure) slowMask |= 1u << cnt;
      cnt++;
      anySlow |= !sure;
      if (n + cnt >= M) { lastApi = tc; done = true; break; }   // :264 — the row is full: everything my scan accepts up to this slot is in it
    }
    const uint32_t first = n;                             // row position of my first staged entry
    n += cnt;
    // ---- staged entries whose partner differs in visualRange or sits on the rim ---------------------
    if (anySlow) {
      for (uint32_t k = 0; k < cnt; k++) {
        if (!(slowMask >> k & 1u)) continue;
        const uint32_t wd = myW[k];
        const uint32_t tc = wd & NS_SLOT_MASK;
        if (scan_accepts(g, s, tc, myX, myY, myCol, myRow, vrBits)) myW[k] = wd | NS_BACK;
        // pair (id, jid) is in P but the partner cannot infer it from its own row
        else if (wd & NS_OUT) explicit_push(s, ctr, tc, (first + k) * g.Npad + e);
      }
    }
    __syncwarp();
    // ---- warp flush of the internal row words: lane after lane at the same k -> full lines --------------
    const uint32_t kmax = __reduce_max_sync(0xffffffffu, cnt);
    for (uint32_t k = 0; k < kmax; k++)
      if (k < cnt) s.NST[(first + k) * g.Npad + e] = myW[k];
    __syncwarp();
  } while (__any_sync(0xffffffffu, !done));
  if (!live) return;
  row_finish(g, s, ctr, e, n, lastApi);
}

// ---- K4 (wide variant): one WARP per entity -------------------------------------------------------
// For scenes with long rows (maxNeighbors in the hundreds, windows of 25+ cells: the reference's
// own demos) a thread per entity leaves the GPU mostly idle and serialises hundreds of candidates.
// Here the 32 lanes test 32 consecutive candidates at a time; an ordered ballot compaction keeps
// the reference's scan order and cap, and every accepting lane writes its own row words (they are
// consecutive within the row, so the stores coalesce).
template <bool WRITE_ROWS>
__global__ void __launch_bounds__(256)
k_neighbors_wide(GridDims g, BySlot s, const uint32_t* __restrict__ cellStart, int32_t* __restrict__ nd,
                 float* __restrict__ dd, Counters* ctr) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= cellStart[g.cells]) return;
  const uint32_t M = g.M;
  const float2 q = s.QXY[e];
  const float4 hi = s.SA[2 * (size_t)e + 1];
  const float vr = hi.z;
  const uint32_t id = __float_as_uint(hi.w);
  const int4 win = s.WIN[e];
  const double myX = q.x, myY = q.y;
  const double vrSq = dmul((double)vr, (double)vr);
  const float vrSqF = vr * vr * 1.00001f;
  int32_t myCol, myRow;
  cell_of(g, q.x, q.y, myCol, myRow);
  uint32_t n = 0;
  uint32_t lastApi = SLOT_NONE;
  for (int32_t row = win.x; row <= win.y && n < M && M > 0; row++) {
    const uint32_t a = cellStart[(uint32_t)row * g.cols + win.z];
    const uint32_t b = cellStart[(uint32_t)row * g.cols + win.w + 1];
    for (uint32_t t0 = a; t0 < b && n < M; t0 += 32) {
      const uint32_t t = t0 + lane;
      bool acc = false;
      double d2 = 0;
      if (t < b) {
        const float2 c = s.QXY[t];
        const float fx = c.x - q.x, fy = c.y - q.y;
        if (!(__fmaf_rn(fx, fx, fy * fy) > vrSqF)) {
          const double dX = dsub((double)c.x, myX), dY = dsub((double)c.y, myY);   // :252-254
          d2 = dadd(dmul(dX, dX), dmul(dY, dY));
          acc = d2 < vrSq && d2 > 0;                                              // :257, :249
        }
      }
      const uint32_t bits = __ballot_sync(0xffffffffu, acc);
      const uint32_t pos = n + __popc(bits & ((1u << lane) - 1));
      const bool take = acc && pos < M;
      if (take) {
        const uint4 pw = s.PW[t];
        const float vrT = __uint_as_float(pw.x);
        bool back = (uint32_t)myRow >= (pw.z & 0xFFFFu) && (uint32_t)myRow <= (pw.z >> 16) &&
                    (uint32_t)myCol >= (pw.w & 0xFFFFu) && (uint32_t)myCol <= (pw.w >> 16);
        if (back && vrT != vr) back = d2 < dmul((double)vrT, (double)vrT);
        const bool out = pw.y > id;
        s.NST[(size_t)pos * g.Npad + e] = t | (out ? NS_OUT : 0u) | (back ? NS_BACK : 0u);
        if (WRITE_ROWS) {
          nd[(size_t)pos * g.Npad + e] = (int32_t)pw.y;  // :259
          dd[(size_t)pos * g.Npad + e] = fround(d2);     // :260
        }
        if (out && !back) explicit_push(s, ctr, t, pos * g.Npad + e);
      }
      const uint32_t taken = __ballot_sync(0xffffffffu, take);
      n += (uint32_t)__popc(taken);                     // :264
      if (n >= M) lastApi = t0 + (uint32_t)(31 - __clz((int)taken));   // the slot that closes the row
    }
  }
  if (lane == 0) row_finish(g, s, ctr, e, n, lastApi);
}

static constexpr uint32_t XPOOL_ROW = 512;   // entries of one overflow-pool row

// ---- K4b: lower-id partners past the cap ---------------------------------------------------------------
// A capped row may have lost LOWER-id partners that do list this entity in THEIR rows: those pairs are in
// P, and the entity has to apply them.  Once every row and every LSLOT is known, each capped entity
// resumes its scan after the slot that closed its row and appends what qualifies — lower id, mutual
// acceptance, and my slot inside the partner's row: not capped, or my slot <= its LSLOT — to its
// INTERNAL row (positions maxNeighbors .. Mint - 1), in scan order, membership already decided.  The
// few entities a whole pile lists early in its scans ("popular" ones: hundreds of such partners) continue
// in a row of the overflow pool (F_XPOOL: XPOOL_ROW more entries, drawn with one atomic); past that, or
// with the pool exhausted, the entity is marked F_XOVER and the sweeps resume the scan themselves.
// (Round 1 did this with one warp per capped entity and a binary search of the partner's row per
// candidate: 0.3 ms for 55 k capped rows, 149 ms for the 4 M capped rows of a settled bed.)
// Two forms, chosen on the device by how many rows are capped.  Few (the piles on the walls of the opening
// frames): one WARP per capped entity from the list row_finish built, 32 candidates per step, ordered
// ballot compaction — the handful of entities with hundreds of candidates left do not become a tail.
// Many (a settled bed: most rows capped, some dozens of candidates left each): one THREAD per entity,
// which keeps every lane busy.  Measured at 16M: 55 k capped rows 0.33 (thread) / 0.14 ms (warp);
// 3.9 M capped rows 17.7 (thread) / 27.9 ms (warp).
__device__ __forceinline__ bool beyond_dense_regime(uint32_t nCapped, uint32_t A) { return (unsigned long long)nCapped * 8ull > A; }
__device__ __forceinline__ void back_count_rows(const GridDims& g, const BySlot& s, uint32_t A);
static constexpr int K4B_BLOCKS = 148 * 8;
__global__ void __launch_bounds__(256)
k_beyond_cap(GridDims g, BySlot s, const uint32_t* __restrict__ cellStart, Counters* ctr) {
  const uint32_t nCapped = ctr->nCapped;
  if (beyond_dense_regime(nCapped, cellStart[g.cells])) {            // a settled bed: reverse edges (below), this launch counts them
    back_count_rows(g, s, cellStart[g.cells]);
    return;
  }
  const uint32_t lane = threadIdx.x & 31, below = (1u << lane) - 1u;
  const uint32_t warpsTotal = gridDim.x * (blockDim.x >> 5);
  for (uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < nCapped; w += warpsTotal) {
    const uint32_t e = s.CAPLIST[w];
    const uint32_t last = s.LSLOT[e];
    const float4 me = s.CXY[e];
    const uint32_t id = __float_as_uint(me.w) & ~CX_EDGE;
    const uint32_t vrBits = __float_as_uint(me.z);
    const double myX = me.x, myY = me.y, vrSq = dmul((double)me.z, (double)me.z);
    const float vrSqF = me.z * me.z * 1.00001f;
    int32_t myCol, myRow;
    cell_of(g, me.x, me.y, myCol, myRow);
    const int4 win = s.WIN[e];
    uint32_t n = g.M;                                   // entries of the internal row + the pool row
    uint32_t pid = SLOT_NONE;
    const uint32_t room = g.Mint + XPOOL_ROW;
    bool xover = false;
    // One warp, one chain of dependent loads: keep it short.  The window rows' slot ranges are fetched by 32 lanes
    // at once, the next 32 candidate records are in flight while the current ones are tested, and a candidate in
    // range fetches its LSLOT and WIN together.
    for (int32_t rowBase = win.x; rowBase <= win.y && !xover; rowBase += 32) {
      const int32_t myR = rowBase + (int32_t)lane;
      uint32_t ra = 0, rb = 0;
      if (myR <= win.y) {
        ra = cellStart[(uint32_t)myR * g.cols + win.z];
        rb = cellStart[(uint32_t)myR * g.cols + win.w + 1];
      }
      const int32_t nr = min(32, win.y - rowBase + 1);
      for (int32_t r = 0; r < nr && !xover; r++) {
        const uint32_t a = max(__shfl_sync(0xffffffffu, ra, r), last + 1u);
        const uint32_t b = __shfl_sync(0xffffffffu, rb, r);
        float4 cNext = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a + lane < b) cNext = __ldg(s.CXY + a + lane);
        for (uint32_t t0 = a; t0 < b && !xover; t0 += 32) {
          const uint32_t tc = t0 + lane;
          const float4 c = cNext;
          if (tc + 32 < b) cNext = __ldg(s.CXY + tc + 32);
          bool ok = false;
          if (tc < b) {
            const float fx = c.x - me.x, fy = c.y - me.y;
            // lower ids only (higher ids past my cap are my own pairs, lost as in the reference)
            if ((__float_as_uint(c.w) & ~CX_EDGE) < id && !(__fmaf_rn(fx, fx, fy * fy) > vrSqF)) {
              const uint32_t lk = s.LSLOT[tc];
              const int4 wt = s.WIN[tc];
              const double dX = dsub((double)c.x, myX), dY = dsub((double)c.y, myY);
              const double d2 = dadd(dmul(dX, dX), dmul(dY, dY));
              // in my range; it is not capped, or its row closed after it reached me — and its scan accepts me
              if (d2 < vrSq && d2 > 0 && (lk == SLOT_NONE || e <= lk)) {
                ok = myRow >= wt.x && myRow <= wt.y && myCol >= wt.z && myCol <= wt.w;
                if (ok && __float_as_uint(c.z) != vrBits) ok = d2 < dmul((double)c.z, (double)c.z);   // scan_accepts
              }
            }
          }
          const uint32_t bits = __ballot_sync(0xffffffffu, ok);
          if (!bits) continue;
          const uint32_t cntNew = (uint32_t)__popc(bits);
          if (n + cntNew > g.Mint && pid == SLOT_NONE) {   // the internal row is full: continue in the overflow pool
            if (lane == 0) pid = atomicAdd(&ctr->xpoolUsed, 1u);
            pid = __shfl_sync(0xffffffffu, pid, 0);
            if (pid >= g.xpoolRows) { pid = SLOT_NONE; xover = true; }
          }
          const uint32_t pos = n + (uint32_t)__popc(bits & below);
          if (ok) {
            if (pos < g.Mint) s.NST[pos * g.Npad + e] = tc | NS_BACK;
            else if (pid != SLOT_NONE && pos < room) s.XR[(size_t)pid * XPOOL_ROW + (pos - g.Mint)] = tc | NS_BACK;
          }
          n += cntNew;
          if (n > room || (n > g.Mint && pid == SLOT_NONE)) xover = true;
        }
      }
    }
    if (lane == 0) {
      const uint32_t inRow = min(n, g.Mint);
      uint32_t add = xover ? F_XOVER : 0u;
      if (pid != SLOT_NONE) {
        const uint32_t m = min(n, room) - g.Mint;
        add |= F_XPOOL;
        s.XPID[e] = pid;
        s.XRCNT[pid] = m;
        for (uint32_t k = m; k < ((m + 3u) & ~3u); k++) s.XR[(size_t)pid * XPOOL_ROW + k] = e;   // padding: no membership bit
      }
      if (add) {
        reinterpret_cast<uint32_t*>(s.SA + 2 * (size_t)e)[3] |= add;
        reinterpret_cast<uint32_t*>(s.GA + e)[3] |= add;
        s.HEAVY[atomicAdd(&ctr->nHeavy, 1u)] = e;
        if (xover) atomicAdd(&ctr->xoverRows, 1u);
      }
      if (inRow > g.M) {
        s.NCNT[e] = g.M | (inRow << 16);
        row_tail_fill(g, s, e, inRow);
      }
    }
  }
}

// ---- the dense regime (a settled bed: most rows capped): reverse edges instead of a search ---------------------
// The search above visits every candidate left in the window of every capped entity: 1e10 candidates for the
// 4 M capped rows of config 4 after 300 frames (cells of 500-1900 entities), 16 ms however the loop is
// arranged (per-cell skips, LSLOT tested first, heavy slots first: all measured, profiles/README.md).  But the
// pairs it looks for are already written down, from the other side: t is a lost lower-id partner of e exactly
// when e sits in t's API row with both membership bits (NS_OUT: t has the lower id; NS_BACK: e's scan accepts
// t) and t lies past the slot that closed e's row.  So every entity walks its own row once — 4e8 words instead
// of 1e10 candidates — and reports itself to those partners:
//   k_beyond_cap (dense branch)  counts the reports per capped entity            (BCNT)
//   k_back_alloc                 sizes the internal row, draws a pool row, or marks F_XOVER (nothing stored:
//                                the sweeps resume the scan from LSLOT, as they do for the search's overflow)
//   k_back_write                 stores the reports through a cursor                (BCUR), in arrival order
//   k_back_sort                  one warp per capped entity sorts them ascending: the order of the scan
// Same stored sets, same order, hence the same sweep results as the search, bit for bit.
__device__ __forceinline__ void back_count_rows(const GridDims& g, const BySlot& s, uint32_t A) {
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < A; t += gridDim.x * blockDim.x) {
    const uint32_t n = s.NCNT[t] & 0xFFFFu;
    for (uint32_t k = 0; k < n; k += 4) {                // four row words, then their partners' LSLOT, in flight together
      uint32_t wd[4], lk[4];
#pragma unroll
      for (uint32_t u = 0; u < 4; u++) wd[u] = k + u < n ? s.NST[(k + u) * g.Npad + t] : 0u;
#pragma unroll
      for (uint32_t u = 0; u < 4; u++)                   // NS_OUT and NS_BACK; SLOT_NONE (row not capped) is never below t
        lk[u] = (wd[u] >> 30) == 3u ? __ldg(s.LSLOT + (wd[u] & NS_SLOT_MASK)) : SLOT_NONE;
#pragma unroll
      for (uint32_t u = 0; u < 4; u++)
        if (lk[u] < t) atomicAdd(&s.BCNT[wd[u] & NS_SLOT_MASK], 1u);
    }
  }
}

static constexpr uint32_t BACK_XOVER = 0x80000000u;   // BCNT: more reports than the internal row and a pool row hold
__global__ void __launch_bounds__(256)
k_back_alloc(GridDims g, BySlot s, const uint32_t* __restrict__ cellStart, Counters* ctr) {
  const uint32_t nCapped = ctr->nCapped;
  if (!beyond_dense_regime(nCapped, cellStart[g.cells])) return;
  const uint32_t room = g.Mint - g.M;
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < nCapped; w += gridDim.x * blockDim.x) {
    const uint32_t e = s.CAPLIST[w];
    const uint32_t c = s.BCNT[e];
    if (c == 0) continue;                                // NCNT and the padding are row_finish's
    uint32_t add = 0, inRow = g.M + min(c, room);
    if (c > room) {
      uint32_t pid = SLOT_NONE;
      if (c - room <= XPOOL_ROW) {
        pid = atomicAdd(&ctr->xpoolUsed, 1u);
        if (pid >= g.xpoolRows) pid = SLOT_NONE;
      }
      if (pid != SLOT_NONE) {
        const uint32_t m = c - room;
        add = F_XPOOL;
        s.XPID[e] = pid;
        s.XRCNT[pid] = m;
        for (uint32_t k = m; k < ((m + 3u) & ~3u); k++) s.XR[(size_t)pid * XPOOL_ROW + k] = e;   // padding: no membership bit
      } else {
        add = F_XOVER;
        inRow = g.M;
        s.BCNT[e] = c | BACK_XOVER;
        atomicAdd(&ctr->xoverRows, 1u);
      }
      reinterpret_cast<uint32_t*>(s.SA + 2 * (size_t)e)[3] |= add;
      reinterpret_cast<uint32_t*>(s.GA + e)[3] |= add;
      s.HEAVY[atomicAdd(&ctr->nHeavy, 1u)] = e;
    }
    if (inRow > g.M) {
      s.NCNT[e] = g.M | (inRow << 16);
      row_tail_fill(g, s, e, inRow);
    }
  }
}

__global__ void __launch_bounds__(256)
k_back_write(GridDims g, BySlot s, const uint32_t* __restrict__ cellStart, const Counters* __restrict__ ctr) {
  const uint32_t A = cellStart[g.cells];
  if (!beyond_dense_regime(ctr->nCapped, A)) return;
  const uint32_t room = g.Mint - g.M;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < A; t += gridDim.x * blockDim.x) {
    const uint32_t n = s.NCNT[t] & 0xFFFFu;
    for (uint32_t k = 0; k < n; k += 4) {
      uint32_t wd[4], lk[4];
#pragma unroll
      for (uint32_t u = 0; u < 4; u++) wd[u] = k + u < n ? s.NST[(k + u) * g.Npad + t] : 0u;
#pragma unroll
      for (uint32_t u = 0; u < 4; u++)
        lk[u] = (wd[u] >> 30) == 3u ? __ldg(s.LSLOT + (wd[u] & NS_SLOT_MASK)) : SLOT_NONE;
#pragma unroll
      for (uint32_t u = 0; u < 4; u++) {
        if (!(lk[u] < t)) continue;
        const uint32_t e = wd[u] & NS_SLOT_MASK;
        if (s.BCNT[e] & BACK_XOVER) continue;            // nothing is stored for it
        const uint32_t pos = atomicAdd(&s.BCUR[e], 1u);
        if (pos < room) s.NST[(g.M + pos) * g.Npad + e] = t | NS_BACK;
        else s.XR[(size_t)s.XPID[e] * XPOOL_ROW + (pos - room)] = t | NS_BACK;
      }
    }
  }
}

static constexpr int BSORT_WARPS = 4, BSORT_CAP = 1024;      // room (at most 128 + the padding of M to a multiple of 4) + XPOOL_ROW entries at most
static_assert(BSORT_CAP >= 132 + (int)XPOOL_ROW, "k_back_sort stages a whole list (internal-row extension + pool row) in shared memory");
// Latency-bound (gather the entries, sort, scatter them back: three dependent round trips per entity): as many
// resident warps as the registers allow, one wave.  A thread-local network for the short lists was slower
// (2.8 vs 2.0 ms at 16 M, frame 300): the lists of a pile are not short.
static constexpr int BSORT_BLOCKS_PER_SM = 12;
__global__ void __launch_bounds__(BSORT_WARPS * 32, BSORT_BLOCKS_PER_SM)
k_back_sort(GridDims g, BySlot s, const uint32_t* __restrict__ cellStart, const Counters* __restrict__ ctr) {
  __shared__ uint32_t sm[BSORT_WARPS][BSORT_CAP];
  const uint32_t nCapped = ctr->nCapped;
  if (!beyond_dense_regime(nCapped, cellStart[g.cells])) return;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t room = g.Mint - g.M;
  const uint32_t warpsTotal = gridDim.x * BSORT_WARPS;
  // one list per warp and step (32 list heads per step, then their sorts one after the other, was no faster at
  // 16 M and slower at 1 M: the sorts, not the heads, are the work)
  for (uint32_t w = blockIdx.x * BSORT_WARPS + warp; w < nCapped; w += warpsTotal) {
    const uint32_t e = s.CAPLIST[w];
    const uint32_t c = s.BCNT[e];
    if (c < 2 || (c & BACK_XOVER)) continue;
    const uint32_t* pool = c > room ? s.XR + (size_t)s.XPID[e] * XPOOL_ROW : nullptr;
    auto get = [&](uint32_t k) { return k < room ? s.NST[(g.M + k) * g.Npad + e] : pool[k - room]; };
    auto put = [&](uint32_t k, uint32_t v) { if (k < room) s.NST[(g.M + k) * g.Npad + e] = v; else const_cast<uint32_t*>(pool)[k - room] = v; };
    if (c <= 32) {                                       // one entry per lane, bitonic network over the warp
      uint32_t v = lane < c ? get(lane) : 0xFFFFFFFFu;
      for (uint32_t k = 2; k <= 32; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          const uint32_t o = __shfl_xor_sync(0xffffffffu, v, j);
          const bool keepMin = ((lane & k) == 0) == ((lane & j) == 0);
          v = keepMin ? min(v, o) : max(v, o);
        }
      if (lane < c) put(lane, v);
    } else {
      uint32_t P = 64;
      while (P < c) P <<= 1;
      uint32_t* a = sm[warp];
      for (uint32_t k = lane; k < P; k += 32) a[k] = k < c ? get(k) : 0xFFFFFFFFu;
      __syncwarp();
      for (uint32_t k = 2; k <= P; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          for (uint32_t idx = lane; idx < P; idx += 32) {
            const uint32_t ixj = idx ^ j;
            if (ixj > idx) {
              const uint32_t x0 = a[idx], x1 = a[ixj];
              if ((x0 > x1) == ((idx & k) == 0)) { a[idx] = x1; a[ixj] = x0; }
            }
          }
          __syncwarp();
        }
      for (uint32_t k = lane; k < c; k += 32) put(k, a[k]);
      __syncwarp();
    }
  }
}

// ---- K4c: put every explicit list in ascending source-slot order (adaptive insertion) ---------
__global__ void __launch_bounds__(256)
k_sort_lists(GridDims g, BySlot s, const uint32_t* __restrict__ cellStart, const Counters* __restrict__ ctr) {
  const uint32_t nSort = ctr->nSort;                 // slots that received explicit pairs this frame
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < nSort; w += gridDim.x * blockDim.x) {
  const uint32_t e = s.SORTLIST[w];
  uint32_t p = s.XHEAD[e];
  if (p == 0 || s.XNEXT[p - 1] == 0) continue;
  uint32_t head = 0, tail = 0, tailKey = 0;
  while (p != 0) {
    const uint32_t nxt = s.XNEXT[p - 1];
    const uint32_t key = (p - 1) % g.Npad;
    if (head == 0) { head = tail = p; tailKey = key; s.XNEXT[p - 1] = 0; }
    else if (key >= tailKey) { s.XNEXT[tail - 1] = p; s.XNEXT[p - 1] = 0; tail = p; tailKey = key; }
    else if (key < (head - 1) % g.Npad) { s.XNEXT[p - 1] = head; head = p; }
    else {
      uint32_t qn = head;
      while (true) {
        const uint32_t qx = s.XNEXT[qn - 1];
        if (qx == 0 || (qx - 1) % g.Npad > key) break;
        qn = qx;
      }
      s.XNEXT[p - 1] = s.XNEXT[qn - 1];
      s.XNEXT[qn - 1] = p;
    }
    p = nxt;
  }
  s.XHEAD[e] = head;
  }
}

// ---- K6: one constraint substep (physics_worker.js:323-395, 405-568), J-order ---------------
// Stored positions already had this sweep's boundary pass (k_slot_prep applies the first one,
// each sweep applies the next one to its own result).  Each entity evaluates every pair of P it
// belongs to on the start-of-sweep positions and accumulates its own corrections in ascending partner-slot order, rounding to float32 after
// each one exactly like the reference's `x[i] += ...` on a Float32Array.
//
// Pair membership (P = {(i,j): i<j, j in row(i), both active colliders}), seen from entity e:
//   - row entry with NS_OUT (partner id higher, always inside my API row): the pair is mine.
//   - partner k with lower id: the pair exists iff I am in row(k).
//       k sees me, I do not see k      -> k pushed the pair on my explicit list in K4
//       mutual                         -> k is in my INTERNAL row (k_beyond_cap appends the lower ids my
//                                         capped row lost) with NS_BACK set; a row is the first
//                                         maxNeighbors candidates its scan accepts in ascending slot
//                                         order, so if k's row is capped I am in it iff my slot <= LSLOT[k]
//       my internal row overflowed too -> F_XOVER: the sweep resumes the scan after its last entry
//
// Two phases per entity: phase 1 walks the row with a float32 pre-filter of the overlap test
// (physics_worker.js:455) and stages the few partners that may overlap; phase 2 runs the
// exact binary64 pair code on the staged ones, in order.
struct SubstepAcc { float x, y; uint32_t hits, outHits; };

// partner position: stored positions already had the boundary pass of the sweep that reads them
__device__ __forceinline__ void partner_pos(const GridDims&, float4 gt, float& xt, float& yt) {
  xt = gt.x; yt = gt.y;
}

// float32 pre-filter of :455 (dist2 >= minDist^2): true = certainly no overlap
__device__ __forceinline__ bool surely_apart(float x, float y, float r, float xt, float yt, float rt) {
  const float fx = x - xt, fy = y - yt, md = r + rt;
  return __fmaf_rn(fx, fx, fy * fy) > md * md * 1.00001f;
}

__device__ __forceinline__ void exact_pair(const Params& p, const BySlot& s, uint32_t frame, uint32_t substep,
                                           uint32_t e, float x, float y, float r, uint32_t fw, uint32_t t,
                                           float xt, float yt, float rt, uint32_t ft, bool lower, SubstepAcc& acc) {
  PairMove m;
  if (lower) m = pair_eval(p, frame, substep, s.SA, e, t, x, y, r, fw, xt, yt, rt, ft);
  else       m = pair_eval(p, frame, substep, s.SA, t, e, xt, yt, rt, ft, x, y, r, fw);
  if (!m.hit) return;
  acc.hits++;
  if (lower) acc.outHits++;
  if (lower ? m.moveI : m.moveJ) {
    acc.x = fround(dadd((double)acc.x, lower ? m.mx : -m.mx));
    acc.y = fround(dadd((double)acc.y, lower ? m.my : -m.my));
  }
}

// am I (slot e) in the row of the entity in slot t, whose scan accepts me?  (ft: its flag word)
__device__ __forceinline__ bool in_row_of(const BySlot& s, uint32_t ft, uint32_t t, uint32_t e) {
  return !(ft & F_CAPPED) || e <= s.LSLOT[t];
}

// The lower-id partners past the end of a full internal row (F_XOVER): K4's scan, resumed after the
// last stored slot.  next() returns their slots in ascending order, SLOT_NONE at the end.
struct BeyondScan {
  uint32_t id, vrBits, after, t, b;
  double myX, myY, vrSq;
  int32_t myCol, myRow, row;
  int4 win;
  __device__ void start(const GridDims& g, const BySlot& s, const uint32_t* __restrict__ cellStart, uint32_t e, uint32_t lastStored) {
    const float4 me = s.CXY[e];
    myX = me.x; myY = me.y; vrBits = __float_as_uint(me.z);
    vrSq = dmul((double)me.z, (double)me.z);
    id = __float_as_uint(me.w) & ~CX_EDGE;
    cell_of(g, me.x, me.y, myCol, myRow);
    win = s.WIN[e];
    after = lastStored;
    row = win.x - 1; t = 0; b = 0;
  }
  __device__ uint32_t next(const GridDims& g, const BySlot& s, const uint32_t* __restrict__ cellStart) {
    while (true) {
      if (t >= b) {
        if (++row > win.y) return SLOT_NONE;
        t = max(cellStart[(uint32_t)row * g.cols + win.z], after + 1u);
        b = cellStart[(uint32_t)row * g.cols + win.w + 1];
        continue;
      }
      const uint32_t tc = t++;
      const float4 c = s.CXY[tc];
      if ((__float_as_uint(c.w) & ~CX_EDGE) >= id) continue;            // higher ids: my own pairs, lost to the cap
      const double dX = dsub((double)c.x, myX), dY = dsub((double)c.y, myY);
      const double d2 = dadd(dmul(dX, dX), dmul(dY, dY));
      if (!(d2 < vrSq && d2 > 0)) continue;
      if (!scan_accepts(g, s, tc, myX, myY, myCol, myRow, vrBits)) continue;
      return tc;
    }
  }
};

// slow path: entities with explicit incoming pairs (sorted by k_sort_lists).  A linear merge of ascending
// streams: the row entries, then the overflow-pool row, then the resumed scan — and the explicit sources.
__device__ __noinline__ void substep_slow(const FrameConst* __restrict__ fc, const Params* __restrict__ pp,
                                          const float4* __restrict__ Gin, const uint32_t* __restrict__ cellStart,
                                          uint32_t frame, uint32_t substep, uint32_t e, float x, float y, float r,
                                          uint32_t fw, uint32_t cnt, uint32_t head, SubstepAcc& acc) {
  const GridDims& g = fc->g;
  const BySlot& s = fc->s;
  const Params& p = *pp;
  const bool xover = (fw & F_XOVER) != 0;
  const uint32_t* pool = nullptr;
  uint32_t pcnt = 0;
  if (fw & F_XPOOL) { const uint32_t pid = s.XPID[e]; pool = s.XR + (size_t)pid * XPOOL_ROW; pcnt = s.XRCNT[pid]; }
  uint32_t lastStored = 0;
  if (pcnt) lastStored = pool[pcnt - 1] & NS_SLOT_MASK;
  else if (cnt) lastStored = s.NST[(size_t)(cnt - 1) * g.Npad + e] & NS_SLOT_MASK;
  BeyondScan bs;
  if (xover) bs.start(g, s, cellStart, e, lastStored);
  uint32_t a = 0, pa = 0, pl = head;
  uint32_t wa = 0, ta = SLOT_NONE;           // head of the row / pool / resumed-scan stream
  auto advance = [&]() {
    if (a < cnt) { wa = s.NST[(size_t)a * g.Npad + e]; ta = wa & NS_SLOT_MASK; a++; }
    else if (pa < pcnt) { wa = pool[pa++]; ta = wa & NS_SLOT_MASK; }
    else if (xover) { ta = bs.next(g, s, cellStart); wa = ta | NS_BACK; }
    else ta = SLOT_NONE;
  };
  advance();
  while (ta != SLOT_NONE || pl != 0) {
    const uint32_t tb = pl != 0 ? (pl - 1) % g.Npad : SLOT_NONE;
    uint32_t t; bool lower, viaList;
    if (tb <= ta) {               // explicit incoming: partner is i, I am j
      pl = s.XNEXT[pl - 1];
      t = tb; lower = false; viaList = true;
      if (ta == tb) advance();
    } else {
      t = ta; lower = (wa & NS_OUT) != 0; viaList = false;
      const bool back = (wa & NS_BACK) != 0;
      advance();
      if (!lower && !back) continue;
    }
    const float4 gt = Gin[t];
    const uint32_t ft = __float_as_uint(gt.w);
    if ((ft & F_COLLIDER) != F_COLLIDER) continue;                 // :441
    if (!lower && !viaList && !in_row_of(s, ft, t, e)) continue;
    float xt, yt;
    partner_pos(g, gt, xt, yt);
    if (surely_apart(x, y, r, xt, yt, gt.z)) continue;
    exact_pair(p, s, frame, substep, e, x, y, r, fw, t, xt, yt, gt.z, ft, lower, acc);
  }
}

// F_XOVER without an explicit list: the partners past everything that was stored, straight from the scan
__device__ __noinline__ void sweep_resumed_scan(const FrameConst* __restrict__ fc, const Params* __restrict__ pp,
                                                const float4* __restrict__ Gin, const uint32_t* __restrict__ cellStart,
                                                uint32_t frame, uint32_t substep, uint32_t e, float x, float y, float r,
                                                uint32_t fw, uint32_t lastStored, SubstepAcc& acc) {
  const GridDims& g = fc->g;
  const BySlot& s = fc->s;
  const Params& p = *pp;
  BeyondScan bs;
  bs.start(g, s, cellStart, e, lastStored);
  for (uint32_t t = bs.next(g, s, cellStart); t != SLOT_NONE; t = bs.next(g, s, cellStart)) {
    const float4 gt = Gin[t];
    const uint32_t ft = __float_as_uint(gt.w);
    if ((ft & F_COLLIDER) != F_COLLIDER || !in_row_of(s, ft, t, e)) continue;
    if (surely_apart(x, y, r, gt.x, gt.y, gt.z)) continue;
    exact_pair(p, s, frame, substep, e, x, y, r, fw, t, gt.x, gt.y, gt.z, ft, false, acc);
  }
}

// ---- K6, second form: mask walk + converged exact pass -----------------------------------------
// Same arithmetic, same order, same results as k_substep; what changed is the control flow.
//   phase 1  walks the row without a data-dependent branch: row word, 16-byte partner gather,
//            membership bits, float32 "surely apart" test -> ONE bit per row entry in a 64-bit
//            register mask (rows longer than 64 are walked 64 entries at a time).  Four entries in
//            flight; the only divergence left is the row length.
//   phase 2  visits the set bits in ascending row position and runs the binary64 pair code from the
//            entity's own point of view: d = me - partner, so the displacement of BOTH endpoints of
//            a pair is +d/dist * h on its own side (negation is exact, so this is bit for bit the
//            reference's `x[i] += ux` / `x[j] -= ux`, physics_worker.js:519-547); the static /
//            trigger bookkeeping collapses to "do I move" and "does my partner".
// A lane whose partner's row is capped (F_CAPPED) resolves membership with one comparison against the
// slot that closes that row (LSLOT), inside phase 2 and only for entries that passed the float32 test.
__device__ __noinline__ void sweep_pair_coincident(const FrameConst* __restrict__ fc, const Params* __restrict__ pp, uint32_t frame,
                                                   uint32_t substep, uint32_t e, float x, float y, float r, uint32_t fw, uint32_t t,
                                                   float4 gt, bool lower, SubstepAcc& acc) {
  exact_pair(*pp, fc->s, frame, substep, e, x, y, r, fw, t, gt.x, gt.y, gt.z, __float_as_uint(gt.w), lower, acc);
}

__device__ __forceinline__ void sweep_pair(const Params* __restrict__ pp, double strength, const FrameConst* __restrict__ fc,
                                           const Counters* __restrict__ ctr, uint32_t substep, uint32_t e,
                                           float x, float y, float r, uint32_t fw, uint32_t t, float4 gt, bool lower,
                                           SubstepAcc& acc) {
  const double dx = dsub((double)x, (double)gt.x);                   // :447-449, seen from this entity
  const double dy = dsub((double)y, (double)gt.y);
  const double dist2 = dadd(dmul(dx, dx), dmul(dy, dy));
  const double minDist = dadd((double)r, (double)gt.z);              // :452
  if (dist2 >= dmul(minDist, minDist)) return;                       // :455
  const double dist = __dsqrt_rn(dist2);
  if (dist == 0) { sweep_pair_coincident(fc, pp, ctr->frame, substep, e, x, y, r, fw, t, gt, lower, acc); return; }   // :460-507
  const double depth = dsub(minDist, dist);                          // :510
  if (!(depth > 0)) return;
  acc.hits++;                                                        // :551-552
  acc.outHits += lower ? 1u : 0u;
  const uint32_t ft = __float_as_uint(gt.w);
  if (((fw | ft) & F_TRIGGER) || (fw & F_STATIC)) return;            // logged, not moved (:512); a static side stays
  double nx, ny;
  ddiv2(dx, dy, dist, nx, ny);                                       // :519-520
  const double corr = dmul(depth, strength);                         // :528
  const double h = (ft & F_STATIC) ? corr : dmul(corr, 0.5);         // :532-546
  acc.x = fround(dadd((double)acc.x, dmul(nx, h)));
  acc.y = fround(dadd((double)acc.y, dmul(ny, h)));
}

// record of slot t in an array of 16-byte records: one IMAD.WIDE (slots are below 2^30)
__device__ __forceinline__ const float4* slot_rec(const float4* __restrict__ G, uint32_t t) {
  return reinterpret_cast<const float4*>(reinterpret_cast<const char*>(G) + (size_t)t * 16u);
}

// Tuning (config 4, 16M entities, one B200, ms per sweep): 256 threads / 64 registers 1.81; 128 threads
// 1.68; + next batch's row words prefetched 1.63; 48 registers (10 blocks/SM) 1.56; 40 registers
// (12 blocks/SM) 1.50; 36 / 32 registers 1.62 / 1.63 (spills).  The walk is bound by the latency of
// the partner gathers, so occupancy pays until the spills start.
#ifndef WEED_K6V2_THREADS
#define WEED_K6V2_THREADS 128
#endif
#ifndef WEED_K6V2_MINBLOCKS
#define WEED_K6V2_MINBLOCKS 12
#endif
static constexpr int K6V2_THREADS = WEED_K6V2_THREADS;

#ifndef WEED_HEAVY_SPLIT
#define WEED_HEAVY_SPLIT 1
#endif
static constexpr bool HEAVY_SPLIT = WEED_HEAVY_SPLIT != 0;   // 0: k_sweep walks pool rows and resumed scans itself, one thread each

template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(K6V2_THREADS, WEED_K6V2_MINBLOCKS)
k_sweep(GridDims g, const Params* __restrict__ pp, BySlot s, const float4* __restrict__ Gin,
        float4* __restrict__ Gout, const uint32_t* __restrict__ cellStart, const Counters* __restrict__ ctr,
        uint32_t substep, const FrameConst* __restrict__ fc) {
  // Blocks are dispatched in index order and slots ascend with the cell row: a bed settled on the floor (the
  // last rows) would be the last blocks, heavy work with nothing left to run beside it (config 3, frame 100:
  // SMs busy 51 % of the sweep).  Walk the slots from the top.
  const uint32_t e = (gridDim.x - 1u - blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= cellStart[g.cells]) return;
  const float4 gme = Gin[e];
  float2 pxy;
  if (FIRST) { const float4 hi = s.SA[2 * (size_t)e + 1]; pxy = make_float2(hi.x, hi.y); }
  else pxy = s.PXY[e];
  const float x = gme.x, y = gme.y, r = gme.z;
  const uint32_t fw = __float_as_uint(gme.w);
  if (HEAVY_SPLIT && (fw & (F_XPOOL | F_XOVER)) && s.XHEAD[e] == 0) return;   // a "popular" entity of a pile: k_sweep_heavy
  SubstepAcc acc; acc.x = x; acc.y = y; acc.hits = 0; acc.outHits = 0;
  if ((fw & F_COLLIDER) == F_COLLIDER) {                         // :430
    const uint32_t cnt = s.NCNT[e] >> 16;                        // the INTERNAL row: API row + lower-id partners past the cap
    const uint32_t xhead = s.XHEAD[e];
    if (xhead == 0) {
      const double strength = pp->responseStrength;
      // Walks `n` row words rp[0], rp[stride], ...: the internal row (slot-major: stride Npad; Npad * Mint
      // < 2^32, so 32-bit indices) or a row of the overflow pool (stride 1).  Rows are padded to a multiple
      // of four entries with words that carry no membership bit, so a batch never needs a bounds test.
      auto walk = [&](const uint32_t* __restrict__ rp, uint32_t stride, uint32_t n) {
        for (uint32_t kb = 0; kb < n; kb += 64) {
          const uint32_t ke = min(n, kb + 64u);
          // ---- phase 1: one bit per row entry that may overlap --------------------------------------
          unsigned long long mask = 0;
          uint32_t idx = kb * stride;
          uint32_t nx[4];
#pragma unroll
          for (int u = 0; u < 4; u++) nx[u] = rp[idx + (uint32_t)u * stride];
          for (uint32_t k = kb; k < ke; k += 4, idx += 4 * stride) {
            uint32_t wd[4];
            float4 gt[4];
#pragma unroll
            for (int u = 0; u < 4; u++) wd[u] = nx[u];
            if (k + 4 < ke) {
#pragma unroll
              for (int u = 0; u < 4; u++) nx[u] = rp[idx + (uint32_t)(4 + u) * stride];     // the next batch's words, in flight behind the gathers
            }
#pragma unroll
            for (int u = 0; u < 4; u++) gt[u] = __ldg(slot_rec(Gin, wd[u] & NS_SLOT_MASK));
            uint32_t nib = 0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const uint32_t ft = __float_as_uint(gt[u].w);
              const bool apart = surely_apart(x, y, r, gt[u].x, gt[u].y, gt[u].z);
              const bool cand = ((ft & F_COLLIDER) == F_COLLIDER) & (wd[u] >= NS_OUT) & !apart;   // :441; OUT or BACK set
              nib |= cand ? (1u << u) : 0u;
            }
            mask |= (unsigned long long)nib << (k - kb);
          }
          // ---- phase 2: exact pair code on the marked entries, in row order --------------------------
          while (mask) {
            const uint32_t k = kb + (uint32_t)__ffsll((long long)mask) - 1u;
            mask &= mask - 1;
            const uint32_t wd = rp[k * stride];
            const uint32_t t = wd & NS_SLOT_MASK;
            const float4 gt = __ldg(slot_rec(Gin, t));
            const bool lower = (wd & NS_OUT) != 0;
            if (!lower && !in_row_of(s, __float_as_uint(gt.w), t, e)) continue;
            sweep_pair(pp, strength, fc, ctr, substep, e, x, y, r, fw, t, gt, lower, acc);
          }
        }
      };
      walk(s.NST + e, g.Npad, cnt);
      if (fw & (F_XPOOL | F_XOVER)) {                              // a "popular" entity of a dense pile
        uint32_t lastStored = cnt ? (s.NST[(cnt - 1) * g.Npad + e] & NS_SLOT_MASK) : 0u;
        if (fw & F_XPOOL) {
          const uint32_t pid = s.XPID[e], pcnt = s.XRCNT[pid];
          const uint32_t* pool = s.XR + (size_t)pid * XPOOL_ROW;
          walk(pool, 1u, pcnt);
          if (pcnt) lastStored = pool[pcnt - 1] & NS_SLOT_MASK;
        }
        if (fw & F_XOVER) sweep_resumed_scan(fc, pp, Gin, cellStart, ctr->frame, substep, e, x, y, r, fw, lastStored, acc);
      }
    } else {
      substep_slow(fc, pp, Gin, cellStart, ctr->frame, substep, e, x, y, r, fw, cnt, xhead, acc);
    }
  }
  const uint32_t cc = ((fw >> F_CC_SHIFT) + acc.hits) & 0xFFu;    // Uint8 wrap (:551-552)
  if (LAST) {
    float4* o = reinterpret_cast<float4*>(s.OUT + e);
    o[0] = make_float4(acc.x, acc.y, pxy.x, pxy.y);
    o[1] = make_float4(__uint_as_float(cc | ((acc.outHits & 0x7FFFFFu) << 8) | ((fw & F_OWNED) ? 0x80000000u : 0u)), 0.f, 0.f, 0.f);
  } else {
    // boundary pass of the next sweep on my own result (:344-376)
    if ((fw & F_DYNAMIC_MASK) == F_DYNAMIC_VAL && !clear_of_walls(g, acc.x, acc.y, r))
      apply_bounds(g, pp->boundaryElasticity, r, acc.x, acc.y, pxy.x, pxy.y);
    Gout[e] = make_float4(acc.x, acc.y, r, __uint_as_float((fw & 0xFFFF00FFu) | (cc << F_CC_SHIFT)));
    s.PXY[e] = pxy;
  }
}

// ---- K6 for the popular entities of a pile: one warp each ------------------------------------------------
// An entity with an overflow-pool row (F_XPOOL: up to 64 + 512 lower-id partners past its cap) or a resumed scan
// (F_XOVER: more than that; the candidates of its whole window) kept ONE thread of k_sweep busy for
// milliseconds while the rest of the GPU had finished (config 3, frame 300: 5.4 ms for both sweeps, SMs busy
// half of it).  J-order lets a warp share the work: every pair is evaluated on start-of-sweep positions, so 32
// lanes evaluate 32 partners at once and the displacements are then ADDED in ascending partner order, with the
// float32 rounding after each — the same sequence of operations the single thread performs, bit for bit.
// Streams in the order of substep_slow: internal row, pool row, resumed scan.  Entities with an explicit list
// (XHEAD) stay with k_sweep's merge path.
struct LaneMove { double ax, ay; bool hit, out, move; };
__device__ __forceinline__ LaneMove lane_pair(const Params& p, const BySlot& s, uint32_t frame, uint32_t substep, uint32_t e,
                                              float x, float y, float r, uint32_t fw, uint32_t t, float4 gt, bool lower) {
  const uint32_t ft = __float_as_uint(gt.w);
  PairMove m;
  if (lower) m = pair_eval(p, frame, substep, s.SA, e, t, x, y, r, fw, gt.x, gt.y, gt.z, ft);
  else       m = pair_eval(p, frame, substep, s.SA, t, e, gt.x, gt.y, gt.z, ft, x, y, r, fw);
  LaneMove lm;
  lm.hit = m.hit; lm.out = m.hit && lower;
  lm.move = m.hit && (lower ? m.moveI : m.moveJ);
  lm.ax = lower ? m.mx : -m.mx; lm.ay = lower ? m.my : -m.my;
  return lm;
}
// the warp's 32 results into the accumulator, ascending lane = ascending partner slot
__device__ __forceinline__ void lanes_apply(const LaneMove& lm, bool valid, SubstepAcc& acc) {
  acc.hits += (uint32_t)__popc(__ballot_sync(0xffffffffu, valid && lm.hit));
  acc.outHits += (uint32_t)__popc(__ballot_sync(0xffffffffu, valid && lm.out));
  uint32_t mb = __ballot_sync(0xffffffffu, valid && lm.move);
  while (mb) {
    const int l = __ffs((int)mb) - 1;
    mb &= mb - 1;
    const double ax = __shfl_sync(0xffffffffu, lm.ax, l), ay = __shfl_sync(0xffffffffu, lm.ay, l);
    acc.x = fround(dadd((double)acc.x, ax));
    acc.y = fround(dadd((double)acc.y, ay));
  }
}

static constexpr int K6H_THREADS = 256, K6H_BLOCKS = 148 * 4;
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(K6H_THREADS, 4)            // latency-bound: 32 resident warps per SM, one wave
k_sweep_heavy(GridDims g, const Params* __restrict__ pp, BySlot s, const float4* __restrict__ Gin,
              float4* __restrict__ Gout, const uint32_t* __restrict__ cellStart, const Counters* __restrict__ ctr,
              uint32_t substep) {
  const uint32_t nHeavy = ctr->nHeavy;
  if (nHeavy == 0) return;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warpsTotal = gridDim.x * (blockDim.x >> 5);
  const Params& p = *pp;
  const uint32_t frame = ctr->frame;
  for (uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < nHeavy; w += warpsTotal) {
    const uint32_t e = s.HEAVY[w];
    if (s.XHEAD[e] != 0) continue;                                 // explicit list: k_sweep merged it
    const float4 gme = Gin[e];
    float2 pxy;
    if (FIRST) { const float4 hi = s.SA[2 * (size_t)e + 1]; pxy = make_float2(hi.x, hi.y); }
    else pxy = s.PXY[e];
    const float x = gme.x, y = gme.y, r = gme.z;
    const uint32_t fw = __float_as_uint(gme.w);
    SubstepAcc acc; acc.x = x; acc.y = y; acc.hits = 0; acc.outHits = 0;
    if ((fw & F_COLLIDER) == F_COLLIDER) {
      const uint32_t cnt = s.NCNT[e] >> 16;
      uint32_t lastStored = cnt ? (s.NST[(cnt - 1) * g.Npad + e] & NS_SLOT_MASK) : 0u;
      // row words rp[0], rp[stride], ...: 32 at a time, two batches ahead — the words of batch i+2 and the partner
      // records of batch i+1 are in flight while batch i is evaluated (the walk is a chain of dependent gathers)
      auto walk = [&](const uint32_t* __restrict__ rp, uint32_t stride, uint32_t n) {
        auto word = [&](uint32_t kb) { const uint32_t k = kb + lane; return k < n ? rp[k * stride] : 0u; };
        auto rec = [&](uint32_t wd) { return wd >= NS_OUT ? __ldg(slot_rec(Gin, wd & NS_SLOT_MASK)) : make_float4(0.f, 0.f, 0.f, 0.f); };
        uint32_t wd0 = word(0), wd1 = word(32);
        float4 gt0 = rec(wd0);
        for (uint32_t kb = 0; kb < n; kb += 32) {
          const float4 gt1 = rec(wd1);
          const uint32_t wd2 = word(kb + 64);
          const uint32_t wd = wd0;
          const float4 gt = gt0;
          bool cand = wd >= NS_OUT;                                // OUT or BACK set (padding words carry neither)
          const uint32_t t = wd & NS_SLOT_MASK;
          const bool lower = (wd & NS_OUT) != 0;
          if (cand) {
            const uint32_t ft = __float_as_uint(gt.w);
            cand = (ft & F_COLLIDER) == F_COLLIDER && !surely_apart(x, y, r, gt.x, gt.y, gt.z) &&
                   (lower || in_row_of(s, ft, t, e));
          }
          LaneMove lm; lm.hit = lm.out = lm.move = false; lm.ax = lm.ay = 0;
          if (cand) lm = lane_pair(p, s, frame, substep, e, x, y, r, fw, t, gt, lower);
          lanes_apply(lm, cand, acc);
          wd0 = wd1; wd1 = wd2; gt0 = gt1;
        }
      };
      walk(s.NST + e, g.Npad, cnt);
      if (fw & F_XPOOL) {
        const uint32_t pid = s.XPID[e], pcnt = s.XRCNT[pid];
        const uint32_t* pool = s.XR + (size_t)pid * XPOOL_ROW;
        walk(pool, 1u, pcnt);
        if (pcnt) lastStored = pool[pcnt - 1] & NS_SLOT_MASK;
      }
      if (fw & F_XOVER) {                                           // BeyondScan, 32 candidates at a time
        const float4 me = s.CXY[e];
        const uint32_t id = __float_as_uint(me.w) & ~CX_EDGE, vrBits = __float_as_uint(me.z);
        const double myX = me.x, myY = me.y, vrSq = dmul((double)me.z, (double)me.z);
        int32_t myCol, myRow;
        cell_of(g, me.x, me.y, myCol, myRow);
        const int4 win = s.WIN[e];
        for (int32_t row = win.x; row <= win.y; row++) {
          const uint32_t a = max(cellStart[(uint32_t)row * g.cols + win.z], lastStored + 1u);
          const uint32_t b = cellStart[(uint32_t)row * g.cols + win.w + 1];
          for (uint32_t t0 = a; t0 < b; t0 += 32) {
            const uint32_t tc = t0 + lane;
            bool cand = false;
            float4 gt = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tc < b && e <= __ldg(s.LSLOT + tc)) {                 // in_row_of, first: most rows of a pile closed before me
              const float4 c = __ldg(s.CXY + tc);
              if ((__float_as_uint(c.w) & ~CX_EDGE) < id) {          // higher ids: my own pairs, lost to the cap
                const double dX = dsub((double)c.x, myX), dY = dsub((double)c.y, myY);
                const double d2 = dadd(dmul(dX, dX), dmul(dY, dY));
                if (d2 < vrSq && d2 > 0 && scan_accepts(g, s, tc, myX, myY, myCol, myRow, vrBits)) {
                  gt = __ldg(slot_rec(Gin, tc));
                  const uint32_t ft = __float_as_uint(gt.w);
                  cand = (ft & F_COLLIDER) == F_COLLIDER && in_row_of(s, ft, tc, e) && !surely_apart(x, y, r, gt.x, gt.y, gt.z);
                }
              }
            }
            LaneMove lm; lm.hit = lm.out = lm.move = false; lm.ax = lm.ay = 0;
            if (cand) lm = lane_pair(p, s, frame, substep, e, x, y, r, fw, tc, gt, false);
            lanes_apply(lm, cand, acc);
          }
        }
      }
    }
    if (lane == 0) {                                                // k_sweep's epilogue
      const uint32_t cc = ((fw >> F_CC_SHIFT) + acc.hits) & 0xFFu;
      if (LAST) {
        float4* o = reinterpret_cast<float4*>(s.OUT + e);
        o[0] = make_float4(acc.x, acc.y, pxy.x, pxy.y);
        o[1] = make_float4(__uint_as_float(cc | ((acc.outHits & 0x7FFFFFu) << 8) | ((fw & F_OWNED) ? 0x80000000u : 0u)), 0.f, 0.f, 0.f);
      } else {
        if ((fw & F_DYNAMIC_MASK) == F_DYNAMIC_VAL && !clear_of_walls(g, acc.x, acc.y, r))
          apply_bounds(g, pp->boundaryElasticity, r, acc.x, acc.y, pxy.x, pxy.y);
        Gout[e] = make_float4(acc.x, acc.y, r, __uint_as_float((fw & 0xFFFF00FFu) | (cc << F_CC_SHIFT)));
        s.PXY[e] = pxy;
      }
    }
  }
}

// ---- K6, tiled form (measured alternative, WEED_FLAG_K6_TILE): TMA-staged partners and row words ------
// One block per tile of TILE consecutive slots.  Thread 0 issues 1-D bulk copies (cp.async.bulk,
// completion on an mbarrier): the tile's partner ranges of the sweep input (TileDesc: at most three
// contiguous slot ranges hold every partner of every entity of the tile) and the tile's row words,
// 8 rows x TILE words first, the rest of a 16-row group once the longest row of the tile is known.
// The walk of k_sweep then runs out of shared memory.  Tiles whose ranges do not fit (a block that
// straddles two grid rows, an observer with a huge visualRange, hundreds of entities per cell)
// gather from global memory as k_sweep does.
// Measured (config 4, 16M, one B200): 2.12 ms per sweep against 1.50 for k_sweep — the bulk copies
// remove the gather stalls, but a block cannot overlap its own fill with its own walk, 24.6 KB of
// shared memory hold occupancy at 8 blocks, and the slot -> tile position arithmetic adds 40 % to the
// instruction count; a 32-row group (2.57 ms) is worse still.  Kept selectable for the cross-check tests.
static constexpr int K6T_CAP = 1024;        // partner records per tile (16 KB)
static constexpr int K6T_GROUP = 16;        // row entries walked per group: a wave of 8 rows, then the rest

struct TileMap { uint32_t a1, a2, d0, d1, d2; };   // slot -> position in the staged ranges
__device__ __forceinline__ uint32_t tile_pos(const TileMap& m, uint32_t t) {
  return t + (t >= m.a2 ? m.d2 : (t >= m.a1 ? m.d1 : m.d0));
}

template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(TILE, 8)
k_sweep_tile(GridDims g, const Params* __restrict__ pp, BySlot s, const float4* __restrict__ Gin,
             float4* __restrict__ Gout, const uint32_t* __restrict__ cellStart, const Counters* __restrict__ ctr,
             uint32_t substep, const FrameConst* __restrict__ fc) {
  __shared__ __align__(128) float4 tG[K6T_CAP];
  __shared__ __align__(128) uint32_t tRow[K6T_GROUP][TILE];
  __shared__ __align__(8) unsigned long long bars[3];       // 0: partner ranges, 1 / 2: row words, first wave / rest
  __shared__ uint32_t sMax[TILE / 32];
  const uint32_t tid = threadIdx.x;
  const uint32_t tile0 = blockIdx.x * TILE;
  const uint32_t A = cellStart[g.cells];
  if (tile0 >= A) return;                                    // whole block past the last slot
  const uint32_t e = tile0 + tid;
  const TileDesc td = s.TD[blockIdx.x];
  const uint32_t total = td.n[0] + td.n[1] + td.n[2];
  const bool tiled = (td.shape & TD_OK) && total <= (uint32_t)K6T_CAP;
  const uint32_t* __restrict__ NST = s.NST;
  const uint32_t Npad = g.Npad;
  if (tid == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    mbar_init_fence();
    mbar_expect_tx(&bars[1], 8u * TILE * 4u);                // rows 0..7 of this tile: needed by every non-empty row
#pragma unroll
    for (uint32_t j = 0; j < 8; j++) bulk_g2s(&tRow[j][0], NST + j * Npad + tile0, TILE * 4u, &bars[1]);
    if (tiled) {
      mbar_expect_tx(&bars[0], total * 16u);
      uint32_t off = 0;
#pragma unroll
      for (int r = 0; r < 3; r++)
        if (td.n[r]) { bulk_g2s(tG + off, Gin + td.a[r], td.n[r] * 16u, &bars[0]); off += td.n[r]; }
    }
  }
  TileMap tm;
  tm.a1 = td.n[1] ? td.a[1] : 0xFFFFFFFFu;
  tm.a2 = td.n[2] ? td.a[2] : 0xFFFFFFFFu;
  tm.d0 = 0u - td.a[0];
  tm.d1 = td.n[0] - td.a[1];
  tm.d2 = td.n[0] + td.n[1] - td.a[2];
  const bool live = e < A;
  float4 gme = make_float4(0.f, 0.f, 0.f, 0.f);
  float2 pxy = make_float2(0.f, 0.f);
  if (live) {
    gme = Gin[e];
    if (FIRST) { const float4 hi = s.SA[2 * (size_t)e + 1]; pxy = make_float2(hi.x, hi.y); }
    else pxy = s.PXY[e];
  }
  const float x = gme.x, y = gme.y, r = gme.z;
  const uint32_t fw = __float_as_uint(gme.w);
  SubstepAcc acc; acc.x = x; acc.y = y; acc.hits = 0; acc.outHits = 0;
  const bool collider = live && (fw & F_COLLIDER) == F_COLLIDER;     // :430
  uint32_t cnt = 0, xhead = 0;
  if (collider) { cnt = s.NCNT[e] >> 16; xhead = s.XHEAD[e]; }
  const bool slow = xhead != 0 || (collider && (fw & (F_XOVER | F_XPOOL)));
  const uint32_t walk = slow ? 0u : cnt;                     // rows walked here; explicit lists take substep_slow
  const uint32_t wmax = __reduce_max_sync(0xffffffffu, walk);
  if ((tid & 31) == 0) sMax[tid >> 5] = wmax;
  __syncthreads();                                           // barrier inits and the warps' maxima visible to everyone
  uint32_t blockMax = 0;
#pragma unroll
  for (int w = 0; w < TILE / 32; w++) blockMax = max(blockMax, sMax[w]);
  if (tid == 0 && blockMax > 8) {
    const uint32_t nr = min((uint32_t)K6T_GROUP - 8u, (blockMax - 8u + 3u) & ~3u);
    mbar_expect_tx(&bars[2], nr * TILE * 4u);
    for (uint32_t j = 0; j < nr; j++) bulk_g2s(&tRow[8 + j][0], NST + (8 + j) * Npad + tile0, TILE * 4u, &bars[2]);
  }
  const double strength = pp->responseStrength;
  if (tiled) mbar_wait(&bars[0], 0);
  auto walk_rows = [&](auto tiledTag) {
    constexpr bool TILED = decltype(tiledTag)::value;
    uint32_t par = 0;                                          // phase parity of bars[1], bars[2]
    for (uint32_t k0 = 0;; k0 += K6T_GROUP) {
      mbar_wait(&bars[1], par);                                // rows k0 .. k0+7 (always issued for k0 == 0)
      const uint32_t ke = min(walk, k0 + (uint32_t)K6T_GROUP);
      // ---- phase 1: one bit per row entry that may overlap ------------------------------------------
      uint32_t mask = 0;
      for (uint32_t k = k0; k < ke; k += 4) {
        if (k == k0 + 8) mbar_wait(&bars[2], par);             // the rest of the group
        uint32_t wd[4];
        float4 gt[4];
#pragma unroll
        for (int u = 0; u < 4; u++) wd[u] = tRow[k - k0 + (uint32_t)u][tid];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const uint32_t t = wd[u] & NS_SLOT_MASK;
          if (TILED) gt[u] = tG[tile_pos(tm, t)]; else gt[u] = __ldg(Gin + t);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const uint32_t ft = __float_as_uint(gt[u].w);
          const bool apart = surely_apart(x, y, r, gt[u].x, gt[u].y, gt[u].z);
          const bool cand = ((ft & F_COLLIDER) == F_COLLIDER) & (wd[u] >= NS_OUT) & !apart;   // :441; OUT or BACK set
          mask |= cand ? (1u << (k - k0 + (uint32_t)u)) : 0u;
        }
      }
      // ---- phase 2: exact pair code on the marked entries, in row order ------------------------------
      while (mask) {
        const uint32_t j = (uint32_t)__ffs((int)mask) - 1u;
        mask &= mask - 1;
        const uint32_t wd = tRow[j][tid];
        const uint32_t t = wd & NS_SLOT_MASK;
        float4 gt;
        if (TILED) gt = tG[tile_pos(tm, t)]; else gt = __ldg(Gin + t);
        const bool lower = (wd & NS_OUT) != 0;
        if (!lower && !in_row_of(s, __float_as_uint(gt.w), t, e)) continue;
        sweep_pair(pp, strength, fc, ctr, substep, e, x, y, r, fw, t, gt, lower, acc);
      }
      if (k0 + K6T_GROUP >= blockMax) break;
      // ---- next group of rows (dense tiles only) --------------------------------------------------------
      // lanes that never waited on bars[2] in this group still observe its phase before it is re-armed
      if (blockMax > k0 + 8) mbar_wait(&bars[2], par);
      __syncthreads();                                         // everyone is done with tRow
      par ^= 1;
      if (tid == 0) {
        const uint32_t left = blockMax - (k0 + K6T_GROUP);
        const uint32_t n1 = min(8u, (left + 3u) & ~3u);
        mbar_expect_tx(&bars[1], n1 * TILE * 4u);
        for (uint32_t j = 0; j < n1; j++) bulk_g2s(&tRow[j][0], NST + (k0 + K6T_GROUP + j) * Npad + tile0, TILE * 4u, &bars[1]);
        if (left > 8) {
          const uint32_t n2 = min((uint32_t)K6T_GROUP - 8u, (left - 8u + 3u) & ~3u);
          mbar_expect_tx(&bars[2], n2 * TILE * 4u);
          for (uint32_t j = 0; j < n2; j++) bulk_g2s(&tRow[8 + j][0], NST + (k0 + K6T_GROUP + 8 + j) * Npad + tile0, TILE * 4u, &bars[2]);
        }
      }
    }
  };
  if (tiled) walk_rows(std::true_type{}); else walk_rows(std::false_type{});
  if (!live) return;
  if (slow) substep_slow(fc, pp, Gin, cellStart, ctr->frame, substep, e, x, y, r, fw, cnt, xhead, acc);
  const uint32_t cc = ((fw >> F_CC_SHIFT) + acc.hits) & 0xFFu;    // Uint8 wrap (:551-552)
  if (LAST) {
    float4* o = reinterpret_cast<float4*>(s.OUT + e);
    o[0] = make_float4(acc.x, acc.y, pxy.x, pxy.y);
    o[1] = make_float4(__uint_as_float(cc | ((acc.outHits & 0x7FFFFFu) << 8) | ((fw & F_OWNED) ? 0x80000000u : 0u)), 0.f, 0.f, 0.f);
  } else {
    // boundary pass of the next sweep on my own result (:344-376)
    if ((fw & F_DYNAMIC_MASK) == F_DYNAMIC_VAL && !clear_of_walls(g, acc.x, acc.y, r))
      apply_bounds(g, pp->boundaryElasticity, r, acc.x, acc.y, pxy.x, pxy.y);
    Gout[e] = make_float4(acc.x, acc.y, r, __uint_as_float((fw & 0xFFFF00FFu) | (cc << F_CC_SHIFT)));
    s.PXY[e] = pxy;
  }
}

// ---- WB + K7: results back to id order, collisionData -------------------------------------
// k_writeback gathers each entity's result sector by id, writes the by-id state coalesced
// and leaves the per-tile count of outgoing colliding pairs; k_pair_scan turns the tile counts
// into exclusive prefixes (id order) and the pair count; k_pair_emit lets only the tiles whose
// prefix is below maxCollisionPairs re-derive their pairs in row order — the reference's
// emission order (i ascending, then row position; physics_worker.js:555-567).
__device__ __forceinline__ uint32_t wb_out_count(const BySlot& s, const uint32_t* __restrict__ slotOf, uint32_t i,
                                                 uint32_t N, uint32_t& slot) {
  slot = SLOT_NONE;
  if (i >= N) return 0;
  slot = slotOf[i];
  if (slot == SLOT_NONE) return 0;
  const uint32_t meta = __float_as_uint(reinterpret_cast<const float4*>(s.OUT + slot)[1].x);
  return (meta >> 31) ? ((meta >> 8) & 0x7FFFFFu) : 0u;   // pairs are logged by the slab that owns the lower id
}

__global__ void __launch_bounds__(WB_THREADS)
k_writeback(GridDims g, ById d, BySlot s, const uint32_t* __restrict__ slotOf, uint32_t* __restrict__ tileCount) {
  __shared__ uint32_t s_warp[WB_THREADS / 32];
  const uint32_t i = blockIdx.x * WB_THREADS + threadIdx.x;
  uint32_t outCnt = 0;
  if (i < g.N) {
    const uint32_t slot = slotOf[i];
    if (slot != SLOT_NONE) {
      const float4* o = reinterpret_cast<const float4*>(s.OUT + slot);
      const float4 o0 = o[0];
      const uint32_t meta = __float_as_uint(o[1].x);
      d.DP[i] = o0;
      d.CC[i] = (uint8_t)(meta & 0xFFu);
      if (meta >> 31) outCnt = (meta >> 8) & 0x7FFFFFu;
    }
  }
  for (int o = 16; o; o >>= 1) outCnt += __shfl_xor_sync(0xffffffffu, outCnt, o);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = outCnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < WB_THREADS / 32; w++) t += s_warp[w];
    tileCount[blockIdx.x] = t;
  }
}

// one block: exclusive prefix of the tile counts (a few 10^4 values), 4096 at a time: every thread takes
// four consecutive counts (one 16-byte load, the next chunk's already in flight), the block scans the
// 1024 partial sums with shuffles, a running carry links the chunks.  (The first version gave every
// thread a contiguous run of tiles: strided, uncoalesced loads — 0.055 ms at 62 k tiles.)
__global__ void __launch_bounds__(1024)
k_pair_scan(const uint32_t* __restrict__ tileCount, uint32_t* __restrict__ tilePrefix, uint32_t numTiles,
            uint32_t maxPairs, Counters* ctr, int32_t* __restrict__ coll) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  auto load4 = [&](uint32_t base) {
    uint4 v = make_uint4(0, 0, 0, 0);
    const uint32_t i = base + threadIdx.x * 4;
    if (i + 3 < numTiles) v = *reinterpret_cast<const uint4*>(tileCount + i);     // tileCount is cudaMalloc'ed: 16-byte aligned
    else {
      if (i < numTiles) v.x = tileCount[i];
      if (i + 1 < numTiles) v.y = tileCount[i + 1];
      if (i + 2 < numTiles) v.z = tileCount[i + 2];
    }
    return v;
  };
  uint4 nxt = load4(0);
  __syncthreads();
  for (uint32_t base = 0; base < numTiles; base += 4096) {
    const uint4 c = nxt;
    if (base + 4096 < numTiles) nxt = load4(base + 4096);
    const uint32_t mine = c.x + c.y + c.z + c.w;
    uint32_t inc = mine;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (uint32_t)o) inc += v;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    const uint32_t carry = s_carry;
    uint32_t w = s_warp[lane], winc = w;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= (uint32_t)o) winc += v;
    }
    const uint32_t warpExcl = __shfl_sync(0xffffffffu, winc - w, warp);
    const uint32_t total = __shfl_sync(0xffffffffu, winc, 31);
    uint32_t e = carry + warpExcl + (inc - mine);
    const uint32_t i = base + threadIdx.x * 4;
    if (i < numTiles) tilePrefix[i] = e;
    e += c.x; if (i + 1 < numTiles) tilePrefix[i + 1] = e;
    e += c.y; if (i + 2 < numTiles) tilePrefix[i + 2] = e;
    e += c.z; if (i + 3 < numTiles) tilePrefix[i + 3] = e;
    __syncthreads();                               // everybody has read s_carry and s_warp
    if (threadIdx.x == 0) s_carry = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    ctr->collisionPairs = s_carry;
    if (coll) coll[0] = (int32_t)min(s_carry, maxPairs);      // :565-567
  }
}

static constexpr uint32_t PE_BATCH = 8;   // row words / partner records in flight per thread in k_pair_emit

__global__ void __launch_bounds__(WB_THREADS)
k_pair_emit(GridDims g, const Params* __restrict__ pp, ById d, BySlot s, const float4* __restrict__ Glast,
            uint32_t gs, const uint32_t* __restrict__ slotOf, const uint32_t* __restrict__ tilePrefix,
            const Counters* __restrict__ ctr, int32_t* __restrict__ coll, uint32_t lastSubstep, uint32_t numTiles) {
  __shared__ uint32_t s_warp[WB_THREADS / 32];
  // The tile prefixes ascend, so the tiles below the cap are a prefix of the tile sequence: a small grid strides
  // over the tiles and every block leaves at the first tile past the cap (a launch of one block per tile spent
  // 0.06 ms at 16 M entities on blocks that read one word and returned).
  for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
  const uint32_t tileBase = tilePrefix[tile];
  if (tileBase >= g.maxPairs) return;                              // this tile and all later ones are past the cap
  const uint32_t i = tile * WB_THREADS + threadIdx.x;
  uint32_t slot;
  const uint32_t outCnt = wb_out_count(s, slotOf, i, g.N, slot);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = outCnt;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += v;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t base = tileBase + (inc - outCnt);
  for (uint32_t w = 0; w < warp; w++) base += s_warp[w];
  if (outCnt != 0 && base < g.maxPairs) {
  // re-derive my colliding outgoing pairs on the last sweep's start positions
  const Params p = *pp;
  const float4 gme = Glast[(size_t)slot * gs];
  float x = gme.x, y = gme.y;
  const uint32_t fw = __float_as_uint(gme.w);
  const uint32_t cnt = s.NCNT[slot] & 0xFFFFu;                    // outgoing pairs live in the API row
  const uint32_t frame = ctr->frame;
  // The few tiles below the cap decide this kernel's duration, and each thread's walk is a chain of dependent
  // loads: fetch PE_BATCH row words, then their partners, then evaluate in row order; stop at the entity's own
  // count of colliding outgoing pairs (known from the write-back).
  uint32_t found = 0;
  for (uint32_t k0 = 0; k0 < cnt && found < outCnt && base < g.maxPairs; k0 += PE_BATCH) {
    uint32_t wd[PE_BATCH];
    float4 gt[PE_BATCH];
#pragma unroll
    for (uint32_t j = 0; j < PE_BATCH; j++) wd[j] = (k0 + j < cnt) ? s.NST[(size_t)(k0 + j) * g.Npad + slot] : 0u;
#pragma unroll
    for (uint32_t j = 0; j < PE_BATCH; j++)
      if (wd[j] & NS_OUT) gt[j] = Glast[(size_t)(wd[j] & NS_SLOT_MASK) * gs];
#pragma unroll
    for (uint32_t j = 0; j < PE_BATCH; j++) {
      if (!(wd[j] & NS_OUT) || base >= g.maxPairs) continue;
      const uint32_t t = wd[j] & NS_SLOT_MASK;
      const uint32_t ft = __float_as_uint(gt[j].w);
      if ((ft & F_COLLIDER) != F_COLLIDER) continue;
      float xt, yt;
      partner_pos(g, gt[j], xt, yt);
      if (surely_apart(x, y, gme.z, xt, yt, gt[j].z)) continue;
      SubstepAcc acc; acc.x = x; acc.y = y; acc.hits = 0; acc.outHits = 0;
      exact_pair(p, s, frame, lastSubstep, slot, x, y, gme.z, fw, t, xt, yt, gt[j].z, ft, true, acc);
      if (acc.hits) {
        coll[1 + 2 * (size_t)base] = (int32_t)(d.GID ? d.GID[i] : i);  // :556-557
        coll[2 + 2 * (size_t)base] = (int32_t)__float_as_uint(s.SA[2 * (size_t)t + 1].w);
        base++;
        found++;
      }
    }
  }
  }
  __syncthreads();                                              // s_warp is rewritten by the next tile
  }
}

// ---- slabs (multi-GPU): one exchange per frame -------------------------------------------------
// Ownership follows position: a context owns the entities whose clamped cell row lies in
// [slabBegin, slabEnd).  slabHalo rows beyond each cut are replicas; they are recomputed
// redundantly during the frame and thrown away at its end.  After the frame each context
//   1. packs its authoritative (owned at frame start) entities that now lie within slabHalo
//      rows of a cut, or beyond it (migration), into 64-byte records for that neighbour,
//   2. drops every entity it did not own this frame (replicas) and lists the holes,
//   3. unpacks the neighbours' records into the holes (or past the table top).
struct __align__(16) SlabRec {
  uint32_t gid, meta;     // meta: flag byte | collisionCount << 8
  float ax, ay;
  float4 dp;              // x, y, px, py
  float4 at;              // maxVel, radius, visualRange, velocityAngle
  float4 v;               // vx, vy, speed, -
};
static_assert(sizeof(SlabRec) == 64, "slab record must be 64 bytes");

// device-resident state of the exchange: the host never has to read it between frames
struct SlabCounters {
  // words the pack / drop kernels update with atomics: a 128-byte line of their own, so that the
  // per-thread read of `top` below is not served by a line under atomic traffic
  uint32_t nLow, nHigh;     // records packed for the low / high neighbour this frame
  uint32_t nHoles;          // free slots found by the drop pass
  uint32_t owned;           // entities owned during the frame being packed
  uint32_t _hot[28];
  uint32_t top;             // slots in use (persistent)
  uint32_t overflow;        // sticky: bit0 exchange quota exceeded, bit1 entity table full
  uint32_t lastOwned, lastLow, lastHigh, lastFromLow, lastFromHigh;   // previous exchange, for reporting
  // cuts: `cur` decides ownership during a frame, `pend` (known to both sides of a cut before the
  // frame's exchange) selects what is packed and becomes `cur` when the exchange is applied
  int32_t curBegin, curEnd, pendBegin, pendEnd;
  uint32_t maxShift, hysteresisPct, minRows;     // dynamic balancing: 0 rows = static cuts
  uint32_t load;                                 // smoothed device time of this slab's frame kernels
  uint32_t cutMoves;                             // how many times one of my cuts moved
  uint32_t seq;                                  // exchanges completed (peer-to-peer transport: frame parity and arrival flag)
  uint32_t _pad[15];
};
static_assert(sizeof(SlabCounters) == 256, "two 128-byte lines");

__device__ __forceinline__ bool present_row(const GridDims& g, const ById& d, uint32_t i, int32_t& row) {
  const uint32_t f = d.F[i];
  const float4 p = d.DP[i];
  if (!(f & F_T_ACTIVE) || p.x != p.x || p.y != p.y) return false;
  int32_t col;
  cell_of(g, p.x, p.y, col, row);
  return true;
}

// Peer-to-peer transport (weed_slab_exchange_*): the pack kernel writes its records STRAIGHT into the
// neighbour's receive buffer over NVLink (a peer mapping: cudaIpcOpenMemHandle across processes,
// cudaDeviceEnablePeerAccess inside one), k_slab_publish then stores the header and, after a system
// fence, the arrival flag; the neighbour's k_slab_wait spins on that flag before its unpack.  Two
// buffers per side alternate by frame parity: a sender can only be one exchange ahead of its
// neighbour (it needs the neighbour's previous message to get there), so the buffer it writes was
// consumed two exchanges ago.  The table lives in device memory so that captured launches see it.
struct SlabXfer {
  SlabRec* send[2][2];     // [side: 0 low neighbour, 1 high][parity]: that neighbour's receive buffer (peer memory), or null
  SlabRec* recv[2][2];     // [side][parity]: my receive buffers
  uint32_t quota, _pad;
};
static constexpr uint32_t SLAB_OVF_QUOTA = 1u, SLAB_OVF_TABLE = 2u, SLAB_OVF_REACH = 4u, SLAB_OVF_TIMEOUT = 8u;

// Exchange buffers hold quota + 1 records; record 0 is a header whose `gid` word is the count.
// key[] still holds the cell of the frame-START position (ownership during the frame).
__global__ void __launch_bounds__(256)
k_slab_pack(GridDims g, ById d, const uint32_t* __restrict__ key, SlabRec* __restrict__ low,
            SlabRec* __restrict__ high, uint32_t quota, SlabCounters* sc, const SlabXfer* __restrict__ xf, int subSteps) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  if (xf) {                                     // peer-to-peer: my records go straight into the neighbours' buffers
    const uint32_t par = sc->seq & 1u;
    low = xf->send[0][par]; high = xf->send[1][par]; quota = xf->quota;
  }
  bool owned = false;
  if (i < sc->top) {
    const uint32_t k = key[i];
    if (k != KEY_INVALID) {
      const int32_t row0 = (int32_t)(k / (uint32_t)g.cols);
      owned = row0 >= sc->curBegin && row0 < sc->curEnd;          // authoritative here during the frame
    }
  }
  // one counter update per BLOCK: same-address atomics serialise at ~2 ns each, and one per warp
  // is still 30 000 of them per million entities
  __shared__ uint32_t s_owned[256 / 32];
  const uint32_t nOwned = __reduce_add_sync(0xffffffffu, (uint32_t)owned);
  if (lane == 0) s_owned[threadIdx.x >> 5] = nOwned;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < 256 / 32; w++) t += s_owned[w];
    if (t) atomicAdd(&sc->owned, t);
  }
  bool toLow = false, toHigh = false;
  int32_t row;
  if (owned && present_row(g, d, i, row)) {
    // bands around the cuts of the NEXT frame (both sides of a cut know them already)
    toLow = sc->pendBegin > 0 && row < sc->pendBegin + g.slabHalo;
    toHigh = sc->pendEnd < g.rows && row >= sc->pendEnd - g.slabHalo;
    // Guard of the halo depth: the halo is sized for the entities that were near a cut when the slabs
    // were planned.  An entity whose reach (rows of neighbours that must be replicated for it: S + 1
    // visual ranges for one that collides, two for an observer — trigger or no collider — whose capped
    // row must be complete) exceeds the halo must stay clear of the cuts; if one gets there the
    // partition would no longer reproduce the single-context result, so the slab says so.
    const uint32_t f = d.F[i];
    const double crd = ceil(dmul((double)d.AT[i].z, g.inv));
    const int32_t cr = crd > 0 ? (crd < (double)g.rows ? (int32_t)crd : g.rows) : 0;      // NaN: empty window
    const bool observer = (f & (F_C_ACTIVE | F_TRIGGER)) != F_C_ACTIVE;
    const int32_t need = observer ? 2 * cr : (subSteps + 1) * cr;
    const int32_t near = observer ? cr : need;
    if (need > g.slabHalo &&
        ((sc->pendBegin > 0 && row < sc->pendBegin + near) || (sc->pendEnd < g.rows && row >= sc->pendEnd - near)))
      atomicOr(&sc->overflow, SLAB_OVF_REACH);
  }
  // positions in the two exchange buffers: ballot inside the warp, shared memory across the
  // warps, ONE atomic per block and direction
  __shared__ uint32_t s_low[256 / 32], s_high[256 / 32], s_base[2];
  const uint32_t mLow = __ballot_sync(0xffffffffu, toLow), mHigh = __ballot_sync(0xffffffffu, toHigh);
  const uint32_t warp = threadIdx.x >> 5;
  if (lane == 0) { s_low[warp] = (uint32_t)__popc(mLow); s_high[warp] = (uint32_t)__popc(mHigh); }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t tl = 0, th = 0;
    for (int w = 0; w < 256 / 32; w++) { tl += s_low[w]; th += s_high[w]; }
    s_base[0] = tl ? atomicAdd(&sc->nLow, tl) : 0u;
    s_base[1] = th ? atomicAdd(&sc->nHigh, th) : 0u;
  }
  __syncthreads();
  uint32_t baseLow = s_base[0], baseHigh = s_base[1];
  for (uint32_t w = 0; w < warp; w++) { baseLow += s_low[w]; baseHigh += s_high[w]; }
  if (!toLow && !toHigh) return;
  SlabRec r;
  r.gid = d.GID[i];
  r.meta = (uint32_t)d.F[i] | ((uint32_t)d.CC[i] << 8);
  const float2 a = d.ACC[i];
  r.ax = a.x; r.ay = a.y;
  r.dp = d.DP[i]; r.at = d.AT[i]; r.v = d.V[i];
  const uint32_t below = (1u << lane) - 1u;
  if (toLow && low) {
    const uint32_t p = baseLow + (uint32_t)__popc(mLow & below);
    if (p < quota) low[1 + p] = r;
  }
  if (toHigh && high) {
    const uint32_t p = baseHigh + (uint32_t)__popc(mHigh & below);
    if (p < quota) high[1 + p] = r;
  }
}

// Header record (index 0) of a message: gid = record count, meta = the sender's smoothed frame time,
// ax / ay = bit patterns of the sender's cuts for the next frame.
__global__ void k_slab_headers(SlabRec* __restrict__ low, SlabRec* __restrict__ high, uint32_t quota, SlabCounters* sc,
                               const Counters* __restrict__ ctr, const SlabXfer* __restrict__ xf) {
  if (threadIdx.x != 0) return;
  if (xf) {
    const uint32_t par = sc->seq & 1u;
    low = xf->send[0][par]; high = xf->send[1][par]; quota = xf->quota;
  }
  if (sc->nLow > quota || sc->nHigh > quota) sc->overflow |= SLAB_OVF_QUOTA;
  const uint32_t ns = ctr->frameNs;
  sc->load = sc->load ? (uint32_t)((3ull * sc->load + ns) / 4) : ns;
  SlabRec* hdr[2] = {low, high};
  const uint32_t cnt[2] = {min(sc->nLow, quota), min(sc->nHigh, quota)};
  for (int k = 0; k < 2; k++) {
    if (!hdr[k]) continue;
    hdr[k][0].gid = cnt[k];
    hdr[k][0].meta = sc->load;
    hdr[k][0].ax = __int_as_float(sc->pendBegin);
    hdr[k][0].ay = __int_as_float(sc->pendEnd);
  }
  if (xf) {
    // the records were written by the previous kernel, the header just now: make all of it visible to
    // the peer before the arrival flag (the exchange number, never 0) appears
    __threadfence_system();
    for (int k = 0; k < 2; k++)
      if (hdr[k]) *reinterpret_cast<volatile uint32_t*>(&hdr[k][0].dp.x) = sc->seq + 1u;
    __threadfence_system();
  }
}

// Spins until both neighbours' messages of this exchange have arrived (their flag shows the exchange
// number).  Bounded: about two seconds of %globaltimer, then the slab reports SLAB_OVF_TIMEOUT
// instead of hanging the device.
__global__ void k_slab_wait(const SlabXfer* __restrict__ xf, SlabCounters* sc) {
  if (threadIdx.x != 0) return;
  const uint32_t par = sc->seq & 1u, want = sc->seq + 1u;
  const unsigned long long t0 = global_timer_ns();
  for (int side = 0; side < 2; side++) {
    if (!xf->send[side][par]) continue;                        // no neighbour on that side
    const volatile uint32_t* flag = reinterpret_cast<const volatile uint32_t*>(&xf->recv[side][par][0].dp.x);
    while (*flag != want) {
      if (global_timer_ns() - t0 > 2000000000ull) { sc->overflow |= SLAB_OVF_TIMEOUT; return; }
      __nanosleep(200);
    }
  }
  __threadfence_system();
}

__global__ void __launch_bounds__(256)
k_slab_drop(GridDims g, ById d, const uint32_t* __restrict__ key, uint32_t* __restrict__ holes, SlabCounters* sc) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  bool drop = false;
  if (i < sc->top) {
    const uint32_t k = key[i];
    bool keep;
    if (k != KEY_INVALID) {
      const int32_t row0 = (int32_t)(k / (uint32_t)g.cols);
      keep = row0 >= sc->curBegin && row0 < sc->curEnd;
    } else {
      keep = (d.F[i] & F_T_ACTIVE) != 0;    // active but never in the grid (NaN position): stays where it is
    }
    drop = !keep;
  }
  const uint32_t m = __ballot_sync(0xffffffffu, drop);
  if (!m) return;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(&sc->nHoles, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (drop) {
    d.F[i] = 0;
    holes[base + (uint32_t)__popc(m & ((1u << lane) - 1u))] = i;
  }
}

// first `quota` threads: records from the low neighbour, next `quota`: from the high one
__global__ void __launch_bounds__(256)
k_slab_unpack(ById d, const SlabRec* __restrict__ fromLow, const SlabRec* __restrict__ fromHigh, uint32_t quota,
              const uint32_t* __restrict__ holes, uint32_t capacity, SlabCounters* sc, const SlabXfer* __restrict__ xf) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (xf) {
    const uint32_t par = sc->seq & 1u;
    fromLow = xf->send[0][par] ? xf->recv[0][par] : nullptr;
    fromHigh = xf->send[1][par] ? xf->recv[1][par] : nullptr;
    quota = xf->quota;
  }
  const uint32_t nL = fromLow ? min(fromLow[0].gid, quota) : 0u;
  const uint32_t nH = fromHigh ? min(fromHigh[0].gid, quota) : 0u;
  const SlabRec* src;
  uint32_t q;
  if (k < quota) { if (k >= nL) return; src = fromLow + 1 + k; q = k; }
  else { const uint32_t kk = k - quota; if (kk >= nH) return; src = fromHigh + 1 + kk; q = nL + kk; }
  const uint32_t nHoles = sc->nHoles, top = sc->top;
  const uint32_t i = q < nHoles ? holes[q] : top + (q - nHoles);
  if (i >= capacity) return;                                   // flagged by k_slab_finish
  const SlabRec r = *src;
  d.GID[i] = r.gid;
  d.F[i] = (uint8_t)(r.meta & 0xFFu);
  d.CC[i] = (uint8_t)((r.meta >> 8) & 0xFFu);
  d.ACC[i] = make_float2(r.ax, r.ay);
  d.DP[i] = r.dp; d.AT[i] = r.at; d.V[i] = r.v;
}

// One cut between a low slab (load l, rows hl) and a high slab (load h, rows hh): which way it
// moves for the frame after next.  Both neighbours evaluate this on the SAME numbers (their own
// and the other's header), so they always agree.
__device__ __forceinline__ int32_t slab_cut_shift(const SlabCounters* sc, uint32_t l, uint32_t h, int32_t hl, int32_t hh) {
  if (!sc->maxShift || !l || !h) return 0;
  const unsigned long long L = l, H = h, pct = 100 + sc->hysteresisPct;
  const int32_t room = (int32_t)sc->minRows + 2 * (int32_t)sc->maxShift;     // the other cut of a slab may move too
  const bool far = (L > H ? (L - H) : (H - L)) * 100ull > (L + H) * (unsigned long long)(4 * sc->hysteresisPct) / 2;
  const int32_t step = far ? (int32_t)sc->maxShift : 1;
  if (L * 100 > H * pct && hl >= room) return -step;         // the low slab is slower: it shrinks
  if (H * 100 > L * pct && hh >= room) return step;          // the high slab is slower: it shrinks
  return 0;
}

__global__ void k_slab_finish(const SlabRec* __restrict__ fromLow, const SlabRec* __restrict__ fromHigh, uint32_t quota,
                              uint32_t capacity, SlabCounters* sc, const SlabXfer* __restrict__ xf) {
  if (threadIdx.x != 0) return;
  if (xf) {
    const uint32_t par = sc->seq & 1u;
    fromLow = xf->send[0][par] ? xf->recv[0][par] : nullptr;
    fromHigh = xf->send[1][par] ? xf->recv[1][par] : nullptr;
    quota = xf->quota;
  }
  sc->seq++;
  const uint32_t nL = fromLow ? min(fromLow[0].gid, quota) : 0u;
  const uint32_t nH = fromHigh ? min(fromHigh[0].gid, quota) : 0u;
  const unsigned long long total = (unsigned long long)nL + nH;
  unsigned long long nt = sc->top;
  if (total > sc->nHoles) nt += total - sc->nHoles;
  if (nt > capacity) { sc->overflow |= SLAB_OVF_TABLE; nt = capacity; }
  sc->top = (uint32_t)nt;
  sc->lastOwned = sc->owned; sc->lastLow = sc->nLow; sc->lastHigh = sc->nHigh;
  sc->lastFromLow = nL; sc->lastFromHigh = nH;
  sc->nLow = 0; sc->nHigh = 0; sc->nHoles = 0; sc->owned = 0;
  // the cuts the exchange just served become the ownership of the next frame; the ones after
  // that follow from the loads both sides of each cut now know
  sc->curBegin = sc->pendBegin; sc->curEnd = sc->pendEnd;
  const int32_t mine = sc->curEnd - sc->curBegin;
  int32_t db = 0, de = 0;
  if (fromLow) db = slab_cut_shift(sc, fromLow[0].meta, sc->load, __float_as_int(fromLow[0].ay) - __float_as_int(fromLow[0].ax), mine);
  if (fromHigh) de = slab_cut_shift(sc, sc->load, fromHigh[0].meta, mine, __float_as_int(fromHigh[0].ay) - __float_as_int(fromHigh[0].ax));
  sc->pendBegin = sc->curBegin + db;
  sc->pendEnd = sc->curEnd + de;
  sc->cutMoves += (db != 0) + (de != 0);
}

// ---- host <-> device column plumbing ---------------------------------------------------------
// Host columns (the SAB SoA columns) are staged verbatim and packed into the by-id records.
struct Staging {
  const uint8_t* t_active; const float* x; const float* y;
  const uint8_t* rb_active; const uint8_t* rb_static;
  const float* vx; const float* vy; const float* ax; const float* ay; const float* px; const float* py;
  const float* maxVel; const float* velAngle; const float* speed; const uint8_t* collCnt;
  const uint8_t* c_active; const float* radius; const uint8_t* isTrigger; const float* visRange;
  const uint8_t* entityType;
};

__global__ void __launch_bounds__(256) k_pack(uint32_t N, uint32_t mask, Staging st, ById d) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint32_t FLAGCOLS = (1u << 0) | (1u << 3) | (1u << 4) | (1u << 15) | (1u << 17);
  if (mask & FLAGCOLS) {
    uint32_t f = d.F[i];
    if (mask & (1u << 0))  f = (f & ~F_T_ACTIVE)  | (st.t_active[i]  ? F_T_ACTIVE : 0u);
    if (mask & (1u << 3))  f = (f & ~F_RB_ACTIVE) | (st.rb_active[i] ? F_RB_ACTIVE : 0u);
    if (mask & (1u << 4))  f = (f & ~F_STATIC)    | (st.rb_static[i] ? F_STATIC : 0u);
    if (mask & (1u << 15)) f = (f & ~F_C_ACTIVE)  | (st.c_active[i]  ? F_C_ACTIVE : 0u);
    if (mask & (1u << 17)) f = (f & ~F_TRIGGER)   | (st.isTrigger[i] ? F_TRIGGER : 0u);
    d.F[i] = (uint8_t)f;
  }
  const uint32_t DPCOLS = (1u << 1) | (1u << 2) | (1u << 9) | (1u << 10);
  if (mask & DPCOLS) {
    float4 v = ((mask & DPCOLS) == DPCOLS) ? make_float4(0, 0, 0, 0) : d.DP[i];
    if (mask & (1u << 1))  v.x = st.x[i];
    if (mask & (1u << 2))  v.y = st.y[i];
    if (mask & (1u << 9))  v.z = st.px[i];
    if (mask & (1u << 10)) v.w = st.py[i];
    d.DP[i] = v;
  }
  const uint32_t ACOLS = (1u << 7) | (1u << 8);
  if (mask & ACOLS) {
    float2 v = ((mask & ACOLS) == ACOLS) ? make_float2(0, 0) : d.ACC[i];
    if (mask & (1u << 7)) v.x = st.ax[i];
    if (mask & (1u << 8)) v.y = st.ay[i];
    d.ACC[i] = v;
  }
  const uint32_t ATCOLS = (1u << 11) | (1u << 16) | (1u << 18) | (1u << 12);
  if (mask & ATCOLS) {
    float4 v = ((mask & ATCOLS) == ATCOLS) ? make_float4(0, 0, 0, 0) : d.AT[i];
    if (mask & (1u << 11)) v.x = st.maxVel[i];
    if (mask & (1u << 16)) v.y = st.radius[i];
    if (mask & (1u << 18)) v.z = st.visRange[i];
    if (mask & (1u << 12)) v.w = st.velAngle[i];
    d.AT[i] = v;
  }
  const uint32_t VCOLS = (1u << 5) | (1u << 6) | (1u << 13);
  if (mask & VCOLS) {
    float4 v = d.V[i];
    if (mask & (1u << 5))  v.x = st.vx[i];
    if (mask & (1u << 6))  v.y = st.vy[i];
    if (mask & (1u << 13)) v.z = st.speed[i];
    d.V[i] = v;
  }
  if (mask & (1u << 14)) d.CC[i] = st.collCnt[i];
  if (mask & (1u << 19)) d.ET[i] = st.entityType[i];
}

struct StagingOut {
  uint8_t* t_active; float* x; float* y;
  uint8_t* rb_active; uint8_t* rb_static;
  float* vx; float* vy; float* ax; float* ay; float* px; float* py;
  float* maxVel; float* velAngle; float* speed; uint8_t* collCnt;
  uint8_t* c_active; float* radius; uint8_t* isTrigger; float* visRange;
  uint8_t* entityType;
};

__global__ void __launch_bounds__(256) k_unpack(uint32_t N, uint32_t mask, StagingOut st, ById d) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint32_t FLAGCOLS = (1u << 0) | (1u << 3) | (1u << 4) | (1u << 15) | (1u << 17);
  if (mask & FLAGCOLS) {
    const uint32_t f = d.F[i];
    if (mask & (1u << 0))  st.t_active[i]  = (f & F_T_ACTIVE) ? 1 : 0;
    if (mask & (1u << 3))  st.rb_active[i] = (f & F_RB_ACTIVE) ? 1 : 0;
    if (mask & (1u << 4))  st.rb_static[i] = (f & F_STATIC) ? 1 : 0;
    if (mask & (1u << 15)) st.c_active[i]  = (f & F_C_ACTIVE) ? 1 : 0;
    if (mask & (1u << 17)) st.isTrigger[i] = (f & F_TRIGGER) ? 1 : 0;
  }
  if (mask & ((1u << 1) | (1u << 2) | (1u << 9) | (1u << 10))) {
    const float4 v = d.DP[i];
    if (mask & (1u << 1))  st.x[i] = v.x;
    if (mask & (1u << 2))  st.y[i] = v.y;
    if (mask & (1u << 9))  st.px[i] = v.z;
    if (mask & (1u << 10)) st.py[i] = v.w;
  }
  if (mask & ((1u << 7) | (1u << 8))) {
    const float2 v = d.ACC[i];
    if (mask & (1u << 7)) st.ax[i] = v.x;
    if (mask & (1u << 8)) st.ay[i] = v.y;
  }
  if (mask & ((1u << 11) | (1u << 16) | (1u << 18) | (1u << 12))) {
    const float4 v = d.AT[i];
    if (mask & (1u << 11)) st.maxVel[i] = v.x;
    if (mask & (1u << 16)) st.radius[i] = v.y;
    if (mask & (1u << 18)) st.visRange[i] = v.z;
    if (mask & (1u << 12)) st.velAngle[i] = v.w;
  }
  if (mask & ((1u << 5) | (1u << 6) | (1u << 13))) {
    const float4 v = d.V[i];
    if (mask & (1u << 5))  st.vx[i] = v.x;
    if (mask & (1u << 6))  st.vy[i] = v.y;
    if (mask & (1u << 13)) st.speed[i] = v.z;
  }
  if (mask & (1u << 14)) st.collCnt[i] = d.CC[i];
  if (mask & (1u << 19)) st.entityType[i] = d.ET[i];
}

// ---- system: boids flocking tick ------------------------------------------------
// demos/predators/boid.js:137-240 + :318-341, one thread per entity, evaluation order and
// rounding of the JavaScript: accumulators are binary64, every `rbAX[i] += ...` rounds to
// float32.  Reads the API rows where the spatial pass left them.
// Per-class parameters: Boid (boid.js:64-69), Prey (prey.js:37, 55-60) and Predator
// (predator.js:43, 57-62) differ only in these numbers and in their processNeighbor hook.
static constexpr uint32_t FLOCK_ANY_TYPE = 0xFFFFFFFFu;
static constexpr int FLOCK_MAX_CLASSES = 8;
struct FlockClass {
  uint32_t type, role, other, _pad;       // role: 0 boid, 1 prey (flees `other`), 2 predator (hunts `other`)
  double prScale, centering, avoid, matching, turn, margin, roleFactor;
};
struct FlockParams {
  FlockClass cls[FLOCK_MAX_CLASSES];
  uint32_t nClasses, mouseType, mouseDown, _pad;
  double dtRatio;
};

__global__ void __launch_bounds__(128)
k_system_flock(GridDims g, FlockParams fp, ById d, RowView rows,
               const float* __restrict__ protectedRange) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.N || i == 0) return;                   // index 0 is the Mouse: its tick() is empty
  if (!(d.F[i] & F_T_ACTIVE)) return;
  const uint32_t myType = d.ET[i];
  int ci = -1;
  for (uint32_t c = 0; c < fp.nClasses; c++)
    if (fp.cls[c].type == myType || fp.cls[c].type == FLOCK_ANY_TYPE) { ci = (int)c; break; }
  if (ci < 0) return;                               // not a boid: some other tick()
  const FlockClass& k = fp.cls[ci];
  const double dt = fp.dtRatio;
  uint32_t slot;
  const int32_t cnt = rows.count(i, slot);
  const float4 me = d.DP[i];
  const double myX = me.x, myY = me.y;
  float2 acc = d.ACC[i];
  if (cnt > 0) {                                                                 // boid.js:138
    const float pr = protectedRange ? protectedRange[i] : fround(dmul((double)d.AT[i].y, k.prScale));
    const double pr2 = dmul((double)pr, (double)pr);
    double cx = 0, cy = 0, avx = 0, avy = 0, sx = 0, sy = 0;
    double fleeX = 0, fleeY = 0, closest2 = CUDART_INF;
    uint32_t same = 0, predators = 0;
    int32_t closest = -1;
    for (int32_t n = 0; n < cnt; n++) {
      const int32_t j = rows.id(slot, n);
      const uint32_t nt = d.ET[j];
      if (nt == fp.mouseType) continue;                                          // :179-180
      const double d2 = (double)rows.d2(slot, n);
      const float4 pj = d.DP[j];
      const double dx = dsub((double)pj.x, myX), dy = dsub((double)pj.y, myY);
      if (d2 < pr2 && d2 > 0) {                                                  // :192-196
        sx = dsub(sx, ddiv(dx, d2));
        sy = dsub(sy, ddiv(dy, d2));
        continue;
      }
      if (nt == myType) {                                                        // :199-206
        const float4 vj = d.V[j];
        cx = dadd(cx, (double)pj.x); cy = dadd(cy, (double)pj.y);
        avx = dadd(avx, (double)vj.x); avy = dadd(avy, (double)vj.y);
        same++;
      }
      // processNeighbor hooks (:209-217)
      if (k.role == 1) {                                                         // prey.js:154-169
        if (nt == k.other && d2 > 0) {
          fleeX = dadd(fleeX, ddiv(-dx, d2));
          fleeY = dadd(fleeY, ddiv(-dy, d2));
          predators++;
        }
      } else if (k.role == 2) {                                                  // predator.js:172-187
        if (nt == k.other && d2 < closest2) { closest2 = d2; closest = j; }
      }
    }
    if (same) {                                                                  // :221-232
      const float4 vi = d.V[i];
      cx = ddiv(cx, (double)same); cy = ddiv(cy, (double)same);
      acc.x = fround(dadd((double)acc.x, dmul(dmul(dsub(cx, myX), k.centering), dt)));
      acc.y = fround(dadd((double)acc.y, dmul(dmul(dsub(cy, myY), k.centering), dt)));
      avx = ddiv(avx, (double)same); avy = ddiv(avy, (double)same);
      acc.x = fround(dadd((double)acc.x, dmul(dmul(dsub(avx, (double)vi.x), k.matching), dt)));
      acc.y = fround(dadd((double)acc.y, dmul(dmul(dsub(avy, (double)vi.y), k.matching), dt)));
    }
    acc.x = fround(dadd((double)acc.x, dmul(dmul(sx, k.avoid), dt)));             // :235-236
    acc.y = fround(dadd((double)acc.y, dmul(dmul(sy, k.avoid), dt)));
    if (k.role == 1 && predators) {                                              // prey.js:176-189
      acc.x = fround(dadd((double)acc.x, dmul(dmul(fleeX, k.roleFactor), dt)));
      acc.y = fround(dadd((double)acc.y, dmul(dmul(fleeY, k.roleFactor), dt)));
    }
    if (k.role == 2 && closest >= 0) {                                           // predator.js:195-215
      const float4 pp = d.DP[closest];
      const double dx = dsub((double)pp.x, myX), dy = dsub((double)pp.y, myY);
      const double dist = __dsqrt_rn(closest2);
      if (dist > 0) {
        acc.x = fround(dadd((double)acc.x, dmul(dmul(ddiv(dx, dist), k.roleFactor), dt)));
        acc.y = fround(dadd((double)acc.y, dmul(dmul(ddiv(dy, dist), k.roleFactor), dt)));
      }
    }
  }
  if (fp.mouseDown) {                                                            // boid.js:281-316
    for (int32_t n = 0; n < cnt; n++) {
      if (rows.id(slot, n) != 0) continue;                                        // the Mouse is entity 0
      const double d2 = (double)rows.d2(slot, n);
      if (d2 != d2 || d2 == 0) break;                                            // `!dist2`
      const float4 pm = d.DP[0];
      const double dx = dsub((double)pm.x, myX), dy = dsub((double)pm.y, myY);
      acc.x = fround(dsub((double)acc.x, dmul(dmul(ddiv(dx, d2), 1000.0), dt)));
      acc.y = fround(dsub((double)acc.y, dmul(dmul(ddiv(dy, d2), 1000.0), dt)));
      break;
    }
  }
  const double turn = dmul(k.turn, dt);                                          // :334-340
  if (myX < k.margin) acc.x = fround(dadd((double)acc.x, turn));
  if (myX > dsub(g.worldW, k.margin)) acc.x = fround(dsub((double)acc.x, turn));
  if (myY < k.margin) acc.y = fround(dadd((double)acc.y, turn));
  if (myY > dsub(g.worldH, k.margin)) acc.y = fround(dsub((double)acc.y, turn));
  d.ACC[i] = acc;
}

// ---- API rows for the host: slot-major planes -> the reference's rows (gameEngine.js:552-559) ------
// One warp per entity of [first, first + count): header and the `count` entries of its row go to the
// mirror in host layout (stride 1 + maxNeighbors).  Entities that are not in the grid keep whatever
// their row held (spatial_worker.js:148,153 never rewrites it), and so do the words past 1 + count.
__global__ void __launch_bounds__(256)
k_rows_gather(RowView rows, uint32_t first, uint32_t count, uint32_t hostStride, int32_t* __restrict__ mnd,
              float* __restrict__ mdd) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= count) return;
  const uint32_t i = first + w;
  const uint32_t slot = rows.slotOf[i];
  if (slot == SLOT_NONE) return;
  const int32_t cnt = (int32_t)(rows.ncnt[slot] & 0xFFFFu);
  const size_t o = (size_t)i * hostStride;
  if (lane == 0) { mnd[o] = cnt; mdd[o] = (float)cnt; }                         // :274-275
  for (int32_t k = (int32_t)lane; k < cnt; k += 32) { mnd[o + 1 + k] = rows.id(slot, k); mdd[o + 1 + k] = rows.d2(slot, k); }
}

// ---- statistics (only when the host asks) -------------------------------------------------
__global__ void __launch_bounds__(256)
k_stats(GridDims g, const uint32_t* __restrict__ NCNT, const uint32_t* __restrict__ cellStart, Counters* ctr) {
  const uint32_t A = cellStart[g.cells];
  unsigned long long sum = 0; uint32_t capped = 0;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < A; e += gridDim.x * blockDim.x) {
    const uint32_t n = NCNT[e] & 0xFFFFu;
    sum += n; capped += (n >= g.M);
  }
  for (int o = 16; o; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    capped += __shfl_xor_sync(0xffffffffu, capped, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&ctr->neighborsTotal, sum);
    atomicAdd(&ctr->cappedRows, capped);
  }
}

}  // namespace weed
