// weed_kernels.cuh — the per-frame kernels (sm_100a).  See DESIGN.md for the data flow.
//
//   id order   k_cell_key      K1  cell key + arrival rank (warp-aggregated atomics)
//   cells      k_cell_scan     K2  exclusive scan, single pass, decoupled look-back
//   id order   k_scatter_ids   K3a ids into their cell segment (arrival order)
//   id order   k_build_slots   K3b stable position inside the cell (ascending id), Verlet
//                                  integration (K5) + derived speed/angle fused, slot records
//   slot order k_neighbors     K4  capped ordered gather, 8 lanes per entity
//   slot order k_explicit_capped K4b explicit pairs caused by capped partners (rare)
//   slot order k_substep<LAST> K6  bounds + circle-circle correction, J-order, 8 lanes/entity
//   id order   k_writeback     WB  gather results by id, look-back scan of pair counts,
//                                  collisionData emission (K7)
#pragma once
#include <cooperative_groups.h>

#include "weed_device.cuh"

namespace weed {
namespace cg = cooperative_groups;

static constexpr int TILE_W = 8;            // lanes cooperating on one entity in K4 / K6
static constexpr int SCAN_THREADS = 512;
static constexpr int SCAN_ITEMS = 4;
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
    ctr->explicitPairs = 0;
    ctr->explicitOverflowFrame = 0;
    ctr->maxCellFrame = 0;
  }
}
__global__ void k_physics_end(Counters* ctr) {
  if (threadIdx.x == 0) { ctr->frame++; ctr->frames++; }
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
  // warp-aggregated counting: one atomic per distinct cell per warp; lanes keep id order
  const uint32_t amask = __activemask();
  const uint32_t peers = __match_any_sync(amask, cell);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t leader = __ffs(peers) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(&cellCount[cell], __popc(peers));
  base = __shfl_sync(peers, base, leader);
  key[i] = cell;
  rank[i] = base + __popc(peers & ((1u << lane) - 1));
}

// ---- K2: exclusive scan of the cell histogram, also clears it for the next frame ---------
__global__ void __launch_bounds__(SCAN_THREADS)
k_cell_scan(uint32_t* __restrict__ cellCount, uint32_t* __restrict__ cellStart, uint32_t numTiles,
            unsigned long long* status, Counters* ctr) {
  __shared__ uint32_t s_tile, s_excl, s_warp[SCAN_THREADS / 32], s_max[SCAN_THREADS / 32];
  if (threadIdx.x == 0) s_tile = atomicAdd(&ctr->scanTile, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t epoch = ctr->epoch;
  const size_t i0 = (size_t)tile * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  uint4 c = *reinterpret_cast<const uint4*>(cellCount + i0);   // arrays are padded to whole tiles
  *reinterpret_cast<uint4*>(cellCount + i0) = make_uint4(0, 0, 0, 0);
  const uint32_t tsum = c.x + c.y + c.z + c.w;
  uint32_t tmax = max(max(c.x, c.y), max(c.z, c.w));
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
  uint4 o4;
  o4.x = e; e += c.x; o4.y = e; e += c.y; o4.z = e; e += c.z; o4.w = e;
  *reinterpret_cast<uint4*>(cellStart + i0) = o4;
}

// ---- K3a: ids into their cell segment, arrival order ------------------------------------
__global__ void __launch_bounds__(256)
k_scatter_ids(uint32_t N, const uint32_t* __restrict__ key, const uint32_t* __restrict__ rank,
              const uint32_t* __restrict__ cellStart, uint32_t* __restrict__ arrIds) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint32_t k = key[i];
  if (k == KEY_INVALID) return;
  arrIds[cellStart[k] + rank[i]] = i;
}

// ---- K3b + K5: slot records, Verlet integration, derived properties ----------------------
// moveBallsVerlet (physics_worker.js:264-315), collisionCount reset (:174-177) and
// updateDerivedProperties (:591-603; it only reads the vx,vy stored by the integration, so it
// commutes with the constraint substeps).  INTEGRATE=false builds the slot records of the
// unchanged state (weed_spatial on its own).
struct ById {
  float4* DP;        // x, y, px, py
  float2* ACC;       // ax, ay
  float4* AT;        // maxVel, radius, visualRange, velocityAngle
  float4* V;         // vx, vy, speed, -
  uint8_t* F;        // flag bits
  uint8_t* CC;       // collisionCount
};
struct BySlot {
  float2* QXY;       // position at grid-build time (query position)
  float* QVR;        // visualRange
  uint32_t* SID;     // entity id
  float4* G0;        // x, y, radius, flagword   (substep ping)
  float4* G1;        //                           (substep pong)
  float2* PXY;       // px, py
  uint32_t* NCNT;    // neighbor count
  uint32_t* NS;      // internal rows [slot][Mpad]
  uint32_t* XCNT;    // explicit incoming count
  uint32_t* XR;      // explicit incoming rows [slot][xcap]
  OutRec* OUT;       // last-substep result
};

template <bool INTEGRATE>
__global__ void __launch_bounds__(256)
k_build_slots(GridDims g, const Params* __restrict__ pp, int subSteps, bool afterSpatial, ById d, BySlot s,
              const uint32_t* __restrict__ key, const uint32_t* __restrict__ cellStart,
              const uint32_t* __restrict__ arrIds, uint32_t* __restrict__ slotOf) {
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
      float4 v;
      v.x = fround(ddiv(ddx, p.dtRatio));                                      // :309-310
      v.y = fround(ddiv(ddy, p.dtRatio));
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
  const uint32_t s0 = cellStart[k], s1 = cellStart[k + 1];
  uint32_t r = 0;
  for (uint32_t t = s0; t < s1; t++) r += arrIds[t] < i;
  const uint32_t slot = s0 + r;
  slotOf[i] = slot;
  s.QXY[slot] = make_float2(x0, y0);
  s.QVR[slot] = at.z;
  s.SID[slot] = i;
  uint32_t keep = 0;
  if (afterSpatial) keep = __float_as_uint(s.G0[slot].w) & F_CAPPED;   // rows of this frame already exist
  else s.XCNT[slot] = 0;
  s.G0[slot] = make_float4(dp.x, dp.y, at.y, __uint_as_float(f | keep | (cc << F_CC_SHIFT)));
  s.PXY[slot] = make_float2(dp.z, dp.w);
}

// ---- K4: capped, ordered neighbor gather (spatial_worker.js:195-277) ---------------------
// 8 lanes per entity.  Candidates of one window row are ONE contiguous slot range because
// slots are sorted by (cell, id) and cells of a grid row are consecutive; the reference's scan
// order (rows, then columns, then list order) is therefore ascending slot order, and an
// ordered ballot compaction reproduces the row content and the cap exactly.
__device__ __forceinline__ void explicit_append(const GridDims& g, BySlot& s, Counters* ctr, uint32_t dstSlot,
                                                uint32_t srcSlot) {
  const uint32_t pos = atomicAdd(&s.XCNT[dstSlot], 1u);
  if (pos < g.xcap) s.XR[(size_t)dstSlot * g.xcap + pos] = srcSlot;
  else atomicExch(&ctr->explicitOverflowFrame, 1u);
  atomicAdd(&ctr->explicitPairs, 1u);
}

template <bool WRITE_ROWS>
__global__ void __launch_bounds__(256)
k_neighbors(GridDims g, BySlot s, const uint32_t* __restrict__ cellStart, int32_t* __restrict__ nd,
            float* __restrict__ dd, Counters* ctr) {
  cg::thread_block_tile<TILE_W> tile = cg::tiled_partition<TILE_W>(cg::this_thread_block());
  const uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) / TILE_W;
  const uint32_t A = cellStart[g.cells];
  if (e >= A) return;
  const uint32_t lane = tile.thread_rank();
  const float2 q = s.QXY[e];
  const float vr = s.QVR[e];
  const uint32_t id = s.SID[e];
  const double myX = q.x, myY = q.y;
  const double vrSq = dmul((double)vr, (double)vr);
  int32_t myCol, myRow;
  cell_of(g, q.x, q.y, myCol, myRow);                 // my clamped cell (for partners' windows)
  const size_t rowBase = (size_t)id * (1 + (size_t)g.M);
  const size_t nsBase = (size_t)e * g.Mpad;
  const uint32_t M = g.M;
  uint32_t n = 0;
  Window w;
  if (M > 0 && query_window(g, q.x, q.y, vr, w)) {
    for (int32_t row = w.r0; row <= w.r1 && n < M; row++) {
      const uint32_t a = cellStart[(uint32_t)row * g.cols + w.c0];
      const uint32_t b = cellStart[(uint32_t)row * g.cols + w.c1 + 1];
      for (uint32_t t0 = a; t0 < b && n < M; t0 += TILE_W) {
        const uint32_t t = t0 + lane;
        bool acc = false;
        double d2 = 0;
        float2 c = make_float2(0.f, 0.f);
        if (t < b && t != e) {
          c = s.QXY[t];
          const double dX = dsub((double)c.x, myX);
          const double dY = dsub((double)c.y, myY);
          d2 = dadd(dmul(dX, dX), dmul(dY, dY));
          acc = d2 < vrSq && d2 > 0;               // :257
        }
        const uint32_t bits = tile.ballot(acc);
        const uint32_t pos = n + __popc(bits & ((1u << lane) - 1));
        if (acc && pos < M) {
          const uint32_t jid = s.SID[t];
          if (WRITE_ROWS) {
            nd[rowBase + 1 + pos] = (int32_t)jid;    // :259
            dd[rowBase + 1 + pos] = fround(d2);      // :260
          }
          // would partner t's own scan accept me (ignoring its cap)?  Its window must contain
          // my clamped cell and d2 < vr_t^2 (d2 is bitwise symmetric, and d2 > 0 holds).
          const float vrt = s.QVR[t];
          Window wt;
          bool back = d2 < dmul((double)vrt, (double)vrt) && query_window(g, c.x, c.y, vrt, wt) &&
                      myRow >= wt.r0 && myRow <= wt.r1 && myCol >= wt.c0 && myCol <= wt.c1;
          const bool out = jid > id;
          s.NS[nsBase + pos] = t | (out ? NS_OUT : 0u) | (back ? NS_BACK : 0u);
          // pair (id, jid) is in P but the partner cannot infer it from its own row
          if (out && !back) explicit_append(g, s, ctr, t, e);
        }
        n = min(M, n + (uint32_t)__popc(bits));
      }
    }
  }
  if (n >= M && M > 0) {
    // capped row: partners cannot trust their NS_BACK bit for me, so every pair I own
    // becomes explicit (the !back ones were appended above)
    for (uint32_t k = lane; k < n; k += TILE_W) {
      const uint32_t wd = s.NS[nsBase + k];
      if ((wd & NS_OUT) && (wd & NS_BACK)) explicit_append(g, s, ctr, wd & NS_SLOT_MASK, e);
    }
    if (lane == 0) {
      float4* gp = s.G0 + e;
      reinterpret_cast<uint32_t*>(gp)[3] |= F_CAPPED;
      ctr->anyCapped = 1;
    }
  }
  if (lane == 0) {
    if (WRITE_ROWS) {
      nd[rowBase] = (int32_t)n;                      // :274
      dd[rowBase] = (float)n;                        // :275
    }
    s.NCNT[e] = n;
  }
}

// ---- K4b: pairs owned by an uncapped row whose partner is capped ---------------------------
// (the partner may have been truncated before reaching me, so it will not infer the pair)
__global__ void __launch_bounds__(256)
k_explicit_capped(GridDims g, BySlot s, const uint32_t* __restrict__ cellStart, Counters* ctr) {
  if (!ctr->anyCapped) return;
  cg::thread_block_tile<TILE_W> tile = cg::tiled_partition<TILE_W>(cg::this_thread_block());
  const uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) / TILE_W;
  if (e >= cellStart[g.cells]) return;
  const uint32_t fw = __float_as_uint(s.G0[e].w);
  if (fw & F_CAPPED) return;                         // its pairs are already explicit
  const uint32_t n = s.NCNT[e];
  const size_t nsBase = (size_t)e * g.Mpad;
  for (uint32_t k = tile.thread_rank(); k < n; k += TILE_W) {
    const uint32_t wd = s.NS[nsBase + k];
    if ((wd & NS_OUT) && (wd & NS_BACK)) {
      const uint32_t t = wd & NS_SLOT_MASK;
      if (__float_as_uint(s.G0[t].w) & F_CAPPED) explicit_append(g, s, ctr, t, e);
    }
  }
}

// ---- K6: one constraint substep (physics_worker.js:323-395, 405-568), J-order ---------------
// Each entity (8 lanes) applies its own boundary pass, then evaluates every pair of P it
// belongs to on the start-of-sweep positions (partners' boundary pass re-applied on the fly)
// and accumulates its own corrections in ascending partner-slot order, rounding to float32
// after each one exactly like the reference's `x[i] += ...` on a Float32Array.
//
// Pair membership (P = {(i,j): i<j, j in row(i), both active colliders}):
//   - row entry with NS_OUT: I am i, the pair is mine.
//   - row entry without NS_OUT (partner id lower): the pair exists iff I am in the partner's
//     row.  That is inferred (NS_BACK, and neither row capped); every pair that cannot be
//     inferred was appended to my explicit list XR by its owner in K4 / K4b.
struct SubstepAcc { float x, y; uint32_t hits, outHits; };

__device__ __forceinline__ bool partner_pair(const GridDims& g, const Params& p, const BySlot& s,
                                             uint32_t frame, uint32_t substep, uint32_t e, float x, float y,
                                             float r, uint32_t fw, uint32_t t, float4 gt, bool iAmLower,
                                             double& mx, double& my, bool& moves) {
  const uint32_t ft = __float_as_uint(gt.w);
  float xt = gt.x, yt = gt.y;
  if ((ft & F_DYNAMIC_MASK) == F_DYNAMIC_VAL) apply_bounds_pos(g, gt.z, xt, yt);  // partner after ITS boundary pass
  PairMove m;
  if (iAmLower) m = pair_eval(p, frame, substep, s.SID, e, t, x, y, r, fw, xt, yt, gt.z, ft);
  else          m = pair_eval(p, frame, substep, s.SID, t, e, xt, yt, gt.z, ft, x, y, r, fw);
  if (iAmLower) { moves = m.moveI; mx = m.mx; my = m.my; }
  else          { moves = m.moveJ; mx = -m.mx; my = -m.my; }
  moves = moves && m.hit;
  return m.hit;
}

// sequential slow path for entities with explicit incoming pairs: lane 0 merges the row and
// the (sorted) explicit list in ascending slot order
__device__ __noinline__ void substep_explicit(const GridDims& g, const Params& p, const BySlot& s,
                                              const float4* __restrict__ Gin, uint32_t frame, uint32_t substep,
                                              uint32_t e, float x, float y, float r, uint32_t fw, uint32_t cnt,
                                              uint32_t xcnt, SubstepAcc& acc) {
  uint32_t* xr = s.XR + (size_t)e * g.xcap;
  for (uint32_t a = 1; a < xcnt; a++) {      // insertion sort (already sorted after the first substep)
    const uint32_t v = xr[a];
    uint32_t b = a;
    while (b > 0 && xr[b - 1] > v) { xr[b] = xr[b - 1]; b--; }
    xr[b] = v;
  }
  const uint32_t* ns = s.NS + (size_t)e * g.Mpad;
  const bool meCapped = (fw & F_CAPPED) != 0;
  uint32_t a = 0, b = 0;
  while (a < cnt || b < xcnt) {
    const uint32_t wa = a < cnt ? ns[a] : 0xFFFFFFFFu;
    const uint32_t ta = a < cnt ? (wa & NS_SLOT_MASK) : 0xFFFFFFFFu;
    const uint32_t tb = b < xcnt ? xr[b] : 0xFFFFFFFFu;
    uint32_t t; bool lower, inP;
    if (tb <= ta) {             // explicit incoming: partner is i, I am j
      t = tb; lower = false; inP = true; b++;
      if (ta == tb) a++;        // the same partner also sits in my row as a non-inferable entry
    } else {
      t = ta; a++;
      lower = (wa & NS_OUT) != 0;
      const float4 gtmp = Gin[t];
      inP = lower || ((wa & NS_BACK) && !meCapped && !(__float_as_uint(gtmp.w) & F_CAPPED));
    }
    const float4 gt = Gin[t];
    const uint32_t ft = __float_as_uint(gt.w);
    if (!inP || (ft & F_COLLIDER) != F_COLLIDER) continue;
    double mx, my; bool moves;
    if (partner_pair(g, p, s, frame, substep, e, x, y, r, fw, t, gt, lower, mx, my, moves)) {
      acc.hits++;
      if (lower) acc.outHits++;
      if (moves) {
        acc.x = fround(dadd((double)acc.x, mx));
        acc.y = fround(dadd((double)acc.y, my));
      }
    }
  }
}

template <bool LAST>
__global__ void __launch_bounds__(256)
k_substep(GridDims g, const Params* __restrict__ pp, BySlot s, const float4* __restrict__ Gin,
          float4* __restrict__ Gout, const uint32_t* __restrict__ cellStart, const Counters* __restrict__ ctr,
          uint32_t substep) {
  cg::thread_block_tile<TILE_W> tile = cg::tiled_partition<TILE_W>(cg::this_thread_block());
  const uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) / TILE_W;
  if (e >= cellStart[g.cells]) return;
  const uint32_t lane = tile.thread_rank();
  const Params p = *pp;
  const uint32_t frame = ctr->frame;
  const float4 gme = Gin[e];
  float2 pxy = s.PXY[e];
  float x = gme.x, y = gme.y;
  const float r = gme.z;
  const uint32_t fw = __float_as_uint(gme.w);
  if ((fw & F_DYNAMIC_MASK) == F_DYNAMIC_VAL) apply_bounds(g, p.boundaryElasticity, r, x, y, pxy.x, pxy.y);
  SubstepAcc acc; acc.x = x; acc.y = y; acc.hits = 0; acc.outHits = 0;
  if ((fw & F_COLLIDER) == F_COLLIDER) {                         // :430
    const uint32_t cnt = s.NCNT[e];
    const uint32_t xcnt = min(s.XCNT[e], g.xcap);
    if (xcnt == 0) {
      const bool meCapped = (fw & F_CAPPED) != 0;
      const uint32_t* ns = s.NS + (size_t)e * g.Mpad;
      for (uint32_t n0 = 0; n0 < cnt; n0 += TILE_W) {
        const uint32_t k = n0 + lane;
        bool hit = false, moves = false, lower = false;
        double mx = 0, my = 0;
        if (k < cnt) {
          const uint32_t wd = ns[k];
          const uint32_t t = wd & NS_SLOT_MASK;
          const float4 gt = Gin[t];
          const uint32_t ft = __float_as_uint(gt.w);
          lower = (wd & NS_OUT) != 0;
          const bool inP = lower || ((wd & NS_BACK) && !meCapped && !(ft & F_CAPPED));
          if (inP && (ft & F_COLLIDER) == F_COLLIDER)             // :441
            hit = partner_pair(g, p, s, frame, substep, e, x, y, r, fw, t, gt, lower, mx, my, moves);
        }
        const uint32_t hb = tile.ballot(hit);
        acc.hits += __popc(hb);
        acc.outHits += __popc(tile.ballot(hit && lower));
        uint32_t mb = tile.ballot(hit && moves);
        while (mb) {                                              // ascending slot order
          const int src = __ffs(mb) - 1;
          mb &= mb - 1;
          const double ax = tile.shfl(mx, src), ay = tile.shfl(my, src);
          acc.x = fround(dadd((double)acc.x, ax));
          acc.y = fround(dadd((double)acc.y, ay));
        }
      }
    } else {
      if (lane == 0) substep_explicit(g, p, s, Gin, frame, substep, e, x, y, r, fw, cnt, xcnt, acc);
      acc.x = tile.shfl(acc.x, 0); acc.y = tile.shfl(acc.y, 0);
      acc.hits = tile.shfl(acc.hits, 0); acc.outHits = tile.shfl(acc.outHits, 0);
    }
  }
  if (lane == 0) {
    const uint32_t cc = ((fw >> F_CC_SHIFT) + acc.hits) & 0xFFu;  // Uint8 wrap (:551-552)
    if (LAST) {
      OutRec o;
      o.x = acc.x; o.y = acc.y; o.px = pxy.x; o.py = pxy.y;
      o.meta = cc | (acc.outHits << 8);
      o.pad[0] = o.pad[1] = o.pad[2] = 0;
      s.OUT[e] = o;
    } else {
      Gout[e] = make_float4(acc.x, acc.y, r, __uint_as_float((fw & 0xFFFF00FFu) | (cc << F_CC_SHIFT)));
      s.PXY[e] = pxy;
    }
  }
}

// ---- WB + K7: results back to id order, collisionData -------------------------------------
// Gathers each entity's result sector by id, writes the by-id state coalesced, scans the
// per-entity outgoing pair counts in id order (decoupled look-back) and lets the entities
// whose pairs fall below maxCollisionPairs re-derive them in row order — which is the
// reference's emission order (i ascending, then row position; physics_worker.js:555-567).
__global__ void __launch_bounds__(WB_THREADS)
k_writeback(GridDims g, const Params* __restrict__ pp, ById d, BySlot s, const float4* __restrict__ Glast,
            const uint32_t* __restrict__ slotOf, uint32_t numTiles, unsigned long long* status,
            Counters* ctr, int32_t* __restrict__ coll, uint32_t lastSubstep) {
  __shared__ uint32_t s_tile, s_excl, s_warp[WB_THREADS / 32];
  if (threadIdx.x == 0) s_tile = atomicAdd(&ctr->wbTile, 1u);
  __syncthreads();
  const uint32_t tileIdx = s_tile;
  const uint32_t i = tileIdx * WB_THREADS + threadIdx.x;
  uint32_t slot = SLOT_NONE, outCnt = 0;
  if (i < g.N) {
    slot = slotOf[i];
    if (slot != SLOT_NONE) {
      const OutRec o = s.OUT[slot];
      d.DP[i] = make_float4(o.x, o.y, o.px, o.py);
      d.CC[i] = (uint8_t)(o.meta & 0xFFu);
      outCnt = o.meta >> 8;
    }
  }
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = outCnt;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += v;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = lane < WB_THREADS / 32 ? s_warp[lane] : 0;
    uint32_t winc = w;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= (uint32_t)o) winc += v;
    }
    if (lane < WB_THREADS / 32) s_warp[lane] = winc - w;
    const uint32_t agg = __shfl_sync(0xffffffffu, winc, 31);
    const uint32_t excl = lb_exclusive(status, tileIdx, agg, ctr->epoch);
    if (lane == 0) {
      s_excl = excl;
      if (tileIdx == numTiles - 1) {
        ctr->wbTile = 0;
        ctr->collisionPairs = excl + agg;
        if (coll) coll[0] = (int32_t)min(excl + agg, g.maxPairs);   // :565-567
      }
    }
  }
  __syncthreads();
  uint32_t base = s_excl + s_warp[warp] + (inc - outCnt);
  if (coll == nullptr || outCnt == 0 || base >= g.maxPairs) return;
  // re-derive my colliding outgoing pairs on the last sweep's start positions
  const Params p = *pp;
  const float4 gme = Glast[slot];
  float x = gme.x, y = gme.y;
  const uint32_t fw = __float_as_uint(gme.w);
  if ((fw & F_DYNAMIC_MASK) == F_DYNAMIC_VAL) apply_bounds_pos(g, gme.z, x, y);
  const uint32_t cnt = s.NCNT[slot];
  const uint32_t* ns = s.NS + (size_t)slot * g.Mpad;
  const uint32_t frame = ctr->frame;
  for (uint32_t k = 0; k < cnt && base < g.maxPairs; k++) {
    const uint32_t wd = ns[k];
    if (!(wd & NS_OUT)) continue;
    const uint32_t t = wd & NS_SLOT_MASK;
    const float4 gt = Glast[t];
    if ((__float_as_uint(gt.w) & F_COLLIDER) != F_COLLIDER) continue;
    double mx, my; bool moves;
    if (partner_pair(g, p, s, frame, lastSubstep, slot, x, y, gme.z, fw, t, gt, true, mx, my, moves)) {
      coll[1 + 2 * (size_t)base] = (int32_t)i;                       // :556-557
      coll[2 + 2 * (size_t)base] = (int32_t)s.SID[t];
      base++;
    }
  }
}

// ---- host <-> device column plumbing ---------------------------------------------------------
// Host columns (the SAB SoA columns) are staged verbatim and packed into the by-id records.
struct Staging {
  const uint8_t* t_active; const float* x; const float* y;
  const uint8_t* rb_active; const uint8_t* rb_static;
  const float* vx; const float* vy; const float* ax; const float* ay; const float* px; const float* py;
  const float* maxVel; const float* velAngle; const float* speed; const uint8_t* collCnt;
  const uint8_t* c_active; const float* radius; const uint8_t* isTrigger; const float* visRange;
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
}

struct StagingOut {
  uint8_t* t_active; float* x; float* y;
  uint8_t* rb_active; uint8_t* rb_static;
  float* vx; float* vy; float* ax; float* ay; float* px; float* py;
  float* maxVel; float* velAngle; float* speed; uint8_t* collCnt;
  uint8_t* c_active; float* radius; uint8_t* isTrigger; float* visRange;
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
}

// ---- statistics (only when the host asks) -------------------------------------------------
__global__ void __launch_bounds__(256)
k_stats(GridDims g, const uint32_t* __restrict__ NCNT, const uint32_t* __restrict__ cellStart, Counters* ctr) {
  const uint32_t A = cellStart[g.cells];
  unsigned long long sum = 0; uint32_t capped = 0;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < A; e += gridDim.x * blockDim.x) {
    const uint32_t n = NCNT[e];
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
