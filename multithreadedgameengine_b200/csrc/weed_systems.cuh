// Device-side consumers of the hot path's outputs.  They read collisionData,
// the neighbor rows and the entity state where the frame left them, so nothing but their small
// results crosses the PCIe bus.
//
//   f2  collision Enter / Stay / Exit      logic_worker.js:429-526 (processCollisionCallbacks)
//   f3  screen visibility                  particle_worker.js:1012-1062 (updateEntityScreenVisibility)
//       shadow sprites                     particle_worker.js:861-1003 (updateShadowSprites)
//
// Every output keeps the reference's ORDER (pair list order, previous-frame insertion order,
// light order then row order), produced by flag -> tile count -> tile prefix -> ordered emit
// passes instead of the reference's sequential loops.
#pragma once
#include "weed_device.cuh"

namespace weed {

static constexpr int SYS_TILE = 256;

// exclusive prefix over tile counts by ONE block; total (optionally clamped) to *total
__global__ void __launch_bounds__(1024)
k_tile_scan(const uint32_t* __restrict__ tileCount, uint32_t* __restrict__ tilePrefix, uint32_t nTiles,
            uint32_t* __restrict__ total) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nTiles; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < nTiles ? tileCount[i] : 0u;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (uint32_t)o) inc += u;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= (uint32_t)o) w += u;
      }
      s_warp[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    const uint32_t excl = carry + (warp ? s_warp[warp - 1] : 0u) + inc - v;
    if (i < nTiles) tilePrefix[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

// exclusive prefix of `flag` inside a block of SYS_TILE threads; returns my offset, *blockTotal
__device__ __forceinline__ uint32_t block_exclusive(uint32_t v, uint32_t* s_warp, uint32_t& blockTotal) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += u;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t before = 0, tot = 0;
  for (uint32_t w = 0; w < SYS_TILE / 32; w++) {
    const uint32_t c = s_warp[w];
    if (w < warp) before += c;
    tot += c;
  }
  blockTotal = tot;
  return before + inc - v;
}

// =============================================================================================
// f2: collision Enter / Stay / Exit
// =============================================================================================
// The reference keeps two JS Sets of Cantor keys (logic_worker.js:417-421, 446-460); here a pair
// is the exact 64-bit key (a << 32 | b), each frame's pairs go into an open-addressing table, and
// the previous frame's table and ordered list stay on the device.
//   state[k]  of current pair k (collisionData order): 1 = Enter (:471-480), 2 = Stay (:481-489)
//   exits     previous-frame pairs that are gone, in the previous frame's order (:493-516)
struct EvCounters {
  uint32_t cur;       // pairs of the frame being classified
  uint32_t enter, stay, exit;
  uint32_t prev;      // pairs of the previous classified frame
  uint32_t _pad[3];
};
static constexpr unsigned long long EV_EMPTY = ~0ull;

__device__ __forceinline__ uint32_t ev_hash(unsigned long long k, uint32_t mask) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return (uint32_t)k & mask;
}
__device__ __forceinline__ bool ev_contains(const unsigned long long* __restrict__ table, uint32_t mask,
                                            unsigned long long key) {
  for (uint32_t h = ev_hash(key, mask);; h = (h + 1) & mask) {
    const unsigned long long v = table[h];
    if (v == key) return true;
    if (v == EV_EMPTY) return false;
  }
}

__global__ void k_ev_begin(const int32_t* __restrict__ coll, EvCounters* ec, uint32_t forget) {
  if (threadIdx.x == 0) {
    ec->cur = (uint32_t)coll[0];
    ec->enter = 0; ec->stay = 0; ec->exit = 0;
    if (forget) ec->prev = 0;
  }
}

// current pairs -> table + ordered key list
__global__ void __launch_bounds__(SYS_TILE)
k_ev_insert(const int32_t* __restrict__ coll, unsigned long long* __restrict__ table, uint32_t mask,
            unsigned long long* __restrict__ list, const EvCounters* __restrict__ ec) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ec->cur) return;
  const unsigned long long key = ((unsigned long long)(uint32_t)coll[1 + 2 * (size_t)k] << 32) | (uint32_t)coll[2 + 2 * (size_t)k];
  list[k] = key;
  for (uint32_t h = ev_hash(key, mask);; h = (h + 1) & mask) {
    const unsigned long long old = atomicCAS(&table[h], EV_EMPTY, key);
    if (old == EV_EMPTY || old == key) break;
  }
}

// Enter / Stay of the current pairs; gone flags of the previous pairs + their tile counts
__global__ void __launch_bounds__(SYS_TILE)
k_ev_classify(const unsigned long long* __restrict__ curList, const unsigned long long* __restrict__ curTable,
              const unsigned long long* __restrict__ prevList, const unsigned long long* __restrict__ prevTable,
              uint32_t mask, uint8_t* __restrict__ state, uint32_t* __restrict__ tileCount, EvCounters* ec) {
  __shared__ uint32_t s_warp[SYS_TILE / 32];
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nCur = ec->cur, nPrev = ec->prev;
  uint32_t isEnter = 0, isStay = 0, gone = 0;
  if (k < nCur) {
    const bool had = nPrev && ev_contains(prevTable, mask, curList[k]);    // :467
    state[k] = had ? 2 : 1;
    isEnter = !had; isStay = had;
  }
  if (k < nPrev) gone = !ev_contains(curTable, mask, prevList[k]);          // :494
  const uint32_t e = __reduce_add_sync(0xffffffffu, isEnter), s = __reduce_add_sync(0xffffffffu, isStay);
  if ((threadIdx.x & 31) == 0) {
    if (e) atomicAdd(&ec->enter, e);
    if (s) atomicAdd(&ec->stay, s);
  }
  uint32_t tot;
  block_exclusive(gone, s_warp, tot);
  if (threadIdx.x == 0) tileCount[blockIdx.x] = tot;
}

// ordered list of ended pairs: exitData[0] = count, then (a, b) in the previous frame's order
__global__ void __launch_bounds__(SYS_TILE)
k_ev_exits(const unsigned long long* __restrict__ curTable, const unsigned long long* __restrict__ prevList,
           uint32_t mask, const uint32_t* __restrict__ tilePrefix, int32_t* __restrict__ exitData, EvCounters* ec) {
  __shared__ uint32_t s_warp[SYS_TILE / 32];
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nPrev = ec->prev;
  unsigned long long key = 0;
  uint32_t gone = 0;
  if (k < nPrev) { key = prevList[k]; gone = !ev_contains(curTable, mask, key); }
  uint32_t tot;
  const uint32_t off = tilePrefix[blockIdx.x] + block_exclusive(gone, s_warp, tot);
  if (gone) {
    exitData[1 + 2 * (size_t)off] = (int32_t)(uint32_t)(key >> 32);
    exitData[2 + 2 * (size_t)off] = (int32_t)(uint32_t)key;
  }
}
__global__ void k_ev_end(EvCounters* ec, int32_t* exitData) {
  if (threadIdx.x == 0) {
    exitData[0] = (int32_t)ec->exit;
    ec->prev = ec->cur;      // :521-523: current becomes previous
  }
}

// =============================================================================================
// f3a: screen visibility (particle_worker.js:1012-1062)
// =============================================================================================
struct CameraParams { double zoom, cameraX, cameraY, canvasW, canvasH; };

__global__ void __launch_bounds__(256)
k_screen_visibility(uint32_t N, CameraParams c, const float4* __restrict__ DP, const uint8_t* __restrict__ F,
                    float* __restrict__ screenX, float* __restrict__ screenY, uint8_t* __restrict__ onScreen) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (!(F[i] & F_T_ACTIVE)) return;                                   // :1046 (stale values stay)
  const double offX = dmul(c.cameraX, c.zoom), offY = dmul(c.cameraY, c.zoom);   // :1034-1035
  const double marginX = dmul(c.canvasW, 0.15), marginY = dmul(c.canvasH, 0.15); // :1036-1037
  const double minX = -marginX, maxX = dadd(c.canvasW, marginX);
  const double minY = -marginY, maxY = dadd(c.canvasH, marginY);
  const float4 dp = DP[i];
  const double sx = dsub(dmul((double)dp.x, c.zoom), offX);           // :1049-1050
  const double sy = dsub(dmul((double)dp.y, c.zoom), offY);
  screenX[i] = fround(sx);
  screenY[i] = fround(sy);
  onScreen[i] = (sx > minX && sx < maxX && sy > minY && sy < maxY) ? 1 : 0;      // :1055-1056
}

// =============================================================================================
// f3b: shadow sprites (particle_worker.js:861-1003)
// =============================================================================================
// Sequential reference: lights in id order (at most maxLights of them, counted whether or not
// they cast anything), each walks its neighbor row in order and emits at most perLight shadows,
// everything stops at maxSprites.  Parallel: light flag -> rank by id; shadow count per eligible
// light -> prefix; ordered emit truncated at maxSprites (a prefix truncation, like the break).
struct ShadowParams { uint32_t maxLights, perLight, maxSprites, _pad; };
struct ShadowIn {
  const uint8_t* lightActive; const float* lightIntensity;
  const uint8_t* casterActive; const float* casterRadius; const float* casterHeight;
  const uint8_t* onScreen;
};
struct ShadowOut { uint8_t* active; float *radius, *x, *y, *rotation, *scaleX, *scaleY, *alpha; };

__device__ __forceinline__ bool sh_is_light(const ShadowIn& in, const uint8_t* __restrict__ F, uint32_t i) {
  return in.lightActive[i] && (F[i] & F_T_ACTIVE) && in.onScreen[i] && !(in.lightIntensity[i] <= 0);   // :913-918 (`intensity <= 0` skips: NaN stays)
}

// pass 1: light flags per tile
__global__ void __launch_bounds__(SYS_TILE)
k_shadow_lights(uint32_t N, ShadowIn in, const uint8_t* __restrict__ F, uint32_t* __restrict__ tileCount) {
  __shared__ uint32_t s_warp[SYS_TILE / 32];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t isL = i < N && sh_is_light(in, F, i);
  uint32_t tot;
  block_exclusive(isL, s_warp, tot);
  if (threadIdx.x == 0) tileCount[blockIdx.x] = tot;
}

// does neighbor k of the light produce a shadow? (the `continue`s of :934-953)
__device__ __forceinline__ bool sh_casts(const ShadowIn& in, const uint8_t* __restrict__ F, int32_t j, float distSq) {
  if (!in.casterActive[j] || !(F[j] & F_T_ACTIVE) || !in.onScreen[j]) return false;
  return !(__dsqrt_rn((double)distSq) < 1.0);
}

// pass 2: the first maxLights lights (by id) list themselves and count their shadows
__global__ void __launch_bounds__(SYS_TILE)
k_shadow_count(uint32_t N, ShadowParams p, ShadowIn in, const uint8_t* __restrict__ F,
               const uint32_t* __restrict__ tilePrefix, RowView rows,
               uint32_t* __restrict__ lightId, uint32_t* __restrict__ lightCount) {
  __shared__ uint32_t s_warp[SYS_TILE / 32];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t isL = i < N && sh_is_light(in, F, i);
  uint32_t tot;
  const uint32_t rank = tilePrefix[blockIdx.x] + block_exclusive(isL, s_warp, tot);
  if (!isL || rank >= p.maxLights) return;                            // :911
  uint32_t slot;
  const int32_t cnt = rows.count(i, slot);
  uint32_t c = 0;
  for (int32_t k = 0; k < cnt && c < p.perLight; k++)                 // :930-931
    if (sh_casts(in, F, rows.id(slot, k), rows.d2(slot, k))) c++;
  lightId[rank] = i;
  lightCount[rank] = c;
}

// pass 3 (one block): prefix over the <= maxLights counts, ordered emit, clear the unused tail
__global__ void __launch_bounds__(256)
k_shadow_emit(ShadowParams p, ShadowIn in, const uint8_t* __restrict__ F, const float4* __restrict__ DP,
              const uint32_t* __restrict__ nLightsTotal, RowView rows,
              const uint32_t* __restrict__ lightId, const uint32_t* __restrict__ lightCount, ShadowOut out,
              uint32_t* __restrict__ nSprites) {
  __shared__ uint32_t s_total;
  const uint32_t nL = min(*nLightsTotal, p.maxLights);
  // lights are few (reference default 20): a serial prefix by each thread is cheapest
  for (uint32_t l = threadIdx.x; l < nL; l += blockDim.x) {
    uint32_t base = 0;
    for (uint32_t q = 0; q < l; q++) base += lightCount[q];
    const uint32_t i = lightId[l];
    const float4 lp = DP[i];
    const double intensity = (double)in.lightIntensity[i];
    uint32_t slot;
    const int32_t cnt = rows.count(i, slot);
    uint32_t c = 0;
    for (int32_t k = 0; k < cnt && c < p.perLight && base + c < p.maxSprites; k++) {
      const int32_t j = rows.id(slot, k);
      const float distSqF = rows.d2(slot, k);
      if (!sh_casts(in, F, j, distSqF)) continue;
      const double distSq = (double)distSqF;
      const float4 cp = DP[j];
      const float cr0 = in.casterRadius[j], ch0 = in.casterHeight[j];
      const double casterRadius = (cr0 != 0.f && cr0 == cr0) ? (double)cr0 : 10.0;        // `|| 10`  (:942)
      const double casterHeight = (ch0 != 0.f && ch0 == ch0) ? (double)ch0 : casterRadius; // :943
      const double dx = dsub((double)cp.x, (double)lp.x), dy = dsub((double)cp.y, (double)lp.y);
      const double dist = __dsqrt_rn(distSq);
      const double invDist = ddiv(1.0, dist);
      const double dirX = dmul(dx, invDist), dirY = dmul(dy, invDist);
      const double posX = dadd((double)cp.x, dmul(dirX, -casterRadius));
      const double posY = dadd((double)cp.y, dmul(dirY, -casterRadius));
      const double distRatio = dmul(dist, 0.00390625);
      const double clamped = distRatio > 1 ? 1.0 : distRatio;
      const double heightFactor = dmul(casterHeight, 0.025);
      const double lengthScale = dmul(dadd(0.3, dmul(clamped, 0.9)), heightFactor);
      const double widthScale = dmul(casterRadius, 0.0714);
      const double alpha = ddiv(intensity, dmul(distSq, 2.0));
      const double angle = atan2(dy, dx);
      const uint32_t sidx = base + c;
      out.active[sidx] = 1;
      out.radius[sidx] = fround(casterRadius);
      out.x[sidx] = fround(posX);
      out.y[sidx] = fround(posY);
      out.rotation[sidx] = fround(dsub(angle, 1.5707963267948966));
      out.scaleX[sidx] = fround(widthScale);
      out.scaleY[sidx] = fround(lengthScale);
      out.alpha[sidx] = fround(alpha);
      c++;
    }
  }
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (uint32_t q = 0; q < nL; q++) t += lightCount[q];
    s_total = min(t, p.maxSprites);
    *nSprites = s_total;
  }
  __syncthreads();
  for (uint32_t q = s_total + threadIdx.x; q < p.maxSprites; q += blockDim.x) out.active[q] = 0;   // :1002-1004
}

}  // namespace weed

// =============================================================================================
// f4: spawn / despawn pools (src/core/gameObject.js:794-951, 668-690, 1001-1034)
// =============================================================================================
// One pool per entity class (a contiguous index range).  The free list is the reference's LIFO
// stack with its interleaved initial order (:794-833); a batch of n spawns pops what n
// sequential spawn() calls would pop, a batch of despawns pushes what n sequential despawn()
// calls would push (inactive entities and repeats inside the batch are skipped, :669-670).
namespace weed {

struct PoolState { int32_t top; uint32_t overflow; };
static constexpr uint32_t POOL_HAS_RIGIDBODY = 1u, POOL_HAS_COLLIDER = 2u;

// :818-831: offsets 0..7, each walking the range with stride 8
__global__ void __launch_bounds__(256)
k_pool_init(uint32_t start, uint32_t count, int32_t* __restrict__ freeList, PoolState* st) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;     // write index
  if (w == 0) { st->top = (int32_t)count - 1; st->overflow = 0; }
  if (w >= count) return;
  // entries of offset o occupy [base(o), base(o) + len(o)), len(o) = ceil((count - o) / 8)
  uint32_t base = 0, o = 0;
  for (; o < 8; o++) {
    const uint32_t len = count > o ? (count - o + 7) / 8 : 0;
    if (w < base + len) break;
    base += len;
  }
  freeList[w] = (int32_t)(start + o + 8u * (w - base));
}

struct SpawnRec { float x, y, vx, vy; };

// :878-945 for record k of the batch
__global__ void __launch_bounds__(256)
k_pool_spawn(uint32_t n, uint32_t components, const SpawnRec* __restrict__ recs, const int32_t* __restrict__ freeList,
             const PoolState* __restrict__ st, ById d, int32_t* __restrict__ indicesOut) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int32_t slot = st->top - (int32_t)k;                    // :876 freeList[freeListTop--]
  if (slot < 0) { indicesOut[k] = -1; return; }                  // :868-873 pool exhausted
  const int32_t i = freeList[slot];
  indicesOut[k] = i;
  const SpawnRec r = recs[k];
  uint32_t f = d.F[i];
  float4 dp = d.DP[i];
  dp.x = r.x; dp.y = r.y;                                        // :901-905, then spawnConfig (:927-931)
  if (components & POOL_HAS_RIGIDBODY) {                         // :888-899
    f |= F_RB_ACTIVE;
    d.ACC[i] = make_float2(0.f, 0.f);
    d.V[i] = make_float4(r.vx, r.vy, 0.f, 0.f);
    reinterpret_cast<float*>(d.AT + i)[3] = 0.f;                 // velocityAngle
    dp.z = fround(dsub((double)r.x, (double)r.vx));              // :936-939
    dp.w = fround(dsub((double)r.y, (double)r.vy));
  }
  if (components & POOL_HAS_COLLIDER) f |= F_C_ACTIVE;           // :907-909
  d.DP[i] = dp;
  d.F[i] = (uint8_t)(f | F_T_ACTIVE);                            // :948
}
__global__ void k_pool_spawn_end(uint32_t n, PoolState* st) {
  if (threadIdx.x == 0) { const int32_t t = st->top - (int32_t)n; st->top = t < -1 ? -1 : t; }
}

// despawn pass 1/2: first occurrence of every index in the batch (scratch is a by-id u32 array)
__global__ void __launch_bounds__(256)
k_pool_mark(uint32_t n, const int32_t* __restrict__ indices, uint32_t* __restrict__ scratch, bool reset) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  if (reset) scratch[indices[k]] = 0xFFFFFFFFu;
  else atomicMin(&scratch[indices[k]], k);
}
__device__ __forceinline__ bool pool_gone(uint32_t k, uint32_t n, const int32_t* __restrict__ indices, uint32_t start,
                                          const uint32_t* __restrict__ scratch, const uint8_t* __restrict__ F, int32_t& i) {
  if (k >= n) return false;
  i = indices ? indices[k] : (int32_t)(start + k);               // despawnAll walks the range (:1013)
  if (indices && scratch[i] != k) return false;                  // a repeat: inactive by then (:669-670)
  return (F[i] & F_T_ACTIVE) != 0;
}
__global__ void __launch_bounds__(SYS_TILE)
k_pool_despawn_count(uint32_t n, const int32_t* __restrict__ indices, uint32_t start, const uint32_t* __restrict__ scratch,
                     const uint8_t* __restrict__ F, uint32_t* __restrict__ tileCount) {
  __shared__ uint32_t s_warp[SYS_TILE / 32];
  int32_t i;
  const uint32_t g = pool_gone(blockIdx.x * blockDim.x + threadIdx.x, n, indices, start, scratch, F, i);
  uint32_t tot;
  block_exclusive(g, s_warp, tot);
  if (threadIdx.x == 0) tileCount[blockIdx.x] = tot;
}
// :678-689: flags off, index pushed (freeList[++freeListTop] = index) in batch order
__global__ void __launch_bounds__(SYS_TILE)
k_pool_despawn_emit(uint32_t n, uint32_t capacity, uint32_t components, const int32_t* __restrict__ indices, uint32_t start,
                    const uint32_t* __restrict__ scratch, uint8_t* __restrict__ F, const uint32_t* __restrict__ tilePrefix,
                    int32_t* __restrict__ freeList, PoolState* st) {
  __shared__ uint32_t s_warp[SYS_TILE / 32];
  int32_t i = 0;
  const uint32_t g = pool_gone(blockIdx.x * blockDim.x + threadIdx.x, n, indices, start, scratch, F, i);
  uint32_t tot;
  const uint32_t off = tilePrefix[blockIdx.x] + block_exclusive(g, s_warp, tot);
  __syncthreads();                                               // every pool_gone read F before any write
  if (!g) return;
  uint32_t clear = F_T_ACTIVE;
  if (components & POOL_HAS_RIGIDBODY) clear |= F_RB_ACTIVE;
  if (components & POOL_HAS_COLLIDER) clear |= F_C_ACTIVE;
  F[i] = (uint8_t)(F[i] & ~clear);
  const int64_t pos = (int64_t)st->top + 1 + off;
  if (pos < (int64_t)capacity) freeList[pos] = i;
  else st->overflow = 1;                                         // JS: out-of-bounds typed-array store is dropped
}
__global__ void k_pool_despawn_end(const uint32_t* total, uint32_t capacity, PoolState* st, uint32_t* despawnedOut) {
  if (threadIdx.x == 0) {
    *despawnedOut = *total;
    int64_t t = (int64_t)st->top + *total;
    if (t > (int64_t)capacity - 1) { t = (int64_t)capacity - 1; st->overflow = 1; }
    st->top = (int32_t)t;
  }
}

}  // namespace weed
