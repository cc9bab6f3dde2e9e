// weed_ctx.cu — context, memory, launch plumbing and the C ABI of include/weedgpu.h.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
//        -shared -Xcompiler -fPIC   (see __graft_entry__.build()).
// There is no CPU path: without a CUDA device weed_create fails.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/weedgpu.h"
#include "weed_kernels.cuh"
#include "weed_systems.cuh"

using namespace weed;

// =============================================================================================
// layout: Component.initializeArrays / getBufferSize (src/core/Component.js:20-42, 77-93)
// =============================================================================================
namespace {
struct ColDef { const char* name; uint32_t bytes; };
const ColDef kTransform[] = {  // src/components/Transform.js:8-17
    {"active", 1}, {"entityType", 1}, {"x", 4}, {"y", 4}, {"rotation", 4}};
const ColDef kRigidBody[] = {  // src/components/RigidBody.js:9-47
    {"active", 1}, {"static", 1}, {"vx", 4}, {"vy", 4}, {"ax", 4}, {"ay", 4}, {"px", 4}, {"py", 4},
    {"angularVelocity", 4}, {"angularAccel", 4}, {"mass", 4}, {"invMass", 4}, {"inertia", 4},
    {"invInertia", 4}, {"drag", 4}, {"angularDrag", 4}, {"maxVel", 4}, {"maxAcc", 4}, {"minSpeed", 4},
    {"friction", 4}, {"velocityAngle", 4}, {"speed", 4}, {"collisionCount", 1}};
const ColDef kCollider[] = {  // src/components/Collider.js:8-46
    {"active", 1}, {"shapeType", 1}, {"offsetX", 4}, {"offsetY", 4}, {"radius", 4}, {"width", 4},
    {"height", 4}, {"isTrigger", 1}, {"restitution", 4}, {"collisionLayer", 2}, {"collisionMask", 2},
    {"aabbMinX", 4}, {"aabbMinY", 4}, {"aabbMaxX", 4}, {"aabbMaxY", 4}, {"visualRange", 4}};

const ColDef* schema(int id, uint32_t* n) {
  switch (id) {
    case WEED_BUF_TRANSFORM: *n = 5; return kTransform;
    case WEED_BUF_RIGIDBODY: *n = 23; return kRigidBody;
    case WEED_BUF_COLLIDER: *n = 16; return kCollider;
  }
  *n = 0;
  return nullptr;
}

// the 19 mirrored columns, in WEED_COL_* bit order: (buffer, schema index, element bytes)
struct HotCol { int buf; uint32_t col; uint32_t bytes; };
const HotCol kHot[20] = {
    {WEED_BUF_TRANSFORM, 0, 1},  {WEED_BUF_TRANSFORM, 2, 4},  {WEED_BUF_TRANSFORM, 3, 4},
    {WEED_BUF_RIGIDBODY, 0, 1},  {WEED_BUF_RIGIDBODY, 1, 1},  {WEED_BUF_RIGIDBODY, 2, 4},
    {WEED_BUF_RIGIDBODY, 3, 4},  {WEED_BUF_RIGIDBODY, 4, 4},  {WEED_BUF_RIGIDBODY, 5, 4},
    {WEED_BUF_RIGIDBODY, 6, 4},  {WEED_BUF_RIGIDBODY, 7, 4},  {WEED_BUF_RIGIDBODY, 16, 4},
    {WEED_BUF_RIGIDBODY, 20, 4}, {WEED_BUF_RIGIDBODY, 21, 4}, {WEED_BUF_RIGIDBODY, 22, 1},
    {WEED_BUF_COLLIDER, 0, 1},   {WEED_BUF_COLLIDER, 4, 4},   {WEED_BUF_COLLIDER, 7, 1},
    {WEED_BUF_COLLIDER, 15, 4},  {WEED_BUF_TRANSFORM, 1, 1}};

thread_local std::string g_create_error;
}  // namespace

extern "C" size_t weed_column_offset(weed_buffer_id id, uint32_t column, uint32_t entityCount) {
  uint32_t n;
  const ColDef* s = schema(id, &n);
  if (!s || column >= n) return (size_t)-1;
  size_t off = 0;
  for (uint32_t k = 0; k <= column; k++) {
    const size_t rem = off % s[k].bytes;          // Component.js:31-34
    if (rem) off += s[k].bytes - rem;
    if (k == column) return off;
    off += (size_t)entityCount * s[k].bytes;      // Component.js:37
  }
  return (size_t)-1;
}

extern "C" uint32_t weed_column_count(weed_buffer_id id) {
  uint32_t n;
  schema(id, &n);
  return n;
}

extern "C" const char* weed_column_name(weed_buffer_id id, uint32_t column) {
  uint32_t n;
  const ColDef* s = schema(id, &n);
  return (s && column < n) ? s[column].name : nullptr;
}

extern "C" size_t weed_buffer_bytes(weed_buffer_id id, uint32_t entityCount, uint32_t maxNeighbors,
                                    uint32_t maxCollisionPairs) {
  uint32_t n;
  const ColDef* s = schema(id, &n);
  if (s) {  // Component.getBufferSize
    size_t off = 0;
    for (uint32_t k = 0; k < n; k++) {
      const size_t rem = off % s[k].bytes;
      if (rem) off += s[k].bytes - rem;
      off += (size_t)entityCount * s[k].bytes;
    }
    return off;
  }
  switch (id) {
    case WEED_BUF_NEIGHBOR:
    case WEED_BUF_DISTANCE: return (size_t)entityCount * (1 + (size_t)maxNeighbors) * 4;  // gameEngine.js:554-558
    case WEED_BUF_COLLISION: return (1 + (size_t)maxCollisionPairs * 2) * 4;              // gameEngine.js:694
    default: return 0;
  }
}

extern "C" void weed_default_config(weed_config* c) {
  if (!c) return;
  memset(c, 0, sizeof(*c));
  c->struct_size = sizeof(weed_config);
  c->maxNeighbors = 100;          // gameEngine.js:553
  c->maxCollisionPairs = 10000;   // gameEngine.js:689-693
  c->seed = 1.0;
  c->physics.subStepCount = 4;    // gameEngine.js:39-45
  c->physics.boundaryElasticity = 0.8;
  c->physics.collisionResponseStrength = 0.5;
  c->physics.verletDamping = 0.995;
  c->physics.minSpeedForRotation = 0.1;
}

// =============================================================================================
// context
// =============================================================================================
// timed spans of one frame (WEED_FLAG_KERNEL_TIMING): ms[k] = ev[k] -> ev[k+1]
//   0 k_cell_key, 1 k_cell_scan, 2 k_scatter_ids, 3 k_build_slots + k_slot_prep, 4 k_neighbors,
//   5 k_capped_rescan + k_sort_lists, 6 all k_substep launches, 7 k_writeback + k_pair_scan + k_pair_emit
static constexpr int kTimedSpans = 8;

struct weed_ctx {
  weed_config cfg;
  weed_physics_config phys;
  GridDims g;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool ownStream = false;
  bool failed = false;
  std::string err;

  // host buffers (owned by the caller)
  void* host[WEED_BUF_COUNT] = {};
  bool registered[WEED_BUF_COUNT] = {};

  std::vector<void*> allocs;
  // by id
  ById d{};
  uint32_t *key = nullptr, *rank = nullptr, *arrIds = nullptr, *slotOf = nullptr;
  // grid
  uint32_t *cellCount = nullptr, *cellStart = nullptr;
  uint32_t *bigCells = nullptr;   // cells above BIG_CELL entities
  uint32_t bigCap = 0;
  uint32_t scanTiles = 0, wbTiles = 0;
  unsigned long long *scanStatus = nullptr;
  uint32_t *tileCount = nullptr, *tilePrefix = nullptr;
  // by slot
  BySlot s{};
  // outputs
  int32_t* nd = nullptr; float* dd = nullptr; int32_t* coll = nullptr;   // API rows as slot-major planes (RowView)
  size_t rowWords = 0;
  int32_t* mnd = nullptr; float* mdd = nullptr;   // the rows in the reference's layout, filled on demand (k_rows_gather)
  // control
  Params* dParams = nullptr;
  FrameConst* dFrameConst = nullptr;   // GridDims + BySlot in device memory (slow paths of the sweep)
  Counters* dCtr = nullptr;
  Params hParams{};
  double lastDt = -1;
  bool paramsDirty = true;
  // staging for host<->device columns
  void* stage[20] = {};
  float* protRange = nullptr;
  // second stream + events of the pipelined weed_step (column copies overlap the frame)
  cudaStream_t copyStream = nullptr;
  cudaEvent_t evUp = nullptr, evBuilt = nullptr, evCopied = nullptr;
  cudaStream_t sideStream = nullptr;                 // k_sweep_heavy beside k_sweep, k_sort_lists beside the cap path (fork / join, also under capture)
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  // graph of one full frame
  cudaGraphExec_t frameGraph = nullptr;
  int graphSubSteps = -1;
  bool spatialValid = false;  // rows/slots of the current frame exist (weed_spatial ran)
  cudaEvent_t ev[kTimedSpans + 1] = {};
  float ms[12] = {};
  uint32_t launchesPerStep = 0;
  // device-side systems (allocated on first use)
  unsigned long long *evTable[2] = {}, *evList[2] = {};
  uint32_t evMask = 0, evCur = 0;
  uint8_t* evState = nullptr; int32_t* evExit = nullptr; EvCounters* dEv = nullptr;
  uint32_t *sysTileCount = nullptr, *sysTilePrefix = nullptr; size_t sysTiles = 0;
  float *screenX = nullptr, *screenY = nullptr; uint8_t* onScreen = nullptr;
  uint8_t *shLightActive = nullptr, *shCasterActive = nullptr;
  float *shIntensity = nullptr, *shRadius = nullptr, *shHeight = nullptr;
  uint32_t *shLightId = nullptr, *shLightCount = nullptr, *shScalars = nullptr;
  uint8_t* shOutActive = nullptr; float* shOut = nullptr; uint32_t shOutCap = 0, shLightCap = 0;
  struct Pool { uint32_t start, count, components; int32_t* freeList; PoolState* st; };
  std::vector<Pool> pools;
  void* poolStage = nullptr; size_t poolStageBytes = 0;   // device staging for a batch (records / indices)
  uint32_t* poolScalars = nullptr;
  bool k4Wide = false;        // warp-per-entity neighbor scan (long rows: maxNeighbors >= 256; WEED_FLAG_K4_WIDE / _THREAD force one)
  // slabs
  bool slab = false;
  uint32_t* holes = nullptr;
  SlabCounters* dSlab = nullptr;
  // peer-to-peer exchange (weed_slab_exchange_*): my receive buffers, the neighbours' mapped ones
  SlabRec* xRecv = nullptr;            // 4 buffers of xQuota + 1 records: [side][parity]
  uint32_t xQuota = 0;
  SlabXfer hXfer{};
  SlabXfer* dXfer = nullptr;
  void* xPeerIpc[2] = {};              // bases opened with cudaIpcOpenMemHandle (closed in weed_destroy)
  size_t copyCount = 0;                // entities weed_step moves per column: the table top of a slab, else N
};

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ctx->failed = true;                                                                     \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e);                          \
      return WEED_E_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define GUARD(ctx)                                                                            \
  do {                                                                                        \
    if (!(ctx)) return WEED_E_INVALID;                                                        \
    if ((ctx)->failed) return WEED_E_CUDA;                                                    \
    cudaSetDevice((ctx)->device);                                                             \
  } while (0)

static int fail(weed_ctx* ctx, int code, const std::string& msg) {
  ctx->err = msg;
  return code;
}

template <typename T>
static int dalloc(weed_ctx* ctx, T** p, size_t count, bool zero = true) {
  void* q = nullptr;
  const size_t bytes = (count ? count : 1) * sizeof(T);
  cudaError_t e = cudaMalloc(&q, bytes);
  if (e != cudaSuccess) {
    ctx->failed = true;
    ctx->err = "cudaMalloc(" + std::to_string(bytes) + " B): " + cudaGetErrorString(e);
    return e == cudaErrorMemoryAllocation ? WEED_E_NOMEM : WEED_E_CUDA;
  }
  ctx->allocs.push_back(q);
  if (zero) {
    e = cudaMemsetAsync(q, 0, bytes, ctx->stream);
    if (e != cudaSuccess) { ctx->failed = true; ctx->err = cudaGetErrorString(e); return WEED_E_CUDA; }
  }
  *p = (T*)q;
  return WEED_OK;
}

static double clamp01(double v, double fallback) {  // utils.js:16-19
  if (v != v) return fallback;
  return fmax(0.0, fmin(1.0, v));
}

// validatePhysicsConfig (src/core/utils.js:269-301)
static void validate_physics(weed_physics_config* cur, const weed_physics_config* in) {
  weed_physics_config o = *in;
  o.subStepCount = in->subStepCount < 1 ? 1 : in->subStepCount;
  o.boundaryElasticity = clamp01(in->boundaryElasticity, cur->boundaryElasticity);
  o.collisionResponseStrength = clamp01(in->collisionResponseStrength, cur->collisionResponseStrength);
  o.verletDamping = clamp01(in->verletDamping, cur->verletDamping);
  *cur = o;
}

static void refresh_params(weed_ctx* ctx, double dtRatio) {
  Params& p = ctx->hParams;
  const weed_physics_config& ph = ctx->phys;
  // `this.settings.gravity.x || 0` (physics_worker.js:179-180): NaN and 0 give 0
  const double gx = (ph.gravityX == ph.gravityX && ph.gravityX != 0) ? ph.gravityX : 0.0;
  const double gy = (ph.gravityY == ph.gravityY && ph.gravityY != 0) ? ph.gravityY : 0.0;
  const double gs = dtRatio * dtRatio;             // Math.pow(dtRatio, 2), physics_worker.js:261
  p.dtRatio = dtRatio;
  p.gravityScaleX = gs * gx;
  p.gravityScaleY = gs * gy;
  p.damping = ph.verletDamping;
  p.boundaryElasticity = ph.boundaryElasticity;
  p.responseStrength = ph.collisionResponseStrength;
  p.minSpeedForRotation = ph.minSpeedForRotation;
  // ToUint32(seed) for the nudge hash
  double sd = ctx->cfg.seed;
  uint32_t s32 = 0;
  if (std::isfinite(sd)) {
    double m = fmod(trunc(sd), 4294967296.0);
    if (m < 0) m += 4294967296.0;
    s32 = (uint32_t)m;
  }
  p.seed32 = s32;
}

static int push_params(weed_ctx* ctx, double dtRatio) {
  if (!ctx->paramsDirty && dtRatio == ctx->lastDt) return WEED_OK;
  refresh_params(ctx, dtRatio);
  CK(cudaMemcpyAsync(ctx->dParams, &ctx->hParams, sizeof(Params), cudaMemcpyHostToDevice, ctx->stream));
  ctx->lastDt = dtRatio;
  ctx->paramsDirty = false;
  return WEED_OK;
}

extern "C" const char* weed_last_error(weed_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" void weed_destroy(weed_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->copyStream) { cudaStreamSynchronize(ctx->copyStream); cudaStreamDestroy(ctx->copyStream); }
  if (ctx->sideStream) { cudaStreamSynchronize(ctx->sideStream); cudaStreamDestroy(ctx->sideStream); }
  for (cudaEvent_t e : {ctx->evUp, ctx->evBuilt, ctx->evCopied, ctx->evFork, ctx->evJoin})
    if (e) cudaEventDestroy(e);
  if (ctx->frameGraph) cudaGraphExecDestroy(ctx->frameGraph);
  for (int b = 0; b < WEED_BUF_COUNT; b++)
    if (ctx->registered[b]) cudaHostUnregister(ctx->host[b]);
  for (void* p : ctx->xPeerIpc)
    if (p) cudaIpcCloseMemHandle(p);
  for (void* p : ctx->allocs) cudaFree(p);
  for (auto& e : ctx->ev)
    if (e) cudaEventDestroy(e);
  if (ctx->ownStream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" int weed_create(const weed_config* cfg, weed_ctx** out) {
  if (out) *out = nullptr;
  if (!cfg || !out) { g_create_error = "null argument"; return WEED_E_INVALID; }
  if (cfg->struct_size != sizeof(weed_config)) { g_create_error = "weed_config.struct_size mismatch (ABI)"; return WEED_E_INVALID; }
  if (cfg->entityCount == 0 || cfg->entityCount > 0x3FFFFFF0u) { g_create_error = "entityCount out of range"; return WEED_E_INVALID; }
  if (cfg->maxNeighbors > 32000) { g_create_error = "maxNeighbors above 32000"; return WEED_E_INVALID; }
  {
    const double mpad = (double)(((cfg->maxNeighbors + 7) / 8) * 8);
    const double mint = mpad + std::max(16.0, std::min(mpad, 128.0));
    if (((double)cfg->entityCount + 128.0) * mint >= 4294967295.0) {
      g_create_error = "entityCount * (internal row capacity) must stay below 2^32 per context (partition the world into slabs)";
      return WEED_E_INVALID;
    }
  }
  if (!(cfg->cellSize > 0) || !(cfg->worldWidth > 0) || !(cfg->worldHeight > 0)) { g_create_error = "world/cell size must be positive"; return WEED_E_INVALID; }
  const double colsD = ceil(cfg->worldWidth / cfg->cellSize), rowsD = ceil(cfg->worldHeight / cfg->cellSize);  // spatial_worker.js:82-83
  if (!(colsD >= 1) || !(rowsD >= 1) || colsD * rowsD > 1.0e9 || colsD > 65535 || rowsD > 65535) {
    g_create_error = "grid too large (at most 65535 x 65535 cells, 1e9 in total)";
    return WEED_E_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_create_error = "no CUDA device: libweedgpu has no CPU fallback";
    return WEED_E_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "bad device ordinal"; return WEED_E_INVALID; }

  weed_ctx* ctx = new weed_ctx();
  ctx->cfg = *cfg;
  ctx->device = cfg->device;
  weed_physics_config defaults;
  {
    weed_config d; weed_default_config(&d); defaults = d.physics;
  }
  ctx->phys = defaults;
  validate_physics(&ctx->phys, &cfg->physics);
  auto bail = [&](int code) { g_create_error = ctx->err; weed_destroy(ctx); return code; };
  if (cudaSetDevice(ctx->device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return bail(WEED_E_CUDA); }
  if (cfg->stream) ctx->stream = (cudaStream_t)cfg->stream;
  else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { ctx->err = "cudaStreamCreate failed"; return bail(WEED_E_CUDA); }
    ctx->ownStream = true;
  }
  if (cudaStreamCreateWithFlags(&ctx->sideStream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->evJoin, cudaEventDisableTiming) != cudaSuccess) { ctx->err = "side stream / events failed"; return bail(WEED_E_CUDA); }
  GridDims& g = ctx->g;
  g.inv = 1.0 / cfg->cellSize;                     // spatial_worker.js:81
  g.worldW = cfg->worldWidth; g.worldH = cfg->worldHeight;
  g.cols = (int32_t)colsD; g.rows = (int32_t)rowsD;
  g.cells = (uint32_t)(g.cols * g.rows);
  g.N = cfg->entityCount;
  g.M = cfg->maxNeighbors;
  g.Mpad = ((g.M + 7) / 8) * 8; if (g.Mpad == 0) g.Mpad = 8;
  g.Mint = g.Mpad + std::max(16u, std::min(g.Mpad, 128u));   // API row + lower-id partners found past the cap (then F_XOVER)
  g.Npad = ((g.N + TILE - 1) / TILE) * TILE;   // whole tiles: a tile's row words are one aligned bulk copy
  g.maxPairs = cfg->maxCollisionPairs;
  ctx->k4Wide = (cfg->flags & WEED_FLAG_K4_WIDE) ? true : (cfg->flags & WEED_FLAG_K4_THREAD) ? false : cfg->maxNeighbors >= 256;
  g.slabBegin = 0; g.slabEnd = g.rows; g.slabHalo = 0;
  if (cfg->slabRowEnd > 0) {
    if (cfg->slabRowBegin >= cfg->slabRowEnd || (int32_t)cfg->slabRowEnd > g.rows) { ctx->err = "bad slab rows"; return bail(WEED_E_INVALID); }
    ctx->slab = true;
    g.slabBegin = (int32_t)cfg->slabRowBegin; g.slabEnd = (int32_t)cfg->slabRowEnd; g.slabHalo = (int32_t)cfg->slabHaloRows;
  }
  g.Wsafe = nextafterf((float)(g.worldW * (1.0 - 2.4e-7)), 0.0f);
  g.Hsafe = nextafterf((float)(g.worldH * (1.0 - 2.4e-7)), 0.0f);
  const size_t N = g.N;
  int rc;
#define A(ptr, count) if ((rc = dalloc(ctx, &(ptr), (count))) != WEED_OK) return bail(rc)
  A(ctx->d.DP, N); A(ctx->d.ACC, N); A(ctx->d.AT, N); A(ctx->d.V, N); A(ctx->d.F, N); A(ctx->d.CC, N); A(ctx->d.ET, N);
  A(ctx->key, N); A(ctx->rank, N); A(ctx->arrIds, N); A(ctx->slotOf, N);
  ctx->scanTiles = (uint32_t)(((size_t)g.cells + 1 + SCAN_TILE - 1) / SCAN_TILE);
  A(ctx->cellCount, (size_t)ctx->scanTiles * SCAN_TILE);
  A(ctx->cellStart, (size_t)ctx->scanTiles * SCAN_TILE);
  ctx->bigCap = (uint32_t)(N / BIG_CELL + 16);
  A(ctx->bigCells, ctx->bigCap);
  A(ctx->scanStatus, ctx->scanTiles);
  ctx->wbTiles = (uint32_t)((N + WB_THREADS - 1) / WB_THREADS);
  A(ctx->tileCount, ctx->wbTiles); A(ctx->tilePrefix, ctx->wbTiles);
  A(ctx->s.SA, 2 * N); A(ctx->s.QXY, N + 4);     /* k_neighbors2 reads up to three positions past a range */ A(ctx->s.CXY, N); A(ctx->s.WIN, N); if (ctx->k4Wide) A(ctx->s.PW, N); A(ctx->s.GA, N); A(ctx->s.GB, N); A(ctx->s.PXY, N);
  A(ctx->s.NCNT, N); A(ctx->s.XHEAD, N); A(ctx->s.OUT, N); A(ctx->s.LSLOT, N); A(ctx->s.CAPLIST, N); A(ctx->s.SORTLIST, N); A(ctx->s.HEAVY, N); A(ctx->s.BCNT, N); A(ctx->s.BCUR, N);
  if (cfg->flags & WEED_FLAG_K6_TILE) A(ctx->s.TD, ((N + PREP_THREADS - 1) / PREP_THREADS) * (PREP_THREADS / TILE));
  A(ctx->s.NST, (size_t)g.Npad * g.Mint);
  g.xpoolRows = (uint32_t)std::max<size_t>(256, N / 32);     // 64 bytes per entity
  A(ctx->s.XPID, N); A(ctx->s.XR, (size_t)g.xpoolRows * XPOOL_ROW); A(ctx->s.XRCNT, g.xpoolRows);
  A(ctx->s.XNEXT, (size_t)g.Npad * g.Mpad);
  ctx->rowWords = (size_t)g.Npad * g.Mpad;
  if (!(cfg->flags & WEED_FLAG_NO_NEIGHBOR_ROWS)) { A(ctx->nd, ctx->rowWords); A(ctx->dd, ctx->rowWords); }
  A(ctx->coll, 1 + 2 * (size_t)g.maxPairs);
  A(ctx->dParams, 1); A(ctx->dCtr, 1);
  if (ctx->slab) { A(ctx->d.GID, N); A(ctx->s.SLID, N); A(ctx->holes, N); A(ctx->dSlab, 1); }
  for (int c = 0; c < 20; c++) {
    uint8_t* p = nullptr;
    if ((rc = dalloc(ctx, &p, N * kHot[c].bytes)) != WEED_OK) return bail(rc);
    ctx->stage[c] = p;
  }
#undef A
  if ((rc = dalloc(ctx, &ctx->dFrameConst, 1)) != WEED_OK) return bail(rc);
  {
    FrameConst fc{ctx->g, ctx->s};
    if (cudaMemcpyAsync(ctx->dFrameConst, &fc, sizeof(fc), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "FrameConst upload failed"; return bail(WEED_E_CUDA); }
  }
  for (auto& e : ctx->ev)
    if (cudaEventCreate(&e) != cudaSuccess) { ctx->err = "cudaEventCreate failed"; return bail(WEED_E_CUDA); }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "init sync failed"; return bail(WEED_E_CUDA); }
  *out = ctx;
  return WEED_OK;
}

extern "C" int weed_bind(weed_ctx* ctx, weed_buffer_id id, void* host_base, size_t bytes) {
  GUARD(ctx);
  if ((int)id < 0 || id >= WEED_BUF_COUNT) return fail(ctx, WEED_E_INVALID, "bad buffer id");
  const size_t need = weed_buffer_bytes(id, ctx->g.N, ctx->g.M, ctx->g.maxPairs);
  if (host_base && bytes < need)
    return fail(ctx, WEED_E_SIZE, "buffer " + std::to_string((int)id) + ": " + std::to_string(bytes) + " B < " + std::to_string(need) + " B");
  if (ctx->registered[id]) { cudaHostUnregister(ctx->host[id]); ctx->registered[id] = false; }
  ctx->host[id] = host_base;
  if (host_base && id <= WEED_BUF_COLLIDER) {
    // pin the SAB for the context lifetime so column copies are true async DMA; optional
    if (cudaHostRegister(host_base, need, cudaHostRegisterDefault) == cudaSuccess) ctx->registered[id] = true;
    else cudaGetLastError();
  }
  return WEED_OK;
}

static void* host_col(weed_ctx* ctx, int c) {
  return (uint8_t*)ctx->host[kHot[c].buf] + weed_column_offset((weed_buffer_id)kHot[c].buf, kHot[c].col, ctx->g.N);
}

static int check_bound(weed_ctx* ctx, uint32_t mask) {
  for (int c = 0; c < 20; c++)
    if ((mask >> c) & 1u)
      if (!ctx->host[kHot[c].buf]) return fail(ctx, WEED_E_NOT_BOUND, "component buffer " + std::to_string(kHot[c].buf) + " not bound");
  return WEED_OK;
}

static int upload_async(weed_ctx* ctx, uint32_t mask, cudaStream_t stream = nullptr, size_t count = 0) {
  if (!stream) stream = ctx->stream;
  mask &= WEED_COLS_INPUT_ALL;
  if (!mask) return WEED_OK;
  int rc = check_bound(ctx, mask);
  if (rc) return rc;
  const size_t N = count ? count : ctx->g.N;
  for (int c = 0; c < 20; c++)
    if ((mask >> c) & 1u)
      CK(cudaMemcpyAsync(ctx->stage[c], host_col(ctx, c), N * kHot[c].bytes, cudaMemcpyHostToDevice, stream));
  Staging st;
  const void** sp = reinterpret_cast<const void**>(&st);
  for (int c = 0; c < 20; c++) sp[c] = ctx->stage[c];
  k_pack<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>((uint32_t)N, mask, st, ctx->d);
  CK(cudaGetLastError());
  // The grid and the rows of a preceding weed_spatial depend on Transform.active, x, y and
  // Collider.visualRange only: a tick() that read the rows and now uploads ax / ay (or any other column)
  // leaves them valid, so weed_spatial -> upload -> weed_physics works (one worker swapped at a time).
  if (mask & (WEED_COL_T_ACTIVE | WEED_COL_T_X | WEED_COL_T_Y | WEED_COL_C_VISRANGE)) ctx->spatialValid = false;
  return WEED_OK;
}

// API rows device -> host.  The scan kernels keep the rows as slot-major planes (RowView); the
// reference's layout (row i at i * (1 + maxNeighbors), gameEngine.js:552-559) is produced on demand:
// k_rows_gather writes header + entries of the requested rows into a device mirror that has the host
// stride and persists (so the words past 1 + count and the rows of entities that are not in the grid
// keep their old contents, as in the reference), then the mirror rows go to the host in one copy.
// `absolute`: the destination is the whole bound buffer (row `first` lands at its own offset),
// otherwise the destination starts at row `first`.
static RowView row_view(weed_ctx* ctx) {
  return RowView{ctx->nd, ctx->dd, ctx->s.NCNT, ctx->slotOf, ctx->g.Npad};
}

static int copy_rows(weed_ctx* ctx, void* nd_host, void* dd_host, uint32_t first, uint32_t count, bool absolute,
                     cudaStream_t stream) {
  if (!count) return WEED_OK;
  const size_t hostWords = 1 + (size_t)ctx->g.M;
  if (!ctx->mnd) {
    int rc = dalloc(ctx, &ctx->mnd, (size_t)ctx->g.N * hostWords);
    if (rc) return rc;
    rc = dalloc(ctx, &ctx->mdd, (size_t)ctx->g.N * hostWords);
    if (rc) return rc;
    if (stream != ctx->stream) CK(cudaStreamSynchronize(ctx->stream));     // the zero fill ran on the context's stream
  }
  k_rows_gather<<<(unsigned)(((size_t)count * 32 + 255) / 256), 256, 0, stream>>>(row_view(ctx), first, count, (uint32_t)hostWords,
                                                                                 ctx->mnd, ctx->mdd);
  CK(cudaGetLastError());
  const size_t off = (size_t)first * hostWords, bytes = (size_t)count * hostWords * 4;
  const size_t hostOff = absolute ? off * 4 : 0;
  if (nd_host) CK(cudaMemcpyAsync((uint8_t*)nd_host + hostOff, ctx->mnd + off, bytes, cudaMemcpyDeviceToHost, stream));
  if (dd_host) CK(cudaMemcpyAsync((uint8_t*)dd_host + hostOff, ctx->mdd + off, bytes, cudaMemcpyDeviceToHost, stream));
  return WEED_OK;
}

static int download_async(weed_ctx* ctx, uint32_t mask, cudaStream_t stream = nullptr, size_t count = 0) {
  if (!stream) stream = ctx->stream;
  const uint32_t cols = mask & WEED_COLS_INPUT_ALL;
  const size_t N = count ? count : ctx->g.N;
  if (cols) {
    int rc = check_bound(ctx, cols);
    if (rc) return rc;
    StagingOut st;
    void** sp = reinterpret_cast<void**>(&st);
    for (int c = 0; c < 20; c++) sp[c] = ctx->stage[c];
    k_unpack<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>((uint32_t)N, cols, st, ctx->d);
    CK(cudaGetLastError());
    for (int c = 0; c < 20; c++)
      if ((cols >> c) & 1u)
        CK(cudaMemcpyAsync(host_col(ctx, c), ctx->stage[c], N * kHot[c].bytes, cudaMemcpyDeviceToHost, stream));
  }
  if (mask & WEED_COL_NEIGHBORS) {
    if (!ctx->nd) return fail(ctx, WEED_E_STATE, "neighbor rows disabled (WEED_FLAG_NO_NEIGHBOR_ROWS)");
    if (!ctx->host[WEED_BUF_NEIGHBOR] || !ctx->host[WEED_BUF_DISTANCE]) return fail(ctx, WEED_E_NOT_BOUND, "neighbor/distance buffer not bound");
    int rc2 = copy_rows(ctx, ctx->host[WEED_BUF_NEIGHBOR], ctx->host[WEED_BUF_DISTANCE], 0, (uint32_t)ctx->g.N, true, stream);
    if (rc2) return rc2;
  }
  if (mask & WEED_COL_COLLISIONS) {
    if (!ctx->host[WEED_BUF_COLLISION]) return fail(ctx, WEED_E_NOT_BOUND, "collision buffer not bound");
    // pairCount first, then only the pairs that exist
    CK(cudaMemcpyAsync(ctx->host[WEED_BUF_COLLISION], ctx->coll, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int32_t n = *(int32_t*)ctx->host[WEED_BUF_COLLISION];
    if (n > 0)
      CK(cudaMemcpyAsync((int32_t*)ctx->host[WEED_BUF_COLLISION] + 1, ctx->coll + 1, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  return WEED_OK;
}

extern "C" int weed_upload(weed_ctx* ctx, uint32_t mask) {
  GUARD(ctx);
  int rc = upload_async(ctx, mask);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return WEED_OK;
}

extern "C" int weed_download(weed_ctx* ctx, uint32_t mask) {
  GUARD(ctx);
  int rc = download_async(ctx, mask);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return WEED_OK;
}

// ---- launches --------------------------------------------------------------------------------
// slab contexts keep their cuts in device memory (they may move between frames)
static inline const int32_t* slab_cuts(weed_ctx* ctx) { return ctx->slab ? &ctx->dSlab->curBegin : nullptr; }

static inline unsigned blocks_for(size_t threads, unsigned bs) { return (unsigned)((threads + bs - 1) / bs); }

#define TIME_MARK(ctx, timing, k) do { if (timing) cudaEventRecord((ctx)->ev[k], (ctx)->stream); } while (0)

static int launch_spatial(weed_ctx* ctx, bool integrate, bool timing, cudaEvent_t waitBeforeBuild = nullptr,
                          cudaEvent_t recordAfterBuild = nullptr) {
  const GridDims& g = ctx->g;
  cudaStream_t st = ctx->stream;
  const unsigned nb = blocks_for(g.N, 256);
  k_spatial_begin<<<1, 32, 0, st>>>(ctx->dCtr);
  TIME_MARK(ctx, timing, 0);
  k_cell_key<<<nb, 256, 0, st>>>(g, ctx->d.DP, ctx->d.F, ctx->key, ctx->rank, ctx->cellCount);
  TIME_MARK(ctx, timing, 1);
  k_cell_scan<<<ctx->scanTiles, SCAN_THREADS, 0, st>>>(ctx->cellCount, ctx->cellStart, ctx->scanTiles, ctx->scanStatus, ctx->dCtr,
                                                       slab_cuts(ctx), (uint32_t)g.cols, g.slabHalo, ctx->bigCells, ctx->bigCap);
  TIME_MARK(ctx, timing, 2);
  k_scatter_ids<<<nb, 256, 0, st>>>(g.N, ctx->key, ctx->rank, ctx->cellStart, ctx->arrIds, ctx->d.GID);
  k_sort_big_cells<<<148 * 4, 256, 0, st>>>(ctx->bigCells, ctx->bigCap, ctx->dCtr, ctx->cellStart, ctx->arrIds);
  k_slot_rank<<<nb, 256, 0, st>>>(g, ctx->d.F, ctx->d.GID, ctx->key, ctx->cellStart, ctx->arrIds, ctx->slotOf);
  TIME_MARK(ctx, timing, 3);
  if (waitBeforeBuild) CK(cudaStreamWaitEvent(st, waitBeforeBuild, 0));
  if (integrate)
    k_build_slots<true><<<nb, 256, 0, st>>>(g, ctx->dParams, ctx->phys.subStepCount, false, ctx->d, ctx->s, ctx->key, ctx->cellStart, ctx->arrIds, ctx->slotOf, slab_cuts(ctx));
  else
    k_build_slots<false><<<nb, 256, 0, st>>>(g, ctx->dParams, ctx->phys.subStepCount, false, ctx->d, ctx->s, ctx->key, ctx->cellStart, ctx->arrIds, ctx->slotOf, slab_cuts(ctx));
  if (recordAfterBuild) CK(cudaEventRecord(recordAfterBuild, st));
  k_slot_prep<<<blocks_for(g.N, PREP_THREADS), PREP_THREADS, 0, st>>>(g, ctx->dParams, ctx->s, ctx->cellStart);
  TIME_MARK(ctx, timing, 4);
  // long rows (the reference's own demos: maxNeighbors 400-1500) -> one warp per entity;
  // short rows (the large synthetic worlds) -> one thread per entity with staged flush
  const bool wide = ctx->k4Wide;
  if (wide) {
    const unsigned wb = blocks_for((size_t)g.N * 32, 256);
    if (ctx->nd) k_neighbors_wide<true><<<wb, 256, 0, st>>>(g, ctx->s, ctx->cellStart, ctx->nd, ctx->dd, ctx->dCtr);
    else         k_neighbors_wide<false><<<wb, 256, 0, st>>>(g, ctx->s, ctx->cellStart, nullptr, nullptr, ctx->dCtr);
  } else {
    const unsigned kb = blocks_for(g.N, K4V2_THREADS);
    if (ctx->nd) k_neighbors2<true><<<kb, K4V2_THREADS, 0, st>>>(g, ctx->s, ctx->cellStart, ctx->nd, ctx->dd, ctx->dCtr);
    else         k_neighbors2<false><<<kb, K4V2_THREADS, 0, st>>>(g, ctx->s, ctx->cellStart, nullptr, nullptr, ctx->dCtr);
  }
  TIME_MARK(ctx, timing, 5);
  // explicit lists (K4's own pushes) are sorted beside the cap path: nothing in common but their predecessor
  CK(cudaEventRecord(ctx->evFork, st));
  CK(cudaStreamWaitEvent(ctx->sideStream, ctx->evFork, 0));
  k_sort_lists<<<148 * 4, 256, 0, ctx->sideStream>>>(g, ctx->s, ctx->cellStart, ctx->dCtr);
  CK(cudaEventRecord(ctx->evJoin, ctx->sideStream));
  k_beyond_cap<<<K4B_BLOCKS, 256, 0, st>>>(g, ctx->s, ctx->cellStart, ctx->dCtr);            // few capped rows: a warp each
  k_back_alloc<<<148 * 4, 256, 0, st>>>(g, ctx->s, ctx->cellStart, ctx->dCtr);                // many capped rows (a settled bed): the reverse-edge form;
  k_back_write<<<K4B_BLOCKS, 256, 0, st>>>(g, ctx->s, ctx->cellStart, ctx->dCtr);             // each of these returns at once in the other regime
  k_back_sort<<<148 * BSORT_BLOCKS_PER_SM, BSORT_WARPS * 32, 0, st>>>(g, ctx->s, ctx->cellStart, ctx->dCtr);
  CK(cudaStreamWaitEvent(st, ctx->evJoin, 0));
  TIME_MARK(ctx, timing, 6);
  CK(cudaGetLastError());
  return WEED_OK;
}

static int launch_constraints(weed_ctx* ctx, bool timing) {
  const GridDims& g = ctx->g;
  cudaStream_t st = ctx->stream;
  const int S = ctx->phys.subStepCount;
  // sweep 0 reads GA (written by k_slot_prep), then GB / GA alternate
  const float4* in = ctx->s.GA;
  const uint32_t gs = 1;
  float4* bufs[2] = {ctx->s.GB, ctx->s.GA};
  const uint32_t kflags = ctx->cfg.flags;
  for (int step = 0; step < S; step++) {
    float4* out = bufs[step & 1];
    const bool first = step == 0, last = step == S - 1;
#define SWEEP_ARGS g, ctx->dParams, ctx->s, in, out, ctx->cellStart, ctx->dCtr, (uint32_t)step, ctx->dFrameConst
    if (kflags & WEED_FLAG_K6_TILE) {
      const unsigned sb = blocks_for(g.N, TILE);
      if (first && last)       k_sweep_tile<true, true><<<sb, TILE, 0, st>>>(SWEEP_ARGS);
      else if (first)          k_sweep_tile<true, false><<<sb, TILE, 0, st>>>(SWEEP_ARGS);
      else if (last)           k_sweep_tile<false, true><<<sb, TILE, 0, st>>>(SWEEP_ARGS);
      else                     k_sweep_tile<false, false><<<sb, TILE, 0, st>>>(SWEEP_ARGS);
    } else {
      const unsigned sb = blocks_for(g.N, K6V2_THREADS);
      if (HEAVY_SPLIT) {       // the popular entities of piles (F_XPOOL / F_XOVER), a warp each, on a parallel branch: other entities,
                               // same input, so the few long warps run beside k_sweep instead of after it
        cudaStream_t hs = ctx->sideStream;
        CK(cudaEventRecord(ctx->evFork, st));
        CK(cudaStreamWaitEvent(hs, ctx->evFork, 0));
#define HEAVY_ARGS g, ctx->dParams, ctx->s, in, out, ctx->cellStart, ctx->dCtr, (uint32_t)step
        if (first && last)     k_sweep_heavy<true, true><<<K6H_BLOCKS, K6H_THREADS, 0, hs>>>(HEAVY_ARGS);
        else if (first)        k_sweep_heavy<true, false><<<K6H_BLOCKS, K6H_THREADS, 0, hs>>>(HEAVY_ARGS);
        else if (last)         k_sweep_heavy<false, true><<<K6H_BLOCKS, K6H_THREADS, 0, hs>>>(HEAVY_ARGS);
        else                   k_sweep_heavy<false, false><<<K6H_BLOCKS, K6H_THREADS, 0, hs>>>(HEAVY_ARGS);
#undef HEAVY_ARGS
        CK(cudaEventRecord(ctx->evJoin, hs));
      }
      if (first && last)       k_sweep<true, true><<<sb, K6V2_THREADS, 0, st>>>(SWEEP_ARGS);
      else if (first)          k_sweep<true, false><<<sb, K6V2_THREADS, 0, st>>>(SWEEP_ARGS);
      else if (last)           k_sweep<false, true><<<sb, K6V2_THREADS, 0, st>>>(SWEEP_ARGS);
      else                     k_sweep<false, false><<<sb, K6V2_THREADS, 0, st>>>(SWEEP_ARGS);
      if (HEAVY_SPLIT) CK(cudaStreamWaitEvent(st, ctx->evJoin, 0));
    }
#undef SWEEP_ARGS
    if (!last) in = out;     // k_pair_emit re-derives the pairs on the LAST sweep's input
  }
  TIME_MARK(ctx, timing, 7);
  k_writeback<<<ctx->wbTiles, WB_THREADS, 0, st>>>(g, ctx->d, ctx->s, ctx->slotOf, ctx->tileCount);
  k_pair_scan<<<1, 1024, 0, st>>>(ctx->tileCount, ctx->tilePrefix, ctx->wbTiles, g.maxPairs, ctx->dCtr, ctx->coll);
  k_pair_emit<<<std::min<uint32_t>(ctx->wbTiles, 148u * 8u), WB_THREADS, 0, st>>>(g, ctx->dParams, ctx->d, ctx->s, in, gs, ctx->slotOf,
                                                                                  ctx->tilePrefix, ctx->dCtr, ctx->coll, (uint32_t)(S - 1), ctx->wbTiles);
  k_physics_end<<<1, 32, 0, st>>>(ctx->dCtr);
  TIME_MARK(ctx, timing, 8);
  CK(cudaGetLastError());
  return WEED_OK;
}

static int launch_frame(weed_ctx* ctx, bool timing) {
  int rc = launch_spatial(ctx, true, timing);
  if (rc) return rc;
  return launch_constraints(ctx, timing);
}

static int ensure_graph(weed_ctx* ctx) {
  if (ctx->frameGraph && ctx->graphSubSteps == ctx->phys.subStepCount) return WEED_OK;
  if (ctx->frameGraph) { cudaGraphExecDestroy(ctx->frameGraph); ctx->frameGraph = nullptr; }
  cudaGraph_t graph = nullptr;
  CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  int rc = launch_frame(ctx, false);
  cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
  if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
  CK(e);
  e = cudaGraphInstantiate(&ctx->frameGraph, graph, 0);
  cudaGraphDestroy(graph);
  CK(e);
  ctx->graphSubSteps = ctx->phys.subStepCount;
  return WEED_OK;
}

static int run_frames(weed_ctx* ctx, double dtRatio, uint32_t frames) {
  int rc = push_params(ctx, dtRatio);
  if (rc) return rc;
  const bool timing = (ctx->cfg.flags & WEED_FLAG_KERNEL_TIMING) != 0;
  const bool direct = timing || (ctx->cfg.flags & WEED_FLAG_NO_GRAPH);
  ctx->launchesPerStep = 18 + ((ctx->cfg.flags & WEED_FLAG_K6_TILE) ? 1u : 2u) * (uint32_t)ctx->phys.subStepCount;
  if (!direct) {
    rc = ensure_graph(ctx);
    if (rc) return rc;
  }
  for (uint32_t f = 0; f < frames; f++) {
    if (direct) {
      rc = launch_frame(ctx, timing);
      if (rc) return rc;
    } else {
      CK(cudaGraphLaunch(ctx->frameGraph, ctx->stream));
    }
  }
  if (timing && frames) {
    CK(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < kTimedSpans; k++) cudaEventElapsedTime(&ctx->ms[k], ctx->ev[k], ctx->ev[k + 1]);
  }
  ctx->spatialValid = false;
  return WEED_OK;
}

extern "C" int weed_spatial(weed_ctx* ctx) {
  GUARD(ctx);
  int rc = launch_spatial(ctx, false, false);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->spatialValid = true;
  return WEED_OK;
}

extern "C" int weed_physics(weed_ctx* ctx, double dtRatio) {
  GUARD(ctx);
  if (!ctx->spatialValid) return fail(ctx, WEED_E_STATE, "weed_physics needs the rows of a preceding weed_spatial of the same frame (or use weed_step)");
  int rc = push_params(ctx, dtRatio);
  if (rc) return rc;
  const GridDims& g = ctx->g;
  k_build_slots<true><<<blocks_for(g.N, 256), 256, 0, ctx->stream>>>(g, ctx->dParams, ctx->phys.subStepCount, true, ctx->d, ctx->s,
                                                                      ctx->key, ctx->cellStart, ctx->arrIds, ctx->slotOf, slab_cuts(ctx));
  k_slots_to_sweep_input<<<blocks_for(g.N, 256), 256, 0, ctx->stream>>>(g, ctx->s, ctx->cellStart);
  rc = launch_constraints(ctx, false);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->spatialValid = false;
  return WEED_OK;
}

// Columns that are final once k_build_slots ran (velocities, derived properties, the zeroed
// accelerations, every pure input): their device->host copies overlap K4..K7 on a second stream.
// x, y, px, py and collisionCount are final only after the write-back.
static constexpr uint32_t kLateCols = WEED_COL_T_X | WEED_COL_T_Y | WEED_COL_RB_PX | WEED_COL_RB_PY | WEED_COL_RB_COLLCNT;
static constexpr uint32_t kAccCols = WEED_COL_RB_AX | WEED_COL_RB_AY;
static constexpr size_t kPipelineMinEntities = 1u << 18;

static int ensure_copy_stream(weed_ctx* ctx) {
  if (ctx->copyStream) return WEED_OK;
  CK(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&ctx->evUp, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ctx->evBuilt, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ctx->evCopied, cudaEventDisableTiming));
  return WEED_OK;
}

// weed_step for large worlds: the frame is launched kernel by kernel so that
//   * an upload of only ax / ay (what tick() writes) runs beside K1-K3, which do not read them,
//   * the early columns leave for the host as soon as k_build_slots is done.
static int step_pipelined(weed_ctx* ctx, double dtRatio, uint32_t upload_mask, uint32_t download_mask) {
  int rc = ensure_copy_stream(ctx);
  if (rc) return rc;
  rc = push_params(ctx, dtRatio);
  if (rc) return rc;
  const uint32_t upCols = upload_mask & WEED_COLS_INPUT_ALL;
  cudaEvent_t waitUp = nullptr;
  if (upCols && !(upCols & ~kAccCols)) {
    CK(cudaEventRecord(ctx->evUp, ctx->stream));                 // order after everything already queued
    CK(cudaStreamWaitEvent(ctx->copyStream, ctx->evUp, 0));
    rc = upload_async(ctx, upCols, ctx->copyStream, ctx->copyCount);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->evUp, ctx->copyStream));
    waitUp = ctx->evUp;
  } else {
    rc = upload_async(ctx, upload_mask, nullptr, ctx->copyCount);
    if (rc) return rc;
  }
  const uint32_t early = download_mask & WEED_COLS_INPUT_ALL & ~kLateCols;
  ctx->launchesPerStep = 18 + ((ctx->cfg.flags & WEED_FLAG_K6_TILE) ? 1u : 2u) * (uint32_t)ctx->phys.subStepCount;
  rc = launch_spatial(ctx, true, false, waitUp, early ? ctx->evBuilt : nullptr);
  if (rc) return rc;
  if (early) {
    CK(cudaStreamWaitEvent(ctx->copyStream, ctx->evBuilt, 0));
    rc = download_async(ctx, early, ctx->copyStream, ctx->copyCount);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->evCopied, ctx->copyStream));
  }
  rc = launch_constraints(ctx, false);
  if (rc) return rc;
  ctx->spatialValid = false;
  rc = download_async(ctx, download_mask & ~early, nullptr, ctx->copyCount);
  if (rc) return rc;
  if (early) CK(cudaStreamWaitEvent(ctx->stream, ctx->evCopied, 0));
  CK(cudaStreamSynchronize(ctx->stream));
  return WEED_OK;
}

extern "C" int weed_step(weed_ctx* ctx, double dtRatio, uint32_t upload_mask, uint32_t download_mask) {
  GUARD(ctx);
  const bool direct = (ctx->cfg.flags & (WEED_FLAG_KERNEL_TIMING | WEED_FLAG_NO_GRAPH)) != 0;
  const bool hasCopies = ((upload_mask | download_mask) & WEED_COLS_INPUT_ALL) != 0;
  ctx->copyCount = 0;
  if (ctx->slab && hasCopies) {
    // a slab's table is used up to `top`: the columns beyond it hold nothing (weed_upload / weed_download move
    // the whole table; the per-frame call moves what exists).  One 4-byte read per step.
    uint32_t top = 0;
    CK(cudaMemcpyAsync(&top, &ctx->dSlab->top, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->copyCount = std::max<size_t>(1, std::min<size_t>(top, ctx->g.N));
  }
  if (!direct && hasCopies && ctx->g.N >= kPipelineMinEntities) return step_pipelined(ctx, dtRatio, upload_mask, download_mask);
  int rc = upload_async(ctx, upload_mask, nullptr, ctx->copyCount);
  if (rc) return rc;
  rc = run_frames(ctx, dtRatio, 1);
  if (rc) return rc;
  rc = download_async(ctx, download_mask, nullptr, ctx->copyCount);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return WEED_OK;
}

extern "C" int weed_run(weed_ctx* ctx, double dtRatio, uint32_t frames) {
  GUARD(ctx);
  return run_frames(ctx, dtRatio, frames);
}

extern "C" uint32_t weed_entity_count(weed_ctx* ctx) { return ctx ? ctx->g.N : 0u; }

extern "C" int weed_sync(weed_ctx* ctx) {
  GUARD(ctx);
  CK(cudaStreamSynchronize(ctx->stream));
  return WEED_OK;
}

extern "C" int weed_set_physics(weed_ctx* ctx, const weed_physics_config* p) {
  GUARD(ctx);
  if (!p) return WEED_E_INVALID;
  validate_physics(&ctx->phys, p);
  ctx->paramsDirty = true;
  return WEED_OK;
}

extern "C" int weed_get_physics(weed_ctx* ctx, weed_physics_config* out) {
  if (!ctx || !out) return WEED_E_INVALID;
  *out = ctx->phys;
  return WEED_OK;
}

extern "C" int weed_fetch_neighbors(weed_ctx* ctx, uint32_t first, uint32_t count) {
  GUARD(ctx);
  if (!ctx->nd) return fail(ctx, WEED_E_STATE, "neighbor rows disabled (WEED_FLAG_NO_NEIGHBOR_ROWS)");
  if (!ctx->host[WEED_BUF_NEIGHBOR] || !ctx->host[WEED_BUF_DISTANCE]) return fail(ctx, WEED_E_NOT_BOUND, "neighbor/distance buffer not bound");
  if ((size_t)first + count > ctx->g.N) return fail(ctx, WEED_E_INVALID, "row range out of bounds");
  int rc = copy_rows(ctx, ctx->host[WEED_BUF_NEIGHBOR], ctx->host[WEED_BUF_DISTANCE], first, count, true, ctx->stream);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return WEED_OK;
}

extern "C" int weed_fetch_neighbors_to(weed_ctx* ctx, uint32_t first, uint32_t count, int32_t* neighbor_out, float* distance_out) {
  GUARD(ctx);
  if (!ctx->nd) return fail(ctx, WEED_E_STATE, "neighbor rows disabled (WEED_FLAG_NO_NEIGHBOR_ROWS)");
  if ((size_t)first + count > ctx->g.N) return fail(ctx, WEED_E_INVALID, "row range out of bounds");
  int rc = copy_rows(ctx, neighbor_out, distance_out, first, count, false, ctx->stream);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return WEED_OK;
}

extern "C" int weed_get_stats(weed_ctx* ctx, weed_stats* out) {
  GUARD(ctx);
  if (!out) return WEED_E_INVALID;
  memset(out, 0, sizeof(*out));
  CK(cudaMemsetAsync(&ctx->dCtr->neighborsTotal, 0, sizeof(unsigned long long), ctx->stream));
  CK(cudaMemsetAsync(&ctx->dCtr->cappedRows, 0, sizeof(uint32_t), ctx->stream));
  k_stats<<<296, 256, 0, ctx->stream>>>(ctx->g, ctx->s.NCNT, ctx->cellStart, ctx->dCtr);
  Counters c;
  CK(cudaMemcpyAsync(&c, ctx->dCtr, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  out->frames = c.frames;
  out->gridCols = (uint32_t)ctx->g.cols; out->gridRows = (uint32_t)ctx->g.rows;
  out->activeInGrid = c.activeInGrid;
  out->maxCellOccupancy = c.maxCellFrame;
  out->neighborsTotal = c.neighborsTotal;
  out->cappedRows = c.cappedRows;
  out->explicitPairs = c.explicitPairs;
  out->collisionPairs = c.collisionPairs;
  out->kernelLaunchesPerStep = ctx->launchesPerStep ? ctx->launchesPerStep : 18 + 2 * (uint32_t)ctx->phys.subStepCount;
  memcpy(out->ms, ctx->ms, sizeof(out->ms));
  out->ms[8] = (float)c.frameNs * 1e-6f;   // device clock, k_spatial_begin -> k_physics_end of the last frame
  out->ms[9] = (float)c.xoverRows;         // a count, not a time: capped rows whose lost partners overflowed the internal row
  return WEED_OK;
}

extern "C" uint32_t weed_row_pitch(weed_ctx* ctx) { return ctx ? ctx->g.Npad : 0u; }

extern "C" int weed_device_ptr(weed_ctx* ctx, weed_devptr_id which, void** out, size_t* bytes) {
  if (!ctx || !out) return WEED_E_INVALID;
  const size_t N = ctx->g.N;
  void* p = nullptr; size_t b = 0;
  switch (which) {
    case WEED_DEV_NEIGHBOR: p = ctx->nd; b = ctx->rowWords * 4; break;
    case WEED_DEV_DISTANCE: p = ctx->dd; b = ctx->rowWords * 4; break;
    case WEED_DEV_COLLISION: p = ctx->coll; b = (1 + 2 * (size_t)ctx->g.maxPairs) * 4; break;
    case WEED_DEV_STATE: p = ctx->d.DP; b = N * 16; break;
    case WEED_DEV_ATTR: p = ctx->d.AT; b = N * 16; break;
    case WEED_DEV_VEL: p = ctx->d.V; b = N * 16; break;
    case WEED_DEV_NEIGHBOR_COUNT: p = ctx->s.NCNT; b = N * 4; break;
    case WEED_DEV_SLOT_OF: p = ctx->slotOf; b = N * 4; break;
    default: return fail(ctx, WEED_E_INVALID, "bad devptr id");
  }
  *out = p;
  if (bytes) *bytes = b;
  return WEED_OK;
}

// =============================================================================================
// slabs
// =============================================================================================
extern "C" int weed_slab_set_gids(weed_ctx* ctx, const uint32_t* gids, uint32_t count) {
  GUARD(ctx);
  if (!ctx->slab) return fail(ctx, WEED_E_STATE, "not a slab context (slabRowEnd == 0)");
  if (!gids || count > ctx->g.N) return fail(ctx, WEED_E_INVALID, "bad gid list");
  for (uint32_t k = 0; k < count; k++)      // bit 31 of an id word is the CX_EDGE flag of the candidate records
    if (gids[k] & 0x80000000u) return fail(ctx, WEED_E_INVALID, "global entity ids must be below 2^31");
  CK(cudaMemsetAsync(ctx->d.GID, 0xFF, (size_t)ctx->g.N * 4, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d.GID, gids, (size_t)count * 4, cudaMemcpyHostToDevice, ctx->stream));
  SlabCounters sc;
  memset(&sc, 0, sizeof(sc));
  sc.top = count;
  sc.curBegin = sc.pendBegin = ctx->g.slabBegin;
  sc.curEnd = sc.pendEnd = ctx->g.slabEnd;
  sc.minRows = (uint32_t)std::max(2 * ctx->g.slabHalo, 8);
  CK(cudaMemcpyAsync(ctx->dSlab, &sc, sizeof(sc), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return WEED_OK;
}

extern "C" int weed_slab_get_gids(weed_ctx* ctx, uint32_t* gids_out, uint32_t* top_out) {
  GUARD(ctx);
  if (!ctx->slab) return fail(ctx, WEED_E_STATE, "not a slab context (slabRowEnd == 0)");
  SlabCounters sc;
  CK(cudaMemcpyAsync(&sc, ctx->dSlab, sizeof(sc), cudaMemcpyDeviceToHost, ctx->stream));
  if (gids_out) CK(cudaMemcpyAsync(gids_out, ctx->d.GID, (size_t)ctx->g.N * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (top_out) *top_out = sc.top;
  return WEED_OK;
}

extern "C" int weed_slab_pack(weed_ctx* ctx, void* dev_low, void* dev_high, uint32_t quota) {
  GUARD(ctx);
  if (!ctx->slab) return fail(ctx, WEED_E_STATE, "not a slab context (slabRowEnd == 0)");
  if (!dev_low || !dev_high || quota == 0) return fail(ctx, WEED_E_INVALID, "exchange buffers missing");
  k_slab_pack<<<blocks_for(ctx->g.N, 256), 256, 0, ctx->stream>>>(ctx->g, ctx->d, ctx->key, (SlabRec*)dev_low, (SlabRec*)dev_high, quota, ctx->dSlab,
                                                                  nullptr, ctx->phys.subStepCount);
  k_slab_headers<<<1, 32, 0, ctx->stream>>>((SlabRec*)dev_low, (SlabRec*)dev_high, quota, ctx->dSlab, ctx->dCtr, nullptr);
  CK(cudaGetLastError());
  return WEED_OK;
}

extern "C" int weed_slab_apply(weed_ctx* ctx, const void* dev_from_low, const void* dev_from_high, uint32_t quota) {
  GUARD(ctx);
  if (!ctx->slab) return fail(ctx, WEED_E_STATE, "not a slab context (slabRowEnd == 0)");
  if (quota == 0) return fail(ctx, WEED_E_INVALID, "quota must be positive");
  k_slab_drop<<<blocks_for(ctx->g.N, 256), 256, 0, ctx->stream>>>(ctx->g, ctx->d, ctx->key, ctx->holes, ctx->dSlab);
  k_slab_unpack<<<blocks_for(2 * (size_t)quota, 256), 256, 0, ctx->stream>>>(ctx->d, (const SlabRec*)dev_from_low, (const SlabRec*)dev_from_high,
                                                                           quota, ctx->holes, ctx->g.N, ctx->dSlab, nullptr);
  k_slab_finish<<<1, 32, 0, ctx->stream>>>((const SlabRec*)dev_from_low, (const SlabRec*)dev_from_high, quota, ctx->g.N, ctx->dSlab, nullptr);
  CK(cudaGetLastError());
  ctx->spatialValid = false;
  return WEED_OK;
}

// ---- peer-to-peer exchange: no host, no library call inside a frame -------------------------------------
static size_t xfer_offset(uint32_t quota, int side, int parity) { return (size_t)(side * 2 + parity) * ((size_t)quota + 1); }

extern "C" int weed_slab_exchange_create(weed_ctx* ctx, uint32_t quota) {
  GUARD(ctx);
  if (!ctx->slab) return fail(ctx, WEED_E_STATE, "not a slab context (slabRowEnd == 0)");
  if (quota == 0) return fail(ctx, WEED_E_INVALID, "quota must be positive");
  if (ctx->xRecv) return fail(ctx, WEED_E_STATE, "exchange buffers already exist");
  int rc = dalloc(ctx, &ctx->xRecv, 4 * ((size_t)quota + 1));          // zeroed: no arrival flag is set
  if (rc) return rc;
  rc = dalloc(ctx, &ctx->dXfer, 1);
  if (rc) return rc;
  ctx->xQuota = quota;
  memset(&ctx->hXfer, 0, sizeof(ctx->hXfer));
  ctx->hXfer.quota = quota;
  for (int side = 0; side < 2; side++)
    for (int par = 0; par < 2; par++) ctx->hXfer.recv[side][par] = ctx->xRecv + xfer_offset(quota, side, par);
  CK(cudaMemcpyAsync(ctx->dXfer, &ctx->hXfer, sizeof(SlabXfer), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return WEED_OK;
}

extern "C" int weed_slab_exchange_export(weed_ctx* ctx, weed_ipc_handle* handle_out, void** base_out) {
  GUARD(ctx);
  if (!ctx->xRecv) return fail(ctx, WEED_E_STATE, "weed_slab_exchange_create has not been called");
  static_assert(sizeof(weed_ipc_handle) == sizeof(cudaIpcMemHandle_t), "weed_ipc_handle carries a cudaIpcMemHandle_t");
  if (handle_out) CK(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle_out), ctx->xRecv));
  if (base_out) *base_out = ctx->xRecv;
  return WEED_OK;
}

extern "C" int weed_slab_exchange_connect(weed_ctx* ctx, int side, const weed_ipc_handle* peer_handle, void* peer_base) {
  GUARD(ctx);
  if (!ctx->xRecv) return fail(ctx, WEED_E_STATE, "weed_slab_exchange_create has not been called");
  if (side < 0 || side > 1 || (!peer_handle && !peer_base)) return fail(ctx, WEED_E_INVALID, "bad side / no peer buffer");
  SlabRec* base = (SlabRec*)peer_base;
  if (peer_handle) {             // another process: map its allocation (peer access over NVLink is enabled lazily)
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, *reinterpret_cast<const cudaIpcMemHandle_t*>(peer_handle), cudaIpcMemLazyEnablePeerAccess));
    ctx->xPeerIpc[side] = p;
    base = (SlabRec*)p;
  } else {                       // same process: a context on this or another device
    cudaPointerAttributes at;
    CK(cudaPointerGetAttributes(&at, peer_base));
    if (at.device != ctx->device) {
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, ctx->device, at.device));
      if (!can) return fail(ctx, WEED_E_CUDA, "no peer access between devices " + std::to_string(ctx->device) + " and " + std::to_string(at.device));
      cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
      cudaGetLastError();
    }
  }
  // what I send to my LOW neighbour arrives in ITS "from high" buffers, and the other way round
  for (int par = 0; par < 2; par++) ctx->hXfer.send[side][par] = base + xfer_offset(ctx->xQuota, 1 - side, par);
  CK(cudaMemcpyAsync(ctx->dXfer, &ctx->hXfer, sizeof(SlabXfer), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return WEED_OK;
}

extern "C" int weed_slab_exchange_disconnect(weed_ctx* ctx) {
  GUARD(ctx);
  CK(cudaStreamSynchronize(ctx->stream));
  for (int side = 0; side < 2; side++) {
    if (ctx->xPeerIpc[side]) { cudaIpcCloseMemHandle(ctx->xPeerIpc[side]); ctx->xPeerIpc[side] = nullptr; }
    ctx->hXfer.send[side][0] = ctx->hXfer.send[side][1] = nullptr;
  }
  if (ctx->dXfer) {
    CK(cudaMemcpyAsync(ctx->dXfer, &ctx->hXfer, sizeof(SlabXfer), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return WEED_OK;
}

static int slab_frame_begin(weed_ctx* ctx, double dtRatio, bool runFrame) {
  if (!ctx->xRecv) return fail(ctx, WEED_E_STATE, "weed_slab_exchange_create has not been called");
  if (runFrame) {
    int rc = run_frames(ctx, dtRatio, 1);
    if (rc) return rc;
  }
  k_slab_pack<<<blocks_for(ctx->g.N, 256), 256, 0, ctx->stream>>>(ctx->g, ctx->d, ctx->key, nullptr, nullptr, 0, ctx->dSlab, ctx->dXfer,
                                                                  ctx->phys.subStepCount);
  k_slab_headers<<<1, 32, 0, ctx->stream>>>(nullptr, nullptr, 0, ctx->dSlab, ctx->dCtr, ctx->dXfer);
  CK(cudaGetLastError());
  return WEED_OK;
}

static int slab_frame_end(weed_ctx* ctx) {
  const uint32_t quota = ctx->xQuota;
  k_slab_wait<<<1, 32, 0, ctx->stream>>>(ctx->dXfer, ctx->dSlab);
  k_slab_drop<<<blocks_for(ctx->g.N, 256), 256, 0, ctx->stream>>>(ctx->g, ctx->d, ctx->key, ctx->holes, ctx->dSlab);
  k_slab_unpack<<<blocks_for(2 * (size_t)quota, 256), 256, 0, ctx->stream>>>(ctx->d, nullptr, nullptr, quota, ctx->holes, ctx->g.N, ctx->dSlab, ctx->dXfer);
  k_slab_finish<<<1, 32, 0, ctx->stream>>>(nullptr, nullptr, quota, ctx->g.N, ctx->dSlab, ctx->dXfer);
  CK(cudaGetLastError());
  ctx->spatialValid = false;
  return WEED_OK;
}

extern "C" int weed_slab_frame(weed_ctx* ctx, double dtRatio) {
  GUARD(ctx);
  if (!ctx->slab) return fail(ctx, WEED_E_STATE, "not a slab context (slabRowEnd == 0)");
  int rc = slab_frame_begin(ctx, dtRatio, true);
  if (rc) return rc;
  return slab_frame_end(ctx);
}

extern "C" int weed_slab_exchange(weed_ctx* ctx) {          // the exchange alone (after a weed_step of the slab)
  GUARD(ctx);
  if (!ctx->slab) return fail(ctx, WEED_E_STATE, "not a slab context (slabRowEnd == 0)");
  int rc = slab_frame_begin(ctx, 1.0, false);
  if (rc) return rc;
  return slab_frame_end(ctx);
}

extern "C" int weed_slab_frame_begin(weed_ctx* ctx, double dtRatio) {
  GUARD(ctx);
  if (!ctx->slab) return fail(ctx, WEED_E_STATE, "not a slab context (slabRowEnd == 0)");
  return slab_frame_begin(ctx, dtRatio, true);
}

extern "C" int weed_slab_frame_end(weed_ctx* ctx) {
  GUARD(ctx);
  if (!ctx->slab || !ctx->xRecv) return fail(ctx, WEED_E_STATE, "no peer-to-peer exchange on this context");
  return slab_frame_end(ctx);
}

extern "C" int weed_slab_balance(weed_ctx* ctx, uint32_t maxShiftRows, uint32_t hysteresisPercent) {
  GUARD(ctx);
  if (!ctx->slab) return fail(ctx, WEED_E_STATE, "not a slab context (slabRowEnd == 0)");
  if (maxShiftRows > 8) return fail(ctx, WEED_E_INVALID, "maxShiftRows > 8 (the exchange quota is sized for the halo band)");
  const uint32_t v[2] = {maxShiftRows, hysteresisPercent ? hysteresisPercent : 3u};
  CK(cudaMemcpyAsync(&ctx->dSlab->maxShift, v, sizeof(v), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return WEED_OK;
}

extern "C" int weed_slab_status(weed_ctx* ctx, weed_slab_stats* out) {
  GUARD(ctx);
  if (!ctx->slab) return fail(ctx, WEED_E_STATE, "not a slab context (slabRowEnd == 0)");
  if (!out) return WEED_E_INVALID;
  SlabCounters sc;
  CK(cudaMemcpyAsync(&sc, ctx->dSlab, sizeof(sc), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  out->top = sc.top; out->owned = sc.lastOwned; out->sentLow = sc.lastLow; out->sentHigh = sc.lastHigh;
  out->receivedLow = sc.lastFromLow; out->receivedHigh = sc.lastFromHigh; out->overflow = sc.overflow;
  out->capacity = ctx->g.N;
  out->rowBegin = sc.curBegin; out->rowEnd = sc.curEnd; out->cutMoves = sc.cutMoves; out->loadNs = sc.load;
  if (sc.overflow & 1u) return fail(ctx, WEED_E_OVERFLOW, "slab exchange quota exceeded: " + std::to_string(sc.lastLow) + " / " + std::to_string(sc.lastHigh) + " records");
  if (sc.overflow & 2u) return fail(ctx, WEED_E_OVERFLOW, "slab entity table full (capacity " + std::to_string(ctx->g.N) + ")");
  if (sc.overflow & SLAB_OVF_REACH) return fail(ctx, WEED_E_OVERFLOW, "an entity whose reach exceeds the halo (" + std::to_string(ctx->g.slabHalo) + " rows) came near a cut: recreate the slabs with a deeper halo");
  if (sc.overflow & SLAB_OVF_TIMEOUT) return fail(ctx, WEED_E_OVERFLOW, "a neighbour's message did not arrive within two seconds");
  return WEED_OK;
}

// ---- a whole world on several GPUs of ONE process (the Node host of INTEGRATION.md) -----------------------
struct weed_group {
  std::vector<weed_ctx*> slabs;
  std::string err;
};
thread_local std::string g_group_error;

extern "C" const char* weed_group_last_error(weed_group* g) { return g ? g->err.c_str() : g_group_error.c_str(); }

extern "C" void weed_group_destroy(weed_group* g) {
  if (!g) return;
  for (weed_ctx* c : g->slabs) weed_destroy(c);
  delete g;
}

extern "C" int weed_group_create(const weed_config* tmpl, uint32_t nSlabs, const int32_t* devices, const uint32_t* rowCuts,
                                 uint32_t haloRows, const uint32_t* capacities, uint32_t quota, weed_group** out) {
  if (out) *out = nullptr;
  if (!tmpl || !out || nSlabs == 0 || !rowCuts || !capacities || quota == 0) { g_group_error = "null / empty argument"; return WEED_E_INVALID; }
  weed_group* g = new weed_group();
  auto bail = [&](int code, const std::string& why) { g_group_error = why; weed_group_destroy(g); return code; };
  for (uint32_t k = 0; k < nSlabs; k++) {
    weed_config c = *tmpl;
    c.entityCount = capacities[k];
    c.device = devices ? devices[k] : tmpl->device;
    c.stream = nullptr;                                  // every slab runs on a stream of its own
    c.slabRowBegin = rowCuts[k]; c.slabRowEnd = rowCuts[k + 1]; c.slabHaloRows = haloRows;
    weed_ctx* ctx = nullptr;
    int rc = weed_create(&c, &ctx);
    if (rc) return bail(rc, "slab " + std::to_string(k) + ": " + weed_last_error(nullptr));
    g->slabs.push_back(ctx);
    rc = weed_slab_exchange_create(ctx, quota);
    if (rc) return bail(rc, "slab " + std::to_string(k) + ": " + weed_last_error(ctx));
  }
  for (uint32_t k = 0; k < nSlabs; k++) {                // neighbours write straight into each other's buffers
    if (k > 0) {
      int rc = weed_slab_exchange_connect(g->slabs[k], 0, nullptr, g->slabs[k - 1]->xRecv);
      if (rc) return bail(rc, "slab " + std::to_string(k) + ": " + weed_last_error(g->slabs[k]));
    }
    if (k + 1 < nSlabs) {
      int rc = weed_slab_exchange_connect(g->slabs[k], 1, nullptr, g->slabs[k + 1]->xRecv);
      if (rc) return bail(rc, "slab " + std::to_string(k) + ": " + weed_last_error(g->slabs[k]));
    }
  }
  *out = g;
  return WEED_OK;
}

extern "C" uint32_t weed_group_size(weed_group* g) { return g ? (uint32_t)g->slabs.size() : 0u; }
extern "C" weed_ctx* weed_group_slab(weed_group* g, uint32_t k) { return (g && k < g->slabs.size()) ? g->slabs[k] : nullptr; }

extern "C" int weed_group_step(weed_group* g, double dtRatio) {
  if (!g) return WEED_E_INVALID;
  for (size_t k = 0; k < g->slabs.size(); k++) {          // queue every slab's frame + pack first ...
    int rc = weed_slab_frame_begin(g->slabs[k], dtRatio);
    if (rc) { g->err = "slab " + std::to_string(k) + ": " + weed_last_error(g->slabs[k]); return rc; }
  }
  for (size_t k = 0; k < g->slabs.size(); k++) {          // ... then the waits: every message they wait for is queued
    int rc = weed_slab_frame_end(g->slabs[k]);
    if (rc) { g->err = "slab " + std::to_string(k) + ": " + weed_last_error(g->slabs[k]); return rc; }
  }
  return WEED_OK;
}

extern "C" int weed_group_sync(weed_group* g) {
  if (!g) return WEED_E_INVALID;
  for (size_t k = 0; k < g->slabs.size(); k++) {
    weed_slab_stats st;
    int rc = weed_slab_status(g->slabs[k], &st);         // synchronises; reports quota / table / reach / timeout
    if (rc) { g->err = "slab " + std::to_string(k) + ": " + weed_last_error(g->slabs[k]); return rc; }
  }
  return WEED_OK;
}

// =============================================================================================
// device-side systems
// =============================================================================================
static int run_flock(weed_ctx* ctx, const FlockParams& fp, const float* protectedRange) {
  if (!ctx->nd) return fail(ctx, WEED_E_STATE, "neighbor rows disabled (WEED_FLAG_NO_NEIGHBOR_ROWS)");
  if (ctx->slab) return fail(ctx, WEED_E_STATE, "systems are not available on slab contexts yet");
  const size_t N = ctx->g.N;
  if (protectedRange) {
    if (!ctx->protRange) { int rc = dalloc(ctx, &ctx->protRange, N); if (rc) return rc; }
    CK(cudaMemcpyAsync(ctx->protRange, protectedRange, N * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  k_system_flock<<<blocks_for(N, 128), 128, 0, ctx->stream>>>(ctx->g, fp, ctx->d, row_view(ctx),
                                                              protectedRange ? ctx->protRange : nullptr);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return WEED_OK;
}

extern "C" int weed_system_boids(weed_ctx* ctx, const weed_boids_params* p, const float* protectedRange, double dtRatio) {
  GUARD(ctx);
  if (!p) return WEED_E_INVALID;
  FlockParams fp{};
  fp.nClasses = 1; fp.mouseType = p->mouseEntityType; fp.mouseDown = 0; fp.dtRatio = dtRatio;
  FlockClass& c = fp.cls[0];
  c.type = FLOCK_ANY_TYPE; c.role = 0; c.other = 0; c.prScale = 2.0;     // boid.js:64
  c.centering = p->centeringFactor; c.avoid = p->avoidFactor; c.matching = p->matchingFactor;
  c.turn = p->turnFactor; c.margin = p->margin; c.roleFactor = 0;
  return run_flock(ctx, fp, protectedRange);
}

extern "C" int weed_system_flock(weed_ctx* ctx, const weed_flock_class* classes, uint32_t classCount,
                                 const weed_flock_params* params, const float* protectedRange) {
  GUARD(ctx);
  if (!classes || !params || classCount == 0 || classCount > FLOCK_MAX_CLASSES)
    return fail(ctx, WEED_E_INVALID, "weed_system_flock: 1.." + std::to_string(FLOCK_MAX_CLASSES) + " classes");
  FlockParams fp{};
  fp.nClasses = classCount; fp.mouseType = params->mouseEntityType; fp.mouseDown = params->mouseDown ? 1u : 0u;
  fp.dtRatio = params->dtRatio;
  for (uint32_t k = 0; k < classCount; k++) {
    if (classes[k].role > WEED_FLOCK_PREDATOR) return fail(ctx, WEED_E_INVALID, "weed_system_flock: bad role");
    FlockClass& c = fp.cls[k];
    c.type = classes[k].entityType; c.role = classes[k].role; c.other = classes[k].otherEntityType;
    c.prScale = classes[k].protectedRangeScale; c.centering = classes[k].centeringFactor; c.avoid = classes[k].avoidFactor;
    c.matching = classes[k].matchingFactor; c.turn = classes[k].turnFactor; c.margin = classes[k].margin;
    c.roleFactor = classes[k].roleFactor;
  }
  return run_flock(ctx, fp, protectedRange);
}

// tile scratch shared by the ordered-compaction passes of the systems
static int sys_tiles(weed_ctx* ctx, size_t elements) {
  const size_t tiles = (elements + SYS_TILE - 1) / SYS_TILE + 1;
  if (tiles <= ctx->sysTiles) return WEED_OK;
  int rc = dalloc(ctx, &ctx->sysTileCount, tiles); if (rc) return rc;
  rc = dalloc(ctx, &ctx->sysTilePrefix, tiles); if (rc) return rc;
  ctx->sysTiles = tiles;
  return WEED_OK;
}

extern "C" int weed_system_collision_events(weed_ctx* ctx, uint32_t flags, weed_collision_event_counts* counts,
                                            uint8_t* state, int32_t* exitData) {
  GUARD(ctx);
  if (ctx->slab) return fail(ctx, WEED_E_STATE, "systems are not available on slab contexts yet");
  const size_t maxPairs = ctx->g.maxPairs;
  cudaStream_t st = ctx->stream;
  if (!ctx->dEv) {
    size_t cap = 1024;
    while (cap < 2 * maxPairs) cap <<= 1;
    if (cap > (1ull << 31)) return fail(ctx, WEED_E_INVALID, "maxCollisionPairs too large for the event tables");
    ctx->evMask = (uint32_t)(cap - 1);
    for (int t = 0; t < 2; t++) {
      int rc = dalloc(ctx, &ctx->evTable[t], cap, false); if (rc) return rc;
      rc = dalloc(ctx, &ctx->evList[t], maxPairs); if (rc) return rc;
    }
    int rc = dalloc(ctx, &ctx->evState, maxPairs); if (rc) return rc;
    rc = dalloc(ctx, &ctx->evExit, 1 + 2 * maxPairs); if (rc) return rc;
    rc = dalloc(ctx, &ctx->dEv, 1); if (rc) return rc;
    rc = sys_tiles(ctx, maxPairs); if (rc) return rc;
    ctx->evCur = 0;
  }
  const uint32_t cur = ctx->evCur, prev = cur ^ 1u;
  const size_t cap = (size_t)ctx->evMask + 1;
  const unsigned blocks = (unsigned)((maxPairs + SYS_TILE - 1) / SYS_TILE);
  CK(cudaMemsetAsync(ctx->evTable[cur], 0xFF, cap * 8, st));
  k_ev_begin<<<1, 32, 0, st>>>(ctx->coll, ctx->dEv, flags & WEED_EVENTS_FORGET_PREVIOUS);
  if (blocks) {
    k_ev_insert<<<blocks, SYS_TILE, 0, st>>>(ctx->coll, ctx->evTable[cur], ctx->evMask, ctx->evList[cur], ctx->dEv);
    k_ev_classify<<<blocks, SYS_TILE, 0, st>>>(ctx->evList[cur], ctx->evTable[cur], ctx->evList[prev], ctx->evTable[prev],
                                               ctx->evMask, ctx->evState, ctx->sysTileCount, ctx->dEv);
    k_tile_scan<<<1, 1024, 0, st>>>(ctx->sysTileCount, ctx->sysTilePrefix, blocks, &ctx->dEv->exit);
    k_ev_exits<<<blocks, SYS_TILE, 0, st>>>(ctx->evTable[cur], ctx->evList[prev], ctx->evMask, ctx->sysTilePrefix,
                                            ctx->evExit, ctx->dEv);
  }
  EvCounters ec;
  CK(cudaMemcpyAsync(&ec, ctx->dEv, sizeof(ec), cudaMemcpyDeviceToHost, st));
  k_ev_end<<<1, 32, 0, st>>>(ctx->dEv, ctx->evExit);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  ctx->evCur = prev;
  if (counts) { counts->pairs = ec.cur; counts->entered = ec.enter; counts->stayed = ec.stay; counts->exited = ec.exit; }
  if (state && ec.cur) CK(cudaMemcpyAsync(state, ctx->evState, ec.cur, cudaMemcpyDeviceToHost, st));
  if (exitData) {
    exitData[0] = (int32_t)ec.exit;
    if (ec.exit) CK(cudaMemcpyAsync(exitData + 1, ctx->evExit + 1, (size_t)ec.exit * 8, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  return WEED_OK;
}

extern "C" int weed_system_screen_visibility(weed_ctx* ctx, const weed_camera* cam, float* screenX, float* screenY,
                                             uint8_t* isItOnScreen) {
  GUARD(ctx);
  if (!cam) return WEED_E_INVALID;
  if (ctx->slab) return fail(ctx, WEED_E_STATE, "systems are not available on slab contexts yet");
  const size_t N = ctx->g.N;
  cudaStream_t st = ctx->stream;
  if (!ctx->onScreen) {
    int rc = dalloc(ctx, &ctx->screenX, N); if (rc) return rc;
    rc = dalloc(ctx, &ctx->screenY, N); if (rc) return rc;
    rc = dalloc(ctx, &ctx->onScreen, N); if (rc) return rc;
  }
  CameraParams c{cam->zoom, cam->cameraX, cam->cameraY, cam->canvasWidth, cam->canvasHeight};
  k_screen_visibility<<<blocks_for(N, 256), 256, 0, st>>>((uint32_t)N, c, ctx->d.DP, ctx->d.F, ctx->screenX, ctx->screenY, ctx->onScreen);
  CK(cudaGetLastError());
  if (screenX) CK(cudaMemcpyAsync(screenX, ctx->screenX, N * 4, cudaMemcpyDeviceToHost, st));
  if (screenY) CK(cudaMemcpyAsync(screenY, ctx->screenY, N * 4, cudaMemcpyDeviceToHost, st));
  if (isItOnScreen) CK(cudaMemcpyAsync(isItOnScreen, ctx->onScreen, N, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return WEED_OK;
}

extern "C" int weed_system_shadows_upload(weed_ctx* ctx, const weed_shadow_columns* cols) {
  GUARD(ctx);
  if (!cols || !cols->lightActive || !cols->lightIntensity || !cols->casterActive || !cols->shadowRadius || !cols->height)
    return WEED_E_INVALID;
  const size_t N = ctx->g.N;
  if (!ctx->shLightActive) {
    int rc = dalloc(ctx, &ctx->shLightActive, N); if (rc) return rc;
    rc = dalloc(ctx, &ctx->shCasterActive, N); if (rc) return rc;
    rc = dalloc(ctx, &ctx->shIntensity, N); if (rc) return rc;
    rc = dalloc(ctx, &ctx->shRadius, N); if (rc) return rc;
    rc = dalloc(ctx, &ctx->shHeight, N); if (rc) return rc;
    rc = dalloc(ctx, &ctx->shScalars, 4); if (rc) return rc;
  }
  cudaStream_t st = ctx->stream;
  CK(cudaMemcpyAsync(ctx->shLightActive, cols->lightActive, N, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->shCasterActive, cols->casterActive, N, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->shIntensity, cols->lightIntensity, N * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->shRadius, cols->shadowRadius, N * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->shHeight, cols->height, N * 4, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  return WEED_OK;
}

extern "C" int weed_system_shadows(weed_ctx* ctx, uint32_t maxShadowCastingLights, uint32_t maxShadowsPerLight,
                                   uint32_t maxShadowSprites, const weed_shadow_sprites* out, uint32_t* spriteCount) {
  GUARD(ctx);
  if (ctx->slab) return fail(ctx, WEED_E_STATE, "systems are not available on slab contexts yet");
  if (!ctx->nd) return fail(ctx, WEED_E_STATE, "neighbor rows disabled (WEED_FLAG_NO_NEIGHBOR_ROWS)");
  if (!ctx->shLightActive) return fail(ctx, WEED_E_STATE, "weed_system_shadows_upload has not been called");
  if (!ctx->onScreen) return fail(ctx, WEED_E_STATE, "weed_system_screen_visibility has not been called");
  const size_t N = ctx->g.N;
  cudaStream_t st = ctx->stream;
  if (maxShadowSprites > ctx->shOutCap) {
    int rc = dalloc(ctx, &ctx->shOutActive, maxShadowSprites); if (rc) return rc;
    rc = dalloc(ctx, &ctx->shOut, 7 * (size_t)maxShadowSprites); if (rc) return rc;
    ctx->shOutCap = maxShadowSprites;
  }
  if (maxShadowCastingLights > ctx->shLightCap) {
    int rc = dalloc(ctx, &ctx->shLightId, maxShadowCastingLights); if (rc) return rc;
    rc = dalloc(ctx, &ctx->shLightCount, maxShadowCastingLights); if (rc) return rc;
    ctx->shLightCap = maxShadowCastingLights;
  }
  int rc = sys_tiles(ctx, N); if (rc) return rc;
  ShadowParams p{maxShadowCastingLights, maxShadowsPerLight, maxShadowSprites, 0u};
  ShadowIn in{ctx->shLightActive, ctx->shIntensity, ctx->shCasterActive, ctx->shRadius, ctx->shHeight, ctx->onScreen};
  const size_t S = maxShadowSprites;
  ShadowOut o{ctx->shOutActive, ctx->shOut, ctx->shOut + S, ctx->shOut + 2 * S, ctx->shOut + 3 * S,
              ctx->shOut + 4 * S, ctx->shOut + 5 * S, ctx->shOut + 6 * S};
  const unsigned blocks = (unsigned)((N + SYS_TILE - 1) / SYS_TILE);
  k_shadow_lights<<<blocks, SYS_TILE, 0, st>>>((uint32_t)N, in, ctx->d.F, ctx->sysTileCount);
  k_tile_scan<<<1, 1024, 0, st>>>(ctx->sysTileCount, ctx->sysTilePrefix, blocks, ctx->shScalars);
  k_shadow_count<<<blocks, SYS_TILE, 0, st>>>((uint32_t)N, p, in, ctx->d.F, ctx->sysTilePrefix, row_view(ctx),
                                              ctx->shLightId, ctx->shLightCount);
  k_shadow_emit<<<1, 256, 0, st>>>(p, in, ctx->d.F, ctx->d.DP, ctx->shScalars, row_view(ctx), ctx->shLightId,
                                   ctx->shLightCount, o, ctx->shScalars + 1);
  CK(cudaGetLastError());
  uint32_t n = 0;
  CK(cudaMemcpyAsync(&n, ctx->shScalars + 1, 4, cudaMemcpyDeviceToHost, st));
  if (out && S) {
    float* const dst[7] = {out->radius, out->x, out->y, out->rotation, out->scaleX, out->scaleY, out->alpha};
    if (out->active) CK(cudaMemcpyAsync(out->active, ctx->shOutActive, S, cudaMemcpyDeviceToHost, st));
    for (int k = 0; k < 7; k++)
      if (dst[k]) CK(cudaMemcpyAsync(dst[k], ctx->shOut + k * S, S * 4, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  if (spriteCount) *spriteCount = n;
  return WEED_OK;
}

// =============================================================================================
// spawn / despawn pools
// =============================================================================================
static int pool_stage(weed_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->poolStageBytes) return WEED_OK;
  uint8_t* p = nullptr;
  int rc = dalloc(ctx, &p, bytes, false);
  if (rc) return rc;
  ctx->poolStage = p; ctx->poolStageBytes = bytes;
  return WEED_OK;
}
static int pool_get(weed_ctx* ctx, uint32_t pool, weed_ctx::Pool** out) {
  if (ctx->slab) return fail(ctx, WEED_E_STATE, "pools are not available on slab contexts yet");
  if (pool >= ctx->pools.size()) return fail(ctx, WEED_E_INVALID, "no such pool");
  *out = &ctx->pools[pool];
  return WEED_OK;
}

extern "C" int weed_pool_create(weed_ctx* ctx, uint32_t startIndex, uint32_t totalCount, uint32_t components, uint32_t* pool_out) {
  GUARD(ctx);
  if (ctx->slab) return fail(ctx, WEED_E_STATE, "pools are not available on slab contexts yet");
  if (!pool_out || totalCount == 0 || (size_t)startIndex + totalCount > ctx->g.N) return fail(ctx, WEED_E_INVALID, "pool range out of bounds");
  weed_ctx::Pool p{startIndex, totalCount, components, nullptr, nullptr};
  int rc = dalloc(ctx, &p.freeList, totalCount, false); if (rc) return rc;
  rc = dalloc(ctx, &p.st, 1); if (rc) return rc;
  if (!ctx->poolScalars) { rc = dalloc(ctx, &ctx->poolScalars, 4); if (rc) return rc; }
  k_pool_init<<<blocks_for(totalCount, 256), 256, 0, ctx->stream>>>(startIndex, totalCount, p.freeList, p.st);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->pools.push_back(p);
  *pool_out = (uint32_t)ctx->pools.size() - 1;
  return WEED_OK;
}

extern "C" int weed_pool_spawn(weed_ctx* ctx, uint32_t pool, const weed_spawn_record* records, uint32_t n, int32_t* indices_out) {
  GUARD(ctx);
  weed_ctx::Pool* p;
  int rc = pool_get(ctx, pool, &p); if (rc) return rc;
  if (n == 0) return WEED_OK;
  if (!records || !indices_out) return WEED_E_INVALID;
  rc = pool_stage(ctx, (size_t)n * (sizeof(SpawnRec) + 4)); if (rc) return rc;
  SpawnRec* dRecs = (SpawnRec*)ctx->poolStage;
  int32_t* dIdx = (int32_t*)((uint8_t*)ctx->poolStage + (size_t)n * sizeof(SpawnRec));
  cudaStream_t st = ctx->stream;
  CK(cudaMemcpyAsync(dRecs, records, (size_t)n * sizeof(SpawnRec), cudaMemcpyHostToDevice, st));
  k_pool_spawn<<<blocks_for(n, 256), 256, 0, st>>>(n, p->components, dRecs, p->freeList, p->st, ctx->d, dIdx);
  k_pool_spawn_end<<<1, 32, 0, st>>>(n, p->st);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(indices_out, dIdx, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  ctx->spatialValid = false;
  return WEED_OK;
}

static int pool_despawn(weed_ctx* ctx, weed_ctx::Pool* p, const int32_t* indices, uint32_t n, uint32_t* despawned) {
  cudaStream_t st = ctx->stream;
  int32_t* dIdx = nullptr;
  if (indices) {
    for (uint32_t k = 0; k < n; k++)
      if (indices[k] < (int32_t)p->start || indices[k] >= (int32_t)(p->start + p->count))
        return fail(ctx, WEED_E_INVALID, "index " + std::to_string(indices[k]) + " is not in the pool");
    int rc = pool_stage(ctx, (size_t)n * 4); if (rc) return rc;
    dIdx = (int32_t*)ctx->poolStage;
    CK(cudaMemcpyAsync(dIdx, indices, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    // `rank` is by-id scratch that only lives inside a frame
    k_pool_mark<<<blocks_for(n, 256), 256, 0, st>>>(n, dIdx, ctx->rank, true);
    k_pool_mark<<<blocks_for(n, 256), 256, 0, st>>>(n, dIdx, ctx->rank, false);
  }
  int rc = sys_tiles(ctx, n); if (rc) return rc;
  const unsigned blocks = blocks_for(n, SYS_TILE);
  k_pool_despawn_count<<<blocks, SYS_TILE, 0, st>>>(n, dIdx, p->start, ctx->rank, ctx->d.F, ctx->sysTileCount);
  k_tile_scan<<<1, 1024, 0, st>>>(ctx->sysTileCount, ctx->sysTilePrefix, blocks, ctx->poolScalars);
  k_pool_despawn_emit<<<blocks, SYS_TILE, 0, st>>>(n, p->count, p->components, dIdx, p->start, ctx->rank, ctx->d.F,
                                                   ctx->sysTilePrefix, p->freeList, p->st);
  k_pool_despawn_end<<<1, 32, 0, st>>>(ctx->poolScalars, p->count, p->st, ctx->poolScalars + 1);
  CK(cudaGetLastError());
  uint32_t cnt = 0; PoolState ps;
  CK(cudaMemcpyAsync(&cnt, ctx->poolScalars + 1, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&ps, p->st, sizeof(ps), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  ctx->spatialValid = false;
  if (despawned) *despawned = cnt;
  if (ps.overflow) return fail(ctx, WEED_E_OVERFLOW, "free list overflow: entities were despawned that the pool had never handed out");
  return WEED_OK;
}

extern "C" int weed_pool_despawn(weed_ctx* ctx, uint32_t pool, const int32_t* indices, uint32_t n, uint32_t* despawned_out) {
  GUARD(ctx);
  weed_ctx::Pool* p;
  int rc = pool_get(ctx, pool, &p); if (rc) return rc;
  if (despawned_out) *despawned_out = 0;
  if (n == 0) return WEED_OK;
  if (!indices) return WEED_E_INVALID;
  return pool_despawn(ctx, p, indices, n, despawned_out);
}

extern "C" int weed_pool_despawn_all(weed_ctx* ctx, uint32_t pool, uint32_t* despawned_out) {
  GUARD(ctx);
  weed_ctx::Pool* p;
  int rc = pool_get(ctx, pool, &p); if (rc) return rc;
  return pool_despawn(ctx, p, nullptr, p->count, despawned_out);
}

extern "C" int weed_pool_stats(weed_ctx* ctx, uint32_t pool, uint32_t* total, uint32_t* available) {
  GUARD(ctx);
  weed_ctx::Pool* p;
  int rc = pool_get(ctx, pool, &p); if (rc) return rc;
  PoolState ps;
  CK(cudaMemcpyAsync(&ps, p->st, sizeof(ps), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (total) *total = p->count;
  if (available) *available = (uint32_t)(ps.top + 1);
  return WEED_OK;
}
