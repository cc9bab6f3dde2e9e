/*
 * weedgpu.h — C ABI of libweedgpu.so, the B200-native replacement for the WeedJS
 * (brotochola/MultithreadedGameEngine) spatial_worker + physics_worker hot path.
 *
 * The reference has no FFI: its seam is the *worker contract* (an "init" message with
 * SharedArrayBuffers + config, then update() once per frame).  Every entry point below
 * names the reference interface it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - plain C, no torch / CUDA types in any signature; device pointers travel as void*.
 *   - every call returns 0 (WEED_OK) or a negative WEED_E_* code; nothing throws or aborts
 *     across the boundary.  weed_last_error() gives the text of the last failure.
 *   - a context is NOT thread-safe (the reference workers are single-threaded too); the
 *     caller serialises calls.  Host buffers are owned by the caller (the JS engine owns
 *     the SABs: src/core/gameEngine.js:534-777) and are never freed here.
 *   - there is no CPU fallback: without a CUDA device weed_create fails with WEED_E_CUDA.
 */
#ifndef WEEDGPU_H
#define WEEDGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WEED_ABI_VERSION 1

/* ---- status codes ---------------------------------------------------------------- */
enum {
  WEED_OK            =  0,
  WEED_E_INVALID     = -1,  /* bad argument / config                                   */
  WEED_E_CUDA        = -2,  /* CUDA runtime failure (context becomes sticky-failed)    */
  WEED_E_NOT_BOUND   = -3,  /* a required host buffer was not bound                    */
  WEED_E_SIZE        = -4,  /* bound buffer smaller than the layout requires           */
  WEED_E_OVERFLOW    = -5,  /* a capacity was exceeded (slab quota / table / halo reach, pool free list) */
  WEED_E_STATE       = -6,  /* call order violated (e.g. weed_physics before spatial)  */
  WEED_E_NOMEM       = -7
};

/* ---- buffers: the SharedArrayBuffers of the "init" payload ------------------------
 * src/core/gameEngine.js:1049-1125 (buffers.componentData.{Transform,RigidBody,Collider},
 * buffers.neighborData, buffers.distanceData, buffers.collisionData).               */
typedef enum weed_buffer_id {
  WEED_BUF_TRANSFORM = 0,   /* src/components/Transform.js:8-17   (14 B / entity)       */
  WEED_BUF_RIGIDBODY = 1,   /* src/components/RigidBody.js:9-47   (83 B / entity)       */
  WEED_BUF_COLLIDER  = 2,   /* src/components/Collider.js:8-46    (51 B / entity)       */
  WEED_BUF_NEIGHBOR  = 3,   /* Int32   [N*(1+maxNeighbors)]  gameEngine.js:552-555      */
  WEED_BUF_DISTANCE  = 4,   /* Float32 [N*(1+maxNeighbors)]  gameEngine.js:557-559      */
  WEED_BUF_COLLISION = 5,   /* Int32   [1+2*maxCollisionPairs] gameEngine.js:689-696    */
  WEED_BUF_COUNT     = 6
} weed_buffer_id;

/* ---- columns the hot path reads or writes --------------------------
 * Bits for upload_mask / download_mask.  Only these columns are mirrored on the device;
 * the remaining schema columns (mass, drag, friction, aabb*, ...) are declared by the
 * reference but never read by either worker, so they are never touched.               */
enum {
  WEED_COL_T_ACTIVE    = 1u << 0,   /* Transform.active        u8  */
  WEED_COL_T_X         = 1u << 1,   /* Transform.x             f32 */
  WEED_COL_T_Y         = 1u << 2,   /* Transform.y             f32 */
  WEED_COL_RB_ACTIVE   = 1u << 3,   /* RigidBody.active        u8  */
  WEED_COL_RB_STATIC   = 1u << 4,   /* RigidBody.static        u8  */
  WEED_COL_RB_VX       = 1u << 5,
  WEED_COL_RB_VY       = 1u << 6,
  WEED_COL_RB_AX       = 1u << 7,
  WEED_COL_RB_AY       = 1u << 8,
  WEED_COL_RB_PX       = 1u << 9,
  WEED_COL_RB_PY       = 1u << 10,
  WEED_COL_RB_MAXVEL   = 1u << 11,
  WEED_COL_RB_VELANGLE = 1u << 12,  /* RigidBody.velocityAngle f32 */
  WEED_COL_RB_SPEED    = 1u << 13,
  WEED_COL_RB_COLLCNT  = 1u << 14,  /* RigidBody.collisionCount u8 */
  WEED_COL_C_ACTIVE    = 1u << 15,
  WEED_COL_C_RADIUS    = 1u << 16,
  WEED_COL_C_ISTRIGGER = 1u << 17,
  WEED_COL_C_VISRANGE  = 1u << 18,  /* Collider.visualRange    f32 */
  WEED_COL_T_ENTITYTYPE = 1u << 19, /* Transform.entityType    u8 (read by device-side systems only) */
  /* pseudo-columns: whole output buffers (download only) */
  WEED_COL_NEIGHBORS   = 1u << 24,  /* all of neighborData + distanceData (large!)      */
  WEED_COL_COLLISIONS  = 1u << 25   /* collisionData                                    */
};
#define WEED_COLS_INPUT_ALL  0x000FFFFFu  /* every mirrored component column             */
/* what the reference's two workers write every frame (physics_worker.js:301-314,596-601,
 * 176,551) */
#define WEED_COLS_OUTPUT_ALL (WEED_COL_T_X | WEED_COL_T_Y | WEED_COL_RB_VX | WEED_COL_RB_VY | \
                              WEED_COL_RB_AX | WEED_COL_RB_AY | WEED_COL_RB_PX | WEED_COL_RB_PY | \
                              WEED_COL_RB_VELANGLE | WEED_COL_RB_SPEED | WEED_COL_RB_COLLCNT)

/* ---- configuration ----------------------------------------------------------------
 * config.physics of the engine; defaults as src/core/gameEngine.js:39-49 and
 * src/workers/physics_worker.js:33-40.  weed_set_physics applies the clamps of
 * validatePhysicsConfig (src/core/utils.js:269-301).                                   */
typedef struct weed_physics_config {
  int32_t subStepCount;               /* >= 1 after validation; default 4              */
  int32_t _pad0;
  double  boundaryElasticity;         /* clamp01; default 0.8                          */
  double  collisionResponseStrength;  /* clamp01; default 0.5                          */
  double  verletDamping;              /* clamp01; default 0.995                        */
  double  minSpeedForRotation;        /* default 0.1                                   */
  double  gravityX;                   /* default 0                                     */
  double  gravityY;                   /* default 0                                     */
} weed_physics_config;

/* Collision resolution order: always the documented deterministic J-order. */

enum {
  WEED_FLAG_NONE          = 0,
  WEED_FLAG_NO_GRAPH      = 1u << 0,  /* launch kernels directly, no CUDA graph         */
  WEED_FLAG_KERNEL_TIMING = 1u << 1,  /* per-kernel cudaEvent timing into weed_stats    */
  WEED_FLAG_NO_NEIGHBOR_ROWS = 1u << 2, /* do not allocate/write neighborData/distanceData
                                          (physics-only consumers); rows then unavailable */
  /* measurement aid (tools/ab_kernels.py and the cross-check tests): another form of the collision
   * sweep.  Results are bit-identical whichever form runs.                                    */
  WEED_FLAG_K6_TILE       = 1u << 10, /* k_sweep_tile (TMA-staged tiles; measured slower)        */
  WEED_FLAG_K4_WIDE       = 1u << 12, /* neighbor scan: one warp per entity whatever maxNeighbors */
  WEED_FLAG_K4_THREAD     = 1u << 13  /* neighbor scan: one thread per entity whatever maxNeighbors */
};

/* The "init" message: gameEngine.js:1049-1125 (entityCount, config.worldWidth/Height,
 * config.spatial.{cellSize,maxNeighbors}, config.physics.*, config.seed).             */
typedef struct weed_config {
  uint32_t struct_size;          /* sizeof(weed_config), ABI check                      */
  uint32_t entityCount;          /* totalEntityCount, includes Mouse at index 0         */
  double   worldWidth;
  double   worldHeight;
  double   cellSize;             /* config.spatial.cellSize  (spatial_worker.js:80)     */
  uint32_t maxNeighbors;         /* config.spatial.maxNeighbors (spatial_worker.js:85)  */
  uint32_t maxCollisionPairs;    /* physics_worker.js:74-77; the JS `|| 10000`
                                    fall-through for 0 is applied by the host wrapper   */
  double   seed;                 /* config.seed (AbstractWorker.js:287-292)             */
  weed_physics_config physics;
  int32_t  device;               /* CUDA ordinal                                        */
  uint32_t flags;                /* WEED_FLAG_*                                         */
  void*    stream;               /* cudaStream_t to run on, or NULL for a private one   */
  /* world-space slab owned by this context (multi-GPU, SURVEY §8 e): entities whose
   * clamped cell row lies in [slabRowBegin, slabRowEnd) are owned; 0,0 = whole world. */
  uint32_t slabRowBegin;
  uint32_t slabRowEnd;
  uint32_t slabHaloRows;         /* replicated rows beyond each cut: (S+1)*ceil(max visualRange/cellSize) */
  uint32_t _pad1;
} weed_config;

typedef struct weed_stats {
  uint64_t frames;               /* steps executed                                      */
  uint32_t gridCols, gridRows;   /* spatial_worker.js:82-83                             */
  uint32_t activeInGrid;         /* entities inserted in the grid last frame            */
  uint32_t maxCellOccupancy;
  uint64_t neighborsTotal;       /* sum of row counts last frame  (k-bar = /active)     */
  uint32_t cappedRows;           /* rows that hit maxNeighbors last frame               */
  uint32_t explicitPairs;        /* pairs routed through the explicit (asymmetric) path */
  uint32_t collisionPairs;       /* pairs found in the last substep (uncapped)          */
  uint32_t kernelLaunchesPerStep;
  float    ms[12];               /* WEED_FLAG_KERNEL_TIMING: ms of the last frame's spans: 0 k_cell_key,
                                    1 k_cell_scan, 2 k_scatter_ids+k_sort_big_cells+k_slot_rank, 3 k_build_slots+k_slot_prep,
                                    4 k_neighbors2, 5 the cap path (k_beyond_cap, k_back_alloc/_write/_sort, k_sort_lists),
                                    6 all k_sweep launches with their k_sweep_heavy branches,
                                    7 k_writeback+k_pair_scan+k_pair_emit;
                                    always: 8 = the last frame on the device clock
                                    (%globaltimer, first to last kernel); 9 = a COUNT: capped rows
                                    whose lost lower-id partners overflowed the internal row       */
} weed_stats;

typedef struct weed_ctx weed_ctx;

/* ---- layout: Component.initializeArrays / getBufferSize (src/core/Component.js:20-42,
 * 77-93).  Column index = position in the component's ARRAY_SCHEMA.                    */
size_t weed_buffer_bytes(weed_buffer_id id, uint32_t entityCount, uint32_t maxNeighbors,
                         uint32_t maxCollisionPairs);
/* byte offset of schema column `column` of component buffer `id`; (size_t)-1 if invalid */
size_t weed_column_offset(weed_buffer_id id, uint32_t column, uint32_t entityCount);
/* number of schema columns of a component buffer (5 / 23 / 16)                         */
uint32_t weed_column_count(weed_buffer_id id);
/* schema name of a column ("x", "velocityAngle", ...) or NULL                          */
const char* weed_column_name(weed_buffer_id id, uint32_t column);

/* fills *cfg with the engine defaults (gameEngine.js:34-49; maxNeighbors 100,
 * maxCollisionPairs 10000)                                                             */
void weed_default_config(weed_config* cfg);

/* replaces: new Worker(spatial_worker.js) + new Worker(physics_worker.js) and their
 * initialize(data) (spatial_worker.js:49-116, physics_worker.js:50-97)                 */
int  weed_create(const weed_config* cfg, weed_ctx** out);
void weed_destroy(weed_ctx* ctx);

/* replaces: the typed-array views the workers create over the SABs
 * (Component.js:36, AbstractWorker.js:176-285).  bytes must be >= weed_buffer_bytes(). */
int weed_bind(weed_ctx* ctx, weed_buffer_id id, void* host_base, size_t bytes);

/* host SAB columns -> device mirror / device mirror -> host SAB columns               */
int weed_upload(weed_ctx* ctx, uint32_t column_mask);
int weed_download(weed_ctx* ctx, uint32_t column_mask);

/* replaces: SpatialWorker.update (spatial_worker.js:283-294): rebuildGrid+findAllNeighbors */
int weed_spatial(weed_ctx* ctx);
/* replaces: PhysicsWorker.update (physics_worker.js:103-108): updateVerlet.  Uses the
 * neighbor rows of the preceding weed_spatial (lockstep order of SURVEY Appendix B).    */
int weed_physics(weed_ctx* ctx, double dtRatio);
/* one lockstep frame = upload(upload_mask); spatial; physics(dtRatio); download(mask). */
int weed_step(weed_ctx* ctx, double dtRatio, uint32_t upload_mask, uint32_t download_mask);
/* `frames` lockstep frames back to back with no host interaction (device-resident run) */
int weed_run(weed_ctx* ctx, double dtRatio, uint32_t frames);

/* replaces: {msg:"updatePhysicsConfig"} -> applyPhysicsConfig (physics_worker.js:114-129) */
int weed_set_physics(weed_ctx* ctx, const weed_physics_config* p);
int weed_get_physics(weed_ctx* ctx, weed_physics_config* out);

/* replaces: GameObject.updateNeighbors reading row i (src/core/gameObject.js:700-729):
 * copies rows [first, first+count) of neighborData and distanceData to the bound SABs.  */
int weed_fetch_neighbors(weed_ctx* ctx, uint32_t first, uint32_t count);
/* same rows into caller-provided compact buffers of count*(1+maxNeighbors) words each (for
 * worlds whose full neighborData would not fit in host memory); either pointer may be NULL */
int weed_fetch_neighbors_to(weed_ctx* ctx, uint32_t first, uint32_t count, int32_t* neighbor_out, float* distance_out);

int weed_sync(weed_ctx* ctx);
uint32_t weed_entity_count(weed_ctx* ctx);   /* entityCount the context was created with */
int weed_get_stats(weed_ctx* ctx, weed_stats* out);
const char* weed_last_error(weed_ctx* ctx);   /* ctx may be NULL: last create() failure */

/* device-resident access for in-process consumers (bench harness, multi-GPU plumbing,
 * device-side systems): raw device pointers.
 *
 * On the device the rows of neighborData / distanceData are stored as slot-major planes, not
 * in the reference's row layout: entry k of the entity in grid slot s is word
 * k * weed_row_pitch() + s of WEED_DEV_NEIGHBOR / WEED_DEV_DISTANCE, its count is
 * WEED_DEV_NEIGHBOR_COUNT[s], and the slot of entity i this frame is WEED_DEV_SLOT_OF[i]
 * (0xFFFFFFFF: inactive or NaN position, i.e. not in the grid; the reference leaves such a
 * row stale).  Consecutive slots are consecutive grid cells, so the scan kernel's stores are
 * whole cache lines; the row layout of gameEngine.js:552-559 exists in the host buffers,
 * which weed_fetch_neighbors / weed_download(WEED_COL_NEIGHBORS) fill through a gather.   */
typedef enum weed_devptr_id {
  WEED_DEV_NEIGHBOR  = 0,  /* int32  [maxNeighbors rounded up to 8][weed_row_pitch()]   */
  WEED_DEV_DISTANCE  = 1,  /* float  [same]                                              */
  WEED_DEV_COLLISION = 2,  /* int32  [1+2*maxPairs] */
  WEED_DEV_STATE     = 3,  /* float4 [N] {x, y, px, py}, see DESIGN.md                  */
  WEED_DEV_ATTR      = 4,  /* float4 [N] {maxVel, radius, visualRange, velocityAngle}   */
  WEED_DEV_VEL       = 5,  /* float4 [N] {vx, vy, speed, -}                             */
  WEED_DEV_NEIGHBOR_COUNT = 6, /* uint32 [N] by grid slot                               */
  WEED_DEV_SLOT_OF   = 7   /* uint32 [N] by entity index                                */
} weed_devptr_id;
int weed_device_ptr(weed_ctx* ctx, weed_devptr_id which, void** out, size_t* bytes);
/* Words between two planes of the DEVICE copies of neighborData / distanceData (the entity
 * count rounded up to 128).                                                               */
uint32_t weed_row_pitch(weed_ctx* ctx);

/* ---- device-side consumers of the neighbor rows ("systems") ----------------
 * The fixed-stride rows exist to feed GameObject.tick(); at >= 1M entities copying them to
 * the host dominates a frame.  A system is a tick() restated as a kernel that reads the rows
 * where they are and writes RigidBody.ax/ay, exactly like the logic worker would.
 * weed_system_boids restates the boids demo: demos/predators/boid.js:115-124 (tick),
 * :137-240 (applyFlockingBehaviors) and :318-341 (keepWithinBounds); avoidMouse and the
 * Prey/Predator processNeighbor hooks are not included.  Call it between frames, where the
 * logic workers run.                                                                      */
typedef struct weed_boids_params {
  double centeringFactor;   /* boid.js:65  0.001 */
  double avoidFactor;       /* boid.js:66  0.3   */
  double matchingFactor;    /* boid.js:67  0.1   */
  double turnFactor;        /* boid.js:68  0.01  */
  double margin;            /* boid.js:69  20    */
  uint32_t mouseEntityType; /* Mouse.entityType (0): such neighbors are skipped, boid.js:179 */
  uint32_t _pad;
} weed_boids_params;
/* protectedRange: per-entity float[entityCount] host array (Flocking.protectedRange), or
 * NULL for 2 * Collider.radius (boid.js:64).                                              */
int weed_system_boids(weed_ctx* ctx, const weed_boids_params* params, const float* protectedRange,
                      double dtRatio);

/* weed_system_flock: the whole predators demo tick() on the device — Boid (boid.js:115-124),
 * Prey (prey.js:120-137: flocking + applyFleeing :176-189 fed by processNeighbor :154-169) and
 * Predator (predator.js: flocking + applyHunting :195-215 fed by processNeighbor :172-187), then
 * avoidMouse (boid.js:281-316) and keepWithinBounds (:318-341), in that order, with the
 * per-class numbers each constructor sets.  An entity runs the class whose entityType matches
 * its Transform.entityType (WEED_FLOCK_ANY_TYPE matches all); entities of no listed class and
 * entity 0 (the Mouse) are left alone.                                                        */
#define WEED_FLOCK_BOID 0u
#define WEED_FLOCK_PREY 1u
#define WEED_FLOCK_PREDATOR 2u
#define WEED_FLOCK_ANY_TYPE 0xFFFFFFFFu
#define WEED_FLOCK_MAX_CLASSES 8
typedef struct weed_flock_class {
  uint32_t entityType;         /* Transform.entityType of the class                              */
  uint32_t role;               /* WEED_FLOCK_BOID / _PREY / _PREDATOR                            */
  uint32_t otherEntityType;    /* PREY: Predator.entityType (prey.js:164); PREDATOR: Prey's (predator.js:182) */
  uint32_t _pad;
  double protectedRangeScale;  /* protectedRange = fround(scale * Collider.radius) when no array is
                                  given: boid.js:64 2, prey.js:55 1.25, predator.js:57 0          */
  double centeringFactor, avoidFactor, matchingFactor, turnFactor, margin;   /* Flocking.*        */
  double roleFactor;           /* PREY: predatorAvoidFactor (prey.js:37); PREDATOR: huntFactor (predator.js:43) */
} weed_flock_class;
typedef struct weed_flock_params {
  uint32_t mouseEntityType;    /* Mouse.entityType: such neighbors are skipped, boid.js:179-180   */
  uint32_t mouseDown;          /* Mouse.x && Mouse.isDown (boid.js:282-283)                       */
  double dtRatio;
} weed_flock_params;
int weed_system_flock(weed_ctx* ctx, const weed_flock_class* classes, uint32_t classCount,
                      const weed_flock_params* params, const float* protectedRange);

/* ---- collision Enter / Stay / Exit ------------------------------------------
 * Replaces the bookkeeping of LogicWorker.processCollisionCallbacks (src/workers/
 * logic_worker.js:429-526): the pair list the last weed_physics / weed_step left in
 * collisionData is compared ON THE DEVICE with the list of the previous call.
 *   state[k], k < pairs : 1 = the pair k of collisionData is new (onCollisionEnter, :471-480),
 *                         2 = it was there last time (onCollisionStay, :481-489)
 *   exitData[0]         : number of ended pairs; then (a, b) couples in the ORDER of the
 *                         previous frame's list, which is the reference's Set iteration order
 *                         (:493-516).  Each ended pair is listed once; the reference fires
 *                         its Exit callbacks twice (once per Cantor key), see INTEGRATION.md.
 * Pairs are compared as exact 64-bit (a, b) keys; the reference's Cantor keys (:417-421) are
 * doubles and alias above ~2^26 ids.  state has room for maxCollisionPairs bytes, exitData
 * for 1 + 2 * maxCollisionPairs int32; either may be NULL.  Call once per frame.            */
#define WEED_EVENTS_FORGET_PREVIOUS 1u   /* start from an empty previous set (scene reload)   */
typedef struct weed_collision_event_counts {
  uint32_t pairs, entered, stayed, exited;
} weed_collision_event_counts;
int weed_system_collision_events(weed_ctx* ctx, uint32_t flags, weed_collision_event_counts* counts,
                                 uint8_t* state, int32_t* exitData);

/* ---- screen visibility and shadow sprites -----------------------------------
 * weed_system_screen_visibility restates ParticleWorker.updateEntityScreenVisibility
 * (src/workers/particle_worker.js:1012-1062): SpriteRenderer.screenX / screenY /
 * isItOnScreen of every Transform.active entity from the device-resident positions; entries of
 * inactive entities keep their previous values.  Host arrays (entityCount each) may be NULL:
 * the device copies stay for weed_system_shadows.                                           */
typedef struct weed_camera {
  double zoom, cameraX, cameraY;     /* cameraData[0..2], particle_worker.js:1029-1031 */
  double canvasWidth, canvasHeight;  /* :1036-1041 (15 % margin on each side)          */
} weed_camera;
int weed_system_screen_visibility(weed_ctx* ctx, const weed_camera* cam, float* screenX, float* screenY,
                                  uint8_t* isItOnScreen);
/* weed_system_shadows restates ParticleWorker.updateShadowSprites (particle_worker.js:861-1003):
 * lights in id order (LightEmitter.active, Transform.active, on screen, intensity > 0; at most
 * maxShadowCastingLights), each walking its neighbor row in order for ShadowCaster entities,
 * at most maxShadowsPerLight each and maxShadowSprites in all; unused sprite slots get
 * active = 0.  The five component columns are uploaded once with weed_system_shadows_upload
 * (again when they change); isItOnScreen is the one weed_system_screen_visibility left on the
 * device; neighbor rows are the ones of the last weed_spatial / weed_step.                  */
typedef struct weed_shadow_columns {
  const uint8_t* lightActive;     /* LightEmitter.active            */
  const float* lightIntensity;    /* LightEmitter.lightIntensity    */
  const uint8_t* casterActive;    /* ShadowCaster.active            */
  const float* shadowRadius;      /* ShadowCaster.shadowRadius      */
  const float* height;            /* ShadowCaster.height            */
} weed_shadow_columns;
typedef struct weed_shadow_sprites {  /* host arrays of maxShadowSprites elements, any may be NULL */
  uint8_t* active;
  float *radius, *x, *y, *rotation, *scaleX, *scaleY, *alpha;
} weed_shadow_sprites;
int weed_system_shadows_upload(weed_ctx* ctx, const weed_shadow_columns* cols);
int weed_system_shadows(weed_ctx* ctx, uint32_t maxShadowCastingLights, uint32_t maxShadowsPerLight,
                        uint32_t maxShadowSprites, const weed_shadow_sprites* out, uint32_t* spriteCount);

/* ---- spawn / despawn pools ----------------------------------------------------
 * GameObject.spawn / despawn / despawnAll (src/core/gameObject.js:840-951, 668-690, 1001-1034)
 * for whole batches, on the device-resident columns.  A pool is one entity class: the index
 * range [startIndex, startIndex + totalCount) with the reference's LIFO free list in its
 * interleaved initial order (initializeFreeList, :794-833 — like the reference it lists EVERY
 * index of the class as free, whatever Transform.active says).  A batch of n spawns takes the
 * indices n successive spawn() calls would take and leaves each entity as spawn() does:
 * ax = ay = 0, vx, vy, x, y from the record, speed = velocityAngle = 0, px = x - vx,
 * py = y - vy (:936-939), component active flags then Transform.active set.  indices_out[k] is
 * -1 once the pool is empty (:868-873).  A batch of despawns clears the active flags and pushes
 * the indices in batch order; inactive entities and repeats are skipped (:669-670).  Host-only
 * parts of spawn() (sprite state, onSpawned / onDespawned callbacks, other components) stay
 * with the caller, which gets the indices; the host copies of the touched columns are stale
 * until downloaded.                                                                           */
#define WEED_POOL_HAS_RIGIDBODY 1u
#define WEED_POOL_HAS_COLLIDER 2u
typedef struct weed_spawn_record { float x, y, vx, vy; } weed_spawn_record;   /* spawnConfig x, y, vx, vy */
int weed_pool_create(weed_ctx* ctx, uint32_t startIndex, uint32_t totalCount, uint32_t components, uint32_t* pool_out);
int weed_pool_spawn(weed_ctx* ctx, uint32_t pool, const weed_spawn_record* records, uint32_t n, int32_t* indices_out);
int weed_pool_despawn(weed_ctx* ctx, uint32_t pool, const int32_t* indices, uint32_t n, uint32_t* despawned_out);
int weed_pool_despawn_all(weed_ctx* ctx, uint32_t pool, uint32_t* despawned_out);
/* getPoolStats (:959-990): total and available (= free list entries)                          */
int weed_pool_stats(weed_ctx* ctx, uint32_t pool, uint32_t* total, uint32_t* available);

/* ---- multi-GPU slabs (SURVEY §8 e) ----------------------------------------
 * A slab context (slabRowEnd > 0) holds a LOCAL entity table of `entityCount` slots; every
 * slot carries a global entity id.  Component buffers, neighbor rows and column masks are
 * then indexed by local slot.  Ownership follows position: the context owns the entities
 * whose cell row lies in [slabRowBegin, slabRowEnd); slabHaloRows rows beyond each cut are
 * replicas that are recomputed redundantly during the frame and dropped at its end, so ONE
 * exchange per frame carries both the halo refresh and entity migration.  Exchange buffers
 * are plain device memory of 64-byte records; the host moves them between adjacent slabs
 * (torch.distributed / NCCL send-recv, or a device-to-device copy in one process).          */
#define WEED_SLAB_RECORD_BYTES 64
/* global ids of local slots [0,count); slots >= count are free.  Call after weed_upload.   */
int weed_slab_set_gids(weed_ctx* ctx, const uint32_t* gids, uint32_t count);
/* gids_out[capacity] (0xFFFFFFFF for free slots), *top_out = slots in use (synchronises)   */
int weed_slab_get_gids(weed_ctx* ctx, uint32_t* gids_out, uint32_t* top_out);
/* Exchange buffers hold (quota + 1) records: record 0 is a header carrying the record count, so
 * a FIXED-size message per neighbour and frame needs no separate size negotiation and the
 * whole exchange stays asynchronous on the context's stream (no host synchronisation).
 * weed_slab_pack: after a frame, records for the low / high neighbour.                     */
int weed_slab_pack(weed_ctx* ctx, void* dev_low, void* dev_high, uint32_t quota);
/* drop this frame's replicas, then insert the neighbours' records (either may be NULL)     */
int weed_slab_apply(weed_ctx* ctx, const void* dev_from_low, const void* dev_from_high, uint32_t quota);

/* Peer-to-peer exchange: no host and no library call inside a frame.  Every slab owns two receive
 * buffers per side (frame parity) of quota + 1 records; the neighbour's pack kernel writes its
 * records STRAIGHT into them over NVLink (a peer mapping), then the header and, after a system
 * fence, an arrival flag; this slab's apply waits for the flag on the device.  Setup, once:
 *   weed_slab_exchange_create(ctx, quota)             the same quota on every slab
 *   weed_slab_exchange_export(ctx, &handle, &base)    handle: for a neighbour in ANOTHER process
 *                                                     (a cudaIpcMemHandle_t, ship it any way you
 *                                                     like); base: for one in THIS process
 *   weed_slab_exchange_connect(ctx, side, handle|NULL, base|NULL)   side 0: the low neighbour
 *                                                     (rows below mine), 1: the high one
 * then, per frame, weed_slab_frame (frame kernels + pack + wait + apply, all asynchronous on the
 * context's stream) or weed_step followed by weed_slab_exchange.  weed_slab_frame_begin / _end split
 * the frame at the wait: one host thread that drives several slabs calls begin on all of them, then
 * end on all of them, so that no slab waits for a message that has not been queued yet.
 * The halo is sized for the entities near the cuts when the slabs were planned; if an entity whose
 * reach exceeds it (an observer with a large visualRange) later comes near a cut, weed_slab_status
 * returns WEED_E_OVERFLOW instead of letting the partition diverge from the single-context result. */
typedef struct weed_ipc_handle { unsigned char bytes[64]; } weed_ipc_handle;
int weed_slab_exchange_create(weed_ctx* ctx, uint32_t quota);
int weed_slab_exchange_export(weed_ctx* ctx, weed_ipc_handle* handle_out, void** base_out);
int weed_slab_exchange_connect(weed_ctx* ctx, int side, const weed_ipc_handle* peer_handle, void* peer_base);
int weed_slab_frame(weed_ctx* ctx, double dtRatio);
int weed_slab_frame_begin(weed_ctx* ctx, double dtRatio);
int weed_slab_frame_end(weed_ctx* ctx);
int weed_slab_exchange(weed_ctx* ctx);
/* unmaps the neighbours' buffers (call it on every slab, then synchronise the processes, before any
 * of them destroys its context: memory a peer still has mapped must not be freed)              */
int weed_slab_exchange_disconnect(weed_ctx* ctx);

/* One world on several GPUs of ONE process (what a single Node engine uses; call site
 * src/core/gameEngine.js:972-1009 creates one spatial and one physics worker — here: one group).
 * weed_group_create makes one slab context per entry of devices[] (capacity = its local entity
 * table), creates their exchange buffers and connects neighbours (peer access over NVLink).  The
 * host fills every slab's local table through weed_group_slab(g, k): weed_bind / weed_upload /
 * weed_slab_set_gids as for any slab context.  weed_group_step queues one frame on every slab and
 * returns at once; weed_group_sync waits and reports the first slab error (WEED_E_OVERFLOW ...).  */
typedef struct weed_group weed_group;
int weed_group_create(const weed_config* cfg_template, uint32_t nSlabs, const int32_t* devices, const uint32_t* rowCuts,
                      uint32_t haloRows, const uint32_t* capacities, uint32_t quota, weed_group** out);
uint32_t weed_group_size(weed_group* g);
weed_ctx* weed_group_slab(weed_group* g, uint32_t k);
int weed_group_step(weed_group* g, double dtRatio);
int weed_group_sync(weed_group* g);
void weed_group_destroy(weed_group* g);
const char* weed_group_last_error(weed_group* g);   /* g may be NULL: last create() failure */
typedef struct weed_slab_stats {
  uint32_t top, capacity;          /* local slots in use / available                         */
  uint32_t owned;                  /* entities owned during the last packed frame            */
  uint32_t sentLow, sentHigh, receivedLow, receivedHigh;   /* records of the last exchange   */
  uint32_t overflow;               /* sticky: 1 = quota exceeded, 2 = table full             */
  int32_t rowBegin, rowEnd;        /* the rows this context owns in the NEXT frame           */
  uint32_t cutMoves;               /* times one of its cuts moved (weed_slab_balance)        */
  uint32_t loadNs;                 /* smoothed device time of its frame kernels              */
} weed_slab_stats;
/* Dynamic balancing (SURVEY §8 e: cuts follow the load).  Every message header carries the
 * sender's smoothed frame time (device %globaltimer around the frame's kernels) and its cuts,
 * so the two neighbours of a cut decide — on identical numbers, hence identically, without
 * any extra communication or host involvement — to move it by up to maxShiftRows rows toward
 * the slower slab when the two times differ by more than hysteresisPercent (0 = 3 %).  The
 * new cut is known one frame ahead: the exchange that precedes its first frame packs the
 * bands around it, so the halo rule holds unchanged and results stay bit-identical to the
 * unpartitioned world whatever the cuts do.  maxShiftRows = 0 (default) keeps the cuts static. */
int weed_slab_balance(weed_ctx* ctx, uint32_t maxShiftRows, uint32_t hysteresisPercent);
/* synchronises; returns WEED_E_OVERFLOW if a quota or the table overflowed at any time      */
int weed_slab_status(weed_ctx* ctx, weed_slab_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* WEEDGPU_H */
