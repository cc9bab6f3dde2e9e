/*
 * weed_oracle.c — CPU restatement of the WeedJS spatial_worker + physics_worker hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (multithreadedgameengine_b200/,
 * csrc/) may import, link or call this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the reference (brotochola/MultithreadedGameEngine) ships no tests,
 * golden vectors or fixtures for this path, and no JavaScript engine exists in this image
 * to run the original.  Fidelity rests on (1) this line-by-line restatement, (2) an
 * independent numpy restatement (oracle/oracle_np.py) that must agree bit-for-bit, and
 * (3) hand-derived micro-cases under tests/golden/.
 *
 * Numeric model: every typed-array load widens to binary64, every
 * arithmetic op is an individually rounded binary64 op (build with -ffp-contract=off),
 * stores round to the column type.  `x | 0` is ECMAScript ToInt32.
 *
 * All citations are relative to the reference tree (src/workers/...).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/weed_nudge.h"

#define WO_EXPORT __attribute__((visibility("default")))

/* ---- Component layout: src/core/Component.js:20-42 --------------------------------- */
typedef struct { const char* name; int bytes; } wo_col;

static const wo_col TRANSFORM_SCHEMA[] = { /* src/components/Transform.js:8-17 */
  {"active", 1}, {"entityType", 1}, {"x", 4}, {"y", 4}, {"rotation", 4}};
static const wo_col RIGIDBODY_SCHEMA[] = { /* src/components/RigidBody.js:9-47 */
  {"active", 1}, {"static", 1}, {"vx", 4}, {"vy", 4}, {"ax", 4}, {"ay", 4}, {"px", 4},
  {"py", 4}, {"angularVelocity", 4}, {"angularAccel", 4}, {"mass", 4}, {"invMass", 4},
  {"inertia", 4}, {"invInertia", 4}, {"drag", 4}, {"angularDrag", 4}, {"maxVel", 4},
  {"maxAcc", 4}, {"minSpeed", 4}, {"friction", 4}, {"velocityAngle", 4}, {"speed", 4},
  {"collisionCount", 1}};
static const wo_col COLLIDER_SCHEMA[] = { /* src/components/Collider.js:8-46 */
  {"active", 1}, {"shapeType", 1}, {"offsetX", 4}, {"offsetY", 4}, {"radius", 4},
  {"width", 4}, {"height", 4}, {"isTrigger", 1}, {"restitution", 4}, {"collisionLayer", 2},
  {"collisionMask", 2}, {"aabbMinX", 4}, {"aabbMinY", 4}, {"aabbMaxX", 4}, {"aabbMaxY", 4},
  {"visualRange", 4}};

static const wo_col* schema_of(int comp, int* n) {
  switch (comp) {
    case 0: *n = 5;  return TRANSFORM_SCHEMA;
    case 1: *n = 23; return RIGIDBODY_SCHEMA;
    case 2: *n = 16; return COLLIDER_SCHEMA;
  }
  *n = 0; return NULL;
}

/* Component.initializeArrays offset rule (Component.js:24-38) */
WO_EXPORT int64_t wo_column_offset(int comp, const char* name, int64_t count) {
  int n; const wo_col* s = schema_of(comp, &n);
  int64_t off = 0;
  for (int k = 0; k < n; k++) {
    int64_t rem = off % s[k].bytes;
    if (rem != 0) off += s[k].bytes - rem;
    if (strcmp(s[k].name, name) == 0) return off;
    off += count * s[k].bytes;
  }
  return -1;
}
/* Component.getBufferSize (Component.js:77-93) */
WO_EXPORT int64_t wo_buffer_size(int comp, int64_t count) {
  int n; const wo_col* s = schema_of(comp, &n);
  int64_t off = 0;
  for (int k = 0; k < n; k++) {
    int64_t rem = off % s[k].bytes;
    if (rem != 0) off += s[k].bytes - rem;
    off += count * s[k].bytes;
  }
  return off;
}

/* ---- JS helpers --------------------------------------------------------------------- */
static int32_t js_toint32(double v) { /* ECMAScript ToInt32 */
  if (!isfinite(v)) return 0;
  v = trunc(v);
  if (v >= -2147483648.0 && v <= 2147483647.0) return (int32_t)v;
  double m = fmod(v, 4294967296.0);
  if (m < 0) m += 4294967296.0;
  return (int32_t)(uint32_t)m;
}
static uint32_t js_touint32(double v) { return (uint32_t)js_toint32(v); }
static double js_min(double a, double b) {
  if (isnan(a) || isnan(b)) return NAN;
  return a < b ? a : b;
}
static double js_max(double a, double b) {
  if (isnan(a) || isnan(b)) return NAN;
  return a > b ? a : b;
}
static double clamp01(double v) { return js_max(0, js_min(1, v)); } /* utils.js:16-19 */

/* seededRandom: src/core/utils.js:333-342.  `t` is a JS Number, not an int32. */
typedef struct { double t; } wo_rng;
static double wo_rng_next(wo_rng* g) {
  g->t += (double)0x6d2b79f5;
  int32_t t32 = js_toint32(g->t);
  uint32_t tu = (uint32_t)t32;
  /* r = Math.imul(t ^ (t >>> 15), 1 | t) */
  int32_t a = (int32_t)(tu ^ (tu >> 15));
  int32_t b = (int32_t)(1u | tu);
  int32_t r = (int32_t)((uint32_t)a * (uint32_t)b);
  /* r = (r + Math.imul(r ^ (r >>> 7), 61 | r)) ^ r;  `+` is a double add, `^` ToInt32s it */
  uint32_t ru = (uint32_t)r;
  int32_t m = (int32_t)((uint32_t)(int32_t)(ru ^ (ru >> 7)) * (uint32_t)(int32_t)(61u | ru));
  double sum = (double)r + (double)m;
  int32_t r2 = js_toint32(sum) ^ r;
  uint32_t r2u = (uint32_t)r2;
  /* ((r ^ (r >>> 14)) >>> 0) / 4294967296 */
  uint32_t out = r2u ^ (r2u >> 14);
  return (double)out / 4294967296.0;
}

/* ---- context ------------------------------------------------------------------------ */
typedef struct { int32_t* v; int32_t len, cap; } wo_list; /* a JS Array used as a cell */

typedef struct wo_ctx {
  int32_t N;
  double worldWidth, worldHeight, cellSize, invCellSize;
  int32_t gridCols, gridRows, totalCells, maxNeighbors, maxPairs;
  /* physics settings (physics_worker.js:33-40) */
  int32_t subStepCount;
  double boundaryElasticity, collisionResponseStrength, verletDamping, minSpeedForRotation;
  double gravityX, gravityY;
  double seed;
  wo_rng rng;
  uint32_t frame; /* counts wo_physics calls; only the J-order nudge hash uses it */
  /* SAB bases */
  uint8_t *tbuf, *rbuf, *cbuf;
  int32_t* neighborData; float* distanceData; int32_t* collisionData;
  /* views */
  uint8_t *t_active; float *x, *y;
  uint8_t *rb_active, *rb_static, *collisionCount;
  float *vx, *vy, *ax, *ay, *px, *py, *maxVel, *velocityAngle, *speed;
  uint8_t *c_active, *isTrigger; float *radius, *visualRange;
  /* grid (spatial_worker.js:93-100); occupiedCells is 32-bit here */
  wo_list* grid; int32_t* occupiedCells; int32_t occupiedCount;
  int32_t* cellOf; /* clamped cell of each entity at the last rebuild, -1 if not inserted */
} wo_ctx;

WO_EXPORT wo_ctx* wo_create(int32_t N, double worldWidth, double worldHeight, double cellSize,
                            int32_t maxNeighbors, int32_t maxPairs, double seed) {
  wo_ctx* c = (wo_ctx*)calloc(1, sizeof(wo_ctx));
  c->N = N; c->worldWidth = worldWidth; c->worldHeight = worldHeight;
  /* spatial_worker.js:80-86 */
  c->cellSize = cellSize;
  c->invCellSize = 1 / cellSize;
  c->gridCols = (int32_t)ceil(worldWidth / cellSize);
  c->gridRows = (int32_t)ceil(worldHeight / cellSize);
  c->totalCells = c->gridCols * c->gridRows;
  c->maxNeighbors = maxNeighbors;
  c->maxPairs = maxPairs;
  c->grid = (wo_list*)calloc((size_t)c->totalCells, sizeof(wo_list));
  c->occupiedCells = (int32_t*)malloc(sizeof(int32_t) * (size_t)(N < c->totalCells ? N : c->totalCells) + 4);
  c->cellOf = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
  for (int i = 0; i < N; i++) c->cellOf[i] = -1;
  /* physics_worker.js:33-40 */
  c->subStepCount = 4; c->boundaryElasticity = 0.8; c->collisionResponseStrength = 0.5;
  c->verletDamping = 0.995; c->minSpeedForRotation = 0.1; c->gravityX = 0; c->gravityY = 0;
  c->seed = seed; c->rng.t = seed; /* AbstractWorker.js:287-292 */
  return c;
}

WO_EXPORT void wo_destroy(wo_ctx* c) {
  if (!c) return;
  for (int i = 0; i < c->totalCells; i++) free(c->grid[i].v);
  free(c->grid); free(c->occupiedCells); free(c->cellOf); free(c);
}

/* validatePhysicsConfig: src/core/utils.js:269-301 (all fields supplied) */
WO_EXPORT void wo_set_physics(wo_ctx* c, int32_t subStepCount, double boundaryElasticity,
                              double collisionResponseStrength, double verletDamping,
                              double minSpeedForRotation, double gx, double gy) {
  c->subStepCount = subStepCount < 1 ? 1 : subStepCount;
  c->boundaryElasticity = clamp01(boundaryElasticity);
  c->collisionResponseStrength = clamp01(collisionResponseStrength);
  c->verletDamping = clamp01(verletDamping);
  c->minSpeedForRotation = minSpeedForRotation;
  c->gravityX = gx; c->gravityY = gy;
}

#define VIEW(type, base, comp, name) ((type*)((base) + wo_column_offset((comp), (name), c->N)))

WO_EXPORT void wo_bind(wo_ctx* c, int id, void* p) {
  switch (id) {
    case 0: c->tbuf = (uint8_t*)p;
      c->t_active = VIEW(uint8_t, c->tbuf, 0, "active");
      c->x = VIEW(float, c->tbuf, 0, "x"); c->y = VIEW(float, c->tbuf, 0, "y"); break;
    case 1: c->rbuf = (uint8_t*)p;
      c->rb_active = VIEW(uint8_t, c->rbuf, 1, "active");
      c->rb_static = VIEW(uint8_t, c->rbuf, 1, "static");
      c->vx = VIEW(float, c->rbuf, 1, "vx"); c->vy = VIEW(float, c->rbuf, 1, "vy");
      c->ax = VIEW(float, c->rbuf, 1, "ax"); c->ay = VIEW(float, c->rbuf, 1, "ay");
      c->px = VIEW(float, c->rbuf, 1, "px"); c->py = VIEW(float, c->rbuf, 1, "py");
      c->maxVel = VIEW(float, c->rbuf, 1, "maxVel");
      c->velocityAngle = VIEW(float, c->rbuf, 1, "velocityAngle");
      c->speed = VIEW(float, c->rbuf, 1, "speed");
      c->collisionCount = VIEW(uint8_t, c->rbuf, 1, "collisionCount"); break;
    case 2: c->cbuf = (uint8_t*)p;
      c->c_active = VIEW(uint8_t, c->cbuf, 2, "active");
      c->isTrigger = VIEW(uint8_t, c->cbuf, 2, "isTrigger");
      c->radius = VIEW(float, c->cbuf, 2, "radius");
      c->visualRange = VIEW(float, c->cbuf, 2, "visualRange"); break;
    case 3: c->neighborData = (int32_t*)p; break;
    case 4: c->distanceData = (float*)p; break;
    case 5: c->collisionData = (int32_t*)p; break;
  }
}

static void list_push(wo_list* l, int32_t v) {
  if (l->len == l->cap) {
    l->cap = l->cap ? l->cap * 2 : 4;
    l->v = (int32_t*)realloc(l->v, sizeof(int32_t) * (size_t)l->cap);
  }
  l->v[l->len++] = v;
}

/* ---- SpatialWorker.rebuildGrid: spatial_worker.js:122-172 --------------------------- */
static void rebuild_grid(wo_ctx* c) {
  wo_list* grid = c->grid;
  for (int i = 0; i < c->occupiedCount; i++) grid[c->occupiedCells[i]].len = 0;
  c->occupiedCount = 0;
  const uint8_t* active = c->t_active; const float* x = c->x; const float* y = c->y;
  const double inv = c->invCellSize;
  const int32_t gridCols = c->gridCols, maxCol = c->gridCols - 1, maxRow = c->gridRows - 1;
  int32_t occupiedIdx = 0;
  for (int32_t i = 0; i < c->N; i++) {
    c->cellOf[i] = -1;
    if (!active[i]) continue;
    const double posX = x[i], posY = y[i];
    if (posX != posX || posY != posY) continue;
    int32_t col = js_toint32(posX * inv);
    int32_t row = js_toint32(posY * inv);
    col = col < 0 ? 0 : col > maxCol ? maxCol : col;
    row = row < 0 ? 0 : row > maxRow ? maxRow : row;
    const int32_t cellIndex = row * gridCols + col;
    wo_list* cell = &grid[cellIndex];
    if (cell->len == 0) c->occupiedCells[occupiedIdx++] = cellIndex;
    list_push(cell, i);
    c->cellOf[i] = cellIndex;
  }
  c->occupiedCount = occupiedIdx;
}

/* ---- SpatialWorker.findAllNeighbors: spatial_worker.js:178-278 ---------------------- */
static void find_all_neighbors(wo_ctx* c) {
  const float* x = c->x; const float* y = c->y; const float* visualRange = c->visualRange;
  wo_list* grid = c->grid;
  const double inv = c->invCellSize;
  const int32_t gridCols = c->gridCols, gridRows = c->gridRows;
  const int32_t maxNeighbors = c->maxNeighbors;
  const int64_t stride = 1 + (int64_t)maxNeighbors;
  for (int32_t cellIdx = 0; cellIdx < c->occupiedCount; cellIdx++) {
    const wo_list* centerCell = &grid[c->occupiedCells[cellIdx]];
    for (int32_t e = 0; e < centerCell->len; e++) {
      const int32_t i = centerCell->v[e];
      const double myX = x[i], myY = y[i];
      const double myVisualRange = visualRange[i];
      const double visualRangeSq = myVisualRange * myVisualRange;
      const double cellRadius = ceil(myVisualRange * inv);
      const int32_t col = js_toint32(myX * inv);
      const int32_t row = js_toint32(myY * inv);
      const int64_t offset = (int64_t)i * stride;
      int32_t neighborCount = 0;
      const double rowMin = row - cellRadius, rowMax = row + cellRadius;
      const double colMin = col - cellRadius, colMax = col + cellRadius;
      const double startRow = rowMin < 0 ? 0 : rowMin;
      const double endRow = rowMax >= gridRows ? gridRows - 1 : rowMax;
      const double startCol = colMin < 0 ? 0 : colMin;
      const double endCol = colMax >= gridCols ? gridCols - 1 : colMax;
      /* JS loop variables are Numbers; NaN bounds give zero iterations. */
      for (double checkRow = startRow; checkRow <= endRow; checkRow++) {
        const int64_t rowBase = (int64_t)checkRow * gridCols;
        for (double checkCol = startCol; checkCol <= endCol; checkCol++) {
          const wo_list* cell = &grid[rowBase + (int64_t)checkCol];
          const int32_t cellLength = cell->len;
          if (cellLength == 0) continue;
          for (int32_t k = 0; k < cellLength; k++) {
            const int32_t j = cell->v[k];
            if (i == j) continue;
            const double deltaX = x[j] - myX;
            const double deltaY = y[j] - myY;
            const double distSq = deltaX * deltaX + deltaY * deltaY;
            if (distSq < visualRangeSq && distSq > 0) {
              const int64_t writeIdx = offset + 1 + neighborCount;
              c->neighborData[writeIdx] = j;
              c->distanceData[writeIdx] = (float)distSq;
              neighborCount++;
              if (neighborCount >= maxNeighbors) break;
            }
          }
          if (neighborCount >= maxNeighbors) break;
        }
        if (neighborCount >= maxNeighbors) break;
      }
      c->neighborData[offset] = neighborCount;
      c->distanceData[offset] = (float)neighborCount;
    }
  }
}

/* SpatialWorker.update: spatial_worker.js:283-294 */
WO_EXPORT void wo_spatial(wo_ctx* c) {
  rebuild_grid(c);
  find_all_neighbors(c);
}

/* Grid introspection for parity tests: CSR of the per-cell lists in cell order. */
WO_EXPORT void wo_grid_export(wo_ctx* c, int32_t* cellOf, int32_t* cellStart /*[C+1]*/,
                              int32_t* sortedIdx /*[N]*/) {
  memcpy(cellOf, c->cellOf, sizeof(int32_t) * (size_t)c->N);
  /* cells not in occupiedCells may hold stale lengths only if never cleared; the rebuild
     clears every previously occupied cell, so len is exact for all cells. */
  int32_t pos = 0;
  for (int32_t k = 0; k < c->totalCells; k++) {
    cellStart[k] = pos;
    for (int32_t e = 0; e < c->grid[k].len; e++) sortedIdx[pos++] = c->grid[k].v[e];
  }
  cellStart[c->totalCells] = pos;
}
WO_EXPORT void wo_grid_dims(wo_ctx* c, int32_t* cols, int32_t* rows) {
  *cols = c->gridCols; *rows = c->gridRows;
}

/* ---- moveBallsVerlet: physics_worker.js:240-316 ------------------------------------- */
static void move_balls_verlet(wo_ctx* c, double dtRatio, double gx, double gy) {
  const double damping = c->verletDamping;
  const double gravityScale = dtRatio * dtRatio; /* Math.pow(dtRatio, 2) */
  float *x = c->x, *y = c->y, *px = c->px, *py = c->py;
  for (int32_t i = 0; i < c->N; i++) {
    if (!c->t_active[i] || !c->rb_active[i]) continue;
    if (c->rb_static[i]) continue;
    const double oldX = x[i], oldY = y[i];
    double dx = ((double)x[i] - (double)px[i]) * damping;
    double dy = ((double)y[i] - (double)py[i]) * damping;
    dx += gravityScale * gx + (double)c->ax[i] * dtRatio;
    dy += gravityScale * gy + (double)c->ay[i] * dtRatio;
    const double maxSpeed = c->maxVel[i] > 0 ? (double)c->maxVel[i] : 100;
    dx = js_max(-maxSpeed, js_min(maxSpeed, dx));
    dy = js_max(-maxSpeed, js_min(maxSpeed, dy));
    x[i] = (float)(oldX + dx);
    y[i] = (float)(oldY + dy);
    px[i] = (float)oldX;
    py[i] = (float)oldY;
    c->vx[i] = (float)(dx / dtRatio);
    c->vy[i] = (float)(dy / dtRatio);
    c->ax[i] = 0;
    c->ay[i] = 0;
  }
}

/* ---- boundary part of applyConstraintsVerlet: physics_worker.js:344-376 ------------- */
static void apply_bounds(wo_ctx* c) {
  const double e = c->boundaryElasticity;
  const double W = c->worldWidth, H = c->worldHeight;
  float *x = c->x, *y = c->y, *px = c->px, *py = c->py;
  for (int32_t i = 0; i < c->N; i++) {
    if (!c->t_active[i] || !c->rb_active[i]) continue;
    if (c->rb_static[i]) continue;
    const double r = c->radius[i];
    if (x[i] < r) {
      x[i] = (float)r;
      px[i] = (float)((double)x[i] + ((double)x[i] - (double)px[i]) * e);
    }
    if (x[i] > W - r) {
      x[i] = (float)(W - r);
      px[i] = (float)((double)x[i] + ((double)x[i] - (double)px[i]) * e);
    }
    if (y[i] < r) {
      y[i] = (float)r;
      py[i] = (float)((double)y[i] + ((double)y[i] - (double)py[i]) * e);
    }
    if (y[i] > H - r) {
      y[i] = (float)(H - r);
      py[i] = (float)((double)y[i] + ((double)y[i] - (double)py[i]) * e);
    }
  }
}

/* ---- resolveCollisionsVerlet, reference order: physics_worker.js:405-568 ------------
 * In-place Gauss-Seidel sweep, i ascending, row order. */
static void resolve_collisions_reference(wo_ctx* c) {
  const int64_t stride = 1 + (int64_t)c->maxNeighbors;
  const double responseStrength = c->collisionResponseStrength;
  int32_t pairCount = 0;
  int32_t* collisionData = c->collisionData;
  const int32_t maxPairs = c->maxPairs;
  float *x = c->x, *y = c->y;
  const float* radius = c->radius;
  for (int32_t i = 0; i < c->N; i++) {
    if (!c->t_active[i] || !c->c_active[i]) continue;
    const int64_t offset = (int64_t)i * stride;
    const int32_t neighborCount = c->neighborData[offset];
    for (int32_t n = 0; n < neighborCount; n++) {
      const int32_t j = c->neighborData[offset + 1 + n];
      if (i == j || !c->t_active[j] || !c->c_active[j]) continue;
      if (i >= j) continue;
      const double dx = (double)x[i] - (double)x[j];
      const double dy = (double)y[i] - (double)y[j];
      const double dist2 = dx * dx + dy * dy;
      const double minDist = (double)radius[i] + (double)radius[j];
      if (dist2 >= minDist * minDist) continue;
      const double dist = sqrt(dist2);
      if (dist == 0) {
        const int eitherIsTrigger = c->isTrigger[i] || c->isTrigger[j];
        if (!eitherIsTrigger) {
          const int iStatic = c->rb_static[i], jStatic = c->rb_static[j];
          const double angle = wo_rng_next(&c->rng) * M_PI * 2;
          const double separation = 0.001;
          const double cosAngle = cos(angle) * separation;
          const double sinAngle = sin(angle) * separation;
          if (iStatic && jStatic) {
          } else if (iStatic) {
            x[j] = (float)((double)x[j] - cosAngle * 2);
            y[j] = (float)((double)y[j] - sinAngle * 2);
          } else if (jStatic) {
            x[i] = (float)((double)x[i] + cosAngle * 2);
            y[i] = (float)((double)y[i] + sinAngle * 2);
          } else {
            x[i] = (float)((double)x[i] + cosAngle);
            y[i] = (float)((double)y[i] + sinAngle);
            x[j] = (float)((double)x[j] - cosAngle);
            y[j] = (float)((double)y[j] - sinAngle);
          }
        }
        c->collisionCount[i]++;
        c->collisionCount[j]++;
        if (collisionData && pairCount < maxPairs) {
          collisionData[1 + pairCount * 2] = i;
          collisionData[1 + pairCount * 2 + 1] = j;
          pairCount++;
        }
        continue;
      }
      const double depth = minDist - dist;
      if (depth > 0) {
        const int eitherIsTrigger = c->isTrigger[i] || c->isTrigger[j];
        if (!eitherIsTrigger) {
          const double nx = dx / dist;
          const double ny = dy / dist;
          const int iStatic = c->rb_static[i], jStatic = c->rb_static[j];
          const double correction = depth * responseStrength;
          if (iStatic && jStatic) {
          } else if (iStatic) {
            x[j] = (float)((double)x[j] - nx * correction);
            y[j] = (float)((double)y[j] - ny * correction);
          } else if (jStatic) {
            x[i] = (float)((double)x[i] + nx * correction);
            y[i] = (float)((double)y[i] + ny * correction);
          } else {
            const double halfCorrection = correction * 0.5;
            x[i] = (float)((double)x[i] + nx * halfCorrection);
            y[i] = (float)((double)y[i] + ny * halfCorrection);
            x[j] = (float)((double)x[j] - nx * halfCorrection);
            y[j] = (float)((double)y[j] - ny * halfCorrection);
          }
        }
        c->collisionCount[i]++;
        c->collisionCount[j]++;
        if (collisionData && pairCount < maxPairs) {
          collisionData[1 + pairCount * 2] = i;
          collisionData[1 + pairCount * 2 + 1] = j;
          pairCount++;
        }
      }
    }
  }
  if (collisionData) collisionData[0] = pairCount;
}

/* ---- resolveCollisionsVerlet, J-order (the documented deterministic GPU order) --------
 * Same pair set P = {(i,j): i<j, j in row(i), both active colliders} and same per-pair
 * formulas as physics_worker.js:428-562, but every pair is evaluated on the positions at
 * the START of the sweep (after the boundary pass), and each entity applies the
 * corrections it receives one by one (each rounded to float32 like the reference's `+=`
 * on a Float32Array), ordered by the partner's (cell index at the last grid rebuild,
 * entity id).  See DESIGN.md "J-order".  collisionData keeps the reference order
 * (i ascending, row position).  The dist==0 nudge uses include/weed_nudge.h. */
typedef struct { int32_t cell, id; double dx, dy; } wo_corr;
static int corr_cmp(const void* a, const void* b) {
  const wo_corr* p = (const wo_corr*)a; const wo_corr* q = (const wo_corr*)b;
  if (p->cell != q->cell) return p->cell < q->cell ? -1 : 1;
  return p->id < q->id ? -1 : (p->id > q->id ? 1 : 0);
}
typedef struct { wo_corr* v; int32_t len, cap; } wo_corrlist;
static void corr_push(wo_corrlist* l, int32_t cell, int32_t id, double dx, double dy) {
  if (l->len == l->cap) {
    l->cap = l->cap ? l->cap * 2 : 4;
    l->v = (wo_corr*)realloc(l->v, sizeof(wo_corr) * (size_t)l->cap);
  }
  wo_corr* e = &l->v[l->len++];
  e->cell = cell; e->id = id; e->dx = dx; e->dy = dy;
}

static void resolve_collisions_jorder(wo_ctx* c, uint32_t substep) {
  const int64_t stride = 1 + (int64_t)c->maxNeighbors;
  const double responseStrength = c->collisionResponseStrength;
  int32_t pairCount = 0;
  int32_t* collisionData = c->collisionData;
  const int32_t maxPairs = c->maxPairs;
  const int32_t N = c->N;
  float* x0 = (float*)malloc(sizeof(float) * (size_t)N);
  float* y0 = (float*)malloc(sizeof(float) * (size_t)N);
  memcpy(x0, c->x, sizeof(float) * (size_t)N);
  memcpy(y0, c->y, sizeof(float) * (size_t)N);
  wo_corrlist* lists = (wo_corrlist*)calloc((size_t)N, sizeof(wo_corrlist));
  const float* radius = c->radius;
  const uint32_t seed32 = js_touint32(c->seed);
  for (int32_t i = 0; i < N; i++) {
    if (!c->t_active[i] || !c->c_active[i]) continue;
    const int64_t offset = (int64_t)i * stride;
    const int32_t neighborCount = c->neighborData[offset];
    for (int32_t n = 0; n < neighborCount; n++) {
      const int32_t j = c->neighborData[offset + 1 + n];
      if (i == j || !c->t_active[j] || !c->c_active[j]) continue;
      if (i >= j) continue;
      const double dx = (double)x0[i] - (double)x0[j];
      const double dy = (double)y0[i] - (double)y0[j];
      const double dist2 = dx * dx + dy * dy;
      const double minDist = (double)radius[i] + (double)radius[j];
      if (dist2 >= minDist * minDist) continue;
      const double dist = sqrt(dist2);
      const int eitherIsTrigger = c->isTrigger[i] || c->isTrigger[j];
      const int iStatic = c->rb_static[i], jStatic = c->rb_static[j];
      double mvx = 0, mvy = 0; /* unit move of the pair; i gets +, j gets - */
      int hit = 0;
      if (dist == 0) {
        double cs, sn;
        weed_nudge_dir(weed_nudge_hash((uint32_t)i, (uint32_t)j, c->frame, substep, seed32), &cs, &sn);
        mvx = cs * 0.001; mvy = sn * 0.001;
        hit = 1;
        if (!eitherIsTrigger) {
          if (iStatic && jStatic) {
          } else if (iStatic) {
            corr_push(&lists[j], c->cellOf[i], i, -(mvx * 2), -(mvy * 2));
          } else if (jStatic) {
            corr_push(&lists[i], c->cellOf[j], j, mvx * 2, mvy * 2);
          } else {
            corr_push(&lists[i], c->cellOf[j], j, mvx, mvy);
            corr_push(&lists[j], c->cellOf[i], i, -mvx, -mvy);
          }
        }
      } else {
        const double depth = minDist - dist;
        if (depth > 0) {
          hit = 1;
          if (!eitherIsTrigger) {
            const double nx = dx / dist;
            const double ny = dy / dist;
            const double correction = depth * responseStrength;
            if (iStatic && jStatic) {
            } else if (iStatic) {
              corr_push(&lists[j], c->cellOf[i], i, -(nx * correction), -(ny * correction));
            } else if (jStatic) {
              corr_push(&lists[i], c->cellOf[j], j, nx * correction, ny * correction);
            } else {
              const double halfCorrection = correction * 0.5;
              corr_push(&lists[i], c->cellOf[j], j, nx * halfCorrection, ny * halfCorrection);
              corr_push(&lists[j], c->cellOf[i], i, -(nx * halfCorrection), -(ny * halfCorrection));
            }
          }
        }
      }
      if (hit) {
        c->collisionCount[i]++;
        c->collisionCount[j]++;
        if (collisionData && pairCount < maxPairs) {
          collisionData[1 + pairCount * 2] = i;
          collisionData[1 + pairCount * 2 + 1] = j;
          pairCount++;
        }
      }
    }
  }
  for (int32_t e = 0; e < N; e++) {
    wo_corrlist* l = &lists[e];
    if (l->len) {
      qsort(l->v, (size_t)l->len, sizeof(wo_corr), corr_cmp);
      float xe = c->x[e], ye = c->y[e];
      for (int32_t k = 0; k < l->len; k++) {
        xe = (float)((double)xe + l->v[k].dx);
        ye = (float)((double)ye + l->v[k].dy);
      }
      c->x[e] = xe; c->y[e] = ye;
      free(l->v);
    }
  }
  free(lists); free(x0); free(y0);
  if (collisionData) collisionData[0] = pairCount;
}

/* ---- updateDerivedProperties: physics_worker.js:575-604 ----------------------------- */
static void update_derived(wo_ctx* c) {
  for (int32_t i = 0; i < c->N; i++) {
    if (!c->t_active[i] || !c->rb_active[i]) continue;
    const double vx = c->vx[i], vy = c->vy[i];
    const double currentSpeed = sqrt(vx * vx + vy * vy);
    c->speed[i] = (float)currentSpeed;
    if (currentSpeed > c->minSpeedForRotation)
      c->velocityAngle[i] = (float)(atan2(vy, vx) + M_PI / 2);
  }
}

/* ---- updateVerlet: physics_worker.js:145-233.  order: 0 = reference sweep, 1 = J-order */
WO_EXPORT void wo_physics(wo_ctx* c, double dtRatio, int order) {
  for (int32_t i = 0; i < c->N; i++) { /* :174-177 */
    if (!c->t_active[i] || !c->rb_active[i]) continue;
    c->collisionCount[i] = 0;
  }
  /* `this.settings.gravity.x || 0` (:179-180): NaN and 0 both give 0 */
  const double gx = (c->gravityX == c->gravityX && c->gravityX != 0) ? c->gravityX : 0;
  const double gy = (c->gravityY == c->gravityY && c->gravityY != 0) ? c->gravityY : 0;
  move_balls_verlet(c, dtRatio, gx, gy);
  for (int32_t step = 0; step < c->subStepCount; step++) { /* :203-217 */
    apply_bounds(c);
    if (c->neighborData) {
      if (order == 0) resolve_collisions_reference(c);
      else resolve_collisions_jorder(c, (uint32_t)step);
    }
  }
  update_derived(c);
  c->frame++;
}

/* one lockstep frame (spatial -> physics, dtRatio explicit) */
WO_EXPORT void wo_step(wo_ctx* c, double dtRatio, int order) {
  wo_spatial(c);
  wo_physics(c, dtRatio, order);
}

/* ---- CPU baseline timing ---------------------------------------------------------------
 * The reference runs ONE spatial worker thread and ONE physics worker thread, free-running
 * on the same SharedArrayBuffers with no barrier (AbstractWorker.js:114-146,
 * gameEngine.js:978-996).  wo_bench_freerun reproduces that structure with two pthreads
 * (races included, exactly as in the browser) and reports the wall time for each worker to
 * finish `frames` updates; wo_bench_lockstep is the single-thread lockstep equivalent. */
static double now_s(void) {
  struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
typedef struct { wo_ctx* c; int frames; double dtRatio; double secs; } wo_job;
static void* spatial_thread(void* p) {
  wo_job* j = (wo_job*)p; double t0 = now_s();
  for (int f = 0; f < j->frames; f++) wo_spatial(j->c);
  j->secs = now_s() - t0; return NULL;
}
static void* physics_thread(void* p) {
  wo_job* j = (wo_job*)p; double t0 = now_s();
  for (int f = 0; f < j->frames; f++) wo_physics(j->c, j->dtRatio, 0);
  j->secs = now_s() - t0; return NULL;
}
WO_EXPORT void wo_bench_freerun(wo_ctx* c, int frames, double dtRatio, double* spatial_s,
                                double* physics_s) {
  wo_job a = {c, frames, dtRatio, 0}, b = {c, frames, dtRatio, 0};
  pthread_t ta, tb;
  pthread_create(&ta, NULL, spatial_thread, &a);
  pthread_create(&tb, NULL, physics_thread, &b);
  pthread_join(ta, NULL); pthread_join(tb, NULL);
  *spatial_s = a.secs; *physics_s = b.secs;
}
WO_EXPORT void wo_bench_lockstep(wo_ctx* c, int frames, double dtRatio, double* spatial_s,
                                 double* physics_s) {
  double ts = 0, tp = 0;
  for (int f = 0; f < frames; f++) {
    double t0 = now_s(); wo_spatial(c);
    double t1 = now_s(); wo_physics(c, dtRatio, 0);
    double t2 = now_s(); ts += t1 - t0; tp += t2 - t1;
  }
  *spatial_s = ts; *physics_s = tp;
}

/* exposed for unit tests of the JS helpers */
WO_EXPORT int32_t wo_js_toint32(double v) { return js_toint32(v); }
WO_EXPORT double wo_seeded_random(double seed, int n) {
  wo_rng g = {seed}; double r = 0;
  for (int i = 0; i < n; i++) r = wo_rng_next(&g);
  return r;
}
WO_EXPORT void wo_nudge_dir(uint32_t h, double* c, double* s) { weed_nudge_dir(h, c, s); }
WO_EXPORT uint32_t wo_nudge_hash(uint32_t a, uint32_t b, uint32_t f, uint32_t s, uint32_t seed) {
  return weed_nudge_hash(a, b, f, s, seed);
}
