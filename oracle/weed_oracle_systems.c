/*
 * weed_oracle_systems.c — CPU restatement of the reference's consumers of the hot path's
 * outputs that the library also offers on the device:
 *   LogicWorker.processCollisionCallbacks      src/workers/logic_worker.js:417-526
 *   ParticleWorker.updateEntityScreenVisibility src/workers/particle_worker.js:1012-1062
 *   ParticleWorker.updateShadowSprites          src/workers/particle_worker.js:861-1003
 *
 * TEST INFRASTRUCTURE ONLY (same rule as weed_oracle.c).  PARITY UNPINNED: the reference has
 * no tests or fixtures for these functions; oracle/oracle_np.py restates them independently
 * and the two must agree.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define WO_EXPORT __attribute__((visibility("default")))

/* ---- f2: Enter / Stay / Exit ------------------------------------------------------------ */
/* A JS Set of numbers iterates in insertion order; membership is by value.  Restated as an
 * insertion-ordered key array plus a sorted copy for lookups.                               */
typedef struct {
  double* keys; int32_t* a; int32_t* b;   /* insertion order; (a,b) = collisionPairCache entry */
  double* sorted;
  int64_t n, cap;
} wo_keyset;
typedef struct { wo_keyset prev, cur; } wo_events;

static double cantor(double a, double b) { /* logic_worker.js:417-421, evaluated in binary64 */
  return ((a + b) * (a + b + 1)) / 2 + b;
}
static int dcmp(const void* x, const void* y) {
  const double a = *(const double*)x, b = *(const double*)y;
  return a < b ? -1 : a > b;
}
static void ks_clear(wo_keyset* s) { s->n = 0; }
static void ks_reserve(wo_keyset* s, int64_t n) {
  if (n <= s->cap) return;
  s->cap = n;
  s->keys = realloc(s->keys, n * sizeof(double)); s->sorted = realloc(s->sorted, n * sizeof(double));
  s->a = realloc(s->a, n * sizeof(int32_t)); s->b = realloc(s->b, n * sizeof(int32_t));
}
static void ks_seal(wo_keyset* s) {
  memcpy(s->sorted, s->keys, s->n * sizeof(double));
  qsort(s->sorted, s->n, sizeof(double), dcmp);
}
static int ks_has(const wo_keyset* s, double k) {
  return s->n && bsearch(&k, s->sorted, s->n, sizeof(double), dcmp) != NULL;
}

WO_EXPORT wo_events* wo_events_create(void) { return calloc(1, sizeof(wo_events)); }
WO_EXPORT void wo_events_destroy(wo_events* e) {
  if (!e) return;
  wo_keyset* s[2] = {&e->prev, &e->cur};
  for (int k = 0; k < 2; k++) { free(s[k]->keys); free(s[k]->sorted); free(s[k]->a); free(s[k]->b); }
  free(e);
}
/* One call of processCollisionCallbacks with a single logic worker.  calls[] receives the
 * callback sequence as (type, self, other) int32 triples: 1 onCollisionEnter, 2 onCollisionStay,
 * 3 onCollisionExit.  Returns the number of triples (counted even beyond cap).              */
WO_EXPORT int64_t wo_events_process(wo_events* e, const int32_t* collisionData, int32_t* calls, int64_t cap) {
  const int32_t pairCount = collisionData[0];                       /* :431 */
  int64_t n = 0;
#define EMIT(t, s, o) do { if (n < cap) { calls[3*n] = (t); calls[3*n+1] = (s); calls[3*n+2] = (o); } n++; } while (0)
  ks_clear(&e->cur);                                                /* :441 */
  ks_reserve(&e->cur, 2 * (int64_t)pairCount);
  for (int32_t i = 0; i < pairCount; i++) {                         /* :444 */
    const int32_t A = collisionData[1 + i * 2], B = collisionData[1 + i * 2 + 1];
    const double keyAB = cantor(A, B), keyBA = cantor(B, A);        /* :456-457 */
    /* Set.add: a duplicate keeps its first position.  Pairs are unique within one frame. */
    e->cur.keys[e->cur.n] = keyAB; e->cur.a[e->cur.n] = A; e->cur.b[e->cur.n] = B; e->cur.n++;
    e->cur.keys[e->cur.n] = keyBA; e->cur.a[e->cur.n] = B; e->cur.b[e->cur.n] = A; e->cur.n++;
    const int isNew = !ks_has(&e->prev, keyAB);                     /* :467 */
    const int t = isNew ? 1 : 2;
    EMIT(t, A, B);                                                  /* :475 / :484 */
    EMIT(t, B, A);                                                  /* :478 / :487 */
  }
  ks_seal(&e->cur);
  for (int64_t k = 0; k < e->prev.n; k++) {                         /* :493 Set iteration order */
    if (!ks_has(&e->cur, e->prev.keys[k])) {                        /* :494 */
      EMIT(3, e->prev.a[k], e->prev.b[k]);                          /* :507 */
      EMIT(3, e->prev.b[k], e->prev.a[k]);                          /* :510 */
    }
  }
  wo_keyset t = e->prev; e->prev = e->cur; e->cur = t;              /* :521-523 */
  return n;
#undef EMIT
}

/* ---- f3a: screen visibility ------------------------------------------------------------- */
WO_EXPORT void wo_screen_visibility(int32_t entityCount, const uint8_t* active, const float* x, const float* y,
                                    double zoom, double cameraX, double cameraY, double canvasWidth,
                                    double canvasHeight, float* screenX, float* screenY, uint8_t* isItOnScreen) {
  const double cameraOffsetX = cameraX * zoom, cameraOffsetY = cameraY * zoom;   /* :1034-1035 */
  const double marginX = canvasWidth * 0.15, marginY = canvasHeight * 0.15;
  const double minX = -marginX, maxX = canvasWidth + marginX, minY = -marginY, maxY = canvasHeight + marginY;
  for (int32_t i = 0; i < entityCount; i++) {
    if (!active[i]) continue;                                       /* :1046 */
    const double sx = x[i] * zoom - cameraOffsetX, sy = y[i] * zoom - cameraOffsetY;
    screenX[i] = (float)sx; screenY[i] = (float)sy;                 /* Float32Array stores */
    isItOnScreen[i] = (sx > minX && sx < maxX && sy > minY && sy < maxY) ? 1 : 0;
  }
}

/* ---- f3b: shadow sprites ---------------------------------------------------------------- */
WO_EXPORT int32_t wo_shadow_sprites(int32_t entityCount, int32_t maxNeighbors, const int32_t* neighborData,
                                    const float* distanceData, const uint8_t* transformActive, const float* worldX,
                                    const float* worldY, const uint8_t* lightEnabled, const float* lightIntensity,
                                    const uint8_t* shadowCasterActive, const float* entityShadowRadius,
                                    const float* entityShadowHeight, const uint8_t* isOnScreen,
                                    int32_t maxShadowCastingLights, int32_t maxShadowsPerLight, int32_t maxShadowSprites,
                                    uint8_t* shadowActive, float* shadowRadius, float* shadowX, float* shadowY,
                                    float* shadowRotation, float* shadowScaleX, float* shadowScaleY, float* shadowAlpha) {
  const int64_t stride = 1 + (int64_t)maxNeighbors;                 /* :896 */
  int32_t shadowIdx = 0, lightsProcessed = 0;
  for (int32_t lightIdx = 0; lightIdx < entityCount; lightIdx++) {  /* :910 */
    if (shadowIdx >= maxShadowSprites) break;
    if (lightsProcessed >= maxShadowCastingLights) break;
    if (!lightEnabled[lightIdx]) continue;
    if (!transformActive[lightIdx]) continue;
    if (!isOnScreen[lightIdx]) continue;
    const double intensity = lightIntensity[lightIdx];
    if (intensity <= 0) continue;                                   /* :918 (NaN passes) */
    lightsProcessed++;
    const double lightX = worldX[lightIdx], lightY = worldY[lightIdx];
    const int64_t offset = lightIdx * stride;
    const int32_t neighborCount = neighborData[offset];
    int32_t shadowsForThisLight = 0;
    for (int32_t k = 0; k < neighborCount; k++) {                   /* :930 */
      if (shadowsForThisLight >= maxShadowsPerLight) break;
      if (shadowIdx >= maxShadowSprites) break;
      const int32_t j = neighborData[offset + 1 + k];
      if (!shadowCasterActive[j]) continue;
      if (!transformActive[j]) continue;
      if (!isOnScreen[j]) continue;
      const double distSq = distanceData[offset + 1 + k];
      const double casterX = worldX[j], casterY = worldY[j];
      const double r0 = entityShadowRadius[j], h0 = entityShadowHeight[j];
      const double casterRadius = (r0 != 0 && r0 == r0) ? r0 : 10;            /* `|| 10` :942 */
      const double casterHeight = (h0 != 0 && h0 == h0) ? h0 : casterRadius;  /* :943 */
      const double dx = casterX - lightX, dy = casterY - lightY;
      const double dist = sqrt(distSq);
      if (dist < 1) continue;                                       /* :951 */
      const double invDist = 1 / dist;
      const double dirX = dx * invDist, dirY = dy * invDist;
      const double posX = casterX + dirX * -casterRadius, posY = casterY + dirY * -casterRadius;
      const double distRatio = dist * 0.00390625;
      const double clampedDistRatio = distRatio > 1 ? 1 : distRatio;
      const double heightFactor = casterHeight * 0.025;
      const double lengthScale = (0.3 + clampedDistRatio * 0.9) * heightFactor;
      const double widthScale = casterRadius * 0.0714;
      const double alpha = intensity / (distSq * 2);
      const double angle = atan2(dy, dx);
      shadowActive[shadowIdx] = 1;
      shadowRadius[shadowIdx] = (float)casterRadius;
      shadowX[shadowIdx] = (float)posX; shadowY[shadowIdx] = (float)posY;
      shadowRotation[shadowIdx] = (float)(angle - 1.5707963267948966);
      shadowScaleX[shadowIdx] = (float)widthScale; shadowScaleY[shadowIdx] = (float)lengthScale;
      shadowAlpha[shadowIdx] = (float)alpha;
      shadowIdx++; shadowsForThisLight++;
    }
  }
  for (int32_t i = shadowIdx; i < maxShadowSprites; i++) shadowActive[i] = 0;   /* :1002-1004 */
  return shadowIdx;
}

/* ---- f4: spawn / despawn pools (src/core/gameObject.js:794-951, 668-690, 1001-1034) -------- */
/* The oracle works on plain column pointers (the SAB views the reference's accessors write). */
typedef struct {
  int32_t startIndex, totalCount, freeListTop;
  int32_t* freeList;
  int hasRigidBody, hasCollider;
} wo_pool;

WO_EXPORT wo_pool* wo_pool_create(int32_t startIndex, int32_t totalCount, int hasRigidBody, int hasCollider) {
  wo_pool* p = calloc(1, sizeof(wo_pool));
  p->startIndex = startIndex; p->totalCount = totalCount;
  p->hasRigidBody = hasRigidBody; p->hasCollider = hasCollider;
  p->freeList = malloc(sizeof(int32_t) * (size_t)totalCount);          /* :802 */
  p->freeListTop = totalCount - 1;                                     /* :803 */
  const int interleaveFactor = 8;                                      /* :821 */
  int32_t writeIndex = 0;
  for (int offset = 0; offset < interleaveFactor; offset++)            /* :828-832 */
    for (int32_t i = offset; i < totalCount; i += interleaveFactor) p->freeList[writeIndex++] = startIndex + i;
  return p;
}
WO_EXPORT void wo_pool_destroy(wo_pool* p) { if (p) { free(p->freeList); free(p); } }
WO_EXPORT int32_t wo_pool_available(const wo_pool* p) { return p->freeListTop + 1; }   /* :968-969 */

typedef struct {
  uint8_t *tActive, *rbActive, *cActive;
  float *x, *y, *vx, *vy, *ax, *ay, *px, *py, *speed, *velocityAngle;
} wo_pool_cols;

/* GameObject.spawn with spawnConfig {x, y, vx, vy}; returns the index or -1 (:868-873) */
WO_EXPORT int32_t wo_pool_spawn(wo_pool* p, const wo_pool_cols* c, float x, float y, float vx, float vy) {
  if (p->freeListTop < 0) return -1;
  const int32_t i = p->freeList[p->freeListTop--];                     /* :876 */
  if (p->hasRigidBody) {                                               /* :888-899 */
    c->rbActive[i] = 1;
    c->ax[i] = 0; c->ay[i] = 0; c->vx[i] = 0; c->vy[i] = 0; c->speed[i] = 0; c->velocityAngle[i] = 0;
    c->px[i] = 0; c->py[i] = 0;
  }
  c->x[i] = 0; c->y[i] = 0;                                            /* :901-905 */
  if (p->hasCollider) c->cActive[i] = 1;                               /* :907-909 */
  c->x[i] = x; c->y[i] = y;                                            /* :927-931 spawnConfig */
  if (p->hasRigidBody) {
    c->vx[i] = vx; c->vy[i] = vy;
    c->px[i] = (float)((double)c->x[i] - (double)c->vx[i]);            /* :936-939 */
    c->py[i] = (float)((double)c->y[i] - (double)c->vy[i]);
  }
  c->tActive[i] = 1;                                                   /* :948 */
  return i;
}
/* GameObject.despawn (:668-690); returns 1 when the entity was active */
WO_EXPORT int wo_pool_despawn(wo_pool* p, const wo_pool_cols* c, int32_t index) {
  if (c->tActive[index] == 0) return 0;                                /* :670 */
  c->tActive[index] = 0;                                               /* :679-681 */
  if (p->hasRigidBody) c->rbActive[index] = 0;
  if (p->hasCollider) c->cActive[index] = 0;
  ++p->freeListTop;                                                    /* :688 */
  if (p->freeListTop < p->totalCount) p->freeList[p->freeListTop] = index;   /* an out-of-range typed-array store is dropped */
  return 1;
}
WO_EXPORT int32_t wo_pool_despawn_all(wo_pool* p, const wo_pool_cols* c) {   /* :1001-1034 */
  int32_t n = 0;
  for (int32_t i = p->startIndex; i < p->startIndex + p->totalCount; i++)
    if (c->tActive[i]) n += wo_pool_despawn(p, c, i);
  return n;
}
WO_EXPORT void wo_pool_free_list(const wo_pool* p, int32_t* out) { memcpy(out, p->freeList, sizeof(int32_t) * (size_t)p->totalCount); }
