// weed_napi.cc — thin Node N-API addon over the C ABI of include/weedgpu.h.
//
// The reference-side binding a WeedJS maintainer builds with node-gyp: no logic, only argument
// marshalling.  The SharedArrayBuffers the engine allocates (src/core/gameEngine.js:534-777) are
// passed by reference: napi_get_arraybuffer_info yields the base pointer that weed_bind() pins and
// mirrors on the device; the addon holds a napi_reference on every bound buffer until the context is
// destroyed, so the garbage collector cannot free memory the DMA engine still writes.
// This image has no Node: tests/test_addon.py compiles this file against tests/mock/node_api.h (the
// declarations of the real header) and drives create -> bind -> step -> fetchNeighbors through a fake
// napi_env (tests/mock/fake_napi.cc) into libweedgpu.so on the GPU box.
//
//   const weed = require('./build/Release/weed_napi.node');
//   const ctx = weed.create({entityCount, worldWidth, worldHeight, cellSize, maxNeighbors,
//                            maxCollisionPairs, seed, physics: {...}});
//   weed.bind(ctx, weed.BUF_TRANSFORM, buffers.componentData.Transform);   // SharedArrayBuffer
//   ...
//   weed.step(ctx, dtRatio, uploadMask, downloadMask);
#include <node_api.h>

#include <cstring>

#include "../include/weedgpu.h"

#define NAPI_OK(call)                                             \
  do {                                                            \
    if ((call) != napi_ok) {                                      \
      napi_throw_error(env, nullptr, "weed_napi: " #call);        \
      return nullptr;                                             \
    }                                                             \
  } while (0)

// what the external handed to JavaScript points at
struct Handle {
  weed_ctx* ctx;
  napi_ref bufs[WEED_BUF_COUNT];     // keeps every bound (Shared)ArrayBuffer alive
  uint32_t maxPairs;
};

static bool get_double(napi_env env, napi_value obj, const char* key, double* out) {
  napi_value v;
  bool has = false;
  if (napi_has_named_property(env, obj, key, &has) != napi_ok || !has) return false;
  if (napi_get_named_property(env, obj, key, &v) != napi_ok) return false;
  return napi_get_value_double(env, v, out) == napi_ok;
}

static napi_value throw_weed(napi_env env, weed_ctx* ctx, int rc) {
  napi_throw_error(env, nullptr, weed_last_error(ctx));
  (void)rc;
  return nullptr;
}

static void fill_physics(napi_env env, napi_value p, weed_physics_config* out) {
  double d;
  if (get_double(env, p, "subStepCount", &d)) out->subStepCount = (int32_t)d;
  if (get_double(env, p, "boundaryElasticity", &d)) out->boundaryElasticity = d;
  if (get_double(env, p, "collisionResponseStrength", &d)) out->collisionResponseStrength = d;
  if (get_double(env, p, "verletDamping", &d)) out->verletDamping = d;
  if (get_double(env, p, "minSpeedForRotation", &d)) out->minSpeedForRotation = d;
  napi_value g;
  bool has = false;
  if (napi_has_named_property(env, p, "gravity", &has) == napi_ok && has &&
      napi_get_named_property(env, p, "gravity", &g) == napi_ok) {
    if (get_double(env, g, "x", &d)) out->gravityX = d;
    if (get_double(env, g, "y", &d)) out->gravityY = d;
  }
}

// create(config) -> external(ctx)        replaces new Worker(spatial|physics) + "init"
static napi_value Create(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value arg;
  NAPI_OK(napi_get_cb_info(env, info, &argc, &arg, nullptr, nullptr));
  weed_config cfg;
  weed_default_config(&cfg);
  double d;
  if (get_double(env, arg, "entityCount", &d)) cfg.entityCount = (uint32_t)d;
  if (get_double(env, arg, "worldWidth", &d)) cfg.worldWidth = d;
  if (get_double(env, arg, "worldHeight", &d)) cfg.worldHeight = d;
  if (get_double(env, arg, "cellSize", &d)) cfg.cellSize = d;
  if (get_double(env, arg, "maxNeighbors", &d)) cfg.maxNeighbors = (uint32_t)d;
  if (get_double(env, arg, "maxCollisionPairs", &d) && d > 0) cfg.maxCollisionPairs = (uint32_t)d;  // `|| 10000`
  if (get_double(env, arg, "seed", &d)) cfg.seed = d;
  if (get_double(env, arg, "device", &d)) cfg.device = (int32_t)d;
  napi_value p;
  bool has = false;
  if (napi_has_named_property(env, arg, "physics", &has) == napi_ok && has &&
      napi_get_named_property(env, arg, "physics", &p) == napi_ok)
    fill_physics(env, p, &cfg.physics);
  weed_ctx* ctx = nullptr;
  const int rc = weed_create(&cfg, &ctx);
  if (rc != WEED_OK) return throw_weed(env, nullptr, rc);
  Handle* h = new Handle();
  h->ctx = ctx;
  h->maxPairs = cfg.maxCollisionPairs;
  for (napi_ref& r : h->bufs) r = nullptr;
  napi_value ext;
  NAPI_OK(napi_create_external(
      env, h,
      [](napi_env e, void* data, void*) {
        Handle* hh = (Handle*)data;
        weed_destroy(hh->ctx);                                 // unpins the buffers first ...
        for (napi_ref r : hh->bufs) if (r) napi_delete_reference(e, r);   // ... then lets them go
        delete hh;
      },
      nullptr, &ext));
  return ext;
}

static Handle* handle_of(napi_env env, napi_value v) {
  void* p = nullptr;
  if (napi_get_value_external(env, v, &p) != napi_ok || !p) { napi_throw_error(env, nullptr, "weed_napi: not a context"); return nullptr; }
  return (Handle*)p;
}
#define CTX_OR_RETURN(h, v)          \
  Handle* h = handle_of(env, (v));   \
  if (!h) return nullptr;            \
  weed_ctx* ctx = h->ctx

// bind(ctx, bufferId, SharedArrayBuffer)
static napi_value Bind(napi_env env, napi_callback_info info) {
  size_t argc = 3;
  napi_value a[3];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  CTX_OR_RETURN(h, a[0]);
  int32_t id;
  NAPI_OK(napi_get_value_int32(env, a[1], &id));
  if (id < 0 || id >= WEED_BUF_COUNT) { napi_throw_error(env, nullptr, "weed_napi: bad buffer id"); return nullptr; }
  void* base = nullptr;
  size_t bytes = 0;
  NAPI_OK(napi_get_arraybuffer_info(env, a[2], &base, &bytes));   // also accepts SharedArrayBuffer (N-API >= 8)
  const int rc = weed_bind(ctx, (weed_buffer_id)id, base, bytes);
  if (rc != WEED_OK) return throw_weed(env, ctx, rc);
  if (h->bufs[id]) { napi_delete_reference(env, h->bufs[id]); h->bufs[id] = nullptr; }   // rebinding: the old buffer was unpinned by weed_bind
  NAPI_OK(napi_create_reference(env, a[2], 1, &h->bufs[id]));
  return nullptr;
}

// step(ctx, dtRatio, uploadMask, downloadMask)     replaces both workers' update()
static napi_value Step(napi_env env, napi_callback_info info) {
  size_t argc = 4;
  napi_value a[4];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  CTX_OR_RETURN(h, a[0]);
  (void)h;
  double dt;
  uint32_t up, down;
  NAPI_OK(napi_get_value_double(env, a[1], &dt));
  NAPI_OK(napi_get_value_uint32(env, a[2], &up));
  NAPI_OK(napi_get_value_uint32(env, a[3], &down));
  const int rc = weed_step(ctx, dt, up, down);
  if (rc != WEED_OK) return throw_weed(env, ctx, rc);
  return nullptr;
}

// setPhysics(ctx, partialConfig)      replaces {msg:"updatePhysicsConfig"}
static napi_value SetPhysics(napi_env env, napi_callback_info info) {
  size_t argc = 2;
  napi_value a[2];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  CTX_OR_RETURN(h, a[0]);
  (void)h;
  weed_physics_config p;
  weed_get_physics(ctx, &p);
  fill_physics(env, a[1], &p);
  const int rc = weed_set_physics(ctx, &p);
  if (rc != WEED_OK) return throw_weed(env, ctx, rc);
  return nullptr;
}

// fetchNeighbors(ctx, first, count)   rows -> the bound neighborData / distanceData SABs
static napi_value FetchNeighbors(napi_env env, napi_callback_info info) {
  size_t argc = 3;
  napi_value a[3];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  CTX_OR_RETURN(h, a[0]);
  (void)h;
  uint32_t first, count;
  NAPI_OK(napi_get_value_uint32(env, a[1], &first));
  NAPI_OK(napi_get_value_uint32(env, a[2], &count));
  const int rc = weed_fetch_neighbors(ctx, first, count);
  if (rc != WEED_OK) return throw_weed(env, ctx, rc);
  return nullptr;
}

// ---- device-side consumers (weed_system_*) --------------------------------------------------
// TypedArray -> base pointer.  NULL (with *ok still true) when v is not a typed array: the C ABI treats
// a NULL output as "not wanted".  A typed array of the wrong element type or with fewer than `need`
// elements is an error (*ok = false, exception pending): native code must never write past it.
static void* typed_data(napi_env env, napi_value v, napi_typedarray_type want, size_t need, bool* ok) {
  napi_valuetype t;
  if (napi_typeof(env, v, &t) != napi_ok || t != napi_object) return nullptr;
  bool is = false;
  if (napi_is_typedarray(env, v, &is) != napi_ok || !is) return nullptr;
  void* data = nullptr;
  napi_typedarray_type ty; size_t len; napi_value ab; size_t off;
  if (napi_get_typedarray_info(env, v, &ty, &len, &data, &ab, &off) != napi_ok) return nullptr;
  if (ty != want || len < need) {
    napi_throw_error(env, nullptr, "weed_napi: typed array of the wrong type or too short");
    *ok = false;
    return nullptr;
  }
  return data;
}

// collisionEvents(ctx, stateU8, exitI32, forgetPrevious) -> {pairs, entered, stayed, exited}
// replaces the Set bookkeeping of LogicWorker.processCollisionCallbacks (logic_worker.js:429-526)
static napi_value CollisionEvents(napi_env env, napi_callback_info info) {
  size_t argc = 4;
  napi_value a[4];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  CTX_OR_RETURN(h, a[0]);
  (void)h;
  bool forget = false;
  if (argc > 3) napi_get_value_bool(env, a[3], &forget);
  weed_collision_event_counts c;
  bool ok = true;
  uint8_t* state = (uint8_t*)typed_data(env, a[1], napi_uint8_array, h->maxPairs, &ok);             // one byte per logged pair
  int32_t* exits = (int32_t*)typed_data(env, a[2], napi_int32_array, 1 + 2 * (size_t)h->maxPairs, &ok);
  if (!ok) return nullptr;
  const int rc = weed_system_collision_events(ctx, forget ? WEED_EVENTS_FORGET_PREVIOUS : 0u, &c, state, exits);
  if (rc != WEED_OK) return throw_weed(env, ctx, rc);
  napi_value out, v;
  NAPI_OK(napi_create_object(env, &out));
  const struct { const char* k; uint32_t n; } f[] = {{"pairs", c.pairs}, {"entered", c.entered}, {"stayed", c.stayed}, {"exited", c.exited}};
  for (const auto& e : f) { NAPI_OK(napi_create_uint32(env, e.n, &v)); NAPI_OK(napi_set_named_property(env, out, e.k, v)); }
  return out;
}

// screenVisibility(ctx, cameraDataF32, canvasWidth, canvasHeight, screenX, screenY, isItOnScreen)
// replaces ParticleWorker.updateEntityScreenVisibility (particle_worker.js:1012-1062)
static napi_value ScreenVisibility(napi_env env, napi_callback_info info) {
  size_t argc = 7;
  napi_value a[7];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  CTX_OR_RETURN(h, a[0]);
  (void)h;
  bool ok = true;
  const float* cam = (const float*)typed_data(env, a[1], napi_float32_array, 3, &ok);
  if (!ok) return nullptr;
  if (!cam) { napi_throw_error(env, nullptr, "weed_napi: cameraData must be a Float32Array"); return nullptr; }
  weed_camera c{cam[0], cam[1], cam[2], 0, 0};
  NAPI_OK(napi_get_value_double(env, a[2], &c.canvasWidth));
  NAPI_OK(napi_get_value_double(env, a[3], &c.canvasHeight));
  const uint32_t N = weed_entity_count(ctx);
  float* sx = (float*)typed_data(env, a[4], napi_float32_array, N, &ok);
  float* sy = (float*)typed_data(env, a[5], napi_float32_array, N, &ok);
  uint8_t* on = (uint8_t*)typed_data(env, a[6], napi_uint8_array, N, &ok);
  if (!ok) return nullptr;
  const int rc = weed_system_screen_visibility(ctx, &c, sx, sy, on);
  if (rc != WEED_OK) return throw_weed(env, ctx, rc);
  return nullptr;
}

// spawn(ctx, pool, recordsF32 /* x, y, vx, vy per entity */, indicesI32)   GameObject.spawn for a batch
static napi_value Spawn(napi_env env, napi_callback_info info) {
  size_t argc = 4;
  napi_value a[4];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, nullptr, nullptr));
  CTX_OR_RETURN(h, a[0]);
  (void)h;
  uint32_t pool;
  NAPI_OK(napi_get_value_uint32(env, a[1], &pool));
  napi_typedarray_type ty; size_t len = 0; void* recs = nullptr; napi_value ab; size_t off;
  NAPI_OK(napi_get_typedarray_info(env, a[2], &ty, &len, &recs, &ab, &off));
  if (ty != napi_float32_array) { napi_throw_error(env, nullptr, "weed_napi: records must be a Float32Array"); return nullptr; }
  bool ok = true;
  int32_t* idx = (int32_t*)typed_data(env, a[3], napi_int32_array, len / 4, &ok);      // one index per record
  if (!ok) return nullptr;
  if (!idx) { napi_throw_error(env, nullptr, "weed_napi: indices must be an Int32Array"); return nullptr; }
  const int rc = weed_pool_spawn(ctx, pool, (const weed_spawn_record*)recs, (uint32_t)(len / 4), idx);
  if (rc != WEED_OK) return throw_weed(env, ctx, rc);
  return nullptr;
}

static napi_value Init(napi_env env, napi_value exports) {
  napi_property_descriptor props[] = {
      {"create", nullptr, Create, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"bind", nullptr, Bind, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"step", nullptr, Step, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"setPhysics", nullptr, SetPhysics, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"fetchNeighbors", nullptr, FetchNeighbors, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"collisionEvents", nullptr, CollisionEvents, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"screenVisibility", nullptr, ScreenVisibility, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"spawn", nullptr, Spawn, nullptr, nullptr, nullptr, napi_default, nullptr},
  };
  napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
  return exports;
}

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
