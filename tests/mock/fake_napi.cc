// A fake napi_env: just enough of a JavaScript value model (numbers, booleans, objects with named
// properties, externals, (Shared)ArrayBuffers, typed arrays, references) to call the functions
// addon/weed_napi.cc exports the way Node would, and a small driver that runs one scene through
//   create -> bind x6 -> step(upload everything, download everything) x frames -> fetchNeighbors
// and dumps the final columns and rows for tests/test_addon.py to compare with the ctypes binding.
// Test infrastructure only.  Usage:  addon_harness <out.bin> [frames]      (exit 3: create threw)
#include <node_api.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/weedgpu.h"

struct napi_value__ {
  napi_valuetype type = napi_undefined;
  double num = 0;
  bool b = false;
  std::map<std::string, napi_value> props;
  napi_callback fn = nullptr;
  void* data = nullptr;            // external pointer / buffer base
  size_t len = 0;                  // buffer bytes / typed array elements
  bool isArrayBuffer = false, isTyped = false;
  napi_typedarray_type ta = napi_uint8_array;
  napi_finalize fin = nullptr;
  int refs = 0;
};
struct napi_ref__ { napi_value v; };
struct napi_env__ { bool pending = false; std::string message; std::vector<napi_value> all; };
struct napi_callback_info__ { std::vector<napi_value> args; };

static napi_value mk(napi_env env) { napi_value v = new napi_value__(); env->all.push_back(v); return v; }

extern "C" {
napi_status napi_throw_error(napi_env env, const char*, const char* msg) { env->pending = true; env->message = msg ? msg : ""; return napi_ok; }
napi_status napi_has_named_property(napi_env, napi_value o, const char* k, bool* r) { if (!o || o->type != napi_object) return napi_object_expected; *r = o->props.count(k) != 0; return napi_ok; }
napi_status napi_get_named_property(napi_env, napi_value o, const char* k, napi_value* r) { if (!o || o->type != napi_object) return napi_object_expected; auto it = o->props.find(k); if (it == o->props.end()) return napi_generic_failure; *r = it->second; return napi_ok; }
napi_status napi_set_named_property(napi_env, napi_value o, const char* k, napi_value v) { if (!o || o->type != napi_object) return napi_object_expected; o->props[k] = v; return napi_ok; }
napi_status napi_get_value_double(napi_env, napi_value v, double* r) { if (!v || v->type != napi_number) return napi_number_expected; *r = v->num; return napi_ok; }
napi_status napi_get_value_int32(napi_env, napi_value v, int32_t* r) { if (!v || v->type != napi_number) return napi_number_expected; *r = (int32_t)v->num; return napi_ok; }
napi_status napi_get_value_uint32(napi_env, napi_value v, uint32_t* r) { if (!v || v->type != napi_number) return napi_number_expected; *r = (uint32_t)v->num; return napi_ok; }
napi_status napi_get_value_bool(napi_env, napi_value v, bool* r) { if (!v || v->type != napi_boolean) return napi_boolean_expected; *r = v->b; return napi_ok; }
napi_status napi_get_cb_info(napi_env, napi_callback_info info, size_t* argc, napi_value* argv, napi_value*, void**) {
  const size_t want = *argc;
  for (size_t k = 0; k < want; k++) argv[k] = k < info->args.size() ? info->args[k] : nullptr;
  *argc = info->args.size();
  return napi_ok;
}
napi_status napi_create_external(napi_env env, void* data, napi_finalize fin, void*, napi_value* r) { napi_value v = mk(env); v->type = napi_external; v->data = data; v->fin = fin; *r = v; return napi_ok; }
napi_status napi_get_value_external(napi_env, napi_value v, void** r) { if (!v || v->type != napi_external) return napi_invalid_arg; *r = v->data; return napi_ok; }
napi_status napi_get_arraybuffer_info(napi_env, napi_value v, void** data, size_t* bytes) { if (!v || !v->isArrayBuffer) return napi_invalid_arg; *data = v->data; *bytes = v->len; return napi_ok; }
napi_status napi_typeof(napi_env, napi_value v, napi_valuetype* r) { *r = v ? v->type : napi_undefined; return napi_ok; }
napi_status napi_is_typedarray(napi_env, napi_value v, bool* r) { *r = v && v->isTyped; return napi_ok; }
napi_status napi_get_typedarray_info(napi_env, napi_value v, napi_typedarray_type* t, size_t* len, void** data, napi_value* ab, size_t* off) {
  if (!v || !v->isTyped) return napi_invalid_arg;
  if (t) *t = v->ta;
  if (len) *len = v->len;
  if (data) *data = v->data;
  if (ab) *ab = nullptr;
  if (off) *off = 0;
  return napi_ok;
}
napi_status napi_create_object(napi_env env, napi_value* r) { napi_value v = mk(env); v->type = napi_object; *r = v; return napi_ok; }
napi_status napi_create_uint32(napi_env env, uint32_t x, napi_value* r) { napi_value v = mk(env); v->type = napi_number; v->num = x; *r = v; return napi_ok; }
napi_status napi_define_properties(napi_env env, napi_value o, size_t n, const napi_property_descriptor* p) {
  for (size_t k = 0; k < n; k++) { napi_value f = mk(env); f->type = napi_function; f->fn = p[k].method; o->props[p[k].utf8name] = f; }
  return napi_ok;
}
napi_status napi_create_reference(napi_env, napi_value v, uint32_t, napi_ref* r) { if (!v) return napi_invalid_arg; v->refs++; *r = new napi_ref__{v}; return napi_ok; }
napi_status napi_delete_reference(napi_env, napi_ref r) { if (!r) return napi_invalid_arg; r->v->refs--; delete r; return napi_ok; }
napi_value weed_napi_test_init(napi_env env, napi_value exports);
}

// ---- the "JavaScript" side ------------------------------------------------------------------------------
static napi_value num(napi_env e, double x) { napi_value v = mk(e); v->type = napi_number; v->num = x; return v; }
static napi_value obj(napi_env e) { napi_value v = mk(e); v->type = napi_object; return v; }
static napi_value sab(napi_env e, void* p, size_t bytes) { napi_value v = mk(e); v->type = napi_object; v->isArrayBuffer = true; v->data = p; v->len = bytes; return v; }
static napi_value call(napi_env e, napi_value exports, const char* name, std::vector<napi_value> args) {
  napi_callback_info__ info{args};
  return exports->props.at(name)->fn(e, &info);
}
static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }
static float unit(uint32_t& s) { return (float)(lcg(s) >> 8) * (1.0f / 16777216.0f); }

int main(int argc, char** argv) {
  const char* out = argc > 1 ? argv[1] : "addon_out.bin";
  const int frames = argc > 2 ? atoi(argv[2]) : 3;
  const uint32_t N = 2001, M = 24, P = 5000;
  const double W = 1600, H = 800, cs = 50;
  napi_env__ envS; napi_env env = &envS;
  napi_value exports = obj(env);
  weed_napi_test_init(env, exports);

  napi_value cfg = obj(env), phys = obj(env), grav = obj(env);
  cfg->props["entityCount"] = num(env, N); cfg->props["worldWidth"] = num(env, W); cfg->props["worldHeight"] = num(env, H);
  cfg->props["cellSize"] = num(env, cs); cfg->props["maxNeighbors"] = num(env, M); cfg->props["maxCollisionPairs"] = num(env, P);
  cfg->props["seed"] = num(env, 1234);
  phys->props["subStepCount"] = num(env, 2); phys->props["verletDamping"] = num(env, 0.99);
  grav->props["x"] = num(env, 0); grav->props["y"] = num(env, 0.5);
  phys->props["gravity"] = grav; cfg->props["physics"] = phys;
  napi_value ctx = call(env, exports, "create", {cfg});
  if (env->pending) { fprintf(stderr, "create threw: %s\n", env->message.c_str()); return 3; }

  // the SharedArrayBuffers of gameEngine.js:534-777, laid out by Component.initializeArrays
  const weed_buffer_id ids[6] = {WEED_BUF_TRANSFORM, WEED_BUF_RIGIDBODY, WEED_BUF_COLLIDER, WEED_BUF_NEIGHBOR, WEED_BUF_DISTANCE, WEED_BUF_COLLISION};
  std::vector<std::vector<uint8_t>> mem(6);
  std::vector<napi_value> bufs(6);
  for (int k = 0; k < 6; k++) {
    mem[k].assign(weed_buffer_bytes(ids[k], N, M, P) + 64, 0);
    bufs[k] = sab(env, mem[k].data(), mem[k].size() - 64);
    call(env, exports, "bind", {ctx, num(env, ids[k]), bufs[k]});
    if (env->pending) { fprintf(stderr, "bind threw: %s\n", env->message.c_str()); return 4; }
    if (bufs[k]->refs != 1) { fprintf(stderr, "bound buffer %d is not referenced\n", k); return 5; }
  }
  auto col = [&](int buf, uint32_t c) { return mem[buf].data() + weed_column_offset(ids[buf], c, N); };
  uint8_t* tAct = col(0, 0); float* x = (float*)col(0, 2); float* y = (float*)col(0, 3);
  uint8_t* rbAct = col(1, 0); float* px = (float*)col(1, 6); float* py = (float*)col(1, 7); float* maxVel = (float*)col(1, 16);
  uint8_t* cAct = col(2, 0); float* radius = (float*)col(2, 4); uint8_t* trig = col(2, 7); float* vr = (float*)col(2, 15);
  uint32_t s = 12345u;
  tAct[0] = 1; cAct[0] = 1; trig[0] = 1; vr[0] = 150.f;                        // the Mouse (src/core/Mouse.js:139-145)
  for (uint32_t i = 1; i < N; i++) {
    tAct[i] = rbAct[i] = cAct[i] = 1;
    x[i] = unit(s) * (float)W; y[i] = unit(s) * (float)H;
    px[i] = x[i]; py[i] = y[i];
    radius[i] = 10.f + 20.f * unit(s); vr[i] = 66.5f; maxVel[i] = 50.f;
  }
  const double ALL_IN = 0x000FFFFF, ALL_OUT = 0x000FFFFF | (1u << 24) | (1u << 25);
  for (int f = 0; f < frames; f++) {
    call(env, exports, "step", {ctx, num(env, 1.0), num(env, f == 0 ? ALL_IN : 0), num(env, ALL_OUT)});
    if (env->pending) { fprintf(stderr, "step threw: %s\n", env->message.c_str()); return 6; }
  }
  call(env, exports, "fetchNeighbors", {ctx, num(env, 0), num(env, N)});
  if (env->pending) { fprintf(stderr, "fetchNeighbors threw: %s\n", env->message.c_str()); return 7; }
  // a typed array that is too short must be refused, not written past
  std::vector<uint8_t> shortState(8);
  napi_value ta = mk(env); ta->type = napi_object; ta->isTyped = true; ta->ta = napi_uint8_array; ta->data = shortState.data(); ta->len = shortState.size();
  call(env, exports, "collisionEvents", {ctx, ta, nullptr, nullptr});
  if (!env->pending) { fprintf(stderr, "collisionEvents accepted a state array shorter than maxCollisionPairs\n"); return 8; }
  env->pending = false;

  FILE* fo = fopen(out, "wb");
  if (!fo) return 9;
  fwrite(&N, 4, 1, fo); fwrite(&M, 4, 1, fo);
  fwrite(x, 4, N, fo); fwrite(y, 4, N, fo);
  fwrite(mem[3].data(), 4, (size_t)N * (1 + M), fo);
  fwrite(mem[5].data(), 4, 1 + 2 * (size_t)P, fo);
  fclose(fo);
  // garbage collection of the context: the finalizer destroys it and releases the buffer references
  ctx->fin(env, ctx->data, nullptr);
  for (int k = 0; k < 6; k++) if (bufs[k]->refs != 0) { fprintf(stderr, "buffer %d still referenced after destroy\n", k); return 10; }
  for (napi_value v : env->all) delete v;
  printf("addon harness ok: %u entities, %d frames\n", N, frames);
  return 0;
}
