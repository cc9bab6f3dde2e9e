// node_harness.mjs — run the UNMODIFIED WeedJS spatial_worker.js + physics_worker.js under Node,
// in lockstep (spatial.update(); physics.update(dt, dtRatio=1)), on a scene written by Python,
// and dump the resulting SharedArrayBuffers.  Purpose: pin the CPU oracle (oracle/weed_oracle.c,
// order 0 = the reference's own sweep) against the real JavaScript the day a box has Node.
//
// NOT exercised in this image (no JavaScript engine exists here, see SURVEY.md §0): the Python
// side (baseline/run_node_harness.py, tests/test_node_reference.py) skips when `node` is missing.
//
//   node baseline/node_harness.mjs <reference_root> <scene.json> <scene.bin> <out.bin> <frames>
//
// scene.json: {entityCount, config, sizes:{Transform,RigidBody,Collider,neighbor,collision}}
// scene.bin : Transform | RigidBody | Collider buffers back to back (Component.js layout)
// out.bin   : Transform | RigidBody | Collider | neighborData | distanceData | collisionData
import fs from "node:fs";
import path from "node:path";
import { pathToFileURL } from "node:url";

const [refRoot, sceneJson, sceneBin, outBin, framesArg] = process.argv.slice(2);
const frames = parseInt(framesArg || "1", 10);
const scene = JSON.parse(fs.readFileSync(sceneJson, "utf8"));
const N = scene.entityCount;

// ---- browser-worker shims ---------------------------------------------------------------------
const shim = () => ({ postMessage() {}, onmessage: null, addEventListener() {} });
globalThis.performance ??= { now: () => Date.now() };
globalThis.requestAnimationFrame = () => 0;   // frames are driven by hand below
const url = (rel) => pathToFileURL(path.join(refRoot, rel)).href;

const workers = {};
{
  const { AbstractWorker } = await import(url("src/workers/AbstractWorker.js"));
  // capture the module-private singletons when they report ready (AbstractWorker.js:336-339)
  AbstractWorker.prototype.reportReady = function () { workers[this.constructor.name] = this; };
  AbstractWorker.prototype.reportLog = function () {};
}
const selfSpatial = shim(), selfPhysics = shim();
globalThis.self = selfSpatial;
await import(url("src/workers/spatial_worker.js"));
globalThis.self = selfPhysics;
await import(url("src/workers/physics_worker.js"));

// ---- SharedArrayBuffers exactly as gameEngine.js:534-777 allocates them -----------------------------
const sab = (bytes) => new SharedArrayBuffer(bytes);
const bin = fs.readFileSync(sceneBin);
const buffers = { componentData: {} };
let off = 0;
for (const name of ["Transform", "RigidBody", "Collider"]) {
  const b = sab(scene.sizes[name]);
  new Uint8Array(b).set(bin.subarray(off, off + scene.sizes[name]));
  off += scene.sizes[name];
  buffers.componentData[name] = b;
}
buffers.neighborData = sab(scene.sizes.neighbor);
buffers.distanceData = sab(scene.sizes.neighbor);
buffers.collisionData = sab(scene.sizes.collision);
const init = {
  msg: "init", buffers, entityCount: N, config: scene.config,
  componentPools: { Transform: { count: N }, RigidBody: { count: N }, Collider: { count: N } },
  registeredClasses: [], scriptsToLoad: [], workerPorts: {},
};
globalThis.self = selfSpatial;
await selfSpatial.onmessage({ data: init });
globalThis.self = selfPhysics;
await selfPhysics.onmessage({ data: init });
if (!workers.SpatialWorker || !workers.PhysicsWorker) throw new Error("workers did not initialise");

// ---- lockstep frames (SURVEY Appendix B), dtRatio = 1 ----------------------------------------------
const t0 = process.hrtime.bigint();
for (let f = 0; f < frames; f++) {
  workers.SpatialWorker.update(16.67, 1, false);
  workers.PhysicsWorker.update(16.67, 1, false);
}
const seconds = Number(process.hrtime.bigint() - t0) / 1e9;

const parts = ["Transform", "RigidBody", "Collider"].map((n) => Buffer.from(new Uint8Array(buffers.componentData[n])));
parts.push(Buffer.from(new Uint8Array(buffers.neighborData)), Buffer.from(new Uint8Array(buffers.distanceData)),
           Buffer.from(new Uint8Array(buffers.collisionData)));
fs.writeFileSync(outBin, Buffer.concat(parts));
console.log(JSON.stringify({ frames, seconds, entityCount: N, node: process.version }));
