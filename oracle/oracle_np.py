"""oracle_np.py — second, independent CPU restatement of the WeedJS spatial+physics path.

TEST INFRASTRUCTURE ONLY (see oracle/weed_oracle.c for the rules).  PARITY UNPINNED: the
reference has no tests or golden vectors for this path and no JS engine exists here; this
pure-Python/numpy restatement exists so that two independently written oracles must agree
bit-for-bit (tests/test_oracle_cross.py).  It is written from the reference sources, not
from weed_oracle.c, and deliberately uses different data structures (dict-of-lists grid,
Python big-int ToInt32, explicit pair lists).

Number model: loads widen to Python float (binary64); each Python
float operation is one correctly rounded binary64 op; stores go through np.float32 /
np.uint8 / np.int32.

Citations are relative to the reference tree.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32

# src/components/Transform.js:8-17, RigidBody.js:9-47, Collider.js:8-46
SCHEMAS = {
    "Transform": [("active", 1), ("entityType", 1), ("x", 4), ("y", 4), ("rotation", 4)],
    "RigidBody": [("active", 1), ("static", 1)]
    + [(n, 4) for n in ("vx vy ax ay px py angularVelocity angularAccel mass invMass inertia "
                        "invInertia drag angularDrag maxVel maxAcc minSpeed friction "
                        "velocityAngle speed").split()]
    + [("collisionCount", 1)],
    "Collider": [("active", 1), ("shapeType", 1), ("offsetX", 4), ("offsetY", 4), ("radius", 4),
                 ("width", 4), ("height", 4), ("isTrigger", 1), ("restitution", 4),
                 ("collisionLayer", 2), ("collisionMask", 2), ("aabbMinX", 4), ("aabbMinY", 4),
                 ("aabbMaxX", 4), ("aabbMaxY", 4), ("visualRange", 4)],
}


def layout(component: str, count: int):
    """Component.initializeArrays offsets (src/core/Component.js:20-42). -> ({name: off}, size)"""
    off = 0
    out = {}
    for name, b in SCHEMAS[component]:
        if off % b:
            off += b - off % b
        out[name] = off
        off += count * b
    return out, off


def to_int32(v: float) -> int:
    """ECMAScript ToInt32 (the `| 0` of spatial_worker.js:157-158,214-215)."""
    if not math.isfinite(v):
        return 0
    n = math.trunc(v) & 0xFFFFFFFF
    return n - (1 << 32) if n >= (1 << 31) else n


def imul(a: int, b: int) -> int:
    return to_int32(float(((a & 0xFFFFFFFF) * (b & 0xFFFFFFFF)) & 0xFFFFFFFF))


class SeededRandom:
    """seededRandom, src/core/utils.js:333-342 (t is a JS Number)."""

    def __init__(self, seed: float):
        self.t = float(seed)

    def __call__(self) -> float:
        self.t += float(0x6D2B79F5)
        t = to_int32(self.t)
        tu = t & 0xFFFFFFFF
        r = imul(t ^ (tu >> 15), 1 | t)
        ru = r & 0xFFFFFFFF
        r = to_int32(float(r + imul(r ^ (ru >> 7), 61 | r))) ^ r
        ru = r & 0xFFFFFFFF
        return float((ru ^ (ru >> 14)) & 0xFFFFFFFF) / 4294967296.0


def js_min(a, b):
    if a != a or b != b:
        return math.nan
    return a if a < b else b


def js_max(a, b):
    if a != a or b != b:
        return math.nan
    return a if a > b else b


def _mix32(h):
    h &= 0xFFFFFFFF
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & 0xFFFFFFFF
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & 0xFFFFFFFF
    h ^= h >> 16
    return h


def nudge_hash(lo, hi, frame, substep, seed):
    """include/weed_nudge.h weed_nudge_hash (documented deviation, J-order only)."""
    h = _mix32(lo ^ 0x9E3779B9)
    h = _mix32(h ^ ((hi * 0x7FEB352D + 0x846CA68B) & 0xFFFFFFFF))
    h = _mix32(h ^ ((frame * 0x2C1B3C6D + substep * 0x297A2D39) & 0xFFFFFFFF))
    h = _mix32(h ^ (seed & 0xFFFFFFFF))
    return h


def nudge_dir(h):
    """include/weed_nudge.h weed_nudge_dir."""
    hq = (h + 0x20000000) & 0xFFFFFFFF
    q = hq >> 30
    rem = float(hq & 0x3FFFFFFF) * (1.0 / 1073741824.0)
    a = (rem + -0.5) * 1.5707963267948966
    a2 = a * a
    c = -1.1470745597729725e-11
    for k in (2.08767569878681e-09, -2.755731922398589e-07, 2.48015873015873e-05,
              -0.001388888888888889, 0.041666666666666664, -0.5, 1.0):
        c = c * a2 + k
    s = -7.647163731819816e-13
    for k in (1.6059043836821613e-10, -2.505210838544172e-08, 2.7557319223985893e-06,
              -0.0001984126984126984, 0.008333333333333333, -0.16666666666666666, 1.0):
        s = s * a2 + k
    s = s * a
    return [(c, s), (-s, c), (-c, -s), (s, -c)][q]


class OracleNP:
    """Columns are numpy arrays in self.col[...] (keys like 'T.x', 'RB.px', 'C.radius')."""

    COLS = {
        "T.active": np.uint8, "T.x": F32, "T.y": F32,
        "RB.active": np.uint8, "RB.static": np.uint8, "RB.vx": F32, "RB.vy": F32,
        "RB.ax": F32, "RB.ay": F32, "RB.px": F32, "RB.py": F32, "RB.maxVel": F32,
        "RB.velocityAngle": F32, "RB.speed": F32, "RB.collisionCount": np.uint8,
        "C.active": np.uint8, "C.radius": F32, "C.isTrigger": np.uint8, "C.visualRange": F32,
    }

    def __init__(self, N, worldWidth, worldHeight, cellSize, maxNeighbors, maxPairs=10000,
                 seed=1.0, physics=None):
        self.N = N
        self.W = float(worldWidth)
        self.H = float(worldHeight)
        # spatial_worker.js:80-86
        self.cellSize = float(cellSize)
        self.inv = 1 / self.cellSize
        self.gridCols = math.ceil(self.W / self.cellSize)
        self.gridRows = math.ceil(self.H / self.cellSize)
        self.M = maxNeighbors
        self.maxPairs = maxPairs
        self.seed = float(seed)
        self.rng = SeededRandom(seed)
        self.frame = 0
        # physics_worker.js:33-40
        self.phys = dict(subStepCount=4, boundaryElasticity=0.8, collisionResponseStrength=0.5,
                         verletDamping=0.995, minSpeedForRotation=0.1, gravityX=0.0, gravityY=0.0)
        if physics:
            self.set_physics(**physics)
        self.col = {k: np.zeros(N, dtype=t) for k, t in self.COLS.items()}
        self.neighborData = np.zeros(N * (1 + maxNeighbors), dtype=np.int32)
        self.distanceData = np.zeros(N * (1 + maxNeighbors), dtype=F32)
        self.collisionData = np.zeros(1 + 2 * maxPairs, dtype=np.int32)
        self.grid = {}
        self.cellOf = np.full(N, -1, dtype=np.int32)

    def set_physics(self, **kw):
        """validatePhysicsConfig, src/core/utils.js:269-301."""
        p = dict(self.phys)
        p.update(kw)
        p["subStepCount"] = max(1, int(p["subStepCount"]))
        for k in ("boundaryElasticity", "collisionResponseStrength", "verletDamping"):
            p[k] = js_max(0.0, js_min(1.0, float(p[k])))
        self.phys = p

    # ---- spatial_worker.js:122-172 -----------------------------------------------------
    def rebuild_grid(self):
        c = self.col
        self.grid = {}
        self.cellOf[:] = -1
        maxCol, maxRow = self.gridCols - 1, self.gridRows - 1
        for i in range(self.N):
            if not c["T.active"][i]:
                continue
            px, py = float(c["T.x"][i]), float(c["T.y"][i])
            if px != px or py != py:
                continue
            col = to_int32(px * self.inv)
            row = to_int32(py * self.inv)
            col = 0 if col < 0 else (maxCol if col > maxCol else col)
            row = 0 if row < 0 else (maxRow if row > maxRow else row)
            cell = row * self.gridCols + col
            self.grid.setdefault(cell, []).append(i)
            self.cellOf[i] = cell

    # ---- spatial_worker.js:178-278 -----------------------------------------------------
    def find_all_neighbors(self):
        c = self.col
        x, y, vr = c["T.x"], c["T.y"], c["C.visualRange"]
        M = self.M
        stride = 1 + M
        for cell in list(self.grid.keys()):  # processing order does not affect any row
            for i in self.grid[cell]:
                myX, myY = float(x[i]), float(y[i])
                myVr = float(vr[i])
                vrSq = myVr * myVr
                cellRadius = myVr * self.inv
                cellRadius = math.ceil(cellRadius) if math.isfinite(cellRadius) else cellRadius
                col = to_int32(myX * self.inv)
                row = to_int32(myY * self.inv)
                off = i * stride
                n = 0
                rowMin, rowMax = row - cellRadius, row + cellRadius
                colMin, colMax = col - cellRadius, col + cellRadius
                startRow = 0 if rowMin < 0 else rowMin
                endRow = self.gridRows - 1 if rowMax >= self.gridRows else rowMax
                startCol = 0 if colMin < 0 else colMin
                endCol = self.gridCols - 1 if colMax >= self.gridCols else colMax
                full = False
                r = startRow
                while r <= endRow and not full:  # NaN bounds -> no iterations
                    cc = startCol
                    while cc <= endCol and not full:
                        for j in self.grid.get(int(r) * self.gridCols + int(cc), ()):
                            if j == i:
                                continue
                            dX = float(x[j]) - myX
                            dY = float(y[j]) - myY
                            d2 = dX * dX + dY * dY
                            if d2 < vrSq and d2 > 0:
                                self.neighborData[off + 1 + n] = j
                                self.distanceData[off + 1 + n] = F32(d2)
                                n += 1
                                if n >= M:
                                    full = True
                                    break
                        cc += 1
                    r += 1
                self.neighborData[off] = n
                self.distanceData[off] = F32(n)

    def spatial(self):
        self.rebuild_grid()
        self.find_all_neighbors()

    def grid_csr(self):
        C = self.gridCols * self.gridRows
        start = np.zeros(C + 1, dtype=np.int32)
        idx = []
        for cell in range(C):
            start[cell] = len(idx)
            idx.extend(self.grid.get(cell, ()))
        start[C] = len(idx)
        return start, np.array(idx, dtype=np.int32)

    # ---- physics_worker.js:240-316 -----------------------------------------------------
    def move_balls(self, dtRatio, gx, gy):
        c = self.col
        damping = self.phys["verletDamping"]
        gscale = dtRatio * dtRatio
        for i in range(self.N):
            if not c["T.active"][i] or not c["RB.active"][i] or c["RB.static"][i]:
                continue
            oldX, oldY = float(c["T.x"][i]), float(c["T.y"][i])
            dx = (oldX - float(c["RB.px"][i])) * damping
            dy = (oldY - float(c["RB.py"][i])) * damping
            dx = dx + (gscale * gx + float(c["RB.ax"][i]) * dtRatio)
            dy = dy + (gscale * gy + float(c["RB.ay"][i]) * dtRatio)
            mv = float(c["RB.maxVel"][i])
            maxSpeed = mv if mv > 0 else 100.0
            dx = js_max(-maxSpeed, js_min(maxSpeed, dx))
            dy = js_max(-maxSpeed, js_min(maxSpeed, dy))
            with np.errstate(all="ignore"):
                c["T.x"][i] = F32(oldX + dx)
                c["T.y"][i] = F32(oldY + dy)
                c["RB.px"][i] = F32(oldX)
                c["RB.py"][i] = F32(oldY)
                c["RB.vx"][i] = F32(_div(dx, dtRatio))
                c["RB.vy"][i] = F32(_div(dy, dtRatio))
            c["RB.ax"][i] = 0
            c["RB.ay"][i] = 0

    # ---- physics_worker.js:344-376 -----------------------------------------------------
    def bounds(self):
        c = self.col
        e = self.phys["boundaryElasticity"]
        x, y, px, py = c["T.x"], c["T.y"], c["RB.px"], c["RB.py"]
        for i in range(self.N):
            if not c["T.active"][i] or not c["RB.active"][i] or c["RB.static"][i]:
                continue
            r = float(c["C.radius"][i])
            with np.errstate(all="ignore"):
                if float(x[i]) < r:
                    x[i] = F32(r)
                    px[i] = F32(float(x[i]) + (float(x[i]) - float(px[i])) * e)
                if float(x[i]) > self.W - r:
                    x[i] = F32(self.W - r)
                    px[i] = F32(float(x[i]) + (float(x[i]) - float(px[i])) * e)
                if float(y[i]) < r:
                    y[i] = F32(r)
                    py[i] = F32(float(y[i]) + (float(y[i]) - float(py[i])) * e)
                if float(y[i]) > self.H - r:
                    y[i] = F32(self.H - r)
                    py[i] = F32(float(y[i]) + (float(y[i]) - float(py[i])) * e)

    def _pairs(self):
        """P in reference sweep order: physics_worker.js:428-444."""
        c = self.col
        stride = 1 + self.M
        for i in range(self.N):
            if not c["T.active"][i] or not c["C.active"][i]:
                continue
            off = i * stride
            for n in range(int(self.neighborData[off])):
                j = int(self.neighborData[off + 1 + n])
                if i == j or not c["T.active"][j] or not c["C.active"][j]:
                    continue
                if i >= j:
                    continue
                yield i, j

    def _pair_moves(self, i, j, xi, yi, xj, yj, substep, jorder):
        """Per-pair body of physics_worker.js:446-560. Returns (hit, move_i, move_j)."""
        c = self.col
        dx = xi - xj
        dy = yi - yj
        dist2 = dx * dx + dy * dy
        minDist = float(c["C.radius"][i]) + float(c["C.radius"][j])
        if dist2 >= minDist * minDist:
            return False, None, None
        if dist2 != dist2:  # NaN falls through every comparison in JS
            return False, None, None
        dist = math.sqrt(dist2)
        trig = bool(c["C.isTrigger"][i]) or bool(c["C.isTrigger"][j])
        iS, jS = bool(c["RB.static"][i]), bool(c["RB.static"][j])
        if dist == 0:
            if jorder:
                cs, sn = nudge_dir(nudge_hash(i, j, self.frame, substep, to_int32(self.seed) & 0xFFFFFFFF))
            else:
                ang = self.rng() * math.pi * 2 if not trig else None
                if ang is None:
                    cs = sn = 0.0
                else:
                    cs, sn = math.cos(ang), math.sin(ang)
            ca, sa = cs * 0.001, sn * 0.001
            if trig or (iS and jS):
                return True, None, None
            if iS:
                return True, None, (-(ca * 2), -(sa * 2))
            if jS:
                return True, (ca * 2, sa * 2), None
            return True, (ca, sa), (-ca, -sa)
        depth = minDist - dist
        if not depth > 0:
            return False, None, None
        if trig or (iS and jS):
            return True, None, None
        nx, ny = dx / dist, dy / dist
        corr = depth * self.phys["collisionResponseStrength"]
        if iS:
            return True, None, (-(nx * corr), -(ny * corr))
        if jS:
            return True, (nx * corr, ny * corr), None
        h = corr * 0.5
        return True, (nx * h, ny * h), (-(nx * h), -(ny * h))

    # ---- physics_worker.js:405-568 (reference order) and the documented J-order ----------
    def collisions(self, substep, order):
        c = self.col
        x, y, cc = c["T.x"], c["T.y"], c["RB.collisionCount"]
        pairs = 0
        if order == 0:
            for i, j in self._pairs():
                hit, mi, mj = self._pair_moves(i, j, float(x[i]), float(y[i]), float(x[j]),
                                               float(y[j]), substep, False)
                if not hit:
                    continue
                if mi:
                    x[i] = F32(float(x[i]) + mi[0])
                    y[i] = F32(float(y[i]) + mi[1])
                if mj:
                    x[j] = F32(float(x[j]) + mj[0])
                    y[j] = F32(float(y[j]) + mj[1])
                cc[i] = (int(cc[i]) + 1) & 255
                cc[j] = (int(cc[j]) + 1) & 255
                if pairs < self.maxPairs:
                    self.collisionData[1 + 2 * pairs] = i
                    self.collisionData[2 + 2 * pairs] = j
                    pairs += 1
        else:
            x0, y0 = x.copy(), y.copy()
            incoming = {}
            for i, j in self._pairs():
                hit, mi, mj = self._pair_moves(i, j, float(x0[i]), float(y0[i]), float(x0[j]),
                                               float(y0[j]), substep, True)
                if not hit:
                    continue
                if mi:
                    incoming.setdefault(i, []).append((int(self.cellOf[j]), j, mi))
                if mj:
                    incoming.setdefault(j, []).append((int(self.cellOf[i]), i, mj))
                cc[i] = (int(cc[i]) + 1) & 255
                cc[j] = (int(cc[j]) + 1) & 255
                if pairs < self.maxPairs:
                    self.collisionData[1 + 2 * pairs] = i
                    self.collisionData[2 + 2 * pairs] = j
                    pairs += 1
            for e, lst in incoming.items():
                lst.sort(key=lambda t: (t[0], t[1]))
                xe, ye = x[e], y[e]
                for _, _, mv in lst:
                    xe = F32(float(xe) + mv[0])
                    ye = F32(float(ye) + mv[1])
                x[e], y[e] = xe, ye
        self.collisionData[0] = pairs

    # ---- physics_worker.js:575-604 -----------------------------------------------------
    def derived(self):
        c = self.col
        for i in range(self.N):
            if not c["T.active"][i] or not c["RB.active"][i]:
                continue
            vx, vy = float(c["RB.vx"][i]), float(c["RB.vy"][i])
            s2 = vx * vx + vy * vy
            sp = math.sqrt(s2) if s2 == s2 and s2 >= 0 else math.nan
            with np.errstate(all="ignore"):
                c["RB.speed"][i] = F32(sp)
                if sp > self.phys["minSpeedForRotation"]:
                    c["RB.velocityAngle"][i] = F32(math.atan2(vy, vx) + math.pi / 2)

    # ---- physics_worker.js:145-233 -----------------------------------------------------
    def physics(self, dtRatio=1.0, order=0):
        c = self.col
        for i in range(self.N):
            if c["T.active"][i] and c["RB.active"][i]:
                c["RB.collisionCount"][i] = 0
        gx = self.phys["gravityX"] or 0.0
        gy = self.phys["gravityY"] or 0.0
        gx = 0.0 if gx != gx else gx
        gy = 0.0 if gy != gy else gy
        self.move_balls(float(dtRatio), gx, gy)
        for step in range(self.phys["subStepCount"]):
            self.bounds()
            self.collisions(step, order)
        self.derived()
        self.frame += 1

    def step(self, dtRatio=1.0, order=0):
        self.spatial()
        self.physics(dtRatio, order)


def _div(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)


# =========================================================================================
# Consumers of the hot path's outputs — second, independent restatement.
# Uses Python containers (an insertion-ordered dict stands in for the JS Set / Map) where
# weed_oracle_systems.c uses sorted arrays.
# =========================================================================================
class CollisionEventsNP:
    """LogicWorker.processCollisionCallbacks with one logic worker (logic_worker.js:429-526)."""

    def __init__(self):
        self.previous = {}   # Set -> dict keyed by Cantor key (insertion ordered)
        self.cache = {}      # collisionPairCache

    @staticmethod
    def key(a, b):           # :417-421 — float arithmetic like a JS number
        a = float(a); b = float(b)
        return ((a + b) * (a + b + 1.0)) / 2.0 + b

    def process(self, collisionData):
        calls = []
        pairCount = int(collisionData[0])
        current = {}
        for i in range(pairCount):
            A = int(collisionData[1 + 2 * i]); B = int(collisionData[2 + 2 * i])
            kab, kba = self.key(A, B), self.key(B, A)
            current.setdefault(kab, None)
            current.setdefault(kba, None)
            if kab not in self.previous:                     # :462-465
                self.cache[kab] = (A, B)
                self.cache[kba] = (B, A)
            t = 1 if kab not in self.previous else 2         # :467
            calls.append((t, A, B))
            calls.append((t, B, A))
        for k in self.previous:                               # :493
            if k not in current:
                pair = self.cache.get(k)
                if pair is None:
                    continue
                calls.append((3, pair[0], pair[1]))
                calls.append((3, pair[1], pair[0]))
                del self.cache[k]                             # :514
        self.previous = current                               # :521-523
        return calls


def screen_visibility_np(active, x, y, zoom, cameraX, cameraY, canvasWidth, canvasHeight, screenX, screenY, isItOnScreen):
    """particle_worker.js:1012-1062, vectorised in binary64; outputs updated in place."""
    act = np.asarray(active) != 0
    sx = np.asarray(x, np.float64) * float(zoom) - float(cameraX) * float(zoom)
    sy = np.asarray(y, np.float64) * float(zoom) - float(cameraY) * float(zoom)
    mx, my = canvasWidth * 0.15, canvasHeight * 0.15
    vis = (sx > -mx) & (sx < canvasWidth + mx) & (sy > -my) & (sy < canvasHeight + my)
    screenX[act] = sx[act].astype(np.float32)
    screenY[act] = sy[act].astype(np.float32)
    isItOnScreen[act] = vis[act].astype(np.uint8)


def shadow_sprites_np(maxNeighbors, neighborData, distanceData, transformActive, worldX, worldY, lightEnabled,
                      lightIntensity, shadowCasterActive, shadowRadius, shadowHeight, isOnScreen, maxLights=20,
                      perLight=15, maxSprites=None):
    """particle_worker.js:861-1003 as plain Python loops."""
    if maxSprites is None:
        maxSprites = maxLights * perLight
    out = {"active": np.zeros(maxSprites, np.uint8)}
    for k in ("radius", "x", "y", "rotation", "scaleX", "scaleY", "alpha"):
        out[k] = np.zeros(maxSprites, np.float32)
    stride = 1 + maxNeighbors
    idx = 0
    lights = 0
    for li in range(len(transformActive)):
        if idx >= maxSprites or lights >= maxLights:
            break
        if not lightEnabled[li] or not transformActive[li] or not isOnScreen[li]:
            continue
        intensity = float(lightIntensity[li])
        if intensity <= 0:
            continue
        lights += 1
        lx, ly = float(worldX[li]), float(worldY[li])
        off = li * stride
        mine = 0
        for k in range(int(neighborData[off])):
            if mine >= perLight or idx >= maxSprites:
                break
            j = int(neighborData[off + 1 + k])
            if not shadowCasterActive[j] or not transformActive[j] or not isOnScreen[j]:
                continue
            d2 = float(distanceData[off + 1 + k])
            r = float(shadowRadius[j]) or 10.0
            if r != r:
                r = 10.0
            h = float(shadowHeight[j])
            if h == 0 or h != h:
                h = r
            dx, dy = float(worldX[j]) - lx, float(worldY[j]) - ly
            dist = math.sqrt(d2) if d2 >= 0 else math.nan
            if dist < 1:
                continue
            inv = 1 / dist
            ratio = dist * 0.00390625
            ratio = 1 if ratio > 1 else ratio
            out["active"][idx] = 1
            out["radius"][idx] = F32(r)
            out["x"][idx] = F32(float(worldX[j]) + (dx * inv) * -r)
            out["y"][idx] = F32(float(worldY[j]) + (dy * inv) * -r)
            out["rotation"][idx] = F32(math.atan2(dy, dx) - 1.5707963267948966)
            out["scaleX"][idx] = F32(r * 0.0714)
            out["scaleY"][idx] = F32((0.3 + ratio * 0.9) * (h * 0.025))
            out["alpha"][idx] = F32(_div(intensity, d2 * 2))
            idx += 1
            mine += 1
    out["count"] = idx
    return out


class PoolNP:
    """GameObject spawn pool (gameObject.js:794-951, 668-690, 1001-1034) with a Python list as the
    free-list stack."""

    def __init__(self, col, startIndex, totalCount, rigidBody=True, collider=True):
        self.col, self.start, self.total = col, startIndex, totalCount
        self.rb, self.cl = rigidBody, collider
        self.free = [startIndex + i for offset in range(8) for i in range(offset, totalCount, 8)]   # :818-832

    def spawn(self, records):
        out = []
        c = self.col
        for x, y, vx, vy in np.asarray(records, np.float32).reshape(-1, 4):
            if not self.free:
                out.append(-1)
                continue
            i = self.free.pop()
            if self.rb:
                c["RB.active"][i] = 1
                for k in ("RB.ax", "RB.ay", "RB.speed", "RB.velocityAngle"):
                    c[k][i] = 0
                c["RB.vx"][i] = vx
                c["RB.vy"][i] = vy
            if self.cl:
                c["C.active"][i] = 1
            c["T.x"][i] = x
            c["T.y"][i] = y
            if self.rb:
                c["RB.px"][i] = F32(float(c["T.x"][i]) - float(c["RB.vx"][i]))
                c["RB.py"][i] = F32(float(c["T.y"][i]) - float(c["RB.vy"][i]))
            c["T.active"][i] = 1
            out.append(i)
        return np.array(out, np.int32)

    def despawn(self, indices):
        n = 0
        c = self.col
        for i in indices:
            if not c["T.active"][i]:
                continue
            c["T.active"][i] = 0
            if self.rb:
                c["RB.active"][i] = 0
            if self.cl:
                c["C.active"][i] = 0
            if len(self.free) < self.total:
                self.free.append(int(i))
            n += 1
        return n

    def despawn_all(self):
        return self.despawn([i for i in range(self.start, self.start + self.total) if self.col["T.active"][i]])

    def available(self):
        return len(self.free)

    def free_list(self):
        return np.array(self.free, np.int32)
