"""Host-side tick() of the boids demo — the CONSUMER of the neighbor rows.

Restates demos/predators/boid.js:115-124 (tick), :137-240 (applyFlockingBehaviors: cohesion,
alignment, separation in one loop over this.neighbors / this.neighborDistances) and :318-341
(keepWithinBounds) in plain Python with the reference's evaluation order: accumulators are
JS Numbers (binary64), every `rbAX[i] += ...` rounds to float32.  tick_all is the plain Boid with
the mouse button up; tick_classes is the whole predators demo: Prey (prey.js:120-189) and
Predator (predator.js:140-215) with their processNeighbor hooks and per-class numbers, then
avoidMouse (boid.js:281-316) and keepWithinBounds.

This is user game code, not part of the accelerated path: it runs on the host between frames,
reads the rows GameObject.updateNeighbors would read (src/core/gameObject.js:700-729) and
writes RigidBody.ax/ay, exactly what the logic worker does in the reference.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
MOUSE_TYPE = 0


def tick_all(col, entityType, neighborData, distanceData, maxNeighbors, worldWidth, worldHeight, dtRatio=1.0,
             centeringFactor=0.001, avoidFactor=0.3, matchingFactor=0.1, turnFactor=0.01, margin=20.0):
    """col: dict of numpy columns ('T.x', 'T.y', 'RB.vx', 'RB.vy', 'RB.ax', 'RB.ay', 'C.radius', 'T.active')."""
    tX, tY, vX, vY = col["T.x"], col["T.y"], col["RB.vx"], col["RB.vy"]
    aX, aY = col["RB.ax"], col["RB.ay"]
    stride = 1 + maxNeighbors
    N = len(tX)
    for i in range(1, N):                       # index 0 is the Mouse (its tick() is empty)
        if not col["T.active"][i]:
            continue
        off = i * stride
        cnt = int(neighborData[off])
        myX, myY = float(tX[i]), float(tY[i])
        if cnt:
            pr = float(F32(col["C.radius"][i] * F32(2)))          # boid.js:64 stored in a Float32Array
            pr2 = pr * pr
            cx = cy = avx = avy = sx = sy = 0.0
            same = 0
            myType = entityType[i]
            for n in range(cnt):
                j = int(neighborData[off + 1 + n])
                nt = entityType[j]
                if nt == MOUSE_TYPE:
                    continue
                d2 = float(distanceData[off + 1 + n])
                dx = float(tX[j]) - myX
                dy = float(tY[j]) - myY
                if d2 < pr2 and d2 > 0:
                    sx -= dx / d2
                    sy -= dy / d2
                    continue
                if nt == myType:
                    cx += float(tX[j]); cy += float(tY[j])
                    avx += float(vX[j]); avy += float(vY[j])
                    same += 1
            if same:
                cx /= same; cy /= same
                aX[i] = F32(float(aX[i]) + (cx - myX) * centeringFactor * dtRatio)
                aY[i] = F32(float(aY[i]) + (cy - myY) * centeringFactor * dtRatio)
                avx /= same; avy /= same
                aX[i] = F32(float(aX[i]) + (avx - float(vX[i])) * matchingFactor * dtRatio)
                aY[i] = F32(float(aY[i]) + (avy - float(vY[i])) * matchingFactor * dtRatio)
            aX[i] = F32(float(aX[i]) + sx * avoidFactor * dtRatio)
            aY[i] = F32(float(aY[i]) + sy * avoidFactor * dtRatio)
        if myX < margin:
            aX[i] = F32(float(aX[i]) + turnFactor * dtRatio)
        if myX > worldWidth - margin:
            aX[i] = F32(float(aX[i]) - turnFactor * dtRatio)
        if myY < margin:
            aY[i] = F32(float(aY[i]) + turnFactor * dtRatio)
        if myY > worldHeight - margin:
            aY[i] = F32(float(aY[i]) - turnFactor * dtRatio)


def _div(a, b):
    """JS Number division (no ZeroDivisionError)."""
    if b == 0:
        if a == 0 or a != a:
            return float("nan")
        return float("inf") if (a > 0) == (np.copysign(1.0, b) > 0) else float("-inf")
    return a / b


def tick_classes(col, entityType, neighborData, distanceData, maxNeighbors, worldWidth, worldHeight, classes,
                 dtRatio=1.0, mouseDown=False, mouseType=MOUSE_TYPE):
    """classes: list of dicts as scenes.PREDATORS_DEMO_CLASSES (entityType, role, otherEntityType,
    protectedRangeScale, centeringFactor, avoidFactor, matchingFactor, turnFactor, margin, roleFactor)."""
    tX, tY, vX, vY = col["T.x"], col["T.y"], col["RB.vx"], col["RB.vy"]
    aX, aY = col["RB.ax"], col["RB.ay"]
    stride = 1 + maxNeighbors
    by_type = {c["entityType"]: c for c in classes}
    for i in range(1, len(tX)):
        if not col["T.active"][i]:
            continue
        k = by_type.get(int(entityType[i]))
        if k is None:
            continue
        off = i * stride
        cnt = int(neighborData[off])
        myX, myY = float(tX[i]), float(tY[i])
        if cnt:
            pr = float(F32(float(col["C.radius"][i]) * k["protectedRangeScale"]))     # stored in a Float32Array
            pr2 = pr * pr
            cx = cy = avx = avy = sx = sy = 0.0
            fleeX = fleeY = 0.0
            predators = 0
            closest, closest2 = -1, float("inf")
            same = 0
            myType = entityType[i]
            for n in range(cnt):
                j = int(neighborData[off + 1 + n])
                nt = entityType[j]
                if nt == mouseType:
                    continue
                d2 = float(distanceData[off + 1 + n])
                dx = float(tX[j]) - myX
                dy = float(tY[j]) - myY
                if d2 < pr2 and d2 > 0:
                    sx -= dx / d2
                    sy -= dy / d2
                    continue
                if nt == myType:
                    cx += float(tX[j]); cy += float(tY[j])
                    avx += float(vX[j]); avy += float(vY[j])
                    same += 1
                if k["role"] == "prey":                                   # prey.js:154-169
                    if nt == k["otherEntityType"] and d2 > 0:
                        fleeX += -dx / d2
                        fleeY += -dy / d2
                        predators += 1
                elif k["role"] == "predator":                             # predator.js:172-187
                    if nt == k["otherEntityType"] and d2 < closest2:
                        closest2 = d2
                        closest = j
            if same:
                cx /= same; cy /= same
                aX[i] = F32(float(aX[i]) + (cx - myX) * k["centeringFactor"] * dtRatio)
                aY[i] = F32(float(aY[i]) + (cy - myY) * k["centeringFactor"] * dtRatio)
                avx /= same; avy /= same
                aX[i] = F32(float(aX[i]) + (avx - float(vX[i])) * k["matchingFactor"] * dtRatio)
                aY[i] = F32(float(aY[i]) + (avy - float(vY[i])) * k["matchingFactor"] * dtRatio)
            aX[i] = F32(float(aX[i]) + sx * k["avoidFactor"] * dtRatio)
            aY[i] = F32(float(aY[i]) + sy * k["avoidFactor"] * dtRatio)
            if k["role"] == "prey" and predators > 0:                     # prey.js:176-189
                aX[i] = F32(float(aX[i]) + fleeX * k["roleFactor"] * dtRatio)
                aY[i] = F32(float(aY[i]) + fleeY * k["roleFactor"] * dtRatio)
            if k["role"] == "predator" and closest != -1:                 # predator.js:195-215
                dx = float(tX[closest]) - myX
                dy = float(tY[closest]) - myY
                dist = float(np.sqrt(closest2))
                if dist > 0:
                    aX[i] = F32(float(aX[i]) + (dx / dist) * k["roleFactor"] * dtRatio)
                    aY[i] = F32(float(aY[i]) + (dy / dist) * k["roleFactor"] * dtRatio)
        if mouseDown:                                                     # boid.js:281-316
            for n in range(cnt):
                if int(neighborData[off + 1 + n]) != 0:
                    continue
                d2 = float(distanceData[off + 1 + n])
                if not d2 or d2 != d2:
                    break
                dx = float(tX[0]) - myX
                dy = float(tY[0]) - myY
                aX[i] = F32(float(aX[i]) - (dx / d2) * 1000 * dtRatio)
                aY[i] = F32(float(aY[i]) - (dy / d2) * 1000 * dtRatio)
                break
        margin, turn = k["margin"], k["turnFactor"]
        if myX < margin:
            aX[i] = F32(float(aX[i]) + turn * dtRatio)
        if myX > worldWidth - margin:
            aX[i] = F32(float(aX[i]) - turn * dtRatio)
        if myY < margin:
            aY[i] = F32(float(aY[i]) + turn * dtRatio)
        if myY > worldHeight - margin:
            aY[i] = F32(float(aY[i]) - turn * dtRatio)
