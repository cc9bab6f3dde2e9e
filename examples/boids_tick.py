"""Host-side tick() of the boids demo — the CONSUMER of the neighbor rows.

Restates demos/predators/boid.js:115-124 (tick), :137-240 (applyFlockingBehaviors: cohesion,
alignment, separation in one loop over this.neighbors / this.neighborDistances) and :318-341
(keepWithinBounds) in plain Python with the reference's evaluation order: accumulators are
JS Numbers (binary64), every `rbAX[i] += ...` rounds to float32.  avoidMouse (:283-316) is a
no-op while the mouse button is up.  Prey/Predator processNeighbor hooks are not included.

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
