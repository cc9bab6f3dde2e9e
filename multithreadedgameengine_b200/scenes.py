"""Synthetic scenes for the workloads BASELINE.json names (SURVEY §8 d).

The reference demos draw positions and radii from unseeded Math.random(); here every scene
is re-seeded (numpy PCG64) and every value is rounded to float32 before use, so the oracle
and the GPU path start from identical bits.  Each generator returns (config, columns):
``config`` uses the reference's own nesting (worldWidth/worldHeight, spatial{cellSize,
maxNeighbors}, physics{...}) and ``columns`` maps 'T.x', 'RB.px', 'C.radius', ... to arrays of
length entityCount (index 0 is the Mouse entity, src/core/gameEngine.js:280).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def _blank(N):
    z = lambda dt: np.zeros(N, dtype=dt)
    return {
        "T.active": z(np.uint8), "T.x": z(F32), "T.y": z(F32),
        "RB.active": z(np.uint8), "RB.static": z(np.uint8), "RB.vx": z(F32), "RB.vy": z(F32),
        "RB.ax": z(F32), "RB.ay": z(F32), "RB.px": z(F32), "RB.py": z(F32), "RB.maxVel": z(F32),
        "RB.velocityAngle": z(F32), "RB.speed": z(F32), "RB.collisionCount": z(np.uint8),
        "C.active": z(np.uint8), "C.radius": z(F32), "C.isTrigger": z(np.uint8),
        "C.visualRange": z(F32),
    }


def _mouse(c, x=0.0, y=0.0):
    """Entity 0: src/core/Mouse.js:139-145 (Collider only: trigger, radius 0, visualRange 150)."""
    c["T.active"][0] = 1
    c["T.x"][0] = x
    c["T.y"][0] = y
    c["C.active"][0] = 1
    c["C.isTrigger"][0] = 1
    c["C.radius"][0] = 0
    c["C.visualRange"][0] = 150


def _balls(c, sl, x, y, radius, vr, maxVel):
    """GameObject.spawn + Ball.onSpawned (src/core/gameObject.js:840-951, demos/balls/ball.js)."""
    c["T.active"][sl] = 1
    c["T.x"][sl] = x
    c["T.y"][sl] = y
    c["RB.active"][sl] = 1
    c["RB.px"][sl] = c["T.x"][sl]  # px = x - vx with vx = 0
    c["RB.py"][sl] = c["T.y"][sl]
    c["RB.maxVel"][sl] = maxVel
    c["C.active"][sl] = 1
    c["C.radius"][sl] = radius
    c["C.visualRange"][sl] = vr


def balls_readme(n_balls=1000, seed=1234, world=(3000.0, 1500.0), cellSize=50.0, maxNeighbors=400,
                 subStepCount=2):
    """BASELINE config 1: README scene (README.md:148,175-191)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    N = n_balls + 1
    c = _blank(N)
    _mouse(c)
    W, H = world
    x = (rng.random(n_balls) * W).astype(F32)
    y = (rng.random(n_balls) * H).astype(F32)
    radius = (rng.random(n_balls) * 20 + 10).astype(F32)       # ball.js:66
    _balls(c, slice(1, N), x, y, radius, F32(cellSize * 1.33), 50)  # ball.js:23,34
    cfg = dict(entityCount=N, worldWidth=W, worldHeight=H, seed=1234,
               spatial=dict(cellSize=cellSize, maxNeighbors=maxNeighbors),
               physics=dict(subStepCount=subStepCount, gravity=dict(x=0.0, y=0.5), verletDamping=0.99))
    return cfg, c


def balls_demo(n_balls=10000, seed=1234):
    """Config 1b: demos/balls/index.html:97-134 as shipped."""
    cfg, c = balls_readme(n_balls, seed, world=(9000.0, 4000.0), cellSize=50.0, maxNeighbors=900)
    cfg["physics"].update(boundaryElasticity=0.0, collisionResponseStrength=0.8, maxCollisionPairs=0)
    return cfg, c


def boids(n_prey=10000, n_pred=500, seed=1234):
    """BASELINE config 2: demos/predators (index.html:304-380, prey.js:41-47,94-98,
    predator.js:46,54,80-82).  Start velocities are small random vectors (px = x - vx)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    N = 1 + n_prey + n_pred
    c = _blank(N)
    _mouse(c)
    W, H = 5000.0, 2000.0
    n = n_prey + n_pred
    x = (rng.random(n) * W).astype(F32)
    y = (rng.random(n) * H).astype(F32)
    s = (0.85 + 0.3 * rng.random(n_prey))
    radius = np.concatenate([(10 * s * s), np.full(n_pred, 30.0)]).astype(F32)
    vr = np.concatenate([60 + rng.random(n_prey) * 100, np.full(n_pred, 250.0)]).astype(F32)
    maxVel = np.concatenate([1.5 + 2 * rng.random(n_prey), np.full(n_pred, 20.0)]).astype(F32)
    _balls(c, slice(1, N), x, y, radius, vr, maxVel)
    vx = ((rng.random(n) - 0.5) * 2).astype(F32)
    vy = ((rng.random(n) - 0.5) * 2).astype(F32)
    c["RB.vx"][1:] = vx
    c["RB.vy"][1:] = vy
    c["RB.px"][1:] = (c["T.x"][1:].astype(np.float64) - vx).astype(F32)  # gameObject.js:936-939
    c["RB.py"][1:] = (c["T.y"][1:].astype(np.float64) - vy).astype(F32)
    cfg = dict(entityCount=N, worldWidth=W, worldHeight=H, seed=1234,
               spatial=dict(cellSize=128.0, maxNeighbors=1500),
               physics=dict(subStepCount=1, gravity=dict(x=0.0, y=0.0), verletDamping=0.99,
                            boundaryElasticity=0.0, collisionResponseStrength=0.9,
                            maxCollisionPairs=1000000))
    return cfg, c


def balls_synthetic(n_balls, world, cellSize, maxNeighbors, subStepCount, radius, visualRange,
                    seed=1234, clusters=0, cluster_sigma=400.0, cluster_frac=0.5, maxVel=50.0,
                    gravity=(0.0, 0.5), damping=0.99, cluster_edge="clip"):
    """Configs 3-5 (SURVEY §8 d): uniform or mixed radii; optional Gaussian clusters."""
    rng = np.random.Generator(np.random.PCG64(seed))
    N = n_balls + 1
    c = _blank(N)
    _mouse(c)
    W, H = world
    x = rng.random(n_balls) * W
    y = rng.random(n_balls) * H
    if clusters:
        k = int(n_balls * cluster_frac)
        cx = rng.random(clusters) * W
        cy = rng.random(clusters) * H
        which = rng.integers(0, clusters, size=k)
        gx = cx[which] + rng.standard_normal(k) * cluster_sigma
        gy = cy[which] + rng.standard_normal(k) * cluster_sigma
        if cluster_edge == "clip":       # SURVEY §8 d config 4: "clipped to world" (piles entities on the walls)
            x[:k], y[:k] = np.clip(gx, 0, W), np.clip(gy, 0, H)
        else:                            # diagnostic variant: reflect at the walls instead
            gx, gy = np.abs(gx), np.abs(gy)
            x[:k] = np.where(gx > W, 2 * W - gx, gx)
            y[:k] = np.where(gy > H, 2 * H - gy, gy)
    if isinstance(radius, tuple):
        r = (radius[0] + rng.random(n_balls) * (radius[1] - radius[0])).astype(F32)
    else:
        r = np.full(n_balls, radius, dtype=F32)
    _balls(c, slice(1, N), x.astype(F32), y.astype(F32), r, F32(visualRange), maxVel)
    cfg = dict(entityCount=N, worldWidth=float(W), worldHeight=float(H), seed=1234,
               spatial=dict(cellSize=float(cellSize), maxNeighbors=maxNeighbors),
               physics=dict(subStepCount=subStepCount, gravity=dict(x=gravity[0], y=gravity[1]),
                            verletDamping=damping))
    return cfg, c


def config3(n_balls=1_000_000, seed=1234):
    """1M single-GPU roofline case."""
    return balls_synthetic(n_balls, (16384.0, 8192.0), 16.0, 32, 2, 4.0, 16.0, seed)


def config4(n_balls=16_000_000, seed=1234, cluster_edge="clip"):
    """16M mixed radii, dense clustering (the configuration the metric is quoted on)."""
    return balls_synthetic(n_balls, (65536.0, 32768.0), 16.0, 64, 2, (2.0, 6.0), 16.0, seed,
                           clusters=256, cluster_sigma=400.0, cluster_edge=cluster_edge)


def config5(n_balls=128_000_000, seed=1234):
    """128M weak scaling (8 GPUs)."""
    return balls_synthetic(n_balls, (65536.0, 32768.0), 8.0, 32, 4, 1.25, 4.0, seed)


_FULL = {
    # name: (n_balls, world, cellSize, maxNeighbors, subStepCount, radius, visualRange, clusters)
    "config3": (1_000_000, (16384.0, 8192.0), 16.0, 32, 2, 4.0, 16.0, 0),
    "config4": (16_000_000, (65536.0, 32768.0), 16.0, 64, 2, (2.0, 6.0), 16.0, 256),
    "config5": (128_000_000, (65536.0, 32768.0), 8.0, 32, 4, 1.25, 4.0, 0),
}


def scaled(name, n_balls, seed=1234):
    """A smaller (or larger) instance of config3/4/5 at the SAME entity density: the world is
    shrunk 2:1 so that n_balls / area matches the full-size scene (cluster count scales too)."""
    n_full, (W, H), cs, M, S, radius, vr, clusters = _FULL[name]
    f = (n_balls / n_full) ** 0.5
    w = max(cs * 4, round(W * f / cs) * cs)
    h = max(cs * 2, round(H * f / cs) * cs)
    k = max(1, round(clusters * n_balls / n_full)) if clusters else 0
    return balls_synthetic(n_balls, (w, h), cs, M, S, radius, vr, seed, clusters=k)


# Flocking numbers of the predators demo, per class (entityType 1 = Prey, 2 = Predator as laid
# out by scenes.boids): demos/predators/prey.js:37,55-60 and predator.js:43,57-62.
PREDATORS_DEMO_CLASSES = [
    dict(entityType=1, role="prey", otherEntityType=2, protectedRangeScale=1.25, centeringFactor=0.0005, avoidFactor=6.0,
         matchingFactor=0.05, turnFactor=0.001, margin=20.0, roleFactor=10.0),
    dict(entityType=2, role="predator", otherEntityType=1, protectedRangeScale=0.0, centeringFactor=0.0, avoidFactor=0.0,
         matchingFactor=0.0, turnFactor=0.1, margin=20.0, roleFactor=0.2),
]
