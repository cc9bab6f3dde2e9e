"""SoA component columns over shared buffers — host-side mirror of the reference's
Component system (src/core/Component.js, src/components/{Transform,RigidBody,Collider}.js).

Same names and meaning as the reference: ``ARRAY_SCHEMA`` (declaration order = memory
order), ``initializeArrays(buffer, count)`` creates one typed view per column over ONE
buffer with natural alignment (Component.js:20-42), ``getBufferSize(count)``
(Component.js:77-93).  The buffers play the role of the engine's SharedArrayBuffers: the
C ABI (include/weedgpu.h, weed_bind) receives their base address and recomputes the same
offsets natively.
"""
from __future__ import annotations

import numpy as np

Uint8Array = np.uint8
Uint16Array = np.uint16
Float32Array = np.float32
Int32Array = np.int32


class Component:
    ARRAY_SCHEMA: dict = {}
    sharedBuffer = None
    entityCount = 0

    @classmethod
    def initializeArrays(cls, buffer, count):
        """Component.initializeArrays (src/core/Component.js:20-42)."""
        buf = np.frombuffer(buffer, dtype=np.uint8) if not isinstance(buffer, np.ndarray) else buffer.view(np.uint8).reshape(-1)
        if buf.size < cls.getBufferSize(count):
            raise ValueError(f"{cls.__name__}: buffer of {buf.size} B < {cls.getBufferSize(count)} B")
        cls.sharedBuffer = buf
        cls.entityCount = count
        offset = 0
        for name, typ in cls.ARRAY_SCHEMA.items():
            b = np.dtype(typ).itemsize
            rem = offset % b
            if rem:
                offset += b - rem
            setattr(cls, name if name != "static" else "static", buf[offset:offset + count * b].view(typ))
            offset += count * b
        return cls

    @classmethod
    def getBufferSize(cls, count):
        """Component.getBufferSize (src/core/Component.js:77-93)."""
        offset = 0
        for typ in cls.ARRAY_SCHEMA.values():
            b = np.dtype(typ).itemsize
            rem = offset % b
            if rem:
                offset += b - rem
            offset += count * b
        return offset

    @classmethod
    def columnOffset(cls, name, count):
        offset = 0
        for n, typ in cls.ARRAY_SCHEMA.items():
            b = np.dtype(typ).itemsize
            rem = offset % b
            if rem:
                offset += b - rem
            if n == name:
                return offset
            offset += count * b
        raise KeyError(name)


def _fresh(cls_name, schema):
    """Component classes hold their views as class attributes (like the JS statics), so every
    engine instance gets its own subclass."""
    return type(cls_name, (Component,), {"ARRAY_SCHEMA": dict(schema)})


TRANSFORM_SCHEMA = {  # src/components/Transform.js:8-17
    "active": Uint8Array, "entityType": Uint8Array,
    "x": Float32Array, "y": Float32Array, "rotation": Float32Array,
}
RIGIDBODY_SCHEMA = {  # src/components/RigidBody.js:9-47
    "active": Uint8Array, "static": Uint8Array,
    "vx": Float32Array, "vy": Float32Array, "ax": Float32Array, "ay": Float32Array,
    "px": Float32Array, "py": Float32Array,
    "angularVelocity": Float32Array, "angularAccel": Float32Array,
    "mass": Float32Array, "invMass": Float32Array, "inertia": Float32Array, "invInertia": Float32Array,
    "drag": Float32Array, "angularDrag": Float32Array,
    "maxVel": Float32Array, "maxAcc": Float32Array, "minSpeed": Float32Array, "friction": Float32Array,
    "velocityAngle": Float32Array, "speed": Float32Array,
    "collisionCount": Uint8Array,
}
COLLIDER_SCHEMA = {  # src/components/Collider.js:8-46
    "active": Uint8Array, "shapeType": Uint8Array,
    "offsetX": Float32Array, "offsetY": Float32Array, "radius": Float32Array,
    "width": Float32Array, "height": Float32Array,
    "isTrigger": Uint8Array, "restitution": Float32Array,
    "collisionLayer": Uint16Array, "collisionMask": Uint16Array,
    "aabbMinX": Float32Array, "aabbMinY": Float32Array, "aabbMaxX": Float32Array, "aabbMaxY": Float32Array,
    "visualRange": Float32Array,
}


class Transform(Component):
    ARRAY_SCHEMA = TRANSFORM_SCHEMA


class RigidBody(Component):
    ARRAY_SCHEMA = RIGIDBODY_SCHEMA


class Collider(Component):
    ARRAY_SCHEMA = COLLIDER_SCHEMA


def new_component_classes():
    return (_fresh("Transform", TRANSFORM_SCHEMA), _fresh("RigidBody", RIGIDBODY_SCHEMA),
            _fresh("Collider", COLLIDER_SCHEMA))


# keys used by scene generators / tests  ->  (component index, schema name)
COLUMN_KEYS = {
    "T.active": (0, "active"), "T.x": (0, "x"), "T.y": (0, "y"),
    "RB.active": (1, "active"), "RB.static": (1, "static"), "RB.vx": (1, "vx"), "RB.vy": (1, "vy"),
    "RB.ax": (1, "ax"), "RB.ay": (1, "ay"), "RB.px": (1, "px"), "RB.py": (1, "py"),
    "RB.maxVel": (1, "maxVel"), "RB.velocityAngle": (1, "velocityAngle"), "RB.speed": (1, "speed"),
    "RB.collisionCount": (1, "collisionCount"),
    "C.active": (2, "active"), "C.radius": (2, "radius"), "C.isTrigger": (2, "isTrigger"),
    "C.visualRange": (2, "visualRange"), "T.entityType": (0, "entityType"),
}
