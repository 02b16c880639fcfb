"""Binary layouts shared by the C-ABI, the host mirror and the tests.

Every record is the reference's own POD layout (sizes pinned by static_assert on the C++ side and
re-checked here), so arrays can be handed to either implementation unchanged:

* ``RIGID_BODY``   128 B  reference ``RigidBody``          (src/RigidBody.h:12-58)
* ``CONTACT_JOINT`` 20 B  reference ``ContactJoint``       (src/Joints.h:6-23)
* ``CONTACT_POINT`` 32 B  reference ``ContactPoint``       (src/Manifold.h:12-43)
* ``MANIFOLD``      16 B  reference ``Manifold``           (src/Manifold.h:45-68)
* ``BROADPHASE_ENTRY`` 20 B reference ``Collider::BroadphaseEntry`` (src/Collider.h:45-50)
"""
import numpy as np

V2 = (np.float32, (2,))

RIGID_BODY = np.dtype(
    [
        ("index", np.uint32),
        ("size", *V2),  # geom.size (half extents)
        ("geom_xVector", *V2),
        ("geom_yVector", *V2),
        ("geom_pos", *V2),
        ("aabb_min", *V2),
        ("aabb_max", *V2),
        ("velocity", *V2),
        ("acceleration", *V2),
        ("displacingVelocity", *V2),
        ("angularVelocity", np.float32),
        ("angularAcceleration", np.float32),
        ("displacingAngularVelocity", np.float32),
        ("invMass", np.float32),
        ("invInertia", np.float32),
        ("xVector", *V2),  # coords.xVector
        ("yVector", *V2),
        ("pos", *V2),
        ("lastIteration", np.int32),
        ("lastDisplacementIteration", np.int32),
    ]
)

CONTACT_JOINT = np.dtype(
    [
        ("contactPointIndex", np.int32),
        ("body1Index", np.int32),
        ("body2Index", np.int32),
        ("normalImpulse", np.float32),  # normalLimiter_accumulatedImpulse
        ("frictionImpulse", np.float32),  # frictionLimiter_accumulatedImpulse
    ]
)

CONTACT_POINT = np.dtype(
    [
        ("delta1", *V2),
        ("delta2", *V2),
        ("normal", *V2),
        ("isMerged", np.uint8),
        ("isNewlyCreated", np.uint8),
        ("_pad", np.uint8, (2,)),
        ("solverIndex", np.int32),
    ]
)

MANIFOLD = np.dtype(
    [
        ("body1Index", np.int32),
        ("body2Index", np.int32),
        ("pointCount", np.int32),
        ("pointIndex", np.int32),
    ]
)

BROADPHASE_ENTRY = np.dtype(
    [
        ("minx", np.float32),
        ("maxx", np.float32),
        ("centery", np.float32),
        ("extenty", np.float32),
        ("index", np.uint32),
    ]
)

assert RIGID_BODY.itemsize == 128
assert CONTACT_JOINT.itemsize == 20
assert CONTACT_POINT.itemsize == 32
assert MANIFOLD.itemsize == 16
assert BROADPHASE_ENTRY.itemsize == 20

# Fields of RigidBody that carry simulation state (dead ints and the uninitialised padding of a
# freshly constructed record are excluded from comparisons).
BODY_STATE_FIELDS = (
    "velocity",
    "angularVelocity",
    "displacingVelocity",
    "displacingAngularVelocity",
    "acceleration",
    "angularAcceleration",
    "pos",
    "xVector",
    "yVector",
    "geom_pos",
    "geom_xVector",
    "geom_yVector",
    "aabb_min",
    "aabb_max",
    "size",
    "invMass",
    "invInertia",
)

# reference enums (src/Configuration.h:5-18)
SOLVE_SCALAR, SOLVE_SSE2, SOLVE_AVX2 = 0, 1, 2
ISLAND_SINGLE, ISLAND_MULTIPLE, ISLAND_SINGLE_SLOPPY, ISLAND_MULTIPLE_SLOPPY = 0, 1, 2, 3
