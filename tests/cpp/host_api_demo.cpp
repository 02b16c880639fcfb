// A program written the way the reference's demo uses the physics API (reference src/main.cpp:82-93,
// 258-269, 337-364, 372-413), compiled against the host mirror's headers.  It exists to show that the
// surface is source-compatible: same includes, same types, same members, same calls.
#include "World.h"
#include "Configuration.h"
#include "base/WorkQueue.h"

#include <cstdio>
#include <cstdlib>

static void resetWorld(World& world, int rows)
{
    world.bodies.clear();
    world.collider.manifolds.clear();
    world.collider.manifoldMap.clear();
    world.solver.contactJoints.clear();

    RigidBody* groundBody = world.AddBody(Coords2f(Vector2f(0, 0), 0.0f), Vector2f(10000.f, 10.0f));
    groundBody->invInertia = 0.0f;
    groundBody->invMass = 0.0f;

    for (int r = 0; r < rows; ++r)
        for (int i = 0; i < rows - r; ++i)
        {
            Vector2f pos = Vector2f((i - (rows - r) * 0.5f) * 21.f, 15.f + 10.f * r);
            Vector2f size(10, 5);
            world.AddBody(Coords2f(pos, 0.f), size);
        }
}

int main(int argc, char** argv)
{
    int rows = argc > 1 ? atoi(argv[1]) : 10;
    int steps = argc > 2 ? atoi(argv[2]) : 20;

    WorkQueue queue(WorkQueue::getIdealWorkerCount() - 1);
    World world;
    resetWorld(world, rows);
    world.gravity = -200.0f;
    const float physicsTime = 1.0f / 60.0f;

    for (int step = 0; step < steps; ++step)
    {
        // the demo drags a body around by writing its acceleration between steps
        RigidBody* draggedBody = &world.bodies[1];
        Vector2f dstVelocity = (Vector2f(0.f, 40.f) - draggedBody->coords.pos) * 5e-1f;
        draggedBody->acceleration += (dstVelocity - draggedBody->velocity) * 5e0f;

        Configuration configuration = { Configuration::Solve_AVX2, Configuration::Island_Single, 15, 15 };
        world.Update(queue, physicsTime, configuration);
    }

    // what the HUD and the renderer read
    double checksum = 0;
    for (int bodyIndex = 0; bodyIndex < world.bodies.size; bodyIndex++)
    {
        RigidBody* body = &world.bodies[bodyIndex];
        Coords2f bodyCoords = body->coords;
        Vector2f size = body->geom.size;
        checksum += bodyCoords.pos.x * 1e-3 + bodyCoords.pos.y + bodyCoords.xVector.y + size.x * 0 + body->velocity.y * 1e-2;
    }
    int newPoints = 0;
    for (int manifoldIndex = 0; manifoldIndex < world.collider.manifolds.size; manifoldIndex++)
    {
        Manifold& man = world.collider.manifolds[manifoldIndex];
        for (int collisionNumber = 0; collisionNumber < man.pointCount; collisionNumber++)
        {
            ContactPoint& cp = world.collider.contactPoints[man.pointIndex + collisionNumber];
            Vector2f point1 = cp.delta1 + world.bodies[man.body1Index].coords.pos;
            checksum += point1.y * 1e-3;
            newPoints += cp.isNewlyCreated ? 1 : 0;
        }
    }
    printf("bodies %d manifolds %d joints %d islands %d maxsize %d new %d checksum %.6f\n", world.bodies.size, world.collider.manifolds.size,
        world.solver.contactJoints.size, world.solver.islandCount, world.solver.islandMaxSize, newPoints, checksum);
    return 0;
}
