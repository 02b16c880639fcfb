"""Deterministic synthetic scenes (BASELINE.json configs).

Each generator returns a float32 array of rows ``(x, y, angle, half_w, half_h, is_static)`` in body
creation order; body 0 is always the static ground.  No RNG.  The scene shapes follow the
reference demo's box stacks (reference src/main.cpp:82-230: ground + rows of (10,5) half-size
boxes), scaled to the BASELINE sizes.
"""
import numpy as np

BOX = (10.0, 5.0)
GRAVITY = -200.0
DT = 1.0 / 60.0


def _ground(half_width):
    return [(0.0, 0.0, 0.0, float(half_width), 10.0, 1.0)]


def pyramid(rows, pitch=21.0, x0=0.0, ground=True, ground_half_width=1e7):
    """Brick pyramid: row r has rows-r boxes at x=(i-(rows-r)*0.5)*pitch, y=15+10r."""
    out = _ground(ground_half_width) if ground else []
    for r in range(rows):
        n = rows - r
        i = np.arange(n, dtype=np.float64)
        xs = x0 + (i - n * 0.5) * pitch
        for x in xs:
            out.append((float(x), 15.0 + 10.0 * r, 0.0, BOX[0], BOX[1], 0.0))
    return np.asarray(out, dtype=np.float32)


def pyramid_fast(rows, pitch=21.0, x0=0.0, ground_half_width=1e7):
    """Vectorised pyramid() for the 1 M-box scene (rows=1414)."""
    r = np.repeat(np.arange(rows), rows - np.arange(rows))
    n = rows - r
    starts = np.concatenate([[0], np.cumsum(rows - np.arange(rows))[:-1]])
    i = np.arange(r.size) - starts[r]
    xs = x0 + (i - n * 0.5) * pitch
    ys = 15.0 + 10.0 * r
    body = np.zeros((r.size + 1, 6), dtype=np.float32)
    body[0] = (0.0, 0.0, 0.0, ground_half_width, 10.0, 1.0)
    body[1:, 0] = xs
    body[1:, 1] = ys
    body[1:, 3] = BOX[0]
    body[1:, 4] = BOX[1]
    return body


def stack(columns, height, pitch=30.0, ground_half_width=1e7):
    """Horizontal row of box columns: x=(c-columns/2)*pitch, y=15+10h, column-major order."""
    c = np.repeat(np.arange(columns), height)
    h = np.tile(np.arange(height), columns)
    body = np.zeros((columns * height + 1, 6), dtype=np.float32)
    body[0] = (0.0, 0.0, 0.0, ground_half_width, 10.0, 1.0)
    body[1:, 0] = (c - columns / 2) * pitch
    body[1:, 1] = 15.0 + 10.0 * h
    body[1:, 3] = BOX[0]
    body[1:, 4] = BOX[1]
    return body


def brick_wall(width, height, pitch=21.0, ground_half_width=1e7):
    """A low, wide wall in running bond: `height` rows of `width` boxes, odd rows shifted by half a brick.  One island that
    is thousands of columns wide and a few rows high: many strips, every strip boundary cuts manifolds."""
    r = np.repeat(np.arange(height), width)
    i = np.tile(np.arange(width), height)
    body = np.zeros((width * height + 1, 6), dtype=np.float32)
    body[0] = (0.0, 0.0, 0.0, ground_half_width, 10.0, 1.0)
    body[1:, 0] = (i - width / 2) * pitch + (r % 2) * (pitch / 2)
    body[1:, 1] = 15.0 + 10.0 * r
    body[1:, 3] = BOX[0]
    body[1:, 4] = BOX[1]
    return body


def multi_island(count, rows, pitch=21.0, ground_half_width=1e7):
    """`count` separate pyramids of `rows` rows side by side (independent islands: the ground is
    static and does not merge islands, reference src/Solver.cpp:304,316-317)."""
    one = pyramid_fast(rows, pitch)[1:]
    span = rows * pitch + 2 * pitch
    body = np.zeros((count * one.shape[0] + 1, 6), dtype=np.float32)
    body[0] = (0.0, 0.0, 0.0, ground_half_width, 10.0, 1.0)
    for k in range(count):
        blk = one.copy()
        blk[:, 0] = (one[:, 0].astype(np.float64) + (k - count / 2) * span).astype(np.float32)
        body[1 + k * one.shape[0] : 1 + (k + 1) * one.shape[0]] = blk
    return body


def tumble(count, seed=12345, ground_half_width=1e7):
    """Boxes of mixed sizes dropped at mixed angles in a loose pile (deterministic LCG, no RNG module):
    exercises the vertex-face / vertex-vertex branches of the narrowphase that axis-aligned stacks never
    reach (reference src/Collider.cpp:122-160), plus manifold churn (pairs appear and disappear)."""
    body = np.zeros((count + 1, 6), dtype=np.float32)
    body[0] = (0.0, 0.0, 0.0, ground_half_width, 10.0, 1.0)
    s = seed
    cols = max(int(np.sqrt(count)), 1)
    for i in range(count):
        vals = []
        for _ in range(4):
            s = (s * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
            vals.append(((s >> 33) & 0xFFFFFF) / float(1 << 24))
        cx, cy = i % cols, i // cols
        body[i + 1] = ((cx - cols / 2) * 26.0 + (vals[0] - 0.5) * 8.0, 30.0 + cy * 24.0 + vals[1] * 6.0, (vals[2] - 0.5) * 2.4,
                       6.0 + 6.0 * vals[3], 4.0 + 3.0 * vals[0], 0.0)
    return body


def clump(count, ground_half_width=1e7):
    """Every box overlaps every other one: ~count^2/2 pairs out of `count` bodies and hundreds of joints per
    body.  Not physical; it drives the pair table through a rebuild, the pair buffer through a regrow and
    the device colouring past its 64-colour limit (host fallback)."""
    body = np.zeros((count + 1, 6), dtype=np.float32)
    body[0] = (0.0, 0.0, 0.0, ground_half_width, 10.0, 1.0)
    i = np.arange(count)
    body[1:, 0] = (i % 17) * 0.37 - 3.0
    body[1:, 1] = 60.0 + (i // 17) * 0.29
    body[1:, 2] = (i % 5) * 0.11
    body[1:, 3] = BOX[0]
    body[1:, 4] = BOX[1]
    return body


def platforms(count=400, seed=777):
    """The demo's "Stacks" layout in small (reference src/main.cpp:170-186): two platforms with
    invMass = 0 only (row flag 2: immovable but free to rotate, so NOT static for the solver: every
    joint on them conflicts) and boxes raining on them.  Platforms collect dozens of joints on one
    dynamic body: deep dependency chains in replay, more than 64 colours in colour mode."""
    body = np.zeros((count + 3, 6), dtype=np.float32)
    body[0] = (0.0, 0.0, 0.0, 1e7, 10.0, 1.0)
    body[1] = (0.0, 120.0, 0.0, 300.0, 10.0, 2.0)
    body[2] = (420.0, 60.0, 0.0, 200.0, 10.0, 2.0)
    s = seed
    for i in range(count):
        vals = []
        for _ in range(3):
            s = (s * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
            vals.append(((s >> 33) & 0xFFFFFF) / float(1 << 24))
        body[i + 3] = (vals[0] * 500.0 - 100.0, 160.0 + vals[1] * 900.0, vals[2] * 0.8, 4.0, 4.0, 0.0)
    return body


SCENES = {
    # BASELINE.json configs[0..4]
    "pyramid_1k": lambda: pyramid_fast(45),
    "stack_100k": lambda: stack(10000, 10),
    "pyramid_1m": lambda: pyramid_fast(1414),
    "islands_1m": lambda: multi_island(1024, 44),
    "stack_10m": lambda: stack(100000, 100, pitch=21.0),
    # smaller relatives used by tests
    "pyramid_10": lambda: pyramid_fast(10),
    "pyramid_10k": lambda: pyramid_fast(141),
    "pyramid_100k": lambda: pyramid_fast(447),
    "stack_1k": lambda: stack(100, 10),
    "stack_10k": lambda: stack(1000, 10),
    "islands_8x10": lambda: multi_island(8, 10),
    "islands_64x20": lambda: multi_island(64, 20),
    "islands_128k": lambda: multi_island(128, 44),   # one rank's share of islands_1m on 8 devices
    "clump_300": lambda: clump(300),
    "platforms_400": lambda: platforms(400),
    "tumble_300": lambda: tumble(300),
    "tumble_3k": lambda: tumble(3000),
    "wall_18k": lambda: brick_wall(3000, 6),
}


def make(name):
    return SCENES[name]()
