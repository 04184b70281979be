"""``fireflies/sampling/poisson.py`` -- Poisson-disk (blue-noise) initialisation of a laser pattern.

Set-up code: Bridson's dart throwing is sequential by construction and runs once before optimisation starts, on the
host with numpy like the reference.  It consumes ``np.random`` draw for draw in the reference's order, so a seeded
run gives the same pattern (pinned by ``tests/golden/poisson_misc.npz``).
"""
import numpy as np


def getGridCoordinates(coords):
    return np.floor(coords).astype("int")


def _window_occupied(occupied: np.ndarray, cell, half: int) -> bool:
    h, w = occupied.shape
    return bool(occupied[max(cell[0] - half, 0):min(cell[0] + half + 1, h), max(cell[1] - half, 0):min(cell[1] + half + 1, w)].any())


def bridson(radius, k=30, radiusType="default"):
    """Poisson-disk sampling with a spatially varying radius (sampling/poisson.py:16-116).

    ``radius``: 2-D array, the minimum distance per cell; its shape is the sampling box.  Each round picks a random
    active point and throws ``k`` darts into its annulus (distance ``r * (u + 1)`` -- or ``r * N(1.5, 0.2)`` for
    ``radiusType="normDist"`` -- and angle ``2 pi u``); every dart that lands inside the box with no occupied cell within
    ``ceil(r)`` cells becomes a sample and an active point (the round keeps throwing after a hit); a point with no hit
    in a round is retired.  Returns ``(count, coordinates [count, 2])``.
    """
    radius = np.asarray(radius)
    height, width = radius.shape
    occupied = np.zeros((height, width), dtype=bool)
    seed = (np.random.random() * height, np.random.random() * width)
    seed_cell = getGridCoordinates(seed)
    occupied[seed_cell[0], seed_cell[1]] = True
    active, samples = [seed], [seed]
    while active:
        pick = np.random.randint(len(active))
        base_pt = active[pick]
        base_cell = getGridCoordinates(base_pt)
        base_radius = radius[base_cell[0], base_cell[1]]
        hit = False
        for _ in range(k):
            if radiusType == "default":
                dist = base_radius * (np.random.random() + 1)
            elif radiusType == "normDist":
                dist = base_radius * np.random.normal(1.5, 0.2)
            angle = 2 * np.pi * np.random.random()
            dart = np.array([base_pt[0] + dist * np.sin(angle), base_pt[1] + dist * np.cos(angle)])
            if not (0 <= dart[1] <= width and 0 <= dart[0] <= height):
                continue
            cell = getGridCoordinates(dart)
            if _window_occupied(occupied, cell, int(np.ceil(radius[cell[0], cell[1]]))):
                continue
            active.append(dart)
            samples.append(dart)
            occupied[cell[0], cell[1]] = True
            hit = True
        if not hit:
            del active[pick]
    return (len(samples), np.array(samples))
