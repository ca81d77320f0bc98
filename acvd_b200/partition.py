"""Vertex-range partition used across GPUs: contiguous ranges of 32-vertex tiles per rank
(mirrors dist_tile_range in csrc/dist.cuh)."""


def tile_range(num_vertices: int, rank: int, world: int):
    n_tiles = (num_vertices + 31) // 32
    return n_tiles * rank // world, n_tiles * (rank + 1) // world


def vertex_range(num_vertices: int, rank: int, world: int):
    t0, t1 = tile_range(num_vertices, rank, world)
    return 32 * t0, min(num_vertices, 32 * t1)
