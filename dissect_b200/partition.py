"""Host-side mirror of the engine's tile partition and of the packed-row sharding used by the multi-GPU
driver (pure index arithmetic; the kernels consume the same tables built in csrc/dgrm_api.cu:build_tiles)."""


def supertile_order(n_pad, group_rows=8):
    """Lower-triangle 256x256 supertiles (I, J), J <= I, in the engine's raster: groups of `group_rows`
    supertile rows swept column by column."""
    T = n_pad // 256
    out = []
    for g in range(0, T, group_rows):
        imax = min(g + group_rows, T) - 1
        for J in range(0, imax + 1):
            for I in range(max(g, J), imax + 1):
                out.append((I, J))
    return out


def tiles_of_rank(n, rank, world):
    """(row block of 128, column panel of 256) pairs owned by `rank`: the contiguous raster segment
    [rank*K/world, (rank+1)*K/world) of the K supertiles."""
    n_pad = (n + 255) // 256 * 256
    tiles = []
    order = supertile_order(n_pad)
    total = len(order)
    s0, s1 = total * rank // world, total * (rank + 1) // world
    for s, (I, J) in enumerate(order):
        if s0 <= s < s1:
            tiles.append((2 * I, J))
            tiles.append((2 * I + 1, J))
    return tiles


def row_shards(m_file, chunk, world):
    """Per chunk, the contiguous row range each rank keeps resident: [(c0, [(r0, r1) per rank])]."""
    shard = (chunk + world - 1) // world
    out = []
    for c0 in range(0, m_file, chunk):
        c1 = min(c0 + chunk, m_file)
        ranges = []
        for r in range(world):
            a = min(c0 + r * shard, c1)
            ranges.append((a, min(a + shard, c1)))
        out.append((c0, ranges))
    return out
