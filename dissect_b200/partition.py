"""Host-side mirror of the engine's tile partition and of the packed-row sharding used by the multi-GPU
driver (pure index arithmetic; the kernels consume the same tables built in csrc/dgrm_api.cu:build_tiles)."""


_ORDER4 = [(0, 0), (1, 1), (1, 0), (2, 0), (2, 1), (2, 2), (3, 3), (3, 2), (3, 0), (3, 1)]


def supertile_order(n_pad, group_rows=8, world=1):
    """Lower-triangle 256x256 supertiles (I, J), J <= I, in the engine's raster: groups of `group_rows`
    supertile rows swept column by column; with four or more engines inside a 4 x 4 grid of macro blocks visited in
    an order that pairs the diagonal halves (every unit of area is one square or two triangles), so that a contiguous
    segment is a compact region (csrc/dgrm_api.cu:build_tiles)."""
    T = n_pad // 256
    MB = 4 if world >= 4 else 1
    mb = (T + MB - 1) // MB
    out = []
    for bI, bJ in (_ORDER4 if MB == 4 else [(0, 0)]):
        if bI * mb >= T:
            continue
        i_lo, i_hi = bI * mb, min(T, bI * mb + mb)
        j_lo, j_hi = bJ * mb, min(T, bJ * mb + mb)
        for g in range(i_lo, i_hi, group_rows):
            imax = min(g + group_rows, i_hi) - 1
            for J in range(j_lo, min(j_hi, imax + 1)):
                for I in range(max(g, J), imax + 1):
                    out.append((I, J))
    return out


def tiles_of_rank(n, rank, world):
    """(row block of 128, column panel of 256) pairs owned by `rank`: the contiguous raster segment
    [rank*K/world, (rank+1)*K/world) of the K supertiles."""
    n_pad = (n + 255) // 256 * 256
    tiles = []
    order = supertile_order(n_pad, world=world)
    total = len(order)
    s0, s1 = total * rank // world, total * (rank + 1) // world
    for s, (I, J) in enumerate(order):
        if s0 <= s < s1:
            tiles.append((2 * I, J))
            tiles.append((2 * I + 1, J))
    return tiles


def individuals_touched(n, rank, world):
    """Fraction of the (padded) individuals whose operands `rank` needs: rows and columns of its supertiles."""
    n_pad = (n + 255) // 256 * 256
    need = set()
    for rb, cp in tiles_of_rank(n, rank, world):
        need.add(rb // 2)
        need.add(cp)
    return len(need) * 256 / n_pad


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
