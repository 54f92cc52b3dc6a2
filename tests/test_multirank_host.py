"""CPU: index arithmetic of the multi-GPU path — the tile partition covers the lower triangle once, in contiguous
raster segments.  (The exchange loop itself runs over gloo in tests/test_multi_driver_cpu.py.)"""
from dissect_b200.partition import row_shards, supertile_order, tiles_of_rank


def test_tile_partition_is_a_partition():
    for n in (1, 255, 256, 257, 2000, 5000):
        n_pad = (n + 255) // 256 * 256
        T = n_pad // 256
        for world in (1, 8):
            order = supertile_order(n_pad, world=world)
            assert len(order) == T * (T + 1) // 2 and len(set(order)) == len(order)
            assert all(J <= I for I, J in order)
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                seen += tiles_of_rank(n, r, world)
            assert len(seen) == len(set(seen)) == T * (T + 1)          # two 128-row tiles per supertile
            counts = [len(tiles_of_rank(n, r, world)) for r in range(world)]
            assert max(counts) - min(counts) <= 2                       # balanced tile sets


def test_tile_sets_are_contiguous_raster_segments():
    """Engine r owns one contiguous segment of the raster: its concurrently running CTAs then share
    A/B panels like on one GPU.  (Round-robin dealing gave every engine one row of each column.)"""
    for n, world in ((5000, 3), (20000, 8), (56569, 8)):
        n_pad = (n + 255) // 256 * 256
        pos = {st: i for i, st in enumerate(supertile_order(n_pad, world=world))}
        ends = []
        for r in range(world):
            idx = sorted({pos[(rb // 2, cp)] for rb, cp in tiles_of_rank(n, r, world)})
            assert idx == list(range(idx[0], idx[-1] + 1))          # contiguous
            ends.append((idx[0], idx[-1]))
        assert ends[0][0] == 0 and ends[-1][1] == len(pos) - 1
        assert all(ends[r + 1][0] == ends[r][1] + 1 for r in range(world - 1))
        # panel sharing inside a window of 74 consecutive supertiles (one wave of CTA pairs): few distinct rows/columns
        order = supertile_order(n_pad, world=world)
        lo = ends[world // 2][0]
        window = order[lo: lo + 74]
        panels = len({I for I, _ in window}) + len({J for _, J in window})
        dealt = order[world // 2:: world][:74]                       # the same wave under round-robin dealing
        panels_rr = len({I for I, _ in dealt}) + len({J for _, J in dealt})
        assert panels <= 32 and (len(dealt) < 74 or panels < panels_rr)


def test_row_shards_cover_every_row_once():
    for m, chunk, world in ((1000, 256, 2), (8192, 8192, 8), (100000, 8192, 4), (5, 128, 2)):
        rows = []
        for c0, ranges in row_shards(m, chunk, world):
            for a, b in ranges:
                rows += list(range(a, b))
        assert rows == list(range(m))


def test_compact_partition_halves_the_individuals_an_engine_touches():
    """With 8 engines the macro-block raster makes every tile set a compact region: an engine needs the operands of
    about half of the individuals (a row band across all columns needs up to all of them), which halves the decode
    work that is replicated on every engine."""
    from dissect_b200.partition import individuals_touched
    for n in (56569, 100000):
        fr = [individuals_touched(n, r, 8) for r in range(8)]
        assert max(fr) <= 0.76 and sum(fr) / 8 <= 0.62, fr      # a row band needs up to 1.0 (0.77 on average)
        assert min(fr) >= 0.24, fr
    # 2 engines keep the plain raster (the second one needs everybody anyway)
    assert individuals_touched(28284, 1, 2) == 1.0
