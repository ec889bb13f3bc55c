"""CPU model of pyramid_strip_kernel's bookkeeping (csrc/detect_pyramid.cu) checked against the oracle's area resize.

The kernel's arithmetic is exact integer sums followed by two fp32 divisions, so what can go wrong is the BOOKKEEPING: which
CTA owns which output row / column (strips, column tiles, level groups), when a level's window starts, ends and whether the
next window includes the current source row, and the modular u16 running-sum / snapshot trick.  This test restates exactly
that bookkeeping in numpy (same formulas as the kernel, strips and tiles forced small so that every boundary case occurs)
and compares every level with oracle.detect.area_resize + normalize bit for bit.  The CUDA kernel itself is compared with
the same oracle in tests/test_gpu_detect_kernels.py::test_pyramid_resize_bit_exact."""
import numpy as np
import pytest
import torch

from oracle import detect


def strip_model(frame, lhs, lws, n_strips, threads):
    H, W, _ = frame.shape
    assert (W * 3) % 16 == 0
    total_chunks = W * 3 // 16
    kw_max = max((W + lw - 1) // lw + 1 for lw in lws)
    halo_chunks = (3 * kw_max + 15) // 16 + 1
    assert halo_chunks < threads // 2
    own_max = threads - halo_chunks
    tiles = (total_chunks + own_max - 1) // own_max
    tile_chunks = (total_chunks + tiles - 1) // tiles
    sh = (H + n_strips - 1) // n_strips
    strips = (H + sh - 1) // sh
    outs = [np.full((3, lh, lw), np.nan, np.float32) for lh, lw in zip(lhs, lws)]
    rows_u8 = frame.reshape(H, W * 3).astype(np.uint32)
    for strip in range(strips):
        for tile in range(tiles):
            ys, ye = strip * sh, min(H, strip * sh + sh)
            c0, c1 = tile * tile_chunks, min(tile * tile_chunks + tile_chunks, total_chunks)
            px0 = (16 * c0 + 2) // 3
            px1 = W if c1 == total_chunks else (16 * c1 + 2) // 3
            rows = 0
            for lh in lhs:
                oyf, oye = (ys * lh + H - 1) // H, min(lh, (ye * lh + H - 1) // H)
                if oye > oyf:
                    rows = max(rows, (oye * H + lh - 1) // lh - ys)
            ncol = threads * 16
            running = np.zeros(ncol, np.uint32)                      # u16 lanes wrap: modelled with & 0xFFFF
            snap = [np.zeros(ncol, np.uint32) for _ in lhs]
            for r in range(rows):
                y = ys + r
                row = np.zeros(ncol, np.uint32)
                nb = min(ncol, W * 3 - 16 * c0)
                row[:nb] = rows_u8[y, 16 * c0:16 * c0 + nb]
                ev = []
                for l, lh in enumerate(lhs):
                    oyf, oye = (ys * lh + H - 1) // H, min(lh, (ye * lh + H - 1) // H)
                    a, bb = (y * lh) // H, ((y + 1) * lh + H - 1) // H - 1
                    own_a, own_b = oyf <= a < oye, oyf <= bb < oye
                    flush = own_a and ((a + 1) * H + lh - 1) // lh == y + 1
                    start = oye > oyf and (oyf * H) // lh == y
                    ev.append((start, flush, flush and bb > a and own_b, a))
                    if start:
                        snap[l] = running.copy()                     # snapshot BEFORE this row
                running = (running + row) & 0xFFFF
                for l, (start, flush, incl, oy) in enumerate(ev):
                    if not flush:
                        continue
                    lh, lw = lhs[l], lws[l]
                    d = (running - snap[l]) & 0xFFFF                 # exact: a window has <= 257 rows
                    snap[l] = ((running - row) & 0xFFFF) if incl else running.copy()
                    kh = np.float32(y + 1 - (oy * H) // lh)
                    oxf = (px0 * lw + W - 1) // W
                    oxe = lw if px1 == W else (px1 * lw + W - 1) // W
                    for ox in range(oxf, oxe):
                        x0, x1 = (ox * W) // lw, ((ox + 1) * W + lw - 1) // lw
                        for ch in range(3):
                            idx = np.arange(x0, x1) * 3 + ch - 16 * c0
                            assert idx.min() >= 0 and idx.max() < ncol
                            assert np.isnan(outs[l][ch, oy, ox]), "output written twice"
                            v = np.float32(d[idx].sum()) / kh / np.float32(x1 - x0)
                            outs[l][ch, oy, ox] = (v - np.float32(127.5)) * np.float32(0.0078125)
    return outs


@pytest.mark.parametrize("H,W,minsize,strips,threads", [(90, 128, 20, 3, 384), (123, 256, 30, 5, 48), (67, 64, 14, 2, 384),
                                                         (200, 320, 40, 7, 64)])
def test_strip_bookkeeping_matches_area_resize(H, W, minsize, strips, threads):
    rng = np.random.RandomState(H * W)
    fr = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    fr[: H // 3] = 255                                               # saturated rows: the modular sums must still be exact
    scales = detect.scale_pyramid(H, W, minsize, 0.709)
    sizes = [detect.level_size(H, W, s) for s in scales]
    outs = strip_model(fr, [a for a, _ in sizes], [b for _, b in sizes], strips, threads)
    x = torch.from_numpy(fr).permute(2, 0, 1)[None].float()
    for l, (lh, lw) in enumerate(sizes):
        ref = detect.normalize(detect.area_resize(x, (lh, lw)))[0].numpy()
        assert not np.isnan(outs[l]).any(), "level %d: outputs nobody owns" % l
        assert np.array_equal(outs[l], ref), "level %d differs by %g" % (l, np.abs(outs[l] - ref).max())
