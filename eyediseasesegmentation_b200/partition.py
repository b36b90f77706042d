"""(image, tile) partition of the sliding-window path over the GPUs of one box.

The reference runs ``tta_patches`` on one GPU, tile after tile (src/main/tta.py:170,207); its only multi-GPU
mechanism is ``nn.DataParallel`` on the whole-image path (tta.py:101-104).  Here the units of work are
``(image, tile)`` pairs dealt round-robin to the ranks of a ``torch.distributed`` group (one process per GPU):

  * all TTA views of a tile stay on one rank (the view mean is a local reduction);
  * tiles are pasted with last-writer-wins (tta.py:213), so the pixels a tile OWNS are a static function of the
    ``make_grid`` order: its window minus every later window -- ``owned_cells``;
  * a rank writes and histograms only the pixels its tiles own; the per-image integer histograms
    ``[n_images, 2, bins]`` (+ the 19 x 2 straddle counters) of all ranks are summed by ONE all-reduce and equal
    the single-process histograms bin for bin, so every rank scans the same numbers;
  * the probability-map pieces are summed to the rank that writes the image's mask (non-owned pixels are 0).

Pure host logic (tested on CPU with gloo, world size 2); the device work is ``eds_paste_tiles_owned_x2_f32``
and ``eds_pr_hist_rects_f32``.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

Rect = Tuple[int, int, int, int]     # y, x, h, w


def world():
    """(rank, world_size) of the default process group, (0, 1) without one."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def tile_units(n_images: int, n_tiles: int, rank: int, world_size: int) -> List[Tuple[int, int]]:
    """This rank's ``(image, tile)`` units: unit u = image * n_tiles + tile goes to rank u % world_size."""
    return [(u // n_tiles, u % n_tiles) for u in range(rank, n_images * n_tiles, world_size)]


def batches(units: Sequence, max_batch: int) -> List[list]:
    """Split a unit list into ceil(n / max_batch) batches of (nearly) equal size, so that a rank's forward
    passes share one or two batch shapes (one CUDA graph each) instead of ending in a small remainder."""
    n = len(units)
    if n == 0:
        return []
    k = -(-n // max_batch)
    base, extra = divmod(n, k)
    out, pos = [], 0
    for b in range(k):
        size = base + (1 if b < extra else 0)
        out.append(list(units[pos:pos + size]))
        pos += size
    return out


def owned_cells(slices: Sequence[Sequence[int]], shape_hw) -> List[List[Rect]]:
    """For every tile ``(y1, y2, x1, x2)`` of ``slices`` (paste order) the rectangles it owns under
    last-writer-wins, clipped to the image.  Exact for any list of axis-aligned windows (duplicates and
    nested windows included): the plane is cut along every window edge and each cell goes to the LAST
    window that covers it; vertically adjacent cells of one tile are merged."""
    H, W = int(shape_hw[0]), int(shape_hw[1])
    rects = [(max(int(y1), 0), min(int(y2), H), max(int(x1), 0), min(int(x2), W)) for (y1, y2, x1, x2) in slices]
    ys = sorted({0, H} | {r[0] for r in rects} | {r[1] for r in rects})
    xs = sorted({0, W} | {r[2] for r in rects} | {r[3] for r in rects})
    cells: List[List[Rect]] = [[] for _ in rects]
    for xi in range(len(xs) - 1):
        x0, x1 = xs[xi], xs[xi + 1]
        run_owner, run_y0, run_y1 = -1, 0, 0
        for yi in range(len(ys) - 1):
            y0, y1 = ys[yi], ys[yi + 1]
            owner = -1
            for t in range(len(rects) - 1, -1, -1):
                r = rects[t]
                if r[0] <= y0 and y1 <= r[1] and r[2] <= x0 and x1 <= r[3] and r[1] > r[0] and r[3] > r[2]:
                    owner = t
                    break
            if owner == run_owner and y0 == run_y1:
                run_y1 = y1
            else:
                if run_owner >= 0:
                    cells[run_owner].append((run_y0, x0, run_y1 - run_y0, x1 - x0))
                run_owner, run_y0, run_y1 = owner, y0, y1
        if run_owner >= 0:
            cells[run_owner].append((run_y0, x0, run_y1 - run_y0, x1 - x0))
    return cells


def owner_map(slices: Sequence[Sequence[int]], shape_hw) -> np.ndarray:
    """Brute-force restatement of the paste loop (``preds[y1:y2, x1:x2] = tile`` in order): index of the tile
    whose value a pixel ends up with, -1 where no tile writes.  Used by the tests to check ``owned_cells``."""
    H, W = int(shape_hw[0]), int(shape_hw[1])
    own = np.full((H, W), -1, dtype=np.int32)
    for t, (y1, y2, x1, x2) in enumerate(slices):
        own[max(int(y1), 0):max(int(y2), 0), max(int(x1), 0):max(int(x2), 0)] = t
    return own


def allreduce_sum_(*tensors):
    """In-place sum over the ranks of the default group (NCCL over NVLink on the GPU box, gloo in the CPU
    tests).  Integer histograms travel as int32: the u32 counters add modulo 2^32 either way."""
    import torch.distributed as dist
    rank, ws = world()
    if ws > 1:
        for t in tensors:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return tensors


def writer_rank(image_index: int, world_size: int) -> int:
    """Rank that assembles and writes the mask of an image."""
    return image_index % world_size
