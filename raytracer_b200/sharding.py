"""Host-side statement of the multi-GPU film decomposition used by rtb_resize (raytracer_b200/csrc/rtb_kernels.cu):
tile t (row-major over the grid of tile_size x tile_size tiles, RenderingParams::tileSize = 32 in the reference,
Context.h:76) belongs to shard t % num_shards; inside a tile pixels go in 8x4 blocks (row-major over
the blocks and inside a block, one warp per block); every shard is padded to the same
number of slots so ONE all-gather of equal-sized buffers reassembles the frame."""
import numpy as np


def shard_pixel_lists(width, height, num_shards, tile_size=32):
    tiles_x = (width + tile_size - 1) // tile_size
    tiles_y = (height + tile_size - 1) // tile_size
    lists = [[] for _ in range(num_shards)]
    for ty in range(tiles_y):
        for tx in range(tiles_x):
            y1, x1 = min((ty + 1) * tile_size, height), min((tx + 1) * tile_size, width)
            for by in range(ty * tile_size, y1, 4):
                for bx in range(tx * tile_size, x1, 8):
                    ys = np.arange(by, min(by + 4, y1))
                    xs = np.arange(bx, min(bx + 8, x1))
                    lists[(ty * tiles_x + tx) % num_shards].append((ys[:, None] * width + xs[None, :]).ravel())
    out = [np.concatenate(l).astype(np.uint32) if l else np.zeros(0, np.uint32) for l in lists]
    slots = max(len(o) for o in out)
    return out, slots


def unshard(gathered, width, height, num_shards, tile_size=32):
    """gathered: (num_shards * slots, 3) float32, shard buffers back to back -> (height, width, 3) raster."""
    lists, slots = shard_pixel_lists(width, height, num_shards, tile_size)
    gathered = np.asarray(gathered, dtype=np.float32).reshape(num_shards, slots, 3)
    img = np.zeros((height * width, 3), dtype=np.float32)
    for s, l in enumerate(lists):
        img[l] = gathered[s, :len(l)]
    return img.reshape(height, width, 3)
