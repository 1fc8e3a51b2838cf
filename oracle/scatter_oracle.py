"""
ORACLE (test infrastructure, NOT product code) -- numpy restatement of the two radar
point -> pixel scatters on the hot path.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this file.

S1  points_to_depth_map / merge_radar_point_clouds
    setup/setup_dataset_nuscenes_with_denseGT.py:814-840 (plot), :644-656 (main sweep),
    :699-713 (z-buffer rule for extra sweeps), :771-782 (nonzero -> point list).
    Parity status: PINNED.  The defining script cannot be imported (module-level nuScenes
    construction), so tests/golden/make_golden.py parses it, EXECUTES the reference's
    own statements of those line ranges (points_to_depth_map as a function) on seeded sweeps
    and checks this restatement against them bit for bit (tests/golden/s1_merge_64x96.npz).

S2  radarnet_main.forward's paste / max / arg-max / fill
    src/radarnet_main.py:563-591.  PINNED: tests/golden/make_golden.py runs the real
    function with a stub model and stores its outputs.
"""
import numpy as np


def s1_points_to_depth_map(points_xy, depth, height, width):
    """setup/...denseGT.py:814-840: np.round (half-to-even) -> int, ordered
    ``img[y, x] = z`` so the LAST writer wins.  points_xy: 2 x N, depth: N."""
    img = np.zeros((height, width), dtype=np.float64)
    q = np.round(np.asarray(points_xy)).astype(int)
    for i in range(q.shape[1]):
        img[q[1, i], q[0, i]] = depth[i]
    return img


def s1_merge_sweep(img, validity, points_xy, depth):
    """:699-713: a later sweep overwrites a pixel iff it is empty or the new point is
    closer.  In place on img / validity (validity: int array, 1 = occupied)."""
    q = np.round(np.asarray(points_xy)).astype(int)
    for i in range(q.shape[1]):
        x, y = q[0, i], q[1, i]
        if validity[y, x] == 1 and depth[i] < img[y, x]:
            img[y, x] = depth[i]
        elif validity[y, x] != 1:
            img[y, x] = depth[i]
            validity[y, x] = 1
    return img, validity


def s1_merge(sweeps, height, width):
    """merge_radar_point_clouds (:600-782) without the nuScenes projection: sweeps is a
    list of (points_xy 2 x N, depth N); the first is the main sweep."""
    img = s1_points_to_depth_map(sweeps[0][0], sweeps[0][1], height, width)
    validity = np.where(img > 0, 1, 0)          # :659
    for pts, dep in sweeps[1:]:
        s1_merge_sweep(img, validity, pts, dep)
    ys, xs = np.nonzero(img)                    # :771 row-major order
    return img, np.stack([xs, ys], axis=0), img[ys, xs]


def s2_scatter(crops, points, image_width, patch_size, compat=True):
    """src/radarnet_main.py:563-591.

    crops : K x 1 x ph x pw float32 sigmoid responses; points : K x 3 float32 with x
    already shifted by +pad (radarnet_main.py:980-983); image_width: UNPADDED width;
    the padded canvas is image_height(=ph + crop_height) x (image_width + 2*pad).
    Here the crop always spans the full canvas height (start row = H - ph).

    compat=True reproduces the reference bit for bit, including the int64 quirk
    (SURVEY 3.3): ``output`` is the int64 arg-max tensor, each fill truncates z toward
    zero and re-matches already filled pixels whose value aliases a later index.
    compat=False returns depth = z[argmax] in float32.
    Returns (depth 1 x H x W, response 1 x H x W)."""
    crops = np.asarray(crops, dtype=np.float32)
    points = np.asarray(points, dtype=np.float32)
    k, _, ph, pw = crops.shape
    pad = patch_size[1] // 2
    height = ph
    tiles = np.zeros((k, height, image_width + 2 * pad), dtype=np.float32)
    for i in range(k):
        c = np.where(crops[i, 0] < 0.5, np.float32(0), crops[i, 0])   # :566
        x = int(points[i, 0])                                         # :568 int() truncation
        tiles[i, height - ph:, x - pad:x + pad] = c
    tiles = tiles[:, :, pad:-pad]                                     # :572
    response = tiles.max(axis=0, keepdims=True)                       # :575
    arg = tiles.argmax(axis=0)[None].astype(np.int64)                 # first max wins (torch.max)
    if compat:
        out = arg.copy()
        for i in range(k):                                            # :578-582
            out = np.where(out == i, np.int64(points[i, 2]), out)     # full_like(int64) truncates
        depth = np.where(response == 0, np.int64(0), out)             # :585-588
    else:
        depth = np.where(response == 0, np.float32(0), points[arg, 2]).astype(np.float32)
    return depth, response


def png16_roundtrip(depth, response):
    """What the reference's PNG files do to RadarNet's outputs on their way into FusionNet (TEST INFRASTRUCTURE):
    save_depth / save_response (reference src/data_utils.py:271-286, 320-335: np.uint32(v * multiplier), mode 'I' -> 16-bit
    PNG) followed by load_depth / load_response (:238-269, 288-318: / multiplier, depth <= 0 -> 0).
    depth, response: numpy H x W (depth may be int64 as the reference produces it).  Returns float32 H x W each."""
    import numpy as np
    zd = np.uint32(np.float32(depth) * 256.0) & 0xffff          # PNG stores 16 bits
    zr = np.uint32(np.float32(response) * 2 ** 14) & 0xffff
    d = np.array(zd, dtype=np.float32) / 256.0
    d[d <= 0] = 0.0
    r = np.array(zr, dtype=np.float32) / 2 ** 14
    return d, r
