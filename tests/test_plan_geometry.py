"""Host-side work decomposition of the conv/FC kernel (frcnn_conv_plan_geometry: no GPU needed): tile search, block_n,
ragged-round and small-layer K splits, for the layer shapes of BASELINE.json's configs on a 148-SM part."""
import ctypes as C

import pytest

from tf_faster_rcnn_b200 import _native as N

KEYS = ["block_n", "tile_n", "tile_h", "tile_w", "m_tiles", "n_tiles", "tiles", "split_tiles", "splits", "kb_per_split",
        "units", "grid", "k_blocks", "kb_per_chunk", "tiles_h", "tiles_w"]


def geom(n, h, w, cin, cout, k, stride=1, pad=None, ho=None, wo=None, sms=148, **kw):
    pad = (k // 2) if pad is None else pad
    ho = ho if ho is not None else (h + 2 * pad - k) // stride + 1
    wo = wo if wo is not None else (w + 2 * pad - k) // stride + 1
    d = N.ConvDesc(None, None, None, None, None, None, None, n, h, w, cin, cout, k, k, stride, pad, pad, ho, wo, 0,
                   kw.get("block_n", 0), kw.get("kb_per_chunk", 0), kw.get("split_k", 0), kw.get("impl", N.CONV_F16X3), 1.0)
    out = (C.c_int * 16)()
    N.check(N.lib().frcnn_conv_plan_geometry(C.byref(d), sms, out), "geometry")
    return dict(zip(KEYS, list(out)))


def rows(g):
    return g["tile_n"] * g["tile_h"] * g["tile_w"]


def test_pointwise_layers_flatten_to_full_tiles():
    g = geom(300, 7, 7, 2048, 512, 1)                      # ResNet head 1x1: 14700 pixels
    assert (g["tile_n"], g["tile_h"]) == (1, 1) and g["m_tiles"] == 115 and g["block_n"] == 128 and g["n_tiles"] == 4
    assert g["tiles"] == 460 and g["split_tiles"] == 460 - 3 * 148 and g["splits"] == 8 and g["kb_per_split"] == 4
    assert g["units"] == 444 + 16 * 8 and g["grid"] == 148 and g["k_blocks"] == 32 and g["kb_per_chunk"] == 4   # 64-wide k-blocks
    g = geom(300, 7, 7, 2048, 512, 1, impl=N.CONV_TF32X3)  # r01 kernel: 32-wide k-blocks
    assert g["splits"] == 8 and g["kb_per_split"] == 8 and g["k_blocks"] == 64 and g["kb_per_chunk"] == 8


def test_spatial_tiles_cover_the_map_with_at_most_128_rows():
    g = geom(1, 38, 50, 256, 256, 3)                       # ResNet block3 3x3 on the 38x50 map
    assert rows(g) <= 128 and (g["tile_h"], g["tile_w"]) == (5, 25) and g["m_tiles"] == 16
    assert g["tiles"] == 32 and g["split_tiles"] == 32 and g["splits"] == 4 and g["units"] == 128   # too small: every tile split
    g = geom(300, 7, 7, 512, 512, 3)                       # head 3x3 over 300 RoIs: tile = 18 RoIs x 1 row x 7
    assert (g["tile_n"], g["tile_h"], g["tile_w"]) == (18, 1, 7) and g["m_tiles"] == 17 * 7
    assert g["tiles"] == 476 and g["split_tiles"] == 32 and g["splits"] == 4 and g["kb_per_split"] == 18 and g["k_blocks"] == 72
    for shape in [(1, 600, 800, 64, 64, 3), (1, 75, 100, 256, 512, 3), (1, 50, 67, 1024, 256, 1), (1000, 7, 7, 512, 512, 3)]:
        g = geom(*shape)
        assert rows(g) <= 128 and g["tile_w"] * g["tiles_w"] >= (shape[2] if shape[5] == 3 else 1)
        assert g["grid"] == min(g["units"], 148)


def test_strided_conv_box_limit_and_short_k_layers():
    g = geom(1, 150, 200, 64, 64, 3, stride=2, pad=1, ho=75, wo=100)   # block1 last unit, traversal stride 2: box <= 256
    assert g["tile_w"] * 2 <= 256 and g["tile_h"] * 2 <= 256 and rows(g) <= 128
    g = geom(300, 7, 7, 512, 2048, 1)                      # K = 512: the ragged round is NOT worth a split
    assert g["tiles"] == 1840 and g["split_tiles"] == 0 and g["splits"] == 1 and g["units"] == 1840
    g = geom(1, 1, 300, 25088, 4096, 1)                    # VGG fc6: 3 x 32 tiles fill most SMs -> ragged rule needs tiles > SMs
    assert g["tiles"] == 96 and g["split_tiles"] == 0
    g = geom(1, 1, 300, 2048, 408, 1)                      # fused cls|bbox FC: tiny -> split 8 ways
    assert g["tiles"] == 12 and g["splits"] == 8 and g["units"] == 96


def test_forced_options_and_errors():
    assert geom(1, 38, 50, 256, 256, 1, block_n=64)["block_n"] == 64
    assert geom(1, 38, 50, 256, 256, 1, split_k=1)["split_tiles"] == 0
    g = geom(1, 38, 50, 1024, 256, 1, split_k=3)
    assert g["split_tiles"] == g["tiles"] and g["splits"] == 3 and g["kb_per_split"] == 6
    g = geom(1, 38, 50, 96, 64, 3)                         # 27 32-channel blocks -> 14 k-blocks, the last one half empty
    assert g["k_blocks"] == 14
    assert geom(1, 38, 50, 256, 64, 1)["block_n"] == 64    # cout 64: a 128-wide tile would be half empty
    d = N.ConvDesc(None, None, None, None, None, None, None, 1, 8, 8, 48, 64, 1, 1, 1, 0, 0, 8, 8, 0, 0, 0, 0, 0, 1.0)
    out = (C.c_int * 16)()
    assert N.lib().frcnn_conv_plan_geometry(C.byref(d), 148, out) == -2 and "multiple of 32" in N.last_error()
