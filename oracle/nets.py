"""Oracle: the three backbones' TEST-mode forward passes (test infrastructure).

  vgg16      lib/nets/vgg16.py:26-60
  resnet     lib/nets/resnet_v1.py:80-152 + slim resnet_v1.bottleneck / resnet_v1_block
             (external TF-slim source, restated: 1x1 -> 3x3(conv2d_same, stride) -> 1x1,
             BN after each, ReLU after the first two and after the residual add; shortcut =
             identity | 1x1/stride max-pool subsample when depth matches, else 1x1 conv + BN)
  mobilenet  lib/nets/mobilenet_v1.py:63-172,214-250 (BN eps 1e-3, ReLU6)

`w` maps the reference's TF variable names to numpy arrays (HWIO / [in,out]).
Every function returns NHWC fp32.  `tap` (optional dict) collects named intermediates so
CUDA kernels can be checked layer by layer.
"""
import numpy as np
from . import layers as L

F = np.float32

RESNET_UNITS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}
MOBILENET_DEFS = [("conv", 2, 32), ("sep", 1, 64), ("sep", 2, 128), ("sep", 1, 128), ("sep", 2, 256),
                  ("sep", 1, 256), ("sep", 2, 512), ("sep", 1, 512), ("sep", 1, 512), ("sep", 1, 512),
                  ("sep", 1, 512), ("sep", 1, 512), ("sep", 1, 1024), ("sep", 1, 1024)]


def _bn(x, w, name, eps):
    p = name + "/BatchNorm/"
    y, _, _ = L.batch_norm(x, w[p + "gamma"], w[p + "beta"], w[p + "moving_mean"], w[p + "moving_variance"], eps)
    return y


def _tap(tap, k, v):
    if tap is not None:
        tap[k] = v


# ---- VGG16 -------------------------------------------------------------------------------

def vgg16_body(w, x, tap=None):
    for b, n in enumerate([2, 2, 3, 3, 3], start=1):
        for i in range(1, n + 1):
            nm = "vgg_16/conv%d/conv%d_%d" % (b, b, i)
            x = L.relu(L.bias_add(L.conv2d(x, w[nm + "/weights"], 1, "SAME"), w[nm + "/biases"]))
            _tap(tap, nm, x)
        if b < 5:
            x = L.max_pool(x, 2, 2, "SAME")
            _tap(tap, "vgg_16/pool%d" % b, x)
    return x


def vgg16_tail(w, pool5, tap=None):
    flat = pool5.reshape(pool5.shape[0], -1)                      # (h,w,c) order, vgg16.py:50
    fc6 = L.relu(L.fully_connected(flat, w["vgg_16/fc6/weights"]) + w["vgg_16/fc6/biases"])
    _tap(tap, "vgg_16/fc6", fc6)
    fc7 = L.relu(L.fully_connected(fc6, w["vgg_16/fc7/weights"]) + w["vgg_16/fc7/biases"])
    return fc7


# ---- ResNet-v1 ---------------------------------------------------------------------------

def _block_plan(num_layers):
    n1, n2, n3, n4 = RESNET_UNITS[num_layers]
    return [("block1", 64, [1] * (n1 - 1) + [2]), ("block2", 128, [1] * (n2 - 1) + [2]),
            ("block3", 256, [1] * n3), ("block4", 512, [1] * n4)]


def _bottleneck(w, x, p, base, stride, eps=1e-5, tap=None):
    depth = base * 4
    if x.shape[3] == depth:
        sc = x if stride == 1 else L.max_pool(x, 1, stride, "VALID")
    else:
        sc = _bn(L.conv2d(x, w[p + "/shortcut/weights"], stride, "SAME"), w, p + "/shortcut", eps)
    r = L.relu(_bn(L.conv2d(x, w[p + "/conv1/weights"], 1, "SAME"), w, p + "/conv1", eps))
    _tap(tap, p + "/conv1", r)
    r = L.relu(_bn(L.conv2d_same(r, w[p + "/conv2/weights"], stride), w, p + "/conv2", eps))
    _tap(tap, p + "/conv2", r)
    r = _bn(L.conv2d(r, w[p + "/conv3/weights"], 1, "SAME"), w, p + "/conv3", eps)
    out = L.relu(sc + r)
    _tap(tap, p, out)
    return out


def resnet_body(w, x, num_layers, tap=None):
    sc = "resnet_v1_%d" % num_layers
    x = L.relu(_bn(L.conv2d_same(x, w[sc + "/conv1/weights"], 2), w, sc + "/conv1", 1e-5))
    _tap(tap, sc + "/conv1", x)
    x = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)))               # resnet_v1.py:83
    x = L.max_pool(x, 3, 2, "VALID")
    _tap(tap, sc + "/pool1", x)
    for bname, base, strides in _block_plan(num_layers)[:3]:
        for u, s in enumerate(strides, start=1):
            x = _bottleneck(w, x, "%s/%s/unit_%d/bottleneck_v1" % (sc, bname, u), base, s, tap=tap)
    return x


def resnet_tail(w, pool5, num_layers, tap=None):
    sc = "resnet_v1_%d" % num_layers
    bname, base, strides = _block_plan(num_layers)[3]
    x = pool5
    for u, s in enumerate(strides, start=1):
        x = _bottleneck(w, x, "%s/%s/unit_%d/bottleneck_v1" % (sc, bname, u), base, s, tap=tap)
    return x.mean(axis=(1, 2), dtype=F)                            # resnet_v1.py:124


# ---- MobileNet-v1 ------------------------------------------------------------------------

def _mobilenet_layers(w, x, first, last, tap=None):
    sc = "MobilenetV1"
    for i in range(first, last):
        kind, stride, _ = MOBILENET_DEFS[i]
        if kind == "conv":
            nm = "%s/Conv2d_%d" % (sc, i)
            x = L.relu6(_bn(L.conv2d_same(x, w[nm + "/weights"], stride), w, nm, 1e-3))
            _tap(tap, nm, x)
        else:
            nm = "%s/Conv2d_%d_depthwise" % (sc, i)
            dwk = w[nm + "/depthwise_weights"]
            x = L.relu6(_bn(L.conv2d_same(x, dwk, stride, groups=dwk.shape[2]), w, nm, 1e-3))
            _tap(tap, nm, x)
            nm = "%s/Conv2d_%d_pointwise" % (sc, i)
            x = L.relu6(_bn(L.conv2d(x, w[nm + "/weights"], 1, "SAME"), w, nm, 1e-3))
            _tap(tap, nm, x)
    return x


def mobilenet_body(w, x, tap=None):
    return _mobilenet_layers(w, x, 0, 12, tap)


def mobilenet_tail(w, pool5, tap=None):
    return _mobilenet_layers(w, pool5, 12, 14, tap).mean(axis=(1, 2), dtype=F)


# ---- dispatch ----------------------------------------------------------------------------

def scope_of(net):
    return {"vgg16": "vgg_16", "mobile": "MobilenetV1"}.get(net) or "resnet_v1_%d" % int(net[3:])


def image_to_head(net, w, blob, tap=None):
    if net == "vgg16":
        return vgg16_body(w, blob, tap)
    if net == "mobile":
        return mobilenet_body(w, blob, tap)
    return resnet_body(w, blob, int(net[3:]), tap)


def head_to_tail(net, w, pool5, tap=None):
    if net == "vgg16":
        return vgg16_tail(w, pool5, tap)
    if net == "mobile":
        return mobilenet_tail(w, pool5, tap)
    return resnet_tail(w, pool5, int(net[3:]), tap)
