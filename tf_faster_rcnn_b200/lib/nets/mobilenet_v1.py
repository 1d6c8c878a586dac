"""MobileNet-v1 (lib/nets/mobilenet_v1.py:63-172, 214-250): Conv2d_0 + 11 depthwise-separable layers as the
body, layers 12-13 + spatial mean as the per-RoI head.  Depthwise 3x3 is a bandwidth kernel; every pointwise
1x1 runs on the tcgen05 GEMM path.  BN eps 1e-3, ReLU6."""
from model.config import cfg
from nets.network import Network
from tf_faster_rcnn_b200 import _native as N

# (kind, stride, depth)
_DEFS = [("conv", 2, 32), ("sep", 1, 64), ("sep", 2, 128), ("sep", 1, 128), ("sep", 2, 256), ("sep", 1, 256),
         ("sep", 2, 512), ("sep", 1, 512), ("sep", 1, 512), ("sep", 1, 512), ("sep", 1, 512), ("sep", 1, 512),
         ("sep", 1, 1024), ("sep", 1, 1024)]
_EPS = 1e-3


class mobilenetv1(Network):
    def __init__(self):
        Network.__init__(self)
        self._depth_multiplier = cfg.MOBILENET.DEPTH_MULTIPLIER
        self._scope = 'MobilenetV1'

    def _layers_range(self, t, x, first, last):
        for i in range(first, last):
            kind, stride, _ = _DEFS[i]
            if kind == "conv":
                x = t.conv_first(x, "MobilenetV1/Conv2d_%d" % i, 3, stride, "EXPLICIT", N.ACT_RELU6, _EPS)
            else:
                x = t.depthwise(x, "MobilenetV1/Conv2d_%d_depthwise" % i, stride, N.ACT_RELU6, _EPS)
                x = t.conv(x, "MobilenetV1/Conv2d_%d_pointwise" % i, 1, "SAME", N.ACT_RELU6, _EPS)
        return x

    def _image_to_head(self, t, image):
        x = self._layers_range(t, image, 0, 12)
        self._layers['head'] = x
        return x

    def _head_to_tail(self, t, pool5):
        return t.spatial_mean(self._layers_range(t, pool5, 12, 14))
