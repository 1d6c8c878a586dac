"""VGG16 conv5_3 backbone and fc6/fc7 head (lib/nets/vgg16.py:26-60) emitted onto the device tape."""
from nets.network import Network
from tf_faster_rcnn_b200 import _native as N


class vgg16(Network):
    def __init__(self):
        Network.__init__(self)
        self._scope = 'vgg_16'

    def _image_to_head(self, t, image):
        x = image
        for block, reps in enumerate((2, 2, 3, 3, 3), start=1):
            for i in range(1, reps + 1):
                name = "vgg_16/conv%d/conv%d_%d" % (block, block, i)
                if x.shape[3] == 3:
                    x = t.conv_first(x, name, 3, 1, "SAME", N.ACT_RELU)
                else:
                    x = t.conv(x, name, 1, "SAME", N.ACT_RELU)
            if block < 5:
                x = t.max_pool(x, 2, 2, "SAME")
        self._layers['head'] = x
        return x

    def _head_to_tail(self, t, pool5):
        r = pool5.shape[0]
        flat = pool5.view(r, -1)                    # (h, w, c) flatten order, vgg16.py:50
        fc6 = t.fc(flat, "vgg_16/fc6", N.ACT_RELU)
        return t.fc(fc6, "vgg_16/fc7", N.ACT_RELU)
