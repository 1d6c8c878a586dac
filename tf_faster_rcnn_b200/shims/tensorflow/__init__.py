"""Minimal stand-in for the handful of TensorFlow-1.x symbols the reference's tools touch
(tools/demo.py:24,129-144; tools/test_net.py:18,86-117) so that they can drive this engine unchanged.
Only used when a real `tensorflow` is not importable (tf_faster_rcnn_b200.paths.add_lib_path(with_shims=True)).
No graph, no ops: the Session is an opaque token; Saver.restore loads an .npz of TF-named variables."""
import os
import numpy as np

__version__ = "1.x-shim (tf_faster_rcnn_b200)"


class _GPUOptions(object):
    allow_growth = False


class ConfigProto(object):
    def __init__(self, allow_soft_placement=False, **kw):
        self.allow_soft_placement = allow_soft_placement
        self.gpu_options = _GPUOptions()


class _InitOp(object):
    pass


def global_variables_initializer():
    return _InitOp()


def _networks():
    from nets import network
    return list(network._REGISTRY)


class Session(object):
    def __init__(self, config=None, **kw):
        self.config = config

    def run(self, fetches, feed_dict=None):
        if isinstance(fetches, _InitOp):
            # test_net.py:116-117 path (no --model): seeded synthetic initialisation
            from tf_faster_rcnn_b200 import synth
            for net in _networks():
                if net.weights is None:
                    name = {"vgg_16": "vgg16", "MobilenetV1": "mobile"}.get(net.scope) or "res%d" % net._num_layers
                    net.load_weights(synth.make(name, net.num_classes, net.num_anchors))
            return None
        raise NotImplementedError("tensorflow shim: Session.run only supports the variable initializer; "
                                  "inference goes through Network.test_image / im_detect")

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _Saver(object):
    def restore(self, sess, save_path):
        """Loads `<save_path>.npz` (or save_path itself if it is an .npz): TF variable name -> array."""
        path = save_path if save_path.endswith(".npz") else save_path + ".npz"
        if not os.path.isfile(path):
            raise IOError("no weights at %s: TF bundle (.index/.data) reading is not implemented; export variables "
                          "to an .npz keyed by TF variable names (tools/make_synthetic_ckpt.py writes one)" % path)
        with np.load(path) as z:
            tensors = {k: z[k] for k in z.files}
        for net in _networks():
            net.load_weights(tensors)


class _Train(object):
    Saver = _Saver


train = _Train()
