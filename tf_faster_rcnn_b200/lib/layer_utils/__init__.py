"""Mirror of the reference's lib/layer_utils import surface for the B200 path (tf_faster_rcnn_b200/lib)."""
