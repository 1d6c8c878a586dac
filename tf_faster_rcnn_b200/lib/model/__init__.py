"""Mirror of the reference's lib/model import surface for the B200 path (tf_faster_rcnn_b200/lib)."""
