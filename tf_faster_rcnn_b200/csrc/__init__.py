"""CUDA sources (sm_100a) and the in-tree build script of libfrcnn_b200.so."""
