"""B200-native Faster R-CNN inference path (package root; see README.md / DESIGN.md)."""
