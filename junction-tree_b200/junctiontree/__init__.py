"""B200-native sum-product propagation behind the junction-tree Python API."""
