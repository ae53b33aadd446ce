"""izpi_b200 -- B200-native (sm_100a) backend for Izpi's hot path: BVH4 closest-hit traversal and
per-sample path tracing of image tiles.  See DESIGN.md."""
__version__ = "0.1.0"
