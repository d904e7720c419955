"""B200-native drop-in for the `path_tracer` hot path of the Tungsten renderer."""
__version__ = "0.1.0"
