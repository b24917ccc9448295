"""veon_b200 -- B200-native lifting hot path of VISION-SJTU/VEON behind the
reference's operator surface."""
__version__ = "0.1.0"
