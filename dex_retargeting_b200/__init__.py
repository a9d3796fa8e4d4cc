"""B200-native batched hand-retargeting IK engine (drop-in for the dex_retargeting optimizer path)."""
__version__ = "0.1.0"
