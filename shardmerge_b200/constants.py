"""Layer classes of a tensor name (same values as shard/constants.py:4-5)."""
INPUT_LAYER = -1
OUTPUT_LAYER = -2
