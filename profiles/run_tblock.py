"""ncu driver: a few launches of the fused transformer-block kernel at the bench shape (R = 16000 rows)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import profiles.time_kernels as tk

tk.timeit  # noqa
tk.tblock(int(os.environ.get("LS_R", "16000")), 0)
