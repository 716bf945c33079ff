"""ncu driver: a few launches of the attention kernel at the bench shape (B = 32 CFG rows, T = 500, 8 heads)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import profiles.time_kernels as tk

tk.attention(int(os.environ.get("LS_B", "32")), int(os.environ.get("LS_T", "500")))
