"""latency of the packet-decode kernel on a handful of conv-coded frames (one warp per frame): python tools/vit_latency.py
   prints b2_mcrx_last_timing(): [channelizer, synchroniser, decode, call] in ms for 1200- and 300-byte payloads"""
import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import test_gpu_parity as T
from refmc import *
from b2 import pkg
for plen in (1200, 300):
    case = (4, 256, 32, 8, MOD_QAM16, FEC_CONV_V27, FEC_NONE, plen, 1, 0.0)
    x = T.make_input(case)
    g = pkg.MultichannelRx(4, 256, 32, 8, max_batch=len(x))
    for _ in range(3):
        g.reset(); g.execute(x); fr, pl = g.poll()
        print(plen, "frames", len(fr), "valid", int(fr["payload_valid"].sum()), "timing", g.last_timing())
    g.close()
