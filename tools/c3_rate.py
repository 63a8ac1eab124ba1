"""BASELINE configs[2] receive side alone (64 channels, M = 256, 16-QAM, conv r1/2 K=7), device-resident input:
   python tools/c3_rate.py [steps]      (the same leg bench.py reports under "config64")"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import types
import bench
from b2 import pkg
args = types.SimpleNamespace(steps=int(sys.argv[1]) if len(sys.argv) > 1 else 5)
print(bench.measure_config64(args, 0, pkg, lambda: None))
