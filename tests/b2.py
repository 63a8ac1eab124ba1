"""test-side access to the product package (its directory name has a hyphen)"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
pkg = importlib.import_module("liquid-usrp_b200")
