"""Puts the product package directory (hyphenated, so not importable by name) on sys.path."""
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "generalized-class-discovery-for-lidar-semantic-segmentation_b200")
for p in (PKG_DIR, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
