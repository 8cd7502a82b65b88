"""Builds experiment variants of the library next to the product one (lib/variants/<name>.so), for A/B timing on the
GPU box with BG_LIB_PATH.  usage: python tools/build_variants.py name:DEF1,DEF2=val ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from buckgnn_b200 import build
d = os.path.join(build.LIB_DIR, "variants")
os.makedirs(d, exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    out = os.path.join(d, name + ".so")
    build.build(force=True, defines=tuple(x for x in defs.split(",") if x), out=out)
    print("built", out)
