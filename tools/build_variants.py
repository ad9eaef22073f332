"""Builds instrumented / tuning variants of the CUDA library next to the product one:
   python tools/build_variants.py name:DEF1=V1,DEF2=V2 ...   ->  juicer_b200/libjuicer_b200_<name>.so
Select one at run time with JUICER_B200_LIB=juicer_b200/libjuicer_b200_<name>.so."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from juicer_b200 import build as jb
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    out = os.path.join(jb.HERE, f"libjuicer_b200_{name}.so")
    jb.build(force=True, defines=tuple(d for d in defs.split(",") if d), out=out)
    print(out)
